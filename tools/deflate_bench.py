"""dev tool: method id 5 on the GPU (reported separately from the chunk path): encode / decode throughput of the
batch entry points on 4 KiB items of the bench corpus, payload size against zlib level 9 and level 1"""
import ctypes as C, sys, os, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from adaptive_compression_b200 import engine, _lib as L
lib = engine.require_cuda()
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n, chunk = mib << 20, 4096
t = engine.synth(n, 0)
items = n // chunk
offs = torch.arange(0, n + 1, chunk, dtype=torch.int64, device="cuda")
stride = (int(lib.ambc_codec_bound(5, chunk)) + 15) & ~15
out = torch.empty(stride * items, dtype=torch.uint8, device="cuda")
lens = torch.empty(items, dtype=torch.int32, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for rep in range(2):
    ev[0].record()
    L.check(lib.ambc_codec_encode_batch(5, C.c_void_p(t.data_ptr()), C.c_void_p(offs.data_ptr()), items, C.c_void_p(out.data_ptr()),
                                        stride, C.c_void_p(lens.data_ptr()), st))
    ev[1].record()
    torch.cuda.synchronize()
enc_ms = ev[0].elapsed_time(ev[1])
hl = lens.cpu().numpy()
assert (hl > 0).all()
# pack the payloads and decode them back
ho = out.cpu().numpy()
blob = np.concatenate([ho[i * stride:i * stride + hl[i]] for i in range(items)])
poffs = torch.from_numpy(np.concatenate([[0], np.cumsum(hl.astype(np.int64))])).cuda()
tin = torch.from_numpy(blob).cuda()
orig = torch.full((items,), chunk, dtype=torch.int32, device="cuda")
dec = torch.empty(n, dtype=torch.uint8, device="cuda")
dl = torch.empty(items, dtype=torch.int32, device="cuda")
for rep in range(2):
    ev[2].record()
    L.check(lib.ambc_codec_decode_batch(5, C.c_void_p(tin.data_ptr()), C.c_void_p(poffs.data_ptr()), C.c_void_p(orig.data_ptr()), items,
                                        C.c_void_p(dec.data_ptr()), chunk, C.c_void_p(dl.data_ptr()), st))
    ev[3].record()
    torch.cuda.synchronize()
dec_ms = ev[2].elapsed_time(ev[3])
assert torch.equal(dec, t)
host = t.cpu().numpy().tobytes()
sample = 2048
z9 = sum(len(zlib.compress(host[i * chunk:(i + 1) * chunk], 9)) for i in range(sample))
z1 = sum(len(zlib.compress(host[i * chunk:(i + 1) * chunk], 1)) for i in range(sample))
assert all(zlib.decompress(bytes(ho[i * stride:i * stride + hl[i]])) == host[i * chunk:(i + 1) * chunk] for i in range(0, items, 97))
print("DEFLATE id 5, %d MiB of the mixed corpus in 4 KiB items: GPU encode %.2f GB/s, GPU inflate %.2f GB/s; payload / input: GPU %.3f, "
      "zlib level 9 %.3f, level 1 %.3f (first %d items)" % (mib, n / enc_ms / 1e6, n / dec_ms / 1e6, hl.sum() / n,
                                                          z9 / (sample * chunk), z1 / (sample * chunk), sample))
