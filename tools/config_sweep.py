"""BASELINE.json configs measured on one B200 (device-resident, CUDA events around the C-ABI calls):
  configs[0]  64 MiB mixed file through main.py (file -> .ambc -> file, MD5 and file I/O included)
  configs[2]  chunk-size sweep 1024 / 2048 / 4096 / 8192 / 16384 on a 4 GiB log corpus
  configs[4]  high-entropy / run-heavy / low-cardinality segments interleaved at 4 KiB (strict and per-chunk-raw)
prints a markdown table"""
import ctypes as C, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from adaptive_compression_b200 import engine, _lib as L
lib = engine.require_cuda()


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return r, best


rows = []
# configs[2]
n = (4 << 30) - 65536  # (a 4 GiB raw package does not fit the u32 length fields -- the reference raises there too)
t = engine.synth(n, 0, kind_mask=1 << 1)
for chunk in (1024, 2048, 4096, 8192, 16384):
    bound = lib.ambc_compress_bound(n, chunk, 4)
    out = torch.empty(bound, dtype=torch.uint8, device="cuda")
    work = torch.empty(lib.ambc_compress_workspace_bytes(n, chunk), dtype=torch.uint8, device="cuda")
    o, ms_c = timed(lambda: engine.compress_device(t, chunk, out=out, work=work), reps=2)
    dec = torch.empty(n, dtype=torch.uint8, device="cuda")
    (d, st), ms_d = timed(lambda: engine.decompress_device(o.body, n, out=dec), reps=2)
    assert st == [0, 0] and torch.equal(d, t)
    rows.append(("configs[2] 4 GiB - 64 KiB log corpus, chunk %d" % chunk, n / ms_c / 1e6, n / ms_d / 1e6, o.body_len / n,
                 "usage %s%s" % (o.usage, "  (no native method eligible: one raw package)" if chunk > 8192 else "")))
    del out, work, dec, o, d
del t
torch.cuda.empty_cache()
# configs[4]
sys.path.insert(0, os.path.join(ROOT, "tests"))
import inputs
parts = []
for i in range(3 * 2048):
    parts.append(inputs.make(("runs", "lowcard", "rand")[i % 3], 4096, 5000 + i % 97))
data = b"".join(parts)
t = engine.to_device(data)
for pcr in (False, True):
    o, ms_c = timed(lambda: engine.compress_device(t, 4096, flags=L.F_PER_CHUNK_RAW if pcr else 0))
    (d, st), ms_d = timed(lambda: engine.decompress_device(o.body, len(data)))
    assert st == [0, 0] and torch.equal(d, t)
    rows.append(("configs[4] interleaved 24 MiB, %s" % ("per-chunk-raw extension" if pcr else "strict (rest of file raw after chunk %d)" % o.first_raw),
                 len(data) / ms_c / 1e6, len(data) / ms_d / 1e6, o.body_len / len(data), "usage %s" % o.usage))
# configs[0]
with tempfile.TemporaryDirectory() as td:
    src, dst, back = os.path.join(td, "in.bin"), os.path.join(td, "out.ambc"), os.path.join(td, "back.bin")
    engine.synth(64 << 20, 0).cpu().numpy().tofile(src)
    t0 = time.perf_counter()
    subprocess.check_call([sys.executable, os.path.join(ROOT, "main.py"), "compress", src, dst, "--chunk-size", "4096", "--no-history"],
                          stdout=subprocess.DEVNULL)
    t1 = time.perf_counter()
    subprocess.check_call([sys.executable, os.path.join(ROOT, "main.py"), "decompress", dst, back], stdout=subprocess.DEVNULL)
    t2 = time.perf_counter()
    assert open(src, "rb").read() == open(back, "rb").read()
    from adaptive_compression_b200 import AdaptiveCompressor
    c = AdaptiveCompressor(chunk_size=4096)
    c.compress(src, dst)
    t3 = time.perf_counter(); c.compress(src, dst); t4 = time.perf_counter(); c.decompress(dst, back); t5 = time.perf_counter()
    rows.append(("configs[0] AdaptiveCompressor.compress / .decompress on the 64 MiB file, warm process (file I/O, MD5 included)",
                 (64 << 20) / (t4 - t3) / 1e9, (64 << 20) / (t5 - t4) / 1e9, os.path.getsize(dst) / (64 << 20),
                 "wall %.3f s compress, %.3f s decompress" % (t4 - t3, t5 - t4)))
    rows.append(("configs[0] main.py on a 64 MiB file (process start, CUDA init, file I/O, MD5 included)",
                 (64 << 20) / (t1 - t0) / 1e9, (64 << 20) / (t2 - t1) / 1e9, os.path.getsize(dst) / (64 << 20),
                 "wall %.2f s compress, %.2f s decompress" % (t1 - t0, t2 - t1)))
print("| config | compress GB/s | decompress GB/s | size ratio | notes |\n|---|---|---|---|---|")
for name, c, d, ratio, note in rows:
    print("| %s | %.2f | %.2f | %.4f | %s |" % (name, c, d, ratio, note))
