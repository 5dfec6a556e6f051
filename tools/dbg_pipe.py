"""dev tool: repeat the host-pipelined round trip of the mixed corpus with the random kind and report mismatches"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from adaptive_compression_b200 import engine, _lib as L
lib = engine.require_cuda()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for mask, chunk in ((127, 4096), (95, 4096)):
    n = (150 << 20) + 12345
    t = engine.synth(n, 0, kind_mask=mask)
    data = t.cpu().numpy()
    bound = lib.ambc_compress_bound(n, chunk, 4)
    body = np.empty(bound, dtype=np.uint8)
    res = L.CompressResult()
    L.check(lib.ambc_compress_host(C.c_void_p(data.ctypes.data), n, chunk, L.NATIVE_MASK, 0, b"\xff\xff\x00\x00", 4,
                                   C.c_void_p(body.ctypes.data), bound, None, None, C.byref(res)))
    print("mask", mask, "body", res.body_len, "first_raw", res.first_raw, "packages", res.n_packages, flush=True)
    for r in range(reps):
        back = np.full(n + 100, 7, dtype=np.uint8)
        st = (C.c_uint32 * 2)()
        rc = lib.ambc_decompress_host(C.c_void_p(body.ctypes.data), res.body_len, b"\xff\xff\x00\x00", 4, L.NATIVE_MASK,
                                      C.c_void_p(back.ctypes.data), n + 100, st)
        bad = np.nonzero(back[:n] != data)[0]
        if rc or len(bad) or back[n:].any() or list(st) != [0, 0]:
            print(" rep", r, "rc", rc, "status", list(st), "bad bytes", len(bad), "first", bad[:1], "last", bad[-1:],
                  "tail", back[n:].any(), flush=True)
            if len(bad):
                ch = np.unique(bad // 65536)
                print("  64K blocks", ch[:10], len(ch), "values", back[bad[:8]], data[bad[:8]], flush=True)
print("done")
