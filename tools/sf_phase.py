"""dev tool: per-phase clock share of k_select_fast (library built with `make timing`, AMBC_LIB_PATH=.../libambc_timing.so)"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptive_compression_b200 import engine, _lib as L
lib = engine.require_cuda()
n = 64 << 20
names = ['csv', 'log', 'runs', 'lowcard', 'binrec', 'random', 'text']
ph = {1: 'rle', 2: 'stats', 3: 'huff_build', 4: 'index', 5: 'parse', 6: 'lz_emit', 7: 'huff_emit',
      10: 'ix.count', 11: 'ix.scan', 12: 'ix.scatter', 13: 'p.spec', 14: 'p.wait', 15: 'p.stitch', 16: 'ix.fo'}
buf = (C.c_ulonglong * 48)()
lib.ambc_enable_timing(1)
kinds = [int(x) for x in sys.argv[1].split(',')] if len(sys.argv) > 1 else (0, 1, 3, 4, 6)
for k in kinds:
    t = engine.synth(n, 0, kind_mask=1 << k)
    engine.compress_device(t, 4096)
    lib.ambc_sf_phase_read(buf, 1)
    engine.compress_device(t, 4096)
    ms = (C.c_float * 4)(); lib.ambc_last_timing(ms)
    lib.ambc_sf_phase_read(buf, 1)
    v = list(buf)
    nch = n // 4096
    print(names[k], "sel %.2f ms |" % ms[0], " ".join("%s=%.0f" % (ph[i], v[i] / nch) for i in sorted(ph)),
          "| per chunk: evals=%.0f batches=%.0f wordsteps=%.0f warp-iters spec=%.0f fix=%.0f chain-iters spec=%.0f fix=%.0f rounds=%.1f redos=%.1f" % tuple(v[i] / nch for i in (30, 31, 32, 33, 34, 35, 36, 37, 38)), flush=True)
