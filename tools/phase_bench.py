"""dev tool: per-phase clock share of k_select (library built with -DAMBC_PHASE_TIMING)"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptive_compression_b200 import engine, _lib as L
lib = engine.require_cuda()
n = 64 << 20
names = ['csv', 'log', 'runs', 'lowcard', 'binrec', 'random', 'text']
ph = {1: 'features', 2: 'lz level4', 3: 'lz level3', 4: 'double8', 5: 'refine(4,8)', 6: 'double16', 7: 'refine(8,16)',
      8: 'double32', 9: 'refine(16,32)', 10: 'gates', 11: 'LZ total', 12: 'huffman', 13: 'rle encode'}
buf = (C.c_ulonglong * 32)()
for k in (0, 1, 3, 4, 6):
    t = engine.synth(n, 0, kind_mask=1 << k)
    engine.compress_device(t, 4096)
    lib.ambc_phase_read(buf, 1)
    engine.compress_device(t, 4096)
    lib.ambc_phase_read(buf, 1)
    v = list(buf)
    nch = n // 4096
    tot = v[1] + v[10] + v[11] + v[12] + v[13]
    print(names[k], "cycles/chunk %.0f" % (tot / nch), " ".join("%s=%.0f" % (ph[i], v[i] / nch) for i in sorted(ph)),
          "chain+emit=%.0f" % ((v[11] - sum(v[2:10])) / nch),
          "| binary: init=%.0f count/scan=%.0f clear=%.0f insert=%.0f resolve=%.0f copy=%.0f participants=%.0f" %
          tuple(v[i] / nch for i in (14, 15, 16, 17, 18, 19, 20)),
          "| features: hist=%.0f samples=%.0f trigrams=%.0f 3sums=%.0f pairs=%.0f entropy=%.0f" % tuple(v[i] / nch for i in (21, 22, 23, 24, 25, 26)),
          "| chain: exits=%.0f walk=%.0f count+scan=%.0f emit=%.0f" % tuple(v[i] / nch for i in (27, 28, 29, 30)), flush=True)
