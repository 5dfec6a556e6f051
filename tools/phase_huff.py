"""dev tool: Huffman sub-phase clocks (ids reuse 0,31 and 14-17: run on Huffman-winning kinds; LZ refine ids overlap -> read only the deltas shown)"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptive_compression_b200 import engine, _lib as L
lib = engine.require_cuda()
n = 64 << 20
buf = (C.c_ulonglong * 32)()
for k, name in ((0, 'csv'), (3, 'lowcard'), (6, 'text')):
    t = engine.synth(n, 0, kind_mask=1 << k)
    engine.compress_device(t, 4096, mask=0b1000)  # Huffman only: ids 14-17 are not polluted by the LZ refine
    lib.ambc_phase_read(buf, 1)
    engine.compress_device(t, 4096, mask=0b1000)
    lib.ambc_phase_read(buf, 1)
    v = list(buf); nch = n // 4096
    print(name, "build: ranksort=%.0f merge=%.0f | lengths+emit: first_order=%.0f table+bitcount+scan=%.0f pack=%.0f copy=%.0f | huffman total=%.0f" %
          (v[31] / nch, v[0] / nch, v[14] / nch, v[15] / nch, v[16] / nch, v[17] / nch, v[12] / nch), flush=True)
