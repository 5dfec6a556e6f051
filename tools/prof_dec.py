"""dev tool: run the decoders on one corpus kind (for ncu)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptive_compression_b200 import engine
kind = int(sys.argv[1]); mib = int(sys.argv[2])
t = engine.synth(mib << 20, 0, kind_mask=1 << kind)
o = engine.compress_device(t, 4096)
for _ in range(3):
    out, st = engine.decompress_device(o.body, mib << 20)
print(st, o.usage)
