"""dev tool: decode of a body that is one raw package (chunk 16384: no native method eligible) and of the strict
interleaved case"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptive_compression_b200 import engine
n = (1 << 30)
t = engine.synth(n, 0, kind_mask=1 << 1)
o = engine.compress_device(t, 16384)
dec = torch.empty(n, dtype=torch.uint8, device="cuda")
best = 1e9
for _ in range(4):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); d, st = engine.decompress_device(o.body, n, out=dec); b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
assert torch.equal(d, t) and st == [0, 0]
print("raw body, 1 GiB: decompress %.1f GB/s (%.2f ms incl. the GPU package index)" % (n / best / 1e6, best))
