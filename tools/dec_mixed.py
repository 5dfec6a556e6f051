"""dev tool: decode kernels' time on the 1 GiB mixed bench body (device timers of the library), best of 5"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptive_compression_b200 import engine
lib = engine.require_cuda()
n = 1 << 30
t = engine.synth(n, 0)
o = engine.compress_device(t, 4096)
lib.ambc_enable_timing(1)
out = torch.empty(n, dtype=torch.uint8, device="cuda")
best = 1e9
for _ in range(6):
    d, st = engine.decompress_device(o.body, n, out=out)
    ms = (C.c_float * 4)(); lib.ambc_last_timing(ms)
    best = min(best, ms[3])
assert torch.equal(d, t)
print("decode kernels %.3f ms per GiB of the mixed corpus" % best, flush=True)
