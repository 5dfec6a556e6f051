"""dev tool: where the time of the dynamic (multi-candidate) mode goes: launches with their durations (run under
ncu --metrics gpu__time_duration.sum) or plain wall time"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptive_compression_b200 import engine
from adaptive_compression_b200.adaptive_compressor import REFERENCE_CANDIDATES
engine.require_cuda()
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
t = engine.synth(mib << 20, 0)
engine.compress_dynamic_device(t[:1 << 20], REFERENCE_CANDIDATES)
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    o = engine.compress_dynamic_device(t, REFERENCE_CANDIDATES)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("%d MiB: %.1f ms = %.2f GB/s" % (mib, dt * 1e3, (mib << 20) / dt / 1e9), flush=True)
