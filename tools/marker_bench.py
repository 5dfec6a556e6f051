"""dev tool: time of the marker search (ambc_find_marker_dev) on the mixed corpus and on random data"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptive_compression_b200 import engine
engine.require_cuda()
for name, mib, mask in (("mixed", 1024, 0b1011111), ("random", 256, 1 << 5), ("log", 1024, 2)):
    t = engine.synth(mib << 20, 0, kind_mask=mask)
    engine.find_marker_device(t[:1 << 20])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    best = 1e9
    for _ in range(3):
        ev[0].record(); m = engine.find_marker_device(t); ev[1].record(); torch.cuda.synchronize()
        best = min(best, ev[0].elapsed_time(ev[1]))
    print("%s %d MiB: marker %s, %.2f ms = %.1f GB/s" % (name, mib, m, best, (mib << 20) / best / 1e6), flush=True)
