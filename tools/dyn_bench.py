"""dev tool: wall time of the dynamic (multi-candidate) mode on the mixed corpus"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptive_compression_b200 import engine
from adaptive_compression_b200.adaptive_compressor import REFERENCE_CANDIDATES
engine.require_cuda()
for mib, cands in ((16, REFERENCE_CANDIDATES), (64, REFERENCE_CANDIDATES), (64, (4096, 2048, 1024))):
    t = engine.synth(mib << 20, 0)
    engine.compress_dynamic_device(t[:1 << 20], cands)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    o = engine.compress_dynamic_device(t, cands)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    f = engine.compress_device(t, 4096)
    sizes = {}
    for ty, orig, comp in o.packages:
        sizes[orig] = sizes.get(orig, 0) + 1
    print("%d MiB cands %s: %.3f s = %.2f GB/s; body %.4f of input (fixed 4096: %.4f); packages by size %s" %
          (mib, cands[:3], dt, (mib << 20) / dt / 1e9, o.body_len / (mib << 20), f.body_len / (mib << 20),
           dict(sorted(sizes.items(), reverse=True)[:6])), flush=True)
    out, st = engine.decompress_device(o.body, mib << 20)
    assert torch.equal(out, t) and st == [0, 0]
