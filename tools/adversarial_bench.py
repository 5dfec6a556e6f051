"""dev tool: chunk-kernel time and parity on inputs built to hurt a bucket-indexed lazy match search: periodic
data, huge trigram classes without long matches, sawtooth, few-symbol noise, zero-padded records"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ctypes as C
import numpy as np, torch
import oracle as O
from adaptive_compression_b200 import engine
lib = engine.require_cuda()
lib.ambc_enable_timing(1)
mib = 32
n = mib << 20
r = np.random.RandomState(1)
def tile(pat): return np.resize(np.frombuffer(pat, dtype=np.uint8), n).copy()
cases = {
    "abab": tile(b"ab"),
    "period7": tile(b"abcdefg"),
    "abcX (class of 1024, 4th byte random)": (lambda a: (a.__setitem__(slice(3, None, 4), r.randint(0, 256, size=n // 4).astype(np.uint8)), a)[1])(tile(b"abc\0")),
    "abcXY (class of 800, two random bytes)": (lambda a: (a.__setitem__(slice(3, None, 5), r.randint(0, 256, size=len(a[3::5])).astype(np.uint8)), a.__setitem__(slice(4, None, 5), r.randint(0, 256, size=len(a[4::5])).astype(np.uint8)), a)[2])(tile(b"abc\0\0")),
    "two symbols random": r.randint(0, 2, size=n).astype(np.uint8) + 65,
    "four symbols random": r.randint(0, 4, size=n).astype(np.uint8) + 65,
    "sawtooth 0..255": tile(bytes(range(256))),
    "zero padded u64 counters": np.arange(n // 8, dtype="<u8").view(np.uint8),
    "text with 70% blanks": np.where(r.rand(n) < 0.7, 32, r.randint(97, 123, size=n)).astype(np.uint8),
    "random bytes with a repeated 40-byte block every 256": (lambda a, blk: (a.reshape(-1, 256).__setitem__((slice(None), slice(0, 40)), blk), a)[1])(r.randint(0, 256, size=n).astype(np.uint8), r.randint(0, 256, size=40).astype(np.uint8)),
}
for name, a in cases.items():
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for flags in (1,):  # per-chunk raw: every chunk is tried
        o = engine.compress_device(t, 4096, flags=flags)
        o = engine.compress_device(t, 4096, flags=flags)
        ms = (C.c_float * 4)(); lib.ambc_last_timing(ms)
        k = 1 << 20
        want, _ = O.compress_body(a[:k].tobytes(), 4096, per_chunk_raw=True) if "per_chunk_raw" in O.compress_body.__code__.co_varnames else (None, None)
        ok = "n/a"
        if want is not None:
            got = engine.compress_device(t[:k], 4096, flags=flags)
            ok = got.body.cpu().numpy().tobytes() == want
        print("%-58s k_select %7.2f ms per %d MiB = %6.1f GB/s  usage %s parity(1 MiB) %s" % (name, ms[0], mib, n / ms[0] / 1e6, o.usage, ok), flush=True)
