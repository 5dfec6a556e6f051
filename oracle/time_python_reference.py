"""BUILD-CONTAINER ONLY (needs /root/reference): time the UNMODIFIED Python reference on slices of the bench
corpus: 64 chunks (256 KiB at chunk 4096) on one core, and 8 processes over 8 disjoint
64-chunk slices.  Writes profiles/r02_python_reference_timing.json, which bench.py quotes inside cpu_baseline
(the reference cannot travel to the GPU box).  TEST INFRASTRUCTURE.

  python oracle/time_python_reference.py"""
import json
import multiprocessing as mp
import os
import platform
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness as R  # noqa: E402
import synth_ref  # noqa: E402

CHUNK, NCH = 4096, 64


def one_slice(k):
    data = synth_ref.corpus(NCH * CHUNK, k * NCH * CHUNK).tobytes()
    with tempfile.TemporaryDirectory() as d:
        t0 = time.perf_counter()
        out, stats, pm = R.compress_bytes(data, d, chunk_size=CHUNK)
        tc = time.perf_counter() - t0
        c = R.make_compressor(chunk_size=CHUNK)
        src, dst = os.path.join(d, "out.ambc"), os.path.join(d, "back.bin")
        t0 = time.perf_counter()
        R.quiet(c.decompress, src, dst)
        td = time.perf_counter() - t0
        assert open(dst, "rb").read() == data
    return tc, td, len(out)


def main():
    assert R.available(), "needs /root/reference"
    tc, td, sz = one_slice(0)
    n = NCH * CHUNK
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(8) as pool:
        res = pool.map(one_slice, range(8))
    wall8 = time.perf_counter() - t0
    out = {"what": "unmodified reference (/root/reference), candidates [4096], methods {1,2,3,4,255}, stdout discarded",
           "host": platform.processor() or platform.machine(), "cpus": os.cpu_count(), "python": platform.python_version(),
           "corpus": "bench corpus (oracle/synth_ref.corpus, seed 0xA3BC0001), slice k = chunks [64k, 64k+64)",
           "one_core": {"bytes": n, "compress_s": tc, "decompress_s": td, "compress_kb_s": n / tc / 1e3,
                        "decompress_kb_s": n / td / 1e3, "roundtrip_kb_s": n / (tc + td) / 1e3, "ambc_bytes": sz},
           "eight_processes": {"bytes": 8 * n, "wall_s": wall8, "roundtrip_kb_s": 8 * n / wall8 / 1e3,
                               "per_slice_compress_s": [r[0] for r in res]},
           "extrapolated_1GiB_one_core_hours": (1 << 30) / (n / (tc + td)) / 3600}
    path = os.path.join(ROOT, "profiles", "r02_python_reference_timing.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
