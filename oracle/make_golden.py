"""Generate tests/golden/*.json by running the UNMODIFIED reference
(/root/reference, via oracle/ref_harness.py) on tests/inputs.py.  TEST
INFRASTRUCTURE; runs only in the build container.  Re-run with

    python oracle/make_golden.py [codec|decode|container|marker|gates ...]

The reference ships no golden vectors of its own; these files
are the pin for oracle/ambc_oracle.c and, through it, for the CUDA path."""
import hashlib
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import inputs  # noqa: E402
import ref_harness as R  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sha(b):
    return hashlib.sha256(b).hexdigest()


def dump(name, obj):
    with open(os.path.join(GOLD, name), "w") as f:
        json.dump(obj, f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote", name, flush=True)


def gen_codec():
    rows = []
    for name, data in inputs.codec_cases():
        t0 = time.time()
        row = {"name": name, "n": len(data), "sha256": sha(data), "methods": {}}
        for mid in (1, 2, 3, 4):
            p, err = R.method_compress(mid, data)
            ent = {"should_use": R.should_use(mid, data)}
            if p is None:
                ent["error"] = err
            else:
                ent["len"] = len(p)
                ent["sha256"] = sha(p)
                if len(p) <= 96:
                    ent["hex"] = p.hex()
                # round trip through the reference decoder
                d, derr = R.method_decompress(mid, p, len(data))
                ent["roundtrip"] = (d == data) if d is not None else derr
            row["methods"][str(mid)] = ent
        rows.append(row)
        print("codec", name, "%.1fs" % (time.time() - t0), flush=True)
    dump("codec_kat.json", rows)


def gen_decode():
    rows = []
    cases = list(inputs.malformed_payloads())
    # derived from real payloads: truncated / corrupted streams
    S = inputs.survey_inputs()
    hp, _ = R.method_compress(3, S["A"])
    cases.append(("huff_trunc_bits", 3, hp[:len(hp) - 20], len(S["A"])))
    cases.append(("huff_short_orig", 3, hp, 100))
    cases.append(("huff_long_orig", 3, hp, 600))
    k = hp[0]
    dup = bytes([k + 1]) + hp[1:1 + 5 * k] + hp[1:6] + hp[1 + 5 * k:]
    cases.append(("huff_dup_entry", 3, dup, len(S["A"])))
    nb = bytearray(hp)
    nb[1 + 5 * k:5 + 5 * k] = (10 ** 6).to_bytes(4, "little")
    cases.append(("huff_numbits_big", 3, bytes(nb), len(S["A"])))
    lp, _ = R.method_compress(2, S["A"])
    cases.append(("lz_trunc", 2, lp[:len(lp) - 3], len(S["A"])))
    cases.append(("lz_short_orig", 2, lp, 77))
    cases.append(("lz_long_orig", 2, lp, 1000))
    rp, _ = R.method_compress(1, S["D"])
    cases.append(("rle_short_orig", 1, rp, 2500))
    cases.append(("rle_long_orig", 1, rp, 3300))
    for name, mid, payload, orig in cases:
        d, err = R.method_decompress(mid, payload, orig)
        rows.append({"name": name, "method": mid, "payload": payload.hex(), "orig_len": orig,
                     "out": None if d is None else d.hex(), "error": err})
    dump("decode_kat.json", rows)


def gen_container():
    rows = []
    with tempfile.TemporaryDirectory() as td:
        for name, data, cfg in inputs.container_cases():
            t0 = time.time()
            out, stats, pm = R.compress_bytes(data, td, **cfg)
            row = {"name": name, "n": len(data), "sha256": sha(data), "cfg": {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()},
                   "ambc_len": len(out), "ambc_sha256": sha(out), "stored_verbatim": pm is None and out == data,
                   "packages": pm}
            if len(out) <= 400:
                row["ambc_hex"] = out.hex()
            if pm is not None:
                row["header_hex"] = out[:out[5] | (out[6] << 8)].hex()
                back = R.decompress_bytes(out, td)
                row["ref_roundtrip"] = back == data
            cs = stats["chunk_stats"]
            row["stats"] = {"original_size": stats["original_size"], "compressed_size": stats["compressed_size"],
                            "ratio": stats["ratio"], "percent_reduction": stats["percent_reduction"],
                            "overhead_bytes": stats["overhead_bytes"],
                            "compression_efficiency": stats["compression_efficiency"],
                            "chunk_stats": {k: ({str(a): b for a, b in v.items()} if isinstance(v, dict) else v)
                                            for k, v in cs.items()}}
            rows.append(row)
            print("container", name, len(data), "->", len(out), "%.1fs" % (time.time() - t0), flush=True)
    dump("container_kat.json", rows)


def gen_container_dyn():
    """multi-candidate (dynamic chunk size) containers -> container_dyn_kat.json"""
    rows = []
    with tempfile.TemporaryDirectory() as td:
        for name, data, cfg in inputs.dynamic_cases():
            t0 = time.time()
            out, stats, pm = R.compress_bytes(data, td, **cfg)
            row = {"name": name, "n": len(data), "sha256": sha(data), "cfg": {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()},
                   "ambc_len": len(out), "ambc_sha256": sha(out), "stored_verbatim": pm is None and out == data,
                   "packages": pm}
            if pm is not None:
                back = R.decompress_bytes(out, td)
                row["ref_roundtrip"] = back == data
            cs = stats["chunk_stats"]
            row["stats"] = {"original_size": stats["original_size"], "compressed_size": stats["compressed_size"],
                            "ratio": stats["ratio"], "percent_reduction": stats["percent_reduction"],
                            "overhead_bytes": stats["overhead_bytes"],
                            "compression_efficiency": stats["compression_efficiency"],
                            "chunk_stats": {k: ({str(a): b for a, b in v.items()} if isinstance(v, dict) else v)
                                            for k, v in cs.items()}}
            rows.append(row)
            print("container_dyn", name, len(data), "->", len(out), "%.1fs" % (time.time() - t0), flush=True)
    dump("container_dyn_kat.json", rows)


def gen_marker():
    rows = []
    for name, data, max_len, sample in inputs.marker_cases():
        t0 = time.time()
        b, L = R.find_marker(data, max_len, sample)
        rows.append({"name": name, "n": len(data), "sha256": sha(data), "max_len": max_len, "sample_size": sample,
                     "marker": None if b is None else b.hex(), "length": L})
        print("marker", name, L, "%.1fs" % (time.time() - t0), flush=True)
    dump("marker_kat.json", rows)


def gen_gates():
    """should_use of all four methods on many short seeded inputs (cheap: no LZ trial)."""
    rows = []
    for i in range(inputs.N_GATE_CASES):
        k, n, frac, data = inputs.gate_case(i)
        rows.append({"i": i, "kind": k, "n": n, "frac": frac, "sha256": sha(data),
                     "gates": [R.should_use(m, data) for m in (1, 2, 3, 4)]})
    dump("gates_kat.json", rows)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    os.chdir(tempfile.gettempdir())
    which = sys.argv[1:] or ["codec", "decode", "marker", "gates", "container"]
    for w in which:
        globals()["gen_" + w]()
