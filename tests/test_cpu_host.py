"""CPU-only: the C-ABI library loads and exports every symbol include/ambc.h declares, the host
index walk agrees with the oracle's package walk, host-side facade logic (header, CLI parsing,
shard placement) works without a GPU."""
import os
import re
import struct

import numpy as np
import pytest

import inputs
import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from adaptive_compression_b200 import _lib
    lib = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "ambc.h")).read()
    declared = set(re.findall(r"\b(ambc_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert set(_lib.exported_symbols()) == declared
    assert lib.ambc_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from adaptive_compression_b200 import AdaptiveCompressor, RLECompression
    with pytest.raises(RuntimeError):
        RLECompression().compress(b"aaaaabbbbb")
    with pytest.raises(RuntimeError):
        AdaptiveCompressor().compress(__file__, "/tmp/should_not_exist.ambc")


def test_index_host_matches_oracle_walk():
    from adaptive_compression_b200 import engine
    for seed, chunk, pcr in [(1, 4096, False), (2, 1024, True), (3, 512, False)]:
        data = inputs.mixed_file(9, chunk, 300 + seed) + inputs.rand(chunk, seed) + inputs.text(100, seed)
        body, pm = O.compress_body(data, chunk, per_chunk_raw=pcr)
        table, covered = engine.index_host(np.frombuffer(body, dtype=np.uint8), len(data))
        assert covered == len(data)
        # entries tile the output exactly, raw packages in pieces of at most 64 KiB
        pos = 0
        for e in table:
            assert e["dst_off"] == pos and e["out_len"] > 0
            pos += int(e["out_len"])
        assert pos == len(data)
        kinds = [int(e["type"]) for e in table if e["type"] != 255]
        assert kinds == [p[0] for p in pm if p[0] != 255]


def test_index_host_errors_and_truncation():
    from adaptive_compression_b200 import engine
    data = inputs.mixed_file(4, 4096, 77)
    body, pm = O.compress_body(data, 4096)
    bad = bytearray(body)
    off = 18 + pm[0][2]
    bad[off] ^= 0xFF  # second package's marker
    with pytest.raises(ValueError, match="Marker mismatch"):
        engine.index_host(np.frombuffer(bytes(bad), dtype=np.uint8), len(data))
    cut = np.frombuffer(body[:off + 18 + 5], dtype=np.uint8)  # payload of package 2 truncated -> stop
    table, covered = engine.index_host(cut, len(data))
    assert len(table) == 1 and covered == 4096
    table, covered = engine.index_host(np.frombuffer(body, dtype=np.uint8), 5000)  # orig_size smaller
    assert covered == 5000 and int(table[-1]["out_len"]) == 5000 - 4096


def test_header_layout_matches_reference_bytes(golden):
    from adaptive_compression_b200 import AdaptiveCompressor
    c = AdaptiveCompressor()
    row = next(r for r in golden["container_kat"] if r["name"] == "H1")
    hdr = bytes.fromhex(row["header_hex"])
    parsed = c._parse_header(hdr)
    assert parsed["original_size"] == row["n"] and parsed["header_size"] == 47
    assert parsed["marker_bytes"] == b"\xff\xff\x00\x00" and parsed["marker_length"] == 32
    rebuilt = c._build_header(parsed["marker_bytes"], 32, parsed["checksum"], parsed["original_size"],
                              parsed["compressed_size"])
    assert rebuilt == hdr
    with pytest.raises(ValueError, match="Magic mismatch"):
        c._parse_header(b"XXXX" + hdr[4:])
    with pytest.raises(ValueError, match="Unsupported version: 3"):
        c._parse_header(hdr[:4] + b"\x03" + hdr[5:])
    c._init_marker(b"\xe0", 3)
    assert (c.marker_bytes_aligned, c.marker_pattern, c.marker_byte_length) == (b"\xe0", "111", 1)
    c._init_marker(b"\xab\xc0", 10)
    assert c.marker_bytes_aligned == b"\xab\xc0" and c.marker_pattern == "1010101111"


def test_cli_method_tokens():
    import main as cli
    assert cli.parse_methods("rle, Huffman,2") == [1, 3, 2]
    assert cli.parse_methods("deflate,zstd,lz4") == [5, 8, 9]
    with pytest.raises(ValueError):
        cli.parse_methods("snappy")


def test_fold_placement_monoid():
    from adaptive_compression_b200.distributed import fold_placement
    assert fold_placement([(100, -1), (50, -1), (70, -1)]) == [(0, "packed"), (100, "packed"), (150, "packed")]
    assert fold_placement([(100, -1), (30, 7), (70, -1), (5, 11)]) == [(0, "packed"), (100, "raw_starts_here"),
                                                                       (None, "in_raw_tail"), (None, "in_raw_tail")]


def test_synth_numpy_twin_is_seekable():
    import synth_ref
    a = synth_ref.corpus(300000, 0)
    b = synth_ref.corpus(1000, 123457)
    assert (a[123457:124457] == b).all()
    kinds = {synth_ref.seg_kind(synth_ref.DEFAULT_SEED, s, synth_ref.DEFAULT_KINDS) for s in range(200)}
    assert kinds == {0, 1, 2, 3, 4, 6}


def test_marker_carry_windows():
    """bits that precede each shard (and end the stream) from per-shard lengths and 4-byte tails == slicing
    the concatenated stream, also for shards shorter than 4 bytes and empty shards"""
    from adaptive_compression_b200 import distributed as D
    r = np.random.RandomState(3)
    for _ in range(200):
        shards = [bytes(r.randint(0, 256, size=int(r.choice([0, 1, 2, 3, 4, 5, 9, 40]))).astype(np.uint8)) for _ in range(int(r.randint(1, 6)))]
        whole = b"".join(shards)
        bits = "".join(format(b, "08b") for b in whole)
        per, end = D.carry_windows([len(s) for s in shards], [s[-4:] for s in shards])
        pos = 0
        for s, (w, before) in zip(shards, per):
            assert before == pos and w == whole[:pos][-4:]
            for want in (1, 7, 15, 23, 31):
                v, have = D._bits_of(w, before, want)
                assert have == min(want, 8 * pos) and (have == 0 or v == int(bits[8 * pos - have:8 * pos], 2))
            pos += len(s)
        v, have = D._bits_of(end[0], end[1], 31)
        assert end[1] == len(whole) and have == min(31, 8 * len(whole)) and (have == 0 or v == int(bits[-have:], 2))
