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


def _py_walk(body, orig_size, mb=4, marker=b"\xff\xff\x00\x00", known=(1, 2, 3, 4, 255)):
    """plain restatement of adaptive_compressor.py:396-454's package walk -> ('ok', entries, covered) | ('marker',)"""
    pos = o = 0
    hdr = mb + 14
    ent = []
    while pos < len(body):
        if pos + hdr > len(body):
            break
        if body[pos:pos + mb] != marker:
            return ("marker",)
        t = body[pos + mb]
        orig = int.from_bytes(body[pos + mb + 6:pos + mb + 10], "little")
        comp = int.from_bytes(body[pos + mb + 10:pos + mb + 14], "little")
        pos += hdr
        if t == 0 or pos + comp > len(body):
            break
        kn = t in known
        nominal = comp if not kn else (orig if t == 255 else (0 if comp == 0 else (min(comp, orig) if t == 4 else orig)))
        emit = min(nominal, max(0, orig_size - o))
        if emit:
            if not kn or t == 255:
                done = 0
                while done < emit:
                    piece = min(65536, emit - done)
                    have = max(0, comp - done)
                    ent.append((pos + done, o + done, min(have, piece), piece, 255, piece))
                    done += piece
            else:
                ent.append((pos, o, comp, orig, t, emit))
        o += nominal
        pos += comp
        if o >= orig_size:
            break
    return ("ok", ent, min(o, orig_size))


def test_index_host_helper_threads_match_serial_walk():
    """the speculative helper threads of the host walk (large bodies) never change its result: well-formed,
    corrupted and truncated bodies, marker bytes and fake headers inside payloads, 2..8 helpers"""
    from adaptive_compression_b200 import engine, _lib as L
    lib = L.lib()
    r = np.random.RandomState(11)
    mk = b"\xff\xff\x00\x00"
    fake = mk + bytes([2, 0]) + (300).to_bytes(4, "little") * 2 + (40).to_bytes(4, "little")
    parts, total = [], 0
    for i in range(1500):
        n = int(r.randint(1, 3000))
        pay = bytearray(r.randint(0, 256, size=n).astype(np.uint8).tobytes())
        if i % 7 == 0 and n > 60:
            k = int(r.randint(0, n - 30)); pay[k:k + len(fake)] = fake[:max(0, min(len(fake), n - k))]
        t = int(r.choice([255, 255, 255, 9]))
        parts.append(mk + bytes([t, 0]) + n.to_bytes(4, "little") * 2 + n.to_bytes(4, "little") + bytes(pay))
        total += n
    body = b"".join(parts) + mk + b"\x00" * 12
    variants = [(body, total), (body, total // 2), (body[:len(body) // 2 + 17], total), (body, total + 1000)]
    for _ in range(12):
        bad = bytearray(body); bad[int(r.randint(len(bad)))] ^= 1 << int(r.randint(8)); variants.append((bytes(bad), total))
    # an END package and a marker mismatch in the middle
    cut = sum(len(p) for p in parts[:700])
    variants.append((body[:cut] + mk + b"\x00" * 12 + body[cut:], total))
    bad = bytearray(body); bad[cut] ^= 0xFF; variants.append((bytes(bad), total))
    try:
        for threads in (2, 3, 8):
            lib.ambc_set_walk_threads(1024, threads)
            for b, osz in variants:
                want = _py_walk(b, osz)
                try:
                    table, cov = engine.index_host(np.frombuffer(b, dtype=np.uint8), osz)
                    got = ("ok", [tuple(int(e[k]) for k in ("src_off", "dst_off", "comp_len", "orig_len", "type", "out_len")) for e in table], cov)
                except ValueError:
                    got = ("marker",)
                assert got == want, (threads, len(b), osz, got[0], want[0])
    finally:
        lib.ambc_set_walk_threads(32 << 20, 0)


def test_shard_place_c_abi_matches_python_fold():
    """ambc_shard_place (host-only entry of the C-ABI) == distributed.fold_placement, and the offsets of
    ranks inside the raw tail == what shard_fragment computes"""
    from adaptive_compression_b200 import distributed as D
    r = np.random.RandomState(5)
    for _ in range(200):
        world = int(r.randint(1, 9))
        chunk = int(r.choice([1024, 4096]))
        per = int(r.randint(1, 50))
        first_chunks = [k * per for k in range(world)]
        recs = []
        for k in range(world):
            fr = -1 if r.rand() < 0.7 else first_chunks[k] + int(r.randint(0, per))
            recs.append((int(r.randint(0, per * chunk)), fr))
        want = D.fold_placement(recs)
        got = D.fold_placement_native(recs, [c * chunk for c in first_chunks], chunk)
        assert [s for _, s in got] == [s for _, s in want]
        g = next((fr for _, fr in recs if fr >= 0), None)
        for k, ((off, state), (woff, _)) in enumerate(zip(got, want)):
            if state != "in_raw_tail":
                assert off == woff
            else:
                packed_total = sum(nb for nb, _ in recs[:next(i for i, (_, fr) in enumerate(recs) if fr >= 0) + 1])
                assert off == packed_total + 18 + (first_chunks[k] * chunk - g * chunk)


def test_history_record_matches_the_analyzer_schema(tmp_path, golden):
    """main.py appends the record the reference's analyzer writes (main.py:184-194,
    compression_analyzer.py:30-72): same keys, same size labels, replace-by-filename in place.  The schema and the
    (size, label) pairs come from the reference's own compression_results/compression_history.json."""
    import json
    import main as M
    fx = golden["history_schema"]
    for size, label in fx["size_labels"]:
        assert M._format_file_size(size) == label
    assert M._format_file_size(0) == "0 B" and M._format_file_size(512) == "512.0 B"
    stats = {"original_size": 40044, "compressed_size": 30000, "ratio": 0.7492, "percent_reduction": 25.08,
             "elapsed_time": 0.01, "throughput_mb_per_sec": 3.8, "overhead_bytes": 52, "compression_efficiency": 0.9,
             "chunk_stats": {"total_chunks": 3, "compressed_chunks": 2, "raw_chunks": 1, "method_usage": {1: 0, 2: 1, 3: 1, 4: 0, 255: 0},
                             "bytes_saved": 10000, "original_size": 40044, "compressed_size_without_overhead": 29948,
                             "overhead_bytes": 52}}
    d = str(tmp_path / "compression_results")
    path = M._append_history("/some/dir/a.log", stats, d)
    M._append_history("/x/b.bin", stats, d)
    recs = json.load(open(path))
    assert [r["filename"] for r in recs] == ["a.log", "b.bin"]
    assert sorted(recs[0].keys()) == fx["record_keys"]
    assert sorted(recs[0]["chunk_stats"].keys()) == fx["chunk_stats_keys"]
    assert all(isinstance(k, str) for k in recs[0]["chunk_stats"]["method_usage"])
    assert recs[0]["extension"] == ".log" and recs[0]["filename_no_ext"] == "a" and recs[0]["size_label"] == "39.1 KB"
    # the same file again: replaced in place (index 0), not appended
    stats2 = dict(stats, compressed_size=111)
    M._append_history("/other/a.log", stats2, d)
    recs = json.load(open(path))
    assert [r["filename"] for r in recs] == ["a.log", "b.bin"] and recs[0]["compressed_size"] == 111
    # a history with duplicates (older files of the reference have them) is reduced to the latest record per name
    dup = recs + [dict(recs[1], timestamp=recs[1]["timestamp"] - 100, compressed_size=5)]
    json.dump(dup, open(path, "w"))
    M._append_history("/x/c", stats, d)
    recs = json.load(open(path))
    assert [r["filename"] for r in recs] == ["a.log", "b.bin", "c"] and recs[1]["compressed_size"] == 30000
    assert recs[2]["extension"] == "unknown"


def test_deflate_plugin_surface():
    """DeflateCompression (advanced_compression.py:71-107) exists with the reference's id and is opt-in in the
    facade: methods=[..., 5] makes type 5 a known (decodable) method without entering the chunk trial"""
    from adaptive_compression_b200 import _lib as L
    from adaptive_compression_b200.adaptive_compressor import AdaptiveCompressor
    from adaptive_compression_b200.compression_methods import DeflateCompression
    assert DeflateCompression().type_id == 5 == L.DEFLATE
    assert DeflateCompression().compress(b"") == b"" and DeflateCompression().decompress(b"", 9) == b""
    assert not DeflateCompression().should_use(bytes(63))
    c = AdaptiveCompressor(chunk_size=4096, methods=[1, 2, 3, 5])
    assert [m.type_id for m in c.compression_methods] == [1, 2, 3, 5, 255]
    assert c._method_mask() == (1 << 1) | (1 << 2) | (1 << 3)
    assert [m.type_id for m in AdaptiveCompressor(chunk_size=4096).compression_methods] == [1, 2, 3, 4, 255]
