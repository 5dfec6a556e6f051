"""GPU: the reference-facing Python surface (AdaptiveCompressor, CompressionMethod plug-ins, CLI)
produces the reference's own files (golden sha256), stats and errors."""
import hashlib
import os
import subprocess
import sys

import pytest

import inputs
import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sha(b):
    return hashlib.sha256(b).hexdigest()


def _compressor(cfg):
    from adaptive_compression_b200 import AdaptiveCompressor
    return AdaptiveCompressor(chunk_size=cfg["chunk_size"], methods=cfg.get("method_ids"),
                              per_chunk_raw=bool(cfg.get("per_chunk_raw")),
                              use_marker_search=bool(cfg.get("found_marker")))


def test_files_equal_reference_files(golden, tmp_path):
    """.ambc written by the facade == the file the unmodified reference wrote (sha256), stats equal"""
    cases = {c[0]: c for c in inputs.container_cases()}
    for row in golden["container_kat"]:
        name, data, _ = cases[row["name"]]
        src, dst, back = tmp_path / "in.bin", tmp_path / "out.ambc", tmp_path / "back.bin"
        src.write_bytes(data)
        c = _compressor(row["cfg"])
        stats = c.compress(str(src), str(dst))
        out = dst.read_bytes()
        assert len(out) == row["ambc_len"] and sha(out) == row["ambc_sha256"], name
        g = row["stats"]
        for k in ("original_size", "compressed_size", "ratio", "percent_reduction", "overhead_bytes",
                  "compression_efficiency"):
            assert stats[k] == pytest.approx(g[k], rel=1e-12, abs=1e-12), (name, k, stats[k], g[k])
        gcs = g["chunk_stats"]
        for k in ("total_chunks", "compressed_chunks", "raw_chunks", "bytes_saved", "original_size",
                  "compressed_size_without_overhead", "overhead_bytes"):
            assert stats["chunk_stats"][k] == gcs[k], (name, k, stats["chunk_stats"][k], gcs[k])
        assert {str(k): v for k, v in stats["chunk_stats"]["method_usage"].items()} == gcs["method_usage"], name
        if not row["stored_verbatim"]:
            d = c.decompress(str(dst), str(back))
            assert back.read_bytes() == data, name
            assert d["decompressed_size"] == len(data) and d["compressed_size"] == len(out)
            assert O.decompress_file(out) == data  # and the oracle's reference decoder reads our file


def test_error_conventions(tmp_path):
    from adaptive_compression_b200 import AdaptiveCompressor
    data = inputs.mixed_file(4, 4096, 11)
    src, dst, back = tmp_path / "a", tmp_path / "a.ambc", tmp_path / "b"
    src.write_bytes(data)
    c = AdaptiveCompressor()
    c.compress(str(src), str(dst))
    good = bytearray(dst.read_bytes())
    bad = bytearray(good); bad[0] = ord("X")
    dst.write_bytes(bad)
    with pytest.raises(ValueError, match="Magic mismatch"):
        c.decompress(str(dst), str(back))
    bad = bytearray(good); bad[4] = 9
    dst.write_bytes(bad)
    with pytest.raises(ValueError, match="Unsupported version: 9"):
        c.decompress(str(dst), str(back))
    bad = bytearray(good); bad[47] ^= 1
    dst.write_bytes(bad)
    with pytest.raises(ValueError, match="Marker mismatch in chunk header."):
        c.decompress(str(dst), str(back))
    bad = bytearray(good); bad[47 + 18 + 30] ^= 0x55  # payload corruption -> checksum error after the write
    dst.write_bytes(bad)
    with pytest.raises(ValueError, match="Checksum mismatch"):
        c.decompress(str(dst), str(back))
    assert back.exists()
    # candidate lists the GPU search cannot grid (common divisor below 256 bytes) are refused loudly
    from adaptive_compression_b200._lib import AmbcError
    with pytest.raises(AmbcError):
        c2 = AdaptiveCompressor(); c2.CHUNK_SIZE_CANDIDATES = [4096, 1000]; c2.compress(str(src), str(dst))
    # the harness-style configuration of the reference selects the dynamic mode
    c3 = AdaptiveCompressor(); c3.CHUNK_SIZE_CANDIDATES = [8192, 4096]; c3.compress(str(src), str(dst))
    want, _, _ = O.compress_file(src.read_bytes(), (8192, 4096))
    assert dst.read_bytes() == want


def test_plugin_objects():
    from adaptive_compression_b200 import (DeltaCompression, DictionaryCompression, HuffmanCompression,
                                           NoCompression, RLECompression)
    A = inputs.survey_inputs()["A"]  # compression_methods.py:719 self-test input
    sizes = {}
    for cls in (RLECompression, DictionaryCompression, HuffmanCompression, DeltaCompression, NoCompression):
        m = cls()
        p = m.compress(A)
        sizes[m.type_id] = len(p)
        assert m.decompress(p, len(A)) == A
        assert m.should_use(A) in (True, False)
        assert m.calculate_overhead() == 0
    assert sizes == {1: 322, 2: 74, 3: 123, 4: 420, 255: 420}  # SURVEY.md §4 self-test sizes
    with pytest.raises(IndexError):
        HuffmanCompression().compress(b"z" * 300)
    with pytest.raises(ValueError):
        HuffmanCompression().compress(bytes(range(256)) * 4)
    assert RLECompression().compress(b"") == b"" and NoCompression().decompress(b"abc", 5) == b"abc\0\0"


def test_cli_roundtrip(tmp_path):
    data = inputs.mixed_file(6, 4096, 21) + inputs.text(500, 22)
    src, dst, back = tmp_path / "in.csv", tmp_path / "o.ambc", tmp_path / "b.csv"
    src.write_bytes(data)
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), "compress", str(src), str(dst), "--chunk-size",
                        "4096", "--disable-methods", "deflate,bzip2,lzma", "--no-history"], capture_output=True,
                       text=True, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Compression completed successfully." in r.stdout
    want, raw, _ = O.compress_file(data, 4096)
    assert dst.read_bytes() == want
    r = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), "decompress", str(dst), str(back)],
                       capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert r.returncode == 0 and back.read_bytes() == data, r.stdout + r.stderr
    r = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), "decompress", str(src), str(back)],
                       capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert r.returncode == 1 and "Error during decompression: Magic mismatch" in r.stdout
    r = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), "compress", str(src), str(dst), "--methods",
                        "rle", "--chunk-size", "1024", "--no-history"], capture_output=True, text=True, env=env,
                       cwd=str(tmp_path))
    assert r.returncode == 0
    want, raw, _ = O.compress_file(data, 1024, (1,))
    assert dst.read_bytes() == want


def test_host_buffer_abi(tmp_path):
    """ambc_compress_host / ambc_decompress_host (the e2e entry points) on numpy buffers"""
    import ctypes as C
    import numpy as np
    from adaptive_compression_b200 import _lib as L
    lib = L.lib()
    data = np.frombuffer(inputs.mixed_file(20, 4096, 31) + inputs.text(999, 32), dtype=np.uint8)
    bound = lib.ambc_compress_bound(data.size, 4096, 4)
    body = np.empty(bound, dtype=np.uint8)
    mt = np.empty(data.size // 4096 + 2, dtype=np.uint8)
    mc = np.empty(data.size // 4096 + 2, dtype=np.uint32)
    res = L.CompressResult()
    L.check(lib.ambc_compress_host(C.c_void_p(data.ctypes.data), data.size, 4096, L.NATIVE_MASK, 0, b"\xff\xff\x00\x00", 4,
                                   C.c_void_p(body.ctypes.data), bound, C.c_void_p(mt.ctypes.data),
                                   C.c_void_p(mc.ctypes.data), C.byref(res)))
    want, pm = O.compress_body(data.tobytes(), 4096)
    assert body[:res.body_len].tobytes() == want
    k = res.first_raw if res.first_raw >= 0 else res.n_chunks
    assert [int(x) for x in mt[:k]] == [p[0] for p in pm[:k]] and len(pm) == res.n_packages
    back = np.empty(data.size, dtype=np.uint8)
    st = (C.c_uint32 * 2)()
    L.check(lib.ambc_decompress_host(C.c_void_p(body.ctypes.data), res.body_len, b"\xff\xff\x00\x00", 4, L.NATIVE_MASK,
                                     C.c_void_p(back.ctypes.data), data.size, st))
    assert (back == data).all() and list(st) == [0, 0]


@pytest.mark.parametrize("kind_mask,chunk", [(0b1011111, 4096), (0b1111111, 4096), (0b1011111, 1024), (0b1111111, 3000)])
def test_host_buffer_abi_pipelined(kind_mask, chunk):
    """multi-piece pipelines of the host-buffer calls (64 MiB upload pieces overlapping k_select,
    16384-package decode pieces overlapping the host walk): same body as the device-resident path,
    bit-exact round trip; kind_mask with the random kind exercises the rest-of-file-raw rule"""
    import ctypes as C
    import numpy as np
    import torch
    from adaptive_compression_b200 import _lib as L
    from adaptive_compression_b200 import engine
    lib = engine.require_cuda()
    n = (150 << 20) + 12345
    t = engine.synth(n, 0, kind_mask=kind_mask)
    dev = engine.compress_device(t, chunk)
    want = dev.body.cpu().numpy()
    data = t.cpu().numpy()
    bound = lib.ambc_compress_bound(n, chunk, 4)
    body = np.empty(bound, dtype=np.uint8)
    res = L.CompressResult()
    L.check(lib.ambc_compress_host(C.c_void_p(data.ctypes.data), n, chunk, L.NATIVE_MASK, 0, b"\xff\xff\x00\x00", 4,
                                   C.c_void_p(body.ctypes.data), bound, None, None, C.byref(res)))
    assert res.body_len == dev.body_len and res.first_raw == dev.first_raw
    assert np.array_equal(body[:res.body_len], want)
    back = np.full(n + 100, 7, dtype=np.uint8)
    st = (C.c_uint32 * 2)()
    # orig_size 100 bytes larger than the data: the reference zero-pads (adaptive_compressor.py:447-449)
    L.check(lib.ambc_decompress_host(C.c_void_p(body.ctypes.data), res.body_len, b"\xff\xff\x00\x00", 4, L.NATIVE_MASK,
                                     C.c_void_p(back.ctypes.data), n + 100, st))
    assert np.array_equal(back[:n], data) and not back[n:].any() and list(st) == [0, 0]


def test_host_decompress_damaged_bodies():
    """ambc_decompress_host's pipeline (helper-thread walk, granular body upload, table ring) on damaged
    bodies of several upload granules: truncation, an END package and a marker mismatch mid-body, an
    unknown type byte, a short orig_size -- output and error equal to the oracle's decoder
    (adaptive_compressor.py:396-454)"""
    import ctypes as C
    import numpy as np
    from adaptive_compression_b200 import _lib as L
    from adaptive_compression_b200 import engine
    lib = engine.require_cuda()
    n = (96 << 20) + 777
    t = engine.synth(n, 3)
    dev = engine.compress_device(t, 4096)
    body = dev.body.cpu().numpy()[:dev.body_len].copy()
    assert body.size > (36 << 20)
    # package starts: walk the headers on the host
    starts, pos = [], 0
    while pos + 18 <= body.size and body[pos + 4] != 0:
        starts.append(pos)
        pos += 18 + int.from_bytes(body[pos + 14:pos + 18].tobytes(), "little")
    mid = starts[len(starts) // 2]
    late = starts[len(starts) * 7 // 8]
    variants = []
    variants.append(("truncated", body[:late + 5].copy(), n))
    v = body.copy(); v[mid:mid + 16] = np.frombuffer(b"\xff\xff\x00\x00" + b"\x00" * 12, dtype=np.uint8)
    variants.append(("end_mid", v, n))
    v = body.copy(); v[late + 1] ^= 0x40
    variants.append(("marker_mismatch", v, n))
    v = body.copy(); v[mid + 4] = 77
    variants.append(("unknown_type", v, n))
    variants.append(("short_orig", body.copy(), n // 3 + 11))
    variants.append(("long_orig", body.copy(), n + 5000))
    for name, b, osz in variants:
        try:
            want = O.decompress_body(b.tobytes(), osz)
        except ValueError:
            want = None
        back = np.full(osz, 9, dtype=np.uint8)
        st = (C.c_uint32 * 2)()
        rc = lib.ambc_decompress_host(C.c_void_p(b.ctypes.data), b.size, b"\xff\xff\x00\x00", 4, L.NATIVE_MASK,
                                      C.c_void_p(back.ctypes.data), osz, st)
        if want is None:
            assert rc == L.E_MARKER, (name, rc)
        else:
            assert rc == 0, (name, rc, lib.ambc_last_error())
            assert back.tobytes() == want, name


def test_dynamic_mode_files_equal_reference_files(golden, tmp_path):
    """several CHUNK_SIZE_CANDIDATES (the reference's default list and custom ones): the facade's file ==
    the file the unmodified reference wrote; stats equal; both decoders read it"""
    cases = {c[0]: c for c in inputs.dynamic_cases()}
    for row in golden["container_dyn_kat"]:
        name, data, _ = cases[row["name"]]
        src, dst, back = tmp_path / "in.bin", tmp_path / "out.ambc", tmp_path / "back.bin"
        src.write_bytes(data)
        c = _compressor(row["cfg"])
        stats = c.compress(str(src), str(dst))
        out = dst.read_bytes()
        assert len(out) == row["ambc_len"] and sha(out) == row["ambc_sha256"], name
        if "stats" in row:
            g = row["stats"]
            for k in ("original_size", "compressed_size", "ratio", "percent_reduction", "overhead_bytes",
                      "compression_efficiency"):
                assert stats[k] == pytest.approx(g[k], rel=1e-12, abs=1e-12), (name, k, stats[k], g[k])
            gcs = g["chunk_stats"]
            for k in ("total_chunks", "compressed_chunks", "raw_chunks", "bytes_saved", "original_size",
                      "compressed_size_without_overhead", "overhead_bytes"):
                assert stats["chunk_stats"][k] == gcs[k], (name, k, stats["chunk_stats"][k], gcs[k])
            assert {str(k): v for k, v in stats["chunk_stats"]["method_usage"].items()} == gcs["method_usage"], name
        c.decompress(str(dst), str(back))
        assert back.read_bytes() == data, name
        assert O.decompress_file(out) == data


def test_cli_dynamic_chunk_size(tmp_path):
    data = inputs.mixed_file(5, 4096, 91, ("text", "log", "runs")) + inputs.csv(900, 92)
    src, dst, back = tmp_path / "a.bin", tmp_path / "a.ambc", tmp_path / "a.out"
    src.write_bytes(data)
    env = dict(os.environ)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), "compress", str(src), str(dst), "--chunk-size",
                        "dynamic", "--no-history"], capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    want, raw, _ = O.compress_file(data, inputs.REF_DEFAULT_CANDIDATES)
    assert dst.read_bytes() == want
    r = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), "decompress", str(dst), str(back)],
                       capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert r.returncode == 0 and back.read_bytes() == data
