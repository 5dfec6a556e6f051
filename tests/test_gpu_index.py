"""GPU package index (ambc_index_dev) == host walk (ambc_index_host), entry for entry, on well-formed
and malformed bodies: truncations at every kind of boundary, marker corruption, early END, unknown
types, marker bytes inside payloads, short / long orig_size, raw packages split into 64 KiB pieces."""
import numpy as np
import pytest

import inputs
import oracle as O

pytestmark = pytest.mark.gpu

MARKER = b"\xff\xff\x00\x00"


@pytest.fixture(scope="module")
def eng():
    from adaptive_compression_b200 import engine
    engine.require_cuda()
    return engine


def both(eng, body, orig_size, marker=MARKER, known=None):
    """-> ('ok', table rows, covered) or ('marker',) from each implementation"""
    import torch
    from adaptive_compression_b200 import _lib as L
    known = L.NATIVE_MASK if known is None else known
    b = np.frombuffer(bytes(body), dtype=np.uint8)
    res = []
    for impl in ("host", "dev"):
        try:
            if impl == "host":
                table, cov = eng.index_host(b, orig_size, marker, known)
                rows = [tuple(int(e[k]) for k in ("src_off", "dst_off", "comp_len", "orig_len", "type", "out_len")) for e in table]
            else:
                t = torch.from_numpy(b.copy()).to("cuda") if b.size else torch.empty(0, dtype=torch.uint8, device="cuda")
                tt, ne, cov = eng.index_device(t, orig_size, marker, known)
                raw = tt.cpu().numpy()[:ne * 32]
                dt = np.dtype([("src_off", "<u8"), ("dst_off", "<u8"), ("comp_len", "<u4"), ("orig_len", "<u4"),
                               ("type", "<u4"), ("out_len", "<u4")])
                rows = [tuple(int(e[k]) for k in dt.names) for e in raw.view(dt)]
            res.append(("ok", rows, int(cov)))
        except ValueError as e:
            assert "Marker mismatch" in str(e)
            res.append(("marker",))
    return res


def check(eng, body, orig_size, **kw):
    h, d = both(eng, body, orig_size, **kw)
    assert h == d, (len(body), orig_size, h[0], d[0], (h[1][:3], d[1][:3]) if h[0] == d[0] == "ok" else None)
    return h


def pkg(t, orig, payload, marker=MARKER, comp=None):
    c = len(payload) if comp is None else comp
    return marker + bytes([t, 0]) + int(orig).to_bytes(4, "little") * 2 + int(c).to_bytes(4, "little") + payload


END = MARKER + b"\x00" * 12


def test_wellformed_bodies(eng):
    for seed, chunk, pcr in [(1, 4096, False), (2, 1024, True), (3, 512, False), (4, 4096, True)]:
        data = inputs.mixed_file(9, chunk, 300 + seed) + inputs.rand(chunk, seed) + inputs.text(100, seed)
        body, pm = O.compress_body(data, chunk, per_chunk_raw=pcr)
        r = check(eng, body, len(data))
        assert r[0] == "ok" and r[2] == len(data)
        for osz in (0, 1, 4095, 4096, 4097, len(data) - 1, len(data) + 500):
            check(eng, body, osz)


def test_truncations_and_corruption(eng):
    data = inputs.mixed_file(5, 2048, 77) + inputs.text(300, 78)
    body, pm = O.compress_body(data, 2048)
    bounds, pos = [], 0
    for t, orig, comp in pm:
        bounds.append(pos); pos += 18 + comp
    cuts = set()
    for b in bounds + [pos, len(body)]:
        for d in (-20, -5, -1, 0, 1, 3, 4, 5, 17, 18, 19, 40):
            if 0 <= b + d <= len(body):
                cuts.add(b + d)
    for cut in sorted(cuts):
        check(eng, body[:cut], len(data))
    r = np.random.RandomState(5)
    for _ in range(60):  # random single-byte corruption anywhere (headers, markers, payloads)
        bad = bytearray(body)
        bad[r.randint(len(bad))] ^= 1 << r.randint(8)
        check(eng, bad, len(data))
    for b in bounds[1:]:  # every package's marker and length fields
        for off in (0, 3, 4, 6, 10, 13, 14, 17):
            bad = bytearray(body); bad[b + off] ^= 0x5A
            check(eng, bad, len(data))


def test_marker_bytes_inside_payloads_and_odd_packages(eng):
    fake = MARKER + bytes([3, 0]) + (100).to_bytes(4, "little") * 2 + (7).to_bytes(4, "little")  # looks like a header
    raw1 = b"abc" + fake + b"xyz" * 10 + MARKER + MARKER
    body = pkg(255, len(raw1), raw1) + pkg(1, 12, b"A\x06B\x06") + pkg(9, 30, b"unknown-type-payload") + \
        pkg(255, 50, b"short raw, padded") + pkg(4, 10, b"") + pkg(2, 5, b"\x00a\x00b\x00c\x00d\x00e") + END
    for osz in (0, 10, len(raw1), len(raw1) + 12, 200, 1000):
        for known in (None, 0b110, 0b1111111110):
            check(eng, body, osz, known=known)
    # END in the middle, data after it; a body that is only END; an empty body; shorter than a header
    check(eng, pkg(1, 4, b"Z\x04") + END + pkg(1, 4, b"Y\x04") + END, 100)
    check(eng, END, 0); check(eng, END, 10); check(eng, b"", 10); check(eng, MARKER + b"\x01", 10)
    # body that does not start with the marker
    check(eng, b"\x00" + pkg(1, 4, b"Z\x04") + END, 10)
    # chain lands in the middle of nowhere / exactly at the end without END
    check(eng, pkg(1, 4, b"Z\x04") + b"garbage-not-a-marker-but-long-enough", 10)
    check(eng, pkg(1, 4, b"Z\x04"), 10)
    check(eng, pkg(1, 4, b"Z\x04") + b"\xff\xff", 10)
    # other marker lengths
    for mk in (b"\xf8", b"\xab\xcd", b"\x01\x02\x03"):
        b2 = pkg(1, 4, b"Z\x04", marker=mk) + pkg(255, 6, b"qwerty" + mk, marker=mk) + mk + b"\x00" * 12
        check(eng, b2, 10, marker=mk)
        check(eng, b2[:-5], 10, marker=mk)


def test_large_raw_package_is_split(eng):
    big = inputs.rand(300000, 9)
    body = pkg(1, 4, b"Z\x04") + pkg(255, len(big), big) + pkg(1, 4, b"Y\x04") + END
    r = check(eng, body, 4 + len(big) + 4)
    assert r[0] == "ok" and len(r[1]) == 1 + 5 + 1
    check(eng, body, 4 + 100000)
    check(eng, body[:200000], 4 + len(big) + 4)


def test_decode_with_gpu_index_at_size(eng):
    import torch
    n = 96 << 20
    t = eng.synth(n, 0, kind_mask=0b1111111)  # with random segments: rest-of-file-raw tail, 64 KiB pieces
    o = eng.compress_device(t, 4096)
    out, status = eng.decompress_device(o.body, n, gpu_index=True)
    assert status == [0, 0] and torch.equal(out, t)
    t = eng.synth(n, 5 * 65536)
    o = eng.compress_device(t, 4096)
    tt, ne, cov = eng.index_device(o.body, n)
    table, cov_h = eng.index_host(o.body.cpu().numpy(), n)
    assert ne == len(table) and cov == cov_h == n
    assert bytes(tt.cpu().numpy()[:ne * 32]) == table.tobytes()
