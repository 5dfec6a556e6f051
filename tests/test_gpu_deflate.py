"""GPU tests of method id 5, DeflateCompression (advanced_compression.py:71-107).  The reference calls zlib, so the
oracle here is stock zlib itself: streams written on the GPU must be read by zlib.decompress, and the GPU inflater
must read what zlib.compress writes (every level, stored / fixed / dynamic blocks) with the reference's pad /
truncate / zeros-on-error conventions.  Reported separately from the chunk path."""
import zlib

import numpy as np
import pytest

import inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from adaptive_compression_b200 import engine
    engine.require_cuda()
    return engine


def _cases():
    r = np.random.RandomState(55)
    out = []
    for k in sorted(inputs.KINDS):
        for n in (1, 2, 3, 64, 100, 1000, 4096, 8192):
            out.append(inputs.make(k, n, 9000 + n))
    out.append(bytes(8192))
    out.append(bytes(r.randint(0, 256, size=8192).astype(np.uint8)))
    out.append(b"abc" * 2730)
    return out


def test_gpu_streams_are_read_by_stock_zlib(eng):
    datas = _cases()
    got = eng.codec_encode_batch(5, datas)
    bad = []
    total_in = total_out = 0
    for i, (d, p) in enumerate(zip(datas, got)):
        if not isinstance(p, bytes):
            bad.append((i, "code", p)); continue
        try:
            if zlib.decompress(p) != d:
                bad.append((i, "differs"))
        except zlib.error as e:
            bad.append((i, str(e)))
        total_in += len(d); total_out += len(p)
    assert not bad, bad[:10]
    assert total_out < 0.8 * total_in  # it does compress (text, csv, logs, runs ...)


def test_gpu_inflate_reads_stock_zlib(eng):
    """zlib.compress at every level (level 0: stored blocks, 1: fixed and dynamic, 6 / 9: dynamic), raw wbits
    variations, orig_len shorter and longer than the data"""
    datas = [d for d in _cases() if d]
    payloads, origs, want = [], [], []
    for i, d in enumerate(datas):
        for level in (0, 1, 6, 9):
            p = zlib.compress(d, level)
            for o in (len(d), max(1, len(d) - 5), len(d) + 7):
                payloads.append(p); origs.append(o); want.append(d[:o].ljust(o, b"\0"))
        co = zlib.compressobj(9, zlib.DEFLATED, 15, 9, zlib.Z_FIXED)  # fixed-Huffman blocks only
        p = co.compress(d) + co.flush()
        payloads.append(p + b"trailing bytes are ignored"); origs.append(len(d)); want.append(d)
    got = eng.codec_decode_batch(5, payloads, origs)
    bad = [(i, len(payloads[i]), origs[i]) for i, (g, w) in enumerate(zip(got, want)) if g != w]
    assert not bad, bad[:10]


def test_gpu_inflate_errors_give_zeros(eng):
    """where zlib.decompress raises, the reference returns original_length zero bytes (:93-97)"""
    d = inputs.make("text", 3000, 1)
    c = zlib.compress(d, 9)
    broken = [c[:-3], c[:len(c) // 2], c[:-1] + bytes([c[-1] ^ 1]), b"\x79" + c[1:], bytes([c[0], c[1] | 0x20]) + c[2:],
              b"\x78\x9c", b"\x78", c[:2] + b"\x07" + c[3:], b"\x78\x9c\x01\x05\x00\x00\x00abcde"]
    for b in broken:
        with pytest.raises(zlib.error):
            zlib.decompress(b)
    got = eng.codec_decode_batch(5, broken, [3000] * len(broken))
    assert all(g == bytes(3000) for g in got), [i for i, g in enumerate(got) if g != bytes(3000)]
    assert eng.codec_decode_batch(5, [b""], [10]) == [b""]


def test_plugin_object_and_type5_packages(eng):
    """DeflateCompression mirrors the reference's class; packages of type 5 (zlib.compress(level=9) payloads, as
    the reference writes them) decode inside a body when the method is known, and are copied through when not"""
    import torch
    from adaptive_compression_b200.compression_methods import DeflateCompression
    m = DeflateCompression()
    assert m.type_id == 5 and m.calculate_overhead() == 0
    d = inputs.make("log", 5000, 3)
    assert zlib.decompress(m.compress(d)) == d
    assert m.decompress(zlib.compress(d, 9), len(d)) == d
    assert m.compress(b"") == b"" and m.decompress(b"", 5) == b""
    assert m.should_use(d) and not m.should_use(d[:63])
    chunks = [inputs.make(k, 4096, 70 + i) for i, k in enumerate(("text", "csv", "runs", "lowcard", "binrec"))]
    body = bytearray()
    for ch in chunks:
        p = zlib.compress(ch, 9)
        body += b"\xff\xff\x00\x00" + bytes([5, 0]) + len(ch).to_bytes(4, "little") * 2 + len(p).to_bytes(4, "little") + p
    body += b"\xff\xff\x00\x00" + bytes(12)
    t = torch.frombuffer(bytearray(body), dtype=torch.uint8).cuda()
    n = sum(len(c) for c in chunks)
    for gpu_index in (True, False):
        dec, status = eng.decompress_device(t, n, known_mask=eng.method_mask([1, 2, 3, 4, 5]), gpu_index=gpu_index)
        assert status == [0, 0] and bytes(dec.cpu().numpy()) == b"".join(chunks)
    # unknown method: the payload bytes are copied through (adaptive_compressor.py:432-435)
    dec, status = eng.decompress_device(t, n, known_mask=eng.method_mask([1, 2, 3, 4]))
    p0 = zlib.compress(chunks[0], 9)
    assert bytes(dec.cpu().numpy())[:len(p0)] == p0


def test_gpu_inflate_mutated_streams_vs_zlib(eng):
    """valid streams with a flipped byte, a cut or an inserted byte: whatever stock zlib does -- data, or an
    exception that the reference turns into zeros (:93-97) -- the GPU inflater does the same"""
    r = np.random.RandomState(909)
    payloads, origs, want = [], [], []
    kinds = sorted(inputs.KINDS)
    for i in range(240):
        d = inputs.make(kinds[i % len(kinds)], int(r.choice([200, 1500, 4096])), 12000 + i)
        c = bytearray(zlib.compress(d, int(r.choice([1, 6, 9]))))
        what = r.randint(3)
        pos = int(r.randint(2, len(c)))
        if what == 0:
            c[pos] ^= 1 << int(r.randint(8))
        elif what == 1:
            del c[pos:]
        else:
            c.insert(pos, int(r.randint(256)))
        c = bytes(c)
        try:
            out = zlib.decompress(c)
            w = out[:len(d)].ljust(len(d), b"\0")
        except zlib.error:
            w = bytes(len(d))
        payloads.append(c); origs.append(len(d)); want.append(w)
    got = eng.codec_decode_batch(5, payloads, origs)
    bad = [(i, len(payloads[i])) for i, (g, w) in enumerate(zip(got, want)) if g != w]
    assert not bad, bad[:10]
