"""ctypes front-end of the CPU oracle (oracle/ambc_oracle.c).  TEST
INFRASTRUCTURE: importable only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from the product."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

ERR_INDEX = -1
ERR_VALUE = -2
FIXED_MARKER = b"\xff\xff\x00\x00"  # adaptive_compressor.py:303-310
NATIVE = (1, 2, 3, 4)


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("ambc_oracle.c", "oracle_mt.c")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, lp, ip = C.c_void_p, C.c_void_p, C.c_void_p
        for name in ("rle", "lz", "huff", "delta"):
            f = getattr(L, "orc_%s_compress" % name); f.restype = C.c_long; f.argtypes = [u8p, C.c_long, u8p]
            f = getattr(L, "orc_%s_decompress" % name); f.restype = C.c_long; f.argtypes = [u8p, C.c_long, C.c_long, u8p]
            f = getattr(L, "orc_%s_should_use" % name); f.restype = C.c_int; f.argtypes = [u8p, C.c_long]
        L.orc_lz_compress_ex.restype = C.c_long; L.orc_lz_compress_ex.argtypes = [u8p, C.c_long, u8p, C.c_int]
        L.orc_raw_decompress.restype = C.c_long; L.orc_raw_decompress.argtypes = [u8p, C.c_long, C.c_long, u8p]
        L.orc_entropy.restype = C.c_double; L.orc_entropy.argtypes = [u8p, C.c_long]
        L.orc_set_lz_fast.argtypes = [C.c_int]
        L.orc_compress_body.restype = C.c_long
        L.orc_compress_body.argtypes = [u8p, C.c_long, lp, C.c_int, ip, C.c_int, u8p, C.c_int, C.c_int, u8p,
                                        ip, lp, lp, C.c_long, lp]
        L.orc_decompress_body.restype = C.c_long
        L.orc_decompress_body.argtypes = [u8p, C.c_long, C.c_long, u8p, C.c_int, ip, C.c_int, u8p]
        L.orc_find_marker.restype = C.c_int; L.orc_find_marker.argtypes = [u8p, C.c_long, C.c_int, u8p]
        L.orc_marker_sample.restype = C.c_long; L.orc_marker_sample.argtypes = [u8p, C.c_long, C.c_long, u8p]
        L.orc_md5.argtypes = [u8p, C.c_long, u8p]
        L.orc_compress_file.restype = C.c_long
        L.orc_compress_file.argtypes = [u8p, C.c_long, lp, C.c_int, ip, C.c_int, u8p, C.c_int, C.c_int, u8p, ip,
                                        ip, lp, lp, C.c_long, lp]
        L.orc_mt_compress.restype = C.c_long
        L.orc_mt_compress.argtypes = [u8p, C.c_long, C.c_long, ip, C.c_int, C.c_int, C.c_void_p, lp]
        L.orc_mt_decompress.restype = C.c_long
        L.orc_mt_decompress.argtypes = [C.c_void_p, lp, lp, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _in(data):
    a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
    if a.size == 0:
        a = np.zeros(1, dtype=np.uint8)[:0]
    return a


def _ptr(a):
    return a.ctypes.data if a.size else None


_NAMES = {1: "rle", 2: "lz", 3: "huff", 4: "delta"}


def set_lz_fast(on):
    lib().orc_set_lz_fast(1 if on else 0)


def compress(mid, data, lz_fast=False):
    """-> bytes, or ERR_INDEX / ERR_VALUE (the Python exception the reference raises)."""
    a = _in(data)
    n = a.size
    if mid == 255:
        return bytes(data)
    out = np.empty(4 * n + 2048, dtype=np.uint8)
    if mid == 2:
        r = lib().orc_lz_compress_ex(_ptr(a), n, out.ctypes.data, 1 if lz_fast else 0)
    else:
        r = getattr(lib(), "orc_%s_compress" % _NAMES[mid])(_ptr(a), n, out.ctypes.data)
    return r if r < 0 else out[:r].tobytes()


def decompress(mid, payload, orig_len):
    a = _in(payload)
    out = np.empty(max(orig_len, 0) + 512, dtype=np.uint8)
    if mid == 255:
        r = lib().orc_raw_decompress(_ptr(a), a.size, orig_len, out.ctypes.data)
    else:
        r = getattr(lib(), "orc_%s_decompress" % _NAMES[mid])(_ptr(a), a.size, orig_len, out.ctypes.data)
    return r if r < 0 else out[:r].tobytes()


def should_use(mid, data):
    a = _in(data)
    if mid == 255:
        return True
    return bool(getattr(lib(), "orc_%s_should_use" % _NAMES[mid])(_ptr(a), a.size))


def entropy(data):
    a = _in(data)
    return float(lib().orc_entropy(_ptr(a), a.size))


def marker_aligned(marker_bytes, marker_bits):
    """adaptive_compressor.py:196-219"""
    v = int.from_bytes(marker_bytes, "big") >> (8 * len(marker_bytes) - marker_bits) if marker_bits else 0
    nb = (marker_bits + 7) // 8
    return (v << (8 * nb - marker_bits)).to_bytes(nb, "big") if nb else b""


def compress_body(data, chunk=4096, methods=NATIVE, marker=FIXED_MARKER, per_chunk_raw=False, lz_fast=True):
    """-> (body bytes, [(type, orig, comp)])"""
    a = _in(data)
    n = a.size
    cands = np.array(sorted([chunk] if isinstance(chunk, int) else list(chunk), reverse=True), dtype=np.int64)
    meth = np.array([m for m in methods if m != 255], dtype=np.int32)
    mk = _in(marker)
    minc = int(cands.min())
    cap_p = n // max(minc, 1) + 4
    out = np.empty(n + cap_p * (mk.size + 14) + 64, dtype=np.uint8)
    mt = np.zeros(cap_p, dtype=np.int32); mo = np.zeros(cap_p, dtype=np.int64); mc = np.zeros(cap_p, dtype=np.int64)
    npk = C.c_long(0)
    set_lz_fast(lz_fast)
    r = lib().orc_compress_body(_ptr(a), n, cands.ctypes.data, cands.size, _ptr(meth), meth.size, _ptr(mk), mk.size,
                                1 if per_chunk_raw else 0, out.ctypes.data, mt.ctypes.data, mo.ctypes.data,
                                mc.ctypes.data, cap_p, C.byref(npk))
    set_lz_fast(False)
    k = npk.value
    return out[:r].tobytes(), list(zip(mt[:k].tolist(), mo[:k].tolist(), mc[:k].tolist()))


def decompress_body(body, orig_size, marker=FIXED_MARKER, known=(1, 2, 3, 4, 255)):
    a = _in(body)
    mk = _in(marker)
    kn = np.array(known, dtype=np.int32)
    out = np.empty(orig_size + 16, dtype=np.uint8)
    r = lib().orc_decompress_body(_ptr(a), a.size, orig_size, _ptr(mk), mk.size, kn.ctypes.data, kn.size, out.ctypes.data)
    if r == -3:
        raise ValueError("Marker mismatch in chunk header.")
    return out[:orig_size].tobytes()


def compress_file(data, chunk=4096, methods=NATIVE, marker=(FIXED_MARKER, 32), per_chunk_raw=False, lz_fast=True):
    """-> (.ambc bytes (or the input verbatim), stored_verbatim, [(type, orig, comp)])"""
    a = _in(data)
    n = a.size
    cands = np.array(sorted([chunk] if isinstance(chunk, int) else list(chunk), reverse=True), dtype=np.int64)
    meth = np.array([m for m in methods if m != 255], dtype=np.int32)
    mraw = _in(marker[0])
    minc = int(cands.min())
    cap_p = n // max(minc, 1) + 4
    out = np.empty(n + cap_p * 18 + 256, dtype=np.uint8)
    mt = np.zeros(cap_p, dtype=np.int32); mo = np.zeros(cap_p, dtype=np.int64); mc = np.zeros(cap_p, dtype=np.int64)
    npk = C.c_long(0); raw = C.c_int(0)
    set_lz_fast(lz_fast)
    r = lib().orc_compress_file(_ptr(a), n, cands.ctypes.data, cands.size, _ptr(meth), meth.size, _ptr(mraw), marker[1],
                                1 if per_chunk_raw else 0, out.ctypes.data, C.byref(raw), mt.ctypes.data,
                                mo.ctypes.data, mc.ctypes.data, cap_p, C.byref(npk))
    set_lz_fast(False)
    k = npk.value
    return out[:r].tobytes(), bool(raw.value), list(zip(mt[:k].tolist(), mo[:k].tolist(), mc[:k].tolist()))


def decompress_file(ambc, known=(1, 2, 3, 4, 255)):
    """adaptive_compressor.py:286-301 + 332-358 on bytes; raises like the reference."""
    import hashlib
    import struct
    if ambc[:4] != b"AMBC":
        raise ValueError("Magic mismatch")
    if ambc[4] > 2:
        raise ValueError("Unsupported version: %d" % ambc[4])
    hs = struct.unpack("<I", ambc[5:9])[0]
    mbits = ambc[9]
    ms = (mbits + 7) // 8
    mbytes = ambc[10:10 + ms]
    ctype = ambc[10 + ms]
    cs = 16 if ctype == 1 else 0
    csum = ambc[11 + ms:11 + ms + cs]
    orig = struct.unpack("<Q", ambc[11 + ms + cs:19 + ms + cs])[0]
    out = decompress_body(ambc[hs:], orig, marker_aligned(mbytes, mbits), known)
    if hashlib.md5(out).digest() != csum:
        raise ValueError("Checksum mismatch => possibly corrupted file.")
    return out


def find_marker(data, max_len=32, sample_size=None):
    """-> (marker bytes, L) or raises ValueError (marker_finder.py:22-123)"""
    a = _in(data)
    if sample_size and a.size > sample_size:
        s = np.empty(a.size, dtype=np.uint8)
        m = lib().orc_marker_sample(_ptr(a), a.size, sample_size, s.ctypes.data)
        a = s[:m]
    out = np.zeros(8, dtype=np.uint8)
    L = lib().orc_find_marker(_ptr(a), a.size, max_len, out.ctypes.data)
    if L == 0:
        raise ValueError("Could not find a marker of length <= %d bits" % max_len)
    return out[:(L + 7) // 8].tobytes(), L


def md5(data):
    a = _in(data)
    out = np.zeros(16, dtype=np.uint8)
    lib().orc_md5(_ptr(a), a.size, out.ctypes.data)
    return out.tobytes()
