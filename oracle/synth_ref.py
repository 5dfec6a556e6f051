"""numpy twin of adaptive_compression_b200/csrc/synth.cu (the synthetic "mixed CSV / log / binary"
corpus of BASELINE.json).  TEST INFRASTRUCTURE: lets the CPU legs (tests, bench cpu_baseline /
--impl reference) regenerate any byte range of the corpus without a GPU; the GPU test
tests/test_gpu_synth.py proves both generators emit the same bytes."""
import numpy as np

SEG = 65536
GOLD = np.uint64(0x9E3779B97F4A7C15)
MIXB = np.uint64(0xD1B54A32D192ED03)
DEFAULT_SEED = 0xA3BC0001
DEFAULT_KINDS = 0b1011111  # csv, log, runs, lowcard, binrec, text (no random)

_CSV_LIT = b"0000000,000.00,c00,2026-00-00,A\n"
_CSV_FID = [0, 0, 0, 0, 0, 0, 0, 255, 1, 1, 1, 255, 2, 2, 255, 255, 3, 3, 255, 255, 255, 255, 255, 255, 4, 4, 255, 5, 5,
            255, 6, 255]
_CSV_DIV = [1000000, 100000, 10000, 1000, 100, 10, 1, 0, 100, 10, 1, 0, 10, 1, 0, 0, 10, 1, 0, 0, 0, 0, 0, 0, 10, 1, 0,
            10, 1, 0, 0, 0]
_LOG_LIT = b"2026-10-18T00:00:00 LLLLL svc00 GET /v1/items/00000 000 00000ms\n"
_LOG_FID = [255] * 11 + [0, 0, 255, 1, 1, 255, 2, 2, 255, 3, 3, 3, 3, 3, 255, 255, 255, 255, 4, 4, 255] + \
           [255] * 4 + [255] * 10 + [5, 5, 5, 5, 5, 255, 6, 6, 6, 255, 7, 7, 7, 7, 7, 255, 255, 255]
_LOG_DIV = [0] * 11 + [10, 1, 0, 10, 1, 0, 10, 1, 0, 0, 1, 2, 3, 4, 0, 0, 0, 0, 10, 1, 0] + [0] * 4 + [0] * 10 + \
           [10000, 1000, 100, 10, 1, 0, 100, 10, 1, 0, 10000, 1000, 100, 10, 1, 0, 0, 0]
_LEVELS = [b"INFO ", b"WARN ", b"ERROR", b"DEBUG"]
_STATUS = np.array([200, 200, 200, 404, 500, 301, 200, 200], dtype=np.uint64)
_SKEW = np.array([0, 0, 0, 0, 1, 1, 1, 2, 2, 3, 4, 5, 6, 7, 8, 11], dtype=np.uint64)
_WORDS = [b"the     ", b"quick   ", b"brown   ", b"fox     ", b"jumps   ", b"over    ", b"lazy    ", b"dog     ",
          b"adaptive", b"marker  ", b"based   ", b"compress", b"chunk   ", b"method  ", b"huffman ", b"dictiona",
          b"delta   ", b"run     ", b"length  ", b"encoding", b"stream  ", b"header  ", b"package ", b"error   ",
          b"warning ", b"info    ", b"debug   ", b"request ", b"response", b"latency ", b"status  ", b"user    ",
          b"session ", b"and     ", b"of      ", b"to      ", b"in      ", b"is      ", b"that    ", b"for     ",
          b"with    ", b"as      ", b"on      ", b"be      ", b"at      ", b"by      ", b"this    ", b"have    ",
          b"from    ", b"or      ", b"one     ", b"had     ", b"not     ", b"but     ", b"what    ", b"all     ",
          b"were    ", b"when    ", b"we      ", b"there   ", b"can     ", b"an      ", b"your    ", b"which.\n "]
_WORDS_A = np.frombuffer(b"".join(_WORDS), dtype=np.uint8).reshape(64, 8)
assert len(_CSV_LIT) == 32 and len(_LOG_LIT) == 64 and len(_LOG_FID) == 64 and len(_LOG_DIV) == 64


def mix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x + GOLD
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def _u(x):
    return np.uint64(x)


def seg_kind(seed, seg, kind_mask):
    m = kind_mask & 0x7F
    kinds = [k for k in range(7) if (m >> k) & 1]
    with np.errstate(over="ignore"):
        pick = int(mix64(_u(seed) ^ (_u(seg + 1) * GOLD)) % _u(len(kinds)))
    return kinds[pick]


def _rec_hash(seed, seg, rec):
    with np.errstate(over="ignore"):
        return mix64(_u(seed) + _u(seg) * GOLD + rec.astype(np.uint64) * MIXB)


def segment(seed, seg, kind):
    """the 65536 bytes of segment `seg`"""
    if kind == 0:
        rec = np.arange(2048, dtype=np.uint64)
        h = _rec_hash(seed, seg, rec)
        f = [(_u(seg) * _u(2048) + rec) % _u(10000000), h % _u(1000), (h >> _u(10)) % _u(100), (h >> _u(20)) % _u(17),
             _u(1) + (h >> _u(28)) % _u(12), _u(1) + (h >> _u(34)) % _u(28), (h >> _u(40)) % _u(4)]
        out = np.empty((2048, 32), dtype=np.uint8)
        for j in range(32):
            fid = _CSV_FID[j]
            if fid == 255:
                out[:, j] = _CSV_LIT[j]
            elif fid == 6:
                out[:, j] = (65 + f[6]).astype(np.uint8)
            else:
                out[:, j] = (48 + (f[fid] // _u(_CSV_DIV[j])) % _u(10)).astype(np.uint8)
        return out.reshape(-1)
    if kind == 1:
        rec = np.arange(1024, dtype=np.uint64)
        h = _rec_hash(seed, seg, rec)
        t = _u(seg) * _u(1024) + rec
        lvl = h % _u(8)
        lvl = np.where(lvl < 5, 0, lvl - _u(4)).astype(np.int64)
        f = [(t // _u(3600)) % _u(24), (t // _u(60)) % _u(60), t % _u(60), None, (h >> _u(8)) % _u(12),
             (h >> _u(16)) % _u(50000), _STATUS[((h >> _u(36)) % _u(8)).astype(np.int64)], (h >> _u(40)) % _u(100000)]
        lv = np.frombuffer(b"".join(_LEVELS), dtype=np.uint8).reshape(4, 5)
        out = np.empty((1024, 64), dtype=np.uint8)
        for j in range(64):
            fid = _LOG_FID[j]
            if fid == 255:
                out[:, j] = _LOG_LIT[j]
            elif fid == 3:
                out[:, j] = lv[lvl, _LOG_DIV[j]]
            else:
                out[:, j] = (48 + (f[fid] // _u(_LOG_DIV[j])) % _u(10)).astype(np.uint8)
        return out.reshape(-1)
    if kind == 2:
        blk = np.arange(256, dtype=np.uint64)
        h = _rec_hash(seed, seg, blk)
        la = (h % _u(257)).astype(np.int64)[:, None]
        a = ((h >> _u(16)) & _u(255)).astype(np.uint8)[:, None]
        b = ((h >> _u(24)) & _u(255)).astype(np.uint8)[:, None]
        j = np.arange(256)[None, :]
        return np.where(j < la, a, b).astype(np.uint8).reshape(-1)
    if kind == 3:
        u = np.arange(4096, dtype=np.uint64)
        h = _rec_hash(seed, seg, u)
        with np.errstate(over="ignore"):
            sa = mix64(_u(seed) ^ (_u(seg) * MIXB))
        out = np.empty((4096, 16), dtype=np.uint8)
        for k in range(16):
            s = _SKEW[((h >> _u(4 * k)) & _u(15)).astype(np.int64)]
            out[:, k] = (_u(48) + ((sa >> (_u(4) * (s % _u(12)))) & _u(15)) + _u(6) * s).astype(np.uint8)
        return out.reshape(-1)
    if kind == 4:
        rec = np.arange(4096, dtype=np.uint64)
        h = _rec_hash(seed, seg, rec)
        cnt = ((_u(seg) * _u(4096) + rec) & _u(0xFFFFFFFF)).astype(np.uint32)
        ts = ((rec * _u(10) + h % _u(7)) & _u(0xFFFFFFFF)).astype(np.uint32)
        out = np.zeros((4096, 16), dtype=np.uint8)
        out[:, 0:4] = cnt.astype("<u4").view(np.uint8).reshape(-1, 4)
        out[:, 4] = (h % _u(40)).astype(np.uint8)
        out[:, 6] = ((h >> _u(8)) % _u(3)).astype(np.uint8)
        out[:, 8:12] = ts.astype("<u4").view(np.uint8).reshape(-1, 4)
        out[:, 12] = ((h >> _u(20)) & _u(1)).astype(np.uint8)
        out[:, 14] = 0xFF
        return out.reshape(-1)
    if kind == 5:
        u = np.arange(4096, dtype=np.uint64)
        h0 = _rec_hash(seed, seg, _u(2) * u)
        h1 = _rec_hash(seed, seg, _u(2) * u + _u(1))
        out = np.empty((4096, 16), dtype=np.uint8)
        out[:, 0:8] = h0.astype("<u8").view(np.uint8).reshape(-1, 8)
        out[:, 8:16] = h1.astype("<u8").view(np.uint8).reshape(-1, 8)
        return out.reshape(-1)
    u = np.arange(4096, dtype=np.uint64)
    h = _rec_hash(seed, seg, u)
    out = np.empty((4096, 16), dtype=np.uint8)
    for w in range(2):
        idx = ((h >> _u(6 * w)) & _u(63)).astype(np.int64)
        sk = ((h >> _u(20 + w)) & _u(1)).astype(bool)
        idx = np.where(sk, idx & 15, idx)
        out[:, 8 * w:8 * w + 8] = _WORDS_A[idx]
    return out.reshape(-1)


def corpus(n, offset=0, seed=DEFAULT_SEED, kind_mask=DEFAULT_KINDS):
    """bytes [offset, offset+n) of the corpus as a numpy uint8 array"""
    out = np.empty(n, dtype=np.uint8)
    pos = offset
    end = offset + n
    while pos < end:
        seg = pos // SEG
        lo = pos - seg * SEG
        hi = min(SEG, end - seg * SEG)
        s = segment(seed, seg, seg_kind(seed, seg, kind_mask))
        out[pos - offset:pos - offset + (hi - lo)] = s[lo:hi]
        pos += hi - lo
    return out
