"""Drive the UNMODIFIED reference from /root/reference.  TEST INFRASTRUCTURE.

Works only in the build container (the GPU box has no /root/reference); used by
oracle/make_golden.py to produce tests/golden/ and by the container-only tests
that pin the C oracle to the reference.

Configuration follows SURVEY.md §0 D1 without touching reference code:
  c.CHUNK_SIZE_CANDIDATES = [N]     (instance attribute shadows the class list,
                                     adaptive_compressor.py:61-62, read at :548)
  c.compression_methods filtered    (and c.method_lookup rebuilt, :90-94)
"""
import contextlib
import importlib
import io
import os
import sys

REF_DIR = os.environ.get("AMBC_REFERENCE_DIR", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shim")


def available():
    return os.path.isfile(os.path.join(REF_DIR, "adaptive_compressor.py"))


_mods = {}


def _load():
    if _mods:
        return _mods
    if not available():
        raise RuntimeError("reference not present at %s" % REF_DIR)
    for p in (REF_DIR, _SHIM):
        if p not in sys.path:
            sys.path.append(p)
    # the reference checks os.path.exists('brotli_lzham_compression.py') relative to cwd
    with contextlib.redirect_stdout(io.StringIO()):
        for name in ("compression_methods", "marker_finder", "adaptive_compressor"):
            _mods[name] = importlib.import_module(name)
    # guard: make sure we imported the reference and not the product's shims of the same name
    for name, m in _mods.items():
        assert os.path.realpath(m.__file__).startswith(os.path.realpath(REF_DIR)), (name, m.__file__)
    return _mods


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        return fn(*a, **k)


def methods():
    cm = _load()["compression_methods"]
    return {1: cm.RLECompression(), 2: cm.DictionaryCompression(), 3: cm.HuffmanCompression(),
            4: cm.DeltaCompression(), 255: cm.NoCompression()}


def method_compress(mid, data):
    """-> (bytes | None, exception-name | None)"""
    m = methods()[mid]
    try:
        return quiet(m.compress, data), None
    except Exception as e:  # noqa: BLE001 - we record which one
        return None, type(e).__name__


def method_decompress(mid, payload, orig_len):
    m = methods()[mid]
    try:
        return quiet(m.decompress, payload, orig_len), None
    except Exception as e:  # noqa: BLE001
        return None, type(e).__name__


def should_use(mid, data):
    return bool(quiet(methods()[mid].should_use, data))


def make_compressor(chunk_size=None, method_ids=(1, 2, 3, 4), per_chunk_raw=False, found_marker=False):
    ac = _load()["adaptive_compressor"]
    base = ac.AdaptiveCompressor

    class _Cfg(base):
        pass

    if per_chunk_raw:
        # labelled extension: (remain,255) -> (min(N,remain),255)
        def _pick(self, data, position):
            csize, mid = base._pick_best_chunk_and_method(self, data, position)
            if mid == 255:
                n = min(self.CHUNK_SIZE_CANDIDATES[-1], len(data) - position)
                return n, 255
            return csize, mid
        _Cfg._pick_best_chunk_and_method = _pick
    if found_marker:
        def _fm(self, file_data, sample_size):
            return self.marker_finder.find_marker(file_data, None)
        _Cfg._find_marker = _fm
    c = quiet(_Cfg)
    if chunk_size is not None:
        c.CHUNK_SIZE_CANDIDATES = [int(chunk_size)] if isinstance(chunk_size, int) else list(chunk_size)
    keep = set(method_ids) | {255}
    seen = set()
    ms = []
    for m in c.compression_methods:
        if m.type_id in keep and m.type_id not in seen:
            seen.add(m.type_id)
            ms.append(m)
    c.compression_methods = ms
    c.method_lookup = {m.type_id: m for m in ms}
    return c


def compress_bytes(data, tmpdir, **cfg):
    """-> (ambc file bytes, stats, package map [(type, orig, comp)])"""
    c = make_compressor(**cfg)
    src = os.path.join(tmpdir, "in.bin")
    dst = os.path.join(tmpdir, "out.ambc")
    with open(src, "wb") as f:
        f.write(data)
    stats = quiet(c.compress, src, dst)
    with open(dst, "rb") as f:
        out = f.read()
    stored_verbatim = (out == data)  # adaptive_compressor.py:241-247
    return out, stats, (None if stored_verbatim else package_map(out))


def package_map(ambc):
    import struct
    hs = struct.unpack("<I", ambc[5:9])[0]
    mlen = (ambc[9] + 7) // 8
    pos = hs
    pk = []
    while pos + mlen + 14 <= len(ambc):
        t = ambc[pos + mlen]
        orig, comp = struct.unpack("<II", ambc[pos + mlen + 6: pos + mlen + 14])
        if t == 0:
            break
        pk.append((t, orig, comp))
        pos += mlen + 14 + comp
    return pk


def decompress_bytes(ambc, tmpdir, **cfg):
    c = make_compressor(**cfg)
    src = os.path.join(tmpdir, "in.ambc")
    dst = os.path.join(tmpdir, "out.bin")
    with open(src, "wb") as f:
        f.write(ambc)
    quiet(c.decompress, src, dst)
    with open(dst, "rb") as f:
        return f.read()


def find_marker(data, max_len=32, sample_size=None):
    mf = _load()["marker_finder"]
    try:
        b, l = quiet(mf.MarkerFinder(max_len).find_marker, data, sample_size)
        return bytes(b), int(l)
    except ValueError:
        return None, 0
