"""GPU: synthetic-corpus generator == its numpy twin; marker search == reference golden / oracle,
including the sharded (carry + byte-max merge) formulation used across GPUs."""
import ctypes as C

import numpy as np
import pytest

import inputs
import oracle as O
import synth_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from adaptive_compression_b200 import engine
    engine.require_cuda()
    return engine


@pytest.mark.parametrize("offset,n,mask", [(0, 1 << 20, synth_ref.DEFAULT_KINDS), (65536 * 7 + 13, 300001, 0x7F),
                                           (5, 100, 0b0100000), (1 << 30, 1 << 18, synth_ref.DEFAULT_KINDS)])
def test_device_corpus_equals_numpy_twin(eng, offset, n, mask):
    got = eng.synth(n, offset=offset, kind_mask=mask).cpu().numpy()
    want = synth_ref.corpus(n, offset, kind_mask=mask)
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, (bad[:10], got[bad[:10]], want[bad[:10]])


def test_each_kind(eng):
    for k in range(7):
        got = eng.synth(65536 * 3, offset=0, kind_mask=1 << k).cpu().numpy()
        want = synth_ref.corpus(65536 * 3, 0, kind_mask=1 << k)
        assert (got == want).all(), k


def test_marker_golden(eng, golden):
    from adaptive_compression_b200 import MarkerFinder
    cases = {c[0]: c for c in inputs.marker_cases()}
    for row in golden["marker_kat"]:
        _, data, max_len, sample = cases[row["name"]]
        mf = MarkerFinder(max_len)
        if row["marker"] is None:
            with pytest.raises(ValueError):
                mf.find_marker(data, sample)
        else:
            b, L = mf.find_marker(data, sample)
            assert (b.hex(), L) == (row["marker"], row["length"]), row["name"]


def test_marker_fuzz_vs_oracle(eng):
    r = np.random.RandomState(5)
    kinds = sorted(inputs.KINDS)
    for i in range(40):
        k = kinds[r.randint(len(kinds))]
        n = int(r.choice([1, 2, 3, 5, 17, 100, 1000, 4096, 20000, 70001]))
        d = inputs.make(k, n, 60000 + i)
        want = O.find_marker(d, 32)
        got = eng.find_marker_device(eng.to_device(d), 32)
        assert got == want, (k, n, got, want)


def test_marker_large_random_needs_second_level(eng):
    d = inputs.rand(1 << 20, 99)  # ~20-bit marker: exercises the global-flags level
    want = O.find_marker(d, 32)
    assert want[1] > 16
    assert eng.find_marker_device(eng.to_device(d), 32) == want


def test_marker_sharded_equals_whole(eng):
    """per-shard flags with the left-boundary carry, merged by byte-wise max (what the NCCL
    all-reduce does across GPUs), then one pick == the single-pass answer"""
    import torch
    from adaptive_compression_b200 import _lib as Lb
    lib = Lb.lib()
    for seed, n, L, shards in [(1, 5000, 16, 3), (2, 40000, 16, 4), (3, 300000, 20, 2), (4, 999, 12, 5)]:
        d = inputs.rand(n, 700 + seed) if seed != 2 else inputs.text(n, seed)
        t = eng.to_device(d)
        merged = torch.zeros(1 << L, dtype=torch.uint8, device="cuda")
        cuts = [n * s // shards for s in range(shards + 1)]
        bits = "".join(format(b, "08b") for b in d)
        for s in range(shards):
            a, b = cuts[s], cuts[s + 1]
            fl = torch.zeros(1 << L, dtype=torch.uint8, device="cuda")
            cb = min(L - 1, a * 8)
            carry = int(bits[a * 8 - cb:a * 8], 2) if cb else 0
            Lb.check(lib.ambc_marker_flags_dev(C.c_void_p(t[a:b].data_ptr()), b - a, L, carry, cb,
                                               C.c_void_p(fl.data_ptr()), None))
            merged = torch.maximum(merged, fl)
        torch.cuda.synchronize()
        tb = min(31, n * 8)
        ln, val = C.c_uint32(0), C.c_uint64(0)
        Lb.check(lib.ambc_marker_pick_dev(C.c_void_p(merged.data_ptr()), L, 32, n * 8, int(bits[-tb:], 2) if tb else 0, tb,
                                          C.byref(ln), C.byref(val), None))
        assert (eng.marker_bytes(val.value, ln.value), ln.value) == O.find_marker(d, 32), (seed, n, L)
