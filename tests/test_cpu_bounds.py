"""The no-parse lower bound of the Dictionary payload that k_select_fast uses to end hopeless trials
(csrc/select_fast.cuh: sf_lz_bound_ngrams): (4 * D + 2 * F) / 3 with D = distinct 4-grams and F = positions whose
trigram and the trigrams one and two bytes before are all first occurrences.  Restated in numpy (exact counts, i.e.
the largest value the kernel's hashed counts can reach) and held against the oracle's exact greedy parse
(compression_methods.py:195-234, 283-313) on every input kind: the bound never exceeds the payload."""
import numpy as np

import inputs
import oracle as O


def ngram_bound(data):
    d = np.frombuffer(data, dtype=np.uint8).astype(np.uint32)
    n = len(d)
    if n < 4:
        return 0
    w4 = d[:-3] | (d[1:-2] << 8) | (d[2:-1] << 16) | (d[3:] << 24)
    D = len(np.unique(w4))
    tri = d[:-2] | (d[1:-1] << 8) | (d[2:] << 16)
    _, first_idx = np.unique(tri, return_index=True)
    first = np.zeros(len(tri), dtype=bool)
    first[first_idx] = True
    forced = first[2:] & first[1:-1] & first[:-2]     # position p = index + 2
    F = int(forced[:max(0, n - 3 - 2)].sum())          # positions p <= n - 4 (they have a 4-gram)
    return (4 * D + 2 * F) // 3


def test_bound_never_exceeds_the_greedy_payload():
    r = np.random.RandomState(31)
    worst = 1.0
    checked = 0
    for kind in sorted(inputs.KINDS):
        for i in range(12):
            n = int(r.choice([100, 257, 1024, 3000, 4096, 8192]))
            data = inputs.make(kind, n, 600 + 17 * i)
            pay = O.compress(2, data, lz_fast=True)
            assert isinstance(pay, (bytes, bytearray))
            b = ngram_bound(data)
            assert b <= len(pay), (kind, n, b, len(pay))
            worst = min(worst, (len(pay) - b) / max(1, len(pay)))
            checked += 1
    # adversarial shapes: all distinct 4-grams, long literal stretches, runs, periodic data
    extra = [bytes(r.randint(0, 256, size=4096).astype(np.uint8)), bytes(range(256)) * 16, b"ab" * 2048, bytes(4096),
             b"".join(bytes([i, i, i, 255 - i]) for i in range(256)) * 4, bytes(r.randint(0, 3, size=4096).astype(np.uint8))]
    for data in extra:
        pay = O.compress(2, data, lz_fast=True)
        assert ngram_bound(data) <= len(pay), (len(data), ngram_bound(data), len(pay))
        checked += 1
    assert checked > 100 and worst >= 0.0


def test_bound_is_sharp_enough_to_matter():
    """on CSV-like and low-cardinality chunks the bound reaches the Huffman payload (the case the kernel skips)"""
    hits = 0
    for kind in ("csv", "lowcard"):
        for i in range(8):
            data = inputs.make(kind, 4096, 900 + i)
            hf = O.compress(3, data)
            if isinstance(hf, (bytes, bytearray)) and ngram_bound(data) >= len(hf) + 1:
                hits += 1
    assert hits >= 8
