"""dev tool: size distribution of the k-gram classes among the heads at 2k (the participants of the
LZ refine brackets), per corpus kind -- decides whether direct comparison beats the hash rounds."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import synth_ref as S
from collections import defaultdict, Counter

def names(data, L):
    first = {}
    out = []
    n = len(data)
    for p in range(n - L + 1):
        g = data[p:p + L]
        out.append(first.setdefault(g, p))
    return out

KINDS = ["csv", "log", "runs", "lowcard", "binrec", "text"]
for kind in (0, 1, 3, 4, 5):
    tot = Counter(); work = Counter(); parts = Counter(); hist = defaultdict(Counter)
    nch = 0
    for seg in range(3):
        seg_bytes = bytes(S.segment(S.DEFAULT_SEED, seg + 7, kind))
        for c in range(0, 65536, 4096 * 4):
            d = seg_bytes[c:c + 4096]
            nch += 1
            for k in (4, 8, 16):
                nk = names(d, k); n2 = names(d, 2 * k)
                cnt = Counter(nk)
                heads = [p for p in range(len(n2)) if n2[p] == p and cnt[nk[p]] > 1]
                cls = defaultdict(list)
                for p in heads:
                    cls[nk[p]].append(p)
                parts[k] += len(heads)
                for m in cls.values():
                    s = len(m)
                    hist[k][min(s, 9)] += s
                    work[k] += s * (s - 1) // 2
    print(KINDS[kind], "chunks", nch)
    for k in (4, 8, 16):
        h = hist[k]; t = sum(h.values()) or 1
        print("  bracket (%d,%d): participants/chunk %.0f, pair-compares/chunk %.0f, by class size: %s" % (
            k, 2 * k, parts[k] / nch, work[k] / nch, " ".join("%d:%.0f%%" % (s, 100 * h[s] / t) for s in sorted(h))))
