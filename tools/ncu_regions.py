"""dev tool: aggregate an ncu source page (cuda,sass) of k_select by code region"""
import csv, collections, subprocess, sys, os
rep = sys.argv[1]; nchunks = int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None; agg = []
for r in rows:
    if r and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; idx = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < 8: continue
    if r[0] != '':
        try: agg.append((cur, int(r[0]), int(r[4] or 0), int(float(r[idx['Instructions Executed']] or 0)), int(float(r[idx['Thread Instructions Executed']] or 0)), r[1].strip()[:90]))
        except Exception: pass
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, 'adaptive_compression_b200/csrc/chunk_codec.cuh')).read().split('\n')
def find(s): return next(i + 1 for i, l in enumerate(src) if s in l)
marks = [('carve/load', 1), ('features', find('inline void chunk_features')), ('first_order', find('inline void chunk_first_order')), ('rle_encode', find('inline int chunk_rle_encode')), ('delta', find('inline int chunk_delta_encode')), ('lz_hash', find('lz_hash_l(const uint8_t')), ('bucket_sort', find('inline void lz_bucket_sort')), ('match_len', find('lz_match_len(const uint8_t')), ('lz_phaseA', find('inline int chunk_lz_encode')), ('lz_phaseB', find('// long buckets: one warp per slot')), ('lz_chain', find('// ---- token chain')), ('huff', find('struct HuffScratch')), ('end', 99999)]
ts = sum(a[2] for a in agg); ti = sum(a[3] for a in agg)
res = collections.OrderedDict()
for f, l, s, i, t, _ in agg:
    key = f if f != 'chunk_codec.cuh' else [m[0] for m in marks if m[1] <= l][-1]
    d = res.setdefault(key, [0, 0, 0]); d[0] += s; d[1] += i; d[2] += t
for k, (s, i, t) in res.items():
    if i * 200 > ti or s * 200 > ts: print("%-28s samples %5.1f%%  warp-inst %5.1f%% (%7.0f/chunk) avg-active %.1f" % (k, 100 * s / ts, 100 * i / ti, i / nchunks, t / max(i, 1)))
print("%.0f warp-inst per chunk" % (ti / nchunks))
for a in sorted(agg, key=lambda a: -a[2])[:14]: print("%5.1f%%s %5.1f%%i %s:%d | %s" % (100 * a[2] / ts, 100 * a[3] / ti, a[0], a[1], a[5]))
