"""dev tool: per-source-line warp instructions / samples of one file from an ncu report"""
import csv, subprocess, sys
rep, fname, nchunks = sys.argv[1], sys.argv[2], int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None; agg = {}
for r in rows:
    if r and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; idx = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < 8 or cur != fname: continue
    if r[0] != '':
        try:
            ln = int(r[0]); d = agg.setdefault(ln, [0, 0, 0, r[1].strip()[:100]])
            d[0] += int(r[4] or 0); d[1] += int(float(r[idx['Instructions Executed']] or 0)); d[2] += int(float(r[idx['Thread Instructions Executed']] or 0))
        except Exception: pass
ts = sum(a[0] for a in agg.values()) or 1
for ln in sorted(agg):
    s, i, t, txt = agg[ln]
    if i: print("%4d %6.1f%%s %8.0f wi/chunk act %4.1f | %s" % (ln, 100 * s / ts, i / nchunks, t / max(i, 1), txt))
