"""dev tool: source lines of an ncu report ranked by one stall reason (column name, e.g. stall_math, stall_wait)"""
import csv, subprocess, sys
rep, col = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; idx = None; agg = {}
for r in rows:
    if r and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': idx = {h: i for i, h in enumerate(r)}; continue
    if idx is None or len(r) < 8 or r[0] == '': continue
    try:
        v = int(r[idx[col]] or 0); inst = int(float(r[idx['Instructions Executed']] or 0))
    except Exception:
        continue
    d = agg.setdefault((cur, int(r[0])), [0, 0, r[1].strip()[:110]]); d[0] += v; d[1] += inst
tot = sum(d[0] for d in agg.values()) or 1
for (f, ln), (v, inst, txt) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% %-28s %5d inst %9d | %s" % (100 * v / tot, f, ln, inst, txt))
