"""dev tool: warp instructions per chunk of k_select_fast by function of select_fast.cuh, from an ncu report
(captured with --set full --import-source on).  usage: sf_regions.py report.ncu-rep n_chunks [top]"""
import csv, re, subprocess, sys, os
rep, nchunks = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 12
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None; lines = {}; src = {}
for r in rows:
    if r and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; idx = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < 8 or cur != 'select_fast.cuh' or r[0] == '': continue
    try:
        ln = int(r[0]); d = lines.setdefault(ln, [0, 0, 0]); src[ln] = r[1]
        d[0] += int(r[4] or 0); d[1] += int(float(r[idx['Instructions Executed']] or 0)); d[2] += int(float(r[idx['Thread Instructions Executed']] or 0))
    except Exception: pass
# function starts from the embedded source
# function starts from the source file itself (the report only lists lines that carry instructions)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
starts = []
for i, line in enumerate(open(os.path.join(root, 'adaptive_compression_b200/csrc/select_fast.cuh')), 1):
    m = re.match(r'^(?:template <[^>]*>\s*)?(?:__host__ )?__device__ .*?\b(sf_[a-z_0-9]+)\s*\(', line)
    if m: starts.append((i, m.group(1)))
    m2 = re.match(r'^__device__ .*?\b(sf_[a-z_0-9]+)\s*\(', line)
    if m2 and (not starts or starts[-1][0] != i): starts.append((i, m2.group(1)))
starts.append((10 ** 9, 'end'))
agg = {}
for ln, (s, i, t) in lines.items():
    name = 'helpers'
    for k in range(len(starts) - 1):
        if starts[k][0] <= ln < starts[k + 1][0]: name = starts[k][1]
    d = agg.setdefault(name, [0, 0, 0]); d[0] += s; d[1] += i; d[2] += t
ti = sum(a[1] for a in agg.values()); ts = sum(a[0] for a in agg.values()) or 1
for k, (s, i, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-22s samples %5.1f%%  warp-inst %5.1f%% (%7.0f/chunk) avg-active %.1f" % (k, 100 * s / ts, 100 * i / ti, i / nchunks, t / max(i, 1)))
print("%.0f warp-inst per chunk (select_fast.cuh only)" % (ti / nchunks))
for ln in sorted(lines, key=lambda l: -lines[l][1])[:top]:
    s, i, t = lines[ln]; print("%5d %5.1f%%s %7.0f wi/chunk act %4.1f | %s" % (ln, 100 * s / ts, i / nchunks, t / max(i, 1), src[ln].strip()[:100]))
