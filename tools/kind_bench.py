"""dev tool: per-kind k_select / k_decode time (ms per 64 MiB) for a list of LZ level sets"""
import ctypes as C, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptive_compression_b200 import engine, _lib as L
lib = engine.require_cuda()
n = 64 << 20
names = ['csv', 'log', 'runs', 'lowcard', 'binrec', 'random', 'text']
level_sets = [[int(x) for x in s.split(',')] for s in (sys.argv[1:] or ['3', '3,4,8'])]
lib.ambc_enable_timing(1)
T=int(os.environ.get('COOP_T','24')); lib.ambc_set_lz_coop_threshold(T); print('coop T',T)
for k in (0, 1, 2, 3, 4, 6, 5):
    t = engine.synth(n, 0, kind_mask=1 << k)
    row = [names[k]]
    for lv in level_sets:
        engine.set_lz_levels(lv)
        for _ in range(2):
            o = engine.compress_device(t, 4096)
        ms = (C.c_float * 4)(); lib.ambc_last_timing(ms)
        row.append("%s: sel %.2f ms" % (lv, ms[0]))
    out, st = engine.decompress_device(o.body, n)
    ms = (C.c_float * 4)(); lib.ambc_last_timing(ms)
    row.append("dec %.2f ms ratio %.3f usage %s" % (ms[3], o.body_len / n, o.usage))
    print(row, flush=True)
