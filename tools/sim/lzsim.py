"""dev tool (CPU): candidate / compare statistics of the lazy bucket parse per corpus kind"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import synth_ref
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "liblzsim.so"))
names = ['csv', 'log', 'runs', 'lowcard', 'binrec', 'random', 'text']
hb = int(sys.argv[1]) if len(sys.argv) > 1 else 11
G = int(sys.argv[2]) if len(sys.argv) > 2 else 8
nchunks = 64
lib.lzsim_mode(int(sys.argv[3]) if len(sys.argv) > 3 else 0)
for k in (0, 1, 2, 3, 4, 6):
    data = synth_ref.corpus(nchunks * 4096, 0, kind_mask=1 << k)
    tot = np.zeros(16, dtype=np.int64)
    mxb = 0
    for i in range(nchunks):
        out = np.zeros(16, dtype=np.int64)
        ch = np.ascontiguousarray(data[i * 4096:(i + 1) * 4096])
        lib.lzsim(ch.ctypes.data, 4096, hb, G, out.ctypes.data)
        mxb = max(mxb, out[9]); tot += out
    t = tot / nchunks
    print("%-8s lit %5.0f match %5.0f payload %5.0f | cand/visit %6.1f batches(G=%d)/chunk %6.0f stopbatches %6.0f | wordcmp naive %7.0f runbest %7.0f bestwords %5.0f | has3 steps %6.0f collide %4.0f maxbucket %d"
          % (names[k], t[0], t[1], t[2], t[3] / max(1, t[1] + t[8]), G, t[4], t[10], t[5], t[6], t[11], t[7], t[8], mxb))
