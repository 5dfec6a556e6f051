"""dev tool: k_decode(+lz) time per corpus kind, best of 5"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from adaptive_compression_b200 import engine, _lib as L
lib = engine.require_cuda()
n = 64 << 20
names = ['csv', 'log', 'runs', 'lowcard', 'binrec', 'random', 'text']
lib.ambc_enable_timing(1)
row = []
for k in (0, 1, 2, 3, 4, 6):
    t = engine.synth(n, 0, kind_mask=1 << k)
    o = engine.compress_device(t, 4096)
    body_host = o.body.cpu().numpy()
    table, _ = engine.index_host(body_host, n)
    t_table = torch.from_numpy(table.view(np.uint8).reshape(-1).copy()).to("cuda")
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    status = torch.zeros(2, dtype=torch.int32, device="cuda")
    best = 1e9
    for _ in range(6):
        L.check(lib.ambc_decompress_dev(C.c_void_p(o.body.data_ptr()), o.body_len, C.c_void_p(t_table.data_ptr()), len(table),
                                        C.c_void_p(out.data_ptr()), n, C.c_void_p(status.data_ptr()),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        ms = (C.c_float * 4)(); lib.ambc_last_timing(ms)
        best = min(best, ms[3])
    assert torch.equal(out, t)
    row.append("%s %.3f" % (names[k], best))
print(os.environ.get("AMBC_LIB_PATH", "default"), " | ".join(row), flush=True)
