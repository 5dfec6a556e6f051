"""Worker of tests/test_gpu_multirank.py (launched with torch.distributed.run, one rank per GPU): every rank
compresses its contiguous chunk range on its own GPU (CUDA path through the C-ABI), the 16-byte placement
records are all-gathered over NCCL, the fragments are sent straight into their placed slices of ONE body on
rank 0, and rank 0 compares that body with (a) the single-GPU CUDA body of the whole input and (b) the oracle.
Also runs the sharded marker search against the single-GPU one."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

import inputs
import oracle as O
from adaptive_compression_b200 import distributed as D
from adaptive_compression_b200 import engine


def case_data(case, world, chunk):
    parts = [inputs.mixed_file(9, chunk, 700 + i, ("text", "log", "runs", "lowcard", "csv")) for i in range(world)]
    if case == "raw_in_rank0":
        parts[0] = parts[0][:2 * chunk] + inputs.rand(chunk, 1) + parts[0][3 * chunk:]
    elif case == "raw_in_middle":
        parts[world // 2] = inputs.rand(chunk, 3) + parts[world // 2][chunk:]
    elif case == "raw_in_last":
        parts[-1] = parts[-1][:chunk] + inputs.rand(chunk, 2) + parts[-1][2 * chunk:]
    elif case == "short":  # fewer chunks than ranks, partial last chunk, raw chunk first
        return inputs.rand(chunk, 4) + inputs.text(300, 9)
    return b"".join(parts) + inputs.text(300, 9)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    engine.require_cuda()
    ok = True
    for chunk in (1024, 4096):
        for case in ("all_packed", "raw_in_rank0", "raw_in_middle", "raw_in_last", "short"):
            data = case_data(case, world, chunk)
            total = len(data)
            c0, b0, b1 = D.shard_range(total, chunk, rank, world)
            t_in = torch.from_numpy(np.frombuffer(data[b0:b1], dtype=np.uint8).copy()).cuda()
            o = engine.compress_device(t_in, chunk)
            fr = int(o.first_raw)
            pk = D.packed_bytes(int(o.body_len), b1 - b0, fr, chunk)
            (_, state), recs = D.place_shards(pk, fr, c0, world)
            _, frag = D.shard_fragment(o.body, t_in, fr, c0, chunk, total, recs, rank)
            body = D.assemble_body(frag, recs, total, chunk)
            if rank == 0:
                whole = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).cuda()
                single = engine.compress_device(whole, chunk)
                got = body.cpu().numpy().tobytes()
                want_gpu = single.body.cpu().numpy().tobytes()
                want_orc, _ = O.compress_body(data, chunk)
                same = got == want_gpu == want_orc
                dec, st = engine.decompress_device(body, total)
                rt = dec.cpu().numpy().tobytes() == data and st == [0, 0]
                print("case %-14s chunk %4d world %d: assembled %d bytes sha %s %s %s" %
                      (case, chunk, world, len(got), hashlib.sha256(got).hexdigest()[:12],
                       "== single-GPU == oracle" if same else "MISMATCH", "round trip ok" if rt else "ROUND TRIP FAILED"), flush=True)
                ok = ok and same and rt
    # sharded marker search == single-GPU search
    data = inputs.mixed_file(64, 1024, 77, ("text", "log", "csv"))
    total = len(data)
    per = (total + world - 1) // world
    t_sh = torch.from_numpy(np.frombuffer(data[rank * per:(rank + 1) * per], dtype=np.uint8).copy()).cuda()
    mk = D.find_marker_sharded(t_sh)
    if rank == 0:
        whole = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).cuda()
        want = engine.find_marker_device(whole) if hasattr(engine, "find_marker_device") else None
        print("marker sharded %s single %s" % (mk, want), flush=True)
        if want is not None:
            ok = ok and tuple(mk) == tuple(want)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTIRANK OK" if ok else "MULTIRANK FAILED", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
