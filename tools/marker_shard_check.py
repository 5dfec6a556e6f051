"""multi-GPU check (torchrun): sharded marker search (per-shard flags + NCCL MAX all-reduce) == single-GPU
search over the whole stream == CPU oracle"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import torch, torch.distributed as dist
from adaptive_compression_b200 import engine, distributed as D
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for name, n, mask in (("mixed", 3 << 20, 0b1011111), ("random", 1 << 20, 1 << 5), ("odd", 1000003, 0b1111111)):
    per = (n + world - 1) // world
    a, b = min(n, rank * per), min(n, rank * per + per)
    shard = engine.synth(b - a, a, kind_mask=mask)
    got = D.find_marker_sharded(shard, 32)
    whole = engine.synth(n, 0, kind_mask=mask)
    want = engine.find_marker_device(whole, 32)
    if rank == 0:
        import oracle as O
        cpu = O.find_marker(whole[:1 << 20].cpu().numpy().tobytes(), 32) if n <= (1 << 20) else None
        print(name, n, "sharded", got, "single", want, "oracle(1MiB)", cpu, flush=True)
    ok = ok and got == want
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED MARKER", "OK" if t.item() else "MISMATCH", flush=True)
dist.destroy_process_group()
sys.exit(0 if t.item() else 1)
