"""Multi-GPU sharding of the chunk path: contiguous chunk ranges per rank, one
process per GPU, torch.distributed (NCCL over NVLink; gloo in the CPU tests) for the two tiny
exchanges the path has:

  place_shards        all-gather of one 16-byte record per rank -> every rank folds the
                      "rest of file raw" monoid left to right and learns its fragment's byte
                      offset in the global body, or that it lies inside the raw tail.
  merge_marker_flags  byte-wise MAX all-reduce of the per-shard n-gram presence flags
                      (NCCL has no bitwise OR; flags are 0/1 bytes).
"""
import torch
import torch.distributed as dist

NO_RAW = -1


def fold_placement(records):
    """records[r] = (fragment_bytes_before_first_raw, first_raw_global_chunk or -1), rank order.
    -> list of (offset, state): state 'packed' (fragment lands at offset), 'raw_starts_here'
    (fragment up to its first raw chunk lands at offset, the raw package follows) or 'in_raw_tail'.
    Monoid: A.B = A if A has a raw chunk else (A.bytes + B.bytes, B.first_raw)."""
    out = []
    offset = 0
    raw_seen = False
    for nbytes, first_raw in records:
        if raw_seen:
            out.append((None, "in_raw_tail"))
            continue
        if first_raw >= 0:
            out.append((offset, "raw_starts_here"))
            raw_seen = True
        else:
            out.append((offset, "packed"))
        offset += nbytes
    return out


def place_shards(fragment_bytes, first_raw_local, first_chunk_global, world):
    """all-gather the placement records; returns (this rank's (offset, state), all records)"""
    rank = dist.get_rank()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    fr = first_chunk_global + first_raw_local if first_raw_local >= 0 else NO_RAW
    mine = torch.tensor([int(fragment_bytes), int(fr)], dtype=torch.int64, device=dev)
    allr = torch.empty(2 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allr, mine)
    recs = [tuple(x) for x in allr.view(world, 2).cpu().tolist()]
    return fold_placement(recs)[rank], recs


def merge_marker_flags(flags):
    """in-place byte-wise MAX all-reduce of a 2^L-byte presence-flag tensor"""
    dist.all_reduce(flags, op=dist.ReduceOp.MAX)
    return flags
