"""Multi-GPU sharding of the chunk path: contiguous chunk ranges per rank, one
process per GPU, torch.distributed (NCCL over NVLink; gloo in the CPU tests) for the two tiny
exchanges the path has:

  place_shards        all-gather of one 16-byte record per rank -> every rank folds the
                      "rest of file raw" monoid left to right (adaptive_compressor.py:586-590) and
                      learns its fragment's byte offset in the global body, or that it lies inside
                      the raw tail.
  merge_marker_flags  byte-wise MAX all-reduce of the per-shard n-gram presence flags
                      (NCCL has no bitwise OR; flags are 0/1 bytes).

A rank compresses its range as if no earlier rank had hit a chunk without a winner; shard_fragment()
then turns the local body into the bytes this rank contributes to the global body."""
import struct

import torch
import torch.distributed as dist

NO_RAW = -1


def shard_range(total_bytes, chunk, rank, world):
    """contiguous chunk range of `rank`: (first_chunk, byte_begin, byte_end)"""
    n_chunks = (total_bytes + chunk - 1) // chunk
    per = (n_chunks + world - 1) // world
    c0 = min(n_chunks, rank * per)
    c1 = min(n_chunks, c0 + per)
    return c0, min(total_bytes, c0 * chunk), min(total_bytes, c1 * chunk)


def fold_placement(records):
    """records[r] = (fragment_bytes_before_first_raw, first_raw_global_chunk or -1), rank order.
    -> list of (offset, state): 'packed' (whole fragment lands at offset), 'raw_starts_here'
    (packed part at offset, then the global raw package) or 'in_raw_tail'.
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


def fold_placement_native(records, first_bytes, chunk, marker_bytes=4):
    """the same fold through the C-ABI (ambc_shard_place): -> list of (offset, state); for ranks inside
    the raw tail the offset is the place of their input bytes in the global body.  first_bytes[r] = input
    byte offset of rank r's shard (shard_range()[1]: clamped to the input length, so an empty trailing
    rank behind a partial last chunk still lands at the end of the raw data)"""
    import ctypes as C
    from . import _lib as L
    n = len(records)
    recs = (L.ShardRec * n)(*[L.ShardRec(int(b), int(fr)) for b, fr in records])
    fc = (C.c_uint64 * n)(*[int(x) for x in first_bytes])
    out = (L.ShardSlot * n)()
    L.check(L.lib().ambc_shard_place(recs, n, fc, chunk, marker_bytes, out))
    names = ("packed", "raw_starts_here", "in_raw_tail")
    return [(int(o.offset), names[o.state]) for o in out]


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def place_shards(fragment_bytes, first_raw_local, first_chunk_global, world):
    """all-gather the placement records; returns (this rank's (offset, state), all records)"""
    rank = dist.get_rank()
    fr = first_chunk_global + first_raw_local if first_raw_local >= 0 else NO_RAW
    mine = torch.tensor([int(fragment_bytes), int(fr)], dtype=torch.int64, device=_device())
    allr = torch.empty(2 * world, dtype=torch.int64, device=mine.device)
    dist.all_gather_into_tensor(allr, mine)
    recs = [tuple(x) for x in allr.view(world, 2).cpu().tolist()]
    return fold_placement(recs)[rank], recs


def packed_bytes(body_len, n_local, first_raw_local, chunk, marker_bytes=4):
    """bytes of the local body that precede its first raw chunk (END and local raw package removed)"""
    end = marker_bytes + 12
    if first_raw_local < 0:
        return body_len - end
    return body_len - end - (marker_bytes + 14) - (n_local - first_raw_local * chunk)


def fragment_plan(records, total_bytes, chunk, marker_bytes=4):
    """-> ([(offset, length, state)] per rank, body length): where every rank's contribution lands in the
    global body and how long it is, from the all-gathered records alone (every rank computes the same plan;
    the shards are shard_range(total_bytes, chunk, r, world)).  The last rank's length includes the END
    package.  Rule sharded: adaptive_compressor.py:586-590."""
    world = len(records)
    ovh, end = marker_bytes + 14, marker_bytes + 12
    states = fold_placement(records)
    g = next((fr for _, fr in records if fr >= 0), NO_RAW)  # global first raw chunk
    packed_total = 0
    if g >= 0:
        packed_total = sum(nb for nb, _ in records[:next(i for i, (_, fr) in enumerate(records) if fr >= 0) + 1])
    plan = []
    for r, ((off, state), (nbytes, _)) in enumerate(zip(states, records)):
        _, b0, b1 = shard_range(total_bytes, chunk, r, world)
        if state == "packed":
            length = nbytes
        elif state == "raw_starts_here":
            length = nbytes + ovh + (b1 - g * chunk)
        else:  # input bytes, verbatim, inside the raw package (b0 is clamped to the input length)
            off = packed_total + ovh + (b0 - g * chunk)
            length = b1 - b0
        if r == world - 1:
            length += end
        plan.append((off, length, state))
    last_off, last_len, _ = plan[-1]
    return plan, last_off + last_len


def shard_fragment(local_body, local_input, first_raw_local, first_chunk_global, chunk, total_bytes, records, rank,
                   marker=b"\xff\xff\x00\x00"):
    """-> (global byte offset, 1-D uint8 tensor) this rank contributes to the global body (the END
    package is appended by the last rank).  local_body / local_input are tensors on any device."""
    mb = len(marker)
    plan, _ = fragment_plan(records, total_bytes, chunk, mb)
    offset, length, state = plan[rank]
    n_local = local_input.numel()
    pk = packed_bytes(local_body.numel(), n_local, first_raw_local, chunk, mb)
    last = rank == len(records) - 1
    end_pkg = torch.tensor(list(marker + bytes(12)), dtype=torch.uint8, device=local_body.device)
    g = next((fr for _, fr in records if fr >= 0), NO_RAW)  # global first raw chunk
    if state == "packed":
        if last and first_raw_local < 0:  # the local body ends with the same END package: no copy
            assert local_body.numel() == pk + mb + 12
            return offset, local_body
        frag = local_body[:pk]
    elif state == "raw_starts_here":
        rawlen = total_bytes - g * chunk
        if rawlen > 0xFFFFFFFF:
            raise struct.error("'I' format requires 0 <= number <= 4294967295")  # adaptive_compressor.py:617-619
        hdr = marker + bytes([255, 0]) + struct.pack("<III", rawlen, rawlen, rawlen)
        hdr_t = torch.tensor(list(hdr), dtype=torch.uint8, device=local_body.device)
        frag = torch.cat([local_body[:pk], hdr_t, local_input[first_raw_local * chunk:]])
    else:
        frag = local_input
    if last:
        frag = torch.cat([frag, end_pkg])
    assert frag.numel() == length, (frag.numel(), length, state)
    return offset, frag


def assemble_body(frag, records, total_bytes, chunk, marker_bytes=4, dst=0, out=None):
    """Gather every rank's fragment into ONE body on rank `dst` (NCCL send / recv straight into the placed
    slice of the destination buffer -- NVLink peer copies; gloo on CPU).  Returns the body tensor on `dst`
    (None elsewhere).  `out`: optional preallocated destination on `dst` (>= body length)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    plan, body_len = fragment_plan(records, total_bytes, chunk, marker_bytes)
    if rank != dst:
        if frag.numel():
            dist.send(frag.contiguous(), dst)
        return None
    body = out[:body_len] if out is not None else torch.empty(body_len, dtype=torch.uint8, device=frag.device)
    reqs = []
    for r, (off, length, _) in enumerate(plan):
        if r == dst:
            body[off:off + length].copy_(frag)
        elif length:
            reqs.append(dist.irecv(body[off:off + length], r))
    for q in reqs:
        q.wait()
    return body


def merge_marker_flags(flags):
    """in-place byte-wise MAX all-reduce of a 2^L-byte presence-flag tensor"""
    dist.all_reduce(flags, op=dist.ReduceOp.MAX)
    return flags


def carry_windows(lens, tails):
    """lens[r] = bytes in shard r, tails[r] = its last min(4, lens[r]) bytes.  -> for every rank the
    (up to) 4 bytes that precede its shard in the global stream and how many exist, plus the same for
    the end of the whole stream.  Shards may be shorter than 4 bytes (or empty)."""
    out, window, before = [], b"", 0
    for n, t in zip(lens, tails):
        out.append((window, before))
        window = (window + bytes(t))[-4:]
        before += int(n)
    return out, (window, before)


def _bits_of(window, before_bytes, want_bits):
    """the last min(want_bits, 8 * before_bytes) bits of the stream ending with `window`, as (value, count)"""
    have = min(want_bits, 8 * min(before_bytes, len(window)))
    v = int.from_bytes(window, "big") if window else 0
    return (v & ((1 << have) - 1)) if have else 0, have


def find_marker_sharded(t_shard, max_len=32, levels=(16, 24, 32)):
    """MarkerFinder.find_marker (marker_finder.py:22-123) over a stream sharded across the ranks in rank
    order: every rank flags the L-bit windows that start in its shard (with the <= L-1 bits that precede
    it as carry), one byte-wise MAX all-reduce merges the flags (NCCL has no bitwise OR), every rank
    picks the same (marker bytes, length).  Raises ValueError when no marker of at most
    min(max_len, levels[-1]) bits exists."""
    import ctypes as C
    from . import _lib as L
    from . import engine
    lib = engine.require_cuda()
    world, rank = dist.get_world_size(), dist.get_rank()
    n_local = t_shard.numel()
    dev = t_shard.device
    meta = torch.zeros(5, dtype=torch.int64, device=dev)
    meta[0] = n_local
    k = min(4, n_local)
    if k:
        meta[1:1 + k] = t_shard[n_local - k:].to(torch.int64)
    allm = torch.empty(5 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allm, meta)
    rows = allm.view(world, 5).cpu().tolist()
    lens = [int(r[0]) for r in rows]
    tails = [bytes(int(x) for x in r[1:1 + min(4, int(r[0]))]) for r in rows]
    per_rank, (end_window, total_bytes) = carry_windows(lens, tails)
    window, before = per_rank[rank]
    total_bits = 8 * total_bytes
    tail, tail_bits = _bits_of(end_window, total_bytes, 31)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for lv in levels:
        Lb = min(lv, max_len)
        flags = torch.zeros(1 << Lb, dtype=torch.uint8, device=dev)
        carry, cb = _bits_of(window, before, Lb - 1)
        L.check(lib.ambc_marker_flags_dev(C.c_void_p(t_shard.data_ptr() if n_local else 0), n_local, Lb, carry, cb,
                                          C.c_void_p(flags.data_ptr()), stream))
        merge_marker_flags(flags)
        ln, val = C.c_uint32(0), C.c_uint64(0)
        rc = lib.ambc_marker_pick_dev(C.c_void_p(flags.data_ptr()), Lb, max_len, total_bits, tail, tail_bits,
                                      C.byref(ln), C.byref(val), stream)
        if rc == L.OK:
            return engine.marker_bytes(val.value, ln.value), ln.value
        if rc != L.E_NO_MARKER or Lb == max_len:
            L.check(rc)
    raise ValueError("Could not find a marker of length <= %d bits" % min(max_len, levels[-1]))
