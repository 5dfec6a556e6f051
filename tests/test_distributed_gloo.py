"""CPU, world_size 2 and 3 over gloo: the multi-rank host logic of the sharded chunk path --
placement all-gather, 'rest of file raw' monoid across ranks, fragment assembly, marker-flag
merge -- checked against the single-shot oracle.  Local bodies come from the oracle here (no GPU);
tests/test_gpu_* cover the CUDA bodies."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import inputs
import oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, data, chunk, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from adaptive_compression_b200 import distributed as D
    try:
        c0, b0, b1 = D.shard_range(len(data), chunk, rank, world)
        local = data[b0:b1]
        body, pm = O.compress_body(local, chunk)
        first_raw = next((i for i, p in enumerate(pm) if p[0] == 255), -1) if local else -1
        # in strict mode a raw package is always the last one and starts at a chunk boundary
        t_body = torch.from_numpy(np.frombuffer(body, dtype=np.uint8).copy())
        t_in = torch.from_numpy(np.frombuffer(local, dtype=np.uint8).copy()) if local else torch.empty(0, dtype=torch.uint8)
        pk = D.packed_bytes(len(body), len(local), first_raw, chunk)
        (off, state), recs = D.place_shards(pk, first_raw, c0, world)
        goff, frag = D.shard_fragment(t_body, t_in, first_raw, c0, chunk, len(data), recs, rank)
        # the assembled body on rank 0 (send / recv straight into the placed slices)
        whole = D.assemble_body(frag, recs, len(data), chunk)
        if rank == 0:
            q.put((-1, 0, whole.numpy().tobytes(), "assembled", b""))
        # marker flags: each rank marks some values, the merge must be the union
        flags = torch.zeros(64, dtype=torch.uint8)
        flags[rank::world] = 1
        flags[5] = 1 if rank == 0 else 0
        D.merge_marker_flags(flags)
        q.put((rank, goff, frag.numpy().tobytes(), state, flags.numpy().tobytes()))
    finally:
        dist.destroy_process_group()


def _run(world, data, chunk):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, data, chunk, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world + 1))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", ["all_packed", "raw_in_rank0", "raw_in_last", "raw_in_middle"])
def test_sharded_body_equals_single_shot(world, case):
    chunk = 1024
    parts = [inputs.mixed_file(5, chunk, 800 + i, ("text", "log", "runs", "lowcard")) for i in range(world)]
    if case == "raw_in_rank0":
        parts[0] = parts[0][:2 * chunk] + inputs.rand(chunk, 1) + parts[0][3 * chunk:]
    elif case == "raw_in_last":
        parts[-1] = parts[-1][:chunk] + inputs.rand(chunk, 2) + parts[-1][2 * chunk:]
    elif case == "raw_in_middle":
        parts[world // 2] = inputs.rand(chunk, 3) + parts[world // 2][chunk:]
    data = b"".join(parts) + inputs.text(300, 9)
    want, _ = O.compress_body(data, chunk)
    res = _run(world, data, chunk)
    assert res[0][3] == "assembled" and res[0][2] == want, case
    res = res[1:]
    out = bytearray(len(want))
    covered = 0
    for rank, off, frag, state, flags in res:
        out[off:off + len(frag)] = frag
        covered += len(frag)
        exp = np.zeros(64, dtype=np.uint8)
        for r in range(world):
            exp[r::world] = 1
        exp[5] = 1
        assert flags == exp.tobytes()
    assert covered == len(want)
    assert bytes(out) == want, case
    assert O.decompress_body(bytes(out), len(data)) == data


@pytest.mark.parametrize("world,n_chunks,tail,raw_at", [(4, 3, 500, 0), (4, 3, 500, 1), (3, 2, 1, 0), (4, 2, 0, 1), (2, 5, 77, 4),
                                                     (4, 3, 500, None), (3, 1, 300, 0)])
def test_more_ranks_than_chunks_and_partial_last_chunk(world, n_chunks, tail, raw_at):
    """empty trailing ranks behind a partial last chunk, with the raw chunk in an early rank --
    the END package has to land right behind the raw data"""
    chunk = 1024
    data = bytearray(inputs.mixed_file(n_chunks, chunk, 900 + n_chunks, ("text", "log", "runs", "lowcard")))
    if tail:
        data = data[:(n_chunks - 1) * chunk + tail]
    if raw_at is not None:
        a = raw_at * chunk
        b = min(len(data), a + chunk)
        data[a:b] = inputs.rand(b - a, 5)
    data = bytes(data)
    want, _ = O.compress_body(data, chunk)
    res = _run(world, data, chunk)
    assert res[0][3] == "assembled"
    assert res[0][2] == want, (world, n_chunks, tail, raw_at, len(res[0][2]), len(want))
    assert O.decompress_body(res[0][2], len(data)) == data
