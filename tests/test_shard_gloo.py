"""CPU test of the N>1 host logic (world_size 2, gloo): index sharding + result gather to rank 0.
The per-rank compute is stood in for by the oracle (this is a test of the plumbing, not of the kernels)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from circl_b200.shard import RowGather, gather_rows, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 1000, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)           # every rank derives the same global inputs
    polys = (rng.integers(0, 2 * 3329, size=(n, 256)).astype(np.int32) - 3329).astype(np.int16)
    lo, hi = shard_range(n, rank, world)
    local = torch.from_numpy(oracle.kyber_ntt(polys[lo:hi]))   # stand-in for the per-rank GPU shard
    out, _ = gather_rows(local, n, dst=0)
    dist.barrier()
    if rank == 0:
        q.put(bool(np.array_equal(out.numpy(), oracle.kyber_ntt(polys))))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 33])
def test_two_rank_gather_matches_single_process(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def _worker_rows(rank, world, port, n, q):
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # weak scaling as bench.py runs it: every rank owns n ops, global op i = rank*n + local index; two result
    # matrices per op (stand-ins for ct and ss), pushed chunk by chunk while "compute" goes on
    rng = np.random.default_rng(7)
    polys = (rng.integers(0, 2 * 3329, size=(world * n, 256)).astype(np.int32) - 3329).astype(np.int16)
    mine = polys[rank * n:(rank + 1) * n]
    a = torch.from_numpy(oracle.kyber_ntt(mine).view(np.uint8).reshape(n, 512).copy())
    b = torch.from_numpy(np.stack([np.frombuffer(oracle.kyber_pack(oracle.kyber_normalize(p)), dtype=np.uint8) for p in mine]))
    g = RowGather(n, [512, 384], transport="sendrecv")
    cuts = [0, n // 3, n // 3, n - 1, n]  # includes an empty chunk and a single-row chunk
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        g.push([a, b], lo, hi)
    g.flush()
    dist.barrier()
    if rank == 0:
        want_a = oracle.kyber_ntt(polys).view(np.uint8).reshape(world * n, 512)
        want_b = np.stack([np.frombuffer(oracle.kyber_pack(oracle.kyber_normalize(p)), dtype=np.uint8) for p in polys])
        q.put(bool(np.array_equal(g.matrix(0).numpy(), want_a) and np.array_equal(g.matrix(1).numpy(), want_b)))
    g.close()
    dist.destroy_process_group()


def test_two_rank_row_gather_places_rows_in_global_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_rows, args=(r, 2, port, 11, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
