"""world_size-2 gloo tests of the multi-process plumbing (parallel.py) on the CPU: rank sharding,
the one-time weight-arena broadcast, and the max-over-ranks timing reduction used by bench.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audioldm2_b200 import parallel


def test_shard_range_partitions_exactly():
    for n in (1, 7, 8, 64, 65):
        for world in (1, 2, 4, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a1 >= a0
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    # (1) arena broadcast: only rank 0 holds real bytes; the others must receive them bit-exactly
    g = torch.Generator().manual_seed(1234)
    arena = torch.randint(0, 256, (4096,), dtype=torch.uint8, generator=g)
    mine = arena if rank == 0 else torch.zeros_like(arena)
    got = parallel.make_arena_bcast(torch.device("cpu"))("unet", mine, arena.numel())
    ok_bcast = bool(torch.equal(got, arena))
    # (2) independent shards: every rank processes its own units, no data-path collective
    lo, hi = parallel.shard_range(10, rank, world)
    local = torch.arange(lo, hi).sum().item()
    # (3) timing reduction = max over ranks
    t = parallel.max_over_ranks(1.0 + rank, torch.device("cpu"))
    parallel.barrier()
    q.put((rank, ok_bcast, local, t))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    assert sum(l for _, _, l, _ in res) == sum(range(10))
    assert all(t == 2.0 for _, _, _, t in res)


def _gather_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    parallel.init_from_env(backend="gloo")
    n = 5                                              # uneven shards: 3 + 2
    r, w, lo, hi = parallel.current_shard(n)
    full = torch.arange(n * 3, dtype=torch.float32).reshape(n, 1, 3)
    got = parallel.all_gather_rows(full[lo:hi].clone(), n)
    # strided candidate rows of the prompts this rank owns (rows i + k*B, ddpm.py:1560-1562): same values as the full draw
    B, n_gen = n, 2
    rows = [i + k * B for k in range(n_gen) for i in range(lo, hi)]
    sn = parallel.ShardedNoise(B * n_gen, 0, 0, (2, 2), "cpu", seed=9, rows=rows)
    ref = parallel.ShardedNoise(B * n_gen, 0, B * n_gen, (2, 2), "cpu", seed=9)
    ok_noise = bool(torch.equal(sn.x_T(), ref.x_T()[rows]) and torch.equal(sn(0, "step"), ref(0, "step")[rows]))
    q.put((rank, bool(torch.equal(got, full)), ok_noise, (lo, hi)))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_prompt_shards_gather_and_strided_noise_rows():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert all(ok and okn for _, ok, okn, _ in res)
    assert [s for *_, s in res] == [(0, 3), (3, 5)]
