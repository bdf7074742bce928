"""N>1 host logic on CPU (world_size 2, gloo): the shared-memory all-gather that carries the per-round
partial sums between the ranks, and the sharded-vs-whole decomposition of the round sums (oracle level)."""
import os
import random
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import gkr_msm_b200 as g

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    name = f"/gkr_test_{port}"
    ex = None
    if rank == 0:
        ex = g.Exchange(name, rank, world, create=True)
    dist.barrier()
    if rank != 0:
        ex = g.Exchange(name, rank, world, create=False)
    dist.barrier()
    ok = True
    for it in range(200):
        mine = np.full((3, 4), 1000 * it + rank, dtype=np.uint64)
        got = ex.allgather(mine)
        for r in range(world):
            ok &= bool(np.all(got[r] == 1000 * it + r))
    dist.barrier()
    ex.close()
    dist.destroy_process_group()
    q.put((rank, ok))


def test_exchange_allgather_two_ranks_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + random.randrange(2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, True), (1, True)]


def test_top_bit_sharding_decomposes_round_sums():
    """Oracle-level statement of SURVEY.md 8e: splitting by the top index bit, the round sums of the whole
    table are the sums of the shards' round sums for the first n-1 rounds, and the last round runs on the
    two survivors."""
    from oracle.pyref import gates as G
    from oracle.pyref import sumcheck as S
    from oracle.pyref.field import P

    rng = random.Random(4)
    nv = 6
    polys = [[rng.randrange(P) for _ in range(1 << nv)] for _ in range(3)]
    f = G.Prod3()
    claim = sum(f.exec([p[i] for p in polys]) for i in range(1 << nv)) % P
    whole = S.DenseSumcheckObjectSO(polys, f, nv, claim)
    half = 1 << (nv - 1)
    shards = [S.DenseSumcheckObjectSO([p[:half] for p in polys], f, nv - 1, 0),
              S.DenseSumcheckObjectSO([p[half:] for p in polys], f, nv - 1, 0)]
    for r in range(nv - 1):
        whole.unipoly()
        for s in shards:
            s.unipoly()
        tot = [(shards[0].last_evals[k] + shards[1].last_evals[k]) % P for k in range(1, 4)]
        assert tot == whole.last_evals[1:]
        t = rng.randrange(P)
        whole.bind(t)
        for s in shards:
            s.bind(t)
    survivors = [[shards[0].polys[j][0], shards[1].polys[j][0]] for j in range(3)]
    assert survivors == whole.polys
