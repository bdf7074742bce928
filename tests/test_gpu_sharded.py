"""Multi-rank sharded sumcheck on the GPU box: two processes (both on cuda:0, gloo for the
rendezvous) each own one half of the hypercube; the proof they produce must be byte-identical to the
single-process proof of the whole instance and to the oracle's."""
import os
import random
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, log_n_local, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    import gkr_msm_b200 as g
    from gkr_msm_b200.sharded import ShardedProd3Sumcheck

    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = g.Context(0)
    job = ShardedProd3Sumcheck(ctx, log_n_local, rank=rank, world=world, dist=dist, seed=77)
    for _ in range(3):
        (claim, point, fe) = job.prove_resident()
    proof = job.last[1]
    q.put((rank, proof, claim.tolist(), point.tolist(), fe.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_two_rank_sharded_proof_equals_single_rank(ctx, world):
    import torch.multiprocessing as mp

    import gkr_msm_b200 as g
    from gkr_msm_b200.sharded import ShardedProd3Sumcheck
    from oracle import coracle

    log_n_local = 10
    gbits = world.bit_length() - 1
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    port = 29500 + random.randrange(2000)
    procs = [mpctx.Process(target=_worker, args=(r, world, port, log_n_local, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
    # every rank ends with the same proof
    for r in res[1:]:
        assert r[1:] == res[0][1:]
    # single-process proof of the whole 2^(log_n_local + gbits) instance
    whole = ShardedProd3Sumcheck(ctx, log_n_local + gbits, rank=0, world=1, seed=77)
    claim, point, fe = whole.prove_resident()
    assert whole.last[1] == res[0][1]
    assert point.tolist() == res[0][3] and fe.tolist() == res[0][4]
    # and the oracle agrees on every round polynomial (replaying the device's challenges)
    n = log_n_local + gbits
    host = [coracle.synth_table(77 + j, 1 << n) for j in range(3)]
    ocl = coracle.gate_sum(0, 10, host)
    assert np.array_equal(ocl, whole.claim)
    chal = point[::-1].copy()  # the protocol returns the challenges reversed (sumcheck.rs:120)
    oev, ofe = coracle.dense_sumcheck(0, 10, host, n, ocl, chal)
    assert np.array_equal(ofe, fe)


def _msm_worker(rank, world, port, n_local, tau_limbs, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    import gkr_msm_b200 as g
    from gkr_msm_b200 import hostmath as H
    from gkr_msm_b200.sharded import sharded_commit

    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = g.Context(0)
    name = f"/gkr_msm_test_{port}"
    ex = g.Exchange(name, rank, world, create=True) if rank == 0 else None
    dist.barrier()
    if rank != 0:
        ex = g.Exchange(name, rank, world, create=False)
    dist.barrier()
    # every rank regenerates the whole mock SRS and keeps its own point range (a real deployment loads only its range)
    full = g.Srs.mock_setup(ctx, np.array(tau_limbs, dtype=np.uint64), H.g1_to_limbs(H.G1_GEN), n_local * world)
    mine = g.Srs(ctx, full.download_affine()[rank * n_local:(rank + 1) * n_local])
    scalars = ctx.synth(5, n_local, first_index=rank * n_local)
    out = sharded_commit(mine, scalars, ex)
    q.put((rank, out.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_msm_split_by_point_range_equals_single_gpu(ctx):
    """commitment MSM sharded by point range over 2 ranks == the single-GPU commitment, limb for limb"""
    import torch.multiprocessing as mp

    import gkr_msm_b200 as g
    from gkr_msm_b200 import hostmath as H
    from tests.util import to_limb1

    world, n_local = 2, 700
    tau = to_limb1(0x1234567890ABCDEF)
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    port = 31500 + random.randrange(2000)
    procs = [mpctx.Process(target=_msm_worker, args=(r, world, port, n_local, tau.tolist(), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
    assert res[0][1] == res[1][1]
    srs = g.Srs.mock_setup(ctx, tau, H.g1_to_limbs(H.G1_GEN), n_local * world)
    whole = srs.msm(ctx.synth(5, n_local * world))
    assert whole.tolist() == res[0][1]


def _eq_gamma_worker(rank, world, port, log_n_local, gate, n_in, point_limbs, gamma_limbs, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    import gkr_msm_b200 as g
    from gkr_msm_b200.sharded import mont_add_many

    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = g.Context(0)
    name = f"/gkr_eqg_test_{port}"
    ex = g.Exchange(name, rank, world, create=True) if rank == 0 else None
    dist.barrier()
    if rank != 0:
        ex = g.Exchange(name, rank, world, create=False)
    dist.barrier()
    n = 1 << log_n_local
    tabs = [ctx.synth(300 + j, n, first_index=rank * n) for j in range(n_in)]
    # the eq table of the GLOBAL point, sliced by the top index bits like every other table
    eq_full = ctx.eq_table(np.array(point_limbs, dtype=np.uint64)).download()
    tabs.append(ctx.upload(eq_full[rank * n:(rank + 1) * n]))
    consts = np.array(gamma_limbs, dtype=np.uint64)
    local = ctx.gate_sum(g.SO_EQ_GAMMA, gate, tabs, consts=consts)
    allc = ex.allgather(local.reshape(1, 4))
    claim = mont_add_many([allc[r, 0] for r in range(world)])
    so = ctx.dense_so(g.SO_EQ_GAMMA, gate, tabs, log_n_local, claim, consts=consts)
    tr = g.Transcript(b"fgstglsp")
    out = g.sumcheck_prove_sharded(tr, so, ex, log_n_local, g.SO_EQ_GAMMA, gate, claim, consts=consts)
    q.put((rank, tr.proof(), claim.tolist(), out[1].tolist(), out[2].tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("gate_name", ["AFF_L1", "LOGUP_LAYER"])
def test_sharded_eq_gamma_object_equals_single_rank(ctx, gate_name):
    """the sharded driver is not Prod3-shaped: an EqWrapper(GammaWrapper(gate)) object -- eq table sliced like the other tables --
    over 2 ranks gives the single-rank proof, incl. the last log2(world) rounds that run on the host"""
    import torch.multiprocessing as mp

    import gkr_msm_b200 as g
    from oracle.pyref.field import P
    from tests.util import to_limbs

    gate, n_in, n_out = {"AFF_L1": (g.GATE_AFF_L1, 4, 3), "LOGUP_LAYER": (g.GATE_LOGUP_LAYER, 4, 2)}[gate_name]
    world, log_n_local = 2, 9
    n = log_n_local + 1
    rng = random.Random(4242 + gate)
    point = to_limbs([rng.randrange(P) for _ in range(n)])
    gamma = rng.randrange(P)
    consts = to_limbs([pow(gamma, i, P) for i in range(max(n_out, 2))])
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    port = 33500 + random.randrange(2000)
    procs = [mpctx.Process(target=_eq_gamma_worker, args=(r, world, port, log_n_local, gate, n_in, point.tolist(), consts.tolist(), q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
    assert res[0][1:] == res[1][1:]
    tabs = [ctx.synth(300 + j, 1 << n) for j in range(n_in)] + [ctx.eq_table(point)]
    claim = ctx.gate_sum(g.SO_EQ_GAMMA, gate, tabs, consts=consts)
    assert claim.tolist() == res[0][2]
    so = ctx.dense_so(g.SO_EQ_GAMMA, gate, tabs, n, claim, consts=consts)
    tr = g.Transcript(b"fgstglsp")
    out = g.sumcheck_prove_sharded(tr, so, None, n, g.SO_EQ_GAMMA, gate, claim, consts=consts)
    assert tr.proof() == res[0][1]
    assert out[1].tolist() == res[0][3] and out[2].tolist() == res[0][4]


def _vv_problem(seed, rowv, colv, nrows):
    """deterministic ragged PRJ_L1 bundle (6 polynomials, the reference's point padding), point, gamma powers"""
    from oracle.pyref.field import P

    rng = random.Random(seed)
    lens = [rng.randrange(0, (1 << rowv) + 1) for _ in range(nrows)]
    lens[0] = 1 << rowv
    lens[nrows - 1] = max(lens[nrows - 1], 3)
    pads = [(0, 0), (1, 1), (1, 1)] * 2
    data = [[[rng.randrange(P) for _ in range(l)] for l in lens] for _ in range(6)]
    point = [rng.randrange(P) for _ in range(rowv + colv)]
    gamma = rng.randrange(P)
    gp = [pow(gamma, i, P) for i in range(4)]
    return lens, pads, data, point, gp


def _vv_upload(ctx, data, pads, rowv, colv, r0, r1):
    from tests.util import to_limb1, to_limbs

    return [ctx.upload_vecvec([to_limbs(r) if len(r) else np.zeros((0, 4), np.uint64) for r in data[j][r0:r1]], to_limb1(pads[j][0]),
                              to_limb1(pads[j][1]), rowv, colv) for j in range(6)]


def _vv_worker(rank, world, port, seed, rowv, colv, nrows, claim_limbs, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    import gkr_msm_b200 as g
    from tests.util import to_limbs

    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = g.Context(0)
    name = f"/gkr_vv_test_{port}"
    ex = g.Exchange(name, rank, world, create=True) if rank == 0 else None
    dist.barrier()
    if rank != 0:
        ex = g.Exchange(name, rank, world, create=False)
    dist.barrier()
    lens, pads, data, point, gp = _vv_problem(seed, rowv, colv, nrows)
    gl = world.bit_length() - 1
    per = 1 << (colv - gl)
    r0, r1 = rank * per, min((rank + 1) * per, nrows)
    polys = _vv_upload(ctx, data, pads, rowv, colv - gl, r0, r1)
    so = ctx.deg2_vecvec_shard_so(g.GATE_PRJ_L1, polys, to_limbs(gp), to_limbs(point), colv, rank, world)
    tr = g.Transcript(b"fgstglsp")
    claim, pt, fe = g.sumcheck_prove_sharded_vecvec(tr, so, ex, rowv + colv, np.array(claim_limbs, dtype=np.uint64))
    # the witness maps are row-local: this shard's L1 layer from its own rows
    l1 = ctx.map_vecvec([(g.GATE_PRJ_L1, 1)], polys, mode=0)
    rows = [[r.tolist() for r in p.download()[0]] for p in l1]
    q.put((rank, tr.proof(), claim.tolist(), pt.tolist(), fe.tolist(), rows))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,colv,nrows", [(2, 3, 7), (4, 3, 8), (2, 1, 2)])
def test_vecvec_sumcheck_sharded_by_rows_equals_single_gpu(ctx, world, colv, nrows):
    """SURVEY 8e, VecVec by bucket rows: the ragged Deg2 sumcheck with its rows split by the top bits of the row index over
    `world` ranks -- sparse rounds (totals added through the exchange), bind_into_dense, sharded dense tail with the last
    log2(world) rounds on the host -- writes the single-GPU proof bytes and returns its point and final evaluations; the
    single-GPU proof is the oracle's (tests/test_gpu_deg2.py).  The shards' witness maps are the rows of the whole map."""
    import torch.multiprocessing as mp

    import gkr_msm_b200 as g
    from oracle.pyref import gates as G
    from oracle.pyref import sumcheck as S
    from oracle.pyref.field import P
    from tests.util import to_limb1, to_limbs

    rowv, seed = 4, 5150 + 10 * world + colv
    lens, pads, data, point, gp = _vv_problem(seed, rowv, colv, nrows)
    nv = rowv + colv
    gate = G.PrjL1()
    opolys = [S.VecVecPolynomial(data[j], pads[j][0], pads[j][1], rowv, colv) for j in range(6)]
    eqp = S.eq_poly_sequence_last(point)
    dense = [p.vec() for p in opolys]
    evs = [0] * gate.n_outs
    for i in range(len(eqp)):
        o = gate.exec([d[i] for d in dense])
        for k in range(gate.n_outs):
            evs[k] = (evs[k] + o[k] * eqp[i]) % P
    claim = sum(gp[i] * evs[i] for i in range(gate.n_outs)) % P

    polys = _vv_upload(ctx, data, pads, rowv, colv, 0, nrows)
    so = ctx.deg2_vecvec_so(g.GATE_PRJ_L1, polys, to_limbs(gp), to_limb1(claim), to_limbs(point), colv)
    tr = g.Transcript(b"fgstglsp")
    c1, p1, f1 = g.sumcheck_prove(tr, so, nv)
    want = (tr.proof(), c1.tolist(), p1.tolist(), f1.tolist())
    # one shard through the sharded driver
    so1 = ctx.deg2_vecvec_shard_so(g.GATE_PRJ_L1, polys, to_limbs(gp), to_limbs(point), colv, 0, 1)
    tr1 = g.Transcript(b"fgstglsp")
    c2, p2, f2 = g.sumcheck_prove_sharded_vecvec(tr1, so1, None, nv, to_limb1(claim))
    assert (tr1.proof(), c2.tolist(), p2.tolist(), f2.tolist()) == want
    full_l1 = [[r.tolist() for r in p.download()[0]] for p in ctx.map_vecvec([(g.GATE_PRJ_L1, 1)], polys, mode=0)]

    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    port = 35500 + random.randrange(2000)
    procs = [mpctx.Process(target=_vv_worker, args=(r, world, port, seed, rowv, colv, nrows, to_limb1(claim).tolist(), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
    per = 1 << (colv - (world.bit_length() - 1))
    for rk, proof, c, pt, fe, rows in res:
        assert (proof, c, pt, fe) == want, f"rank {rk}"
        for j in range(len(full_l1)):
            assert rows[j] == full_l1[j][rk * per:(rk + 1) * per]


def test_peer_pool_places_tables_on_other_gpus_same_proof(tmp_path):
    """gkr_ctx_peer_pool: with the home GPU declared full beyond 64 MiB of tables (test hook), the prover places its tables on
    GPU 1 and reaches them through NVLink peer access -- the proof, outputs and pairing pair are the ones of the ordinary run.
    Needs two GPUs (gpurun --gpus 2); the single-GPU box of the regular suite skips it."""
    import json
    import subprocess

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    outs = []
    for pool in (0, 2):
        dump = str(tmp_path / f"p{pool}.npz")
        env = dict(os.environ)
        if pool:
            env["GKR_PEER_POOL_LOCAL_LIMIT_MIB"] = "64"
        cmd = [sys.executable, os.path.join(ROOT, "tools", "bench_pippenger.py"), "--x-logsize", "14", "--d-logsize", "7", "--nbits", "128", "--reps", "1",
               "--dump", dump] + (["--peer-pool", "2"] if pool else [])
        res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stderr[-2000:]
        outs.append((np.load(dump), json.loads(res.stdout.strip().splitlines()[-1])))
    a, b = outs[0][0], outs[1][0]
    for k in ("proof", "dense", "evs", "pair"):
        assert np.array_equal(a[k], b[k]), k
    assert outs[1][1]["peer_pool_peak_gib"] > 0.05, "nothing was placed on the peer"
