"""Commitment MSM split by point range over several ranks inside the prover (csrc/msm_team.cu): worker
processes (on cuda:r when the box has that many GPUs, otherwise sharing cuda:0) serve slices of every large gkr_msm_g1 the
leader issues; results must equal the single-GPU results limb for limb, and a whole Pippenger proof made with the team
attached must be byte-identical to the one made without it."""
import os
import random
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAU = 0x1234567890ABCDEF1234567890ABCDEF12345


def _worker(rank, world, name, n_srs, max_n):
    sys.path.insert(0, ROOT)
    import torch

    import gkr_msm_b200 as g
    from gkr_msm_b200 import hostmath as H
    from gkr_msm_b200.fieldutil import to_limb1

    dev = rank % max(torch.cuda.device_count(), 1)
    ctx = g.Context(dev)
    srs = g.Srs.mock_setup(ctx, to_limb1(TAU), H.g1_to_limbs(H.G1_GEN), n_srs)
    team = g.MsmTeam(ctx, name, rank, world, max_n)
    team.serve(srs, idle_timeout_s=120.0)
    team.close()
    ctx.close()


@pytest.mark.parametrize("world", [2, 3])
def test_team_msm_and_proof_equal_single_gpu(ctx, world):
    import torch.multiprocessing as mp

    import gkr_msm_b200 as g
    from gkr_msm_b200 import hostmath as H
    from gkr_msm_b200 import pippenger as DPP
    from gkr_msm_b200.fieldutil import R_MOD, to_limbs

    d, x, nbits, clm = 4, 9, 32, 1
    nv = x + clm
    n_srs = 2 * (1 << nv) - 1
    name = f"/gkr_msm_team_test_{os.getpid()}_{world}"
    mpctx = mp.get_context("spawn")
    procs = [mpctx.Process(target=_worker, args=(r, world, name, n_srs, n_srs)) for r in range(1, world)]
    kzg = DPP.KzgKey.mock_setup(ctx, TAU, H.G1_GEN, n_srs)
    key = DPP.KnucklesKey(ctx, kzg, nv, 2)
    rng = np.random.default_rng(5)
    sc = ctx.upload(to_limbs([int.from_bytes(rng.bytes(32), "little") % R_MOD for _ in range(n_srs)]))
    cfg = DPP.pippenger_config(d, x, nbits, clm)
    pts = H.te_points_arithmetic_progression(77, 0x9E3779B97F4A7C15, 1 << x)
    points_xy = np.stack([to_limbs([p[0] for p in pts]), to_limbs([p[1] for p in pts])])
    raw = np.frombuffer(rng.bytes(32 << x), dtype=np.uint8).reshape(1 << x, 32).copy()
    raw[:, nbits // 8:] = 0
    coefs = raw.view(np.uint64).reshape(1 << x, 4)
    r = [int.from_bytes(rng.bytes(32), "little") % R_MOD for _ in range(cfg["y_logsize"])]

    def prove():
        tr = g.Transcript(b"fgstglsp")
        g.run_pippenger_native(ctx, tr, kzg.srs, kzg.g0, key.dev, points_xy, coefs, d, x, nbits, clm, to_limbs(r))
        return tr.proof()

    # single-GPU references first
    ref_full = kzg.srs.msm(sc)
    ref_part = kzg.srs.msm(sc, n=700, first=13)
    ref_proof = prove()
    launches_single = ctx.launches
    for p in procs:
        p.start()
    team = g.MsmTeam(ctx, name, 0, world, n_srs)
    try:
        team.set_min_n(64)
        team.wait_ready(timeout_s=240.0)
        assert np.array_equal(kzg.srs.msm(sc), ref_full)
        assert np.array_equal(kzg.srs.msm(sc, n=700, first=13), ref_part)  # ragged slices, offset into the SRS
        assert prove() == ref_proof
    finally:
        team.quit()
        for p in procs:
            p.join(timeout=120)
        team.close()
    assert all(p.exitcode == 0 for p in procs)
    assert np.array_equal(kzg.srs.msm(sc), ref_full)  # detached again: local path
