"""GPU parity for the EC-addition GKR circuits end to end: witness generation on the device (maps + splits)
and every sumcheck layer through the C ABI, against the oracle restatement on the same points and the same
Fiat-Shamir transcript -- identical proof bytes and output claims; then the oracle VERIFIER accepts the
device-made proof.  Mirrors bintree_add.rs:401-460, triangle_add.rs:356-393, pippenger_ending.rs:176-275."""
import random

import numpy as np
import pytest

import gkr_msm_b200 as g
from gkr_msm_b200 import protocols as DP
from oracle.pyref import gates as G
from oracle.pyref import gkr as K
from oracle.pyref import polys as OP
from oracle.pyref import sumcheck as S
from oracle.pyref.field import P
from oracle.pyref.transcript import ProofTranscript2
from tests.test_oracle_gkr import dense_of, rand_points_affine
from tests.util import from_limbs, to_limb1, to_limbs

pytestmark = pytest.mark.gpu


def upload_vv(ctx, ov):
    rows = [to_limbs(r) if len(r) else np.zeros((0, 4), np.uint64) for r in ov.data]
    return ctx.upload_vecvec(rows, to_limb1(ov.row_pad), to_limb1(ov.col_pad), ov.row_logsize, ov.col_logsize)


@pytest.mark.parametrize("num_adds,row_logsize,col_logsize,bitcheck", [(5, 4, 2, False), (5, 2, 4, False), (3, 3, 1, False), (4, 4, 3, False)])
def test_bintree_device_vs_oracle(ctx, num_adds, row_logsize, col_logsize, bitcheck):
    rng = random.Random(100 * num_adds + 10 * row_logsize + col_logsize)
    num_vars = row_logsize + col_logsize
    points, _ = rand_points_affine(rng, row_logsize, col_logsize)
    inputs = OP.vecvec_map_split(points, G.Id(2), ("LO", 0), 2)
    oadv = K.bintree_witness(("vv", inputs), row_logsize, num_adds, False)
    olayers = K.bintree_protocol(num_vars, num_adds, row_logsize, False)
    olast = K.bintree_last_step(oadv[-1], num_adds - 1)
    dense_output = dense_of(olast, num_vars - num_adds)
    point = [rng.randrange(P) for _ in range(num_vars - num_adds)]
    claims = (point, [S.evaluate_poly(o, point) for o in dense_output])
    tp = ProofTranscript2.start_prover(b"fgstglsp")
    oclaims = K.simple_gkr_prove(olayers, tp, claims, oadv)
    proof = tp.end()

    dpoints = [upload_vv(ctx, p) for p in points]
    dinputs = ctx.map_vecvec(DP.ID(2)["parts"], dpoints, mode=1, bundle_size=2)
    dadv = DP.bintree_witness(ctx, ("vv", dinputs), row_logsize, num_adds, False)
    # witness parity, layer by layer
    for da, oa in zip(dadv, oadv):
        assert da[0] == oa[0]
        if da[0] == "dense":
            assert [from_limbs(t.download()) for t in da[1]] == [list(c) for c in oa[1]]
        elif da[0] == "vv":
            for dv, ov in zip(da[1], oa[1]):
                rows = dv.download()[0]
                assert [from_limbs(r) if len(r) else [] for r in rows] == ov.data
    dlayers = DP.bintree_protocol(ctx, num_vars, num_adds, row_logsize, False)
    tr = g.Transcript(b"fgstglsp")
    dclaims = DP.simple_gkr_prove(dlayers, tr, claims, dadv)
    assert tr.proof() == proof
    assert (list(dclaims[0]), list(dclaims[1])) == (list(oclaims[0]), list(oclaims[1]))
    tv = ProofTranscript2.start_verifier(b"fgstglsp", tr.proof())
    assert K.simple_gkr_verify(olayers, tv, claims) == oclaims


def test_triangle_device_vs_oracle(ctx):
    rng = random.Random(8)
    from oracle.pyref import curves as CV
    num_vars, split_var, hi = 6, ("HI", 2), 2
    pts = [CV.te_random_point(rng) for _ in range(1 << num_vars)]
    zs = [rng.randrange(1, P) for _ in pts]
    base = [[p[0] * z % P for p, z in zip(pts, zs)], [p[1] * z % P for p, z in zip(pts, zs)], zs]
    oin = OP.dense_map_split(OP.dense_map_split(base, G.Id(3), split_var, 3), G.Id(6), split_var, 3)
    oadv = K.triangle_witness(oin, num_vars - 2, split_var)
    olayers = K.triangle_protocol(num_vars - 2, split_var)
    olast = K.triangle_last_step(oadv[-1][1], num_vars - 2 - hi)
    point = [rng.randrange(P) for _ in range(hi)]
    claims = (point, [S.evaluate_poly(o, point) for o in olast])
    tp = ProofTranscript2.start_prover(b"fgstglsp")
    oclaims = K.simple_gkr_prove(olayers, tp, claims, oadv)

    dbase = [ctx.upload(to_limbs(c)) for c in base]
    din = ctx.map_dense(DP.ID(6)["parts"], ctx.map_dense(DP.ID(3)["parts"], dbase, split=split_var, bundle_size=3), split=split_var, bundle_size=3)
    dadv = DP.triangle_witness(ctx, din, num_vars - 2, split_var)
    dlast = DP.triangle_last_step(ctx, dadv[-1][1], num_vars - 2 - hi)
    assert [from_limbs(t.download()) for t in dlast] == olast
    tr = g.Transcript(b"fgstglsp")
    dclaims = DP.simple_gkr_prove(DP.triangle_protocol(ctx, num_vars - 2, split_var), tr, claims, dadv)
    assert tr.proof() == tp.end()
    assert (list(dclaims[0]), list(dclaims[1])) == (list(oclaims[0]), list(oclaims[1]))


def test_pippenger_ending_device_vs_oracle(ctx):
    rng = random.Random(21)
    multirow_vars, bucket_vars, point_vars = 1, 3, 3
    pre, rows = rand_points_affine(rng, point_vars, multirow_vars + bucket_vars)
    domain = S.VecVecPolynomial([[1] * len(r) for r in pre[0].data], 0, 0, point_vars, multirow_vars + bucket_vars)
    oinputs = OP.vecvec_map_split(pre, G.Id(2), ("LO", 0), 2) + OP.vecvec_map_split([domain], G.Id(1), ("LO", 0), 1)
    owg = K.PippengerEndingWG(multirow_vars, bucket_vars, point_vars, oinputs)
    oending = K.PippengerBucketed(multirow_vars, bucket_vars, point_vars)
    num_vars = multirow_vars + bucket_vars
    dense_output = K.triangle_last_step(owg.last(), num_vars - 2 - multirow_vars)
    point = [rng.randrange(P) for _ in range(multirow_vars)]
    claims = (point, [S.evaluate_poly(o, point) for o in dense_output])
    tp = ProofTranscript2.start_prover(b"fgstglsp")
    oclaims = oending.prove(tp, claims, owg)
    proof = tp.end()

    dimage = [upload_vv(ctx, p) for p in pre + [domain]]
    dinputs = DP.GlueSplit.witness(ctx, dimage)
    dwg = DP.PippengerEndingWG(ctx, multirow_vars, bucket_vars, point_vars, dinputs)
    dlast = DP.triangle_last_step(ctx, dwg.last(), num_vars - 2 - multirow_vars)
    assert [from_limbs(t.download()) for t in dlast] == dense_output
    tr = g.Transcript(b"fgstglsp")
    dclaims = DP.PippengerBucketed(ctx, multirow_vars, bucket_vars, point_vars).prove(tr, claims, dwg)
    assert tr.proof() == proof
    assert (list(dclaims[0]), list(dclaims[1])) == (list(oclaims[0]), list(oclaims[1]))
    tv = ProofTranscript2.start_verifier(b"fgstglsp", proof)
    assert oending.verify(tv, claims) == oclaims
