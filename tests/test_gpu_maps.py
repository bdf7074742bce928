"""GPU parity for the witness maps (trait MapSplit): the reference's `map`, `map_split`, `map_split_to_dense`,
`vec_algfn_map`, `map_split(_lo)` tests (polys/vecvec.rs:711-876, src/utils.rs:365-541) with the device
maps checked against the oracle restatement, on dense and ragged inputs."""
import random

import numpy as np
import pytest

import gkr_msm_b200 as g
from oracle.pyref import gates as G
from oracle.pyref import polys as OP
from oracle.pyref import sumcheck as S
from oracle.pyref.field import P
from tests.test_gpu_deg2 import BASE, oracle_stack
from tests.util import from_limbs, to_limb1, to_limbs

pytestmark = pytest.mark.gpu


def ostack(parts):
    f = None
    for gid, rep in parts:
        part = G.Id(rep) if gid == g.GATE_ID else (BASE[gid]() if rep == 1 else G.Repeated(BASE[gid](), rep))
        f = part if f is None else G.Stacked(f, part)
    return f


STACKS = [
    [(g.GATE_AFF_L1, 1)], [(g.GATE_AFF_L2, 1)], [(g.GATE_AFF_L3, 1)], [(g.GATE_PRJ_L1, 1)], [(g.GATE_PRJ_L2, 1)], [(g.GATE_PRJ_L3, 1)],
    [(g.GATE_ID, 3)], [(g.GATE_ID, 6)], [(g.GATE_TRI_L1, 1), (g.GATE_PRJ_L1, 2)], [(g.GATE_PRJ_L2, 4)], [(g.GATE_PRJ_L3, 5)],
    [(g.GATE_AFF_L1_BITCHECK2, 1)],
]


@pytest.mark.parametrize("parts", STACKS)
def test_dense_map_and_splits(ctx, parts):
    rng = random.Random(len(str(parts)))
    f = ostack(parts)
    nv = 6
    polys = [[rng.randrange(P) for _ in range(1 << nv)] for _ in range(f.n_ins)]
    tabs = [ctx.upload(to_limbs(p)) for p in polys]
    got = [from_limbs(t.download()) for t in ctx.map_dense(parts, tabs)]
    assert got == OP.dense_map(polys, f)
    for var_idx in (("LO", 0), ("LO", 2), ("HI", 0), ("HI", 2), ("HI", nv - 1)):
        for bundle in (1, 3, f.n_outs):
            got = [from_limbs(t.download()) for t in ctx.map_dense(parts, tabs, split=var_idx, bundle_size=bundle)]
            assert got == OP.dense_map_split(polys, f, var_idx, bundle), (var_idx, bundle)
    # map_split_hi == split at HI(0) with one bundle
    l, r = OP.map_split_hi(polys, f)
    got = [from_limbs(t.download()) for t in ctx.map_dense(parts, tabs, split=("HI", 0), bundle_size=f.n_outs)]
    assert got == l + r


def _mk(rng, ctx, n_polys, rowv, colv, dens):
    nrows = (1 << colv) if dens < 2 else rng.randrange(0, 1 << colv) + 1
    lens = [(1 << rowv) if dens == 0 else rng.randrange(0, (1 << rowv) + 1) for _ in range(nrows)]
    pads = [(rng.randrange(P), rng.randrange(P)) for _ in range(n_polys)]
    data = [[[rng.randrange(P) for _ in range(l)] for l in lens] for _ in range(n_polys)]
    opolys = [S.VecVecPolynomial(data[j], pads[j][0], pads[j][1], rowv, colv) for j in range(n_polys)]
    dpolys = [ctx.upload_vecvec([to_limbs(r) if len(r) else np.zeros((0, 4), np.uint64) for r in data[j]], to_limb1(pads[j][0]),
                                to_limb1(pads[j][1]), rowv, colv) for j in range(n_polys)]
    return opolys, dpolys


def _vv_eq(dv, ov):
    rows, rp, cp, rl, cl = dv.download()
    assert [from_limbs(r) if len(r) else [] for r in rows] == ov.data
    assert from_limbs(rp.reshape(1, 4))[0] == ov.row_pad and from_limbs(cp.reshape(1, 4))[0] == ov.col_pad
    assert (rl, cl) == (ov.row_logsize, ov.col_logsize)


VV_STACKS = [[(g.GATE_AFF_L1_BITCHECK2, 1)], [(g.GATE_AFF_L2, 1)], [(g.GATE_AFF_L3, 1)], [(g.GATE_PRJ_L1, 1)], [(g.GATE_PRJ_L2, 1)],
             [(g.GATE_PRJ_L3, 1)], [(g.GATE_ID, 2)], [(g.GATE_ID, 1)]]


@pytest.mark.parametrize("parts", VV_STACKS)
@pytest.mark.parametrize("dens", [0, 1, 2])
def test_vecvec_map_and_split(ctx, parts, dens):
    rng = random.Random(7 * dens + len(str(parts)))
    f = ostack(parts)
    rowv, colv = 4, 3
    opolys, dpolys = _mk(rng, ctx, f.n_ins, rowv, colv, dens)
    for dv, ov in zip(ctx.map_vecvec(parts, dpolys, mode=0), OP.vecvec_map(opolys, f)):
        _vv_eq(dv, ov)
    for bundle in (1, 3):
        for dv, ov in zip(ctx.map_vecvec(parts, dpolys, mode=1, bundle_size=bundle), OP.vecvec_map_split(opolys, f, ("LO", 0), bundle)):
            _vv_eq(dv, ov)


@pytest.mark.parametrize("dens", [0, 1, 2])
def test_vecvec_map_split_to_dense(ctx, dens):
    rng = random.Random(50 + dens)
    parts = [(g.GATE_PRJ_L3, 1)]
    f = ostack(parts)
    opolys, dpolys = _mk(rng, ctx, f.n_ins, 1, 4, dens)
    want = OP.vecvec_map_split_to_dense(opolys, f, ("LO", 0), 3)
    got = [from_limbs(t.download()) for t in ctx.map_vecvec(parts, dpolys, mode=2, bundle_size=3)]
    assert got == want


def test_affine_and_projective_addition_through_the_gates(ctx):
    """bintree_add.rs:401-561 check_affine/projective_point_addition: L1 o L2 o L3 on the device equals the
    twisted Edwards group law (oracle curve arithmetic), for Bandersnatch points."""
    from oracle.pyref import curves as CV

    rng = random.Random(3)
    n = 64
    pts1 = [CV.te_random_point(rng) for _ in range(n)]
    pts2 = [CV.te_random_point(rng) for _ in range(n)]
    cols = [[p[0] for p in pts1], [p[1] for p in pts1], [p[0] for p in pts2], [p[1] for p in pts2]]
    tabs = [ctx.upload(to_limbs(c)) for c in cols]
    l1 = ctx.map_dense([(g.GATE_AFF_L1, 1)], tabs)
    l2 = ctx.map_dense([(g.GATE_AFF_L2, 1)], l1)
    l3 = ctx.map_dense([(g.GATE_AFF_L3, 1)], l2)
    X, Y, Z = [from_limbs(t.download()) for t in l3]
    for i in range(n):
        want = CV.te_add_affine(pts1[i], pts2[i])
        zi = pow(Z[i], -1, P)
        assert (X[i] * zi % P, Y[i] * zi % P) == want
    # projective: (X, Y, Z) + (X, Y, Z) of the sums above with themselves == doubling
    tabs6 = l3 + l3
    p1 = ctx.map_dense([(g.GATE_PRJ_L1, 1)], tabs6)
    p2 = ctx.map_dense([(g.GATE_PRJ_L2, 1)], p1)
    p3 = ctx.map_dense([(g.GATE_PRJ_L3, 1)], p2)
    X2, Y2, Z2 = [from_limbs(t.download()) for t in p3]
    for i in range(n):
        s = CV.te_add_affine(pts1[i], pts2[i])
        want = CV.te_add_affine(s, s)
        zi = pow(Z2[i], -1, P)
        assert (X2[i] * zi % P, Y2[i] * zi % P) == want
