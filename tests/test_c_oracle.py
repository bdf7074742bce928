"""CPU tests pinning the plain-C oracle (oracle/c/gkr_oracle.c, the CPU baseline port) against the
python big-int oracle: arithmetic, every gate through the dense sumcheck object, eq tables, generator."""
import random

import pytest

from oracle import coracle
from oracle.pyref import gates as G
from oracle.pyref import sumcheck as S
from oracle.pyref.field import P, SplitMix64
from tests.util import from_limbs, to_limb1, to_limbs


def test_fr_mul_against_bigint():
    rng = random.Random(1)
    vals = [0, 1, P - 1, P - 2] + [rng.randrange(P) for _ in range(200)]
    for a in vals[:8]:
        for b in vals:
            got = from_limbs(coracle.fr_mul(to_limb1(a), to_limb1(b)).reshape(1, 4))[0]
            assert got == a * b % P


CASES = [
    (0, 10, G.Prod3, 0), (1, 0, G.AffL1, 0), (1, 1, G.AffL2, 0), (1, 2, G.AffL3, 0), (1, 3, G.PrjL1, 0), (1, 4, G.PrjL2, 0),
    (1, 5, G.PrjL3, 0), (1, 13, G.AffL1BitCheck2, 0), (1, 8, G.LogupLayer, 0), (1, 9, G.AddInverses, 0), (1, 6, G.TriL1, 0),
]


@pytest.mark.parametrize("so_kind,gid,cls,param", CASES)
def test_dense_sumcheck_c_vs_python(so_kind, gid, cls, param):
    rng = random.Random(10 + gid)
    nv = 5
    gamma = rng.randrange(P)
    if so_kind == 0:
        f = cls()
        polys = [[rng.randrange(P) for _ in range(1 << nv)] for _ in range(f.n_ins)]
        consts = None
    else:
        gate = cls()
        f = G.EqWrapper(G.GammaWrapper(gate, gamma))
        polys = [[rng.randrange(P) for _ in range(1 << nv)] for _ in range(gate.n_ins + 1)]
        consts = to_limbs(S.make_gamma_pows(gamma, 16))
    claim = sum(f.exec([p[i] for p in polys]) for i in range(1 << nv)) % P
    tabs = [to_limbs(p) for p in polys]
    assert from_limbs(coracle.gate_sum(so_kind, gid, tabs, param, consts).reshape(1, 4))[0] == claim
    chals = [rng.randrange(P) for _ in range(nv)]
    ev, fe = coracle.dense_sumcheck(so_kind, gid, tabs, nv, to_limb1(claim), to_limbs(chals), param, consts)
    so = S.DenseSumcheckObjectSO(polys, f, nv, claim)
    for r in range(nv):
        so.unipoly()
        assert from_limbs(ev[r]) == so.last_evals
        so.bind(chals[r])
    assert from_limbs(fe) == so.final_evals()


def test_folded_prod_c_vs_python():
    rng = random.Random(3)
    nv, nargs = 4, 4
    gamma = rng.randrange(P)
    f = G.FoldedProd(gamma, nargs)
    polys = [[rng.randrange(P) for _ in range(1 << nv)] for _ in range(2 * nargs)]
    claim = sum(f.exec([p[i] for p in polys]) for i in range(1 << nv)) % P
    chals = [rng.randrange(P) for _ in range(nv)]
    ev, fe = coracle.dense_sumcheck(0, 11, [to_limbs(p) for p in polys], nv, to_limb1(claim), to_limbs(chals), nargs, to_limbs(f.gammas))
    so = S.DenseSumcheckObjectSO(polys, f, nv, claim)
    for r in range(nv):
        so.unipoly()
        assert from_limbs(ev[r]) == so.last_evals
        so.bind(chals[r])
    assert from_limbs(fe) == so.final_evals()


def test_eq_and_synth():
    rng = random.Random(4)
    pt = [rng.randrange(P) for _ in range(7)]
    mult = rng.randrange(P)
    assert from_limbs(coracle.eq_table(to_limbs(pt), to_limb1(mult))) == S.eq_poly_sequence_from_multiplier(mult, pt)[-1]
    got = coracle.synth_table(99, 300)
    gen = SplitMix64(99)
    for i in range(300):
        v = gen.fr()
        assert [int(x) for x in got[i]] == [(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)]
