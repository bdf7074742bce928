"""Old API, Shape::full: the round polynomials of FragmentedLincomb (src/protocol/sumcheck.rs:36-156,
eq materialised per round, evaluated at 0..degree+1) are the ones DenseSumcheckObjectSO over EqWrapper(GammaWrapper(f))
produces, so the same device object serves the old SumcheckPolyMap prover (benches/bintree.rs, gkr_msm_simple) on full tables.

CPU: restated FragmentedLincomb == restated new-API object, round by round, for the six twisted-Edwards layer gates of
benches/bintree.rs:49-84 -- and == the plain sum over the hypercube (the old API's own test style, sumcheck.rs:704+).
GPU: the device object's evaluations at 0..3 and final evaluations == the restated FragmentedLincomb."""
import random

import numpy as np
import pytest

from oracle.pyref import gates as G
from oracle.pyref import oldapi as O
from oracle.pyref import sumcheck as S
from oracle.pyref.field import P
from tests.util import from_limbs, to_limb1, to_limbs

BINTREE_GATES = [G.AffL1, G.AffL2, G.AffL3, G.PrjL1, G.PrjL2, G.PrjL3]


def _instance(cls, nv, seed):
    rng = random.Random(seed)
    gate = cls()
    gamma = rng.randrange(P)
    point = [rng.randrange(P) for _ in range(nv)]
    polys = [[rng.randrange(P) for _ in range(1 << nv)] for _ in range(gate.n_ins)]
    evs = [[(o, 0) for o in range(gate.n_outs)]]  # to_multieval: every output claimed at the one point (sumcheck.rs:317-321)
    gamma_pows = O.make_gamma_pows_legacy(gate.n_outs, gamma)
    old = O.FragmentedLincombFull(polys, [point], O.make_folded_f(evs, gamma_pows, gate.exec, gate.n_ins), gate.deg)
    return rng, gate, gamma, point, polys, old


@pytest.mark.parametrize("cls", BINTREE_GATES)
def test_old_api_round_polynomials_equal_new_api(cls):
    nv = 5
    rng, gate, gamma, point, polys, old = _instance(cls, nv, 40 + BINTREE_GATES.index(cls))
    f = G.EqWrapper(G.GammaWrapper(gate, gamma))
    eq = S.eq_poly_sequence_last(point)
    claim = sum(f.exec([p[i] for p in polys] + [eq[i]]) for i in range(1 << nv)) % P
    new = S.DenseSumcheckObjectSO([list(p) for p in polys] + [eq], f, nv, claim)
    for _ in range(nv):
        ev_old = old.unipoly_evals()
        assert (ev_old[0] + ev_old[1]) % P == claim  # the old API evaluates node 0 directly; the new one derives it
        assert new.unipoly() == old.unipoly()
        assert new.last_evals == ev_old
        t = rng.randrange(1 << 128)
        old.bind(t)
        new.bind(t)
        claim = new.claim
    assert new.final_evals()[:-1] == old.final_evals()


@pytest.mark.gpu
@pytest.mark.parametrize("cls", BINTREE_GATES)
def test_device_object_serves_old_api(ctx, cls):
    import gkr_msm_b200 as g

    nv = 7
    rng, gate, gamma, point, polys, old = _instance(cls, nv, 90 + BINTREE_GATES.index(cls))
    tabs = [ctx.upload(to_limbs(p)) for p in polys] + [ctx.eq_table(to_limbs(point))]  # EqPoly::materialize == the eq table
    ev0 = old.unipoly_evals()
    consts = S.make_gamma_pows(gamma, max(gate.n_outs, 2))
    assert consts[:gate.n_outs] == O.make_gamma_pows_legacy(gate.n_outs, gamma)[:gate.n_outs]
    so = ctx.dense_so(g.SO_EQ_GAMMA, gate.gate_id, tabs, nv, to_limb1((ev0[0] + ev0[1]) % P), consts=to_limbs(consts))
    for _ in range(nv):
        assert from_limbs(so.unipoly()) == old.unipoly_evals()
        t = rng.randrange(1 << 128)
        so.bind(to_limb1(t))
        old.bind(t)
    assert from_limbs(so.final_evals())[:-1] == old.final_evals()
