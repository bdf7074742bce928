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


# ---- the round-by-round old-API bintree prover (benches/bintree.rs; src/protocol/bintree.rs:467-583 restated) ------------
def _bintree_instance(log_n, seed):
    rng = random.Random(seed)
    n = 1 << log_n
    tables = [[rng.randrange(P) for _ in range(n)] for _ in range(2)]  # (x, y) columns; the gates are polynomial maps of any values
    layers = O.bintree_layers(log_n)
    trace, output = O.bintree_witness(tables, layers, log_n)
    point = [rng.randrange(P)]
    evs = [O.evaluate_full(p, point) for p in output]
    return tables, layers, trace, output, point, evs


@pytest.mark.parametrize("log_n", [3, 5])
def test_old_bintree_prover_vs_verifier(log_n):
    """prover_vs_verifier (bintree.rs:531-583): the final claims are the INPUT tables evaluated at the final point, the
    verifier accepts and ends in the same claim and the same transcript state; a corrupted round polynomial is rejected"""
    tables, layers, trace, output, point, evs = _bintree_instance(log_n, 500 + log_n)
    assert len(output) == 3 and len(output[0]) == 2  # one variable left: two points that were never added together
    tp = O.OldTranscript(b"test")
    (fpoint, fevs), proofs = O.bintree_prove(tp, point, evs, trace, layers, log_n)
    assert fevs == [O.evaluate_full(t, fpoint) for t in tables]
    tv = O.OldTranscript(b"test")
    vpoint, vevs = O.bintree_verify(tv, point, evs, proofs, layers, log_n)
    assert (vpoint, vevs) == (fpoint, fevs)
    assert tv.challenge_scalar(b"end") == tp.challenge_scalar(b"end")
    bad = [p if p is None else ([list(q) for q in p[0]], list(p[1])) for p in proofs]
    k = next(i for i, p in enumerate(bad) if p is not None and p[0])
    bad[k][0][0][0] = (bad[k][0][0][0] + 1) % P
    with pytest.raises(AssertionError):
        O.bintree_verify(O.OldTranscript(b"test"), point, evs, bad, layers, log_n)


def test_old_bintree_witness_is_the_point_sum():
    """witness_generation (bintree.rs:467-529) in spirit: the layers really add curve points -- with on-curve inputs the
    output columns are the projective sums of the two halves of the point list"""
    from oracle.pyref import curves as CV
    rng = random.Random(9)
    log_n = 3
    pts = [CV.te_random_point(rng) for _ in range(1 << log_n)]
    layers = O.bintree_layers(log_n)
    _, out = O.bintree_witness([[p[0] for p in pts], [p[1] for p in pts]], layers, log_n)
    for k in range(2):  # even/odd splits: output k sums the points whose index has top... collect by construction below
        got = CV.te_to_affine((out[0][k], out[1][k], out[2][k]))
        # after log_n - 1 even/odd splits entry k of the output holds the points with index = k mod 2 ... in bit-reversed order:
        # split i pairs (2j, 2j+1) of the CURRENT table, so the first addition joins neighbours; entry k sums indices [4k, 4k+4)
        acc = (0, 1, 1)
        for p in pts[4 * k:4 * k + 4]:
            acc = CV.te_add_proj(acc, (p[0], p[1], 1))
        assert got == CV.te_to_affine(acc)


@pytest.mark.gpu
@pytest.mark.parametrize("log_n", [4, 9])
def test_device_old_bintree_equals_oracle(ctx, log_n):
    """BASELINE config[4] flow on the device: witness tables, every round polynomial, the final claim and the transcript
    state equal the restated old API"""
    import gkr_msm_b200 as g
    from gkr_msm_b200 import oldapi as DO

    tables, layers, trace, output, point, evs = _bintree_instance(log_n, 700 + log_n)
    dtabs = [ctx.upload(to_limbs(t)) for t in tables]
    dlayers = DO.bintree_layers(log_n)
    dtrace, dout = DO.bintree_witness(ctx, dtabs, dlayers, log_n)
    assert [from_limbs(t.download()) for t in dout] == output
    assert [[from_limbs(t.download()) for t in lay] for lay in dtrace[:5]] == trace[:5]
    tp = O.OldTranscript(b"test")
    (fpoint, fevs), proofs = O.bintree_prove(tp, point, evs, trace, layers, log_n)
    tr = g.Transcript(b"test")
    (dpoint, devs), dproofs = DO.bintree_prove(ctx, tr, to_limbs(point), to_limbs(evs), dtrace, dlayers, log_n)
    assert (dpoint, devs) == (fpoint, fevs)
    assert dproofs == proofs
    assert from_limbs(tr.challenge_scalar_old(b"end").reshape(1, 4))[0] == tp.challenge_scalar(b"end")
