"""GPU parity: DenseSumcheckObjectSO on the device (through the C ABI) vs the oracle, bit-exact on
canonical Montgomery limbs -- every round polynomial, the running claim, every final evaluation and
the serialized proof bytes.  Mirrors src/cleanup/protocols/sumcheck.rs:941-1078 and
src/cleanup/protocols/sumchecks/dense_eq.rs:259-343."""
import random

import numpy as np
import pytest

import gkr_msm_b200 as g
from oracle.pyref import gates as G
from oracle.pyref import sumcheck as S
from oracle.pyref.field import P, SplitMix64, synth_fr
from oracle.pyref.transcript import ProofTranscript2
from tests.util import from_limbs, rand_chal128, to_limb1, to_limbs

pytestmark = pytest.mark.gpu


def _run_rounds(ctx, so_kind, gate_id, oracle_gate, polys, nv, claim, rng, gate_param=0, consts=None, check_tables=False, prelaunch=False):
    tables = [ctx.upload(to_limbs(p)) for p in polys]
    dso = ctx.dense_so(so_kind, gate_id, tables, nv, to_limb1(claim), gate_param=gate_param,
                       consts=None if consts is None else to_limbs(consts))
    if prelaunch:  # small rounds enqueued one round ahead: 128-bit challenges release the waiting kernel, full-width ones cancel it
        dso.set_prelaunch(True)
    oso = S.DenseSumcheckObjectSO(polys, oracle_gate, nv, claim)
    assert dso.degree == oracle_gate.deg and dso.num_polys == oracle_gate.n_ins
    for r in range(nv):
        oso.unipoly()
        ev = from_limbs(dso.unipoly())
        assert ev == oso.last_evals, f"round {r}"
        assert from_limbs(dso.unipoly()) == ev  # second call returns the cache (sumcheck.rs:280-281)
        t = rand_chal128(rng) if r % 2 == 0 else rng.randrange(P)
        oso.bind(t)
        dso.bind(to_limb1(t))
        assert from_limbs(dso.claim.reshape(1, 4))[0] == oso.claim
        assert dso.round == r + 1
    assert from_limbs(dso.final_evals()) == oso.final_evals()
    # the caller's tables are left untouched on the device
    if check_tables:
        for t, p in zip(tables, polys):
            assert from_limbs(t.download()) == p
    dso.destroy()


@pytest.mark.parametrize("nv", [1, 2, 3, 7, 10, 13])
def test_prod3_rounds(ctx, nv):
    rng = random.Random(100 + nv)
    polys = [[rng.randrange(P) for _ in range(1 << nv)] for _ in range(3)]
    f = G.Prod3()
    claim = sum(f.exec([p[i] for p in polys]) for i in range(1 << nv)) % P
    got = from_limbs(ctx.gate_sum(g.SO_PLAIN, g.GATE_PROD3, [ctx.upload(to_limbs(p)) for p in polys]).reshape(1, 4))[0]
    assert got == claim
    _run_rounds(ctx, g.SO_PLAIN, g.GATE_PROD3, f, polys, nv, claim, rng, check_tables=True)


@pytest.mark.parametrize("nargs", [1, 2, 3, 4])
def test_folded_prod_rounds(ctx, nargs):
    rng = random.Random(200 + nargs)
    nv = 8
    gamma = rng.randrange(P)
    f = G.FoldedProd(gamma, nargs)
    polys = [[rng.randrange(P) for _ in range(1 << nv)] for _ in range(2 * nargs)]
    claim = sum(f.exec([p[i] for p in polys]) for i in range(1 << nv)) % P
    _run_rounds(ctx, g.SO_PLAIN, g.GATE_FOLDED_PROD, f, polys, nv, claim, rng, gate_param=nargs, consts=f.gammas)


EQ_GAMMA_GATES = [
    (g.GATE_AFF_L1, G.AffL1), (g.GATE_AFF_L2, G.AffL2), (g.GATE_AFF_L3, G.AffL3),
    (g.GATE_PRJ_L1, G.PrjL1), (g.GATE_PRJ_L2, G.PrjL2), (g.GATE_PRJ_L3, G.PrjL3),
    (g.GATE_AFF_L1_BITCHECK2, G.AffL1BitCheck2), (g.GATE_LOGUP_LAYER, G.LogupLayer), (g.GATE_ADD_INVERSES, G.AddInverses),
]


@pytest.mark.parametrize("gid,cls", EQ_GAMMA_GATES)
def test_eq_gamma_rounds(ctx, gid, cls):
    rng = random.Random(300 + gid)
    nv = 6
    gate = cls()
    gamma = rng.randrange(P)
    f = G.EqWrapper(G.GammaWrapper(gate, gamma))
    point = [rng.randrange(P) for _ in range(nv)]
    polys = [[rng.randrange(P) for _ in range(1 << nv)] for _ in range(gate.n_ins)]
    polys.append(S.eq_poly_sequence_last(point))
    # wrong claims are fine for round-by-round parity (claim only enters p(0)); use the true one
    claim = sum(f.exec([p[i] for p in polys]) for i in range(1 << nv)) % P
    consts = S.make_gamma_pows(gamma, max(gate.n_outs, 2))
    _run_rounds(ctx, g.SO_EQ_GAMMA, gid, f, polys, nv, claim, rng, consts=consts)


def test_zero_and_extreme_values(ctx):
    rng = random.Random(7)
    nv = 5
    specials = [0, 1, P - 1, P - 2, (P - 1) // 2]
    polys = [[rng.choice(specials) for _ in range(1 << nv)] for _ in range(3)]
    f = G.Prod3()
    claim = sum(f.exec([p[i] for p in polys]) for i in range(1 << nv)) % P
    _run_rounds(ctx, g.SO_PLAIN, g.GATE_PROD3, f, polys, nv, claim, rng)


def test_eq_table_matches_oracle(ctx):
    rng = random.Random(11)
    for n in (0, 1, 5, 10, 11, 14):
        pt = [rng.randrange(P) for _ in range(n)]
        mult = rng.randrange(P)
        t = ctx.eq_table(to_limbs(pt) if n else np.zeros((0, 4), np.uint64), to_limb1(mult))
        want = S.eq_poly_sequence_from_multiplier(mult, pt)[-1]
        assert from_limbs(t.download()) == want


def test_synth_table_matches_oracle_stream(ctx):
    t = ctx.synth(1234, 1000)
    got = t.download()
    gen = SplitMix64(1234)
    for i in range(1000):
        v = gen.fr()
        m = int(got[i, 0]) | (int(got[i, 1]) << 64) | (int(got[i, 2]) << 128) | (int(got[i, 3]) << 192)
        assert m == v
        if i % 97 == 0:
            assert synth_fr(1234, i) == v


def test_protocol_errors_like_reference_panics(ctx):
    rng = random.Random(5)
    nv = 3
    polys = [[rng.randrange(P) for _ in range(1 << nv)] for _ in range(3)]
    tabs = [ctx.upload(to_limbs(p)) for p in polys]
    so = ctx.dense_so(g.SO_PLAIN, g.GATE_PROD3, tabs, nv, to_limb1(0))
    with pytest.raises(g.GkrError) as e:  # bind before unipoly panics in the reference (sumcheck.rs:271-274)
        so.bind(to_limb1(5))
    assert e.value.code == g.GKR_ERR_PROTOCOL
    with pytest.raises(g.GkrError):  # final_evals before the end (sumcheck.rs:336)
        so.final_evals()
    for _ in range(nv):
        so.unipoly()
        so.bind(to_limb1(rng.randrange(P)))
    with pytest.raises(g.GkrError):  # "the protocol has already ended" (sumcheck.rs:278)
        so.unipoly()
    # wrong table count / length (sumcheck.rs:255-258)
    with pytest.raises(g.GkrError) as e:
        ctx.dense_so(g.SO_PLAIN, g.GATE_PROD3, tabs[:2], nv, to_limb1(0))
    assert e.value.code == g.GKR_ERR_ARG
    with pytest.raises(g.GkrError):
        ctx.dense_so(g.SO_PLAIN, g.GATE_PROD3, tabs, nv + 1, to_limb1(0))


def test_generic_sumcheck_prove_proof_bytes(ctx):
    """GenericSumcheckProtocol::prove end to end: device object + C++ host transcript vs the oracle's
    prover -- identical proof bytes, point, claim, final evals; then the oracle VERIFIER accepts the
    device-produced proof (sumcheck.rs:941-969)."""
    rng = random.Random(77)
    nv = 9
    polys = [[rng.randrange(P) for _ in range(1 << nv)] for _ in range(3)]
    f = G.Prod3()
    claim = sum(f.exec([p[i] for p in polys]) for i in range(1 << nv)) % P

    tp = ProofTranscript2.start_prover(b"fgstglsp")
    oso = S.DenseSumcheckObjectSO(polys, f, nv, claim)
    (oclaim, opoint), ofinal = S.generic_sumcheck_prove(tp, [3] * nv, claim, oso)
    tp.write_scalars(ofinal)

    tr = g.Transcript(b"fgstglsp")
    dso = ctx.dense_so(g.SO_PLAIN, g.GATE_PROD3, [ctx.upload(to_limbs(p)) for p in polys], nv, to_limb1(claim))
    dclaim, dpoint, dfinal = g.sumcheck_prove(tr, dso, nv)
    tr.write_scalars(dfinal)
    assert from_limbs(dclaim.reshape(1, 4))[0] == oclaim
    assert from_limbs(dpoint) == opoint
    assert from_limbs(dfinal) == ofinal
    proof = tr.proof()
    assert proof == tp.end()
    tv = ProofTranscript2.start_verifier(b"fgstglsp", proof)
    point, evs = S.BareSumcheckSO(f, nv).verify(tv, claim)
    assert evs == [S.evaluate_poly(p, point) for p in polys]


def test_large_synthetic_sumcheck_properties(ctx):
    """Full-size property check (2^20 x 3 tables, too big for the big-int oracle): the device prover's
    messages must satisfy the verifier equations round by round, p_k(0)+p_k(1) == claim_k with
    claim_{k+1} = p_k(r_k), and the final claim must equal prod of the final evals."""
    nv = 20
    tabs = [ctx.synth(1000 + j, 1 << nv) for j in range(3)]
    claim = from_limbs(ctx.gate_sum(g.SO_PLAIN, g.GATE_PROD3, tabs).reshape(1, 4))[0]
    so = ctx.dense_so(g.SO_PLAIN, g.GATE_PROD3, tabs, nv, to_limb1(claim))
    rng = random.Random(1)
    cur = claim
    for _ in range(nv):
        ev = from_limbs(so.unipoly())
        assert (ev[0] + ev[1]) % P == cur
        coeffs = S.unipoly_from_evals(ev)
        t = rand_chal128(rng)
        so.bind(to_limb1(t))
        cur = S.evaluate_univar(coeffs, t)
    fe = from_limbs(so.final_evals())
    assert fe[0] * fe[1] % P * fe[2] % P == cur


@pytest.mark.parametrize("kind", ["prod3", "folded4", "logup", "add_inverses"])
def test_fast_fold_equals_montgomery_fold_at_scale(ctx, kind):
    """The 128-bit-challenge fold (fr_fold128 + host-side rescaling) and the plain Montgomery fold must give
    identical round polynomials and final evaluations -- checked at 2^18, beyond the big-int oracle's reach."""
    nv = 18
    rng = random.Random(4242)
    n = 1 << nv
    if kind == "prod3":
        so_kind, gate, param, consts, ntab = g.SO_PLAIN, g.GATE_PROD3, 0, None, 3
    elif kind == "folded4":
        so_kind, gate, param, ntab = g.SO_PLAIN, g.GATE_FOLDED_PROD, 4, 8
        consts = to_limbs([rng.randrange(P) for _ in range(4)])
    elif kind == "logup":
        so_kind, gate, param, ntab = g.SO_EQ_GAMMA, g.GATE_LOGUP_LAYER, 0, 5
        consts = to_limbs([1, rng.randrange(P)])
    else:
        so_kind, gate, param, ntab = g.SO_EQ_GAMMA, g.GATE_ADD_INVERSES, 0, 3
        consts = to_limbs([1, rng.randrange(P)])
    tabs = [ctx.synth(500 + j, n) for j in range(ntab)]
    claim = ctx.gate_sum(so_kind, gate, tabs, gate_param=param, consts=consts)
    chals = [rand_chal128(rng) for _ in range(nv)]
    chals[5] = rng.randrange(P)  # one full-width challenge in the middle: mixed fast / general folds
    runs = []
    for fast in (True, False):
        ctx.set_fast_fold(fast)
        so = ctx.dense_so(so_kind, gate, tabs, nv, claim, gate_param=param, consts=consts)
        evs = []
        for r in range(nv):
            evs.append(so.unipoly().copy())
            so.bind(to_limb1(chals[r]))
        runs.append((evs, so.final_evals().copy(), so.claim.copy()))
        so.destroy()
    ctx.set_fast_fold(True)
    for a, b in zip(runs[0][0], runs[1][0]):
        assert np.array_equal(a, b)
    assert np.array_equal(runs[0][1], runs[1][1])
    assert np.array_equal(runs[0][2], runs[1][2])


@pytest.mark.parametrize("so_kind,gate,ntab", [(g.SO_PLAIN, g.GATE_PROD3, 3), (g.SO_EQ_GAMMA, g.GATE_AFF_L1, 5), (g.SO_EQ_GAMMA, g.GATE_PRJ_L3, 5)])
def test_c_oracle_parity_across_kernel_switch(ctx, so_kind, gate, ntab):
    """2^15 entries: the first rounds run the throughput kernel, the rounds of <= 4096 items the block-cooperative small-round
    kernel (dense_kernel.cuh) -- every round polynomial and the final evaluations bit-exact against the C oracle."""
    from oracle import coracle

    nv = 15
    tabs = [ctx.synth(40 + j, 1 << nv) for j in range(ntab)]
    host = [coracle.synth_table(40 + j, 1 << nv) for j in range(ntab)]
    gam = coracle.synth_table(5, 16) if so_kind == g.SO_EQ_GAMMA else None
    claim = ctx.gate_sum(so_kind, gate, tabs, consts=gam)
    assert np.array_equal(claim, coracle.gate_sum(1 if so_kind == g.SO_EQ_GAMMA else 0, gate, host, consts=gam))
    so = ctx.dense_so(so_kind, gate, tabs, nv, claim, consts=gam)
    ch = coracle.synth_table(6, nv)
    ch[:, 2:] = 0  # 128-bit challenges: the fast folds
    ch[7] = coracle.synth_table(8, 1)[0]  # and one full-width challenge
    evs = []
    for r in range(nv):
        evs.append(so.unipoly().copy())
        so.bind(ch[r])
    oev, ofe = coracle.dense_sumcheck(1 if so_kind == g.SO_EQ_GAMMA else 0, gate, host, nv, claim, ch, consts=gam)
    for r in range(nv):
        assert np.array_equal(evs[r], oev[r]), f"round {r}"
    assert np.array_equal(so.final_evals(), ofe)



@pytest.mark.parametrize("nv", [3, 7, 10, 13])
def test_prelaunched_rounds_release_and_cancel(ctx, nv):
    """gkr_so_set_prelaunch on a dense object driven from here: the loop alternates 128-bit challenges (the pre-launched kernel is
    released through the mailbox) and full-width ones (it is cancelled and the round is launched the ordinary way) -- every round
    polynomial, claim and final evaluation against the big-int oracle; then the same for an EqWrapper(GammaWrapper(..)) object"""
    rng = random.Random(4100 + nv)
    polys = [[rng.randrange(P) for _ in range(1 << nv)] for _ in range(3)]
    f = G.Prod3()
    claim = sum(f.exec([p[i] for p in polys]) for i in range(1 << nv)) % P
    _run_rounds(ctx, g.SO_PLAIN, g.GATE_PROD3, f, polys, nv, claim, rng, check_tables=True, prelaunch=True)
    gid, cls = EQ_GAMMA_GATES[3]
    gate = cls()
    gamma = rng.randrange(P)
    w = G.EqWrapper(G.GammaWrapper(gate, gamma))
    point = [rng.randrange(P) for _ in range(nv)]
    data = [[rng.randrange(P) for _ in range(1 << nv)] for _ in range(gate.n_ins)]
    data.append(S.eq_poly_sequence_last(point))
    claim = sum(w.exec([p[i] for p in data]) for i in range(1 << nv)) % P
    _run_rounds(ctx, g.SO_EQ_GAMMA, gid, w, data, nv, claim, rng, consts=S.make_gamma_pows(gamma, max(gate.n_outs, 2)), prelaunch=True)
