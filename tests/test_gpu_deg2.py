"""GPU parity for the degree-2 gate sumchecks with factored eq (dense and ragged VecVec), bit-exact vs the
oracle: the reference's own `check_univars` tests (dense_eq.rs:259-343, vecvec_eq.rs:511-600) restated with
the device object in place of the optimised CPU object, plus the prover/verifier round trip
(vecvec_eq.rs:602-660)."""
import random

import numpy as np
import pytest

import gkr_msm_b200 as g
from oracle.pyref import gates as G
from oracle.pyref import sumcheck as S
from oracle.pyref.field import P
from oracle.pyref.transcript import ProofTranscript2
from tests.util import from_limbs, to_limb1, to_limbs

pytestmark = pytest.mark.gpu

BASE = {
    g.GATE_AFF_L1: G.AffL1, g.GATE_AFF_L2: G.AffL2, g.GATE_AFF_L3: G.AffL3, g.GATE_PRJ_L1: G.PrjL1,
    g.GATE_PRJ_L2: G.PrjL2, g.GATE_PRJ_L3: G.PrjL3, g.GATE_TRI_L1: G.TriL1, g.GATE_BITCHECK: G.BitCheck,
    g.GATE_AFF_L1_BITCHECK2: G.AffL1BitCheck2,
}


def oracle_stack(parts):
    f = None
    for gid, rep in parts:
        part = BASE[gid]() if rep == 1 else G.Repeated(BASE[gid](), rep)
        f = part if f is None else G.Stacked(f, part)
    return f


def claims_of(gate, dense_inputs, eqp):
    claims = [0] * gate.n_outs
    for i in range(len(eqp)):
        o = gate.exec([d[i] for d in dense_inputs])
        for k in range(gate.n_outs):
            claims[k] = (claims[k] + o[k] * eqp[i]) % P
    return claims


DENSE_STACKS = [
    [(g.GATE_PRJ_L1, 1)], [(g.GATE_AFF_L1, 1)], [(g.GATE_AFF_L2, 1)], [(g.GATE_AFF_L3, 1)], [(g.GATE_PRJ_L2, 1)], [(g.GATE_PRJ_L3, 1)],
    [(g.GATE_TRI_L1, 1)],                            # triangle layer 0, L1
    [(g.GATE_TRI_L1, 1), (g.GATE_PRJ_L1, 2)],        # triangle layer 2, L1 (triangle_add.rs:199-212)
    [(g.GATE_PRJ_L2, 5)], [(g.GATE_PRJ_L3, 4)],      # triangle L2 / L3 (Repeated(.., layer+3))
    [(g.GATE_AFF_L1_BITCHECK2, 1)],
]


@pytest.mark.parametrize("parts", DENSE_STACKS)
@pytest.mark.parametrize("nv", [1, 4, 7])
def test_dense_deg2_rounds(ctx, parts, nv):
    rng = random.Random(hash(str(parts)) % 1000 + nv)
    gate = oracle_stack(parts)
    data = [[rng.randrange(P) for _ in range(1 << nv)] for _ in range(gate.n_ins)]
    point = [rng.randrange(P) for _ in range(nv)]
    gamma = rng.randrange(P)
    eqp = S.eq_poly_sequence_last(point)
    claims = claims_of(gate, data, eqp)
    oso = S.DenseDeg2SumcheckObjectSO.rlc(data, gate, claims, point, gamma)
    tabs = [ctx.upload(to_limbs(p)) for p in data]
    dso = ctx.deg2_dense_so(parts, tabs, to_limbs(S.make_gamma_pows(gamma, gate.n_outs)), to_limb1(oso.claim), to_limbs(point))
    assert dso.num_polys == gate.n_ins and dso.degree == 3
    for r in range(nv):
        oso.unipoly()
        assert from_limbs(dso.unipoly()) == oso.last_evals, f"round {r}"
        with pytest.raises(g.GkrError):  # second unipoly in a round panics in the reference (dense_eq.rs:109-111)
            dso.unipoly()
        t = rng.randrange(1 << 128) if r % 2 else rng.randrange(P)
        oso.bind(t)
        dso.bind(to_limb1(t))
        assert from_limbs(dso.claim.reshape(1, 4))[0] == oso.claim
    assert from_limbs(dso.final_evals()) == oso.final_evals()
    # inputs untouched
    assert from_limbs(tabs[0].download()) == data[0]
    dso.destroy()  # result slots are a bounded resource: do not wait for the collector


def make_vecvec(rng, dens, rowv, colv, n_polys, pads):
    nrows = (1 << colv) if dens < 2 else rng.randrange(0, 1 << colv) + 1
    lens = [(1 << rowv) if dens == 0 else rng.randrange(0, (1 << rowv) + 1) for _ in range(nrows)]
    if max(lens) == 0:
        lens[0] = 1
    data = [[[rng.randrange(P) for _ in range(l)] for l in lens] for _ in range(n_polys)]
    opolys = [S.VecVecPolynomial(data[j], pads[j][0], pads[j][1], rowv, colv) for j in range(n_polys)]
    return data, opolys, lens


VV_GATES = [g.GATE_PRJ_L1, g.GATE_AFF_L1, g.GATE_AFF_L2, g.GATE_AFF_L3, g.GATE_PRJ_L2, g.GATE_PRJ_L3, g.GATE_AFF_L1_BITCHECK2]


@pytest.mark.parametrize("dens", [0, 1, 2])
@pytest.mark.parametrize("colv", [0, 1, 3])
@pytest.mark.parametrize("gid", VV_GATES)
def test_vecvec_deg2_rounds(ctx, dens, colv, gid):
    rng = random.Random(1000 * dens + 10 * colv + gid)
    nv = 6
    rowv = nv - colv
    gate = BASE[gid]()
    pads = [(rng.randrange(P), rng.randrange(P)) for _ in range(gate.n_ins)]
    if gid == g.GATE_PRJ_L1:
        pads = [(0, 0), (1, 1), (1, 1)] * 2  # the reference's point padding (vecvec.rs:241-263)
    data, opolys, lens = make_vecvec(rng, dens, rowv, colv, gate.n_ins, pads)
    point = [rng.randrange(P) for _ in range(nv)]
    gamma = rng.randrange(P)
    eqp = S.eq_poly_sequence_last(point)
    dense = [p.vec() for p in opolys]
    claims = claims_of(gate, dense, eqp)
    oso = S.VecVecDeg2SumcheckObjectSO.rlc(opolys, gate, claims, point, colv, gamma)
    dpolys = [ctx.upload_vecvec([to_limbs(r) if len(r) else np.zeros((0, 4), np.uint64) for r in data[j]], to_limb1(pads[j][0]),
                                to_limb1(pads[j][1]), rowv, colv) for j in range(gate.n_ins)]
    # upload == VecVecPolynomial::new (odd rows padded)
    rows, rp, cp, rl, cl = dpolys[0].download()
    assert [from_limbs(r) if len(r) else [] for r in rows] == opolys[0].data and (rl, cl) == (rowv, colv)
    dso = ctx.deg2_vecvec_so(gid, dpolys, to_limbs(S.make_gamma_pows(gamma, max(gate.n_outs, 2))), to_limb1(oso.claim), to_limbs(point), colv)
    for r in range(nv):
        oso.unipoly()
        assert from_limbs(dso.unipoly()) == oso.last_evals, f"round {r} (dens={dens}, colv={colv})"
        t = rng.randrange(1 << 128) if r % 2 else rng.randrange(P)
        oso.bind(t)
        dso.bind(to_limb1(t))
        assert from_limbs(dso.claim.reshape(1, 4))[0] == oso.claim
    assert from_limbs(dso.final_evals()) == oso.final_evals()
    dso.destroy()


def test_vecvec_sumcheck_prover_verifier(ctx):
    """vecvec_eq.rs:602-660: VecVecDeg2Sumcheck::prove through the device object and the C++ host transcript,
    proof bytes identical to the oracle prover, oracle verifier accepts, outputs == MLE evaluations."""
    rng = random.Random(31)
    nv, colv = 7, 2
    rowv = nv - colv
    gid, gate = g.GATE_PRJ_L1, G.PrjL1()
    pads = [(0, 0), (1, 1), (1, 1)] * 2
    data, opolys, lens = make_vecvec(rng, 2, rowv, colv, 6, pads)
    point = [rng.randrange(P) for _ in range(nv)]
    dense = [p.vec() for p in opolys]
    eqp = S.eq_poly_sequence_last(point)
    evs = claims_of(gate, dense, eqp)
    prot = S.VecVecDeg2Sumcheck(gate, nv, colv)
    tp = ProofTranscript2.start_prover(b"fgstglsp")
    out_point, out_evs = prot.prove(tp, (point, evs), opolys)
    proof = tp.end()

    tr = g.Transcript(b"fgstglsp")
    gamma = from_limbs(tr.challenge(128).reshape(1, 4))[0]
    gp = S.make_gamma_pows(gamma, gate.n_outs)
    claim = sum(gp[i] * evs[i] for i in range(gate.n_outs)) % P
    dpolys = [ctx.upload_vecvec([to_limbs(r) if len(r) else np.zeros((0, 4), np.uint64) for r in data[j]], to_limb1(pads[j][0]),
                                to_limb1(pads[j][1]), rowv, colv) for j in range(6)]
    dso = ctx.deg2_vecvec_so(gid, dpolys, to_limbs(gp), to_limb1(claim), to_limbs(point), colv)
    dclaim, dpoint, dfinal = g.sumcheck_prove(tr, dso, nv)
    tr.write_scalars(dfinal[:-1])  # poly_evs.pop() drops the eq evaluation (vecvec_eq.rs:445)
    assert tr.proof() == proof
    assert from_limbs(dpoint) == out_point and from_limbs(dfinal[:-1]) == out_evs
    tv = ProofTranscript2.start_verifier(b"fgstglsp", proof)
    vpoint, vevs = prot.verify(tv, (point, evs))
    assert vevs == [S.evaluate_poly(d, vpoint) for d in dense]


@pytest.mark.parametrize("colv", [0, 2])
def test_deg2_prelaunched_rounds_release_and_cancel(ctx, colv):
    """gkr_so_set_prelaunch on the Deg2 objects driven from here with alternating 128-bit challenges (the queued kernel is released
    through the mailbox) and full-width ones (cancelled, ordinary launch): dense object and ragged VecVec object incl. its dense
    tail, every round against the oracle"""
    rng = random.Random(7700 + colv)
    nv = 7
    # dense Deg2
    parts = [(g.GATE_PRJ_L1, 1)]
    gate = oracle_stack(parts)
    data = [[rng.randrange(P) for _ in range(1 << nv)] for _ in range(gate.n_ins)]
    point = [rng.randrange(P) for _ in range(nv)]
    gamma = rng.randrange(P)
    claims = claims_of(gate, data, S.eq_poly_sequence_last(point))
    oso = S.DenseDeg2SumcheckObjectSO.rlc(data, gate, claims, point, gamma)
    dso = ctx.deg2_dense_so(parts, [ctx.upload(to_limbs(p)) for p in data], to_limbs(S.make_gamma_pows(gamma, gate.n_outs)), to_limb1(oso.claim), to_limbs(point))
    dso.set_prelaunch(True)
    for r in range(nv):
        oso.unipoly()
        assert from_limbs(dso.unipoly()) == oso.last_evals, f"dense round {r}"
        t = rng.randrange(1 << 128) if r % 3 else rng.randrange(P)
        oso.bind(t)
        dso.bind(to_limb1(t))
    assert from_limbs(dso.final_evals()) == oso.final_evals()
    dso.destroy()
    # ragged VecVec
    rowv = nv - colv
    gid, vgate = g.GATE_PRJ_L1, BASE[g.GATE_PRJ_L1]()
    pads = [(0, 0), (1, 1), (1, 1)] * 2
    vdata, opolys, lens = make_vecvec(rng, 2, rowv, colv, vgate.n_ins, pads)
    vclaims = claims_of(vgate, [p.vec() for p in opolys], S.eq_poly_sequence_last(point))
    voso = S.VecVecDeg2SumcheckObjectSO.rlc(opolys, vgate, vclaims, point, colv, gamma)
    dpolys = [ctx.upload_vecvec([to_limbs(r) if len(r) else np.zeros((0, 4), np.uint64) for r in vdata[j]], to_limb1(pads[j][0]),
                                to_limb1(pads[j][1]), rowv, colv) for j in range(vgate.n_ins)]
    vdso = ctx.deg2_vecvec_so(gid, dpolys, to_limbs(S.make_gamma_pows(gamma, max(vgate.n_outs, 2))), to_limb1(voso.claim), to_limbs(point), colv)
    vdso.set_prelaunch(True)
    for r in range(nv):
        voso.unipoly()
        assert from_limbs(vdso.unipoly()) == voso.last_evals, f"vecvec round {r}"
        t = rng.randrange(1 << 128) if r % 3 else rng.randrange(P)
        voso.bind(t)
        vdso.bind(to_limb1(t))
    assert from_limbs(vdso.final_evals()) == voso.final_evals()
    vdso.destroy()
