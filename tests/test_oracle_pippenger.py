"""Oracle self-checks for the top-level protocol (no GPU): the restated prover's proof is accepted by the restated
verifier and the proved result equals an independent MSM -- the reference's own end-to-end test
(src/cleanup/protocols/pippenger.rs:621-645), plus its pushforward / logup / opening unit tests
(pushforward.rs:990-1189, logup_mainphase.rs:252-338, opening.rs:159-195, verifier_polys.rs:150-197)."""
import random

import pytest

from oracle.pyref import curves as CV
from oracle.pyref import pippenger as PP
from oracle.pyref import sumcheck as S
from oracle.pyref.field import P
from oracle.pyref.transcript import ProofTranscript2


def make_key(rng, num_vars):
    tau = rng.randrange(1, P)
    g0 = CV.g1_mul(rng.randrange(1, P), CV.G1_GEN)
    return PP.KnucklesKey(PP.KzgKey(tau, g0, 2 * (1 << num_vars) - 1), num_vars, 2)


def make_instance(rng, d_logsize, x_logsize, num_bits, clm):
    cfg = PP.pippenger_config(d_logsize, x_logsize, num_bits, clm)
    points = [CV.te_random_point(rng) for _ in range(1 << x_logsize)]
    coefs = [rng.randrange(1 << num_bits) for _ in range(1 << x_logsize)]
    r = [rng.randrange(P) for _ in range(cfg["y_logsize"])]
    key = make_key(rng, x_logsize + clm)
    return cfg, points, coefs, r, key


@pytest.mark.parametrize("d,x,nbits,clm", [(2, 3, 6, 0), (2, 3, 8, 1), (3, 4, 7, 0), (2, 2, 8, 2)])
def test_pippenger_prove_verify(d, x, nbits, clm):
    rng = random.Random(1000 * d + 100 * x + 10 * nbits + clm)
    cfg, points, coefs, r, key = make_instance(rng, d, x, nbits, clm)
    tp = ProofTranscript2.start_prover(b"fgstglsp")
    dense_output, claims = PP.run_pippenger(tp, points, coefs, cfg, r, key)
    proof = tp.end()
    tv = ProofTranscript2.start_verifier(b"fgstglsp", proof)
    expected = CV.te_msm(points, coefs)
    got = PP.verify_pippenger(tv, cfg, dense_output, claims, key, expected)
    assert tv.ctr == len(proof)
    assert got == expected
    # a corrupted proof must be rejected
    bad = bytearray(proof)
    bad[len(bad) // 2] ^= 1
    with pytest.raises(AssertionError):
        PP.verify_pippenger(ProofTranscript2.start_verifier(b"fgstglsp", bytes(bad)), cfg, dense_output, claims, key, expected)


def test_commit_shortcut_equals_srs_route():
    """the oracle's (sum coeff tau^i) g0 shortcut == MSM over explicit SRS points == bucket sums + running sums"""
    rng = random.Random(5)
    cfg, points, coefs, r, key = make_instance(rng, 2, 3, 6, 1)
    a = PP.PushForwardState(points, coefs, cfg["y_size"], cfg["y_logsize"], 2, 3, 1, key)
    b = PP.PushForwardState(points, coefs, cfg["y_size"], cfg["y_logsize"], 2, 3, 1, key, literal_commits=True)
    assert a.c_comm == b.c_comm and a.d_comm == b.d_comm
    poly = [rng.randrange(P) for _ in range(9)]
    assert key.kzg.commit(poly) == key.kzg.commit_literal(poly)
    for pt in a.c_comm + [None, CV.G1_GEN]:
        assert PP.g1_deserialize(PP.g1_serialize(pt)) == pt


def test_verifier_polys():
    rng = random.Random(6)
    for nv in (1, 3, 4):
        r = [rng.randrange(P) for _ in range(nv)]
        pt = [rng.randrange(P) for _ in range(nv)]
        for k in range(0, (1 << nv) + 1):
            ev = PP.eq_trunc_evals(nv, k, r)
            assert S.evaluate_poly(ev, pt) == PP.eq_trunc_evaluate(nv, k, r, pt)
            sel = [1] * k + [0] * ((1 << nv) - k)
            assert S.evaluate_poly(sel, pt) == PP.selector_evaluate(nv, k, pt)


def test_logup_mainphase_and_knuckles_opening():
    rng = random.Random(9)
    logsizes = [3, 3, 2, 1]
    proto = PP.LogupMainphase(logsizes)
    inp = [[[rng.randrange(P) for _ in range(1 << ls)], [rng.randrange(1, P) for _ in range(1 << ls)]] for ls in logsizes]
    total = sum(n * pow(dn, -1, P) for arr in inp for n, dn in zip(arr[0], arr[1])) % P
    tp = ProofTranscript2.start_prover(b"x")
    pc = proto.prove(tp, total, [[list(a[0]), list(a[1])] for a in inp])
    tv = ProofTranscript2.start_verifier(b"x", tp.end())
    assert proto.verify(tv, total) == pc
    # claims are evaluations of the inputs: first = the two halves merged (HI split still to be applied by the caller)
    for (point, evs), arr in zip(pc[1:], inp[2:]):
        assert evs == [S.evaluate_poly(arr[0], point), S.evaluate_poly(arr[1], point)]
    # Knuckles
    nv = 3
    key = make_key(rng, nv)
    poly = [rng.randrange(P) for _ in range(1 << nv)]
    point = [rng.randrange(P) for _ in range(nv)]
    claim = (key.commit(poly), point, S.evaluate_poly(poly, point))
    tp = ProofTranscript2.start_prover(b"k")
    pair = PP.KnucklesOpening(key).prove(tp, claim, poly)
    key.kzg.verify_pair(pair)
    tv = ProofTranscript2.start_verifier(b"k", tp.end())
    assert PP.KnucklesOpening(key).verify(tv, claim) == pair
