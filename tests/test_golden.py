"""Golden fixtures (tests/golden/, minted by tests/golden/make_golden.py from the python oracle).

CPU part: the oracle still reproduces every committed fixture (regression pin of the checker itself) and the committed
proofs are accepted by the oracle verifier.  GPU part: the device prover / kernels, driven through the C ABI, reproduce the
committed bytes WITHOUT running the python prover -- the same comparison a Rust-side parity harness would make against proof
bytes dumped by the reference."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

from oracle.pyref import curves as CV
from oracle.pyref import gates as G
from oracle.pyref import pippenger as PP
from oracle.pyref import sumcheck as S
from oracle.pyref.field import P
from oracle.pyref.transcript import ProofTranscript2
from tests.golden import make_golden as MG
from tests.util import from_limbs, to_limbs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def ints(xs):
    return [int(v, 16) for v in xs]


PIP = load("pippenger.json")


@pytest.mark.parametrize("name", sorted(PIP))
def test_oracle_reproduces_pippenger_golden(name):
    e = PIP[name]
    with open(os.path.join(GOLD, name + ".proof"), "rb") as f:
        gold = f.read()
    assert len(gold) == e["proof_len"] and hashlib.sha256(gold).hexdigest() == e["proof_sha256"]
    d, x, nbits, clm = e["d_logsize"], e["x_logsize"], e["num_bits"], e["commitment_log_multiplicity"]
    cfg, points, coefs, r, key = MG.pippenger_instance(d, x, nbits, clm)
    # the committed inputs are the seeded instance
    assert [[MG.hx(p[0]), MG.hx(p[1])] for p in points] == e["points"] and [MG.hx(c) for c in coefs] == e["coefs"]
    assert MG.hx(key.kzg.tau) == e["tau"]
    tp = ProofTranscript2.start_prover(b"fgstglsp")
    dense, claims = PP.run_pippenger(tp, points, coefs, cfg, r, key)
    assert tp.end() == gold
    assert [[MG.hx(v) for v in t] for t in dense] == e["dense_output"]
    assert [MG.hx(v) for v in claims[1]] == e["claims_evs"]
    # and the committed proof verifies against the committed MSM result
    expected = tuple(ints(e["msm_result"]))
    tv = ProofTranscript2.start_verifier(b"fgstglsp", gold)
    assert PP.verify_pippenger(tv, cfg, dense, claims, key, expected) == expected


def test_oracle_reproduces_dense_golden():
    gold = load("dense_sumcheck.json")
    for name, e in gold.items():
        nv = e["num_vars"]
        tabs = [MG.synth_table(s, 1 << nv) for s in e["table_seeds"]]
        so = S.DenseSumcheckObjectSO(tabs, G.Prod3(), nv, int(e["claim"], 16))
        for rnd, t in zip(e["rounds"], ints(e["challenges"])):
            assert [MG.hx(c) for c in so.unipoly()] == rnd["coeffs"]
            assert [MG.hx(c) for c in so.last_evals] == rnd["evals_0_to_deg"]
            so.bind(t)
        assert [MG.hx(v) for v in so.final_evals()] == e["final_evals"]


def test_c_oracle_reproduces_dense_golden():
    from oracle import coracle
    gold = load("dense_sumcheck.json")
    for name, e in gold.items():
        nv = e["num_vars"]
        tabs = [coracle.synth_table(s, 1 << nv) for s in e["table_seeds"]]
        ch = to_limbs(ints(e["challenges"]))
        ev, fe = coracle.dense_sumcheck(0, G.GATE_PROD3, tabs, nv, to_limbs([int(e["claim"], 16)])[0], ch)
        for r, rnd in enumerate(e["rounds"]):
            assert from_limbs(ev[r]) == ints(rnd["evals_0_to_deg"]), (name, r)
        assert from_limbs(fe) == ints(e["final_evals"])


def test_oracle_reproduces_msm_eq_golden():
    e = load("msm_g1.json")
    key = PP.KzgKey(int(e["tau"], 16), tuple(ints(e["g0"])), e["n"])
    for kind in ("full", "small"):
        sc = ints(e[kind + "_scalars"])
        c = key.commit_literal(sc)
        assert c == tuple(ints(e[kind + "_commit"])) == key.commit(sc)
        assert PP.g1_serialize(c).hex() == e[kind + "_commit_bytes"]
        assert PP.g1_deserialize(bytes.fromhex(e[kind + "_commit_bytes"])) == c
    q = load("eq_table.json")
    eq = S.eq_poly_sequence_last(ints(q["point"]))
    assert [MG.hx(v) for v in eq[:8]] == q["first8"] and MG.hx(eq[-1]) == q["last"] and MG.hx(sum(eq) % P) == q["sum"]
    assert hashlib.sha256("\n".join(MG.hx(v) for v in eq).encode()).hexdigest() == q["sha256_of_be_hex_lines"]


# ------------------------------------------------------------------------------------------------- GPU
def _coefs_to_u64(coefs):
    return np.array([[(c >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)] for c in coefs], dtype=np.uint64)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(PIP))
def test_device_proof_equals_golden_bytes(ctx, name):
    """inputs and expected bytes come from the committed files only (no oracle prover in the loop)"""
    import gkr_msm_b200 as g
    from gkr_msm_b200 import pippenger as DPP

    e = PIP[name]
    with open(os.path.join(GOLD, name + ".proof"), "rb") as f:
        gold = f.read()
    d, x, nbits, clm = e["d_logsize"], e["x_logsize"], e["num_bits"], e["commitment_log_multiplicity"]
    pts = [ints(p) for p in e["points"]]
    points_xy = np.stack([to_limbs([p[0] for p in pts]), to_limbs([p[1] for p in pts])])
    coefs = _coefs_to_u64(ints(e["coefs"]))
    nv = x + clm
    kzg = DPP.KzgKey.mock_setup(ctx, int(e["tau"], 16), tuple(ints(e["g0"])), 2 * (1 << nv) - 1)
    key = DPP.KnucklesKey(ctx, kzg, nv, 2)
    tr = g.Transcript(b"fgstglsp")
    ndense, nevs, npair = g.run_pippenger_native(ctx, tr, kzg.srs, kzg.g0, key.dev, points_xy, coefs, d, x, nbits, clm, to_limbs(ints(e["r"])))
    proof = tr.proof()
    assert hashlib.sha256(proof).hexdigest() == e["proof_sha256"]
    assert proof == gold
    assert [[MG.hx(v) for v in from_limbs(t)] for t in ndense] == e["dense_output"]
    assert [MG.hx(v) for v in from_limbs(nevs)] == e["claims_evs"]


@pytest.mark.gpu
def test_device_dense_sumcheck_equals_golden(ctx):
    import gkr_msm_b200 as g

    gold = load("dense_sumcheck.json")
    for name, e in gold.items():
        nv = e["num_vars"]
        tabs = [ctx.synth(s, 1 << nv) for s in e["table_seeds"]]
        so = ctx.dense_so(g.SO_PLAIN, g.GATE_PROD3, tabs, nv, to_limbs([int(e["claim"], 16)])[0])
        for rnd, t in zip(e["rounds"], ints(e["challenges"])):
            assert [MG.hx(v) for v in from_limbs(so.unipoly())] == rnd["evals_0_to_deg"], name
            so.bind(to_limbs([t])[0])
        assert [MG.hx(v) for v in from_limbs(so.final_evals())] == e["final_evals"]


@pytest.mark.gpu
def test_device_msm_eq_equals_golden(ctx):
    import gkr_msm_b200 as g
    from gkr_msm_b200 import hostmath as H
    from tests.test_gpu_msm import res_to_point

    e = load("msm_g1.json")
    g0 = tuple(ints(e["g0"]))
    srs = g.Srs.mock_setup(ctx, to_limbs([int(e["tau"], 16)])[0], H.g1_to_limbs(g0), e["n"])
    for kind in ("full", "small"):
        got = res_to_point(srs.msm(ctx.upload(to_limbs(ints(e[kind + "_scalars"])))))
        assert got == tuple(ints(e[kind + "_commit"]))
        assert H.g1_serialize(got).hex() == e[kind + "_commit_bytes"]
    q = load("eq_table.json")
    eq = from_limbs(ctx.eq_table(to_limbs(ints(q["point"]))).download())
    assert hashlib.sha256("\n".join(MG.hx(v) for v in eq).encode()).hexdigest() == q["sha256_of_be_hex_lines"]
