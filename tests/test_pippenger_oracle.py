"""The independent C++ prover (oracle/c/pippenger_oracle.cpp: benchutils::run_pippenger, src/cleanup/protocols/pippenger.rs:499-559,
restated from the reference sources) is pinned bit-for-bit to the python oracle (oracle/pyref) -- proof bytes, output tables, claims and
the deferred pairing pair -- on the committed golden proofs and on live instances covering the reference's edge cases (degenerate
scalars, x == d, y_size not a power of two, clm == y_logsize).  It is the byte-level target of the device prover at the BASELINE sizes
(tests/golden/pippenger_large.json, tests/test_gpu_pippenger.py) and the CPU baseline of bench.py.  No GPU needed."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

from oracle import pippenger_oracle as PO
from oracle.pyref import curves as CV
from oracle.pyref import pippenger as PP
from oracle.pyref.field import P, fq_vec_from_mont_u64, fq_vec_to_mont_u64
from oracle.pyref.transcript import ProofTranscript2
from tests.test_oracle_pippenger import make_instance
from tests.util import from_limbs, to_limbs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def coefs_to_u64(coefs):
    return np.array([[(c >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)] for c in coefs], dtype=np.uint64)


def g1_limbs(pt):
    return fq_vec_to_mont_u64([pt[0], pt[1]]).reshape(12)


def oracle_key(okey, nv):
    return PO.Key(to_limbs([okey.kzg.tau])[0], g1_limbs(okey.kzg.g0), nv, to_limbs([okey.k])[0])


def cpp_prove(points, coefs, r, okey, d, x, nbits, clm):
    key = oracle_key(okey, x + clm)
    points_xy = np.stack([to_limbs([p[0] for p in points]), to_limbs([p[1] for p in points])])
    out = PO.run_pippenger(key, points_xy, coefs_to_u64(coefs), to_limbs(r), d, x, nbits, clm)
    key.close()
    return out


def pair_point(limbs12):
    x, y = fq_vec_from_mont_u64(limbs12.reshape(2, 6))
    return None if (x, y) == (0, 0) else (x, y)


def test_cpp_oracle_reproduces_the_committed_golden_proofs():
    """inputs and expected bytes come from tests/golden/pippenger.json / *.proof (minted by oracle/pyref)"""
    gold = json.load(open(os.path.join(GOLDEN, "pippenger.json")))
    assert len(gold) == 4
    for name, g in gold.items():
        d, x, nbits, clm = g["d_logsize"], g["x_logsize"], g["num_bits"], g["commitment_log_multiplicity"]
        points = [(int(a, 16), int(b, 16)) for a, b in g["points"]]
        coefs = [int(c, 16) for c in g["coefs"]]
        r = [int(v, 16) for v in g["r"]]
        okey = PP.KnucklesKey(PP.KzgKey(int(g["tau"], 16), (int(g["g0"][0], 16), int(g["g0"][1], 16)), 2 * (1 << (x + clm)) - 1), x + clm, 2)
        out = cpp_prove(points, coefs, r, okey, d, x, nbits, clm)
        proof = open(os.path.join(GOLDEN, name + ".proof"), "rb").read()
        assert out["proof"] == proof, name
        assert hashlib.sha256(out["proof"]).hexdigest() == g["proof_sha256"]
        assert [from_limbs(t) for t in out["dense_output"]] == [[int(v, 16) for v in t] for t in g["dense_output"]]
        assert from_limbs(out["claim_evs"]) == [int(v, 16) for v in g["claims_evs"]]
        okey.kzg.verify_pair((pair_point(out["pair"][0]), pair_point(out["pair"][1])))  # A == tau * B (mock setup)


@pytest.mark.parametrize("d,x,nbits,clm", [(3, 5, 16, 1), (3, 3, 6, 0), (4, 4, 8, 1), (2, 5, 9, 0), (3, 3, 24, 3), (2, 4, 15, 2)])
def test_cpp_oracle_equals_python_oracle_live(d, x, nbits, clm):
    rng = random.Random(77000 + 1000 * d + 100 * x + 10 * nbits + clm)
    cfg, points, coefs, r, okey = make_instance(rng, d, x, nbits, clm)
    tp = ProofTranscript2.start_prover(b"fgstglsp")
    odense, oclaims = PP.run_pippenger(tp, points, coefs, cfg, r, okey)
    out = cpp_prove(points, coefs, r, okey, d, x, nbits, clm)
    assert out["proof"] == tp.end()
    assert [from_limbs(t) for t in out["dense_output"]] == odense
    assert from_limbs(out["claim_evs"]) == list(oclaims[1])


@pytest.mark.parametrize("kind", ["zeros", "ones", "same", "one-hot"])
def test_cpp_oracle_degenerate_scalars(kind):
    """collisions and empty buckets: every point of a digit row lands in ONE bucket (counters up to 2^x - 1)"""
    d, x, nbits, clm = 3, 5, 16, 1
    rng = random.Random(7100 + len(kind))
    cfg = PP.pippenger_config(d, x, nbits, clm)
    n = 1 << x
    points = [CV.te_random_point(rng) for _ in range(n)]
    coefs = {"zeros": [0] * n, "ones": [(1 << nbits) - 1] * n, "same": [rng.randrange(1 << nbits)] * n, "one-hot": [0] * (n - 1) + [5]}[kind]
    r = [rng.randrange(P) for _ in range(cfg["y_logsize"])]
    okey = PP.KnucklesKey(PP.KzgKey(rng.randrange(1, P), CV.g1_mul(rng.randrange(1, P), CV.G1_GEN), 2 * (1 << (x + clm)) - 1), x + clm, 2)
    tp = ProofTranscript2.start_prover(b"fgstglsp")
    PP.run_pippenger(tp, points, coefs, cfg, r, okey)
    out = cpp_prove(points, coefs, r, okey, d, x, nbits, clm)
    assert out["proof"] == tp.end()


def test_cpp_oracle_rejects_a_single_digit_row_like_the_reference():
    """nbits <= d_logsize: LogupMainphaseProtocol::new panics ("logsizes must be non-increasing", logup_mainphase.rs:75-77)"""
    rng = random.Random(3)
    d, x, nbits, clm = 3, 4, 3, 0
    cfg, points, coefs, r, okey = make_instance(rng, d, x, nbits, clm)
    with pytest.raises(PO.OracleError, match="non-increasing"):
        cpp_prove(points, coefs, r, okey, d, x, nbits, clm)


def test_key_and_msm_against_the_python_oracle():
    """mock_setup powers of tau (kzg.rs:84-97), commit == MSM (kzg.rs:123-126) incl. the compressed encoding, synthetic points"""
    rng = random.Random(11)
    tau = rng.randrange(1, P)
    g0 = CV.g1_mul(rng.randrange(1, P), CV.G1_GEN)
    okey = PP.KnucklesKey(PP.KzgKey(tau, g0, 2 * (1 << 6) - 1), 6, 2)
    key = oracle_key(okey, 6)
    for i in (0, 1, 2, 63, 126):
        assert pair_point(key.point(i)) == CV.g1_mul(pow(tau, i, P), g0)
    for n in (1, 2, 33, 100, 127):  # below and above the MSM's slicing threshold, full-width and small scalars
        poly = [rng.randrange(P) if i % 3 else (P - rng.randrange(50)) for i in range(n)]
        assert key.commit_bytes(to_limbs(poly)) == PP.g1_serialize(okey.kzg.commit(poly))
    assert key.commit_bytes(to_limbs([0, 0, 0])) == PP.g1_serialize(None)
    key.close()
    k0, step = 0x1234567 + 5, 0x9E3779B97F4A7C15
    pts = PO.te_arithmetic_progression(k0, step, 40)
    xs, ys = from_limbs(pts[0]), from_limbs(pts[1])
    for i in (0, 1, 2, 39):
        assert (xs[i], ys[i]) == CV.te_mul((k0 + i * step) % CV.TE_SUBGROUP_ORDER, CV.TE_GEN)


def test_large_golden_is_reproducible_from_its_recipe():
    """tests/golden/pippenger_large.json (the device prover's byte target at the BASELINE sizes) really is what
    tests/golden/make_golden_large.py mints: re-mint the smallest entry here (a few seconds of CPU)."""
    from tests.golden import make_golden_large as M
    gold = json.load(open(os.path.join(GOLDEN, "pippenger_large.json")))
    for cfg in M.CONFIGS:
        assert M.name_of(*cfg) in gold
    got = M.mint(5, 10, 64, 2)
    want = gold[M.name_of(5, 10, 64, 2)]
    for k in ("proof_len", "proof_sha256", "dense_output_sha256", "claim_evs_sha256", "pair_sha256"):
        assert got[k] == want[k], k
