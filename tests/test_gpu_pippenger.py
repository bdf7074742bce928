"""GPU parity for the TOP-LEVEL protocol (everything examples/pippenger runs between build_pippenger_data and
verify_pippenger): the device prover's proof bytes, output tables, claims and final pairing pair are identical to the
oracle restatement's on the same points / scalars / SRS, and the oracle VERIFIER accepts the device-made proof and
recovers the expected MSM (the reference's own end-to-end test, src/cleanup/protocols/pippenger.rs:621-645)."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

import gkr_msm_b200 as g
from gkr_msm_b200 import hostmath as H
from gkr_msm_b200 import pippenger as DPP
from oracle.pyref import curves as CV
from oracle.pyref import pippenger as PP
from oracle.pyref.field import P
from oracle.pyref.transcript import ProofTranscript2
from tests.test_gpu_msm import aff_to_limbs, res_to_point
from tests.test_oracle_pippenger import make_instance
from tests.util import from_limbs, to_limbs

pytestmark = pytest.mark.gpu

# digests minted by the INDEPENDENT C++ prover (oracle/c/pippenger_oracle.cpp via tests/golden/make_golden_large.py, which never
# imports the product package): the byte-level target of the device prover at the BASELINE sizes
LARGE_GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pippenger_large.json")))


def coefs_to_u64(coefs):
    return np.array([[(c >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)] for c in coefs], dtype=np.uint64)


def test_mock_setup_matches_powers_of_tau(ctx):
    rng = random.Random(3)
    tau = rng.randrange(1, P)
    g0 = CV.g1_mul(rng.randrange(1, P), CV.G1_GEN)
    srs = g.Srs.mock_setup(ctx, to_limbs([tau])[0], H.g1_to_limbs(g0), 37)
    got = [res_to_point(r) for r in srs.download_affine()]
    assert got == [CV.g1_mul(pow(tau, i, P), g0) for i in range(37)]
    assert H.g1_from_limbs(aff_to_limbs([g0])[0]) == g0
    for pt in got[:4] + [None]:
        assert H.g1_serialize(pt) == PP.g1_serialize(pt)


def test_scalar_digits():
    rng = random.Random(4)
    for d, y_size, nbits in [(3, 5, 15), (8, 16, 128), (10, 13, 128), (7, 36, 252), (8, 32, 253)]:
        coefs = [rng.randrange(1 << nbits) for _ in range(50)]
        got = DPP.scalar_digits(coefs_to_u64(coefs), y_size, d)
        for y in range(y_size):
            assert [int(v) for v in got[y]] == [(c >> (y * d)) & ((1 << d) - 1) for c in coefs]


@pytest.mark.parametrize("d,x,nbits,clm", [(2, 3, 6, 0), (2, 3, 8, 1), (3, 4, 7, 0), (2, 2, 8, 2), (3, 5, 16, 1),
                                           (3, 3, 6, 0), (4, 4, 8, 1), (2, 5, 9, 0), (3, 3, 24, 3)])  # x == d, y_size not a power of two, clm == y_logsize
def test_pippenger_device_vs_oracle(ctx, d, x, nbits, clm):
    rng = random.Random(1000 * d + 100 * x + 10 * nbits + clm)
    cfg, points, coefs, r, okey = make_instance(rng, d, x, nbits, clm)
    tp = ProofTranscript2.start_prover(b"fgstglsp")
    odense, oclaims = PP.run_pippenger(tp, points, coefs, cfg, r, okey)
    oproof = tp.end()

    nv = x + clm
    kzg = DPP.KzgKey.mock_setup(ctx, okey.kzg.tau, okey.kzg.g0, 2 * (1 << nv) - 1)
    key = DPP.KnucklesKey(ctx, kzg, nv, 2)
    points_xy = np.stack([to_limbs([p[0] for p in points]), to_limbs([p[1] for p in points])])
    tr = g.Transcript(b"fgstglsp")
    ddense, dclaims, dpair = DPP.run_pippenger(ctx, tr, points_xy, coefs_to_u64(coefs), cfg, r, key)
    assert [from_limbs(t.download()) for t in ddense] == odense
    assert (list(dclaims[0]), list(dclaims[1])) == (list(oclaims[0]), list(oclaims[1]))
    proof = tr.proof()
    assert proof == oproof
    # the oracle verifier accepts the device-made proof and the proved result is the MSM
    tv = ProofTranscript2.start_verifier(b"fgstglsp", proof)
    expected = CV.te_msm(points, coefs)
    assert PP.verify_pippenger(tv, cfg, odense, oclaims, okey, expected) == expected
    # the pair returned by the prover satisfies the (mock-setup) pairing equation A == tau * B
    a, b = res_to_point(dpair[0]), res_to_point(dpair[1])
    okey.kzg.verify_pair((a, b))
    # the C++ host orchestration (gkr_run_pippenger, csrc/protocol.cu) produces the same bytes
    tr2 = g.Transcript(b"fgstglsp")
    ndense, nevs, npair = g.run_pippenger_native(ctx, tr2, kzg.srs, kzg.g0, key.dev, points_xy, coefs_to_u64(coefs), d, x, nbits, clm, to_limbs(r))
    assert tr2.proof() == oproof
    assert [from_limbs(t) for t in ndense] == odense
    assert from_limbs(nevs) == list(oclaims[1])
    assert np.array_equal(npair[0], dpair[0]) and np.array_equal(npair[1], dpair[1])


@pytest.mark.parametrize("d,x,nbits,clm", [(6, 12, 128, 0), (5, 10, 64, 2), (8, 16, 128, 0), (10, 20, 128, 0), (8, 14, 253, 2)])
def test_pippenger_full_size_properties(ctx, d, x, nbits, clm):
    """Sizes the python prover cannot reach (BASELINE config[0], x = 16, and the 2^20-point shape of config[2], x = 20): size-independent properties instead of a
    byte comparison -- (1) the ORACLE VERIFIER accepts the device-made proof (every sumcheck round, every claim reduction,
    the opening equation and the pairing check A == tau B of the mock setup), (2) the proved MSM result, recovered by the
    verifier from the output tables, equals the true MSM, known in closed form because the synthetic points are the
    arithmetic progression (k0 + i step) G:  sum_i c_i P_i = (sum_i c_i (k0 + i step)) G."""
    rng = np.random.default_rng(1000 * d + x)
    cfg = DPP.pippenger_config(d, x, nbits, clm)
    n = 1 << x
    k0, step = 0x1234567 + x, 0x9E3779B97F4A7C15
    pts = H.te_points_arithmetic_progression(k0, step, n)
    points_xy = np.stack([to_limbs([p[0] for p in pts]), to_limbs([p[1] for p in pts])])
    raw = np.frombuffer(rng.bytes(32 * n), dtype=np.uint8).reshape(n, 32).copy()
    raw[:, nbits // 8:] = 0
    coefs_u64 = raw.view(np.uint64).reshape(n, 4)
    coefs = [int.from_bytes(raw[i].tobytes(), "little") for i in range(n)]
    r = [int.from_bytes(rng.bytes(32), "little") % P for _ in range(cfg["y_logsize"])]
    tau = int.from_bytes(rng.bytes(32), "little") % P
    nv = x + clm
    key = DPP.KnucklesKey(ctx, DPP.KzgKey.mock_setup(ctx, tau, CV.G1_GEN, 2 * (1 << nv) - 1), nv, 2)
    tr = g.Transcript(b"fgstglsp")
    ddense, dclaims, dpair = DPP.run_pippenger(ctx, tr, points_xy, coefs_u64, cfg, r, key)
    proof = tr.proof()
    tr2 = g.Transcript(b"fgstglsp")  # C++ host orchestration: same proof, outputs and pairing pair
    ndense, nevs, npair = g.run_pippenger_native(ctx, tr2, key.kzg.srs, key.kzg.g0, key.dev, points_xy, coefs_u64, d, x, nbits, clm, to_limbs(r))
    assert tr2.proof() == proof
    # BYTE PARITY at this size: the proof, the output tables, the claims and the deferred pairing pair are exactly what the
    # independent CPU prover produced from the same seeded inputs (tests/golden/pippenger_large.json)
    gold = LARGE_GOLDEN[f"pippenger_d{d}_x{x}_n{nbits}_c{clm}"]
    assert len(proof) == gold["proof_len"]
    assert hashlib.sha256(proof).hexdigest() == gold["proof_sha256"], "device proof differs from the independent CPU prover's"
    assert hashlib.sha256(np.ascontiguousarray(ndense).tobytes()).hexdigest() == gold["dense_output_sha256"]
    assert hashlib.sha256(np.ascontiguousarray(nevs).tobytes()).hexdigest() == gold["claim_evs_sha256"]
    assert hashlib.sha256(np.ascontiguousarray(npair).tobytes()).hexdigest() == gold["pair_sha256"]
    assert all(np.array_equal(a, t.download()) for a, t in zip(ndense, ddense))
    assert from_limbs(nevs) == list(dclaims[1])
    assert np.array_equal(npair[0], dpair[0]) and np.array_equal(npair[1], dpair[1])
    okey = PP.KnucklesKey(PP.KzgKey(tau, CV.G1_GEN, 2 * (1 << nv) - 1), nv, 2)
    dense_output = [from_limbs(t.download()) for t in ddense]
    total = sum(c * (k0 + i * step) for i, c in enumerate(coefs)) % CV.TE_SUBGROUP_ORDER
    expected = CV.te_to_affine(H.te_mul(total, CV.TE_GEN))
    tv = ProofTranscript2.start_verifier(b"fgstglsp", proof)
    got = PP.verify_pippenger(tv, cfg, dense_output, (list(dclaims[0]), list(dclaims[1])), okey, expected)
    assert tv.ctr == len(proof) and got == expected
    okey.kzg.verify_pair((res_to_point(dpair[0]), res_to_point(dpair[1])))
    # a corrupted proof is rejected
    bad = bytearray(proof)
    bad[len(bad) // 3] ^= 4
    with pytest.raises(AssertionError):
        PP.verify_pippenger(ProofTranscript2.start_verifier(b"fgstglsp", bytes(bad)), cfg, dense_output,
                            (list(dclaims[0]), list(dclaims[1])), okey, expected)


@pytest.mark.parametrize("kind", ["zeros", "ones", "same", "one-hot"])
@pytest.mark.parametrize("d,x,nbits,clm", [(2, 3, 6, 0), (3, 5, 16, 1)])
def test_pippenger_degenerate_scalars(ctx, kind, d, x, nbits, clm):
    """Collisions and empty buckets: all scalars zero / all-ones / identical / a single non-zero one -- every point of a digit
    row lands in ONE bucket (counters run up to 2^x - 1, all other bucket rows are empty).  Device proof == oracle proof, and
    the oracle verifier recovers the MSM (the identity for the zero scalars)."""
    rng = random.Random(7000 + 10 * d + x)
    cfg = PP.pippenger_config(d, x, nbits, clm)
    n = 1 << x
    points = [CV.te_random_point(rng) for _ in range(n)]
    coefs = {"zeros": [0] * n, "ones": [(1 << nbits) - 1] * n, "same": [rng.randrange(1 << nbits)] * n, "one-hot": [0] * (n - 1) + [5]}[kind]
    r = [rng.randrange(P) for _ in range(cfg["y_logsize"])]
    nv = x + clm
    tau = rng.randrange(1, P)
    g0 = CV.g1_mul(rng.randrange(1, P), CV.G1_GEN)
    okey = PP.KnucklesKey(PP.KzgKey(tau, g0, 2 * (1 << nv) - 1), nv, 2)
    tp = ProofTranscript2.start_prover(b"fgstglsp")
    odense, oclaims = PP.run_pippenger(tp, points, coefs, cfg, r, okey)
    oproof = tp.end()
    kzg = DPP.KzgKey.mock_setup(ctx, tau, g0, 2 * (1 << nv) - 1)
    key = DPP.KnucklesKey(ctx, kzg, nv, 2)
    points_xy = np.stack([to_limbs([p[0] for p in points]), to_limbs([p[1] for p in points])])
    tr = g.Transcript(b"fgstglsp")
    ndense, nevs, npair = g.run_pippenger_native(ctx, tr, kzg.srs, kzg.g0, key.dev, points_xy, coefs_to_u64(coefs), d, x, nbits, clm, to_limbs(r))
    assert tr.proof() == oproof
    assert [from_limbs(t) for t in ndense] == odense
    expected = CV.te_msm(points, coefs)
    tv = ProofTranscript2.start_verifier(b"fgstglsp", tr.proof())
    assert PP.verify_pippenger(tv, cfg, odense, oclaims, okey, expected) == expected


def test_pippenger_single_digit_row_is_rejected_like_the_reference(ctx):
    """nbits <= d_logsize gives y_size = 1, y_logsize = 0: the reference panics in LogupMainphase::new ("logsizes must be
    non-increasing", logup_mainphase.rs) -- the oracle raises, the device entry returns an error status instead of a proof"""
    rng = random.Random(3)
    d, x, nbits, clm = 3, 4, 3, 0
    with pytest.raises(AssertionError):
        cfg, points, coefs, r, okey = make_instance(random.Random(3), d, x, nbits, clm)
        PP.run_pippenger(ProofTranscript2.start_prover(b"fgstglsp"), points, coefs, cfg, r, okey)
    cfg, points, coefs, r, okey = make_instance(rng, d, x, nbits, clm)
    kzg = DPP.KzgKey.mock_setup(ctx, okey.kzg.tau, okey.kzg.g0, 2 * (1 << x) - 1)
    key = DPP.KnucklesKey(ctx, kzg, x, 2)
    points_xy = np.stack([to_limbs([p[0] for p in points]), to_limbs([p[1] for p in points])])
    with pytest.raises(g.GkrError):
        g.run_pippenger_native(ctx, g.Transcript(b"fgstglsp"), kzg.srs, kzg.g0, key.dev, points_xy, coefs_to_u64(coefs), d, x, nbits, clm, to_limbs(r))


@pytest.mark.parametrize("d,x,nbits,clm", [(8, 10, 253, 2), (10, 12, 128, 0), (4, 9, 64, 1), (7, 8, 252, 0), (6, 11, 100, 3), (9, 9, 31, 1)])
def test_pippenger_device_vs_cpp_oracle_live(ctx, d, x, nbits, clm):
    """Sizes and shapes the python prover cannot reach, checked BYTE FOR BYTE against the independent C++ prover run live
    (oracle/c/pippenger_oracle.cpp, pinned to oracle/pyref in tests/test_pippenger_oracle.py): full-width scalars with
    y_size = 32 and clm = 2 (the shape of BASELINE config[3]), d = 10 (config[2]), x == d, y_size not a power of two, clm = 3,
    byte-truncated odd bit widths (pippenger.rs:464-466)."""
    from oracle import pippenger_oracle as PO
    from oracle.pyref.field import fq_vec_to_mont_u64
    rng = np.random.default_rng(5000 + 1000 * d + 10 * x + clm)
    n = 1 << x
    cfg = DPP.pippenger_config(d, x, nbits, clm)
    k0, step = 0xABCDEF + x, 0x9E3779B97F4A7C15
    points_xy = PO.te_arithmetic_progression(k0, step, n)
    raw = np.frombuffer(rng.bytes(32 * n), dtype=np.uint8).reshape(n, 32).copy()
    raw[:, nbits // 8:] = 0
    coefs_u64 = raw.view(np.uint64).reshape(n, 4)
    r = [int.from_bytes(rng.bytes(32), "little") % P for _ in range(cfg["y_logsize"])]
    tau = int.from_bytes(rng.bytes(32), "little") % P
    nv = x + clm
    okey = PO.Key(to_limbs([tau])[0], fq_vec_to_mont_u64([CV.G1_GEN[0], CV.G1_GEN[1]]).reshape(12), nv, to_limbs([2])[0])
    want = PO.run_pippenger(okey, points_xy, coefs_u64, to_limbs(r), d, x, nbits, clm)
    key = DPP.KnucklesKey(ctx, DPP.KzgKey.mock_setup(ctx, tau, CV.G1_GEN, 2 * (1 << nv) - 1), nv, 2)
    # the two keys hold the same SRS points
    got_srs = key.kzg.srs.download_affine()
    for i in (0, 1, n, 2 * (1 << nv) - 2):
        assert np.array_equal(np.asarray(got_srs[i]).reshape(12), okey.point(i))
    okey.close()
    tr = g.Transcript(b"fgstglsp")
    ndense, nevs, npair = g.run_pippenger_native(ctx, tr, key.kzg.srs, key.kzg.g0, key.dev, points_xy, coefs_u64, d, x, nbits, clm, to_limbs(r))
    assert tr.proof() == want["proof"], "device proof differs from the independent CPU prover's"
    assert np.array_equal(ndense, want["dense_output"])
    assert np.array_equal(nevs, want["claim_evs"])
    assert np.array_equal(npair, want["pair"])


@pytest.mark.parametrize("d,x,nbits,clm", [(6, 12, 128, 0), (5, 10, 64, 2), (8, 14, 253, 2)])
def test_pippenger_witness_recompute_plan_same_proof(ctx, d, x, nbits, clm, monkeypatch):
    """memory plan of large instances (protocol.cu, bintree_witness): only the input of every addition layer stays resident,
    the L1 / L2 images are recomputed when the prover reaches the layer -- same tables, so the proof, the outputs and the
    pairing pair are the golden ones of the independent CPU prover"""
    monkeypatch.setenv("GKR_WITNESS_RECOMPUTE", "1")
    rng = np.random.default_rng(1000 * d + x)
    n = 1 << x
    k0, step = 0x1234567 + x, 0x9E3779B97F4A7C15
    pts = H.te_points_arithmetic_progression(k0, step, n)
    points_xy = np.stack([to_limbs([p[0] for p in pts]), to_limbs([p[1] for p in pts])])
    raw = np.frombuffer(rng.bytes(32 * n), dtype=np.uint8).reshape(n, 32).copy()
    raw[:, nbits // 8:] = 0
    coefs_u64 = raw.view(np.uint64).reshape(n, 4)
    cfg = DPP.pippenger_config(d, x, nbits, clm)
    r = [int.from_bytes(rng.bytes(32), "little") % P for _ in range(cfg["y_logsize"])]
    tau = int.from_bytes(rng.bytes(32), "little") % P
    nv = x + clm
    key = DPP.KnucklesKey(ctx, DPP.KzgKey.mock_setup(ctx, tau, CV.G1_GEN, 2 * (1 << nv) - 1), nv, 2)
    tr = g.Transcript(b"fgstglsp")
    ndense, nevs, npair = g.run_pippenger_native(ctx, tr, key.kzg.srs, key.kzg.g0, key.dev, points_xy, coefs_u64, d, x, nbits, clm, to_limbs(r))
    gold = LARGE_GOLDEN[f"pippenger_d{d}_x{x}_n{nbits}_c{clm}"]
    assert hashlib.sha256(tr.proof()).hexdigest() == gold["proof_sha256"]
    assert hashlib.sha256(np.ascontiguousarray(ndense).tobytes()).hexdigest() == gold["dense_output_sha256"]
    assert hashlib.sha256(np.ascontiguousarray(nevs).tobytes()).hexdigest() == gold["claim_evs_sha256"]
    assert hashlib.sha256(np.ascontiguousarray(npair).tobytes()).hexdigest() == gold["pair_sha256"]


@pytest.mark.skipif(not os.environ.get("GKR_TEST_LARGE_X"), reason="minutes of host-side input generation: set GKR_TEST_LARGE_X=<x_logsize> (profiles/ keeps the log)")
def test_pippenger_config3_shape_largest_single_gpu(ctx):
    """BASELINE config[3] shape (full-width 253-bit scalars, d = 8, commitment-log-multiplicity 2) at the largest x one B200 holds:
    the oracle verifier accepts the device proof and recovers the true MSM (closed form over the arithmetic-progression points)"""
    x, d, nbits, clm = int(os.environ["GKR_TEST_LARGE_X"]), 8, 253, 2
    rng = np.random.default_rng(1000 * d + x)
    cfg = DPP.pippenger_config(d, x, nbits, clm)
    n = 1 << x
    k0, step = 0x1234567 + x, 0x9E3779B97F4A7C15
    pts = H.te_points_arithmetic_progression(k0, step, n)
    points_xy = np.stack([to_limbs([p[0] for p in pts]), to_limbs([p[1] for p in pts])])
    raw = np.frombuffer(rng.bytes(32 * n), dtype=np.uint8).reshape(n, 32).copy()
    raw[:, nbits // 8:] = 0
    coefs_u64 = raw.view(np.uint64).reshape(n, 4)
    r = [int.from_bytes(rng.bytes(32), "little") % P for _ in range(cfg["y_logsize"])]
    tau = int.from_bytes(rng.bytes(32), "little") % P
    nv = x + clm
    key = DPP.KnucklesKey(ctx, DPP.KzgKey.mock_setup(ctx, tau, CV.G1_GEN, 2 * (1 << nv) - 1), nv, 2)
    tr = g.Transcript(b"fgstglsp")
    import time
    t0 = time.perf_counter()
    ndense, nevs, npair = g.run_pippenger_native(ctx, tr, key.kzg.srs, key.kzg.g0, key.dev, points_xy, coefs_u64, d, x, nbits, clm, to_limbs(r))
    print(f"x = {x}: run_pippenger {time.perf_counter() - t0:.2f} s (first call), proof {len(tr.proof())} bytes")
    proof = tr.proof()
    okey = PP.KnucklesKey(PP.KzgKey(tau, CV.G1_GEN, 2 * (1 << nv) - 1), nv, 2)
    dense_output = [from_limbs(t) for t in ndense]
    words = coefs_u64.astype(object)
    total = 0
    for i in range(n):
        c = int(words[i, 0]) | (int(words[i, 1]) << 64) | (int(words[i, 2]) << 128) | (int(words[i, 3]) << 192)
        total += c * (k0 + i * step)
    expected = CV.te_to_affine(H.te_mul(total % CV.TE_SUBGROUP_ORDER, CV.TE_GEN))
    tv = ProofTranscript2.start_verifier(b"fgstglsp", proof)
    got = PP.verify_pippenger(tv, cfg, dense_output, (list(r), from_limbs(nevs)), okey, expected)
    assert tv.ctr == len(proof) and got == expected
    okey.kzg.verify_pair((res_to_point(npair[0]), res_to_point(npair[1])))
    bad = bytearray(proof)
    bad[len(bad) // 2] ^= 1
    with pytest.raises(AssertionError):
        PP.verify_pippenger(ProofTranscript2.start_verifier(b"fgstglsp", bytes(bad)), cfg, dense_output, (list(r), from_limbs(nevs)), okey, expected)


@pytest.mark.parametrize("d,x,nbits,clm", [(8, 18, 253, 2), (10, 22, 128, 0)])
def test_pippenger_large_golden_digests(ctx, d, x, nbits, clm):
    """byte parity beyond the BASELINE sizes: the config[3] shape at x = 18 (2^23 point-digit incidences, 253-bit scalars, clm 2)
    and 2^22 points at 128 bit -- proof, output tables, claims and pairing pair of the device prover hash to the digests the
    independent CPU prover minted (tests/golden/make_golden_large.py; minutes of CPU time each).  The 2^22-point case needs
    ~45 s of host-side input generation and runs under GKR_TEST_LARGE_GOLDEN=1 only."""
    name = f"pippenger_d{d}_x{x}_n{nbits}_c{clm}"
    if name not in LARGE_GOLDEN:
        pytest.skip("digest not minted")
    if x >= 20 and not os.environ.get("GKR_TEST_LARGE_GOLDEN"):
        pytest.skip("set GKR_TEST_LARGE_GOLDEN=1")
    rng = np.random.default_rng(1000 * d + x)
    n = 1 << x
    k0, step = 0x1234567 + x, 0x9E3779B97F4A7C15
    pts = H.te_points_arithmetic_progression(k0, step, n)
    points_xy = np.stack([to_limbs([p[0] for p in pts]), to_limbs([p[1] for p in pts])])
    raw = np.frombuffer(rng.bytes(32 * n), dtype=np.uint8).reshape(n, 32).copy()
    raw[:, nbits // 8:] = 0
    coefs_u64 = raw.view(np.uint64).reshape(n, 4)
    cfg = DPP.pippenger_config(d, x, nbits, clm)
    r = [int.from_bytes(rng.bytes(32), "little") % P for _ in range(cfg["y_logsize"])]
    tau = int.from_bytes(rng.bytes(32), "little") % P
    nv = x + clm
    key = DPP.KnucklesKey(ctx, DPP.KzgKey.mock_setup(ctx, tau, CV.G1_GEN, 2 * (1 << nv) - 1), nv, 2)
    tr = g.Transcript(b"fgstglsp")
    ndense, nevs, npair = g.run_pippenger_native(ctx, tr, key.kzg.srs, key.kzg.g0, key.dev, points_xy, coefs_u64, d, x, nbits, clm, to_limbs(r))
    gold = LARGE_GOLDEN[name]
    assert len(tr.proof()) == gold["proof_len"]
    assert hashlib.sha256(tr.proof()).hexdigest() == gold["proof_sha256"]
    assert hashlib.sha256(np.ascontiguousarray(ndense).tobytes()).hexdigest() == gold["dense_output_sha256"]
    assert hashlib.sha256(np.ascontiguousarray(nevs).tobytes()).hexdigest() == gold["claim_evs_sha256"]
    assert hashlib.sha256(np.ascontiguousarray(npair).tobytes()).hexdigest() == gold["pair_sha256"]
