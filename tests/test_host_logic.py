"""CPU tests of the host side of the product: the C-ABI library loads and exports every symbol the
header declares, the C++ merlin transcript matches the oracle / published vector, and a context
refuses to be created without a GPU (no CPU fallback)."""
import os
import random
import re

import numpy as np
import pytest

import gkr_msm_b200 as g
from oracle.pyref.field import P
from oracle.pyref.transcript import ProofTranscript2
from tests.util import from_limbs, to_limbs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = g.load_library()
    hdr = open(os.path.join(ROOT, "include", "gkr_msm_b200.h")).read()
    names = set(re.findall(r"\b(gkr_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    for n in sorted(names):
        assert hasattr(lib, n), f"symbol {n} declared in include/gkr_msm_b200.h is not exported"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(g.GkrError) as e:
        g.Context(0)
    assert e.value.code == g.GKR_ERR_CUDA


def test_cpp_transcript_matches_oracle_and_merlin_vector():
    rng = random.Random(1)
    t = g.Transcript(b"fgstglsp")
    o = ProofTranscript2.start_prover(b"fgstglsp")
    for step in range(6):
        xs = [rng.randrange(P) for _ in range(1 + step % 3)]
        t.write_scalars(to_limbs(xs))
        o.write_scalars(xs)
        bits = (128, 512, 8, 255)[step % 4]
        assert from_limbs(t.challenge(bits).reshape(1, 4))[0] == o.challenge(bits)
    t.write_raw(b"\x01\x02\x03" * 100)  # crosses the 166-byte STROBE rate
    o.write_raw_msg(b"\x01\x02\x03" * 100)
    assert t.raw_challenge(200) == o.raw_challenge(200)
    assert t.proof() == o.end()


def test_cpp_transcript_rejects_non_canonical_scalar():
    t = g.Transcript(b"x")
    bad = np.array([[0xFFFFFFFFFFFFFFFF] * 4], dtype=np.uint64)
    with pytest.raises(g.GkrError):
        t.write_scalars(bad)


def test_host_g1_horner_matches_group_law():
    """the CPU tail of every MSM (csrc/host_g1.hpp): Horner over extended-Jacobian window sums + normalisation"""
    import ctypes as C

    from oracle.pyref import curves as CV
    from oracle.pyref.field import FQ_MODULUS as Q
    from oracle.pyref.field import fq_vec_from_mont_u64, fq_vec_to_mont_u64

    lib = g.load_library()
    lib.gkr_host_g1_horner.restype = C.c_int
    lib.gkr_host_g1_horner.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    rng = random.Random(7)
    for c, n_windows in [(4, 1), (5, 3), (13, 20), (16, 16), (0, 1)]:
        pts = [CV.g1_mul(rng.randrange(1, P), CV.G1_GEN) for _ in range(n_windows)]
        if n_windows > 2:
            pts[1] = None  # an empty window
            pts[-1] = pts[0]
        ws = np.zeros((n_windows, 24), np.uint64)
        for i, pt in enumerate(pts):
            if pt is None:
                continue
            z = rng.randrange(1, Q)
            zz, zzz = z * z % Q, z * z * z % Q
            ws[i] = fq_vec_to_mont_u64([pt[0] * zz % Q, pt[1] * zzz % Q, zz, zzz]).reshape(24)
        out = np.zeros(12, np.uint64)
        assert lib.gkr_host_g1_horner(ws.ctypes.data, c, n_windows, out.ctypes.data) == 0
        want = None
        for i, pt in enumerate(pts):
            want = CV.g1_add(want, CV.g1_mul(1 << (c * i), pt) if pt is not None else None)
        x, y = fq_vec_from_mont_u64(out.reshape(2, 6))
        assert (None if (x, y) == (0, 0) else (x, y)) == want
    # P + (-P) and doubling inside the Horner chain
    pt = CV.g1_mul(5, CV.G1_GEN)
    neg = (pt[0], Q - pt[1])
    ws = np.zeros((2, 24), np.uint64)
    ws[0] = fq_vec_to_mont_u64([neg[0], neg[1], 1, 1]).reshape(24)
    ws[1] = fq_vec_to_mont_u64([CV.g1_mul(5 * pow(2, -1, P) % P, CV.G1_GEN)[0], CV.g1_mul(5 * pow(2, -1, P) % P, CV.G1_GEN)[1], 1, 1]).reshape(24)
    out = np.ones(12, np.uint64)
    assert lib.gkr_host_g1_horner(ws.ctypes.data, 1, 2, out.ctypes.data) == 0
    assert not out.any()  # 2 * (5/2 G) - 5 G = infinity


def test_pushforward_bucketize_matches_reference_bookkeeping():
    """digits / counters / buckets of PushForwardState::new (pushforward.rs:351-396) from the host library"""
    rng = random.Random(11)
    for d, y_size, nbits, n in [(3, 5, 15, 64), (8, 16, 128, 500), (10, 13, 128, 300), (7, 36, 252, 100), (8, 32, 253, 128)]:
        coefs = [rng.randrange(1 << nbits) for _ in range(n)]
        co = np.array([[(c >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)] for c in coefs], dtype=np.uint64)
        digits, counter, order, lens = g.pushforward_bucketize(co, y_size, d)
        for y in range(y_size):
            want = [(c >> (y * d)) & ((1 << d) - 1) for c in coefs]
            assert [int(v) for v in digits[y]] == want
            buckets = [[] for _ in range(1 << d)]
            cnt = []
            for x, dg in enumerate(want):
                cnt.append(len(buckets[dg]))
                buckets[dg].append(x)
            assert [int(v) for v in counter[y]] == cnt
            assert [int(v) for v in order[y]] == [x for b in buckets for x in b]
            assert [int(v) for v in lens[y]] == [len(b) for b in buckets]
    with pytest.raises(g.GkrError):
        g.pushforward_bucketize(np.zeros((4, 4), np.uint64), 40, 8)


def test_g1_sum_combines_partial_commitments():
    """host combine step of the MSM split by point range: sum of G affine partial results"""
    from gkr_msm_b200 import hostmath as H
    from oracle.pyref import curves as CV

    rng = random.Random(13)
    pts = [CV.g1_mul(rng.randrange(1, P), CV.G1_GEN) for _ in range(5)] + [None]
    pts.append((pts[0][0], CV.Q - pts[0][1]))  # cancels the first one
    got = g.g1_sum(np.stack([H.g1_to_limbs(p) for p in pts]))
    want = None
    for p in pts:
        want = CV.g1_add(want, p)
    assert H.g1_from_limbs(got) == want
    assert not g.g1_sum(np.zeros((3, 12), np.uint64)).any()


def test_cpp_g1_wire_format_matches_published_generator_encoding():
    """csrc/host_g1.hpp::serialize_compressed (what gkr_run_pippenger writes into the proof) against the published compressed
    encoding of the BLS12-381 G1 generator, the infinity encoding, and the oracle on random points"""
    import ctypes as C

    from gkr_msm_b200 import hostmath as H
    from oracle.pyref import curves as CV
    from oracle.pyref import pippenger as PP

    lib = g.load_library()
    lib.gkr_host_g1_serialize.restype = C.c_int
    lib.gkr_host_g1_serialize.argtypes = [C.c_void_p, C.c_void_p]

    def ser(pt):
        xy = np.ascontiguousarray(H.g1_to_limbs(pt), dtype=np.uint64).reshape(12)
        out = np.zeros(48, np.uint8)
        assert lib.gkr_host_g1_serialize(xy.ctypes.data, out.ctypes.data) == 0
        return out.tobytes()

    assert ser(CV.G1_GEN).hex() == "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"
    assert ser(None) == bytes([0xC0]) + bytes(47)
    rng = random.Random(12)
    for _ in range(8):
        pt = CV.g1_mul(rng.randrange(1, P), CV.G1_GEN)
        assert ser(pt) == PP.g1_serialize(pt) == H.g1_serialize(pt)
        assert ser(CV.g1_neg(pt)) == PP.g1_serialize(CV.g1_neg(pt))
