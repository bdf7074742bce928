"""The reduced-radix (carry-free IMAD.WIDE) field code of gkr-msm_b200/csrc/rr_field.cuh / dense29_item.cuh, compiled for the
CPU from the SAME source the device kernels use (tests/native/rr_host.cpp), against python big integers: radix conversion,
Montgomery products at the extremes of the documented limb / value bounds, the short fold, and the whole per-item
arithmetic of the Prod3 round kernels (DenseSumcheckObjectSO::unipoly + bind_dense_poly, sumcheck.rs:160-163, 277-332)."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
P = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
M29 = (1 << 29) - 1
R261 = 1 << 261


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(HERE, "native", "librr_host.so")
    src = os.path.join(HERE, "native", "rr_host.cpp")
    deps = [src] + [os.path.join(HERE, "..", "gkr-msm_b200", "csrc", "lab", f) for f in ("rr_field.cuh", "dense29_item.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-Wno-unknown-pragmas", "-o", so, src])
    return C.CDLL(so)


def words(x, n=8):
    return np.array([(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)], dtype=np.uint32)


def from_words(w):
    return sum(int(v) << (32 * i) for i, v in enumerate(w))


def limbs(x):
    out = [(x >> (29 * i)) & M29 for i in range(8)] + [x >> 232]
    assert out[8] < (1 << 32)
    return np.array(out, dtype=np.uint32)


def from_limbs(l):
    return sum(int(v) << (29 * i) for i, v in enumerate(l))


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def test_radix_conversion_round_trip(lib):
    rng = random.Random(1)
    for x in [0, 1, P - 1, P, (1 << 256) - 1, 1 << 232, (1 << 232) - 1] + [rng.getrandbits(256) for _ in range(200)]:
        l = np.zeros(9, dtype=np.uint32)
        lib.rr_t_load(ptr(words(x)), ptr(l))
        assert from_limbs(l) == x and all(int(v) <= M29 for v in l[:8])
        w = np.zeros(8, dtype=np.uint32)
        lib.rr_t_store(ptr(l), ptr(w))
        assert from_words(w) == x
    for x in [0, P - 1, P, P + 1, 2 * P - 1, 2 * P, 2 * P + 5, (1 << 256) - 1]:
        w = words(x)
        lib.rr_t_canonical(ptr(w), 3)
        assert from_words(w) == x % P or x >= 4 * P


def loose(rng, bound_limb, top_bound):
    """a limb vector with limbs up to bound_limb (exclusive) -- the value is whatever it sums to"""
    l = [rng.randrange(bound_limb) for _ in range(8)] + [rng.randrange(top_bound)]
    return np.array(l, dtype=np.uint32)


def test_montgomery_product_bounds_and_value(lib):
    rng = random.Random(2)
    cases = []
    for _ in range(300):
        cases.append((limbs(rng.getrandbits(256)), limbs(rng.getrandbits(256))))
    # extremes of the documented discipline: one operand with limbs < 1.5 * 2^30 against a tight one, both < 2^30,
    # all-ones limbs, values up to the 261-bit container
    ones30 = np.array([(1 << 30) - 1] * 8 + [(1 << 26) - 1], dtype=np.uint32)
    tight_max = np.array([M29] * 9, dtype=np.uint32)
    loose15 = np.array([3 * (1 << 29) - 1] * 8 + [(1 << 27) - 1], dtype=np.uint32)
    cases += [(ones30, ones30), (loose15, tight_max), (tight_max, loose15), (tight_max, tight_max)]
    for _ in range(200):
        cases.append((loose(rng, 1 << 30, 1 << 26), loose(rng, 1 << 30, 1 << 26)))
        cases.append((loose(rng, 3 << 29, 1 << 27), loose(rng, 1 << 29, 1 << 29)))
    for a, b in cases:
        out = np.zeros(9, dtype=np.uint32)
        lib.rr_t_mul(ptr(a), ptr(b), ptr(out))
        va, vb, vr = from_limbs(a), from_limbs(b), from_limbs(out)
        assert all(int(v) <= M29 for v in out[:8])
        assert (vr * R261 - va * vb) % P == 0
        assert vr < va * vb // R261 + P + 1
    for _ in range(100):
        a = loose(rng, 1 << 30, 1 << 26)
        out = np.zeros(9, dtype=np.uint32)
        lib.rr_t_sqr(ptr(a), ptr(out))
        va, vr = from_limbs(a), from_limbs(out)
        assert (vr * R261 - va * va) % P == 0 and vr < va * va // R261 + P + 1


def test_sub_and_fold(lib):
    rng = random.Random(3)
    for _ in range(300):
        a, b = rng.getrandbits(256), rng.getrandbits(256)
        out = np.zeros(9, dtype=np.uint32)
        lib.rr_t_sub_norm(ptr(limbs(a)), ptr(limbs(b)), ptr(out))
        assert from_limbs(out) == a - b + 4 * P and all(int(v) <= M29 for v in out[:8])
    ts = [0, 1, (1 << 128) - 1] + [rng.getrandbits(128) for _ in range(200)]
    for t in ts:
        e0, e1 = rng.randrange(P), rng.randrange(P)
        if t == (1 << 128) - 1:
            e0, e1 = 0, P - 1  # largest difference
        out = np.zeros(9, dtype=np.uint32)
        lib.rr_t_fold(ptr(words(e0)), ptr(words(e1)), ptr(words(t, 4)), ptr(out))
        v = from_limbs(out)
        assert all(int(x) <= M29 for x in out[:8])
        assert (v * (1 << 145) - (e0 + t * (e1 - e0))) % P == 0
        assert v < 2 * P


@pytest.mark.parametrize("n", [1, 3, 4, 37, 256])
def test_prod3_round_items_match_big_int(lib, n):
    """sums at the nodes 1..3 of Prod3 over n pairs (times 2^-10 in Montgomery-261 terms), and the fused fold + next round"""
    rng = random.Random(100 + n)
    tabs = [[rng.randrange(P) for _ in range(2 * n)] for _ in range(3)]
    if n >= 3:  # extremes
        tabs[0][0], tabs[0][1] = 0, P - 1
        tabs[1][2], tabs[1][3] = P - 1, 0
        tabs[2][4], tabs[2][5] = P - 1, P - 1
    flat = np.concatenate([words(v) for t in tabs for v in t])
    sums = np.zeros(24, dtype=np.uint32)
    lib.rr_t_prod3_eval(ptr(flat), C.c_uint64(n), ptr(sums))
    inv = pow(R261, -1, P)
    for s in range(3):
        want = 0
        for i in range(n):
            term = 1
            for j in range(3):
                lo, hi = tabs[j][2 * i], tabs[j][2 * i + 1]
                term = term * (hi + s * (hi - lo)) % P
            want += term
        want = want * inv * inv % P
        got = from_words(sums[8 * s:8 * s + 8])
        assert got == want and got < P

    # fused: fold quads by a 128-bit challenge, then the same sums over the folded pairs
    t = rng.getrandbits(128)
    q = [[rng.randrange(P) for _ in range(4 * n)] for _ in range(3)]
    flat = np.concatenate([words(v) for tb in q for v in tb])
    folded = np.zeros(3 * 2 * n * 8, dtype=np.uint32)
    lib.rr_t_prod3_fold_eval(ptr(flat), C.c_uint64(n), ptr(words(t, 4)), ptr(folded), ptr(sums))
    i145 = pow(1 << 145, -1, P)
    ft = [[(tb[2 * k] + t * (tb[2 * k + 1] - tb[2 * k])) * i145 % P for k in range(2 * n)] for tb in q]
    for j in range(3):
        for k in range(2 * n):
            off = (j * 2 * n + k) * 8
            assert from_words(folded[off:off + 8]) == ft[j][k]
    for s in range(3):
        want = 0
        for i in range(n):
            term = 1
            for j in range(3):
                lo, hi = ft[j][2 * i], ft[j][2 * i + 1]
                term = term * (hi + s * (hi - lo)) % P
            want += term
        want = want * inv * inv % P
        assert from_words(sums[8 * s:8 * s + 8]) == want
