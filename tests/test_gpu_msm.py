"""GPU parity for the commitment MSM over BLS12-381 G1 (KzgProvingKey::commit, kzg.rs:123-126; msm_nonaff,
msm_nonaffine.rs:34-38) against an independent textbook group-law oracle -- the way the reference's own tests
compare against `G::msm` (pullback.rs:86-107, binary_msm.rs:63-96).  The result is a unique group element, so
affine coordinates must match bit for bit."""
import random

import numpy as np
import pytest

import gkr_msm_b200 as g
from oracle.pyref import curves as CV
from oracle.pyref.field import FQ_MODULUS, P, fq_vec_from_mont_u64, fq_vec_to_mont_u64
from tests.util import to_limbs

pytestmark = pytest.mark.gpu


def aff_to_limbs(points):
    out = np.zeros((len(points), 12), np.uint64)
    for i, p in enumerate(points):
        if p is not None:
            out[i] = fq_vec_to_mont_u64([p[0], p[1]]).reshape(12)
    return out


def res_to_point(xy):
    x, y = fq_vec_from_mont_u64(xy.reshape(2, 6))
    return None if (x, y) == (0, 0) else (x, y)


def rand_g1(rng):
    return CV.g1_mul(rng.randrange(1, CV.G1_ORDER), CV.G1_GEN)


@pytest.mark.parametrize("n", [1, 2, 17, 300])
def test_msm_random(ctx, n):
    rng = random.Random(n)
    pts = [rand_g1(rng) for _ in range(min(n, 24))]
    pts = [pts[i % len(pts)] for i in range(n)]  # repeated bases exercise the doubling path inside buckets
    sc = [rng.randrange(P) for _ in range(n)]
    srs = g.Srs(ctx, aff_to_limbs(pts))
    got = res_to_point(srs.msm(ctx.upload(to_limbs(sc))))
    assert got == CV.g1_msm(pts, sc)
    assert CV.g1_on_curve(got)


def test_msm_edge_cases(ctx):
    rng = random.Random(5)
    a, b = rand_g1(rng), rand_g1(rng)
    # zeros, ones, r-1, infinity bases, P + (-P), equal scalars on equal points
    pts = [a, CV.g1_neg(a), b, None, b, a, b]
    sc = [7, 7, 0, 12345, 1, P - 1, 1 << 200]
    srs = g.Srs(ctx, aff_to_limbs(pts))
    assert res_to_point(srs.msm(ctx.upload(to_limbs(sc)))) == CV.g1_msm(pts, sc)
    # everything cancels -> infinity
    pts2, sc2 = [a, CV.g1_neg(a)], [99, 99]
    assert res_to_point(g.Srs(ctx, aff_to_limbs(pts2)).msm(ctx.upload(to_limbs(sc2)))) is None
    # small scalars (the c/d counter tables of the pushforward commit are small integers)
    pts3 = [rand_g1(rng) for _ in range(8)] * 8
    sc3 = [rng.randrange(16) for _ in range(64)]
    assert res_to_point(g.Srs(ctx, aff_to_limbs(pts3)).msm(ctx.upload(to_limbs(sc3)))) == CV.g1_msm(pts3, sc3)
    # prefix / offset of a longer SRS (kzg.rs:125 `&self.ptau_1[..poly.len()]`), and the too-long assert
    srs3 = g.Srs(ctx, aff_to_limbs(pts3))
    assert res_to_point(srs3.msm(ctx.upload(to_limbs(sc3[:10])), n=10, first=3)) == CV.g1_msm(pts3[3:13], sc3[:10])
    with pytest.raises(g.GkrError):
        srs3.msm(ctx.upload(to_limbs(sc3)), n=64, first=1)


def test_msm_projective_bases(ctx):
    rng = random.Random(9)
    pts = [rand_g1(rng) for _ in range(20)]
    sc = [rng.randrange(P) for _ in range(20)]
    jac = np.zeros((20, 18), np.uint64)
    for i, p in enumerate(pts):
        z = rng.randrange(1, FQ_MODULUS)
        jac[i] = fq_vec_to_mont_u64([p[0] * z * z % FQ_MODULUS, p[1] * z * z * z % FQ_MODULUS, z]).reshape(18)
    jac[5] = 0  # Z == 0: infinity
    want = CV.g1_msm([p for i, p in enumerate(pts) if i != 5], [s for i, s in enumerate(sc) if i != 5])
    srs = g.Srs(ctx, jac, projective=True)
    assert res_to_point(srs.msm(ctx.upload(to_limbs(sc)))) == want


def test_msm_linearity_large(ctx):
    """size-independent property at 2^16: msm(bases, a*s + t) == a*msm(bases, s) + msm(bases, t)."""
    rng = random.Random(11)
    n = 1 << 16
    base_pts = [rand_g1(rng) for _ in range(16)]
    pts = aff_to_limbs(base_pts)[np.arange(n) % 16]
    srs = g.Srs(ctx, pts)
    s_tab, t_tab = ctx.synth(1, n), ctx.synth(2, n)
    s = ctx.upload(s_tab.download())
    ms, mt = res_to_point(srs.msm(s_tab)), res_to_point(srs.msm(t_tab))
    # u = s + t computed on the host in Montgomery form (addition commutes with the Montgomery map)
    from oracle.pyref.field import fr_vec_from_mont_u64
    sv, tv = fr_vec_from_mont_u64(s_tab.download()), fr_vec_from_mont_u64(t_tab.download())
    u = ctx.upload(to_limbs([(x + y) % P for x, y in zip(sv, tv)]))
    assert res_to_point(srs.msm(u)) == CV.g1_add(ms, mt)
    # and against the oracle through the 16 distinct bases
    agg = [sum(sv[i] for i in range(j, n, 16)) % P for j in range(16)]
    assert ms == CV.g1_msm(base_pts, agg)


@pytest.mark.parametrize("kind", ["equal", "small", "mixed"])
def test_msm_skewed_buckets(ctx, kind):
    """bucket sizes far from uniform (>= 256 entries: the block-per-bucket path; counting-sorted light buckets)"""
    rng = random.Random(len(kind))
    n = 3000
    base_pts = [rand_g1(rng) for _ in range(16)]
    pts = [base_pts[rng.randrange(16)] for _ in range(n)]
    if kind == "equal":
        s0 = rng.randrange(P)
        sc = [s0] * n
    elif kind == "small":
        sc = [rng.randrange(4) for _ in range(n)]
    else:
        sc = [rng.randrange(P) if i % 3 else P - 1 for i in range(n)]
    srs = g.Srs(ctx, aff_to_limbs(pts))
    got = res_to_point(srs.msm(ctx.upload(to_limbs(sc))))
    # group by base point: sum_j (sum of the scalars on base j) * base_j
    tot = {}
    for p, s in zip(pts, sc):
        tot[p] = (tot.get(p, 0) + s) % P
    assert got == CV.g1_msm(list(tot.keys()), list(tot.values()))


@pytest.mark.parametrize("log_n", [18, 20])
def test_msm_large_closed_form(ctx, log_n):
    """Sizes the textbook oracle cannot reach (2^20 points: window reduction in two levels, >= 2^20 buckets): over the mock
    SRS tau^i * g0 (kzg.rs:84-97) the commitment is (sum_i s_i tau^i) * g0 -- the shortcut the reference's own mock setup
    makes possible -- so the device result is checked against one host scalar multiplication."""
    from gkr_msm_b200 import hostmath as H

    n = 1 << log_n
    tau = 0x5DEECE66D1234567890ABCDEF0123456789ABCDEF0FEDCBA9876543210F00D
    srs = g.Srs.mock_setup(ctx, to_limbs([tau])[0], H.g1_to_limbs(CV.G1_GEN), n)
    sc = ctx.synth(1234 + log_n, n)
    got = res_to_point(srs.msm(sc))
    if log_n == 20:  # the same through the fixed-base window table (c = 20: what the 2^20-point prover uses)
        srs.precompute(20)
        assert res_to_point(srs.msm(sc)) == got
    vals = from_limbs_fast(sc.download())
    acc, pw = 0, 1
    for v in vals:
        acc = (acc + v * pw) % P
        pw = pw * tau % P
    assert got == CV.g1_mul(acc, CV.G1_GEN)
    # a sub-range with an offset into the SRS: sum_{i < m} s_i tau^(first + i)
    first, m = 12345, n - 20000
    got2 = res_to_point(srs.msm(sc, n=m, first=first))
    acc2, pw = 0, pow(tau, first, P)
    for v in vals[:m]:
        acc2 = (acc2 + v * pw) % P
        pw = pw * tau % P
    assert got2 == CV.g1_mul(acc2, CV.G1_GEN)


@pytest.mark.parametrize("n,c", [(5000, 12), (70000, 14), (300, 12)])
def test_msm_fixed_base_table_equals_plain(ctx, n, c):
    """gkr_srs_precompute: the one-bucket-set MSM over T[k][i] = 2^(c k) P_i gives the same point as the windowed MSM and
    as the closed form over the mock SRS; small / offset calls fall back or index the table correctly."""
    from gkr_msm_b200 import hostmath as H

    tau = 0x1F2E3D4C5B6A79880123456789ABCDEF
    srs = g.Srs.mock_setup(ctx, to_limbs([tau])[0], H.g1_to_limbs(CV.G1_GEN), n)
    sc = ctx.synth(99 + n, n)
    plain = srs.msm(sc)
    plain_part = srs.msm(sc, n=n - 7, first=5)
    srs.precompute(c)
    assert np.array_equal(srs.msm(sc), plain)
    assert np.array_equal(srs.msm(sc, n=n - 7, first=5), plain_part)
    vals = from_limbs_fast(sc.download())
    acc, pw = 0, 1
    for v in vals:
        acc = (acc + v * pw) % P
        pw = pw * tau % P
    assert res_to_point(plain) == CV.g1_mul(acc, CV.G1_GEN)
    # zeros, ones and r - 1 as scalars
    edge = ctx.upload(to_limbs([0, 1, P - 1] * (n // 3) + [0] * (n % 3)))
    srs2 = g.Srs.mock_setup(ctx, to_limbs([tau])[0], H.g1_to_limbs(CV.G1_GEN), n)
    assert np.array_equal(srs.msm(edge), srs2.msm(edge))


def from_limbs_fast(arr):
    """Montgomery u64 limbs -> ints (vectorised split, python big-int reduction)"""
    a = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
    rinv = pow(1 << 256, -1, P)
    return [((int(r[0]) | (int(r[1]) << 64) | (int(r[2]) << 128) | (int(r[3]) << 192)) * rinv) % P for r in a]


def test_msm_multi_equals_single_calls(ctx):
    """gkr_msm_g1_multi: k scalar tables of different lengths over the same bases == k gkr_msm_g1 calls"""
    import ctypes as C

    from gkr_msm_b200 import hostmath as H

    n = 3000
    srs = g.Srs.mock_setup(ctx, to_limbs([0xABCDEF12345])[0], H.g1_to_limbs(CV.G1_GEN), n)
    tabs = [ctx.synth(1, n), ctx.synth(2, n), ctx.upload(to_limbs([i % 7 for i in range(n)])), ctx.synth(3, 64)]
    lens = [n, n - 5, n, 64]
    want = np.stack([srs.msm(t, n=ln, first=2 if ln < n else 0) if False else srs.msm(t, n=ln) for t, ln in zip(tabs, lens)])
    lib = ctx.lib
    lib.gkr_msm_g1_multi.restype = C.c_int
    lib.gkr_msm_g1_multi.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), C.c_void_p, C.c_uint32, C.c_void_p]
    arr = (C.c_void_p * 4)(*[t.h for t in tabs])
    ln = np.array(lens, dtype=np.uint64)
    out = np.zeros((4, 12), np.uint64)
    ctx.check(lib.gkr_msm_g1_multi(ctx.h, srs.h, 0, arr, ln.ctypes.data_as(C.c_void_p), 4, out.ctypes.data_as(C.c_void_p)))
    assert np.array_equal(out, want)


@pytest.fixture(params=[2, 0], ids=["signed", "unsigned"])
def recoding(ctx, request):
    """both window recodings of the bucket MSM (msm_recode, msm.cu): signed digits (forced: the default uses them from 2^15
    points) and unsigned"""
    ctx.set_tuning("msm_signed", request.param)
    yield request.param
    ctx.set_tuning("msm_signed", 1)


@pytest.mark.parametrize("n", [300, 3000])
def test_msm_signed_digit_boundaries(ctx, recoding, n):
    """scalars built from the boundary digits of the signed recoding -- windows equal to 2^(c-1) (stay positive, bucket 0 =
    the magnitude 2^(c-1)), 2^(c-1) + 1 (first negative), 2^c - 1 (carry chains through every window), r - 1 and values whose
    top window takes a carry -- for every window width the size picks (c = 8 .. 12 here), against the textbook oracle"""
    rng = random.Random(1000 + n)
    base = [rand_g1(rng) for _ in range(16)]
    pts = [base[i % 16] for i in range(n)]
    sc = []
    for c in (8, 9, 10, 11, 12, 13):
        half = 1 << (c - 1)
        for digit in (half, half + 1, half - 1, (1 << c) - 1, 1):
            v = 0
            for w in range(0, 255, c):
                v |= digit << w
            sc.append(v % P)
        sc.append(sum((half if (w // c) % 2 == 0 else (1 << c) - 1) << w for w in range(0, 255, c)) % P)
    sc += [P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, (1 << 254) - 1, (1 << 254), (1 << 254) + (1 << 253), 0, 1]
    sc += [rng.randrange(P) for _ in range(n - len(sc))]
    srs = g.Srs(ctx, aff_to_limbs(pts))
    got = res_to_point(srs.msm(ctx.upload(to_limbs(sc))))
    assert got == CV.g1_msm(pts, sc)


def test_msm_signed_equals_unsigned_at_scale(ctx):
    """2^17 and 2^19 points over the mock SRS (windows of 14 / 15 bits, one- and two-level bucket reduction): both recodings
    return the same affine point, which is the closed form sum_i s_i tau^i G; same for the fixed-base table path"""
    from gkr_msm_b200 import hostmath as H
    from gkr_msm_b200.fieldutil import to_limb1

    tau = 0x1234567890ABCDEF1234567
    for log_n in (17, 19):
        n = 1 << log_n
        srs = g.Srs.mock_setup(ctx, to_limb1(tau), H.g1_to_limbs(H.G1_GEN), n)
        sc = ctx.synth(4000 + log_n, n)
        res = []
        for sg in (1, 0):
            ctx.set_tuning("msm_signed", sg)
            res.append(srs.msm(sc).tolist())
        ctx.set_tuning("msm_signed", 1)
        assert res[0] == res[1]
        if log_n == 17:
            srs.precompute(16)  # 16 x 16 = 256 bits: the signed fixed-base path
            assert srs.msm(sc).tolist() == res[0]
            vals = from_limbs_fast(sc.download())
            acc, t = 0, 1
            for v in vals:
                acc = (acc + v * t) % P
                t = t * tau % P
            assert res_to_point(np.array(res[0], dtype=np.uint64)) == CV.g1_mul(acc, CV.G1_GEN)
        srs.free()


def test_msm_signed_projective_and_batch(ctx, recoding):
    """negative digits negate Jacobian and XYZZ bases too (msm_nonaff over bucket sums, gkr_msm_g1_batch)"""
    rng = random.Random(77)
    n = 600
    base = [rand_g1(rng) for _ in range(12)]
    pts = [base[i % 12] for i in range(n)]
    sc = [rng.randrange(P) for _ in range(n)]
    sc[:4] = [P - 1, (1 << 7) | (1 << 15), (P + 1) // 2, 255]
    zs = [rng.randrange(1, FQ_MODULUS) for _ in range(n)]
    proj = np.zeros((n, 18), np.uint64)
    for i, (p, z) in enumerate(zip(pts, zs)):
        z2 = z * z % FQ_MODULUS
        proj[i] = fq_vec_to_mont_u64([p[0] * z2 % FQ_MODULUS, p[1] * z2 * z % FQ_MODULUS, z]).reshape(18)
    srs = g.Srs(ctx, proj, projective=True)
    assert res_to_point(srs.msm(ctx.upload(to_limbs(sc)))) == CV.g1_msm(pts, sc)
