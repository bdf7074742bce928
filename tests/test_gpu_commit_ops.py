"""GPU parity for the commitment-side algebra vs the oracle restatement."""
import random

import numpy as np
import pytest

import gkr_msm_b200 as g
from oracle.pyref import commitments as OC
from oracle.pyref import curves as CV
from oracle.pyref import sumcheck as S
from oracle.pyref.field import P
from tests.test_gpu_msm import aff_to_limbs, rand_g1, res_to_point
from tests.util import from_limbs, to_limb1, to_limbs

pytestmark = pytest.mark.gpu


def test_bucket_sums_running_sum_and_msm_nonaff(ctx):
    """pushforward.rs:398-429, 504-524, 598-604 on a miniature instance: digits -> bucket sums -> d commitment ==
    commit(d table) (the reference's own commented-out assertion :527-530), then msm_nonaff(buckets, eq_d)."""
    rng = random.Random(1)
    x_size, d_log = 64, 3
    bases = [rand_g1(rng) for _ in range(x_size)]
    digits = [rng.randrange(1 << d_log) for _ in range(x_size)]
    srs = g.Srs(ctx, aff_to_limbs(bases))
    bk = srs.bucket_sums(np.arange(x_size), digits, 1 << d_log)
    want = OC.bucket_sums(bases, range(x_size), digits, 1 << d_log)
    got = [res_to_point(r) for r in bk.download_affine()]
    assert got == want
    d_comm = res_to_point(bk.weighted_sum())
    assert d_comm == OC.running_sum_commit(want)
    assert d_comm == CV.g1_msm(bases, digits)  # == KzgProvingKey::commit(d)
    dtab = g.U32Buf(ctx, digits).to_field()
    assert from_limbs(dtab.download()) == digits
    assert res_to_point(srs.msm(dtab)) == d_comm
    # second phase: msm_nonaff over the bucket bases with eq_d as scalars
    r_d = [rng.randrange(P) for _ in range(d_log)]
    eq_d = S.eq_poly_sequence_last(r_d)
    got2 = res_to_point(bk.msm(ctx.eq_table(to_limbs(r_d))))
    assert got2 == CV.g1_msm([b for b in want], eq_d)
    # == commit(d_pull) with d_pull[x] = eq_d[digit[x]]  (pushforward.rs:606-609)
    d_pull = ctx.gather(ctx.eq_table(to_limbs(r_d)), g.U32Buf(ctx, digits))
    assert from_limbs(d_pull.download()) == [eq_d[d] for d in digits]
    assert res_to_point(srs.msm(d_pull)) == got2
    # access counts are negated counts
    ac = [0] * (1 << d_log)
    for d in digits:
        ac[d] += 1
    assert from_limbs(g.U32Buf(ctx, ac).to_field(negate=True).download()) == [(-c) % P for c in ac]
    with pytest.raises(g.GkrError):
        srs.bucket_sums([0, 1], [0, 99], 8)


@pytest.mark.parametrize("n", [1, 2, 5, 64, 511, 512, 513, 1000, 5000, 32768, 32769, 40001])  # 1, 2 and 3 chunk levels
def test_poly_eval_and_div_by_linear(ctx, n):
    rng = random.Random(n)
    poly = [rng.randrange(P) for _ in range(n)]
    x = rng.randrange(P)
    tab = ctx.upload(to_limbs(poly))
    assert from_limbs(ctx.poly_eval(tab, to_limb1(x)).reshape(1, 4))[0] == OC.ev(poly, x)
    q, rem = ctx.div_by_linear(tab, to_limb1(x))
    oq, orem = OC.div_by_linear(poly, x)
    assert from_limbs(rem.reshape(1, 4))[0] == orem
    assert (from_limbs(q.download()) if n > 1 else []) == oq


@pytest.mark.parametrize("num_vars,plen", [(1, 2), (3, 8), (5, 20), (8, 256)])
def test_knuckles_compute_t(ctx, num_vars, plen):
    rng = random.Random(num_vars)
    k = 2
    inv = OC.knuckles_inverses(num_vars, k)
    poly = [rng.randrange(P) for _ in range(plen)]
    point = [rng.randrange(P) for _ in range(num_vars)]
    key = g.Knuckles(ctx, num_vars, to_limb1(k))
    t, opening = key.compute_t(ctx.upload(to_limbs(poly)), to_limbs(point))
    ot, oopen = OC.compute_t(num_vars, inv, poly, point)
    assert from_limbs(opening.reshape(1, 4))[0] == oopen
    assert from_limbs(t.download()) == ot
    # knuckles.rs:326-340: the opening is the multilinear evaluation of the (zero-padded) table
    padded = poly + [0] * ((1 << num_vars) - plen)
    assert oopen == S.evaluate_poly(padded, point)


def test_lincomb_slices(ctx):
    rng = random.Random(4)
    a = [rng.randrange(P) for _ in range(16)]
    b = [rng.randrange(P) for _ in range(8)]
    ca, cb = rng.randrange(P), rng.randrange(P)
    ta, tb = ctx.upload(to_limbs(a)), ctx.upload(to_limbs(b))
    # zero-extended lambda*t + p  (opening.rs:65-75)
    got = from_limbs(ctx.lincomb([(ta, to_limb1(ca), 0, 0, 16), (tb, to_limb1(1), 0, 0, 8)], 16).download())
    assert got == [(ca * a[i] + (b[i] if i < 8 else 0)) % P for i in range(16)]
    # strided accumulation like combined_witness (pippenger.rs:209-223): rows 0 and 2 of a 4x4 matrix into one row
    got = from_limbs(ctx.lincomb([(ta, to_limb1(ca), 0, 0, 4), (ta, to_limb1(cb), 8, 0, 4)], 4).download())
    assert got == [(ca * a[i] + cb * a[8 + i]) % P for i in range(4)]
    with pytest.raises(g.GkrError):
        ctx.lincomb([(tb, to_limb1(1), 4, 0, 8)], 8)


def test_batched_bucket_commitments(ctx):
    """all commitment chunks in one pass: combined bucket ids, grouped running sums, msm_nonaff per chunk (same scalars)"""
    rng = random.Random(12)
    x_size, d_log, chunks = 48, 3, 3
    bases = [rand_g1(rng) for _ in range(x_size)]
    srs = g.Srs(ctx, aff_to_limbs(bases))
    digits = [[rng.randrange(1 << d_log) for _ in range(x_size)] for _ in range(chunks)]
    digits[1] = [5] * x_size  # one heavy bucket, the others empty
    pidx = np.tile(np.arange(x_size, dtype=np.uint32), chunks)
    bidx = np.concatenate([np.array(digits[k], dtype=np.uint32) + (k << d_log) for k in range(chunks)])
    allb = srs.bucket_sums(pidx, bidx, chunks << d_log)
    want = [OC.bucket_sums(bases, range(x_size), digits[k], 1 << d_log) for k in range(chunks)]
    got = [res_to_point(r) for r in allb.download_affine()]
    assert got == [b for w in want for b in w]
    comms = [res_to_point(r) for r in allb.weighted_sums(d_log, chunks)]
    assert comms == [OC.running_sum_commit(w) for w in want]
    assert comms == [CV.g1_msm(bases, digits[k]) for k in range(chunks)]
    r_d = [rng.randrange(P) for _ in range(d_log)]
    eq_d = S.eq_poly_sequence_last(r_d)
    pulls = [res_to_point(r) for r in allb.msm_batch(ctx.eq_table(to_limbs(r_d)), 1 << d_log, 0, 1 << d_log, chunks)]
    assert pulls == [CV.g1_msm(want[k], eq_d) for k in range(chunks)]
    # a sub-range of the groups, and range checking
    assert [res_to_point(r) for r in allb.weighted_sums(d_log, 2, first=1 << d_log)] == comms[1:]
    with pytest.raises(g.GkrError):
        allb.weighted_sums(d_log, chunks + 2)
    with pytest.raises(g.GkrError):
        allb.msm_batch(ctx.eq_table(to_limbs(r_d)), 1 << d_log, 0, 1 << d_log, chunks + 1)


@pytest.mark.parametrize("clm", [0, 1, 2])
def test_bucket_sums_rows_matches_host_index_route(ctx, clm):
    """device-resident digit matrix -> bucket sums == the explicit (point index, bucket index) route, incl. a ragged last chunk"""
    rng = random.Random(20 + clm)
    xl, d_log, y_size = 4, 2, 5
    x_size, cm = 1 << xl, 1 << clm
    bases = [rand_g1(rng) for _ in range(x_size * cm)]
    srs = g.Srs(ctx, aff_to_limbs(bases))
    digits = np.array([[rng.randrange(1 << d_log) for _ in range(x_size)] for _ in range(y_size)], dtype=np.uint32)
    got = srs.bucket_sums_rows(g.U32Buf(ctx, digits.reshape(-1)), xl, clm, d_log)
    n_comms = -(-y_size // cm)
    assert got.n == n_comms << d_log
    pidx = np.concatenate([np.arange(x_size, dtype=np.uint32) + x_size * (y % cm) for y in range(y_size)])
    bidx = np.concatenate([digits[y] + ((y // cm) << d_log) for y in range(y_size)])
    want = srs.bucket_sums(pidx, bidx, n_comms << d_log)
    assert np.array_equal(got.download_affine(), want.download_affine())
    comms = [res_to_point(r) for r in got.weighted_sums(d_log, n_comms)]
    for k in range(n_comms):
        ys = range(k * cm, min((k + 1) * cm, y_size))
        assert comms[k] == CV.g1_msm([bases[x + x_size * (y % cm)] for y in ys for x in range(x_size)], [int(digits[y][x]) for y in ys for x in range(x_size)])
    with pytest.raises(g.GkrError):  # a digit that does not fit group_log bits
        srs.bucket_sums_rows(g.U32Buf(ctx, digits.reshape(-1)), xl, clm, 1)


@pytest.mark.parametrize("n,y_size,d", [(1, 1, 1), (37, 5, 3), (2048, 3, 8), (2049, 16, 8), (5000, 13, 10), (70001, 4, 13), (4097, 32, 8)])
def test_bucketize_device_equals_host(ctx, n, y_size, d):
    """gkr_pushforward_bucketize_dev (stable counting sort on the device) == the host bookkeeping (pushforward.rs:351-396):
    digits, in-bucket ranks in input order, bucket sizes, and the even-padded bucket contents."""
    rng = np.random.default_rng(n * 31 + d)
    co = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    if n > 100:
        co[: n // 3] = co[0]  # heavy buckets: many equal scalars in a row
        co[n // 2] = 0
    hd, hc, ho, hl = g.pushforward_bucketize(co, y_size, d)
    dd, dc, dpo, dl = g.pushforward_bucketize_dev(ctx, co, y_size, d)
    assert np.array_equal(dl, hl) and np.array_equal(dd, hd) and np.array_equal(dc, hc)
    # host padded order: every bucket's x list, padded to even length with 0xffffffff
    exp = []
    for y in range(y_size):
        off = 0
        for b in range(1 << d):
            ln = int(hl[y, b])
            exp.append(ho[y, off:off + ln])
            if ln & 1:
                exp.append(np.array([0xFFFFFFFF], np.uint32))
            off += ln
    exp = np.concatenate(exp) if exp else np.zeros(0, np.uint32)
    assert np.array_equal(dpo, exp)


def test_ctx_trim_keeps_the_context_usable(ctx):
    """gkr_ctx_trim hands cached blocks back to the driver; objects created afterwards work as before"""
    import ctypes as C

    t = ctx.synth(3, 1 << 18)  # 8 MiB: goes through the large-block cache
    before = t.download()[:4].copy()
    t.free()
    ctx.lib.gkr_ctx_trim.restype = C.c_int
    ctx.lib.gkr_ctx_trim.argtypes = [C.c_void_p]
    ctx.check(ctx.lib.gkr_ctx_trim(ctx.h))
    t2 = ctx.synth(3, 1 << 18)
    assert np.array_equal(t2.download()[:4], before)
