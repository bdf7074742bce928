"""Old-API bit-column commitment: prepare_bases / binary_msm (src/binary_msm.rs:19-54) on the device
against the oracle restatement and -- like the reference's own tests bin_msm / bin_msm_gamma_3 (binary_msm.rs:62-96) --
against the plain sum of the bases whose bit is set."""
import random

import numpy as np
import pytest

import gkr_msm_b200 as g
from oracle.pyref import commitments as OC
from oracle.pyref import curves as CV
from tests.test_gpu_msm import aff_to_limbs, rand_g1, res_to_point

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("num,gamma", [(100, 8), (100, 3), (7, 4), (64, 1), (1, 5)])
def test_binary_msm_matches_reference_tests(ctx, num, gamma):
    rng = random.Random(num * 10 + gamma)
    bits = [rng.random() < 0.5 for _ in range(num)]
    pool = [rand_g1(rng) for _ in range(min(num, 12))]
    bases = [pool[rng.randrange(len(pool))] for _ in range(num)]  # repeated bases: subset sums hit the doubling path
    pcoefs = OC.prepare_coefs(bits, gamma)
    srs = g.Srs(ctx, aff_to_limbs(bases))
    prepared = g.binary_msm_prepare(ctx, srs, gamma)
    # the prepared tables equal prepare_bases entry for entry
    want = [p for chunk in OC.prepare_bases(bases, gamma) for p in chunk]
    got = [res_to_point(r) for r in prepared.download_affine()]
    assert got == want
    res = res_to_point(g.binary_msm(ctx, prepared, gamma, pcoefs))
    assert res == OC.binary_msm(pcoefs, OC.prepare_bases(bases, gamma))
    expected = None
    for c, b in zip(bits, bases):
        if c:
            expected = CV.g1_add(expected, b)
    assert res == expected  # binary_msm.rs:76-77
    # all-zero coefficients select nothing: the point at infinity
    assert res_to_point(g.binary_msm(ctx, prepared, gamma, [0] * len(pcoefs))) is None
    with pytest.raises(g.GkrError):  # coefs.len() != bases.len() (binary_msm.rs:21)
        g.binary_msm(ctx, prepared, gamma, pcoefs + [0])


def test_binary_msm_large_sum(ctx):
    """2^16 chunks: the strided block partial sums and the final tree"""
    rng = random.Random(5)
    pool = [rand_g1(rng) for _ in range(8)]
    n = 1 << 16
    bases = [pool[i % 8] for i in range(n)]
    gamma = 1
    srs = g.Srs(ctx, aff_to_limbs(bases))
    prepared = g.binary_msm_prepare(ctx, srs, gamma)
    coefs = np.array([rng.randrange(2) for _ in range(n)], dtype=np.uint8)
    cnt = [int(coefs[j::8].sum()) for j in range(8)]
    expected = CV.g1_msm(pool, cnt)
    assert res_to_point(g.binary_msm(ctx, prepared, gamma, coefs)) == expected
