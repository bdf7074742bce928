"""CPU tests pinning the oracle's GKR restatement with the reference's own tests:
bintree_add.rs:401-460 prove_and_verify, triangle_add.rs:277-393 witness_gen / prove_and_verify,
pippenger_ending.rs:176-275 integration (claims == MLE of the inputs, outputs == sum_b b * bucket_sum)."""
import random

import pytest

from oracle.pyref import curves as CV
from oracle.pyref import gates as G
from oracle.pyref import gkr as K
from oracle.pyref import polys as OP
from oracle.pyref import sumcheck as S
from oracle.pyref.field import P
from oracle.pyref.transcript import ProofTranscript2


def rand_points_affine(rng, row_logsize, col_logsize, full=False):
    """VecVecPolynomial::rand_points_affine (vecvec.rs:347-378): x padded with 0, y with 1 (the identity)."""
    nrows = (1 << col_logsize) if full else rng.randrange(1 << col_logsize) + 1
    rows = [[CV.te_random_point(rng) for _ in range((1 << row_logsize) if full else rng.randrange(1 << row_logsize) + 1)]
            for _ in range(nrows)]
    xs = S.VecVecPolynomial([[p[0] for p in r] for r in rows], 0, 0, row_logsize, col_logsize)
    ys = S.VecVecPolynomial([[p[1] for p in r] for r in rows], 1, 1, row_logsize, col_logsize)
    return [xs, ys], rows


def dense_of(advice, hint):
    if advice[0] == "vv":
        return [p.vec() for p in advice[1]]
    return [list(c) + [0] * ((1 << hint) - len(c)) for c in advice[1]]


@pytest.mark.parametrize("num_adds,row_logsize,col_logsize", [(5, 4, 2), (5, 2, 4), (3, 3, 1)])
def test_bintree_prove_and_verify(num_adds, row_logsize, col_logsize):
    rng = random.Random(10 * num_adds + row_logsize)
    num_vars = row_logsize + col_logsize
    points, _ = rand_points_affine(rng, row_logsize, col_logsize)
    inputs = OP.vecvec_map_split(points, G.Id(2), ("LO", 0), 2)
    advices = K.bintree_witness(("vv", inputs), row_logsize, num_adds, False)
    layers = K.bintree_protocol(num_vars, num_adds, row_logsize, False)
    last = K.bintree_last_step(advices[-1], num_adds - 1)
    dense_output = dense_of(last, num_vars - num_adds)
    point = [rng.randrange(P) for _ in range(num_vars - num_adds)]
    claims = (point, [S.evaluate_poly(o, point) for o in dense_output])
    tp = ProofTranscript2.start_prover(b"fgstglsp")
    out_claims = K.simple_gkr_prove(layers, tp, claims, advices)
    tv = ProofTranscript2.start_verifier(b"fgstglsp", tp.end())
    assert K.simple_gkr_verify(layers, tv, claims) == out_claims
    # the reduced claims are evaluations of the circuit inputs
    dense_in = [p.vec() for p in inputs]
    assert out_claims[1] == [S.evaluate_poly(d, out_claims[0]) for d in dense_in]


def test_bintree_witness_is_the_group_law():
    # bintree_add.rs:462-505 witness_gen: every output is the sum of its 2^num_adds leaves
    rng = random.Random(5)
    row_logsize, col_logsize, num_adds = 3, 1, 3
    points, rows = rand_points_affine(rng, row_logsize, col_logsize, full=True)
    inputs = OP.vecvec_map_split(points, G.Id(2), ("LO", 0), 2)
    advices = K.bintree_witness(("vv", inputs), row_logsize, num_adds, False)
    out = dense_of(K.bintree_last_step(advices[-1], num_adds - 1), row_logsize + col_logsize - num_adds)
    flat = [p for r in rows for p in r]
    for idx in range(len(out[0])):
        acc = CV.TE_IDENTITY
        for c in range(1 << num_adds):
            acc = CV.te_add_affine(acc, flat[idx * (1 << num_adds) + c])
        zi = pow(out[2][idx], -1, P)
        assert (out[0][idx] * zi % P, out[1][idx] * zi % P) == acc


def test_triangle_prove_and_verify_and_weights():
    # triangle_add.rs:277-393
    rng = random.Random(8)
    num_vars, split_var = 6, ("HI", 2)
    pts = [CV.te_random_point(rng) for _ in range(1 << num_vars)]
    zs = [rng.randrange(1, P) for _ in pts]
    inputs = [[p[0] * z % P for p, z in zip(pts, zs)], [p[1] * z % P for p, z in zip(pts, zs)], zs]
    inputs = OP.dense_map_split(inputs, G.Id(3), split_var, 3)
    inputs = OP.dense_map_split(inputs, G.Id(6), split_var, 3)
    advices = K.triangle_witness(inputs, num_vars - 2, split_var)
    layers = K.triangle_protocol(num_vars - 2, split_var)
    hi = 2
    last = K.triangle_last_step(advices[-1][1], num_vars - 2 - hi)
    # weights: sum_{i >= 1} 2^(i-1) * result_i == sum_i i * P_i per chunk
    chunk = 1 << (num_vars - hi)
    for idx in range(1 << hi):
        want = CV.te_msm(pts[idx * chunk:(idx + 1) * chunk], list(range(chunk)))
        acc, coef = CV.TE_IDENTITY, 1
        for i in range(1, len(last) // 3):
            z = last[3 * i + 2][idx]
            zi = pow(z, -1, P)
            acc = CV.te_add_affine(acc, CV.te_mul(coef, (last[3 * i][idx] * zi % P, last[3 * i + 1][idx] * zi % P)))
            coef *= 2
        assert acc == want
    point = [rng.randrange(P) for _ in range(hi)]
    claims = (point, [S.evaluate_poly(o, point) for o in last])
    tp = ProofTranscript2.start_prover(b"fgstglsp")
    out_claims = K.simple_gkr_prove(layers, tp, claims, advices)
    tv = ProofTranscript2.start_verifier(b"fgstglsp", tp.end())
    assert K.simple_gkr_verify(layers, tv, claims) == out_claims
    assert out_claims[1] == [S.evaluate_poly(d, out_claims[0]) for d in inputs]


def test_pippenger_ending_integration():
    # pippenger_ending.rs:176-275
    rng = random.Random(21)
    multirow_vars, bucket_vars, point_vars = 1, 3, 3
    pre, rows = rand_points_affine(rng, point_vars, multirow_vars + bucket_vars)
    domain = S.VecVecPolynomial([[1] * len(r) for r in pre[0].data], 0, 0, point_vars, multirow_vars + bucket_vars)
    inputs = OP.vecvec_map_split(pre, G.Id(2), ("LO", 0), 2) + OP.vecvec_map_split([domain], G.Id(1), ("LO", 0), 1)
    dense_input = [p.vec() for p in inputs]
    wg = K.PippengerEndingWG(multirow_vars, bucket_vars, point_vars, inputs)
    ending = K.PippengerBucketed(multirow_vars, bucket_vars, point_vars)
    num_vars = multirow_vars + bucket_vars
    dense_output = K.triangle_last_step(wg.last(), num_vars - 2 - multirow_vars)
    point = [rng.randrange(P) for _ in range(multirow_vars)]
    claims = (point, [S.evaluate_poly(o, point) for o in dense_output])
    tp = ProofTranscript2.start_prover(b"fgstglsp")
    out_claims = ending.prove(tp, claims, wg)
    tv = ProofTranscript2.start_verifier(b"fgstglsp", tp.end())
    assert ending.verify(tv, claims) == out_claims
    assert out_claims[1] == [S.evaluate_poly(d, out_claims[0]) for d in dense_input]
    # outputs == sum_bucket bucket_idx * bucket_sum
    nb = 1 << bucket_vars
    sums = []
    for r in rows:
        acc = CV.TE_IDENTITY
        for p in r:
            acc = CV.te_add_affine(acc, p)
        sums.append(acc)
    sums += [CV.TE_IDENTITY] * ((1 << num_vars) - len(sums))
    for m in range(1 << multirow_vars):
        want = CV.te_msm(sums[m * nb:(m + 1) * nb], list(range(nb)))
        acc, coef = CV.TE_IDENTITY, 1
        for b in range(1, bucket_vars + 1):
            z = dense_output[3 * b + 2][m]
            zi = pow(z, -1, P)
            acc = CV.te_add_affine(acc, CV.te_mul(coef, (dense_output[3 * b][m] * zi % P, dense_output[3 * b + 1][m] * zi % P)))
            coef *= 2
        assert acc == want
