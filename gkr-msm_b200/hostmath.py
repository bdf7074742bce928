"""O(1)-per-proof host arithmetic of the top-level protocol over python ints (nothing table-sized): verifier polynomials the
PROVER also evaluates, univariate interpolation on the nodes 0..deg, the ark-serialize G1 wire format and the Bandersnatch
group law used to generate synthetic inputs.  In the reference all of this is host-side Rust and stays there (north_star).

  eq_poly_sequence_last / eq_sum                     src/utils.rs:189-291
  EqTruncPoly / SelectorPoly ::evaluate              src/cleanup/protocols/verifier_polys.rs:60-137
  UniPoly::from_evals                                liblasso (interpolation on 0..deg)
  compress_coefficients / evaluate_univar            src/cleanup/protocols/sumcheck.rs:27-44
  G1 compressed encoding                             ark-bls12-381 0.4.0 (zcash / IETF format), proof_transcript.rs:52-69
"""
from __future__ import annotations

import numpy as np

from .fieldutil import R_MOD

P = R_MOD
Q_MOD = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
_QR = (1 << 384) % Q_MOD
_QRINV = pow(1 << 384, -1, Q_MOD)
_M64 = 0xFFFFFFFFFFFFFFFF

# Bandersnatch (twisted Edwards a x^2 + y^2 = 1 + d x^2 y^2 over BLS12-381 Fr), src/utils.rs:32-49
TE_A = P - 5
TE_D = 45022363124591815672509500913686876175488063829319466900776701791074614335719
TE_GEN = (18886178867200960497001835917649091219057080094937609519140440539760939937304,
          19188667384257783945677642223292697773471335439753913231509108946878080696678)
G1_GEN = (0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
          0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1)


def inv(x: int) -> int:
    return pow(x % P, -1, P)


def eq_poly_sequence_last(pt):  # utils.rs:222-262 (last level); pt[-1] <-> least-significant index bit
    ret = [1]
    for r in pt:
        nxt = []
        for w in ret:
            hi = w * r % P
            nxt.append((w - hi) % P)
            nxt.append(hi)
        ret = nxt
    return ret


def eq_sum(pt, k):  # utils.rs:265-291: sum of eq(pt, i) over i < k
    n = len(pt)
    if k >= (1 << n):
        assert k == 1 << n
        return 1
    mult, acc = 1, 0
    for i in range(n):
        left_bit = k >> (n - i - 1)
        old = mult
        if left_bit == 1:
            mult = mult * pt[i] % P
            acc = (acc + old - mult) % P
        else:
            mult = mult * (1 - pt[i]) % P
        k -= left_bit << (n - i - 1)
    return acc


def eq_trunc_evals(num_vars, k, r):  # verifier_polys.rs:90-96
    ret = eq_poly_sequence_last(r)
    for i in range(k, 1 << num_vars):
        ret[i] = 0
    return ret


def eq_trunc_evaluate(num_vars, k, r, pt):  # verifier_polys.rs:98-136
    assert len(pt) == num_vars
    partial = [1]
    for i in range(num_vars):
        j = num_vars - i - 1
        partial.append(partial[-1] * ((1 - pt[j] - r[j] + 2 * r[j] * pt[j]) % P) % P)
    if k >= (1 << num_vars):
        assert k == 1 << num_vars
        return partial[num_vars]
    multiplier, acc = 1, 0
    for i in range(num_vars):
        left_bit = k >> (num_vars - i - 1)
        m_ = multiplier
        if left_bit == 1:
            multiplier = multiplier * pt[i] % P * r[i] % P
            acc = (acc + m_ * (1 - pt[i]) % P * (1 - r[i]) % P * partial[num_vars - i - 1]) % P
        else:
            multiplier = multiplier * (1 - pt[i]) % P * (1 - r[i]) % P
        k -= left_bit << (num_vars - i - 1)
    return acc


def evaluate_poly(poly, pt):  # cleanup/utils/arith.rs:6-9
    e = eq_poly_sequence_last(pt)
    assert len(e) == len(poly)
    return sum(a * b for a, b in zip(poly, e)) % P


def gamma_rlc(gamma, vals):  # sumcheck.rs:591-602
    if not vals:
        return 0
    ret = vals[-1]
    for v in reversed(vals[:-1]):
        ret = (ret * gamma + v) % P
    return ret


def from_evals(evals):
    """coefficients (low -> high) of the polynomial with the given values on 0..deg (liblasso UniPoly::from_evals)."""
    n = len(evals)
    coeffs = [0] * n
    for i, y in enumerate(evals):
        num, den = [1], 1  # prod_{j != i} (X - j) / (i - j)
        for j in range(n):
            if j == i:
                continue
            nxt = [0] * (len(num) + 1)
            for k, c in enumerate(num):
                nxt[k] = (nxt[k] - j * c) % P
                nxt[k + 1] = (nxt[k + 1] + c) % P
            num = nxt
            den = den * (i - j) % P
        s = y * inv(den) % P
        for k, c in enumerate(num):
            coeffs[k] = (coeffs[k] + s * c) % P
    return coeffs


def evaluate_univar(coeffs, x):  # sumcheck.rs:33-44
    ret = 0
    for c in reversed(coeffs):
        ret = (ret * x + c) % P
    return ret


def compress_coefficients(coeffs):  # sumcheck.rs:27-31
    return [coeffs[0]] + list(coeffs[2:])


def from_le_bytes_mod_order(b: bytes) -> int:
    return int.from_bytes(b, "little") % P


# ---- G1 points at the boundary: (12,) uint64 = x | y in Fq Montgomery limbs, all zero = infinity -----------------
def g1_from_limbs(a):
    a = [int(v) for v in np.asarray(a, dtype=np.uint64).reshape(12)]
    x = sum(a[i] << (64 * i) for i in range(6)) * _QRINV % Q_MOD
    y = sum(a[6 + i] << (64 * i) for i in range(6)) * _QRINV % Q_MOD
    return None if x == 0 and y == 0 else (x, y)


def g1_to_limbs(pt) -> np.ndarray:
    out = np.zeros(12, np.uint64)
    if pt is not None:
        for k, v in enumerate(pt):
            m = v % Q_MOD * _QR % Q_MOD
            for i in range(6):
                out[6 * k + i] = (m >> (64 * i)) & _M64
    return out


def g1_serialize(pt) -> bytes:
    """48-byte big-endian x; top bits of byte 0 = (compressed, infinity, y is the lexicographically larger root)."""
    if pt is None:
        return bytes([0xC0]) + bytes(47)
    x, y = pt
    b = bytearray(x.to_bytes(48, "big"))
    b[0] |= 0x80
    if y > (Q_MOD - 1) // 2:
        b[0] |= 0x20
    return bytes(b)


# ---- synthetic inputs (build_pippenger_data, pippenger.rs:462-497) ------------------------------------------------
def te_add_proj(p1, p2):  # add-2008-bbjlp
    X1, Y1, Z1 = p1
    X2, Y2, Z2 = p2
    A = Z1 * Z2 % P
    B = A * A % P
    C = X1 * X2 % P
    D = Y1 * Y2 % P
    E = TE_D * C % P * D % P
    F = (B - E) % P
    G = (B + E) % P
    return (A * F % P * ((X1 + Y1) * (X2 + Y2) - C - D) % P, A * G % P * (D - TE_A * C) % P, F * G % P)


def te_mul(k, pt):
    acc, base = (0, 1, 1), (pt[0], pt[1], 1)
    while k:
        if k & 1:
            acc = te_add_proj(acc, base)
        base = te_add_proj(base, base)
        k >>= 1
    return acc


def te_points_arithmetic_progression(k0: int, step: int, n: int):
    """n affine points (k0 + i*step) * G of the prime-order subgroup: one projective addition per point and one shared
    inversion (Montgomery's trick).  On-curve, pairwise distinct -- the distribution SURVEY 8d asks for, without sqrt."""
    cur, q = te_mul(k0, TE_GEN), te_mul(step, TE_GEN)
    proj = []
    for _ in range(n):
        proj.append(cur)
        cur = te_add_proj(cur, q)
    pref, acc = [], 1
    for p in proj:
        pref.append(acc)
        acc = acc * p[2] % P
    ia = inv(acc)
    out = [None] * n
    for i in range(n - 1, -1, -1):
        zi = ia * pref[i] % P
        ia = ia * proj[i][2] % P
        out[i] = (proj[i][0] * zi % P, proj[i][1] * zi % P)
    return out
