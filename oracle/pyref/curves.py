"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

Independent group-law oracles (textbook formulas over python ints), used the way the reference's tests use
arkworks' `Affine + Affine` / `Projective + Projective` / `G::msm` (bintree_add.rs:401-638, pullback.rs:86-107):

* Bandersnatch twisted Edwards curve  a x^2 + y^2 = 1 + d x^2 y^2  over BLS12-381 Fr, a = -5
  (ark-ed-on-bls12-381-bandersnatch 0.4.0; src/utils.rs:32-49 hard-codes a and d).
* BLS12-381 G1  y^2 = x^3 + 4  over Fq (ark-bls12-381 0.4.0), the commitment group (src/commitments/kzg.rs).
"""
from __future__ import annotations

from .field import FQ_MODULUS, FR_MODULUS, TE_A, TE_D

P = FR_MODULUS
Q = FQ_MODULUS

# ---- Bandersnatch (twisted Edwards) -----------------------------------------------------------------------
# prime-order subgroup generator of ark-ed-on-bls12-381-bandersnatch (checked on-curve at import)
TE_GEN = (18886178867200960497001835917649091219057080094937609519140440539760939937304,
          19188667384257783945677642223292697773471335439753913231509108946878080696678)
TE_SUBGROUP_ORDER = 13108968793781547619861935127046491459309155893440570251786403306729687672801
TE_IDENTITY = (0, 1)


def te_on_curve(pt) -> bool:
    x, y = pt
    return (TE_A * x * x + y * y - 1 - TE_D * x * x % P * y * y) % P == 0


def te_add_affine(p1, p2):
    x1, y1 = p1
    x2, y2 = p2
    dxy = TE_D * x1 % P * x2 % P * y1 % P * y2 % P
    x3 = (x1 * y2 + x2 * y1) % P * pow((1 + dxy) % P, -1, P) % P
    y3 = (y1 * y2 - TE_A * x1 % P * x2) % P * pow((1 - dxy) % P, -1, P) % P
    return (x3, y3)


def te_add_proj(p1, p2):
    """add-2008-bbjlp, projective (X:Y:Z)."""
    X1, Y1, Z1 = p1
    X2, Y2, Z2 = p2
    A = Z1 * Z2 % P
    B = A * A % P
    C = X1 * X2 % P
    D = Y1 * Y2 % P
    E = TE_D * C % P * D % P
    F = (B - E) % P
    G = (B + E) % P
    X3 = A * F % P * ((X1 + Y1) * (X2 + Y2) - C - D) % P
    Y3 = A * G % P * (D - TE_A * C) % P
    Z3 = F * G % P
    return (X3, Y3, Z3)


def te_to_affine(p):
    X, Y, Z = p
    zi = pow(Z, -1, P)
    return (X * zi % P, Y * zi % P)


def te_mul(k: int, pt):
    acc = (0, 1, 1)
    base = (pt[0], pt[1], 1)
    while k:
        if k & 1:
            acc = te_add_proj(acc, base)
        base = te_add_proj(base, base)
        k >>= 1
    return te_to_affine(acc)


def te_random_point(rng):
    """uniform element of the prime-order subgroup as k*G (avoids square roots)."""
    return te_mul(rng.randrange(1, TE_SUBGROUP_ORDER), TE_GEN)


def te_msm(points, scalars):
    acc = (0, 1, 1)
    for pt, k in zip(points, scalars):
        x, y = te_mul(k, pt)
        acc = te_add_proj(acc, (x, y, 1))
    return te_to_affine(acc)


assert te_on_curve(TE_GEN), "Bandersnatch generator is not on the curve"

# ---- BLS12-381 G1 (short Weierstrass, a = 0, b = 4) ----------------------------------------------------------
G1_B = 4
G1_GEN = (0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
          0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1)
G1_ORDER = FR_MODULUS


def g1_on_curve(pt) -> bool:
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - G1_B) % Q == 0


def g1_add(p1, p2):
    """affine, None = point at infinity."""
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if (y1 + y2) % Q == 0:
            return None
        lam = 3 * x1 * x1 % Q * pow(2 * y1, -1, Q) % Q
    else:
        lam = (y2 - y1) * pow((x2 - x1) % Q, -1, Q) % Q
    x3 = (lam * lam - x1 - x2) % Q
    y3 = (lam * (x1 - x3) - y1) % Q
    return (x3, y3)


def g1_neg(p):
    return None if p is None else (p[0], (-p[1]) % Q)


def g1_jac_add(p1, p2):
    """Jacobian (X, Y, Z), Z == 0 is infinity; complete via explicit doubling check."""
    X1, Y1, Z1 = p1
    X2, Y2, Z2 = p2
    if Z1 == 0:
        return p2
    if Z2 == 0:
        return p1
    Z1Z1 = Z1 * Z1 % Q
    Z2Z2 = Z2 * Z2 % Q
    U1 = X1 * Z2Z2 % Q
    U2 = X2 * Z1Z1 % Q
    S1 = Y1 * Z2 % Q * Z2Z2 % Q
    S2 = Y2 * Z1 % Q * Z1Z1 % Q
    if U1 == U2:
        if S1 != S2:
            return (1, 1, 0)
        return g1_jac_dbl(p1)
    H = (U2 - U1) % Q
    R = (S2 - S1) % Q
    HH = H * H % Q
    HHH = H * HH % Q
    V = U1 * HH % Q
    X3 = (R * R - HHH - 2 * V) % Q
    Y3 = (R * (V - X3) - S1 * HHH) % Q
    Z3 = Z1 * Z2 % Q * H % Q
    return (X3, Y3, Z3)


def g1_jac_dbl(p):
    X, Y, Z = p
    if Z == 0 or Y == 0:
        return (1, 1, 0)
    A = X * X % Q
    B = Y * Y % Q
    C = B * B % Q
    D = 2 * ((X + B) * (X + B) - A - C) % Q
    E = 3 * A % Q
    F = E * E % Q
    X3 = (F - 2 * D) % Q
    Y3 = (E * (D - X3) - 8 * C) % Q
    Z3 = 2 * Y * Z % Q
    return (X3, Y3, Z3)


def g1_from_jac(p):
    X, Y, Z = p
    if Z == 0:
        return None
    zi = pow(Z, -1, Q)
    zi2 = zi * zi % Q
    return (X * zi2 % Q, Y * zi2 % Q * zi % Q)


def g1_mul(k: int, pt):
    if pt is None:
        return None
    k %= G1_ORDER
    acc = (1, 1, 0)
    base = (pt[0], pt[1], 1)
    while k:
        if k & 1:
            acc = g1_jac_add(acc, base)
        base = g1_jac_dbl(base)
        k >>= 1
    return g1_from_jac(acc)


def g1_msm(points, scalars):
    acc = (1, 1, 0)
    for pt, k in zip(points, scalars):
        if pt is None:
            continue
        r = g1_mul(k, pt)
        if r is not None:
            acc = g1_jac_add(acc, (r[0], r[1], 1))
    return g1_from_jac(acc)


assert g1_on_curve(G1_GEN), "BLS12-381 G1 generator is not on the curve"
