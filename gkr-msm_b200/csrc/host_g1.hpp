// Host-side BLS12-381 Fq (6 x u64 Montgomery limbs) and G1 extended-Jacobian arithmetic for the O(windows) tail of an
// MSM: Horner over the <= 64 window sums (255 dependent doublings) and the normalisation to affine.  A latency chain of
// ~3000 dependent field multiplications runs ~20x faster on one CPU core (64-bit mulx) than on one GPU thread; the
// O(n) bucket work stays on the device.  Reference: the tail of liblasso VariableBaseMSM::msm (called from
// KzgProvingKey::commit, src/commitments/kzg.rs:123-126) and of msm_nonaff (src/msm_nonaffine.rs:144-161).
#pragma once
#include <cstdint>
#include <cstring>

namespace gkr {

struct FqH {
    uint64_t v[6];
};

namespace fqh {

typedef unsigned __int128 u128;
static const uint64_t MOD[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                                0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
static const uint64_t INV = 0x89f3fffcfffcfffdULL;  // -q^-1 mod 2^64
static const FqH ONE = {{0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL,
                         0x77ce585370525745ULL, 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL}};  // R mod q
static const FqH ZERO = {{0, 0, 0, 0, 0, 0}};

static inline bool is_zero(const FqH& a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5]) == 0; }
static inline bool eq(const FqH& a, const FqH& b) { return std::memcmp(a.v, b.v, 48) == 0; }
static inline bool geq_mod(const uint64_t a[6]) {
    for (int i = 5; i >= 0; i--) {
        if (a[i] > MOD[i]) return true;
        if (a[i] < MOD[i]) return false;
    }
    return true;
}
static inline void sub_mod_inplace(uint64_t a[6]) {
    u128 br = 0;
    for (int i = 0; i < 6; i++) {
        u128 d = (u128)a[i] - MOD[i] - br;
        a[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
}
static inline FqH add(const FqH& a, const FqH& b) {
    FqH r;
    u128 c = 0;
    for (int i = 0; i < 6; i++) {
        c += (u128)a.v[i] + b.v[i];
        r.v[i] = (uint64_t)c;
        c >>= 64;
    }
    if (geq_mod(r.v)) sub_mod_inplace(r.v);  // q < 2^381: no carry out of the top limb
    return r;
}
static inline FqH sub(const FqH& a, const FqH& b) {
    FqH r;
    u128 br = 0;
    for (int i = 0; i < 6; i++) {
        u128 d = (u128)a.v[i] - b.v[i] - br;
        r.v[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
    if (br) {
        u128 c = 0;
        for (int i = 0; i < 6; i++) {
            c += (u128)r.v[i] + MOD[i];
            r.v[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    return r;
}
static inline FqH dbl(const FqH& a) { return add(a, a); }
static inline FqH mul(const FqH& a, const FqH& b) {  // CIOS
    uint64_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 6; i++) {
        u128 c = 0;
        for (int j = 0; j < 6; j++) {
            c += (u128)a.v[j] * b.v[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[6];
        t[6] = (uint64_t)c;
        t[7] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * INV;
        c = (u128)m * MOD[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 6; j++) {
            c += (u128)m * MOD[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[6];
        t[5] = (uint64_t)c;
        t[6] = t[7] + (uint64_t)(c >> 64);
    }
    FqH r = {{t[0], t[1], t[2], t[3], t[4], t[5]}};
    if (t[6] || geq_mod(r.v)) sub_mod_inplace(r.v);
    return r;
}
static inline FqH sqr(const FqH& a) { return mul(a, a); }
static inline FqH inv(const FqH& a) {  // a^(q-2)
    uint64_t e[6];
    for (int i = 0; i < 6; i++) e[i] = MOD[i];
    e[0] -= 2;
    FqH r = ONE;
    for (int i = 380; i >= 0; i--) {
        r = sqr(r);
        if ((e[i >> 6] >> (i & 63)) & 1) r = mul(r, a);
    }
    return r;
}

}  // namespace fqh

// extended Jacobian (X, Y, ZZ, ZZZ): x = X/ZZ, y = Y/ZZZ; ZZ == 0 is the point at infinity (same layout as the device G1X)
struct G1XH {
    FqH X, Y, ZZ, ZZZ;
};

namespace g1h {
using namespace fqh;

static inline G1XH inf() { return G1XH{ZERO, ZERO, ZERO, ZERO}; }
static inline bool is_inf(const G1XH& p) { return is_zero(p.ZZ); }

static inline G1XH dbl(const G1XH& p) {  // dbl-2008-s-1, a = 0
    if (is_inf(p) || is_zero(p.Y)) return inf();
    FqH U = fqh::dbl(p.Y), V = sqr(U), W = mul(U, V), S = mul(p.X, V), XX = sqr(p.X);
    FqH M = add(fqh::dbl(XX), XX);
    G1XH r;
    r.X = sub(sqr(M), fqh::dbl(S));
    r.Y = sub(mul(M, sub(S, r.X)), mul(W, p.Y));
    r.ZZ = mul(V, p.ZZ);
    r.ZZZ = mul(W, p.ZZZ);
    return r;
}
static inline G1XH add(const G1XH& a, const G1XH& b) {  // add-2008-s
    if (is_inf(b)) return a;
    if (is_inf(a)) return b;
    FqH U1 = mul(a.X, b.ZZ), U2 = mul(b.X, a.ZZ), S1 = mul(a.Y, b.ZZZ), S2 = mul(b.Y, a.ZZZ);
    FqH Pp = sub(U2, U1), R = sub(S2, S1);
    if (is_zero(Pp)) return is_zero(R) ? dbl(a) : inf();
    FqH PP = sqr(Pp), PPP = mul(Pp, PP), Q = mul(U1, PP);
    G1XH r;
    r.X = sub(sub(sqr(R), PPP), fqh::dbl(Q));
    r.Y = sub(mul(R, sub(Q, r.X)), mul(S1, PPP));
    r.ZZ = mul(mul(a.ZZ, b.ZZ), PP);
    r.ZZZ = mul(mul(a.ZZZ, b.ZZZ), PPP);
    return r;
}
// affine (x, y) as 12 u64 (all zero = infinity); one inversion: 1/ZZ = (ZZ / ZZZ)^2 because ZZ^3 == ZZZ^2
static inline void to_affine(const G1XH& p, uint64_t out_xy[12]) {
    if (is_inf(p)) {
        std::memset(out_xy, 0, 96);
        return;
    }
    FqH iz3 = inv(p.ZZZ);
    FqH iz2 = sqr(mul(p.ZZ, iz3));
    FqH x = mul(p.X, iz2), y = mul(p.Y, iz3);
    std::memcpy(out_xy, x.v, 48);
    std::memcpy(out_xy + 6, y.v, 48);
}
// sum_w 2^(c w) * window_sums[w]  (Horner from the top window), normalised to affine
static inline void horner_windows(const G1XH* window_sums, int c, int n_windows, uint64_t out_xy[12]) {
    G1XH acc = inf();
    for (int w = n_windows - 1; w >= 0; w--) {
        if (!is_inf(acc))
            for (int k = 0; k < c; k++) acc = dbl(acc);
        acc = add(acc, window_sums[w]);
    }
    to_affine(acc, out_xy);
}

// ark-bls12-381 0.4.0 compressed G1 (zcash / IETF format, proof_transcript.rs:52-69): 48-byte big-endian x; top bits of
// byte 0 = (compressed, infinity, y is the lexicographically larger root).  xy: 12 u64 Montgomery limbs, all zero = infinity.
static inline void serialize_compressed(const uint64_t xy[12], uint8_t out[48]) {
    bool inf = true;
    for (int i = 0; i < 12; i++) inf = inf && xy[i] == 0;
    std::memset(out, 0, 48);
    if (inf) {
        out[0] = 0xC0;
        return;
    }
    FqH x, y, one_raw = {{1, 0, 0, 0, 0, 0}};
    std::memcpy(x.v, xy, 48);
    std::memcpy(y.v, xy + 6, 48);
    x = mul(x, one_raw);  // leave Montgomery form
    y = mul(y, one_raw);
    for (int i = 0; i < 6; i++)
        for (int b = 0; b < 8; b++) out[47 - (8 * i + b)] = (uint8_t)(x.v[i] >> (8 * b));
    out[0] |= 0x80;
    // y > (q - 1) / 2  <=>  2y > q - 1  <=>  2y >= q + 1 (q odd)  <=>  2y > q
    uint64_t t[7];
    uint64_t carry = 0;
    for (int i = 0; i < 6; i++) {
        t[i] = (y.v[i] << 1) | carry;
        carry = y.v[i] >> 63;
    }
    t[6] = carry;
    bool greater = t[6] != 0;
    if (!greater) {
        for (int i = 5; i >= 0; i--) {
            if (t[i] != MOD[i]) {
                greater = t[i] > MOD[i];
                break;
            }
        }
    }
    if (greater) out[0] |= 0x20;
}

}  // namespace g1h
}  // namespace gkr
