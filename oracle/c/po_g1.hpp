/* TEST INFRASTRUCTURE ONLY -- CPU oracle, never linked into or called from the product path.
 *
 * BLS12-381 G1 (y^2 = x^3 + 4 over Fq), the commitment group of the reference (ark-bls12-381 0.4.0 / ark-ec 0.4.2, not
 * vendored under /root/reference).  Group elements are unique, so textbook Jacobian formulas restate every result of
 *   KzgProvingKey::{mock_setup, commit, open}     src/commitments/kzg.rs:84-133  (-> liblasso::msm::VariableBaseMSM::msm, kzg.rs:13)
 *   VariableBaseMsmNonaffine::msm_nonaff          src/msm_nonaffine.rs:34-38
 *   bucket accumulation `+= point`                src/cleanup/protocols/pushforward/pushforward.rs:398-429
 * exactly.  The compressed wire format (48-byte big-endian x, flag bits compressed / infinity / y-is-larger) is the
 * zcash / IETF encoding ark-bls12-381 0.4.0 implements; pinned by the published generator encoding in
 * tests/test_oracle_pins.py.
 */
#pragma once
#include <algorithm>
#include <cmath>
#include "po_field.hpp"
#ifdef _OPENMP
#include <omp.h>
#endif

namespace po {

struct G1A {
    Fq x, y;
    bool inf;
    static G1A infinity() { return G1A{Fq::zero(), Fq::zero(), true}; }
};
struct G1J {  // Jacobian (X/Z^2, Y/Z^3); Z == 0 is the point at infinity
    Fq X, Y, Z;
    static G1J infinity() { return G1J{Fq::one(), Fq::one(), Fq::zero()}; }
    bool is_inf() const { return Z.is_zero(); }
    static G1J from_affine(const G1A& a) { return a.inf ? infinity() : G1J{a.x, a.y, Fq::one()}; }
};

static inline G1J g1_dbl(const G1J& p) {  // dbl-2009-l (a = 0)
    if (p.is_inf()) return p;
    Fq A = p.X.sqr(), B = p.Y.sqr(), C = B.sqr();
    Fq D = ((p.X + B).sqr() - A - C).dbl();
    Fq E = A.dbl() + A, F = E.sqr();
    G1J r;
    r.X = F - D.dbl();
    r.Y = E * (D - r.X) - C.dbl().dbl().dbl();
    r.Z = (p.Y * p.Z).dbl();
    return r;
}
static inline G1J g1_add(const G1J& p, const G1J& q) {  // add-2007-bl, complete by the explicit doubling / inverse checks
    if (p.is_inf()) return q;
    if (q.is_inf()) return p;
    Fq Z1Z1 = p.Z.sqr(), Z2Z2 = q.Z.sqr();
    Fq U1 = p.X * Z2Z2, U2 = q.X * Z1Z1;
    Fq S1 = p.Y * q.Z * Z2Z2, S2 = q.Y * p.Z * Z1Z1;
    if (U1 == U2) {
        if (S1 == S2) return g1_dbl(p);
        return G1J::infinity();
    }
    Fq H = U2 - U1, R = S2 - S1;
    Fq HH = H.sqr(), HHH = H * HH, V = U1 * HH;
    G1J r;
    r.X = R.sqr() - HHH - V.dbl();
    r.Y = R * (V - r.X) - S1 * HHH;
    r.Z = p.Z * q.Z * H;
    return r;
}
static inline G1J g1_madd(const G1J& p, const G1A& q) {  // mixed addition (Z2 = 1)
    if (q.inf) return p;
    if (p.is_inf()) return G1J::from_affine(q);
    Fq Z1Z1 = p.Z.sqr();
    Fq U2 = q.x * Z1Z1, S2 = q.y * p.Z * Z1Z1;
    if (p.X == U2) {
        if (p.Y == S2) return g1_dbl(p);
        return G1J::infinity();
    }
    Fq H = U2 - p.X, R = S2 - p.Y;
    Fq HH = H.sqr(), HHH = H * HH, V = p.X * HH;
    G1J r;
    r.X = R.sqr() - HHH - V.dbl();
    r.Y = R * (V - r.X) - p.Y * HHH;
    r.Z = p.Z * H;
    return r;
}
static inline G1A g1_neg(const G1A& a) { return G1A{a.x, -a.y, a.inf}; }
static inline G1J g1_neg(const G1J& a) { return G1J{a.X, -a.Y, a.Z}; }
static inline G1A g1_to_affine(const G1J& p) {
    if (p.is_inf()) return G1A::infinity();
    Fq zi = p.Z.inverse(), zi2 = zi.sqr();
    return G1A{p.X * zi2, p.Y * zi2 * zi, false};
}
/* ark_ec::CurveGroup::normalize_batch */
static inline void g1_batch_to_affine(const G1J* in, G1A* out, size_t n) {
    std::vector<Fq> z(n);
    for (size_t i = 0; i < n; i++) z[i] = in[i].Z;
    batch_inverse(z.data(), n);
    for (size_t i = 0; i < n; i++) {
        if (in[i].is_inf()) {
            out[i] = G1A::infinity();
            continue;
        }
        Fq zi2 = z[i].sqr();
        out[i] = G1A{in[i].X * zi2, in[i].Y * zi2 * z[i], false};
    }
}
static inline bool g1_eq(const G1J& a, const G1J& b) {
    if (a.is_inf() || b.is_inf()) return a.is_inf() && b.is_inf();
    Fq za = a.Z.sqr(), zb = b.Z.sqr();
    return a.X * zb == b.X * za && a.Y * zb * b.Z == b.Y * za * a.Z;
}
/* scalar multiplication by a field element (`comm * coeff`, pippenger.rs:194-201) */
static inline G1J g1_mul(const G1J& p, const Fr& k) {
    uint64_t e[4];
    k.to_raw(e);
    G1J acc = G1J::infinity();
    for (int i = 3; i >= 0; i--)
        for (int b = 63; b >= 0; b--) {
            acc = g1_dbl(acc);
            if ((e[i] >> b) & 1) acc = g1_add(acc, p);
        }
    return acc;
}

/* ark-bls12-381 0.4.0 compressed G1: 48-byte big-endian x, top bits of byte 0 = (compressed, infinity, y > (q-1)/2) */
static inline void g1_serialize(const G1A& p, uint8_t out[48]) {
    memset(out, 0, 48);
    if (p.inf) {
        out[0] = 0xC0;
        return;
    }
    uint64_t x[6], y[6], ny[6];
    p.x.to_raw(x);
    p.y.to_raw(y);
    (-p.y).to_raw(ny);
    for (int i = 0; i < 6; i++)
        for (int b = 0; b < 8; b++) out[47 - (8 * i + b)] = (uint8_t)(x[i] >> (8 * b));
    out[0] |= 0x80;
    bool larger = false;  // y > q - y
    for (int i = 5; i >= 0; i--) {
        if (y[i] != ny[i]) {
            larger = y[i] > ny[i];
            break;
        }
    }
    if (larger) out[0] |= 0x20;
}

/* ---- multi-scalar multiplication (result unique; algorithm: signed-digit bucket method, one slice of the points per
 * thread like a rayon `par_chunks` reduction) -------------------------------------------------------------------------- */
struct MsmDigits {
    int c, n_windows;
    std::vector<int32_t> d;  // [n][n_windows]
};
static inline int msm_window(size_t n) {  // ark-ec: ln(n) + 2 for n >= 32, else 3
    if (n < 32) return 3;
    return (int)std::min(18.0, std::floor(std::log((double)n)) + 2);
}
/* signed c-bit digits of the canonical scalar (make_digits, src/msm_nonaffine.rs:275-314) */
static inline void msm_signed_digits(const uint64_t raw[4], int c, int n_windows, int32_t* out) {
    int64_t carry = 0;
    const int64_t radix = (int64_t)1 << c, half = radix >> 1;
    for (int w = 0; w < n_windows; w++) {
        int bit = w * c;
        int64_t v = 0;
        if (bit < 256) {
            int limb = bit >> 6, off = bit & 63;
            uint64_t lo = raw[limb] >> off;
            if (off + c > 64 && limb + 1 < 4) lo |= raw[limb + 1] << (64 - off);
            v = (int64_t)(lo & (uint64_t)(radix - 1));
        }
        v += carry;
        carry = 0;
        if (v > half || (v == half && w + 1 < n_windows)) {  // keep the top window unsigned
            v -= radix;
            carry = 1;
        }
        out[w] = (int32_t)v;
    }
}

template <class Base, class AddFn>
static inline G1J msm_slice(const Base* bases, const Fr* scalars, size_t n, AddFn add_base) {
    if (n == 0) return G1J::infinity();
    const int c = msm_window(n);
    const int n_windows = (255 + c - 1) / c + 1;
    std::vector<int32_t> digs(n * (size_t)n_windows);
    for (size_t i = 0; i < n; i++) {
        uint64_t raw[4];
        scalars[i].to_raw(raw);
        msm_signed_digits(raw, c, n_windows, &digs[i * n_windows]);
    }
    const size_t nb = (size_t)1 << (c - 1);
    std::vector<G1J> buckets(nb);
    G1J total = G1J::infinity();
    for (int w = n_windows - 1; w >= 0; w--) {
        for (int k = 0; k < c; k++) total = g1_dbl(total);
        bool any = false;
        for (size_t b = 0; b < nb; b++) buckets[b] = G1J::infinity();
        for (size_t i = 0; i < n; i++) {
            int32_t d = digs[i * n_windows + w];
            if (d == 0) continue;
            any = true;
            if (d > 0) buckets[d - 1] = add_base(buckets[d - 1], bases[i], false);
            else buckets[-d - 1] = add_base(buckets[-d - 1], bases[i], true);
        }
        if (!any) continue;
        G1J running = G1J::infinity(), acc = G1J::infinity();
        for (size_t b = nb; b-- > 0;) {
            running = g1_add(running, buckets[b]);
            acc = g1_add(acc, running);
        }
        total = g1_add(total, acc);
    }
    return total;
}

template <class Base, class AddFn>
static inline G1J msm_generic(const Base* bases, const Fr* scalars, size_t n, AddFn add_base) {
    int threads = 1;
#ifdef _OPENMP
    threads = omp_get_max_threads();
#endif
    size_t n_slices = std::max<size_t>(1, std::min<size_t>((size_t)threads, n / 256));
    std::vector<G1J> part(n_slices);
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t s = 0; s < n_slices; s++) {
        size_t lo = n * s / n_slices, hi = n * (s + 1) / n_slices;
        part[s] = msm_slice(bases + lo, scalars + lo, hi - lo, add_base);
    }
    G1J total = G1J::infinity();
    for (size_t s = 0; s < n_slices; s++) total = g1_add(total, part[s]);
    return total;
}
/* KzgProvingKey::commit -> VariableBaseMSM::msm over affine SRS points (kzg.rs:123-126) */
static inline G1J g1_msm(const G1A* bases, const Fr* scalars, size_t n) {
    return msm_generic(bases, scalars, n, [](const G1J& acc, const G1A& b, bool neg) { return g1_madd(acc, neg ? g1_neg(b) : b); });
}
/* msm_nonaff: projective bases (src/msm_nonaffine.rs:34-38; pushforward.rs:598-604) */
static inline G1J g1_msm_proj(const G1J* bases, const Fr* scalars, size_t n) {
    return msm_generic(bases, scalars, n, [](const G1J& acc, const G1J& b, bool neg) { return g1_add(acc, neg ? g1_neg(b) : b); });
}

static inline G1J g1_msm_proj_serial(const G1J* bases, const Fr* scalars, size_t n) {
    return msm_slice(bases, scalars, n, [](const G1J& acc, const G1J& b, bool neg) { return g1_add(acc, neg ? g1_neg(b) : b); });
}

/* KzgProvingKey::mock_setup (kzg.rs:84-97): ptau_1[i] = g0 * tau^i, by a fixed-base byte-window table of g0 */
static inline std::vector<G1A> g1_powers_of_tau(const Fr& tau, const G1A& g0, size_t size) {
    // table[k][j] = (j + 1) * 2^(8k) * g0
    std::vector<G1A> table(32 * 255);
    {
        std::vector<G1J> tj(32 * 255);
        G1J base = G1J::from_affine(g0);
        for (int k = 0; k < 32; k++) {
            G1J acc = base;
            for (int j = 0; j < 255; j++) {
                tj[k * 255 + j] = acc;
                acc = g1_add(acc, base);
            }
            base = acc;  // 256 * base
        }
        g1_batch_to_affine(tj.data(), table.data(), tj.size());
    }
    std::vector<Fr> pows(size);
    Fr p = Fr::one();
    for (size_t i = 0; i < size; i++) {
        pows[i] = p;
        p = p * tau;
    }
    std::vector<G1A> out(size);
    const size_t CH = 4096;
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t lo = 0; lo < size; lo += CH) {
        size_t hi = std::min(size, lo + CH);
        std::vector<G1J> tmp(hi - lo);
        for (size_t i = lo; i < hi; i++) {
            uint64_t raw[4];
            pows[i].to_raw(raw);
            G1J acc = G1J::infinity();
            for (int k = 0; k < 32; k++) {
                unsigned byte = (unsigned)((raw[k / 8] >> (8 * (k % 8))) & 0xff);
                if (byte) acc = g1_madd(acc, table[k * 255 + byte - 1]);
            }
            tmp[i - lo] = acc;
        }
        g1_batch_to_affine(tmp.data(), out.data() + lo, hi - lo);
    }
    return out;
}
}  // namespace po
