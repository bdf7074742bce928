// Carry-free reduced-radix Montgomery arithmetic ("RR"): field elements as N limbs of B < 32 bits in 32-bit registers,
// products accumulated in 64-bit COLUMNS by plain IMAD.WIDE.U32 (32x32+64 -> 64) with no carry flag anywhere.
//
// Why (measured on the B200, profiles/README.md r02, tools/kernel_lab.py): the carry-chained multiply-add the 32-bit-limb
// routines of field_gen.cuh are built from (mad.lo.cc / madc.hi.cc = IMAD.WIDE.U32.X) retires at 9.1e12 /s, HALF the rate
// of the carry-less IMAD.WIDE.U32 (1.85e13 /s = 64 per clock and SM).  With B-bit limbs a column of up to
// 2^(64 - 2B - slack) products cannot overflow, so every product goes through the full-rate form; the carries are
// propagated once per operation by shifts / masks on the ALU pipe, which the carry-chained kernels leave half idle.
//   Fr (255 bits): 9 limbs x 29 bits, Montgomery radix 2^261:  81 + 72 full-rate multiply-adds per product instead of
//                  112 half-rate ones;
//   Fq (381 bits): 14 limbs x 28 bits, Montgomery radix 2^392: 196 + 196 + 14 instead of 288 half-rate ones.
// Everything here is plain C++ on uint32_t / uint64_t (no inline PTX), __host__ __device__, so the same code is executed on
// the CPU against python big integers (tests/test_rr_field.py) before it runs on the device.
//
// Discipline (checked by the bounds in the comments, exercised at the extremes by the CPU tests):
//   * "tight" limbs are < 2^B; "loose" limbs are sums of a few tight ones.  rr_mul_acc needs
//     N * max(a_i) * max(b_j) + N * 2^(2B) + 2^36 < 2^64.
//   * values are only bounded by the container (N*B bits); a Montgomery product of a < A, b < Bv returns
//     < A*Bv / 2^(N*B) + p with tight limbs.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RR_FN __host__ __device__ __forceinline__
#else
#define RR_FN inline
#endif

// ---- configurations ----------------------------------------------------------------------------------------
// BLS12-381 Fr, r = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001 (== 1 mod 2^32, so -r^-1 == -1 mod 2^29)
struct RrFr {
    static constexpr int N = 9, B = 29;
    static constexpr bool P0_IS_ONE = true;
    static constexpr uint32_t NINV = 0x1fffffffu;  // -p^-1 mod 2^29
    RR_FN static constexpr uint32_t p(int i) {
        constexpr uint32_t v[9] = {0x1u, 0x1ffffff8u, 0x1f96ffbfu, 0x1b4805ffu, 0x1d80553bu, 0x0c0404d0u, 0x1520cce7u, 0x0a6533afu, 0x73eda7u};
        return v[i];
    }
    // 4p with every limb but the top one lifted by a borrow from its upper neighbour: a + SUBC - b has non-negative
    // limbs for tight b < 2^256 (top limb of b < 2^24 < SUBC[8])
    RR_FN static constexpr uint32_t subc(int i) {
        constexpr uint32_t v[9] = {0x20000004u, 0x3fffffdfu, 0x3e5bfefeu, 0x2d2017feu, 0x360154eeu, 0x30101342u, 0x3483339cu, 0x2994cebdu, 0x1cfb69cu};
        return v[i];
    }
    // 2^261 mod r (the Montgomery one of this radix) and 2^522 mod r
    RR_FN static constexpr uint32_t one(int i) {
        constexpr uint32_t v[9] = {0x1fffffbau, 0x22fu, 0x1cb61180u, 0x0a4e5c00u, 0x0ee8b1a2u, 0x16e6aedfu, 0x1907f8bbu, 0x0853ddf7u, 0x4d043fu};
        return v[i];
    }
    RR_FN static constexpr uint32_t r2(int i) {
        constexpr uint32_t v[9] = {0x0a71b3c0u, 0x1d32207eu, 0x1663d999u, 0x1c5abc93u, 0x03b58c44u, 0x0be37438u, 0x0829f771u, 0x1660139eu, 0x27fd91u};
        return v[i];
    }
};

template <class C>
struct RrElem {
    uint32_t l[C::N];
};
using F29 = RrElem<RrFr>;

template <class C>
RR_FN constexpr uint32_t rr_mask() { return (1u << C::B) - 1u; }

// ---- radix conversion ----------------------------------------------------------------------------------------
// W 32-bit words (little endian, any value) -> N tight limbs (the top limb takes whatever is left)
template <class C, int W>
RR_FN RrElem<C> rr_from_words(const uint32_t* w) {
    RrElem<C> r;
#pragma unroll
    for (int i = 0; i < C::N; i++) {
        const int o = C::B * i, k = o >> 5, s = o & 31;
        uint32_t v = 0;
        if (k < W) {
            v = w[k] >> s;
            if (s != 0 && s + C::B > 32 && k + 1 < W) v |= w[k + 1] << (32 - s);
        }
        r.l[i] = (i == C::N - 1) ? v : (v & rr_mask<C>());
    }
    return r;
}
// tight limbs of a value < 2^(32 W) -> W words
template <class C, int W>
RR_FN void rr_to_words(const RrElem<C>& a, uint32_t* w) {
#pragma unroll
    for (int k = 0; k < W; k++) {
        const int i0 = (32 * k) / C::B, r = 32 * k - C::B * i0;
        uint32_t v = a.l[i0] >> r;
        if (i0 + 1 < C::N) v |= a.l[i0 + 1] << (C::B - r);
        if (i0 + 2 < C::N && 2 * C::B - r < 32) v |= a.l[i0 + 2] << (2 * C::B - r);
        w[k] = v;
    }
}

// ---- limb-wise linear operations ----------------------------------------------------------------------------------------
template <class C>
RR_FN RrElem<C> rr_add(const RrElem<C>& a, const RrElem<C>& b) {  // limbs add up, no carries
    RrElem<C> r;
#pragma unroll
    for (int i = 0; i < C::N; i++) r.l[i] = a.l[i] + b.l[i];
    return r;
}
// a - b + 4p for TIGHT b < 2^(top-limb bound of subc); limbs < max(a_i) + 1.5 * 2^(B+1), never negative
template <class C>
RR_FN RrElem<C> rr_sub(const RrElem<C>& a, const RrElem<C>& b) {
    RrElem<C> r;
#pragma unroll
    for (int i = 0; i < C::N; i++) r.l[i] = a.l[i] + C::subc(i) - b.l[i];
    return r;
}
// propagate carries: loose (32-bit) limbs -> tight; the top limb keeps the rest
template <class C>
RR_FN RrElem<C> rr_norm(const RrElem<C>& a) {
    RrElem<C> r;
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < C::N; i++) {
        const uint32_t v = a.l[i] + c;
        if (i == C::N - 1) {
            r.l[i] = v;
        } else {
            r.l[i] = v & rr_mask<C>();
            c = v >> C::B;
        }
    }
    return r;
}

// ---- products ----------------------------------------------------------------------------------------
// t[i + j] += a_i * b_j for i < NA, j < NB   (NA * NB IMAD.WIDE.U32, no carries)
template <int NA, int NB>
RR_FN void rr_mul_acc(uint64_t* t, const uint32_t* a, const uint32_t* b) {
#pragma unroll
    for (int i = 0; i < NA; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) t[i + j] += (uint64_t)a[i] * b[j];
}
// a^2 into the columns: cross products once, against the doubled operand
template <int NA>
RR_FN void rr_sqr_acc(uint64_t* t, const uint32_t* a) {
    uint32_t a2[NA];
#pragma unroll
    for (int i = 0; i < NA; i++) a2[i] = a[i] << 1;
#pragma unroll
    for (int i = 0; i < NA; i++) {
        t[2 * i] += (uint64_t)a[i] * a[i];
#pragma unroll
        for (int j = i + 1; j < NA; j++) t[i + j] += (uint64_t)a[i] * a2[j];
    }
}

// ROUNDS word-serial Montgomery rounds on the columns t[0 .. ROUNDS + N]: afterwards the value / 2^(B * ROUNDS) sits in
// t[ROUNDS ..] (columns still 64-bit, not yet normalised).  t needs ROUNDS + N + 1 entries... the carry of round i goes
// to t[i + 1] <= t[ROUNDS], which exists.
template <class C, int ROUNDS>
RR_FN void rr_redc_rounds(uint64_t* t) {
#pragma unroll
    for (int i = 0; i < ROUNDS; i++) {
        uint32_t m;
        if (C::P0_IS_ONE) m = (0u - (uint32_t)t[i]) & rr_mask<C>();
        else m = ((uint32_t)t[i] * C::NINV) & rr_mask<C>();
        if (C::P0_IS_ONE) t[i] += m;
        else t[i] += (uint64_t)m * C::p(0);
#pragma unroll
        for (int j = 1; j < C::N; j++) t[i + j] += (uint64_t)m * C::p(j);
        t[i + 1] += t[i] >> C::B;  // the low B bits of t[i] are zero now
    }
}
// columns -> tight limbs (the top limb takes the rest; the caller's value bound keeps it < 2^32)
template <class C>
RR_FN RrElem<C> rr_cols_to_elem(const uint64_t* t) {
    RrElem<C> r;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < C::N; i++) {
        const uint64_t v = t[i] + c;
        if (i == C::N - 1) {
            r.l[i] = (uint32_t)v;
        } else {
            r.l[i] = (uint32_t)v & rr_mask<C>();
            c = v >> C::B;
        }
    }
    return r;
}

// Montgomery product a * b / 2^(N B) mod p; result tight, < a * b / 2^(N B) + p
template <class C>
RR_FN RrElem<C> rr_mul(const RrElem<C>& a, const RrElem<C>& b) {
    uint64_t t[2 * C::N];
#pragma unroll
    for (int i = 0; i < 2 * C::N; i++) t[i] = 0;
    rr_mul_acc<C::N, C::N>(t, a.l, b.l);
    rr_redc_rounds<C, C::N>(t);
    return rr_cols_to_elem<C>(t + C::N);
}
template <class C>
RR_FN RrElem<C> rr_sqr(const RrElem<C>& a) {
    uint64_t t[2 * C::N];
#pragma unroll
    for (int i = 0; i < 2 * C::N; i++) t[i] = 0;
    rr_sqr_acc<C::N>(t, a.l);
    rr_redc_rounds<C, C::N>(t);
    return rr_cols_to_elem<C>(t + C::N);
}

// (c + a * t) / 2^(B NT) mod p for a short plain multiplier t of NT tight limbs (the 128-bit Fiat-Shamir challenge of a
// sumcheck fold: NT = 5): NT * N + NT * (N - 1) multiply-adds instead of a full product.
// c tight, a limbs <= 2^(B+2) (the difference of rr_sub is fine).  Result tight, < (c + a t) / 2^(B NT) + p.
template <class C, int NT>
RR_FN RrElem<C> rr_fold_short(const RrElem<C>& c, const RrElem<C>& a, const uint32_t* t) {
    uint64_t col[C::N + NT + 1];
#pragma unroll
    for (int i = 0; i < C::N + NT + 1; i++) col[i] = i < C::N ? (uint64_t)c.l[i] : 0;
    rr_mul_acc<C::N, NT>(col, a.l, t);
    rr_redc_rounds<C, NT>(col);
    return rr_cols_to_elem<C>(col + NT);
}

// ---- Fr boundary: canonical 8 x u32 Montgomery-2^256 words <-> F29 ----------------------------------------------------------------------------------------
// The RR domain keeps the SAME residues as the canonical tables (x R256 mod r); a Montgomery product in radix 2^261 therefore
// returns x y R256 * 2^-5: the callers (dense29 kernels) are homogeneous and their host side multiplies the known power of
// 2^5 back, exactly like the 2^-128 of fr_fold128.
RR_FN F29 f29_load(const uint32_t* w8) { return rr_from_words<RrFr, 8>(w8); }

// subtract r while the value (8 words, < 2^256) is >= r; `times` subtractions at most
RR_FN void fr_words_canonical(uint32_t* w, int times) {
    const uint32_t P[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
    for (int k = 0; k < times; k++) {
        uint32_t s[8];
        uint64_t borrow = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint64_t d = (uint64_t)w[i] - P[i] - borrow;
            s[i] = (uint32_t)d;
            borrow = (d >> 32) & 1;
        }
        if (!borrow) {
#pragma unroll
            for (int i = 0; i < 8; i++) w[i] = s[i];
        }
    }
}
