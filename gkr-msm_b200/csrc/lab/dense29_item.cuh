// Per-item arithmetic of the reduced-radix dense round kernels (dense29_kernel.cuh), as __host__ __device__ functions so
// that the CPU tests execute exactly the code the device runs (tests/test_rr_field.py through tests/rr_host.cpp).
//   DenseSumcheckObjectSO::unipoly   src/cleanup/protocols/sumcheck.rs:277-332  (args = p[2i+1], difs = p[2i+1]-p[2i], args += difs per node)
//   bind_dense_poly                  src/cleanup/protocols/sumcheck.rs:160-163  (p'[i] = p[2i] + t (p[2i+1] - p[2i]))
//
// Scaling.  Tables hold residues x R (R = 2^256, the reference's Montgomery form), possibly times a common factor sigma
// left by earlier folds.  In radix 2^29 a product returns x y R * 2^-5 and a short fold (c + a t) * 2^-145; a gate of
// homogeneous degree h therefore yields its true value times sigma^h * 2^(-5 (h - 1)), which the host multiplies away
// (DenseSO::unscale_sum, dense_sumcheck.cu) -- the round polynomials stay bit-exact.
#pragma once
#include "rr_field.cuh"

#define F29_FOLD_LIMBS 5  // 128-bit challenge = 5 limbs of 29 bits

// per-thread sum of tight products: limbs 0..8 collect up to 4 tight values between normalisations, limb 9 the overflow
struct Acc29 {
    uint32_t l[10];
};
RR_FN void acc29_zero(Acc29& s) {
#pragma unroll
    for (int i = 0; i < 10; i++) s.l[i] = 0;
}
RR_FN void acc29_add(Acc29& s, const F29& v) {
#pragma unroll
    for (int i = 0; i < 9; i++) s.l[i] += v.l[i];
}
RR_FN void acc29_norm(Acc29& s) {
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const uint32_t v = s.l[i] + c;
        s.l[i] = v & 0x1fffffffu;
        c = v >> 29;
    }
    s.l[9] += c;
}
// the accumulated value mod r as 8 canonical words: (V_low * 2^261 + l9 * 2^522) / 2^261
RR_FN void acc29_finish(const Acc29& s0, uint32_t* w8) {
    Acc29 s = s0;
    acc29_norm(s);
    uint64_t t[18];
#pragma unroll
    for (int i = 0; i < 18; i++) t[i] = 0;
    uint32_t one[9], r2[9];
#pragma unroll
    for (int i = 0; i < 9; i++) { one[i] = RrFr::one(i); r2[i] = RrFr::r2(i); }
    rr_mul_acc<9, 9>(t, s.l, one);
    // l9 < 2^32 is not tight: split it so that every factor stays below 2^29
    uint32_t top[2] = {s.l[9] & 0x1fffffffu, s.l[9] >> 29};
    rr_mul_acc<2, 9>(t, top, r2);
    rr_redc_rounds<RrFr, 9>(t);
    const F29 r = rr_cols_to_elem<RrFr>(t + 9);
    rr_to_words<RrFr, 8>(r, w8);
    fr_words_canonical(w8, 3);
}

// challenge words (128-bit plain integer) -> 5 tight limbs
RR_FN void f29_challenge(const uint32_t* t128, uint32_t* t5) {
    t5[0] = t128[0] & 0x1fffffffu;
    t5[1] = ((t128[0] >> 29) | (t128[1] << 3)) & 0x1fffffffu;
    t5[2] = ((t128[1] >> 26) | (t128[2] << 6)) & 0x1fffffffu;
    t5[3] = ((t128[2] >> 23) | (t128[3] << 9)) & 0x1fffffffu;
    t5[4] = t128[3] >> 20;
}

// fold one pair: (e0 + t (e1 - e0)) * 2^-145, tight, < 2^241 + r
RR_FN F29 f29_fold(const F29& e0, const F29& e1, const uint32_t* t5) {
    return rr_fold_short<RrFr, F29_FOLD_LIMBS>(e0, rr_sub<RrFr>(e1, e0), t5);
}

// Prod3Fn (pushforward.rs:266-281) at the nodes 1, 2, 3 of one pair per table: lo[j], hi[j] tight, < 2^256.
// acc[s] += prod_j (hi_j + s (hi_j - lo_j)), each product scaled by 2^-10.
RR_FN void prod3_nodes29(const F29* lo, const F29* hi, Acc29* acc) {
    F29 a0 = hi[0], a1 = hi[1], a2 = hi[2];
    // differences: tight, < 2^256 + 4r
    const F29 d0 = rr_norm<RrFr>(rr_sub<RrFr>(hi[0], lo[0]));
    const F29 d1 = rr_norm<RrFr>(rr_sub<RrFr>(hi[1], lo[1]));
    const F29 d2 = rr_norm<RrFr>(rr_sub<RrFr>(hi[2], lo[2]));
    acc29_add(acc[0], rr_mul<RrFr>(rr_mul<RrFr>(a0, a1), a2));
    // node 2: a0, a2 loose (< 2^30), a1 tight
    a0 = rr_add<RrFr>(a0, d0);
    a1 = rr_norm<RrFr>(rr_add<RrFr>(a1, d1));
    a2 = rr_add<RrFr>(a2, d2);
    acc29_add(acc[1], rr_mul<RrFr>(rr_mul<RrFr>(a0, a1), a2));
    // node 3: a0, a2 < 1.5 * 2^30 against tight partners (9 * 1.5 * 2^59 + 9 * 2^58 < 2^63)
    a0 = rr_add<RrFr>(a0, d0);
    a1 = rr_norm<RrFr>(rr_add<RrFr>(a1, d1));
    a2 = rr_add<RrFr>(a2, d2);
    acc29_add(acc[2], rr_mul<RrFr>(rr_mul<RrFr>(a0, a1), a2));
}
