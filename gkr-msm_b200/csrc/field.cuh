// Device-side field types for the GKR-MSM hot path (sm_100a).
//
//   Fr = BLS12-381 scalar field = Bandersnatch base field: every sumcheck / GKR table element.
//        Reference type: ark_bls12_381::Fr = Fp256<MontBackend<FrConfig,4>> (4 x u64 LE limbs,
//        Montgomery form, R = 2^256), e.g. src/utils.rs:32-49, src/cleanup/protocols/pippenger.rs:519.
//   Fq = BLS12-381 base field (G1 commitments): Fp384<MontBackend<FqConfig,6>>, src/commitments/kzg.rs.
//
// Memory layout == the reference's boundary layout (32-byte / 48-byte AoS elements of canonical
// Montgomery limbs).  sm_100a has 256-bit global loads/stores (LDG.E.256 / STG.E.256), so one Fr is
// exactly one vector access and one 32-byte DRAM sector: a warp reading 32 consecutive elements
// issues one fully coalesced 1 KiB request and no re-layout is needed at upload/download.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "field_gen.cuh"

struct __align__(32) Fr {
    uint32_t l[8];
};

struct __align__(16) Fq {
    uint32_t l[12];
};

// r = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
#define FR_P0 0x00000001u
#define FR_P1 0xffffffffu
#define FR_P2 0xfffe5bfeu
#define FR_P3 0x53bda402u
#define FR_P4 0x09a1d805u
#define FR_P5 0x3339d808u
#define FR_P6 0x299d7d48u
#define FR_P7 0x73eda753u

__device__ __forceinline__ Fr fr_zero() {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = 0;
    return r;
}

// Montgomery form of 1: R mod r = 0x1824b159acc5056f998c4fefecbc4ff55884b7fa0003480200000001fffffffe
__device__ __forceinline__ Fr fr_one() {
    Fr r;
    r.l[0] = 0xfffffffeu; r.l[1] = 0x00000001u; r.l[2] = 0x00034802u; r.l[3] = 0x5884b7fau;
    r.l[4] = 0xecbc4ff5u; r.l[5] = 0x998c4fefu; r.l[6] = 0xacc5056fu; r.l[7] = 0x1824b159u;
    return r;
}

// Montgomery form of the twisted-Edwards d of Bandersnatch == COEFF_D of src/utils.rs:34-37
// (u64 limbs 12167860994669987632, 4043113551995129031, 6052647550941614584, 3904213385886034240).
__device__ __forceinline__ Fr fr_te_d() {
    Fr r;
    r.l[0] = 0x47a2c730u; r.l[1] = 0xa8dced1bu; r.l[2] = 0xad3cccc7u; r.l[3] = 0x381c065au;
    r.l[4] = 0x188351f8u; r.l[5] = 0x53ff52e1u; r.l[6] = 0x990fe940u; r.l[7] = 0x362e8d63u;
    return r;
}

// GKR_COMPACT_FIELD: translation units whose kernels are latency-bound on small tables (one block, cold instruction
// cache) compile the multiplier as an out-of-line call: ~10x less code to fetch per launch.
#ifdef GKR_COMPACT_FIELD
#define GKR_MUL_INLINE __noinline__
#else
#define GKR_MUL_INLINE __forceinline__
#endif
__device__ GKR_MUL_INLINE Fr fr_mul(const Fr& a, const Fr& b) {
    Fr r;
    fr_mul_asm(r.l, a.l, b.l);
    return r;
}
__device__ GKR_MUL_INLINE Fr fr_sqr(const Fr& a) {
    Fr r;
    fr_sqr_asm(r.l, a.l);
    return r;
}
__device__ __forceinline__ Fr fr_add(const Fr& a, const Fr& b) {
    Fr r;
    fr_add_asm(r.l, a.l, b.l);
    return r;
}
__device__ __forceinline__ Fr fr_sub(const Fr& a, const Fr& b) {
    Fr r;
    fr_sub_asm(r.l, a.l, b.l);
    return r;
}
__device__ __forceinline__ Fr fr_dbl(const Fr& a) { return fr_add(a, a); }

// ---- lazy reduction --------------------------------------------------------------------------------
// FrWide: a 544-bit accumulator of UNREDUCED 512-bit products.  Montgomery reduction is linear, so a thread sums the
// last multiplication of every gate evaluation here (64 wide multiply-adds instead of 112) and reduces once at the
// end: sum_i REDC(x_i y_i) == REDC(sum_i x_i y_i) (mod r), bit-exact after canonicalisation.
struct FrWide {
    uint32_t l[17];
};
__device__ __forceinline__ void frw_zero(FrWide& w) {
#pragma unroll
    for (int i = 0; i < 17; i++) w.l[i] = 0;
}
__device__ __forceinline__ void frw_mac(FrWide& w, const Fr& a, const Fr& b) { fr_mac_wide_asm(w.l, a.l, b.l); }
// canonical Montgomery reduction of the accumulator: w = L + H 2^256 + T 2^512  ->  L R^-1 + H + T R  (mod r)
__device__ __forceinline__ Fr frw_reduce(const FrWide& w) {
    Fr lo, hi, one_plain, top, r2;
#pragma unroll
    for (int i = 0; i < 8; i++) { lo.l[i] = w.l[i]; hi.l[i] = w.l[8 + i]; one_plain.l[i] = 0; top.l[i] = 0; }
    one_plain.l[0] = 1;
    top.l[0] = w.l[16];
    // R^2 mod r = 0x0748d9d99f59ff1105d314967254398f2b6cedcb87925c23c999e990f3f29c6d
    r2.l[0] = 0xf3f29c6du; r2.l[1] = 0xc999e990u; r2.l[2] = 0x87925c23u; r2.l[3] = 0x2b6cedcbu;
    r2.l[4] = 0x7254398fu; r2.l[5] = 0x05d31496u; r2.l[6] = 0x9f59ff11u; r2.l[7] = 0x0748d9d9u;
    Fr a = fr_mul(lo, one_plain);  // L R^-1: the CIOS bound only needs L * 1 < r R
    Fr b;
    fr_reduce2_asm(b.l, hi.l);     // H < 2^256 < 3 r
    Fr c = fr_mul(top, r2);        // T * R^2 * R^-1
    return fr_add(fr_add(a, b), c);
}
// (c + t * a) * 2^-128 mod r for a 128-bit t (4 plain limbs): the sumcheck fold e0 + t (e1 - e0) by a Fiat-Shamir
// challenge of transcript.challenge(128), at half the multiplier work of a Montgomery product.  The result carries
// an extra factor 2^-128 which the host tracks (DenseSO::fast_folds).
__device__ __forceinline__ Fr fr_fold128(const Fr& c, const Fr& a, const uint32_t* t) {
    Fr r;
    fr_fold128_asm(r.l, c.l, a.l, t);
    return r;
}
__device__ __forceinline__ Fr fr_neg(const Fr& a) { return fr_sub(fr_zero(), a); }

// a = -5 on Bandersnatch: mul_by_a(x) = -(4x + x)   (src/utils.rs:40-43)
__device__ __forceinline__ Fr fr_mul_by_a(const Fr& x) {
    Fr t = fr_dbl(fr_dbl(x));
    return fr_neg(fr_add(t, x));
}
// y - a*x = y + 5x
__device__ __forceinline__ Fr fr_add_5x(const Fr& y, const Fr& x) {
    Fr t = fr_dbl(fr_dbl(x));
    return fr_add(y, fr_add(t, x));
}

__device__ __forceinline__ bool fr_is_zero(const Fr& a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.l[i];
    return o == 0;
}

__device__ __forceinline__ Fr fr_shfl_down(const Fr& a, int delta) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_down_sync(0xffffffffu, a.l[i], delta);
    return r;
}

__device__ __forceinline__ Fr fr_shfl_xor(const Fr& a, int mask) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_xor_sync(0xffffffffu, a.l[i], mask);
    return r;
}

// Streaming (read-once) 256-bit load: bypass L1 allocation so tables do not thrash it.
__device__ __forceinline__ Fr fr_ldg_stream(const Fr* p) {
    Fr r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
                 : "l"(p));
    return r;
}

// ---- Fq -------------------------------------------------------------------------------------
__device__ __forceinline__ Fq fq_mul(const Fq& a, const Fq& b) {
    Fq r;
    fq_mul_asm(r.l, a.l, b.l);
    return r;
}
__device__ __forceinline__ Fq fq_sqr(const Fq& a) {
    Fq r;
    fq_sqr_asm(r.l, a.l);
    return r;
}
__device__ __forceinline__ Fq fq_add(const Fq& a, const Fq& b) {
    Fq r;
    fq_add_asm(r.l, a.l, b.l);
    return r;
}
__device__ __forceinline__ Fq fq_sub(const Fq& a, const Fq& b) {
    Fq r;
    fq_sub_asm(r.l, a.l, b.l);
    return r;
}
__device__ __forceinline__ Fq fq_zero() {
    Fq r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = 0;
    return r;
}
__device__ __forceinline__ bool fq_is_zero(const Fq& a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) o |= a.l[i];
    return o == 0;
}
