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

__device__ __forceinline__ Fr fr_mul(const Fr& a, const Fr& b) {
    Fr r;
    fr_mul_asm(r.l, a.l, b.l);
    return r;
}
__device__ __forceinline__ Fr fr_sqr(const Fr& a) {
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
