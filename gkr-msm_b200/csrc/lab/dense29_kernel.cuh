// Reduced-radix flavour of the dense round kernel (dense_kernel.cuh) for Prod3Fn: the same round -- fold quads by the
// 128-bit challenge, write the half-size tables, evaluate the fresh pairs at the nodes 1..3 -- with every multiplication in
// the carry-free 9 x 29-bit form of rr_field.cuh (full-rate IMAD.WIDE.U32 instead of half-rate IMAD.WIDE.U32.X).
//   DenseSumcheckObjectSO::unipoly / bind_dense_poly   src/cleanup/protocols/sumcheck.rs:277-332, 160-163
// Tables stay in the reference's layout (32-byte canonical Montgomery words); the radix changes in registers only.  The folded
// tables are written canonical (< r), so every other kernel can read them; they carry 2^-145 per fold instead of the 2^-128 of
// fr_fold128, and the sums an extra 2^-10 (two products in radix 2^261) -- DenseSO tracks both (dense_sumcheck.cu).
#pragma once
#include "../common.cuh"
#include "../dense_kernel.cuh"
#include "dense29_item.cuh"

__device__ __forceinline__ F29 f29_ldg(const Fr* p) {
    const Fr v = *p;
    return f29_load(v.l);
}
// tight value < 3r -> canonical words, one 256-bit store
__device__ __forceinline__ void f29_stg_canonical(Fr* p, const F29& v) {
    Fr w, c;
    rr_to_words<RrFr, 8>(v, w.l);
    fr_reduce2_asm(c.l, w.l);
    *p = c;
}

// MODE 0: evaluate pairs; MODE 1: fold quads + evaluate.  Each thread normalises its accumulators every 4th item.
template <int MODE, int MINB>
__global__ void __launch_bounds__(GKR_REDUCE_THREADS, MINB) dense29_prod3_kernel(const __grid_constant__ DenseRoundArgs A) {
    __shared__ Fr smem[3 * (GKR_REDUCE_THREADS / 32)];
    Acc29 acc[3];
#pragma unroll
    for (int s = 0; s < 3; s++) acc29_zero(acc[s]);
    uint32_t t5[F29_FOLD_LIMBS];
    f29_challenge(A.t128, t5);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t it = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < A.n_items; i += stride, it++) {
        F29 lo[3], hi[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (MODE == 1) {
                const Fr* src = A.in[j] + 4 * i;
                const F29 e0 = f29_ldg(src), e1 = f29_ldg(src + 1), e2 = f29_ldg(src + 2), e3 = f29_ldg(src + 3);
                lo[j] = f29_fold(e0, e1, t5);
                hi[j] = f29_fold(e2, e3, t5);
                Fr* dst = A.out[j] + 2 * i;
                f29_stg_canonical(dst, lo[j]);
                f29_stg_canonical(dst + 1, hi[j]);
            } else {
                const Fr* src = A.in[j] + 2 * i;
                lo[j] = f29_ldg(src);
                hi[j] = f29_ldg(src + 1);
            }
        }
        prod3_nodes29(lo, hi, acc);
        if ((it & 3) == 3) {
#pragma unroll
            for (int s = 0; s < 3; s++) acc29_norm(acc[s]);
        }
    }
    Fr out[3];
#pragma unroll
    for (int s = 0; s < 3; s++) acc29_finish(acc[s], out[s].l);
    grid_reduce_to_host<3>(out, smem, A.o);
}
