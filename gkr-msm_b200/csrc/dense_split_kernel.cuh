// Node-split round kernel of DenseSumcheckObjectSO: one warp per evaluation node.
// (reference: src/cleanup/protocols/sumcheck.rs:160-163 bind_dense_poly + :277-332 unipoly, fused)
//
// The register kernel (dense_kernel.cuh) keeps 2 P table values and DEG 544-bit accumulators live per thread (166
// registers for Prod3 -> 12 warps / SM, 39 % issue-active).  Here a block is DEG warps working on the SAME 32 items:
// warp w folds the tables w, w + DEG, ... of those items (one quad in registers at a time), parks the fresh (lo, hi)
// pairs in shared memory, and after one block barrier evaluates the gate at node w + 1 only -- one accumulator and P
// values per thread.  Same arithmetic and the same sums as the register kernel (field addition is associative and
// every partial is canonical), at about half the registers, twice the resident warps and a third of the instruction
// footprint per warp.
// Shared-memory layout: value (table j, half h, item l) as two 16-byte planes indexed by l, so a warp's LDS.128 /
// STS.128 touch 512 contiguous bytes (conflict-free); double-buffered so one barrier per tile is enough.
#pragma once
#include "dense_kernel.cuh"

// grid stage: `vals[0..N)` are this block's sums (shared memory); same protocol as grid_reduce_to_host (the host folds
// <= GKR_HOST_FOLD_MAX_BLOCKS partials, otherwise the last block folds them: warp w folds accumulator w).
template <int N>
__device__ __forceinline__ void split_grid_reduce(Fr* vals, const RoundOut& o) {
    const unsigned int n_blocks = gridDim.x, bid = blockIdx.x;
    if (n_blocks <= GKR_HOST_FOLD_MAX_BLOCKS) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int s = 0; s < N; s++) o.part[(size_t)bid * N + s] = vals[s];
            __threadfence_system();
            bool last = true;
            if (n_blocks > 1) {
                unsigned int tk = atomicAdd(o.ticket, 1u);
                last = (tk == n_blocks - 1);
                if (last) *o.ticket = 0;
            }
            if (last) {
                __threadfence_system();
                *(volatile uint32_t*)o.flag = o.seq;
            }
        }
        return;
    }
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < N; s++) o.dev_part[(size_t)bid * N + s] = vals[s];
        __threadfence();
        unsigned int tk = atomicAdd(o.ticket, 1u);
        is_last = (tk == n_blocks - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    Fr v = fr_zero();
    for (unsigned int b = lane; b < n_blocks; b += 32) {
        const Fr* p = &o.dev_part[(size_t)b * N + w];
        Fr x;
        asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(x.l[0]), "=r"(x.l[1]), "=r"(x.l[2]), "=r"(x.l[3]), "=r"(x.l[4]), "=r"(x.l[5]), "=r"(x.l[6]), "=r"(x.l[7])
                     : "l"(p));
        v = fr_add(v, x);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fr_add(v, fr_shfl_down(v, off));
    if (lane == 0) o.part[w] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        *o.ticket = 0;
        __threadfence_system();
        *(volatile uint32_t*)o.flag = o.seq;
    }
}

__device__ __forceinline__ void sv_put(uint4* plane0, uint4* plane1, const Fr& v) {
    *plane0 = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    *plane1 = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ Fr sv_get(const uint4* plane0, const uint4* plane1) {
    const uint4 a = *plane0, b = *plane1;
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}

template <class SO, int MODE, bool FAST, int MINB = 5>
__global__ void __launch_bounds__(32 * SO::DEG, MINB) dense_round_split_kernel(const __grid_constant__ DenseRoundArgs A) {
    static_assert(MODE == 0 || MODE == 1, "split kernel: eval / fold+eval rounds only");
    constexpr int P = SO::P, DEG = SO::DEG;
    __shared__ Fr smem[DEG];              // one partial per warp (= per node)
    __shared__ uint4 sv[2][P][2][2][32];  // [buffer][table][half][16-byte plane][item]
    const uint32_t lane = threadIdx.x & 31, node = threadIdx.x >> 5;
    FrWide acc;
    frw_zero(acc);
    const uint64_t n_tiles = (A.n_items + 31) >> 5;
    uint32_t buf = 0;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
        const uint64_t i = (tile << 5) + lane;
        const bool live = i < A.n_items;
#pragma unroll
        for (int j0 = 0; j0 < P; j0 += DEG) {
            const int j = j0 + (int)node;
            if (j < P && live) {
                Fr lo, hi;
                if constexpr (MODE == 1) {
                    const Fr* src = A.in[j] + 4 * i;
                    const Fr e0 = src[0], e1 = src[1], e2 = src[2], e3 = src[3];
                    if (FAST) {
                        lo = fr_fold128(e0, fr_sub(e1, e0), A.t128);
                        hi = fr_fold128(e2, fr_sub(e3, e2), A.t128);
                    } else {
                        lo = fr_add(e0, fr_mul(A.t, fr_sub(e1, e0)));
                        hi = fr_add(e2, fr_mul(A.t, fr_sub(e3, e2)));
                    }
                    Fr* dst = A.out[j] + 2 * i;
                    dst[0] = lo;
                    dst[1] = hi;
                } else {
                    const Fr* src = A.in[j] + 2 * i;
                    lo = src[0];
                    hi = src[1];
                }
                sv_put(&sv[buf][j][0][0][lane], &sv[buf][j][0][1][lane], lo);
                sv_put(&sv[buf][j][1][0][lane], &sv[buf][j][1][1][lane], hi);
            }
        }
        __syncthreads();
        if (live) {
            Fr a[P];
#pragma unroll
            for (int j = 0; j < P; j++) {
                const Fr hi = sv_get(&sv[buf][j][1][0][lane], &sv[buf][j][1][1][lane]);
                // node w evaluates at 1 + w: args = p[2i+1] + w (p[2i+1] - p[2i])   (sumcheck.rs:295-313)
                Fr x = hi;
                if (node >= 1) {
                    const Fr lo = sv_get(&sv[buf][j][0][0][lane], &sv[buf][j][0][1][lane]);
                    const Fr d = fr_sub(hi, lo);
                    x = fr_add(x, d);
                    if (node >= 2) x = fr_add(x, d);
                    if (DEG > 3 && node >= 3) x = fr_add(x, d);
                }
                a[j] = x;
            }
            SO::mac(acc, a, A.consts);
        }
    }
    // every warp holds ONE accumulator (its node): reduce over the warp, then over the grid
    Fr mine = frw_reduce(acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mine = fr_add(mine, fr_shfl_down(mine, off));
    if (lane == 0) smem[node] = mine;
    __syncthreads();
    split_grid_reduce<DEG>(smem, A.o);
}
