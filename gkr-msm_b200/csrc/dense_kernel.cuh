// The round kernel of DenseSumcheckObjectSO (shared by dense_sumcheck.cu and the kernel lab, kernel_lab.cu).
#pragma once
#include "common.cuh"
#include "gates.cuh"

struct DenseRoundArgs {
    const Fr* in[GKR_MAX_POLYS];
    Fr* out[GKR_MAX_POLYS];
    uint64_t n_items;  // MODE 0/1: number of pairs evaluated; MODE 2: number of elements
    Fr t;              // challenge, Montgomery form (general fold)
    uint32_t t128[4];  // challenge as a plain 128-bit integer (FAST fold)
    GateConsts consts;
    RoundOut o;
    MailboxRef mbox;   // pre-launched small round (dense_small_kernel, FAST fold): t128 comes through the mailbox
};

// MODE 0: evaluate pairs (2i, 2i+1) of `in`                      (first round: nothing to fold yet)
// MODE 1: fold quads (4i..4i+3) of `in` into `out` (2i, 2i+1), then evaluate that fresh pair
// MODE 2: plain sum of f over all elements (claim_hint computation), one accumulator
//
// Integer-pipe economy (this kernel is bound by the 32x32->64 multiplier, not by HBM: DESIGN.md section 3):
//   * FAST folds: a Fiat-Shamir challenge of transcript.challenge(128) is a 128-bit integer, so
//     e0 + t (e1 - e0) is a 4x8-limb product plus a 4-round Montgomery reduction (fr_fold128, 56 wide
//     multiply-adds instead of 112).  The folded table then carries a factor 2^-128 per fast fold; every gate
//     on this path is homogeneous in the tables, so the round sums come out scaled by a known power of it
//     and the host multiplies it away (DenseSO::unscale_*): the round polynomials stay bit-exact.
//   * the last multiplication of every gate evaluation is accumulated UNREDUCED (FrWide, 64 instead of 112)
//     and each thread reduces its accumulators once.
//
// MINB: minimum resident blocks per SM (register cap).  ACC_SMEM: keep the NACC x 544-bit accumulators in shared
// memory (limb-major, conflict-free) instead of registers -- 51 registers less for a degree-3 gate.
template <int NACC, bool ACC_SMEM>
struct WideAccs {
    FrWide r[ACC_SMEM ? 1 : NACC];
    uint32_t* sm;
    __device__ __forceinline__ void init(uint32_t* smem_base) {
        sm = smem_base + threadIdx.x;
        if (ACC_SMEM) {
#pragma unroll
            for (int k = 0; k < NACC * 17; k++) sm[k * GKR_REDUCE_THREADS] = 0;
        } else {
#pragma unroll
            for (int s = 0; s < NACC; s++) frw_zero(r[s]);
        }
    }
    template <class SO>
    __device__ __forceinline__ void mac(int s, const Fr* a, const GateConsts& c) {
        if (ACC_SMEM) {
            FrWide w;
#pragma unroll
            for (int k = 0; k < 17; k++) w.l[k] = sm[(s * 17 + k) * GKR_REDUCE_THREADS];
            SO::mac(w, a, c);
#pragma unroll
            for (int k = 0; k < 17; k++) sm[(s * 17 + k) * GKR_REDUCE_THREADS] = w.l[k];
        } else {
            SO::mac(r[s], a, c);
        }
    }
    __device__ __forceinline__ Fr reduce(int s) {
        if (ACC_SMEM) {
            FrWide w;
#pragma unroll
            for (int k = 0; k < 17; k++) w.l[k] = sm[(s * 17 + k) * GKR_REDUCE_THREADS];
            return frw_reduce(w);
        }
        return frw_reduce(r[s]);
    }
};

// PF: software prefetch distance in grid-stride iterations -- every thread touches the cache lines of its iteration
// i + PF * stride with prefetch.global.L2 (no registers, no shared memory), so the demand loads of that iteration
// hit L2 instead of HBM and the long-scoreboard stalls of this low-occupancy kernel shrink.
template <class SO, int MODE, bool FAST, int MINB = 3, bool ACC_SMEM = false, int PF = 0>
__global__ void __launch_bounds__(GKR_REDUCE_THREADS, MINB) dense_round_kernel(const __grid_constant__ DenseRoundArgs A) {
    constexpr int P = SO::P;
    constexpr int NACC = (MODE == 2) ? 1 : SO::DEG;
    __shared__ Fr smem[NACC * (GKR_REDUCE_THREADS / 32)];
    __shared__ uint32_t acc_sm[ACC_SMEM ? NACC * 17 * GKR_REDUCE_THREADS : 1];
    WideAccs<NACC, ACC_SMEM> W;
    W.init(acc_sm);

    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < A.n_items; i += stride) {
        Fr a[P];
        if (PF > 0 && MODE != 2) {
            const uint64_t ip = i + (uint64_t)PF * stride;
            if (ip < A.n_items) {
#pragma unroll
                for (int j = 0; j < P; j++) {
                    const Fr* nxt = A.in[j] + (MODE == 1 ? 4 : 2) * ip;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
                }
            }
        }
        if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < P; j++) a[j] = A.in[j][i];
            W.template mac<SO>(0, a, A.consts);
        } else {
            Fr d[P];
#pragma unroll
            for (int j = 0; j < P; j++) {
                Fr lo, hi;
                if (MODE == 1) {
                    const Fr* src = A.in[j] + 4 * i;
                    Fr e0 = src[0], e1 = src[1], e2 = src[2], e3 = src[3];
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
                a[j] = hi;
                d[j] = fr_sub(hi, lo);
            }
            W.template mac<SO>(0, a, A.consts);
#pragma unroll
            for (int s = 1; s < SO::DEG; s++) {
#pragma unroll
                for (int j = 0; j < P; j++) a[j] = fr_add(a[j], d[j]);
                W.template mac<SO>(s, a, A.consts);
            }
        }
    }
    Fr acc[NACC];
#pragma unroll
    for (int s = 0; s < NACC; s++) acc[s] = W.reduce(s);
    grid_reduce_to_host<NACC>(acc, smem, A.o);
}


// ---- small rounds ----------------------------------------------------------------------------------------------------
// ~90 % of the rounds of a proof run on tables of at most a few thousand entries, where one launch of the kernel above is
// bounded by the chain of 2P folds + DEG gate evaluations every thread walks alone (and by streaming that much straight-line
// code through a cold instruction cache).  This kernel spreads one item over the block instead:
//   phase A: one thread per (item, table, half): fold (or load) ONE element, store it, park it in shared memory;
//   phase B: one thread per (item, evaluation node): one gate evaluation from the parked values (warp w = node w + 1).
// Same arithmetic, same results; the dependent chain per thread drops from 2P folds + DEG gates to ~ceil(2P/4) folds + 1 gate.
#define GKR_DENSE_SMALL_QB 32        // items per block iteration
#define GKR_DENSE_SMALL_MAX 4096     // items up to which launch_dense_round picks this kernel

template <class SO, int MODE, bool FAST>
__global__ void __launch_bounds__(GKR_REDUCE_THREADS) dense_small_kernel(const __grid_constant__ DenseRoundArgs A) {
    constexpr int P = SO::P, DEG = SO::DEG, QB = GKR_DENSE_SMALL_QB;
    static_assert(DEG * QB <= GKR_REDUCE_THREADS, "one warp per evaluation node");
    __shared__ Fr smem[DEG * (GKR_REDUCE_THREADS / 32)];
    __shared__ Fr sv[2 * P][QB];  // [2 * table + half][item]
    Fr mine = fr_zero();
    const uint32_t node = threadIdx.x >> 5, quad = threadIdx.x & 31;
    // pre-launched round: the challenge of the fold arrives through the mailbox (a cancelled launch publishes nothing)
    uint32_t t128[4] = {A.t128[0], A.t128[1], A.t128[2], A.t128[3]};
    if (MODE == 1 && FAST && A.mbox.box) {
        if (!gkr_mailbox_wait(A.mbox, t128)) return;
    }
    for (uint64_t base = (uint64_t)blockIdx.x * QB; base < A.n_items; base += (uint64_t)gridDim.x * QB) {
        for (uint32_t task = threadIdx.x; task < 2 * P * QB; task += GKR_REDUCE_THREADS) {
            const uint32_t k = task / QB, j = k >> 1, half = k & 1;
            const uint64_t i = base + (task % QB);
            if (i < A.n_items) {
                Fr v;
                if (MODE == 1) {
                    const Fr* src = A.in[j] + 4 * i + 2 * half;
                    const Fr e0 = src[0], e1 = src[1];
                    if (FAST) v = fr_fold128(e0, fr_sub(e1, e0), t128);
                    else v = fr_add(e0, fr_mul(A.t, fr_sub(e1, e0)));
                    A.out[j][2 * i + half] = v;
                } else {
                    v = A.in[j][2 * i + half];
                }
                sv[k][task % QB] = v;
            }
        }
        __syncthreads();
        if (node < DEG && base + quad < A.n_items) {
            Fr a[P];
#pragma unroll
            for (int j = 0; j < P; j++) {
                const Fr lo = sv[2 * j][quad], hi = sv[2 * j + 1][quad];
                const Fr d = fr_sub(hi, lo);
                Fr x = hi;  // node 0 evaluates at 1 (args = p[2i+1]), node s at 1 + s (args += difs), sumcheck.rs:295-313
                if (node >= 1) x = fr_add(x, d);
                if (node >= 2) x = fr_add(x, d);
                a[j] = x;
            }
            mine = fr_add(mine, SO::evalx(a, A.consts));
        }
        __syncthreads();
    }
    Fr acc[DEG];
#pragma unroll
    for (int s = 0; s < DEG; s++)
#pragma unroll
        for (int k = 0; k < 8; k++) acc[s].l[k] = (node == (uint32_t)s) ? mine.l[k] : 0u;
    grid_reduce_to_host<DEG>(acc, smem, A.o);
}


// ---- staged variant: tables flow HBM -> shared memory by cp.async (LDGSTS), one tile ahead ----------------------------
// The register kernel above runs at 12 warps / SM, and every warp stalls on its own 256-bit table loads (ncu: long
// scoreboard 1.6 cycles per issue at 39 % issue-active).  Here the loads never touch the register file: each WARP owns a
// private shared-memory tile per table (32 items x ELEMS elements = 4 KiB for quads), filled by 16-byte cp.async copies
// that are issued one whole iteration before the data is consumed -- table j of the next tile is requested as soon as
// table j of the current one has been read into registers, so the copy overlaps the remaining folds and all gate
// evaluations of the current tile.  No block-wide barrier: a tile is produced and consumed by the same warp
// (cp.async.wait_group + __syncwarp).  Shared-memory layout: 16-byte chunk c of a tile sits at position
// c ^ ((c >> 3) & 7) (the 128-byte XOR swizzle), which makes both the lane-contiguous cp.async writes and the
// "lane t reads ITS 128 / 64 bytes" LDS.128 reads bank-conflict free.
// Requires n_items % 32 == 0 (whole tiles); launch_dense_round falls back to the register kernel otherwise.
#define GKR_STAGED_WARPS (GKR_REDUCE_THREADS / 32)

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int ELEMS>
__device__ __forceinline__ void staged_issue_tile(uint32_t tile_smem, const Fr* gsrc, uint32_t lane) {
    // tile = 32 * ELEMS elements = 64 * ELEMS chunks of 16 B; lane copies chunks lane, lane + 32, ...
    const char* g = reinterpret_cast<const char*>(gsrc);
#pragma unroll
    for (int k = 0; k < 2 * ELEMS; k++) {
        const uint32_t c = lane + 32 * k;
        cp_async16(tile_smem + 16 * (c ^ ((c >> 3) & 7)), g + 16 * c);
    }
}
__device__ __forceinline__ Fr staged_read_elem(const uint4* tile, uint32_t elem) {
    const uint32_t c0 = 2 * elem, c1 = c0 + 1;
    const uint32_t sw = (c0 >> 3) & 7;
    const uint4 lo = tile[c0 ^ sw], hi = tile[c1 ^ sw];
    Fr r;
    r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
    r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
    return r;
}

template <class SO, int MODE, bool FAST, int MINB = 3>
__global__ void __launch_bounds__(GKR_REDUCE_THREADS, MINB) dense_round_staged_kernel(const __grid_constant__ DenseRoundArgs A) {
    static_assert(MODE == 0 || MODE == 1, "staged kernel: eval / fold+eval rounds only");
    constexpr int P = SO::P, NACC = SO::DEG;
    constexpr int ELEMS = MODE == 1 ? 4 : 2;              // table elements per item
    constexpr uint32_t TILE_BYTES = 32 * ELEMS * 32;      // one warp, one table
    extern __shared__ __align__(128) unsigned char staged_smem[];
    __shared__ Fr smem[NACC * (GKR_REDUCE_THREADS / 32)];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* my = staged_smem + (size_t)warp * P * TILE_BYTES;
    const uint32_t my_addr = (uint32_t)__cvta_generic_to_shared(my);

    WideAccs<NACC, false> W;
    W.init(nullptr);

    const uint64_t n_tiles = A.n_items >> 5;
    const uint64_t tile_stride = (uint64_t)gridDim.x * GKR_STAGED_WARPS;
    uint64_t tile = (uint64_t)blockIdx.x * GKR_STAGED_WARPS + warp;
    if (tile < n_tiles) {
#pragma unroll
        for (int j = 0; j < P; j++) {
            staged_issue_tile<ELEMS>(my_addr + j * TILE_BYTES, A.in[j] + (uint64_t)ELEMS * 32 * tile, lane);
            cp_async_commit();
        }
    }
    for (; tile < n_tiles; tile += tile_stride) {
        const uint64_t next = tile + tile_stride;
        const uint64_t i = (tile << 5) + lane;
        Fr a[P], d[P];
#pragma unroll
        for (int j = 0; j < P; j++) {
            cp_async_wait<P - 1>();
            __syncwarp();
            const uint4* t4 = reinterpret_cast<const uint4*>(my + j * TILE_BYTES);
            Fr e[ELEMS];
#pragma unroll
            for (int k = 0; k < ELEMS; k++) e[k] = staged_read_elem(t4, ELEMS * lane + k);
            __syncwarp();
            if (next < n_tiles) staged_issue_tile<ELEMS>(my_addr + j * TILE_BYTES, A.in[j] + (uint64_t)ELEMS * 32 * next, lane);
            cp_async_commit();
            Fr lo, hi;
            if constexpr (MODE == 1) {
                if (FAST) {
                    lo = fr_fold128(e[0], fr_sub(e[1], e[0]), A.t128);
                    hi = fr_fold128(e[ELEMS - 2], fr_sub(e[ELEMS - 1], e[ELEMS - 2]), A.t128);
                } else {
                    lo = fr_add(e[0], fr_mul(A.t, fr_sub(e[1], e[0])));
                    hi = fr_add(e[ELEMS - 2], fr_mul(A.t, fr_sub(e[ELEMS - 1], e[ELEMS - 2])));
                }
                Fr* dst = A.out[j] + 2 * i;
                dst[0] = lo;
                dst[1] = hi;
            } else {
                lo = e[0];
                hi = e[1];
            }
            a[j] = hi;
            d[j] = fr_sub(hi, lo);
        }
        W.template mac<SO>(0, a, A.consts);
#pragma unroll
        for (int s = 1; s < SO::DEG; s++) {
#pragma unroll
            for (int j = 0; j < P; j++) a[j] = fr_add(a[j], d[j]);
            W.template mac<SO>(s, a, A.consts);
        }
    }
    cp_async_wait<0>();
    Fr acc[NACC];
#pragma unroll
    for (int s = 0; s < NACC; s++) acc[s] = W.reduce(s);
    grid_reduce_to_host<NACC>(acc, smem, A.o);
}
