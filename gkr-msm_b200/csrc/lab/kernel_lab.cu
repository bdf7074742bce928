// MEASUREMENT ONLY -- built into lib/libgkr_lab.so by `make lab`, never part of the product library.
// Kernel lab (tools/kernel_lab.py): times variants of the dense round kernel -- register cap /
// resident blocks per SM, accumulators in registers or shared memory -- on the same resident tables, so that the
// configuration used by dense_sumcheck.cu is chosen from measurements on the B200 rather than guessed.
#include <algorithm>
#include "../common.cuh"
#include "../gates.cuh"
#include "../dense_kernel.cuh"
#include "../dense_split_kernel.cuh"
#include "dense29_kernel.cuh"

template <int MODE, bool FAST, int MINB, bool ACC_SMEM, int PF = 0>
static int lab_run(gkr_ctx* ctx, DenseRoundArgs& a, int iters, float* ms, int* blocks_per_sm, int grid_mult) {
    auto kern = dense_round_kernel<SoProd3, MODE, FAST, MINB, ACC_SMEM, PF>;
    int b = 0;
    GKR_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, GKR_REDUCE_THREADS, 0));
    *blocks_per_sm = b;
    unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)ctx->num_sms * b * grid_mult, GKR_MAX_BLOCKS);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 2; w++) {
        a.o = ctx->round_out(0);
        kern<<<grid, GKR_REDUCE_THREADS, 0, ctx->stream>>>(a);
    }
    cudaEventRecord(e0, ctx->stream);
    for (int i = 0; i < iters; i++) {
        a.o = ctx->round_out(0);
        kern<<<grid, GKR_REDUCE_THREADS, 0, ctx->stream>>>(a);
        ctx->launches++;
    }
    cudaEventRecord(e1, ctx->stream);
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    GKR_CUDA_OK(ctx, cudaGetLastError());
    cudaEventElapsedTime(ms, e0, e1);
    *ms /= iters;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return GKR_OK;
}

template <int MODE, bool FAST, int MINB>
static int lab_run_staged(gkr_ctx* ctx, DenseRoundArgs& a, int iters, float* ms, int* blocks_per_sm, int grid_mult) {
    auto kern = dense_round_staged_kernel<SoProd3, MODE, FAST, MINB>;
    const size_t smem = (size_t)GKR_STAGED_WARPS * SoProd3::P * 32 * (MODE == 1 ? 4 : 2) * 32;
    GKR_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int b = 0;
    GKR_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, GKR_REDUCE_THREADS, smem));
    *blocks_per_sm = b;
    if (a.n_items % 32) return ctx->fail(GKR_ERR_ARG, "staged kernel: whole tiles only");
    unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)ctx->num_sms * b * grid_mult, GKR_MAX_BLOCKS);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 2; w++) {
        a.o = ctx->round_out(0);
        kern<<<grid, GKR_REDUCE_THREADS, smem, ctx->stream>>>(a);
    }
    cudaEventRecord(e0, ctx->stream);
    for (int i = 0; i < iters; i++) {
        a.o = ctx->round_out(0);
        kern<<<grid, GKR_REDUCE_THREADS, smem, ctx->stream>>>(a);
        ctx->launches++;
    }
    cudaEventRecord(e1, ctx->stream);
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    GKR_CUDA_OK(ctx, cudaGetLastError());
    cudaEventElapsedTime(ms, e0, e1);
    *ms /= iters;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return GKR_OK;
}

template <int MODE, bool FAST, int MINB>
static int lab_run_split(gkr_ctx* ctx, DenseRoundArgs& a, int iters, float* ms, int* blocks_per_sm, int grid_mult) {
    auto kern = dense_round_split_kernel<SoProd3, MODE, FAST, MINB>;
    const int threads = 32 * SoProd3::DEG;
    int b = 0;
    GKR_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, threads, 0));
    *blocks_per_sm = b;
    unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)ctx->num_sms * b * grid_mult, GKR_MAX_BLOCKS);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 2; w++) {
        a.o = ctx->round_out(0);
        kern<<<grid, threads, 0, ctx->stream>>>(a);
    }
    cudaEventRecord(e0, ctx->stream);
    for (int i = 0; i < iters; i++) {
        a.o = ctx->round_out(0);
        kern<<<grid, threads, 0, ctx->stream>>>(a);
        ctx->launches++;
    }
    cudaEventRecord(e1, ctx->stream);
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    GKR_CUDA_OK(ctx, cudaGetLastError());
    cudaEventElapsedTime(ms, e0, e1);
    *ms /= iters;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return GKR_OK;
}

template <int MODE, int MINB>
static int lab_run29(gkr_ctx* ctx, DenseRoundArgs& a, int iters, float* ms, int* blocks_per_sm, int grid_mult) {
    auto kern = dense29_prod3_kernel<MODE, MINB>;
    int b = 0;
    GKR_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, GKR_REDUCE_THREADS, 0));
    *blocks_per_sm = b;
    unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)ctx->num_sms * b * grid_mult, GKR_MAX_BLOCKS);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 2; w++) {
        a.o = ctx->round_out(0);
        kern<<<grid, GKR_REDUCE_THREADS, 0, ctx->stream>>>(a);
    }
    cudaEventRecord(e0, ctx->stream);
    for (int i = 0; i < iters; i++) {
        a.o = ctx->round_out(0);
        kern<<<grid, GKR_REDUCE_THREADS, 0, ctx->stream>>>(a);
        ctx->launches++;
    }
    cudaEventRecord(e1, ctx->stream);
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    GKR_CUDA_OK(ctx, cudaGetLastError());
    cudaEventElapsedTime(ms, e0, e1);
    *ms /= iters;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return GKR_OK;
}

// mode 0: eval only over pairs; mode 1: FAST fold + eval (writes `out` tables of n/2).  n = table length.
extern "C" int gkr_lab_dense_prod3(gkr_ctx* ctx, int variant, int mode, gkr_table* const* tables, gkr_table* const* out, uint64_t n,
                                   int iters, int grid_mult, float* ms, int* blocks_per_sm) {
    if (!ctx || !tables || !ms || !blocks_per_sm) return GKR_ERR_ARG;
    DenseRoundArgs a;
    for (int j = 0; j < 3; j++) {
        a.in[j] = tables[j]->d;
        a.out[j] = out ? out[j]->d : nullptr;
    }
    a.n_items = mode == 1 ? n / 4 : n / 2;
    a.t128[0] = 0x12345678u; a.t128[1] = 0x9abcdef0u; a.t128[2] = 0x0fedcba9u; a.t128[3] = 0x87654321u;
    a.t = fr_from_host(gkr::frh::ONE);
    for (int i = 0; i < GKR_MAX_GATE_CONSTS; i++) a.consts.g[i] = fr_from_host(gkr::frh::ONE);
#define LAB(M, F, B, S) return lab_run<M, F, B, S>(ctx, a, iters, ms, blocks_per_sm, grid_mult)
#define LABP(M, F, B, S, PFD) return lab_run<M, F, B, S, PFD>(ctx, a, iters, ms, blocks_per_sm, grid_mult)
    if (mode == 0) {
        switch (variant) {
            case 0: LAB(0, false, 3, false);
            case 1: LAB(0, false, 4, false);
            case 2: LAB(0, false, 4, true);
            case 3: LAB(0, false, 5, true);
            case 4: LAB(0, false, 2, false);
            case 5: LABP(0, false, 3, false, 1);
            case 6: LABP(0, false, 3, false, 2);
            case 7: LABP(0, false, 3, false, 4);
            case 8: return lab_run_staged<0, false, 3>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
            case 9: return lab_run_staged<0, false, 4>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
            case 10: return lab_run_split<0, false, 4>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
            case 11: return lab_run_split<0, false, 5>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
            case 12: return lab_run_split<0, false, 6>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
            case 13: return lab_run_split<0, false, 7>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
            case 14: return lab_run_split<0, false, 8>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
            case 15: return lab_run_split<0, false, 10>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
            case 20: return lab_run29<0, 3>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
            case 21: return lab_run29<0, 2>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
            case 22: return lab_run29<0, 4>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
        }
    } else {
        switch (variant) {
            case 0: LAB(1, true, 3, false);
            case 1: LAB(1, true, 4, false);
            case 2: LAB(1, true, 4, true);
            case 3: LAB(1, true, 5, true);
            case 4: LAB(1, true, 2, false);
            case 5: LABP(1, true, 3, false, 1);
            case 6: LABP(1, true, 3, false, 2);
            case 7: LABP(1, true, 3, false, 4);
            case 8: return lab_run_staged<1, true, 3>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
            case 9: return lab_run_staged<1, true, 4>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
            case 10: return lab_run_split<1, true, 4>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
            case 11: return lab_run_split<1, true, 5>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
            case 12: return lab_run_split<1, true, 6>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
            case 13: return lab_run_split<1, true, 7>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
            case 14: return lab_run_split<1, true, 8>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
            case 15: return lab_run_split<1, true, 10>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
            case 20: return lab_run29<1, 3>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
            case 21: return lab_run29<1, 2>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
            case 22: return lab_run29<1, 4>(ctx, a, iters, ms, blocks_per_sm, grid_mult);
        }
    }
#undef LAB
#undef LABP
    return ctx->fail(GKR_ERR_ARG, "unknown lab variant");
}


// ---- integer-pipe microbenchmark: ILP independent chains of dependent Montgomery multiplications -----------
template <int ILP>
__global__ void modmul_bench_kernel(Fr* out, int iters) {
    Fr x[ILP], y;
    for (int k = 0; k < ILP; k++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[k].l[i] = (threadIdx.x + 1) * (i + 3 + k) + blockIdx.x;
        x[k].l[7] &= 0x3fffffffu;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) y.l[i] = 0x9e3779b9u * (i + 1) + threadIdx.x;
    y.l[7] &= 0x3fffffffu;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) x[k] = fr_mul(x[k], y);
    }
    Fr acc = x[0];
    for (int k = 1; k < ILP; k++) acc = fr_add(acc, x[k]);
    if (acc.l[0] == 0x12345678u && acc.l[5] == 77u) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;  // keep the work alive
}

// Runs `iters` x ILP multiplications per thread on a grid filling the device; returns modmul/s.
extern "C" int gkr_bench_modmul(gkr_ctx* ctx, int ilp, int threads, int blocks_per_sm, int iters, double* modmul_per_s) {
    if (!ctx || !modmul_per_s) return GKR_ERR_ARG;
    Fr* out = nullptr;
    int grid = ctx->num_sms * blocks_per_sm;
    GKR_CUDA_OK(ctx, cudaMalloc(&out, sizeof(Fr) * (size_t)grid * threads));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(a, ctx->stream);
        if (ilp == 1) modmul_bench_kernel<1><<<grid, threads, 0, ctx->stream>>>(out, iters);
        else if (ilp == 2) modmul_bench_kernel<2><<<grid, threads, 0, ctx->stream>>>(out, iters);
        else modmul_bench_kernel<4><<<grid, threads, 0, ctx->stream>>>(out, iters);
        cudaEventRecord(b, ctx->stream);
        ctx->launches++;
    }
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(out);
    int eff_ilp = ilp == 1 ? 1 : (ilp == 2 ? 2 : 4);
    *modmul_per_s = (double)grid * threads * (double)iters * eff_ilp / (ms * 1e-3);
    return GKR_OK;
}

// ---- multiplier-pipe peak: ILP independent 32x32+64 -> 64 multiply-adds per thread (IMAD.WIDE.U32), no carries, no memory --
// The roofline denominator of every Montgomery kernel: how many wide multiply-adds per second the device can retire when
// nothing else limits it.  bench.py divides the wide multiply-adds a dense round actually executes (static SASS count x
// items / measured kernel time) by this number; by construction the fraction cannot exceed 1.
template <int ILP>
__global__ void imad_wide_peak_kernel(unsigned long long* out, int iters, uint32_t b0) {
    unsigned long long acc[ILP];
    const uint32_t a = threadIdx.x * 2654435761u + 12345u, b = b0 + blockIdx.x * 40503u;
#pragma unroll
    for (int k = 0; k < ILP; k++) acc[k] = 0x9e3779b97f4a7c15ull * (k + 1) + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int k = 0; k < ILP; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a), "r"(b));
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s ^= acc[k];
    if (s == 0x123456789abcdef0ull) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // keep the work alive
}

// The same probe with DISTINCT multiplier registers per instruction (8 a's x 8 b's, as a multi-precision product has): the
// probe above re-reads the same two registers in every instruction, which the operand reuse cache serves without touching
// the register file; a product of two 8..14-limb numbers cannot.  This is the rate a field multiplication can reach.
__global__ void imad_wide_rot_kernel(unsigned long long* out, int iters, uint32_t b0) {
    unsigned long long acc[16];
    uint32_t a[8], b[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        a[k] = threadIdx.x * 2654435761u + 12345u + 77u * k;
        b[k] = b0 + blockIdx.x * 40503u + 1013u * k;
    }
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = 0x9e3779b97f4a7c15ull * (k + 1) + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i + j]) : "r"(a[i]), "r"(b[j]));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) s ^= acc[k];
    if (s == 0x123456789abcdef0ull) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
extern "C" int gkr_lab_imad_wide_rot_peak(gkr_ctx* ctx, int threads, int blocks_per_sm, int iters, double* mads_per_s) {
    if (!ctx || !mads_per_s) return GKR_ERR_ARG;
    unsigned long long* out = nullptr;
    int grid = ctx->num_sms * blocks_per_sm;
    GKR_CUDA_OK(ctx, cudaMalloc(&out, sizeof(unsigned long long) * (size_t)grid * threads));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a, ctx->stream);
        imad_wide_rot_kernel<<<grid, threads, 0, ctx->stream>>>(out, iters, 77u);
        cudaEventRecord(b, ctx->stream);
        ctx->launches++;
        GKR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(out);
    *mads_per_s = (double)grid * threads * (double)iters * 64.0 / (best * 1e-3);
    return GKR_OK;
}

extern "C" int gkr_lab_imad_wide_peak(gkr_ctx* ctx, int ilp, int threads, int blocks_per_sm, int iters, double* mads_per_s) {
    if (!ctx || !mads_per_s) return GKR_ERR_ARG;
    unsigned long long* out = nullptr;
    int grid = ctx->num_sms * blocks_per_sm;
    GKR_CUDA_OK(ctx, cudaMalloc(&out, sizeof(unsigned long long) * (size_t)grid * threads));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a, ctx->stream);
        if (ilp <= 4) imad_wide_peak_kernel<4><<<grid, threads, 0, ctx->stream>>>(out, iters, 77u);
        else if (ilp <= 8) imad_wide_peak_kernel<8><<<grid, threads, 0, ctx->stream>>>(out, iters, 77u);
        else imad_wide_peak_kernel<16><<<grid, threads, 0, ctx->stream>>>(out, iters, 77u);
        cudaEventRecord(b, ctx->stream);
        ctx->launches++;
        GKR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(out);
    const int eff = ilp <= 4 ? 4 : (ilp <= 8 ? 8 : 16);
    *mads_per_s = (double)grid * threads * (double)iters * 4.0 * eff / (best * 1e-3);
    return GKR_OK;
}

// ---- the same probe in the form the field routines use: mad.lo.cc / madc.hi.cc pairs chained through the carry flag (ptxas
// fuses each pair into one IMAD.WIDE.U32.X).  ILP independent 8-pair chains per thread.  Answers whether the carry-chained form
// retires at the rate of the plain IMAD.WIDE.U32 above.
template <int ILP>
__global__ void imad_wide_x_peak_kernel(unsigned int* out, int iters, uint32_t b0) {
    uint32_t acc[ILP][17];
    const uint32_t a = threadIdx.x * 2654435761u + 12345u, b = b0 + blockIdx.x * 40503u;
#pragma unroll
    for (int k = 0; k < ILP; k++)
#pragma unroll
        for (int i = 0; i < 17; i++) acc[k][i] = (k + 1) * 0x9e3779b9u + i + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) {
            asm volatile(
                "mad.lo.cc.u32 %0, %17, %18, %0;\n\t"
                "madc.hi.cc.u32 %1, %17, %18, %1;\n\t"
                "madc.lo.cc.u32 %2, %17, %18, %2;\n\t"
                "madc.hi.cc.u32 %3, %17, %18, %3;\n\t"
                "madc.lo.cc.u32 %4, %17, %18, %4;\n\t"
                "madc.hi.cc.u32 %5, %17, %18, %5;\n\t"
                "madc.lo.cc.u32 %6, %17, %18, %6;\n\t"
                "madc.hi.cc.u32 %7, %17, %18, %7;\n\t"
                "madc.lo.cc.u32 %8, %17, %18, %8;\n\t"
                "madc.hi.cc.u32 %9, %17, %18, %9;\n\t"
                "madc.lo.cc.u32 %10, %17, %18, %10;\n\t"
                "madc.hi.cc.u32 %11, %17, %18, %11;\n\t"
                "madc.lo.cc.u32 %12, %17, %18, %12;\n\t"
                "madc.hi.cc.u32 %13, %17, %18, %13;\n\t"
                "madc.lo.cc.u32 %14, %17, %18, %14;\n\t"
                "madc.hi.cc.u32 %15, %17, %18, %15;\n\t"
                "addc.u32 %16, %16, 0;\n\t"
                : "+r"(acc[k][0]), "+r"(acc[k][1]), "+r"(acc[k][2]), "+r"(acc[k][3]), "+r"(acc[k][4]), "+r"(acc[k][5]), "+r"(acc[k][6]), "+r"(acc[k][7]),
                  "+r"(acc[k][8]), "+r"(acc[k][9]), "+r"(acc[k][10]), "+r"(acc[k][11]), "+r"(acc[k][12]), "+r"(acc[k][13]), "+r"(acc[k][14]),
                  "+r"(acc[k][15]), "+r"(acc[k][16])
                : "r"(a), "r"(b));
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++)
#pragma unroll
        for (int i = 0; i < 17; i++) s ^= acc[k][i];
    if (s == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int gkr_lab_imad_wide_x_peak(gkr_ctx* ctx, int ilp, int threads, int blocks_per_sm, int iters, double* mads_per_s) {
    if (!ctx || !mads_per_s) return GKR_ERR_ARG;
    unsigned int* out = nullptr;
    int grid = ctx->num_sms * blocks_per_sm;
    GKR_CUDA_OK(ctx, cudaMalloc(&out, sizeof(unsigned int) * (size_t)grid * threads));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a, ctx->stream);
        if (ilp <= 1) imad_wide_x_peak_kernel<1><<<grid, threads, 0, ctx->stream>>>(out, iters, 77u);
        else if (ilp <= 2) imad_wide_x_peak_kernel<2><<<grid, threads, 0, ctx->stream>>>(out, iters, 77u);
        else imad_wide_x_peak_kernel<4><<<grid, threads, 0, ctx->stream>>>(out, iters, 77u);
        cudaEventRecord(b, ctx->stream);
        ctx->launches++;
        GKR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(out);
    const int eff = ilp <= 1 ? 1 : (ilp <= 2 ? 2 : 4);
    *mads_per_s = (double)grid * threads * (double)iters * 8.0 * eff / (best * 1e-3);  // 8 wide multiply-adds per chain
    return GKR_OK;
}
