// Kernel lab (measurement only, tools/kernel_lab.py): times variants of the dense round kernel -- register cap /
// resident blocks per SM, accumulators in registers or shared memory -- on the same resident tables, so that the
// configuration used by dense_sumcheck.cu is chosen from measurements on the B200 rather than guessed.
#include <algorithm>
#include "common.cuh"
#include "gates.cuh"
#include "dense_kernel.cuh"

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
        }
    }
#undef LAB
#undef LABP
    return ctx->fail(GKR_ERR_ARG, "unknown lab variant");
}
