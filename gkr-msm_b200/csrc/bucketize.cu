// PushForwardState::new index bookkeeping on the device (src/cleanup/protocols/pushforward/pushforward.rs:351-396):
//   digits[y][x]  = (coef_x >> (y d)) & (2^d - 1)
//   counter[y][x] = rank of x inside its bucket (y, digit), in input order   (a STABLE counting sort per digit row)
//   lens[y][b]    = bucket sizes
//   pidx          = the bucket contents back to back, every bucket padded to even length with 0xffffffff -- the gather
//                   index of the bucket images (VecVecPolynomial::new pads odd rows, vecvec.rs:179-189)
// The host version (gkr_pushforward_bucketize, capi.cu) stays as the index-only host variant for digit widths above 13 bits (pure integer bookkeeping, no field arithmetic) and as the test oracle of
// this one; on the device only the scalars (32 B each) cross PCIe instead of three y_size x n index matrices.
//
// Stable ranks without sorting: the x range is cut into chunks of CHUNK consecutive scalars; one WARP walks one chunk in
// input order, 32 scalars per step -- __match_any_sync groups the lanes of equal digit, a lane's rank inside the step is the
// number of lower lanes in its group, and a per-warp shared-memory counter per digit carries the running count across steps.
// A column scan over the chunk histograms then gives every chunk its starting rank per digit.
#include <algorithm>
#include <vector>
#include "common.cuh"

#define BKT_CHUNK 2048u  // scalars per warp

__device__ __forceinline__ uint32_t bkt_digit(const uint64_t* c, uint32_t bit, uint32_t d_logsize) {
    const uint32_t limb = bit >> 6, sh = bit & 63;
    uint64_t v = c[limb] >> sh;
    if (sh && sh + d_logsize > 64 && limb + 1 < 4) v |= c[limb + 1] << (64 - sh);
    return (uint32_t)v & ((1u << d_logsize) - 1u);
}

// grid = (ceil(chunks / warps_per_block), y_size); dynamic shared memory = warps_per_block * 2^d counters
__global__ void bkt_rank_kernel(const uint64_t* coefs, uint64_t n, uint32_t d_logsize, uint32_t n_chunks, uint32_t* digits, uint32_t* counter,
                                uint32_t* hist /* [y][chunk][2^d] */) {
    extern __shared__ uint32_t bins_all[];
    const uint32_t nb = 1u << d_logsize, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, y = blockIdx.y;
    const uint32_t chunk = blockIdx.x * (blockDim.x >> 5) + warp;
    uint32_t* bins = bins_all + (size_t)warp * nb;
    for (uint32_t b = lane; b < nb; b += 32) bins[b] = 0;
    __syncwarp();
    if (chunk >= n_chunks) return;
    const uint64_t x0 = (uint64_t)chunk * BKT_CHUNK;
    uint32_t* dg = digits + (size_t)y * n;
    uint32_t* ct = counter + (size_t)y * n;
    for (uint32_t it = 0; it < BKT_CHUNK / 32; it++) {
        const uint64_t x = x0 + it * 32 + lane;
        const bool live = x < n;
        const unsigned act = __ballot_sync(0xffffffffu, live);
        if (act == 0) break;
        if (live) {
            const uint32_t d = bkt_digit(coefs + 4 * x, y * d_logsize, d_logsize);
            const unsigned same = __match_any_sync(act, d);
            const uint32_t before = __popc(same & ((1u << lane) - 1u));
            const uint32_t base = bins[d];
            __syncwarp(act);
            if (before == 0) bins[d] = base + __popc(same);  // the lowest lane of every group advances the counter
            __syncwarp(act);
            dg[x] = d;
            ct[x] = base + before;  // rank inside the chunk
        }
    }
    __syncwarp();
    uint32_t* h = hist + ((size_t)y * n_chunks + chunk) * nb;
    for (uint32_t b = lane; b < nb; b += 32) h[b] = bins[b];
}

// one thread per (y, digit): exclusive scan of the chunk histograms down the chunk axis; the total is the bucket size
__global__ void bkt_scan_kernel(uint32_t* hist, uint32_t n_chunks, uint32_t nb, uint32_t y_size, uint32_t* lens) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= y_size * nb) return;
    const uint32_t y = t / nb, b = t % nb;
    uint32_t run = 0;
    uint32_t* h = hist + (size_t)y * n_chunks * nb + b;
    for (uint32_t c = 0; c < n_chunks; c++) {
        const uint32_t v = h[(size_t)c * nb];
        h[(size_t)c * nb] = run;
        run += v;
    }
    lens[t] = run;
}

// counter += starting rank of the chunk; pidx[even_off[y][digit] + counter] = x; odd buckets get their pad entry
__global__ void bkt_finish_kernel(uint64_t n, uint32_t d_logsize, uint32_t n_chunks, uint32_t y_size, const uint32_t* digits, uint32_t* counter,
                                  const uint32_t* hist, const uint32_t* lens, const uint32_t* even_off, uint32_t* pidx) {
    const uint32_t nb = 1u << d_logsize;
    const uint64_t total = (uint64_t)y_size * n, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint32_t y = (uint32_t)(i / n);
        const uint64_t x = i - (uint64_t)y * n;
        const uint32_t d = digits[i], chunk = (uint32_t)(x / BKT_CHUNK);
        const uint32_t c = counter[i] + hist[((size_t)y * n_chunks + chunk) * nb + d];
        counter[i] = c;
        pidx[even_off[(size_t)y * nb + d] + c] = (uint32_t)x;
    }
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < (uint64_t)y_size * nb; r += stride)
        if (lens[r] & 1u) pidx[even_off[r] + lens[r]] = 0xffffffffu;
}

static gkr_u32buf* new_u32(gkr_ctx* ctx, uint64_t n, cudaError_t* e) {
    gkr_u32buf* b = new gkr_u32buf();
    b->ctx = ctx;
    b->n = n;
    *e = gkr_malloc_async(&b->d, sizeof(uint32_t) * std::max<uint64_t>(n, 1), ctx->stream);
    return b;
}

extern "C" int gkr_pushforward_bucketize_dev(gkr_ctx* ctx, const uint64_t* coefs, uint64_t n, uint32_t y_size, uint32_t d_logsize,
                                             gkr_u32buf** digits, gkr_u32buf** counter, gkr_u32buf** padded_order, uint32_t* lens) {
    if (!ctx) return GKR_ERR_ARG;
    if (!coefs || !digits || !counter || !padded_order || !lens || n == 0 || y_size == 0) return ctx->fail(GKR_ERR_ARG, "null argument");
    if (d_logsize == 0 || d_logsize > 13 || (uint64_t)y_size * d_logsize > 256 || n >= ((uint64_t)1 << 32))
        return ctx->fail(GKR_ERR_UNSUPPORTED, "device bucketize: 1 <= d_logsize <= 13, y_size * d_logsize <= 256, n < 2^32");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint32_t nb = 1u << d_logsize;
    const uint32_t n_chunks = (uint32_t)((n + BKT_CHUNK - 1) / BKT_CHUNK);
    const uint64_t m = (uint64_t)y_size * n, rows = (uint64_t)y_size * nb;
    uint64_t* d_coefs = nullptr;
    uint32_t *d_hist = nullptr, *d_lens = nullptr, *d_even = nullptr;
    cudaError_t e = cudaSuccess;
    gkr_u32buf *dg = new_u32(ctx, m, &e), *ct = nullptr, *po = nullptr;
    auto cleanup = [&](int rc) {
        if (d_coefs) gkr_free_async(d_coefs, st);
        if (d_hist) gkr_free_async(d_hist, st);
        if (d_lens) gkr_free_async(d_lens, st);
        if (d_even) gkr_free_async(d_even, st);
        if (rc) {
            gkr_u32_free(dg);
            gkr_u32_free(ct);
            gkr_u32_free(po);
        }
        return rc;
    };
    if (e == cudaSuccess) ct = new_u32(ctx, m, &e);
    if (e == cudaSuccess) e = gkr_malloc_async(&d_coefs, 32 * n, st);
    if (e == cudaSuccess) e = gkr_malloc_async(&d_hist, sizeof(uint32_t) * rows * n_chunks, st);
    if (e == cudaSuccess) e = gkr_malloc_async(&d_lens, sizeof(uint32_t) * rows, st);
    if (e == cudaSuccess) e = gkr_malloc_async(&d_even, sizeof(uint32_t) * rows, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_coefs, coefs, 32 * n, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cleanup(ctx->fail(GKR_ERR_CUDA, cudaGetErrorString(e)));
    {
        const uint32_t wpb = std::max<uint32_t>(1, std::min<uint32_t>(8, (48u << 10) / (nb * 4)));
        dim3 grid((n_chunks + wpb - 1) / wpb, y_size);
        bkt_rank_kernel<<<grid, wpb * 32, (size_t)wpb * nb * 4, st>>>(d_coefs, n, d_logsize, n_chunks, dg->d, ct->d, d_hist);
        bkt_scan_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(d_hist, n_chunks, nb, y_size, d_lens);
        ctx->launches += 2;
    }
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(lens, d_lens, sizeof(uint32_t) * rows, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cleanup(ctx->fail(GKR_ERR_CUDA, cudaGetErrorString(e)));
    // even-padded bucket offsets (host: y_size * 2^d entries)
    std::vector<uint32_t> even_off(rows);
    uint64_t total = 0;
    for (uint64_t r = 0; r < rows; r++) {
        even_off[r] = (uint32_t)total;
        total += (lens[r] + 1) & ~1u;
    }
    if (total >= ((uint64_t)1 << 32)) return cleanup(ctx->fail(GKR_ERR_UNSUPPORTED, "more than 2^32 bucket entries"));
    po = new_u32(ctx, total, &e);
    if (e != cudaSuccess) return cleanup(ctx->fail(GKR_ERR_CUDA, cudaGetErrorString(e)));
    {
        int rc = gkr_stage_upload(ctx, d_even, even_off.data(), sizeof(uint32_t) * rows);
        if (rc) return cleanup(rc);
    }
    {
        unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((m + 255) / 256, (uint64_t)ctx->num_sms * 8));
        bkt_finish_kernel<<<grid, 256, 0, st>>>(n, d_logsize, n_chunks, y_size, dg->d, ct->d, d_hist, d_lens, d_even, po->d);
        ctx->launches++;
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return cleanup(ctx->fail(GKR_ERR_CUDA, cudaGetErrorString(e)));
    *digits = dg;
    *counter = ct;
    *padded_order = po;
    return cleanup(GKR_OK);
}

extern "C" int gkr_u32_download(gkr_ctx* ctx, const gkr_u32buf* b, uint32_t* out) {
    if (!ctx || !b || !out) return GKR_ERR_ARG;
    if (b->n) GKR_CUDA_OK(ctx, cudaMemcpyAsync(out, b->d, sizeof(uint32_t) * b->n, cudaMemcpyDeviceToHost, ctx->stream));
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return GKR_OK;
}
extern "C" uint64_t gkr_u32_len(const gkr_u32buf* b) { return b ? b->n : 0; }
