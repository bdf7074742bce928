// Commitment MSM split by POINT RANGE over the GPUs of one box (KzgProvingKey::commit,
// src/commitments/kzg.rs:123-126, is a sum over independent points, so any partition of the SRS range works).
//
// One process per GPU.  The LEADER (rank 0) runs the prover; every large gkr_msm_g1 it issues is cut into `world` contiguous
// slices of the coefficient vector.  The workers hold the same SRS and sit in gkr_msm_team_serve.  Per MSM:
//   leader : D2H of the n scalars into a pinned POSIX shared-memory segment -> command (n, first) -> its own slice on its GPU
//   worker : H2D of its slice of the scalars -> local bucket MSM over bases [first + lo, first + hi) -> 96-byte affine result
//   leader : adds the `world` partial results on the host (group addition commutes: same point, same canonical limbs).
// Nothing but the scalars (32 B per point, once) and the partial results crosses a bus; there is no device collective because
// the result has to reach the host-side transcript anyway.  Every wait has a deadline: a dead peer is an error, not a hang.
//
// Second command: the c / d BUCKET SUMS of PushForwardState::new (pushforward.rs:398-429) split by x-range ("shard by
// x, not y, for balance").  The leader posts the digit / counter matrix (4 B per incidence) in the shared segment; GPU g
// accumulates the points g * 2^x / G <= x < (g + 1) * 2^x / G of every digit row into its own copy of the bucket array and
// returns it (192 B per bucket) through the second half of the segment; the leader adds the G arrays bucket by bucket.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <immintrin.h>
#include <atomic>
#include <cstring>
#include <string>
#include "common.cuh"
#include "host_g1.hpp"

#define GKR_TEAM_MAX_RANKS 16

int gkr_msm_g1_local(gkr_ctx* ctx, const gkr_srs* srs, uint64_t first, const Fr* d_scalars, uint64_t n, uint64_t* out_xy);  // msm.cu
int gkr_g1_bucket_sums_rows_range(gkr_ctx* ctx, const gkr_srs* srs, const uint32_t* d_idx, uint64_t n, uint32_t x_logsize, uint32_t clm,
                                  uint32_t group_log, uint32_t x_lo, uint32_t x_hi, bool allow_team, gkr_srs** out);  // msm.cu
int gkr_g1x_accumulate(gkr_ctx* ctx, void* d_acc, const void* d_part, uint64_t n);                                      // msm.cu
size_t gkr_g1x_bytes();
void* gkr_srs_device_ptr(gkr_srs* s);
extern "C" void gkr_srs_free(gkr_srs* s);
extern "C" uint64_t gkr_srs_len(const gkr_srs* s);

struct TeamShared {
    std::atomic<uint64_t> cmd_seq;
    uint64_t n, first;
    uint32_t quit, op;  // op 0: MSM slice; 1: bucket sums over an x-range
    uint32_t x_logsize, clm, group_log, pad_;
    std::atomic<uint64_t> done_seq[GKR_TEAM_MAX_RANKS];
    int32_t status[GKR_TEAM_MAX_RANKS];
    std::atomic<uint32_t> ready[GKR_TEAM_MAX_RANKS];  // worker r is inside gkr_msm_team_serve
    uint64_t result[GKR_TEAM_MAX_RANKS][12];
};
static const size_t TEAM_HDR = 8192;  // the scalars follow the header, page aligned
static_assert(sizeof(TeamShared) <= TEAM_HDR, "header does not fit");

struct gkr_msm_team {
    TeamShared* sh = nullptr;
    unsigned char* scalars = nullptr;
    size_t bytes = 0;
    int rank = 0, world = 1;
    uint64_t max_n = 0, seen = 0;
    std::string name;
    bool creator = false, registered = false;
};

extern "C" int gkr_msm_team_open(gkr_ctx* ctx, const char* name, int rank, int world, uint64_t max_n, int create, gkr_msm_team** out) {
    if (!ctx) return GKR_ERR_ARG;
    if (!name || !out || world < 1 || world > GKR_TEAM_MAX_RANKS || rank < 0 || rank >= world || max_n == 0)
        return ctx->fail(GKR_ERR_ARG, "gkr_msm_team_open: bad arguments");
    const size_t bytes = TEAM_HDR + (size_t)max_n * 64;  // first half: scalars / index matrix; second half: partial bucket sums
    int fd = shm_open(name, create ? (O_CREAT | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) return ctx->fail(GKR_ERR_ARG, "gkr_msm_team_open: shm_open failed (the leader creates the segment first)");
    if (create && ftruncate(fd, (off_t)bytes) != 0) {
        close(fd);
        return ctx->fail(GKR_ERR_ARG, "gkr_msm_team_open: ftruncate failed");
    }
    if (!create) {  // the segment must have its final size already
        struct stat st;
        if (fstat(fd, &st) != 0 || (size_t)st.st_size < bytes) {
            close(fd);
            return ctx->fail(GKR_ERR_ARG, "gkr_msm_team_open: segment not ready");
        }
    }
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return ctx->fail(GKR_ERR_ARG, "gkr_msm_team_open: mmap failed");
    gkr_msm_team* t = new gkr_msm_team();
    t->sh = (TeamShared*)p;
    t->scalars = (unsigned char*)p + TEAM_HDR;
    t->bytes = bytes;
    t->rank = rank;
    t->world = world;
    t->max_n = max_n;
    t->name = name;
    t->creator = create != 0;
    if (create) std::memset(p, 0, sizeof(TeamShared));
    cudaSetDevice(ctx->device);
    // pinned for DMA in this process (the other ranks register their own mapping); not fatal if it fails
    t->registered = cudaHostRegister(t->scalars, (size_t)max_n * 64, cudaHostRegisterPortable) == cudaSuccess;
    if (!t->registered) (void)cudaGetLastError();
    t->seen = t->sh->cmd_seq.load(std::memory_order_acquire);
    if (rank == 0) ctx->team = t;
    *out = t;
    return GKR_OK;
}

extern "C" void gkr_msm_team_close(gkr_ctx* ctx, gkr_msm_team* t) {
    if (!t) return;
    if (ctx && ctx->team == t) ctx->team = nullptr;
    if (t->registered) cudaHostUnregister(t->scalars);
    munmap(t->sh, t->bytes);
    if (t->creator) shm_unlink(t->name.c_str());
    delete t;
}

extern "C" int gkr_msm_team_world(const gkr_msm_team* t) { return t ? t->world : 0; }
// smallest MSM the leader shares with the team (default 2^18 points; tests lower it)
extern "C" void gkr_msm_team_set_min_n(gkr_ctx* ctx, uint64_t n) {
    if (ctx) ctx->team_min_n = n;
}

// leader: tell the workers to leave gkr_msm_team_serve
extern "C" void gkr_msm_team_quit(gkr_msm_team* t) {
    if (!t || t->rank != 0) return;
    t->sh->quit = 1;
    t->sh->cmd_seq.fetch_add(1, std::memory_order_release);
}

static double team_deadline_s() {  // how long the leader waits for a worker (default 30 s)
    const char* v = getenv("GKR_TEAM_TIMEOUT_S");
    const double d = v ? atof(v) : 0.0;
    return d > 0.0 ? d : 30.0;
}
static inline void team_slice(uint64_t n, int world, int rank, uint64_t* lo, uint64_t* hi) {
    const uint64_t chunk = (n + world - 1) / world;
    *lo = std::min<uint64_t>(n, chunk * rank);
    *hi = std::min<uint64_t>(n, *lo + chunk);
}

// worker loop: returns GKR_OK after the leader's quit, an error after `idle_timeout_s` without a command or on a CUDA failure
extern "C" int gkr_msm_team_serve(gkr_ctx* ctx, gkr_msm_team* t, const gkr_srs* srs, double idle_timeout_s) {
    if (!ctx) return GKR_ERR_ARG;
    if (!t || !srs || t->rank == 0) return ctx->fail(GKR_ERR_ARG, "gkr_msm_team_serve: workers only");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    Fr* d_sc = nullptr;
    uint64_t cap = 0;
    uint64_t idle0 = gkr_now_ns();
    t->seen = t->sh->cmd_seq.load(std::memory_order_acquire);
    t->sh->ready[t->rank].store(1, std::memory_order_release);
    for (;;) {
        const uint64_t c = t->sh->cmd_seq.load(std::memory_order_acquire);
        if (c == t->seen) {
            if ((gkr_now_ns() - idle0) * 1e-9 > idle_timeout_s) {
                if (d_sc) gkr_free_async(d_sc, ctx->stream);
                return ctx->fail(GKR_ERR_PROTOCOL, "gkr_msm_team_serve: no command from the leader (timeout)");
            }
            _mm_pause();
            continue;
        }
        t->seen = c;
        if (t->sh->quit) break;
        const uint64_t n = t->sh->n, first = t->sh->first;
        if (t->sh->op == 1) {  // bucket sums over this rank's x-range
            int rc = GKR_OK;
            const uint32_t xl = t->sh->x_logsize;
            uint64_t lo, hi;
            team_slice((uint64_t)1 << xl, t->world, t->rank, &lo, &hi);
            uint32_t* d_idx = nullptr;
            gkr_srs* part = nullptr;
            if (gkr_malloc_async(&d_idx, 4 * n, ctx->stream) != cudaSuccess) rc = ctx->fail(GKR_ERR_CUDA, "team worker: out of memory");
            if (rc == GKR_OK && cudaMemcpyAsync(d_idx, t->scalars, 4 * n, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
                rc = ctx->fail(GKR_ERR_CUDA, "team worker: H2D of the index matrix failed");
            if (rc == GKR_OK) rc = gkr_g1_bucket_sums_rows_range(ctx, srs, d_idx, n, xl, t->sh->clm, t->sh->group_log, (uint32_t)lo, (uint32_t)hi, false, &part);
            if (rc == GKR_OK) {
                const size_t nb = gkr_srs_len(part) * gkr_g1x_bytes();
                unsigned char* dst = t->scalars + (size_t)t->max_n * 32 + (size_t)(t->rank - 1) * nb;
                if ((size_t)(t->world - 1) * nb > (size_t)t->max_n * 32) rc = ctx->fail(GKR_ERR_UNSUPPORTED, "team worker: bucket array does not fit the segment");
                else if (cudaMemcpyAsync(dst, gkr_srs_device_ptr(part), nb, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                         cudaStreamSynchronize(ctx->stream) != cudaSuccess)
                    rc = ctx->fail(GKR_ERR_CUDA, "team worker: D2H of the bucket sums failed");
            }
            if (part) gkr_srs_free(part);
            if (d_idx) gkr_free_async(d_idx, ctx->stream);
            t->sh->status[t->rank] = rc;
            t->sh->done_seq[t->rank].store(c, std::memory_order_release);
            idle0 = gkr_now_ns();
            continue;
        }
        uint64_t lo, hi;
        team_slice(n, t->world, t->rank, &lo, &hi);
        int rc = GKR_OK;
        uint64_t out[12] = {0};
        if (hi > lo) {
            if (hi - lo > cap) {
                if (d_sc) gkr_free_async(d_sc, ctx->stream);
                cap = hi - lo;
                if (gkr_malloc_async(&d_sc, sizeof(Fr) * cap, ctx->stream) != cudaSuccess) rc = ctx->fail(GKR_ERR_CUDA, "team worker: out of memory");
            }
            if (rc == GKR_OK && cudaMemcpyAsync(d_sc, t->scalars + 32 * lo, 32 * (hi - lo), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
                rc = ctx->fail(GKR_ERR_CUDA, "team worker: H2D of the scalars failed");
            if (rc == GKR_OK) rc = gkr_msm_g1_local(ctx, srs, first + lo, d_sc, hi - lo, out);
        }
        std::memcpy(t->sh->result[t->rank], out, 96);
        t->sh->status[t->rank] = rc;
        t->sh->done_seq[t->rank].store(c, std::memory_order_release);
        idle0 = gkr_now_ns();
    }
    t->sh->ready[t->rank].store(0, std::memory_order_release);
    if (d_sc) gkr_free_async(d_sc, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    return GKR_OK;
}

// leader: wait until every worker sits in gkr_msm_team_serve (commands posted earlier would be missed)
extern "C" int gkr_msm_team_wait_ready(gkr_ctx* ctx, gkr_msm_team* t, double timeout_s) {
    if (!ctx) return GKR_ERR_ARG;
    if (!t || t->rank != 0) return ctx->fail(GKR_ERR_ARG, "gkr_msm_team_wait_ready: leader only");
    const uint64_t t0 = gkr_now_ns();
    for (int r = 1; r < t->world; r++)
        while (!t->sh->ready[r].load(std::memory_order_acquire)) {
            if ((gkr_now_ns() - t0) * 1e-9 > timeout_s) return ctx->fail(GKR_ERR_PROTOCOL, "gkr_msm_team_wait_ready: a worker did not show up");
            usleep(200);
        }
    return GKR_OK;
}

// leader side of one MSM (called from msm_g1_impl when a team is attached and the MSM is large enough)
int gkr_msm_team_run(gkr_ctx* ctx, const gkr_srs* srs, uint64_t first, const Fr* d_scalars, uint64_t n, uint64_t* out_xy) {
    gkr_msm_team* t = ctx->team;
    if (n > t->max_n || t->world == 1) return gkr_msm_g1_local(ctx, srs, first, d_scalars, n, out_xy);  // does not fit the shared buffer
    GKR_CUDA_OK(ctx, cudaMemcpyAsync(t->scalars, d_scalars, 32 * n, cudaMemcpyDeviceToHost, ctx->stream));
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    t->sh->n = n;
    t->sh->first = first;
    t->sh->op = 0;
    const uint64_t c = t->sh->cmd_seq.fetch_add(1, std::memory_order_acq_rel) + 1;
    uint64_t lo, hi;
    team_slice(n, t->world, 0, &lo, &hi);
    std::vector<gkr::G1XH> parts(t->world, gkr::g1h::inf());
    uint64_t mine[12] = {0};
    int rc = hi > lo ? gkr_msm_g1_local(ctx, srs, first + lo, d_scalars + lo, hi - lo, mine) : GKR_OK;
    auto put = [&](int r, const uint64_t* xy) {
        bool inf = true;
        for (int i = 0; i < 12; i++) inf = inf && xy[i] == 0;
        if (inf) return;
        std::memcpy(parts[r].X.v, xy, 48);
        std::memcpy(parts[r].Y.v, xy + 6, 48);
        parts[r].ZZ = gkr::fqh::ONE;
        parts[r].ZZZ = gkr::fqh::ONE;
    };
    put(0, mine);
    const uint64_t t0 = gkr_now_ns();
    for (int r = 1; r < t->world; r++) {  // always drain every worker, even after a local failure
        while (t->sh->done_seq[r].load(std::memory_order_acquire) < c) {
            if ((gkr_now_ns() - t0) * 1e-9 > team_deadline_s()) {
                ctx->team = nullptr;  // poisoned: a worker may still be reading the segment, never post into it again
                return ctx->fail(GKR_ERR_PROTOCOL, "team MSM: a worker did not answer in time (team detached; GKR_TEAM_TIMEOUT_S)");
            }
            _mm_pause();
        }
        if (t->sh->status[r] != GKR_OK && rc == GKR_OK) rc = ctx->fail(t->sh->status[r], "team MSM: a worker failed");
        put(r, t->sh->result[r]);
    }
    if (rc) return rc;
    gkr::g1h::horner_windows(parts.data(), 0, t->world, out_xy);  // plain sum + one inversion
    return GKR_OK;
}


// leader side of the c / d bucket sums (called from gkr_g1_bucket_sums_rows_range when a team is attached).  *handled = false:
// the matrix does not fit the shared segment or is too small to be worth it -- the caller runs it locally.
int gkr_team_bucket_sums(gkr_ctx* ctx, const gkr_srs* srs, const uint32_t* d_idx, uint64_t n, uint32_t x_logsize, uint32_t clm, uint32_t group_log,
                         gkr_srs** out, bool* handled) {
    gkr_msm_team* t = ctx->team;
    *handled = false;
    const uint64_t x_size = (uint64_t)1 << x_logsize;
    if (!t || t->world == 1 || 4 * n > (uint64_t)t->max_n * 32 || n < ctx->team_min_n || x_size < (uint64_t)t->world) return GKR_OK;
    const uint64_t rows = n / x_size, n_buckets = ((rows + ((uint64_t)1 << clm) - 1) >> clm) << group_log;
    const size_t nb = (size_t)n_buckets * gkr_g1x_bytes();
    if ((size_t)(t->world - 1) * nb > (size_t)t->max_n * 32) return GKR_OK;
    *handled = true;
    GKR_CUDA_OK(ctx, cudaMemcpyAsync(t->scalars, d_idx, 4 * n, cudaMemcpyDeviceToHost, ctx->stream));
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    t->sh->n = n;
    t->sh->first = 0;
    t->sh->op = 1;
    t->sh->x_logsize = x_logsize;
    t->sh->clm = clm;
    t->sh->group_log = group_log;
    const uint64_t c = t->sh->cmd_seq.fetch_add(1, std::memory_order_acq_rel) + 1;
    uint64_t lo, hi;
    team_slice(x_size, t->world, 0, &lo, &hi);
    gkr_srs* mine = nullptr;
    int rc = gkr_g1_bucket_sums_rows_range(ctx, srs, d_idx, n, x_logsize, clm, group_log, (uint32_t)lo, (uint32_t)hi, false, &mine);
    void* d_part = nullptr;
    if (rc == GKR_OK && gkr_malloc_async(&d_part, nb, ctx->stream) != cudaSuccess) rc = ctx->fail(GKR_ERR_CUDA, "team bucket sums: out of memory");
    const uint64_t t0 = gkr_now_ns();
    for (int r = 1; r < t->world; r++) {  // always drain every worker
        while (t->sh->done_seq[r].load(std::memory_order_acquire) < c) {
            if ((gkr_now_ns() - t0) * 1e-9 > team_deadline_s()) {
                ctx->team = nullptr;
                if (mine) gkr_srs_free(mine);
                if (d_part) gkr_free_async(d_part, ctx->stream);
                return ctx->fail(GKR_ERR_PROTOCOL, "team bucket sums: a worker did not answer in time (team detached)");
            }
            _mm_pause();
        }
        if (t->sh->status[r] != GKR_OK && rc == GKR_OK) rc = ctx->fail(t->sh->status[r], "team bucket sums: a worker failed");
        if (rc == GKR_OK) {
            const unsigned char* src = t->scalars + (size_t)t->max_n * 32 + (size_t)(r - 1) * nb;
            if (cudaMemcpyAsync(d_part, src, nb, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = ctx->fail(GKR_ERR_CUDA, "team bucket sums: H2D failed");
            if (rc == GKR_OK) rc = gkr_g1x_accumulate(ctx, gkr_srs_device_ptr(mine), d_part, n_buckets);
        }
    }
    if (d_part) gkr_free_async(d_part, ctx->stream);
    if (rc) {
        if (mine) gkr_srs_free(mine);
        return rc;
    }
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));  // the segment may be reused by the next command
    *out = mine;
    return GKR_OK;
}
