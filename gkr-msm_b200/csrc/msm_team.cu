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
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <atomic>
#include <cstring>
#include <string>
#include "common.cuh"
#include "host_g1.hpp"

#define GKR_TEAM_MAX_RANKS 16

int gkr_msm_g1_local(gkr_ctx* ctx, const gkr_srs* srs, uint64_t first, const Fr* d_scalars, uint64_t n, uint64_t* out_xy);  // msm.cu

struct TeamShared {
    std::atomic<uint64_t> cmd_seq;
    uint64_t n, first;
    uint32_t quit, pad_;
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
    const size_t bytes = TEAM_HDR + (size_t)max_n * 32;
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
    t->registered = cudaHostRegister(t->scalars, (size_t)max_n * 32, cudaHostRegisterPortable) == cudaSuccess;
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
            continue;
        }
        t->seen = c;
        if (t->sh->quit) break;
        const uint64_t n = t->sh->n, first = t->sh->first;
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
            if ((gkr_now_ns() - t0) * 1e-9 > 30.0) return ctx->fail(GKR_ERR_PROTOCOL, "team MSM: a worker did not answer within 30 s");
        }
        if (t->sh->status[r] != GKR_OK && rc == GKR_OK) rc = ctx->fail(t->sh->status[r], "team MSM: a worker failed");
        put(r, t->sh->result[r]);
    }
    if (rc) return rc;
    gkr::g1h::horner_windows(parts.data(), 0, t->world, out_xy);  // plain sum + one inversion
    return GKR_OK;
}
