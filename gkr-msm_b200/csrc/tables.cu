// Context, device tables and table builders (eq tables, synthetic inputs).
//   eq tables: eq_poly_sequence_from_multiplier / eq_poly_sequence_last  src/utils.rs:222-262,
//              EqPoly::evals src/cleanup/protocols/verifier_polys.rs:31-33
#include <cstdlib>
#include <algorithm>
#include <mutex>
#include "common.cuh"

static std::mutex g_slot_mutex;
struct SlotPool {
    std::vector<int> free_list;
};
static std::vector<std::pair<gkr_ctx*, SlotPool*>> g_pools;

static SlotPool* pool_of(gkr_ctx* ctx) {
    for (auto& p : g_pools)
        if (p.first == ctx) return p.second;
    return nullptr;
}

int gkr_result_slot_acquire(gkr_ctx* ctx) {
    std::lock_guard<std::mutex> lk(g_slot_mutex);
    SlotPool* p = pool_of(ctx);
    if (!p || p->free_list.empty()) return -1;
    int s = p->free_list.back();
    p->free_list.pop_back();
    return s;
}

void gkr_result_slot_release(gkr_ctx* ctx, int slot) {
    std::lock_guard<std::mutex> lk(g_slot_mutex);
    SlotPool* p = pool_of(ctx);
    if (p) p->free_list.push_back(slot);
}

// Spin on the slot's flag (written by the last block into mapped host memory) and fold the per-block partials.
// Every ~64k spins the stream is queried so that a faulted or vanished kernel turns into an error, never a hang.
// ---- large-block cache in front of the stream-ordered pool (see common.cuh) ---------------------------------------------
#include <map>
#include <unordered_map>
namespace {
struct BigCache {
    std::multimap<size_t, void*> free_blocks;
    std::unordered_map<void*, size_t> live;  // large blocks handed out
};
std::mutex g_big_mutex;
std::map<cudaStream_t, BigCache> g_big;
// bytes of large blocks handed out (live) and their high-water mark: the footprint of the tables of a proof without the
// idle blocks the cache keeps for reuse (GKR_TRACE prints both per span, protocol.cu)
uint64_t g_big_live = 0, g_big_peak = 0;
inline void big_account(int64_t delta) {
    g_big_live = (uint64_t)((int64_t)g_big_live + delta);
    if (g_big_live > g_big_peak) g_big_peak = g_big_live;
}
// Peer pool (gkr_ctx_peer_pool): when the home GPU is full, large blocks are placed in the HBM of the other GPUs of the box and
// the kernels -- which all run on the home GPU -- reach them through NVLink peer access.  Nothing else changes: same kernels,
// same pointers, same proof; tables that spill over run at NVLink instead of HBM speed.  This is what lets one prover hold
// the ~430 GiB of BASELINE config[3] at x = 24; it is memory pooling, not compute scaling.
struct PeerPool {
    int home = -1;
    std::vector<int> peers;
    std::unordered_map<void*, int> owner;  // blocks that live on a peer
    void* ballast = nullptr;               // 1 GiB held back on the home GPU for the small allocations of the stream-ordered pool
    uint64_t remote_live = 0, remote_peak = 0;
};
PeerPool g_peer;  // guarded by g_big_mutex

cudaError_t peer_malloc(void** p, size_t n) {
    int best = -1;
    size_t best_free = 0;
    for (int d : g_peer.peers) {
        size_t f = 0, t = 0;
        if (cudaSetDevice(d) == cudaSuccess && cudaMemGetInfo(&f, &t) == cudaSuccess && f > best_free) {
            best_free = f;
            best = d;
        }
    }
    // what a peer keeps for itself (GKR_PEER_RESERVE_GIB, default 0.25): raise it when the peers also run MSM team workers
    static const double reserve_gib = getenv("GKR_PEER_RESERVE_GIB") ? atof(getenv("GKR_PEER_RESERVE_GIB")) : 0.25;
    cudaError_t e = cudaErrorMemoryAllocation;
    if (best >= 0 && (double)best_free > (double)n + reserve_gib * 1073741824.0) {
        cudaSetDevice(best);
        e = cudaMalloc(p, n);
        if (e == cudaSuccess) {
            g_peer.owner[*p] = best;
            g_peer.remote_live += n;
            if (g_peer.remote_live > g_peer.remote_peak) g_peer.remote_peak = g_peer.remote_live;
        }
    }
    cudaSetDevice(g_peer.home);
    if (e != cudaSuccess) (void)cudaGetLastError();
    return e;
}
// give a block back to whoever owns it: the stream-ordered pool of the home GPU, or the peer it was placed on
void release_block(void* p, size_t n, cudaStream_t s) {
    auto it = g_peer.owner.find(p);
    if (it == g_peer.owner.end()) {
        cudaFreeAsync(p, s);
        return;
    }
    cudaStreamSynchronize(s);  // kernels of the home GPU may still be reading it
    cudaSetDevice(it->second);
    cudaFree(p);
    cudaSetDevice(g_peer.home);
    g_peer.remote_live -= n;
    g_peer.owner.erase(it);
}
}  // namespace

// out[0] = live bytes in large blocks, out[1] = their peak since the last reset
void gkr_big_mem_stats(uint64_t out[2], bool reset_peak) {
    std::lock_guard<std::mutex> lk(g_big_mutex);
    out[0] = g_big_live;
    out[1] = g_big_peak;
    if (reset_peak) g_big_peak = g_big_live;
}

cudaError_t gkr_malloc_async_impl(void** p, size_t n, cudaStream_t s) {
    if (n < GKR_BIG_BLOCK) {
        cudaError_t es = cudaMallocAsync(p, n, s);
        if (es != cudaSuccess) fprintf(stderr, "[gkr-msm-b200] small allocation of %zu bytes failed: %s\n", n, cudaGetErrorString(es));
        return es;
    }
    std::lock_guard<std::mutex> lk(g_big_mutex);
    BigCache& c = g_big[s];
    auto it = c.free_blocks.lower_bound(n);
    if (it != c.free_blocks.end() && it->first <= 2 * n) {
        *p = it->second;
        c.live[*p] = it->first;
        big_account((int64_t)it->first);
        c.free_blocks.erase(it);
        return cudaSuccess;
    }
    // test hook (GKR_PEER_POOL_LOCAL_LIMIT_MIB): pretend the home GPU is full beyond this many MiB of live tables, so that the
    // peer placement is exercised by instances of any size
    static const long long local_limit = getenv("GKR_PEER_POOL_LOCAL_LIMIT_MIB") ? atoll(getenv("GKR_PEER_POOL_LOCAL_LIMIT_MIB")) : -1;
    cudaError_t e = cudaErrorMemoryAllocation;
    bool force_peer = local_limit >= 0 && !g_peer.peers.empty() && g_big_live - g_peer.remote_live + n > ((uint64_t)local_limit << 20);
    if (!force_peer && !g_peer.peers.empty()) {
        // with a peer pool the home GPU keeps 2 GiB free for the small allocations of the stream-ordered pool (parameter arrays,
        // result buffers): a large block that would eat into that goes to a peer
        size_t f = 0, t = 0;
        if (cudaMemGetInfo(&f, &t) == cudaSuccess && f < n + ((size_t)2 << 30)) force_peer = true;
    }
    if (!force_peer) e = cudaMallocAsync(p, n, s);
    if (e != cudaSuccess && !force_peer) {  // out of memory: give the cached blocks back and retry once
        for (auto& kv : c.free_blocks) release_block(kv.second, kv.first, s);
        c.free_blocks.clear();
        (void)cudaGetLastError();
        e = cudaMallocAsync(p, n, s);
    }
    if (e != cudaSuccess && !g_peer.peers.empty()) {  // the home GPU is full: place the block on a peer (NVLink peer access)
        (void)cudaGetLastError();
        if (g_peer.ballast) {  // first time: hand the reserve to the stream-ordered pool for the small allocations to come
            cudaFree(g_peer.ballast);
            g_peer.ballast = nullptr;
        }
        e = peer_malloc(p, n);
    }
    if (e == cudaSuccess) {
        c.live[*p] = n;
        big_account((int64_t)n);
    } else {
        fprintf(stderr, "[gkr-msm-b200] allocation of %.1f MiB failed: %.2f GiB of tables live (%.2f GiB of them on %zu peer GPUs)\n", n / 1048576.0,
                g_big_live / 1073741824.0, g_peer.remote_live / 1073741824.0, g_peer.peers.size());
    }
    return e;
}

cudaError_t gkr_free_async(void* p, cudaStream_t s) {
    if (!p) return cudaSuccess;
    {
        std::lock_guard<std::mutex> lk(g_big_mutex);
        auto ci = g_big.find(s);
        if (ci != g_big.end()) {
            auto it = ci->second.live.find(p);
            if (it != ci->second.live.end()) {
                ci->second.free_blocks.emplace(it->second, p);
                big_account(-(int64_t)it->second);
                ci->second.live.erase(it);
                return cudaSuccess;
            }
        }
    }
    return cudaFreeAsync(p, s);
}

void gkr_big_cache_release(cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_big_mutex);
    auto ci = g_big.find(s);
    if (ci == g_big.end()) return;
    for (auto& kv : ci->second.free_blocks) release_block(kv.second, kv.first, s);
    g_big.erase(ci);
}

// Lend this context the memory of the other GPUs of the box: devices [0, n_devices) except its own become the peer pool.
// out_stats (optional, 2 x u64): bytes currently placed on peers and their peak.  n_devices <= 1 only reads the statistics.
extern "C" int gkr_ctx_peer_pool(gkr_ctx* ctx, int n_devices, uint64_t* out_stats) {
    if (!ctx) return GKR_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_big_mutex);
    if (n_devices > 1 && g_peer.peers.empty()) {
        int count = 0;
        GKR_CUDA_OK(ctx, cudaGetDeviceCount(&count));
        if (n_devices > count) return ctx->fail(GKR_ERR_ARG, "gkr_ctx_peer_pool: fewer devices than requested");
        GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
        for (int d = 0; d < n_devices; d++) {
            if (d == ctx->device) continue;
            int can = 0;
            GKR_CUDA_OK(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, d));
            if (!can) return ctx->fail(GKR_ERR_UNSUPPORTED, "gkr_ctx_peer_pool: no peer access between the GPUs");
            cudaError_t e = cudaDeviceEnablePeerAccess(d, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return ctx->fail(GKR_ERR_CUDA, cudaGetErrorString(e));
            (void)cudaGetLastError();
            g_peer.peers.push_back(d);
        }
        g_peer.home = ctx->device;
        if (cudaMalloc(&g_peer.ballast, (size_t)1 << 30) != cudaSuccess) {
            g_peer.ballast = nullptr;
            (void)cudaGetLastError();
        }
    }
    if (out_stats) {
        out_stats[0] = g_peer.remote_live;
        out_stats[1] = g_peer.remote_peak;
    }
    return GKR_OK;
}

int gkr_slot_wait_seq(gkr_ctx* ctx, int slot, uint32_t seq, uint32_t n_blocks, int n_acc, gkr::FrH* out) {
    if (n_blocks > GKR_HOST_FOLD_MAX_BLOCKS) n_blocks = 1;  // large launches fold their partials on the device (grid_reduce_to_host)
    GkrSlot* s = &ctx->slots_host[slot];
    uint64_t spins = 0;
    const uint64_t t0 = gkr_now_ns();
    while (s->flag != seq) {
        if ((++spins & 0xffff) == 0) {
            cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q == cudaSuccess) {
                if (s->flag == seq) break;
                return ctx->fail(GKR_ERR_CUDA, "round kernel finished without publishing its result");
            }
            if (q != cudaErrorNotReady) return ctx->fail(GKR_ERR_CUDA, std::string("round kernel failed: ") + cudaGetErrorString(q));
        }
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    const uint64_t dt = gkr_now_ns() - t0;
    ctx->ns_wait += dt;
    ctx->n_waits++;
    ctx->wait_hist_ns[ctx->wait_kind % 3][ctx->wait_log % 40] += dt;
    ctx->wait_hist_n[ctx->wait_kind % 3][ctx->wait_log % 40]++;
    for (int a = 0; a < n_acc; a++) {
        gkr::FrH acc = gkr::frh::ZERO;
        for (uint32_t b = 0; b < n_blocks; b++) acc = gkr::frh::add(acc, fr_to_host(s->part[(size_t)b * n_acc + a]));
        out[a] = acc;
    }
    return GKR_OK;
}

__global__ void fetch_firsts_kernel(const __grid_constant__ GkrFirsts f, RoundOut o) {
    if (threadIdx.x < f.n) o.part[threadIdx.x] = f.p[threadIdx.x][0];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        *(volatile uint32_t*)o.flag = o.seq;
    }
}
int gkr_fetch_firsts(gkr_ctx* ctx, int slot, const GkrFirsts& f, gkr::FrH* out) {
    RoundOut o = ctx->round_out(slot);
    fetch_firsts_kernel<<<1, 32, 0, ctx->stream>>>(f, o);
    ctx->launches++;
    GKR_CUDA_OK(ctx, cudaGetLastError());
    return gkr_slot_wait(ctx, slot, 1, f.n, out);
}

__global__ void fetch_firsts_dev_kernel(const Fr* const* ptrs, int n, RoundOut o) {
    for (int j = threadIdx.x; j < n; j += blockDim.x) o.part[j] = ptrs[j][0];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        *(volatile uint32_t*)o.flag = o.seq;
    }
}
int gkr_fetch_firsts_dev(gkr_ctx* ctx, int slot, const Fr* const* d_ptrs, int n, gkr::FrH* out) {
    if (n > GKR_MAX_BLOCKS * GKR_MAX_DEG) return ctx->fail(GKR_ERR_UNSUPPORTED, "too many tables");
    RoundOut o = ctx->round_out(slot);
    fetch_firsts_dev_kernel<<<1, 64, 0, ctx->stream>>>(d_ptrs, n, o);
    ctx->launches++;
    GKR_CUDA_OK(ctx, cudaGetLastError());
    return gkr_slot_wait(ctx, slot, 1, n, out);
}

extern "C" int gkr_version(void) { return 1; }

extern "C" int gkr_ctx_create(int device, gkr_ctx** out) {
    if (!out) return GKR_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || device < 0 || device >= count) return GKR_ERR_CUDA;  // no CPU fallback
    gkr_ctx* ctx = new gkr_ctx();
    ctx->device = device;
    auto bail = [&](cudaError_t err) {
        fprintf(stderr, "gkr_ctx_create: %s\n", cudaGetErrorString(err));
        delete ctx;
        return (int)GKR_ERR_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail(e);
    ctx->num_sms = prop.multiProcessorCount;
    { const char* nf = getenv("GKR_NO_FAST_FOLD"); ctx->no_fast_fold = nf && nf[0] == '1'; }
    { const char* v = getenv("GKR_DENSE_FLAVOR"); if (v) ctx->dense_flavor = atoi(v); }
    { const char* v = getenv("GKR_MSM_SIGNED"); if (v) ctx->msm_signed = atoi(v); }
    { const char* v = getenv("GKR_MSM_LIGHT_MINB"); if (v) ctx->msm_light_minb = atoi(v); }
    { const char* v = getenv("GKR_DENSE_STAGED_MIN"); if (v && atoll(v) >= 0) ctx->dense_staged_min = (uint64_t)atoll(v); }
    { const char* v = getenv("GKR_DENSE_SMALL_MAX"); if (v && atoll(v) >= 0) ctx->dense_small_max = (uint64_t)atoll(v); }
    { const char* v = getenv("GKR_DEG2_COMPACT_MAX"); if (v && atoll(v) >= 0) ctx->deg2_compact_max = (uint64_t)atoll(v); }
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e);
    {
        // stream-ordered allocator with an unbounded release threshold: after warm-up a table / ping-pong slab
        // allocation is a pointer bump, not a driver call (a proof creates ~10^3 short-lived objects)
        cudaMemPool_t pool;
        if ((e = cudaDeviceGetDefaultMemPool(&pool, device)) != cudaSuccess) return bail(e);
        uint64_t thr = UINT64_MAX;
        if ((e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr)) != cudaSuccess) return bail(e);
    }
    if ((e = cudaMalloc(&ctx->partials, sizeof(Fr) * GKR_MAX_BLOCKS * GKR_MAX_DEG)) != cudaSuccess) return bail(e);
    if ((e = cudaMalloc(&ctx->ticket, sizeof(unsigned int))) != cudaSuccess) return bail(e);
    if ((e = cudaMemset(ctx->ticket, 0, sizeof(unsigned int))) != cudaSuccess) return bail(e);
    if ((e = cudaHostAlloc(&ctx->result_host, sizeof(Fr) * GKR_RESULT_SLOTS * GKR_MAX_DEG, cudaHostAllocMapped)) != cudaSuccess) return bail(e);
    if ((e = cudaHostGetDevicePointer(&ctx->result_dev, ctx->result_host, 0)) != cudaSuccess) return bail(e);
    if ((e = cudaHostAlloc(&ctx->slots_host, sizeof(GkrSlot) * GKR_RESULT_SLOTS, cudaHostAllocMapped)) != cudaSuccess) return bail(e);
    if ((e = cudaHostGetDevicePointer(&ctx->slots_dev, ctx->slots_host, 0)) != cudaSuccess) return bail(e);
    for (int i = 0; i < GKR_RESULT_SLOTS; i++) ctx->slots_host[i].flag = 0;
    if ((e = cudaHostAlloc(&ctx->mbox_host, sizeof(GkrMailbox) * GKR_RESULT_SLOTS, cudaHostAllocMapped)) != cudaSuccess) return bail(e);
    if ((e = cudaHostGetDevicePointer(&ctx->mbox_dev, ctx->mbox_host, 0)) != cudaSuccess) return bail(e);
    std::memset((void*)ctx->mbox_host, 0, sizeof(GkrMailbox) * GKR_RESULT_SLOTS);
    if ((e = cudaMalloc(&ctx->mbox_bcast, sizeof(uint32_t) * 8 * GKR_RESULT_SLOTS)) != cudaSuccess) return bail(e);
    if ((e = cudaMemset(ctx->mbox_bcast, 0, sizeof(uint32_t) * 8 * GKR_RESULT_SLOTS)) != cudaSuccess) return bail(e);
    { const char* v = getenv("GKR_PRELAUNCH"); if (v) ctx->prelaunch = v[0] != '0'; }
    if ((e = cudaMalloc(&ctx->slot_tickets, sizeof(unsigned int) * GKR_RESULT_SLOTS)) != cudaSuccess) return bail(e);
    if ((e = cudaMemset(ctx->slot_tickets, 0, sizeof(unsigned int) * GKR_RESULT_SLOTS)) != cudaSuccess) return bail(e);
    {
        std::lock_guard<std::mutex> lk(g_slot_mutex);
        SlotPool* p = new SlotPool();
        for (int i = GKR_RESULT_SLOTS - 1; i >= 0; i--) p->free_list.push_back(i);  // indices into ctx->slots_*
        g_pools.push_back({ctx, p});
    }
    *out = ctx;
    return GKR_OK;
}

// Long-running provers: hand the cached large blocks (gkr_malloc_async) and the idle part of the stream-ordered pool back to
// the driver, e.g. between proofs of very different sizes.  Synchronises the context stream.
extern "C" int gkr_ctx_trim(gkr_ctx* ctx) {
    if (!ctx) return GKR_ERR_ARG;
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    ctx->deg2_layout.reset();
    gkr_big_cache_release(ctx->stream);
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    cudaMemPool_t pool;
    GKR_CUDA_OK(ctx, cudaDeviceGetDefaultMemPool(&pool, ctx->device));
    GKR_CUDA_OK(ctx, cudaMemPoolTrimTo(pool, 0));
    return GKR_OK;
}

extern "C" void gkr_ctx_destroy(gkr_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ctx->deg2_layout.reset();  // frees its device arrays on the stream
    if (ctx->stream) gkr_big_cache_release(ctx->stream);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->stage_host) cudaFreeHost(ctx->stage_host);
    {
        std::lock_guard<std::mutex> lk(g_slot_mutex);
        for (size_t i = 0; i < g_pools.size(); i++)
            if (g_pools[i].first == ctx) {
                delete g_pools[i].second;
                g_pools.erase(g_pools.begin() + i);
                break;
            }
    }
    if (ctx->partials) cudaFree(ctx->partials);
    if (ctx->ticket) cudaFree(ctx->ticket);
    if (ctx->result_host) cudaFreeHost(ctx->result_host);
    if (ctx->slots_host) cudaFreeHost(ctx->slots_host);
    if (ctx->mbox_host) cudaFreeHost((void*)ctx->mbox_host);
    if (ctx->mbox_bcast) cudaFree(ctx->mbox_bcast);
    if (ctx->slot_tickets) cudaFree(ctx->slot_tickets);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* gkr_last_error(const gkr_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int gkr_ctx_sync(gkr_ctx* ctx) {
    if (!ctx) return GKR_ERR_ARG;
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return GKR_OK;
}

extern "C" uint64_t gkr_ctx_launch_count(const gkr_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" void* gkr_ctx_stream(gkr_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

// ---- tables ----------------------------------------------------------------------------------------
extern "C" int gkr_table_alloc(gkr_ctx* ctx, uint64_t n, gkr_table** out) {
    if (!ctx || !out) return GKR_ERR_ARG;
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    gkr_table* t = new gkr_table();
    t->ctx = ctx;
    t->n = n;
    cudaError_t e = gkr_malloc_async(&t->d, sizeof(Fr) * std::max<uint64_t>(n, 1), ctx->stream);
    if (e != cudaSuccess) {
        delete t;
        return ctx->fail(GKR_ERR_CUDA, std::string("gkr_malloc_async(table): ") + cudaGetErrorString(e));
    }
    *out = t;
    return GKR_OK;
}

extern "C" int gkr_table_upload(gkr_ctx* ctx, const uint64_t* limbs, uint64_t n, gkr_table** out) {
    if (!ctx) return GKR_ERR_ARG;
    if (!limbs && n) return ctx->fail(GKR_ERR_ARG, "null host buffer");
    int rc = gkr_table_alloc(ctx, n, out);
    if (rc) return rc;
    if (n && sizeof(Fr) * n <= ((size_t)1 << 20)) return gkr_stage_upload(ctx, (*out)->d, limbs, sizeof(Fr) * n);  // small: pinned ring, truly asynchronous
    if (n) GKR_CUDA_OK(ctx, cudaMemcpyAsync((*out)->d, limbs, sizeof(Fr) * n, cudaMemcpyHostToDevice, ctx->stream));
    return GKR_OK;
}

extern "C" int gkr_table_download(gkr_ctx* ctx, const gkr_table* t, uint64_t* limbs_out) {
    if (!ctx) return GKR_ERR_ARG;
    if (!t || !limbs_out) return ctx->fail(GKR_ERR_ARG, "null argument");
    if (t->n) GKR_CUDA_OK(ctx, cudaMemcpyAsync(limbs_out, t->d, sizeof(Fr) * t->n, cudaMemcpyDeviceToHost, ctx->stream));
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return GKR_OK;
}

extern "C" uint64_t gkr_table_len(const gkr_table* t) { return t ? t->n : 0; }
extern "C" void* gkr_table_device_ptr(gkr_table* t) { return t ? (void*)t->d : nullptr; }

extern "C" void gkr_table_free(gkr_table* t) {
    if (!t) return;
    if (t->owned && t->d) gkr_free_async(t->d, t->ctx->stream);  // stream-ordered: queued kernels finish first
    delete t;
}

// ---- synthetic input generator ------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64_at(uint64_t seed, uint64_t k) {
    uint64_t z = seed + k * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__device__ __forceinline__ bool fr_geq_p(const Fr& a) {
    const uint32_t p[8] = {FR_P0, FR_P1, FR_P2, FR_P3, FR_P4, FR_P5, FR_P6, FR_P7};
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        if (a.l[i] > p[i]) return true;
        if (a.l[i] < p[i]) return false;
    }
    return true;
}

__device__ __forceinline__ Fr fr_sub_p_raw(const Fr& a) {
    const uint32_t p[8] = {FR_P0, FR_P1, FR_P2, FR_P3, FR_P4, FR_P5, FR_P6, FR_P7};
    Fr r;
    uint64_t br = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a.l[i] - p[i] - br;
        r.l[i] = (uint32_t)d;
        br = (d >> 63) & 1;
    }
    return r;
}

__global__ void synth_kernel(Fr* out, uint64_t n, uint64_t seed, uint64_t first) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Fr v;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint64_t z = splitmix64_at(seed, 4 * (first + i) + j + 1);
            v.l[2 * j] = (uint32_t)z;
            v.l[2 * j + 1] = (uint32_t)(z >> 32);
        }
        if (fr_geq_p(v)) v = fr_sub_p_raw(v);
        if (fr_geq_p(v)) v = fr_sub_p_raw(v);
        out[i] = v;
    }
}

extern "C" int gkr_table_synth(gkr_ctx* ctx, uint64_t seed, uint64_t first_index, uint64_t n, gkr_table** out) {
    int rc = gkr_table_alloc(ctx, n, out);
    if (rc) return rc;
    if (n) {
        unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->num_sms * 8);
        synth_kernel<<<grid, 256, 0, ctx->stream>>>((*out)->d, n, seed, first_index);
        ctx->launches++;
        GKR_CUDA_OK(ctx, cudaGetLastError());
    }
    return GKR_OK;
}

// ---- eq tables -------------------------------------------------------------------------------------------
// Small table (n <= 10): one block runs the reference's doubling construction level by level
// ([w - w r_i, w r_i], src/utils.rs:239-244) in shared memory.
__global__ void eq_small_kernel(Fr* out, const Fr* point, int n, Fr mult) {
    extern __shared__ Fr sh[];  // 2 * 2^n
    Fr* a = sh;
    Fr* b = sh + ((size_t)1 << n);
    if (threadIdx.x == 0) a[0] = mult;
    __syncthreads();
    for (int lvl = 1; lvl <= n; lvl++) {
        Fr r = point[lvl - 1];
        int half = 1 << (lvl - 1);
        for (int j = threadIdx.x; j < half; j += blockDim.x) {
            Fr w = a[j];
            Fr m = fr_mul(r, w);
            b[2 * j] = fr_sub(w, m);
            b[2 * j + 1] = m;
        }
        __syncthreads();
        Fr* tmp = a;
        a = b;
        b = tmp;
    }
    for (int j = threadIdx.x; j < (1 << n); j += blockDim.x) out[j] = a[j];
}

// out[i] = hi[i >> k] * lo[i & (2^k - 1)]: eq over the concatenated point factors into the two halves.
__global__ void eq_outer_kernel(Fr* out, const Fr* hi, const Fr* lo, int k, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t mask = ((uint64_t)1 << k) - 1;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        out[i] = fr_mul(hi[i >> k], lo[i & mask]);
    }
}

static int eq_build(gkr_ctx* ctx, const Fr* d_point, uint32_t n, const Fr& mult, Fr* d_out) {
    if (n <= 10) {
        size_t sh = sizeof(Fr) * 2 * ((size_t)1 << n);
        if (sh > 48 * 1024) GKR_CUDA_OK(ctx, cudaFuncSetAttribute(eq_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        eq_small_kernel<<<1, 256, sh, ctx->stream>>>(d_out, d_point, (int)n, mult);
        ctx->launches++;
        GKR_CUDA_OK(ctx, cudaGetLastError());
        return GKR_OK;
    }
    uint32_t h = n / 2, k = n - h;
    Fr *hi = nullptr, *lo = nullptr;
    GKR_CUDA_OK(ctx, gkr_malloc_async(&hi, sizeof(Fr) << h, ctx->stream));
    GKR_CUDA_OK(ctx, gkr_malloc_async(&lo, sizeof(Fr) << k, ctx->stream));
    int rc = eq_build(ctx, d_point, h, mult, hi);
    if (rc == GKR_OK) rc = eq_build(ctx, d_point + h, k, fr_from_host(gkr::frh::ONE), lo);
    if (rc == GKR_OK) {
        uint64_t total = (uint64_t)1 << n;
        unsigned grid = (unsigned)std::min<uint64_t>((total + 255) / 256, (uint64_t)ctx->num_sms * 8);
        eq_outer_kernel<<<grid, 256, 0, ctx->stream>>>(d_out, hi, lo, (int)k, total);
        ctx->launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = ctx->fail(GKR_ERR_CUDA, cudaGetErrorString(e));
    }
    gkr_free_async(hi, ctx->stream);
    gkr_free_async(lo, ctx->stream);
    return rc;
}

int gkr_eq_build_device(gkr_ctx* ctx, const Fr* d_point, uint32_t n, const Fr& mult, Fr* d_out) { return eq_build(ctx, d_point, n, mult, d_out); }

int gkr_stage_upload(gkr_ctx* ctx, void* d_dst, const void* src, size_t n) {
    if (n == 0) return GKR_OK;
    const size_t RING = (size_t)64 << 20;  // a 2^20-point proof stages ~20 MB of parameters: the ring wraps (one stream synchronisation) every few proofs
    if (!ctx->stage_host) {
        GKR_CUDA_OK(ctx, cudaHostAlloc(&ctx->stage_host, RING, cudaHostAllocDefault));
        ctx->stage_size = RING;
        ctx->stage_pos = 0;
    }
    if (n > ctx->stage_size / 2) {  // too large for the ring: plain copy, synchronised so that the source may be released
        GKR_CUDA_OK(ctx, cudaMemcpyAsync(d_dst, src, n, cudaMemcpyHostToDevice, ctx->stream));
        GKR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        return GKR_OK;
    }
    if (ctx->stage_pos + n > ctx->stage_size) {  // wrap: everything staged so far must have left the ring
        GKR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->stage_pos = 0;
    }
    std::memcpy(ctx->stage_host + ctx->stage_pos, src, n);
    GKR_CUDA_OK(ctx, cudaMemcpyAsync(d_dst, ctx->stage_host + ctx->stage_pos, n, cudaMemcpyHostToDevice, ctx->stream));
    ctx->stage_pos += (n + 255) & ~(size_t)255;
    return GKR_OK;
}

extern "C" int gkr_eq_table(gkr_ctx* ctx, const uint64_t* point, uint32_t n, const uint64_t mult[4], gkr_table** out) {
    if (!ctx || !out || (!point && n) || !mult) return GKR_ERR_ARG;
    if (n >= 40) return ctx->fail(GKR_ERR_ARG, "eq table: too many variables");
    gkr::FrH m = frh_from_limbs(mult);
    if (!frh_canonical(m)) return ctx->fail(GKR_ERR_ARG, "eq table: multiplier not canonical");
    for (uint32_t i = 0; i < n; i++)
        if (!frh_canonical(frh_from_limbs(point + 4 * i))) return ctx->fail(GKR_ERR_ARG, "eq table: point not canonical");
    int rc = gkr_table_alloc(ctx, (uint64_t)1 << n, out);
    if (rc) return rc;
    Fr* d_point = nullptr;
    if (gkr_malloc_async(&d_point, sizeof(Fr) * std::max<uint32_t>(n, 1), ctx->stream) != cudaSuccess) {
        gkr_table_free(*out);
        *out = nullptr;
        return ctx->fail(GKR_ERR_CUDA, "eq table: out of device memory");
    }
    if (n) {
        int rcs = gkr_stage_upload(ctx, d_point, point, sizeof(Fr) * n);
        if (rcs) {  // nothing leaks on the staging failure path
            gkr_free_async(d_point, ctx->stream);
            gkr_table_free(*out);
            *out = nullptr;
            return rcs;
        }
    }
    rc = eq_build(ctx, d_point, n, fr_from_host(m), (*out)->d);
    gkr_free_async(d_point, ctx->stream);
    if (rc) {
        gkr_table_free(*out);
        *out = nullptr;
    }
    return rc;
}

// ---- per-launch timing (used only by bench.py's roofline leg) ----------------------------------------------
extern "C" int gkr_ctx_host_stats(gkr_ctx* ctx, uint64_t out[4], int reset) {
    if (!ctx || !out) return GKR_ERR_ARG;
    out[0] = ctx->ns_launch; out[1] = ctx->ns_wait; out[2] = ctx->n_waits; out[3] = ctx->launches;
    if (reset) ctx->ns_launch = ctx->ns_wait = ctx->n_waits = 0;
    return GKR_OK;
}

extern "C" int gkr_ctx_set_fast_fold(gkr_ctx* ctx, int on) {
    if (!ctx) return GKR_ERR_ARG;
    ctx->no_fast_fold = on == 0;
    return GKR_OK;
}

extern "C" int gkr_ctx_set_tuning(gkr_ctx* ctx, const char* key, long long value) {
    if (!ctx) return GKR_ERR_ARG;
    if (!key || value < 0) return ctx->fail(GKR_ERR_ARG, "gkr_ctx_set_tuning: null key / negative value");
    const std::string k(key);
    if (k == "dense_small_max") ctx->dense_small_max = (uint64_t)value;
    else if (k == "deg2_compact_max") ctx->deg2_compact_max = (uint64_t)value;
    else if (k == "msm_signed") ctx->msm_signed = (int)value;
    else if (k == "prelaunch") ctx->prelaunch = value != 0;
    else if (k == "msm_light_minb") ctx->msm_light_minb = (int)value;
    else return ctx->fail(GKR_ERR_ARG, "gkr_ctx_set_tuning: unknown key");
    return GKR_OK;
}

extern "C" int gkr_ctx_timing_enable(gkr_ctx* ctx, int on) {
    if (!ctx) return GKR_ERR_ARG;
    ctx->timing = on != 0;
    return GKR_OK;
}

// Drains the recorded launches: kernel_id[i], n_items[i], ms[i].  Returns the number written (<= max_n).
extern "C" int gkr_ctx_timing_read(gkr_ctx* ctx, int* kernel_id, uint64_t* n_items, float* ms, int max_n) {
    if (!ctx) return GKR_ERR_ARG;
    cudaStreamSynchronize(ctx->stream);
    int n = 0;
    for (auto& t : ctx->timed) {
        if (n < max_n) {
            float v = 0.f;
            cudaEventElapsedTime(&v, t.start, t.stop);
            kernel_id[n] = t.kernel_id;
            n_items[n] = t.n_items;
            ms[n] = v;
            n++;
        }
        ctx->event_pool.push_back(t.start);
        ctx->event_pool.push_back(t.stop);
    }
    ctx->timed.clear();
    return n;
}
