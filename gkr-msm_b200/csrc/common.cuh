// Shared host/device plumbing for the C-ABI implementation: context, tables, error handling and the
// block / grid reduction of per-thread field accumulators (warp shuffles -> shared memory -> last block).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <ctime>
#include <memory>
#include <string>
#include <vector>
#include "../../include/gkr_msm_b200.h"
#include "field.cuh"
#include "host_field.hpp"

#define GKR_MAX_POLYS 16        // max tables per sumcheck object (triangle L1 has 12 inputs + eq)
#define GKR_MAX_DEG 4           // max number of accumulated evaluation points per round
#define GKR_REDUCE_THREADS 128  // block size of the sumcheck round kernels
#define GKR_MAX_BLOCKS 1024     // upper bound on the grid of a round kernel (one partial per block and accumulator)
#define GKR_RESULT_SLOTS 128    // live sumcheck objects per context (128 KiB of pinned, mapped host memory each)

// Result channel of one sumcheck object, in pinned host memory mapped into the device address space.  Every block of a
// round kernel writes its partial sums here; the last block (device ticket) publishes `flag = seq`.  The host spins on
// the flag (no stream synchronisation, no memcpy launch) and folds the <= GKR_MAX_BLOCKS partials itself.
struct GkrSlot {
    volatile uint32_t flag;
    uint32_t pad_[7];
    Fr part[GKR_MAX_BLOCKS * GKR_MAX_DEG];
};

// Launches of more than this many blocks fold their per-block partials on the DEVICE (last block, second stage) and publish
// one result; smaller launches let the host fold the few partials (one PCIe write per block, no second stage).
#define GKR_HOST_FOLD_MAX_BLOCKS 8

struct RoundOut {  // kernel-side view of a slot
    Fr* part;
    Fr* dev_part;  // device scratch for the per-block partials of large launches
    uint32_t* flag;
    unsigned int* ticket;
    uint32_t seq;
};

// Challenge mailbox of one sumcheck object, in pinned host memory mapped into the device address space (one per result slot).
// The round kernel of round k + 1 is ENQUEUED while round k still runs -- before its challenge exists -- and waits for the host,
// which hashes round k's message and writes the challenge here: the launch latency leaves the critical path.  One 32-byte
// sector {seq, cmd, t[0..4), 0, 0} that the device fetches with ONE 256-bit system-scope load per poll (a PCIe read is ~2 us:
// everything must arrive in one); only thread 0 of block 0 polls the host, the other blocks wait on a copy in device memory.
// Host writes cmd / t first and seq last.  A launch that waits longer than GKR_MAILBOX_TIMEOUT_NS gives up without
// publishing a result and says so in `timed_out` (the host relaunches the round the ordinary way): a pre-launched kernel can
// never hang the device.  The challenge travels as the plain 128-bit integer of transcript.challenge(128).
struct __attribute__((aligned(64))) GkrMailbox {
    volatile uint32_t seq;        // host -> device, written after cmd / t
    volatile uint32_t cmd;        // 1: go, the challenge is in t; 2: cancelled
    volatile uint32_t t[4];       // the 128-bit challenge as a plain integer
    uint32_t pad_;
    volatile uint32_t seq2;       // the same sequence number at the END of the 32-byte sector: should the 256-bit poll ever be served as
                                  // two 16-byte reads, a torn snapshot shows two different numbers and the poll is simply repeated
    uint32_t pad2_[7];
    volatile uint32_t timed_out;  // device -> host: seq of a launch that gave up
};
static_assert(offsetof(GkrMailbox, seq2) == 28 && sizeof(GkrMailbox) == 64, "mailbox layout: one 32-byte sector + the reply word");
#define GKR_MAILBOX_TIMEOUT_NS 2000000000ull
struct MailboxRef {  // kernel-side view; box == nullptr: the challenge travels in the kernel arguments
    GkrMailbox* box = nullptr;
    uint32_t* bcast = nullptr;  // device memory, 8 words per slot: {seq, cmd, t[0..4)} re-published by the polling thread
    uint32_t seq = 0;
};

struct Deg2Layout;  // deg2.cu: row layout of the most recent ragged sumcheck bundle
struct gkr_msm_team;  // msm_team.cu: commitment MSMs split by point range over the GPUs of one box

struct gkr_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    uint64_t ns_launch = 0, ns_wait = 0, n_waits = 0;  // host-side latency accounting (gkr_ctx_host_stats)
    // GKR_TRACE: waits by object kind (0 dense, 1 Deg2 dense, 2 Deg2 ragged) and log2 of the pairs of the round
    int wait_kind = 0, wait_log = 0;
    uint64_t wait_hist_ns[3][40] = {{0}}, wait_hist_n[3][40] = {{0}};
    bool no_fast_fold = false;      // test hook (GKR_NO_FAST_FOLD=1): always fold with the full Montgomery product
    Fr* partials = nullptr;         // [GKR_MAX_BLOCKS * GKR_MAX_DEG] device scratch (device-side two-stage reductions)
    unsigned int* ticket = nullptr; // device counter for the last-block pattern
    Fr* result_host = nullptr;      // pinned + mapped scratch for one-shot reductions (gate sums)
    Fr* result_dev = nullptr;       // device alias of result_host
    GkrSlot* slots_host = nullptr;  // [GKR_RESULT_SLOTS] pinned + mapped result channels
    GkrSlot* slots_dev = nullptr;
    unsigned int* slot_tickets = nullptr;  // [GKR_RESULT_SLOTS] device
    uint32_t slot_seq[GKR_RESULT_SLOTS] = {0};
    GkrMailbox* mbox_host = nullptr;  // [GKR_RESULT_SLOTS] pinned + mapped challenge mailboxes (pre-launched rounds)
    GkrMailbox* mbox_dev = nullptr;
    uint32_t* mbox_bcast = nullptr;   // [GKR_RESULT_SLOTS][8] device memory
    uint32_t mbox_seq[GKR_RESULT_SLOTS] = {0};
    bool prelaunch = true;            // GKR_PRELAUNCH=0: never enqueue a round before its challenge is known
    MailboxRef next_mailbox(int slot) {
        MailboxRef m;
        m.box = mbox_dev + slot;
        m.bcast = mbox_bcast + 8 * slot;
        m.seq = ++mbox_seq[slot];
        return m;
    }
    // host side of the mailbox: cmd 1 releases the pre-launched kernel with the challenge words, cmd 2 cancels it
    void post_mailbox(int slot, uint32_t seq, uint32_t cmd, const uint32_t* t4) {
        GkrMailbox* m = mbox_host + slot;
        for (int k = 0; k < 4; k++) m->t[k] = t4 ? t4[k] : 0u;
        m->cmd = cmd;
        __atomic_store_n((uint32_t*)&m->seq, seq, __ATOMIC_RELEASE);
        __atomic_store_n((uint32_t*)&m->seq2, seq, __ATOMIC_RELEASE);
    }
    bool mailbox_timed_out(int slot, uint32_t seq) const { return mbox_host[slot].timed_out == seq; }
    RoundOut round_out(int slot) {  // next launch on this slot
        RoundOut o;
        o.part = slots_dev[slot].part;
        o.dev_part = partials;
        o.flag = (uint32_t*)&slots_dev[slot].flag;
        o.ticket = slot_tickets + slot;
        o.seq = ++slot_seq[slot];
        return o;
    }
    // optional per-launch timing (bench.py's roofline leg): CUDA events on the launching stream
    bool timing = false;
    struct TimedLaunch {
        int kernel_id;
        uint64_t n_items;
        cudaEvent_t start, stop;
    };
    std::vector<TimedLaunch> timed;
    std::vector<cudaEvent_t> event_pool;  // recycled events so a timed launch costs two cudaEventRecord only
    cudaEvent_t take_event() {
        if (event_pool.empty()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            return e;
        }
        cudaEvent_t e = event_pool.back();
        event_pool.pop_back();
        return e;
    }
    // pinned staging ring for small parameter uploads (one truly asynchronous H2D copy per object instead of a dozen
    // pageable ones, no synchronisation until the ring wraps)
    // kernel-selection thresholds (defaults measured on B200; GKR_DENSE_SMALL_MAX / GKR_DEG2_COMPACT_MAX override them for experiments)
    uint64_t dense_small_max = 4096, deg2_compact_max = 32768;
    // large dense rounds (GKR_DENSE_FLAVOR): 2 (default) cp.async-staged kernel for fused rounds of <= 4 tables, register kernel otherwise;
    // 0 register kernel only; 1 node-split kernel (dense_split_kernel.cuh, measured slower: kept for the lab)
    int dense_flavor = 2;
    // MSM window recoding (GKR_MSM_SIGNED): 1 = signed digits from 2^15 points (half the buckets per window), 2 = always, 0 = unsigned
    int msm_signed = 1;
    int msm_light_minb = 3;  // GKR_MSM_LIGHT_MINB: register cap of the one-thread-per-bucket accumulation kernel (blocks per SM)
    uint64_t dense_staged_min = (uint64_t)1 << 13;  // items (quads) from which the staged kernel is used (GKR_DENSE_STAGED_MIN)
    gkr_msm_team* team = nullptr;        // leader only: large gkr_msm_g1 calls are shared with the worker ranks
    uint64_t team_min_n = (uint64_t)1 << 18;
    std::shared_ptr<Deg2Layout> deg2_layout;  // reused by consecutive VecVec objects over the same rows
    unsigned char* stage_host = nullptr;
    size_t stage_size = 0, stage_pos = 0;
    int fail(int code, const std::string& msg) {
        err = msg;
        return code;
    }
};

struct gkr_table {
    gkr_ctx* ctx = nullptr;
    Fr* d = nullptr;
    uint64_t n = 0;
    bool owned = true;
};

// digit / counter arrays resident on the device (gkr_u32_upload)
struct gkr_u32buf {
    gkr_ctx* ctx = nullptr;
    uint32_t* d = nullptr;
    uint64_t n = 0;
};

// VecVecPolynomial<F> (src/cleanup/polys/vecvec.rs:149-160) in CSR form: rows back to back, every row even-length
struct gkr_vecvec {
    gkr_ctx* ctx = nullptr;
    Fr* d = nullptr;
    uint64_t total = 0;
    std::vector<uint32_t> row_len;  // host copy (even lengths)
    gkr::FrH row_pad, col_pad;
    uint32_t row_logsize = 0, col_logsize = 0;
};

// kernel ids reported by gkr_ctx_timing_read
enum GkrKernelId { GKR_K_DENSE_EVAL = 0, GKR_K_DENSE_FOLD_EVAL = 1, GKR_K_DENSE_SUM = 2, GKR_K_DENSE_FOLD = 3 };

static inline uint64_t gkr_now_ns() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

struct GkrLaunchTimer {
    gkr_ctx* ctx;
    bool on;
    uint64_t t0;
    gkr_ctx::TimedLaunch t;
    // events = false: host-side accounting only (a pre-launched kernel spins on its mailbox: its event time is not kernel time)
    GkrLaunchTimer(gkr_ctx* c, int kernel_id, uint64_t n_items, bool events = true) : ctx(c), on(c->timing && events), t0(gkr_now_ns()) {
        if (!on) return;
        t.kernel_id = kernel_id;
        t.n_items = n_items;
        t.start = ctx->take_event();
        t.stop = ctx->take_event();
        cudaEventRecord(t.start, ctx->stream);
    }
    ~GkrLaunchTimer() {
        ctx->ns_launch += gkr_now_ns() - t0;
        if (!on) return;
        cudaEventRecord(t.stop, ctx->stream);
        ctx->timed.push_back(t);
    }
};

#define GKR_CUDA_OK(ctx, call)                                                                          \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            return (ctx)->fail(GKR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));      \
        }                                                                                               \
    } while (0)

static inline gkr::FrH frh_from_limbs(const uint64_t* p) {
    gkr::FrH r;
    for (int i = 0; i < 4; i++) r.v[i] = p[i];
    return r;
}
static inline void frh_to_limbs(const gkr::FrH& a, uint64_t* p) {
    for (int i = 0; i < 4; i++) p[i] = a.v[i];
}
static inline bool frh_canonical(const gkr::FrH& a) { return !gkr::frh::geq_mod(a.v); }

static inline Fr fr_from_host(const gkr::FrH& a) {
    Fr r;
    for (int i = 0; i < 4; i++) {
        r.l[2 * i] = (uint32_t)a.v[i];
        r.l[2 * i + 1] = (uint32_t)(a.v[i] >> 32);
    }
    return r;
}
static inline gkr::FrH fr_to_host(const Fr& a) {
    gkr::FrH r;
    for (int i = 0; i < 4; i++) r.v[i] = (uint64_t)a.l[2 * i] | ((uint64_t)a.l[2 * i + 1] << 32);
    return r;
}

// Device memory: the stream-ordered pool (cudaMallocAsync) for everything small, and a per-stream cache of LARGE blocks
// (>= GKR_BIG_BLOCK) in front of it.  A proof at x = 20 allocates and frees hundreds of 0.1 - 1.5 GB tables and slabs; handing
// those back to the pool lets it fragment, and the driver then re-maps physical memory to satisfy the next large request --
// sporadic stalls of hundreds of ms.  Sizes repeat exactly from sumcheck to sumcheck and from proof to proof, so freed large
// blocks are kept and handed out again (best fit, at most 2x the request); reuse is ordered by the single context stream.
#define GKR_BIG_BLOCK ((size_t)4 << 20)
cudaError_t gkr_malloc_async_impl(void** p, size_t n, cudaStream_t s);
cudaError_t gkr_free_async(void* p, cudaStream_t s);
void gkr_big_cache_release(cudaStream_t s);  // return every cached block of this stream to the pool
template <class T>
static inline cudaError_t gkr_malloc_async(T** p, size_t n, cudaStream_t s) {
    return gkr_malloc_async_impl((void**)p, n, s);
}

// copy n bytes of host data to the device through the context's pinned staging ring (asynchronous on ctx->stream; the
// source may be freed as soon as the call returns)
int gkr_stage_upload(gkr_ctx* ctx, void* d_dst, const void* src, size_t n);

// final_evals of a sumcheck object: element 0 of each of the n <= GKR_MAX_POLYS tables, delivered through the object's
// result slot (one tiny kernel + a flag spin instead of n device-to-host copies and a stream synchronisation)
struct GkrFirsts {
    const Fr* p[GKR_MAX_POLYS];
    int n;
};
int gkr_fetch_firsts(gkr_ctx* ctx, int slot, const GkrFirsts& f, gkr::FrH* out);
// the same for any number of tables whose pointers already sit in a DEVICE array
int gkr_fetch_firsts_dev(gkr_ctx* ctx, int slot, const Fr* const* d_ptrs, int n, gkr::FrH* out);

// host side: wait for the launch `seq` on `slot` and fold its per-block partials (n_acc accumulators per block)
int gkr_slot_wait_seq(gkr_ctx* ctx, int slot, uint32_t seq, uint32_t n_blocks, int n_acc, gkr::FrH* out);
static inline int gkr_slot_wait(gkr_ctx* ctx, int slot, uint32_t n_blocks, int n_acc, gkr::FrH* out) {  // the latest launch on the slot
    return gkr_slot_wait_seq(ctx, slot, ctx->slot_seq[slot], n_blocks, n_acc, out);
}

#ifdef __CUDACC__
// Device side of the mailbox.  Returns false when the launch is cancelled or gave up (all threads of the grid agree; nothing
// may be published then).  t_out: the 4 challenge words, for every thread.
__device__ __forceinline__ bool gkr_mailbox_wait(const MailboxRef& m, uint32_t* t_out) {
    __shared__ uint32_t mb_sh[5];
    if (threadIdx.x == 0) {
        uint32_t w[8];
        w[1] = 0;
        if (blockIdx.x == 0 && blockIdx.y == 0) {  // the one thread of the grid that talks to the host
            unsigned long long t0, t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            uint32_t polls = 0;
            for (;;) {
                asm volatile("ld.volatile.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                             : "l"(m.box)
                             : "memory");
                if (w[0] == m.seq && w[7] == m.seq) break;
                if ((++polls & 15u) == 0) {
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    if (t1 - t0 > GKR_MAILBOX_TIMEOUT_NS) {
                        w[1] = 3;
                        m.box->timed_out = m.seq;
                        __threadfence_system();
                        break;
                    }
                }
            }
            if (gridDim.x * gridDim.y > 1) {  // re-publish in device memory for the other blocks: payload, fence, seq
#pragma unroll
                for (int k = 1; k < 6; k++) ((volatile uint32_t*)m.bcast)[k] = w[k];
                __threadfence();
                ((volatile uint32_t*)m.bcast)[0] = m.seq;
            }
        } else {
            while (((volatile uint32_t*)m.bcast)[0] != m.seq) {}
            __threadfence();
#pragma unroll
            for (int k = 1; k < 6; k++) w[k] = ((volatile uint32_t*)m.bcast)[k];
        }
        mb_sh[4] = w[1];
#pragma unroll
        for (int k = 0; k < 4; k++) mb_sh[k] = w[2 + k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) t_out[k] = mb_sh[k];
    return mb_sh[4] == 1;
}

// Sum `acc[0..N)` over all threads of the grid.  Field addition is associative and commutative and every
// partial is canonical, so the result is bit-identical to the reference's sequential / rayon sum whatever
// the order.  Block: shuffle tree inside each warp, one shared-memory hop, warp 0 finishes.  Grid: every
// block publishes its partial, the last block to take a ticket folds them (threadfence reduction) and
// writes the N results to `result` (pinned host memory mapped into the device address space).
template <int N>
__device__ __forceinline__ void block_reduce_fr(Fr* acc, Fr* smem /* [N * warps] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
    for (int s = 0; s < N; s++) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc[s] = fr_add(acc[s], fr_shfl_down(acc[s], off));
        if (lane == 0) smem[s * nwarps + warp] = acc[s];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int s = 0; s < N; s++) {
            Fr v = lane < nwarps ? smem[s * nwarps + lane] : fr_zero();
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v = fr_add(v, fr_shfl_down(v, off));
            acc[s] = v;
        }
    }
    __syncthreads();
}

// Block-level sums go straight to the host-mapped slot; the last block to take a ticket publishes the sequence number.
// Launches of more than GKR_HOST_FOLD_MAX_BLOCKS blocks park their partials in device memory instead and the last block
// folds them (hundreds of partials cost the host ~10 us per round and one PCIe write + system fence per block).
template <int N>
__device__ __forceinline__ void grid_reduce_to_host(Fr* acc, Fr* smem, const RoundOut& o) {
    block_reduce_fr<N>(acc, smem);
    const unsigned int n_blocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (n_blocks <= GKR_HOST_FOLD_MAX_BLOCKS) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int s = 0; s < N; s++) o.part[(size_t)bid * N + s] = acc[s];
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
        for (int s = 0; s < N; s++) o.dev_part[(size_t)bid * N + s] = acc[s];
        __threadfence();
        unsigned int tk = atomicAdd(o.ticket, 1u);
        is_last = (tk == n_blocks - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
#pragma unroll
    for (int s = 0; s < N; s++) acc[s] = fr_zero();
    for (unsigned int b = threadIdx.x; b < n_blocks; b += blockDim.x) {
#pragma unroll
        for (int s = 0; s < N; s++) {
            const Fr* p = &o.dev_part[(size_t)b * N + s];
            Fr v;
            asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v.l[0]), "=r"(v.l[1]), "=r"(v.l[2]), "=r"(v.l[3]), "=r"(v.l[4]), "=r"(v.l[5]), "=r"(v.l[6]), "=r"(v.l[7])
                         : "l"(p));
            acc[s] = fr_add(acc[s], v);
        }
    }
    block_reduce_fr<N>(acc, smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < N; s++) o.part[s] = acc[s];
        *o.ticket = 0;
        __threadfence_system();
        *(volatile uint32_t*)o.flag = o.seq;
    }
}

template <int N>
__device__ __forceinline__ void grid_reduce_fr(Fr* acc, Fr* smem, Fr* partials, unsigned int* ticket, Fr* result) {
    block_reduce_fr<N>(acc, smem);
    const unsigned int n_blocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (n_blocks == 1) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int s = 0; s < N; s++) result[s] = acc[s];
        }
        return;
    }
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < N; s++) partials[(size_t)bid * N + s] = acc[s];
        __threadfence();
        unsigned int tk = atomicAdd(ticket, 1u);
        is_last = (tk == n_blocks - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
#pragma unroll
    for (int s = 0; s < N; s++) acc[s] = fr_zero();
    for (unsigned int b = threadIdx.x; b < n_blocks; b += blockDim.x) {
#pragma unroll
        for (int s = 0; s < N; s++) {
            const Fr* p = &partials[(size_t)b * N + s];
            Fr v;
            asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v.l[0]), "=r"(v.l[1]), "=r"(v.l[2]), "=r"(v.l[3]), "=r"(v.l[4]), "=r"(v.l[5]), "=r"(v.l[6]), "=r"(v.l[7])
                         : "l"(p));
            acc[s] = fr_add(acc[s], v);
        }
    }
    block_reduce_fr<N>(acc, smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < N; s++) result[s] = acc[s];
        *ticket = 0;
    }
}
#endif
