// DenseSumcheckObjectSO on the device  (reference: src/cleanup/protocols/sumcheck.rs:241-347).
//
// One kernel per sumcheck round.  `bind(t)` of round k and `unipoly()` of round k+1 are FUSED: the kernel
// reads every table once (4 consecutive elements per thread and table), folds them with the challenge
// (bind_dense_poly, sumcheck.rs:160-163: p'[i] = p[2i] + t (p[2i+1] - p[2i])), writes the half-size table
// and, from the two fresh elements still in registers, accumulates the next round's evaluations at
// 1..deg exactly like sumcheck.rs:295-313 (args = p[2i+1]; difs = p[2i+1]-p[2i]; args += difs per node).
// Algorithmic traffic: 32 B read + 16 B written per table element per round (the reference's separate
// unipoly + bind passes move 64 + 16).  The per-round result (deg field elements) is written by the last
// block straight into pinned host memory; the challenge travels as a kernel argument, so a round costs
// one launch and one stream synchronisation and nothing table-sized ever crosses PCIe.
#include <algorithm>
#include <atomic>
#include "common.cuh"
#include "gates.cuh"
#include "dense_kernel.cuh"
#include "dense_split_kernel.cuh"
#include "so.hpp"

// out[j][i] = in[j][2i] + t (in[j][2i+1] - in[j][2i])   -- used for the last round (one pair -> one value)
__global__ void dense_fold_kernel(const __grid_constant__ DenseRoundArgs A, int n_polys) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < A.n_items; i += stride) {
        for (int j = 0; j < n_polys; j++) {
            Fr e0 = A.in[j][2 * i], e1 = A.in[j][2 * i + 1];
            A.out[j][i] = fr_add(e0, fr_mul(A.t, fr_sub(e1, e0)));
        }
    }
}

// ---- gate dispatch --------------------------------------------------------------------------------
template <class F>
static int dispatch_dense_so(gkr_ctx* ctx, int so_kind, int gate, uint32_t param, F&& f) {
    if (so_kind == GKR_SO_PLAIN) {
        if (gate == GKR_GATE_PROD3) return f(SoProd3{});
        if (gate == GKR_GATE_FOLDED_PROD) {
            switch (param) {
                case 1: return f(SoFoldedProd<1>{});
                case 2: return f(SoFoldedProd<2>{});
                case 3: return f(SoFoldedProd<3>{});
                case 4: return f(SoFoldedProd<4>{});
                default: return ctx->fail(GKR_ERR_UNSUPPORTED, "FOLDED_PROD: nargs must be 1..4");
            }
        }
        return ctx->fail(GKR_ERR_UNSUPPORTED, "GKR_SO_PLAIN supports PROD3 and FOLDED_PROD");
    }
    if (so_kind == GKR_SO_EQ_GAMMA) {
        switch (gate) {
            case GKR_GATE_AFF_L1: return f(SoEqGamma<GATE_AFF_L1>{});
            case GKR_GATE_AFF_L2: return f(SoEqGamma<GATE_AFF_L2>{});
            case GKR_GATE_AFF_L3: return f(SoEqGamma<GATE_AFF_L3>{});
            case GKR_GATE_PRJ_L1: return f(SoEqGamma<GATE_PRJ_L1>{});
            case GKR_GATE_PRJ_L2: return f(SoEqGamma<GATE_PRJ_L2>{});
            case GKR_GATE_PRJ_L3: return f(SoEqGamma<GATE_PRJ_L3>{});
            case GKR_GATE_AFF_L1_BITCHECK2: return f(SoEqGamma<GATE_AFF_L1_BITCHECK2>{});
            case GKR_GATE_LOGUP_LAYER: return f(SoEqGamma<GATE_LOGUP_LAYER>{});
            case GKR_GATE_ADD_INVERSES: return f(SoEqGamma<GATE_ADD_INVERSES>{});
            default: return ctx->fail(GKR_ERR_UNSUPPORTED, "GKR_SO_EQ_GAMMA: unsupported gate");
        }
    }
    return ctx->fail(GKR_ERR_ARG, "unknown so_kind");
}

template <class SO, int MODE, bool FAST = false>
static int launch_dense_round(gkr_ctx* ctx, const DenseRoundArgs& args, uint32_t* n_blocks_out) {
    static std::atomic<int> blocks_per_sm{0};  // same value whichever thread computes it first
    if (blocks_per_sm == 0) {
        int b = 0;
        GKR_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, dense_round_kernel<SO, MODE, FAST>, GKR_REDUCE_THREADS, 0));
        blocks_per_sm.store(std::max(b, 1));
    }
    if constexpr (MODE != 2) {
        if (args.n_items <= ctx->dense_small_max) {  // small round: the block-cooperative kernel
            unsigned grid = (unsigned)std::max<uint64_t>(1, (args.n_items + GKR_DENSE_SMALL_QB - 1) / GKR_DENSE_SMALL_QB);
            *n_blocks_out = grid;
            {
                GkrLaunchTimer timer(ctx, MODE == 0 ? GKR_K_DENSE_EVAL : GKR_K_DENSE_FOLD_EVAL, args.n_items, args.mbox.box == nullptr);
                dense_small_kernel<SO, MODE, FAST><<<grid, GKR_REDUCE_THREADS, 0, ctx->stream>>>(args);
            }
            ctx->launches++;
            GKR_CUDA_OK(ctx, cudaGetLastError());
            return GKR_OK;
        }
    }
    if constexpr (MODE == 1 && SO::P <= 4) {
        // fused fold+eval rounds over whole 32-item tiles: tables staged HBM -> shared memory by cp.async one tile ahead
        // (dense_round_staged_kernel; measured 0.672 -> 0.649 ms on the 2^24 x 3 Prod3 round).  dense_flavor 0 disables it.
        if (ctx->dense_flavor != 0 && ctx->dense_flavor != 1 && args.n_items % 32 == 0 && args.n_items >= ctx->dense_staged_min) {
            static std::atomic<int> staged_blocks_per_sm{0};
            constexpr size_t smem = (size_t)GKR_STAGED_WARPS * SO::P * 32 * 4 * 32;
            if (staged_blocks_per_sm == 0) {
                GKR_CUDA_OK(ctx, cudaFuncSetAttribute(dense_round_staged_kernel<SO, MODE, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                int b = 0;
                GKR_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, dense_round_staged_kernel<SO, MODE, FAST>, GKR_REDUCE_THREADS, smem));
                staged_blocks_per_sm.store(std::max(b, 1));
            }
            // two waves of blocks: the grid-stride tiles of a block that starts late even out the tail
            uint64_t want = (args.n_items / 32 + GKR_STAGED_WARPS - 1) / GKR_STAGED_WARPS;
            uint64_t cap = std::min<uint64_t>((uint64_t)ctx->num_sms * staged_blocks_per_sm * 2, GKR_MAX_BLOCKS);
            unsigned grid = (unsigned)std::max<uint64_t>(1, std::min(want, cap));
            *n_blocks_out = grid;
            {
                GkrLaunchTimer timer(ctx, GKR_K_DENSE_FOLD_EVAL, args.n_items);
                dense_round_staged_kernel<SO, MODE, FAST><<<grid, GKR_REDUCE_THREADS, smem, ctx->stream>>>(args);
            }
            ctx->launches++;
            GKR_CUDA_OK(ctx, cudaGetLastError());
            return GKR_OK;
        }
    }
    if constexpr (MODE != 2 && SO::DEG <= 4) {
        if (ctx->dense_flavor == 1) {  // node-split kernel: one warp per evaluation node
            static std::atomic<int> split_blocks_per_sm{0};
            constexpr int threads = 32 * SO::DEG;
            if (split_blocks_per_sm == 0) {
                int b = 0;
                GKR_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, dense_round_split_kernel<SO, MODE, FAST>, threads, 0));
                split_blocks_per_sm.store(std::max(b, 1));
            }
            uint64_t want = (args.n_items + 31) / 32;
            uint64_t cap = std::min<uint64_t>((uint64_t)ctx->num_sms * split_blocks_per_sm, GKR_MAX_BLOCKS);
            unsigned grid = (unsigned)std::max<uint64_t>(1, std::min(want, cap));
            *n_blocks_out = grid;
            {
                GkrLaunchTimer timer(ctx, MODE == 0 ? GKR_K_DENSE_EVAL : GKR_K_DENSE_FOLD_EVAL, args.n_items);
                dense_round_split_kernel<SO, MODE, FAST><<<grid, threads, 0, ctx->stream>>>(args);
            }
            ctx->launches++;
            GKR_CUDA_OK(ctx, cudaGetLastError());
            return GKR_OK;
        }
    }
    uint64_t want = (args.n_items + GKR_REDUCE_THREADS - 1) / GKR_REDUCE_THREADS;
    uint64_t cap = std::min<uint64_t>((uint64_t)ctx->num_sms * blocks_per_sm, GKR_MAX_BLOCKS);
    unsigned grid = (unsigned)std::max<uint64_t>(1, std::min(want, cap));
    // small rounds: one block no wider than the work, so the shuffle/shared-memory reduction stays shallow
    unsigned threads = GKR_REDUCE_THREADS;
    if (grid == 1) threads = (unsigned)std::max<uint64_t>(32, std::min<uint64_t>(GKR_REDUCE_THREADS, (args.n_items + 31) / 32 * 32));
    *n_blocks_out = grid;
    {
        GkrLaunchTimer timer(ctx, MODE == 0 ? GKR_K_DENSE_EVAL : (MODE == 1 ? GKR_K_DENSE_FOLD_EVAL : GKR_K_DENSE_SUM), args.n_items);
        dense_round_kernel<SO, MODE, FAST><<<grid, threads, 0, ctx->stream>>>(args);
    }
    ctx->launches++;
    GKR_CUDA_OK(ctx, cudaGetLastError());
    return GKR_OK;
}

struct DenseSoInfo {
    int P, DEG, HDEG, N_OUTS;
    int gamma_shift[GKR_MAX_GATE_CONSTS];
};

static int dense_so_info(gkr_ctx* ctx, int so_kind, int gate, uint32_t param, DenseSoInfo* info) {
    return dispatch_dense_so(ctx, so_kind, gate, param, [&](auto so) {
        using SO = decltype(so);
        info->P = SO::P;
        info->DEG = SO::DEG;
        info->HDEG = SO::HDEG;
        info->N_OUTS = SO::N_OUTS;
        for (int i = 0; i < GKR_MAX_GATE_CONSTS; i++) info->gamma_shift[i] = i < SO::N_OUTS ? SO::gamma_shift(i) : 0;
        return (int)GKR_OK;
    });
}

int gkr_result_slot_acquire(gkr_ctx* ctx);
void gkr_result_slot_release(gkr_ctx* ctx, int slot);

// 2^128 and 2^-128 in Montgomery form (host side of the FAST folds)
static const gkr::FrH& pow128_m() {
    static const gkr::FrH v = gkr::frh::mul(gkr::FrH{{0, 0, 1, 0}}, gkr::frh::R2);
    return v;
}
static const gkr::FrH& inv128_m() {
    static const gkr::FrH v = gkr::frh::inverse(pow128_m());
    return v;
}

class DenseSO : public gkr_so {
   public:
    int so_kind, gate;
    uint32_t gate_param;
    GateConsts consts;                      // device view for the CURRENT scale of the tables
    gkr::FrH base_consts[GKR_MAX_GATE_CONSTS];  // the caller's gate constants
    DenseSoInfo info;
    int P = 0, DEG = 0;
    // FAST folds leave the tables scaled by sigma = 2^(-128 fast_folds); sums come out scaled by sigma^HDEG
    uint32_t fast_folds = 0;
    gkr::FrH sigma = gkr::frh::ONE, unscale_sum = gkr::frh::ONE, unscale_val = gkr::frh::ONE;
    uint32_t num_vars = 0, round_idx = 0;
    gkr::FrH claim_;
    const Fr* cur[GKR_MAX_POLYS];  // current tables (round 0: the caller's tables, untouched)
    Fr* slab = nullptr;            // P * (n/2 + n/4) ping-pong storage
    Fr* buf[2][GKR_MAX_POLYS];
    int next_buf = 0;
    bool sums_pending = false;  // a launched kernel will deliver the sums of the current round
    uint32_t pending_blocks = 0;
    uint32_t pending_seq = 0;   // result-slot sequence number of that kernel
    // pre-launched next round (set_prelaunch): enqueued by unipoly() of round k before the challenge exists, released by bind(t_k)
    bool allow_prelaunch = false, pre_active = false;
    uint32_t pre_mbox_seq = 0, pre_blocks = 0, pre_slot_seq = 0;
    void set_prelaunch(bool on) override { allow_prelaunch = on; }
    bool cached = false;
    gkr::FrH evals[GKR_MAX_DEG + 1];
    int slot = -1;

    ~DenseSO() override {
        if (pre_active) ctx->post_mailbox(slot, pre_mbox_seq, 2, nullptr);  // a spinning pre-launched kernel exits without publishing
        if (slab) gkr_free_async(slab, ctx->stream);
        if (slot >= 0) gkr_result_slot_release(ctx, slot);
    }

    void fill_common(DenseRoundArgs& a) {
        a.consts = consts;
        a.o = ctx->round_out(slot);
    }

    // gate constants for tables scaled by sigma: gamma^i * sigma^gamma_shift(i)   (gates.cuh, gamma_eval_scaled)
    void consts_for(const gkr::FrH& sg, GateConsts* out) const {
        for (int i = 0; i < GKR_MAX_GATE_CONSTS; i++) {
            gkr::FrH g = base_consts[i];
            if (i < info.N_OUTS && info.gamma_shift[i] > 0) {
                if (i == 0) g = gkr::frh::ONE;  // output 0 has coefficient one in the reference (sumcheck.rs:724-731)
                for (int k = 0; k < info.gamma_shift[i]; k++) g = gkr::frh::mul(g, sg);
            }
            out->g[i] = fr_from_host(g);
        }
    }
    void rescale_consts() { consts_for(sigma, &consts); }

    int ensure_slab(uint64_t new_len) {
        if (slab) return GKR_OK;
        uint64_t per = new_len + (new_len >> 1);
        GKR_CUDA_OK(ctx, gkr_malloc_async(&slab, sizeof(Fr) * per * P, ctx->stream));
        for (int j = 0; j < P; j++) {
            buf[0][j] = slab + (size_t)j * per;
            buf[1][j] = slab + (size_t)j * per + new_len;
        }
        return GKR_OK;
    }

    // Enqueue the fused kernel of the NEXT round (fold by the challenge of this round, evaluate the round after) before that
    // challenge exists: small rounds only (the block-cooperative kernel carries the mailbox prologue), 128-bit-challenge folds
    // only (what transcript.challenge(128) produces; anything else cancels the launch in bind()).
    int maybe_prelaunch() {
        if (!allow_prelaunch || !ctx->prelaunch || pre_active || info.HDEG == 0 || ctx->no_fast_fold) return GKR_OK;
        if (round_idx + 1 >= num_vars) return GKR_OK;
        const uint64_t new_len = (uint64_t)1 << (num_vars - round_idx - 1);
        if (new_len < 2 || (new_len >> 1) > ctx->dense_small_max) return GKR_OK;
        int rc = ensure_slab(new_len);
        if (rc) return rc;
        DenseRoundArgs a;
        for (int j = 0; j < P; j++) { a.in[j] = cur[j]; a.out[j] = buf[next_buf][j]; }
        a.t = fr_from_host(gkr::frh::ZERO);
        a.t128[0] = a.t128[1] = a.t128[2] = a.t128[3] = 0;
        consts_for(gkr::frh::mul(sigma, inv128_m()), &a.consts);  // the tables after one more fast fold
        a.o = ctx->round_out(slot);
        a.mbox = ctx->next_mailbox(slot);
        a.n_items = new_len >> 1;
        uint32_t nb = 0;
        rc = dispatch_dense_so(ctx, so_kind, gate, gate_param, [&](auto so) {
            using SO = decltype(so);
            if constexpr (SO::HDEG > 0) return launch_dense_round<SO, 1, true>(ctx, a, &nb);
            return (int)GKR_ERR_UNSUPPORTED;
        });
        if (rc) return rc;
        pre_active = true;
        pre_mbox_seq = a.mbox.seq;
        pre_slot_seq = a.o.seq;
        pre_blocks = nb;
        return GKR_OK;
    }

    int unipoly(gkr::FrH* out, uint32_t* n_evals) override {
        if (round_idx >= num_vars) return ctx->fail(GKR_ERR_PROTOCOL, "unipoly: the protocol has already ended");
        if (!cached) {
            if (!sums_pending) {
                DenseRoundArgs a;
                for (int j = 0; j < P; j++) { a.in[j] = cur[j]; a.out[j] = nullptr; }
                a.n_items = (uint64_t)1 << (num_vars - round_idx - 1);
                fill_common(a);
                int rc = dispatch_dense_so(ctx, so_kind, gate, gate_param, [&](auto so) {
                    return launch_dense_round<decltype(so), 0>(ctx, a, &pending_blocks);
                });
                if (rc) return rc;
                pending_seq = a.o.seq;
            }
            {
                int rcp = maybe_prelaunch();  // the next round's kernel queues up behind the one we are about to wait for
                if (rcp) return rcp;
            }
            ctx->wait_kind = 0;
            ctx->wait_log = (int)(num_vars - round_idx - 1);
            int rcw = gkr_slot_wait_seq(ctx, slot, pending_seq, pending_blocks, DEG, evals + 1);
            if (rcw) return rcw;
            if (fast_folds)
                for (int s = 1; s <= DEG; s++) evals[s] = gkr::frh::mul(evals[s], unscale_sum);
            sums_pending = false;
            evals[0] = gkr::frh::sub(claim_, evals[1]);  // sumcheck.rs:325
            cached = true;
        }
        for (int s = 0; s <= DEG; s++) out[s] = evals[s];
        if (n_evals) *n_evals = DEG + 1;
        return GKR_OK;
    }

    int bind(const gkr::FrH& t) override {
        if (round_idx >= num_vars) return ctx->fail(GKR_ERR_PROTOCOL, "bind: the protocol has already ended");
        if (!cached) return ctx->fail(GKR_ERR_PROTOCOL, "bind: should evaluate unipoly before binding");
        if (!frh_canonical(t)) return ctx->fail(GKR_ERR_ARG, "bind: challenge is not a canonical field element");
        const uint64_t cur_len = (uint64_t)1 << (num_vars - round_idx);
        const uint64_t new_len = cur_len >> 1;
        {
            int rcs = ensure_slab(new_len);
            if (rcs) return rcs;
        }
        DenseRoundArgs a;
        for (int j = 0; j < P; j++) { a.in[j] = cur[j]; a.out[j] = buf[next_buf][j]; }
        a.t = fr_from_host(t);
        // transcript.challenge(128) is a 128-bit integer: fold with fr_fold128 when the gate is homogeneous
        const gkr::FrH t_plain = gkr::frh::mul(t, gkr::FrH{{1, 0, 0, 0}});
        const bool fast = info.HDEG > 0 && new_len >= 2 && t_plain.v[2] == 0 && t_plain.v[3] == 0 && !ctx->no_fast_fold;
        bool released = false;
        if (pre_active) {
            pre_active = false;
            if (fast && !ctx->mailbox_timed_out(slot, pre_mbox_seq)) {
                const uint32_t tw[4] = {(uint32_t)t_plain.v[0], (uint32_t)(t_plain.v[0] >> 32), (uint32_t)t_plain.v[1], (uint32_t)(t_plain.v[1] >> 32)};
                ctx->post_mailbox(slot, pre_mbox_seq, 1, tw);  // the kernel is already resident (or next in the stream): go
                released = true;
            } else {
                // a launch that gave up means something serialises launches behind our back (a profiler replaying kernels, a
                // debugger): stop pre-launching on this context instead of paying the watchdog again
                if (ctx->mailbox_timed_out(slot, pre_mbox_seq)) ctx->prelaunch = false;
                ctx->post_mailbox(slot, pre_mbox_seq, 2, nullptr);  // full-width challenge (or the launch gave up): the ordinary launch below
            }
        }
        if (fast) {
            a.t128[0] = (uint32_t)t_plain.v[0]; a.t128[1] = (uint32_t)(t_plain.v[0] >> 32);
            a.t128[2] = (uint32_t)t_plain.v[1]; a.t128[3] = (uint32_t)(t_plain.v[1] >> 32);
            fast_folds++;
            sigma = gkr::frh::mul(sigma, inv128_m());
            unscale_val = gkr::frh::mul(unscale_val, pow128_m());
            for (int k = 0; k < info.HDEG; k++) unscale_sum = gkr::frh::mul(unscale_sum, pow128_m());
            rescale_consts();
        }
        if (released) {
            sums_pending = true;
            pending_blocks = pre_blocks;
            pending_seq = pre_slot_seq;
        } else if (new_len >= 2) {
            fill_common(a);
            a.n_items = new_len >> 1;
            int rc = dispatch_dense_so(ctx, so_kind, gate, gate_param, [&](auto so) {
                using SO = decltype(so);
                if constexpr (SO::HDEG > 0) {
                    if (fast) return launch_dense_round<SO, 1, true>(ctx, a, &pending_blocks);
                }
                return launch_dense_round<SO, 1, false>(ctx, a, &pending_blocks);
            });
            if (rc) return rc;
            sums_pending = true;
            pending_seq = a.o.seq;
        } else {
            fill_common(a);
            a.n_items = new_len;
            dense_fold_kernel<<<1, 32, 0, ctx->stream>>>(a, P);
            ctx->launches++;
            GKR_CUDA_OK(ctx, cudaGetLastError());
        }
        for (int j = 0; j < P; j++) cur[j] = buf[next_buf][j];
        next_buf ^= 1;
        claim_ = gkr::frh::interpolate_eval(evals, DEG + 1, t);  // u.evaluate(&t), sumcheck.rs:273
        cached = false;
        round_idx++;
        return GKR_OK;
    }

    int final_evals(gkr::FrH* out) override {
        if (round_idx != num_vars) return ctx->fail(GKR_ERR_PROTOCOL, "final_evals: can only be called after the last round");
        GkrFirsts f;
        f.n = P;
        for (int j = 0; j < P; j++) f.p[j] = cur[j];
        int rc = gkr_fetch_firsts(ctx, slot, f, out);
        if (rc) return rc;
        if (fast_folds)
            for (int j = 0; j < P; j++) out[j] = gkr::frh::mul(out[j], unscale_val);
        return GKR_OK;
    }

    gkr::FrH claim() const override { return claim_; }
    uint32_t degree() const override { return DEG; }
    uint32_t num_polys() const override { return P; }
    uint32_t round() const override { return round_idx; }
};

static int fill_consts(gkr_ctx* ctx, const gkr::FrH* consts, uint32_t n_consts, GateConsts* out) {
    if (n_consts > GKR_MAX_GATE_CONSTS) return ctx->fail(GKR_ERR_ARG, "too many gate constants");
    for (uint32_t i = 0; i < GKR_MAX_GATE_CONSTS; i++) {
        if (i < n_consts) {
            if (!frh_canonical(consts[i])) return ctx->fail(GKR_ERR_ARG, "gate constant is not canonical");
            out->g[i] = fr_from_host(consts[i]);
        } else {
            for (int k = 0; k < 8; k++) out->g[i].l[k] = 0;
        }
    }
    return GKR_OK;
}

int gkr_make_dense_so(gkr_ctx* ctx, int so_kind, int gate, uint32_t gate_param, const gkr::FrH* consts, uint32_t n_consts,
                      gkr_table* const* tables, uint32_t n_polys, uint32_t num_vars, const gkr::FrH& claim, gkr_so** out) {
    DenseSoInfo info;
    int rc = dense_so_info(ctx, so_kind, gate, gate_param, &info);
    if (rc) return rc;
    if ((int)n_polys != info.P) return ctx->fail(GKR_ERR_ARG, "number of tables != f.n_ins()");  // sumcheck.rs:255
    if (num_vars >= 40) return ctx->fail(GKR_ERR_ARG, "num_vars too large");
    for (uint32_t j = 0; j < n_polys; j++) {
        if (!tables[j] || tables[j]->n != ((uint64_t)1 << num_vars))
            return ctx->fail(GKR_ERR_ARG, "table length != 1 << num_vars");  // sumcheck.rs:257
    }
    if (!frh_canonical(claim)) return ctx->fail(GKR_ERR_ARG, "claim is not canonical");
    DenseSO* so = new DenseSO();
    so->ctx = ctx;
    so->so_kind = so_kind;
    so->gate = gate;
    so->gate_param = gate_param;
    rc = fill_consts(ctx, consts, n_consts, &so->consts);
    if (rc) { delete so; return rc; }
    for (uint32_t i = 0; i < GKR_MAX_GATE_CONSTS; i++) so->base_consts[i] = i < n_consts ? consts[i] : gkr::frh::ZERO;
    so->info = info;
    so->rescale_consts();
    so->P = info.P;
    so->DEG = info.DEG;
    so->num_vars = num_vars;
    so->claim_ = claim;
    for (uint32_t j = 0; j < n_polys; j++) so->cur[j] = tables[j]->d;
    so->slot = gkr_result_slot_acquire(ctx);
    if (so->slot < 0) { delete so; return ctx->fail(GKR_ERR_UNSUPPORTED, "too many live sumcheck objects"); }
    *out = so;
    return GKR_OK;
}

// sum_i f(tables[.][i])
int gkr_dense_gate_sum_impl(gkr_ctx* ctx, int so_kind, int gate, uint32_t gate_param, const gkr::FrH* consts, uint32_t n_consts,
                            gkr_table* const* tables, uint32_t n_polys, gkr::FrH* out) {
    DenseSoInfo info;
    int rc = dense_so_info(ctx, so_kind, gate, gate_param, &info);
    if (rc) return rc;
    if ((int)n_polys != info.P) return ctx->fail(GKR_ERR_ARG, "number of tables != f.n_ins()");
    DenseRoundArgs a;
    rc = fill_consts(ctx, consts, n_consts, &a.consts);
    if (rc) return rc;
    if (info.gamma_shift[0] > 0) a.consts.g[0] = fr_from_host(gkr::frh::ONE);  // gamma_eval_scaled reads g[0]
    for (uint32_t j = 0; j < n_polys; j++) {
        if (!tables[j] || tables[j]->n != tables[0]->n) return ctx->fail(GKR_ERR_ARG, "tables must have equal length");
        a.in[j] = tables[j]->d;
        a.out[j] = nullptr;
    }
    a.n_items = tables[0]->n;
    int slot = gkr_result_slot_acquire(ctx);
    if (slot < 0) return ctx->fail(GKR_ERR_UNSUPPORTED, "no free result slot");
    a.o = ctx->round_out(slot);
    uint32_t nb = 0;
    rc = dispatch_dense_so(ctx, so_kind, gate, gate_param, [&](auto so) { return launch_dense_round<decltype(so), 2>(ctx, a, &nb); });
    if (rc == GKR_OK) rc = gkr_slot_wait(ctx, slot, nb, 1, out);
    gkr_result_slot_release(ctx, slot);
    return rc;
}
