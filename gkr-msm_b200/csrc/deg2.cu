// Degree-2 gate sumchecks with the eq factor pulled out (Gruen-style), dense and ragged:
//   DenseDeg2SumcheckObjectSO      src/cleanup/protocols/sumchecks/dense_eq.rs:62-173
//   VecVecDeg2(Lo)SumcheckObjectSO src/cleanup/protocols/sumchecks/vecvec_eq.rs:74-398
//   EQPolyData / EQPolyPointParts  src/cleanup/polys/vecvec.rs:20-147
//   VecVecPolynomial, bind_21      src/cleanup/polys/vecvec.rs:149-206, 420-441
//   UnivarFormat::from12           src/cleanup/protocols/sumchecks/vecvec_eq.rs:197-216
//
// Device layout of a VecVecPolynomial: CSR -- one flat Fr array (rows concatenated, every row even-length
// as VecVecPolynomial::new pads them) plus per-round row offsets precomputed on the host for ALL sparse
// rounds at construction (row lengths halve and re-pad deterministically, vecvec.rs:432-437).
// The reference's in-place `make_21` (p[2i] <- 2 p[2i+1] - p[2i]) is never materialised: the eval kernel
// forms the value at "2" in registers, and the fold p[2i+1] + (t-1)(p2 - p[2i+1]) == p[2i] + t (p[2i+1]-p[2i]).
// Per round the device returns S1 = sum_pairs w*G(p at 1), S2 = sum_pairs w*G(p at 2) with
// w = eq_row[idx] * row_eq_coefs[row] and G = sum_o gamma^o f_o, plus T = sum_rows row_eq_coefs[row] *
// (1 - sum_{idx < len/2} eq_row[idx]) for the closed-form padding term; the O(1) rest (pad_results, col pad,
// multiplier, from12) is host arithmetic exactly as in the reference.
#include <algorithm>
#include <memory>
#include "common.cuh"
#include "gates.cuh"
#include "host_gates.hpp"
#include "deg2_kernel.cuh"
#include "so.hpp"
#include "transcript.hpp"

#define GKR_DEG2_COMPACT_MAX_PAIRS 32768  // pairs x gate blocks up to which a round runs the compact kernel

int gkr_result_slot_acquire(gkr_ctx* ctx);
void gkr_result_slot_release(gkr_ctx* ctx, int slot);
int gkr_eq_build_device(gkr_ctx* ctx, const Fr* d_point, uint32_t n, const Fr& mult, Fr* d_out);

// defined in deg2_compact.cu: the same kernel built with the out-of-line multiplier
int gkr_launch_deg2_round_compact(int uniform_gate, const Deg2RoundArgs& a, dim3 grid, unsigned threads, cudaStream_t stream);

// ragged fold (VecVecPolynomial::bind_21): grid.y = table
struct VvFoldArgs {
    const Fr* const* in;
    Fr* const* out;
    const uint32_t* off_old;  // element offsets [nrows + 1]
    const uint32_t* off_new;
    uint32_t nrows;
    uint64_t n_new;
    Fr t;
    const Fr* row_pads;  // [P]
};

__global__ void vv_fold_kernel(const __grid_constant__ VvFoldArgs A) {
    const int j = blockIdx.y;
    const Fr* in = A.in[j];
    Fr* out = A.out[j];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < A.n_new; e += stride) {
        uint32_t lo = 0, hi = A.nrows;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if ((uint64_t)A.off_new[mid] <= e) lo = mid; else hi = mid;
        }
        uint64_t i = e - A.off_new[lo];
        uint64_t half_old = (A.off_old[lo + 1] - A.off_old[lo]) >> 1;
        Fr v;
        if (i < half_old) {
            const Fr* src = in + A.off_old[lo] + 2 * i;
            Fr e0 = src[0], e1 = src[1];
            v = fr_add(e0, fr_mul(A.t, fr_sub(e1, e0)));
        } else {
            v = A.row_pads[j];  // odd half re-padded with row_pad (vecvec.rs:432-436)
        }
        out[e] = v;
    }
}

// bind_into_dense (vecvec_eq.rs:157-175): one value per bucket row, col_pad beyond the last row
struct VvToDenseArgs {
    const Fr* const* in;
    Fr* const* out;
    const uint32_t* off_old;
    uint32_t nrows;
    uint64_t n_out;
    Fr t;
    const Fr* row_pads;
    const Fr* col_pads;
};

__global__ void vv_to_dense_kernel(const __grid_constant__ VvToDenseArgs A) {
    const int j = blockIdx.y;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < A.n_out; r += (uint64_t)gridDim.x * blockDim.x) {
        Fr v;
        if (r < A.nrows) {
            uint32_t len = A.off_old[r + 1] - A.off_old[r];
            if (len == 0) {
                v = A.row_pads[j];
            } else {
                const Fr* src = A.in[j] + A.off_old[r];
                Fr e0 = src[0], e1 = src[1];
                v = fr_add(e0, fr_mul(A.t, fr_sub(e1, e0)));
            }
        } else {
            v = A.col_pads[j];
        }
        A.out[j][r] = v;
    }
}

// all eq levels after the first in ONE launch (one block walks the levels in order): an object over n row variables used to
// issue n - 1 tiny launches / device copies at construction
struct EqLevels {
    uint64_t off[33], size[33];
    uint8_t single[33];  // level is the one-entry table `singles[b]` (all remaining row variables are padding)
    uint32_t n;
    const Fr* singles;
};
__global__ void __launch_bounds__(1024) eq_levels_kernel(Fr* d_eq, const __grid_constant__ EqLevels L) {
    for (uint32_t b = 1; b < L.n; b++) {
        if (L.single[b]) {
            if (threadIdx.x == 0) d_eq[L.off[b]] = L.singles[b];
        } else {
            const Fr* in = d_eq + L.off[b - 1];
            Fr* out = d_eq + L.off[b];
            for (uint64_t i = threadIdx.x; i < L.size[b]; i += blockDim.x) out[i] = fr_add(in[2 * i], in[2 * i + 1]);
        }
        __syncthreads();
    }
}

__global__ void eq_halve_kernel(Fr* out, const Fr* in, uint64_t n_out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = fr_add(in[2 * i], in[2 * i + 1]);
}

// ---- host side ---------------------------------------------------------------------------------------------
static uint32_t log2_ceil_lasso(uint64_t n) {  // liblasso Math::log_2: exact for powers of two, ceil otherwise
    if (n <= 1) return 0;
    uint32_t l = 0;
    while (((uint64_t)1 << l) < n) l++;
    return l;
}

static gkr::FrH host_eq_sum(const gkr::FrH* pt, uint32_t n, uint64_t k) {  // src/utils.rs:265-291
    using namespace gkr::frh;
    if (k >= ((uint64_t)1 << n)) return ONE;
    gkr::FrH mult = ONE, acc = ZERO;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t bit = (uint32_t)(k >> (n - i - 1)) & 1u;
        if (bit) {
            gkr::FrH nm = mul(mult, pt[i]);
            acc = add(acc, sub(mult, nm));
            mult = nm;
        } else {
            mult = mul(mult, sub(ONE, pt[i]));
        }
    }
    return acc;
}

static std::vector<Deg2Block> expand_blocks(const gkr::GateStack& gs) {
    std::vector<Deg2Block> out;
    int in_off = 0, out_off = 0;
    for (size_t p = 0; p < gs.gate.size(); p++) {
        int ni = 0, no = 0;
        gkr::base_gate_io(gs.gate[p], &ni, &no);
        for (int k = 0; k < gs.repeat[p]; k++) {
            auto push = [&](int gate, std::initializer_list<int> idx, int ooff) {
                Deg2Block b;
                b.gate = gate;
                int c = 0;
                for (int x : idx) b.in_idx[c++] = in_off + x;
                for (; c < 6; c++) b.in_idx[c] = 0;
                b.out_off = out_off + ooff;
                b.own_mask = 0;
                out.push_back(b);
            };
            switch (gs.gate[p]) {
                case GKR_GATE_TRI_L1:  // three projective L1 on (a,c), (b,d), (c,d)   twisted_edwards_ops.rs:67-80
                    push(GATE_PRJ_L1, {0, 1, 2, 6, 7, 8}, 0);
                    push(GATE_PRJ_L1, {3, 4, 5, 9, 10, 11}, 4);
                    push(GATE_PRJ_L1, {6, 7, 8, 9, 10, 11}, 8);
                    break;
                case GKR_GATE_AFF_L1_BITCHECK2:
                    push(GATE_AFF_L1, {0, 1, 2, 3}, 0);
                    push(GATE_BITCHECK, {4}, 3);
                    push(GATE_BITCHECK, {5}, 4);
                    break;
                case GKR_GATE_AFF_L1: push(GATE_AFF_L1, {0, 1, 2, 3}, 0); break;
                case GKR_GATE_AFF_L2: push(GATE_AFF_L2, {0, 1, 2}, 0); break;
                case GKR_GATE_AFF_L3: push(GATE_AFF_L3, {0, 1, 2}, 0); break;
                case GKR_GATE_PRJ_L1: push(GATE_PRJ_L1, {0, 1, 2, 3, 4, 5}, 0); break;
                case GKR_GATE_PRJ_L2: push(GATE_PRJ_L2, {0, 1, 2, 3}, 0); break;
                case GKR_GATE_PRJ_L3: push(GATE_PRJ_L3, {0, 1, 2, 3}, 0); break;
                case GKR_GATE_BITCHECK: push(GATE_BITCHECK, {0}, 0); break;
                case GKR_GATE_LOGUP_LAYER: push(GATE_LOGUP_LAYER, {0, 1, 2, 3}, 0); break;
                case GKR_GATE_ADD_INVERSES: push(GATE_ADD_INVERSES, {0, 1}, 0); break;
                default: break;
            }
            in_off += ni;
            out_off += no;
        }
    }
    // every table is folded and written by exactly one block: the first that reads it
    std::vector<char> owned(in_off, 0);
    for (auto& b : out) {
        int ni = 0, no = 0;
        gkr::base_gate_io(b.gate, &ni, &no);
        for (int k = 0; k < ni; k++)
            if (!owned[b.in_idx[k]]) {
                owned[b.in_idx[k]] = 1;
                b.own_mask |= 1 << k;
            }
    }
    return out;
}

// from12 (vecvec_eq.rs:197-216) with the inverse of eq0 supplied (batch-inverted once per object)
static void from12(const gkr::FrH& p1, const gkr::FrH& p2, const gkr::FrH& eq1, const gkr::FrH& eq0_inv, const gkr::FrH& prev_claim, gkr::FrH* evals) {
    using namespace gkr::frh;
    gkr::FrH eq0 = sub(ONE, eq1);
    gkr::FrH eq2 = sub(dbl(eq1), eq0);
    gkr::FrH eq3 = sub(dbl(eq2), eq1);
    gkr::FrH prod1 = mul(p1, eq1);
    gkr::FrH prod0 = sub(prev_claim, prod1);
    gkr::FrH p0 = mul(prod0, eq0_inv);
    gkr::FrH p3 = add(sub(add(dbl(p2), p2), add(dbl(p1), p1)), p0);
    evals[0] = prod0;
    evals[1] = prod1;
    evals[2] = mul(p2, eq2);
    evals[3] = mul(p3, eq3);
}

// Row layout of a ragged bundle over ALL of its sparse rounds: lens[k] = row lengths after k folds (they halve and re-pad
// deterministically, vecvec.rs:432-437), element / pair offsets of every level on the device.  The three sumchecks of one
// bintree layer run over the same row structure, and the rows of layer i+1 (vecvec_map_split halves every row) are the
// level-1 rows of layer i -- so one layout, computed and uploaded once, serves every VecVec object of a proof; an object
// whose rows equal level k of the cached layout uses it from level k on.
struct Deg2Layout {
    gkr_ctx* ctx = nullptr;
    uint32_t nrows = 0;
    bool ragged = false;
    std::vector<std::vector<uint32_t>> lens;  // [level][row] element counts (even when ragged)
    std::vector<uint64_t> totals;             // [level]
    std::vector<char> one_pair;               // [level] every row holds exactly one pair
    uint32_t* d_off = nullptr;                // [levels * (nrows + 1)] ELEMENT offsets
    uint32_t* d_poff = nullptr;               // the same in PAIRS
    ~Deg2Layout() {
        if (d_off) gkr_free_async(d_off, ctx->stream);  // one allocation: d_poff follows d_off
    }
    size_t levels() const { return lens.size(); }
};

static int deg2_layout_get(gkr_ctx* ctx, const std::vector<uint32_t>& lens0, bool ragged, uint32_t n_levels, std::shared_ptr<Deg2Layout>* out,
                           uint32_t* base) {
    std::shared_ptr<Deg2Layout>& cache = ctx->deg2_layout;
    if (cache && cache->ragged == ragged && cache->nrows == lens0.size()) {
        for (size_t k = 0; k + n_levels <= cache->levels(); k++)
            if (cache->lens[k] == lens0) {
                *out = cache;
                *base = (uint32_t)k;
                return GKR_OK;
            }
    }
    auto L = std::make_shared<Deg2Layout>();
    L->ctx = ctx;
    L->nrows = (uint32_t)lens0.size();
    L->ragged = ragged;
    const uint32_t nrows = L->nrows;
    L->lens.resize(n_levels);
    L->totals.assign(n_levels, 0);
    L->one_pair.assign(n_levels, 0);
    L->lens[0] = lens0;
    std::vector<uint32_t> h_off((size_t)2 * n_levels * (nrows + 1));
    uint32_t* h_poff = h_off.data() + (size_t)n_levels * (nrows + 1);
    for (uint32_t b = 0; b < n_levels; b++) {
        if (b > 0) {
            L->lens[b].resize(nrows);
            const uint32_t* prev = L->lens[b - 1].data();
            uint32_t* cur = L->lens[b].data();
            for (uint32_t r = 0; r < nrows; r++) {
                uint32_t h = prev[r] / 2;
                cur[r] = ragged ? ((h + 1) & ~1u) : h;
            }
        }
        const uint32_t* cur = L->lens[b].data();
        uint32_t* o = h_off.data() + (size_t)b * (nrows + 1);
        uint32_t* po = h_poff + (size_t)b * (nrows + 1);
        uint64_t acc = 0;
        bool all_two = true;
        for (uint32_t r = 0; r < nrows; r++) {
            o[r] = (uint32_t)acc;
            po[r] = (uint32_t)(acc / 2);
            acc += cur[r];
            all_two &= cur[r] == 2;
        }
        L->one_pair[b] = all_two ? 1 : 0;
        if (acc >= ((uint64_t)1 << 32)) return ctx->fail(GKR_ERR_UNSUPPORTED, "more than 2^32 elements per polynomial");
        o[nrows] = (uint32_t)acc;
        po[nrows] = (uint32_t)(acc / 2);
        L->totals[b] = acc;
    }
    GKR_CUDA_OK(ctx, gkr_malloc_async(&L->d_off, sizeof(uint32_t) * h_off.size(), ctx->stream));
    L->d_poff = L->d_off + (size_t)n_levels * (nrows + 1);
    int rc = gkr_stage_upload(ctx, L->d_off, h_off.data(), sizeof(uint32_t) * h_off.size());
    if (rc) return rc;
    if (ragged) cache = L;  // single-row dense objects are trivial: not worth evicting a ragged layout for
    *out = L;
    *base = 0;
    return GKR_OK;
}

// Shared machinery of the two Deg2 objects: P ragged (or single-row dense) tables, per-round offsets and eq levels.
class Deg2SO : public gkr_so {
   public:
    bool is_vecvec = false;
    gkr::GateStack gs;
    int tail_gate = -1;  // public gate id for the dense tail after bind_into_dense
    int P = 0;
    uint32_t n_vars = 0, col = 0, row_logsize = 0, nrows = 0;
    uint32_t n_sparse = 0;   // number of rounds run by the Deg2 kernels (dense object: all n_vars)
    uint32_t round_idx = 0;
    std::vector<gkr::FrH> point, gamma_pows, eq0_inv;  // eq0_inv[i] = 1 / (1 - point[i])
    std::vector<gkr::FrH> row_pads, col_pads;
    gkr::FrH claim_, multiplier, padG, colpadG, col_tail;
    bool has_col_tail = false;
    // per-round layout (shared, see Deg2Layout): round b of this object is level layout_base + b
    std::vector<uint32_t> lens0;              // row lengths at construction
    std::shared_ptr<Deg2Layout> layout;
    const uint64_t* totals = nullptr;         // [round] total elements
    const char* one_pair = nullptr;           // [round] every row holds exactly one pair
    uint32_t* d_off = nullptr;                // [(n_sparse + 1) * (nrows + 1)] ELEMENT offsets per round
    uint32_t* d_poff = nullptr;               // same in PAIRS
    // eq levels
    Fr* d_eq = nullptr;
    std::vector<uint64_t> eq_off;  // offset of the level used in round b
    Fr* d_rowcoef = nullptr;
    Fr* d_pt_row = nullptr;  // device copy of the row variables (for the pad kernel)
    uint32_t m_row = 0;      // number of row variables excluding the binding one at round 0
    // data
    unsigned char* d_params = nullptr;  // one allocation + one staged upload for every small parameter array
    Fr* slab[2] = {nullptr, nullptr};
    const Fr** d_tabs[3] = {nullptr, nullptr, nullptr};  // device pointer arrays: [0] inputs, [1]/[2] ping-pong
    std::vector<const Fr*> h_sets[3];  // host copies of the three pointer arrays
    int cur_set = 0;
    Deg2Block* d_blocks = nullptr;
    int n_blocks = 0;
    int uniform_gate = -1;  // base gate shared by every gate block of the stack (-1: mixed -> generic kernel)
    Fr* d_gammas = nullptr;
    Fr* d_pads = nullptr;  // [2P]: row pads then col pads
    bool cached = false, sums_pending = false;
    gkr::FrH evals[4];
    int slot = -1;
    gkr_so* dense = nullptr;  // after bind_into_dense
    std::vector<gkr_table*> dense_tables;

    ~Deg2SO() override {
        if (pre_active) ctx->post_mailbox(slot, pre_mbox_seq, 2, nullptr);  // a spinning pre-launched kernel exits without publishing
        delete dense;
        for (auto* t : dense_tables) gkr_table_free(t);
        cudaStream_t s = ctx->stream;
        if (d_params) gkr_free_async(d_params, s);  // offsets, gate blocks, gammas, pads, points, pointer arrays
        if (d_eq) gkr_free_async(d_eq, s);
        if (d_rowcoef) gkr_free_async(d_rowcoef, s);
        for (int i = 0; i < 2; i++) if (slab[i]) gkr_free_async(slab[i], s);
        if (slot >= 0) gkr_result_slot_release(ctx, slot);
    }

    uint32_t binding_idx() const { return n_vars - 1 - round_idx; }
    uint32_t pending_blocks = 0, pending_seq = 0;
    // pre-launched next round (set_prelaunch; common.cuh, GkrMailbox): enqueued while the current round runs, released by bind(t)
    bool allow_prelaunch = false, pre_active = false;
    uint32_t pre_mbox_seq = 0, pre_blocks = 0, pre_slot_seq = 0;
    int pre_dst_set = 0;
    void set_prelaunch(bool on) override {
        allow_prelaunch = on;
        if (dense) dense->set_prelaunch(on);
    }
    bool round_is_compact(uint32_t b) const { return (totals[b] / 2) * (uint64_t)n_blocks <= ctx->deg2_compact_max; }

    // launches the round kernel for round `b` (= round_idx at call time); with fold_t != nullptr it first folds round b-1.
    // prelaunch: the fused kernel is enqueued BEFORE the challenge exists and takes it from the mailbox (fold_t is ignored).
    int launch_round(uint32_t b, const gkr::FrH* fold_t, bool prelaunch = false) {
        const uint64_t n_pairs = totals[b] / 2;
        Deg2RoundArgs a;
        int dst_set = cur_set;
        if (fold_t || prelaunch) {
            dst_set = (cur_set == 1) ? 2 : 1;
            a.in = d_tabs[cur_set];
            a.out = (Fr* const*)d_tabs[dst_set];
            a.off_old = d_off + (size_t)(b - 1) * (nrows + 1);
            a.fold = 1;
            a.t = fr_from_host(prelaunch ? gkr::frh::ZERO : *fold_t);
            if (prelaunch) a.mbox = ctx->next_mailbox(slot);
        } else {
            a.in = d_tabs[cur_set];
            a.out = nullptr;
            a.off_old = nullptr;
            a.fold = 0;
            a.t = fr_from_host(gkr::frh::ZERO);
        }
        a.blocks = d_blocks;
        a.gammas = d_gammas;
        a.pair_off = is_vecvec ? d_poff + (size_t)b * (nrows + 1) : nullptr;
        a.nrows = nrows;
        a.eq = d_eq + eq_off[b];
        a.rowcoef = d_rowcoef;
        a.n_pairs = n_pairs;
        a.row_pads = d_pads;
        a.pt = d_pt_row;
        a.n_pt = m_row >= b ? m_row - b : 0;
        a.do_pad = is_vecvec ? 1 : 0;
        a.one_pair_rows = (is_vecvec && one_pair[b]) ? 1 : 0;
        a.o = ctx->round_out(slot);
        const uint64_t want = std::max<uint64_t>(1, (2 * n_pairs + GKR_REDUCE_THREADS - 1) / GKR_REDUCE_THREADS);  // two lanes per pair
        // latency flavour for rounds that fit in about one wave of blocks, throughput flavour above
        const bool compact = n_pairs * (uint64_t)n_blocks <= ctx->deg2_compact_max;
        // compact + ragged: extra blocks at the tail of every slice for the padding term, one row per thread (only the y == 0
        // ones work), so that T overlaps the gate evaluations instead of following them
        const uint64_t pad_blocks = (is_vecvec && compact) ? std::min<uint64_t>(128, (nrows + GKR_REDUCE_THREADS - 1) / GKR_REDUCE_THREADS) : 0;
        const uint64_t cap = std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)ctx->num_sms * 3, GKR_MAX_BLOCKS - 3 * 128) / n_blocks);
        uint64_t work_blocks = std::min(want, cap);
        if (is_vecvec && !compact) work_blocks = std::max<uint64_t>(work_blocks, std::min<uint64_t>(cap, (nrows + GKR_REDUCE_THREADS - 1) / GKR_REDUCE_THREADS));
        a.work_blocks_x = (uint32_t)work_blocks;
        dim3 grid((unsigned)(work_blocks + pad_blocks), (unsigned)n_blocks);
        unsigned threads = GKR_REDUCE_THREADS;
        if (grid.x == 1) threads = (unsigned)std::max<uint64_t>(32, std::min<uint64_t>(GKR_REDUCE_THREADS, (2 * n_pairs + 31) / 32 * 32));
        if (compact) {
            gkr_launch_deg2_round_compact(uniform_gate, a, grid, threads, ctx->stream);
        } else {
            deg2_inline::launch_deg2_round(uniform_gate, a, grid, threads, ctx->stream);
        }
        ctx->launches++;
        GKR_CUDA_OK(ctx, cudaGetLastError());
        if (prelaunch) {  // takes effect when bind() releases it
            pre_active = true;
            pre_mbox_seq = a.mbox.seq;
            pre_slot_seq = a.o.seq;
            pre_blocks = grid.x * grid.y;
            pre_dst_set = dst_set;
        } else {
            pending_blocks = grid.x * grid.y;
            pending_seq = a.o.seq;
            cur_set = dst_set;
        }
        return GKR_OK;
    }

    // the two eq-weighted totals of this round (vecvec_eq.rs:302-388 up to the call of from12): linear in the rows, so the row
    // shards of one object (gkr_so_create_deg2_vecvec_shard) add theirs up
    int round_totals(gkr::FrH* total1, gkr::FrH* total2) {
        if (round_idx >= n_sparse) return ctx->fail(GKR_ERR_PROTOCOL, "unipoly: the protocol has already ended");
        if (cached) return ctx->fail(GKR_ERR_PROTOCOL, "unipoly called twice in a round");  // dense_eq.rs:109-111
        using namespace gkr::frh;
        if (!sums_pending) {
            int rc = launch_round(round_idx, nullptr);
            if (rc) return rc;
        }
        // the next fused round (small rounds only: the latency-bound ones) queues up behind the kernel we are about to wait for
        if (allow_prelaunch && ctx->prelaunch && !pre_active && round_idx + 1 < n_sparse && round_is_compact(round_idx + 1)) {
            int rc = launch_round(round_idx + 1, nullptr, true);
            if (rc) return rc;
        }
        gkr::FrH r[3];
        ctx->wait_kind = is_vecvec ? 2 : 1;
        ctx->wait_log = 0;
        while (((uint64_t)2 << ctx->wait_log) <= totals[round_idx]) ctx->wait_log++;
        int rcw = gkr_slot_wait_seq(ctx, slot, pending_seq, pending_blocks, 3, r);
        if (rcw) return rcw;
        sums_pending = false;
        gkr::FrH padterm = ZERO;
        if (is_vecvec) {
            padterm = mul(padG, r[2]);
            if (has_col_tail) padterm = add(padterm, mul(colpadG, col_tail));
        }  // dense object: tables are full, trailing_sum (dense_eq.rs:141) is zero
        *total1 = mul(add(r[0], padterm), multiplier);
        *total2 = mul(add(r[1], padterm), multiplier);
        return GKR_OK;
    }

    int unipoly(gkr::FrH* out, uint32_t* n_evals) override {
        if (dense) return dense->unipoly(out, n_evals);
        gkr::FrH total1, total2;
        int rc = round_totals(&total1, &total2);
        if (rc) return rc;
        const uint32_t b = binding_idx();
        from12(total1, total2, point[b], eq0_inv[b], claim_, evals);
        cached = true;
        for (int i = 0; i < 4; i++) out[i] = evals[i];
        if (n_evals) *n_evals = 4;
        return GKR_OK;
    }

    int partial_sums(gkr::FrH* out, uint32_t* n) override {
        if (dense) return dense->partial_sums(out, n);
        int rc = round_totals(&out[0], &out[1]);
        if (rc) return rc;
        // the round polynomial belongs to the driver that adds the shards up: this object's own claim is not tracked from here on
        for (int i = 0; i < 4; i++) evals[i] = gkr::frh::ZERO;
        cached = true;
        *n = 2;
        return GKR_OK;
    }

    // plain fold of the current round into the next layout (last round of the dense object: nothing left to evaluate)
    int fold_to_next(const gkr::FrH& t) {
        const uint32_t b = round_idx;
        int dst_set = (cur_set == 1) ? 2 : 1;
        VvFoldArgs f;
        f.in = d_tabs[cur_set];
        f.out = (Fr* const*)d_tabs[dst_set];
        f.off_old = d_off + (size_t)b * (nrows + 1);
        f.off_new = d_off + (size_t)(b + 1) * (nrows + 1);
        f.nrows = nrows;
        f.n_new = totals[b + 1];
        f.t = fr_from_host(t);
        f.row_pads = d_pads;
        if (f.n_new > 0) {
            dim3 grid((unsigned)std::max<uint64_t>(1, std::min<uint64_t>((f.n_new + 255) / 256, (uint64_t)ctx->num_sms * 4)), (unsigned)P);
            vv_fold_kernel<<<grid, 256, 0, ctx->stream>>>(f);
            ctx->launches++;
            GKR_CUDA_OK(ctx, cudaGetLastError());
        }
        cur_set = dst_set;
        return GKR_OK;
    }

    int bind(const gkr::FrH& t) override {
        if (dense) return dense->bind(t);
        if (round_idx >= n_sparse) return ctx->fail(GKR_ERR_PROTOCOL, "bind: the protocol has already ended");
        if (!cached) return ctx->fail(GKR_ERR_PROTOCOL, "bind: should evaluate unipoly before binding");
        if (!frh_canonical(t)) return ctx->fail(GKR_ERR_ARG, "bind: challenge is not canonical");
        using namespace gkr::frh;
        const uint32_t b = binding_idx();
        gkr::FrH new_claim = interpolate_eval(evals, 4, t);
        gkr::FrH new_mult = mul(multiplier, eq1(point[b], t));
        if (is_vecvec && round_idx + 1 == n_sparse) return bind_into_dense(t, new_claim, new_mult);
        int rc = GKR_OK;
        if (pre_active) {
            pre_active = false;
            const gkr::FrH t_plain = mul(t, gkr::FrH{{1, 0, 0, 0}});
            if (t_plain.v[2] == 0 && t_plain.v[3] == 0 && !ctx->mailbox_timed_out(slot, pre_mbox_seq)) {  // release the queued kernel
                const uint32_t tw[4] = {(uint32_t)t_plain.v[0], (uint32_t)(t_plain.v[0] >> 32), (uint32_t)t_plain.v[1], (uint32_t)(t_plain.v[1] >> 32)};
                ctx->post_mailbox(slot, pre_mbox_seq, 1, tw);
                pending_blocks = pre_blocks;
                pending_seq = pre_slot_seq;
                cur_set = pre_dst_set;
                sums_pending = true;
            } else {  // a full-width challenge, or the launch gave up (the host was held up for seconds): cancel, ordinary launch
                if (ctx->mailbox_timed_out(slot, pre_mbox_seq)) ctx->prelaunch = false;  // see DenseSO::bind
                ctx->post_mailbox(slot, pre_mbox_seq, 2, nullptr);
                rc = launch_round(round_idx + 1, &t);
                sums_pending = true;
            }
        } else if (round_idx + 1 < n_sparse) {
            rc = launch_round(round_idx + 1, &t);  // fused: fold round b, evaluate round b+1
            sums_pending = true;
        } else {
            rc = fold_to_next(t);
        }
        if (rc) return rc;
        multiplier = new_mult;
        claim_ = new_claim;
        cached = false;
        round_idx++;
        return GKR_OK;
    }

    int bind_into_dense(const gkr::FrH& t, const gkr::FrH& new_claim, const gkr::FrH& new_mult);

    int final_evals(gkr::FrH* out) override {
        if (dense) return dense->final_evals(out);
        if (is_vecvec) return ctx->fail(GKR_ERR_PROTOCOL, "final_evals: sparse stage has no final evals (vecvec_eq.rs:390-393)");
        if (round_idx != n_sparse) return ctx->fail(GKR_ERR_PROTOCOL, "final_evals: can only be called after the last round");
        return gkr_fetch_firsts_dev(ctx, slot, d_tabs[cur_set], P, out);
    }

    gkr::FrH claim() const override { return dense ? dense->claim() : claim_; }
    uint32_t degree() const override { return 3; }
    // number of final evaluations: the VecVec object ends on the dense tail, which carries the eq table as well
    uint32_t num_polys() const override { return dense ? dense->num_polys() : (uint32_t)(is_vecvec ? P + 1 : P); }
    uint32_t round() const override { return dense ? n_sparse + dense->round() : round_idx; }

    // common construction once lens[0], data pointers, point, gammas are known
    int setup(const std::vector<const Fr*>& inputs);
};

int Deg2SO::setup(const std::vector<const Fr*>& inputs) {
    using namespace gkr::frh;
    cudaStream_t s = ctx->stream;
    // batch inversion of (1 - point[i])
    {
        const size_t n = point.size();
        eq0_inv.assign(n, ZERO);
        std::vector<gkr::FrH> pref(n + 1, ONE);
        for (size_t i = 0; i < n; i++) {
            gkr::FrH e0 = sub(ONE, point[i]);
            if (is_zero(e0)) return ctx->fail(GKR_ERR_ARG, "point coordinate equal to one: eq0 is not invertible (from12 would panic)");
            pref[i + 1] = mul(pref[i], e0);
        }
        gkr::FrH inv_all = inverse(pref[n]);
        for (size_t i = n; i-- > 0;) {
            eq0_inv[i] = mul(inv_all, pref[i]);
            inv_all = mul(inv_all, sub(ONE, point[i]));
        }
    }
    // per-round row lengths and offsets: shared with the other objects over the same rows
    {
        uint32_t base = 0;
        int rc = deg2_layout_get(ctx, lens0, is_vecvec, n_sparse + 1, &layout, &base);
        if (rc) return rc;
        totals = layout->totals.data() + base;
        one_pair = layout->one_pair.data() + base;
        d_off = layout->d_off + (size_t)base * (nrows + 1);
        d_poff = layout->d_poff + (size_t)base * (nrows + 1);
    }
    // gate program, gammas, pads
    std::vector<Deg2Block> blocks = expand_blocks(gs);
    n_blocks = (int)blocks.size();
    uniform_gate = blocks.empty() ? -1 : blocks[0].gate;
    for (const auto& b : blocks)
        if (b.gate != uniform_gate) uniform_gate = -1;
    { const char* v = getenv("GKR_DEG2_GENERIC"); if (v && v[0] == '1') uniform_gate = -1; }  // test hook: force the generic kernel
    std::vector<Fr> g(gs.n_outs);
    for (int i = 0; i < gs.n_outs; i++) g[i] = fr_from_host(gamma_pows[i]);
    std::vector<Fr> pads(2 * P);
    for (int j = 0; j < P; j++) {
        pads[j] = fr_from_host(row_pads[j]);
        pads[P + j] = fr_from_host(col_pads[j]);
    }
    // pad_results / col_pad_results folded with gamma (vecvec_eq.rs:309-315, 372-378)
    {
        std::vector<gkr::FrH> o(gs.n_outs);
        gs.eval(row_pads.data(), o.data());
        padG = o[0];
        for (int i = 1; i < gs.n_outs; i++) padG = add(padG, mul(o[i], gamma_pows[i]));
        gs.eval(col_pads.data(), o.data());
        colpadG = o[0];
        for (int i = 1; i < gs.n_outs; i++) colpadG = add(colpadG, mul(o[i], gamma_pows[i]));
    }

    // eq levels.  Row variables: point[col .. n_vars-1); the last one (binding variable of round 0) is excluded.
    m_row = (n_vars - col) - 1;  // == row_logsize - 1
    uint64_t max_len = 0;
    for (uint32_t r = 0; r < nrows; r++) max_len = std::max<uint64_t>(max_len, lens0[r]);
    uint32_t max_seg_log = log2_ceil_lasso(max_len);
    if (!is_vecvec) max_seg_log = n_vars;  // dense tables span all variables
    // number of leading row variables over which every row is padding (EQPolyPointParts::padded_vars_range)
    uint32_t seg_idx = n_vars - std::min(max_seg_log, n_vars);
    uint32_t pad_lo = col, pad_hi = std::min(seg_idx, n_vars - 1);
    uint32_t npad = pad_hi > pad_lo ? pad_hi - pad_lo : 0;
    const gkr::FrH* pt_row = point.data() + col;
    std::vector<gkr::FrH> prefix(m_row + 1, ONE);  // prefix[i] = prod_{k<i} (1 - pt_row[k])
    for (uint32_t i = 0; i < m_row; i++) prefix[i + 1] = mul(prefix[i], sub(ONE, pt_row[i]));
    eq_off.assign(n_sparse + 1, 0);
    uint64_t eq_total = 0;
    std::vector<uint64_t> lvl_size(n_sparse + 1, 1);
    for (uint32_t b = 0; b < n_sparse; b++) {
        uint32_t lvl = m_row >= b ? m_row - b : 0;  // number of variables in this round's eq table
        lvl_size[b] = lvl > npad ? ((uint64_t)1 << (lvl - npad)) : 1;
        eq_off[b] = eq_total;
        eq_total += lvl_size[b];
    }
    std::vector<Fr> p_row(std::max<uint32_t>(m_row, 1), fr_from_host(ZERO)), p_col(std::max<uint32_t>(col, 1), fr_from_host(ZERO));
    for (uint32_t i = 0; i < m_row; i++) p_row[i] = fr_from_host(pt_row[i]);
    for (uint32_t i = 0; i < col; i++) p_col[i] = fr_from_host(point[i]);
    std::vector<Fr> singles(n_sparse + 1, fr_from_host(ZERO));  // one-entry eq levels (all remaining row variables are padding)
    for (uint32_t b = 0; b < n_sparse; b++) {
        uint32_t lvl = m_row >= b ? m_row - b : 0;
        if (lvl <= npad) singles[b] = fr_from_host(prefix[lvl]);
    }

    // data: ping-pong slabs sized for round 1 and round 2, pointer arrays
    uint64_t sz1 = n_sparse >= 1 ? totals[1] : 0, sz2 = n_sparse >= 2 ? totals[2] : 0;
    std::vector<const Fr*> p1(P), p2(P);
    GKR_CUDA_OK(ctx, gkr_malloc_async(&slab[0], sizeof(Fr) * std::max<uint64_t>(sz1 * P, 1), s));
    GKR_CUDA_OK(ctx, gkr_malloc_async(&slab[1], sizeof(Fr) * std::max<uint64_t>(sz2 * P, 1), s));
    for (int j = 0; j < P; j++) {
        p1[j] = slab[0] + (size_t)j * sz1;
        p2[j] = slab[1] + (size_t)j * sz2;
    }
    h_sets[0] = inputs;
    h_sets[1] = p1;
    h_sets[2] = p2;

    // ONE device allocation and ONE staged upload for all of the above
    std::vector<unsigned char> arena;
    auto put = [&](const void* src, size_t n) {
        size_t off = (arena.size() + 31) & ~(size_t)31;
        arena.resize(off + std::max<size_t>(n, 1));
        if (n) std::memcpy(arena.data() + off, src, n);
        return off;
    };
    const size_t o_blocks = put(blocks.data(), sizeof(Deg2Block) * blocks.size());
    const size_t o_g = put(g.data(), sizeof(Fr) * g.size());
    const size_t o_pads = put(pads.data(), sizeof(Fr) * pads.size());
    const size_t o_prow = put(p_row.data(), sizeof(Fr) * p_row.size());
    const size_t o_pcol = put(p_col.data(), sizeof(Fr) * p_col.size());
    const size_t o_single = put(singles.data(), sizeof(Fr) * singles.size());
    const size_t o_t0 = put(inputs.data(), sizeof(Fr*) * P);
    const size_t o_t1 = put(p1.data(), sizeof(Fr*) * P);
    const size_t o_t2 = put(p2.data(), sizeof(Fr*) * P);
    GKR_CUDA_OK(ctx, gkr_malloc_async(&d_params, arena.size(), s));
    {
        int rc = gkr_stage_upload(ctx, d_params, arena.data(), arena.size());
        if (rc) return rc;
    }
    d_blocks = (Deg2Block*)(d_params + o_blocks);
    d_gammas = (Fr*)(d_params + o_g);
    d_pads = (Fr*)(d_params + o_pads);
    d_pt_row = (Fr*)(d_params + o_prow);
    Fr* d_pt_col = (Fr*)(d_params + o_pcol);
    const Fr* d_single = (const Fr*)(d_params + o_single);
    d_tabs[0] = (const Fr**)(d_params + o_t0);
    d_tabs[1] = (const Fr**)(d_params + o_t1);
    d_tabs[2] = (const Fr**)(d_params + o_t2);

    GKR_CUDA_OK(ctx, gkr_malloc_async(&d_eq, sizeof(Fr) * std::max<uint64_t>(eq_total, 1), s));
    const bool fused_levels = n_sparse >= 2 && n_sparse <= 33 && lvl_size[1] <= 16384;
    for (uint32_t b = 0; b < n_sparse; b++) {
        if (b >= 1 && fused_levels) break;
        uint32_t lvl = m_row >= b ? m_row - b : 0;
        if (lvl <= npad) {
            GKR_CUDA_OK(ctx, cudaMemcpyAsync(d_eq + eq_off[b], d_single + b, sizeof(Fr), cudaMemcpyDeviceToDevice, s));
        } else if (b == 0) {
            int rc = gkr_eq_build_device(ctx, d_pt_row + npad, lvl - npad, fr_from_host(prefix[npad]), d_eq + eq_off[0]);
            if (rc) return rc;
        } else {
            uint64_t n_out = lvl_size[b];
            unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n_out + 255) / 256, (uint64_t)ctx->num_sms * 4));
            eq_halve_kernel<<<grid, 256, 0, s>>>(d_eq + eq_off[b], d_eq + eq_off[b - 1], n_out);
            ctx->launches++;
            GKR_CUDA_OK(ctx, cudaGetLastError());
        }
    }
    if (fused_levels) {
        EqLevels L;
        L.n = n_sparse;
        L.singles = d_single;
        for (uint32_t b = 0; b < 33; b++) {
            const uint32_t lvl = (b < n_sparse && m_row >= b) ? m_row - b : 0;
            L.off[b] = b < n_sparse ? eq_off[b] : 0;
            L.size[b] = b < n_sparse ? lvl_size[b] : 0;
            L.single[b] = (b < n_sparse && lvl <= npad) ? 1 : 0;
        }
        eq_levels_kernel<<<1, 1024, 0, s>>>(d_eq, L);
        ctx->launches++;
        GKR_CUDA_OK(ctx, cudaGetLastError());
    }
    if (is_vecvec) {
        GKR_CUDA_OK(ctx, gkr_malloc_async(&d_rowcoef, sizeof(Fr) << col, s));
        int rc = gkr_eq_build_device(ctx, d_pt_col, col, fr_from_host(ONE), d_rowcoef);
        if (rc) return rc;
        has_col_tail = nrows < ((uint64_t)1 << col);
        col_tail = sub(ONE, host_eq_sum(point.data(), col, nrows));  // row_eq_coefs_tail_sums[row_count]
    }
    cur_set = 0;
    multiplier = ONE;
    slot = gkr_result_slot_acquire(ctx);
    if (slot < 0) return ctx->fail(GKR_ERR_UNSUPPORTED, "too many live sumcheck objects");
    return GKR_OK;
}

int Deg2SO::bind_into_dense(const gkr::FrH& t, const gkr::FrH& new_claim, const gkr::FrH& new_mult) {
    using namespace gkr::frh;
    cudaStream_t s = ctx->stream;
    const uint64_t n_out = (uint64_t)1 << col;
    dense_tables.assign(P + 1, nullptr);
    std::vector<Fr*> outs(P);
    for (int j = 0; j < P; j++) {
        int rc = gkr_table_alloc(ctx, n_out, &dense_tables[j]);
        if (rc) return rc;
        outs[j] = dense_tables[j]->d;
    }
    Fr** d_outs = nullptr;
    GKR_CUDA_OK(ctx, gkr_malloc_async(&d_outs, sizeof(Fr*) * P, s));
    {
        int rc = gkr_stage_upload(ctx, d_outs, outs.data(), sizeof(Fr*) * P);
        if (rc) return rc;
    }
    VvToDenseArgs a;
    a.in = d_tabs[cur_set];
    a.out = d_outs;
    a.off_old = d_off + (size_t)round_idx * (nrows + 1);
    a.nrows = nrows;
    a.n_out = n_out;
    a.t = fr_from_host(t);
    a.row_pads = d_pads;
    a.col_pads = d_pads + P;
    dim3 grid((unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n_out + 255) / 256, (uint64_t)ctx->num_sms * 4)), (unsigned)P);
    vv_to_dense_kernel<<<grid, 256, 0, s>>>(a);
    ctx->launches++;
    GKR_CUDA_OK(ctx, cudaGetLastError());
    gkr_free_async(d_outs, s);
    // eq table over the vertical variables scaled by the multiplier of all bound variables (vecvec_eq.rs:176-179)
    std::vector<uint64_t> pt(4 * std::max<uint32_t>(col, 1));
    for (uint32_t i = 0; i < col; i++) frh_to_limbs(point[i], pt.data() + 4 * i);
    uint64_t m[4];
    frh_to_limbs(new_mult, m);
    int rc = gkr_eq_table(ctx, pt.data(), col, m, &dense_tables[P]);
    if (rc) return rc;
    // DenseSumcheckObjectSO over EqWrapper(GammaWrapper(func, gamma)) (vecvec_eq.rs:182-189)
    std::vector<gkr::FrH> consts(gamma_pows.begin(), gamma_pows.end());
    rc = gkr_make_dense_so(ctx, GKR_SO_EQ_GAMMA, tail_gate, 0, consts.data(), (uint32_t)std::min<size_t>(consts.size(), GKR_MAX_GATE_CONSTS),
                           dense_tables.data(), (uint32_t)P + 1, col, new_claim, &dense);
    if (rc) return rc;
    dense->set_prelaunch(allow_prelaunch);
    multiplier = new_mult;
    claim_ = new_claim;
    cached = false;
    round_idx++;
    return GKR_OK;
}

static int load_stack(gkr_ctx* ctx, const int* part_gate, const uint32_t* part_repeat, uint32_t n_parts, gkr::GateStack* gs) {
    if (!part_gate || !part_repeat || !gs->init(part_gate, part_repeat, n_parts)) return ctx->fail(GKR_ERR_ARG, "invalid gate stack");
    return GKR_OK;
}

// DenseDeg2SumcheckObjectSO::new   dense_eq.rs:75-95
extern "C" int gkr_so_create_deg2_dense(gkr_ctx* ctx, const int* part_gate, const uint32_t* part_repeat, uint32_t n_parts,
                                        gkr_table* const* tables, uint32_t n_polys, const uint64_t* gamma_pows,
                                        const uint64_t claim[4], const uint64_t* point, uint32_t num_vars, gkr_so** out) {
    if (!ctx) return GKR_ERR_ARG;
    if (!tables || !gamma_pows || !claim || !point || !out) return ctx->fail(GKR_ERR_ARG, "null argument");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    Deg2SO* so = new Deg2SO();
    so->ctx = ctx;
    int rc = load_stack(ctx, part_gate, part_repeat, n_parts, &so->gs);
    if (rc) { delete so; return rc; }
    if ((int)n_polys != so->gs.n_ins) { delete so; return ctx->fail(GKR_ERR_ARG, "number of tables != f.n_ins()"); }
    if (num_vars == 0 || num_vars >= 32) { delete so; return ctx->fail(GKR_ERR_ARG, "bad num_vars"); }
    for (uint32_t j = 0; j < n_polys; j++) {
        if (!tables[j] || tables[j]->n != ((uint64_t)1 << num_vars)) {
            delete so;
            // the reference accepts shorter tables only in its non-`parallel` build (dense.rs:39-61); the README build
            // (`--features parallel`) indexes out of bounds on them, so full tables are required here
            return ctx->fail(GKR_ERR_UNSUPPORTED, "Deg2 dense object: every table must have 1 << num_vars entries");
        }
    }
    so->is_vecvec = false;
    so->P = (int)n_polys;
    so->n_vars = num_vars;
    so->col = 0;
    so->row_logsize = num_vars;
    so->nrows = 1;
    so->n_sparse = num_vars;
    so->point.resize(num_vars);
    for (uint32_t i = 0; i < num_vars; i++) so->point[i] = frh_from_limbs(point + 4 * i);
    so->gamma_pows.resize(so->gs.n_outs);
    for (int i = 0; i < so->gs.n_outs; i++) so->gamma_pows[i] = frh_from_limbs(gamma_pows + 4 * i);
    so->claim_ = frh_from_limbs(claim);
    so->row_pads.assign(n_polys, gkr::frh::ZERO);
    so->col_pads.assign(n_polys, gkr::frh::ZERO);
    so->lens0.assign(1, (uint32_t)((uint64_t)1 << num_vars));
    std::vector<const Fr*> in(n_polys);
    for (uint32_t j = 0; j < n_polys; j++) in[j] = tables[j]->d;
    rc = so->setup(in);
    if (rc) { delete so; return rc; }
    *out = so;
    return GKR_OK;
}

// VecVecDeg2SumcheckObjectSO::new   vecvec_eq.rs:94-118
// weight: multiplier the object starts from (ONE; a row shard starts from eq(top column coordinates, shard index))
static int make_vecvec_so(gkr_ctx* ctx, int gate, gkr_vecvec* const* polys, uint32_t n_polys, const uint64_t* gamma_pows,
                          const uint64_t claim[4], const uint64_t* point, uint32_t num_vars, uint32_t col_logsize, const gkr::FrH& weight,
                          gkr_so** out) {
    if (!ctx) return GKR_ERR_ARG;
    if (!polys || !gamma_pows || !claim || !point || !out) return ctx->fail(GKR_ERR_ARG, "null argument");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    Deg2SO* so = new Deg2SO();
    so->ctx = ctx;
    uint32_t one = 1;
    int rc = load_stack(ctx, &gate, &one, 1, &so->gs);
    if (rc) { delete so; return rc; }
    if ((int)n_polys != so->gs.n_ins) { delete so; return ctx->fail(GKR_ERR_ARG, "number of polynomials != f.n_ins()"); }
    const gkr_vecvec* p0 = polys[0];
    if (!p0 || p0->col_logsize != col_logsize || p0->row_logsize + col_logsize != num_vars || p0->row_logsize == 0) {
        delete so;
        return ctx->fail(GKR_ERR_ARG, "row_logsize + col_logsize must equal the point length");
    }
    for (uint32_t j = 0; j < n_polys; j++) {
        if (!polys[j] || polys[j]->row_len != p0->row_len || polys[j]->row_logsize != p0->row_logsize || polys[j]->col_logsize != col_logsize) {
            delete so;
            return ctx->fail(GKR_ERR_ARG, "all polynomials of a bundle must share the row structure");
        }
    }
    if (p0->row_len.empty()) { delete so; return ctx->fail(GKR_ERR_ARG, "empty polynomial (reference: max() of empty iterator panics)"); }
    so->is_vecvec = true;
    so->tail_gate = gate;
    so->P = (int)n_polys;
    so->n_vars = num_vars;
    so->col = col_logsize;
    so->row_logsize = p0->row_logsize;
    so->nrows = (uint32_t)p0->row_len.size();
    so->n_sparse = p0->row_logsize;
    so->point.resize(num_vars);
    for (uint32_t i = 0; i < num_vars; i++) so->point[i] = frh_from_limbs(point + 4 * i);
    so->gamma_pows.resize(std::max(so->gs.n_outs, 2));
    for (size_t i = 0; i < so->gamma_pows.size(); i++) so->gamma_pows[i] = frh_from_limbs(gamma_pows + 4 * i);
    so->claim_ = frh_from_limbs(claim);
    so->row_pads.resize(n_polys);
    so->col_pads.resize(n_polys);
    std::vector<const Fr*> in(n_polys);
    for (uint32_t j = 0; j < n_polys; j++) {
        so->row_pads[j] = polys[j]->row_pad;
        so->col_pads[j] = polys[j]->col_pad;
        in[j] = polys[j]->d;
    }
    so->lens0 = p0->row_len;
    rc = so->setup(in);
    if (rc) { delete so; return rc; }
    so->multiplier = weight;
    *out = so;
    return GKR_OK;
}

extern "C" int gkr_so_create_deg2_vecvec(gkr_ctx* ctx, int gate, gkr_vecvec* const* polys, uint32_t n_polys, const uint64_t* gamma_pows,
                                         const uint64_t claim[4], const uint64_t* point, uint32_t num_vars, uint32_t col_logsize,
                                         gkr_so** out) {
    return make_vecvec_so(ctx, gate, polys, n_polys, gamma_pows, claim, point, num_vars, col_logsize, gkr::frh::ONE, out);
}

// ---- VecVec sumcheck sharded by bucket rows ---------------------------------------------------------------------
// The column (bucket-index) variables are the MOST significant ones and are only bound in the dense tail (vecvec.rs:156-159), so
// the rows split by the top log2(G) bits of the row index: shard g holds the rows [g R / G, (g + 1) R / G) as a VecVec bundle
// with col_logsize - log2(G) column variables.  Its row multipliers are eq(point_col_low, local row); the missing factor
// e_g = eq(point[0 .. log2 G), g) is what the shard's multiplier starts from, so everything it reports -- the two totals of a
// sparse round, the eq table of its slice of bind_into_dense -- is its exact share of the whole object's value.
//   point: the WHOLE object's point (num_vars = row_logsize + col_logsize coordinates); polys: this shard's rows.
extern "C" int gkr_so_create_deg2_vecvec_shard(gkr_ctx* ctx, int gate, gkr_vecvec* const* polys, uint32_t n_polys, const uint64_t* gamma_pows,
                                               const uint64_t* point, uint32_t num_vars, uint32_t col_logsize, uint32_t shard, uint32_t n_shards,
                                               gkr_so** out) {
    if (!ctx) return GKR_ERR_ARG;
    if (!point || !out) return ctx->fail(GKR_ERR_ARG, "null argument");
    uint32_t g = 0;
    while ((1u << g) < n_shards) g++;
    if (n_shards == 0 || (1u << g) != n_shards || shard >= n_shards || g > col_logsize)
        return ctx->fail(GKR_ERR_ARG, "the number of shards must be a power of two, at most the number of rows");
    using namespace gkr::frh;
    gkr::FrH weight = ONE;  // eq(point[0 .. g), shard): point[0] pairs with the top bit of the row index
    for (uint32_t i = 0; i < g; i++) {
        const gkr::FrH p = frh_from_limbs(point + 4 * i);
        weight = mul(weight, ((shard >> (g - 1 - i)) & 1) ? p : sub(ONE, p));
    }
    const uint64_t zero[4] = {0, 0, 0, 0};  // a shard does not track the claim: the driver owns it
    return make_vecvec_so(ctx, gate, polys, n_polys, gamma_pows, zero, point + 4 * g, num_vars - g, col_logsize - g, weight, out);
}

extern "C" int gkr_exchange_world(const gkr_exchange* ex);  // sharded.cu

// VecVecDeg2Sumcheck::prove over the row shards: every rank calls this with its shard, the shared transcript state and the WHOLE
// object's claim.  Sparse rounds: the shards' totals are added (one all-gather of two field elements per round through the
// exchange), from12 + Fiat-Shamir run replicated on every rank; then the dense tail over the column variables is the dense
// sharded sumcheck (sharded.cu: local rounds, final gather, last log2(G) rounds on the host).  ex == NULL: one shard.
// out_point: all num_vars challenges reversed (sumcheck.rs:120); out_final_evals: n_polys + 1 values (the eq table last).
extern "C" int gkr_sumcheck_prove_sharded_vecvec(gkr_transcript* t, gkr_so* so, gkr_exchange* ex, const uint64_t global_claim[4],
                                                 uint64_t out_claim[4], uint64_t* out_point, uint64_t* out_final_evals) {
    if (!t || !so || !global_claim) return GKR_ERR_ARG;
    gkr_ctx* ctx = so->ctx;
    Deg2SO* d = dynamic_cast<Deg2SO*>(so);
    if (!d || !d->is_vecvec || d->round_idx != 0 || d->dense) return ctx->fail(GKR_ERR_ARG, "expects a fresh VecVec Deg2 object");
    using namespace gkr::frh;
    const uint32_t world = ex ? (uint32_t)gkr_exchange_world(ex) : 1;
    so->set_prelaunch(true);  // strict partial_sums -> bind alternation below; the dense driver switches it off at the end
    gkr::FrH claim = frh_from_limbs(global_claim);
    std::vector<gkr::FrH> r;
    std::vector<uint64_t> mine(8), all((size_t)8 * world);
    const uint32_t n_sparse = d->n_sparse;
    for (uint32_t k = 0; k < n_sparse; k++) {
        gkr::FrH tot[2];
        uint32_t n = 0;
        const uint32_t b = d->binding_idx();
        int rc = d->partial_sums(tot, &n);
        if (rc) return rc;
        if (world > 1) {
            frh_to_limbs(tot[0], mine.data());
            frh_to_limbs(tot[1], mine.data() + 4);
            rc = gkr_exchange_allgather(ex, mine.data(), 2, all.data());
            if (rc) return ctx->fail(rc, "partial-sum exchange failed");
            tot[0] = tot[1] = ZERO;
            for (uint32_t q = 0; q < world; q++) {
                tot[0] = add(tot[0], frh_from_limbs(all.data() + (size_t)q * 8));
                tot[1] = add(tot[1], frh_from_limbs(all.data() + (size_t)q * 8 + 4));
            }
        }
        gkr::FrH ev[4];
        from12(tot[0], tot[1], d->point[b], d->eq0_inv[b], claim, ev);
        std::vector<gkr::FrH> poly = interpolate_coeffs(ev, 4);
        gkr::FrH msg[3] = {poly[0], poly[2], poly[3]};  // compress_coefficients: the linear term is dropped
        t->t.write_scalars(msg, 3);
        const gkr::FrH x = t->t.challenge(128);
        r.push_back(x);
        claim = evaluate_univar(poly, x);
        rc = d->bind(x);
        if (rc) return rc;
    }
    // dense tail: EqWrapper(GammaWrapper(func, gamma)) over this shard's rows, eq slice already scaled by the shard weight
    std::vector<uint64_t> consts(4 * d->gamma_pows.size()), tail_point((size_t)4 * (d->col + 8));
    for (size_t i = 0; i < d->gamma_pows.size(); i++) frh_to_limbs(d->gamma_pows[i], consts.data() + 4 * i);
    uint64_t cl[4];
    frh_to_limbs(claim, cl);
    uint32_t g = 0;
    while ((1u << g) < world) g++;
    int rc = gkr_sumcheck_prove_sharded(t, so, ex, d->col, GKR_SO_EQ_GAMMA, d->tail_gate, 0, consts.data(), (uint32_t)d->gamma_pows.size(), cl, out_claim,
                                        tail_point.data(), out_final_evals);
    if (rc) return rc;
    if (out_point) {
        const uint32_t n_tail = d->col + g;
        std::memcpy(out_point, tail_point.data(), (size_t)32 * n_tail);  // the last challenges come first
        for (uint32_t k = 0; k < n_sparse; k++) frh_to_limbs(r[n_sparse - 1 - k], out_point + 4 * ((size_t)n_tail + k));
    }
    return GKR_OK;
}

// ---- VecVecPolynomial handles ------------------------------------------------------------------------------
// VecVecPolynomial::new (vecvec.rs:179-189): rows of odd length are padded with row_pad
extern "C" int gkr_vecvec_upload(gkr_ctx* ctx, const uint64_t* flat, const uint32_t* row_len, uint32_t n_rows, const uint64_t row_pad[4],
                                 const uint64_t col_pad[4], uint32_t row_logsize, uint32_t col_logsize, gkr_vecvec** out) {
    if (!ctx) return GKR_ERR_ARG;
    if (!out || !row_pad || !col_pad || (n_rows && !row_len)) return ctx->fail(GKR_ERR_ARG, "null argument");
    if (col_logsize >= 32 || row_logsize >= 32 || n_rows > ((uint64_t)1 << col_logsize)) return ctx->fail(GKR_ERR_ARG, "too many rows for col_logsize");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    gkr_vecvec* v = new gkr_vecvec();
    v->ctx = ctx;
    v->row_pad = frh_from_limbs(row_pad);
    v->col_pad = frh_from_limbs(col_pad);
    v->row_logsize = row_logsize;
    v->col_logsize = col_logsize;
    v->row_len.resize(n_rows);
    uint64_t total = 0, src_total = 0;
    for (uint32_t r = 0; r < n_rows; r++) {
        if (row_len[r] > ((uint64_t)1 << row_logsize)) { delete v; return ctx->fail(GKR_ERR_ARG, "row longer than 1 << row_logsize"); }
        v->row_len[r] = (row_len[r] + 1) & ~1u;
        total += v->row_len[r];
        src_total += row_len[r];
    }
    v->total = total;
    std::vector<uint64_t> padded((size_t)4 * std::max<uint64_t>(total, 1));
    uint64_t so = 0, dof = 0;
    for (uint32_t r = 0; r < n_rows; r++) {
        if (row_len[r]) std::memcpy(padded.data() + 4 * dof, flat + 4 * so, (size_t)32 * row_len[r]);
        if (row_len[r] & 1) std::memcpy(padded.data() + 4 * (dof + row_len[r]), row_pad, 32);
        so += row_len[r];
        dof += v->row_len[r];
    }
    cudaError_t e = gkr_malloc_async(&v->d, sizeof(Fr) * std::max<uint64_t>(total, 1), ctx->stream);
    if (e == cudaSuccess && total) e = cudaMemcpyAsync(v->d, padded.data(), sizeof(Fr) * total, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { delete v; return ctx->fail(GKR_ERR_CUDA, cudaGetErrorString(e)); }
    *out = v;
    return GKR_OK;
}

// VecVecPolynomial::new over rows gathered from a resident table: row r holds src[idx[..]] for its row_len[r] consecutive
// entries of `idx` (src == NULL: the all-ones table), odd rows padded with row_pad.  This is how PushForwardState::new builds
// the bucket images of the point coordinates (pushforward.rs:363-396) without moving the coordinates through the host.
__global__ void vecvec_gather_kernel(Fr* out, const Fr* src, uint64_t src_n, const uint32_t* pidx, uint64_t total, Fr row_pad, int* bad) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t k = pidx[i];
        Fr v;
        if (k == 0xffffffffu) v = row_pad;
        else if (!src) v = fr_one();
        else if (k < src_n) v = src[k];
        else { *bad = 1; v = fr_zero(); }
        out[i] = v;
    }
}
// shared tail of the two gather entries: `d_idx` is the PADDED gather index on the device (0xffffffff = row_pad entry)
static int vecvec_gather_impl(gkr_ctx* ctx, const gkr_table* const* srcs, uint32_t n_src, const uint32_t* d_idx, uint64_t total,
                              const std::vector<uint32_t>& even, const uint64_t* row_pads, const uint64_t* col_pads, uint32_t row_logsize,
                              uint32_t col_logsize, gkr_vecvec** outs) {
    cudaStream_t st = ctx->stream;
    int* d_bad = nullptr;
    GKR_CUDA_OK(ctx, gkr_malloc_async(&d_bad, sizeof(int), st));
    cudaError_t e = cudaMemsetAsync(d_bad, 0, sizeof(int), st);
    for (uint32_t k = 0; k < n_src; k++) outs[k] = nullptr;
    for (uint32_t k = 0; k < n_src && e == cudaSuccess; k++) {
        gkr_vecvec* v = new gkr_vecvec();
        outs[k] = v;
        v->ctx = ctx;
        v->row_pad = frh_from_limbs(row_pads + 4 * k);
        v->col_pad = frh_from_limbs(col_pads + 4 * k);
        v->row_logsize = row_logsize;
        v->col_logsize = col_logsize;
        v->row_len = even;
        v->total = total;
        e = gkr_malloc_async(&v->d, sizeof(Fr) * std::max<uint64_t>(total, 1), st);
        if (e == cudaSuccess && total) {
            unsigned g = (unsigned)std::min<uint64_t>((total + 255) / 256, (uint64_t)ctx->num_sms * 8);
            vecvec_gather_kernel<<<g, 256, 0, st>>>(v->d, srcs[k] ? srcs[k]->d : nullptr, srcs[k] ? srcs[k]->n : 0, d_idx, total,
                                                    fr_from_host(v->row_pad), d_bad);
            ctx->launches++;
            e = cudaGetLastError();
        }
    }
    int bad = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    gkr_free_async(d_bad, st);
    if (e != cudaSuccess || bad) {
        for (uint32_t k = 0; k < n_src; k++) {
            gkr_vecvec_free(outs[k]);
            outs[k] = nullptr;
        }
        return e != cudaSuccess ? ctx->fail(GKR_ERR_CUDA, cudaGetErrorString(e)) : ctx->fail(GKR_ERR_ARG, "gather index out of range");
    }
    return GKR_OK;
}

extern "C" int gkr_vecvec_gather_multi(gkr_ctx* ctx, const gkr_table* const* srcs, uint32_t n_src, const uint32_t* idx, const uint32_t* row_len,
                                       uint32_t n_rows, const uint64_t* row_pads, const uint64_t* col_pads, uint32_t row_logsize,
                                       uint32_t col_logsize, gkr_vecvec** outs) {
    if (!ctx) return GKR_ERR_ARG;
    if (!outs || !srcs || n_src == 0 || !row_pads || !col_pads || (n_rows && (!row_len || !idx))) return ctx->fail(GKR_ERR_ARG, "null argument");
    if (col_logsize >= 32 || row_logsize >= 32 || n_rows > ((uint64_t)1 << col_logsize)) return ctx->fail(GKR_ERR_ARG, "too many rows for col_logsize");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    std::vector<uint32_t> even(n_rows);
    uint64_t total = 0;
    for (uint32_t r = 0; r < n_rows; r++) {
        if (row_len[r] > ((uint64_t)1 << row_logsize)) return ctx->fail(GKR_ERR_ARG, "row longer than 1 << row_logsize");
        even[r] = (row_len[r] + 1) & ~1u;
        total += even[r];
    }
    // padded gather index, built and uploaded once for all sources
    std::vector<uint32_t> pidx(std::max<uint64_t>(total, 1));
    uint64_t so = 0, dof = 0;
    for (uint32_t r = 0; r < n_rows; r++) {
        if (row_len[r]) std::memcpy(pidx.data() + dof, idx + so, sizeof(uint32_t) * row_len[r]);
        if (row_len[r] & 1) pidx[dof + row_len[r]] = 0xffffffffu;
        so += row_len[r];
        dof += even[r];
    }
    cudaStream_t st = ctx->stream;
    uint32_t* d_idx = nullptr;
    GKR_CUDA_OK(ctx, gkr_malloc_async(&d_idx, sizeof(uint32_t) * std::max<uint64_t>(total, 1), st));
    if (total) GKR_CUDA_OK(ctx, cudaMemcpyAsync(d_idx, pidx.data(), sizeof(uint32_t) * total, cudaMemcpyHostToDevice, st));
    int rc = vecvec_gather_impl(ctx, srcs, n_src, d_idx, total, even, row_pads, col_pads, row_logsize, col_logsize, outs);  // synchronises
    gkr_free_async(d_idx, st);
    return rc;
}

// the same with the padded gather index already on the device (gkr_pushforward_bucketize_dev): `row_len` are the UNPADDED
// row lengths, the index holds every row padded to even length
extern "C" int gkr_vecvec_gather_multi_dev(gkr_ctx* ctx, const gkr_table* const* srcs, uint32_t n_src, const gkr_u32buf* padded_idx,
                                           const uint32_t* row_len, uint32_t n_rows, const uint64_t* row_pads, const uint64_t* col_pads,
                                           uint32_t row_logsize, uint32_t col_logsize, gkr_vecvec** outs) {
    if (!ctx) return GKR_ERR_ARG;
    if (!outs || !srcs || n_src == 0 || !row_pads || !col_pads || !padded_idx || (n_rows && !row_len)) return ctx->fail(GKR_ERR_ARG, "null argument");
    if (col_logsize >= 32 || row_logsize >= 32 || n_rows > ((uint64_t)1 << col_logsize)) return ctx->fail(GKR_ERR_ARG, "too many rows for col_logsize");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    std::vector<uint32_t> even(n_rows);
    uint64_t total = 0;
    for (uint32_t r = 0; r < n_rows; r++) {
        if (row_len[r] > ((uint64_t)1 << row_logsize)) return ctx->fail(GKR_ERR_ARG, "row longer than 1 << row_logsize");
        even[r] = (row_len[r] + 1) & ~1u;
        total += even[r];
    }
    if (total != padded_idx->n) return ctx->fail(GKR_ERR_ARG, "padded index length != sum of the even-padded row lengths");
    return vecvec_gather_impl(ctx, srcs, n_src, padded_idx->d, total, even, row_pads, col_pads, row_logsize, col_logsize, outs);
}
extern "C" int gkr_vecvec_gather(gkr_ctx* ctx, const gkr_table* src, const uint32_t* idx, const uint32_t* row_len, uint32_t n_rows,
                                 const uint64_t row_pad[4], const uint64_t col_pad[4], uint32_t row_logsize, uint32_t col_logsize, gkr_vecvec** out) {
    return gkr_vecvec_gather_multi(ctx, &src, 1, idx, row_len, n_rows, row_pad, col_pad, row_logsize, col_logsize, out);
}

extern "C" uint32_t gkr_vecvec_num_rows(const gkr_vecvec* v) { return v ? (uint32_t)v->row_len.size() : 0; }
extern "C" uint64_t gkr_vecvec_total_len(const gkr_vecvec* v) { return v ? v->total : 0; }

extern "C" int gkr_vecvec_download(gkr_ctx* ctx, const gkr_vecvec* v, uint64_t* flat_out, uint32_t* row_len_out, uint64_t row_pad[4],
                                   uint64_t col_pad[4], uint32_t* row_logsize, uint32_t* col_logsize) {
    if (!ctx || !v) return GKR_ERR_ARG;
    if (flat_out && v->total) GKR_CUDA_OK(ctx, cudaMemcpyAsync(flat_out, v->d, sizeof(Fr) * v->total, cudaMemcpyDeviceToHost, ctx->stream));
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    if (row_len_out) std::memcpy(row_len_out, v->row_len.data(), sizeof(uint32_t) * v->row_len.size());
    if (row_pad) frh_to_limbs(v->row_pad, row_pad);
    if (col_pad) frh_to_limbs(v->col_pad, col_pad);
    if (row_logsize) *row_logsize = v->row_logsize;
    if (col_logsize) *col_logsize = v->col_logsize;
    return GKR_OK;
}

extern "C" void gkr_vecvec_free(gkr_vecvec* v) {
    if (!v) return;
    if (v->d) gkr_free_async(v->d, v->ctx->stream);
    delete v;
}
