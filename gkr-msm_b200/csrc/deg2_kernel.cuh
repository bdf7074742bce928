// Round kernel of the Deg2 sumcheck objects (see deg2.cu for the object layer and the reference citations).  Shared by two
// translation units: deg2.cu builds it with the multiplier inlined (throughput flavour, large tables) and
// deg2_compact.cu with GKR_COMPACT_FIELD, i.e. the multiplier as an out-of-line call (latency flavour: ~4x less code to
// stream through a cold instruction cache, which is what bounds the small rounds of a proof).
#pragma once
#include "common.cuh"
#include "gates.cuh"

#ifdef GKR_COMPACT_FIELD
#define GKR_DEG2_NS deg2_compact
#else
#define GKR_DEG2_NS deg2_inline
#endif

struct Deg2Block {
    int gate;
    int in_idx[6];
    int out_off;
    int own_mask;  // bit k: this block writes the folded table in_idx[k] (each table has exactly one owner)
};

// One round of a Deg2 object: FOLD the previous round's tables with the challenge (unless this is round 0), write the
// new tables, evaluate the gate stack at 1 and "2" on the fresh pairs, weight by eq, and add the closed-form padding sum.
struct Deg2RoundArgs {
    const Fr* const* in;       // [P] round b-1 tables when fold != 0, else the current tables
    Fr* const* out;            // [P] round b tables (written when fold != 0)
    const Deg2Block* blocks;   // [gridDim.y]
    const Fr* gammas;          // [n_outs], gammas[0] == 1
    const uint32_t* off_old;   // element offsets of round b-1 [nrows + 1] (fold only)
    const uint32_t* pair_off;  // PAIR offsets of round b [nrows + 1]; nullptr: one dense row
    uint32_t nrows;
    const Fr* eq;              // eq table of round b (row-local part)
    const Fr* rowcoef;         // [nrows] or nullptr
    uint64_t n_pairs;          // pairs of round b
    int fold;
    Fr t;
    const Fr* row_pads;        // [P]
    const Fr* pt;              // row variables of this round's eq table, for the padding term
    uint32_t n_pt;
    int do_pad;                // ragged objects only: compute the padding term T (compact flavour: by the blocks with
                               // blockIdx.x >= work_blocks_x of the y == 0 slice; inline flavour: epilogue of the y == 0 blocks)
    uint32_t work_blocks_x;    // blocks per gate slice that evaluate pairs
    int one_pair_rows;         // every row of this round holds exactly one pair: row == pair index, no search
    RoundOut o;                // 3 accumulators: S1, S2, T
    MailboxRef mbox;           // pre-launched fused round: the challenge (plain 128-bit integer) comes through the mailbox
};

// Two lanes per pair.  Lane role r = lane & 1 owns element 2 idx + r of the (new) row: it folds (or loads) that element of
// every input, stores it, and evaluates the gate stack at ONE point -- role 1 at "1" (a = p1), role 0 at "2"
// (a = 2 p1 - p0, with p1 taken from the partner lane by shuffle).  Compared with one thread per pair this halves the
// dependent multiplication chain and the code executed per launch (the small rounds of a proof are latency / instruction
// fetch bound), and every load and store is fully coalesced (consecutive lanes touch consecutive 64-byte / 32-byte pieces).
namespace GKR_DEG2_NS {

template <int G>
__device__ __forceinline__ Fr deg2_block_round(const Deg2RoundArgs& A, const Fr& t_fold, const Deg2Block& blk, bool active, uint64_t q, uint32_t role,
                                               uint32_t row, uint64_t idx, const Fr& w) {
    constexpr int NI = MoGate<G>::N_INS, NO = MoGate<G>::N_OUTS;
    Fr a[NI];
    uint64_t old_base = 0, half_old = 0;
    if (A.fold && active) {
        if (A.pair_off) {
            old_base = A.off_old[row];
            half_old = (A.off_old[row + 1] - old_base) >> 1;
        } else {
            half_old = 2 * A.n_pairs;  // dense: the old table has 4 * n_pairs entries
        }
    }
    const uint64_t i_self = 2 * idx + role;  // position of this lane's element inside the row
#pragma unroll
    for (int j = 0; j < NI; j++) {
        const int tj = blk.in_idx[j];
        Fr self = fr_zero();
        if (active) {
            if (A.fold) {
                if (i_self < half_old) {
                    const Fr* src = A.in[tj] + old_base + 2 * i_self;
                    Fr e0 = src[0], e1 = src[1];
                    self = fr_add(e0, fr_mul(t_fold, fr_sub(e1, e0)));
                } else {
                    self = A.row_pads[tj];  // odd half re-padded with row_pad (vecvec.rs:432-436)
                }
                if ((blk.own_mask >> j) & 1) A.out[tj][2 * q + role] = self;
            } else {
                self = A.in[tj][2 * q + role];
            }
        }
        const Fr other = fr_shfl_xor(self, 1);
        const Fr two = fr_sub(fr_dbl(other), self);  // role 0: the "2-1" value 2 p(1) - p(0) of make_21, never stored
#pragma unroll
        for (int k = 0; k < 8; k++) a[j].l[k] = role ? self.l[k] : two.l[k];
    }
    Fr o[NO];
    MoGate<G>::eval(a, o);
    Fr g;
    if (blk.out_off == 0) {
        g = o[0];
    } else {
        g = fr_mul(o[0], A.gammas[blk.out_off]);
    }
#pragma unroll
    for (int k = 1; k < NO; k++) g = fr_add(g, fr_mul(o[k], A.gammas[blk.out_off + k]));
    return fr_mul(g, w);
}

// grid = (x: pair lanes, y: gate blocks).  All blocks reduce into the same three sums.
// GSEL >= 0: every gate block of the launch is the same base gate (true for all stacks of the reference except
// Stacked(affine L1, Repeated(BitCheck, 2)): TRI_L1 decomposes into three PRJ_L1 blocks, Repeated(g, n) into n blocks of
// g), so the kernel holds ONE inlined gate instead of a 9-way switch over all of them -- round 1's generic kernel was
// ~40 k SASS instructions and spent 4.6 cycles per issue waiting for instruction fetch on mid-size rounds
// (profiles/r01_ncu_full_deg2_inline_round.csv).  GSEL < 0: the generic kernel (mixed stacks).
template <int GSEL>
__global__ void __launch_bounds__(GKR_REDUCE_THREADS) deg2_round_kernel(const __grid_constant__ Deg2RoundArgs A) {
    __shared__ Fr smem[3 * (GKR_REDUCE_THREADS / 32)];
    Fr mine = fr_zero();
    Fr t_fold = A.t;
    if (A.mbox.box) {  // pre-launched: wait for the challenge (a cancelled launch publishes nothing)
        uint32_t tw[4];
        if (!gkr_mailbox_wait(A.mbox, tw)) return;
        // the plain 128-bit integer -> Montgomery form: one product with R^2 mod r
        Fr tp = fr_zero(), r2;
#pragma unroll
        for (int k = 0; k < 4; k++) tp.l[k] = tw[k];
        r2.l[0] = 0xf3f29c6du; r2.l[1] = 0xc999e990u; r2.l[2] = 0x87925c23u; r2.l[3] = 0x2b6cedcbu;
        r2.l[4] = 0x7254398fu; r2.l[5] = 0x05d31496u; r2.l[6] = 0x9f59ff11u; r2.l[7] = 0x0748d9d9u;
        t_fold = fr_mul(tp, r2);
    }
    const Deg2Block blk = A.blocks[blockIdx.y];
#ifdef GKR_COMPACT_FIELD
    // latency flavour: the tail of the grid (blockIdx.x >= work_blocks_x) only computes T, overlapping the gate evaluations
    const bool pad_slice = blockIdx.x >= A.work_blocks_x;
    const uint64_t stride = (uint64_t)A.work_blocks_x * blockDim.x, n_lanes = pad_slice ? 0 : 2 * A.n_pairs;
#else
    // throughput flavour: T is a negligible epilogue of the y == 0 blocks (launched with work_blocks_x == gridDim.x)
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, n_lanes = 2 * A.n_pairs;
#endif
    const uint32_t lane = threadIdx.x & 31, role = lane & 1;
    for (uint64_t wb = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x - lane); wb < n_lanes; wb += stride) {
        const uint64_t t = wb + lane, q = t >> 1;
        const bool active = t < n_lanes;  // n_lanes is even: both lanes of a pair agree
        Fr w = fr_zero();
        uint32_t row = 0;
        uint64_t idx = q;
        if (active) {
            if (A.pair_off && A.one_pair_rows) {  // every row holds one pair: row == pair index, no search (both flavours)
                row = (uint32_t)q;
                idx = 0;
                w = fr_mul(A.eq[0], A.rowcoef[row]);
            } else if (A.pair_off) {
                uint32_t lo = 0, hi = A.nrows;  // largest r with pair_off[r] <= q (empty rows repeat an offset)
                while (hi - lo > 1) {
                    uint32_t mid = (lo + hi) >> 1;
                    if ((uint64_t)A.pair_off[mid] <= q) lo = mid; else hi = mid;
                }
                row = lo;
                idx = q - A.pair_off[lo];
                w = fr_mul(A.eq[idx], A.rowcoef[lo]);
            } else {
                w = A.eq[q];
            }
        }
        Fr v = fr_zero();
        if constexpr (GSEL >= 0) {
            v = deg2_block_round<GSEL>(A, t_fold, blk, active, q, role, row, idx, w);
        } else
        switch (blk.gate) {
            case GATE_AFF_L1: v = deg2_block_round<GATE_AFF_L1>(A, t_fold, blk, active, q, role, row, idx, w); break;
            case GATE_AFF_L2: v = deg2_block_round<GATE_AFF_L2>(A, t_fold, blk, active, q, role, row, idx, w); break;
            case GATE_AFF_L3: v = deg2_block_round<GATE_AFF_L3>(A, t_fold, blk, active, q, role, row, idx, w); break;
            case GATE_PRJ_L1: v = deg2_block_round<GATE_PRJ_L1>(A, t_fold, blk, active, q, role, row, idx, w); break;
            case GATE_PRJ_L2: v = deg2_block_round<GATE_PRJ_L2>(A, t_fold, blk, active, q, role, row, idx, w); break;
            case GATE_PRJ_L3: v = deg2_block_round<GATE_PRJ_L3>(A, t_fold, blk, active, q, role, row, idx, w); break;
            case GATE_BITCHECK: v = deg2_block_round<GATE_BITCHECK>(A, t_fold, blk, active, q, role, row, idx, w); break;
            case GATE_LOGUP_LAYER: v = deg2_block_round<GATE_LOGUP_LAYER>(A, t_fold, blk, active, q, role, row, idx, w); break;
            case GATE_ADD_INVERSES: v = deg2_block_round<GATE_ADD_INVERSES>(A, t_fold, blk, active, q, role, row, idx, w); break;
            default: break;
        }
        mine = fr_add(mine, v);  // inactive lanes carry w == 0
    }
    Fr acc[3];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        acc[0].l[k] = role ? mine.l[k] : 0u;  // S1: lanes that evaluated at 1
        acc[1].l[k] = role ? 0u : mine.l[k];  // S2: lanes that evaluated at "2"
        acc[2].l[k] = 0u;
    }
    // T = sum_rows rowcoef[row] * (1 - eq_sum(pt, len_row / 2)): closed form of src/utils.rs:265-291 per row
    // (vecvec_eq.rs:344-369), computed by its own slice of the grid so that it overlaps the gate evaluations
#ifdef GKR_COMPACT_FIELD
    if (pad_slice && A.do_pad && blockIdx.y == 0) {
        const Fr one = fr_one();
        const uint64_t pstride = (uint64_t)(gridDim.x - A.work_blocks_x) * blockDim.x;
        for (uint64_t r = (uint64_t)(blockIdx.x - A.work_blocks_x) * blockDim.x + threadIdx.x; r < A.nrows; r += pstride) {
#else
    if (A.do_pad && blockIdx.y == 0) {
        const Fr one = fr_one();
        for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < A.nrows; r += stride) {
#endif
            uint64_t k = A.pair_off[r + 1] - A.pair_off[r];
            Fr s;
            if (k >= ((uint64_t)1 << A.n_pt)) {
                s = one;
            } else {
                Fr mult = one;
                s = fr_zero();
                for (uint32_t i = 0; i < A.n_pt; i++) {
                    uint32_t bit = (uint32_t)(k >> (A.n_pt - i - 1)) & 1u;
                    Fr p = A.pt[i];
                    if (bit) {
                        Fr nm = fr_mul(mult, p);
                        s = fr_add(s, fr_sub(mult, nm));
                        mult = nm;
                    } else {
                        mult = fr_mul(mult, fr_sub(one, p));
                    }
                }
            }
            acc[2] = fr_add(acc[2], fr_mul(A.rowcoef[r], fr_sub(one, s)));
        }
    }
    grid_reduce_to_host<3>(acc, smem, A.o);
}

// host side: pick the instantiation (uniform_gate < 0: generic)
static inline void launch_deg2_round(int uniform_gate, const Deg2RoundArgs& a, dim3 grid, unsigned threads, cudaStream_t stream) {
    switch (uniform_gate) {
#define GKR_DEG2_CASE(G) case G: deg2_round_kernel<G><<<grid, threads, 0, stream>>>(a); break;
        GKR_DEG2_CASE(GATE_AFF_L1)
        GKR_DEG2_CASE(GATE_AFF_L2)
        GKR_DEG2_CASE(GATE_AFF_L3)
        GKR_DEG2_CASE(GATE_PRJ_L1)
        GKR_DEG2_CASE(GATE_PRJ_L2)
        GKR_DEG2_CASE(GATE_PRJ_L3)
        GKR_DEG2_CASE(GATE_BITCHECK)
        GKR_DEG2_CASE(GATE_LOGUP_LAYER)
        GKR_DEG2_CASE(GATE_ADD_INVERSES)
#undef GKR_DEG2_CASE
        default: deg2_round_kernel<-1><<<grid, threads, 0, stream>>>(a); break;
    }
}

}  // namespace GKR_DEG2_NS
