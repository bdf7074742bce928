// Witness generation: element-wise polynomial gate maps over dense and ragged (VecVec) tables, optionally fused
// with the even/odd (or high-bit) split that pairs neighbours for the next GKR layer.
//   trait MapSplit {algfn_map, algfn_map_split}          src/cleanup/polys/common.rs:23-35
//   Vec<F>:            algfn_map / algfn_map_split        src/cleanup/polys/dense.rs:114-185
//   VecVecPolynomial:  vecvec_map / vecvec_map_split      src/cleanup/polys/vecvec.rs:480-606
//                      vecvec_map_split_to_dense          src/cleanup/polys/vecvec.rs:608-654
//   AlgFnUtils::{map, map_split_hi}                       src/cleanup/utils/algfn.rs:49-90
// The reference's split variants are serial with a heap allocation per element (dense.rs:131, vecvec.rs:581);
// here every variant is one grid launch: thread = input element, blockIdx.y = base-gate block of the stack.
#include <algorithm>
#include "common.cuh"
#include "gates.cuh"
#include "host_gates.hpp"

#define GATE_ID1 100  // internal: one input copied to one output (IdAlgFn expands into n of these)

struct MapBlock {
    int gate;
    int in_idx[6];
    int out_idx[2][4];  // destination table per side (side 1 only for split maps) and output
    int n_out;
};

struct MapArgs {
    const Fr* const* in;
    Fr* const* out;
    const MapBlock* blocks;
    uint64_t n;  // input elements
    // split description
    int split;            // 0: plain map, 1: dense split with segment `seg`, 2: ragged split (segment 1), 3: ragged -> dense
    uint64_t seg;         // dense: segment size (power of two)
    const uint32_t* off_old;  // ragged: element offsets of the input rows [nrows + 1]
    const uint32_t* off_new;  // ragged split: element offsets of the output rows
    uint32_t nrows;
    const Fr* out_row_pad;  // [n_out_tables] pad value of every OUTPUT table (ragged split re-padding / empty rows)
};

template <int G>
__device__ __forceinline__ void map_block(const MapArgs& A, const MapBlock& b, uint64_t e, int side, uint64_t dst) {
    constexpr int NI = MoGate<G>::N_INS, NO = MoGate<G>::N_OUTS;
    Fr a[NI], o[NO];
#pragma unroll
    for (int j = 0; j < NI; j++) a[j] = A.in[b.in_idx[j]][e];
    MoGate<G>::eval(a, o);
#pragma unroll
    for (int k = 0; k < NO; k++) A.out[b.out_idx[side][k]][dst] = o[k];
}

__global__ void __launch_bounds__(256) map_kernel(const __grid_constant__ MapArgs A) {
    const MapBlock b = A.blocks[blockIdx.y];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < A.n; e += stride) {
        int side = 0;
        uint64_t dst = e;
        if (A.split == 1) {
            side = (int)((e / A.seg) & 1);
            dst = (e / (2 * A.seg)) * A.seg + (e & (A.seg - 1));
        } else if (A.split >= 2) {
            uint32_t lo = 0, hi = A.nrows;
            while (hi - lo > 1) {
                uint32_t mid = (lo + hi) >> 1;
                if ((uint64_t)A.off_old[mid] <= e) lo = mid; else hi = mid;
            }
            uint64_t i = e - A.off_old[lo];
            side = (int)(i & 1);
            if (A.split == 2) {
                dst = A.off_new[lo] + (i >> 1);
                // odd half: the output row gets one pad element on both sides (vecvec.rs:586-593)
                uint64_t half = (A.off_old[lo + 1] - A.off_old[lo]) >> 1;
                if ((half & 1) && i + 1 == 2 * half && side == 1) {
                    for (int s = 0; s < 2; s++)
                        for (int k = 0; k < b.n_out; k++) {
                            int t = b.out_idx[s][k];
                            A.out[t][A.off_new[lo] + half] = A.out_row_pad[t];
                        }
                }
            } else {
                dst = lo;  // rows of length two: one value per row and side
            }
        }
        switch (b.gate) {
            case GATE_AFF_L1: map_block<GATE_AFF_L1>(A, b, e, side, dst); break;
            case GATE_AFF_L2: map_block<GATE_AFF_L2>(A, b, e, side, dst); break;
            case GATE_AFF_L3: map_block<GATE_AFF_L3>(A, b, e, side, dst); break;
            case GATE_PRJ_L1: map_block<GATE_PRJ_L1>(A, b, e, side, dst); break;
            case GATE_PRJ_L2: map_block<GATE_PRJ_L2>(A, b, e, side, dst); break;
            case GATE_PRJ_L3: map_block<GATE_PRJ_L3>(A, b, e, side, dst); break;
            case GATE_BITCHECK: map_block<GATE_BITCHECK>(A, b, e, side, dst); break;
            case GATE_LOGUP_LAYER: map_block<GATE_LOGUP_LAYER>(A, b, e, side, dst); break;
            case GATE_ADD_INVERSES: map_block<GATE_ADD_INVERSES>(A, b, e, side, dst); break;
            case GATE_ID1: A.out[b.out_idx[side][0]][dst] = A.in[b.in_idx[0]][e]; break;
            default: break;
        }
    }
}

// fills out[t][r] for empty input rows (ragged -> dense, vecvec.rs:640-646) and r >= nrows (col pad, :650-652)
struct FillArgs {
    Fr* const* out;
    const uint32_t* off_old;
    uint32_t nrows;
    uint64_t n_out;
    const Fr* row_pad;
    const Fr* col_pad;
};
__global__ void map_to_dense_fill_kernel(const __grid_constant__ FillArgs A) {
    const int t = blockIdx.y;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < A.n_out; r += (uint64_t)gridDim.x * blockDim.x) {
        if (r >= A.nrows) A.out[t][r] = A.col_pad[t];
        else if (A.off_old[r + 1] == A.off_old[r]) A.out[t][r] = A.row_pad[t];
    }
}

// ---- host side ---------------------------------------------------------------------------------------------
namespace {

struct Stack {
    std::vector<int> gate, repeat;  // public gate ids; GKR_GATE_ID allowed (repeat = n)
    int n_ins = 0, n_outs = 0;
};

bool stack_init(Stack* s, const int* g, const uint32_t* r, uint32_t n) {
    if (!g || !r || n == 0) return false;
    for (uint32_t i = 0; i < n; i++) {
        int ni = 0, no = 0;
        if (g[i] == GKR_GATE_ID) { ni = no = 1; }
        else if (!gkr::base_gate_io(g[i], &ni, &no)) return false;
        if (r[i] == 0) return false;
        s->gate.push_back(g[i]);
        s->repeat.push_back((int)r[i]);
        s->n_ins += ni * (int)r[i];
        s->n_outs += no * (int)r[i];
    }
    return true;
}

void stack_eval_host(const Stack& s, const gkr::FrH* a, gkr::FrH* o) {
    for (size_t i = 0; i < s.gate.size(); i++) {
        int ni = 1, no = 1;
        if (s.gate[i] != GKR_GATE_ID) gkr::base_gate_io(s.gate[i], &ni, &no);
        for (int k = 0; k < s.repeat[i]; k++) {
            if (s.gate[i] == GKR_GATE_ID) o[0] = a[0]; else gkr::base_gate_eval(s.gate[i], a, o);
            a += ni;
            o += no;
        }
    }
}

// table index of (side, output o) after interleaving chunks of `bundle` (dense.rs:137-138, vecvec.rs:596-597)
int split_out_index(int o, int side, int bundle, int n_outs) {
    int c = o / bundle, pos = o % bundle;
    int size_c = std::min(bundle, n_outs - c * bundle);
    return 2 * c * bundle + side * size_c + pos;
}

std::vector<MapBlock> expand(const Stack& s, bool split, int bundle) {
    std::vector<MapBlock> out;
    int in_off = 0, out_off = 0;
    auto push = [&](int gate, std::initializer_list<int> idx, int ooff, int nout) {
        MapBlock b;
        b.gate = gate;
        b.n_out = nout;
        int c = 0;
        for (int x : idx) b.in_idx[c++] = in_off + x;
        for (; c < 6; c++) b.in_idx[c] = 0;
        for (int k = 0; k < 4; k++) {
            int o = out_off + ooff + std::min(k, nout - 1);
            b.out_idx[0][k] = split ? split_out_index(o, 0, bundle, s.n_outs) : o;
            b.out_idx[1][k] = split ? split_out_index(o, 1, bundle, s.n_outs) : o;
        }
        out.push_back(b);
    };
    for (size_t p = 0; p < s.gate.size(); p++) {
        int ni = 1, no = 1;
        if (s.gate[p] != GKR_GATE_ID) gkr::base_gate_io(s.gate[p], &ni, &no);
        for (int k = 0; k < s.repeat[p]; k++) {
            switch (s.gate[p]) {
                case GKR_GATE_TRI_L1:
                    push(GATE_PRJ_L1, {0, 1, 2, 6, 7, 8}, 0, 4);
                    push(GATE_PRJ_L1, {3, 4, 5, 9, 10, 11}, 4, 4);
                    push(GATE_PRJ_L1, {6, 7, 8, 9, 10, 11}, 8, 4);
                    break;
                case GKR_GATE_AFF_L1_BITCHECK2:
                    push(GATE_AFF_L1, {0, 1, 2, 3}, 0, 3);
                    push(GATE_BITCHECK, {4}, 3, 1);
                    push(GATE_BITCHECK, {5}, 4, 1);
                    break;
                case GKR_GATE_AFF_L1: push(GATE_AFF_L1, {0, 1, 2, 3}, 0, 3); break;
                case GKR_GATE_AFF_L2: push(GATE_AFF_L2, {0, 1, 2}, 0, 3); break;
                case GKR_GATE_AFF_L3: push(GATE_AFF_L3, {0, 1, 2}, 0, 3); break;
                case GKR_GATE_PRJ_L1: push(GATE_PRJ_L1, {0, 1, 2, 3, 4, 5}, 0, 4); break;
                case GKR_GATE_PRJ_L2: push(GATE_PRJ_L2, {0, 1, 2, 3}, 0, 4); break;
                case GKR_GATE_PRJ_L3: push(GATE_PRJ_L3, {0, 1, 2, 3}, 0, 3); break;
                case GKR_GATE_BITCHECK: push(GATE_BITCHECK, {0}, 0, 1); break;
                case GKR_GATE_LOGUP_LAYER: push(GATE_LOGUP_LAYER, {0, 1, 2, 3}, 0, 2); break;
                case GKR_GATE_ADD_INVERSES: push(GATE_ADD_INVERSES, {0, 1}, 0, 2); break;
                case GKR_GATE_ID: push(GATE_ID1, {0}, 0, 1); break;
                default: break;
            }
            in_off += ni;
            out_off += no;
        }
    }
    return out;
}

// small device-side argument arrays of one map launch: ONE allocation and ONE staged (pinned ring) upload, so that a map call
// neither issues five pageable copies nor has to synchronise before its host vectors go out of scope
struct DevArrays {
    gkr_ctx* ctx;
    unsigned char* base = nullptr;
    std::vector<unsigned char> host;
    const Fr** in = nullptr;
    Fr** out = nullptr;
    MapBlock* blocks = nullptr;
    Fr* pads = nullptr;
    uint32_t* offs = nullptr;
    size_t o_in = 0, o_out = 0, o_blocks = 0, o_pads = 0, o_offs = 0;
    explicit DevArrays(gkr_ctx* c) : ctx(c) {}
    ~DevArrays() {
        if (base) gkr_free_async(base, ctx->stream);
    }
    template <class T>
    size_t put(const std::vector<T>& v) {
        size_t off = (host.size() + 31) & ~(size_t)31;
        host.resize(off + std::max<size_t>(sizeof(T) * v.size(), 1));
        if (!v.empty()) std::memcpy(host.data() + off, v.data(), sizeof(T) * v.size());
        return off;
    }
    int upload() {
        GKR_CUDA_OK(ctx, gkr_malloc_async(&base, host.size(), ctx->stream));
        int rc = gkr_stage_upload(ctx, base, host.data(), host.size());
        if (rc) return rc;
        in = (const Fr**)(base + o_in);
        out = (Fr**)(base + o_out);
        blocks = (MapBlock*)(base + o_blocks);
        pads = (Fr*)(base + o_pads);
        offs = (uint32_t*)(base + o_offs);
        return GKR_OK;
    }
};

int launch_map(gkr_ctx* ctx, MapArgs& a, int n_blocks) {
    if (a.n == 0) return GKR_OK;
    dim3 grid((unsigned)std::max<uint64_t>(1, std::min<uint64_t>((a.n + 255) / 256, (uint64_t)ctx->num_sms * 8)), (unsigned)n_blocks);
    map_kernel<<<grid, 256, 0, ctx->stream>>>(a);
    ctx->launches++;
    GKR_CUDA_OK(ctx, cudaGetLastError());
    return GKR_OK;
}

}  // namespace

// Vec::algfn_map (dense.rs:141-184) and, with split_kind >= 0, Vec::algfn_map_split (dense.rs:115-139).
// split_kind: -1 none, 0 = SplitIdx::LO(var_idx), 1 = SplitIdx::HI(var_idx).  out receives n_outs (or 2*n_outs)
// fresh tables in the reference's output order.
extern "C" int gkr_map_dense(gkr_ctx* ctx, const int* part_gate, const uint32_t* part_repeat, uint32_t n_parts,
                             gkr_table* const* in, uint32_t n_in, int split_kind, uint32_t var_idx, uint32_t bundle_size,
                             gkr_table** out, uint32_t* n_out) {
    if (!ctx) return GKR_ERR_ARG;
    if (!in || !out) return ctx->fail(GKR_ERR_ARG, "null argument");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    Stack st;
    if (!stack_init(&st, part_gate, part_repeat, n_parts)) return ctx->fail(GKR_ERR_ARG, "invalid gate stack");
    if ((int)n_in != st.n_ins) return ctx->fail(GKR_ERR_ARG, "number of tables != f.n_ins()");
    const uint64_t len = in[0]->n;
    for (uint32_t j = 0; j < n_in; j++)
        if (!in[j] || in[j]->n != len) return ctx->fail(GKR_ERR_ARG, "tables must have equal length");
    const bool split = split_kind >= 0;
    uint64_t seg = 1;
    if (split) {
        if (bundle_size == 0) return ctx->fail(GKR_ERR_ARG, "bundle_size must be positive");
        uint32_t nv = 0;
        while (((uint64_t)1 << nv) < len) nv++;
        if (((uint64_t)1 << nv) != len || nv == 0) return ctx->fail(GKR_ERR_ARG, "split map needs a power-of-two table of at least 2 entries");
        if (var_idx >= nv) return ctx->fail(GKR_ERR_ARG, "split variable out of range");
        uint32_t lo = split_kind == 0 ? var_idx : nv - 1 - var_idx;
        seg = (uint64_t)1 << lo;
    }
    const int n_tables = split ? 2 * st.n_outs : st.n_outs;
    const uint64_t out_len = split ? len / 2 : len;
    std::vector<Fr*> outs(n_tables);
    for (int t = 0; t < n_tables; t++) {
        out[t] = nullptr;
        int rc = gkr_table_alloc(ctx, out_len, &out[t]);
        if (rc) return rc;
        outs[t] = out[t]->d;
    }
    std::vector<const Fr*> ins(n_in);
    for (uint32_t j = 0; j < n_in; j++) ins[j] = in[j]->d;
    std::vector<MapBlock> blocks = expand(st, split, (int)bundle_size);
    DevArrays d(ctx);
    d.o_in = d.put(ins);
    d.o_out = d.put(outs);
    d.o_blocks = d.put(blocks);
    int rc = d.upload();
    if (rc) return rc;
    MapArgs a{};
    a.in = d.in;
    a.out = d.out;
    a.blocks = d.blocks;
    a.n = len;
    a.split = split ? 1 : 0;
    a.seg = seg;
    rc = launch_map(ctx, a, (int)blocks.size());
    if (rc) return rc;
    if (n_out) *n_out = (uint32_t)n_tables;
    return GKR_OK;
}

// vecvec_map (vecvec.rs:480-540), vecvec_map_split (:542-606) and vecvec_map_split_to_dense (:608-654).
// mode 0: map -> n_outs gkr_vecvec;  mode 1: split at LO(0) -> 2*n_outs gkr_vecvec (row_logsize - 1);
// mode 2: split to dense (requires row_logsize == 1) -> 2*n_outs gkr_table of 1 << col_logsize entries.
extern "C" int gkr_map_vecvec(gkr_ctx* ctx, const int* part_gate, const uint32_t* part_repeat, uint32_t n_parts,
                              gkr_vecvec* const* in, uint32_t n_in, int mode, uint32_t bundle_size, void** out, uint32_t* n_out) {
    if (!ctx) return GKR_ERR_ARG;
    if (!in || !out) return ctx->fail(GKR_ERR_ARG, "null argument");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    Stack st;
    if (!stack_init(&st, part_gate, part_repeat, n_parts)) return ctx->fail(GKR_ERR_ARG, "invalid gate stack");
    if ((int)n_in != st.n_ins) return ctx->fail(GKR_ERR_ARG, "number of polynomials != f.n_ins()");
    const gkr_vecvec* p0 = in[0];
    for (uint32_t j = 0; j < n_in; j++)
        if (!in[j] || in[j]->row_len != p0->row_len || in[j]->row_logsize != p0->row_logsize || in[j]->col_logsize != p0->col_logsize)
            return ctx->fail(GKR_ERR_ARG, "all polynomials of a bundle must share the row structure");
    if (mode < 0 || mode > 2) return ctx->fail(GKR_ERR_ARG, "bad mode");
    if (mode >= 1 && bundle_size == 0) return ctx->fail(GKR_ERR_ARG, "bundle_size must be positive");
    if (mode >= 1 && p0->row_logsize == 0) return ctx->fail(GKR_ERR_ARG, "no row variable left to split");
    if (mode == 2 && p0->row_logsize != 1) return ctx->fail(GKR_ERR_ARG, "split to dense requires row_logsize == 1 (vecvec.rs:618)");
    const uint32_t nrows = (uint32_t)p0->row_len.size();
    const bool split = mode >= 1;
    const int n_tables = split ? 2 * st.n_outs : st.n_outs;
    // pads of the outputs: f(row_pads), f(col_pads)  (vecvec.rs:491-499)
    std::vector<gkr::FrH> rp_in(n_in), cp_in(n_in), rp_out(st.n_outs), cp_out(st.n_outs);
    for (uint32_t j = 0; j < n_in; j++) { rp_in[j] = in[j]->row_pad; cp_in[j] = in[j]->col_pad; }
    stack_eval_host(st, rp_in.data(), rp_out.data());
    stack_eval_host(st, cp_in.data(), cp_out.data());
    std::vector<Fr> pad_tab(2 * (size_t)n_tables);  // [row pads per output table | col pads per output table]
    std::vector<gkr::FrH> rp_tab(n_tables), cp_tab(n_tables);
    for (int o = 0; o < st.n_outs; o++)
        for (int s = 0; s < (split ? 2 : 1); s++) {
            int t = split ? split_out_index(o, s, (int)bundle_size, st.n_outs) : o;
            rp_tab[t] = rp_out[o];
            cp_tab[t] = cp_out[o];
            pad_tab[t] = fr_from_host(rp_out[o]);
            pad_tab[n_tables + t] = fr_from_host(cp_out[o]);
        }
    // row structure of the outputs
    std::vector<uint32_t> new_len(nrows), off_old(nrows + 1), off_new(nrows + 1);
    uint64_t acc_o = 0, acc_n = 0;
    for (uint32_t r = 0; r < nrows; r++) {
        off_old[r] = (uint32_t)acc_o;
        off_new[r] = (uint32_t)acc_n;
        uint32_t h = p0->row_len[r] / 2;
        new_len[r] = mode == 0 ? p0->row_len[r] : ((h + 1) & ~1u);
        acc_o += p0->row_len[r];
        acc_n += new_len[r];
    }
    off_old[nrows] = (uint32_t)acc_o;
    off_new[nrows] = (uint32_t)acc_n;
    if (mode == 2)
        for (uint32_t r = 0; r < nrows; r++)
            if (p0->row_len[r] != 0 && p0->row_len[r] != 2) return ctx->fail(GKR_ERR_ARG, "split to dense: rows must have length 0 or 2");

    std::vector<Fr*> outs(n_tables);
    const uint64_t dense_len = (uint64_t)1 << p0->col_logsize;
    for (int t = 0; t < n_tables; t++) {
        out[t] = nullptr;
        if (mode == 2) {
            gkr_table* tb = nullptr;
            int rc = gkr_table_alloc(ctx, dense_len, &tb);
            if (rc) return rc;
            out[t] = tb;
            outs[t] = tb->d;
        } else {
            gkr_vecvec* v = new gkr_vecvec();
            v->ctx = ctx;
            v->total = mode == 0 ? acc_o : acc_n;
            v->row_len = mode == 0 ? p0->row_len : new_len;
            v->row_pad = rp_tab[t];
            v->col_pad = cp_tab[t];
            v->row_logsize = mode == 0 ? p0->row_logsize : p0->row_logsize - 1;
            v->col_logsize = p0->col_logsize;
            cudaError_t e = gkr_malloc_async(&v->d, sizeof(Fr) * std::max<uint64_t>(v->total, 1), ctx->stream);
            if (e != cudaSuccess) { delete v; return ctx->fail(GKR_ERR_CUDA, cudaGetErrorString(e)); }
            out[t] = v;
            outs[t] = v->d;
        }
    }
    std::vector<const Fr*> ins(n_in);
    for (uint32_t j = 0; j < n_in; j++) ins[j] = in[j]->d;
    std::vector<MapBlock> blocks = expand(st, split, (int)bundle_size);
    std::vector<uint32_t> offs(off_old);
    offs.insert(offs.end(), off_new.begin(), off_new.end());
    DevArrays d(ctx);
    d.o_in = d.put(ins);
    d.o_out = d.put(outs);
    d.o_blocks = d.put(blocks);
    d.o_pads = d.put(pad_tab);
    d.o_offs = d.put(offs);
    int rc = d.upload();
    if (rc) return rc;
    MapArgs a{};
    a.in = d.in;
    a.out = d.out;
    a.blocks = d.blocks;
    a.n = acc_o;
    a.split = mode == 0 ? 0 : (mode == 1 ? 2 : 3);
    a.seg = 1;
    a.off_old = d.offs;
    a.off_new = d.offs + (nrows + 1);
    a.nrows = nrows;
    a.out_row_pad = d.pads;
    rc = launch_map(ctx, a, (int)blocks.size());
    if (rc) return rc;
    if (mode == 2) {
        FillArgs f;
        f.out = d.out;
        f.off_old = d.offs;
        f.nrows = nrows;
        f.n_out = dense_len;
        f.row_pad = d.pads;
        f.col_pad = d.pads + n_tables;
        dim3 grid((unsigned)std::max<uint64_t>(1, std::min<uint64_t>((dense_len + 255) / 256, 1024)), (unsigned)n_tables);
        map_to_dense_fill_kernel<<<grid, 256, 0, ctx->stream>>>(f);
        ctx->launches++;
        GKR_CUDA_OK(ctx, cudaGetLastError());
    }
    if (n_out) *n_out = (uint32_t)n_tables;
    return GKR_OK;
}
