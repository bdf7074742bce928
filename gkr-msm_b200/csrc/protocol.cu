// Host-side protocol layer in C++: the stand-in for the reference's Rust host (no Rust toolchain in this image), mirroring
// it file by file on top of the C ABI of this library.  Orchestration, Fiat-Shamir and O(1) claim algebra run here on the
// CPU exactly as in the reference (north_star); every table-sized step is a device call.
//
//   SplitAt / GlueSplit / ZeroCheck          src/cleanup/protocols/splits.rs:120-203, zero_check.rs:17-33
//   DenseDeg2Sumcheck::prove                 src/cleanup/protocols/sumchecks/dense_eq.rs:192-221
//   VecVecDeg2Sumcheck::prove                src/cleanup/protocols/sumchecks/vecvec_eq.rs:418-450
//   DenseEqSumcheck::prove                   src/cleanup/protocols/sumcheck.rs:843-872
//   SimpleGKR::prove                         src/cleanup/protocols/gkrs/gkr.rs:45-50
//   bintree / triangle witness + protocols   src/cleanup/protocols/gkrs/bintree_add.rs:124-375, triangle_add.rs:76-232
//   PippengerEndingWG / PippengerBucketed    src/cleanup/protocols/pippenger_ending.rs:26-157
//   PushForwardState::{new, second_phase}    src/cleanup/protocols/pushforward/pushforward.rs:329-622
//   PushforwardProtocol::prove               src/cleanup/protocols/pushforward/pushforward.rs:631-847
//   LogupMainphaseProtocol::prove            src/cleanup/protocols/pushforward/logup_mainphase.rs:85-200
//   MultiOpenReduction::prove                src/cleanup/protocols/multiopen_reduction.rs:65-93
//   KnucklesOpeningProtocol::prove           src/cleanup/protocols/opening.rs:39-98
//   PippengerWG::new, Pippenger::prove       src/cleanup/protocols/pippenger.rs:30-70, 122-294
//   benchutils::run_pippenger                src/cleanup/protocols/pippenger.rs:499-559
//
// gkr-msm_b200/{protocols,pippenger}.py hold the same logic in python (the executable specification the parity tests
// compare against); gkr_run_pippenger below is the entry point a Rust `examples/pippenger` would call instead of
// `run_pippenger`.
#include <algorithm>
#include <array>
#include <cstdlib>
#include <map>
#include <memory>
#include "common.cuh"
#include "host_g1.hpp"
#include "so.hpp"
#include "transcript.hpp"

void gkr_big_mem_stats(uint64_t out[2], bool reset_peak);  // tables.cu

namespace {

using gkr::FrH;
namespace F = gkr::frh;
typedef std::array<uint64_t, 12> G1P;

struct Fail {
    int code;
};
inline void ck(int rc) {
    if (rc) throw Fail{rc};
}
[[noreturn]] inline void fail(gkr_ctx* ctx, const char* msg) {
    ctx->fail(GKR_ERR_PROTOCOL, msg);
    throw Fail{GKR_ERR_PROTOCOL};
}

// optional span tree on stderr (GKR_TRACE=1), like the tracing spans of examples/pippenger.rs:75-89; every span
// synchronises the stream on both sides, so it is for looking at a breakdown only
struct Span {
    gkr_ctx* ctx;
    const char* name;
    uint64_t t0 = 0;
    bool on, sync;
    Span(gkr_ctx* c, const char* n) : ctx(c), name(n), on(getenv("GKR_TRACE") != nullptr) {
        sync = on && getenv("GKR_TRACE")[0] != '2';  // GKR_TRACE=2: host timestamps only, no synchronisation
        if (on) {
            if (sync) cudaStreamSynchronize(ctx->stream);
            t0 = gkr_now_ns();
        }
    }
    ~Span() {
        if (on) {
            if (sync) cudaStreamSynchronize(ctx->stream);
            uint64_t m[2];
            gkr_big_mem_stats(m, true);  // table memory now / its peak since the previous span line
            fprintf(stderr, "  [gkr_run_pippenger] %9.2f ms  %-44s  tables %7.2f GiB live, peak %7.2f GiB\n", (gkr_now_ns() - t0) / 1e6, name,
                    m[0] / 1073741824.0, m[1] / 1073741824.0);
        }
    }
};

extern "C" void gkr_sumcheck_prove_stats_dump(void);
// flat accumulators next to the span tree: host time per sumcheck kind, split into object construction and rounds
struct TraceAcc {
    struct Row {
        uint64_t ns = 0, calls = 0, rounds = 0;
    };
    static std::map<std::string, Row>& rows() {
        static std::map<std::string, Row> r;
        return r;
    }
    static bool on() {
        static const bool v = getenv("GKR_TRACE") != nullptr;
        return v;
    }
    const char* name;
    uint64_t t0 = 0, rounds;
    TraceAcc(const char* n, uint64_t nr = 0) : name(n), rounds(nr) {
        if (on()) t0 = gkr_now_ns();
    }
    ~TraceAcc() {
        if (!on()) return;
        Row& r = rows()[name];
        r.ns += gkr_now_ns() - t0;
        r.calls++;
        r.rounds += rounds;
    }
    static void dump(gkr_ctx* ctx) {
        if (!on()) return;
        gkr_sumcheck_prove_stats_dump();
        static const char* kinds[3] = {"dense", "deg2 dense", "deg2 ragged"};
        for (int k = 0; k < 3; k++)
            for (int l = 0; l < 40; l++)
                if (ctx->wait_hist_n[k][l]) {
                    fprintf(stderr, "  [gkr_run_pippenger wait] %-12s 2^%-2d pairs: %5llu waits %9.2f ms  (%7.1f us each)\n", kinds[k], l,
                            (unsigned long long)ctx->wait_hist_n[k][l], ctx->wait_hist_ns[k][l] / 1e6,
                            ctx->wait_hist_ns[k][l] / 1e3 / ctx->wait_hist_n[k][l]);
                    ctx->wait_hist_n[k][l] = ctx->wait_hist_ns[k][l] = 0;
                }
        for (auto& kv : rows())
            fprintf(stderr, "  [gkr_run_pippenger acc] %9.2f ms  %6llu calls %6llu rounds  %s\n", kv.second.ns / 1e6, (unsigned long long)kv.second.calls,
                    (unsigned long long)kv.second.rounds, kv.first.c_str());
        rows().clear();
    }
};

struct TabH {
    gkr_table* h = nullptr;
    explicit TabH(gkr_table* p) : h(p) {}
    ~TabH() { gkr_table_free(h); }
};
struct VvH {
    gkr_vecvec* h = nullptr;
    explicit VvH(gkr_vecvec* p) : h(p) {}
    ~VvH() { gkr_vecvec_free(h); }
};
struct SrsH {
    gkr_srs* h = nullptr;
    explicit SrsH(gkr_srs* p) : h(p) {}
    ~SrsH() { gkr_srs_free(h); }
};
struct U32H {
    gkr_u32buf* h = nullptr;
    explicit U32H(gkr_u32buf* p) : h(p) {}
    ~U32H() { gkr_u32_free(h); }
};
struct SoH {
    gkr_so* h = nullptr;
    explicit SoH(gkr_so* p) : h(p) {}
    ~SoH() { gkr_so_destroy(h); }
};
typedef std::shared_ptr<TabH> Tab;
typedef std::shared_ptr<VvH> Vv;
typedef std::shared_ptr<SrsH> Srs;
typedef std::shared_ptr<U32H> U32;

struct Claims {
    std::vector<FrH> point, evs;
};
struct Gate;
struct Advice {
    int kind = 0;  // 0 empty, 1 VecVec polys, 2 dense tables  (SplitVecVecMapGKRAdvice, split_map_gkr.rs:65-71)
    std::vector<Vv> vv;
    std::vector<Tab> dense;
    // kind 3: not stored -- recomputed as chain.back()(.. chain[0](*base)) when the prover reaches it (bintree_witness, memory plan
    // of large instances: only the INPUT of every addition layer stays resident, its L1 / L2 images are two maps away)
    std::shared_ptr<Advice> base;
    std::vector<const Gate*> chain;
};
struct Gate {
    int gid;  // public gate id for single-gate objects (-1: stack only)
    std::vector<std::pair<int, uint32_t>> parts;
    int n_ins, n_outs;
};

const Gate AFF_L1{GKR_GATE_AFF_L1, {{GKR_GATE_AFF_L1, 1}}, 4, 3};
const Gate AFF_L1_BC2{GKR_GATE_AFF_L1_BITCHECK2, {{GKR_GATE_AFF_L1_BITCHECK2, 1}}, 6, 5};
const Gate AFF_L2{GKR_GATE_AFF_L2, {{GKR_GATE_AFF_L2, 1}}, 3, 3};
const Gate AFF_L3{GKR_GATE_AFF_L3, {{GKR_GATE_AFF_L3, 1}}, 3, 3};
const Gate PRJ_L1{GKR_GATE_PRJ_L1, {{GKR_GATE_PRJ_L1, 1}}, 6, 4};
const Gate PRJ_L2{GKR_GATE_PRJ_L2, {{GKR_GATE_PRJ_L2, 1}}, 4, 4};
const Gate PRJ_L3{GKR_GATE_PRJ_L3, {{GKR_GATE_PRJ_L3, 1}}, 4, 3};
const Gate LOGUP{GKR_GATE_LOGUP_LAYER, {{GKR_GATE_LOGUP_LAYER, 1}}, 4, 2};
const Gate ADD_INV{GKR_GATE_ADD_INVERSES, {{GKR_GATE_ADD_INVERSES, 1}}, 2, 2};
Gate ID(int n) { return Gate{-1, {{GKR_GATE_ID, (uint32_t)n}}, n, n}; }
Gate tri_l1(int layer_idx) {  // Stacked(triangle_l1, Repeated(prj_l1, layer_idx))   triangle_add.rs:128-135
    Gate g{-1, {{GKR_GATE_TRI_L1, 1}}, 12 + 6 * layer_idx, 12 + 4 * layer_idx};
    if (layer_idx) g.parts.push_back({GKR_GATE_PRJ_L1, (uint32_t)layer_idx});
    return g;
}
Gate repeated(const Gate& g, int k) { return Gate{-1, {{g.parts[0].first, (uint32_t)k}}, g.n_ins * k, g.n_outs * k}; }

// ---- small host helpers --------------------------------------------------------------------------------------------
std::vector<uint64_t> limbs_of(const std::vector<FrH>& v) {
    std::vector<uint64_t> out(4 * std::max<size_t>(v.size(), 1));
    for (size_t i = 0; i < v.size(); i++) frh_to_limbs(v[i], out.data() + 4 * i);
    return out;
}
std::vector<FrH> make_gamma_pows(const FrH& gamma, size_t count) {  // src/utils.rs:126-135 (at least [1, gamma])
    std::vector<FrH> g{F::ONE, gamma};
    for (size_t i = 2; i < count; i++) g.push_back(F::mul(g[i - 1], gamma));
    return g;
}
FrH gamma_rlc(const FrH& gamma, const std::vector<FrH>& vals) {  // sumcheck.rs:591-602
    if (vals.empty()) return F::ZERO;
    FrH ret = vals.back();
    for (size_t i = vals.size() - 1; i-- > 0;) ret = F::add(F::mul(ret, gamma), vals[i]);
    return ret;
}
std::vector<FrH> eq_poly_sequence_last(const std::vector<FrH>& pt) {  // utils.rs:222-262, last level
    std::vector<FrH> ret{F::ONE};
    for (const FrH& r : pt) {
        std::vector<FrH> nxt;
        nxt.reserve(2 * ret.size());
        for (const FrH& w : ret) {
            FrH hi = F::mul(w, r);
            nxt.push_back(F::sub(w, hi));
            nxt.push_back(hi);
        }
        ret.swap(nxt);
    }
    return ret;
}
FrH eq_sum(const std::vector<FrH>& pt, uint64_t k) {  // utils.rs:265-291
    const size_t n = pt.size();
    if (k >= ((uint64_t)1 << n)) return F::ONE;
    FrH mult = F::ONE, acc = F::ZERO;
    for (size_t i = 0; i < n; i++) {
        uint64_t left_bit = k >> (n - i - 1);
        FrH old = mult;
        if (left_bit == 1) {
            mult = F::mul(mult, pt[i]);
            acc = F::add(acc, F::sub(old, mult));
        } else {
            mult = F::mul(mult, F::sub(F::ONE, pt[i]));
        }
        k -= left_bit << (n - i - 1);
    }
    return acc;
}
std::vector<FrH> eq_trunc_evals(uint32_t num_vars, uint64_t k, const std::vector<FrH>& r) {  // verifier_polys.rs:90-96
    std::vector<FrH> ret = eq_poly_sequence_last(r);
    for (uint64_t i = k; i < ((uint64_t)1 << num_vars); i++) ret[i] = F::ZERO;
    return ret;
}
FrH eq_trunc_evaluate(uint32_t num_vars, uint64_t k, const std::vector<FrH>& r, const std::vector<FrH>& pt) {  // :98-136
    std::vector<FrH> partial{F::ONE};
    for (uint32_t i = 0; i < num_vars; i++) {
        uint32_t j = num_vars - i - 1;
        FrH t = F::add(F::sub(F::sub(F::ONE, pt[j]), r[j]), F::dbl(F::mul(r[j], pt[j])));
        partial.push_back(F::mul(partial.back(), t));
    }
    if (k >= ((uint64_t)1 << num_vars)) return partial[num_vars];
    FrH multiplier = F::ONE, acc = F::ZERO;
    for (uint32_t i = 0; i < num_vars; i++) {
        uint64_t left_bit = k >> (num_vars - i - 1);
        FrH m_ = multiplier;
        FrH omp = F::sub(F::ONE, pt[i]), omr = F::sub(F::ONE, r[i]);
        if (left_bit == 1) {
            multiplier = F::mul(F::mul(multiplier, pt[i]), r[i]);
            acc = F::add(acc, F::mul(F::mul(F::mul(m_, omp), omr), partial[num_vars - i - 1]));
        } else {
            multiplier = F::mul(F::mul(multiplier, omp), omr);
        }
        k -= left_bit << (num_vars - i - 1);
    }
    return acc;
}

// ---- thin C++ view of the C ABI -------------------------------------------------------------------------------------
struct Dev {
    gkr_ctx* ctx;
    gkr::ProofTranscript2* tr;

    Tab upload(const uint64_t* limbs, uint64_t n) {
        gkr_table* t = nullptr;
        ck(gkr_table_upload(ctx, limbs, n, &t));
        return std::make_shared<TabH>(t);
    }
    Tab upload(const std::vector<FrH>& v) { return upload(limbs_of(v).data(), v.size()); }
    std::vector<FrH> download(const Tab& t) {
        uint64_t n = gkr_table_len(t->h);
        std::vector<uint64_t> buf(4 * std::max<uint64_t>(n, 1));
        ck(gkr_table_download(ctx, t->h, buf.data()));
        std::vector<FrH> out(n);
        for (uint64_t i = 0; i < n; i++) out[i] = frh_from_limbs(buf.data() + 4 * i);
        return out;
    }
    Tab eq_table(const std::vector<FrH>& point) {
        uint64_t one[4];
        frh_to_limbs(F::ONE, one);
        gkr_table* t = nullptr;
        ck(gkr_eq_table(ctx, limbs_of(point).data(), (uint32_t)point.size(), one, &t));
        return std::make_shared<TabH>(t);
    }
    static void parts_arrays(const Gate& g, std::vector<int>& pg, std::vector<uint32_t>& pr) {
        for (auto& p : g.parts) {
            pg.push_back(p.first);
            pr.push_back(p.second);
        }
    }
    // Vec::algfn_map (split_kind < 0) / algfn_map_split (0 = LO(var), 1 = HI(var))
    std::vector<Tab> map_dense(const Gate& g, const std::vector<Tab>& tabs, int split_kind = -1, uint32_t var = 0, uint32_t bundle = 1) {
        std::vector<int> pg;
        std::vector<uint32_t> pr;
        parts_arrays(g, pg, pr);
        std::vector<gkr_table*> in;
        for (int i = 0; i < g.n_ins; i++) in.push_back(tabs[i]->h);
        gkr_table* out[256] = {nullptr};
        uint32_t n = 0;
        ck(gkr_map_dense(ctx, pg.data(), pr.data(), (uint32_t)pg.size(), in.data(), (uint32_t)in.size(), split_kind, var, bundle, out, &n));
        std::vector<Tab> res;
        for (uint32_t i = 0; i < n; i++) res.push_back(std::make_shared<TabH>(out[i]));
        return res;
    }
    // mode 0: vecvec_map, 1: vecvec_map_split at LO(0), 2: vecvec_map_split_to_dense
    void map_vecvec(const Gate& g, const std::vector<Vv>& polys, int mode, uint32_t bundle, std::vector<Vv>* out_vv, std::vector<Tab>* out_dense) {
        std::vector<int> pg;
        std::vector<uint32_t> pr;
        parts_arrays(g, pg, pr);
        std::vector<gkr_vecvec*> in;
        for (int i = 0; i < g.n_ins; i++) in.push_back(polys[i]->h);
        void* out[256] = {nullptr};
        uint32_t n = 0;
        ck(gkr_map_vecvec(ctx, pg.data(), pr.data(), (uint32_t)pg.size(), in.data(), (uint32_t)in.size(), mode, bundle, out, &n));
        for (uint32_t i = 0; i < n; i++) {
            if (mode == 2) out_dense->push_back(std::make_shared<TabH>((gkr_table*)out[i]));
            else out_vv->push_back(std::make_shared<VvH>((gkr_vecvec*)out[i]));
        }
    }
    // terms: (table or null = all ones, coef, src_off, dst_off, len)
    struct Term {
        Tab t;
        FrH coef;
        uint64_t src_off, dst_off, len;
    };
    Tab lincomb(const std::vector<Term>& terms, uint64_t out_len) {
        std::vector<gkr_table*> src;
        std::vector<uint64_t> coefs(4 * terms.size()), so, dof, ln;
        for (size_t i = 0; i < terms.size(); i++) {
            src.push_back(terms[i].t ? terms[i].t->h : nullptr);
            frh_to_limbs(terms[i].coef, coefs.data() + 4 * i);
            so.push_back(terms[i].src_off);
            dof.push_back(terms[i].dst_off);
            ln.push_back(terms[i].len);
        }
        gkr_table* t = nullptr;
        ck(gkr_table_lincomb(ctx, (uint32_t)terms.size(), src.data(), coefs.data(), so.data(), dof.data(), ln.data(), out_len, &t));
        return std::make_shared<TabH>(t);
    }
    G1P msm(gkr_srs* srs, const Tab& scalars, uint64_t n, uint64_t first = 0) {
        G1P out;
        ck(gkr_msm_g1(ctx, srs, first, scalars->h, n, out.data()));
        return out;
    }
    // sum_i coefs[i] * pts[i] for a handful of points: a tiny MSM on the device
    G1P g1_lincomb(const std::vector<FrH>& coefs, const std::vector<G1P>& pts) {
        std::vector<uint64_t> flat(12 * pts.size());
        for (size_t i = 0; i < pts.size(); i++) std::memcpy(flat.data() + 12 * i, pts[i].data(), 96);
        gkr_srs* s = nullptr;
        ck(gkr_srs_upload(ctx, flat.data(), pts.size(), 0, &s));
        Srs srs = std::make_shared<SrsH>(s);
        Tab sc = upload(coefs);
        return msm(srs->h, sc, coefs.size());
    }
    void write_points(const std::vector<G1P>& pts) {  // proof_transcript.rs:64-69
        std::vector<uint8_t> buf(48 * pts.size());
        for (size_t i = 0; i < pts.size(); i++) gkr::g1h::serialize_compressed(pts[i].data(), buf.data() + 48 * i);
        tr->write_raw_msg(buf.data(), buf.size());
    }
    // GenericSumcheckProtocol::prove through the C entry (point comes back reversed like the reference)
    void sumcheck_prove(gkr_so* so, uint32_t num_rounds, std::vector<FrH>* point, std::vector<FrH>* final_evals) {
        gkr_transcript* th = reinterpret_cast<gkr_transcript*>(tr);  // gkr_transcript holds exactly one ProofTranscript2
        std::vector<uint64_t> pt(4 * std::max<uint32_t>(num_rounds, 1)), fe(4 * gkr_so_num_polys(so));
        uint64_t claim[4];
        ck(gkr_sumcheck_prove(th, so, num_rounds, claim, pt.data(), fe.data()));
        point->resize(num_rounds);
        for (uint32_t i = 0; i < num_rounds; i++) (*point)[i] = frh_from_limbs(pt.data() + 4 * i);
        final_evals->resize(gkr_so_num_polys(so));
        for (size_t i = 0; i < final_evals->size(); i++) (*final_evals)[i] = frh_from_limbs(fe.data() + 4 * i);
    }
};

// ---- sumcheck layers -------------------------------------------------------------------------------------------------
struct Layer {
    virtual ~Layer() {}
    virtual Claims prove(Dev& d, const Claims& claims, const Advice& advice) = 0;
};
typedef std::vector<std::unique_ptr<Layer>> Layers;

struct DenseDeg2Sumcheck : Layer {  // dense_eq.rs:192-221
    Gate gate;
    uint32_t num_vars;
    DenseDeg2Sumcheck(const Gate& g, uint32_t nv) : gate(g), num_vars(nv) {}
    Claims prove(Dev& d, const Claims& claims, const Advice& advice) override {
        if ((int)advice.dense.size() != gate.n_ins) fail(d.ctx, "DenseDeg2Sumcheck: wrong number of input tables");
        FrH gamma = d.tr->challenge(128);
        std::vector<FrH> gp = make_gamma_pows(gamma, gate.n_outs);
        FrH claim = claims.evs[0];
        for (size_t i = 1; i < claims.evs.size(); i++) claim = F::add(claim, F::mul(gp[i], claims.evs[i]));
        std::vector<int> pg;
        std::vector<uint32_t> pr;
        Dev::parts_arrays(gate, pg, pr);
        std::vector<gkr_table*> in;
        for (auto& t : advice.dense) in.push_back(t->h);
        uint64_t cl[4];
        frh_to_limbs(claim, cl);
        gkr_so* so = nullptr;
        {
            TraceAcc ta("deg2 dense: create");
            ck(gkr_so_create_deg2_dense(d.ctx, pg.data(), pr.data(), (uint32_t)pg.size(), in.data(), (uint32_t)in.size(), limbs_of(gp).data(), cl,
                                        limbs_of(claims.point).data(), (uint32_t)claims.point.size(), &so));
        }
        SoH guard(so);
        Claims out;
        {
            TraceAcc ta("deg2 dense: rounds", num_vars);
            d.sumcheck_prove(so, num_vars, &out.point, &out.evs);
        }
        d.tr->write_scalars(out.evs.data(), out.evs.size());
        return out;
    }
};

struct VecVecDeg2Sumcheck : Layer {  // vecvec_eq.rs:418-450
    Gate gate;
    uint32_t num_vars, nvv;
    VecVecDeg2Sumcheck(const Gate& g, uint32_t nv, uint32_t vertical) : gate(g), num_vars(nv), nvv(vertical) {}
    Claims prove(Dev& d, const Claims& claims, const Advice& advice) override {
        if ((int)advice.vv.size() != gate.n_ins) fail(d.ctx, "VecVecDeg2Sumcheck: wrong number of input polynomials");
        FrH gamma = d.tr->challenge(128);
        std::vector<FrH> gp = make_gamma_pows(gamma, gate.n_outs);
        FrH claim = claims.evs[0];
        for (size_t i = 1; i < claims.evs.size(); i++) claim = F::add(claim, F::mul(gp[i], claims.evs[i]));
        std::vector<gkr_vecvec*> in;
        for (auto& p : advice.vv) in.push_back(p->h);
        uint64_t cl[4];
        frh_to_limbs(claim, cl);
        gkr_so* so = nullptr;
        {
            TraceAcc ta("deg2 vecvec: create");
            ck(gkr_so_create_deg2_vecvec(d.ctx, gate.gid, in.data(), (uint32_t)in.size(), limbs_of(gp).data(), cl, limbs_of(claims.point).data(),
                                         (uint32_t)claims.point.size(), nvv, &so));
        }
        SoH guard(so);
        Claims out;
        {
            TraceAcc ta("deg2 vecvec: rounds (sparse + dense tail)", num_vars);
            d.sumcheck_prove(so, num_vars, &out.point, &out.evs);
        }
        out.evs.pop_back();  // poly_evs.pop(): the eq evaluation is not sent (vecvec_eq.rs:445)
        d.tr->write_scalars(out.evs.data(), out.evs.size());
        return out;
    }
};

struct SplitAt : Layer {  // splits.rs:120-170
    bool hi;
    uint32_t var;
    uint32_t bundle;
    SplitAt(bool is_hi, uint32_t v, uint32_t b) : hi(is_hi), var(v), bundle(b) {}
    Claims prove(Dev& d, const Claims& claims, const Advice&) override {
        FrH r = d.tr->challenge(128);
        Claims out;
        out.point = claims.point;
        std::vector<FrH> l, rr;
        for (size_t i = 0; i < claims.evs.size(); i++) (((i / bundle) & 1) ? rr : l).push_back(claims.evs[i]);
        for (size_t i = 0; i < std::min(l.size(), rr.size()); i++) out.evs.push_back(F::add(l[i], F::mul(r, F::sub(rr[i], l[i]))));
        size_t pos = hi ? var : out.point.size() - var;
        out.point.insert(out.point.begin() + pos, r);
        return out;
    }
};
struct GlueSplit : Layer {  // splits.rs:172-203
    Claims prove(Dev& d, const Claims& claims, const Advice&) override {
        FrH r = d.tr->challenge(128);
        const std::vector<FrH>& e = claims.evs;
        Claims out;
        out.point = claims.point;
        out.evs = {F::add(e[0], F::mul(r, F::sub(e[2], e[0]))), F::add(e[1], F::mul(r, F::sub(e[3], e[1]))),
                   F::add(e[4], F::mul(r, F::sub(e[5], e[4])))};
        out.point.push_back(r);
        return out;
    }
};
struct ZeroCheck : Layer {  // zero_check.rs:17-33
    Claims prove(Dev&, const Claims& claims, const Advice&) override {
        Claims out = claims;
        out.evs.push_back(F::ZERO);
        out.evs.push_back(F::ZERO);
        return out;
    }
};

Advice advice_map(Dev& d, const Advice& a, const Gate& gate);
Claims simple_gkr_prove(Dev& d, Layers& layers, Claims claims, std::vector<Advice>& advices) {  // gkr.rs:45-50
    if (advices.size() != layers.size()) fail(d.ctx, "SimpleGKR: advice / layer count mismatch");
    // recomputed advices (kind 3): the layers of one addition are visited L3, L2, L1 -- the L1 image computed on the way to the
    // L2 image is kept for the next visit and dropped after it
    const Advice* stash_base = nullptr;
    Advice stash;
    for (size_t i = layers.size(); i-- > 0;) {
        if (advices.back().kind == 3) {
            const Advice lazy = advices.back();
            Advice cur;
            size_t from = 0;
            if (stash_base == lazy.base.get() && lazy.chain.size() == 1) {
                cur = stash;
                from = 1;
            } else {
                cur = *lazy.base;
            }
            stash_base = nullptr;
            stash = Advice();
            for (size_t k = from; k < lazy.chain.size(); k++) {
                Span sp(d.ctx, "    bintree map (recomputed)");
                cur = advice_map(d, cur, *lazy.chain[k]);
                if (k == 0 && lazy.chain.size() > 1) {
                    stash_base = lazy.base.get();
                    stash = cur;
                }
            }
            claims = layers[i]->prove(d, claims, cur);
        } else {
            claims = layers[i]->prove(d, claims, advices.back());
        }
        advices.pop_back();
    }
    return claims;
}

// ---- witness builders ------------------------------------------------------------------------------------------------
Advice advice_map(Dev& d, const Advice& a, const Gate& gate) {
    Advice out;
    if (a.kind == 1) {
        out.kind = 1;
        d.map_vecvec(gate, a.vv, 0, 1, &out.vv, nullptr);
    } else {
        out.kind = 2;
        out.dense = d.map_dense(gate, a.dense);
    }
    return out;
}
Advice advice_map_split(Dev& d, const Advice& a, const Gate& gate, uint32_t layer_idx, uint32_t row_logsize, uint32_t bundle) {
    Advice out;
    if (a.kind == 1) {
        if (layer_idx + 2 == row_logsize) {
            out.kind = 2;
            d.map_vecvec(gate, a.vv, 2, bundle, nullptr, &out.dense);
        } else {
            out.kind = 1;
            d.map_vecvec(gate, a.vv, 1, bundle, &out.vv, nullptr);
        }
    } else {
        out.kind = 2;
        out.dense = d.map_dense(gate, a.dense, 0, 0, bundle);
    }
    return out;
}
// recompute: keep only the input of every addition layer resident and leave its L1 / L2 images as recipes (Advice kind 3) that
// simple_gkr_prove replays when it reaches them -- the reference keeps all three per layer (bintree_add.rs:149-170), which is
// 70 % of the prover's memory; the images are the same tables either way, so the proof does not change.
std::vector<Advice> bintree_witness(Dev& d, Advice advice, uint32_t row_logsize, uint32_t num_adds, bool do_bitcheck, bool recompute) {  // bintree_add.rs:173-202
    std::vector<Advice> advices;
    for (uint32_t add_idx = 0; add_idx < num_adds; add_idx++) {
        const Gate& g1 = add_idx == 0 ? AFF_L1 : PRJ_L1;
        const Gate& g2 = add_idx == 0 ? AFF_L2 : PRJ_L2;
        std::shared_ptr<Advice> base;
        for (int step = 0; step < 3; step++) {
            const bool last = add_idx + 1 == num_adds;
            Advice nxt;
            Span sp(d.ctx, step == 0 ? "    bintree L1 map" : (step == 1 ? "    bintree L2 map" : "    bintree L3 map+split"));
            if (step == 0) nxt = advice_map(d, advice, g1);
            else if (step == 1) nxt = advice_map(d, advice, g2);
            else if (!last) nxt = advice_map_split(d, advice, add_idx == 0 ? AFF_L3 : PRJ_L3, add_idx, row_logsize, 3);
            if (recompute && step == 0) base = std::make_shared<Advice>(advice);
            if (recompute && step > 0) {
                Advice lazy;
                lazy.kind = 3;
                lazy.base = base;
                lazy.chain.push_back(&g1);
                if (step == 2) lazy.chain.push_back(&g2);
                advices.push_back(lazy);
            } else {
                advices.push_back(advice);
            }
            if (add_idx == 0 && step == 0 && do_bitcheck) advices.push_back(Advice());
            if (!(step == 2 && last)) advice = nxt;
        }
        if (add_idx + 1 != num_adds) advices.push_back(Advice());
    }
    return advices;
}
Layers bintree_protocol(uint32_t num_vars, uint32_t num_adds, uint32_t row_logsize, bool do_bitcheck) {  // bintree_add.rs:228-375
    Layers layers;
    const uint32_t nvv = num_vars - row_logsize;
    for (uint32_t i = 0; i < num_adds; i++) {
        for (int step = 0; step < 3; step++) {
            const uint32_t nv = num_vars - i - 1;
            if (i == 0) {
                const Gate& g = step == 0 ? (do_bitcheck ? AFF_L1_BC2 : AFF_L1) : (step == 1 ? AFF_L2 : AFF_L3);
                layers.emplace_back(new VecVecDeg2Sumcheck(g, nv, nvv));
            } else {
                const Gate& g = step == 0 ? PRJ_L1 : (step == 1 ? PRJ_L2 : PRJ_L3);
                if (i + 1 < row_logsize) layers.emplace_back(new VecVecDeg2Sumcheck(g, nv, nvv));
                else layers.emplace_back(new DenseDeg2Sumcheck(g, nv));
            }
            if (i == 0 && step == 0 && do_bitcheck) layers.emplace_back(new ZeroCheck());
        }
        if (i != num_adds - 1) layers.emplace_back(new SplitAt(false, 0, 3));
    }
    return layers;
}
std::vector<Advice> triangle_witness(Dev& d, std::vector<Tab> tables, uint32_t num_vars, uint32_t hi) {  // triangle_add.rs:88-157
    const uint32_t num_layers = num_vars - hi;
    std::vector<Advice> advices;
    std::vector<Tab> advice = tables;
    for (uint32_t layer_idx = 0; layer_idx <= num_layers; layer_idx++) {
        for (int step = 0; step < 3; step++) {
            std::vector<Tab> nxt;
            bool have_next = true;
            if (step == 0) nxt = d.map_dense(tri_l1(layer_idx), advice);
            else if (step == 1) nxt = d.map_dense(repeated(PRJ_L2, layer_idx + 3), advice);
            else if (num_layers == layer_idx) have_next = false;
            else nxt = d.map_dense(repeated(PRJ_L3, layer_idx + 3), advice, 1, hi, 3);
            Advice a;
            a.kind = 2;
            a.dense = advice;
            advices.push_back(a);
            if (have_next) advice = nxt;
        }
        if (layer_idx < num_layers) advices.push_back(Advice());
    }
    return advices;
}
Layers triangle_protocol(uint32_t num_vars, uint32_t hi) {  // triangle_add.rs:199-232
    const uint32_t num_layers = num_vars - hi;
    Layers layers;
    for (uint32_t layer_idx = 0; layer_idx <= num_layers; layer_idx++) {
        const uint32_t nv = num_vars - layer_idx;
        layers.emplace_back(new DenseDeg2Sumcheck(tri_l1(layer_idx), nv));
        layers.emplace_back(new DenseDeg2Sumcheck(repeated(PRJ_L2, layer_idx + 3), nv));
        layers.emplace_back(new DenseDeg2Sumcheck(repeated(PRJ_L3, layer_idx + 3), nv));
        if (layer_idx < num_layers) layers.emplace_back(new SplitAt(true, hi, 3));
    }
    return layers;
}

struct PippengerEndingWG {  // pippenger_ending.rs:32-95 (the reference builds the bintree witness twice; once suffices)
    std::vector<Advice> bintree_advices, triangle_advices;
    PippengerEndingWG(Dev& d, uint32_t multirow_vars, uint32_t bucket_vars, uint32_t horizontal_vars, const std::vector<Vv>& inputs, bool recompute) {
        Advice in;
        in.kind = 1;
        in.vv = inputs;
        Advice l2_last;  // input of the last L3 step (the last stored advice unless it is a recipe)
        bintree_advices = bintree_witness(d, in, horizontal_vars, horizontal_vars, true, recompute);
        if (bintree_advices.back().kind == 3) {
            l2_last = *bintree_advices.back().base;
            for (const Gate* g : bintree_advices.back().chain) l2_last = advice_map(d, l2_last, *g);
        } else {
            l2_last = bintree_advices.back();
        }
        Advice last = advice_map(d, l2_last, horizontal_vars - 1 == 0 ? AFF_L3 : PRJ_L3);
        l2_last = Advice();
        std::vector<Tab> split_l1 = d.map_dense(ID(3), last.dense, 1, multirow_vars, 3);
        std::vector<Tab> split_l2 = d.map_dense(ID(6), split_l1, 1, multirow_vars, 3);
        triangle_advices = triangle_witness(d, split_l2, multirow_vars + bucket_vars - 2, multirow_vars);
    }
    const std::vector<Tab>& last() const { return triangle_advices.back().dense; }
};

// ---- commitment keys -------------------------------------------------------------------------------------------------
struct Keys {
    gkr_srs* srs;
    G1P g0;
    const gkr_knuckles* knuckles;
    uint32_t num_vars;
    FrH k;
};

// ---- pushforward state (pushforward.rs:329-622) ------------------------------------------------------------------------
struct PushForwardState {
    uint32_t y_size, y_logsize, d_logsize, x_logsize, clm, n_comms, c_log;
    uint64_t x_size;
    Tab p_0, p_1, d, c, ac_d, ac_c, eq_c, eq_d, c_pull, d_pull;
    U32 d_idx, c_idx;
    std::vector<Vv> image;
    Srs d_all, c_all;
    std::vector<G1P> c_comm, d_comm, c_pull_comm, d_pull_comm;
    G1P p_0_comm, p_1_comm, ac_c_comm, ac_d_comm;

    PushForwardState(Dev& dv, const Keys& key, const uint64_t* px, const uint64_t* py, const uint64_t* coefs, uint32_t y_size_, uint32_t y_logsize_,
                     uint32_t d_logsize_, uint32_t x_logsize_, uint32_t clm_)
        : y_size(y_size_), y_logsize(y_logsize_), d_logsize(d_logsize_), x_logsize(x_logsize_), clm(clm_) {
        gkr_ctx* ctx = dv.ctx;
        if (key.num_vars != x_logsize + clm) fail(ctx, "commitment key has the wrong number of variables");
        x_size = (uint64_t)1 << x_logsize;
        const uint32_t nb = 1u << d_logsize;
        const uint64_t m = (uint64_t)y_size * x_size;
        std::vector<uint32_t> lens((size_t)y_size * nb);
        uint64_t zero[4] = {0, 0, 0, 0}, one[4];
        frh_to_limbs(F::ONE, one);
        uint64_t pads[12];
        std::memcpy(pads, zero, 32);
        std::memcpy(pads + 4, one, 32);
        std::memcpy(pads + 8, zero, 32);
        auto u32p = [&](const uint32_t* v, uint64_t n) {
            gkr_u32buf* b = nullptr;
            ck(gkr_u32_upload(ctx, v, n, &b));
            return std::make_shared<U32H>(b);
        };
        auto u32 = [&](const std::vector<uint32_t>& v) { return u32p(v.data(), v.size()); };
        std::unique_ptr<Span> s2;
        if (d_logsize <= 13) {
            // digits, in-bucket ranks and bucket contents by a stable counting sort on the device: only the scalars are uploaded
            U32 order;
            {
                Span s1(ctx, "  state: bucketize (device)");
                gkr_u32buf *dg = nullptr, *ct = nullptr, *po = nullptr;
                ck(gkr_pushforward_bucketize_dev(ctx, coefs, x_size, y_size, d_logsize, &dg, &ct, &po, lens.data()));
                d_idx = std::make_shared<U32H>(dg);
                c_idx = std::make_shared<U32H>(ct);
                order = std::make_shared<U32H>(po);
            }
            s2.reset(new Span(ctx, "  state: images + tables"));
            p_0 = dv.upload(px, x_size);
            p_1 = dv.upload(py, x_size);
            const gkr_table* srcs[3] = {p_0->h, p_1->h, nullptr};
            gkr_vecvec* outs[3] = {nullptr, nullptr, nullptr};
            ck(gkr_vecvec_gather_multi_dev(ctx, srcs, 3, order->h, lens.data(), (uint32_t)lens.size(), pads, pads, x_logsize, y_logsize + d_logsize, outs));
            for (int k = 0; k < 3; k++) image.push_back(std::make_shared<VvH>(outs[k]));
        } else {
            // index matrices: uninitialised storage, first touched (page-faulted) by the bucketize threads themselves
            std::unique_ptr<uint32_t[]> digits_(new uint32_t[m]), counter_(new uint32_t[m]), order_(new uint32_t[m]);
            uint32_t *digits = digits_.get(), *counter = counter_.get(), *order = order_.get();
            {
                Span s1(ctx, "  state: bucketize (host)");
                ck(gkr_pushforward_bucketize(coefs, x_size, y_size, d_logsize, digits, counter, order, lens.data()));
            }
            s2.reset(new Span(ctx, "  state: images + tables"));
            p_0 = dv.upload(px, x_size);
            p_1 = dv.upload(py, x_size);
            // bucket images of (x, y, 1): row (y, digit) holds the coordinates of the bucket's points in input order
            const gkr_table* srcs[3] = {p_0->h, p_1->h, nullptr};
            gkr_vecvec* outs[3] = {nullptr, nullptr, nullptr};
            ck(gkr_vecvec_gather_multi(ctx, srcs, 3, order, lens.data(), (uint32_t)lens.size(), pads, pads, x_logsize, y_logsize + d_logsize, outs));
            for (int k = 0; k < 3; k++) image.push_back(std::make_shared<VvH>(outs[k]));
            d_idx = u32p(digits, m);
            c_idx = u32p(counter, m);
        }
        auto to_field = [&](const U32& b, int negate) {
            gkr_table* t = nullptr;
            ck(gkr_table_from_u32(ctx, b->h, negate, &t));
            return std::make_shared<TabH>(t);
        };
        d = to_field(d_idx, 0);
        c = to_field(c_idx, 0);
        // access counts from the bucket sizes: ac_d[v] = #incidences with digit v; ac_c[v] = #buckets longer than v
        std::vector<uint32_t> acd(nb, 0), acc_(x_size, 0), len_hist(x_size + 2, 0);
        uint32_t max_len = 0;
        for (uint32_t y = 0; y < y_size; y++)
            for (uint32_t b = 0; b < nb; b++) {
                uint32_t l = lens[(size_t)y * nb + b];
                acd[b] += l;
                len_hist[std::min<uint64_t>(l, x_size + 1)]++;
                max_len = std::max(max_len, l);
            }
        {
            uint64_t cum = 0;
            for (uint64_t v = 0; v < x_size; v++) {
                cum += len_hist[v];
                acc_[v] = (uint32_t)(lens.size() - cum);
            }
        }
        ac_d = to_field(u32(acd), 1);
        ac_c = to_field(u32(acc_), 1);
        // c / d commitments: ONE bucket accumulation for all chunks, then running sums (pushforward.rs:398-456, 504-524)
        s2.reset(new Span(ctx, "  state: c/d bucket sums + running sums"));
        n_comms = (y_size + (1u << clm) - 1) >> clm;
        c_log = 1;
        while (((uint64_t)1 << c_log) < max_len) c_log++;  // counters run up to max_len - 1
        gkr_srs* s = nullptr;
        ck(gkr_g1_bucket_sums_rows(ctx, key.srs, d_idx->h, x_logsize, clm, d_logsize, &s));
        d_all = std::make_shared<SrsH>(s);
        ck(gkr_g1_bucket_sums_rows(ctx, key.srs, c_idx->h, x_logsize, clm, c_log, &s));
        c_all = std::make_shared<SrsH>(s);
        d_comm.resize(n_comms);
        c_comm.resize(n_comms);
        ck(gkr_g1_weighted_bucket_sums(ctx, d_all->h, 0, d_logsize, n_comms, d_comm[0].data()));
        ck(gkr_g1_weighted_bucket_sums(ctx, c_all->h, 0, c_log, n_comms, c_comm[0].data()));
        s2.reset(new Span(ctx, "  state: 4 MSM commitments"));
        {  // four commitments over the same bases: one shared sort / accumulation / reduction below 2^19 points
            const gkr_table* tabs[4] = {p_0->h, p_1->h, ac_c->h, ac_d->h};
            const uint64_t lens4[4] = {x_size, x_size, x_size, nb};
            uint64_t out4[48];
            ck(gkr_msm_g1_multi(ctx, key.srs, 0, tabs, lens4, 4, out4));
            std::memcpy(p_0_comm.data(), out4, 96);
            std::memcpy(p_1_comm.data(), out4 + 12, 96);
            std::memcpy(ac_c_comm.data(), out4 + 24, 96);
            std::memcpy(ac_d_comm.data(), out4 + 36, 96);
        }
    }

    void second_phase(Dev& dv, const std::vector<FrH>& r) {  // pushforward.rs:572-622
        gkr_ctx* ctx = dv.ctx;
        if (r.size() != y_logsize + d_logsize + x_logsize) fail(ctx, "second_phase: wrong point length");
        eq_d = dv.eq_table(std::vector<FrH>(r.begin() + y_logsize, r.begin() + y_logsize + d_logsize));
        eq_c = dv.eq_table(std::vector<FrH>(r.begin() + y_logsize + d_logsize, r.end()));
        gkr_table* t = nullptr;
        ck(gkr_table_gather(ctx, eq_c->h, c_idx->h, &t));
        c_pull = std::make_shared<TabH>(t);
        ck(gkr_table_gather(ctx, eq_d->h, d_idx->h, &t));
        d_pull = std::make_shared<TabH>(t);
        // msm_nonaff over the bucket bases with eq as scalars (pushforward.rs:598-604) == commit(c_pull chunk)
        c_pull_comm.resize(n_comms);
        d_pull_comm.resize(n_comms);
        ck(gkr_msm_g1_batch(ctx, c_all->h, 0, (uint64_t)1 << c_log, n_comms, eq_c->h, (uint64_t)1 << c_log, c_pull_comm[0].data()));
        ck(gkr_msm_g1_batch(ctx, d_all->h, 0, (uint64_t)1 << d_logsize, n_comms, eq_d->h, (uint64_t)1 << d_logsize, d_pull_comm[0].data()));
    }
};

// ---- DenseEqSumcheck (sumcheck.rs:831-872) -----------------------------------------------------------------------------
gkr_so* dense_eq_so(Dev& d, const Gate& gate, const std::vector<Tab>& tables, const std::vector<FrH>& point, const std::vector<FrH>& evs,
                    const FrH& gamma, std::vector<Tab>* keep) {
    Tab eq = point.empty() ? d.upload(std::vector<FrH>{F::ONE}) : d.eq_table(point);
    keep->push_back(eq);
    std::vector<FrH> gp = make_gamma_pows(gamma, gate.n_outs);
    std::vector<gkr_table*> in;
    for (auto& t : tables) in.push_back(t->h);
    in.push_back(eq->h);
    uint64_t cl[4];
    frh_to_limbs(gamma_rlc(gamma, evs), cl);
    gkr_so* so = nullptr;
    ck(gkr_so_create_dense(d.ctx, GKR_SO_EQ_GAMMA, gate.gid, 0, limbs_of(gp).data(), (uint32_t)gp.size(), in.data(), (uint32_t)in.size(),
                           (uint32_t)point.size(), cl, &so));
    return so;
}
Claims dense_eq_prove(Dev& d, const Gate& gate, uint32_t num_vars, const Claims& claims, const std::vector<Tab>& advice) {
    FrH gamma = d.tr->challenge(128);
    Claims out;
    if (num_vars == 0) {  // no rounds: the final evaluations are the single entries themselves
        for (auto& t : advice) out.evs.push_back(d.download(t)[0]);
        d.tr->write_scalars(out.evs.data(), out.evs.size());
        return out;
    }
    std::vector<Tab> keep;
    gkr_so* raw = nullptr;
    {
        TraceAcc ta("dense eq: create");
        raw = dense_eq_so(d, gate, advice, claims.point, claims.evs, gamma, &keep);
    }
    SoH so(raw);
    {
        TraceAcc ta("dense eq: rounds", num_vars);
        d.sumcheck_prove(so.h, num_vars, &out.point, &out.evs);
    }
    out.evs.pop_back();
    d.tr->write_scalars(out.evs.data(), out.evs.size());
    return out;
}

// ---- logup main phase (logup_mainphase.rs:64-200) ------------------------------------------------------------------------
typedef std::array<Tab, 2> Frac;  // [numerators, denominators]
std::vector<Claims> logup_mainphase_prove(Dev& d, std::vector<uint32_t> logsizes, const FrH& claims, std::vector<Frac> inp) {
    gkr_ctx* ctx = d.ctx;
    if (logsizes.size() < 2 || logsizes[0] != logsizes[1]) fail(ctx, "logup: bad logsizes");
    // make_witness (:85-133)
    std::reverse(inp.begin(), inp.end());
    std::vector<Frac> layers;
    layers.push_back(inp.back());
    inp.pop_back();
    layers.push_back(inp.back());
    inp.pop_back();
    size_t i = 0;
    for (;;) {
        uint64_t next_size = inp.empty() ? 1 : gkr_table_len(inp.back()[0]->h);
        uint64_t curr_size = gkr_table_len(layers[i][0]->h);
        std::vector<Tab> ins{layers[i][0], layers[i][1], layers[i + 1][0], layers[i + 1][1]};
        if (curr_size == next_size) {
            std::vector<Tab> o = d.map_dense(LOGUP, ins);
            layers.push_back(Frac{o[0], o[1]});
            if (!inp.empty()) {
                layers.push_back(inp.back());
                inp.pop_back();
            } else {
                break;
            }
            i += 2;
        } else {
            if (curr_size < next_size) fail(ctx, "logup: logsizes must be non-increasing");
            std::vector<Tab> o = d.map_dense(LOGUP, ins, 1, 0, 2);  // AlgFnUtils::map_split_hi
            layers.push_back(Frac{o[0], o[1]});
            layers.push_back(Frac{o[2], o[3]});
            i += 2;
        }
    }
    Frac top = layers.back();
    layers.pop_back();
    FrH num = d.download(top[0])[0], denom = d.download(top[1])[0];
    if (F::is_zero(denom) || num != F::mul(denom, claims)) fail(ctx, "logup: the fraction sum does not match the claim");
    FrH nd[2] = {num, denom};
    d.tr->write_scalars(nd, 2);
    // prove (:135-200)
    Claims running;
    running.evs = {num, denom};
    uint32_t curr = 0;
    std::vector<Claims> accumulated;
    Claims tmp;
    for (;;) {
        uint32_t incoming = logsizes.back();
        Frac adv_r = layers.back();
        layers.pop_back();
        Frac adv_l = layers.back();
        layers.pop_back();
        Claims claim_4 = dense_eq_prove(d, LOGUP, curr, running, {adv_l[0], adv_l[1], adv_r[0], adv_r[1]});
        if (incoming == curr) {
            if (logsizes.size() == 2) {
                tmp = claim_4;
                break;
            }
            running.point = claim_4.point;
            running.evs = {claim_4.evs[0], claim_4.evs[1]};
            Claims acc;
            acc.point = claim_4.point;
            acc.evs = {claim_4.evs[2], claim_4.evs[3]};
            accumulated.push_back(acc);
            logsizes.pop_back();
        } else {
            SplitAt s(true, 0, 2);
            running = s.prove(d, claim_4, Advice());
            curr += 1;
        }
    }
    accumulated.push_back(tmp);
    std::reverse(accumulated.begin(), accumulated.end());
    return accumulated;
}

// ---- pushforward protocol (pushforward.rs:631-847) -------------------------------------------------------------------------
struct FinalClaims {
    FrH gamma;
    Claims matrix, ac_c, ac_d;
};
FinalClaims pushforward_prove(Dev& d, uint32_t xl, uint32_t yl, uint32_t y_size, uint32_t dl, const Claims& claims, PushForwardState& st) {
    gkr_ctx* ctx = d.ctx;
    std::vector<FrH> point = claims.point, evs = claims.evs;
    evs[1] = F::sub(evs[1], F::ONE);
    if (point.size() != yl + dl + xl || evs.size() != 3) fail(ctx, "pushforward: malformed input claims");
    std::vector<FrH> r_y(point.begin(), point.begin() + yl);
    const uint64_t x_size = (uint64_t)1 << xl, matrix_size = x_size * y_size, full = (uint64_t)1 << (xl + yl);
    const uint32_t matrix_logsize = xl + yl;
    uint8_t raw[256];
    d.tr->raw_challenge(raw, 256);  // challenge_vec(4, 512), pushforward.rs:689
    FrH psi = F::from_le_bytes_mod_order(raw, 64), tau_c = F::from_le_bytes_mod_order(raw + 64, 64),
        tau_d = F::from_le_bytes_mod_order(raw + 128, 64), tau_s = F::from_le_bytes_mod_order(raw + 192, 64);
    FrH gamma = d.tr->challenge(128);

    auto adj = [&](const Tab& pull, const Tab& tab, const FrH& tau) {  // pull + psi * tab - tau, padded with tau_s (:700-710)
        std::vector<Dev::Term> t{{pull, F::ONE, 0, 0, matrix_size}, {tab, psi, 0, 0, matrix_size}, {nullptr, F::neg(tau), 0, 0, matrix_size}};
        if (full > matrix_size) t.push_back({nullptr, tau_s, 0, matrix_size, full - matrix_size});
        return d.lincomb(t, full);
    };
    Tab c_adj = adj(st.c_pull, st.c, tau_c), d_adj = adj(st.d_pull, st.d, tau_d);
    Tab c_pull_p = d.lincomb({{st.c_pull, F::ONE, 0, 0, matrix_size}}, full);
    Tab d_pull_p = d.lincomb({{st.d_pull, F::ONE, 0, 0, matrix_size}}, full);
    std::vector<Tab> halves = d.map_dense(ADD_INV, {c_adj, d_adj}, 1, 0, 2);  // map_split_hi, :719
    Tab iota;
    {
        std::vector<uint32_t> io(x_size);
        for (uint64_t i = 0; i < x_size; i++) io[i] = (uint32_t)i;
        gkr_u32buf* b = nullptr;
        ck(gkr_u32_upload(ctx, io.data(), io.size(), &b));
        U32 bh = std::make_shared<U32H>(b);
        gkr_table* t = nullptr;
        ck(gkr_table_from_u32(ctx, b, 0, &t));
        iota = std::make_shared<TabH>(t);
    }
    const uint64_t dsz = (uint64_t)1 << dl;
    Tab table_c = d.lincomb({{st.eq_c, F::ONE, 0, 0, x_size}, {iota, psi, 0, 0, x_size}, {nullptr, F::neg(tau_c), 0, 0, x_size}}, x_size);
    Tab table_d = d.lincomb({{st.eq_d, F::ONE, 0, 0, dsz}, {iota, psi, 0, 0, dsz}, {nullptr, F::neg(tau_d), 0, 0, dsz}}, dsz);
    FrH suppression_total = F::ZERO;
    if (!F::is_zero(tau_s)) suppression_total = F::mul(F::dbl(F::from_u64(full - matrix_size)), F::inverse(tau_s));
    const uint32_t m = xl + yl - 1;
    std::vector<Claims> mp = logup_mainphase_prove(d, {m, m, xl, dl}, suppression_total,
                                                   {Frac{halves[0], halves[1]}, Frac{halves[2], halves[3]}, Frac{st.ac_c, table_c}, Frac{st.ac_d, table_d}});
    if (mp.size() != 3) fail(ctx, "pushforward: logup returned the wrong number of claims");
    SplitAt s(true, 0, 2);
    Claims cd_claims = s.prove(d, mp[0], Advice());
    std::vector<FrH> gammas = make_gamma_pows(gamma, 5);
    // p_folded = p_0 + gamma (p_1 - 1) + gamma^2 ; p_selector_prod[y, x] = eq_trunc(r_y)[y] * p_folded[x]  (:740-758)
    Tab p_folded = d.lincomb({{st.p_0, F::ONE, 0, 0, x_size}, {st.p_1, gammas[1], 0, 0, x_size}, {nullptr, F::sub(gammas[2], gammas[1]), 0, 0, x_size}}, x_size);
    std::vector<FrH> eq_sel_y = eq_trunc_evals(yl, y_size, r_y);
    std::vector<Dev::Term> sel;
    for (uint32_t y = 0; y < y_size; y++) sel.push_back({p_folded, eq_sel_y[y], 0, (uint64_t)y << xl, x_size});
    Tab p_selector_prod = d.lincomb(sel, full);
    FrH ev_folded = F::add(F::add(evs[0], F::mul(gammas[1], evs[1])), F::mul(gammas[2], evs[2]));
    gkr_so* prod3_raw = nullptr;
    {
        gkr_table* in[3] = {p_selector_prod->h, c_pull_p->h, d_pull_p->h};
        uint64_t cl[4];
        frh_to_limbs(ev_folded, cl);
        ck(gkr_so_create_dense(ctx, GKR_SO_PLAIN, GKR_GATE_PROD3, 0, nullptr, 0, in, 3, matrix_logsize, cl, &prod3_raw));
    }
    SoH prod3(prod3_raw);
    if (cd_claims.evs.size() != 2) fail(ctx, "pushforward: malformed cd claims");
    FrH claim = F::add(F::add(cd_claims.evs[0], F::mul(gammas[1], cd_claims.evs[1])), F::mul(gammas[2], ev_folded));
    std::vector<Tab> keep;
    SoH frac(dense_eq_so(d, ADD_INV, {c_adj, d_adj}, cd_claims.point, cd_claims.evs, gamma, &keep));
    std::vector<FrH> output_point;
    for (uint32_t k = 0; k < matrix_logsize; k++) {  // the combined loop, pushforward.rs:781-806
        FrH pe[GKR_MAX_DEG + 1], fe[GKR_MAX_DEG + 1];
        uint32_t np = 0, nf = 0;
        ck(prod3.h->unipoly(pe, &np));
        ck(frac.h->unipoly(fe, &nf));
        if (np != 4 || nf != 4) fail(ctx, "pushforward: round polynomials must have degree 3");
        std::vector<FrH> pr = F::interpolate_coeffs(pe, 4), fr = F::interpolate_coeffs(fe, 4), combined(4);
        for (int j = 0; j < 4; j++) combined[j] = F::add(fr[j], F::mul(gammas[2], pr[j]));
        FrH chk = F::add(F::add(F::dbl(combined[0]), combined[1]), F::add(combined[2], combined[3]));
        if (chk != claim) fail(ctx, "pushforward: combined round polynomial does not match the running claim");
        FrH msg[3] = {combined[0], combined[2], combined[3]};
        d.tr->write_scalars(msg, 3);
        FrH t = d.tr->challenge(128);
        claim = F::evaluate_univar(combined, t);
        output_point.push_back(t);
        ck(prod3.h->bind(t));
        ck(frac.h->bind(t));
    }
    std::reverse(output_point.begin(), output_point.end());
    FrH pfe[3], ffe[3];
    ck(prod3.h->final_evals(pfe));
    ck(frac.h->final_evals(ffe));
    std::vector<FrH> out_y(output_point.begin(), output_point.begin() + yl);
    FrH adj_p_folded_ev = F::mul(pfe[0], F::inverse(eq_trunc_evaluate(yl, y_size, r_y, out_y)));
    FrH p_folded_ev = F::add(adj_p_folded_ev, gamma);
    FrH sel_ev = eq_sum(out_y, y_size);  // SelectorPoly::evaluate, verifier_polys.rs:68-71
    FrH tmp = F::mul(tau_s, F::sub(F::ONE, sel_ev));
    FrH psi_inv = F::inverse(psi);
    FrH c_ev = F::mul(psi_inv, F::sub(F::add(F::sub(ffe[0], pfe[1]), F::mul(tau_c, sel_ev)), tmp));
    FrH d_ev = F::mul(psi_inv, F::sub(F::add(F::sub(ffe[1], pfe[2]), F::mul(tau_d, sel_ev)), tmp));
    FinalClaims fc;
    fc.gamma = gamma;
    fc.matrix.point = output_point;
    fc.matrix.evs = {p_folded_ev, pfe[1], pfe[2], c_ev, d_ev};
    d.tr->write_scalars(fc.matrix.evs.data(), 5);
    fc.ac_c = mp[1];
    fc.ac_d = mp[2];
    return fc;
}

// ---- multiopen reduction (multiopen_reduction.rs:43-93) ----------------------------------------------------------------------
Claims multiopen_prove(Dev& d, uint32_t nvars, const std::vector<std::pair<std::vector<FrH>, FrH>>& claims, const std::vector<Tab>& advice) {
    const uint32_t nargs = (uint32_t)claims.size();
    FrH gamma = d.tr->challenge(128);
    std::vector<FrH> evs;
    for (auto& c : claims) evs.push_back(c.second);
    FrH folded = gamma_rlc(gamma, evs);
    std::vector<Tab> tables = advice;
    for (auto& c : claims) tables.push_back(d.eq_table(c.first));
    std::vector<FrH> gp = make_gamma_pows(gamma, nargs);
    std::vector<gkr_table*> in;
    for (auto& t : tables) in.push_back(t->h);
    uint64_t cl[4];
    frh_to_limbs(folded, cl);
    gkr_so* so = nullptr;
    ck(gkr_so_create_dense(d.ctx, GKR_SO_PLAIN, GKR_GATE_FOLDED_PROD, nargs, limbs_of(gp).data(), (uint32_t)gp.size(), in.data(), (uint32_t)in.size(),
                           nvars, cl, &so));
    SoH guard(so);
    Claims out;
    std::vector<FrH> fe;
    d.sumcheck_prove(so, nvars, &out.point, &fe);
    out.evs.assign(fe.begin(), fe.begin() + nargs);
    d.tr->write_scalars(out.evs.data(), out.evs.size());
    return out;
}

// ---- Knuckles opening (opening.rs:39-98) ----------------------------------------------------------------------------------------
struct Kzg {
    Dev& d;
    const Keys& key;
    G1P commit(const Tab& t) { return d.msm(key.srs, t, gkr_table_len(t->h)); }
    G1P open(const Tab& t, const FrH& pt, FrH* rem) {  // kzg.rs:129-132
        uint64_t p[4], r[4];
        frh_to_limbs(pt, p);
        gkr_table* q = nullptr;
        ck(gkr_poly_div_by_linear(d.ctx, t->h, p, &q, r));
        Tab qh = std::make_shared<TabH>(q);
        if (rem) *rem = frh_from_limbs(r);
        return commit(qh);
    }
    std::pair<G1P, G1P> verify_reduce_to_pair(const G1P& poly_comm, const G1P& quot_comm, const FrH& opening_at, const FrH& opening) {  // kzg.rs:49-60
        return {d.g1_lincomb({opening_at, F::neg(opening), F::ONE}, {quot_comm, key.g0, poly_comm}), quot_comm};
    }
};
std::pair<G1P, G1P> knuckles_opening_prove(Dev& d, const Keys& key, const G1P& comm, const std::vector<FrH>& point, const FrH& ev_claim, const Tab& advice) {
    gkr_ctx* ctx = d.ctx;
    Kzg kzg{d, key};
    gkr_table* t_raw = nullptr;
    uint64_t op[4];
    ck(gkr_knuckles_compute_t(ctx, key.knuckles, advice->h, limbs_of(point).data(), (uint32_t)point.size(), &t_raw, op));
    Tab t = std::make_shared<TabH>(t_raw);
    if (frh_from_limbs(op) != ev_claim) fail(ctx, "knuckles: opening != claimed evaluation (opening.rs:49)");
    G1P t_comm = kzg.commit(t);
    d.write_points({t_comm});
    FrH x = d.tr->challenge(128);
    FrH kx = F::mul(x, key.k);
    uint64_t xl[4], o1[4], o2[4];
    frh_to_limbs(x, xl);
    ck(gkr_poly_eval(ctx, t->h, xl, o1));
    ck(gkr_poly_eval(ctx, advice->h, xl, o2));
    FrH t_x = frh_from_limbs(o1), p_x = frh_from_limbs(o2);
    FrH two[2] = {t_x, p_x};
    d.tr->write_scalars(two, 2);
    FrH lam = d.tr->challenge(128);
    const uint64_t tl = gkr_table_len(t->h), al = gkr_table_len(advice->h);
    Tab p_lt = d.lincomb({{t, lam, 0, 0, tl}, {advice, F::ONE, 0, 0, al}}, tl);  // opening.rs:65-75
    // the two openings (kzg.rs:129-132) are separated by no challenge: both quotients first, ONE two-problem commitment, then
    // the transcript writes in the reference's order (opening.rs:77-87)
    FrH t_kx;
    G1P p_lt_x_proof, t_kx_proof;
    {
        uint64_t p[4], r1[4], r2[4];
        gkr_table *q1 = nullptr, *q2 = nullptr;
        frh_to_limbs(x, p);
        ck(gkr_poly_div_by_linear(ctx, p_lt->h, p, &q1, r1));
        Tab q1h = std::make_shared<TabH>(q1);
        frh_to_limbs(kx, p);
        ck(gkr_poly_div_by_linear(ctx, t->h, p, &q2, r2));
        Tab q2h = std::make_shared<TabH>(q2);
        t_kx = frh_from_limbs(r2);
        const gkr_table* tabs[2] = {q1, q2};
        const uint64_t lens2[2] = {gkr_table_len(q1), gkr_table_len(q2)};
        uint64_t out2[24];
        ck(gkr_msm_g1_multi(ctx, key.srs, 0, tabs, lens2, 2, out2));
        std::memcpy(p_lt_x_proof.data(), out2, 96);
        std::memcpy(t_kx_proof.data(), out2 + 12, 96);
    }
    d.write_points({p_lt_x_proof});
    d.tr->write_scalars(&t_kx, 1);
    d.write_points({t_kx_proof});
    FrH fin = d.tr->challenge(128);
    G1P p_lt_comm = d.g1_lincomb({lam, F::ONE}, {t_comm, comm});
    FrH p_lt_open = F::add(F::mul(t_x, lam), p_x);
    auto ab0 = kzg.verify_reduce_to_pair(p_lt_comm, p_lt_x_proof, x, p_lt_open);
    auto ab1 = kzg.verify_reduce_to_pair(t_comm, t_kx_proof, kx, t_kx);
    return {d.g1_lincomb({F::ONE, fin}, {ab0.first, ab1.first}), d.g1_lincomb({F::ONE, fin}, {ab0.second, ab1.second})};
}

FrH evaluate_poly(const std::vector<FrH>& poly, const std::vector<FrH>& pt) {  // cleanup/utils/arith.rs:6-9
    std::vector<FrH> e = eq_poly_sequence_last(pt);
    FrH acc = F::ZERO;
    for (size_t i = 0; i < poly.size(); i++) acc = F::add(acc, F::mul(poly[i], e[i]));
    return acc;
}

}  // namespace

// benchutils::run_pippenger (pippenger.rs:499-559): witness generation + phase-1 commitments + the whole proof.
extern "C" int gkr_run_pippenger(gkr_ctx* ctx, gkr_transcript* transcript, const gkr_srs* srs, const uint64_t* g0_xy, const gkr_knuckles* knuckles,
                                 const uint64_t* points_x, const uint64_t* points_y, const uint64_t* coefs, uint32_t d_logsize, uint32_t x_logsize,
                                 uint32_t num_bits, uint32_t clm, const uint64_t* r, uint64_t* dense_output, uint64_t* claim_evs, uint64_t* pair_xy) {
    if (!ctx) return GKR_ERR_ARG;
    if (!transcript || !srs || !g0_xy || !knuckles || !points_x || !points_y || !coefs || !r || d_logsize == 0 || num_bits == 0)
        return ctx->fail(GKR_ERR_ARG, "null argument");
    try {
        ck(cudaSetDevice(ctx->device) == cudaSuccess ? 0 : GKR_ERR_CUDA);
        Dev d{ctx, &transcript->t};
        const uint32_t y_size = (num_bits + d_logsize - 1) / d_logsize;
        uint32_t yl = 0;
        while ((1u << yl) < y_size) yl++;  // ark_std::log2 = ceil(log2)
        const uint32_t dl = d_logsize, xl = x_logsize;
        if (xl < dl || yl < clm) return ctx->fail(GKR_ERR_ARG, "x_logsize >= d_logsize and y_logsize >= clm required (pippenger.rs:98-99)");
        if (dl < 2) return ctx->fail(GKR_ERR_ARG, "d_logsize >= 2 required by the triangle circuit");
        Keys key;
        key.srs = const_cast<gkr_srs*>(srs);
        std::memcpy(key.g0.data(), g0_xy, 96);
        key.knuckles = knuckles;
        key.num_vars = gkr_knuckles_num_vars(knuckles);
        uint64_t kk[4];
        gkr_knuckles_k(knuckles, kk);
        key.k = frh_from_limbs(kk);
        if (gkr_srs_len(srs) < 2 * ((uint64_t)1 << key.num_vars) - 1) return ctx->fail(GKR_ERR_ARG, "SRS is too short.");  // knuckles.rs:67
        // PippengerWG::new (pippenger.rs:30-70)
        std::unique_ptr<Span> sp(new Span(ctx, "PushForwardState::new"));
        PushForwardState st(d, key, points_x, points_y, coefs, y_size, yl, dl, xl, clm);
        sp.reset(new Span(ctx, "witness: bintree + triangle"));
        std::vector<Vv> glue;  // GlueSplit::witness: split (x, y) as a bundle of 2 and the domain polynomial alone
        d.map_vecvec(ID(2), {st.image[0], st.image[1]}, 1, 2, &glue, nullptr);
        d.map_vecvec(ID(1), {st.image[2]}, 1, 1, &glue, nullptr);
        // memory plan: with more than ~35 % of the device memory in bintree witness (about 640 bytes per point-digit incidence when
        // every layer keeps its L1 / L2 images, as the reference does) only the layer inputs stay resident (GKR_WITNESS_RECOMPUTE=0/1 forces)
        bool recompute = false;
        {
            const double witness_bytes = 640.0 * (double)y_size * (double)((uint64_t)1 << xl);
            const char* v = getenv("GKR_WITNESS_RECOMPUTE");
            if (v) {
                recompute = v[0] == '1';
            } else if (witness_bytes > 16e9) {  // cudaMemGetInfo costs milliseconds: only asked when the answer can be yes
                size_t free_b = 0, total_b = 0;
                cudaMemGetInfo(&free_b, &total_b);
                recompute = witness_bytes > 0.35 * (double)total_b;
            }
        }
        PippengerEndingWG ending(d, yl, dl, xl, glue, recompute);
        sp.reset(new Span(ctx, "output claims"));
        // claims on the output of the triangle (pippenger.rs:528-539)
        std::vector<Tab> dense_out = d.map_dense(repeated(PRJ_L3, (dl - 2) + 3), ending.last());
        std::vector<FrH> rr(yl);
        for (uint32_t i = 0; i < yl; i++) rr[i] = frh_from_limbs(r + 4 * i);
        Claims claims;
        claims.point = rr;
        const uint64_t out_len = (uint64_t)1 << yl;
        for (size_t k = 0; k < dense_out.size(); k++) {
            std::vector<FrH> o = d.download(dense_out[k]);
            if (o.size() != out_len) fail(ctx, "unexpected output table length");
            claims.evs.push_back(evaluate_poly(o, rr));
            if (dense_output)
                for (uint64_t i = 0; i < out_len; i++) frh_to_limbs(o[i], dense_output + 4 * (k * out_len + i));
            if (claim_evs) frh_to_limbs(claims.evs.back(), claim_evs + 4 * k);
        }
        // Pippenger::prove (pippenger.rs:122-294)
        sp.reset(new Span(ctx, "prove: ending GKR"));
        d.write_points(st.c_comm);
        d.write_points(st.d_comm);
        d.write_points({st.p_0_comm});
        d.write_points({st.p_1_comm});
        d.write_points({st.ac_c_comm});
        d.write_points({st.ac_d_comm});
        {  // PippengerBucketed::prove (pippenger_ending.rs:102-157)
            Layers tri = triangle_protocol(yl + dl - 2, yl);
            claims = simple_gkr_prove(d, tri, claims, ending.triangle_advices);
            SplitAt s(true, yl, 3);
            claims = s.prove(d, claims, Advice());
            claims = s.prove(d, claims, Advice());
            Layers bt = bintree_protocol(yl + dl + xl, xl, xl, true);
            claims = simple_gkr_prove(d, bt, claims, ending.bintree_advices);
        }
        claims = GlueSplit().prove(d, claims, Advice());
        sp.reset(new Span(ctx, "prove: second phase"));
        st.second_phase(d, claims.point);
        d.write_points(st.c_pull_comm);
        d.write_points(st.d_pull_comm);
        sp.reset(new Span(ctx, "prove: pushforward"));
        FinalClaims fc = pushforward_prove(d, xl, yl, y_size, dl, claims, st);
        sp.reset(new Span(ctx, "prove: opening inputs + multiopen"));
        const FrH gamma = fc.gamma;
        // opening claims (pippenger.rs:166-205)
        const std::vector<FrH>& mpt = fc.matrix.point;
        const FrH p_folded_ev = fc.matrix.evs[0], c_pull_ev = fc.matrix.evs[1], d_pull_ev = fc.matrix.evs[2], c_ev = fc.matrix.evs[3], d_ev = fc.matrix.evs[4];
        std::vector<FrH> p_folded_point(clm, F::ZERO), ac_c_point(clm, F::ZERO), ac_d_point(xl + clm - dl, F::ZERO);
        p_folded_point.insert(p_folded_point.end(), mpt.begin() + yl, mpt.end());
        ac_c_point.insert(ac_c_point.end(), fc.ac_c.point.begin(), fc.ac_c.point.end());
        ac_d_point.insert(ac_d_point.end(), fc.ac_d.point.begin(), fc.ac_d.point.end());
        std::vector<FrH> combined_point(mpt.begin() + (yl - clm), mpt.end());
        std::vector<FrH> multirow_evs = eq_poly_sequence_last(std::vector<FrH>(mpt.begin(), mpt.begin() + (yl - clm)));
        uint8_t raw[64];
        d.tr->raw_challenge(raw, 64);  // challenge(512)
        FrH u = F::from_le_bytes_mod_order(raw, 64);
        std::vector<FrH> us = make_gamma_pows(u, 4);
        FrH combined_ev = F::add(F::add(c_ev, F::mul(d_ev, us[1])), F::add(F::mul(c_pull_ev, us[2]), F::mul(d_pull_ev, us[3])));
        std::vector<FrH> comm_coefs;
        std::vector<G1P> comm_pts;
        const std::vector<G1P>* groups[4] = {&st.c_comm, &st.d_comm, &st.c_pull_comm, &st.d_pull_comm};
        for (int j = 0; j < 4; j++)
            for (size_t k = 0; k < groups[j]->size(); k++) {
                comm_coefs.push_back(F::mul(multirow_evs[k], us[j]));
                comm_pts.push_back((*groups[j])[k]);
            }
        G1P combined_comm = d.g1_lincomb(comm_coefs, comm_pts);
        std::vector<std::pair<std::vector<FrH>, FrH>> oclaims{{p_folded_point, F::sub(p_folded_ev, F::mul(gamma, gamma))},
                                                             {ac_c_point, fc.ac_c.evs[0]},
                                                             {ac_d_point, fc.ac_d.evs[0]},
                                                             {combined_point, combined_ev}};
        // combined witness (pippenger.rs:209-223): row y of c, d, c_pull, d_pull lands in slot (y mod 2^clm)
        const uint64_t x_size = (uint64_t)1 << xl, cm = (uint64_t)1 << clm;
        const uint32_t nv = xl + clm;
        const uint64_t nvl = (uint64_t)1 << nv;
        std::vector<Dev::Term> terms;
        const Tab* tabs4[4] = {&st.c, &st.d, &st.c_pull, &st.d_pull};
        for (uint32_t y = 0; y < y_size; y++)
            for (int j = 0; j < 4; j++) terms.push_back({*tabs4[j], F::mul(multirow_evs[y / cm], us[j]), x_size * y, x_size * (y % cm), x_size});
        Tab combined_witness = d.lincomb(terms, nvl);
        std::vector<Tab> mw{d.lincomb({{st.p_0, F::ONE, 0, 0, x_size}, {st.p_1, gamma, 0, 0, x_size}}, nvl),
                            d.lincomb({{st.ac_c, F::ONE, 0, 0, gkr_table_len(st.ac_c->h)}}, nvl),
                            d.lincomb({{st.ac_d, F::ONE, 0, 0, gkr_table_len(st.ac_d->h)}}, nvl), combined_witness};
        Claims mo = multiopen_prove(d, nv, oclaims, mw);
        FrH q = d.tr->challenge(128);
        std::vector<FrH> qs = make_gamma_pows(q, 4);
        G1P folded_comm = d.g1_lincomb({qs[0], F::mul(qs[0], gamma), qs[1], qs[2], qs[3]},
                                       {st.p_0_comm, st.p_1_comm, st.ac_c_comm, st.ac_d_comm, combined_comm});
        Tab folded_witness = d.lincomb({{mw[0], qs[0], 0, 0, nvl}, {mw[1], qs[1], 0, 0, nvl}, {mw[2], qs[2], 0, 0, nvl}, {mw[3], qs[3], 0, 0, nvl}}, nvl);
        sp.reset(new Span(ctx, "prove: knuckles opening"));
        auto pair = knuckles_opening_prove(d, key, folded_comm, mo.point, gamma_rlc(q, mo.evs), folded_witness);
        sp.reset();
        if (pair_xy) {
            std::memcpy(pair_xy, pair.first.data(), 96);
            std::memcpy(pair_xy + 12, pair.second.data(), 96);
        }
        TraceAcc::dump(ctx);
        return GKR_OK;
    } catch (const Fail& f) {
        return f.code;
    } catch (const std::exception& e) {
        return ctx->fail(GKR_ERR_PROTOCOL, e.what());
    }
}
