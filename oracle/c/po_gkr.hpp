/* TEST INFRASTRUCTURE ONLY -- CPU oracle, never linked into or called from the product path.
 *
 * The EC-addition GKR circuits of the reference restated in C++ (from the reference sources, not from csrc/protocol.cu):
 *   SimpleGKR::prove                                  src/cleanup/protocols/gkrs/gkr.rs:39-60
 *   SplitAt, GlueSplit                                src/cleanup/protocols/splits.rs:120-203
 *   ZeroCheck                                         src/cleanup/protocols/zero_check.rs:17-33
 *   bintree witness `build` / `make_step`, protocol   src/cleanup/protocols/gkrs/bintree_add.rs:124-375
 *   triangle witness / protocol                       src/cleanup/protocols/gkrs/triangle_add.rs:76-232
 *   PippengerEndingWG, PippengerBucketed              src/cleanup/protocols/pippenger_ending.rs:26-157
 */
#pragma once
#include "po_sumcheck.hpp"

namespace po {

struct Advice {  // SplitVecVecMapGKRAdvice, split_map_gkr.rs:65-71
    enum Kind { EMPTY, VV, DENSE } kind = EMPTY;
    std::vector<VecVec> vv;
    std::vector<Vec> dense;
    static Advice of(std::vector<VecVec> v) {
        Advice a;
        a.kind = VV;
        a.vv = std::move(v);
        return a;
    }
    static Advice of(std::vector<Vec> v) {
        Advice a;
        a.kind = DENSE;
        a.dense = std::move(v);
        return a;
    }
};

struct Layer {  // one GKRLayer of a SimpleGKR
    enum Kind { VV_SUMCHECK, DENSE_SUMCHECK, ZERO_CHECK, SPLIT_AT } kind;
    GateP gate;
    size_t num_vars = 0, num_vertical_vars = 0;
    SplitIdx idx{true, 0};
    size_t bundle = 1;
};

/* SplitAt::prove, splits.rs:127-143 */
static inline Claims split_at_prove(Transcript& tr, const Claims& c, SplitIdx idx, size_t bundle) {
    Fr r = tr.challenge(128);
    Claims out;
    out.point = c.point;
    const size_t n_chunks = (c.evs.size() + bundle - 1) / bundle;
    Vec l, rr;
    for (size_t ch = 0; ch < n_chunks; ch++)
        for (size_t k = ch * bundle; k < std::min(c.evs.size(), (ch + 1) * bundle); k++) ((ch % 2) ? rr : l).push_back(c.evs[k]);
    for (size_t i = 0; i < std::min(l.size(), rr.size()); i++) out.evs.push_back(l[i] + r * (rr[i] - l[i]));
    const size_t pos = idx.lo ? out.point.size() - idx.k : idx.k;
    out.point.insert(out.point.begin() + pos, r);
    return out;
}
/* GlueSplit, splits.rs:172-203 */
static inline std::vector<VecVec> glue_split_witness(const std::vector<VecVec>& polys) {
    std::vector<VecVec> first(polys.begin(), polys.begin() + 2), second(polys.begin() + 2, polys.begin() + 3);
    std::vector<VecVec> out = vecvec_map_split(first, IdGate(2), SplitIdx::LO(0), 2);
    std::vector<VecVec> o2 = vecvec_map_split(second, IdGate(1), SplitIdx::LO(0), 1);
    for (auto& v : o2) out.push_back(std::move(v));
    return out;
}
static inline Claims glue_split_prove(Transcript& tr, const Claims& c) {
    Fr r = tr.challenge(128);
    Claims out;
    out.point = c.point;
    out.evs = Vec{c.evs[0] + r * (c.evs[2] - c.evs[0]), c.evs[1] + r * (c.evs[3] - c.evs[1]), c.evs[4] + r * (c.evs[5] - c.evs[4])};
    out.point.push_back(r);
    return out;
}

/* SimpleGKR::prove, gkr.rs:45-50: layers in reverse, advices popped from the end */
static inline Claims simple_gkr_prove(const std::vector<Layer>& layers, Transcript& tr, Claims claims, std::vector<Advice>& advices) {
    PO_ASSERT(advices.size() == layers.size(), "SimpleGKR: advices / layers mismatch");
    for (size_t li = layers.size(); li-- > 0;) {
        const Layer& L = layers[li];
        Advice adv = std::move(advices.back());
        advices.pop_back();
        switch (L.kind) {
            case Layer::VV_SUMCHECK:
                PO_ASSERT(adv.kind == Advice::VV && (int)adv.vv.size() == L.gate->n_ins, "VecVecDeg2Sumcheck: advice shape");
                claims = vecvec_deg2_sumcheck_prove(tr, L.gate, L.num_vars, L.num_vertical_vars, claims, std::move(adv.vv));
                break;
            case Layer::DENSE_SUMCHECK:
                PO_ASSERT(adv.kind == Advice::DENSE && (int)adv.dense.size() == L.gate->n_ins, "DenseDeg2Sumcheck: advice shape");
                claims = dense_deg2_sumcheck_prove(tr, L.gate, L.num_vars, claims, std::move(adv.dense));
                break;
            case Layer::ZERO_CHECK:  // zero_check.rs:24-28: two zero claims appended
                claims.evs.push_back(Fr::zero());
                claims.evs.push_back(Fr::zero());
                break;
            case Layer::SPLIT_AT:
                claims = split_at_prove(tr, claims, L.idx, L.bundle);
                break;
        }
    }
    return claims;
}

/* ---- bintree (bintree_add.rs) ---------------------------------------------------------------------------------------- */
static inline GateP bt_gate(int step, bool affine) {
    if (affine) {
        if (step == 0) return std::make_shared<AffL1>();
        if (step == 1) return std::make_shared<AffL2>();
        return std::make_shared<AffL3>();
    }
    if (step == 0) return std::make_shared<PrjL1>();
    if (step == 1) return std::make_shared<PrjL2>();
    return std::make_shared<PrjL3>();
}
static inline Advice advice_map(const Advice& a, const Gate& f) {  // :173-183
    if (a.kind == Advice::VV) {
        std::vector<VecVec> in(a.vv.begin(), a.vv.begin() + f.n_ins);
        return Advice::of(vecvec_map(in, f));
    }
    PO_ASSERT(a.kind == Advice::DENSE, "advice_map on EMPTY");
    return Advice::of(dense_map(ptrs(a.dense, f.n_ins), f));
}
static inline Advice advice_map_split(const Advice& a, const Gate& f, size_t layer_idx, size_t row_logsize, SplitIdx idx, size_t bundle) {  // :185-202
    if (a.kind == Advice::VV) {
        std::vector<VecVec> in(a.vv.begin(), a.vv.begin() + f.n_ins);
        if (layer_idx + 2 == row_logsize) return Advice::of(vecvec_map_split_to_dense(in, f, idx, bundle));
        return Advice::of(vecvec_map_split(in, f, idx, bundle));
    }
    PO_ASSERT(a.kind == Advice::DENSE, "advice_map_split on EMPTY");
    std::vector<Vec> in(a.dense.begin(), a.dense.begin() + f.n_ins);
    return Advice::of(dense_map_split(in, f, idx, bundle));
}
static inline std::vector<Advice> bintree_witness(Advice advice, size_t row_logsize, size_t num_adds, bool do_bitcheck) {  // :137-171
    PO_ASSERT(num_adds > 0, "bintree: num_adds");
    std::vector<Advice> advices;
    for (size_t add_idx = 0; add_idx < num_adds; add_idx++) {
        for (int step = 0; step < 3; step++) {
            const bool last = add_idx + 1 == num_adds;
            Advice next;
            bool have_next = true;
            GateP g = bt_gate(step, add_idx == 0);
            if (step < 2) next = advice_map(advice, *g);
            else if (last) have_next = false;
            else next = advice_map_split(advice, *g, add_idx, row_logsize, SplitIdx::LO(0), 3);
            advices.push_back(std::move(advice));
            if (add_idx == 0 && step == 0 && do_bitcheck) advices.push_back(Advice());
            if (have_next) advice = std::move(next);
            else advice = Advice();
        }
        if (add_idx + 1 != num_adds) advices.push_back(Advice());
    }
    return advices;
}
static inline std::vector<Layer> bintree_protocol(size_t num_vars, size_t num_adds, size_t row_logsize, bool do_bitcheck) {  // :247-375
    std::vector<Layer> layers;
    const size_t nvv = num_vars - row_logsize;
    for (size_t i = 0; i < num_adds; i++) {
        for (int step = 0; step < 3; step++) {
            Layer L;
            L.num_vars = num_vars - i - 1;
            L.num_vertical_vars = nvv;
            if (i == 0) {
                L.kind = Layer::VV_SUMCHECK;
                L.gate = bt_gate(step, true);
                if (step == 0 && do_bitcheck) L.gate = std::make_shared<Stacked>(std::make_shared<AffL1>(), std::make_shared<Repeated>(std::make_shared<BitCheck>(), 2));
            } else {
                L.kind = (i + 1 < row_logsize) ? Layer::VV_SUMCHECK : Layer::DENSE_SUMCHECK;
                L.gate = bt_gate(step, false);
            }
            layers.push_back(L);
            if (i == 0 && step == 0 && do_bitcheck) {
                Layer Z;
                Z.kind = Layer::ZERO_CHECK;
                layers.push_back(Z);
            }
        }
        if (i != num_adds - 1) {
            Layer S;
            S.kind = Layer::SPLIT_AT;
            S.idx = SplitIdx::LO(0);
            S.bundle = 3;
            layers.push_back(S);
        }
    }
    return layers;
}

/* ---- triangle (triangle_add.rs) ---------------------------------------------------------------------------------------- */
static inline GateP tri_l1(size_t layer_idx) {
    return std::make_shared<Stacked>(std::make_shared<TriL1>(), std::make_shared<Repeated>(std::make_shared<PrjL1>(), (int)layer_idx));
}
static inline GateP tri_l2(size_t layer_idx) { return std::make_shared<Repeated>(std::make_shared<PrjL2>(), (int)layer_idx + 3); }
static inline GateP tri_l3(size_t layer_idx) { return std::make_shared<Repeated>(std::make_shared<PrjL3>(), (int)layer_idx + 3); }
static inline std::vector<Vec> triangle_last_step(const std::vector<Vec>& advice, size_t layer_idx) {  // :88-99
    GateP g = tri_l3(layer_idx);
    return dense_map(ptrs(advice, g->n_ins), *g);
}
static inline std::vector<Advice> triangle_witness(std::vector<Vec> advice, size_t num_vars, SplitIdx split_idx) {  // :101-158
    const size_t hi = split_idx.hi_usize(num_vars);
    const SplitIdx split_hi = SplitIdx::HI(hi);
    const size_t num_layers = num_vars - hi;
    std::vector<Advice> advices;
    for (size_t layer_idx = 0; layer_idx <= num_layers; layer_idx++) {
        for (int step = 0; step < 3; step++) {
            std::vector<Vec> next;
            if (step == 0) {
                GateP g = tri_l1(layer_idx);
                next = dense_map(ptrs(advice, g->n_ins), *g);
            } else if (step == 1) {
                GateP g = tri_l2(layer_idx);
                next = dense_map(ptrs(advice, g->n_ins), *g);
            } else if (layer_idx != num_layers) {
                GateP g = tri_l3(layer_idx);
                next = dense_map_split(advice, *g, split_hi, 3);
            }
            advices.push_back(Advice::of(std::move(advice)));
            advice = std::move(next);
        }
        if (layer_idx < num_layers) advices.push_back(Advice());
    }
    return advices;
}
static inline std::vector<Layer> triangle_protocol(size_t num_vars, SplitIdx split_idx) {  // :173-232
    const size_t hi = split_idx.hi_usize(num_vars), num_layers = num_vars - hi;
    std::vector<Layer> layers;
    for (size_t layer_idx = 0; layer_idx <= num_layers; layer_idx++) {
        for (int step = 0; step < 3; step++) {
            Layer L;
            L.kind = Layer::DENSE_SUMCHECK;
            L.num_vars = num_vars - layer_idx;
            L.gate = step == 0 ? tri_l1(layer_idx) : (step == 1 ? tri_l2(layer_idx) : tri_l3(layer_idx));
            layers.push_back(L);
        }
        if (layer_idx < num_layers) {
            Layer S;
            S.kind = Layer::SPLIT_AT;
            S.idx = SplitIdx::HI(hi);
            S.bundle = 3;
            layers.push_back(S);
        }
    }
    return layers;
}

/* ---- PippengerEndingWG / PippengerBucketed (pippenger_ending.rs:26-157) -------------------------------------------------- */
struct PippengerEndingWG {
    std::vector<Advice> bintree_advices, triangle_advices;
    /* The reference builds the bintree witness twice (:40-45 and :67-72) and keeps one copy; once is enough for identical
     * outputs (and is the cheaper CPU baseline). */
    PippengerEndingWG(size_t multirow_vars, size_t bucket_vars, size_t horizontal_vars, std::vector<VecVec> inputs) {
        PO_ASSERT(inputs.size() == 6, "PippengerEndingWG: 6 inputs");
        bintree_advices = bintree_witness(Advice::of(std::move(inputs)), horizontal_vars, horizontal_vars, true);
        GateP l3 = bt_gate(2, horizontal_vars - 1 == 0);
        Advice last = advice_map(bintree_advices.back(), *l3);  // bintree last_step, :124-135
        PO_ASSERT(last.kind == Advice::DENSE, "bintree output must be dense");
        std::vector<Vec> split_l1 = dense_map_split(last.dense, IdGate(3), SplitIdx::HI(multirow_vars), 3);
        std::vector<Vec> split_l2 = dense_map_split(split_l1, Repeated(std::make_shared<IdGate>(3), 2), SplitIdx::HI(multirow_vars), 3);
        triangle_advices = triangle_witness(std::move(split_l2), multirow_vars + bucket_vars - 2, SplitIdx::HI(multirow_vars));
    }
    const std::vector<Vec>& last() const { return triangle_advices.back().dense; }
};
static inline Claims pippenger_bucketed_prove(Transcript& tr, Claims claims, PippengerEndingWG& wg, size_t multirow_vars, size_t bucket_vars, size_t horizontal_vars) {
    std::vector<Layer> triangle = triangle_protocol(multirow_vars + bucket_vars - 2, SplitIdx::HI(multirow_vars));
    std::vector<Layer> bintree = bintree_protocol(multirow_vars + bucket_vars + horizontal_vars, horizontal_vars, horizontal_vars, true);
    claims = simple_gkr_prove(triangle, tr, claims, wg.triangle_advices);
    claims = split_at_prove(tr, claims, SplitIdx::HI(multirow_vars), 3);
    claims = split_at_prove(tr, claims, SplitIdx::HI(multirow_vars), 3);
    return simple_gkr_prove(bintree, tr, claims, wg.bintree_advices);
}
}  // namespace po
