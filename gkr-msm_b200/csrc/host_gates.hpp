// Host-side evaluation of the gate set on single points -- O(1) glue only: the closed-form padding terms of
// the Deg2 objects (`pad_results`, `col_pad_results`: dense_eq.rs:114-116, vecvec_eq.rs:309-315) and the
// row/col pad values of mapped VecVec polynomials (vecvec.rs:491-499).  Tables never pass through here.
#pragma once
#include <vector>
#include "../../include/gkr_msm_b200.h"
#include "host_field.hpp"

namespace gkr {

static const FrH TE_D_MONT = {{12167860994669987632ULL, 4043113551995129031ULL, 6052647550941614584ULL, 3904213385886034240ULL}};  // src/utils.rs:34-37

static inline bool base_gate_io(int gate, int* n_in, int* n_out) {
    switch (gate) {
        case GKR_GATE_AFF_L1: *n_in = 4; *n_out = 3; return true;
        case GKR_GATE_AFF_L2: *n_in = 3; *n_out = 3; return true;
        case GKR_GATE_AFF_L3: *n_in = 3; *n_out = 3; return true;
        case GKR_GATE_PRJ_L1: *n_in = 6; *n_out = 4; return true;
        case GKR_GATE_PRJ_L2: *n_in = 4; *n_out = 4; return true;
        case GKR_GATE_PRJ_L3: *n_in = 4; *n_out = 3; return true;
        case GKR_GATE_TRI_L1: *n_in = 12; *n_out = 12; return true;
        case GKR_GATE_BITCHECK: *n_in = 1; *n_out = 1; return true;
        case GKR_GATE_LOGUP_LAYER: *n_in = 4; *n_out = 2; return true;
        case GKR_GATE_ADD_INVERSES: *n_in = 2; *n_out = 2; return true;
        case GKR_GATE_AFF_L1_BITCHECK2: *n_in = 6; *n_out = 5; return true;
        default: return false;
    }
}

static inline FrH h_add5(const FrH& y, const FrH& x) {  // y - a x with a = -5
    FrH t = frh::dbl(frh::dbl(x));
    return frh::add(y, frh::add(t, x));
}

static inline void base_gate_eval(int gate, const FrH* a, FrH* o) {
    using namespace frh;
    switch (gate) {
        case GKR_GATE_AFF_L1:
            o[0] = mul(a[0], a[3]); o[1] = mul(a[2], a[1]); o[2] = h_add5(mul(a[1], a[3]), mul(a[0], a[2]));
            break;
        case GKR_GATE_AFF_L2:
            o[0] = add(a[0], a[1]); o[1] = a[2]; o[2] = mul(a[0], a[1]);
            break;
        case GKR_GATE_AFF_L3: {
            FrH dxy = mul(a[2], TE_D_MONT), m = sub(ONE, dxy), p = add(ONE, dxy);
            o[0] = mul(m, a[0]); o[1] = mul(p, a[1]); o[2] = mul(m, p);
            break;
        }
        case GKR_GATE_PRJ_L1:
            o[0] = mul(a[0], a[4]); o[1] = mul(a[3], a[1]); o[2] = h_add5(mul(a[1], a[4]), mul(a[0], a[3])); o[3] = mul(a[2], a[5]);
            break;
        case GKR_GATE_PRJ_L2:
            o[0] = mul(add(a[0], a[1]), a[3]); o[1] = mul(a[2], a[3]); o[2] = mul(a[3], a[3]); o[3] = mul(a[0], a[1]);
            break;
        case GKR_GATE_PRJ_L3: {
            FrH dxy = mul(a[3], TE_D_MONT), m = sub(a[2], dxy), p = add(a[2], dxy);
            o[0] = mul(m, a[0]); o[1] = mul(p, a[1]); o[2] = mul(m, p);
            break;
        }
        case GKR_GATE_TRI_L1: {
            FrH t[6];
            for (int i = 0; i < 3; i++) { t[i] = a[i]; t[3 + i] = a[6 + i]; }
            base_gate_eval(GKR_GATE_PRJ_L1, t, o);
            for (int i = 0; i < 3; i++) { t[i] = a[3 + i]; t[3 + i] = a[9 + i]; }
            base_gate_eval(GKR_GATE_PRJ_L1, t, o + 4);
            base_gate_eval(GKR_GATE_PRJ_L1, a + 6, o + 8);
            break;
        }
        case GKR_GATE_BITCHECK: o[0] = sub(mul(a[0], a[0]), a[0]); break;
        case GKR_GATE_LOGUP_LAYER: o[0] = add(mul(a[0], a[3]), mul(a[1], a[2])); o[1] = mul(a[1], a[3]); break;
        case GKR_GATE_ADD_INVERSES: o[0] = add(a[0], a[1]); o[1] = mul(a[0], a[1]); break;
        case GKR_GATE_AFF_L1_BITCHECK2:
            base_gate_eval(GKR_GATE_AFF_L1, a, o);
            o[3] = sub(mul(a[4], a[4]), a[4]); o[4] = sub(mul(a[5], a[5]), a[5]);
            break;
        default: break;
    }
}

// A composite gate: Stacked(Repeated(g_0, r_0), Repeated(g_1, r_1), ...)  (algfn.rs:187-259) -- covers every
// composition the reference builds (bintree_add.rs:259-273, triangle_add.rs:126-157, 199-231).
struct GateStack {
    std::vector<int> gate, repeat;
    int n_ins = 0, n_outs = 0;
    bool init(const int* g, const uint32_t* r, uint32_t n_parts) {
        gate.clear(); repeat.clear(); n_ins = n_outs = 0;
        for (uint32_t i = 0; i < n_parts; i++) {
            int ni, no;
            if (!base_gate_io(g[i], &ni, &no) || r[i] == 0) return false;
            gate.push_back(g[i]); repeat.push_back((int)r[i]);
            n_ins += ni * (int)r[i]; n_outs += no * (int)r[i];
        }
        return n_parts > 0;
    }
    void eval(const FrH* a, FrH* o) const {
        for (size_t i = 0; i < gate.size(); i++) {
            int ni = 0, no = 0;
            base_gate_io(gate[i], &ni, &no);
            for (int k = 0; k < repeat[i]; k++) { base_gate_eval(gate[i], a, o); a += ni; o += no; }
        }
    }
};

}  // namespace gkr
