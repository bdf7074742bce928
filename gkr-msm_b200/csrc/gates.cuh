// The closed set of polynomial gates the reference instantiates on the hot path, as device functors.
// `AlgFn` is an open Rust generic (src/cleanup/utils/algfn.rs:20-34); a C ABI needs a closed enum, so the
// ids below are the ones exported in include/gkr_msm_b200.h (enum gkr_gate_id).
//
//   twisted-Edwards addition gates      src/cleanup/utils/twisted_edwards_ops.rs:10-81
//   BitCheckFn, Stacked, Repeated       src/cleanup/utils/algfn.rs:187-291
//   LogupLayerFn                        src/cleanup/protocols/pushforward/logup_mainphase.rs:42-61
//   AddInversesFn, Prod3Fn              src/cleanup/protocols/pushforward/pushforward.rs:38-50, 266-281
//   FoldedProdAlgFn                     src/cleanup/protocols/multiopen_reduction.rs:13-41
//   GammaWrapper, EqWrapper             src/cleanup/protocols/sumcheck.rs:706-741, 802-829
#pragma once
#include "field.cuh"

enum GateId : int {
    GATE_AFF_L1 = 0,
    GATE_AFF_L2 = 1,
    GATE_AFF_L3 = 2,
    GATE_PRJ_L1 = 3,
    GATE_PRJ_L2 = 4,
    GATE_PRJ_L3 = 5,
    GATE_TRI_L1 = 6,
    GATE_BITCHECK = 7,
    GATE_LOGUP_LAYER = 8,
    GATE_ADD_INVERSES = 9,
    GATE_PROD3 = 10,
    GATE_FOLDED_PROD = 11,
    GATE_ID = 12,
    GATE_AFF_L1_BITCHECK2 = 13,
};

#define GKR_MAX_GATE_CONSTS 16
// gamma powers: g[i] multiplies output i (g[0] is never read: output 0 has coefficient one).
struct GateConsts {
    Fr g[GKR_MAX_GATE_CONSTS];
};

template <int G>
struct MoGate;

template <>
struct MoGate<GATE_AFF_L1> {
    static constexpr int HOMOG = 2;  // degree when every output is homogenised (0: not homogeneous)
    __host__ __device__ static constexpr int out_deg(int i) { return 2; }
    static constexpr int N_INS = 4, N_OUTS = 3;
    __device__ __forceinline__ static void eval(const Fr* a, Fr* o) {
        o[0] = fr_mul(a[0], a[3]);
        o[1] = fr_mul(a[2], a[1]);
        o[2] = fr_add_5x(fr_mul(a[1], a[3]), fr_mul(a[0], a[2]));
    }
};

template <>
struct MoGate<GATE_AFF_L2> {
    static constexpr int HOMOG = 2;  // degree when every output is homogenised (0: not homogeneous)
    __host__ __device__ static constexpr int out_deg(int i) { return i == 2 ? 2 : 1; }
    static constexpr int N_INS = 3, N_OUTS = 3;
    __device__ __forceinline__ static void eval(const Fr* a, Fr* o) {
        o[0] = fr_add(a[0], a[1]);
        o[1] = a[2];
        o[2] = fr_mul(a[0], a[1]);
    }
};

template <>
struct MoGate<GATE_AFF_L3> {
    static constexpr int HOMOG = 0;  // degree when every output is homogenised (0: not homogeneous)
    __host__ __device__ static constexpr int out_deg(int i) { return 0; }
    static constexpr int N_INS = 3, N_OUTS = 3;
    __device__ __forceinline__ static void eval(const Fr* a, Fr* o) {
        Fr dxy = fr_mul(a[2], fr_te_d());
        Fr m = fr_sub(fr_one(), dxy);
        Fr p = fr_add(fr_one(), dxy);
        o[0] = fr_mul(m, a[0]);
        o[1] = fr_mul(p, a[1]);
        o[2] = fr_mul(m, p);
    }
};

template <>
struct MoGate<GATE_PRJ_L1> {
    static constexpr int HOMOG = 2;  // degree when every output is homogenised (0: not homogeneous)
    __host__ __device__ static constexpr int out_deg(int i) { return 2; }
    static constexpr int N_INS = 6, N_OUTS = 4;
    __device__ __forceinline__ static void eval(const Fr* a, Fr* o) {
        o[0] = fr_mul(a[0], a[4]);
        o[1] = fr_mul(a[3], a[1]);
        o[2] = fr_add_5x(fr_mul(a[1], a[4]), fr_mul(a[0], a[3]));
        o[3] = fr_mul(a[2], a[5]);
    }
};

template <>
struct MoGate<GATE_PRJ_L2> {
    static constexpr int HOMOG = 2;  // degree when every output is homogenised (0: not homogeneous)
    __host__ __device__ static constexpr int out_deg(int i) { return 2; }
    static constexpr int N_INS = 4, N_OUTS = 4;
    __device__ __forceinline__ static void eval(const Fr* a, Fr* o) {
        o[0] = fr_mul(fr_add(a[0], a[1]), a[3]);
        o[1] = fr_mul(a[2], a[3]);
        o[2] = fr_sqr(a[3]);
        o[3] = fr_mul(a[0], a[1]);
    }
};

template <>
struct MoGate<GATE_PRJ_L3> {
    static constexpr int HOMOG = 2;  // degree when every output is homogenised (0: not homogeneous)
    __host__ __device__ static constexpr int out_deg(int i) { return 2; }
    static constexpr int N_INS = 4, N_OUTS = 3;
    __device__ __forceinline__ static void eval(const Fr* a, Fr* o) {
        Fr dxy = fr_mul(a[3], fr_te_d());
        Fr m = fr_sub(a[2], dxy);
        Fr p = fr_add(a[2], dxy);
        o[0] = fr_mul(m, a[0]);
        o[1] = fr_mul(p, a[1]);
        o[2] = fr_mul(m, p);
    }
};

// three projective L1 on (a,c), (b,d), (c,d); inputs a,b,c,d = 4 points x (x,y,z)
template <>
struct MoGate<GATE_TRI_L1> {
    static constexpr int HOMOG = 2;  // degree when every output is homogenised (0: not homogeneous)
    __host__ __device__ static constexpr int out_deg(int i) { return 2; }
    static constexpr int N_INS = 12, N_OUTS = 12;
    __device__ __forceinline__ static void eval(const Fr* p, Fr* o) {
        Fr t[6];
#pragma unroll
        for (int i = 0; i < 3; i++) { t[i] = p[i]; t[3 + i] = p[6 + i]; }
        MoGate<GATE_PRJ_L1>::eval(t, o);
#pragma unroll
        for (int i = 0; i < 3; i++) { t[i] = p[3 + i]; t[3 + i] = p[9 + i]; }
        MoGate<GATE_PRJ_L1>::eval(t, o + 4);
        MoGate<GATE_PRJ_L1>::eval(p + 6, o + 8);
    }
};

template <>
struct MoGate<GATE_BITCHECK> {
    static constexpr int HOMOG = 0;  // degree when every output is homogenised (0: not homogeneous)
    __host__ __device__ static constexpr int out_deg(int i) { return 0; }
    static constexpr int N_INS = 1, N_OUTS = 1;
    __device__ __forceinline__ static void eval(const Fr* a, Fr* o) { o[0] = fr_sub(fr_sqr(a[0]), a[0]); }
};

// Stacked(affine_l1, Repeated(BitCheck, 2))   (src/cleanup/protocols/gkrs/bintree_add.rs:259-273)
template <>
struct MoGate<GATE_AFF_L1_BITCHECK2> {
    static constexpr int HOMOG = 0;  // degree when every output is homogenised (0: not homogeneous)
    __host__ __device__ static constexpr int out_deg(int i) { return 0; }
    static constexpr int N_INS = 6, N_OUTS = 5;
    __device__ __forceinline__ static void eval(const Fr* a, Fr* o) {
        MoGate<GATE_AFF_L1>::eval(a, o);
        o[3] = fr_sub(fr_sqr(a[4]), a[4]);
        o[4] = fr_sub(fr_sqr(a[5]), a[5]);
    }
};

template <>
struct MoGate<GATE_LOGUP_LAYER> {
    static constexpr int HOMOG = 2;  // degree when every output is homogenised (0: not homogeneous)
    __host__ __device__ static constexpr int out_deg(int i) { return 2; }
    static constexpr int N_INS = 4, N_OUTS = 2;
    __device__ __forceinline__ static void eval(const Fr* a, Fr* o) {
        o[0] = fr_add(fr_mul(a[0], a[3]), fr_mul(a[1], a[2]));
        o[1] = fr_mul(a[1], a[3]);
    }
};

template <>
struct MoGate<GATE_ADD_INVERSES> {
    static constexpr int HOMOG = 2;  // degree when every output is homogenised (0: not homogeneous)
    __host__ __device__ static constexpr int out_deg(int i) { return i == 1 ? 2 : 1; }
    static constexpr int N_INS = 2, N_OUTS = 2;
    __device__ __forceinline__ static void eval(const Fr* a, Fr* o) {
        o[0] = fr_add(a[0], a[1]);
        o[1] = fr_mul(a[0], a[1]);
    }
};

// sum_i gamma^i * out_i  (GammaWrapper::exec / the gamma_pows fold of the Deg2 objects)
template <int G>
__device__ __forceinline__ Fr gamma_eval(const Fr* a, const GateConsts& c) {
    Fr o[MoGate<G>::N_OUTS];
    MoGate<G>::eval(a, o);
    Fr ret = o[0];
#pragma unroll
    for (int i = 1; i < MoGate<G>::N_OUTS; i++) ret = fr_add(ret, fr_mul(o[i], c.g[i]));
    return ret;
}

// Same, for tables that carry a common scale factor sigma (DenseSO's 128-bit folds leave 2^-128 per fold, see
// dense_sumcheck.cu): an output of degree below the gate's degree gets its gamma power pre-multiplied by sigma on
// the host, INCLUDING output 0, so that the sum is homogeneous (= sigma^HOMOG * the true value).
template <int G>
__device__ __forceinline__ Fr gamma_eval_scaled(const Fr* a, const GateConsts& c) {
    Fr o[MoGate<G>::N_OUTS];
    MoGate<G>::eval(a, o);
    Fr ret = (MoGate<G>::out_deg(0) < MoGate<G>::HOMOG) ? fr_mul(o[0], c.g[0]) : o[0];
#pragma unroll
    for (int i = 1; i < MoGate<G>::N_OUTS; i++) ret = fr_add(ret, fr_mul(o[i], c.g[i]));
    return ret;
}

// ---- single-output gates for DenseSumcheckObjectSO --------------------------------------------
// eval(): the gate value.  mac(): adds the gate value to an unreduced accumulator (the last multiplication is not
// reduced, FrWide).  HDEG: total degree under a common scaling of all tables (0 = not homogeneous, no fast folds).
struct SoProd3 {
    static constexpr int P = 3, DEG = 3, HDEG = 3, N_OUTS = 1;
    __host__ __device__ static constexpr int gamma_shift(int) { return 0; }
    __device__ __forceinline__ static Fr eval(const Fr* a, const GateConsts&) { return fr_mul(fr_mul(a[0], a[1]), a[2]); }
    __device__ __forceinline__ static void mac(FrWide& w, const Fr* a, const GateConsts&) { frw_mac(w, fr_mul(a[0], a[1]), a[2]); }
    // evalx(): the value mac() accumulates, reduced (same gate constants as mac: the small-round kernel, dense_kernel.cuh)
    __device__ __forceinline__ static Fr evalx(const Fr* a, const GateConsts& c) { return eval(a, c); }
};

template <int NARGS>
struct SoFoldedProd {
    static constexpr int P = 2 * NARGS, DEG = 2, HDEG = 2, N_OUTS = NARGS;
    __host__ __device__ static constexpr int gamma_shift(int) { return 0; }
    __device__ __forceinline__ static Fr eval(const Fr* a, const GateConsts& c) {
        Fr ret = fr_mul(a[0], a[NARGS]);  // gammas[0] == 1
#pragma unroll
        for (int i = 1; i < NARGS; i++) ret = fr_add(ret, fr_mul(fr_mul(a[i], a[i + NARGS]), c.g[i]));
        return ret;
    }
    __device__ __forceinline__ static void mac(FrWide& w, const Fr* a, const GateConsts& c) {
        frw_mac(w, a[0], a[NARGS]);
#pragma unroll
        for (int i = 1; i < NARGS; i++) frw_mac(w, fr_mul(a[i], c.g[i]), a[i + NARGS]);
    }
    __device__ __forceinline__ static Fr evalx(const Fr* a, const GateConsts& c) { return eval(a, c); }
};

// EqWrapper(GammaWrapper(G, gamma)): last input is the eq table
template <int G>
struct SoEqGamma {
    static constexpr int P = MoGate<G>::N_INS + 1, DEG = 3, N_OUTS = MoGate<G>::N_OUTS;
    static constexpr int HDEG = MoGate<G>::HOMOG ? MoGate<G>::HOMOG + 1 : 0;
    // power of sigma the host multiplies gamma^i with
    __host__ __device__ static constexpr int gamma_shift(int i) { return MoGate<G>::HOMOG - MoGate<G>::out_deg(i); }
    __device__ __forceinline__ static Fr eval(const Fr* a, const GateConsts& c) {
        return fr_mul(gamma_eval<G>(a, c), a[MoGate<G>::N_INS]);
    }
    __device__ __forceinline__ static void mac(FrWide& w, const Fr* a, const GateConsts& c) {
        if (MoGate<G>::HOMOG)
            frw_mac(w, gamma_eval_scaled<G>(a, c), a[MoGate<G>::N_INS]);
        else
            frw_mac(w, gamma_eval<G>(a, c), a[MoGate<G>::N_INS]);
    }
    __device__ __forceinline__ static Fr evalx(const Fr* a, const GateConsts& c) {
        if (MoGate<G>::HOMOG) return fr_mul(gamma_eval_scaled<G>(a, c), a[MoGate<G>::N_INS]);
        return fr_mul(gamma_eval<G>(a, c), a[MoGate<G>::N_INS]);
    }
};
