/* TEST INFRASTRUCTURE ONLY -- CPU oracle, never linked into or called from the product path.
 *
 * The reference's sumcheck engine restated in C++ over Montgomery limbs (written from the reference sources below, NOT from
 * the product's csrc/protocol.cu):
 *   gates (AlgFn / AlgFnSO)          src/cleanup/utils/algfn.rs:11-34,130-291; utils/twisted_edwards_ops.rs:10-81; src/utils.rs:32-49;
 *                                    pushforward/pushforward.rs:28-50,255-281; pushforward/logup_mainphase.rs:32-61;
 *                                    multiopen_reduction.rs:13-41; sumcheck.rs:706-741,802-829
 *   univariate helpers               sumcheck.rs:14-44 (compress / evaluate_univar), liblasso UniPoly::from_evals (unique interpolant)
 *   eq tables, gamma powers          src/utils.rs:126-154,189-291
 *   dense / ragged table ops         polys/dense.rs:39-61,99-185; polys/vecvec.rs:20-206,393-654; utils/algfn.rs:49-90
 *   sumcheck objects                 sumcheck.rs:241-347 (DenseSumcheckObjectSO); sumchecks/dense_eq.rs:43-173; sumchecks/vecvec_eq.rs:53-398
 *   protocols                        sumcheck.rs:101-123 (GenericSumcheckProtocol::prove), :831-889 (DenseEqSumcheck);
 *                                    dense_eq.rs:176-237; vecvec_eq.rs:400-467
 * The reference parallelises with rayon under `--features parallel`; this port uses OpenMP in the same places and
 * additionally in loops the reference leaves serial (dense_eq.rs:121-139, vecvec_eq.rs:320-361, the split maps) -- the sums
 * are exact field sums, so the order does not change a single bit, and a faster CPU baseline is the conservative choice.
 */
#pragma once
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include "po_transcript.hpp"

namespace po {
typedef std::vector<Fr> Vec;

#define PO_ASSERT(c, msg)                                                    \
    do {                                                                     \
        if (!(c)) throw std::runtime_error(std::string("oracle: ") + (msg)); \
    } while (0)
#define PO_MAX_ARITY 160

static const Fr TE_D_M = Fr{{12167860994669987632ULL, 4043113551995129031ULL, 6052647550941614584ULL, 3904213385886034240ULL}};  // COEFF_D, src/utils.rs:34-37
static inline Fr add5(const Fr& y, const Fr& x) { return y + (x.dbl().dbl() + x); }  // y - a x with a = -5 (mul_by_a, src/utils.rs:40-43)

/* ---- gates --------------------------------------------------------------------------------------------------------- */
struct Gate {  // AlgFn: several outputs
    int n_ins = 0, n_outs = 0, deg = 0;
    virtual void exec(const Fr* a, Fr* o) const = 0;
    virtual ~Gate() {}
};
typedef std::shared_ptr<const Gate> GateP;
struct GateSO {  // AlgFnSO: one output
    int n_ins = 0, deg = 0;
    virtual Fr exec(const Fr* a) const = 0;
    virtual ~GateSO() {}
};
typedef std::shared_ptr<const GateSO> GateSOP;

struct AffL1 : Gate {  // twisted_edwards_ops.rs:10-14
    AffL1() { n_ins = 4, n_outs = 3, deg = 2; }
    void exec(const Fr* a, Fr* o) const override {
        o[0] = a[0] * a[3];
        o[1] = a[2] * a[1];
        o[2] = add5(a[1] * a[3], a[0] * a[2]);
    }
};
struct AffL2 : Gate {  // :16-20
    AffL2() { n_ins = 3, n_outs = 3, deg = 2; }
    void exec(const Fr* a, Fr* o) const override {
        o[0] = a[0] + a[1];
        o[1] = a[2];
        o[2] = a[0] * a[1];
    }
};
struct AffL3 : Gate {  // :22-29
    AffL3() { n_ins = 3, n_outs = 3, deg = 2; }
    void exec(const Fr* a, Fr* o) const override {
        Fr dxy = a[2] * TE_D_M, m = Fr::one() - dxy, p = Fr::one() + dxy;
        o[0] = m * a[0];
        o[1] = p * a[1];
        o[2] = m * p;
    }
};
struct PrjL1 : Gate {  // :31-40
    PrjL1() { n_ins = 6, n_outs = 4, deg = 2; }
    void exec(const Fr* a, Fr* o) const override {
        o[0] = a[0] * a[4];
        o[1] = a[3] * a[1];
        o[2] = add5(a[1] * a[4], a[0] * a[3]);
        o[3] = a[2] * a[5];
    }
};
struct PrjL2 : Gate {  // :43-52
    PrjL2() { n_ins = 4, n_outs = 4, deg = 2; }
    void exec(const Fr* a, Fr* o) const override {
        o[0] = (a[0] + a[1]) * a[3];
        o[1] = a[2] * a[3];
        o[2] = a[3] * a[3];
        o[3] = a[0] * a[1];
    }
};
struct PrjL3 : Gate {  // :54-65
    PrjL3() { n_ins = 4, n_outs = 3, deg = 2; }
    void exec(const Fr* a, Fr* o) const override {
        Fr dxy = a[3] * TE_D_M, m = a[2] - dxy, p = a[2] + dxy;
        o[0] = m * a[0];
        o[1] = p * a[1];
        o[2] = m * p;
    }
};
struct TriL1 : Gate {  // :67-80: three projective L1 on (a, c), (b, d), (c, d)
    TriL1() { n_ins = 12, n_outs = 12, deg = 2; }
    void exec(const Fr* p, Fr* o) const override {
        PrjL1 l1;
        Fr in[6];
        const Fr *a = p, *b = p + 3, *c = p + 6, *d = p + 9;
        for (int i = 0; i < 3; i++) in[i] = a[i], in[3 + i] = c[i];
        l1.exec(in, o);
        for (int i = 0; i < 3; i++) in[i] = b[i], in[3 + i] = d[i];
        l1.exec(in, o + 4);
        for (int i = 0; i < 3; i++) in[i] = c[i], in[3 + i] = d[i];
        l1.exec(in, o + 8);
    }
};
struct BitCheck : Gate {  // algfn.rs:262-291
    BitCheck() { n_ins = 1, n_outs = 1, deg = 2; }
    void exec(const Fr* a, Fr* o) const override { o[0] = a[0] * a[0] - a[0]; }
};
struct IdGate : Gate {  // algfn.rs:130-164
    explicit IdGate(int n) { n_ins = n_outs = n, deg = 1; }
    void exec(const Fr* a, Fr* o) const override {
        for (int i = 0; i < n_ins; i++) o[i] = a[i];
    }
};
struct Repeated : Gate {  // algfn.rs:187-225
    GateP f;
    int count;
    Repeated(GateP f_, int c) : f(f_), count(c) { n_ins = f->n_ins * c, n_outs = f->n_outs * c, deg = f->deg; }
    void exec(const Fr* a, Fr* o) const override {
        for (int i = 0; i < count; i++) f->exec(a + i * f->n_ins, o + i * f->n_outs);
    }
};
struct Stacked : Gate {  // algfn.rs:227-259
    GateP f1, f2;
    Stacked(GateP a, GateP b) : f1(a), f2(b) { n_ins = a->n_ins + b->n_ins, n_outs = a->n_outs + b->n_outs, deg = std::max(a->deg, b->deg); }
    void exec(const Fr* a, Fr* o) const override {
        f1->exec(a, o);
        f2->exec(a + f1->n_ins, o + f1->n_outs);
    }
};
struct LogupLayer : Gate {  // logup_mainphase.rs:42-61
    LogupLayer() { n_ins = 4, n_outs = 2, deg = 2; }
    void exec(const Fr* a, Fr* o) const override {
        o[0] = a[0] * a[3] + a[1] * a[2];
        o[1] = a[1] * a[3];
    }
};
struct AddInverses : Gate {  // pushforward.rs:266-281
    AddInverses() { n_ins = 2, n_outs = 2, deg = 2; }
    void exec(const Fr* a, Fr* o) const override {
        o[0] = a[0] + a[1];
        o[1] = a[0] * a[1];
    }
};
struct Prod3 : GateSO {  // pushforward.rs:38-50
    Prod3() { n_ins = 3, deg = 3; }
    Fr exec(const Fr* a) const override { return a[0] * a[1] * a[2]; }
};
static inline Vec make_gamma_pows(const Fr& gamma, size_t count) {  // src/utils.rs:126-135 (always at least [1, gamma])
    Vec g{Fr::one(), gamma};
    for (size_t i = 2; i < count; i++) g.push_back(g[i - 1] * gamma);
    return g;
}
struct FoldedProd : GateSO {  // multiopen_reduction.rs:13-41
    Vec gammas;
    int nargs;
    FoldedProd(const Fr& gamma, int n) : gammas(make_gamma_pows(gamma, n)), nargs(n) { n_ins = 2 * n, deg = 2; }
    Fr exec(const Fr* a) const override {
        Fr s = Fr::zero();
        for (int i = 0; i < nargs; i++) s += a[i] * a[i + nargs] * gammas[i];
        return s;
    }
};
struct GammaWrapper : GateSO {  // sumcheck.rs:706-741: out_0 + sum_i out_i gamma^i
    GateP f;
    Vec gamma_pows;
    GammaWrapper(GateP f_, const Fr& gamma) : f(f_) {
        PO_ASSERT(f->n_outs > 1, "GammaWrapper needs several outputs");
        gamma_pows.push_back(gamma);
        for (int i = 0; i + 2 < f->n_outs; i++) gamma_pows.push_back(gamma * gamma_pows.back());
        n_ins = f->n_ins, deg = f->deg;
    }
    Fr exec(const Fr* a) const override {
        Fr o[PO_MAX_ARITY];
        f->exec(a, o);
        Fr r = o[0];
        for (int i = 1; i < f->n_outs; i++) r += o[i] * gamma_pows[i - 1];
        return r;
    }
};
struct EqWrapper : GateSO {  // sumcheck.rs:802-829
    GateSOP f;
    explicit EqWrapper(GateSOP f_) : f(f_) { n_ins = f->n_ins + 1, deg = f->deg + 1; }
    Fr exec(const Fr* a) const override { return f->exec(a) * a[f->n_ins]; }
};

/* ---- univariate helpers ---------------------------------------------------------------------------------------------- */
/* coefficients (low -> high) of the unique polynomial of degree < n with p(i) = evals[i]  (UniPoly::from_evals) */
static inline Vec unipoly_from_evals(const Vec& evals) {
    const size_t n = evals.size();
    Vec coeffs(n, Fr::zero());
    for (size_t i = 0; i < n; i++) {
        Vec num{Fr::one()};
        Fr den = Fr::one();
        for (size_t j = 0; j < n; j++) {
            if (j == i) continue;
            Vec nw(num.size() + 1, Fr::zero());
            Fr fj = Fr::from_u64(j);
            for (size_t k = 0; k < num.size(); k++) {
                nw[k] -= fj * num[k];
                nw[k + 1] += num[k];
            }
            num = nw;
            den *= (i > j) ? Fr::from_u64(i - j) : -Fr::from_u64(j - i);
        }
        Fr s = evals[i] * den.inverse();
        for (size_t k = 0; k < num.size(); k++) coeffs[k] += num[k] * s;
    }
    return coeffs;
}
static inline Fr evaluate_univar(const Vec& c, const Fr& x) {  // sumcheck.rs:33-44
    Fr r = Fr::zero();
    for (size_t i = c.size(); i-- > 0;) r = r * x + c[i];
    return r;
}
static inline Vec compress_coefficients(const Vec& c) {  // sumcheck.rs:27-31: drop the linear term
    Vec r{c[0]};
    r.insert(r.end(), c.begin() + 2, c.end());
    return r;
}
static inline Fr gamma_rlc(const Fr& gamma, const Vec& v) {  // sumcheck.rs:591-602 == utils.rs zip_with_gamma
    if (v.empty()) return Fr::zero();
    Fr r = v.back();
    for (size_t i = v.size() - 1; i-- > 0;) r = r * gamma + v[i];
    return r;
}
static inline Fr eq1(const Fr& q, const Fr& t) { return Fr::one() - q - t + (q * t).dbl(); }

/* ---- eq tables (src/utils.rs:189-262) ---------------------------------------------------------------------------------- */
static inline std::vector<Vec> eq_poly_sequence_from_multiplier(const Fr& mult, const Fr* pt, size_t l) {
    std::vector<Vec> ret;
    ret.push_back(Vec{mult});
    for (size_t i = 1; i <= l; i++) {
        const Vec& last = ret[i - 1];
        const Fr m_ = pt[i - 1];
        Vec inc((size_t)1 << i);
        const size_t half = (size_t)1 << (i - 1);
#pragma omp parallel for schedule(static) if (half >= 4096)
        for (size_t j = 0; j < half; j++) {
            Fr m = m_ * last[j];
            inc[2 * j] = last[j] - m;
            inc[2 * j + 1] = m;
        }
        ret.push_back(std::move(inc));
    }
    return ret;
}
static inline Vec eq_poly_last_from_multiplier(const Fr& mult, const Fr* pt, size_t l) {
    Vec cur{mult};
    for (size_t i = 1; i <= l; i++) {
        const Fr m_ = pt[i - 1];
        const size_t half = (size_t)1 << (i - 1);
        Vec inc(2 * half);
#pragma omp parallel for schedule(static) if (half >= 4096)
        for (size_t j = 0; j < half; j++) {
            Fr m = m_ * cur[j];
            inc[2 * j] = cur[j] - m;
            inc[2 * j + 1] = m;
        }
        cur.swap(inc);
    }
    return cur;
}
static inline Vec eq_poly_last(const Fr* pt, size_t l) { return eq_poly_last_from_multiplier(Fr::one(), pt, l); }
static inline Vec eq_poly_last(const Vec& pt) { return eq_poly_last(pt.data(), pt.size()); }
static inline std::vector<Vec> padded_eq_poly_sequence(size_t padding, const Fr* pt, size_t l) {  // utils.rs:189-220
    std::vector<Vec> ret;
    ret.push_back(Vec{Fr::one()});
    for (size_t i = 1; i <= padding; i++) ret.push_back(Vec{ret[i - 1][0] * (Fr::one() - pt[i - 1])});
    for (size_t i = padding + 1; i <= l; i++) {
        const Vec& last = ret[i - 1];
        const Fr m_ = pt[i - 1];
        const size_t half = (size_t)1 << (i - 1 - padding);
        Vec inc(2 * half);
        for (size_t j = 0; j < half; j++) {
            Fr m = m_ * last[j];
            inc[2 * j] = last[j] - m;
            inc[2 * j + 1] = m;
        }
        ret.push_back(std::move(inc));
    }
    return ret;
}
static inline Fr evaluate_poly(const Vec& poly, const Vec& pt) {  // cleanup/utils/arith.rs:6-9
    Vec e = eq_poly_last(pt);
    PO_ASSERT(e.size() == poly.size(), "evaluate_poly: length mismatch");
    Fr s = Fr::zero();
    for (size_t i = 0; i < e.size(); i++) s += poly[i] * e[i];
    return s;
}
static inline size_t log_2(size_t n) {  // liblasso Math::log_2: exact for powers of two, ceil otherwise
    PO_ASSERT(n != 0, "log_2(0)");
    size_t l = 0;
    while (((size_t)1 << l) < n) l++;
    return l;
}

/* ---- ragged matrices (polys/vecvec.rs:149-206) ---------------------------------------------------------------------------- */
struct VecVec {
    std::vector<Vec> data;
    Fr row_pad, col_pad;
    size_t row_logsize = 0, col_logsize = 0;
    VecVec() {}
    VecVec(std::vector<Vec> d, const Fr& rp, const Fr& cp, size_t rl, size_t cl, bool unchecked = false)
        : data(std::move(d)), row_pad(rp), col_pad(cp), row_logsize(rl), col_logsize(cl) {
        PO_ASSERT(data.size() <= ((size_t)1 << cl), "VecVecPolynomial: too many rows");
        if (!unchecked)
            for (auto& r : data) {
                PO_ASSERT(r.size() <= ((size_t)1 << rl), "VecVecPolynomial: row too long");
                if (r.size() % 2 == 1) r.push_back(row_pad);  // vecvec.rs:183-187
            }
    }
    void make_21() {  // vecvec.rs:400-413
#pragma omp parallel for schedule(dynamic, 64) if (data.size() >= 256)
        for (size_t k = 0; k < data.size(); k++) {
            Vec& r = data[k];
            for (size_t i = 0; i < r.size() / 2; i++) r[2 * i] = r[2 * i + 1].dbl() - r[2 * i];
        }
    }
    void bind_21(const Fr& t) {  // vecvec.rs:420-441
        const Fr tm1 = t - Fr::one();
#pragma omp parallel for schedule(dynamic, 64) if (data.size() >= 256)
        for (size_t k = 0; k < data.size(); k++) {
            Vec& r = data[k];
            const size_t h = r.size() / 2;
            Vec nw(h);
            for (size_t i = 0; i < h; i++) nw[i] = r[2 * i + 1] + tm1 * (r[2 * i] - r[2 * i + 1]);
            if (h % 2 == 1) nw.push_back(row_pad);
            r.swap(nw);
        }
        row_logsize -= 1;
    }
};

/* ---- witness maps --------------------------------------------------------------------------------------------------------- */
struct SplitIdx {  // splits.rs:13-50
    bool lo;
    size_t k;
    static SplitIdx LO(size_t k) { return SplitIdx{true, k}; }
    static SplitIdx HI(size_t k) { return SplitIdx{false, k}; }
    size_t lo_usize(size_t num_vars) const { return lo ? k : num_vars - k - 1; }
    size_t hi_usize(size_t num_vars) const { return lo ? num_vars - k - 1 : k; }
};
template <class T>
static inline std::vector<T> interleave_bundles(std::vector<T>& l, std::vector<T>& r, size_t bundle) {  // dense.rs:137-138
    std::vector<T> out;
    size_t nl = (l.size() + bundle - 1) / bundle, nr = (r.size() + bundle - 1) / bundle;
    for (size_t i = 0; i < std::max(nl, nr); i++) {
        for (size_t k = i * bundle; k < std::min(l.size(), (i + 1) * bundle); k++) out.push_back(std::move(l[k]));
        for (size_t k = i * bundle; k < std::min(r.size(), (i + 1) * bundle); k++) out.push_back(std::move(r[k]));
    }
    return out;
}
static inline std::vector<Vec> dense_map(const std::vector<const Vec*>& polys, const Gate& f) {  // dense.rs:141-184
    const size_t n = polys[0]->size();
    std::vector<Vec> outs(f.n_outs, Vec(n));
#pragma omp parallel for schedule(static) if (n >= 256)
    for (size_t i = 0; i < n; i++) {
        Fr a[PO_MAX_ARITY], o[PO_MAX_ARITY];
        for (int j = 0; j < f.n_ins; j++) a[j] = (*polys[j])[i];
        f.exec(a, o);
        for (int k = 0; k < f.n_outs; k++) outs[k][i] = o[k];
    }
    return outs;
}
static inline std::vector<const Vec*> ptrs(const std::vector<Vec>& v, size_t n) {
    std::vector<const Vec*> p;
    for (size_t i = 0; i < n; i++) p.push_back(&v[i]);
    return p;
}
static inline std::vector<Vec> dense_map_split(const std::vector<Vec>& polys, const Gate& f, SplitIdx idx, size_t bundle) {  // dense.rs:115-139
    const size_t n = polys[0].size(), num_vars = log_2(n), seg = (size_t)1 << idx.lo_usize(num_vars);
    std::vector<Vec> full = dense_map(ptrs(polys, f.n_ins), f);
    std::vector<Vec> l(f.n_outs), r(f.n_outs);
    for (int k = 0; k < f.n_outs; k++) {
        l[k].reserve(n / 2);
        r[k].reserve(n / 2);
        for (size_t i = 0; i < n; i++) ((i / seg) % 2 ? r : l)[k].push_back(full[k][i]);
    }
    return interleave_bundles(l, r, bundle);
}
static inline void map_split_hi(const std::vector<const Vec*>& polys, const Gate& f, std::vector<Vec>* out0, std::vector<Vec>* out1) {  // algfn.rs:82-89
    const size_t half = polys[0]->size() / 2;
    std::vector<Vec> lo(f.n_outs, Vec(half)), hi(f.n_outs, Vec(half));
#pragma omp parallel for schedule(static) if (half >= 256)
    for (size_t i = 0; i < half; i++) {
        Fr a[PO_MAX_ARITY], o[PO_MAX_ARITY];
        for (int j = 0; j < f.n_ins; j++) a[j] = (*polys[j])[i];
        f.exec(a, o);
        for (int k = 0; k < f.n_outs; k++) lo[k][i] = o[k];
        for (int j = 0; j < f.n_ins; j++) a[j] = (*polys[j])[half + i];
        f.exec(a, o);
        for (int k = 0; k < f.n_outs; k++) hi[k][i] = o[k];
    }
    *out0 = std::move(lo);
    *out1 = std::move(hi);
}
static inline void gate_pads(const std::vector<VecVec>& polys, const Gate& f, Fr* rp, Fr* cp) {
    Fr a[PO_MAX_ARITY];
    for (int j = 0; j < f.n_ins; j++) a[j] = polys[j].row_pad;
    f.exec(a, rp);
    for (int j = 0; j < f.n_ins; j++) a[j] = polys[j].col_pad;
    f.exec(a, cp);
}
static inline std::vector<VecVec> vecvec_map(const std::vector<VecVec>& polys, const Gate& f) {  // vecvec.rs:480-540
    Fr rp[PO_MAX_ARITY], cp[PO_MAX_ARITY];
    gate_pads(polys, f, rp, cp);
    const size_t rows = polys[0].data.size();
    std::vector<std::vector<Vec>> datas(f.n_outs, std::vector<Vec>(rows));
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t r = 0; r < rows; r++) {
        const size_t len = polys[0].data[r].size();
        for (int k = 0; k < f.n_outs; k++) datas[k][r].resize(len);
        Fr a[PO_MAX_ARITY], o[PO_MAX_ARITY];
        for (size_t i = 0; i < len; i++) {
            for (int j = 0; j < f.n_ins; j++) a[j] = polys[j].data[r][i];
            f.exec(a, o);
            for (int k = 0; k < f.n_outs; k++) datas[k][r][i] = o[k];
        }
    }
    std::vector<VecVec> out;
    for (int k = 0; k < f.n_outs; k++) out.emplace_back(std::move(datas[k]), rp[k], cp[k], polys[0].row_logsize, polys[0].col_logsize);
    return out;
}
static inline std::vector<VecVec> vecvec_map_split(const std::vector<VecVec>& polys, const Gate& f, SplitIdx idx, size_t bundle) {  // vecvec.rs:542-606
    Fr rp[PO_MAX_ARITY], cp[PO_MAX_ARITY];
    gate_pads(polys, f, rp, cp);
    const size_t rl = polys[0].row_logsize, cl = polys[0].col_logsize, seg = (size_t)1 << idx.lo_usize(rl + cl);
    const size_t rows = polys[0].data.size();
    std::vector<std::vector<Vec>> dl(f.n_outs, std::vector<Vec>(rows)), dr(f.n_outs, std::vector<Vec>(rows));
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t r = 0; r < rows; r++) {
        const size_t len = polys[0].data[r].size();
        Fr a[PO_MAX_ARITY], o[PO_MAX_ARITY];
        for (size_t i = 0; i < len; i++) {
            for (int j = 0; j < f.n_ins; j++) a[j] = polys[j].data[r][i];
            f.exec(a, o);
            auto& side = ((i / seg) % 2) ? dr : dl;
            for (int k = 0; k < f.n_outs; k++) side[k][r].push_back(o[k]);
        }
        if (dl[0][r].size() % 2 == 1)  // vecvec.rs:585-592: both halves are re-padded to an even length
            for (int k = 0; k < f.n_outs; k++) {
                dl[k][r].push_back(rp[k]);
                dr[k][r].push_back(rp[k]);
            }
    }
    std::vector<VecVec> l, rr;
    for (int k = 0; k < f.n_outs; k++) {
        l.emplace_back(std::move(dl[k]), rp[k], cp[k], rl - 1, cl, true);
        rr.emplace_back(std::move(dr[k]), rp[k], cp[k], rl - 1, cl, true);
    }
    return interleave_bundles(l, rr, bundle);
}
static inline std::vector<Vec> vecvec_map_split_to_dense(const std::vector<VecVec>& polys, const Gate& f, SplitIdx idx, size_t bundle) {  // vecvec.rs:608-654
    Fr rp[PO_MAX_ARITY], cp[PO_MAX_ARITY];
    gate_pads(polys, f, rp, cp);
    const size_t rl = polys[0].row_logsize, cl = polys[0].col_logsize, seg = (size_t)1 << idx.lo_usize(rl + cl);
    PO_ASSERT(rl == 1, "vecvec_map_split_to_dense: row_logsize must be 1");
    std::vector<Vec> l(f.n_outs), r(f.n_outs);
    for (size_t row = 0; row < polys[0].data.size(); row++) {
        Fr a[PO_MAX_ARITY], o[PO_MAX_ARITY];
        for (size_t i = 0; i < polys[0].data[row].size(); i++) {
            for (int j = 0; j < f.n_ins; j++) a[j] = polys[j].data[row][i];
            f.exec(a, o);
            for (int k = 0; k < f.n_outs; k++) (((i / seg) % 2) ? r : l)[k].push_back(o[k]);
        }
        if (l[0].size() < row + 1)
            for (int k = 0; k < f.n_outs; k++) {
                l[k].push_back(rp[k]);
                r[k].push_back(rp[k]);
            }
    }
    const size_t n = (size_t)1 << cl;
    for (int k = 0; k < f.n_outs; k++) {
        l[k].resize(n, cp[k]);
        r[k].resize(n, cp[k]);
    }
    return interleave_bundles(l, r, bundle);
}

/* ---- sumcheck objects ------------------------------------------------------------------------------------------------- */
struct Sumcheckable {  // vecvec_eq.rs:218-225
    virtual Vec unipoly() = 0;
    virtual void bind(const Fr& t) = 0;
    virtual Vec final_evals() = 0;
    virtual Fr claim() const = 0;
    virtual ~Sumcheckable() {}
};

static inline void bind_dense_poly(Vec& p, const Fr& t) {  // sumcheck.rs:160-163
    const size_t h = p.size() / 2;
    Vec nw(h);
#pragma omp parallel for schedule(static) if (h >= 2048)
    for (size_t i = 0; i < h; i++) nw[i] = p[2 * i] + t * (p[2 * i + 1] - p[2 * i]);
    p.swap(nw);
}

struct DenseSO : Sumcheckable {  // sumcheck.rs:241-347
    std::vector<Vec> polys;
    GateSOP f;
    size_t num_vars, round_idx = 0;
    Fr claim_;
    bool cached = false;
    Vec cached_unipoly;
    DenseSO(std::vector<Vec> p, GateSOP f_, size_t nv, const Fr& claim_hint) : polys(std::move(p)), f(f_), num_vars(nv), claim_(claim_hint) {
        PO_ASSERT((int)polys.size() == f->n_ins, "DenseSumcheckObjectSO: number of tables != n_ins");
        for (auto& v : polys) PO_ASSERT(v.size() == ((size_t)1 << nv), "DenseSumcheckObjectSO: table length");
    }
    Fr claim() const override { return claim_; }
    Vec unipoly() override {
        PO_ASSERT(round_idx < num_vars, "the protocol has already ended");
        if (cached) return cached_unipoly;
        const size_t half = (size_t)1 << (num_vars - round_idx - 1);
        const int P = (int)polys.size(), D = f->deg;
        Vec total(D + 1, Fr::zero());
#pragma omp parallel if (half >= 512)
        {
            Fr acc[8];
            for (int s = 0; s < D; s++) acc[s] = Fr::zero();
            Fr args[PO_MAX_ARITY], difs[PO_MAX_ARITY];
#pragma omp for schedule(static) nowait
            for (size_t i = 0; i < half; i++) {
                for (int j = 0; j < P; j++) args[j] = polys[j][2 * i + 1];
                acc[0] += f->exec(args);
                for (int j = 0; j < P; j++) difs[j] = polys[j][2 * i + 1] - polys[j][2 * i];
                for (int s = 1; s < D; s++) {
                    for (int j = 0; j < P; j++) args[j] += difs[j];
                    acc[s] += f->exec(args);
                }
            }
#pragma omp critical
            for (int s = 0; s < D; s++) total[s + 1] += acc[s];
        }
        total[0] = claim_ - total[1];
        cached_unipoly = unipoly_from_evals(total);
        cached = true;
        return cached_unipoly;
    }
    void bind(const Fr& t) override {
        PO_ASSERT(round_idx < num_vars, "the protocol has already ended");
        PO_ASSERT(cached, "should evaluate unipoly before binding");
        for (auto& p : polys) bind_dense_poly(p, t);
        round_idx++;
        claim_ = evaluate_univar(cached_unipoly, t);
        cached = false;
    }
    Vec final_evals() override {
        PO_ASSERT(round_idx == num_vars, "can only call final evals after the last round");
        Vec r;
        for (auto& p : polys) r.push_back(p[0]);
        return r;
    }
};

static inline Vec univar_from12(const Fr& p1, const Fr& p2, const Fr& e1, const Fr& previous_claim) {  // vecvec_eq.rs:197-216
    Fr e0 = Fr::one() - e1, e2 = e1.dbl() - e0, e3 = e2.dbl() - e1;
    Fr prod1 = p1 * e1, prod0 = previous_claim - prod1;
    PO_ASSERT(!e0.is_zero(), "from12: eq0 is not invertible");
    Fr p0 = prod0 * e0.inverse();
    Fr p3 = p2.dbl() + p2 - p1.dbl() - p1 + p0;
    return unipoly_from_evals(Vec{prod0, prod1, p2 * e2, p3 * e3});
}

struct DenseDeg2SO : Sumcheckable {  // dense_eq.rs:62-173; tables may be shorter than 2^n (implicit zero padding)
    std::vector<Vec> polys;
    GateP func;
    Vec gamma_pows, point;
    Fr claim_, multiplier = Fr::one();
    std::vector<Vec> eq_poly_data;
    bool cached = false;
    Vec cached_unipoly;
    DenseDeg2SO(std::vector<Vec> p, GateP f, Vec gp, const Fr& claim, Vec pt)
        : polys(std::move(p)), func(f), gamma_pows(std::move(gp)), point(std::move(pt)), claim_(claim) {
        eq_poly_data = eq_poly_sequence_from_multiplier(Fr::one(), point.data(), point.size() - 1);
    }
    Fr claim() const override { return claim_; }
    Vec unipoly() override {
        PO_ASSERT(!cached, "unipoly called twice");
        for (auto& v : polys)
            for (size_t i = 0; i < v.size() / 2; i++) v[2 * i] = v[2 * i + 1].dbl() - v[2 * i];  // make_21, dense.rs:99-112
        const int P = (int)polys.size(), NO = func->n_outs;
        Fr zero_in[PO_MAX_ARITY], pad_results[PO_MAX_ARITY];
        for (int j = 0; j < P; j++) zero_in[j] = Fr::zero();
        func->exec(zero_in, pad_results);
        const Vec& eq = eq_poly_data.back();
        const size_t half = polys[0].size() / 2;
        Vec sum2(NO, Fr::zero()), sum1(NO, Fr::zero());
        Fr eq_sum = Fr::zero();
#pragma omp parallel if (half >= 256)
        {
            Fr l2[PO_MAX_ARITY], l1[PO_MAX_ARITY], a[PO_MAX_ARITY], o[PO_MAX_ARITY], es = Fr::zero();
            for (int i = 0; i < NO; i++) l2[i] = l1[i] = Fr::zero();
#pragma omp for schedule(static) nowait
            for (size_t idx = 0; idx < half; idx++) {
                for (int j = 0; j < P; j++) a[j] = polys[j][2 * idx];
                func->exec(a, o);
                for (int i = 0; i < NO; i++) l2[i] += o[i] * eq[idx];
                for (int j = 0; j < P; j++) a[j] = polys[j][2 * idx + 1];
                func->exec(a, o);
                for (int i = 0; i < NO; i++) l1[i] += o[i] * eq[idx];
                es += eq[idx];
            }
#pragma omp critical
            {
                for (int i = 0; i < NO; i++) sum2[i] += l2[i], sum1[i] += l1[i];
                eq_sum += es;
            }
        }
        const Fr trailing = Fr::one() - eq_sum;
        for (int i = 0; i < NO; i++) {
            sum2[i] += pad_results[i] * trailing;
            sum1[i] += pad_results[i] * trailing;
        }
        Fr total2 = sum2[0], total1 = sum1[0];
        for (int i = 1; i < NO; i++) {
            total2 += sum2[i] * gamma_pows[i];
            total1 += sum1[i] * gamma_pows[i];
        }
        total2 *= multiplier;
        total1 *= multiplier;
        cached_unipoly = univar_from12(total1, total2, point.back(), claim_);
        cached = true;
        return cached_unipoly;
    }
    void bind(const Fr& t) override {
        multiplier *= eq1(point.back(), t);
        const Fr tm1 = t - Fr::one();
        for (auto& v : polys) {  // bind_21, dense.rs:54-61
            PO_ASSERT(v.size() % 2 == 0, "bind_21: odd length");
            const size_t h = v.size() / 2;
            Vec nw(h);
#pragma omp parallel for schedule(static) if (h >= 2048)
            for (size_t i = 0; i < h; i++) nw[i] = v[2 * i + 1] + tm1 * (v[2 * i] - v[2 * i + 1]);
            v.swap(nw);
        }
        eq_poly_data.pop_back();
        point.pop_back();
        claim_ = evaluate_univar(cached_unipoly, t);
        cached = false;
    }
    Vec final_evals() override {
        Vec r;
        for (auto& p : polys) r.push_back(p[0]);
        return r;
    }
};

struct EQPolyData {  // vecvec.rs:20-147
    size_t padded_vars_idx, segment_vars_idx;
    long binding_var_idx;  // -1 == None
    Vec point, row_eq_coefs, row_eq_coefs_tail_sums;
    std::vector<Vec> row_eq_poly_seq, row_eq_poly_prefix_seq;
    Fr multiplier = Fr::one();
    size_t already_bound_vars = 0;
    EQPolyData() {}
    EQPolyData(const Vec& pt, size_t col_logsize, size_t max_row_len) : point(pt) {
        const size_t max_segment_logsize = log_2(max_row_len);
        padded_vars_idx = col_logsize;
        segment_vars_idx = pt.size() - max_segment_logsize;
        binding_var_idx = (long)pt.size() - 1;
        row_eq_coefs = eq_poly_last(point.data(), col_logsize);
        row_eq_coefs_tail_sums.resize(row_eq_coefs.size());
        Fr acc = Fr::zero();
        for (size_t i = row_eq_coefs.size(); i-- > 0;) {
            acc += row_eq_coefs[i];
            row_eq_coefs_tail_sums[i] = acc;
        }
        const size_t pad_hi = std::min(segment_vars_idx, (size_t)binding_var_idx), row_hi = std::max(segment_vars_idx, (size_t)binding_var_idx);
        const size_t padding = pad_hi > padded_vars_idx ? pad_hi - padded_vars_idx : 0;
        row_eq_poly_seq = padded_eq_poly_sequence(padding, point.data() + padded_vars_idx, row_hi - padded_vars_idx);
        for (auto& v : row_eq_poly_seq) {
            Vec pre(v.size() + 1);
            pre[0] = Fr::zero();
            for (size_t i = 0; i < v.size(); i++) pre[i + 1] = pre[i] + v[i];
            row_eq_poly_prefix_seq.push_back(std::move(pre));
        }
    }
    void bind(const Fr& t) {
        multiplier *= eq1(point[binding_var_idx], t);
        if (binding_var_idx >= 0) binding_var_idx -= 1;  // Some(0) -> None
        already_bound_vars++;
    }
    const Vec& current_evals() const { return row_eq_poly_seq[row_eq_poly_seq.size() - 1 - already_bound_vars]; }
    Fr trailing_sum(size_t segment_len) const { return Fr::one() - row_eq_poly_prefix_seq[row_eq_poly_prefix_seq.size() - 1 - already_bound_vars][segment_len]; }
};

struct VecVecDeg2SO : Sumcheckable {  // vecvec_eq.rs:74-398: sparse stage, then DenseSumcheckObjectSO over EqWrapper(GammaWrapper(func))
    std::vector<VecVec> polys;
    GateP func;
    Vec gamma_pows;
    Fr claim_;
    EQPolyData eq;
    bool cached = false;
    Vec cached_unipoly;
    std::unique_ptr<DenseSO> dense;
    VecVecDeg2SO(std::vector<VecVec> p, GateP f, Vec gp, const Fr& claim, const Vec& point, size_t col_logsize)
        : polys(std::move(p)), func(f), gamma_pows(std::move(gp)), claim_(claim) {
        size_t mx = 0;
        for (auto& r : polys[0].data) mx = std::max(mx, r.size());
        eq = EQPolyData(point, col_logsize, mx);
    }
    Fr claim() const override { return dense ? dense->claim() : claim_; }
    Vec unipoly() override {
        if (dense) return dense->unipoly();
        PO_ASSERT(!cached, "unipoly called twice");
        for (auto& p : polys) p.make_21();
        const int P = (int)polys.size(), NO = func->n_outs;
        Fr pad_results[PO_MAX_ARITY], col_pad_results[PO_MAX_ARITY];
        gate_pads(polys, *func, pad_results, col_pad_results);
        Vec sum2(NO, Fr::zero()), sum1(NO, Fr::zero());
        const size_t row_count = polys[0].data.size();
        const Vec& eqv = eq.current_evals();
#pragma omp parallel if (row_count * eqv.size() >= 2048)
        {
            Fr t2[PO_MAX_ARITY], t1[PO_MAX_ARITY], l2[PO_MAX_ARITY], l1[PO_MAX_ARITY], a[PO_MAX_ARITY], o[PO_MAX_ARITY];
            for (int i = 0; i < NO; i++) t2[i] = t1[i] = Fr::zero();
#pragma omp for schedule(dynamic, 16) nowait
            for (size_t row = 0; row < row_count; row++) {
                for (int i = 0; i < NO; i++) l2[i] = l1[i] = Fr::zero();
                const size_t segment_len = polys[0].data[row].size() / 2;
                for (size_t idx = 0; idx < segment_len; idx++) {
                    for (int j = 0; j < P; j++) a[j] = polys[j].data[row][2 * idx];
                    func->exec(a, o);
                    for (int i = 0; i < NO; i++) l2[i] += o[i] * eqv[idx];
                    for (int j = 0; j < P; j++) a[j] = polys[j].data[row][2 * idx + 1];
                    func->exec(a, o);
                    for (int i = 0; i < NO; i++) l1[i] += o[i] * eqv[idx];
                }
                const Fr trailing = eq.trailing_sum(segment_len), vmul = eq.row_eq_coefs[row];
                for (int i = 0; i < NO; i++) {
                    t2[i] += (l2[i] + pad_results[i] * trailing) * vmul;
                    t1[i] += (l1[i] + pad_results[i] * trailing) * vmul;
                }
            }
#pragma omp critical
            for (int i = 0; i < NO; i++) sum2[i] += t2[i], sum1[i] += t1[i];
        }
        if (row_count < ((size_t)1 << eq.padded_vars_idx))
            for (int i = 0; i < NO; i++) {
                Fr res = col_pad_results[i] * eq.row_eq_coefs_tail_sums[row_count];
                sum2[i] += res;
                sum1[i] += res;
            }
        Fr total2 = sum2[0], total1 = sum1[0];
        for (int i = 1; i < NO; i++) {
            total2 += sum2[i] * gamma_pows[i];
            total1 += sum1[i] * gamma_pows[i];
        }
        total2 *= eq.multiplier;
        total1 *= eq.multiplier;
        cached_unipoly = univar_from12(total1, total2, eq.point[eq.binding_var_idx], claim_);
        cached = true;
        return cached_unipoly;
    }
    void bind(const Fr& t) override {
        if (dense) {
            dense->bind(t);
            return;
        }
        if ((size_t)eq.binding_var_idx > eq.padded_vars_idx) {  // vecvec_eq.rs:235
            for (auto& p : polys) p.bind_21(t);
            eq.bind(t);
            claim_ = evaluate_univar(cached_unipoly, t);
            cached = false;
            return;
        }
        // bind_into_dense, vecvec_eq.rs:157-190
        const Fr tm1 = t - Fr::one();
        const size_t n = (size_t)1 << eq.padded_vars_idx;
        std::vector<Vec> dp;
        for (auto& p : polys) {
            Vec col;
            col.reserve(n);
            for (auto& r : p.data) {
                if (col.size() == n) break;
                if (r.size() == 0) col.push_back(p.row_pad);
                else if (r.size() == 2) col.push_back(r[1] + tm1 * (r[0] - r[1]));
                else PO_ASSERT(false, "bind_into_dense: unreachable row length");
            }
            col.resize(n, p.col_pad);
            dp.push_back(std::move(col));
        }
        const Fr mult = eq.multiplier * eq1(eq.point[eq.binding_var_idx], t);
        dp.push_back(eq_poly_last_from_multiplier(mult, eq.point.data(), eq.padded_vars_idx));
        GateSOP g = std::make_shared<EqWrapper>(std::make_shared<GammaWrapper>(func, gamma_pows[1]));
        dense.reset(new DenseSO(std::move(dp), g, eq.padded_vars_idx, evaluate_univar(cached_unipoly, t)));
        cached = false;
        polys.clear();
    }
    Vec final_evals() override {
        PO_ASSERT(dense != nullptr, "final_evals in the sparse stage");
        return dense->final_evals();
    }
};

/* ---- protocols --------------------------------------------------------------------------------------------------------- */
struct Claims {  // EvalClaim / SinglePointClaims: one point, several evaluations
    Vec point, evs;
};

/* GenericSumcheckProtocol::prove, sumcheck.rs:101-123: returns the output point (reversed challenges) */
static inline Vec generic_sumcheck_prove(Transcript& tr, size_t rounds, size_t msg_len, Sumcheckable& so, Vec* final_evals) {
    Vec r;
    for (size_t i = 0; i < rounds; i++) {
        Vec poly = so.unipoly();
        Vec msg = compress_coefficients(poly);
        PO_ASSERT(msg.size() == msg_len, "round message length != degree");
        tr.write_scalars(msg);
        Fr x = tr.challenge(128);
        r.push_back(x);
        so.bind(x);
    }
    std::reverse(r.begin(), r.end());
    *final_evals = so.final_evals();
    return r;
}
static inline Fr rlc_claim(const Vec& gp, const Vec& claims) {  // dense_eq.rs:43-60 / vecvec_eq.rs:53-71
    Fr c = claims[0];
    for (size_t i = 1; i < claims.size(); i++) c += gp[i] * claims[i];
    return c;
}
/* DenseEqSumcheck::prove, sumcheck.rs:849-872 */
static inline Claims dense_eq_sumcheck_prove(Transcript& tr, GateP f, size_t num_vars, const Claims& claims, std::vector<Vec> advice) {
    Fr gamma = tr.challenge(128);
    advice.push_back(eq_poly_last(claims.point));
    GateSOP g = std::make_shared<EqWrapper>(std::make_shared<GammaWrapper>(f, gamma));
    DenseSO so(std::move(advice), g, claims.point.size(), gamma_rlc(gamma, claims.evs));
    Vec fe;
    Vec pt = generic_sumcheck_prove(tr, num_vars, f->deg + 1, so, &fe);
    fe.pop_back();
    tr.write_scalars(fe);
    return Claims{pt, fe};
}
/* DenseDeg2Sumcheck::prove, dense_eq.rs:199-214 */
static inline Claims dense_deg2_sumcheck_prove(Transcript& tr, GateP f, size_t num_vars, const Claims& claims, std::vector<Vec> advice) {
    PO_ASSERT(f->deg == 2, "DenseDeg2Sumcheck: degree");
    Fr gamma = tr.challenge(128);
    Vec gp = make_gamma_pows(gamma, f->n_outs);
    DenseDeg2SO so(std::move(advice), f, gp, rlc_claim(gp, claims.evs), claims.point);
    Vec fe;
    Vec pt = generic_sumcheck_prove(tr, num_vars, 3, so, &fe);
    tr.write_scalars(fe);
    return Claims{pt, fe};
}
/* VecVecDeg2Sumcheck::prove, vecvec_eq.rs:425-443 */
static inline Claims vecvec_deg2_sumcheck_prove(Transcript& tr, GateP f, size_t num_vars, size_t num_vertical_vars, const Claims& claims, std::vector<VecVec> advice) {
    PO_ASSERT(f->deg == 2, "VecVecDeg2Sumcheck: degree");
    Fr gamma = tr.challenge(128);
    Vec gp = make_gamma_pows(gamma, f->n_outs);
    VecVecDeg2SO so(std::move(advice), f, gp, rlc_claim(gp, claims.evs), claims.point, num_vertical_vars);
    Vec fe;
    Vec pt = generic_sumcheck_prove(tr, num_vars, 3, so, &fe);
    fe.pop_back();
    tr.write_scalars(fe);
    return Claims{pt, fe};
}
}  // namespace po
