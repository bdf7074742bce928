/* TEST INFRASTRUCTURE ONLY -- CPU oracle (plain C restatement of the reference algorithm).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * It is never linked into, imported by, or called from the product path (gkr-msm_b200/).
 *
 * Restates, over 4 x u64 Montgomery limbs (the reference's own representation, ark-ff 0.4.2
 * Fp256<MontBackend<FrConfig,4>>, not vendored under /root/reference):
 *   DenseSumcheckObjectSO::{unipoly,bind}      src/cleanup/protocols/sumcheck.rs:263-332
 *   bind_dense_poly                            src/cleanup/protocols/sumcheck.rs:160-163
 *   eq_poly_sequence_from_multiplier           src/utils.rs:222-250
 *   gates                                      src/cleanup/utils/twisted_edwards_ops.rs:10-81,
 *                                              pushforward.rs:38-50,266-281, logup_mainphase.rs:42-61,
 *                                              multiopen_reduction.rs:13-41, sumcheck.rs:706-741,802-829
 * The reference parallelises with rayon (`--features parallel`); this port uses OpenMP the same way
 * (index-range tasks, per-task accumulators, sums folded at the end) so it can serve as the CPU
 * baseline ("kind": "port").  Parity unpinned by reference golden vectors (the reference has none):
 * pinned instead against the python big-int oracle (tests/test_c_oracle.py) and the literal COEFF_D.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef struct { uint64_t v[4]; } fr_t;

static const uint64_t MOD[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
static const uint64_t INV = 0xfffffffeffffffffULL;
static const fr_t FR_ONE = {{0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL}};
/* COEFF_D, src/utils.rs:34-37 */
static const fr_t FR_D = {{12167860994669987632ULL, 4043113551995129031ULL, 6052647550941614584ULL, 3904213385886034240ULL}};

static inline int geq_mod(const uint64_t a[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > MOD[i]) return 1;
        if (a[i] < MOD[i]) return 0;
    }
    return 1;
}
static inline void sub_mod(uint64_t a[4]) {
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - MOD[i] - br;
        a[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
}
static inline fr_t fr_add(fr_t a, fr_t b) {
    fr_t r;
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a.v[i] + b.v[i];
        r.v[i] = (uint64_t)c;
        c >>= 64;
    }
    if (geq_mod(r.v)) sub_mod(r.v);
    return r;
}
static inline fr_t fr_sub(fr_t a, fr_t b) {
    fr_t r;
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a.v[i] - b.v[i] - br;
        r.v[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
    if (br) {
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)r.v[i] + MOD[i];
            r.v[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    return r;
}
static inline fr_t fr_mul(fr_t a, fr_t b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a.v[j] * b.v[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * INV;
        c = (u128)m * MOD[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * MOD[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    fr_t r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || geq_mod(r.v)) sub_mod(r.v);
    return r;
}
static inline fr_t fr_dbl(fr_t a) { return fr_add(a, a); }
static const fr_t FR_ZERO = {{0, 0, 0, 0}};
/* mul_by_a: a = -5  (src/utils.rs:40-43) ; y - a x = y + 5x */
static inline fr_t add5(fr_t y, fr_t x) { return fr_add(y, fr_add(fr_dbl(fr_dbl(x)), x)); }

/* ---- gates (ids == include/gkr_msm_b200.h) ---- */
enum { G_AFF_L1 = 0, G_AFF_L2, G_AFF_L3, G_PRJ_L1, G_PRJ_L2, G_PRJ_L3, G_TRI_L1, G_BITCHECK, G_LOGUP, G_ADDINV, G_PROD3, G_FOLDED, G_ID, G_AFF_L1_BC2 };

static void gate_io(int g, uint32_t param, int* n_in, int* n_out) {
    switch (g) {
        case G_AFF_L1: *n_in = 4; *n_out = 3; break;
        case G_AFF_L2: *n_in = 3; *n_out = 3; break;
        case G_AFF_L3: *n_in = 3; *n_out = 3; break;
        case G_PRJ_L1: *n_in = 6; *n_out = 4; break;
        case G_PRJ_L2: *n_in = 4; *n_out = 4; break;
        case G_PRJ_L3: *n_in = 4; *n_out = 3; break;
        case G_TRI_L1: *n_in = 12; *n_out = 12; break;
        case G_BITCHECK: *n_in = 1; *n_out = 1; break;
        case G_LOGUP: *n_in = 4; *n_out = 2; break;
        case G_ADDINV: *n_in = 2; *n_out = 2; break;
        case G_PROD3: *n_in = 3; *n_out = 1; break;
        case G_FOLDED: *n_in = 2 * (int)param; *n_out = 1; break;
        case G_ID: *n_in = (int)param; *n_out = (int)param; break;
        case G_AFF_L1_BC2: *n_in = 6; *n_out = 5; break;
        default: *n_in = 0; *n_out = 0;
    }
}

static void prj_l1(const fr_t* a, fr_t* o) {
    o[0] = fr_mul(a[0], a[4]);
    o[1] = fr_mul(a[3], a[1]);
    o[2] = add5(fr_mul(a[1], a[4]), fr_mul(a[0], a[3]));
    o[3] = fr_mul(a[2], a[5]);
}

static void gate_mo(int g, uint32_t param, const fr_t* a, fr_t* o) {
    switch (g) {
        case G_AFF_L1:
            o[0] = fr_mul(a[0], a[3]); o[1] = fr_mul(a[2], a[1]); o[2] = add5(fr_mul(a[1], a[3]), fr_mul(a[0], a[2]));
            break;
        case G_AFF_L2:
            o[0] = fr_add(a[0], a[1]); o[1] = a[2]; o[2] = fr_mul(a[0], a[1]);
            break;
        case G_AFF_L3: {
            fr_t dxy = fr_mul(a[2], FR_D), m = fr_sub(FR_ONE, dxy), p = fr_add(FR_ONE, dxy);
            o[0] = fr_mul(m, a[0]); o[1] = fr_mul(p, a[1]); o[2] = fr_mul(m, p);
            break;
        }
        case G_PRJ_L1: prj_l1(a, o); break;
        case G_PRJ_L2:
            o[0] = fr_mul(fr_add(a[0], a[1]), a[3]); o[1] = fr_mul(a[2], a[3]); o[2] = fr_mul(a[3], a[3]); o[3] = fr_mul(a[0], a[1]);
            break;
        case G_PRJ_L3: {
            fr_t dxy = fr_mul(a[3], FR_D), m = fr_sub(a[2], dxy), p = fr_add(a[2], dxy);
            o[0] = fr_mul(m, a[0]); o[1] = fr_mul(p, a[1]); o[2] = fr_mul(m, p);
            break;
        }
        case G_TRI_L1: {
            fr_t t[6];
            memcpy(t, a, 3 * sizeof(fr_t)); memcpy(t + 3, a + 6, 3 * sizeof(fr_t)); prj_l1(t, o);
            memcpy(t, a + 3, 3 * sizeof(fr_t)); memcpy(t + 3, a + 9, 3 * sizeof(fr_t)); prj_l1(t, o + 4);
            prj_l1(a + 6, o + 8);
            break;
        }
        case G_BITCHECK: o[0] = fr_sub(fr_mul(a[0], a[0]), a[0]); break;
        case G_LOGUP:
            o[0] = fr_add(fr_mul(a[0], a[3]), fr_mul(a[1], a[2])); o[1] = fr_mul(a[1], a[3]);
            break;
        case G_ADDINV: o[0] = fr_add(a[0], a[1]); o[1] = fr_mul(a[0], a[1]); break;
        case G_ID: for (uint32_t i = 0; i < param; i++) o[i] = a[i]; break;
        case G_AFF_L1_BC2:
            gate_mo(G_AFF_L1, 0, a, o);
            o[3] = fr_sub(fr_mul(a[4], a[4]), a[4]); o[4] = fr_sub(fr_mul(a[5], a[5]), a[5]);
            break;
        default: break;
    }
}

/* so_kind 0: PROD3 / FOLDED_PROD(param) ; so_kind 1: EqWrapper(GammaWrapper(gate)) */
typedef struct { int so_kind, gate; uint32_t param; const fr_t* consts; int P, deg; } so_desc;

static int so_init(so_desc* d, int so_kind, int gate, uint32_t param, const fr_t* consts) {
    d->so_kind = so_kind; d->gate = gate; d->param = param; d->consts = consts;
    if (so_kind == 0) {
        if (gate == G_PROD3) { d->P = 3; d->deg = 3; return 0; }
        if (gate == G_FOLDED) { d->P = 2 * (int)param; d->deg = 2; return 0; }
        return -1;
    }
    int ni, no; gate_io(gate, param, &ni, &no);
    if (!ni) return -1;
    d->P = ni + 1; d->deg = 3;
    return 0;
}

static inline fr_t so_eval(const so_desc* d, const fr_t* a) {
    if (d->so_kind == 0) {
        if (d->gate == G_PROD3) return fr_mul(fr_mul(a[0], a[1]), a[2]);
        int n = (int)d->param;
        fr_t r = FR_ZERO;
        for (int i = 0; i < n; i++) r = fr_add(r, fr_mul(fr_mul(a[i], a[i + n]), d->consts[i]));
        return r;
    }
    int ni, no; gate_io(d->gate, d->param, &ni, &no);
    fr_t o[16];
    gate_mo(d->gate, d->param, a, o);
    fr_t r = o[0];
    for (int i = 1; i < no; i++) r = fr_add(r, fr_mul(o[i], d->consts[i]));
    return fr_mul(r, a[ni]);
}

/* sumcheck.rs:283-323: sums at nodes 1..deg over pairs of the current tables (length 2*half). */
static void dense_round_sums(const so_desc* d, fr_t* const* polys, uint64_t half, fr_t* out) {
    const int P = d->P, deg = d->deg;
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    fr_t* part = (fr_t*)calloc((size_t)nthreads * 4, sizeof(fr_t));
#pragma omp parallel
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        fr_t acc[4] = {FR_ZERO, FR_ZERO, FR_ZERO, FR_ZERO};
        fr_t args[16], difs[16];
#pragma omp for schedule(static)
        for (uint64_t i = 0; i < half; i++) {
            for (int j = 0; j < P; j++) args[j] = polys[j][2 * i + 1];
            acc[0] = fr_add(acc[0], so_eval(d, args));
            for (int j = 0; j < P; j++) difs[j] = fr_sub(polys[j][2 * i + 1], polys[j][2 * i]);
            for (int s = 1; s < deg; s++) {
                for (int j = 0; j < P; j++) args[j] = fr_add(args[j], difs[j]);
                acc[s] = fr_add(acc[s], so_eval(d, args));
            }
        }
        for (int s = 0; s < deg; s++) part[(size_t)tid * 4 + s] = acc[s];
    }
    for (int s = 0; s < deg; s++) {
        fr_t t = FR_ZERO;
        for (int k = 0; k < nthreads; k++) t = fr_add(t, part[(size_t)k * 4 + s]);
        out[s] = t;
    }
    free(part);
}

/* sumcheck.rs:160-163 (allocates a fresh half-size vector like the reference) */
static fr_t* bind_dense(const fr_t* p, uint64_t half, fr_t t) {
    fr_t* r = (fr_t*)malloc(sizeof(fr_t) * (half ? half : 1));
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < half; i++) r[i] = fr_add(p[2 * i], fr_mul(t, fr_sub(p[2 * i + 1], p[2 * i])));
    return r;
}

/* Lagrange evaluation on nodes 0..n-1 (UniPoly::from_evals + evaluate) */
static fr_t fr_from_u64(uint64_t x) {
    static const fr_t R2 = {{0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL, 0x0748d9d99f59ff11ULL}};
    fr_t a = {{x, 0, 0, 0}};
    return fr_mul(a, R2);
}
static fr_t fr_pow(fr_t a, const uint64_t e[4]) {
    fr_t r = FR_ONE;
    for (int i = 255; i >= 0; i--) {
        r = fr_mul(r, r);
        if ((e[i / 64] >> (i % 64)) & 1) r = fr_mul(r, a);
    }
    return r;
}
static fr_t fr_inv(fr_t a) {
    uint64_t e[4] = {MOD[0] - 2, MOD[1], MOD[2], MOD[3]};
    return fr_pow(a, e);
}
static fr_t interp_eval(const fr_t* ev, int n, fr_t x) {
    fr_t res = FR_ZERO;
    for (int i = 0; i < n; i++) {
        fr_t num = FR_ONE, den = FR_ONE;
        for (int j = 0; j < n; j++) {
            if (j == i) continue;
            num = fr_mul(num, fr_sub(x, fr_from_u64((uint64_t)j)));
            fr_t dd = i > j ? fr_from_u64((uint64_t)(i - j)) : fr_sub(FR_ZERO, fr_from_u64((uint64_t)(j - i)));
            den = fr_mul(den, dd);
        }
        res = fr_add(res, fr_mul(ev[i], fr_mul(num, fr_inv(den))));
    }
    return res;
}

/* Whole DenseSumcheckObjectSO run with externally supplied challenges.
 * tables: P pointers to 2^nv elements (not modified).  evals_out: nv*(deg+1) elements (nodes 0..deg);
 * final_out: P elements.  rounds_to_run <= nv lets the CPU baseline time a bounded prefix. */
int oracle_dense_sumcheck(int so_kind, int gate, uint32_t param, const uint64_t* consts, int P_in, uint32_t nv,
                          const uint64_t* const* tables, const uint64_t* claim, const uint64_t* challenges,
                          uint32_t rounds_to_run, uint64_t* evals_out, uint64_t* final_out) {
    so_desc d;
    if (so_init(&d, so_kind, gate, param, (const fr_t*)consts)) return -1;
    if (d.P != P_in) return -2;
    fr_t* cur[16];
    int owned = 0;
    for (int j = 0; j < d.P; j++) cur[j] = (fr_t*)tables[j];
    fr_t cl;
    memcpy(&cl, claim, 32);
    for (uint32_t r = 0; r < rounds_to_run && r < nv; r++) {
        uint64_t half = (uint64_t)1 << (nv - r - 1);
        fr_t ev[5];
        dense_round_sums(&d, cur, half, ev + 1);
        ev[0] = fr_sub(cl, ev[1]);
        if (evals_out) memcpy(evals_out + (size_t)r * (d.deg + 1) * 4, ev, sizeof(fr_t) * (d.deg + 1));
        fr_t t;
        memcpy(&t, challenges + 4 * r, 32);
        for (int j = 0; j < d.P; j++) {
            fr_t* nx = bind_dense(cur[j], half, t);
            if (owned) free(cur[j]);
            cur[j] = nx;
        }
        owned = 1;
        cl = interp_eval(ev, d.deg + 1, t);
    }
    if (final_out && rounds_to_run >= nv)
        for (int j = 0; j < d.P; j++) memcpy(final_out + 4 * j, &cur[j][0], 32);
    if (owned)
        for (int j = 0; j < d.P; j++) free(cur[j]);
    return 0;
}

/* eq_poly_sequence_from_multiplier(mult, pt).last()   src/utils.rs:222-250 */
int oracle_eq_table(const uint64_t* point, uint32_t n, const uint64_t* mult, uint64_t* out) {
    fr_t* a = (fr_t*)malloc(sizeof(fr_t) << n);
    fr_t* b = (fr_t*)malloc(sizeof(fr_t) << n);
    memcpy(&a[0], mult, 32);
    for (uint32_t i = 1; i <= n; i++) {
        fr_t r;
        memcpy(&r, point + 4 * (i - 1), 32);
        uint64_t half = (uint64_t)1 << (i - 1);
#pragma omp parallel for schedule(static)
        for (uint64_t j = 0; j < half; j++) {
            fr_t m = fr_mul(r, a[j]);
            b[2 * j] = fr_sub(a[j], m);
            b[2 * j + 1] = m;
        }
        fr_t* t = a; a = b; b = t;
    }
    memcpy(out, a, sizeof(fr_t) << n);
    free(a); free(b);
    return 0;
}

/* synthetic table generator shared with the device (SplitMix64, 4 outputs per element, mod r) */
static inline uint64_t splitmix_at(uint64_t seed, uint64_t k) {
    uint64_t z = seed + k * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
int oracle_synth_table(uint64_t seed, uint64_t n, uint64_t* out) {
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < n; i++) {
        uint64_t v[4];
        for (int j = 0; j < 4; j++) v[j] = splitmix_at(seed, 4 * i + j + 1);
        while (geq_mod(v)) sub_mod(v);
        memcpy(out + 4 * i, v, 32);
    }
    return 0;
}

int oracle_gate_sum(int so_kind, int gate, uint32_t param, const uint64_t* consts, int P_in, uint64_t n,
                    const uint64_t* const* tables, uint64_t* out) {
    so_desc d;
    if (so_init(&d, so_kind, gate, param, (const fr_t*)consts)) return -1;
    if (d.P != P_in) return -2;
    fr_t total = FR_ZERO;
#pragma omp parallel
    {
        fr_t acc = FR_ZERO, args[16];
#pragma omp for schedule(static)
        for (uint64_t i = 0; i < n; i++) {
            for (int j = 0; j < d.P; j++) memcpy(&args[j], tables[j] + 4 * i, 32);
            acc = fr_add(acc, so_eval(&d, args));
        }
#pragma omp critical
        total = fr_add(total, acc);
    }
    memcpy(out, &total, 32);
    return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* single multiplication, for pinning the arithmetic against python big ints */
void oracle_fr_mul(const uint64_t* a, const uint64_t* b, uint64_t* out) {
    fr_t x, y;
    memcpy(&x, a, 32); memcpy(&y, b, 32);
    fr_t r = fr_mul(x, y);
    memcpy(out, &r, 32);
}
