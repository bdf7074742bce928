// Host-side BLS12-381 Fr arithmetic (4 x u64 Montgomery limbs) for the O(1)-per-round glue that the
// reference also keeps on the CPU: UniPoly::from_evals / evaluate (liblasso), UnivarFormat::from12
// (src/cleanup/protocols/sumchecks/vecvec_eq.rs:197-216), compress_coefficients / evaluate_univar
// (src/cleanup/protocols/sumcheck.rs:14-44), gamma powers (src/utils.rs:126-135), eq multipliers.
// Nothing table-sized ever goes through this file.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace gkr {

struct FrH {
    uint64_t v[4];
    bool operator==(const FrH& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2] && v[3] == o.v[3]; }
    bool operator!=(const FrH& o) const { return !(*this == o); }
};

namespace frh {

static const uint64_t MOD[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
static const uint64_t INV = 0xfffffffeffffffffULL;  // -r^-1 mod 2^64
static const FrH ONE = {{0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL}};  // R mod r
static const FrH R2 = {{0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL, 0x0748d9d99f59ff11ULL}};   // R^2 mod r
static const FrH ZERO = {{0, 0, 0, 0}};

typedef unsigned __int128 u128;

static inline bool geq_mod(const uint64_t a[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > MOD[i]) return true;
        if (a[i] < MOD[i]) return false;
    }
    return true;
}

static inline void sub_mod_inplace(uint64_t a[4]) {
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - MOD[i] - br;
        a[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
}

static inline FrH add(const FrH& a, const FrH& b) {
    FrH r;
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a.v[i] + b.v[i];
        r.v[i] = (uint64_t)c;
        c >>= 64;
    }
    if (geq_mod(r.v)) sub_mod_inplace(r.v);
    return r;
}

static inline FrH sub(const FrH& a, const FrH& b) {
    FrH r;
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

static inline FrH neg(const FrH& a) { return sub(ZERO, a); }
static inline FrH dbl(const FrH& a) { return add(a, a); }

static inline FrH mul(const FrH& a, const FrH& b) {
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
    FrH r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || geq_mod(r.v)) sub_mod_inplace(r.v);
    return r;
}

static inline FrH from_u64(uint64_t x) {
    FrH r = {{x, 0, 0, 0}};
    return mul(r, R2);
}

static inline FrH pow(const FrH& a, const uint64_t e[4]) {
    FrH r = ONE;
    for (int i = 255; i >= 0; i--) {
        r = mul(r, r);
        if ((e[i / 64] >> (i % 64)) & 1) r = mul(r, a);
    }
    return r;
}

static inline FrH inverse(const FrH& a) {
    uint64_t e[4] = {MOD[0] - 2, MOD[1], MOD[2], MOD[3]};
    return pow(a, e);
}

static inline bool is_zero(const FrH& a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }

// canonical (non-Montgomery) little-endian bytes  <->  Montgomery limbs
static inline void to_bytes_le(const FrH& a, uint8_t out[32]) {
    FrH one_raw = {{1, 0, 0, 0}};
    FrH c = mul(a, one_raw);  // a * R^-1
    std::memcpy(out, c.v, 32);
}

// F::from_le_bytes_mod_order for inputs of at most 64 bytes (src/cleanup/proof_transcript.rs:33-41): the value is
// lo + hi 2^256 with 256-bit halves; each half is brought below r by at most two subtractions (2^256 < 3 r) and into
// Montgomery form by one multiplication with R^2 (two for the high half).
static inline FrH from_le_bytes_mod_order(const uint8_t* b, size_t n) {
    uint8_t buf[64] = {0};
    std::memcpy(buf, b, n > 64 ? 64 : n);
    FrH lo, hi;
    std::memcpy(lo.v, buf, 32);
    std::memcpy(hi.v, buf + 32, 32);
    auto reduce = [](FrH& a) {
        while (geq_mod(a.v)) {
            unsigned __int128 br = 0;
            for (int i = 0; i < 4; i++) {
                unsigned __int128 d = (unsigned __int128)a.v[i] - MOD[i] - (uint64_t)br;
                a.v[i] = (uint64_t)d;
                br = (d >> 64) & 1;
            }
        }
    };
    reduce(lo);
    FrH acc = mul(lo, R2);
    if (n > 32) {
        reduce(hi);
        acc = add(acc, mul(mul(hi, R2), R2));
    }
    return acc;
}

// Lagrange data of the nodes 0..n-1 for every n <= 8, built ONCE by a function-local static (thread-safe initialisation since
// C++11: two threads proving concurrently on first use cannot observe a half-written table):
//   inv_den[n][i]  = 1 / prod_{j != i} (i - j)
//   basis[n][i][k] = coefficient of X^k in prod_{j != i} (X - j) / prod_{j != i} (i - j)
struct LagrangeTables {
    FrH inv_den[9][8];
    FrH basis[9][8][8];
    LagrangeTables() {
        for (int n = 1; n <= 8; n++) {
            for (int i = 0; i < n; i++) {
                FrH den = ONE;
                for (int j = 0; j < n; j++) {
                    if (j == i) continue;
                    FrH d = (i > j) ? from_u64((uint64_t)(i - j)) : neg(from_u64((uint64_t)(j - i)));
                    den = mul(den, d);
                }
                inv_den[n][i] = inverse(den);
            }
            for (int i = 0; i < n; i++) {
                std::vector<FrH> num(1, ONE);
                for (int j = 0; j < n; j++) {
                    if (j == i) continue;
                    std::vector<FrH> nw(num.size() + 1, ZERO);
                    FrH fj = from_u64((uint64_t)j);
                    for (size_t k = 0; k < num.size(); k++) {
                        nw[k] = sub(nw[k], mul(fj, num[k]));
                        nw[k + 1] = add(nw[k + 1], num[k]);
                    }
                    num.swap(nw);
                }
                for (int k = 0; k < n; k++) basis[n][i][k] = mul(num[k], inv_den[n][i]);
            }
        }
    }
};
static inline const LagrangeTables& lagrange_tables() {
    static const LagrangeTables t;
    return t;
}
static inline const FrH* lagrange_inv_denominators(int n) { return lagrange_tables().inv_den[n]; }
static inline const FrH (*lagrange_basis_coeffs(int n))[8] { return lagrange_tables().basis[n]; }

// coefficients (low -> high) of the unique polynomial of degree < n through (i, evals[i]): UniPoly::from_evals(..).as_vec()
static inline void interpolate_coeffs_into(const FrH* evals, int n, FrH* coeffs) {
    const FrH(*basis)[8] = lagrange_basis_coeffs(n);
    for (int k = 0; k < n; k++) {
        FrH c = ZERO;
        for (int i = 0; i < n; i++) c = add(c, mul(evals[i], basis[i][k]));
        coeffs[k] = c;
    }
}
static inline std::vector<FrH> interpolate_coeffs(const FrH* evals, int n) {
    std::vector<FrH> coeffs(n, ZERO);
    interpolate_coeffs_into(evals, n, coeffs.data());
    return coeffs;
}

// value of that polynomial at x
static inline FrH interpolate_eval(const FrH* evals, int n, const FrH& x) {
    FrH c[8];
    interpolate_coeffs_into(evals, n, c);
    FrH ret = ZERO;
    for (int i = n; i-- > 0;) ret = add(mul(ret, x), c[i]);
    return ret;
}

static inline FrH evaluate_univar(const std::vector<FrH>& coeffs, const FrH& x) {  // sumcheck.rs:33-44
    FrH ret = ZERO;
    for (size_t i = coeffs.size(); i-- > 0;) ret = add(mul(ret, x), coeffs[i]);
    return ret;
}

// eq1(q, t) = 1 - q - t + 2qt
static inline FrH eq1(const FrH& q, const FrH& t) { return add(sub(sub(ONE, q), t), dbl(mul(q, t))); }

}  // namespace frh
}  // namespace gkr
