/* TEST INFRASTRUCTURE ONLY -- CPU oracle (C++ restatement of the reference algorithm), never linked into or called from the
 * product path (gkr-msm_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load it.
 *
 * Prime fields of the hot path in the reference's own representation (ark-ff 0.4.2 `Fp<MontBackend<_, N>, N>`, not vendored
 * under /root/reference -- Cargo.lock:112-248): N little-endian u64 limbs of x * R mod p, R = 2^(64 N).
 *   Fr = BLS12-381 scalar field (N = 4)  -- every sumcheck / GKR table element        (src/cleanup/protocols/pippenger.rs:519)
 *   Fq = BLS12-381 base field   (N = 6)  -- coordinates of the G1 commitments         (src/commitments/kzg.rs)
 * Published parameters only; all results are mathematically unique, so this is an exact restatement.
 */
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace po {
typedef unsigned __int128 u128;

template <int N>
struct FpParams;

template <>
struct FpParams<4> {  // r = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    static constexpr uint64_t MOD[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
    static constexpr uint64_t INV = 0xfffffffeffffffffULL;  // -r^-1 mod 2^64
    static constexpr uint64_t ONE[4] = {0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL};
    static constexpr uint64_t R2[4] = {0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL, 0x0748d9d99f59ff11ULL};
};
template <>
struct FpParams<6> {  // q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    static constexpr uint64_t MOD[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                                        0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
    static constexpr uint64_t INV = 0x89f3fffcfffcfffdULL;
    static constexpr uint64_t ONE[6] = {0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL,
                                        0x77ce585370525745ULL, 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL};
    static constexpr uint64_t R2[6] = {0xf4df1f341c341746ULL, 0x0a76e6a609d104f1ULL, 0x8de5476c4c95b6d5ULL,
                                       0x67eb88a9939d83c0ULL, 0x9a793e85b519952dULL, 0x11988fe592cae3aaULL};
};

template <int N>
struct Fp {
    uint64_t v[N];
    typedef FpParams<N> PP;

    static Fp zero() {
        Fp r;
        for (int i = 0; i < N; i++) r.v[i] = 0;
        return r;
    }
    static Fp one() {
        Fp r;
        for (int i = 0; i < N; i++) r.v[i] = PP::ONE[i];
        return r;
    }
    static Fp r2() {
        Fp r;
        for (int i = 0; i < N; i++) r.v[i] = PP::R2[i];
        return r;
    }
    bool is_zero() const {
        uint64_t o = 0;
        for (int i = 0; i < N; i++) o |= v[i];
        return o == 0;
    }
    bool operator==(const Fp& b) const {
        uint64_t o = 0;
        for (int i = 0; i < N; i++) o |= v[i] ^ b.v[i];
        return o == 0;
    }
    bool operator!=(const Fp& b) const { return !(*this == b); }

    static inline bool geq_mod(const uint64_t* a) {
        for (int i = N - 1; i >= 0; i--) {
            if (a[i] > PP::MOD[i]) return true;
            if (a[i] < PP::MOD[i]) return false;
        }
        return true;
    }
    static inline void sub_mod(uint64_t* a) {
        u128 br = 0;
        for (int i = 0; i < N; i++) {
            u128 d = (u128)a[i] - PP::MOD[i] - br;
            a[i] = (uint64_t)d;
            br = (d >> 64) & 1;
        }
    }
    Fp operator+(const Fp& b) const {
        Fp r;
        u128 c = 0;
        for (int i = 0; i < N; i++) {
            c += (u128)v[i] + b.v[i];
            r.v[i] = (uint64_t)c;
            c >>= 64;
        }
        if (geq_mod(r.v)) sub_mod(r.v);  // p < 2^(64N-1): no carry out
        return r;
    }
    Fp operator-(const Fp& b) const {
        Fp r;
        u128 br = 0;
        for (int i = 0; i < N; i++) {
            u128 d = (u128)v[i] - b.v[i] - br;
            r.v[i] = (uint64_t)d;
            br = (d >> 64) & 1;
        }
        if (br) {
            u128 c = 0;
            for (int i = 0; i < N; i++) {
                c += (u128)r.v[i] + PP::MOD[i];
                r.v[i] = (uint64_t)c;
                c >>= 64;
            }
        }
        return r;
    }
    Fp operator-() const { return zero() - *this; }
    Fp dbl() const { return *this + *this; }
    /* Montgomery product a b R^-1 (CIOS); valid whenever a * b < p R (operands need not be canonical), result canonical */
    Fp operator*(const Fp& b) const {
        uint64_t t[N + 2];
#pragma GCC unroll 8
        for (int i = 0; i < N + 2; i++) t[i] = 0;
#pragma GCC unroll 8
        for (int i = 0; i < N; i++) {
            u128 c = 0;
#pragma GCC unroll 8
            for (int j = 0; j < N; j++) {
                c += (u128)v[j] * b.v[i] + t[j];
                t[j] = (uint64_t)c;
                c >>= 64;
            }
            c += t[N];
            t[N] = (uint64_t)c;
            t[N + 1] = (uint64_t)(c >> 64);
            uint64_t m = t[0] * PP::INV;
            c = (u128)m * PP::MOD[0] + t[0];
            c >>= 64;
#pragma GCC unroll 8
            for (int j = 1; j < N; j++) {
                c += (u128)m * PP::MOD[j] + t[j];
                t[j - 1] = (uint64_t)c;
                c >>= 64;
            }
            c += t[N];
            t[N - 1] = (uint64_t)c;
            t[N] = t[N + 1] + (uint64_t)(c >> 64);
        }
        Fp r;
#pragma GCC unroll 8
        for (int i = 0; i < N; i++) r.v[i] = t[i];
        if (t[N] || geq_mod(r.v)) sub_mod(r.v);
        return r;
    }
    Fp& operator+=(const Fp& b) { return *this = *this + b; }
    Fp& operator-=(const Fp& b) { return *this = *this - b; }
    Fp& operator*=(const Fp& b) { return *this = *this * b; }
    Fp sqr() const { return *this * *this; }

    static Fp from_u64(uint64_t x) {
        Fp r = zero();
        r.v[0] = x;
        return r * r2();
    }
    /* raw N-limb integer (any value < 2^(64N)) -> Montgomery form of (x mod p) */
    static Fp from_raw(const uint64_t* limbs) {
        Fp r;
        for (int i = 0; i < N; i++) r.v[i] = limbs[i];
        return r * r2();  // x * R^2 * R^-1; x R2 < 2^(64N) p = p R
    }
    /* canonical (non-Montgomery) value */
    void to_raw(uint64_t* out) const {
        Fp o = zero();
        o.v[0] = 1;
        Fp r = *this * o;
        for (int i = 0; i < N; i++) out[i] = r.v[i];
    }
    Fp pow(const uint64_t* e, int n_limbs) const {
        Fp acc = one();
        for (int i = n_limbs - 1; i >= 0; i--)
            for (int b = 63; b >= 0; b--) {
                acc = acc.sqr();
                if ((e[i] >> b) & 1) acc = acc * *this;
            }
        return acc;
    }
    /* x^(p-2); 0 -> 0 (callers assert non-zero where the reference unwraps) */
    Fp inverse() const {
        uint64_t e[N];
        for (int i = 0; i < N; i++) e[i] = PP::MOD[i];
        e[0] -= 2;  // p is odd and p[0] >= 2
        return pow(e, N);
    }
};

typedef Fp<4> Fr;
typedef Fp<6> Fq;

/* ark_ff::PrimeField::from_le_bytes_mod_order for up to 64 bytes (transcript challenges, proof_transcript.rs:33-41) */
static inline Fr fr_from_le_bytes_mod_order(const uint8_t* b, size_t n) {
    uint64_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
    for (size_t i = 0; i < n && i < 32; i++) lo[i / 8] |= (uint64_t)b[i] << (8 * (i % 8));
    for (size_t i = 32; i < n && i < 64; i++) hi[(i - 32) / 8] |= (uint64_t)b[i] << (8 * ((i - 32) % 8));
    Fr l = Fr::from_raw(lo);
    if (n <= 32) return l;
    Fr h = Fr::from_raw(hi);
    return l + h * Fr::r2();  // h * 2^256: Montgomery form of 2^256 is R * R = R2
}
/* ark-serialize compressed Fr: 32 bytes little-endian of the canonical value (proof_transcript.rs:52-57) */
static inline void fr_serialize(const Fr& x, uint8_t* out) {
    uint64_t r[4];
    x.to_raw(r);
    memcpy(out, r, 32);
}

/* Montgomery's simultaneous inversion (ark_ff::batch_inversion); zeros stay zero */
template <int N>
static inline void batch_inverse(Fp<N>* a, size_t n) {
    std::vector<Fp<N>> pre(n);
    Fp<N> acc = Fp<N>::one();
    for (size_t i = 0; i < n; i++) {
        pre[i] = acc;
        if (!a[i].is_zero()) acc = acc * a[i];
    }
    Fp<N> inv = acc.inverse();
    for (size_t i = n; i-- > 0;) {
        if (a[i].is_zero()) continue;
        Fp<N> t = inv * pre[i];
        inv = inv * a[i];
        a[i] = t;
    }
}
}  // namespace po
