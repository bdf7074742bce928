// Multi-scalar multiplication over BLS12-381 G1 (the multilinear-commitment MSM).
//   KzgProvingKey::commit / open     src/commitments/kzg.rs:123-133  (-> liblasso::msm::VariableBaseMSM::msm)
//   msm_nonaff                       src/msm_nonaffine.rs:34-38       (projective bases)
// The MSM value is a unique group element, so any algorithm that returns its affine coordinates in canonical
// Montgomery form is bit-exact with the reference.  Bucket method, c-bit unsigned windows:
//   1. digits:      scalars leave Montgomery form (one multiplication by the raw integer 1), every (point, window)
//                   digit is histogrammed;
//   2. scatter:     per-window counting sort of the point indices by digit (offsets from an exclusive scan);
//   3. accumulate:  one thread per (window, bucket) sums its points with XYZZ mixed additions (8M + 2S, no inversion);
//   4. reduce:      per window sum_d d*B_d with the running-sum trick split over the threads of one block;
//   5. combine:     Horner over the windows (c doublings each), one inversion to return to affine.
// Point arithmetic is integer-pipe (IMAD) bound: a mixed add is ~10 Fq multiplications of 12x12 32-bit limbs.
#include <algorithm>
#include "common.cuh"
#include "host_g1.hpp"
#include <thread>

// q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
__device__ __constant__ uint32_t FQ_P[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                             0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};

__device__ __forceinline__ Fq fq_one() {  // R mod q
    Fq r;
    r.l[0] = 0x0002fffdu; r.l[1] = 0x76090000u; r.l[2] = 0xc40c0002u; r.l[3] = 0xebf4000bu; r.l[4] = 0x53c758bau; r.l[5] = 0x5f489857u;
    r.l[6] = 0x70525745u; r.l[7] = 0x77ce5853u; r.l[8] = 0xa256ec6du; r.l[9] = 0x5c071a97u; r.l[10] = 0xfa80e493u; r.l[11] = 0x15f65ec3u;
    return r;
}
__device__ __forceinline__ Fq fq_dbl(const Fq& a) { return fq_add(a, a); }
__device__ __forceinline__ bool fq_eq(const Fq& a, const Fq& b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) o |= a.l[i] ^ b.l[i];
    return o == 0;
}

struct G1Aff {  // (0, 0) encodes the point at infinity (not on y^2 = x^3 + 4)
    Fq x, y;
};
struct G1X {  // extended Jacobian: x = X/ZZ, y = Y/ZZZ, ZZ^3 == ZZZ^2; ZZ == 0 is infinity
    Fq X, Y, ZZ, ZZZ;
};

__device__ __forceinline__ G1X g1x_inf() {
    G1X r;
    r.X = fq_zero(); r.Y = fq_zero(); r.ZZ = fq_zero(); r.ZZZ = fq_zero();
    return r;
}
__device__ __forceinline__ bool g1x_is_inf(const G1X& p) { return fq_is_zero(p.ZZ); }
__device__ __forceinline__ bool g1a_is_inf(const G1Aff& p) { return fq_is_zero(p.x) && fq_is_zero(p.y); }

// dbl-2008-s-1 (a = 0)
__device__ __noinline__ G1X g1x_dbl(const G1X& p) {
    if (g1x_is_inf(p) || fq_is_zero(p.Y)) return g1x_inf();
    Fq U = fq_dbl(p.Y);
    Fq V = fq_sqr(U);
    Fq W = fq_mul(U, V);
    Fq S = fq_mul(p.X, V);
    Fq XX = fq_sqr(p.X);
    Fq M = fq_add(fq_dbl(XX), XX);
    G1X r;
    r.X = fq_sub(fq_sqr(M), fq_dbl(S));
    r.Y = fq_sub(fq_mul(M, fq_sub(S, r.X)), fq_mul(W, p.Y));
    r.ZZ = fq_mul(V, p.ZZ);
    r.ZZZ = fq_mul(W, p.ZZZ);
    return r;
}

// madd-2008-s: XYZZ += affine.  The _i variants are inlined into the bucket-accumulation loops (the accumulator stays in
// registers); the plain ones are out-of-line calls for the cold paths.
__device__ __forceinline__ void g1x_madd_i(G1X& a, const G1Aff& b) {
    if (g1a_is_inf(b)) return;
    if (g1x_is_inf(a)) {
        a.X = b.x; a.Y = b.y; a.ZZ = fq_one(); a.ZZZ = fq_one();
        return;
    }
    Fq U2 = fq_mul(b.x, a.ZZ);
    Fq S2 = fq_mul(b.y, a.ZZZ);
    Fq Pp = fq_sub(U2, a.X);
    Fq R = fq_sub(S2, a.Y);
    if (fq_is_zero(Pp)) {
        if (fq_is_zero(R)) {  // same point: double the affine operand
            G1X t;
            t.X = b.x; t.Y = b.y; t.ZZ = fq_one(); t.ZZZ = fq_one();
            a = g1x_dbl(t);
        } else {
            a = g1x_inf();
        }
        return;
    }
    Fq PP = fq_sqr(Pp);
    Fq PPP = fq_mul(Pp, PP);
    Fq Q = fq_mul(a.X, PP);
    Fq X3 = fq_sub(fq_sub(fq_sqr(R), PPP), fq_dbl(Q));
    Fq Y3 = fq_sub(fq_mul(R, fq_sub(Q, X3)), fq_mul(a.Y, PPP));
    a.X = X3;
    a.Y = Y3;
    a.ZZ = fq_mul(a.ZZ, PP);
    a.ZZZ = fq_mul(a.ZZZ, PPP);
}

__device__ __noinline__ void g1x_madd(G1X& a, const G1Aff& b) { g1x_madd_i(a, b); }

// add-2008-s: XYZZ += XYZZ
__device__ __forceinline__ void g1x_add_i(G1X& a, const G1X& b) {
    if (g1x_is_inf(b)) return;
    if (g1x_is_inf(a)) { a = b; return; }
    Fq U1 = fq_mul(a.X, b.ZZ);
    Fq U2 = fq_mul(b.X, a.ZZ);
    Fq S1 = fq_mul(a.Y, b.ZZZ);
    Fq S2 = fq_mul(b.Y, a.ZZZ);
    Fq Pp = fq_sub(U2, U1);
    Fq R = fq_sub(S2, S1);
    if (fq_is_zero(Pp)) {
        if (fq_is_zero(R)) a = g1x_dbl(a); else a = g1x_inf();
        return;
    }
    Fq PP = fq_sqr(Pp);
    Fq PPP = fq_mul(Pp, PP);
    Fq Q = fq_mul(U1, PP);
    Fq X3 = fq_sub(fq_sub(fq_sqr(R), PPP), fq_dbl(Q));
    Fq Y3 = fq_sub(fq_mul(R, fq_sub(Q, X3)), fq_mul(S1, PPP));
    a.X = X3;
    a.Y = Y3;
    a.ZZ = fq_mul(fq_mul(a.ZZ, b.ZZ), PP);
    a.ZZZ = fq_mul(fq_mul(a.ZZZ, b.ZZZ), PPP);
}

__device__ __noinline__ void g1x_add(G1X& a, const G1X& b) { g1x_add_i(a, b); }

// a^(q-2)
__device__ Fq fq_inv(const Fq& a) {
    Fq r = fq_one();
    uint32_t e[12];
    for (int i = 0; i < 12; i++) e[i] = FQ_P[i];
    e[0] -= 2;  // low limb 0xffffaaab, no borrow
    for (int i = 380; i >= 0; i--) {
        r = fq_sqr(r);
        if ((e[i >> 5] >> (i & 31)) & 1) r = fq_mul(r, a);
    }
    return r;
}

// ---- kernels ------------------------------------------------------------------------------------------
// Window digits.  A digit entry is MSM_SKIP (nothing to add) or  bucket | sign << 31.
//   unsigned (cb == c):   bucket = digit in [1, 2^c)
//   signed   (cb == c-1): v = raw digit + carry; v > 2^cb is recoded as v - 2^c with a carry into the next window, so the
//                         magnitudes are 0 .. 2^cb and a window needs HALF the buckets.  bucket = magnitude mod 2^cb: bucket
//                         0 holds the magnitude 2^cb (the reduction gives it that weight, msm_window_sums `wrap`), and the
//                         negative digits add the NEGATED base.  The window count covers 256 bits, so the last carry is zero.
#define MSM_SKIP 0xffffffffu
#define MSM_NEG 0x80000000u
__device__ __forceinline__ uint32_t msm_raw_digit(const Fr& s, int w, int c) {
    const int bit = w * c;
    if (bit >= 256) return 0;
    const int limb = bit >> 5, sh = bit & 31;
    uint64_t v = s.l[limb];
    if (limb + 1 < 8) v |= (uint64_t)s.l[limb + 1] << 32;
    return (uint32_t)(v >> sh) & ((1u << c) - 1);
}
__device__ __forceinline__ uint32_t msm_recode(uint32_t raw, int c, int cb, uint32_t& carry) {
    if (cb == c) return raw ? raw : MSM_SKIP;
    const uint32_t half = 1u << cb, v = raw + carry;
    if (v > half) {
        carry = 1;
        const uint32_t m = (1u << c) - v;  // 0 .. half - 1
        return m ? (m | MSM_NEG) : MSM_SKIP;
    }
    carry = 0;
    return v ? (v & (half - 1)) : MSM_SKIP;  // v == half -> bucket 0
}
// scalars: Fr in Montgomery form -> plain integers; per (window, bucket) histogram
__global__ void msm_digits_kernel(const Fr* scalars, uint32_t n, int c, int cb, int n_windows, uint32_t* digits /* [W][n] */,
                                  uint32_t* counts /* [W][2^cb] */) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Fr one_raw = fr_zero();
        one_raw.l[0] = 1;
        Fr s = fr_mul(scalars[i], one_raw);  // s * R^-1: the canonical integer
        uint32_t carry = 0;
        for (int w = 0; w < n_windows; w++) {
            const uint32_t e = msm_recode(msm_raw_digit(s, w, c), c, cb, carry);
            digits[(size_t)w * n + i] = e;
            if (e != MSM_SKIP) atomicAdd(&counts[((size_t)w << cb) + (e & ~MSM_NEG)], 1u);
        }
    }
}

// the same for up to MSM_MULTI_MAX scalar tables over the SAME bases (gkr_msm_g1_multi): problem j owns windows [j W, (j+1) W)
#define MSM_MULTI_MAX 8
struct MsmMultiScalars {
    const Fr* s[MSM_MULTI_MAX];
    uint32_t n[MSM_MULTI_MAX];
    uint32_t k;
};
__global__ void msm_digits_multi_kernel(const __grid_constant__ MsmMultiScalars S, uint32_t n_max, int c, int cb, int n_windows,
                                        uint32_t* digits /* [k W][n_max] */, uint32_t* counts /* [k W][2^cb] */) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_max; i += gridDim.x * blockDim.x) {
        for (uint32_t j = 0; j < S.k; j++) {
            const size_t w0 = (size_t)j * n_windows;
            if (i >= S.n[j]) {  // beyond this problem's length: never accumulated
                for (int w = 0; w < n_windows; w++) digits[(w0 + w) * n_max + i] = MSM_SKIP;
                continue;
            }
            Fr one_raw = fr_zero();
            one_raw.l[0] = 1;
            Fr s = fr_mul(S.s[j][i], one_raw);
            uint32_t carry = 0;
            for (int w = 0; w < n_windows; w++) {
                const uint32_t e = msm_recode(msm_raw_digit(s, w, c), c, cb, carry);
                digits[(w0 + w) * n_max + i] = e;
                if (e != MSM_SKIP) atomicAdd(&counts[((w0 + w) << cb) + (e & ~MSM_NEG)], 1u);
            }
        }
    }
}

// exclusive scan of every window's histogram (one block per window)
__global__ void msm_scan_kernel(const uint32_t* counts, uint32_t* offsets, int c) {
    __shared__ uint32_t carry;
    __shared__ uint32_t tmp[1024];
    const uint32_t nb = 1u << c;
    const uint32_t* cnt = counts + ((size_t)blockIdx.x << c);
    uint32_t* off = offsets + ((size_t)blockIdx.x << c);
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += blockDim.x) {
        uint32_t v = base + threadIdx.x < nb ? cnt[base + threadIdx.x] : 0;
        tmp[threadIdx.x] = v;
        __syncthreads();
        for (uint32_t s = 1; s < blockDim.x; s <<= 1) {
            uint32_t t = threadIdx.x >= s ? tmp[threadIdx.x - s] : 0;
            __syncthreads();
            tmp[threadIdx.x] += t;
            __syncthreads();
        }
        if (base + threadIdx.x < nb) off[base + threadIdx.x] = carry + tmp[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry += tmp[threadIdx.x];
        __syncthreads();
    }
}

// exclusive scan of ONE long histogram (the 2^c buckets of the fixed-base path) in 1024-entry tiles:
// tile-local scan + tile totals, msm_scan_kernel over the totals, then the tile bases are added
__global__ void msm_scan_tiles_kernel(const uint32_t* counts, uint32_t* offsets, uint32_t* tile_sums) {
    __shared__ uint32_t tmp[1024];
    const size_t i = (size_t)blockIdx.x * 1024 + threadIdx.x;
    const uint32_t v = counts[i];
    tmp[threadIdx.x] = v;
    __syncthreads();
    for (uint32_t s = 1; s < 1024; s <<= 1) {
        uint32_t t = threadIdx.x >= s ? tmp[threadIdx.x - s] : 0;
        __syncthreads();
        tmp[threadIdx.x] += t;
        __syncthreads();
    }
    offsets[i] = tmp[threadIdx.x] - v;
    if (threadIdx.x == 1023) tile_sums[blockIdx.x] = tmp[1023];
}
__global__ void msm_scan_add_kernel(uint32_t* offsets, const uint32_t* tile_offs) {
    offsets[(size_t)blockIdx.x * 1024 + threadIdx.x] += tile_offs[blockIdx.x];
}

// sorted entries: point index | sign << 31 (the accumulation negates the base of a negative digit)
__global__ void msm_scatter_kernel(const uint32_t* digits, uint32_t n, int cb, int n_windows, const uint32_t* offsets, uint32_t* cursor,
                                   uint32_t* sorted /* [W][n] */) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        for (int w = 0; w < n_windows; w++) {
            const uint32_t e = digits[(size_t)w * n + i];
            if (e == MSM_SKIP) continue;
            const size_t b = ((size_t)w << cb) + (e & ~MSM_NEG);
            const uint32_t pos = atomicAdd(&cursor[b], 1u);
            sorted[(size_t)w * n + offsets[b] + pos] = i | (e & MSM_NEG);
        }
    }
}

// ---- balanced bucket accumulation ----------------------------------------------------------------------------
// Bucket sizes are data dependent (a degenerate top window, small-integer scalar tables, digit buckets of 2^(x-d) points):
// one thread per bucket would serialise thousands of additions.  Buckets are therefore counting-sorted by size
// (descending) into three tiers: HUGE (>= MSM_HUGE entries) are summed by a whole block each, MEDIUM (>= cap) by one
// warp each (strided partial sums + shared-memory tree), LIGHT (< cap) by one thread each -- and because neighbouring
// threads own buckets of equal size, warps of the light tier do not diverge.  `cap` is chosen by the host from the
// average bucket size and the number of buckets (enough threads to fill the machine either way).
#define MSM_MAXCAP 1024
#define MSM_HUGE 4096
#define MSM_NBINS (MSM_MAXCAP + 2)  // bins 0..cap-1: exact light sizes; bin cap: medium; bin cap+1: huge

__device__ __forceinline__ uint32_t msm_bin_of(uint32_t cnt, uint32_t cap) { return cnt < cap ? cnt : (cnt < MSM_HUGE ? cap : cap + 1); }

__global__ void msm_bin_hist_kernel(const uint32_t* counts, uint64_t nbk, uint32_t cap, uint32_t* bins /* [MSM_NBINS] */) {
    __shared__ uint32_t sh[MSM_NBINS];
    for (int i = threadIdx.x; i < MSM_NBINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nbk; b += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(&sh[msm_bin_of(counts[b], cap)], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < MSM_NBINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&bins[i], sh[i]);
}
// descending exclusive offsets: huge first, then medium, then light by decreasing size, empty buckets last
__global__ void msm_bin_scan_kernel(const uint32_t* bins, uint32_t cap, uint32_t* bin_off) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t acc = 0;
    for (int k = (int)cap + 1; k >= 0; k--) {
        bin_off[k] = acc;
        acc += bins[k];
    }
}
__global__ void __launch_bounds__(256) msm_bin_scatter_kernel(const uint32_t* counts, uint64_t nbk, uint32_t cap, const uint32_t* bin_off,
                                                               uint32_t* bin_cursor, uint32_t* order) {
    __shared__ uint32_t sh_cnt[MSM_NBINS], sh_base[MSM_NBINS];
    const uint64_t tiles = (nbk + blockDim.x - 1) / blockDim.x;
    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (int i = threadIdx.x; i < MSM_NBINS; i += blockDim.x) sh_cnt[i] = 0;
        __syncthreads();
        const uint64_t b = tile * blockDim.x + threadIdx.x;
        uint32_t k = 0, pos = 0;
        if (b < nbk) {
            k = msm_bin_of(counts[b], cap);
            pos = atomicAdd(&sh_cnt[k], 1u);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < MSM_NBINS; i += blockDim.x)
            if (sh_cnt[i]) sh_base[i] = bin_off[i] + atomicAdd(&bin_cursor[i], sh_cnt[i]);
        __syncthreads();
        if (b < nbk) order[sh_base[k] + pos] = (uint32_t)b;
        __syncthreads();
    }
}

// KIND 0: affine bases, 1: Jacobian (X, Y, Z), 2: extended Jacobian (X, Y, ZZ, ZZZ) as produced by gkr_g1_bucket_sums
// `e` = base index | sign << 31 (sorted entries of the signed-digit MSM; the bucket-sum callers never set the sign);
// shift: base offset of the problem (gkr_msm_g1_batch)
template <int KIND>
__device__ __forceinline__ void msm_add_base(G1X& acc, const void* bases, uint32_t e, uint32_t shift = 0) {
    const uint32_t i = (e & ~MSM_NEG) + shift;
    const bool neg = (e & MSM_NEG) != 0;
    if (KIND == 2) {
        G1X t = ((const G1X*)bases)[i];
        if (neg) t.Y = fq_sub(fq_zero(), t.Y);
        g1x_add_i(acc, t);
    } else if (KIND == 1) {
        // Jacobian (X, Y, Z) base: ZZ = Z^2, ZZZ = Z^3
        const Fq* p = (const Fq*)bases + (size_t)3 * i;
        Fq Z = p[2];
        if (fq_is_zero(Z)) return;
        G1X t;
        t.X = p[0]; t.Y = neg ? fq_sub(fq_zero(), p[1]) : p[1]; t.ZZ = fq_sqr(Z); t.ZZZ = fq_mul(t.ZZ, Z);
        g1x_add_i(acc, t);
    } else {
        G1Aff p = ((const G1Aff*)bases)[i];
        if (neg) p.y = fq_sub(fq_zero(), p.y);  // (0, 0) = infinity stays (0, 0)
        g1x_madd_i(acc, p);
    }
}

struct MsmAccArgs {
    const void* bases;
    const uint32_t *sorted, *counts, *offsets, *order, *bins;
    uint32_t n, cap;
    int c;
    uint64_t total;  // buckets per problem
    G1X* buckets;
    uint32_t n_problems, problem_stride;  // independent MSMs sharing the scalars: bases shifted by p * problem_stride
};

// light tier (< cap entries): one thread per bucket, in size order
// MINB: resident blocks per SM the register allocation is capped for (2: 176 registers; 3: 168, 12 warps per SM)
template <int KIND, int MINB>
__global__ void __launch_bounds__(128, MINB) msm_accumulate_light_kernel(const __grid_constant__ MsmAccArgs A) {
    const uint64_t n_heavy = (uint64_t)A.bins[A.cap] + A.bins[A.cap + 1];
    const uint64_t n_light = A.total - n_heavy;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_light * A.n_problems; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t p = (uint32_t)(i / n_light);
        const uint32_t b = A.order[n_heavy + i % n_light];
        const uint32_t cnt = A.counts[b];
        const uint32_t shift = p * A.problem_stride;
        G1X acc = g1x_inf();
        const uint32_t* idx = A.sorted + (size_t)(b >> A.c) * A.n + A.offsets[b];
        for (uint32_t k = 0; k < cnt; k++) msm_add_base<KIND>(acc, A.bases, idx[k], shift);
        A.buckets[(size_t)p * A.total + b] = acc;
    }
}
// tree-sum the G partial sums of a group of G consecutive threads (G = 32: one warp, G = blockDim: the block)
__device__ __forceinline__ void msm_group_tree(G1X* sh, uint32_t lane, uint32_t G, bool whole_block) {
    for (uint32_t s = G >> 1; s > 0; s >>= 1) {
        if (lane < s) {
            G1X a = sh[lane];
            g1x_add(a, sh[lane + s]);
            sh[lane] = a;
        }
        if (whole_block) __syncthreads(); else __syncwarp();
    }
}
// medium tier ([cap, MSM_HUGE) entries): one warp per bucket
template <int KIND>
__global__ void __launch_bounds__(128) msm_accumulate_medium_kernel(const __grid_constant__ MsmAccArgs A) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    G1X* sh = reinterpret_cast<G1X*>(smem_raw) + warp * 32;
    const uint32_t n_huge = A.bins[A.cap + 1], n_med = A.bins[A.cap];
    for (uint64_t h = (uint64_t)blockIdx.x * wpb + warp; h < (uint64_t)n_med * A.n_problems; h += (uint64_t)gridDim.x * wpb) {
        const uint32_t p = (uint32_t)(h / n_med);
        const uint32_t b = A.order[n_huge + h % n_med];
        const uint32_t cnt = A.counts[b];
        const uint32_t shift = p * A.problem_stride;
        const uint32_t* idx = A.sorted + (size_t)(b >> A.c) * A.n + A.offsets[b];
        G1X acc = g1x_inf();
        for (uint32_t k = lane; k < cnt; k += 32) msm_add_base<KIND>(acc, A.bases, idx[k], shift);
        sh[lane] = acc;
        __syncwarp();
        msm_group_tree(sh, lane, 32, false);
        if (lane == 0) A.buckets[(size_t)p * A.total + b] = sh[0];
        __syncwarp();
    }
}
// huge tier: one block per bucket
template <int KIND>
__global__ void __launch_bounds__(256) msm_accumulate_huge_kernel(const __grid_constant__ MsmAccArgs A) {
    extern __shared__ unsigned char smem_raw[];
    G1X* sh = reinterpret_cast<G1X*>(smem_raw);
    const uint32_t n_huge = A.bins[A.cap + 1];
    for (uint64_t h = blockIdx.x; h < (uint64_t)n_huge * A.n_problems; h += gridDim.x) {
        const uint32_t p = (uint32_t)(h / n_huge);
        const uint32_t b = A.order[h % n_huge];
        const uint32_t cnt = A.counts[b];
        const uint32_t shift = p * A.problem_stride;
        const uint32_t* idx = A.sorted + (size_t)(b >> A.c) * A.n + A.offsets[b];
        G1X acc = g1x_inf();
        for (uint32_t k = threadIdx.x; k < cnt; k += blockDim.x) msm_add_base<KIND>(acc, A.bases, idx[k], shift);
        sh[threadIdx.x] = acc;
        __syncthreads();
        msm_group_tree(sh, threadIdx.x, blockDim.x, true);
        if (threadIdx.x == 0) A.buckets[(size_t)p * A.total + b] = sh[0];
        __syncthreads();
    }
}

// per window: sum_d d * B_d.  A thread owns the segment [lo, lo + L) of one window; descending running sums give
// tot = sum (d - lo + 1) B_d and run = sum B_d, so the segment contributes tot + (lo - 1) * run.
// 2^k * p by k doublings
__device__ __forceinline__ G1X g1x_mul_pow2(G1X p, int k) {
    for (int j = 0; j < k; j++) p = g1x_dbl(p);
    return p;
}
// wrap: bucket 0 of every window weighs 2^c instead of 0 (the magnitude 2^c of the signed-digit recoding, msm_recode)
__global__ void __launch_bounds__(128) msm_segment_kernel(const G1X* buckets, int c, int seg_log, uint64_t n_threads, G1X* seg_out,
                                                          uint32_t out_stride /* entries per window in seg_out */, bool wrap) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_threads) return;
    const uint32_t segs = 1u << (c - seg_log);
    const uint32_t w = (uint32_t)(t / segs), sgm = (uint32_t)(t % segs);
    const G1X* B = buckets + ((size_t)w << c);
    const uint32_t lo = sgm << seg_log, hi = lo + (1u << seg_log);
    G1X run = g1x_inf(), tot = g1x_inf();
    for (uint32_t d = hi; d-- > lo;) {
        g1x_add(run, B[d]);
        g1x_add(tot, run);
    }
    if (lo >= 2) {  // tot += (lo - 1) * run
        uint32_t k = lo - 1;
        G1X acc = g1x_inf(), base = run;
        while (k) {
            if (k & 1) g1x_add(acc, base);
            k >>= 1;
            if (k) base = g1x_dbl(base);
        }
        g1x_add(tot, acc);
    } else if (lo == 0) {
        // tot counted (d + 1) * B_d for the segment starting at zero: remove one run
        G1X neg = run;
        neg.Y = fq_sub(fq_zero(), neg.Y);
        g1x_add(tot, neg);
        if (wrap) g1x_add(tot, g1x_mul_pow2(B[0], c));
    }
    seg_out[(size_t)w * out_stride + sgm] = tot;
}
// Two-level variant for large bucket arrays (the scalar multiplication by lo - 1 above is 3/5 of a thread's work at c = 16):
// level 0 leaves tot_s = sum_{d in segment} (d - lo) B_d and hands run8_s = 2^seg_log * run_s to level 1, because
//   sum_d d B_d = sum_s tot_s + sum_s s * (2^seg_log run_s),
// the second sum being the same weighted sum over an array 2^seg_log times shorter (msm_segment_kernel on it).
// Per bucket: (2 L + seg_log) / L additions here + 5 / L above, instead of 5.
__global__ void __launch_bounds__(128) msm_segment_level0_kernel(const G1X* buckets, int c, int seg_log, uint64_t n_threads, G1X* tot_out,
                                                                 uint32_t tot_stride, G1X* run_out, bool wrap) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_threads) return;
    const uint32_t segs = 1u << (c - seg_log);
    const uint32_t w = (uint32_t)(t / segs), sgm = (uint32_t)(t % segs);
    const G1X* B = buckets + ((size_t)w << c);
    const uint32_t lo = sgm << seg_log, hi = lo + (1u << seg_log);
    G1X run = g1x_inf(), tot = g1x_inf();
    for (uint32_t d = hi; d-- > lo + 1;) {
        g1x_add(run, B[d]);
        g1x_add(tot, run);
    }
    g1x_add(run, B[lo]);  // weight zero inside the segment
    for (int k = 0; k < seg_log; k++) run = g1x_dbl(run);
    if (wrap && lo == 0) g1x_add(tot, g1x_mul_pow2(B[0], c));  // run_0 has weight zero in level 1, so B[0] counts only here
    tot_out[(size_t)w * tot_stride + sgm] = tot;
    run_out[t] = run;
}

// Tiny windows (c == 4: the commitments-times-coefficients combinations of a proof, a handful of points): one WARP per
// window instead of a 40-addition chain in one thread.  sum_d d B_d = sum_j 2^j (sum of the 8 buckets with bit j set):
// lane (j, p) fetches the p-th such bucket, three shuffle levels give the four bit sums, lane 0 combines them with
// three doublings -- a dependent chain of 9 group operations.
__global__ void __launch_bounds__(128) msm_window_bits_kernel(const G1X* buckets, uint32_t n_windows, G1X* window_sums) {
    const uint32_t lane = threadIdx.x & 31, w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= n_windows) return;
    const uint32_t j = lane >> 3, p = lane & 7;
    // p-th 4-bit value with bit j set: insert a one at position j into the 3-bit number p
    const uint32_t low = p & ((1u << j) - 1u), d = ((p >> j) << (j + 1)) | (1u << j) | low;
    G1X v = buckets[((size_t)w << 4) + d];
    for (int off = 4; off > 0; off >>= 1) {
        G1X o;
#pragma unroll
        for (int k = 0; k < 12; k++) {
            o.X.l[k] = __shfl_down_sync(0xffffffffu, v.X.l[k], off);
            o.Y.l[k] = __shfl_down_sync(0xffffffffu, v.Y.l[k], off);
            o.ZZ.l[k] = __shfl_down_sync(0xffffffffu, v.ZZ.l[k], off);
            o.ZZZ.l[k] = __shfl_down_sync(0xffffffffu, v.ZZZ.l[k], off);
        }
        if (p < (uint32_t)off) g1x_add(v, o);
    }
    // lanes 0, 8, 16, 24 hold the bit sums b0..b3
    G1X acc = g1x_inf();
    for (int bit = 3; bit >= 0; bit--) {
        G1X b;
#pragma unroll
        for (int k = 0; k < 12; k++) {
            b.X.l[k] = __shfl_sync(0xffffffffu, v.X.l[k], bit * 8);
            b.Y.l[k] = __shfl_sync(0xffffffffu, v.Y.l[k], bit * 8);
            b.ZZ.l[k] = __shfl_sync(0xffffffffu, v.ZZ.l[k], bit * 8);
            b.ZZZ.l[k] = __shfl_sync(0xffffffffu, v.ZZZ.l[k], bit * 8);
        }
        if (lane == 0) {
            if (bit != 3) acc = g1x_dbl(acc);
            g1x_add(acc, b);
        }
    }
    if (lane == 0) window_sums[w] = acc;
}

// ---- fixed-base windows (gkr_srs_precompute) ----------------------------------------------------------------------------
// The SRS of a proving key is fixed, so the window shifts can be paid once: with T[k][i] = 2^(c k) P_i resident, digit k of
// scalar i selects T[k][i] and ALL windows fall into ONE set of 2^c buckets,
//     sum_i s_i P_i = sum_d d * (sum of the T[k][i] with digit_k(s_i) == d).
// That removes the per-window bucket sets (c can grow to 20: 13 instead of 16 additions per 255-bit scalar), the per-window
// running-sum reductions and the Horner tail over the windows.
__global__ void __launch_bounds__(128) srs_precompute_kernel(const G1Aff* P, uint64_t n, int c, int n_windows, G1Aff* T) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const G1Aff p = P[i];
        T[i] = p;
        if (g1a_is_inf(p)) {
            for (int k = 1; k < n_windows; k++) T[(size_t)k * n + i] = p;
            continue;
        }
        G1X cur;
        cur.X = p.x; cur.Y = p.y; cur.ZZ = fq_one(); cur.ZZZ = fq_one();
        for (int k = 1; k < n_windows; k++) {
            for (int j = 0; j < c; j++) cur = g1x_dbl(cur);
            G1Aff r;
            if (g1x_is_inf(cur)) {
                r.x = fq_zero(); r.y = fq_zero();
            } else {
                // one inversion: 1/ZZZ, then 1/ZZ = ZZZ^-1 * ZZZ / ZZ ... use 1/ZZ = (ZZ * ZZZ^-1)^2 (ZZ^3 == ZZZ^2)
                const Fq izzz = fq_inv(cur.ZZZ);
                const Fq t = fq_mul(cur.ZZ, izzz);  // = 1 / sqrt(ZZ) up to the curve relation: (ZZ / ZZZ)^2 = 1 / ZZ
                r.x = fq_mul(cur.X, fq_sqr(t));
                r.y = fq_mul(cur.Y, izzz);
            }
            T[(size_t)k * n + i] = r;
        }
    }
}

// digits of all windows into one histogram; entry e = k * n + i
__global__ void msm_digits_pre_kernel(const Fr* scalars, uint32_t n, int c, int cb, int n_windows, uint32_t* digits /* [W][n] */,
                                      uint32_t* counts /* [2^cb] */) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Fr one_raw = fr_zero();
        one_raw.l[0] = 1;
        Fr s = fr_mul(scalars[i], one_raw);  // the canonical integer
        uint32_t carry = 0;
        for (int w = 0; w < n_windows; w++) {
            const uint32_t e = msm_recode(msm_raw_digit(s, w, c), c, cb, carry);
            digits[(size_t)w * n + i] = e;
            if (e != MSM_SKIP) atomicAdd(&counts[e & ~MSM_NEG], 1u);
        }
    }
}
// sorted[offsets[d] + pos] = index of T[k][first + i] in the table of `srs_n` points per window
__global__ void msm_scatter_pre_kernel(const uint32_t* digits, uint32_t n, int n_windows, uint32_t first, uint32_t srs_n, const uint32_t* offsets,
                                       uint32_t* cursor, uint32_t* sorted) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        for (int w = 0; w < n_windows; w++) {
            const uint32_t e = digits[(size_t)w * n + i];
            if (e == MSM_SKIP) continue;
            const uint32_t d = e & ~MSM_NEG;
            const uint32_t pos = atomicAdd(&cursor[d], 1u);
            sorted[offsets[d] + pos] = ((uint32_t)w * srs_n + first + i) | (e & MSM_NEG);
        }
    }
}

// window_sums[w] = sum of the window's segment contributions (strided partial sums + shared-memory tree)
__global__ void __launch_bounds__(256) msm_window_tree_kernel(const G1X* seg_out, uint32_t segs, G1X* window_sums) {
    extern __shared__ unsigned char smem_raw[];
    G1X* sh = reinterpret_cast<G1X*>(smem_raw);
    const G1X* S = seg_out + (size_t)blockIdx.x * segs;
    G1X acc = g1x_inf();
    for (uint32_t k = threadIdx.x; k < segs; k += blockDim.x) g1x_add(acc, S[k]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (uint32_t s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            G1X a = sh[threadIdx.x];
            g1x_add(a, sh[threadIdx.x + s]);
            sh[threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) window_sums[blockIdx.x] = sh[0];
}

// ---- host -------------------------------------------------------------------------------------------------
struct gkr_srs {
    gkr_ctx* ctx = nullptr;
    void* d = nullptr;
    uint64_t n = 0;
    int kind = 0;  // 0: affine (x, y) 2x6 u64; 1: Jacobian (X, Y, Z) 3x6 u64; 2: XYZZ 4x6 u64 (device-produced bucket sums)
    // gkr_srs_precompute: pre[k * n + i] = 2^(pre_c * k) * P_i for k = 0 .. pre_w - 1 (affine; k = 0 is a copy of d)
    G1Aff* pre = nullptr;
    int pre_c = 0, pre_w = 0;
    size_t stride() const { return kind == 0 ? sizeof(G1Aff) : (kind == 1 ? 3 * sizeof(Fq) : sizeof(G1X)); }
};

extern "C" int gkr_srs_upload(gkr_ctx* ctx, const uint64_t* points, uint64_t n, int projective, gkr_srs** out) {
    if (!ctx) return GKR_ERR_ARG;
    if (!out || (!points && n)) return ctx->fail(GKR_ERR_ARG, "null argument");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    gkr_srs* s = new gkr_srs();
    s->ctx = ctx;
    s->n = n;
    s->kind = projective ? 1 : 0;
    size_t bytes = (size_t)n * s->stride();
    cudaError_t e = gkr_malloc_async(&s->d, std::max<size_t>(bytes, 16), ctx->stream);
    if (e == cudaSuccess && bytes && bytes <= ((size_t)1 << 20)) {
        if (gkr_stage_upload(ctx, s->d, points, bytes)) e = cudaErrorUnknown;  // a handful of points (commitment combinations): pinned ring
    } else {
        if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(s->d, points, bytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
    if (e != cudaSuccess) {
        delete s;
        return ctx->fail(GKR_ERR_CUDA, cudaGetErrorString(e));
    }
    *out = s;
    return GKR_OK;
}

extern "C" uint64_t gkr_srs_len(const gkr_srs* s) { return s ? s->n : 0; }

extern "C" void gkr_srs_free(gkr_srs* s) {
    if (!s) return;
    if (s->pre) gkr_free_async(s->pre, s->ctx->stream);
    if (s->d) gkr_free_async(s->d, s->ctx->stream);
    delete s;
}

// window width: minimise W * (n + 3 * 2^c) bucket additions (accumulate + running-sum reduce) over c, and avoid a
// degenerate top window (255 - (W - 1) c < 7 bits would put n / 2^bits points into each of a handful of buckets)
// Signed digits (msm_recode) halve the buckets of a window at the same number of additions; they need c >= 8 and the windows
// to cover 256 bits (the last carry).
struct MsmWindow {
    int c, cb, W;  // window bits, bucket bits (c - 1 when signed), windows
    bool is_signed() const { return cb != c; }
};
// mode (gkr_ctx::msm_signed): 0 never, 1 from 2^15 points (below that an MSM is bound by the latency of its passes and the
// signed recoding measured slower: the batched 2^10-point commitments of second_phase 7.9 -> 13 ms), 2 whenever c >= 8 (tests)
static MsmWindow pick_window(uint64_t n, int mode) {
    const bool allow_signed = mode >= 2 || (mode == 1 && n >= ((uint64_t)1 << 15));
    int lg = 0;
    while (((uint64_t)1 << lg) < n) lg++;
    MsmWindow best = {4, 4, 64};
    double best_cost = 1e300;
    for (int c = 4; c <= 18; c++) {
        if (c > lg + 1 && c > 4) break;
        const bool sg = allow_signed && c >= 8;
        const int W = sg ? (256 + c - 1) / c : (255 + c - 1) / c;
        const int cb = sg ? c - 1 : c;
        const int top_bits = 255 - (W - 1) * c;  // scalar bits in the top window (a signed top window also takes a carry)
        double cost = (double)W * ((double)n + 3.0 * (double)((uint64_t)1 << cb));
        if (top_bits < 7 && n > 4096) cost += 4.0 * (double)n;
        if (cost < best_cost) {
            best_cost = cost;
            best = {c, cb, W};
        }
    }
    return best;
}

// size-ordered bucket accumulation: counts/offsets [nbk], sorted [W][n] -> buckets [n_problems][nbk].  `work` holds
// order [nbk] followed by bins / bin_off / bin_cursor [3 * MSM_NBINS].  n_entries: total sorted entries (for the average).
static int msm_accumulate(gkr_ctx* ctx, const void* bases, int kind, const uint32_t* sorted, const uint32_t* counts, const uint32_t* offsets,
                          uint32_t n, int c, uint64_t nbk, uint64_t n_entries, uint32_t* work, G1X* buckets, uint32_t n_problems = 1,
                          uint32_t problem_stride = 0) {
    cudaStream_t st = ctx->stream;
    uint32_t* order = work;
    uint32_t* bins = work + nbk;
    uint32_t* bin_off = bins + MSM_NBINS;
    uint32_t* bin_cursor = bin_off + MSM_NBINS;
    // light/medium threshold: with few buckets every bucket needs many threads; with many buckets one thread per bucket
    // already fills the machine and only outliers (>= 4x the average) are worth a warp
    uint32_t cap = 32;
    if (nbk * n_problems >= 65536) {
        uint64_t avg = n_entries / nbk;
        cap = (uint32_t)std::min<uint64_t>(MSM_MAXCAP, std::max<uint64_t>(32, 4 * avg));
    }
    GKR_CUDA_OK(ctx, cudaMemsetAsync(bins, 0, sizeof(uint32_t) * 3 * MSM_NBINS, st));
    unsigned gb = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((nbk + 255) / 256, (uint64_t)ctx->num_sms * 4));
    msm_bin_hist_kernel<<<gb, 256, 0, st>>>(counts, nbk, cap, bins);
    msm_bin_scan_kernel<<<1, 32, 0, st>>>(bins, cap, bin_off);
    msm_bin_scatter_kernel<<<gb, 256, 0, st>>>(counts, nbk, cap, bin_off, bin_cursor, order);
    MsmAccArgs A;
    A.bases = bases; A.sorted = sorted; A.counts = counts; A.offsets = offsets; A.order = order; A.bins = bins;
    A.n = n; A.cap = cap; A.c = c; A.total = nbk; A.buckets = buckets; A.n_problems = n_problems; A.problem_stride = problem_stride;
    unsigned gl = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((nbk * n_problems + 127) / 128, (uint64_t)ctx->num_sms * 16));
    unsigned gm = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((nbk * n_problems + 3) / 4, (uint64_t)ctx->num_sms * 8));
    unsigned gh = (unsigned)ctx->num_sms * 2;
    const size_t sh_huge = sizeof(G1X) * 256, sh_med = sizeof(G1X) * 128;
#define GKR_MSM_ACC(K)                                                   \
    msm_accumulate_huge_kernel<K><<<gh, 256, sh_huge, st>>>(A);          \
    msm_accumulate_medium_kernel<K><<<gm, 128, sh_med, st>>>(A);         \
    if (K == 0 && ctx->msm_light_minb >= 3) msm_accumulate_light_kernel<K, 3><<<gl, 128, 0, st>>>(A);   \
    else msm_accumulate_light_kernel<K, 2><<<gl, 128, 0, st>>>(A);  /* projective bases need 222 registers: no cap */
    if (kind == 2) { GKR_MSM_ACC(2) } else if (kind == 1) { GKR_MSM_ACC(1) } else { GKR_MSM_ACC(0) }
#undef GKR_MSM_ACC
    ctx->launches += 6;
    GKR_CUDA_OK(ctx, cudaGetLastError());
    return GKR_OK;
}

// window sums S_w = sum_d d * buckets[w][d] for W consecutive groups of 2^c buckets: segment running sums and the
// per-group tree on the device, result (extended Jacobian) copied to the host.
// wrap: bucket 0 of every group weighs 2^c (signed-digit windows).
static int msm_window_sums(gkr_ctx* ctx, const G1X* buckets, int c, uint32_t W, std::vector<gkr::G1XH>& h, bool wrap = false) {
    cudaStream_t st = ctx->stream;
    G1X* seg_out = nullptr;
    G1X* wsums = nullptr;
    if (c == 4 && !wrap) {  // tiny windows: one warp per window
        GKR_CUDA_OK(ctx, gkr_malloc_async(&seg_out, sizeof(G1X) * W, st));
        wsums = seg_out;
        msm_window_bits_kernel<<<(W + 3) / 4, 128, 0, st>>>(buckets, W, wsums);
        ctx->launches++;
    } else {
        const int seg_log = c < 3 ? c : 3;
        const uint32_t segs = 1u << (c - seg_log);
        const uint64_t n_threads = (uint64_t)W * segs;
        if (((uint64_t)W << c) >= ((uint64_t)1 << 20) && c >= 9) {  // measured: faster from 2^20 buckets up (MSMs of >= 2^20 points), slower below
            // two levels: running sums per segment, then the weighted sum of the (pre-scaled) segment totals
            const int c1 = c - seg_log, seg_log1 = 3;
            const uint32_t segs1 = 1u << (c1 - seg_log1), stride = segs + segs1;
            const uint64_t n_threads1 = (uint64_t)W * segs1;
            GKR_CUDA_OK(ctx, gkr_malloc_async(&seg_out, sizeof(G1X) * ((uint64_t)W * stride + n_threads + W + 9 * (uint64_t)W), st));
            G1X* runs = seg_out + (uint64_t)W * stride;
            wsums = runs + n_threads;
            G1X* parts = wsums + W;
            msm_segment_level0_kernel<<<(unsigned)((n_threads + 127) / 128), 128, 0, st>>>(buckets, c, seg_log, n_threads, seg_out, stride, runs, wrap);
            msm_segment_kernel<<<(unsigned)((n_threads1 + 127) / 128), 128, 0, st>>>(runs, c1, seg_log1, n_threads1, seg_out + segs, stride, false);
            // stride = 9 * segs1 entries per window: nine blocks per window sum segs1 entries each, then one block the nine partials
            msm_window_tree_kernel<<<9 * W, 256, sizeof(G1X) * 256, st>>>(seg_out, segs1, parts);
            msm_window_tree_kernel<<<W, 256, sizeof(G1X) * 256, st>>>(parts, 9, wsums);
            ctx->launches += 4;
        } else {
            GKR_CUDA_OK(ctx, gkr_malloc_async(&seg_out, sizeof(G1X) * (n_threads + W), st));
            wsums = seg_out + n_threads;
            msm_segment_kernel<<<(unsigned)((n_threads + 127) / 128), 128, 0, st>>>(buckets, c, seg_log, n_threads, seg_out, segs, wrap);
            msm_window_tree_kernel<<<W, 256, sizeof(G1X) * 256, st>>>(seg_out, segs, wsums);
            ctx->launches += 2;
        }
    }
    GKR_CUDA_OK(ctx, cudaGetLastError());
    h.resize(W);
    GKR_CUDA_OK(ctx, cudaMemcpyAsync(h.data(), wsums, sizeof(G1X) * W, cudaMemcpyDeviceToHost, st));
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(st));
    gkr_free_async(seg_out, st);
    return GKR_OK;
}

// host tail for `n_problems` MSMs of W windows each (window sums back to back): Horner + inversion per problem
// (host_g1.hpp), one thread per problem when there are several.
static void msm_host_tail(const std::vector<gkr::G1XH>& h, int c, int W, uint32_t n_problems, uint64_t* out_xy) {
    if (n_problems == 1) {
        gkr::g1h::horner_windows(h.data(), c, W, out_xy);
        return;
    }
    unsigned hw = std::thread::hardware_concurrency();
    unsigned nt = std::max(1u, std::min(hw ? hw : 4u, n_problems));
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
        th.emplace_back([&, t]() {
            for (uint32_t p = t; p < n_problems; p += nt) gkr::g1h::horner_windows(h.data() + (size_t)p * W, c, W, out_xy + 12 * (size_t)p);
        });
    for (auto& t : th) t.join();
}

int gkr_msm_team_run(gkr_ctx* ctx, const gkr_srs* srs, uint64_t first, const Fr* d_scalars, uint64_t n, uint64_t* out_xy);  // msm_team.cu

// Fixed-base window table of an affine SRS (proving-key preprocessing, not part of a proof): (n_windows - 1) * n more points.
extern "C" int gkr_srs_precompute(gkr_ctx* ctx, gkr_srs* srs, int c) {
    if (!ctx) return GKR_ERR_ARG;
    if (!srs || srs->kind != 0 || c < 12 || c > 22 || srs->n == 0) return ctx->fail(GKR_ERR_ARG, "gkr_srs_precompute: affine SRS and 12 <= c <= 22");
    const int W = (255 + c - 1) / c;
    if ((uint64_t)W * srs->n >= ((uint64_t)1 << 31)) return ctx->fail(GKR_ERR_UNSUPPORTED, "gkr_srs_precompute: table index does not fit 31 bits");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    if (srs->pre) gkr_free_async(srs->pre, ctx->stream);
    srs->pre = nullptr;
    GKR_CUDA_OK(ctx, gkr_malloc_async(&srs->pre, sizeof(G1Aff) * (size_t)W * srs->n, ctx->stream));
    unsigned g = (unsigned)std::min<uint64_t>((srs->n + 127) / 128, (uint64_t)ctx->num_sms * 16);
    srs_precompute_kernel<<<g, 128, 0, ctx->stream>>>((const G1Aff*)srs->d, srs->n, c, W, srs->pre);
    ctx->launches++;
    GKR_CUDA_OK(ctx, cudaGetLastError());
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    srs->pre_c = c;
    srs->pre_w = W;
    return GKR_OK;
}

// One bucket set for all windows over the fixed-base table.  The 2^c buckets are reduced as V = 2^(c - 14) chunks of 2^14:
// chunk j contributes S_j = sum_{d in chunk} (d - j 2^14) B_d (the device window sums with c' = 14) plus j 2^14 R_j with
// R_j = sum_{d in chunk} B_d; the V pairs go to the host, which finishes with a running sum over the R_j.
static int msm_g1_pre(gkr_ctx* ctx, const gkr_srs* srs, uint64_t first, const Fr* d_scalars, uint64_t n, uint64_t* out_xy) {
    cudaStream_t st = ctx->stream;
    // signed digits (msm_recode) when the windows cover 256 bits: half the buckets to reduce
    const int c = srs->pre_c, W = srs->pre_w;
    const bool sg = ctx->msm_signed != 0 && W * c >= 256;
    const int cb = sg ? c - 1 : c, cc = cb < 14 ? cb : 14;
    const size_t nbk = (size_t)1 << cb;
    const uint32_t V = 1u << (cb - cc);
    uint32_t *digits = nullptr, *sorted = nullptr, *counts = nullptr;
    G1X* buckets = nullptr;
    GKR_CUDA_OK(ctx, gkr_malloc_async(&digits, sizeof(uint32_t) * W * n, st));
    GKR_CUDA_OK(ctx, gkr_malloc_async(&sorted, sizeof(uint32_t) * W * n, st));
    GKR_CUDA_OK(ctx, gkr_malloc_async(&counts, sizeof(uint32_t) * (nbk * 4 + 3 * MSM_NBINS), st));
    uint32_t *offsets = counts + nbk, *cursor = counts + 2 * nbk, *work = counts + 3 * nbk;
    const uint32_t parts = (uint32_t)(nbk >> 10) ? (uint32_t)(nbk >> 10) : 1;  // partial sums of 1024 buckets each
    GKR_CUDA_OK(ctx, gkr_malloc_async(&buckets, sizeof(G1X) * (nbk + V + parts), st));
    G1X* chunk_sums = buckets + nbk;
    G1X* part_sums = chunk_sums + V;
    GKR_CUDA_OK(ctx, cudaMemsetAsync(counts, 0, sizeof(uint32_t) * nbk * 3, st));
    unsigned g1 = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->num_sms * 8);
    msm_digits_pre_kernel<<<g1, 256, 0, st>>>(d_scalars, (uint32_t)n, c, cb, W, digits, counts);
    if (cb >= 12) {  // tiled scan: `cursor` (still zero) doubles as scratch for the tile totals / bases and is cleared again
        const uint32_t tiles = 1u << (cb - 10);
        uint32_t *tile_sums = cursor, *tile_offs = cursor + tiles;
        msm_scan_tiles_kernel<<<tiles, 1024, 0, st>>>(counts, offsets, tile_sums);
        msm_scan_kernel<<<1, 1024, 0, st>>>(tile_sums, tile_offs, cb - 10);
        msm_scan_add_kernel<<<tiles, 1024, 0, st>>>(offsets, tile_offs);
        GKR_CUDA_OK(ctx, cudaMemsetAsync(cursor, 0, sizeof(uint32_t) * 2 * tiles, st));
        ctx->launches += 2;
    } else {
        msm_scan_kernel<<<1, 1024, 0, st>>>(counts, offsets, cb);
    }
    msm_scatter_pre_kernel<<<g1, 256, 0, st>>>(digits, (uint32_t)n, W, (uint32_t)first, (uint32_t)srs->n, offsets, cursor, sorted);
    ctx->launches += 3;
    std::vector<gkr::G1XH> hs, hr(V);
    gkr::G1XH b0;  // bucket 0 = the magnitude 2^cb of the signed recoding
    int rc = msm_accumulate(ctx, srs->pre, 0, sorted, counts, offsets, (uint32_t)((uint64_t)W * n), cb, nbk, (uint64_t)W * n, work, buckets);
    if (rc == GKR_OK) {
        if (nbk >= 1024 && parts >= V) {  // plain chunk sums R_j in two stages, so that the first one fills the machine
            msm_window_tree_kernel<<<parts, 256, sizeof(G1X) * 256, st>>>(buckets, 1024, part_sums);
            msm_window_tree_kernel<<<V, 256, sizeof(G1X) * 256, st>>>(part_sums, parts / V, chunk_sums);
            ctx->launches += 2;
        } else {
            msm_window_tree_kernel<<<V, 256, sizeof(G1X) * 256, st>>>(buckets, 1u << cc, chunk_sums);
            ctx->launches++;
        }
        rc = msm_window_sums(ctx, buckets, cc, V, hs);  // synchronises
    }
    if (rc == GKR_OK) {
        cudaError_t e = cudaMemcpyAsync(hr.data(), chunk_sums, sizeof(G1X) * V, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&b0, buckets, sizeof(G1X), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = ctx->fail(GKR_ERR_CUDA, cudaGetErrorString(e));
    }
    gkr_free_async(digits, st);
    gkr_free_async(sorted, st);
    gkr_free_async(counts, st);
    gkr_free_async(buckets, st);
    if (rc) return rc;
    using namespace gkr::g1h;
    gkr::G1XH total = inf(), run = inf(), wsum = inf();
    for (uint32_t j = V; j-- > 1;) {  // wsum = sum_j j R_j by descending running sums
        run = add(run, hr[j]);
        wsum = add(wsum, run);
    }
    for (int k = 0; k < cc; k++) wsum = dbl(wsum);
    for (uint32_t j = 0; j < V; j++) total = add(total, hs[j]);
    total = add(total, wsum);
    if (sg) {
        for (int k = 0; k < cb; k++) b0 = dbl(b0);
        total = add(total, b0);
    }
    to_affine(total, out_xy);
    return GKR_OK;
}

static int msm_g1_impl(gkr_ctx* ctx, const gkr_srs* srs, uint64_t first, uint64_t problem_stride, uint32_t n_problems, const Fr* d_scalars,
                       uint64_t n, uint64_t* out_xy, bool allow_team = true) {
    if (!ctx) return GKR_ERR_ARG;
    if (!srs || !d_scalars || !out_xy || n_problems == 0) return ctx->fail(GKR_ERR_ARG, "null argument");
    if (first + (uint64_t)(n_problems - 1) * problem_stride + n > srs->n) return ctx->fail(GKR_ERR_ARG, "Vector is too large.");  // kzg.rs:124
    // MSM split by point range over the GPUs of the box (msm_team.cu): affine SRS bases, one problem, large enough to pay
    if (allow_team && ctx->team && n_problems == 1 && srs->kind == 0 && n >= ctx->team_min_n) return gkr_msm_team_run(ctx, srs, first, d_scalars, n, out_xy);
    // fixed-base window table present and enough entries per bucket to pay for reducing 2^c buckets
    if (srs->pre && n_problems == 1 && n > 0 && (uint64_t)srs->pre_w * n >= ((uint64_t)4 << srs->pre_c)) {
        GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
        return msm_g1_pre(ctx, srs, first, d_scalars, n, out_xy);
    }
    if (srs->n >= ((uint64_t)1 << 31)) return ctx->fail(GKR_ERR_UNSUPPORTED, "MSM larger than 2^31 points");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (n == 0) {
        std::memset(out_xy, 0, 96 * (size_t)n_problems);
        return GKR_OK;
    }
    const MsmWindow mw = pick_window(n, ctx->msm_signed);
    const int c = mw.c, cb = mw.cb, W = mw.W;
    const size_t nbk = (size_t)W << cb;
    uint32_t *digits = nullptr, *sorted = nullptr, *counts = nullptr, *offsets = nullptr, *cursor = nullptr, *work = nullptr;
    G1X* buckets = nullptr;
    GKR_CUDA_OK(ctx, gkr_malloc_async(&digits, sizeof(uint32_t) * W * n, st));
    GKR_CUDA_OK(ctx, gkr_malloc_async(&sorted, sizeof(uint32_t) * W * n, st));
    GKR_CUDA_OK(ctx, gkr_malloc_async(&counts, sizeof(uint32_t) * (nbk * 4 + 3 * MSM_NBINS), st));
    offsets = counts + nbk;
    cursor = counts + 2 * nbk;
    work = counts + 3 * nbk;
    GKR_CUDA_OK(ctx, gkr_malloc_async(&buckets, sizeof(G1X) * nbk * n_problems, st));
    GKR_CUDA_OK(ctx, cudaMemsetAsync(counts, 0, sizeof(uint32_t) * nbk * 3, st));
    unsigned g1 = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->num_sms * 8);
    msm_digits_kernel<<<g1, 256, 0, st>>>(d_scalars, (uint32_t)n, c, cb, W, digits, counts);
    msm_scan_kernel<<<W, 1024, 0, st>>>(counts, offsets, cb);
    msm_scatter_kernel<<<g1, 256, 0, st>>>(digits, (uint32_t)n, cb, W, offsets, cursor, sorted);
    ctx->launches += 3;
    const void* bases = (const unsigned char*)srs->d + first * srs->stride();
    std::vector<gkr::G1XH> h;
    int rc = msm_accumulate(ctx, bases, srs->kind, sorted, counts, offsets, (uint32_t)n, cb, nbk, (uint64_t)W * n, work, buckets, n_problems,
                            (uint32_t)problem_stride);
    if (rc == GKR_OK) rc = msm_window_sums(ctx, buckets, cb, (uint32_t)W * n_problems, h, mw.is_signed());
    gkr_free_async(digits, st);
    gkr_free_async(sorted, st);
    gkr_free_async(counts, st);
    gkr_free_async(buckets, st);
    if (rc == GKR_OK) msm_host_tail(h, c, W, n_problems, out_xy);
    return rc;
}

// <bases[first .. first+n), scalars>   scalars: device table of n Fr (Montgomery).  out_xy: affine result, 12 u64
// (x then y, Montgomery form; all zero for the point at infinity).
extern "C" int gkr_msm_g1(gkr_ctx* ctx, const gkr_srs* srs, uint64_t first, const gkr_table* scalars, uint64_t n, uint64_t* out_xy) {
    if (!ctx) return GKR_ERR_ARG;
    if (!scalars) return ctx->fail(GKR_ERR_ARG, "null argument");
    if (scalars->n < n) return ctx->fail(GKR_ERR_ARG, "fewer scalars than requested");
    return msm_g1_impl(ctx, srs, first, 0, 1, scalars->d, n, out_xy);
}
// the local bucket MSM on this GPU over raw device scalars, never delegated to a team (used by both sides of msm_team.cu)
int gkr_msm_g1_local(gkr_ctx* ctx, const gkr_srs* srs, uint64_t first, const Fr* d_scalars, uint64_t n, uint64_t* out_xy) {
    return msm_g1_impl(ctx, srs, first, 0, 1, d_scalars, n, out_xy, false);
}
// `n_problems` MSMs with the SAME scalars over the base ranges [first + p * problem_stride, + n): the c_pull / d_pull
// commitments of PushForwardState::second_phase (one msm_nonaff per commitment chunk over that chunk's bucket sums with
// eq_c / eq_d as scalars, pushforward.rs:598-604).  The digit sort is shared; out_xy: n_problems x 12 u64.
extern "C" int gkr_msm_g1_batch(gkr_ctx* ctx, const gkr_srs* srs, uint64_t first, uint64_t problem_stride, uint32_t n_problems,
                                const gkr_table* scalars, uint64_t n, uint64_t* out_xy) {
    if (!ctx) return GKR_ERR_ARG;
    if (!scalars) return ctx->fail(GKR_ERR_ARG, "null argument");
    if (scalars->n < n) return ctx->fail(GKR_ERR_ARG, "fewer scalars than requested");
    return msm_g1_impl(ctx, srs, first, problem_stride, n_problems, scalars->d, n, out_xy);
}


// k commitments with DIFFERENT scalar tables over the SAME base range [first, first + n_j): the phase-1 commitments to p_0, p_1,
// ac_c, ac_d (pushforward.rs:534-537) and the two openings of KnucklesOpeningProtocol::prove that no challenge separates
// (opening.rs:77-83).  Below 2^19 points an MSM is bound by the latency of its reduction passes, not by additions, so the k
// problems share ONE digit sort, ONE accumulation and ONE reduction over k W windows; larger ones (or a team / fixed-base table)
// run one by one.  out_xy: k x 12 u64.
extern "C" int gkr_msm_g1_multi(gkr_ctx* ctx, const gkr_srs* srs, uint64_t first, const gkr_table* const* scalars, const uint64_t* n,
                                uint32_t k, uint64_t* out_xy) {
    if (!ctx) return GKR_ERR_ARG;
    if (!srs || !scalars || !n || !out_xy || k == 0) return ctx->fail(GKR_ERR_ARG, "null argument");
    uint64_t n_max = 0;
    for (uint32_t j = 0; j < k; j++) {
        if (!scalars[j] || scalars[j]->n < n[j]) return ctx->fail(GKR_ERR_ARG, "fewer scalars than requested");
        if (first + n[j] > srs->n) return ctx->fail(GKR_ERR_ARG, "Vector is too large.");  // kzg.rs:124
        n_max = std::max(n_max, n[j]);
    }
    if (k == 1 || k > MSM_MULTI_MAX || n_max == 0 || n_max >= ((uint64_t)1 << 19) || srs->kind != 0) {
        for (uint32_t j = 0; j < k; j++) {
            int rc = msm_g1_impl(ctx, srs, first, 0, 1, scalars[j]->d, n[j], out_xy + 12 * (size_t)j);
            if (rc) return rc;
        }
        return GKR_OK;
    }
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const MsmWindow mw = pick_window(n_max, ctx->msm_signed);
    const int c = mw.c, cb = mw.cb, W = mw.W;
    const uint32_t KW = k * (uint32_t)W;
    const size_t nbk = (size_t)KW << cb;
    uint32_t *digits = nullptr, *sorted = nullptr, *counts = nullptr;
    G1X* buckets = nullptr;
    GKR_CUDA_OK(ctx, gkr_malloc_async(&digits, sizeof(uint32_t) * KW * n_max, st));
    GKR_CUDA_OK(ctx, gkr_malloc_async(&sorted, sizeof(uint32_t) * KW * n_max, st));
    GKR_CUDA_OK(ctx, gkr_malloc_async(&counts, sizeof(uint32_t) * (nbk * 4 + 3 * MSM_NBINS), st));
    uint32_t *offsets = counts + nbk, *cursor = counts + 2 * nbk, *work = counts + 3 * nbk;
    GKR_CUDA_OK(ctx, gkr_malloc_async(&buckets, sizeof(G1X) * nbk, st));
    GKR_CUDA_OK(ctx, cudaMemsetAsync(counts, 0, sizeof(uint32_t) * nbk * 3, st));
    MsmMultiScalars S;
    S.k = k;
    uint64_t entries = 0;
    for (uint32_t j = 0; j < MSM_MULTI_MAX; j++) {
        S.s[j] = j < k ? scalars[j]->d : nullptr;
        S.n[j] = j < k ? (uint32_t)n[j] : 0;
        if (j < k) entries += (uint64_t)W * n[j];
    }
    unsigned g1 = (unsigned)std::min<uint64_t>((n_max + 255) / 256, (uint64_t)ctx->num_sms * 8);
    msm_digits_multi_kernel<<<g1, 256, 0, st>>>(S, (uint32_t)n_max, c, cb, W, digits, counts);
    msm_scan_kernel<<<KW, 1024, 0, st>>>(counts, offsets, cb);
    msm_scatter_kernel<<<g1, 256, 0, st>>>(digits, (uint32_t)n_max, cb, (int)KW, offsets, cursor, sorted);
    ctx->launches += 3;
    const void* bases = (const unsigned char*)srs->d + first * srs->stride();
    std::vector<gkr::G1XH> h;
    int rc = msm_accumulate(ctx, bases, 0, sorted, counts, offsets, (uint32_t)n_max, cb, nbk, entries, work, buckets);
    if (rc == GKR_OK) rc = msm_window_sums(ctx, buckets, cb, KW, h, mw.is_signed());
    gkr_free_async(digits, st);
    gkr_free_async(sorted, st);
    gkr_free_async(counts, st);
    gkr_free_async(buckets, st);
    if (rc == GKR_OK) msm_host_tail(h, c, W, k, out_xy);
    return rc;
}

// ---- bucket accumulation for the c / d commitments ----------------------------------------
//   PushForwardState::new       src/cleanup/protocols/pushforward/pushforward.rs:398-429, 433-456
//   Pullback::bucketed_msm      src/pullback.rs:28-59
// B[b] = sum over incidences k with bucket_idx[k] == b of bases[point_idx[k]].  Merging the rows of one commitment
// chunk (pushforward.rs:433-456) is implicit: all incidences of the chunk are accumulated into one bucket array.
__global__ void g1_hist_kernel(const uint32_t* bidx, uint32_t n, uint32_t n_buckets, uint32_t* counts, int* bad) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t b = bidx[i];
        if (b >= n_buckets) { *bad = 1; continue; }
        atomicAdd(&counts[b], 1u);
    }
}
__global__ void g1_scatter_kernel(const uint32_t* bidx, const uint32_t* pidx, uint32_t n, uint32_t n_buckets, const uint32_t* offsets,
                                  uint32_t* cursor, uint32_t* sorted) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t b = bidx[i];
        if (b >= n_buckets) continue;
        uint32_t pos = atomicAdd(&cursor[b], 1u);
        sorted[offsets[b] + pos] = pidx[i];
    }
}

extern "C" int gkr_g1_bucket_sums(gkr_ctx* ctx, const gkr_srs* srs, const uint32_t* point_idx, const uint32_t* bucket_idx, uint64_t n,
                                  uint32_t n_buckets, gkr_srs** out) {
    if (!ctx) return GKR_ERR_ARG;
    if (!srs || !out || (n && (!point_idx || !bucket_idx)) || n_buckets == 0) return ctx->fail(GKR_ERR_ARG, "null argument");
    if (srs->kind != 0) return ctx->fail(GKR_ERR_ARG, "bucket sums need affine bases");
    if (n >= ((uint64_t)1 << 31)) return ctx->fail(GKR_ERR_UNSUPPORTED, "too many incidences");
    for (uint64_t k = 0; k < n; k++)
        if (point_idx[k] >= srs->n) return ctx->fail(GKR_ERR_ARG, "point index out of range");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int c = 0;
    while ((1u << c) < n_buckets) c++;
    const size_t nbk = (size_t)1 << c;
    uint32_t *d_p = nullptr, *d_b = nullptr, *sorted = nullptr, *counts = nullptr;
    int* d_bad = nullptr;
    gkr_srs* res = new gkr_srs();
    res->ctx = ctx;
    res->n = n_buckets;
    res->kind = 2;
    GKR_CUDA_OK(ctx, gkr_malloc_async(&res->d, sizeof(G1X) * nbk, st));
    GKR_CUDA_OK(ctx, gkr_malloc_async(&d_p, sizeof(uint32_t) * std::max<uint64_t>(n, 1), st));
    GKR_CUDA_OK(ctx, gkr_malloc_async(&d_b, sizeof(uint32_t) * std::max<uint64_t>(n, 1), st));
    GKR_CUDA_OK(ctx, gkr_malloc_async(&sorted, sizeof(uint32_t) * std::max<uint64_t>(n, 1), st));
    GKR_CUDA_OK(ctx, gkr_malloc_async(&counts, sizeof(uint32_t) * (nbk * 4 + 3 * MSM_NBINS + 1), st));
    uint32_t* offsets = counts + nbk;
    uint32_t* cursor = counts + 2 * nbk;
    d_bad = (int*)(counts + 3 * nbk);
    uint32_t* work = counts + 3 * nbk + 1;
    GKR_CUDA_OK(ctx, cudaMemsetAsync(counts, 0, sizeof(uint32_t) * (nbk * 3 + 1), st));
    if (n) {
        GKR_CUDA_OK(ctx, cudaMemcpyAsync(d_p, point_idx, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
        GKR_CUDA_OK(ctx, cudaMemcpyAsync(d_b, bucket_idx, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
        unsigned g1 = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->num_sms * 8);
        g1_hist_kernel<<<g1, 256, 0, st>>>(d_b, (uint32_t)n, n_buckets, counts, d_bad);
        msm_scan_kernel<<<1, 1024, 0, st>>>(counts, offsets, c);
        g1_scatter_kernel<<<g1, 256, 0, st>>>(d_b, d_p, (uint32_t)n, n_buckets, offsets, cursor, sorted);
        ctx->launches += 3;
    }
    {
        int rc = msm_accumulate(ctx, srs->d, 0, sorted, counts, offsets, (uint32_t)n, c, nbk, n, work, (G1X*)res->d);
        if (rc) return rc;
    }
    int bad = 0;
    GKR_CUDA_OK(ctx, cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(st));  // host index arrays may be released by the caller now
    gkr_free_async(d_p, st);
    gkr_free_async(d_b, st);
    gkr_free_async(sorted, st);
    gkr_free_async(counts, st);
    if (bad) {
        gkr_srs_free(res);
        return ctx->fail(GKR_ERR_ARG, "bucket index out of range");
    }
    *out = res;
    return GKR_OK;
}

// ---- the same with the digit / counter matrix already on the device (PushForwardState::new) ------------------------------
__device__ __forceinline__ uint32_t rows_bucket(const uint32_t* idx, uint64_t k, uint32_t xl, uint32_t clm, uint32_t group_log, uint32_t* pidx) {
    const uint32_t y = (uint32_t)(k >> xl), x = (uint32_t)(k & (((uint64_t)1 << xl) - 1));
    *pidx = x + ((y & ((1u << clm) - 1)) << xl);
    return ((y >> clm) << group_log) | idx[k];
}
// x_lo / x_hi: only the points x_lo <= x < x_hi of every digit row (the x-range one GPU of a team owns, msm_team.cu)
__device__ __forceinline__ bool rows_in_range(uint64_t k, uint32_t xl, uint32_t x_lo, uint32_t x_hi) {
    const uint32_t x = (uint32_t)(k & (((uint64_t)1 << xl) - 1));
    return x >= x_lo && x < x_hi;
}
__global__ void g1_rows_hist_kernel(const uint32_t* idx, uint64_t n, uint32_t xl, uint32_t clm, uint32_t group_log, uint32_t* counts, int* bad,
                                    uint32_t x_lo, uint32_t x_hi) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
        if (!rows_in_range(k, xl, x_lo, x_hi)) continue;
        if (idx[k] >> group_log) { *bad = 1; continue; }
        uint32_t p;
        atomicAdd(&counts[rows_bucket(idx, k, xl, clm, group_log, &p)], 1u);
    }
}
__global__ void g1_rows_scatter_kernel(const uint32_t* idx, uint64_t n, uint32_t xl, uint32_t clm, uint32_t group_log, const uint32_t* offsets,
                                       uint32_t* cursor, uint32_t* sorted, uint32_t x_lo, uint32_t x_hi) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
        if (!rows_in_range(k, xl, x_lo, x_hi)) continue;
        if (idx[k] >> group_log) continue;
        uint32_t p;
        const uint32_t b = rows_bucket(idx, k, xl, clm, group_log, &p);
        sorted[offsets[b] + atomicAdd(&cursor[b], 1u)] = p;
    }
}
int gkr_team_bucket_sums(gkr_ctx* ctx, const gkr_srs* srs, const uint32_t* d_idx, uint64_t n, uint32_t x_logsize, uint32_t clm, uint32_t group_log,
                         gkr_srs** out, bool* handled);  // msm_team.cu
__global__ void g1x_accumulate_kernel(G1X* acc, const G1X* part, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        G1X a = acc[i];
        g1x_add(a, part[i]);
        acc[i] = a;
    }
}
// acc[i] += part[i] over n XYZZ bucket sums resident on the device (team leader: partial sums of the other GPUs)
int gkr_g1x_accumulate(gkr_ctx* ctx, void* d_acc, const void* d_part, uint64_t n) {
    unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + 127) / 128, (uint64_t)ctx->num_sms * 8));
    g1x_accumulate_kernel<<<grid, 128, 0, ctx->stream>>>((G1X*)d_acc, (const G1X*)d_part, n);
    ctx->launches++;
    GKR_CUDA_OK(ctx, cudaGetLastError());
    return GKR_OK;
}
size_t gkr_g1x_bytes() { return sizeof(G1X); }
void* gkr_srs_device_ptr(gkr_srs* s) { return s->d; }

int gkr_g1_bucket_sums_rows_range(gkr_ctx* ctx, const gkr_srs* srs, const uint32_t* d_idx, uint64_t n, uint32_t x_logsize, uint32_t clm,
                                  uint32_t group_log, uint32_t x_lo, uint32_t x_hi, bool allow_team, gkr_srs** out);

extern "C" int gkr_g1_bucket_sums_rows(gkr_ctx* ctx, const gkr_srs* srs, const gkr_u32buf* idx, uint32_t x_logsize, uint32_t clm, uint32_t group_log,
                                       gkr_srs** out) {
    if (!ctx) return GKR_ERR_ARG;
    if (!srs || !idx || !out || x_logsize > 30 || clm > 16 || group_log > 30) return ctx->fail(GKR_ERR_ARG, "bad argument");
    return gkr_g1_bucket_sums_rows_range(ctx, srs, idx->d, idx->n, x_logsize, clm, group_log, 0, 1u << x_logsize, true, out);
}

// The bucket accumulation of PushForwardState::new (pushforward.rs:398-429) restricted to the points x_lo <= x < x_hi of every
// digit row.  allow_team: when a team is attached (msm_team.cu) the x-range is split over its GPUs and the partial bucket
// sums are added here -- group addition commutes, so the sums are the same points.
int gkr_g1_bucket_sums_rows_range(gkr_ctx* ctx, const gkr_srs* srs, const uint32_t* d_idx, uint64_t n, uint32_t x_logsize, uint32_t clm,
                                  uint32_t group_log, uint32_t x_lo, uint32_t x_hi, bool allow_team, gkr_srs** out) {
    if (srs->kind != 0) return ctx->fail(GKR_ERR_ARG, "bucket sums need affine bases");
    const uint64_t x_size = (uint64_t)1 << x_logsize;
    if (n == 0 || n % x_size) return ctx->fail(GKR_ERR_ARG, "index matrix is not a whole number of rows");
    if (n >= ((uint64_t)1 << 31)) return ctx->fail(GKR_ERR_UNSUPPORTED, "too many incidences");
    const uint64_t rows = n / x_size;
    if (x_size * std::min<uint64_t>(rows, (uint64_t)1 << clm) > srs->n) return ctx->fail(GKR_ERR_ARG, "point index out of range");
    const uint64_t n_groups = (rows + ((uint64_t)1 << clm) - 1) >> clm;
    const uint64_t n_buckets = n_groups << group_log;
    if (n_buckets >= ((uint64_t)1 << 31)) return ctx->fail(GKR_ERR_UNSUPPORTED, "too many buckets");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    if (allow_team && ctx->team && x_lo == 0 && x_hi == x_size) {  // split the x-range over the GPUs of the team
        bool handled = false;
        gkr_srs* r = nullptr;
        int trc = gkr_team_bucket_sums(ctx, srs, d_idx, n, x_logsize, clm, group_log, &r, &handled);
        if (handled) {
            if (trc) return trc;
            *out = r;
            return GKR_OK;
        }
    }
    cudaStream_t st = ctx->stream;
    int c = 0;
    while (((uint64_t)1 << c) < n_buckets) c++;
    const size_t nbk = (size_t)1 << c;
    uint32_t *sorted = nullptr, *counts = nullptr;
    gkr_srs* res = new gkr_srs();
    res->ctx = ctx;
    res->n = n_buckets;
    res->kind = 2;
    GKR_CUDA_OK(ctx, gkr_malloc_async(&res->d, sizeof(G1X) * nbk, st));
    GKR_CUDA_OK(ctx, gkr_malloc_async(&sorted, sizeof(uint32_t) * n, st));
    GKR_CUDA_OK(ctx, gkr_malloc_async(&counts, sizeof(uint32_t) * (nbk * 4 + 3 * MSM_NBINS + 1), st));
    uint32_t* offsets = counts + nbk;
    uint32_t* cursor = counts + 2 * nbk;
    int* d_bad = (int*)(counts + 3 * nbk);
    uint32_t* work = counts + 3 * nbk + 1;
    GKR_CUDA_OK(ctx, cudaMemsetAsync(counts, 0, sizeof(uint32_t) * (nbk * 3 + 1), st));
    unsigned g1 = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->num_sms * 8);
    g1_rows_hist_kernel<<<g1, 256, 0, st>>>(d_idx, n, x_logsize, clm, group_log, counts, d_bad, x_lo, x_hi);
    msm_scan_kernel<<<1, 1024, 0, st>>>(counts, offsets, c);
    g1_rows_scatter_kernel<<<g1, 256, 0, st>>>(d_idx, n, x_logsize, clm, group_log, offsets, cursor, sorted, x_lo, x_hi);
    ctx->launches += 3;
    int rc = msm_accumulate(ctx, srs->d, 0, sorted, counts, offsets, (uint32_t)n, c, nbk, n, work, (G1X*)res->d);
    int bad = 0;
    if (rc == GKR_OK) {
        cudaError_t e = cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = ctx->fail(GKR_ERR_CUDA, cudaGetErrorString(e));
    }
    gkr_free_async(sorted, st);
    gkr_free_async(counts, st);
    if (rc == GKR_OK && bad) rc = ctx->fail(GKR_ERR_ARG, "bucket index out of range");
    if (rc) {
        gkr_srs_free(res);
        return rc;
    }
    *out = res;
    return GKR_OK;
}

// sum_{i=1}^{len-1} i * B[i]: the running-sum commitment of pushforward.rs:504-524 (== commit of the digit / counter table),
// for `count` consecutive groups of 2^group_log buckets starting at bucket `first` (one group per commitment chunk).
extern "C" int gkr_g1_weighted_bucket_sums(gkr_ctx* ctx, const gkr_srs* buckets, uint64_t first, uint32_t group_log, uint32_t count,
                                           uint64_t* out_xy) {
    if (!ctx) return GKR_ERR_ARG;
    if (!buckets || !out_xy || buckets->kind != 2 || count == 0 || group_log > 30) return ctx->fail(GKR_ERR_ARG, "expects bucket sums");
    int cap_log = 0;
    while (((uint64_t)1 << cap_log) < buckets->n) cap_log++;
    if (first + ((uint64_t)count << group_log) > ((uint64_t)1 << cap_log)) return ctx->fail(GKR_ERR_ARG, "bucket range out of bounds");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    std::vector<gkr::G1XH> h;
    // buckets beyond n (up to the power of two) were written as infinity by the accumulate kernels
    int rc = msm_window_sums(ctx, (const G1X*)buckets->d + first, (int)group_log, count, h);
    if (rc) return rc;
    for (uint32_t k = 0; k < count; k++) gkr::g1h::to_affine(h[k], out_xy + 12 * (size_t)k);
    return GKR_OK;
}
extern "C" int gkr_g1_weighted_bucket_sum(gkr_ctx* ctx, const gkr_srs* buckets, uint64_t* out_xy) {
    if (!ctx) return GKR_ERR_ARG;
    if (!buckets) return ctx->fail(GKR_ERR_ARG, "expects bucket sums");
    int c = 0;
    while (((uint64_t)1 << c) < buckets->n) c++;
    return gkr_g1_weighted_bucket_sums(ctx, buckets, 0, (uint32_t)c, 1, out_xy);
}

// download bucket sums / any point set as affine (x, y) pairs (tests, and `c_comm`-style per-bucket inspection)
__global__ void g1_to_affine_kernel(const G1X* in, uint64_t n, G1Aff* out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        G1X p = in[i];
        G1Aff r;
        if (g1x_is_inf(p)) { r.x = fq_zero(); r.y = fq_zero(); }
        else { r.x = fq_mul(p.X, fq_inv(p.ZZ)); r.y = fq_mul(p.Y, fq_inv(p.ZZZ)); }
        out[i] = r;
    }
}

extern "C" int gkr_g1_download_affine(gkr_ctx* ctx, const gkr_srs* pts, uint64_t* out_xy) {
    if (!ctx) return GKR_ERR_ARG;
    if (!pts || !out_xy) return ctx->fail(GKR_ERR_ARG, "null argument");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (pts->kind == 0) {
        GKR_CUDA_OK(ctx, cudaMemcpyAsync(out_xy, pts->d, sizeof(G1Aff) * pts->n, cudaMemcpyDeviceToHost, st));
        GKR_CUDA_OK(ctx, cudaStreamSynchronize(st));
        return GKR_OK;
    }
    if (pts->kind != 2) return ctx->fail(GKR_ERR_UNSUPPORTED, "download of Jacobian bases is not needed by the path");
    G1Aff* tmp = nullptr;
    GKR_CUDA_OK(ctx, gkr_malloc_async(&tmp, sizeof(G1Aff) * std::max<uint64_t>(pts->n, 1), st));
    unsigned g = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((pts->n + 63) / 64, (uint64_t)ctx->num_sms * 8));
    g1_to_affine_kernel<<<g, 64, 0, st>>>((const G1X*)pts->d, pts->n, tmp);
    ctx->launches++;
    GKR_CUDA_OK(ctx, cudaMemcpyAsync(out_xy, tmp, sizeof(G1Aff) * pts->n, cudaMemcpyDeviceToHost, st));
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(st));
    gkr_free_async(tmp, st);
    return GKR_OK;
}

// ---- KzgProvingKey::mock_setup  (src/commitments/kzg.rs:84-97) ------------------------------------------------
// ptau_1[i] = tau^i * g0.  Fixed-base: T[j] = 2^j g0 (255 affine points), then every thread adds the T[j] selected by the
// bits of tau^i with mixed additions and normalises its own point (one inversion: 1/ZZ = (ZZ / ZZZ)^2).
__global__ void srs_doublings_kernel(G1Aff g0, G1X* T) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    G1X p;
    p.X = g0.x; p.Y = g0.y; p.ZZ = fq_one(); p.ZZZ = fq_one();
    if (g1a_is_inf(g0)) p = g1x_inf();
    for (int j = 0; j < 255; j++) {
        T[j] = p;
        p = g1x_dbl(p);
    }
}
__global__ void __launch_bounds__(128) srs_fixed_base_kernel(const G1Aff* T, Fr tau, uint64_t n, G1Aff* out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        Fr s = fr_one(), b = tau;  // tau^i
        for (uint64_t e = i; e; e >>= 1) {
            if (e & 1) s = fr_mul(s, b);
            b = fr_sqr(b);
        }
        Fr one_raw = fr_zero();
        one_raw.l[0] = 1;
        s = fr_mul(s, one_raw);  // leave Montgomery form
        G1X acc = g1x_inf();
        for (int j = 0; j < 255; j++)
            if ((s.l[j >> 5] >> (j & 31)) & 1) g1x_madd(acc, T[j]);
        G1Aff r;
        if (g1x_is_inf(acc)) {
            r.x = fq_zero(); r.y = fq_zero();
        } else {
            Fq iz3 = fq_inv(acc.ZZZ);
            Fq iz2 = fq_sqr(fq_mul(acc.ZZ, iz3));
            r.x = fq_mul(acc.X, iz2);
            r.y = fq_mul(acc.Y, iz3);
        }
        out[i] = r;
    }
}

extern "C" int gkr_srs_mock_setup(gkr_ctx* ctx, const uint64_t tau[4], const uint64_t* g0_xy, uint64_t n, gkr_srs** out) {
    if (!ctx) return GKR_ERR_ARG;
    if (!tau || !g0_xy || !out) return ctx->fail(GKR_ERR_ARG, "null argument");
    gkr::FrH t = frh_from_limbs(tau);
    if (!frh_canonical(t)) return ctx->fail(GKR_ERR_ARG, "tau not canonical");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    gkr_srs* s = new gkr_srs();
    s->ctx = ctx;
    s->n = n;
    s->kind = 0;
    G1X* Tx = nullptr;
    G1Aff* Ta = nullptr;
    GKR_CUDA_OK(ctx, gkr_malloc_async(&s->d, std::max<size_t>(sizeof(G1Aff) * n, 16), st));
    GKR_CUDA_OK(ctx, gkr_malloc_async(&Tx, sizeof(G1X) * 255, st));
    GKR_CUDA_OK(ctx, gkr_malloc_async(&Ta, sizeof(G1Aff) * 255, st));
    G1Aff g0;
    std::memcpy(&g0, g0_xy, sizeof(G1Aff));
    srs_doublings_kernel<<<1, 32, 0, st>>>(g0, Tx);
    g1_to_affine_kernel<<<4, 64, 0, st>>>(Tx, 255, Ta);
    if (n) {
        unsigned g = (unsigned)std::min<uint64_t>((n + 127) / 128, (uint64_t)ctx->num_sms * 8);
        srs_fixed_base_kernel<<<g, 128, 0, st>>>(Ta, fr_from_host(t), n, (G1Aff*)s->d);
    }
    ctx->launches += 3;
    GKR_CUDA_OK(ctx, cudaGetLastError());
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(st));
    gkr_free_async(Tx, st);
    gkr_free_async(Ta, st);
    *out = s;
    return GKR_OK;
}

// test hook (no device needed): the host tail of an MSM -- Horner over `n_windows` extended-Jacobian window sums
// (24 u64 each: X, Y, ZZ, ZZZ) with c doublings per window, normalised to affine.
extern "C" int gkr_host_g1_horner(const uint64_t* window_sums, int c, int n_windows, uint64_t* out_xy) {
    if (!window_sums || !out_xy || c < 0 || n_windows < 0) return GKR_ERR_ARG;
    std::vector<gkr::G1XH> h(n_windows);
    if (n_windows) std::memcpy(h.data(), window_sums, sizeof(gkr::G1XH) * n_windows);
    gkr::g1h::horner_windows(h.data(), c, n_windows, out_xy);
    return GKR_OK;
}


// ---- binary_msm (old API): commitments to bit columns  (src/binary_msm.rs:19-54, gkr_msm_simple.rs:62-68) ---
// prepare_bases: every chunk of `gamma` consecutive bases becomes the table of its 2^gamma - 1 non-empty subset sums (affine),
//   entry i - 1 = sum of chunk[len - 1 - idx] over the set bits idx of i   (prepare_chunk zips 0..gamma with chunk.iter().rev());
// binary_msm: coefs[k] in [0, 2^gamma) selects entry coefs[k] - 1 of chunk k (0: nothing); the result is the plain sum.
__global__ void __launch_bounds__(128) binmsm_prepare_kernel(const G1Aff* bases, uint64_t n, uint32_t gamma, uint64_t n_chunks, G1Aff* out) {
    const uint32_t per = (1u << gamma) - 1;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_chunks * per; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = t / per;
        const uint32_t i = (uint32_t)(t % per) + 1;
        const uint64_t b0 = k * gamma;
        const uint32_t len = (uint32_t)(b0 + gamma <= n ? gamma : n - b0);
        G1X acc = g1x_inf();
        for (uint32_t idx = 0; idx < len; idx++)
            if ((i >> idx) & 1u) g1x_madd_i(acc, bases[b0 + len - 1 - idx]);
        G1Aff r;
        if (g1x_is_inf(acc)) {
            r.x = fq_zero(); r.y = fq_zero();
        } else {
            const Fq izzz = fq_inv(acc.ZZZ);
            const Fq u = fq_mul(acc.ZZ, izzz);  // (ZZ / ZZZ)^2 == 1 / ZZ
            r.x = fq_mul(acc.X, fq_sqr(u));
            r.y = fq_mul(acc.Y, izzz);
        }
        out[t] = r;
    }
}
// strided partial sums of the selected table entries, one partial per block
__global__ void __launch_bounds__(256) binmsm_sum_kernel(const G1Aff* table, uint32_t per, const uint8_t* coefs, uint64_t n_chunks, G1X* partial) {
    extern __shared__ unsigned char smem_raw[];
    G1X* sh = reinterpret_cast<G1X*>(smem_raw);
    G1X acc = g1x_inf();
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_chunks; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t c = coefs[k];
        if (c) g1x_madd_i(acc, table[k * per + (c - 1)]);
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    msm_group_tree(sh, threadIdx.x, blockDim.x, true);
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

extern "C" int gkr_binary_msm_prepare(gkr_ctx* ctx, const gkr_srs* bases, uint32_t gamma, gkr_srs** prepared) {
    if (!ctx) return GKR_ERR_ARG;
    if (!bases || !prepared || bases->kind != 0 || gamma == 0 || gamma > 8) return ctx->fail(GKR_ERR_ARG, "gkr_binary_msm_prepare: affine bases, 1 <= gamma <= 8");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const uint64_t n_chunks = (bases->n + gamma - 1) / gamma, per = ((uint64_t)1 << gamma) - 1;
    gkr_srs* s = new gkr_srs();
    s->ctx = ctx;
    s->kind = 0;
    s->n = n_chunks * per;
    cudaError_t e = gkr_malloc_async(&s->d, sizeof(G1Aff) * std::max<uint64_t>(s->n, 1), ctx->stream);
    if (e != cudaSuccess) { delete s; return ctx->fail(GKR_ERR_CUDA, cudaGetErrorString(e)); }
    if (s->n) {
        unsigned g = (unsigned)std::min<uint64_t>((s->n + 127) / 128, (uint64_t)ctx->num_sms * 16);
        binmsm_prepare_kernel<<<g, 128, 0, ctx->stream>>>((const G1Aff*)bases->d, bases->n, gamma, n_chunks, (G1Aff*)s->d);
        ctx->launches++;
    }
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { gkr_srs_free(s); return ctx->fail(GKR_ERR_CUDA, cudaGetErrorString(e)); }
    *prepared = s;
    return GKR_OK;
}

extern "C" int gkr_binary_msm(gkr_ctx* ctx, const gkr_srs* prepared, uint32_t gamma, const uint8_t* coefs, uint64_t n_chunks, uint64_t* out_xy) {
    if (!ctx) return GKR_ERR_ARG;
    if (!prepared || !out_xy || (!coefs && n_chunks) || prepared->kind != 0 || gamma == 0 || gamma > 8) return ctx->fail(GKR_ERR_ARG, "gkr_binary_msm: bad arguments");
    const uint32_t per = (1u << gamma) - 1;
    if (n_chunks * per != prepared->n) return ctx->fail(GKR_ERR_ARG, "gkr_binary_msm: coefs.len() != bases.len()");  // binary_msm.rs:21
    for (uint64_t k = 0; k < n_chunks; k++)
        if (coefs[k] > per) return ctx->fail(GKR_ERR_ARG, "gkr_binary_msm: coefficient out of range (the reference indexes out of bounds)");
    std::memset(out_xy, 0, 96);
    if (n_chunks == 0) return GKR_OK;
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n_chunks + 255) / 256, (uint64_t)ctx->num_sms * 2));
    uint8_t* d_coefs = nullptr;
    G1X* partial = nullptr;
    GKR_CUDA_OK(ctx, gkr_malloc_async(&d_coefs, n_chunks, st));
    GKR_CUDA_OK(ctx, gkr_malloc_async(&partial, sizeof(G1X) * (blocks + 1), st));
    int rc = gkr_stage_upload(ctx, d_coefs, coefs, n_chunks);
    gkr::G1XH h;
    if (rc == GKR_OK) {
        binmsm_sum_kernel<<<blocks, 256, sizeof(G1X) * 256, st>>>((const G1Aff*)prepared->d, per, d_coefs, n_chunks, partial);
        msm_window_tree_kernel<<<1, 256, sizeof(G1X) * 256, st>>>(partial, blocks, partial + blocks);
        ctx->launches += 2;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(&h, partial + blocks, sizeof(G1X), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = ctx->fail(GKR_ERR_CUDA, cudaGetErrorString(e));
    }
    gkr_free_async(d_coefs, st);
    gkr_free_async(partial, st);
    if (rc) return rc;
    gkr::g1h::to_affine(h, out_xy);
    return GKR_OK;
}
