// Univariate / element-wise table algebra around the commitments:
//   KnucklesProvingKey::new (inverses)     src/commitments/knuckles.rs:65-81
//   KnucklesProvingKey::compute_t          src/commitments/knuckles.rs:111-154
//   div_by_linear, ev                      src/commitments/kzg.rs:73-81, 142-150
//   second_phase gathers                   src/cleanup/protocols/pushforward/pushforward.rs:585-595
//   c / d / ac tables from u32             pushforward.rs:479-500
//   combined_witness, folded_witness, p_lt src/cleanup/protocols/pippenger.rs:209-223, 274-279; opening.rs:65-75
// Every value is a field element determined by exact arithmetic, so the parallel evaluation orders used here
// (chunked Horner, scans of affine maps) reproduce the reference's serial loops bit for bit.
#include <algorithm>
#include "common.cuh"

// R^2 mod r (Montgomery form of a small integer v is mont_mul(v, R^2))
__device__ __forceinline__ Fr fr_r2() {
    Fr r;
    r.l[0] = 0xf3f29c6du; r.l[1] = 0xc999e990u; r.l[2] = 0x87925c23u; r.l[3] = 0x2b6cedcbu;
    r.l[4] = 0x7254398fu; r.l[5] = 0x05d31496u; r.l[6] = 0x9f59ff11u; r.l[7] = 0x0748d9d9u;
    return r;
}
__device__ __forceinline__ Fr fr_from_u64(uint64_t v) {
    Fr a = fr_zero();
    a.l[0] = (uint32_t)v;
    a.l[1] = (uint32_t)(v >> 32);
    return fr_mul(a, fr_r2());
}
__device__ Fr fr_pow_u64(Fr base, uint64_t e) {
    Fr r = fr_one();
    while (e) {
        if (e & 1) r = fr_mul(r, base);
        base = fr_sqr(base);
        e >>= 1;
    }
    return r;
}
__device__ Fr fr_inv(const Fr& a) {  // a^(r-2)
    const uint32_t e[8] = {FR_P0 - 2, FR_P1, FR_P2, FR_P3, FR_P4, FR_P5, FR_P6, FR_P7};  // low limb 1 - 2 borrows:
    // r - 2 = ...ffffffff00000001 - 2 = ...fffffffeffffffff
    uint32_t ee[8];
    for (int i = 0; i < 8; i++) ee[i] = e[i];
    ee[0] = 0xffffffffu;
    ee[1] = 0xfffffffeu;
    Fr r = fr_one();
    for (int i = 254; i >= 0; i--) {
        r = fr_sqr(r);
        if ((ee[i >> 5] >> (i & 31)) & 1) r = fr_mul(r, a);
    }
    return r;
}

// ---- u32 buffers, conversions, gathers ----------------------------------------------------------------
extern "C" int gkr_u32_upload(gkr_ctx* ctx, const uint32_t* vals, uint64_t n, gkr_u32buf** out) {
    if (!ctx) return GKR_ERR_ARG;
    if (!out || (!vals && n)) return ctx->fail(GKR_ERR_ARG, "null argument");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    gkr_u32buf* b = new gkr_u32buf();
    b->ctx = ctx;
    b->n = n;
    cudaError_t e = gkr_malloc_async(&b->d, sizeof(uint32_t) * std::max<uint64_t>(n, 1), ctx->stream);
    if (e == cudaSuccess && n && sizeof(uint32_t) * n <= ((size_t)1 << 20)) {
        if (gkr_stage_upload(ctx, b->d, vals, sizeof(uint32_t) * n)) e = cudaErrorUnknown;  // small: pinned ring, no synchronisation
    } else {
        if (e == cudaSuccess && n) e = cudaMemcpyAsync(b->d, vals, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
    if (e != cudaSuccess) { delete b; return ctx->fail(GKR_ERR_CUDA, cudaGetErrorString(e)); }
    *out = b;
    return GKR_OK;
}
extern "C" void gkr_u32_free(gkr_u32buf* b) {
    if (!b) return;
    if (b->d) gkr_free_async(b->d, b->ctx->stream);
    delete b;
}

__global__ void from_u32_kernel(Fr* out, const uint32_t* v, uint64_t n, int negate) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        Fr x = fr_from_u64(v[i]);
        out[i] = negate ? fr_neg(x) : x;
    }
}
// F::from(v) per entry (pushforward.rs:479-483), or its negation (access counts, :499-500)
extern "C" int gkr_table_from_u32(gkr_ctx* ctx, const gkr_u32buf* v, int negate, gkr_table** out) {
    if (!ctx) return GKR_ERR_ARG;
    if (!v || !out) return ctx->fail(GKR_ERR_ARG, "null argument");
    int rc = gkr_table_alloc(ctx, v->n, out);
    if (rc) return rc;
    if (v->n) {
        unsigned g = (unsigned)std::min<uint64_t>((v->n + 255) / 256, (uint64_t)ctx->num_sms * 8);
        from_u32_kernel<<<g, 256, 0, ctx->stream>>>((*out)->d, v->d, v->n, negate);
        ctx->launches++;
        GKR_CUDA_OK(ctx, cudaGetLastError());
    }
    return GKR_OK;
}

__global__ void gather_kernel(Fr* out, const Fr* src, uint64_t src_n, const uint32_t* idx, uint64_t n, int* bad) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t k = idx[i];
        if (k >= src_n) { *bad = 1; continue; }
        out[i] = src[k];
    }
}
// out[i] = src[idx[i]]   (c_pull / d_pull, pushforward.rs:585-595)
extern "C" int gkr_table_gather(gkr_ctx* ctx, const gkr_table* src, const gkr_u32buf* idx, gkr_table** out) {
    if (!ctx) return GKR_ERR_ARG;
    if (!src || !idx || !out) return ctx->fail(GKR_ERR_ARG, "null argument");
    int rc = gkr_table_alloc(ctx, idx->n, out);
    if (rc) return rc;
    int* d_bad = nullptr;
    GKR_CUDA_OK(ctx, gkr_malloc_async(&d_bad, sizeof(int), ctx->stream));
    GKR_CUDA_OK(ctx, cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
    if (idx->n) {
        unsigned g = (unsigned)std::min<uint64_t>((idx->n + 255) / 256, (uint64_t)ctx->num_sms * 8);
        gather_kernel<<<g, 256, 0, ctx->stream>>>((*out)->d, src->d, src->n, idx->d, idx->n, d_bad);
        ctx->launches++;
    }
    int bad = 0;
    GKR_CUDA_OK(ctx, cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    gkr_free_async(d_bad, ctx->stream);
    if (bad) {
        gkr_table_free(*out);
        *out = nullptr;
        return ctx->fail(GKR_ERR_ARG, "gather index out of range");
    }
    return GKR_OK;
}

// ---- linear combinations of (slices of) tables ---------------------------------------------------------
struct LinTerm {
    const Fr* src;
    uint64_t dst_off, len;
    Fr coef;
    int coef_is_one;
};
__global__ void lincomb_kernel(Fr* out, uint64_t n, const LinTerm* terms, int n_terms) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        Fr acc = fr_zero();
        for (int k = 0; k < n_terms; k++) {
            const LinTerm t = terms[k];
            if (i >= t.dst_off && i - t.dst_off < t.len) {
                if (!t.src) {  // constant term: the all-ones table times coef
                    acc = fr_add(acc, t.coef);
                    continue;
                }
                Fr v = t.src[i - t.dst_off];
                acc = fr_add(acc, t.coef_is_one ? v : fr_mul(v, t.coef));
            }
        }
        out[i] = acc;
    }
}
// out = zeros(out_len); out[dst_off_k + i] += coef_k * src_k[src_off_k + i], i < len_k.
// Covers `x + gamma*y`, the zero-extended `lambda*t + p`, folded_witness and the strided combined_witness.
// src_k == NULL stands for the all-ones table: constant shifts and pads (c_adj / d_adj, pushforward.rs:700-710).
extern "C" int gkr_table_lincomb(gkr_ctx* ctx, uint32_t n_terms, gkr_table* const* src, const uint64_t* coefs, const uint64_t* src_off,
                                 const uint64_t* dst_off, const uint64_t* len, uint64_t out_len, gkr_table** out) {
    if (!ctx) return GKR_ERR_ARG;
    if (!out || (n_terms && (!src || !coefs || !src_off || !dst_off || !len))) return ctx->fail(GKR_ERR_ARG, "null argument");
    std::vector<LinTerm> terms(n_terms);
    for (uint32_t k = 0; k < n_terms; k++) {
        if ((src[k] && src_off[k] + len[k] > src[k]->n) || dst_off[k] + len[k] > out_len) return ctx->fail(GKR_ERR_ARG, "slice out of range");
        gkr::FrH c = frh_from_limbs(coefs + 4 * k);
        if (!frh_canonical(c)) return ctx->fail(GKR_ERR_ARG, "coefficient not canonical");
        terms[k].src = src[k] ? src[k]->d + src_off[k] : nullptr;  // NULL source = the constant-one table
        terms[k].dst_off = dst_off[k];
        terms[k].len = len[k];
        terms[k].coef = fr_from_host(c);
        terms[k].coef_is_one = c == gkr::frh::ONE;
    }
    int rc = gkr_table_alloc(ctx, out_len, out);
    if (rc) return rc;
    LinTerm* d_terms = nullptr;
    GKR_CUDA_OK(ctx, gkr_malloc_async(&d_terms, sizeof(LinTerm) * std::max<uint32_t>(n_terms, 1), ctx->stream));
    if (n_terms) {  // staged through the pinned ring: no pageable copy, no synchronisation before `terms` goes out of scope
        int rcs = gkr_stage_upload(ctx, d_terms, terms.data(), sizeof(LinTerm) * n_terms);
        if (rcs) return rcs;
    }
    if (out_len) {
        unsigned g = (unsigned)std::min<uint64_t>((out_len + 255) / 256, (uint64_t)ctx->num_sms * 8);
        lincomb_kernel<<<g, 256, 0, ctx->stream>>>((*out)->d, out_len, d_terms, (int)n_terms);
        ctx->launches++;
        GKR_CUDA_OK(ctx, cudaGetLastError());
    }
    gkr_free_async(d_terms, ctx->stream);
    return GKR_OK;
}

// ---- Horner evaluation and synthetic division -----------------------------------------------------------
// chunk sums S_c = sum_{j in chunk} p[b + j] x^j
__global__ void horner_chunks_kernel(const Fr* p, uint64_t n, uint64_t chunk, Fr x, Fr* S, uint64_t n_chunks) {
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t b = c * chunk, e = b + chunk < n ? b + chunk : n;
        Fr acc = fr_zero();
        for (uint64_t j = e; j-- > b;) acc = fr_add(fr_mul(acc, x), p[j]);
        S[c] = acc;
    }
}
// H[c] = sum_{c' > c} S[c'] x^{chunk * (c' - c - 1)}: value carried into chunk c from everything above it.
// One block; n_chunks <= 512.  Suffix scan of the affine maps h -> S + x^chunk * h.
__global__ void __launch_bounds__(512) horner_scan_kernel(const Fr* S, uint64_t n_chunks, uint64_t chunk, Fr x, Fr* H, Fr* total) {
    __shared__ Fr A[512];
    __shared__ Fr B[512];
    const uint32_t t = threadIdx.x;
    const Fr xc = fr_pow_u64(x, chunk);
    // element t represents the map of chunk index (n_chunks - 1 - t): reversed so that a PREFIX scan composes from the top
    uint64_t c = n_chunks - 1 - t;
    bool live = t < n_chunks;
    A[t] = live ? xc : fr_one();
    B[t] = live ? S[c] : fr_zero();
    __syncthreads();
    // inclusive prefix composition: F_t = f_t o F_{t-1} where f(h) = B + A h  and F_{-1} = identity... we want
    // G_t(0) = value after applying maps of chunks (top .. c): G_t = f_t(G_{t-1}), G_{-1} = 0.
    for (uint32_t s = 1; s < 512; s <<= 1) {
        Fr a2 = A[t], b2 = B[t];
        Fr a1 = fr_one(), b1 = fr_zero();
        bool has = t >= s;
        if (has) { a1 = A[t - s]; b1 = B[t - s]; }
        __syncthreads();
        if (has) {
            // (a2, b2) o (a1, b1): h -> b2 + a2 (b1 + a1 h)
            B[t] = fr_add(b2, fr_mul(a2, b1));
            A[t] = fr_mul(a2, a1);
        }
        __syncthreads();
    }
    // B[t] = G_t(0) = sum_{c' >= c} S[c'] x^{chunk (c' - c)}; carry INTO chunk c is G_{t-1}(0)
    if (live) H[c] = t == 0 ? fr_zero() : B[t - 1];
    if (t == 0 && total) *total = B[n_chunks - 1];
}
// quotient of p(X) / (X - x): q[i] = sum_{j > i} p[j] x^{j - i - 1}   (kzg.rs:73-81)
// (q_len = n - 1 for the quotient itself; q_len = n when q is the carry array of the level below, see poly_levels)
__global__ void divlin_chunks_kernel(const Fr* p, uint64_t n, uint64_t chunk, Fr x, const Fr* H, Fr* q, uint64_t n_chunks, uint64_t q_len) {
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t b = c * chunk, e = b + chunk < n ? b + chunk : n;
        Fr rem = H[c];  // = sum_{j >= e} p[j] x^{j - e}
        for (uint64_t j = e; j-- > b;) {
            if (j < q_len) q[j] = rem;
            rem = fr_add(p[j], fr_mul(rem, x));
        }
    }
}

// Both entries walk the same hierarchy: level 0 is the polynomial; level k+1 holds the Horner sums of the POLY_CHUNK-element
// chunks of level k, a polynomial in x_{k+1} = x_k^POLY_CHUNK.  One thread per chunk (a chain of POLY_CHUNK dependent
// products, 2^15 threads for 2^21 coefficients) until at most 512 values are left, which one block scans.  The carry INTO
// element j of level k+1 (= the quotient coefficient of that level) is the carry into chunk j of level k, so the quotient is
// built top-down with the same divlin kernel.  (The first version cut the polynomial into 512 chunks whatever its length:
// 512 threads walking 4096 dependent products each, 2.6 / 3.9 ms per call at 2^21 under ncu.)
#define POLY_CHUNK 64
struct PolyLevels {
    std::vector<const Fr*> data;  // level arrays (level 0 = the input)
    std::vector<uint64_t> len;
    std::vector<gkr::FrH> x;
    Fr* buf = nullptr;            // all level arrays + top carries + total
    Fr* top_carry = nullptr;      // [len.back()] carries into the elements of the top level
    Fr* total = nullptr;          // p(x)
};

static int poly_levels_build(gkr_ctx* ctx, const Fr* p, uint64_t n, const gkr::FrH& x, PolyLevels* L) {
    cudaStream_t st = ctx->stream;
    L->data.assign(1, p);
    L->len.assign(1, n);
    L->x.assign(1, x);
    uint64_t extra = 0;
    for (uint64_t m = n; m > 512;) {
        m = (m + POLY_CHUNK - 1) / POLY_CHUNK;
        extra += m;
    }
    uint64_t top = n;
    while (top > 512) top = (top + POLY_CHUNK - 1) / POLY_CHUNK;
    GKR_CUDA_OK(ctx, gkr_malloc_async(&L->buf, sizeof(Fr) * (extra + top + 1), st));
    Fr* cursor = L->buf;
    while (L->len.back() > 512) {
        const uint64_t m = L->len.back(), n_chunks = (m + POLY_CHUNK - 1) / POLY_CHUNK;
        horner_chunks_kernel<<<(unsigned)((n_chunks + 127) / 128), 128, 0, st>>>(L->data.back(), m, POLY_CHUNK, fr_from_host(L->x.back()), cursor, n_chunks);
        ctx->launches++;
        gkr::FrH y = L->x.back();
        for (int i = 0; i < 6; i++) y = gkr::frh::mul(y, y);  // x^64
        L->data.push_back(cursor);
        L->len.push_back(n_chunks);
        L->x.push_back(y);
        cursor += n_chunks;
    }
    L->top_carry = cursor;
    L->total = cursor + L->len.back();
    // top level: chunks of ONE element, so the scan yields the carry into every element and the total
    horner_scan_kernel<<<1, 512, 0, st>>>(L->data.back(), L->len.back(), 1, fr_from_host(L->x.back()), L->top_carry, L->total);
    ctx->launches++;
    GKR_CUDA_OK(ctx, cudaGetLastError());
    return GKR_OK;
}

// ev(poly, x) = sum poly[i] x^i   (kzg.rs:142-150)
extern "C" int gkr_poly_eval(gkr_ctx* ctx, const gkr_table* poly, const uint64_t x[4], uint64_t out[4]) {
    if (!ctx) return GKR_ERR_ARG;
    if (!poly || !x || !out) return ctx->fail(GKR_ERR_ARG, "null argument");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    if (poly->n == 0) { std::memset(out, 0, 32); return GKR_OK; }
    cudaStream_t st = ctx->stream;
    PolyLevels L;
    int rc = poly_levels_build(ctx, poly->d, poly->n, frh_from_limbs(x), &L);
    if (rc) return rc;
    Fr r;
    GKR_CUDA_OK(ctx, cudaMemcpyAsync(&r, L.total, sizeof(Fr), cudaMemcpyDeviceToHost, st));
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(st));
    gkr_free_async(L.buf, st);
    frh_to_limbs(fr_to_host(r), out);
    return GKR_OK;
}

// div_by_linear(poly, pt) -> (quotient of len-1 entries, remainder)   (kzg.rs:73-81)
extern "C" int gkr_poly_div_by_linear(gkr_ctx* ctx, const gkr_table* poly, const uint64_t pt[4], gkr_table** quotient, uint64_t rem[4]) {
    if (!ctx) return GKR_ERR_ARG;
    if (!poly || !pt || !quotient || poly->n == 0) return ctx->fail(GKR_ERR_ARG, "null / empty argument");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int rc = gkr_table_alloc(ctx, poly->n - 1, quotient);
    if (rc) return rc;
    PolyLevels L;
    rc = poly_levels_build(ctx, poly->d, poly->n, frh_from_limbs(pt), &L);
    if (rc) return rc;
    // top-down: the carries into the elements of level k+1 are the carries into the chunks of level k
    const Fr* carry = L.top_carry;
    Fr* scratch = nullptr;  // carry arrays of the intermediate levels
    uint64_t scratch_len = 0;
    for (size_t k = 1; k + 1 < L.len.size(); k++) scratch_len += L.len[k];
    if (scratch_len) GKR_CUDA_OK(ctx, gkr_malloc_async(&scratch, sizeof(Fr) * scratch_len, st));
    Fr* sc = scratch;
    for (size_t k = L.len.size() - 1; k-- > 0;) {
        const uint64_t m = L.len[k], n_chunks = (m + POLY_CHUNK - 1) / POLY_CHUNK;
        Fr* q = k == 0 ? (*quotient)->d : sc;
        const uint64_t q_len = k == 0 ? m - 1 : m;
        divlin_chunks_kernel<<<(unsigned)((n_chunks + 127) / 128), 128, 0, st>>>(L.data[k], m, POLY_CHUNK, fr_from_host(L.x[k]), carry, q, n_chunks, q_len);
        ctx->launches++;
        carry = q;
        if (k != 0) sc += m;
    }
    if (L.len.size() == 1 && poly->n > 1)  // at most 512 coefficients: the top-level carries ARE the quotient
        GKR_CUDA_OK(ctx, cudaMemcpyAsync((*quotient)->d, L.top_carry, sizeof(Fr) * (poly->n - 1), cudaMemcpyDeviceToDevice, st));
    GKR_CUDA_OK(ctx, cudaGetLastError());
    Fr r;
    GKR_CUDA_OK(ctx, cudaMemcpyAsync(&r, L.total, sizeof(Fr), cudaMemcpyDeviceToHost, st));
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(st));
    gkr_free_async(L.buf, st);
    if (scratch) gkr_free_async(scratch, st);
    if (rem) frh_to_limbs(fr_to_host(r), rem);
    return GKR_OK;
}

// ---- Knuckles ---------------------------------------------------------------------------------------------
struct gkr_knuckles {
    gkr_ctx* ctx = nullptr;
    uint32_t num_vars = 0;
    gkr::FrH k;
    Fr* inverses = nullptr;  // [2n - 1]: 1 / (k^s - k^(n-1)), entry n-1 is 1 (knuckles.rs:65-81)
};

__global__ void knuckles_kpows_kernel(Fr* out, uint64_t len, uint64_t n, Fr k) {
    const Fr kn = fr_pow_u64(k, n - 1);
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < len; s += (uint64_t)gridDim.x * blockDim.x) {
        Fr v = fr_sub(fr_pow_u64(k, s), kn);
        if (s == n - 1) v = fr_add(v, fr_one());  // "so inversion doesn't fail"
        out[s] = v;
    }
}
// in-place batch inversion, zeros stay zero (ark_ff::batch_inversion); one Fermat inversion per 32 entries
__global__ void batch_inverse_kernel(Fr* v, uint64_t n) {
    const int CH = 32;
    const uint64_t n_chunks = (n + CH - 1) / CH;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t b = c * CH, e = b + CH < n ? b + CH : n;
        Fr pref[CH];
        Fr acc = fr_one();
        for (uint64_t j = b; j < e; j++) {
            pref[j - b] = acc;
            Fr x = v[j];
            if (!fr_is_zero(x)) acc = fr_mul(acc, x);
        }
        Fr inv = fr_inv(acc);
        for (uint64_t j = e; j-- > b;) {
            Fr x = v[j];
            if (fr_is_zero(x)) continue;
            v[j] = fr_mul(inv, pref[j - b]);
            inv = fr_mul(inv, x);
        }
    }
}

extern "C" int gkr_knuckles_create(gkr_ctx* ctx, uint32_t num_vars, const uint64_t k[4], gkr_knuckles** out) {
    if (!ctx) return GKR_ERR_ARG;
    if (!k || !out || num_vars >= 31) return ctx->fail(GKR_ERR_ARG, "bad argument");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    gkr_knuckles* key = new gkr_knuckles();
    key->ctx = ctx;
    key->num_vars = num_vars;
    key->k = frh_from_limbs(k);
    const uint64_t n = (uint64_t)1 << num_vars, len = 2 * n - 1;
    cudaError_t e = gkr_malloc_async(&key->inverses, sizeof(Fr) * len, ctx->stream);
    if (e != cudaSuccess) { delete key; return ctx->fail(GKR_ERR_CUDA, cudaGetErrorString(e)); }
    unsigned g = (unsigned)std::min<uint64_t>((len + 127) / 128, (uint64_t)ctx->num_sms * 8);
    knuckles_kpows_kernel<<<g, 128, 0, ctx->stream>>>(key->inverses, len, n, fr_from_host(key->k));
    unsigned g2 = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((len / 32 + 63) / 64, (uint64_t)ctx->num_sms * 8));
    batch_inverse_kernel<<<g2, 64, 0, ctx->stream>>>(key->inverses, len);
    ctx->launches += 2;
    GKR_CUDA_OK(ctx, cudaGetLastError());
    *out = key;
    return GKR_OK;
}
extern "C" uint32_t gkr_knuckles_num_vars(const gkr_knuckles* key) { return key ? key->num_vars : 0; }
extern "C" int gkr_knuckles_k(const gkr_knuckles* key, uint64_t out[4]) {
    if (!key || !out) return GKR_ERR_ARG;
    frh_to_limbs(key->k, out);
    return GKR_OK;
}
extern "C" void gkr_knuckles_free(gkr_knuckles* key) {
    if (!key) return;
    if (key->inverses) gkr_free_async(key->inverses, key->ctx->stream);
    delete key;
}

// one pass of knuckles.rs:131-146: t <- t * (pt + (1 - pt) X^offset)
__global__ void compute_t_pass_kernel(Fr* out, const Fr* in, uint64_t new_size, uint64_t offset, Fr one_minus_pt) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < new_size; i += (uint64_t)gridDim.x * blockDim.x) {
        Fr x = in[i];
        Fr v = fr_sub(x, fr_mul(x, one_minus_pt));
        if (i >= offset) v = fr_add(v, fr_mul(in[i - offset], one_minus_pt));
        out[i] = v;
    }
}
__global__ void compute_t_finish_kernel(Fr* t, const Fr* inverses, uint64_t len, uint64_t n, Fr* opening) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (uint64_t)gridDim.x * blockDim.x) {
        Fr x = t[i];
        if (i == n - 1) {
            *opening = x;
            x = fr_zero();
        }
        t[i] = fr_mul(x, inverses[i]);
    }
}

// KnucklesProvingKey::compute_t(poly, point) -> (t of 2n-1 entries, opening)   knuckles.rs:111-154
extern "C" int gkr_knuckles_compute_t(gkr_ctx* ctx, const gkr_knuckles* key, const gkr_table* poly, const uint64_t* point, uint32_t n_point,
                                      gkr_table** t_out, uint64_t opening[4]) {
    if (!ctx) return GKR_ERR_ARG;
    if (!key || !poly || !point || !t_out) return ctx->fail(GKR_ERR_ARG, "null argument");
    if (n_point != key->num_vars) return ctx->fail(GKR_ERR_ARG, "point.len() != num_vars");  // knuckles.rs:112
    const uint64_t n = (uint64_t)1 << key->num_vars, len = 2 * n - 1;
    if (poly->n > n) return ctx->fail(GKR_ERR_ARG, "poly.len() > n");  // knuckles.rs:118
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    gkr_table *a = nullptr, *b = nullptr;
    int rc = gkr_table_alloc(ctx, len, &a);
    if (rc) return rc;
    rc = gkr_table_alloc(ctx, len, &b);
    if (rc) { gkr_table_free(a); return rc; }
    GKR_CUDA_OK(ctx, cudaMemsetAsync(a->d, 0, sizeof(Fr) * len, st));
    GKR_CUDA_OK(ctx, cudaMemsetAsync(b->d, 0, sizeof(Fr) * len, st));
    if (poly->n) GKR_CUDA_OK(ctx, cudaMemcpyAsync(a->d, poly->d, sizeof(Fr) * poly->n, cudaMemcpyDeviceToDevice, st));
    uint64_t curr = n;
    for (uint32_t i = 0; i < key->num_vars; i++) {
        // pt.reverse(): pass i uses point[num_vars - 1 - i]
        gkr::FrH p = frh_from_limbs(point + 4 * (key->num_vars - 1 - i));
        if (!frh_canonical(p)) { gkr_table_free(a); gkr_table_free(b); return ctx->fail(GKR_ERR_ARG, "point not canonical"); }
        gkr::FrH omp = gkr::frh::sub(gkr::frh::ONE, p);
        uint64_t offset = (uint64_t)1 << i;
        curr += offset;
        unsigned g = (unsigned)std::min<uint64_t>((curr + 255) / 256, (uint64_t)ctx->num_sms * 8);
        compute_t_pass_kernel<<<g, 256, 0, st>>>(b->d, a->d, curr, offset, fr_from_host(omp));
        ctx->launches++;
        std::swap(a, b);
    }
    Fr* d_open = nullptr;
    GKR_CUDA_OK(ctx, gkr_malloc_async(&d_open, sizeof(Fr), st));
    unsigned g = (unsigned)std::min<uint64_t>((len + 255) / 256, (uint64_t)ctx->num_sms * 8);
    compute_t_finish_kernel<<<g, 256, 0, st>>>(a->d, key->inverses, len, n, d_open);
    ctx->launches++;
    GKR_CUDA_OK(ctx, cudaGetLastError());
    Fr o;
    GKR_CUDA_OK(ctx, cudaMemcpyAsync(&o, d_open, sizeof(Fr), cudaMemcpyDeviceToHost, st));
    GKR_CUDA_OK(ctx, cudaStreamSynchronize(st));
    gkr_free_async(d_open, st);
    gkr_table_free(b);
    if (opening) frh_to_limbs(fr_to_host(o), opening);
    *t_out = a;
    return GKR_OK;
}
