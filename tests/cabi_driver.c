/* A plain-C host driving the C ABI (include/gkr_msm_b200.h) -- compiled with gcc, no C++, no python, no ctypes.
 *
 * It replays BareSumcheckSO::prove (src/cleanup/protocols/sumcheck.rs:646-691), i.e. GenericSumcheckProtocol::prove
 * (sumcheck.rs:101-123) round by round through the trait-shaped entries exactly as the reference's Rust host would:
 *     so.unipoly()  ->  UniPoly::from_evals  ->  compress_coefficients  ->  transcript.write_scalars  ->
 *     transcript.challenge(128)  ->  so.bind(x)                                        ... then write_scalars(final_evals)
 * over a Prod3 DenseSumcheckObjectSO on synthetic device tables, then runs the same proof through the ABI's own host loop
 * (gkr_sumcheck_prove) and checks that both transcripts hold the same bytes.  The proof is printed as hex; the pytest
 * (tests/test_cabi_driver.py) compares it with the python oracle's proof.
 *
 *   usage: cabi_driver <num_vars> <seed0> <seed1> <seed2>
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "gkr_msm_b200.h"

/* ---- the host side needs a little field arithmetic of its own (interpolation on nodes 0..3): BLS12-381 Fr, 4 x u64 Montgomery */
__extension__ typedef unsigned __int128 u128;
typedef struct { uint64_t v[4]; } fr;
static const uint64_t MOD[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
static const uint64_t NINV = 0xfffffffeffffffffULL;
static const fr ONE = {{0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL}};

static int geq(const uint64_t* a) {
    int i;
    for (i = 3; i >= 0; i--) {
        if (a[i] > MOD[i]) return 1;
        if (a[i] < MOD[i]) return 0;
    }
    return 1;
}
static void submod(uint64_t* a) {
    u128 br = 0;
    int i;
    for (i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - MOD[i] - br;
        a[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
}
static fr fr_add(fr a, fr b) {
    fr r;
    u128 c = 0;
    int i;
    for (i = 0; i < 4; i++) {
        c += (u128)a.v[i] + b.v[i];
        r.v[i] = (uint64_t)c;
        c >>= 64;
    }
    if (geq(r.v)) submod(r.v);
    return r;
}
static fr fr_sub(fr a, fr b) {
    fr r;
    u128 br = 0, c = 0;
    int i;
    for (i = 0; i < 4; i++) {
        u128 d = (u128)a.v[i] - b.v[i] - br;
        r.v[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
    if (br)
        for (i = 0; i < 4; i++) {
            c += (u128)r.v[i] + MOD[i];
            r.v[i] = (uint64_t)c;
            c >>= 64;
        }
    return r;
}
static fr fr_mul(fr a, fr b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    int i, j;
    fr r;
    for (i = 0; i < 4; i++) {
        u128 c = 0;
        uint64_t m;
        for (j = 0; j < 4; j++) {
            c += (u128)a.v[j] * b.v[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        m = t[0] * NINV;
        c = ((u128)m * MOD[0] + t[0]) >> 64;
        for (j = 1; j < 4; j++) {
            c += (u128)m * MOD[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    memcpy(r.v, t, 32);
    if (t[4] || geq(r.v)) submod(r.v);
    return r;
}
static fr fr_inv(fr a) { /* a^(r-2) */
    uint64_t e[4];
    fr acc = ONE;
    int i, b;
    memcpy(e, MOD, 32);
    e[0] -= 2;
    for (i = 3; i >= 0; i--)
        for (b = 63; b >= 0; b--) {
            acc = fr_mul(acc, acc);
            if ((e[i] >> b) & 1) acc = fr_mul(acc, a);
        }
    return acc;
}
/* UniPoly::from_evals on nodes 0..3 (Newton form expanded), coefficients low -> high */
static void from_evals4(const fr e[4], fr c[4]) {
    fr two = fr_add(ONE, ONE), three = fr_add(two, ONE), six = fr_add(three, three);
    fr i2 = fr_inv(two), i3 = fr_inv(three), i6 = fr_inv(six);
    fr d1 = fr_sub(e[1], e[0]);
    fr d2 = fr_add(fr_sub(e[2], fr_add(e[1], e[1])), e[0]);
    fr d3 = fr_sub(fr_add(fr_sub(e[3], fr_mul(three, e[2])), fr_mul(three, e[1])), e[0]);
    c[0] = e[0];
    c[3] = fr_mul(d3, i6);
    c[2] = fr_mul(fr_sub(d2, d3), i2);
    c[1] = fr_add(fr_sub(d1, fr_mul(d2, i2)), fr_mul(d3, i3));
}

#define CHECK(call)                                                                                   \
    do {                                                                                              \
        int rc__ = (call);                                                                            \
        if (rc__ != GKR_OK) {                                                                         \
            fprintf(stderr, "%s failed: %d (%s)\n", #call, rc__, ctx ? gkr_last_error(ctx) : "");   \
            return 2;                                                                                 \
        }                                                                                             \
    } while (0)

int main(int argc, char** argv) {
    gkr_ctx* ctx = NULL;
    gkr_table* tabs[3];
    gkr_so* so = NULL;
    gkr_transcript *tr = NULL, *tr2 = NULL;
    uint64_t claim[4], evals[16], chal[4], fin[12], fin2[12], out_claim[4], point[4 * 64];
    uint32_t nv, n_evals = 0, r;
    size_t len, len2, i;
    uint8_t *proof, *proof2;
    int j;

    if (argc != 5) {
        fprintf(stderr, "usage: %s num_vars seed0 seed1 seed2\n", argv[0]);
        return 2;
    }
    nv = (uint32_t)atoi(argv[1]);
    if (nv < 1 || nv > 28) return 2;
    CHECK(gkr_ctx_create(0, &ctx));
    for (j = 0; j < 3; j++) CHECK(gkr_table_synth(ctx, strtoull(argv[2 + j], NULL, 10), 0, (uint64_t)1 << nv, &tabs[j]));
    CHECK(gkr_dense_gate_sum(ctx, GKR_SO_PLAIN, GKR_GATE_PROD3, 0, NULL, 0, tabs, 3, claim));

    /* -- by hand, through the trait-shaped entries ------------------------------------------------------------------ */
    CHECK(gkr_so_create_dense(ctx, GKR_SO_PLAIN, GKR_GATE_PROD3, 0, NULL, 0, tabs, 3, nv, claim, &so));
    CHECK(gkr_transcript_new((const uint8_t*)"fgstglsp", 8, &tr));
    for (r = 0; r < nv; r++) {
        fr e[4], c[4];
        uint64_t msg[12];
        CHECK(gkr_so_unipoly(so, evals, &n_evals));
        if (n_evals != 4) return 3;
        for (j = 0; j < 4; j++) memcpy(e[j].v, evals + 4 * j, 32);
        from_evals4(e, c);
        memcpy(msg, c[0].v, 32); /* compress_coefficients: the linear term is dropped (sumcheck.rs:27-31) */
        memcpy(msg + 4, c[2].v, 32);
        memcpy(msg + 8, c[3].v, 32);
        CHECK(gkr_transcript_write_scalars(tr, msg, 3));
        CHECK(gkr_transcript_challenge(tr, 128, chal));
        CHECK(gkr_so_bind(so, chal));
    }
    CHECK(gkr_so_final_evals(so, fin));
    CHECK(gkr_transcript_write_scalars(tr, fin, 3));
    gkr_so_destroy(so);

    /* -- the same proof through the ABI's own host loop ---------------------------------------------------------------- */
    CHECK(gkr_so_create_dense(ctx, GKR_SO_PLAIN, GKR_GATE_PROD3, 0, NULL, 0, tabs, 3, nv, claim, &so));
    CHECK(gkr_transcript_new((const uint8_t*)"fgstglsp", 8, &tr2));
    CHECK(gkr_sumcheck_prove(tr2, so, nv, out_claim, point, fin2));
    CHECK(gkr_transcript_write_scalars(tr2, fin2, 3));
    gkr_so_destroy(so);

    len = gkr_transcript_proof_len(tr);
    len2 = gkr_transcript_proof_len(tr2);
    proof = (uint8_t*)malloc(len);
    proof2 = (uint8_t*)malloc(len2);
    CHECK(gkr_transcript_proof(tr, proof));
    CHECK(gkr_transcript_proof(tr2, proof2));
    if (len != len2 || memcmp(proof, proof2, len) != 0 || memcmp(fin, fin2, sizeof fin) != 0) {
        fprintf(stderr, "hand-driven rounds and gkr_sumcheck_prove disagree\n");
        return 4;
    }
    for (i = 0; i < len; i++) printf("%02x", proof[i]);
    printf("\n");
    fprintf(stderr, "cabi_driver: %u rounds, %zu proof bytes, %llu kernel launches\n", nv, len, (unsigned long long)gkr_ctx_launch_count(ctx));
    free(proof);
    free(proof2);
    gkr_transcript_free(tr);
    gkr_transcript_free(tr2);
    for (j = 0; j < 3; j++) gkr_table_free(tabs[j]);
    gkr_ctx_destroy(ctx);
    return 0;
}
