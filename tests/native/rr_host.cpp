// CPU build of the reduced-radix field code (gkr-msm_b200/csrc/rr_field.cuh, dense29_item.cuh): the SAME source the device
// kernels compile, exported for tests/test_rr_field.py which checks it against python big integers.  Test infrastructure.
#include <cstring>
#include "../../gkr-msm_b200/csrc/lab/dense29_item.cuh"

extern "C" {
void rr_t_load(const uint32_t* w8, uint32_t* l9) {
    F29 a = f29_load(w8);
    std::memcpy(l9, a.l, 36);
}
void rr_t_store(const uint32_t* l9, uint32_t* w8) {
    F29 a;
    std::memcpy(a.l, l9, 36);
    rr_to_words<RrFr, 8>(a, w8);
}
void rr_t_canonical(uint32_t* w8, int times) { fr_words_canonical(w8, times); }
void rr_t_mul(const uint32_t* a9, const uint32_t* b9, uint32_t* out9) {
    F29 a, b;
    std::memcpy(a.l, a9, 36);
    std::memcpy(b.l, b9, 36);
    F29 r = rr_mul<RrFr>(a, b);
    std::memcpy(out9, r.l, 36);
}
void rr_t_sqr(const uint32_t* a9, uint32_t* out9) {
    F29 a;
    std::memcpy(a.l, a9, 36);
    F29 r = rr_sqr<RrFr>(a);
    std::memcpy(out9, r.l, 36);
}
void rr_t_sub_norm(const uint32_t* a9, const uint32_t* b9, uint32_t* out9) {
    F29 a, b;
    std::memcpy(a.l, a9, 36);
    std::memcpy(b.l, b9, 36);
    F29 r = rr_norm<RrFr>(rr_sub<RrFr>(a, b));
    std::memcpy(out9, r.l, 36);
}
void rr_t_fold(const uint32_t* e0w, const uint32_t* e1w, const uint32_t* t128, uint32_t* out9) {
    uint32_t t5[5];
    f29_challenge(t128, t5);
    F29 r = f29_fold(f29_load(e0w), f29_load(e1w), t5);
    std::memcpy(out9, r.l, 36);
}
// n pairs per table, tables [3][2 n][8] words (canonical or any value < 2^256): sums at the nodes 1..3 as canonical words
void rr_t_prod3_eval(const uint32_t* tabs, uint64_t n, uint32_t* sums /* [3][8] */) {
    Acc29 acc[3];
    for (int s = 0; s < 3; s++) acc29_zero(acc[s]);
    for (uint64_t i = 0; i < n; i++) {
        F29 lo[3], hi[3];
        for (int j = 0; j < 3; j++) {
            lo[j] = f29_load(tabs + ((size_t)j * 2 * n + 2 * i) * 8);
            hi[j] = f29_load(tabs + ((size_t)j * 2 * n + 2 * i + 1) * 8);
        }
        prod3_nodes29(lo, hi, acc);
        if ((i & 3) == 3)
            for (int s = 0; s < 3; s++) acc29_norm(acc[s]);
    }
    for (int s = 0; s < 3; s++) acc29_finish(acc[s], sums + 8 * s);
}
// n quads per table, tables [3][4 n][8]: folded tables [3][2 n][8] (canonical words) and the sums of the next round
void rr_t_prod3_fold_eval(const uint32_t* tabs, uint64_t n, const uint32_t* t128, uint32_t* folded, uint32_t* sums) {
    uint32_t t5[5];
    f29_challenge(t128, t5);
    Acc29 acc[3];
    for (int s = 0; s < 3; s++) acc29_zero(acc[s]);
    for (uint64_t i = 0; i < n; i++) {
        F29 lo[3], hi[3];
        for (int j = 0; j < 3; j++) {
            const uint32_t* src = tabs + ((size_t)j * 4 * n + 4 * i) * 8;
            lo[j] = f29_fold(f29_load(src), f29_load(src + 8), t5);
            hi[j] = f29_fold(f29_load(src + 16), f29_load(src + 24), t5);
            uint32_t* dst = folded + ((size_t)j * 2 * n + 2 * i) * 8;
            rr_to_words<RrFr, 8>(lo[j], dst);
            fr_words_canonical(dst, 2);
            rr_to_words<RrFr, 8>(hi[j], dst + 8);
            fr_words_canonical(dst + 8, 2);
        }
        prod3_nodes29(lo, hi, acc);
        if ((i & 3) == 3)
            for (int s = 0; s < 3; s++) acc29_norm(acc[s]);
    }
    for (int s = 0; s < 3; s++) acc29_finish(acc[s], sums + 8 * s);
}
}
