/* TEST INFRASTRUCTURE ONLY -- CPU oracle, never linked into or called from the product path.
 *
 * Fiat-Shamir transcript of the reference:  src/cleanup/proof_transcript.rs:17-147  (TProofTranscript2 / ProofTranscript2)
 *     write_raw_msg   = merlin append_message(b"", msg) + proof.extend(msg)                          :128-131
 *     raw_challenge   = merlin challenge_bytes(b"", buf)                                               :109-113
 *     challenge(bits) = F::from_le_bytes_mod_order(raw_challenge((bits + 7) / 8))                     :33-41
 *     write_scalars   = ark-serialize compressed Fr (32 B little-endian canonical value) each         :52-57
 *     write_points    = ark-serialize compressed G1 (48 B) each                                       :64-69
 * merlin 3.0.0 / keccak 0.1.5 are third-party crates absent from /root/reference (Cargo.lock); their published
 * construction (STROBE-128 over Keccak-f[1600], "Merlin v1.0" domain separator) is restated here and pinned by merlin's
 * own documented test vector in tests/test_oracle_pins.py.
 */
#pragma once
#include <string>
#include <vector>
#include "po_g1.hpp"

namespace po {

static inline uint64_t rol64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

static inline void keccak_f1600(uint64_t st[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL, 0x0000000080000001ULL,
        0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
        0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
        0x000000000000800AULL, 0x800000008000000AULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const int ROT[5][5] = {{0, 36, 3, 41, 18}, {1, 44, 10, 45, 2}, {62, 6, 43, 15, 61}, {28, 55, 25, 21, 56}, {27, 20, 39, 8, 14}};
    for (int rnd = 0; rnd < 24; rnd++) {
        uint64_t c[5], d[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20];
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) st[x + 5 * y] ^= d[x];
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(st[x + 5 * y], ROT[x][y]);
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) st[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        st[0] ^= RC[rnd];
    }
}

class Strobe128 {
    static constexpr int R = 166;
    enum { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };
    union {
        uint64_t w[25];
        uint8_t b[200];
    } st;
    int pos = 0, pos_begin = 0, cur_flags = 0;
    void run_f() {
        st.b[pos] ^= (uint8_t)pos_begin;
        st.b[pos + 1] ^= 0x04;
        st.b[R + 1] ^= 0x80;
        keccak_f1600(st.w);
        pos = 0;
        pos_begin = 0;
    }
    void absorb(const uint8_t* data, size_t n) {
        for (size_t i = 0; i < n; i++) {
            st.b[pos] ^= data[i];
            if (++pos == R) run_f();
        }
    }
    void squeeze(uint8_t* out, size_t n) {
        for (size_t i = 0; i < n; i++) {
            out[i] = st.b[pos];
            st.b[pos] = 0;
            if (++pos == R) run_f();
        }
    }
    void begin_op(int flags, bool more) {
        if (more) return;  // continuation of the same operation
        int old_begin = pos_begin;
        pos_begin = pos + 1;
        cur_flags = flags;
        uint8_t hdr[2] = {(uint8_t)old_begin, (uint8_t)flags};
        absorb(hdr, 2);
        if ((flags & (FLAG_C | FLAG_K)) && pos != 0) run_f();
    }

   public:
    explicit Strobe128(const char* label) {
        memset(st.b, 0, 200);
        const uint8_t init[6] = {1, R + 2, 1, 0, 1, 96};
        memcpy(st.b, init, 6);
        memcpy(st.b + 6, "STROBEv1.0.2", 12);
        keccak_f1600(st.w);
        meta_ad((const uint8_t*)label, strlen(label), false);
    }
    void meta_ad(const uint8_t* d, size_t n, bool more) {
        begin_op(FLAG_M | FLAG_A, more);
        absorb(d, n);
    }
    void ad(const uint8_t* d, size_t n, bool more) {
        begin_op(FLAG_A, more);
        absorb(d, n);
    }
    void prf(uint8_t* out, size_t n, bool more) {
        begin_op(FLAG_I | FLAG_A | FLAG_C, more);
        squeeze(out, n);
    }
};

class Merlin {
    Strobe128 s;

   public:
    explicit Merlin(const std::string& label) : s("Merlin v1.0") { append_message("dom-sep", (const uint8_t*)label.data(), label.size()); }
    void append_message(const char* label, const uint8_t* msg, size_t n) {
        uint32_t len = (uint32_t)n;
        uint8_t le[4] = {(uint8_t)len, (uint8_t)(len >> 8), (uint8_t)(len >> 16), (uint8_t)(len >> 24)};
        s.meta_ad((const uint8_t*)label, strlen(label), false);
        s.meta_ad(le, 4, true);
        s.ad(msg, n, false);
    }
    void challenge_bytes(const char* label, uint8_t* out, size_t n) {
        uint32_t len = (uint32_t)n;
        uint8_t le[4] = {(uint8_t)len, (uint8_t)(len >> 8), (uint8_t)(len >> 16), (uint8_t)(len >> 24)};
        s.meta_ad((const uint8_t*)label, strlen(label), false);
        s.meta_ad(le, 4, true);
        s.prf(out, n, false);
    }
};

/* ProofTranscript2, prover side (proof_transcript.rs:76-147) */
class Transcript {
    Merlin m;

   public:
    std::vector<uint8_t> proof;
    explicit Transcript(const std::string& pparam) : m(pparam) {}
    void write_raw_msg(const uint8_t* msg, size_t n) {
        m.append_message("", msg, n);
        proof.insert(proof.end(), msg, msg + n);
    }
    Fr challenge(int bitsize) {
        uint8_t buf[64];
        size_t n = (size_t)(bitsize + 7) / 8;
        m.challenge_bytes("", buf, n);
        return fr_from_le_bytes_mod_order(buf, n);
    }
    /* challenge_vec(n, bitsize): ONE raw challenge of n * bytesize bytes cut into n pieces (proof_transcript.rs:43-50) */
    std::vector<Fr> challenge_vec(int n, int bitsize) {
        size_t bs = (size_t)(bitsize + 7) / 8;
        std::vector<uint8_t> buf(bs * n);
        m.challenge_bytes("", buf.data(), buf.size());
        std::vector<Fr> out;
        for (int i = 0; i < n; i++) out.push_back(fr_from_le_bytes_mod_order(buf.data() + bs * i, bs));
        return out;
    }
    void write_scalars(const Fr* v, size_t n) {
        std::vector<uint8_t> buf(32 * n);
        for (size_t i = 0; i < n; i++) fr_serialize(v[i], buf.data() + 32 * i);
        write_raw_msg(buf.data(), buf.size());
    }
    void write_scalars(const std::vector<Fr>& v) { write_scalars(v.data(), v.size()); }
    void write_points(const G1A* p, size_t n) {
        std::vector<uint8_t> buf(48 * n);
        for (size_t i = 0; i < n; i++) g1_serialize(p[i], buf.data() + 48 * i);
        write_raw_msg(buf.data(), buf.size());
    }
    void write_points(const std::vector<G1A>& p) { write_points(p.data(), p.size()); }
};
}  // namespace po
