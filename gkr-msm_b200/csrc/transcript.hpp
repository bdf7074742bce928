// Host-side Fiat-Shamir transcript: the C++ stand-in for the reference's Rust host while no Rust
// toolchain is available (north_star keeps the transcript on the host).
//   ProofTranscript2            src/cleanup/proof_transcript.rs:76-147
//   merlin 3.0.0 Transcript     third-party crate (Cargo.lock), STROBE-128 over Keccak-f[1600]
// Pinned against merlin's published test vector and hashlib.sha3_256 in tests/test_oracle_pins.py /
// tests/test_host_logic.py.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "host_field.hpp"

namespace gkr {

static inline uint64_t rotl64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

static inline void keccak_f1600(uint64_t st[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL,
        0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL,
        0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL,
        0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    // fully unrolled by hand (nvcc's front end drops GCC unroll pragmas): theta, rho+pi along the 24-cycle of pi, chi, iota
    for (int rnd = 0; rnd < 24; rnd++) {
        uint64_t c0 = st[0] ^ st[5] ^ st[10] ^ st[15] ^ st[20], c1 = st[1] ^ st[6] ^ st[11] ^ st[16] ^ st[21];
        uint64_t c2 = st[2] ^ st[7] ^ st[12] ^ st[17] ^ st[22], c3 = st[3] ^ st[8] ^ st[13] ^ st[18] ^ st[23];
        uint64_t c4 = st[4] ^ st[9] ^ st[14] ^ st[19] ^ st[24];
        const uint64_t d0 = c4 ^ rotl64(c1, 1), d1 = c0 ^ rotl64(c2, 1), d2 = c1 ^ rotl64(c3, 1), d3 = c2 ^ rotl64(c4, 1), d4 = c3 ^ rotl64(c0, 1);
        st[0] ^= d0; st[1] ^= d1; st[2] ^= d2; st[3] ^= d3; st[4] ^= d4;
        st[5] ^= d0; st[6] ^= d1; st[7] ^= d2; st[8] ^= d3; st[9] ^= d4;
        st[10] ^= d0; st[11] ^= d1; st[12] ^= d2; st[13] ^= d3; st[14] ^= d4;
        st[15] ^= d0; st[16] ^= d1; st[17] ^= d2; st[18] ^= d3; st[19] ^= d4;
        st[20] ^= d0; st[21] ^= d1; st[22] ^= d2; st[23] ^= d3; st[24] ^= d4;
        uint64_t t = st[1], b0;
        b0 = st[10]; st[10] = rotl64(t, 1); t = b0;
        b0 = st[7]; st[7] = rotl64(t, 3); t = b0;
        b0 = st[11]; st[11] = rotl64(t, 6); t = b0;
        b0 = st[17]; st[17] = rotl64(t, 10); t = b0;
        b0 = st[18]; st[18] = rotl64(t, 15); t = b0;
        b0 = st[3]; st[3] = rotl64(t, 21); t = b0;
        b0 = st[5]; st[5] = rotl64(t, 28); t = b0;
        b0 = st[16]; st[16] = rotl64(t, 36); t = b0;
        b0 = st[8]; st[8] = rotl64(t, 45); t = b0;
        b0 = st[21]; st[21] = rotl64(t, 55); t = b0;
        b0 = st[24]; st[24] = rotl64(t, 2); t = b0;
        b0 = st[4]; st[4] = rotl64(t, 14); t = b0;
        b0 = st[15]; st[15] = rotl64(t, 27); t = b0;
        b0 = st[23]; st[23] = rotl64(t, 41); t = b0;
        b0 = st[19]; st[19] = rotl64(t, 56); t = b0;
        b0 = st[13]; st[13] = rotl64(t, 8); t = b0;
        b0 = st[12]; st[12] = rotl64(t, 25); t = b0;
        b0 = st[2]; st[2] = rotl64(t, 43); t = b0;
        b0 = st[20]; st[20] = rotl64(t, 62); t = b0;
        b0 = st[14]; st[14] = rotl64(t, 18); t = b0;
        b0 = st[22]; st[22] = rotl64(t, 39); t = b0;
        b0 = st[9]; st[9] = rotl64(t, 61); t = b0;
        b0 = st[6]; st[6] = rotl64(t, 20); t = b0;
        b0 = st[1]; st[1] = rotl64(t, 44); t = b0;
        c0 = st[0]; c1 = st[1]; c2 = st[2]; c3 = st[3]; c4 = st[4];
        st[0] = c0 ^ (~c1 & c2); st[1] = c1 ^ (~c2 & c3); st[2] = c2 ^ (~c3 & c4); st[3] = c3 ^ (~c4 & c0); st[4] = c4 ^ (~c0 & c1);
        c0 = st[5]; c1 = st[6]; c2 = st[7]; c3 = st[8]; c4 = st[9];
        st[5] = c0 ^ (~c1 & c2); st[6] = c1 ^ (~c2 & c3); st[7] = c2 ^ (~c3 & c4); st[8] = c3 ^ (~c4 & c0); st[9] = c4 ^ (~c0 & c1);
        c0 = st[10]; c1 = st[11]; c2 = st[12]; c3 = st[13]; c4 = st[14];
        st[10] = c0 ^ (~c1 & c2); st[11] = c1 ^ (~c2 & c3); st[12] = c2 ^ (~c3 & c4); st[13] = c3 ^ (~c4 & c0); st[14] = c4 ^ (~c0 & c1);
        c0 = st[15]; c1 = st[16]; c2 = st[17]; c3 = st[18]; c4 = st[19];
        st[15] = c0 ^ (~c1 & c2); st[16] = c1 ^ (~c2 & c3); st[17] = c2 ^ (~c3 & c4); st[18] = c3 ^ (~c4 & c0); st[19] = c4 ^ (~c0 & c1);
        c0 = st[20]; c1 = st[21]; c2 = st[22]; c3 = st[23]; c4 = st[24];
        st[20] = c0 ^ (~c1 & c2); st[21] = c1 ^ (~c2 & c3); st[22] = c2 ^ (~c3 & c4); st[23] = c3 ^ (~c4 & c0); st[24] = c4 ^ (~c0 & c1);
        st[0] ^= RC[rnd];
    }
}

class Strobe128 {
    static constexpr int R = 166;
    static constexpr uint8_t FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32;
    union {
        uint64_t w[25];
        uint8_t b[200];
    } st;
    uint8_t pos = 0, pos_begin = 0, cur_flags = 0;

    void run_f() {
        st.b[pos] ^= pos_begin;
        st.b[pos + 1] ^= 0x04;
        st.b[R + 1] ^= 0x80;
        keccak_f1600(st.w);
        pos = 0;
        pos_begin = 0;
    }
    void absorb(const uint8_t* d, size_t n) {
        for (size_t i = 0; i < n; i++) {
            st.b[pos] ^= d[i];
            if (++pos == R) run_f();
        }
    }
    void squeeze(uint8_t* d, size_t n) {
        for (size_t i = 0; i < n; i++) {
            d[i] = st.b[pos];
            st.b[pos] = 0;
            if (++pos == R) run_f();
        }
    }
    void begin_op(uint8_t flags, bool more) {
        if (more) return;
        uint8_t old_begin = pos_begin;
        pos_begin = pos + 1;
        cur_flags = flags;
        uint8_t hdr[2] = {old_begin, flags};
        absorb(hdr, 2);
        bool force_f = (flags & (FLAG_C | FLAG_K)) != 0;
        if (force_f && pos != 0) run_f();
    }

   public:
    explicit Strobe128(const char* protocol_label) {
        std::memset(st.b, 0, 200);
        const uint8_t init[6] = {1, R + 2, 1, 0, 1, 96};
        std::memcpy(st.b, init, 6);
        std::memcpy(st.b + 6, "STROBEv1.0.2", 12);
        keccak_f1600(st.w);
        meta_ad((const uint8_t*)protocol_label, std::strlen(protocol_label), false);
    }
    void meta_ad(const uint8_t* d, size_t n, bool more) {
        begin_op(FLAG_M | FLAG_A, more);
        absorb(d, n);
    }
    void ad(const uint8_t* d, size_t n, bool more) {
        begin_op(FLAG_A, more);
        absorb(d, n);
    }
    void prf(uint8_t* d, size_t n, bool more) {
        begin_op(FLAG_I | FLAG_A | FLAG_C, more);
        squeeze(d, n);
    }
};

class MerlinTranscript {
    Strobe128 strobe;

   public:
    MerlinTranscript(const uint8_t* label, size_t n) : strobe("Merlin v1.0") { append_message((const uint8_t*)"dom-sep", 7, label, n); }
    void append_message(const uint8_t* label, size_t ln, const uint8_t* msg, size_t n) {
        uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
        strobe.meta_ad(label, ln, false);
        strobe.meta_ad(len, 4, true);
        strobe.ad(msg, n, false);
    }
    void challenge_bytes(const uint8_t* label, size_t ln, uint8_t* out, size_t n) {
        uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
        strobe.meta_ad(label, ln, false);
        strobe.meta_ad(len, 4, true);
        strobe.prf(out, n, false);
    }
};

// ProofTranscript2 in prover mode (proof_transcript.rs:76-147)
class ProofTranscript2 {
    MerlinTranscript merlin;

   public:
    std::vector<uint8_t> proof;
    ProofTranscript2(const uint8_t* pparam, size_t n) : merlin(pparam, n) {}
    void write_raw_msg(const uint8_t* msg, size_t n) {
        merlin.append_message(nullptr, 0, msg, n);
        proof.insert(proof.end(), msg, msg + n);
    }
    void raw_challenge(uint8_t* out, size_t n) { merlin.challenge_bytes(nullptr, 0, out, n); }
    // old API (src/transcript.rs:78-101): labelled merlin messages that are NOT part of a proof byte string, and labelled
    // 64-byte challenges
    void append_labeled(const uint8_t* label, size_t ln, const uint8_t* msg, size_t n) { merlin.append_message(label, ln, msg, n); }
    void challenge_labeled(const uint8_t* label, size_t ln, uint8_t* out, size_t n) { merlin.challenge_bytes(label, ln, out, n); }
    void write_scalars(const FrH* v, size_t n) {  // ark-serialize compressed Fr: 32 B LE canonical value
        uint8_t small[32 * 8] = {0};
        std::vector<uint8_t> big;
        uint8_t* buf = small;
        if (n > 8) {
            big.resize(32 * n);
            buf = big.data();
        }
        for (size_t i = 0; i < n; i++) frh::to_bytes_le(v[i], buf + 32 * i);
        write_raw_msg(buf, 32 * n);
    }
    FrH challenge(uint32_t bitsize) {  // F::from_le_bytes_mod_order(raw_challenge((bits+7)/8))
        uint8_t buf[64];
        size_t n = (bitsize + 7) / 8;
        if (n > 64) n = 64;
        raw_challenge(buf, n);
        return frh::from_le_bytes_mod_order(buf, n);
    }
};

}  // namespace gkr

// the opaque handle of the C ABI
struct gkr_transcript {
    gkr::ProofTranscript2 t;
    gkr_transcript(const uint8_t* l, size_t n) : t(l, n) {}
};
