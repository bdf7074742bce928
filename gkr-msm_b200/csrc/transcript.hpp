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
    static const int ROT[5][5] = {{0, 36, 3, 41, 18}, {1, 44, 10, 45, 2}, {62, 6, 43, 15, 61}, {28, 55, 25, 21, 56}, {27, 20, 39, 8, 14}};
    for (int rnd = 0; rnd < 24; rnd++) {
        uint64_t c[5], d[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20];
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) st[i] ^= d[i % 5];
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(st[x + 5 * y], ROT[x][y]);
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) st[x + 5 * y] = b[x + 5 * y] ^ ((~b[(x + 1) % 5 + 5 * y]) & b[(x + 2) % 5 + 5 * y]);
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
    void write_scalars(const FrH* v, size_t n) {  // ark-serialize compressed Fr: 32 B LE canonical value
        std::vector<uint8_t> buf(32 * n);
        for (size_t i = 0; i < n; i++) frh::to_bytes_le(v[i], buf.data() + 32 * i);
        write_raw_msg(buf.data(), buf.size());
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
