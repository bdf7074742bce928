// extern "C" surface declared in include/gkr_msm_b200.h: thin argument checking + dispatch.
#include <atomic>
#include "common.cuh"
#include "so.hpp"
#include "transcript.hpp"
#include "host_g1.hpp"

int gkr_dense_gate_sum_impl(gkr_ctx* ctx, int so_kind, int gate, uint32_t gate_param, const gkr::FrH* consts, uint32_t n_consts,
                            gkr_table* const* tables, uint32_t n_polys, gkr::FrH* out);

static std::vector<gkr::FrH> load_frs(const uint64_t* p, uint32_t n) {
    std::vector<gkr::FrH> v(n);
    for (uint32_t i = 0; i < n; i++) v[i] = frh_from_limbs(p + 4 * i);
    return v;
}

extern "C" int gkr_so_create_dense(gkr_ctx* ctx, int so_kind, int gate, uint32_t gate_param, const uint64_t* gate_consts,
                                   uint32_t n_consts, gkr_table* const* tables, uint32_t n_polys, uint32_t num_vars,
                                   const uint64_t claim[4], gkr_so** out) {
    if (!ctx) return GKR_ERR_ARG;
    if (!out || !tables || !claim || (!gate_consts && n_consts)) return ctx->fail(GKR_ERR_ARG, "null argument");
    if (n_polys == 0 || n_polys > GKR_MAX_POLYS) return ctx->fail(GKR_ERR_ARG, "bad number of tables");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    std::vector<gkr::FrH> c = load_frs(gate_consts, n_consts);
    return gkr_make_dense_so(ctx, so_kind, gate, gate_param, c.data(), n_consts, tables, n_polys, num_vars, frh_from_limbs(claim), out);
}

extern "C" int gkr_dense_gate_sum(gkr_ctx* ctx, int so_kind, int gate, uint32_t gate_param, const uint64_t* gate_consts,
                                  uint32_t n_consts, gkr_table* const* tables, uint32_t n_polys, uint64_t out[4]) {
    if (!ctx) return GKR_ERR_ARG;
    if (!out || !tables || (!gate_consts && n_consts)) return ctx->fail(GKR_ERR_ARG, "null argument");
    if (n_polys == 0 || n_polys > GKR_MAX_POLYS) return ctx->fail(GKR_ERR_ARG, "bad number of tables");
    GKR_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    std::vector<gkr::FrH> c = load_frs(gate_consts, n_consts);
    gkr::FrH r;
    int rc = gkr_dense_gate_sum_impl(ctx, so_kind, gate, gate_param, c.data(), n_consts, tables, n_polys, &r);
    if (rc == GKR_OK) frh_to_limbs(r, out);
    return rc;
}

extern "C" int gkr_so_unipoly(gkr_so* so, uint64_t* evals_out, uint32_t* n_evals) {
    if (!so || !evals_out) return GKR_ERR_ARG;
    gkr::FrH ev[GKR_MAX_DEG + 1];
    uint32_t n = 0;
    int rc = so->unipoly(ev, &n);
    if (rc) return rc;
    for (uint32_t i = 0; i < n; i++) frh_to_limbs(ev[i], evals_out + 4 * i);
    if (n_evals) *n_evals = n;
    return GKR_OK;
}

extern "C" int gkr_so_bind(gkr_so* so, const uint64_t t[4]) {
    if (!so || !t) return GKR_ERR_ARG;
    return so->bind(frh_from_limbs(t));
}

extern "C" int gkr_so_final_evals(gkr_so* so, uint64_t* out) {
    if (!so || !out) return GKR_ERR_ARG;
    std::vector<gkr::FrH> v(so->num_polys());
    int rc = so->final_evals(v.data());
    if (rc) return rc;
    for (size_t i = 0; i < v.size(); i++) frh_to_limbs(v[i], out + 4 * i);
    return GKR_OK;
}

extern "C" int gkr_so_claim(const gkr_so* so, uint64_t out[4]) {
    if (!so || !out) return GKR_ERR_ARG;
    frh_to_limbs(so->claim(), out);
    return GKR_OK;
}

extern "C" uint32_t gkr_so_degree(const gkr_so* so) { return so ? so->degree() : 0; }
extern "C" uint32_t gkr_so_num_polys(const gkr_so* so) { return so ? so->num_polys() : 0; }
extern "C" uint32_t gkr_so_round(const gkr_so* so) { return so ? so->round() : 0; }
extern "C" void gkr_so_destroy(gkr_so* so) { delete so; }

// ---- transcript ----------------------------------------------------------------------------------------
extern "C" int gkr_transcript_new(const uint8_t* label, size_t label_len, gkr_transcript** out) {
    if (!out || (!label && label_len)) return GKR_ERR_ARG;
    *out = new gkr_transcript(label, label_len);
    return GKR_OK;
}
extern "C" void gkr_transcript_free(gkr_transcript* t) { delete t; }

extern "C" int gkr_transcript_write_scalars(gkr_transcript* t, const uint64_t* limbs, uint32_t n) {
    if (!t || (!limbs && n)) return GKR_ERR_ARG;
    std::vector<gkr::FrH> v = load_frs(limbs, n);
    for (auto& x : v)
        if (!frh_canonical(x)) return GKR_ERR_ARG;
    t->t.write_scalars(v.data(), v.size());
    return GKR_OK;
}

extern "C" int gkr_transcript_write_raw(gkr_transcript* t, const uint8_t* msg, size_t len) {
    if (!t || (!msg && len)) return GKR_ERR_ARG;
    t->t.write_raw_msg(msg, len);
    return GKR_OK;
}

extern "C" int gkr_transcript_challenge(gkr_transcript* t, uint32_t bitsize, uint64_t out[4]) {
    if (!t || !out || bitsize == 0 || bitsize > 512) return GKR_ERR_ARG;
    frh_to_limbs(t->t.challenge(bitsize), out);
    return GKR_OK;
}

extern "C" int gkr_transcript_raw_challenge(gkr_transcript* t, uint8_t* out, size_t len) {
    if (!t || (!out && len)) return GKR_ERR_ARG;
    t->t.raw_challenge(out, len);
    return GKR_OK;
}

// old API transcript (src/transcript.rs:78-101, `impl TranscriptReceiver / TranscriptSender for merlin::Transcript`)
extern "C" int gkr_transcript_append_message(gkr_transcript* t, const uint8_t* label, size_t label_len, const uint8_t* msg, size_t len) {
    if (!t || (!label && label_len) || (!msg && len)) return GKR_ERR_ARG;
    t->t.append_labeled(label, label_len, msg, len);
    return GKR_OK;
}
extern "C" int gkr_transcript_append_scalars_old(gkr_transcript* t, const uint64_t* limbs, uint32_t n) {
    if (!t || (!limbs && n)) return GKR_ERR_ARG;
    std::vector<gkr::FrH> v = load_frs(limbs, n);
    for (auto& x : v) {  // append_scalars: one message per scalar, label b"" (transcript.rs:78-89)
        if (!frh_canonical(x)) return GKR_ERR_ARG;
        uint8_t buf[32];
        gkr::frh::to_bytes_le(x, buf);
        t->t.append_labeled(nullptr, 0, buf, 32);
    }
    return GKR_OK;
}
extern "C" int gkr_transcript_challenge_scalar_old(gkr_transcript* t, const uint8_t* label, size_t label_len, uint64_t out[4]) {
    if (!t || !out || (!label && label_len)) return GKR_ERR_ARG;
    uint8_t buf[64];  // challenge_scalar: 64 bytes -> from_le_bytes_mod_order (transcript.rs:96-101)
    t->t.challenge_labeled(label, label_len, buf, 64);
    frh_to_limbs(gkr::frh::from_le_bytes_mod_order(buf, 64), out);
    return GKR_OK;
}

extern "C" size_t gkr_transcript_proof_len(const gkr_transcript* t) { return t ? t->t.proof.size() : 0; }

extern "C" int gkr_transcript_proof(const gkr_transcript* t, uint8_t* out) {
    if (!t || !out) return GKR_ERR_ARG;
    std::memcpy(out, t->t.proof.data(), t->t.proof.size());
    return GKR_OK;
}

// GenericSumcheckProtocol::prove  (src/cleanup/protocols/sumcheck.rs:101-123)
extern "C" int gkr_so_set_prelaunch(gkr_so* so, int on) {
    if (!so) return GKR_ERR_ARG;
    so->set_prelaunch(on != 0);
    return GKR_OK;
}

// GKR_TRACE: where the host time of the round loop goes (printed by gkr_sumcheck_prove_stats_dump)
static std::atomic<uint64_t> g_sc_ns[4];  // unipoly (launch-to-result wait included), interpolation + transcript, bind, final_evals
static std::atomic<uint64_t> g_sc_rounds{0};
extern "C" void gkr_sumcheck_prove_stats_dump(void) {
    if (!g_sc_rounds) return;
    fprintf(stderr, "  [gkr_sumcheck_prove] %llu rounds: unipoly %.2f ms, interpolate+transcript %.2f ms, bind %.2f ms, final_evals %.2f ms\n",
            (unsigned long long)g_sc_rounds.load(), g_sc_ns[0].load() / 1e6, g_sc_ns[1].load() / 1e6, g_sc_ns[2].load() / 1e6, g_sc_ns[3].load() / 1e6);
    g_sc_rounds = 0;
    for (int i = 0; i < 4; i++) g_sc_ns[i] = 0;
}

extern "C" int gkr_sumcheck_prove(gkr_transcript* t, gkr_so* so, uint32_t num_rounds, uint64_t out_claim[4],
                                  uint64_t* out_point, uint64_t* out_final_evals) {
    if (!t || !so) return GKR_ERR_ARG;
    static const bool trace = getenv("GKR_TRACE") != nullptr;
    gkr::FrH claim = so->claim();
    std::vector<gkr::FrH> r;
    r.reserve(num_rounds);
    // this loop is the strict unipoly -> bind alternation with nothing else on the stream: the object may pre-launch its small rounds
    struct Prelaunch {
        gkr_so* so;
        explicit Prelaunch(gkr_so* s) : so(s) { so->set_prelaunch(true); }
        ~Prelaunch() { so->set_prelaunch(false); }
    } prelaunch_guard(so);
    uint64_t t0 = trace ? gkr_now_ns() : 0;
    for (uint32_t k = 0; k < num_rounds; k++) {
        gkr::FrH ev[GKR_MAX_DEG + 1];
        uint32_t n = 0;
        int rc = so->unipoly(ev, &n);
        if (rc) return rc;
        uint64_t t1 = trace ? gkr_now_ns() : 0;
        gkr::FrH poly[GKR_MAX_DEG + 1], msg[GKR_MAX_DEG + 1];
        gkr::frh::interpolate_coeffs_into(ev, (int)n, poly);  // unipoly().as_vec()
        uint32_t nm = 0;                                       // compress_coefficients: drop the linear term
        msg[nm++] = poly[0];
        for (uint32_t i = 2; i < n; i++) msg[nm++] = poly[i];
        t->t.write_scalars(msg, nm);
        gkr::FrH x = t->t.challenge(128);
        r.push_back(x);
        uint64_t t2 = trace ? gkr_now_ns() : 0;
        rc = so->bind(x);
        if (rc) return rc;
        gkr::FrH c = gkr::frh::ZERO;  // evaluate_univar(poly, x)
        for (uint32_t i = n; i-- > 0;) c = gkr::frh::add(gkr::frh::mul(c, x), poly[i]);
        claim = c;
        if (trace) {
            uint64_t t3 = gkr_now_ns();
            g_sc_ns[0] += t1 - t0;
            g_sc_ns[1] += t2 - t1;
            g_sc_ns[2] += t3 - t2;
            g_sc_rounds++;
            t0 = t3;
        }
    }
    if (out_claim) frh_to_limbs(claim, out_claim);
    if (out_point)
        for (uint32_t k = 0; k < num_rounds; k++) frh_to_limbs(r[num_rounds - 1 - k], out_point + 4 * k);  // r.reverse()
    if (out_final_evals) {
        std::vector<gkr::FrH> fe(so->num_polys());
        int rc = so->final_evals(fe.data());
        if (rc) return rc;
        for (size_t i = 0; i < fe.size(); i++) frh_to_limbs(fe[i], out_final_evals + 4 * i);
        if (trace) g_sc_ns[3] += gkr_now_ns() - t0;
    }
    return GKR_OK;
}

// ---- PushForwardState::new, index bookkeeping (pushforward.rs:351-396) --------------------------------------------
// digits[y][x] = (coef_x >> (y d)) & (2^d - 1); counter[y][x] = rank of x inside its bucket (input order);
// order[y] = the x of row y sorted by digit (stable) = the bucket contents back to back; lens[y][b] = bucket sizes.
// Pure host work (a counting sort per digit row, one thread per row) -- no field arithmetic, nothing for the device.
#include <thread>
extern "C" int gkr_pushforward_bucketize(const uint64_t* coefs, uint64_t n, uint32_t y_size, uint32_t d_logsize, uint32_t* digits,
                                         uint32_t* counter, uint32_t* order, uint32_t* lens) {
    if (!coefs || !digits || !counter || !order || !lens || d_logsize == 0 || d_logsize > 24 || (uint64_t)y_size * d_logsize > 256 ||
        n >= ((uint64_t)1 << 32))
        return GKR_ERR_ARG;
    const uint32_t nb = 1u << d_logsize, mask = nb - 1;
    auto do_row = [&](uint32_t y) {
        uint32_t* dg = digits + (size_t)y * n;
        uint32_t* ct = counter + (size_t)y * n;
        uint32_t* od = order + (size_t)y * n;
        uint32_t* ln = lens + (size_t)y * nb;
        const uint32_t bit = y * d_logsize, limb = bit >> 6, sh = bit & 63;
        std::vector<uint32_t> cnt(nb, 0), off(nb, 0);
        for (uint64_t x = 0; x < n; x++) {
            const uint64_t* c = coefs + 4 * x;
            uint64_t v = c[limb] >> sh;
            if (sh && sh + d_logsize > 64 && limb + 1 < 4) v |= c[limb + 1] << (64 - sh);
            const uint32_t d = (uint32_t)v & mask;
            dg[x] = d;
            ct[x] = cnt[d]++;
        }
        uint32_t acc = 0;
        for (uint32_t b = 0; b < nb; b++) {
            ln[b] = cnt[b];
            off[b] = acc;
            acc += cnt[b];
        }
        for (uint64_t x = 0; x < n; x++) od[off[dg[x]] + ct[x]] = (uint32_t)x;
    };
    unsigned hw = std::thread::hardware_concurrency();
    unsigned nt = std::max(1u, std::min(hw ? hw : 4u, y_size));
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
        th.emplace_back([&, t]() {
            for (uint32_t y = t; y < y_size; y += nt) do_row(y);
        });
    for (auto& t : th) t.join();
    return GKR_OK;
}


// host only: the 48-byte compressed G1 encoding the prover writes into the proof (ark-bls12-381 0.4 = zcash / IETF format,
// proof_transcript.rs:52-69) of an affine point given as 12 u64 Montgomery limbs (all zero = infinity).  Exposed so that the
// wire format of the C++ host layer can be pinned against the published encoding of the generator without a device.
extern "C" int gkr_host_g1_serialize(const uint64_t* xy, uint8_t* out48) {
    if (!xy || !out48) return GKR_ERR_ARG;
    gkr::g1h::serialize_compressed(xy, out48);
    return GKR_OK;
}
