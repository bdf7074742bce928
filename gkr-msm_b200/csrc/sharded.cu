// Multi-GPU sharding of the dense sumcheck.
//
// Folding binds the LEAST significant index bit first (pairs 2i, 2i+1 -- sumcheck.rs:160-163), so splitting
// the hypercube by its TOP log2(G) index bits gives every GPU a contiguous slice that stays independent for
// the first n - log2(G) rounds: one process per GPU, each runs the same fused fold+eval kernel on its slice.
// The only exchange is the per-round partial sums (deg field elements per rank).  They have to reach the
// HOST anyway -- Fiat-Shamir stays on the host -- so the ranks of one box exchange them through a small
// POSIX shared-memory segment (sequence-numbered slots, ~1 us) instead of a device collective: no kernel,
// no NVLink traffic and no extra launch sits on the per-round critical path.  Every rank then runs the
// identical transcript and derives the identical challenge.  After the local rounds each rank holds one
// value per table; these G x P values are all-gathered the same way and the last log2(G) rounds run
// replicated on every rank.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <atomic>
#include <cstring>
#include "common.cuh"
#include "host_gates.hpp"
#include "so.hpp"
#include "transcript.hpp"

#define GKR_EX_MAX_RANKS 16
#define GKR_EX_MAX_ELEMS 32

struct ExShared {
    std::atomic<uint64_t> seq[GKR_EX_MAX_RANKS];
    uint64_t pad[8];
    uint64_t data[2][GKR_EX_MAX_RANKS][GKR_EX_MAX_ELEMS * 4];
};

struct gkr_exchange {
    ExShared* sh = nullptr;
    int rank = 0, world = 1;
    uint64_t counter = 0;
    std::string name;
    bool creator = false;
};

extern "C" int gkr_exchange_open(const char* name, int rank, int world, int create, gkr_exchange** out) {
    if (!name || !out || world < 1 || world > GKR_EX_MAX_RANKS || rank < 0 || rank >= world) return GKR_ERR_ARG;
    int fd = shm_open(name, create ? (O_CREAT | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) return GKR_ERR_ARG;
    if (create && ftruncate(fd, sizeof(ExShared)) != 0) {
        close(fd);
        return GKR_ERR_ARG;
    }
    if (!create) {  // opened before the creator's ftruncate: touching the mapping would SIGBUS -- report "not ready" instead
        struct stat st;
        if (fstat(fd, &st) != 0 || (size_t)st.st_size < sizeof(ExShared)) {
            close(fd);
            return GKR_ERR_ARG;
        }
    }
    void* p = mmap(nullptr, sizeof(ExShared), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return GKR_ERR_ARG;
    gkr_exchange* ex = new gkr_exchange();
    ex->sh = (ExShared*)p;
    ex->rank = rank;
    ex->world = world;
    ex->name = name;
    ex->creator = create != 0;
    if (create) std::memset(p, 0, sizeof(ExShared));
    *out = ex;
    return GKR_OK;
}

extern "C" int gkr_exchange_world(const gkr_exchange* ex) { return ex ? ex->world : 1; }

extern "C" void gkr_exchange_close(gkr_exchange* ex) {
    if (!ex) return;
    munmap(ex->sh, sizeof(ExShared));
    if (ex->creator) shm_unlink(ex->name.c_str());
    delete ex;
}

// all-gather of n_elems field elements per rank; `all` receives world * n_elems elements ordered by rank
extern "C" int gkr_exchange_allgather(gkr_exchange* ex, const uint64_t* mine, uint32_t n_elems, uint64_t* all) {
    if (!ex || !mine || !all || n_elems > GKR_EX_MAX_ELEMS) return GKR_ERR_ARG;
    const uint64_t c = ++ex->counter;
    const int par = (int)(c & 1);
    std::memcpy(ex->sh->data[par][ex->rank], mine, sizeof(uint64_t) * 4 * n_elems);
    ex->sh->seq[ex->rank].store(c, std::memory_order_release);
    for (int r = 0; r < ex->world; r++) {
        uint64_t spins = 0;
        while (ex->sh->seq[r].load(std::memory_order_acquire) < c) {
            if (++spins > (1ull << 34)) return GKR_ERR_PROTOCOL;  // a peer died
        }
        std::memcpy(all + (size_t)r * 4 * n_elems, ex->sh->data[par][r], sizeof(uint64_t) * 4 * n_elems);
    }
    return GKR_OK;
}

// GenericSumcheckProtocol::prove (sumcheck.rs:101-123) over a hypercube sharded by its top index bits.
//   so: this rank's DenseSumcheckObjectSO over its slice (local_rounds variables)
//   global_claim: claim of the whole sum.  out_point: local_rounds + log2(world) challenges, reversed (:120).
//   out_final_evals: n_polys elements (identical on every rank).
extern "C" int gkr_sumcheck_prove_sharded(gkr_transcript* t, gkr_so* so, gkr_exchange* ex, uint32_t local_rounds,
                                          int so_kind, int gate, uint32_t gate_param, const uint64_t* gate_consts, uint32_t n_consts,
                                          const uint64_t global_claim[4], uint64_t out_claim[4], uint64_t* out_point,
                                          uint64_t* out_final_evals) {
    if (!t || !so || !global_claim) return GKR_ERR_ARG;
    gkr_ctx* ctx = so->ctx;
    const int world = ex ? ex->world : 1;
    int g = 0;
    while ((1 << g) < world) g++;
    if ((1 << g) != world) return ctx->fail(GKR_ERR_ARG, "world size must be a power of two");
    const uint32_t deg = so->degree(), P = so->num_polys();
    gkr::FrH claim = frh_from_limbs(global_claim);
    std::vector<gkr::FrH> r;
    so->set_prelaunch(true);  // strict partial_sums -> bind alternation below (so.hpp)
    struct PrelaunchOff {
        gkr_so* so;
        ~PrelaunchOff() { so->set_prelaunch(false); }
    } prelaunch_guard{so};
    auto round_io = [&](const gkr::FrH* sums_total) {  // sums at nodes 1..deg of the WHOLE hypercube
        gkr::FrH ev[GKR_MAX_DEG + 1];
        for (uint32_t s = 0; s < deg; s++) ev[s + 1] = sums_total[s];
        ev[0] = gkr::frh::sub(claim, ev[1]);
        std::vector<gkr::FrH> poly = gkr::frh::interpolate_coeffs(ev, (int)deg + 1);
        std::vector<gkr::FrH> msg;
        msg.push_back(poly[0]);
        for (size_t i = 2; i < poly.size(); i++) msg.push_back(poly[i]);
        t->t.write_scalars(msg.data(), msg.size());
        gkr::FrH x = t->t.challenge(128);
        r.push_back(x);
        claim = gkr::frh::evaluate_univar(poly, x);
        return x;
    };
    std::vector<uint64_t> mine(4 * GKR_EX_MAX_ELEMS), all((size_t)4 * GKR_EX_MAX_ELEMS * GKR_EX_MAX_RANKS);
    for (uint32_t k = 0; k < local_rounds; k++) {
        gkr::FrH ev[GKR_MAX_DEG + 1];
        uint32_t n = 0;
        int rc = so->unipoly(ev, &n);  // ev[1..deg] are this shard's partial sums (ev[0] is not used here)
        if (rc) return rc;
        gkr::FrH tot[GKR_MAX_DEG];
        if (world > 1) {
            for (uint32_t s = 0; s < deg; s++) frh_to_limbs(ev[s + 1], mine.data() + 4 * s);
            rc = gkr_exchange_allgather(ex, mine.data(), deg, all.data());
            if (rc) return ctx->fail(rc, "partial-sum exchange failed");
            for (uint32_t s = 0; s < deg; s++) {
                tot[s] = gkr::frh::ZERO;
                for (int q = 0; q < world; q++) tot[s] = gkr::frh::add(tot[s], frh_from_limbs(all.data() + ((size_t)q * deg + s) * 4));
            }
        } else {
            for (uint32_t s = 0; s < deg; s++) tot[s] = ev[s + 1];
        }
        gkr::FrH x = round_io(tot);
        rc = so->bind(x);
        if (rc) return rc;
    }
    std::vector<gkr::FrH> fe(P);
    int rc = so->final_evals(fe.data());
    if (rc) return rc;
    if (world > 1) {
        // gather the G x P surviving values; table j over the remaining g variables is [rank 0, rank 1, ...]
        for (uint32_t j = 0; j < P; j++) frh_to_limbs(fe[j], mine.data() + 4 * j);
        rc = gkr_exchange_allgather(ex, mine.data(), P, all.data());
        if (rc) return ctx->fail(rc, "final gather failed");
        // The last log2(world) rounds run over world x P gathered values on the HOST of every rank -- the same arithmetic as the
        // device object (sumcheck.rs:277-332, 160-163) for whatever single-output gate the object wraps, without uploads, a
        // second object and log2(world) more launches.
        std::vector<gkr::FrH> consts(n_consts);
        for (uint32_t i = 0; i < n_consts; i++) consts[i] = frh_from_limbs(gate_consts + 4 * i);
        int gi = 0, go = 0;
        const bool eq_gamma = so_kind == GKR_SO_EQ_GAMMA;
        if (eq_gamma && (!gkr::base_gate_io(gate, &gi, &go) || (uint32_t)gi + 1 != P || (go > 1 && n_consts < (uint32_t)go)))
            return ctx->fail(GKR_ERR_ARG, "sharded sumcheck: gate / constants do not match the object");
        if (!eq_gamma && !(gate == GKR_GATE_PROD3 && P == 3) && !(gate == GKR_GATE_FOLDED_PROD && P == 2 * gate_param && n_consts >= gate_param))
            return ctx->fail(GKR_ERR_UNSUPPORTED, "sharded sumcheck: unsupported single-output gate");
        auto f = [&](const gkr::FrH* a) {  // the object's single-output function on one point
            using namespace gkr::frh;
            if (!eq_gamma) {
                if (gate == GKR_GATE_PROD3) return mul(mul(a[0], a[1]), a[2]);
                gkr::FrH s = ZERO;  // FoldedProdAlgFn, multiopen_reduction.rs:28-32
                for (uint32_t k = 0; k < gate_param; k++) s = add(s, mul(mul(a[k], a[k + gate_param]), consts[k]));
                return s;
            }
            gkr::FrH o[16];
            gkr::base_gate_eval(gate, a, o);
            gkr::FrH s = o[0];  // GammaWrapper: out_0 + sum_i gamma^i out_i, then EqWrapper: times the eq value
            for (int k = 1; k < go; k++) s = add(s, mul(o[k], consts[k]));
            return mul(s, a[gi]);
        };
        std::vector<std::vector<gkr::FrH>> tb(P, std::vector<gkr::FrH>(world));
        for (uint32_t jj = 0; jj < P; jj++)
            for (int q = 0; q < world; q++) tb[jj][q] = frh_from_limbs(all.data() + ((size_t)q * P + jj) * 4);
        std::vector<gkr::FrH> a(P), d(P);
        for (int k = 0; k < g; k++) {
            const size_t half = tb[0].size() / 2;
            gkr::FrH sums[GKR_MAX_DEG];
            for (uint32_t sidx = 0; sidx < deg; sidx++) sums[sidx] = gkr::frh::ZERO;
            for (size_t i = 0; i < half; i++) {
                for (uint32_t jj = 0; jj < P; jj++) {
                    a[jj] = tb[jj][2 * i + 1];
                    d[jj] = gkr::frh::sub(tb[jj][2 * i + 1], tb[jj][2 * i]);
                }
                for (uint32_t sidx = 0; sidx < deg; sidx++) {
                    if (sidx)
                        for (uint32_t jj = 0; jj < P; jj++) a[jj] = gkr::frh::add(a[jj], d[jj]);
                    sums[sidx] = gkr::frh::add(sums[sidx], f(a.data()));
                }
            }
            gkr::FrH x = round_io(sums);
            for (uint32_t jj = 0; jj < P; jj++) {
                for (size_t i = 0; i < half; i++)
                    tb[jj][i] = gkr::frh::add(tb[jj][2 * i], gkr::frh::mul(x, gkr::frh::sub(tb[jj][2 * i + 1], tb[jj][2 * i])));
                tb[jj].resize(half);
            }
        }
        for (uint32_t jj = 0; jj < P; jj++) fe[jj] = tb[jj][0];
    }
    if (out_claim) frh_to_limbs(claim, out_claim);
    if (out_point)
        for (size_t k = 0; k < r.size(); k++) frh_to_limbs(r[r.size() - 1 - k], out_point + 4 * k);
    if (out_final_evals)
        for (uint32_t j = 0; j < P; j++) frh_to_limbs(fe[j], out_final_evals + 4 * j);
    return GKR_OK;
}
