/* TEST INFRASTRUCTURE ONLY -- CPU oracle + CPU baseline of the WHOLE `examples/pippenger` prover.
 * Only tests/, tests/golden/make_golden_large.py, __graft_entry__.smoke() and bench.py's CPU legs may load this library; it is
 * never linked into, imported by, or called from the product path (gkr-msm_b200/).
 *
 * benchutils::run_pippenger (src/cleanup/protocols/pippenger.rs:499-559) restated in C++ / OpenMP, written from the reference
 * sources file by file -- NOT from the product's host orchestration (gkr-msm_b200/csrc/protocol.cu) -- so that a device proof
 * can be byte-compared against an independent prover at sizes the python oracle (oracle/pyref) cannot reach:
 *   PushForwardState::{new, second_phase}      src/cleanup/protocols/pushforward/pushforward.rs:329-622
 *   PushforwardProtocol::prove                 src/cleanup/protocols/pushforward/pushforward.rs:641-847
 *   LogupMainphaseProtocol::{make_witness,prove}  src/cleanup/protocols/pushforward/logup_mainphase.rs:85-200
 *   EqTruncPoly / SelectorPoly                 src/cleanup/protocols/verifier_polys.rs:68-137, src/utils.rs:265-291
 *   MultiOpenReduction::prove                  src/cleanup/protocols/multiopen_reduction.rs:65-93
 *   KnucklesOpeningProtocol::prove             src/cleanup/protocols/opening.rs:39-98
 *   KnucklesProvingKey::{new, compute_t}       src/commitments/knuckles.rs:65-81, 111-154
 *   KzgProvingKey::{mock_setup, commit, open}, div_by_linear, ev, verify_reduce_to_pair   src/commitments/kzg.rs:49-150
 *   PippengerWG::new, Pippenger::prove         src/cleanup/protocols/pippenger.rs:36-70, 122-294
 * (sumcheck engine: po_sumcheck.hpp; GKR circuits: po_gkr.hpp; transcript: po_transcript.hpp; fields / G1: po_field.hpp, po_g1.hpp).
 *
 * Differences from the reference, none of which can change a proof byte: commitments use this file's own bucket MSM (group
 * elements are unique); the bintree witness is built once instead of twice (pippenger_ending.rs:40-45,67-72); the unused
 * `BandersnatchConfig::msm` inside the timed region (pippenger.rs:516) is skipped; loops the reference leaves serial are
 * OpenMP-parallel.  All four make this CPU baseline FASTER than a faithful port would be.
 *
 * PARITY: pinned bit-for-bit against oracle/pyref on the four committed golden proofs and on live small instances
 * (tests/test_pippenger_oracle.py); the reference itself holds no golden vectors and cannot be built in this image
 * (nightly Rust + un-vendored git dependencies), so byte parity with a real reference run stays "parity unpinned".
 */
#include <array>
#include <chrono>
#include "po_gkr.hpp"

namespace po {

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

/* ---- commitment keys ------------------------------------------------------------------------------------------------- */
struct KnucklesKey {  // KzgProvingKey (mock setup) + KnucklesProvingKey
    std::vector<G1A> ptau;  // ptau_1[i] = tau^i g0
    G1A g0;
    size_t num_vars;
    Fr k;
    Vec inverses;
    KnucklesKey(const Fr& tau, const G1A& g0_, size_t nv, const Fr& k_) : g0(g0_), num_vars(nv), k(k_) {
        const size_t n = (size_t)1 << nv;
        ptau = g1_powers_of_tau(tau, g0, 2 * n - 1);
        // KnucklesProvingKey::new, knuckles.rs:65-81
        Vec k_pows(2 * n - 1);
        Fr power = Fr::one();
        for (size_t i = 0; i < 2 * n - 1; i++) {
            k_pows[i] = power;
            power *= k;
        }
        const Fr k_n = k_pows[n - 1];
        for (auto& x : k_pows) x -= k_n;
        k_pows[n - 1] += Fr::one();
        batch_inverse(k_pows.data(), k_pows.size());
        inverses = std::move(k_pows);
    }
    G1A commit(const Vec& poly) const {  // kzg.rs:123-126
        PO_ASSERT(poly.size() <= ptau.size(), "Vector is too large.");
        return g1_to_affine(g1_msm(ptau.data(), poly.data(), poly.size()));
    }
};
static inline void div_by_linear(const Vec& poly, const Fr& pt, Vec* quotient, Fr* rem_out) {  // kzg.rs:73-81
    quotient->assign(poly.size() - 1, Fr::zero());
    Fr rem = poly.back();
    for (size_t i = quotient->size(); i-- > 0;) {
        (*quotient)[i] = rem;
        rem = poly[i] + rem * pt;
    }
    *rem_out = rem;
}
static inline Fr ev(const Vec& poly, const Fr& x) {  // kzg.rs:142-150
    Fr power = Fr::one(), acc = Fr::zero();
    for (size_t i = 0; i < poly.size(); i++) {
        acc += poly[i] * power;
        power *= x;
    }
    return acc;
}
static inline G1A kzg_open(const KnucklesKey& key, const Vec& poly, const Fr& pt, Fr* rem) {  // kzg.rs:129-132
    Vec q;
    div_by_linear(poly, pt, &q, rem);
    return key.commit(q);
}
static inline void compute_t(const KnucklesKey& key, const Vec& poly, const Vec& point, Vec* t_out, Fr* opening) {  // knuckles.rs:111-154
    PO_ASSERT(point.size() == key.num_vars, "compute_t: point length");
    const size_t nv = key.num_vars, n = (size_t)1 << nv;
    PO_ASSERT(poly.size() <= n, "compute_t: poly too long");
    Vec pt(point.rbegin(), point.rend());
    Vec t(2 * n - 1, Fr::zero()), t_scaled(2 * n - 1, Fr::zero());
    for (size_t i = 0; i < poly.size(); i++) t[i] = poly[i];
    size_t curr = n;
    for (size_t i = 0; i < nv; i++) {
        const Fr pr = Fr::one() - pt[i];
#pragma omp parallel for schedule(static) if (curr >= 4096)
        for (size_t idx = 0; idx < curr; idx++) t_scaled[idx] = t[idx] * pr;
        const size_t offset = (size_t)1 << i;
        curr += offset;
#pragma omp parallel for schedule(static) if (curr >= 4096)
        for (size_t idx = 0; idx < curr; idx++) {
            if (idx < offset) t[idx] -= t_scaled[idx];
            else t[idx] = t[idx] - t_scaled[idx] + t_scaled[idx - offset];
        }
    }
    *opening = t[n - 1];
    t[n - 1] = Fr::zero();
#pragma omp parallel for schedule(static) if (t.size() >= 4096)
    for (size_t idx = 0; idx < t.size(); idx++) t[idx] *= key.inverses[idx];
    *t_out = std::move(t);
}

/* ---- verifier polys (verifier_polys.rs) ---------------------------------------------------------------------------------- */
static inline Fr eq_sum(const Vec& pt, size_t k) {  // src/utils.rs:265-291 (SelectorPoly::evaluate)
    const size_t n = pt.size();
    if (k >= ((size_t)1 << n)) {
        PO_ASSERT(k == ((size_t)1 << n), "eq_sum: k out of range");
        return Fr::one();
    }
    Fr mult = Fr::one(), acc = Fr::zero();
    for (size_t i = 0; i < n; i++) {
        const size_t left_bit = k >> (n - i - 1);
        const Fr old = mult;
        if (left_bit == 1) {
            mult *= pt[i];
            acc += old - mult;
        } else {
            mult *= Fr::one() - pt[i];
        }
        k -= left_bit << (n - i - 1);
    }
    return acc;
}
static inline Vec eq_trunc_evals(size_t num_vars, size_t k, const Vec& r) {  // verifier_polys.rs:90-96
    Vec ret = eq_poly_last(r);
    for (size_t i = k; i < ((size_t)1 << num_vars); i++) ret[i] = Fr::zero();
    return ret;
}
static inline Fr eq_trunc_evaluate(size_t num_vars, size_t k, const Vec& r, const Vec& pt) {  // verifier_polys.rs:98-136
    PO_ASSERT(pt.size() == num_vars, "EqTruncPoly::evaluate: point length");
    Vec partial{Fr::one()};
    for (size_t i = 0; i < num_vars; i++) {
        const size_t j = num_vars - i - 1;
        partial.push_back(partial.back() * (Fr::one() - pt[j] - r[j] + (r[j] * pt[j]).dbl()));
    }
    if (k >= ((size_t)1 << num_vars)) return partial[num_vars];
    Fr multiplier = Fr::one(), acc = Fr::zero();
    for (size_t i = 0; i < num_vars; i++) {
        const size_t left_bit = k >> (num_vars - i - 1);
        const Fr m_ = multiplier;
        if (left_bit == 1) {
            multiplier = multiplier * pt[i] * r[i];
            acc += m_ * (Fr::one() - pt[i]) * (Fr::one() - r[i]) * partial[num_vars - i - 1];
        } else {
            multiplier = multiplier * (Fr::one() - pt[i]) * (Fr::one() - r[i]);
        }
        k -= left_bit << (num_vars - i - 1);
    }
    return acc;
}
static inline void pad_vector(Vec& v, size_t logsize, const Fr& with) {  // src/utils.rs:324-329
    PO_ASSERT(v.size() <= ((size_t)1 << logsize), "pad_vector: too long");
    v.resize((size_t)1 << logsize, with);
}

/* ---- PushForwardState (pushforward.rs:329-622) ----------------------------------------------------------------------------- */
struct PushForwardState {
    size_t y_size, y_logsize, d_logsize, x_logsize, x_size, clm;
    const KnucklesKey* key;
    std::vector<std::vector<uint32_t>> digits, counter;
    std::vector<VecVec> image;
    Vec c, d, p_0, p_1, ac_c, ac_d, c_pull, d_pull;
    std::vector<G1A> c_comm, d_comm, c_pull_comm, d_pull_comm;
    G1A p_0_comm, p_1_comm, ac_c_comm, ac_d_comm;
    std::vector<std::vector<G1J>> d_outer_buckets, c_outer_buckets;  // per commitment chunk

    /* coefs: canonical little-endian integers (the Bandersnatch scalars), 4 limbs each */
    PushForwardState(const Vec& px, const Vec& py, const uint64_t* coefs, size_t y_size_, size_t yl, size_t dl, size_t xl, size_t clm_, const KnucklesKey* key_)
        : y_size(y_size_), y_logsize(yl), d_logsize(dl), x_logsize(xl), x_size((size_t)1 << xl), clm(clm_), key(key_) {
        PO_ASSERT(key->num_vars == xl + clm, "commitment key: num_vars != x_logsize + commitment_log_multiplicity");
        PO_ASSERT(px.size() == x_size && py.size() == x_size, "points.len() != 1 << x_logsize");
        PO_ASSERT(y_size * dl <= 256, "y_size * d_logsize > 256 (to_bits_le() index out of range in the reference)");
        PO_ASSERT(((size_t)1 << yl) >= y_size, "1 << y_logsize < y_size");
        const Fr one = Fr::one(), zero = Fr::zero();
        std::vector<Vec> polys{px, py, Vec(x_size, one)};
        digits.assign(y_size, std::vector<uint32_t>(x_size, 0));
        for (size_t x = 0; x < x_size; x++)
            for (size_t y = 0; y < y_size; y++) {
                uint32_t dg = 0;
                for (size_t i = 0; i < dl; i++) {
                    const size_t bit = y * dl + i;
                    dg += (uint32_t)((coefs[4 * x + bit / 64] >> (bit % 64)) & 1) << i;
                }
                digits[y][x] = dg;
            }
        const Fr row_pad[3] = {zero, one, zero}, col_pad[3] = {zero, one, zero};
        counter.assign(y_size, std::vector<uint32_t>(x_size, 0));
        const size_t n_buckets = y_size << dl;
        std::vector<std::vector<Vec>> buckets(3, std::vector<Vec>(n_buckets));  // [poly][bucket]
        const size_t comm_mul = (size_t)1 << clm, n_comms = (y_size + comm_mul - 1) / comm_mul;
        std::vector<std::vector<G1J>> d_outer(y_size), c_outer(y_size);
        std::vector<size_t> c_upper(y_size);
#pragma omp parallel for schedule(dynamic, 1)
        for (size_t y = 0; y < y_size; y++) {  // :401-429 (rayon over digit rows)
            d_outer[y].assign((size_t)1 << dl, G1J::infinity());
            c_outer[y].assign(x_size, G1J::infinity());
            size_t max_c = 0;
            for (size_t x = 0; x < x_size; x++) {
                const size_t dg = digits[y][x];
                const size_t b = (y << dl) + dg;
                const size_t cc = buckets[0][b].size();
                max_c = std::max(max_c, cc);
                const G1A& point = key->ptau[x + x_size * (y % comm_mul)];
                d_outer[y][dg] = g1_madd(d_outer[y][dg], point);
                c_outer[y][cc] = g1_madd(c_outer[y][cc], point);
                counter[y][x] = (uint32_t)cc;
                for (int pid = 0; pid < 3; pid++) buckets[pid][b].push_back(polys[pid][x]);
            }
            c_upper[y] = max_c + 1;
        }
        for (size_t k = 0; k < n_comms; k++) {  // :433-456: merge the rows of one commitment chunk
            const size_t y0 = k * comm_mul, y1 = std::min(y_size, y0 + comm_mul);
            size_t max_c = 0;
            for (size_t y = y0; y < y1; y++) max_c = std::max(max_c, c_upper[y]);
            std::vector<G1J> dd((size_t)1 << dl), cc(max_c);
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < dd.size(); i++) {
                G1J acc = d_outer[y0][i];
                for (size_t y = y0 + 1; y < y1; y++) acc = g1_add(acc, d_outer[y][i]);
                dd[i] = acc;
            }
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < max_c; i++) {
                G1J acc = c_outer[y0][i];
                for (size_t y = y0 + 1; y < y1; y++) acc = g1_add(acc, c_outer[y][i]);
                cc[i] = acc;
            }
            d_outer_buckets.push_back(std::move(dd));
            c_outer_buckets.push_back(std::move(cc));
        }
        d_outer.clear();
        c_outer.clear();
        for (int pid = 0; pid < 3; pid++) image.emplace_back(std::move(buckets[pid]), row_pad[pid], col_pad[pid], xl, yl + dl);
        c.resize(y_size * x_size);
        d.resize(y_size * x_size);
        std::vector<uint64_t> cnt_d((size_t)1 << dl, 0), cnt_c(x_size, 0);
        for (size_t y = 0; y < y_size; y++)
            for (size_t x = 0; x < x_size; x++) {
                d[y * x_size + x] = Fr::from_u64(digits[y][x]);
                c[y * x_size + x] = Fr::from_u64(counter[y][x]);
                cnt_d[digits[y][x]]++;
                cnt_c[counter[y][x]]++;
            }
        for (auto v : cnt_c) ac_c.push_back(-Fr::from_u64(v));
        for (auto v : cnt_d) ac_d.push_back(-Fr::from_u64(v));
        p_0 = px;
        p_1 = py;
        auto running_sum_commit = [](const std::vector<G1J>& b) {  // :504-524
            G1J acc = G1J::infinity(), running = G1J::infinity();
            const size_t len = b.size();
            for (size_t i = 0; i + 1 < len; i++) {
                running = g1_add(running, b[len - i - 1]);
                acc = g1_add(acc, running);
            }
            return acc;
        };
        std::vector<G1J> dj(n_comms), cj(n_comms);
#pragma omp parallel for schedule(dynamic, 1)
        for (size_t k = 0; k < 2 * n_comms; k++) {
            if (k < n_comms) dj[k] = running_sum_commit(d_outer_buckets[k]);
            else cj[k - n_comms] = running_sum_commit(c_outer_buckets[k - n_comms]);
        }
        d_comm.resize(n_comms);
        c_comm.resize(n_comms);
        g1_batch_to_affine(dj.data(), d_comm.data(), n_comms);
        g1_batch_to_affine(cj.data(), c_comm.data(), n_comms);
        p_0_comm = key->commit(p_0);  // :534-537
        p_1_comm = key->commit(p_1);
        ac_c_comm = key->commit(ac_c);
        ac_d_comm = key->commit(ac_d);
    }

    void second_phase(const Vec& r) {  // :572-622
        PO_ASSERT(c_pull.empty(), "second_phase called twice");
        PO_ASSERT(r.size() == y_logsize + d_logsize + x_logsize, "second_phase: point length");
        Vec r_d(r.begin() + y_logsize, r.begin() + y_logsize + d_logsize), r_c(r.begin() + y_logsize + d_logsize, r.end());
        Vec eq_c = eq_poly_last(r_c), eq_d = eq_poly_last(r_d);
        c_pull.resize(y_size * x_size);
        d_pull.resize(y_size * x_size);
#pragma omp parallel for schedule(static)
        for (size_t y = 0; y < y_size; y++)
            for (size_t x = 0; x < x_size; x++) {
                c_pull[y * x_size + x] = eq_c[counter[y][x]];
                d_pull[y * x_size + x] = eq_d[digits[y][x]];
            }
        // one msm_nonaff per commitment chunk (rayon par_iter over the chunks in the reference, :598-604)
        const size_t n_comms = d_outer_buckets.size();
        std::vector<G1J> dj(n_comms), cj(n_comms);
        if (n_comms >= 4) {
#pragma omp parallel for schedule(dynamic, 1)
            for (size_t k = 0; k < 2 * n_comms; k++) {
                if (k < n_comms) dj[k] = g1_msm_proj_serial(d_outer_buckets[k].data(), eq_d.data(), d_outer_buckets[k].size());
                else cj[k - n_comms] = g1_msm_proj_serial(c_outer_buckets[k - n_comms].data(), eq_c.data(), c_outer_buckets[k - n_comms].size());
            }
        } else {
            for (size_t k = 0; k < n_comms; k++) {
                dj[k] = g1_msm_proj(d_outer_buckets[k].data(), eq_d.data(), d_outer_buckets[k].size());
                cj[k] = g1_msm_proj(c_outer_buckets[k].data(), eq_c.data(), c_outer_buckets[k].size());
            }
        }
        d_pull_comm.resize(n_comms);
        c_pull_comm.resize(n_comms);
        g1_batch_to_affine(dj.data(), d_pull_comm.data(), n_comms);
        g1_batch_to_affine(cj.data(), c_pull_comm.data(), n_comms);
    }
};

/* ---- LogupMainphaseProtocol (logup_mainphase.rs:64-200) ------------------------------------------------------------------- */
typedef std::array<Vec, 2> Frac;
struct LogupMainphase {
    std::vector<size_t> logsizes;
    explicit LogupMainphase(std::vector<size_t> ls) : logsizes(std::move(ls)) {
        PO_ASSERT(logsizes.size() > 1, "logup: at least two inputs");
        for (size_t i = 0; i + 1 < logsizes.size(); i++) PO_ASSERT(logsizes[i] >= logsizes[i + 1], "logsizes must be non-increasing");
        PO_ASSERT(logsizes[0] == logsizes[1], "logup: the first two inputs must have equal size");
    }
    void make_witness(std::vector<Frac> input, std::vector<Frac>* layers_out, Fr* num, Fr* den) const {
        for (size_t i = 0; i < input.size(); i++)
            PO_ASSERT(input[i][0].size() == ((size_t)1 << logsizes[i]) && input[i][1].size() == ((size_t)1 << logsizes[i]), "logup: input size");
        std::reverse(input.begin(), input.end());
        std::vector<Frac> layers;
        layers.push_back(std::move(input.back()));
        input.pop_back();
        layers.push_back(std::move(input.back()));
        input.pop_back();
        size_t i = 0;
        LogupLayer f;
        for (;;) {
            const size_t next_size = input.empty() ? 1 : input.back()[0].size();
            const size_t curr_size = layers[i][0].size();
            std::vector<const Vec*> in{&layers[i][0], &layers[i][1], &layers[i + 1][0], &layers[i + 1][1]};
            if (curr_size == next_size) {
                std::vector<Vec> out = dense_map(in, f);
                layers.push_back(Frac{std::move(out[0]), std::move(out[1])});
                if (input.empty()) break;
                layers.push_back(std::move(input.back()));
                input.pop_back();
                i += 2;
            } else {
                PO_ASSERT(curr_size > next_size, "logup: unreachable");
                std::vector<Vec> o0, o1;
                map_split_hi(in, f, &o0, &o1);
                layers.push_back(Frac{std::move(o0[0]), std::move(o0[1])});
                layers.push_back(Frac{std::move(o1[0]), std::move(o1[1])});
                i += 2;
            }
        }
        Frac tmp = std::move(layers.back());
        layers.pop_back();
        PO_ASSERT(tmp[0].size() == 1 && tmp[1].size() == 1, "logup: root size");
        *num = tmp[0][0];
        *den = tmp[1][0];
        *layers_out = std::move(layers);
    }
    std::vector<Claims> prove(Transcript& tr, const Fr& claim, std::vector<Frac> advice) const {
        std::vector<Frac> witness;
        Fr num, denom;
        make_witness(std::move(advice), &witness, &num, &denom);
        PO_ASSERT(!denom.is_zero(), "logup: zero denominator");
        PO_ASSERT(num == denom * claim, "logup: total sum mismatch");
        tr.write_scalars(Vec{num, denom});
        std::vector<size_t> ls = logsizes;
        size_t curr = 0;
        Claims running{Vec{}, Vec{num, denom}};
        std::vector<Claims> accumulated;
        GateP f = std::make_shared<LogupLayer>();
        Claims tmp;
        for (;;) {
            const size_t incoming = ls.back();
            Frac adv_r = std::move(witness.back());
            witness.pop_back();
            Frac adv_l = std::move(witness.back());
            witness.pop_back();
            std::vector<Vec> adv;
            adv.push_back(std::move(adv_l[0]));
            adv.push_back(std::move(adv_l[1]));
            adv.push_back(std::move(adv_r[0]));
            adv.push_back(std::move(adv_r[1]));
            Claims claim_4 = dense_eq_sumcheck_prove(tr, f, curr, running, std::move(adv));
            if (incoming == curr) {
                if (ls.size() == 2) {
                    tmp = claim_4;
                    break;
                }
                running = Claims{claim_4.point, Vec{claim_4.evs[0], claim_4.evs[1]}};
                accumulated.push_back(Claims{claim_4.point, Vec{claim_4.evs[2], claim_4.evs[3]}});
                ls.pop_back();
            } else {
                running = split_at_prove(tr, claim_4, SplitIdx::HI(0), 2);
                curr += 1;
            }
        }
        accumulated.push_back(tmp);
        std::reverse(accumulated.begin(), accumulated.end());
        return accumulated;
    }
};

/* ---- PushforwardProtocol::prove (pushforward.rs:641-847) ----------------------------------------------------------------- */
struct PushforwardFinalClaims {
    Fr gamma;
    Claims matrix, ac_c, ac_d;
};
static PushforwardFinalClaims pushforward_prove(Transcript& tr, Claims claims, PushForwardState& st) {
    claims.evs[1] -= Fr::one();
    const size_t xl = st.x_logsize, yl = st.y_logsize, dl = st.d_logsize, y_size = st.y_size, x_size = st.x_size;
    PO_ASSERT(claims.point.size() == yl + dl + xl, "pushforward: point length");
    Vec r_y(claims.point.begin(), claims.point.begin() + yl), r_d(claims.point.begin() + yl, claims.point.begin() + yl + dl),
        r_c(claims.point.begin() + yl + dl, claims.point.end());
    const size_t matrix_logsize = xl + yl, matrix_size = x_size * y_size, full = (size_t)1 << matrix_logsize;
    Vec adj_p_1(x_size);
    for (size_t i = 0; i < x_size; i++) adj_p_1[i] = st.p_1[i] - Fr::one();

    std::vector<Fr> ch = tr.challenge_vec(4, 512);
    const Fr psi = ch[0], tau_c = ch[1], tau_d = ch[2], tau_s = ch[3];
    const Fr gamma = tr.challenge(128);

    Vec c_adj(matrix_size), d_adj(matrix_size);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < matrix_size; i++) {
        c_adj[i] = st.c_pull[i] + psi * st.c[i] - tau_c;
        d_adj[i] = st.d_pull[i] + psi * st.d[i] - tau_d;
    }
    pad_vector(c_adj, matrix_logsize, tau_s);
    pad_vector(d_adj, matrix_logsize, tau_s);
    Vec c_pull = st.c_pull, d_pull = st.d_pull;
    pad_vector(c_pull, matrix_logsize, Fr::zero());
    pad_vector(d_pull, matrix_logsize, Fr::zero());

    AddInverses f_addinv;
    std::vector<Vec> left, right;
    map_split_hi(std::vector<const Vec*>{&c_adj, &d_adj}, f_addinv, &left, &right);
    Vec eq_c = eq_poly_last(r_c), eq_d = eq_poly_last(r_d);
    Vec table_c(x_size), table_d((size_t)1 << dl);
    for (size_t i = 0; i < x_size; i++) table_c[i] = eq_c[i] + psi * Fr::from_u64(i) - tau_c;
    for (size_t i = 0; i < table_d.size(); i++) table_d[i] = eq_d[i] + psi * Fr::from_u64(i) - tau_d;
    PO_ASSERT(!tau_s.is_zero(), "pushforward: tau_suppression_term has no inverse");
    const Fr suppression_total = Fr::from_u64(2 * (full - matrix_size)) * tau_s.inverse();

    LogupMainphase mainphase({xl + yl - 1, xl + yl - 1, xl, dl});
    std::vector<Frac> adv;
    adv.push_back(Frac{std::move(left[0]), std::move(left[1])});
    adv.push_back(Frac{std::move(right[0]), std::move(right[1])});
    adv.push_back(Frac{st.ac_c, table_c});
    adv.push_back(Frac{st.ac_d, table_d});
    std::vector<Claims> mp = mainphase.prove(tr, suppression_total, std::move(adv));
    PO_ASSERT(mp.size() == 3, "pushforward: three mainphase claims");
    Claims cd_claims = split_at_prove(tr, mp[0], SplitIdx::HI(0), 2);
    const Claims ac_c_claims = mp[1], ac_d_claims = mp[2];

    Vec gammas = make_gamma_pows(gamma, 5);
    Vec p_folded(x_size);
    for (size_t i = 0; i < x_size; i++) p_folded[i] = st.p_0[i] + gammas[1] * adj_p_1[i] + gammas[2];
    Vec eq_sel_y = eq_trunc_evals(yl, y_size, r_y);
    Vec p_selector_prod(full);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < full; i++) p_selector_prod[i] = eq_sel_y[i >> xl] * p_folded[i & (x_size - 1)];
    PO_ASSERT(claims.evs.size() == 3, "pushforward: three input claims");
    const Fr ev_folded = claims.evs[0] + gammas[1] * claims.evs[1] + gammas[2] * claims.evs[2];

    std::vector<Vec> p3;
    p3.push_back(std::move(p_selector_prod));
    p3.push_back(std::move(c_pull));
    p3.push_back(std::move(d_pull));
    DenseSO prod3(std::move(p3), std::make_shared<Prod3>(), matrix_logsize, ev_folded);
    PO_ASSERT(cd_claims.evs.size() == 2, "pushforward: two cd claims");
    Fr claim = (cd_claims.evs[0] + gammas[1] * cd_claims.evs[1]) + gammas[2] * ev_folded;
    std::vector<Vec> fr_in;
    fr_in.push_back(std::move(c_adj));
    fr_in.push_back(std::move(d_adj));
    fr_in.push_back(eq_poly_last(cd_claims.point));
    GateSOP frac_gate = std::make_shared<EqWrapper>(std::make_shared<GammaWrapper>(std::make_shared<AddInverses>(), gamma));
    DenseSO frac(std::move(fr_in), frac_gate, cd_claims.point.size(), gamma_rlc(gamma, cd_claims.evs));

    Vec output_point;
    for (size_t i = 0; i < matrix_logsize; i++) {  // :781-801: two sumchecks driven by one combined message
        Vec pr = prod3.unipoly(), fr = frac.unipoly();
        PO_ASSERT(pr.size() == 4 && fr.size() == 4, "pushforward: degree-3 responses");
        Vec combined(4);
        for (int k = 0; k < 4; k++) combined[k] = fr[k] + gammas[2] * pr[k];
        PO_ASSERT(combined[0].dbl() + combined[1] + combined[2] + combined[3] == claim, "pushforward: combined round check");
        tr.write_scalars(compress_coefficients(combined));
        Fr t = tr.challenge(128);
        claim = evaluate_univar(combined, t);
        output_point.push_back(t);
        prod3.bind(t);
        frac.bind(t);
    }
    std::reverse(output_point.begin(), output_point.end());
    Vec pe = prod3.final_evals(), fe = frac.final_evals();
    const Fr p_selector_prod_ev = pe[0], c_pull_ev = pe[1], d_pull_ev = pe[2], c_adj_ev = fe[0], d_adj_ev = fe[1];
    Vec out_y(output_point.begin(), output_point.begin() + yl);
    const Fr trunc_ev = eq_trunc_evaluate(yl, y_size, r_y, out_y);
    PO_ASSERT(!trunc_ev.is_zero(), "pushforward: eq_sel_y evaluation has no inverse");
    const Fr adj_p_folded_ev = p_selector_prod_ev * trunc_ev.inverse();
    const Fr p_folded_ev = adj_p_folded_ev + gamma;
    const Fr sel_ev = eq_sum(out_y, y_size);
    const Fr tmp = tau_s * (Fr::one() - sel_ev);
    PO_ASSERT(!psi.is_zero(), "pushforward: psi has no inverse");
    const Fr psi_inv = psi.inverse();
    const Fr c_ev = psi_inv * (c_adj_ev - c_pull_ev + tau_c * sel_ev - tmp);
    const Fr d_ev = psi_inv * (d_adj_ev - d_pull_ev + tau_d * sel_ev - tmp);
    Vec output_evs{p_folded_ev, c_pull_ev, d_pull_ev, c_ev, d_ev};
    tr.write_scalars(output_evs);
    return PushforwardFinalClaims{gamma, Claims{output_point, output_evs}, ac_c_claims, ac_d_claims};
}

/* ---- MultiOpenReduction::prove (multiopen_reduction.rs:65-93) ---------------------------------------------------------------- */
static Claims multiopen_prove(Transcript& tr, size_t nvars, const std::vector<Vec>& points, const Vec& evs, std::vector<Vec> advice) {
    const int nargs = (int)points.size();
    Fr gamma = tr.challenge(128);
    GateSOP fun = std::make_shared<FoldedProd>(gamma, nargs);
    Fr folded = gamma_rlc(gamma, evs);
    for (auto& p : points) advice.push_back(eq_poly_last(p));
    DenseSO so(std::move(advice), fun, nvars, folded);
    Vec fe;
    Vec pt = generic_sumcheck_prove(tr, nvars, 2, so, &fe);
    fe.resize(nargs);
    tr.write_scalars(fe);
    return Claims{pt, fe};
}

/* ---- KnucklesOpeningProtocol::prove (opening.rs:39-98) --------------------------------------------------------------------- */
static void knuckles_open_prove(Transcript& tr, const KnucklesKey& pk, const G1J& commitment, const Vec& point, const Fr& ev_claim, const Vec& advice,
                                G1A* pair_a, G1A* pair_b) {
    Vec t;
    Fr opening;
    compute_t(pk, advice, point, &t, &opening);
    PO_ASSERT(opening == ev_claim, "opening: compute_t disagrees with the claimed evaluation");
    G1A t_comm = pk.commit(t);
    tr.write_points(&t_comm, 1);
    Fr x = tr.challenge(128);
    Fr kx = x * pk.k;
    Fr t_x = ev(t, x), p_x = ev(advice, x);
    tr.write_scalars(Vec{t_x, p_x});
    Fr lambda = tr.challenge(128);
    Vec p_lt(t.size());
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < t.size(); i++) p_lt[i] = lambda * t[i] + (i < advice.size() ? advice[i] : Fr::zero());
    Fr rem_unused, t_kx;
    G1A p_lt_x_proof = kzg_open(pk, p_lt, x, &rem_unused);
    tr.write_points(&p_lt_x_proof, 1);
    G1A t_kx_proof = kzg_open(pk, t, kx, &t_kx);
    tr.write_scalars(Vec{t_kx});
    tr.write_points(&t_kx_proof, 1);
    Fr fin = tr.challenge(128);
    // verify_reduce_to_pair, kzg.rs:49-60: A = opening_at * quot - opening * g0 + poly_comm, B = quot
    G1J g0 = G1J::from_affine(pk.g0);
    G1J p_lt_comm = g1_add(g1_mul(G1J::from_affine(t_comm), lambda), commitment);
    Fr p_lt_open = t_x * lambda + p_x;
    G1J a0 = g1_add(g1_add(g1_mul(G1J::from_affine(p_lt_x_proof), x), g1_neg(g1_mul(g0, p_lt_open))), p_lt_comm);
    G1J a1 = g1_add(g1_add(g1_mul(G1J::from_affine(t_kx_proof), kx), g1_neg(g1_mul(g0, t_kx))), G1J::from_affine(t_comm));
    *pair_a = g1_to_affine(g1_add(a0, g1_mul(a1, fin)));
    *pair_b = g1_to_affine(g1_add(G1J::from_affine(p_lt_x_proof), g1_mul(G1J::from_affine(t_kx_proof), fin)));
}

static G1J g1_lincomb(const Vec& coefs, const std::vector<G1A>& pts) {
    G1J acc = G1J::infinity();
    for (size_t i = 0; i < std::min(coefs.size(), pts.size()); i++) acc = g1_add(acc, g1_mul(G1J::from_affine(pts[i]), coefs[i]));
    return acc;
}

/* ---- Pippenger::prove (pippenger.rs:122-294) ---------------------------------------------------------------------------------- */
struct Timings {
    double witness = 0, ending = 0, second_phase = 0, pushforward = 0, open = 0, total = 0;
};
static void pippenger_prove(Transcript& tr, Claims claims, PushForwardState& st, PippengerEndingWG& ending, const KnucklesKey& key, Timings* tm, G1A* pair_a,
                            G1A* pair_b) {
    const size_t clm = st.clm, xl = st.x_logsize, yl = st.y_logsize, dl = st.d_logsize, y_size = st.y_size, x_size = st.x_size;
    PO_ASSERT(xl >= dl && yl >= clm, "Pippenger::new: x_logsize >= d_logsize and y_logsize >= commitment_log_multiplicity");
    const size_t n_comms = (y_size + ((size_t)1 << clm) - 1) >> clm;
    PO_ASSERT(st.c_comm.size() == n_comms && st.d_comm.size() == n_comms, "phase-1 commitments");
    double t0 = now_s();
    tr.write_points(st.c_comm);
    tr.write_points(st.d_comm);
    tr.write_points(&st.p_0_comm, 1);
    tr.write_points(&st.p_1_comm, 1);
    tr.write_points(&st.ac_c_comm, 1);
    tr.write_points(&st.ac_d_comm, 1);
    claims = pippenger_bucketed_prove(tr, claims, ending, yl, dl, xl);
    claims = glue_split_prove(tr, claims);
    tm->ending = now_s() - t0;
    t0 = now_s();
    st.second_phase(claims.point);
    tm->second_phase = now_s() - t0;
    t0 = now_s();
    tr.write_points(st.c_pull_comm);
    tr.write_points(st.d_pull_comm);
    PushforwardFinalClaims fc = pushforward_prove(tr, claims, st);
    tm->pushforward = now_s() - t0;
    t0 = now_s();
    const Fr gamma = fc.gamma;
    const Vec& matrix_pt = fc.matrix.point;
    const Fr p_folded_ev = fc.matrix.evs[0], c_pull_ev = fc.matrix.evs[1], d_pull_ev = fc.matrix.evs[2], c_ev = fc.matrix.evs[3], d_ev = fc.matrix.evs[4];
    Vec p_folded_point(clm, Fr::zero()), ac_c_point(clm, Fr::zero()), ac_d_point(xl + clm - dl, Fr::zero());
    p_folded_point.insert(p_folded_point.end(), matrix_pt.begin() + yl, matrix_pt.end());
    ac_c_point.insert(ac_c_point.end(), fc.ac_c.point.begin(), fc.ac_c.point.end());
    ac_d_point.insert(ac_d_point.end(), fc.ac_d.point.begin(), fc.ac_d.point.end());
    Vec combined_point(matrix_pt.begin() + (yl - clm), matrix_pt.end());
    Vec multirow_evs = eq_poly_last(matrix_pt.data(), yl - clm);
    G1J c_comb = g1_lincomb(multirow_evs, st.c_comm), d_comb = g1_lincomb(multirow_evs, st.d_comm);
    G1J cp_comb = g1_lincomb(multirow_evs, st.c_pull_comm), dp_comb = g1_lincomb(multirow_evs, st.d_pull_comm);
    const Fr u = tr.challenge(512);
    Vec us = make_gamma_pows(u, 4);
    G1J combined_comm = g1_add(g1_add(c_comb, g1_mul(d_comb, us[1])), g1_add(g1_mul(cp_comb, us[2]), g1_mul(dp_comb, us[3])));
    const Fr combined_ev = c_ev + d_ev * us[1] + c_pull_ev * us[2] + d_pull_ev * us[3];
    const size_t cm = (size_t)1 << clm;
    Vec combined_witness(x_size * cm);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < x_size * cm; i++) {  // :209-223
        const size_t x = i % x_size, y_rem = i >> xl;
        Fr ret = Fr::zero();
        for (size_t y = 0; y < y_size; y++)
            if (y % cm == y_rem) {
                const size_t idx = x + x_size * y;
                ret += multirow_evs[y / cm] * (st.c[idx] + st.d[idx] * us[1] + st.c_pull[idx] * us[2] + st.d_pull[idx] * us[3]);
            }
        combined_witness[i] = ret;
    }
    const size_t nv = xl + clm;
    std::vector<Vec> mw(4);
    mw[0].resize(x_size);
    for (size_t i = 0; i < x_size; i++) mw[0][i] = st.p_0[i] + gamma * st.p_1[i];
    mw[1] = st.ac_c;
    mw[2] = st.ac_d;
    mw[3] = std::move(combined_witness);
    for (auto& a : mw) pad_vector(a, nv, Fr::zero());
    std::vector<Vec> pts{p_folded_point, ac_c_point, ac_d_point, combined_point};
    Vec evs{p_folded_ev - gamma * gamma, fc.ac_c.evs[0], fc.ac_d.evs[0], combined_ev};
    Claims mo = multiopen_prove(tr, nv, pts, evs, mw);
    const Fr q = tr.challenge(128);
    Vec qs = make_gamma_pows(q, 4);
    G1J folded_comm = g1_mul(g1_add(G1J::from_affine(st.p_0_comm), g1_mul(G1J::from_affine(st.p_1_comm), gamma)), qs[0]);
    folded_comm = g1_add(folded_comm, g1_mul(G1J::from_affine(st.ac_c_comm), qs[1]));
    folded_comm = g1_add(folded_comm, g1_mul(G1J::from_affine(st.ac_d_comm), qs[2]));
    folded_comm = g1_add(folded_comm, g1_mul(combined_comm, qs[3]));
    Vec folded_witness((size_t)1 << nv);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < folded_witness.size(); i++) folded_witness[i] = mw[0][i] * qs[0] + mw[1][i] * qs[1] + mw[2][i] * qs[2] + mw[3][i] * qs[3];
    knuckles_open_prove(tr, key, folded_comm, mo.point, gamma_rlc(q, mo.evs), folded_witness, pair_a, pair_b);
    tm->open = now_s() - t0;
}
}  // namespace po

/* ---- synthetic inputs: Bandersnatch points in arithmetic progression ---------------------------------------------------------
 * build_pippenger_data draws `Affine::rand` points (pippenger.rs:472-476); the repo's synthetic workloads use the on-curve
 * progression P_i = (k0 + i * step) G instead, whose MSM is known in closed form.  Twisted Edwards
 * a x^2 + y^2 = 1 + d x^2 y^2 with a = -5 and d = COEFF_D (src/utils.rs:32-49); projective addition add-2008-bbjlp. */
namespace po {
struct TeP {
    Fr X, Y, Z;
};
static inline TeP te_add(const TeP& p, const TeP& q) {
    Fr A = p.Z * q.Z, B = A.sqr(), C = p.X * q.X, D = p.Y * q.Y, E = TE_D_M * C * D, F = B - E, G = B + E;
    TeP r;
    r.X = A * F * ((p.X + p.Y) * (q.X + q.Y) - C - D);
    r.Y = A * G * add5(D, C);  // D - a C
    r.Z = F * G;
    return r;
}
static inline TeP te_mul_u64(const TeP& g, uint64_t k) {
    TeP acc{Fr::zero(), Fr::one(), Fr::one()}, base = g;
    while (k) {
        if (k & 1) acc = te_add(acc, base);
        base = te_add(base, base);
        k >>= 1;
    }
    return acc;
}
}  // namespace po

/* ==== C ABI (ctypes) ========================================================================================================= */
using namespace po;
static thread_local std::string g_err;

static Fr fr_from_mont_limbs(const uint64_t* p) {
    Fr r;
    for (int i = 0; i < 4; i++) r.v[i] = p[i];
    return r;
}
extern "C" {
const char* po_last_error() { return g_err.c_str(); }
int po_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* KzgProvingKey::mock_setup(tau, g0, _, 2 * 2^num_vars - 1) + KnucklesProvingKey::new(.., num_vars, k); all limbs Montgomery */
void* po_key_create(const uint64_t tau[4], const uint64_t g0_xy[12], uint32_t num_vars, const uint64_t k[4]) {
    try {
        G1A g0;
        for (int i = 0; i < 6; i++) g0.x.v[i] = g0_xy[i], g0.y.v[i] = g0_xy[6 + i];
        g0.inf = false;
        return new KnucklesKey(fr_from_mont_limbs(tau), g0, num_vars, fr_from_mont_limbs(k));
    } catch (std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void po_key_destroy(void* key) { delete (KnucklesKey*)key; }
/* i-th SRS point, affine Montgomery limbs (x, y) -- lets tests compare the key with the device's gkr_srs_mock_setup */
int po_key_point(void* key, uint64_t i, uint64_t out_xy[12]) {
    KnucklesKey* k = (KnucklesKey*)key;
    if (!k || i >= k->ptau.size() || k->ptau[i].inf) return 1;
    for (int j = 0; j < 6; j++) out_xy[j] = k->ptau[i].x.v[j], out_xy[6 + j] = k->ptau[i].y.v[j];
    return 0;
}
/* commit(poly) through the key's MSM: compressed 48-byte encoding (tests: MSM == python oracle) */
int po_key_commit(void* key, const uint64_t* poly, uint64_t n, uint8_t out48[48]) {
    try {
        KnucklesKey* k = (KnucklesKey*)key;
        Vec p(n);
        for (uint64_t i = 0; i < n; i++) p[i] = fr_from_mont_limbs(poly + 4 * i);
        G1A c = k->commit(p);
        g1_serialize(c, out48);
        return 0;
    } catch (std::exception& e) {
        g_err = e.what();
        return 1;
    }
}
/* out_xy: [2][n][4] Montgomery limbs of the affine points (k0 + i * step) * G, G = the Bandersnatch subgroup generator */
int po_te_arithmetic_progression(uint64_t k0, uint64_t step, uint64_t n, uint64_t* out_xy) {
    try {
        // subgroup generator of ark-ed-on-bls12-381-bandersnatch 0.4.0 (canonical values; pinned on-curve in tests/test_oracle_pins.py)
        static const uint64_t GX[4] = {0xe1e71866a252ae18ULL, 0x2b79c022ad998465ULL, 0x743711777bbe42f3ULL, 0x29c132cc2c0b34c5ULL};
        static const uint64_t GY[4] = {0x5e3167b6cc974166ULL, 0x358cad81eee46460ULL, 0x157d8b50badcd586ULL, 0x2a6c669eda123e0fULL};
        const TeP g{Fr::from_raw(GX), Fr::from_raw(GY), Fr::one()};
        const TeP p0 = te_mul_u64(g, k0), s1 = te_mul_u64(g, step);
        int threads = po_num_threads();
        const uint64_t n_chunks = std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)threads * 4, n / 1024));
#pragma omp parallel for schedule(dynamic, 1)
        for (uint64_t ch = 0; ch < n_chunks; ch++) {
            const uint64_t lo = n * ch / n_chunks, hi = n * (ch + 1) / n_chunks;
            if (lo == hi) continue;
            std::vector<TeP> pts(hi - lo);
            // (k0 + lo * step) G = p0 + lo * (step G): lo < 2^64 and step < 2^64 multiply as two scalar multiplications
            TeP cur = te_add(p0, te_mul_u64(s1, lo));
            for (uint64_t i = lo; i < hi; i++) {
                pts[i - lo] = cur;
                cur = te_add(cur, s1);
            }
            Vec z(hi - lo);
            for (uint64_t i = 0; i < hi - lo; i++) z[i] = pts[i].Z;
            batch_inverse(z.data(), z.size());
            for (uint64_t i = lo; i < hi; i++) {
                Fr x = pts[i - lo].X * z[i - lo], y = pts[i - lo].Y * z[i - lo];
                memcpy(out_xy + 4 * i, x.v, 32);
                memcpy(out_xy + 4 * (n + i), y.v, 32);
            }
        }
        return 0;
    } catch (std::exception& e) {
        g_err = e.what();
        return 1;
    }
}
/* benchutils::run_pippenger.  points_xy: [2][n][4] Montgomery limbs (all x, then all y); coefs: [n][4] plain little-endian
 * integers already truncated to num_bits / 8 bytes (pippenger.rs:464-466); r: [y_logsize][4] Montgomery.
 * Outputs: proof bytes; dense_output tables ((d_logsize + 1) * 3 tables of 2^y_logsize Montgomery elements) and their claimed
 * evaluations; the deferred pairing pair (A, B) as 2 x 12 Montgomery limbs (affine; zeros when infinity); seconds[6] =
 * {witness + phase-1 commitments, ending GKR, second phase, pushforward, open, total}.  Returns 0, or 1 with po_last_error(). */
int po_run_pippenger(void* key_, const uint64_t* points_xy, const uint64_t* coefs, const uint64_t* r_limbs, uint32_t d_logsize, uint32_t x_logsize,
                     uint32_t num_bits, uint32_t clm, uint8_t* proof_out, uint64_t proof_cap, uint64_t* proof_len, uint64_t* dense_out, uint64_t dense_cap_elems,
                     uint64_t* n_dense_tables, uint64_t* claim_evs_out, uint64_t* pair_out, double* seconds) {
    try {
        KnucklesKey* key = (KnucklesKey*)key_;
        PO_ASSERT(key != nullptr, "null key");
        const size_t n = (size_t)1 << x_logsize;
        const size_t y_size = (num_bits + d_logsize - 1) / d_logsize;  // build_pippenger_data, pippenger.rs:468-470
        size_t y_logsize = 0;
        while (((size_t)1 << y_logsize) < y_size) y_logsize++;
        // LogupMainphaseProtocol::new([x + y - 1, x + y - 1, x, d]) panics for y_logsize = 0 (logup_mainphase.rs:75-77); the
        // reference reaches that assert only after building the witness -- this port reports it up front
        PO_ASSERT(y_logsize >= 1, "logsizes must be non-increasing (y_logsize = 0: num_bits <= d_logsize)");
        PO_ASSERT(x_logsize >= d_logsize && y_logsize >= clm, "Pippenger::new: x_logsize >= d_logsize and y_logsize >= commitment_log_multiplicity");
        Vec px(n), py(n);
        for (size_t i = 0; i < n; i++) {
            px[i] = fr_from_mont_limbs(points_xy + 4 * i);
            py[i] = fr_from_mont_limbs(points_xy + 4 * (n + i));
        }
        Vec r(y_logsize);
        for (size_t i = 0; i < y_logsize; i++) r[i] = fr_from_mont_limbs(r_limbs + 4 * i);
        Timings tm;
        const double t_start = now_s();
        Transcript tr("fgstglsp");
        PushForwardState st(px, py, coefs, y_size, y_logsize, d_logsize, x_logsize, clm, key);
        PippengerEndingWG ending(y_logsize, d_logsize, x_logsize, glue_split_witness(st.image));
        st.image.clear();
        // claim computation, pippenger.rs:531-539
        const size_t nvt = y_logsize + d_logsize - 2;
        std::vector<Vec> dense_output = triangle_last_step(ending.last(), nvt - y_logsize);
        Claims claims;
        claims.point = r;
        for (auto& o : dense_output) claims.evs.push_back(evaluate_poly(o, r));
        tm.witness = now_s() - t_start;
        G1A pa, pb;
        pippenger_prove(tr, claims, st, ending, *key, &tm, &pa, &pb);
        tm.total = now_s() - t_start;
        if (proof_len) *proof_len = tr.proof.size();
        if (proof_out) {
            PO_ASSERT(tr.proof.size() <= proof_cap, "proof buffer too small");
            memcpy(proof_out, tr.proof.data(), tr.proof.size());
        }
        if (n_dense_tables) *n_dense_tables = dense_output.size();
        if (dense_out) {
            PO_ASSERT(dense_output.size() * dense_output[0].size() <= dense_cap_elems, "dense output buffer too small");
            size_t k = 0;
            for (auto& t : dense_output)
                for (auto& v : t) {
                    memcpy(dense_out + 4 * k, v.v, 32);
                    k++;
                }
        }
        if (claim_evs_out)
            for (size_t i = 0; i < claims.evs.size(); i++) memcpy(claim_evs_out + 4 * i, claims.evs[i].v, 32);
        if (pair_out) {
            memset(pair_out, 0, 24 * 8);
            if (!pa.inf) memcpy(pair_out, pa.x.v, 48), memcpy(pair_out + 6, pa.y.v, 48);
            if (!pb.inf) memcpy(pair_out + 12, pb.x.v, 48), memcpy(pair_out + 18, pb.y.v, 48);
        }
        if (seconds) {
            seconds[0] = tm.witness, seconds[1] = tm.ending, seconds[2] = tm.second_phase, seconds[3] = tm.pushforward, seconds[4] = tm.open, seconds[5] = tm.total;
        }
        return 0;
    } catch (std::exception& e) {
        g_err = e.what();
        return 1;
    }
}
}
