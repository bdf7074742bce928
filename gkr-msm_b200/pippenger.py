"""Host-side mirror of the reference's TOP-LEVEL protocol, driving the device through the C ABI: everything
`examples/pippenger` runs between `build_pippenger_data` and `verify_pippenger`.  Orchestration, Fiat-Shamir and O(1)
claim algebra stay on the host (north_star); every table-sized step is a device call.

  KzgProvingKey / KnucklesProvingKey                 src/commitments/kzg.rs:17-133, knuckles.rs:42-154
  PushForwardState::{new, second_phase}              src/cleanup/protocols/pushforward/pushforward.rs:329-622
  PushforwardProtocol::prove                         src/cleanup/protocols/pushforward/pushforward.rs:631-847
  LogupMainphaseProtocol::prove                      src/cleanup/protocols/pushforward/logup_mainphase.rs:85-200
  DenseEqSumcheck::prove                             src/cleanup/protocols/sumcheck.rs:843-872
  MultiOpenReduction::prove                          src/cleanup/protocols/multiopen_reduction.rs:65-93
  KnucklesOpeningProtocol::prove                     src/cleanup/protocols/opening.rs:39-98
  PippengerWG::new, Pippenger::prove                 src/cleanup/protocols/pippenger.rs:30-70, 122-294
  benchutils::{build_pippenger_data, run_pippenger}  src/cleanup/protocols/pippenger.rs:462-559

Claims are (point, evs) over python ints; tables, SRS and bucket sums never leave the device.
"""
from __future__ import annotations

import numpy as np

from . import binding as g
from . import hostmath as H
from . import protocols as DP
from .fieldutil import R_MOD, from_limbs, make_gamma_pows, to_limb1, to_limbs

P = R_MOD
ONE = None  # lincomb source standing for the all-ones table

from .profiling import span  # noqa: E402


def write_points(tr, pts):  # proof_transcript.rs:64-69; pts: (12,) limb arrays
    tr.write_raw(b"".join(H.g1_serialize(H.g1_from_limbs(p)) for p in pts))


def g1_lincomb(ctx, coefs, pts) -> np.ndarray:
    """sum_i coefs[i] * pts[i] for a handful of points: a tiny MSM on the device (keeps G1 arithmetic off the host)."""
    srs = g.Srs(ctx, np.stack([np.asarray(p, dtype=np.uint64).reshape(12) for p in pts]))
    sc = ctx.upload(to_limbs(coefs))
    out = srs.msm(sc)
    srs.free()
    sc.free()
    return out


# ---------------------------------------------------------------- commitment keys ------------------------
class KzgKey:
    """KzgProvingKey (kzg.rs:17-22): the SRS resident in HBM."""

    def __init__(self, ctx, srs: g.Srs, g0_xy):
        self.ctx, self.srs, self.g0 = ctx, srs, np.asarray(g0_xy, dtype=np.uint64).reshape(12)

    @staticmethod
    def mock_setup(ctx, tau: int, g0, size: int, precompute_c: int = 0) -> "KzgKey":  # kzg.rs:84-97
        """precompute_c > 0: also build the fixed-base window table of the SRS (gkr_srs_precompute) -- proving-key
        preprocessing that pays from about 2^19 points per commitment"""
        g0_xy = H.g1_to_limbs(g0)
        srs = g.Srs.mock_setup(ctx, to_limb1(tau), g0_xy, size)
        if precompute_c:
            srs.precompute(precompute_c)
        return KzgKey(ctx, srs, g0_xy)

    @property
    def size(self):
        return self.srs.n

    def commit(self, table, n=None):  # kzg.rs:123-126
        return self.srs.msm(table, n=n)

    def open(self, table, pt_limbs):  # kzg.rs:129-132
        q, rem = self.ctx.div_by_linear(table, pt_limbs)
        comm = self.commit(q)
        q.free()
        return comm, rem

    def verify_reduce_to_pair(self, poly_comm, quot_comm, opening_at: int, opening: int):  # kzg.rs:49-60
        a = g1_lincomb(self.ctx, [opening_at, (-opening) % P, 1], [quot_comm, self.g0, poly_comm])
        return a, quot_comm


class KnucklesKey:
    """KnucklesProvingKey::new (knuckles.rs:65-81)."""

    def __init__(self, ctx, kzg: KzgKey, num_vars: int, k: int = 2):
        assert kzg.size >= 2 * (1 << num_vars) - 1, "SRS is too short."
        self.ctx, self.kzg, self.num_vars, self.k = ctx, kzg, num_vars, k % P
        self.dev = g.Knuckles(ctx, num_vars, to_limb1(self.k))

    def commit(self, table, n=None):
        return self.kzg.commit(table, n)

    def compute_t(self, table, point):
        return self.dev.compute_t(table, to_limbs(point))


# ---------------------------------------------------------------- pushforward state ----------------------
def scalar_digits(coefs_u64: np.ndarray, y_size: int, d_logsize: int) -> np.ndarray:
    """digits[y][x] = (coef_x >> (y * d)) & (2^d - 1)   (pushforward.rs:351-361); coefs_u64: (n, 4) little-endian limbs."""
    n = coefs_u64.shape[0]
    ext = np.concatenate([coefs_u64.astype(np.uint64), np.zeros((n, 1), np.uint64)], axis=1)
    out = np.empty((y_size, n), np.uint32)
    mask = np.uint64((1 << d_logsize) - 1)
    for y in range(y_size):
        bit = y * d_logsize
        limb, sh = bit >> 6, bit & 63
        v = ext[:, limb] >> np.uint64(sh)
        if sh and sh + d_logsize > 64:
            v = v | (ext[:, limb + 1] << np.uint64(64 - sh))
        out[y] = (v & mask).astype(np.uint32)
    return out


class PushForwardState:
    """PushForwardState::new (pushforward.rs:329-570).  points_xy: (2, n, 4) Montgomery limbs of the affine Bandersnatch
    coordinates; coefs_u64: (n, 4) plain little-endian scalars.  Bucketing (a counting sort of the digits) is index
    bookkeeping on the host; tables, images and commitments are built on the device."""

    def __init__(self, ctx, points_xy, coefs_u64, y_size, y_logsize, d_logsize, x_logsize, clm, key: KnucklesKey):
        assert key.num_vars == x_logsize + clm
        x_size = 1 << x_logsize
        assert points_xy.shape[1] == x_size and y_size * d_logsize <= 256
        self.ctx, self.key = ctx, key
        self.y_size, self.y_logsize, self.d_logsize, self.x_logsize, self.x_size, self.clm = y_size, y_logsize, d_logsize, x_logsize, x_size, clm
        nb = 1 << d_logsize
        sp = span(ctx, "state: host bucketing (digits, counters, order)")
        sp.__enter__()
        digits, counter, order, lens = g.pushforward_bucketize(coefs_u64, y_size, d_logsize)
        self.digits, self.counter = digits, counter
        sp.__exit__()
        sp = span(ctx, "state: upload images / tables")
        sp.__enter__()
        # image polynomials: row (y, digit) holds the coordinates of the points of that bucket in input order
        self.p_0, self.p_1 = ctx.upload(points_xy[0]), ctx.upload(points_xy[1])
        flat_order = order.reshape(-1)
        flat_lens = lens.reshape(-1)
        one = g.MONT_ONE
        zero = np.zeros(4, np.uint64)
        cl = y_logsize + d_logsize
        self.image = ctx.vecvec_gather_multi([self.p_0, self.p_1, None], flat_order, flat_lens, [zero, one, zero], [zero, one, zero],
                                             x_logsize, cl)
        self.d_idx, self.c_idx = g.U32Buf(ctx, digits.reshape(-1)), g.U32Buf(ctx, counter.reshape(-1))
        self.d, self.c = self.d_idx.to_field(), self.c_idx.to_field()
        # access counts from the bucket sizes: ac_d[v] = #incidences with digit v; ac_c[v] = #buckets longer than v
        ac_d = lens.sum(axis=0, dtype=np.uint64).astype(np.uint32)
        len_hist = np.bincount(flat_lens, minlength=x_size + 1)[:x_size + 1]
        ac_c = (flat_lens.shape[0] - np.cumsum(len_hist)[:x_size]).astype(np.uint32)
        self.ac_d, self.ac_c = g.U32Buf(ctx, ac_d).to_field(negate=True), g.U32Buf(ctx, ac_c).to_field(negate=True)
        sp.__exit__()
        sp = span(ctx, "state: c/d bucket sums + running-sum commitments")
        sp.__enter__()
        # c / d commitments: bucket sums over the SRS then running sums (pushforward.rs:398-456, 504-524)
        # All commitment chunks go through ONE bucket accumulation: bucket id = (chunk << group_log) | digit.
        comm_mul = 1 << clm
        n_comms = -(-y_size // comm_mul)
        self.n_comms = n_comms
        self.c_log = max((int(lens.max()) - 1).bit_length(), 1)  # counters run up to the longest bucket
        self.d_all = key.kzg.srs.bucket_sums_rows(self.d_idx, x_logsize, clm, d_logsize)
        self.c_all = key.kzg.srs.bucket_sums_rows(self.c_idx, x_logsize, clm, self.c_log)
        self.d_comm = list(self.d_all.weighted_sums(d_logsize, n_comms))
        self.c_comm = list(self.c_all.weighted_sums(self.c_log, n_comms))
        sp.__exit__()
        with span(ctx, "state: 4 MSM commitments (p_0, p_1, ac_c, ac_d)"):
            self.p_0_comm, self.p_1_comm = key.commit(self.p_0), key.commit(self.p_1)
            self.ac_c_comm, self.ac_d_comm = key.commit(self.ac_c), key.commit(self.ac_d)
        self.c_pull = self.d_pull = None

    def second_phase(self, r):  # pushforward.rs:572-622
        assert self.c_pull is None
        yl, dl, xl = self.y_logsize, self.d_logsize, self.x_logsize
        assert len(r) == yl + dl + xl
        ctx = self.ctx
        self.eq_d, self.eq_c = ctx.eq_table(to_limbs(r[yl:yl + dl])), ctx.eq_table(to_limbs(r[yl + dl:]))
        self.c_pull, self.d_pull = ctx.gather(self.eq_c, self.c_idx), ctx.gather(self.eq_d, self.d_idx)
        # msm_nonaff over the bucket bases with eq as scalars (pushforward.rs:598-604) == commit(c_pull chunk)
        self.c_pull_comm = list(self.c_all.msm_batch(self.eq_c, 1 << self.c_log, 0, 1 << self.c_log, self.n_comms))
        self.d_pull_comm = list(self.d_all.msm_batch(self.eq_d, 1 << dl, 0, 1 << dl, self.n_comms))


# ---------------------------------------------------------------- dense eq sumcheck ----------------------
LOGUP = dict(gid=g.GATE_LOGUP_LAYER, parts=[(g.GATE_LOGUP_LAYER, 1)], n_ins=4, n_outs=2)
ADD_INV = dict(gid=g.GATE_ADD_INVERSES, parts=[(g.GATE_ADD_INVERSES, 1)], n_ins=2, n_outs=2)


def dense_eq_so(ctx, gate, tables, point, evs, gamma):
    """DenseEqSumcheckObject::rlc (sumcheck.rs:831-841): DenseSumcheckObjectSO over EqWrapper(GammaWrapper(f, gamma))."""
    eq = ctx.eq_table(to_limbs(point)) if len(point) else ctx.upload(g.MONT_ONE.reshape(1, 4))
    gp = make_gamma_pows(gamma, gate["n_outs"])
    so = ctx.dense_so(g.SO_EQ_GAMMA, gate["gid"], list(tables) + [eq], len(point), to_limb1(H.gamma_rlc(gamma, evs)), consts=to_limbs(gp))
    return so


class DenseEqSumcheck:
    """sumcheck.rs:843-872"""

    def __init__(self, ctx, gate, num_vars):
        self.ctx, self.gate, self.num_vars = ctx, gate, num_vars

    def prove(self, tr, claims, advice):
        point, evs = claims
        gamma = from_limbs(tr.challenge(128))[0]
        if self.num_vars == 0:  # no rounds: the final evaluations are the single entries themselves
            fe = np.stack([t.download()[0] for t in advice])
            tr.write_scalars(fe)
            return ([], from_limbs(fe))
        so = dense_eq_so(self.ctx, self.gate, advice, point, evs, gamma)
        _, out_point, fe = g.sumcheck_prove(tr, so, self.num_vars)
        so.destroy()
        fe = fe[:-1]
        tr.write_scalars(fe)
        return (from_limbs(out_point), from_limbs(fe))


# ---------------------------------------------------------------- logup main phase -----------------------
class LogupMainphase:
    """logup_mainphase.rs:64-200.  Inputs: [num, denom] table pairs with non-increasing logsizes."""

    def __init__(self, ctx, logsizes):
        assert len(logsizes) > 1 and logsizes[0] == logsizes[1]
        assert all(logsizes[i] >= logsizes[i + 1] for i in range(len(logsizes) - 1)), "logsizes must be non-increasing"
        self.ctx, self.logsizes = ctx, list(logsizes)

    def make_witness(self, inp):  # logup_mainphase.rs:85-133
        ctx = self.ctx
        for arr, ls in zip(inp, self.logsizes):
            assert len(arr[0]) == 1 << ls and len(arr[1]) == 1 << ls
        inp = list(reversed(inp))
        layers = [inp.pop(), inp.pop()]
        i = 0
        while True:
            next_size = len(inp[-1][0]) if inp else 1
            curr_size = len(layers[i][0])
            a0, a1 = layers[i], layers[i + 1]
            ins = [a0[0], a0[1], a1[0], a1[1]]
            if curr_size == next_size:
                layers.append(ctx.map_dense(LOGUP["parts"], ins))
                if inp:
                    layers.append(inp.pop())
                else:
                    break
                i += 2
            else:
                assert curr_size > next_size
                o = ctx.map_dense(LOGUP["parts"], ins, split=("HI", 0), bundle_size=2)  # AlgFnUtils::map_split_hi
                layers.append(o[0:2])
                layers.append(o[2:4])
                i += 2
        tmp = layers.pop()
        assert len(tmp[0]) == 1 and len(tmp[1]) == 1
        return layers, (from_limbs(tmp[0].download())[0], from_limbs(tmp[1].download())[0])

    def prove(self, tr, claims, advice):  # logup_mainphase.rs:135-200
        with span(self.ctx, "logup: witness"):
            witness, (num, denom) = self.make_witness(advice)
        assert denom != 0 and num == denom * claims % P
        tr.write_scalars(to_limbs([num, denom]))
        running = ([], [num, denom])
        logsizes = list(self.logsizes)
        curr = 0
        accumulated = []
        while True:
            incoming = logsizes[-1]
            adv_r = witness.pop()
            adv_l = witness.pop()
            claim_4 = DenseEqSumcheck(self.ctx, LOGUP, curr).prove(tr, running, [adv_l[0], adv_l[1], adv_r[0], adv_r[1]])
            if incoming == curr:
                if len(logsizes) == 2:
                    tmp = claim_4
                    break
                running = (list(claim_4[0]), [claim_4[1][0], claim_4[1][1]])
                accumulated.append((list(claim_4[0]), [claim_4[1][2], claim_4[1][3]]))
                logsizes.pop()
            else:
                running = DP.SplitAt(("HI", 0), 2).prove(tr, claim_4)
                curr += 1
        accumulated.append(tmp)
        accumulated.reverse()
        return accumulated


# ---------------------------------------------------------------- pushforward protocol -------------------
class PushforwardProtocol:
    """pushforward.rs:300-326, 631-847"""

    def __init__(self, ctx, x_logsize, y_logsize, y_size, d_logsize):
        assert y_size <= 1 << y_logsize
        self.ctx, self.x_logsize, self.y_logsize, self.y_size, self.d_logsize = ctx, x_logsize, y_logsize, y_size, d_logsize

    def prove(self, tr, claims, st: PushForwardState):
        ctx = self.ctx
        point, evs = list(claims[0]), list(claims[1])
        evs[1] = (evs[1] - 1) % P
        xl, yl, dl, y_size = self.x_logsize, self.y_logsize, self.d_logsize, self.y_size
        r_y = point[:yl]
        assert len(point) == yl + dl + xl
        x_size = 1 << xl
        matrix_logsize, matrix_size = xl + yl, x_size * y_size
        full = 1 << matrix_logsize
        assert len(st.c) == matrix_size and len(st.c_pull) == matrix_size

        raw = tr.raw_challenge(4 * 64)  # challenge_vec(4, 512), pushforward.rs:689
        psi, tau_c, tau_d, tau_s = (H.from_le_bytes_mod_order(raw[64 * i:64 * i + 64]) for i in range(4))
        gamma = from_limbs(tr.challenge(128))[0]
        L1 = to_limb1

        def adj(pull, tab, tau):  # pull + psi * tab - tau, padded with tau_s (pushforward.rs:700-710)
            terms = [(pull, L1(1), 0, 0, matrix_size), (tab, L1(psi), 0, 0, matrix_size), (ONE, L1(-tau), 0, 0, matrix_size)]
            if full > matrix_size:
                terms.append((ONE, L1(tau_s), 0, matrix_size, full - matrix_size))
            return ctx.lincomb(terms, full)

        sp = span(ctx, "pushforward: table algebra")
        sp.__enter__()
        c_adj, d_adj = adj(st.c_pull, st.c, tau_c), adj(st.d_pull, st.d, tau_d)
        c_pull_p = ctx.lincomb([(st.c_pull, L1(1), 0, 0, matrix_size)], full)
        d_pull_p = ctx.lincomb([(st.d_pull, L1(1), 0, 0, matrix_size)], full)

        halves = ctx.map_dense(ADD_INV["parts"], [c_adj, d_adj], split=("HI", 0), bundle_size=2)  # map_split_hi, :719
        left, right = halves[0:2], halves[2:4]
        iota_c = g.U32Buf(ctx, np.arange(x_size, dtype=np.uint32)).to_field()
        table_c = ctx.lincomb([(st.eq_c, L1(1), 0, 0, x_size), (iota_c, L1(psi), 0, 0, x_size), (ONE, L1(-tau_c), 0, 0, x_size)], x_size)
        table_d = ctx.lincomb([(st.eq_d, L1(1), 0, 0, 1 << dl), (iota_c, L1(psi), 0, 0, 1 << dl), (ONE, L1(-tau_d), 0, 0, 1 << dl)], 1 << dl)
        suppression_total = 2 * (full - matrix_size) % P * H.inv(tau_s) % P if tau_s else 0

        m = xl + yl - 1
        sp.__exit__()
        with span(ctx, "pushforward: logup main phase"):
            mainphase_claims = LogupMainphase(ctx, [m, m, xl, dl]).prove(tr, suppression_total,
                                                                        [left, right, [st.ac_c, table_c], [st.ac_d, table_d]])
        assert len(mainphase_claims) == 3
        cd_claims, ac_c_claims, ac_d_claims = mainphase_claims
        cd_claims = DP.SplitAt(("HI", 0), 2).prove(tr, cd_claims)
        gammas = make_gamma_pows(gamma, 5)
        sp = span(ctx, "pushforward: combined prod3 + frac sumcheck")
        sp.__enter__()
        # p_folded = p_0 + gamma (p_1 - 1) + gamma^2 ; p_selector_prod[y, x] = eq_trunc(r_y)[y] * p_folded[x]  (:740-758)
        p_folded = ctx.lincomb([(st.p_0, L1(1), 0, 0, x_size), (st.p_1, L1(gammas[1]), 0, 0, x_size),
                                (ONE, L1(gammas[2] - gammas[1]), 0, 0, x_size)], x_size)
        eq_sel_y = H.eq_trunc_evals(yl, y_size, r_y)
        p_selector_prod = ctx.lincomb([(p_folded, L1(eq_sel_y[y]), 0, y << xl, x_size) for y in range(y_size)], full)
        assert len(evs) == 3
        ev_folded = (evs[0] + gammas[1] * evs[1] + gammas[2] * evs[2]) % P
        prod3 = ctx.dense_so(g.SO_PLAIN, g.GATE_PROD3, [p_selector_prod, c_pull_p, d_pull_p], matrix_logsize, L1(ev_folded))
        cd_point, cd_evs = cd_claims
        assert len(cd_evs) == 2
        claim = (cd_evs[0] + gammas[1] * cd_evs[1] + gammas[2] * ev_folded) % P
        frac = dense_eq_so(ctx, ADD_INV, [c_adj, d_adj], cd_point, cd_evs, gamma)
        output_point = []
        for _ in range(matrix_logsize):  # the combined loop, pushforward.rs:781-806
            pr = H.from_evals(from_limbs(prod3.unipoly()))
            fr = H.from_evals(from_limbs(frac.unipoly()))
            assert len(pr) == 4 and len(fr) == 4
            combined = [(fr[i] + gammas[2] * pr[i]) % P for i in range(4)]
            assert (2 * combined[0] + combined[1] + combined[2] + combined[3]) % P == claim
            tr.write_scalars(to_limbs(H.compress_coefficients(combined)))
            t_l = tr.challenge(128)
            t = from_limbs(t_l)[0]
            claim = H.evaluate_univar(combined, t)
            output_point.append(t)
            prod3.bind(t_l)
            frac.bind(t_l)
        output_point.reverse()
        p_selector_prod_ev, c_pull_ev, d_pull_ev = from_limbs(prod3.final_evals())
        c_adj_ev, d_adj_ev, _ = from_limbs(frac.final_evals())
        prod3.destroy()
        frac.destroy()
        sp.__exit__()
        adj_p_folded_ev = p_selector_prod_ev * H.inv(H.eq_trunc_evaluate(yl, y_size, r_y, output_point[:yl])) % P
        p_folded_ev = (adj_p_folded_ev + gamma) % P
        sel_ev = H.eq_sum(output_point[:yl], y_size)  # SelectorPoly::evaluate, verifier_polys.rs:68-71
        tmp = tau_s * (1 - sel_ev) % P
        psi_inv = H.inv(psi)
        c_ev = psi_inv * (c_adj_ev - c_pull_ev + tau_c * sel_ev - tmp) % P
        d_ev = psi_inv * (d_adj_ev - d_pull_ev + tau_d * sel_ev - tmp) % P
        output_evs = [p_folded_ev, c_pull_ev, d_pull_ev, c_ev, d_ev]
        tr.write_scalars(to_limbs(output_evs))
        return dict(gamma=gamma, matrix=(output_point, output_evs), ac_c=ac_c_claims, ac_d=ac_d_claims)


# ---------------------------------------------------------------- multiopen reduction --------------------
class MultiOpenReduction:
    """multiopen_reduction.rs:43-93.  claims = [(point, ev)] * nargs; advice = nargs device tables of 2^nvars."""

    def __init__(self, ctx, nvars, nargs):
        self.ctx, self.nvars, self.nargs = ctx, nvars, nargs

    def prove(self, tr, claims, advice):
        ctx = self.ctx
        gamma = from_limbs(tr.challenge(128))[0]
        folded = H.gamma_rlc(gamma, [c[1] for c in claims])
        tables = list(advice) + [ctx.eq_table(to_limbs(c[0])) for c in claims]
        gp = make_gamma_pows(gamma, self.nargs)
        so = ctx.dense_so(g.SO_PLAIN, g.GATE_FOLDED_PROD, tables, self.nvars, to_limb1(folded), gate_param=self.nargs, consts=to_limbs(gp))
        _, out_point, fe = g.sumcheck_prove(tr, so, self.nvars)
        so.destroy()
        evs = fe[:self.nargs]
        tr.write_scalars(evs)
        return (from_limbs(out_point), from_limbs(evs))


# ---------------------------------------------------------------- Knuckles opening -----------------------
class KnucklesOpening:
    """opening.rs:13-98.  claim = (commitment limbs, point, ev); advice = the committed device table."""

    def __init__(self, ctx, key: KnucklesKey):
        self.ctx, self.key = ctx, key

    def prove(self, tr, claim, advice):
        ctx, pk = self.ctx, self.key
        comm, point, ev_claim = claim
        with span(ctx, "knuckles: compute_t"):
            t, opening = pk.compute_t(advice, point)
        assert from_limbs(opening)[0] == ev_claim
        with span(ctx, "knuckles: commit t (MSM 2N-1)"):
            t_comm = pk.kzg.commit(t)
        write_points(tr, [t_comm])
        x_l = tr.challenge(128)
        x = from_limbs(x_l)[0]
        kx = x * pk.k % P
        t_x, p_x = ctx.poly_eval(t, x_l), ctx.poly_eval(advice, x_l)
        tr.write_scalars(np.stack([t_x, p_x]))
        lam_l = tr.challenge(128)
        lam = from_limbs(lam_l)[0]
        p_lt = ctx.lincomb([(t, lam_l, 0, 0, len(t)), (advice, to_limb1(1), 0, 0, len(advice))], len(t))  # opening.rs:65-75
        with span(ctx, "knuckles: open p_lt (div + MSM)"):
            p_lt_x_proof, _ = pk.kzg.open(p_lt, x_l)
        write_points(tr, [p_lt_x_proof])
        with span(ctx, "knuckles: open t (div + MSM)"):
            t_kx_proof, t_kx = pk.kzg.open(t, to_limb1(kx))
        tr.write_scalars(t_kx.reshape(1, 4))
        write_points(tr, [t_kx_proof])
        fin = from_limbs(tr.challenge(128))[0]
        t_x_i, p_x_i, t_kx_i = from_limbs(t_x)[0], from_limbs(p_x)[0], from_limbs(t_kx)[0]
        p_lt_comm = g1_lincomb(ctx, [lam, 1], [t_comm, comm])
        p_lt_open = (t_x_i * lam + p_x_i) % P
        a0, b0 = pk.kzg.verify_reduce_to_pair(p_lt_comm, p_lt_x_proof, x, p_lt_open)
        a1, b1 = pk.kzg.verify_reduce_to_pair(t_comm, t_kx_proof, kx, t_kx_i)
        return (g1_lincomb(ctx, [1, fin], [a0, a1]), g1_lincomb(ctx, [1, fin], [b0, b1]))


# ---------------------------------------------------------------- top level ------------------------------
class PippengerWG:
    """pippenger.rs:30-70"""

    def __init__(self, ctx, points_xy, coefs_u64, y_size, y_logsize, d_logsize, x_logsize, clm, key):
        self.beginning = PushForwardState(ctx, points_xy, coefs_u64, y_size, y_logsize, d_logsize, x_logsize, clm, key)
        with span(ctx, "witness: bintree + triangle (EC-add maps)"):
            self.ending = DP.PippengerEndingWG(ctx, y_logsize, d_logsize, x_logsize, DP.GlueSplit.witness(ctx, self.beginning.image))


class Pippenger:
    """pippenger.rs:72-294"""

    def __init__(self, ctx, y_size, y_logsize, d_logsize, x_logsize, key: KnucklesKey, clm):
        assert x_logsize >= d_logsize and y_logsize >= clm
        self.ctx, self.key, self.clm = ctx, key, clm
        self.beginning = PushforwardProtocol(ctx, x_logsize, y_logsize, y_size, d_logsize)
        self.ending = DP.PippengerBucketed(ctx, y_logsize, d_logsize, x_logsize)

    def prove(self, tr, claims, state: PippengerWG):
        ctx, b, clm = self.ctx, self.beginning, self.clm
        st = state.beginning
        n_comms = -(-b.y_size // (1 << clm))
        assert len(st.c_comm) == n_comms and len(st.d_comm) == n_comms
        write_points(tr, st.c_comm)
        write_points(tr, st.d_comm)
        for pt in (st.p_0_comm, st.p_1_comm, st.ac_c_comm, st.ac_d_comm):
            write_points(tr, [pt])
        with span(ctx, "prove: ending GKR (triangle + bintree sumchecks)"):
            claims = self.ending.prove(tr, claims, state.ending)
        claims = DP.GlueSplit().prove(tr, claims)
        with span(ctx, "prove: second phase (pulls + msm_nonaff)"):
            st.second_phase(claims[0])
        write_points(tr, st.c_pull_comm)
        write_points(tr, st.d_pull_comm)
        with span(ctx, "prove: pushforward"):
            fc = b.prove(tr, claims, st)
        gamma = fc["gamma"]
        sp = span(ctx, "prove: opening inputs (lincombs)")
        sp.__enter__()
        # opening claims (pippenger.rs:166-205)
        matrix_pt, matrix_evs = fc["matrix"]
        p_folded_ev, c_pull_ev, d_pull_ev, c_ev, d_ev = matrix_evs
        p_folded_point = [0] * clm + list(matrix_pt[b.y_logsize:])
        ac_c_point = [0] * clm + list(fc["ac_c"][0])
        ac_d_point = [0] * (b.x_logsize + clm - b.d_logsize) + list(fc["ac_d"][0])
        combined_point = list(matrix_pt[b.y_logsize - clm:])
        multirow_evs = H.eq_poly_sequence_last(matrix_pt[:b.y_logsize - clm])
        u = H.from_le_bytes_mod_order(tr.raw_challenge(64))  # challenge(512)
        us = make_gamma_pows(u, 4)
        combined_ev = (c_ev + d_ev * us[1] + c_pull_ev * us[2] + d_pull_ev * us[3]) % P
        comm_coefs, comm_pts = [], []
        for j, comms in enumerate((st.c_comm, st.d_comm, st.c_pull_comm, st.d_pull_comm)):
            for k, pt in enumerate(comms):
                comm_coefs.append(multirow_evs[k] * us[j] % P)
                comm_pts.append(pt)
        combined_comm = g1_lincomb(ctx, comm_coefs, comm_pts)
        oclaims = [(p_folded_point, (p_folded_ev - gamma * gamma) % P), (ac_c_point, fc["ac_c"][1][0]),
                   (ac_d_point, fc["ac_d"][1][0]), (combined_point, combined_ev)]
        # combined witness (pippenger.rs:209-223): row y of c, d, c_pull, d_pull lands in slot (y mod 2^clm)
        x_size, y_size, xl = 1 << b.x_logsize, b.y_size, b.x_logsize
        cm = 1 << clm
        nv = xl + clm
        terms = []
        for y in range(y_size):
            for j, tab in enumerate((st.c, st.d, st.c_pull, st.d_pull)):
                terms.append((tab, to_limb1(multirow_evs[y // cm] * us[j] % P), x_size * y, x_size * (y % cm), x_size))
        combined_witness = ctx.lincomb(terms, 1 << nv)
        one = to_limb1(1)
        mw = [ctx.lincomb([(st.p_0, one, 0, 0, x_size), (st.p_1, to_limb1(gamma), 0, 0, x_size)], 1 << nv),
              ctx.lincomb([(st.ac_c, one, 0, 0, len(st.ac_c))], 1 << nv),
              ctx.lincomb([(st.ac_d, one, 0, 0, len(st.ac_d))], 1 << nv),
              combined_witness]
        sp.__exit__()
        with span(ctx, "prove: multiopen reduction"):
            mo_point, mo_evs = MultiOpenReduction(ctx, nv, 4).prove(tr, oclaims, mw)
        q = from_limbs(tr.challenge(128))[0]
        qs = make_gamma_pows(q, 4)
        folded_comm = g1_lincomb(ctx, [qs[0], qs[0] * gamma % P, qs[1], qs[2], qs[3]],
                                 [st.p_0_comm, st.p_1_comm, st.ac_c_comm, st.ac_d_comm, combined_comm])
        folded_witness = ctx.lincomb([(mw[i], to_limb1(qs[i]), 0, 0, 1 << nv) for i in range(4)], 1 << nv)
        with span(ctx, "prove: knuckles opening"):
            return KnucklesOpening(ctx, self.key).prove(tr, (folded_comm, mo_point, H.gamma_rlc(q, mo_evs)), folded_witness)


def pippenger_config(d_logsize, x_logsize, num_bits, clm):  # build_pippenger_data, pippenger.rs:462-497
    y_size = (num_bits + d_logsize - 1) // d_logsize
    y_logsize = (y_size - 1).bit_length()  # ark_std::log2 = ceil(log2)
    return dict(y_size=y_size, y_logsize=y_logsize, d_logsize=d_logsize, x_logsize=x_logsize, clm=clm)


def run_pippenger(ctx, tr, points_xy, coefs_u64, cfg, r, key):
    """benchutils::run_pippenger (pippenger.rs:499-559): witness + phase-1 commitments + proof.
    Returns (dense_output device tables, claims, the opening pair)."""
    y_size, yl, dl, xl, clm = cfg["y_size"], cfg["y_logsize"], cfg["d_logsize"], cfg["x_logsize"], cfg["clm"]
    wg = PippengerWG(ctx, points_xy, coefs_u64, y_size, yl, dl, xl, clm, key)
    dense_output = DP.triangle_last_step(ctx, wg.ending.last(), yl + dl - 2 - yl)
    claims = (list(r), [H.evaluate_poly(from_limbs(o.download()), r) for o in dense_output])
    pair = Pippenger(ctx, y_size, yl, dl, xl, key, clm).prove(tr, claims, wg)
    return dense_output, claims, pair
