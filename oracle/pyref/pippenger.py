"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

The top of the reference's proof system restated over python ints: everything `examples/pippenger` runs between
`build_pippenger_data` and `verify_pippenger`.

  KzgProvingKey / KzgVerifyingKey                    src/commitments/kzg.rs:17-150
  KnucklesProvingKey                                 src/commitments/knuckles.rs:42-154
  EqPoly, SelectorPoly, EqTruncPoly                  src/cleanup/protocols/verifier_polys.rs:14-137
  PushForwardState::{new, second_phase}              src/cleanup/protocols/pushforward/pushforward.rs:329-622
  PushforwardProtocol::{prove, verify}               src/cleanup/protocols/pushforward/pushforward.rs:631-966
  LogupMainphaseProtocol                             src/cleanup/protocols/pushforward/logup_mainphase.rs:64-250
  MultiOpenReduction                                 src/cleanup/protocols/multiopen_reduction.rs:43-117
  KnucklesOpeningProtocol                            src/cleanup/protocols/opening.rs:13-141
  PippengerWG, Pippenger::{prove, verify}            src/cleanup/protocols/pippenger.rs:30-406
  benchutils::{run_pippenger, verify_pippenger}      src/cleanup/protocols/pippenger.rs:499-606

The SRS is the reference's own mock setup (kzg.rs:84-97): ptau_1[i] = tau^i * g0 with a KNOWN tau, h1 = tau * h0.
Every commitment in the protocol is a linear combination of SRS points, so the oracle computes it as
(sum_i coeff_i tau^i) * g0 -- the unique group element any MSM algorithm must return -- and the final pairing check
<A, h0> == <B, h1> (kzg.rs:63-67) reduces to A == tau * B in G1.
"""
from __future__ import annotations

from . import curves as CV
from . import gates as G
from . import gkr as K
from . import polys as OP
from .commitments import compute_t, div_by_linear, ev, knuckles_inverses
from .field import FQ_MODULUS as Q
from .field import P
from .sumcheck import (DenseEqSumcheck, DenseSumcheckObjectSO, compress_coefficients, decompress_coefficients, eq_eval,
                       eq_poly_sequence_last, eq_sum, evaluate_poly, evaluate_univar, gamma_rlc, generic_sumcheck_prove,
                       generic_sumcheck_verify, make_gamma_pows, VecVecPolynomial)


def inv(x):
    return pow(x % P, -1, P)


# ---------------------------------------------------------------- G1 wire format -------------------------
def g1_serialize(pt) -> bytes:
    """ark-bls12-381 0.4.0 compressed G1 (third-party, not in the reference tree): the zcash / IETF encoding --
    48-byte big-endian x with the three top bits of byte 0 = (compressed, infinity, y is the larger root)."""
    if pt is None:
        return bytes([0xC0]) + bytes(47)
    x, y = pt
    b = bytearray(x.to_bytes(48, "big"))
    b[0] |= 0x80
    if y > (Q - 1) // 2:
        b[0] |= 0x20
    return bytes(b)


def g1_deserialize(b: bytes):
    assert len(b) == 48 and b[0] & 0x80
    if b[0] & 0x40:
        return None
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    y = pow((x * x * x + CV.G1_B) % Q, (Q + 1) // 4, Q)
    assert (y * y - x * x * x - CV.G1_B) % Q == 0, "not on the curve"
    if (y > (Q - 1) // 2) != bool(b[0] & 0x20):
        y = Q - y
    return (x, y)


def write_points(transcript, pts):  # proof_transcript.rs:64-69
    transcript.write_raw_msg(b"".join(g1_serialize(p) for p in pts))


def read_points(transcript, n):  # proof_transcript.rs:58-61
    raw = transcript.read_raw_msg(48 * n)
    return [g1_deserialize(raw[48 * i:48 * i + 48]) for i in range(n)]


def g1_sub(a, b):
    return CV.g1_add(a, CV.g1_neg(b))


# ---------------------------------------------------------------- commitment keys ------------------------
class KzgKey:
    """KzgProvingKey::mock_setup(tau, g0, h0, size)  kzg.rs:84-97 (G2 side represented by tau itself)."""

    def __init__(self, tau, g0, size):
        self.tau, self.g0, self.size = tau % P, g0, size
        self._basis = {}

    def basis(self, i):
        assert i < self.size
        if i not in self._basis:
            self._basis[i] = CV.g1_mul(pow(self.tau, i, P), self.g0)
        return self._basis[i]

    def commit(self, poly):  # kzg.rs:123-126
        assert len(poly) <= self.size, "Vector is too large."
        return CV.g1_mul(ev(poly, self.tau), self.g0)

    def commit_literal(self, poly):
        """the same through the SRS points (slow; used to pin the shortcut in tests)"""
        return CV.g1_msm([self.basis(i) for i in range(len(poly))], poly)

    def open(self, poly, pt):  # kzg.rs:129-132
        q, rem = div_by_linear(poly, pt)
        return self.commit(q), rem

    def verify_reduce_to_pair(self, poly_comm, quot_comm, opening_at, opening):  # kzg.rs:49-60
        a = CV.g1_add(g1_sub(CV.g1_mul(opening_at, quot_comm), CV.g1_mul(opening, self.g0)), poly_comm)
        return a, quot_comm

    def verify_pair(self, pair):  # kzg.rs:63-67 with h1 = tau h0
        assert pair[0] == CV.g1_mul(self.tau, pair[1]), "pairing check failed"


class KnucklesKey:
    """KnucklesProvingKey::new  knuckles.rs:65-81."""

    def __init__(self, kzg: KzgKey, num_vars, k=2):
        n = 1 << num_vars
        assert kzg.size >= 2 * n - 1, "SRS is too short."
        self.kzg, self.num_vars, self.k = kzg, num_vars, k % P
        self.inverses = knuckles_inverses(num_vars, self.k)

    def commit(self, poly):
        assert len(poly) <= 1 << self.num_vars
        return self.kzg.commit(poly)

    def compute_t(self, poly, point):
        return compute_t(self.num_vars, self.inverses, poly, point)


# ---------------------------------------------------------------- verifier polys -------------------------
def eq_trunc_evals(num_vars, k, r):  # verifier_polys.rs:90-96
    ret = eq_poly_sequence_last(r)
    for i in range(k, 1 << num_vars):
        ret[i] = 0
    return ret


def eq_trunc_evaluate(num_vars, k, r, pt):  # verifier_polys.rs:98-136
    assert len(pt) == num_vars
    partial = [1]
    for i in range(num_vars):
        j = num_vars - i - 1
        partial.append(partial[-1] * ((1 - pt[j] - r[j] + 2 * r[j] * pt[j]) % P) % P)
    if k >= (1 << num_vars):
        assert k == 1 << num_vars
        return partial[num_vars]
    multiplier, acc = 1, 0
    for i in range(num_vars):
        left_bit = k >> (num_vars - i - 1)
        m_ = multiplier
        if left_bit == 1:
            multiplier = multiplier * pt[i] % P * r[i] % P
            acc = (acc + m_ * (1 - pt[i]) % P * (1 - r[i]) % P * partial[num_vars - i - 1]) % P
        else:
            multiplier = multiplier * (1 - pt[i]) % P * (1 - r[i]) % P
        k -= left_bit << (num_vars - i - 1)
    return acc


def selector_evaluate(num_vars, k, pt):  # verifier_polys.rs:68-71
    assert len(pt) == num_vars
    return eq_sum(pt, k)


def pad_vector(v, logsize, with_):  # utils.rs:324-329
    assert len(v) <= 1 << logsize
    return list(v) + [with_] * ((1 << logsize) - len(v))


# ---------------------------------------------------------------- pushforward state ----------------------
class PushForwardState:
    """PushForwardState::new  pushforward.rs:329-570."""

    def __init__(self, points, coefs, y_size, y_logsize, d_logsize, x_logsize, clm, key: KnucklesKey, literal_commits=False):
        assert key.num_vars == x_logsize + clm
        x_size = 1 << x_logsize
        assert len(points) == x_size and y_size * d_logsize <= 256
        self.y_size, self.y_logsize, self.d_logsize, self.x_logsize, self.x_size = y_size, y_logsize, d_logsize, x_logsize, x_size
        self.clm, self.key = clm, key
        polys = [[p[0] for p in points], [p[1] for p in points], [1] * x_size]
        mask = (1 << d_logsize) - 1
        self.digits = [[(coefs[x] >> (y * d_logsize)) & mask for x in range(x_size)] for y in range(y_size)]
        row_pad, col_pad = [0, 1, 0], [0, 1, 0]
        self.counter = [[0] * x_size for _ in range(y_size)]
        buckets = [[[] for _ in polys] for _ in range(y_size << d_logsize)]
        comm_mul = 1 << clm
        for y in range(y_size):
            chunk = buckets[y << d_logsize:(y + 1) << d_logsize]
            for x in range(x_size):
                d = self.digits[y][x]
                self.counter[y][x] = len(chunk[d][0])
                for pid in range(3):
                    chunk[d][pid].append(polys[pid][x])
        self.image = [VecVecPolynomial([buckets[r][pid] for r in range(y_size << d_logsize)], row_pad[pid], col_pad[pid],
                                       x_logsize, y_logsize + d_logsize) for pid in range(3)]
        self.d = [v for row in self.digits for v in row]
        self.c = [v for row in self.counter for v in row]
        ac_d, ac_c = [0] * (1 << d_logsize), [0] * x_size
        for v in self.d:
            ac_d[v] += 1
        for v in self.c:
            ac_c[v] += 1
        self.ac_c = [(-v) % P for v in ac_c]
        self.ac_d = [(-v) % P for v in ac_d]
        self.p_0, self.p_1 = polys[0], polys[1]
        chunk_len = x_size * comm_mul
        n_comms = -(-y_size // comm_mul)
        if literal_commits:
            # the reference's route: bucket sums over SRS points, then running sums (pushforward.rs:398-456, 504-524)
            from .commitments import bucket_sums, running_sum_commit
            self.c_comm, self.d_comm = [], []
            for k in range(n_comms):
                ys = range(k * comm_mul, min((k + 1) * comm_mul, y_size))
                pidx = [x + x_size * (y % comm_mul) for y in ys for x in range(x_size)]
                bases = {i: key.kzg.basis(i) for i in set(pidx)}
                dd = [self.digits[y][x] for y in ys for x in range(x_size)]
                cc = [self.counter[y][x] for y in ys for x in range(x_size)]
                self.d_comm.append(running_sum_commit(bucket_sums(bases, pidx, dd, 1 << d_logsize)))
                self.c_comm.append(running_sum_commit(bucket_sums(bases, pidx, cc, max(cc) + 1)))
        else:
            self.c_comm = [key.commit(self.c[k * chunk_len:(k + 1) * chunk_len]) for k in range(n_comms)]
            self.d_comm = [key.commit(self.d[k * chunk_len:(k + 1) * chunk_len]) for k in range(n_comms)]
        self.p_0_comm, self.p_1_comm = key.commit(self.p_0), key.commit(self.p_1)
        self.ac_c_comm, self.ac_d_comm = key.commit(self.ac_c), key.commit(self.ac_d)
        self.c_pull = self.d_pull = None

    def second_phase(self, r):  # pushforward.rs:572-622
        assert self.c_pull is None
        yl, dl, xl = self.y_logsize, self.d_logsize, self.x_logsize
        assert len(r) == yl + dl + xl
        r_d, r_c = r[yl:yl + dl], r[yl + dl:]
        eq_c, eq_d = eq_poly_sequence_last(r_c), eq_poly_sequence_last(r_d)
        self.c_pull = [eq_c[v] for v in self.c]
        self.d_pull = [eq_d[v] for v in self.d]
        chunk_len = self.x_size << self.clm
        n_comms = -(-self.y_size // (1 << self.clm))
        self.c_pull_comm = [self.key.commit(self.c_pull[k * chunk_len:(k + 1) * chunk_len]) for k in range(n_comms)]
        self.d_pull_comm = [self.key.commit(self.d_pull[k * chunk_len:(k + 1) * chunk_len]) for k in range(n_comms)]


# ---------------------------------------------------------------- logup main phase -----------------------
class LogupMainphase:
    """logup_mainphase.rs:64-250."""

    def __init__(self, logsizes):
        assert len(logsizes) > 1 and logsizes[0] == logsizes[1]
        assert all(logsizes[i] >= logsizes[i + 1] for i in range(len(logsizes) - 1)), "logsizes must be non-increasing"
        self.logsizes = list(logsizes)

    def make_witness(self, inp):
        for arr, ls in zip(inp, self.logsizes):
            assert len(arr[0]) == 1 << ls and len(arr[1]) == 1 << ls
        inp = list(reversed(inp))
        layers = [inp.pop(), inp.pop()]
        i = 0
        f = G.LogupLayer()
        while True:
            next_size = len(inp[-1][0]) if inp else 1
            curr_size = len(layers[i][0])
            a0, a1 = layers[i], layers[i + 1]
            if curr_size == next_size:
                layers.append(OP.dense_map([a0[0], a0[1], a1[0], a1[1]], f))
                if inp:
                    layers.append(inp.pop())
                else:
                    break
                i += 2
            else:
                assert curr_size > next_size
                out0, out1 = OP.map_split_hi([a0[0], a0[1], a1[0], a1[1]], f)
                layers.append(out0)
                layers.append(out1)
                i += 2
        tmp = layers.pop()
        assert len(tmp[0]) == 1 and len(tmp[1]) == 1
        return layers, (tmp[0][0], tmp[1][0])

    def _run(self, transcript, running_claim, step):
        f = G.LogupLayer()
        logsizes = list(self.logsizes)
        curr = 0
        accumulated = []
        while True:
            incoming = logsizes[-1]
            claim_4 = step(DenseEqSumcheck(f, curr), running_claim)
            if incoming == curr:
                if len(logsizes) == 2:
                    tmp = claim_4
                    break
                running_claim = (list(claim_4[0]), [claim_4[1][0], claim_4[1][1]])
                accumulated.append((list(claim_4[0]), [claim_4[1][2], claim_4[1][3]]))
                logsizes.pop()
            else:
                running_claim = K.SplitAt(("HI", 0), 2).prove(transcript, claim_4)
                curr += 1
        accumulated.append(tmp)
        accumulated.reverse()
        return accumulated

    def prove(self, transcript, claims, advice):
        witness, (num, denom) = self.make_witness(advice)
        assert denom != 0 and num == denom * claims % P
        transcript.write_scalars([num, denom])

        def step(proto, running):
            adv_r = witness.pop()
            adv_l = witness.pop()
            return proto.prove(transcript, running, [adv_l[0], adv_l[1], adv_r[0], adv_r[1]])

        return self._run(transcript, ([], [num, denom]), step)

    def verify(self, transcript, claims):
        num, denom = transcript.read_scalars(2)
        assert denom != 0 and num == denom * claims % P
        return self._run(transcript, ([], [num, denom]), lambda proto, running: proto.verify(transcript, running))


# ---------------------------------------------------------------- pushforward protocol -------------------
class PushforwardProtocol:
    """pushforward.rs:300-326, 631-966."""

    def __init__(self, x_logsize, y_logsize, y_size, d_logsize):
        assert y_size <= 1 << y_logsize
        self.x_logsize, self.y_logsize, self.y_size, self.d_logsize = x_logsize, y_logsize, y_size, d_logsize

    def _mainphase(self):
        m = self.x_logsize + self.y_logsize - 1
        return LogupMainphase([m, m, self.x_logsize, self.d_logsize])

    def prove(self, transcript, claims, st: PushForwardState):
        point, evs = list(claims[0]), list(claims[1])
        evs[1] = (evs[1] - 1) % P
        xl, yl, dl, y_size = self.x_logsize, self.y_logsize, self.d_logsize, self.y_size
        r_y, r_d, r_c = point[:yl], point[yl:yl + dl], point[yl + dl:]
        assert len(r_c) == xl
        x_size = 1 << xl
        matrix_logsize, matrix_size = xl + yl, x_size * y_size
        c, d, p_0, p_1, ac_c, ac_d = st.c, st.d, st.p_0, st.p_1, st.ac_c, st.ac_d
        c_pull, d_pull = st.c_pull, st.d_pull
        adj_p_1 = [(v - 1) % P for v in p_1]
        assert len(c) == matrix_size and len(c_pull) == matrix_size

        psi, tau_c, tau_d, tau_s = transcript.challenge_vec(4, 512)
        gamma = transcript.challenge(128)
        c_adj = pad_vector([(cp + psi * cv - tau_c) % P for cp, cv in zip(c_pull, c)], matrix_logsize, tau_s)
        d_adj = pad_vector([(dp + psi * dv - tau_d) % P for dp, dv in zip(d_pull, d)], matrix_logsize, tau_s)
        c_pull_p = pad_vector(c_pull, matrix_logsize, 0)
        d_pull_p = pad_vector(d_pull, matrix_logsize, 0)

        f_addinv = G.AddInverses()
        left, right = OP.map_split_hi([c_adj, d_adj], f_addinv)
        eq_c, eq_d = eq_poly_sequence_last(r_c), eq_poly_sequence_last(r_d)
        table_c = [(eq_c[i] + psi * i - tau_c) % P for i in range(x_size)]
        table_d = [(eq_d[i] + psi * i - tau_d) % P for i in range(1 << dl)]
        suppression_total = 2 * ((1 << matrix_logsize) - matrix_size) % P * inv(tau_s) % P if tau_s else 0

        mainphase_claims = self._mainphase().prove(transcript, suppression_total,
                                                   [left, right, [list(ac_c), table_c], [list(ac_d), table_d]])
        assert len(mainphase_claims) == 3
        cd_claims, ac_c_claims, ac_d_claims = mainphase_claims
        cd_claims = K.SplitAt(("HI", 0), 2).prove(transcript, cd_claims)
        gammas = make_gamma_pows(gamma, 5)
        p_folded = [(a + gammas[1] * b + gammas[2]) % P for a, b in zip(p_0, adj_p_1)]
        eq_sel_y = eq_trunc_evals(yl, y_size, r_y)
        p_selector_prod = [eq_sel_y[i >> xl] * p_folded[i & (x_size - 1)] % P for i in range(1 << matrix_logsize)]
        assert len(evs) == 3
        ev_folded = (evs[0] + gammas[1] * evs[1] + gammas[2] * evs[2]) % P
        prod3 = DenseSumcheckObjectSO([p_selector_prod, c_pull_p, d_pull_p], G.Prod3(), matrix_logsize, ev_folded)
        cd_point, cd_evs = cd_claims
        assert len(cd_evs) == 2
        claim = (cd_evs[0] + gammas[1] * cd_evs[1] + gammas[2] * ev_folded) % P
        frac = DenseEqSumcheck(f_addinv, matrix_logsize).make_so([c_adj, d_adj], cd_point, cd_evs, gamma)
        output_point = []
        for _ in range(matrix_logsize):
            pr = prod3.unipoly()
            fr = frac.unipoly()
            assert len(pr) == 4 and len(fr) == 4
            combined = [(fr[i] + gammas[2] * pr[i]) % P for i in range(4)]
            assert (2 * combined[0] + combined[1] + combined[2] + combined[3]) % P == claim
            transcript.write_scalars(compress_coefficients(combined))
            t = transcript.challenge(128)
            claim = evaluate_univar(combined, t)
            output_point.append(t)
            prod3.bind(t)
            frac.bind(t)
        output_point.reverse()
        p_selector_prod_ev, c_pull_ev, d_pull_ev = prod3.final_evals()
        c_adj_ev, d_adj_ev, _ = frac.final_evals()
        adj_p_folded_ev = p_selector_prod_ev * inv(eq_trunc_evaluate(yl, y_size, r_y, output_point[:yl])) % P
        p_folded_ev = (adj_p_folded_ev + gamma) % P
        sel_ev = selector_evaluate(yl, y_size, output_point[:yl])
        tmp = tau_s * (1 - sel_ev) % P
        psi_inv = inv(psi)
        c_ev = psi_inv * (c_adj_ev - c_pull_ev + tau_c * sel_ev - tmp) % P
        d_ev = psi_inv * (d_adj_ev - d_pull_ev + tau_d * sel_ev - tmp) % P
        output_evs = [p_folded_ev, c_pull_ev, d_pull_ev, c_ev, d_ev]
        transcript.write_scalars(output_evs)
        return dict(gamma=gamma, matrix=(output_point, output_evs), ac_c=ac_c_claims, ac_d=ac_d_claims)

    def verify(self, transcript, claims):
        point, evs = list(claims[0]), list(claims[1])
        evs[1] = (evs[1] - 1) % P
        xl, yl, dl, y_size = self.x_logsize, self.y_logsize, self.d_logsize, self.y_size
        r_y = point[:yl]
        assert len(point) == yl + dl + xl
        matrix_logsize, matrix_size = xl + yl, (1 << xl) * y_size
        psi, tau_c, tau_d, tau_s = transcript.challenge_vec(4, 512)
        gamma = transcript.challenge(128)
        suppression_total = 2 * ((1 << matrix_logsize) - matrix_size) % P * inv(tau_s) % P if tau_s else 0
        cd_claims, ac_c_claims, ac_d_claims = self._mainphase().verify(transcript, suppression_total)
        cd_claims = K.SplitAt(("HI", 0), 2).prove(transcript, cd_claims)
        gammas = make_gamma_pows(gamma, 5)
        ev_folded = (evs[0] + gammas[1] * evs[1] + gammas[2] * evs[2]) % P
        cd_point, cd_evs = cd_claims
        claim = (cd_evs[0] + gammas[1] * cd_evs[1] + gammas[2] * ev_folded) % P
        output_point = []
        for _ in range(matrix_logsize):
            combined = decompress_coefficients(transcript.read_scalars(3), claim)
            t = transcript.challenge(128)
            claim = evaluate_univar(combined, t)
            output_point.append(t)
        output_point.reverse()
        p_folded_ev, c_pull_ev, d_pull_ev, c_ev, d_ev = transcript.read_scalars(5)
        adj_p_folded_ev = (p_folded_ev - gamma) % P
        p_selector_prod_ev = adj_p_folded_ev * eq_trunc_evaluate(yl, y_size, r_y, output_point[:yl]) % P
        sel_ev = selector_evaluate(yl, y_size, output_point[:yl])
        tmp = tau_s * (1 - sel_ev) % P
        c_adj_ev = (c_pull_ev + psi * c_ev - tau_c * sel_ev + tmp) % P
        d_adj_ev = (d_pull_ev + psi * d_ev - tau_d * sel_ev + tmp) % P
        lhs = (eq_eval(cd_point, output_point) * ((c_adj_ev + d_adj_ev + gammas[1] * c_adj_ev % P * d_adj_ev) % P)
               + gammas[2] * (c_pull_ev * d_pull_ev % P * p_selector_prod_ev % P)) % P
        assert lhs == claim, "pushforward final check failed"
        return dict(gamma=gamma, matrix=(output_point, [p_folded_ev, c_pull_ev, d_pull_ev, c_ev, d_ev]), ac_c=ac_c_claims, ac_d=ac_d_claims)


# ---------------------------------------------------------------- multiopen reduction --------------------
class MultiOpenReduction:
    """multiopen_reduction.rs:43-117.  claims = [(point, ev)] * nargs."""

    def __init__(self, nvars, nargs):
        self.nvars, self.nargs = nvars, nargs

    def prove(self, transcript, claims, advice):
        gamma = transcript.challenge(128)
        fun = G.FoldedProd(gamma, self.nargs)
        folded = gamma_rlc(gamma, [c[1] for c in claims])
        advice = [list(a) for a in advice] + [eq_poly_sequence_last(c[0]) for c in claims]
        so = DenseSumcheckObjectSO(advice, fun, self.nvars, folded)
        (_, out_point), poly_evs = generic_sumcheck_prove(transcript, [2] * self.nvars, so.claim, so)
        evs = poly_evs[:self.nargs]
        transcript.write_scalars(evs)
        return (out_point, evs)

    def verify(self, transcript, claims):
        assert len(claims) == self.nargs
        gamma = transcript.challenge(128)
        fun = G.FoldedProd(gamma, self.nargs)
        folded = gamma_rlc(gamma, [c[1] for c in claims])
        claim, out_point = generic_sumcheck_verify(transcript, [2] * self.nvars, folded)
        evs = transcript.read_scalars(self.nargs)
        ext = list(evs) + [eq_eval(c[0], out_point) for c in claims]
        assert claim == fun.exec(ext) % P, "Final combinator check has failed."
        return (out_point, evs)


# ---------------------------------------------------------------- Knuckles opening -----------------------
class KnucklesOpening:
    """opening.rs:13-141.  claim = (commitment, point, ev)."""

    def __init__(self, key: KnucklesKey):
        self.key = key

    def prove(self, transcript, claim, advice):
        comm, point, ev_claim = claim
        pk = self.key
        t, opening = pk.compute_t(advice, point)
        assert opening == ev_claim
        t_comm = pk.kzg.commit(t)
        write_points(transcript, [t_comm])
        x = transcript.challenge(128)
        kx = x * pk.k % P
        t_x, p_x = ev(t, x), ev(advice, x)
        transcript.write_scalars([t_x, p_x])
        lam = transcript.challenge(128)
        padded = list(advice) + [0] * (len(t) - len(advice))
        p_lt = [(lam * b + a) % P for a, b in zip(padded, t)]
        p_lt_x_proof, _ = pk.kzg.open(p_lt, x)
        write_points(transcript, [p_lt_x_proof])
        t_kx_proof, t_kx = pk.kzg.open(t, kx)
        transcript.write_scalars([t_kx])
        write_points(transcript, [t_kx_proof])
        fin = transcript.challenge(128)
        p_lt_comm = CV.g1_add(CV.g1_mul(lam, t_comm), comm)
        p_lt_open = (t_x * lam + p_x) % P
        a0, b0 = pk.kzg.verify_reduce_to_pair(p_lt_comm, p_lt_x_proof, x, p_lt_open)
        a1, b1 = pk.kzg.verify_reduce_to_pair(t_comm, t_kx_proof, kx, t_kx)
        return (CV.g1_add(a0, CV.g1_mul(fin, a1)), CV.g1_add(b0, CV.g1_mul(fin, b1)))

    def verify(self, transcript, claim):
        comm, point, ev_claim = claim
        vk = self.key
        t_comm = read_points(transcript, 1)[0]
        x = transcript.challenge(128)
        kx = x * vk.k % P
        t_x, p_x = transcript.read_scalars(2)
        lam = transcript.challenge(128)
        p_lt_comm = CV.g1_add(CV.g1_mul(lam, t_comm), comm)
        p_lt_open = (t_x * lam + p_x) % P
        p_lt_x_proof = read_points(transcript, 1)[0]
        a0, b0 = vk.kzg.verify_reduce_to_pair(p_lt_comm, p_lt_x_proof, x, p_lt_open)
        t_kx = transcript.read_scalars(1)[0]
        t_kx_proof = read_points(transcript, 1)[0]
        a1, b1 = vk.kzg.verify_reduce_to_pair(t_comm, t_kx_proof, kx, t_kx)
        k_pow_n_1 = pow(vk.k, (1 << vk.num_vars) - 1, P)
        xpow, eq_ev = x, 1
        for i in range(vk.num_vars):
            r = point[vk.num_vars - i - 1]
            eq_ev = eq_ev * ((r + (1 - r) * xpow) % P) % P
            xpow = xpow * xpow % P
        lhs = (x * (t_kx - k_pow_n_1 * t_x) + xpow * ev_claim) % P
        rhs = x * p_x % P * eq_ev % P
        assert lhs == rhs, "knuckles opening equation failed"
        fin = transcript.challenge(128)
        return (CV.g1_add(a0, CV.g1_mul(fin, a1)), CV.g1_add(b0, CV.g1_mul(fin, b1)))


# ---------------------------------------------------------------- top level ------------------------------
class PippengerWG:
    """pippenger.rs:30-70"""

    def __init__(self, points, coefs, y_size, y_logsize, d_logsize, x_logsize, clm, key):
        self.beginning = PushForwardState(points, coefs, y_size, y_logsize, d_logsize, x_logsize, clm, key)
        self.ending = K.PippengerEndingWG(y_logsize, d_logsize, x_logsize, K.GlueSplit.witness(self.beginning.image))


def _g1_lincomb(coefs, pts):
    acc = None
    for c, p in zip(coefs, pts):
        acc = CV.g1_add(acc, CV.g1_mul(c, p))
    return acc


class Pippenger:
    """pippenger.rs:72-406"""

    def __init__(self, y_size, y_logsize, d_logsize, x_logsize, key: KnucklesKey, clm):
        assert x_logsize >= d_logsize and y_logsize >= clm
        self.key, self.clm = key, clm
        self.beginning = PushforwardProtocol(x_logsize, y_logsize, y_size, d_logsize)
        self.ending = K.PippengerBucketed(y_logsize, d_logsize, x_logsize)

    def _opening_inputs(self, transcript, fc, c, d, c_pull, d_pull):
        b, clm = self.beginning, self.clm
        matrix_pt, matrix_evs = fc["matrix"]
        p_folded_ev, c_pull_ev, d_pull_ev, c_ev, d_ev = matrix_evs
        gamma = fc["gamma"]
        p_folded_point = [0] * clm + list(matrix_pt[b.y_logsize:])
        ac_c_point = [0] * clm + list(fc["ac_c"][0])
        ac_d_point = [0] * (b.x_logsize + clm - b.d_logsize) + list(fc["ac_d"][0])
        combined_point = list(matrix_pt[b.y_logsize - clm:])
        multirow_evs = eq_poly_sequence_last(matrix_pt[:b.y_logsize - clm])
        c_comb, d_comb = _g1_lincomb(multirow_evs, c), _g1_lincomb(multirow_evs, d)
        cp_comb, dp_comb = _g1_lincomb(multirow_evs, c_pull), _g1_lincomb(multirow_evs, d_pull)
        u = transcript.challenge(512)
        us = make_gamma_pows(u, 4)
        combined_comm = _g1_lincomb([1, us[1], us[2], us[3]], [c_comb, d_comb, cp_comb, dp_comb])
        combined_ev = (c_ev + d_ev * us[1] + c_pull_ev * us[2] + d_pull_ev * us[3]) % P
        claims = [(p_folded_point, (p_folded_ev - gamma * gamma) % P), (ac_c_point, fc["ac_c"][1][0]),
                  (ac_d_point, fc["ac_d"][1][0]), (combined_point, combined_ev)]
        return claims, multirow_evs, us, combined_comm

    def prove(self, transcript, claims, state: PippengerWG):
        b, clm = self.beginning, self.clm
        st = state.beginning
        n_comms = -(-b.y_size // (1 << clm))
        assert len(st.c_comm) == n_comms and len(st.d_comm) == n_comms
        write_points(transcript, st.c_comm)
        write_points(transcript, st.d_comm)
        for pt in (st.p_0_comm, st.p_1_comm, st.ac_c_comm, st.ac_d_comm):
            write_points(transcript, [pt])
        claims = self.ending.prove(transcript, claims, state.ending)
        claims = K.GlueSplit().prove(transcript, claims)
        st.second_phase(claims[0])
        write_points(transcript, st.c_pull_comm)
        write_points(transcript, st.d_pull_comm)
        fc = b.prove(transcript, claims, st)
        gamma = fc["gamma"]
        oclaims, multirow_evs, us, combined_comm = self._opening_inputs(transcript, fc, st.c_comm, st.d_comm, st.c_pull_comm, st.d_pull_comm)
        x_size, y_size, xl = 1 << b.x_logsize, b.y_size, b.x_logsize
        cm = 1 << clm
        combined_witness = []
        for i in range(x_size * cm):
            x, y_rem = i % x_size, i >> xl
            ret = 0
            for y in range(y_size):
                if y % cm == y_rem:
                    idx = x + x_size * y
                    ret += multirow_evs[y // cm] * ((st.c[idx] + st.d[idx] * us[1] + st.c_pull[idx] * us[2] + st.d_pull[idx] * us[3]) % P)
            combined_witness.append(ret % P)
        nv = xl + clm
        mw = [[(a + gamma * bb) % P for a, bb in zip(st.p_0, st.p_1)], list(st.ac_c), list(st.ac_d), combined_witness]
        mw = [pad_vector(a, nv, 0) for a in mw]
        mo_point, mo_evs = MultiOpenReduction(nv, 4).prove(transcript, oclaims, mw)
        q = transcript.challenge(128)
        qs = make_gamma_pows(q, 4)
        folded_comm = _g1_lincomb(qs, [CV.g1_add(st.p_0_comm, CV.g1_mul(gamma, st.p_1_comm)), st.ac_c_comm, st.ac_d_comm, combined_comm])
        folded_witness = [(mw[0][i] * qs[0] + mw[1][i] * qs[1] + mw[2][i] * qs[2] + mw[3][i] * qs[3]) % P for i in range(1 << nv)]
        return KnucklesOpening(self.key).prove(transcript, (folded_comm, mo_point, gamma_rlc(q, mo_evs)), folded_witness)

    def verify(self, transcript, claims):
        b, clm = self.beginning, self.clm
        n_comms = -(-b.y_size // (1 << clm))
        c = read_points(transcript, n_comms)
        d = read_points(transcript, n_comms)
        p_0, p_1, ac_c, ac_d = (read_points(transcript, 1)[0] for _ in range(4))
        claims = self.ending.verify(transcript, claims)
        claims = K.GlueSplit().prove(transcript, claims)
        c_pull = read_points(transcript, n_comms)
        d_pull = read_points(transcript, n_comms)
        fc = b.verify(transcript, claims)
        gamma = fc["gamma"]
        oclaims, _, _, combined_comm = self._opening_inputs(transcript, fc, c, d, c_pull, d_pull)
        mo_point, mo_evs = MultiOpenReduction(b.x_logsize + clm, 4).verify(transcript, oclaims)
        q = transcript.challenge(128)
        qs = make_gamma_pows(q, 4)
        folded_comm = _g1_lincomb(qs, [CV.g1_add(p_0, CV.g1_mul(gamma, p_1)), ac_c, ac_d, combined_comm])
        pair = KnucklesOpening(self.key).verify(transcript, (folded_comm, mo_point, gamma_rlc(q, mo_evs)))
        self.key.kzg.verify_pair(pair)


def pippenger_config(d_logsize, x_logsize, num_bits, clm):  # build_pippenger_data, pippenger.rs:462-497
    y_size = (num_bits + d_logsize - 1) // d_logsize
    y_logsize = (y_size - 1).bit_length()  # ark_std::log2 = ceil(log2)
    return dict(y_size=y_size, y_logsize=y_logsize, d_logsize=d_logsize, x_logsize=x_logsize, clm=clm)


def run_pippenger(transcript, points, coefs, cfg, r, key):
    """benchutils::run_pippenger  pippenger.rs:499-559.  Returns (dense_output, claims)."""
    y_size, yl, dl, xl, clm = cfg["y_size"], cfg["y_logsize"], cfg["d_logsize"], cfg["x_logsize"], cfg["clm"]
    wg = PippengerWG(points, coefs, y_size, yl, dl, xl, clm, key)
    dense_output = K.triangle_last_step(wg.ending.last(), yl + dl - 2 - yl)
    claims = (list(r), [evaluate_poly(o, r) for o in dense_output])
    Pippenger(y_size, yl, dl, xl, key, clm).prove(transcript, claims, wg)
    return dense_output, claims


def verify_pippenger(transcript, cfg, dense_output, claims, key, expected=None):
    """benchutils::verify_pippenger  pippenger.rs:561-606 (+ build_points, src/utils.rs:290-322)."""
    y_size, yl, dl, xl, clm = cfg["y_size"], cfg["y_logsize"], cfg["d_logsize"], cfg["x_logsize"], cfg["clm"]
    Pippenger(y_size, yl, dl, xl, key, clm).verify(transcript, claims)
    assert (dl + 1) * 3 == len(dense_output)
    chunks = [dense_output[i:i + 3] for i in range(0, len(dense_output), 3)]
    pts = [[(ch[0][i], ch[1][i], ch[2][i]) for i in range(len(ch[0]))] for ch in chunks]  # projective (X, Y, Z)
    transposed = []
    for idx in range(len(pts[0])):
        for i in range(1, len(pts)):
            transposed.append(pts[i][idx])
    acc = (0, 1, 1)
    for pt in reversed(transposed):
        acc = CV.te_add_proj(acc, acc)
        acc = CV.te_add_proj(acc, pt)
    res = CV.te_to_affine(acc)
    if expected is not None:
        assert res == expected, "proved MSM result differs from the expected one"
    return res
