"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

Big-int restatement of the reference's sumcheck engine:

  src/cleanup/protocols/sumcheck.rs:14-44      compress/decompress_coefficients, evaluate_univar
  src/cleanup/protocols/sumcheck.rs:101-128    GenericSumcheckProtocol::{prove,verify}
  src/cleanup/protocols/sumcheck.rs:136-235    ExampleSumcheckObjectSO (the reference's own naive oracle)
  src/cleanup/protocols/sumcheck.rs:241-347    DenseSumcheckObjectSO
  src/cleanup/protocols/sumcheck.rs:591-602    gamma_rlc
  src/cleanup/protocols/sumcheck.rs:831-889    DenseEqSumcheck
  src/cleanup/protocols/sumchecks/dense_eq.rs:43-237     DenseDeg2SumcheckObject(SO), DenseDeg2Sumcheck
  src/cleanup/protocols/sumchecks/vecvec_eq.rs:53-467    VecVecDeg2 objects, UnivarFormat::from12, VecVecDeg2Sumcheck
  src/cleanup/polys/vecvec.rs:20-147, 178-206, 393-441   EQPolyPointParts, EQPolyData, VecVecPolynomial, make_21, bind_21
  src/cleanup/polys/dense.rs:39-61, 99-112               bind_21, make_21 (dense)
  src/utils.rs:126-154, 189-262                          make_gamma_pows, zip_with_gamma, eq_eval, eq tables
  liblasso@925a7a74 UniPoly::{from_evals, as_vec, evaluate} (not vendored): interpolation on nodes
  0..deg, coefficients low -> high -- the interpolant is unique, so Lagrange is an exact restatement.

Elements are python ints in [0, r) (standard form).
"""
from __future__ import annotations

from .field import P


def inv(x):
    return pow(x % P, -1, P)


# ------------------------------------------------------------------ univariate helpers ----
def unipoly_from_evals(evals):
    """coefficients (low -> high) of the unique poly of degree < len(evals) with p(i) = evals[i]."""
    n = len(evals)
    coeffs = [0] * n
    for i in range(n):
        # Lagrange basis l_i(X) = prod_{j != i} (X - j)/(i - j)
        num = [1]
        den = 1
        for j in range(n):
            if j == i:
                continue
            new = [0] * (len(num) + 1)
            for k, c in enumerate(num):
                new[k] = (new[k] - j * c) % P
                new[k + 1] = (new[k + 1] + c) % P
            num = new
            den = den * (i - j) % P
        s = evals[i] * inv(den) % P
        for k, c in enumerate(num):
            coeffs[k] = (coeffs[k] + c * s) % P
    return coeffs


def evaluate_univar(coeffs, x):  # sumcheck.rs:33-44
    ret = 0
    for c in reversed(coeffs):
        ret = (ret * x + c) % P
    return ret


def compress_coefficients(coeffs):  # sumcheck.rs:27-31
    return [coeffs[0]] + list(coeffs[2:])


def decompress_coefficients(c, s):  # sumcheck.rs:14-25
    sum_minus_l = (2 * c[0] + sum(c[1:])) % P
    return [c[0], (s - sum_minus_l) % P] + list(c[1:])


def gamma_rlc(gamma, vals):  # sumcheck.rs:591-602 == utils.rs:137-148 zip_with_gamma
    if not vals:
        return 0
    ret = vals[-1]
    for v in reversed(vals[:-1]):
        ret = (ret * gamma + v) % P
    return ret


zip_with_gamma = gamma_rlc


def make_gamma_pows(gamma, count):  # utils.rs:126-135  (always at least [1, gamma])
    g = [1, gamma % P]
    for i in range(2, count):
        g.append(g[i - 1] * gamma % P)
    return g


def eq_eval(p1, p2):  # utils.rs:150-154
    assert len(p1) == len(p2)
    r = 1
    for a, b in zip(p1, p2):
        r = r * ((1 - a - b + 2 * a * b) % P) % P
    return r


def eq_poly_sequence_from_multiplier(mult, pt):  # utils.rs:222-250
    ret = [[mult % P]]
    for i in range(1, len(pt) + 1):
        last, m_ = ret[i - 1], pt[i - 1]
        inc = [0] * (1 << i)
        for j, w in enumerate(last):
            m = m_ * w % P
            inc[2 * j] = (w - m) % P
            inc[2 * j + 1] = m
        ret.append(inc)
    return ret


def eq_poly_sequence(pt):
    return eq_poly_sequence_from_multiplier(1, pt)


def eq_poly_sequence_last(pt):
    return eq_poly_sequence(pt)[-1]


def padded_eq_poly_sequence(padding_size, pt):  # utils.rs:189-220
    l = len(pt)
    ret = [[1]]
    for i in range(1, padding_size + 1):
        ret.append([ret[i - 1][0] * (1 - pt[i - 1]) % P])
    for i in range(padding_size + 1, l + 1):
        last, m_ = ret[i - 1], pt[i - 1]
        inc = [0] * (1 << (i - padding_size))
        for j in range(1 << (i - 1 - padding_size)):
            w = last[j]
            m = m_ * w % P
            inc[2 * j] = (w - m) % P
            inc[2 * j + 1] = m
        ret.append(inc)
    return ret


def eq_sum(pt, k):  # utils.rs:265-291
    n = len(pt)
    if k >= (1 << n):
        assert k == 1 << n
        return 1
    mult, acc = 1, 0
    for i in range(n):
        left_bit = k >> (n - i - 1)
        old = mult
        if left_bit == 1:
            mult = mult * pt[i] % P
            acc = (acc + old - mult) % P
        else:
            mult = mult * (1 - pt[i]) % P
        k -= left_bit << (n - i - 1)
    return acc


def evaluate_poly(poly, pt):  # cleanup/utils/arith.rs:6-9
    e = eq_poly_sequence_last(pt)
    assert len(e) == len(poly)
    return sum(a * b for a, b in zip(poly, e)) % P


def log_2(n: int) -> int:
    """liblasso Math::log_2: exact for powers of two, ceil otherwise (0 -> 0)."""
    if n <= 1:
        return 0
    return (n - 1).bit_length()


def bind_dense_poly(poly, t):  # sumcheck.rs:160-163
    return [(poly[2 * i] + t * (poly[2 * i + 1] - poly[2 * i])) % P for i in range(len(poly) // 2)]


# ------------------------------------------------------------------ sumcheck objects ------
class ExampleSumcheckObjectSO:
    """sumcheck.rs:136-235 -- the reference's own naive object (evaluates at 0..deg)."""

    def __init__(self, polys, f, num_vars):
        assert len(polys) == f.n_ins
        for p in polys:
            assert len(p) == 1 << num_vars
        self.polys = [list(p) for p in polys]
        self.f, self.num_vars, self.round_idx = f, num_vars, 0
        self.cached = None
        self.chals = []

    def claim(self):
        n = 1 << (self.num_vars - self.round_idx)
        return sum(self.f.exec([p[i] for p in self.polys]) for i in range(n)) % P

    def unipoly(self):
        assert self.round_idx < self.num_vars
        if self.cached is not None:
            return list(self.cached)
        half = 1 << (self.num_vars - self.round_idx - 1)
        d = self.f.deg
        acc = [0] * (d + 1)
        for i in range(half):
            a0 = [p[2 * i] for p in self.polys]
            a1 = [p[2 * i + 1] for p in self.polys]
            acc[0] += self.f.exec(a0)
            acc[1] += self.f.exec(a1)
            dif = [(y - x) % P for x, y in zip(a0, a1)]
            args = a1
            for s in range(2, d + 1):
                args = [(x + y) % P for x, y in zip(args, dif)]
                acc[s] += self.f.exec(args)
        self.cached = unipoly_from_evals([a % P for a in acc])
        return list(self.cached)

    def bind(self, t):
        assert self.round_idx < self.num_vars
        if self.cached is None:
            raise RuntimeError("should evaluate unipoly before binding")
        self.chals.append(t)
        self.polys = [bind_dense_poly(p, t) for p in self.polys]
        self.round_idx += 1
        self.cached = None

    def final_evals(self):
        assert self.round_idx == self.num_vars
        return [p[0] for p in self.polys]


class DenseSumcheckObjectSO:
    """sumcheck.rs:241-347.  evals at 1..deg, eval(0) = claim - eval(1)."""

    def __init__(self, polys, f, num_vars, claim_hint):
        assert len(polys) == f.n_ins
        for p in polys:
            assert len(p) == 1 << num_vars
        self.polys = [list(p) for p in polys]
        self.f, self.num_vars, self.round_idx = f, num_vars, 0
        self.cached = None
        self.claim = claim_hint % P
        self.chals = []
        self.last_evals = None  # [p(0), ..., p(deg)] of the last unipoly (what the device returns)

    def unipoly(self):
        assert self.round_idx < self.num_vars
        if self.cached is not None:
            return list(self.cached)
        half = 1 << (self.num_vars - self.round_idx - 1)
        d = self.f.deg
        acc = [0] * d
        for i in range(half):
            a0 = [p[2 * i] for p in self.polys]
            args = [p[2 * i + 1] for p in self.polys]
            acc[0] += self.f.exec(args)
            dif = [(y - x) % P for x, y in zip(a0, args)]
            for s in range(1, d):
                args = [(x + y) % P for x, y in zip(args, dif)]
                acc[s] += self.f.exec(args)
        total = [0] + [a % P for a in acc]
        total[0] = (self.claim - total[1]) % P
        self.last_evals = total
        self.cached = unipoly_from_evals(total)
        return list(self.cached)

    def bind(self, t):
        assert self.round_idx < self.num_vars
        if self.cached is None:
            raise RuntimeError("should evaluate unipoly before binding")
        self.chals.append(t)
        self.polys = [bind_dense_poly(p, t) for p in self.polys]
        self.round_idx += 1
        self.claim = evaluate_univar(self.cached, t)
        self.cached = None

    def final_evals(self):
        assert self.round_idx == self.num_vars
        return [p[0] for p in self.polys]


def univar_from12(p1, p2, eq1, previous_claim):
    """vecvec_eq.rs:197-216 UnivarFormat::from12."""
    eq0 = (1 - eq1) % P
    eq2 = (2 * eq1 - eq0) % P
    eq3 = (2 * eq2 - eq1) % P
    prod1 = p1 * eq1 % P
    prod0 = (previous_claim - prod1) % P
    p0 = prod0 * inv(eq0) % P
    p3 = (3 * p2 - 3 * p1 + p0) % P
    evals = [prod0, prod1, p2 * eq2 % P, p3 * eq3 % P]
    return unipoly_from_evals(evals), evals


def dense_make_21(v):  # dense.rs:99-112
    for i in range(len(v) // 2):
        v[2 * i] = (2 * v[2 * i + 1] - v[2 * i]) % P


def dense_bind_21(v, t):
    """dense.rs:54-61 (feature `parallel`, the README build): output length len/2, no re-padding.
    (The non-parallel variant dense.rs:39-52 additionally zero-pads odd lengths; both agree whenever
    every intermediate length is even, which holds for the power-of-two tables the protocols pass.)"""
    tm1 = (t - 1) % P
    assert len(v) % 2 == 0
    return [(v[2 * i + 1] + tm1 * (v[2 * i] - v[2 * i + 1])) % P for i in range(len(v) // 2)]


class DenseDeg2SumcheckObjectSO:
    """dense_eq.rs:62-173.  `polys` may be shorter than 2^n (implicit zero padding)."""

    def __init__(self, polys, func, gamma_pows, claim, point):
        self.polys = [list(p) for p in polys]
        self.func, self.gamma_pows = func, list(gamma_pows)
        self.claim = claim % P
        self.point = list(point)
        self.eq_poly_data = eq_poly_sequence(self.point[:-1])
        self.multiplier = 1
        self.current_point = []
        self.cached = None
        self.last_evals = None

    @classmethod
    def rlc(cls, polys, func, claims, point, gamma):  # dense_eq.rs:43-60
        gp = make_gamma_pows(gamma, func.n_outs)
        claim = claims[0]
        for i in range(1, len(claims)):
            claim = (claim + gp[i] * claims[i]) % P
        return cls(polys, func, gp, claim, point)

    def unipoly(self):
        if self.cached is not None:
            raise RuntimeError("unipoly called twice")
        for v in self.polys:
            dense_make_21(v)
        n_out = self.func.n_outs
        pad_results = self.func.exec([0] * len(self.polys))
        sum2, sum1 = [0] * n_out, [0] * n_out
        eq = self.eq_poly_data[-1]
        eq_sum_ = 0
        for idx in range(len(self.polys[0]) // 2):
            o2 = self.func.exec([p[2 * idx] for p in self.polys])
            o1 = self.func.exec([p[2 * idx + 1] for p in self.polys])
            for i in range(n_out):
                sum2[i] += o2[i] * eq[idx]
                sum1[i] += o1[i] * eq[idx]
            eq_sum_ += eq[idx]
        trailing = (1 - eq_sum_) % P
        for i in range(n_out):
            sum2[i] = (sum2[i] + pad_results[i] * trailing) % P
            sum1[i] = (sum1[i] + pad_results[i] * trailing) % P
        total2, total1 = sum2[0], sum1[0]
        for i in range(1, n_out):
            total2 += sum2[i] * self.gamma_pows[i]
            total1 += sum1[i] * self.gamma_pows[i]
        total2 = total2 * self.multiplier % P
        total1 = total1 * self.multiplier % P
        self.last_p12 = (total1, total2)
        self.cached, self.last_evals = univar_from12(total1, total2, self.point[-1], self.claim)
        return list(self.cached)

    def bind(self, t):
        q = self.point[-1]
        self.multiplier = self.multiplier * ((1 - q - t + 2 * q * t) % P) % P
        self.polys = [dense_bind_21(v, t) for v in self.polys]
        self.current_point.append(t)
        self.eq_poly_data.pop()
        self.point.pop()
        self.claim = evaluate_univar(self.cached, t)
        self.cached = None

    def final_evals(self):
        return [p[0] for p in self.polys]


class VecVecPolynomial:
    """polys/vecvec.rs:149-206."""

    def __init__(self, data, row_pad, col_pad, row_logsize, col_logsize, unchecked=False):
        assert len(data) <= (1 << col_logsize)
        self.data = [list(r) for r in data]
        if not unchecked:
            for r in self.data:
                assert len(r) <= 1 << row_logsize
                if len(r) % 2 == 1:
                    r.append(row_pad)
        self.row_pad, self.col_pad = row_pad % P, col_pad % P
        self.row_logsize, self.col_logsize = row_logsize, col_logsize

    def clone(self):
        return VecVecPolynomial(self.data, self.row_pad, self.col_pad, self.row_logsize, self.col_logsize, unchecked=True)

    def num_vars(self):
        return self.row_logsize + self.col_logsize

    def make_21(self):  # vecvec.rs:400-413
        for r in self.data:
            for i in range(len(r) // 2):
                r[2 * i] = (2 * r[2 * i + 1] - r[2 * i]) % P

    def bind_21(self, t):  # vecvec.rs:420-441
        tm1 = (t - 1) % P
        for k, r in enumerate(self.data):
            h = len(r) // 2
            new = [(r[2 * i + 1] + tm1 * (r[2 * i] - r[2 * i + 1])) % P for i in range(h)]
            if h % 2 == 1:
                new.append(self.row_pad)
            self.data[k] = new
        self.row_logsize -= 1

    def vec(self):  # vecvec.rs:446-461
        ret = []
        for r in range(1 << self.col_logsize):
            for c in range(1 << self.row_logsize):
                if r >= len(self.data):
                    ret.append(self.col_pad)
                elif c >= len(self.data[r]):
                    ret.append(self.row_pad)
                else:
                    ret.append(self.data[r][c])
        return ret


class EQPolyData:
    """polys/vecvec.rs:20-147 (EQPolyPointParts + EQPolyData)."""

    def __init__(self, point, col_logsize, max_row_len):
        max_segment_logsize = log_2(max_row_len)
        self.padded_vars_idx = col_logsize
        self.segment_vars_idx = len(point) - max_segment_logsize
        self.binding_var_idx = len(point) - 1
        self.point = list(point)
        self.row_eq_coefs = eq_poly_sequence_last(self.point[0:col_logsize])
        tails, acc = [], 0
        for v in reversed(self.row_eq_coefs):
            acc = (acc + v) % P
            tails.append(acc)
        tails.reverse()
        self.row_eq_coefs_tail_sums = tails
        lo, hi = self.padded_vars_range()
        rlo, rhi = self.row_vars_range()
        self.row_eq_poly_seq = padded_eq_poly_sequence(max(0, hi - lo), self.point[rlo:rhi])
        self.row_eq_poly_prefix_seq = []
        for v in self.row_eq_poly_seq:
            acc = [0]
            for x in v:
                acc.append((acc[-1] + x) % P)
            self.row_eq_poly_prefix_seq.append(acc)
        self.multiplier = 1
        self.already_bound_vars = 0

    def padded_vars_range(self):
        return self.padded_vars_idx, min(self.segment_vars_idx, self.binding_var_idx)

    def row_vars_range(self):
        return self.padded_vars_idx, max(self.segment_vars_idx, self.binding_var_idx)

    def bind(self, t):
        q = self.point[self.binding_var_idx]
        self.multiplier = self.multiplier * ((1 - q - t + 2 * q * t) % P) % P
        if self.binding_var_idx is not None:
            self.binding_var_idx = None if self.binding_var_idx == 0 else self.binding_var_idx - 1
        self.already_bound_vars += 1

    def get_segment_evals(self, segment_len):
        return self.row_eq_poly_seq[len(self.row_eq_poly_seq) - 1 - self.already_bound_vars][0:segment_len]

    def get_trailing_sum(self, segment_len):
        s = self.row_eq_poly_prefix_seq[len(self.row_eq_poly_prefix_seq) - 1 - self.already_bound_vars][segment_len]
        return (1 - s) % P


class VecVecDeg2SumcheckObjectSO:
    """vecvec_eq.rs:74-398: sparse stage (VecVecDeg2LoSumcheckObjectSO) then hand-off to the dense
    object over EqWrapper(GammaWrapper(func, gamma))."""

    def __init__(self, polys, func, gamma_pows, claim, point, col_logsize):
        from .gates import EqWrapper, GammaWrapper  # noqa: F401
        self.polys = [p.clone() for p in polys]
        self.func, self.gamma_pows = func, list(gamma_pows)
        self.claim_ = claim % P
        self.eq = EQPolyData(point, col_logsize, max(len(r) for r in self.polys[0].data))
        self.current_point = []
        self.cached = None
        self.dense = None  # DenseSumcheckObjectSO once handed off
        self.last_evals = None

    @classmethod
    def rlc(cls, polys, func, claims, point, num_vertical_vars, gamma):  # vecvec_eq.rs:53-71
        gp = make_gamma_pows(gamma, func.n_outs)
        claim = claims[0]
        for i in range(1, len(claims)):
            claim = (claim + gp[i] * claims[i]) % P
        return cls(polys, func, gp, claim, point, num_vertical_vars)

    @property
    def claim(self):
        return self.dense.claim if self.dense is not None else self.claim_

    def unipoly(self):
        if self.dense is not None:
            u = self.dense.unipoly()
            self.last_evals = self.dense.last_evals
            return u
        if self.cached is not None:
            raise RuntimeError("unipoly called twice")
        for p in self.polys:
            p.make_21()
        n_out = self.func.n_outs
        pad_results = self.func.exec([p.row_pad for p in self.polys])
        col_pad_results = self.func.exec([p.col_pad for p in self.polys])
        sum2, sum1 = [0] * n_out, [0] * n_out
        row_count = len(self.polys[0].data)
        for row_idx in range(row_count):
            l2, l1 = [0] * n_out, [0] * n_out
            segment_len = len(self.polys[0].data[row_idx]) // 2
            eq = self.eq.get_segment_evals(segment_len)
            for idx in range(segment_len):
                o2 = self.func.exec([p.data[row_idx][2 * idx] for p in self.polys])
                o1 = self.func.exec([p.data[row_idx][2 * idx + 1] for p in self.polys])
                for i in range(n_out):
                    l2[i] += o2[i] * eq[idx]
                    l1[i] += o1[i] * eq[idx]
            trailing = self.eq.get_trailing_sum(segment_len)
            vmul = self.eq.row_eq_coefs[row_idx]
            for i in range(n_out):
                sum2[i] = (sum2[i] + (l2[i] + pad_results[i] * trailing) * vmul) % P
                sum1[i] = (sum1[i] + (l1[i] + pad_results[i] * trailing) * vmul) % P
        if row_count < (1 << self.eq.padded_vars_idx):
            for i, out in enumerate(col_pad_results):
                res = out * self.eq.row_eq_coefs_tail_sums[row_count] % P
                sum2[i] = (sum2[i] + res) % P
                sum1[i] = (sum1[i] + res) % P
        total2, total1 = sum2[0], sum1[0]
        for i in range(1, n_out):
            total2 += sum2[i] * self.gamma_pows[i]
            total1 += sum1[i] * self.gamma_pows[i]
        total2 = total2 * self.eq.multiplier % P
        total1 = total1 * self.eq.multiplier % P
        self.last_p12 = (total1, total2)
        self.cached, self.last_evals = univar_from12(total1, total2, self.eq.point[self.eq.binding_var_idx], self.claim_)
        return list(self.cached)

    def bind(self, t):
        from .gates import EqWrapper, GammaWrapper
        if self.dense is not None:
            self.dense.bind(t)
            return
        if self.eq.binding_var_idx > self.eq.padded_vars_idx:  # vecvec_eq.rs:235
            for p in self.polys:
                p.bind_21(t)
            self.current_point.append(t)
            self.eq.bind(t)
            self.claim_ = evaluate_univar(self.cached, t)
            self.cached = None
            return
        # bind_into_dense (vecvec_eq.rs:157-190)
        tm1 = (t - 1) % P
        n = 1 << self.eq.padded_vars_idx
        polys = []
        for p in self.polys:
            col = []
            for r in p.data:
                if len(r) == 0:
                    col.append(p.row_pad)
                elif len(r) == 2:
                    col.append((r[1] + tm1 * (r[0] - r[1])) % P)
                else:
                    raise AssertionError("unreachable")
            col += [p.col_pad] * (n - len(col))
            polys.append(col[:n])
        q = self.eq.point[self.eq.binding_var_idx]
        mult = self.eq.multiplier * ((1 - q - t + 2 * q * t) % P) % P
        polys.append(eq_poly_sequence_from_multiplier(mult, self.eq.point[0:self.eq.padded_vars_idx])[-1])
        f = EqWrapper(GammaWrapper(self.func, self.gamma_pows[1]))
        self.dense = DenseSumcheckObjectSO(polys, f, self.eq.padded_vars_idx, evaluate_univar(self.cached, t))
        self.cached = None

    def final_evals(self):
        assert self.dense is not None
        return self.dense.final_evals()


# ------------------------------------------------------------------ protocols -------------
def generic_sumcheck_prove(transcript, degrees, claim, so):
    """sumcheck.rs:101-123.  Returns ((claim, point), final_evals)."""
    r = []
    for d in degrees:
        poly = so.unipoly()
        msg = compress_coefficients(poly)
        assert len(msg) == d
        transcript.write_scalars(msg)
        x = transcript.challenge(128)
        r.append(x)
        so.bind(x)
        claim = evaluate_univar(poly, x)
    r.reverse()
    return (claim, r), so.final_evals()


def generic_sumcheck_verify(transcript, degrees, claim):
    """sumcheck.rs:63-77."""
    r = []
    for d in degrees:
        msg = transcript.read_scalars(d)
        poly = decompress_coefficients(msg, claim)
        x = transcript.challenge(128)
        r.append(x)
        claim = evaluate_univar(poly, x)
    r.reverse()
    return claim, r


class DenseEqSumcheck:
    """sumcheck.rs:831-889: DenseEqSumcheckObject::rlc -> DenseSumcheckObjectSO over
    EqWrapper(GammaWrapper(f, gamma)) with the materialised eq table appended."""

    def __init__(self, f, num_vars):
        self.f, self.num_vars = f, num_vars

    def make_so(self, polys, point, evs, gamma):
        from .gates import EqWrapper, GammaWrapper
        polys = [list(p) for p in polys] + [eq_poly_sequence_last(point)]
        return DenseSumcheckObjectSO(polys, EqWrapper(GammaWrapper(self.f, gamma)), len(point), gamma_rlc(gamma, evs))

    def prove(self, transcript, claims, advice):
        point, evs = claims
        gamma = transcript.challenge(128)
        so = self.make_so(advice, point, evs, gamma)
        (_, out_point), poly_evs = generic_sumcheck_prove(transcript, [self.f.deg + 1] * self.num_vars, so.claim, so)
        poly_evs = poly_evs[:-1]
        transcript.write_scalars(poly_evs)
        return (out_point, poly_evs)

    def verify(self, transcript, claims):
        point, evs = claims
        gamma = transcript.challenge(128)
        folded = gamma_rlc(gamma, evs)
        ev, out_point = generic_sumcheck_verify(transcript, [self.f.deg + 1] * self.num_vars, folded)
        poly_evs = transcript.read_scalars(self.f.n_ins)
        lhs = gamma_rlc(gamma, self.f.exec(poly_evs)) * eq_eval(point, out_point) % P
        assert lhs == ev, "Final combinator check has failed."
        return (out_point, poly_evs)


class DenseDeg2Sumcheck:
    """dense_eq.rs:176-237."""

    def __init__(self, f, num_vars):
        self.f, self.num_vars = f, num_vars

    def prove(self, transcript, claims, advice):
        assert self.f.deg == 2
        point, evs = claims
        gamma = transcript.challenge(128)
        so = DenseDeg2SumcheckObjectSO.rlc(advice, self.f, evs, point, gamma)
        (_, out_point), poly_evs = generic_sumcheck_prove(transcript, [3] * self.num_vars, so.claim, so)
        transcript.write_scalars(poly_evs)
        return (out_point, poly_evs)

    def verify(self, transcript, claims):
        point, evs = claims
        gamma = transcript.challenge(128)
        folded = zip_with_gamma(gamma, evs)
        ev, out_point = generic_sumcheck_verify(transcript, [3] * self.num_vars, folded)
        poly_evs = transcript.read_scalars(self.f.n_ins)
        lhs = zip_with_gamma(gamma, self.f.exec(poly_evs)) * eq_eval(point, out_point) % P
        assert lhs == ev, "Final combinator check has failed."
        return (out_point, poly_evs)


class VecVecDeg2Sumcheck:
    """vecvec_eq.rs:400-467."""

    def __init__(self, f, num_vars, num_vertical_vars):
        self.f, self.num_vars, self.num_vertical_vars = f, num_vars, num_vertical_vars

    def prove(self, transcript, claims, advice):
        assert self.f.deg == 2
        point, evs = claims
        gamma = transcript.challenge(128)
        so = VecVecDeg2SumcheckObjectSO.rlc(advice, self.f, evs, point, self.num_vertical_vars, gamma)
        (_, out_point), poly_evs = generic_sumcheck_prove(transcript, [3] * self.num_vars, so.claim, so)
        poly_evs = poly_evs[:-1]
        transcript.write_scalars(poly_evs)
        return (out_point, poly_evs)

    def verify(self, transcript, claims):
        point, evs = claims
        gamma = transcript.challenge(128)
        folded = zip_with_gamma(gamma, evs)
        ev, out_point = generic_sumcheck_verify(transcript, [3] * self.num_vars, folded)
        poly_evs = transcript.read_scalars(self.f.n_ins)
        lhs = zip_with_gamma(gamma, self.f.exec(poly_evs)) * eq_eval(point, out_point) % P
        assert lhs == ev, "Final combinator check has failed."
        return (out_point, poly_evs)


class BareSumcheckSO:
    """sumcheck.rs:646-691 (single-output gate, no eq)."""

    def __init__(self, f, num_vars):
        self.f, self.num_vars = f, num_vars

    def prove(self, transcript, sum_claim, so):
        (_, point), poly_evs = generic_sumcheck_prove(transcript, [self.f.deg] * self.num_vars, sum_claim, so)
        transcript.write_scalars(poly_evs)
        return (point, poly_evs)

    def verify(self, transcript, sum_claim):
        ev, point = generic_sumcheck_verify(transcript, [self.f.deg] * self.num_vars, sum_claim)
        poly_evs = transcript.read_scalars(self.f.n_ins)
        assert self.f.exec(poly_evs) == ev, "Final combinator check has failed."
        return (point, poly_evs)
