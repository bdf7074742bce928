"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

Old API of the reference, restated for `Shape::full` tables (one Data fragment, no constants
-- the shape benches/bintree.rs:37-47 builds):

  FragmentedLincomb::{split, bind, unipoly, final_evals}   src/protocol/sumcheck.rs:36-156
  FragmentedPoly::{split, bind_from}                        src/polynomial/fragmented.rs:676-741, 745-751
  EqPoly::{materialize_split, bind}                         src/copoly.rs:581-635
  make_gamma_pows_legacy / make_folded_f                    src/utils.rs:104-113, src/protocol/sumcheck.rs:674-701

Differences to the new API that matter for parity: the eq factor is a MATERIALISED half table re-built from scratch every
round (one inversion per round), the round polynomial is evaluated at ALL nodes 0..degree+1 (no claim shortcut) and ALL its
coefficients go to the transcript.  The polynomial itself is the one DenseSumcheckObjectSO over
EqWrapper(GammaWrapper(f)) produces, which is what tests/test_gpu_oldapi.py checks on the device.
"""
from __future__ import annotations

from .field import P
from .sumcheck import eq_poly_sequence_from_multiplier, unipoly_from_evals


def make_gamma_pows_legacy(num_claims, gamma):  # utils.rs:104-113
    pows = [1, gamma % P]
    for i in range(2, num_claims):
        pows.append(pows[i - 1] * gamma % P)
    return pows


def make_folded_f(claim_evs, gamma_pows, exec_f, num_i):
    """claim_evs: per claimed point the list of (output index, value) pairs (MultiEvalClaim::evs); sumcheck.rs:674-701"""

    def folded(args):
        ins, eqs = args[:num_i], args[num_i:]
        out = exec_f(ins)
        acc, i = 0, 0
        for j, evs in enumerate(claim_evs):
            inner = 0
            for (o, _) in evs:
                inner += out[o] * gamma_pows[i]
                i += 1
            acc += inner % P * eqs[j]
        return acc % P

    return folded


class EqPolyFull:
    """EqPoly over a full shape: `values` tables only (copoly.rs:570-635)"""

    def __init__(self, point):
        self.point, self.multiplier = [p % P for p in point], 1

    def materialize_split(self):  # copoly.rs:600-635
        point = list(self.point)
        m1 = point.pop()
        m0 = (1 - m1) % P
        n = 1 << len(point)
        if m0 == 0:
            b = eq_poly_sequence_from_multiplier(m1 * self.multiplier % P, point)[-1] if point else [m1 * self.multiplier % P]
            return [0] * n, b
        m = m1 * pow(m0, -1, P) % P
        a = eq_poly_sequence_from_multiplier(m0 * self.multiplier % P, point)[-1] if point else [m0 * self.multiplier % P]
        return a, [x * m % P for x in a]

    def bind(self, value):  # copoly.rs:581-588
        p0 = self.point.pop()
        self.multiplier = self.multiplier * ((p0 * value + (1 - p0) * (1 - value)) % P) % P


class FragmentedLincombFull:
    def __init__(self, polys, points, folded_f, degree):
        self.polys = [[v % P for v in p] for p in polys]
        self.copolys = [EqPolyFull(pt) for pt in points]
        self.folded_f, self.degree = folded_f, degree

    def _split(self):  # fragmented.rs:676-741 on one Data fragment: even entries left, odd entries right
        l = [p[0::2] for p in self.polys]
        r = [p[1::2] for p in self.polys]
        lc, rc = zip(*[c.materialize_split() for c in self.copolys])
        return l, r, list(lc), list(rc)

    def unipoly_evals(self):
        """values at the nodes 0 .. degree + 1 (sumcheck.rs:96-146)"""
        l, r, lc, rc = self._split()
        pd = [[(y - x) % P for x, y in zip(a, b)] for a, b in zip(l, r)]
        cd = [[(y - x) % P for x, y in zip(a, b)] for a, b in zip(lc, rc)]
        exts_p, exts_c = [l, r], [lc, rc]
        for _ in range(self.degree):
            exts_p.append([[(x + d) % P for x, d in zip(a, dd)] for a, dd in zip(exts_p[-1], pd)])
            exts_c.append([[(x + d) % P for x, d in zip(a, dd)] for a, dd in zip(exts_c[-1], cd)])
        out = []
        for polys, eqs in zip(exts_p, exts_c):
            s = 0
            for i in range(len(polys[0])):
                s += self.folded_f([p[i] for p in polys] + [e[i] for e in eqs])
            out.append(s % P)
        return out

    def unipoly(self):
        return unipoly_from_evals(self.unipoly_evals())

    def bind(self, f):  # sumcheck.rs:82-94
        l, r, _, _ = self._split()
        self.polys = [[(x + f * (y - x)) % P for x, y in zip(a, b)] for a, b in zip(l, r)]
        for c in self.copolys:
            c.bind(f)

    def final_evals(self):
        return [p[0] for p in self.polys]
