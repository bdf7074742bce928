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


# ---------------------------------------------------------------------------------------------------------------------
# Old-API protocol flow on Shape::full tables (BASELINE config[4]: benches/bintree.rs, gkr_msm_simple):
#   merlin transcript with labels                      src/transcript.rs:78-101
#   to_multieval, make_folded_claim, SumcheckPolyMap::{witness}, SumcheckPolyMapProver::{start, round},
#   SumcheckPolyMapVerifier::{start, round}            src/protocol/sumcheck.rs:160-257, 317-321, 524-672
#   Split::witness, SplitProver::round                 src/protocol/split.rs:37-85
#   Layer, BintreeParams::unroll, BintreeProtocol::witness, BintreeProver::round, BintreeVerifier::round
#                                                      src/protocol/bintree.rs:14-397
#   the layer list and the driver loop                 benches/bintree.rs:20-118, 160-190
from .field import fr_serialize, from_le_bytes_mod_order  # noqa: E402
from .sumcheck import compress_coefficients, decompress_coefficients, eq_eval, evaluate_univar  # noqa: E402
from .transcript import MerlinTranscript  # noqa: E402


class OldTranscript:
    """`impl TranscriptReceiver / TranscriptSender for merlin::Transcript` (src/transcript.rs:78-101)"""

    def __init__(self, label: bytes):
        self.m = MerlinTranscript(label)

    def append_scalars(self, label, scalars):  # one message per scalar, label b"" whatever the caller passes
        for s in scalars:
            self.m.append_message(b"", fr_serialize(s))

    def challenge_scalar(self, label: bytes) -> int:  # 64 bytes -> from_le_bytes_mod_order
        return from_le_bytes_mod_order(self.m.challenge_bytes(label, 64))


def split_full(p):  # FragmentedPoly::split on one Data fragment: even entries left, odd entries right (fragmented.rs:676-741)
    return p[0::2], p[1::2]


def evaluate_full(p, point):  # FragmentedPoly::evaluate (fragmented.rs:748-761): binds the LAST coordinate first
    cur = list(p)
    for f in reversed(point):
        l, r = split_full(cur)
        cur = [(x + f * (y - x)) % P for x, y in zip(l, r)]
    return cur[0]


class MapLayer:  # Layer::Mapping(PolynomialMapping)
    def __init__(self, gate):
        self.gate, self.num_i, self.num_o, self.degree = gate, gate.n_ins, gate.n_outs, gate.deg
        self.is_split = False


class SplitLayer:  # Layer::Split(n)
    def __init__(self, n):
        self.num_i, self.num_o, self.is_split = n, 2 * n, True


def bintree_layers(log_num_points):
    """benches/bintree.rs:86-108: split(2), affine L1/L2/L3, then (split(3), projective L1/L2/L3) x (log_num_points - 2)"""
    from . import gates as G
    layers = [SplitLayer(2), MapLayer(G.AffL1()), MapLayer(G.AffL2()), MapLayer(G.AffL3())]
    for _ in range(log_num_points - 2):
        layers += [SplitLayer(3), MapLayer(G.PrjL1()), MapLayer(G.PrjL2()), MapLayer(G.PrjL3())]
    return layers


def unroll(layers, num_vars):  # BintreeParams::unroll (bintree.rs:78-122)
    out, last_o = [], None
    for layer in layers:
        if last_o is not None:
            assert last_o == layer.num_i, "Amount of inputs differs from amount of outputs"
        out.append((layer, num_vars))
        if layer.is_split:
            assert num_vars > 0, "Can not split 0-variable vector."
            num_vars -= 1
        last_o = layer.num_o
    assert not out[-1][0].is_split, "Technical condition: split can not be last operation."
    return out


def bintree_witness(args, layers, num_vars):
    """BintreeProtocol::witness (bintree.rs:168-185): trace = the input of every layer"""
    assert len(args[0]) == 1 << num_vars
    trace, output = [], [list(a) for a in args]
    for layer, _nv in unroll(layers, num_vars):
        trace.append(output)
        if layer.is_split:  # Split::witness (split.rs:37-48): all left halves, then all right halves
            ls, rs = zip(*[split_full(p) for p in output])
            output = [list(x) for x in ls] + [list(x) for x in rs]
        else:  # SumcheckPolyMap::witness = map_over_poly
            n = len(output[0])
            outs = [[0] * n for _ in range(layer.num_o)]
            for i in range(n):
                o = layer.gate.exec([p[i] for p in output])
                for k in range(layer.num_o):
                    outs[k][i] = o[k] % P
            output = outs
    return trace, output


def make_folded_claim(evs, gamma_pows):  # sumcheck.rs:658-672 (one claimed point)
    return sum(e * g for e, g in zip(evs, gamma_pows)) % P


def bintree_prove(transcript: OldTranscript, point, evs, trace, layers, num_vars, label=b"challenge_nextround"):
    """BintreeProver::{start, round} driven like benches/bintree.rs:177-183.  Returns (final EvalClaim (point, evs),
    proof = per layer None (split) | (compressed round polynomials, final evaluations))."""
    trace = list(trace)
    params = unroll(layers, num_vars)
    claim_point, claim_evs = list(point), list(evs)
    proofs = []
    while params:
        layer, nv = params.pop()
        polys = trace.pop()
        if layer.is_split:  # SplitProver::round (split.rs:64-84)
            r = transcript.challenge_scalar(label)
            h = len(claim_evs) // 2
            claim_evs = [(x + r * (y - x)) % P for x, y in zip(claim_evs[:h], claim_evs[h:])]
            claim_point = claim_point + [r]  # fix_var_top
            proofs.append(None)
            continue
        # SumcheckPolyMapProver (sumcheck.rs:178-257); to_multieval: every output claimed at the one point
        assert len(polys) == layer.num_i and len(claim_point) == nv
        gamma = transcript.challenge_scalar(label)
        gamma_pows = make_gamma_pows_legacy(len(claim_evs), gamma)
        claim_struct = [[(o, v) for o, v in enumerate(claim_evs)]]
        so = FragmentedLincombFull(polys, [claim_point], make_folded_f(claim_struct, gamma_pows, layer.gate.exec, layer.num_i), layer.degree)
        rs, round_polys = [], []
        while True:
            if len(rs) == nv:
                fe = so.final_evals()[:layer.num_i]
                transcript.append_scalars(b"sumcheck_final_evals", fe)
                break
            poly = so.unipoly()
            transcript.append_scalars(b"poly", poly)
            round_polys.append(compress_coefficients(poly))
            r_j = transcript.challenge_scalar(label)
            rs.insert(0, r_j)  # fix_var_bot
            so.bind(r_j)
        proofs.append((round_polys, fe))
        claim_point, claim_evs = rs, fe
    return (claim_point, claim_evs), proofs


def bintree_verify(transcript: OldTranscript, point, evs, proofs, layers, num_vars, label=b"challenge_nextround"):
    """BintreeVerifier::round over SumcheckPolyMapVerifier / SplitVerifier (bintree.rs:300-397, sumcheck.rs:524-652)"""
    params = unroll(layers, num_vars)
    proofs = list(proofs)
    claim_point, claim_evs = list(point), list(evs)
    while params:
        layer, nv = params.pop()
        proof = proofs.pop(0)
        if layer.is_split:
            assert proof is None
            r = transcript.challenge_scalar(label)
            h = len(claim_evs) // 2
            claim_evs = [(x + r * (y - x)) % P for x, y in zip(claim_evs[:h], claim_evs[h:])]
            claim_point = claim_point + [r]
            continue
        round_polys, fe = proof
        assert len(round_polys) == nv and len(fe) == layer.num_i and len(claim_point) == nv
        gamma = transcript.challenge_scalar(label)
        gamma_pows = make_gamma_pows_legacy(len(claim_evs), gamma)
        current = make_folded_claim(claim_evs, gamma_pows)
        folded = make_folded_f([[(o, v) for o, v in enumerate(claim_evs)]], gamma_pows, layer.gate.exec, layer.num_i)
        rs = []
        for k in range(nv):
            poly = decompress_coefficients(round_polys[k], current)
            assert len(poly) == layer.degree + 2, "Verifier failure: polynomial degree incorrect"
            transcript.append_scalars(b"poly", poly)
            r_j = transcript.challenge_scalar(label)
            rs.insert(0, r_j)
            current = evaluate_univar(poly, r_j)
        transcript.append_scalars(b"sumcheck_final_evals", fe)
        assert folded(list(fe) + [eq_eval(claim_point, rs)]) == current, "Verifier failure: final check incorrect"
        claim_point, claim_evs = rs, list(fe)
    return claim_point, claim_evs
