"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

The EC-addition GKR circuits of the reference restated over python ints:
  SimpleGKR::{prove,verify}                          src/cleanup/protocols/gkrs/gkr.rs:39-60
  SplitAt, GlueSplit                                 src/cleanup/protocols/splits.rs:120-203
  ZeroCheck                                          src/cleanup/protocols/zero_check.rs:17-33
  bintree witness `build` / `make_step`, protocol    src/cleanup/protocols/gkrs/bintree_add.rs:124-375
  triangle witness / protocol                        src/cleanup/protocols/gkrs/triangle_add.rs:76-232
  PippengerEndingWG, PippengerBucketed               src/cleanup/protocols/pippenger_ending.rs:26-157

Advice = ("vv", [VecVecPolynomial]) | ("dense", [list]) | ("empty",).  Claims = (point, evs).
"""
from __future__ import annotations

from . import gates as G
from . import polys as OP
from .field import P
from .sumcheck import DenseDeg2Sumcheck, VecVecDeg2Sumcheck


# ---------------------------------------------------------------- claim-only layers ------------------
class SplitAt:
    def __init__(self, var_idx, bundle_size):
        self.var_idx, self.bundle_size = var_idx, bundle_size

    def prove(self, transcript, claims, advice=None):
        r = transcript.challenge(128)
        point, evs = list(claims[0]), list(claims[1])
        chunks = [evs[i:i + self.bundle_size] for i in range(0, len(evs), self.bundle_size)]
        evs_l = [x for c in chunks[0::2] for x in c]
        evs_r = [x for c in chunks[1::2] for x in c]
        evs_new = [(x + r * (y - x)) % P for x, y in zip(evs_l, evs_r)]
        kind, x = self.var_idx
        point.insert(len(point) - x if kind == "LO" else x, r)
        return (point, evs_new)

    verify = prove


class GlueSplit:
    @staticmethod
    def witness(polys):
        out = OP.vecvec_map_split(polys[0:2], G.Id(2), ("LO", 0), 2)
        out += OP.vecvec_map_split(polys[2:3], G.Id(1), ("LO", 0), 1)
        return out

    def prove(self, transcript, claims, advice=None):
        r = transcript.challenge(128)
        point, evs = list(claims[0]), list(claims[1])
        evs_new = [(evs[0] + r * (evs[2] - evs[0])) % P, (evs[1] + r * (evs[3] - evs[1])) % P, (evs[4] + r * (evs[5] - evs[4])) % P]
        point.append(r)
        return (point, evs_new)

    verify = prove


class ZeroCheck:
    def prove(self, transcript, claims, advice=None):
        return (list(claims[0]), list(claims[1]) + [0, 0])

    verify = prove


class _SumcheckLayer:
    """adapter: a *Deg2Sumcheck used as a GKRLayer (split_map_gkr.rs:106-148); extra advice columns beyond
    f.n_ins are never read by the gate but ARE carried through the sumcheck object in the reference -- the
    witness builder always hands over exactly the gate's inputs, which is asserted here."""

    def __init__(self, proto):
        self.proto = proto

    def prove(self, transcript, claims, advice):
        assert len(advice[1]) == self.proto.f.n_ins
        return self.proto.prove(transcript, claims, advice[1])

    def verify(self, transcript, claims):
        return self.proto.verify(transcript, claims)


def simple_gkr_prove(layers, transcript, claims, advices):
    """advices is consumed from the END (the WG iterators pop), layers are walked in reverse (gkr.rs:45-50)."""
    advices = list(advices)
    assert len(advices) == len(layers)
    for layer in reversed(layers):
        claims = layer.prove(transcript, claims, advices.pop())
    return claims


def simple_gkr_verify(layers, transcript, claims):
    for layer in reversed(layers):
        claims = layer.verify(transcript, claims)
    return claims


# ---------------------------------------------------------------- bintree --------------------------
def advice_map(advice, f):
    if advice[0] == "vv":
        return ("vv", OP.vecvec_map(advice[1][:f.n_ins], f))
    return ("dense", OP.dense_map(advice[1][:f.n_ins], f))


def advice_map_split(advice, f, layer_idx, row_logsize, idx, bundle_size):
    if advice[0] == "vv":
        if layer_idx + 2 == row_logsize:
            return ("dense", OP.vecvec_map_split_to_dense(advice[1][:f.n_ins], f, idx, bundle_size))
        return ("vv", OP.vecvec_map_split(advice[1][:f.n_ins], f, idx, bundle_size))
    return ("dense", OP.dense_map_split(advice[1][:f.n_ins], f, idx, bundle_size))


def bintree_last_step(advice, layer_idx):
    return advice_map(advice, G.AffL3() if layer_idx == 0 else G.PrjL3())


def bintree_witness(advice, row_logsize, num_adds, do_bitcheck):
    """bintree_add.rs:137-171"""
    assert num_adds > 0
    advices = []
    for add_idx in range(num_adds):
        for step in ("L1", "L2", "L3"):
            last = add_idx + 1 == num_adds
            if step == "L1":
                nxt = advice_map(advice, G.AffL1() if add_idx == 0 else G.PrjL1())
            elif step == "L2":
                nxt = advice_map(advice, G.AffL2() if add_idx == 0 else G.PrjL2())
            elif last:
                nxt = None
            else:
                nxt = advice_map_split(advice, G.AffL3() if add_idx == 0 else G.PrjL3(), add_idx, row_logsize, ("LO", 0), 3)
            advices.append(advice)
            if add_idx == 0 and step == "L1" and do_bitcheck:
                advices.append(("empty",))
            advice = nxt
        if add_idx + 1 != num_adds:
            advices.append(("empty",))
    return advices


def bintree_protocol(num_vars, num_adds, row_logsize, do_bitcheck):
    """bintree_add.rs:247-375"""
    layers = []
    nvv = num_vars - row_logsize
    for i in range(num_adds):
        for step in ("L1", "L2", "L3"):
            nv = num_vars - i - 1
            if i == 0:
                gate = {"L1": G.AffL1BitCheck2() if do_bitcheck else G.AffL1(), "L2": G.AffL2(), "L3": G.AffL3()}[step]
                layers.append(_SumcheckLayer(VecVecDeg2Sumcheck(gate, nv, nvv)))
            else:
                gate = {"L1": G.PrjL1(), "L2": G.PrjL2(), "L3": G.PrjL3()}[step]
                if i + 1 < row_logsize:
                    layers.append(_SumcheckLayer(VecVecDeg2Sumcheck(gate, nv, nvv)))
                else:
                    layers.append(_SumcheckLayer(DenseDeg2Sumcheck(gate, nv)))
            if i == 0 and step == "L1" and do_bitcheck:
                layers.append(ZeroCheck())
        if i != num_adds - 1:
            layers.append(SplitAt(("LO", 0), 3))
    return layers


# ---------------------------------------------------------------- triangle -------------------------
def _tri_l1(layer_idx):
    return G.TriL1() if layer_idx == 0 else G.Stacked(G.TriL1(), G.Repeated(G.PrjL1(), layer_idx))


def triangle_last_step(advice, layer_idx):
    return OP.dense_map(advice, G.Repeated(G.PrjL3(), layer_idx + 3))


def triangle_witness(advice, num_vars, split_idx):
    """triangle_add.rs:101-158"""
    hi = split_idx[1] if split_idx[0] == "HI" else num_vars - split_idx[1] - 1
    split_hi = ("HI", hi)
    num_layers = num_vars - hi
    advices = []
    for layer_idx in range(num_layers + 1):
        for step in ("L1", "L2", "L3"):
            if step == "L1":
                nxt = OP.dense_map(advice, _tri_l1(layer_idx))
            elif step == "L2":
                nxt = OP.dense_map(advice, G.Repeated(G.PrjL2(), layer_idx + 3))
            elif num_layers == layer_idx:
                nxt = None
            else:
                nxt = OP.dense_map_split(advice, G.Repeated(G.PrjL3(), layer_idx + 3), split_hi, 3)
            advices.append(("dense", advice))
            advice = nxt
        if layer_idx < num_layers:
            advices.append(("empty",))
    return advices


def triangle_protocol(num_vars, split_idx):
    """triangle_add.rs:173-232"""
    hi = split_idx[1] if split_idx[0] == "HI" else num_vars - split_idx[1] - 1
    num_layers = num_vars - hi
    layers = []
    for layer_idx in range(num_layers + 1):
        nv = num_vars - layer_idx
        layers.append(_SumcheckLayer(DenseDeg2Sumcheck(_tri_l1(layer_idx), nv)))
        layers.append(_SumcheckLayer(DenseDeg2Sumcheck(G.Repeated(G.PrjL2(), layer_idx + 3), nv)))
        layers.append(_SumcheckLayer(DenseDeg2Sumcheck(G.Repeated(G.PrjL3(), layer_idx + 3), nv)))
        if layer_idx < num_layers:
            layers.append(SplitAt(("HI", hi), 3))
    return layers


# ---------------------------------------------------------------- pippenger ending -------------------
class PippengerEndingWG:
    """pippenger_ending.rs:32-95 (the reference builds the bintree witness twice; once is enough here)."""

    def __init__(self, multirow_vars, bucket_vars, horizontal_vars, inputs):
        assert len(inputs) == 6
        self.bintree_advices = bintree_witness(("vv", inputs), horizontal_vars, horizontal_vars, True)
        last = bintree_last_step(self.bintree_advices[-1], horizontal_vars - 1)[1]
        split_l1 = OP.dense_map_split(last, G.Id(3), ("HI", multirow_vars), 3)
        split_l2 = OP.dense_map_split(split_l1, G.Repeated(G.Id(3), 2), ("HI", multirow_vars), 3)
        self.triangle_advices = triangle_witness(split_l2, multirow_vars + bucket_vars - 2, ("HI", multirow_vars))

    def last(self):
        return self.triangle_advices[-1][1]


class PippengerBucketed:
    """pippenger_ending.rs:102-157"""

    def __init__(self, multirow_vars, bucket_vars, horizontal_vars):
        self.bintree = bintree_protocol(multirow_vars + bucket_vars + horizontal_vars, horizontal_vars, horizontal_vars, True)
        self.splits = SplitAt(("HI", multirow_vars), 3)
        self.triangle = triangle_protocol(multirow_vars + bucket_vars - 2, ("HI", multirow_vars))

    def prove(self, transcript, claims, wg: PippengerEndingWG):
        claims = simple_gkr_prove(self.triangle, transcript, claims, wg.triangle_advices)
        claims = self.splits.prove(transcript, claims)
        claims = self.splits.prove(transcript, claims)
        return simple_gkr_prove(self.bintree, transcript, claims, wg.bintree_advices)

    def verify(self, transcript, claims):
        claims = simple_gkr_verify(self.triangle, transcript, claims)
        claims = self.splits.verify(transcript, claims)
        claims = self.splits.verify(transcript, claims)
        return simple_gkr_verify(self.bintree, transcript, claims)
