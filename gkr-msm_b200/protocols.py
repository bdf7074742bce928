"""Host-side mirror of the reference's protocol layer for the EC-addition GKR circuits, driving the device
objects through the C ABI.  In the reference this orchestration is Rust and stays on the host (north_star);
with no Rust toolchain in this image this module is the stand-in, written to read like the reference:

  DenseDeg2Sumcheck::prove                 src/cleanup/protocols/sumchecks/dense_eq.rs:192-221
  VecVecDeg2Sumcheck::prove                src/cleanup/protocols/sumchecks/vecvec_eq.rs:418-450
  DenseEqSumcheck::prove                   src/cleanup/protocols/sumcheck.rs:843-872
  SplitAt / GlueSplit / ZeroCheck          src/cleanup/protocols/splits.rs:120-203, zero_check.rs:17-33
  SimpleGKR::prove                         src/cleanup/protocols/gkrs/gkr.rs:45-50
  bintree witness / protocol builders      src/cleanup/protocols/gkrs/bintree_add.rs:124-375
  triangle witness / protocol builders     src/cleanup/protocols/gkrs/triangle_add.rs:76-232
  PippengerEndingWG / PippengerBucketed    src/cleanup/protocols/pippenger_ending.rs:26-157

Claims are (point, evs) with python ints; tables never leave the device.  Advice is ("vv", [VecVec]),
("dense", [Table]) or ("empty",) like SplitVecVecMapGKRAdvice (split_map_gkr.rs:65-71).
"""
from __future__ import annotations

from . import binding as g
from .fieldutil import R_MOD, from_limbs, make_gamma_pows, to_limb1, to_limbs
from .profiling import span

P = R_MOD

# gate descriptors: (public gate id for single-gate objects, stack parts, n_ins, n_outs)
AFF_L1 = dict(gid=g.GATE_AFF_L1, parts=[(g.GATE_AFF_L1, 1)], n_ins=4, n_outs=3)
AFF_L1_BC2 = dict(gid=g.GATE_AFF_L1_BITCHECK2, parts=[(g.GATE_AFF_L1_BITCHECK2, 1)], n_ins=6, n_outs=5)
AFF_L2 = dict(gid=g.GATE_AFF_L2, parts=[(g.GATE_AFF_L2, 1)], n_ins=3, n_outs=3)
AFF_L3 = dict(gid=g.GATE_AFF_L3, parts=[(g.GATE_AFF_L3, 1)], n_ins=3, n_outs=3)
PRJ_L1 = dict(gid=g.GATE_PRJ_L1, parts=[(g.GATE_PRJ_L1, 1)], n_ins=6, n_outs=4)
PRJ_L2 = dict(gid=g.GATE_PRJ_L2, parts=[(g.GATE_PRJ_L2, 1)], n_ins=4, n_outs=4)
PRJ_L3 = dict(gid=g.GATE_PRJ_L3, parts=[(g.GATE_PRJ_L3, 1)], n_ins=4, n_outs=3)


def ID(n):
    return dict(gid=None, parts=[(g.GATE_ID, n)], n_ins=n, n_outs=n)


def tri_l1(layer_idx):  # Stacked(triangle_l1, Repeated(prj_l1, layer_idx))   triangle_add.rs:128-135
    parts = [(g.GATE_TRI_L1, 1)] + ([(g.GATE_PRJ_L1, layer_idx)] if layer_idx else [])
    return dict(gid=None, parts=parts, n_ins=12 + 6 * layer_idx, n_outs=12 + 4 * layer_idx)


def repeated(gate, k):
    return dict(gid=None, parts=[(gate["parts"][0][0], k)], n_ins=gate["n_ins"] * k, n_outs=gate["n_outs"] * k)


# ---------------------------------------------------------------- sumcheck layers -------------------
class DenseDeg2Sumcheck:
    def __init__(self, ctx, gate, num_vars):
        self.ctx, self.gate, self.num_vars = ctx, gate, num_vars

    def prove(self, tr, claims, advice):
        point, evs = claims
        tables = advice[1]
        assert len(tables) == self.gate["n_ins"]
        gamma = from_limbs(tr.challenge(128))[0]
        gp = make_gamma_pows(gamma, self.gate["n_outs"])
        claim = evs[0]
        for i in range(1, len(evs)):
            claim = (claim + gp[i] * evs[i]) % P
        with span(self.ctx, "gkr: deg2 dense object setup"):
            so = self.ctx.deg2_dense_so(self.gate["parts"], tables, to_limbs(gp), to_limb1(claim), to_limbs(point))
        with span(self.ctx, "gkr: deg2 dense rounds"):
            _, out_point, fe = g.sumcheck_prove(tr, so, self.num_vars)
        so.destroy()
        tr.write_scalars(fe)
        return (from_limbs(out_point), from_limbs(fe))


class VecVecDeg2Sumcheck:
    def __init__(self, ctx, gate, num_vars, num_vertical_vars):
        self.ctx, self.gate, self.num_vars, self.nvv = ctx, gate, num_vars, num_vertical_vars

    def prove(self, tr, claims, advice):
        point, evs = claims
        polys = advice[1]
        assert len(polys) == self.gate["n_ins"]
        gamma = from_limbs(tr.challenge(128))[0]
        gp = make_gamma_pows(gamma, self.gate["n_outs"])
        claim = evs[0]
        for i in range(1, len(evs)):
            claim = (claim + gp[i] * evs[i]) % P
        with span(self.ctx, "gkr: deg2 vecvec object setup"):
            so = self.ctx.deg2_vecvec_so(self.gate["gid"], polys, to_limbs(gp), to_limb1(claim), to_limbs(point), self.nvv)
        with span(self.ctx, "gkr: deg2 vecvec rounds (sparse + dense tail)"):
            _, out_point, fe = g.sumcheck_prove(tr, so, self.num_vars)
        so.destroy()
        fe = fe[:-1]  # poly_evs.pop(): the eq evaluation is not sent (vecvec_eq.rs:445)
        tr.write_scalars(fe)
        return (from_limbs(out_point), from_limbs(fe))


class SplitAt:
    def __init__(self, var_idx, bundle_size):
        self.var_idx, self.bundle_size = var_idx, bundle_size

    def prove(self, tr, claims, advice=None):
        r = from_limbs(tr.challenge(128))[0]
        point, evs = list(claims[0]), list(claims[1])
        b = self.bundle_size
        chunks = [evs[i:i + b] for i in range(0, len(evs), b)]
        evs_l = [x for c in chunks[0::2] for x in c]
        evs_r = [x for c in chunks[1::2] for x in c]
        evs_new = [(x + r * (y - x)) % P for x, y in zip(evs_l, evs_r)]
        kind, x = self.var_idx
        point.insert(len(point) - x if kind == "LO" else x, r)
        return (point, evs_new)


class GlueSplit:
    @staticmethod
    def witness(ctx, polys):
        out = ctx.map_vecvec(ID(2)["parts"], polys[0:2], mode=1, bundle_size=2)
        out += ctx.map_vecvec(ID(1)["parts"], polys[2:3], mode=1, bundle_size=1)
        return out

    def prove(self, tr, claims, advice=None):
        r = from_limbs(tr.challenge(128))[0]
        point, evs = list(claims[0]), list(claims[1])
        evs_new = [(evs[0] + r * (evs[2] - evs[0])) % P, (evs[1] + r * (evs[3] - evs[1])) % P, (evs[4] + r * (evs[5] - evs[4])) % P]
        point.append(r)
        return (point, evs_new)


class ZeroCheck:
    def prove(self, tr, claims, advice=None):
        return (list(claims[0]), list(claims[1]) + [0, 0])


def simple_gkr_prove(layers, tr, claims, advices):
    advices = list(advices)
    assert len(advices) == len(layers)
    for layer in reversed(layers):
        claims = layer.prove(tr, claims, advices.pop())
    return claims


# ---------------------------------------------------------------- witness builders ------------------
def advice_map(ctx, advice, gate):
    if advice[0] == "vv":
        return ("vv", ctx.map_vecvec(gate["parts"], advice[1][:gate["n_ins"]], mode=0))
    return ("dense", ctx.map_dense(gate["parts"], advice[1][:gate["n_ins"]]))


def advice_map_split(ctx, advice, gate, layer_idx, row_logsize, bundle_size):
    ins = advice[1][:gate["n_ins"]]
    if advice[0] == "vv":
        if layer_idx + 2 == row_logsize:
            return ("dense", ctx.map_vecvec(gate["parts"], ins, mode=2, bundle_size=bundle_size))
        return ("vv", ctx.map_vecvec(gate["parts"], ins, mode=1, bundle_size=bundle_size))
    return ("dense", ctx.map_dense(gate["parts"], ins, split=("LO", 0), bundle_size=bundle_size))


def bintree_last_step(ctx, advice, layer_idx):
    return advice_map(ctx, advice, AFF_L3 if layer_idx == 0 else PRJ_L3)


def bintree_witness(ctx, advice, row_logsize, num_adds, do_bitcheck):
    advices = []
    for add_idx in range(num_adds):
        for step in ("L1", "L2", "L3"):
            last = add_idx + 1 == num_adds
            if step == "L1":
                nxt = advice_map(ctx, advice, AFF_L1 if add_idx == 0 else PRJ_L1)
            elif step == "L2":
                nxt = advice_map(ctx, advice, AFF_L2 if add_idx == 0 else PRJ_L2)
            elif last:
                nxt = None
            else:
                nxt = advice_map_split(ctx, advice, AFF_L3 if add_idx == 0 else PRJ_L3, add_idx, row_logsize, 3)
            advices.append(advice)
            if add_idx == 0 and step == "L1" and do_bitcheck:
                advices.append(("empty",))
            advice = nxt
        if add_idx + 1 != num_adds:
            advices.append(("empty",))
    return advices


def bintree_protocol(ctx, num_vars, num_adds, row_logsize, do_bitcheck):
    layers = []
    nvv = num_vars - row_logsize
    for i in range(num_adds):
        for step in ("L1", "L2", "L3"):
            nv = num_vars - i - 1
            if i == 0:
                gate = {"L1": AFF_L1_BC2 if do_bitcheck else AFF_L1, "L2": AFF_L2, "L3": AFF_L3}[step]
                layers.append(VecVecDeg2Sumcheck(ctx, gate, nv, nvv))
            else:
                gate = {"L1": PRJ_L1, "L2": PRJ_L2, "L3": PRJ_L3}[step]
                layers.append(VecVecDeg2Sumcheck(ctx, gate, nv, nvv) if i + 1 < row_logsize else DenseDeg2Sumcheck(ctx, gate, nv))
            if i == 0 and step == "L1" and do_bitcheck:
                layers.append(ZeroCheck())
        if i != num_adds - 1:
            layers.append(SplitAt(("LO", 0), 3))
    return layers


def triangle_last_step(ctx, tables, layer_idx):
    return ctx.map_dense(repeated(PRJ_L3, layer_idx + 3)["parts"], tables)


def triangle_witness(ctx, tables, num_vars, split_idx):
    hi = split_idx[1] if split_idx[0] == "HI" else num_vars - split_idx[1] - 1
    num_layers = num_vars - hi
    advices = []
    advice = tables
    for layer_idx in range(num_layers + 1):
        for step in ("L1", "L2", "L3"):
            if step == "L1":
                nxt = ctx.map_dense(tri_l1(layer_idx)["parts"], advice)
            elif step == "L2":
                nxt = ctx.map_dense(repeated(PRJ_L2, layer_idx + 3)["parts"], advice)
            elif num_layers == layer_idx:
                nxt = None
            else:
                nxt = ctx.map_dense(repeated(PRJ_L3, layer_idx + 3)["parts"], advice, split=("HI", hi), bundle_size=3)
            advices.append(("dense", advice))
            advice = nxt
        if layer_idx < num_layers:
            advices.append(("empty",))
    return advices


def triangle_protocol(ctx, num_vars, split_idx):
    hi = split_idx[1] if split_idx[0] == "HI" else num_vars - split_idx[1] - 1
    num_layers = num_vars - hi
    layers = []
    for layer_idx in range(num_layers + 1):
        nv = num_vars - layer_idx
        layers.append(DenseDeg2Sumcheck(ctx, tri_l1(layer_idx), nv))
        layers.append(DenseDeg2Sumcheck(ctx, repeated(PRJ_L2, layer_idx + 3), nv))
        layers.append(DenseDeg2Sumcheck(ctx, repeated(PRJ_L3, layer_idx + 3), nv))
        if layer_idx < num_layers:
            layers.append(SplitAt(("HI", hi), 3))
    return layers


class PippengerEndingWG:
    """pippenger_ending.rs:32-95.  (The reference builds the whole bintree witness twice; once suffices.)"""

    def __init__(self, ctx, multirow_vars, bucket_vars, horizontal_vars, inputs):
        assert len(inputs) == 6
        self.bintree_advices = bintree_witness(ctx, ("vv", inputs), horizontal_vars, horizontal_vars, True)
        last = bintree_last_step(ctx, self.bintree_advices[-1], horizontal_vars - 1)[1]
        split_l1 = ctx.map_dense(ID(3)["parts"], last, split=("HI", multirow_vars), bundle_size=3)
        split_l2 = ctx.map_dense(ID(6)["parts"], split_l1, split=("HI", multirow_vars), bundle_size=3)
        self.triangle_advices = triangle_witness(ctx, split_l2, multirow_vars + bucket_vars - 2, ("HI", multirow_vars))

    def last(self):
        return self.triangle_advices[-1][1]


class PippengerBucketed:
    """pippenger_ending.rs:102-157"""

    def __init__(self, ctx, multirow_vars, bucket_vars, horizontal_vars):
        self.bintree = bintree_protocol(ctx, multirow_vars + bucket_vars + horizontal_vars, horizontal_vars, horizontal_vars, True)
        self.splits = SplitAt(("HI", multirow_vars), 3)
        self.triangle = triangle_protocol(ctx, multirow_vars + bucket_vars - 2, ("HI", multirow_vars))

    def prove(self, tr, claims, wg: PippengerEndingWG):
        claims = simple_gkr_prove(self.triangle, tr, claims, wg.triangle_advices)
        claims = self.splits.prove(tr, claims)
        claims = self.splits.prove(tr, claims)
        return simple_gkr_prove(self.bintree, tr, claims, wg.bintree_advices)
