"""Old API of the reference on the device (BASELINE config[4]): the round-by-round prover of
`benches/bintree.rs` -- BintreeProtocol::witness + BintreeProver::round (src/protocol/bintree.rs:168-288) over
SumcheckPolyMapProver::round (src/protocol/sumcheck.rs:178-257) and SplitProver::round (src/protocol/split.rs:64-84) -- for
`Shape::full` tables (one Data fragment, the shape the benches build, benches/bintree.rs:37-47).

Host orchestration only (python, like protocols.py): every table-sized step runs on the device through the C ABI --
  FragmentedPoly::map_over_poly (fragmented.rs:811-817)            -> gkr_map_dense
  FragmentedPoly::split, even / odd (fragmented.rs:676-741)        -> gkr_map_dense(ID gate, SplitIdx::LO(0), bundle = all)
  FragmentedLincomb::{unipoly, bind, final_evals} with the materialised EqPoly (sumcheck.rs:66-156, copoly.rs:600-635)
                                                                   -> gkr_so_create_dense(GKR_SO_EQ_GAMMA) + gkr_eq_table
  merlin transcript with labels, 64-byte challenges (src/transcript.rs:78-101) -> gkr_transcript_*_old
The old prover evaluates the round polynomial at 0..degree+1 and appends ALL its coefficients to the transcript; the device
object returns exactly those evaluations (eval(0) = claim - eval(1)), so the transcript and the proof are identical.
"""
from __future__ import annotations

import numpy as np

from . import binding as g
from . import hostmath as H
from .fieldutil import R_MOD, from_limbs, to_limbs

LABEL = b"challenge_nextround"


def bintree_layers(log_num_points: int):
    """benches/bintree.rs:86-108 as (kind, gate id | n) pairs"""
    layers = [("split", 2), ("map", g.GATE_AFF_L1), ("map", g.GATE_AFF_L2), ("map", g.GATE_AFF_L3)]
    for _ in range(log_num_points - 2):
        layers += [("split", 3), ("map", g.GATE_PRJ_L1), ("map", g.GATE_PRJ_L2), ("map", g.GATE_PRJ_L3)]
    return layers


GATE_IO = {g.GATE_AFF_L1: (4, 3), g.GATE_AFF_L2: (3, 3), g.GATE_AFF_L3: (3, 3), g.GATE_PRJ_L1: (6, 4), g.GATE_PRJ_L2: (4, 4), g.GATE_PRJ_L3: (4, 3)}


def unroll(layers, num_vars):
    """BintreeParams::unroll (bintree.rs:78-122)"""
    out = []
    for kind, v in layers:
        out.append((kind, v, num_vars))
        if kind == "split":
            if num_vars == 0:
                raise ValueError("Can not split 0-variable vector.")
            num_vars -= 1
    if out[-1][0] == "split":
        raise ValueError("Technical condition: split can not be last operation.")
    return out


def bintree_witness(ctx: g.Context, tables, layers, num_vars):
    """BintreeProtocol::witness: (trace = input tables of every layer, output tables), all resident"""
    trace, output = [], list(tables)
    for kind, v, _nv in unroll(layers, num_vars):
        trace.append(output)
        if kind == "split":  # all left (even) halves, then all right (odd) halves (split.rs:45-46)
            output = ctx.map_dense([(g.GATE_ID, len(output))], output, split=("LO", 0), bundle_size=len(output))
        else:
            output = ctx.map_dense([(v, 1)], output[:GATE_IO[v][0]])
    return trace, output


def interpolate4(evals_limbs):
    """UniPoly::from_evals on nodes 0..3 -> coefficients low -> high (host, O(1) per round)"""
    return H.from_evals(from_limbs(evals_limbs))


def bintree_prove(ctx: g.Context, transcript: g.Transcript, point_limbs, evs_limbs, trace, layers, num_vars):
    """BintreeProver driven like benches/bintree.rs:177-183: one labelled 64-byte challenge before every round call.
    Returns ((point, evs) as python ints, proofs = per layer None | (compressed round polys, final evals))."""
    params = unroll(layers, num_vars)
    trace = list(trace)
    claim_point = from_limbs(np.asarray(point_limbs).reshape(-1, 4))
    claim_evs = from_limbs(np.asarray(evs_limbs).reshape(-1, 4))
    proofs = []
    while params:
        kind, v, nv = params.pop()
        tabs = trace.pop()
        if kind == "split":  # SplitProver::round
            r = from_limbs(transcript.challenge_scalar_old(LABEL).reshape(1, 4))[0]
            h = len(claim_evs) // 2
            claim_evs = [(x + r * (y - x)) % R_MOD for x, y in zip(claim_evs[:h], claim_evs[h:])]
            claim_point = claim_point + [r]  # fix_var_top
            proofs.append(None)
            continue
        num_i, num_o = GATE_IO[v]
        assert len(tabs) == num_i and len(claim_point) == nv
        gamma = from_limbs(transcript.challenge_scalar_old(LABEL).reshape(1, 4))[0]
        gamma_pows = [1, gamma]
        for i in range(2, max(len(claim_evs), 2)):
            gamma_pows.append(gamma_pows[-1] * gamma % R_MOD)  # make_gamma_pows_legacy (utils.rs:104-113)
        claim = sum(e * gp for e, gp in zip(claim_evs, gamma_pows)) % R_MOD  # make_folded_claim
        eq = ctx.eq_table(to_limbs(claim_point)) if nv > 0 else ctx.upload(to_limbs([1]))
        so = ctx.dense_so(g.SO_EQ_GAMMA, v, list(tabs) + [eq], nv, to_limbs([claim])[0], consts=to_limbs(gamma_pows))
        rs, round_polys = [], []
        for _ in range(nv):
            coeffs = interpolate4(so.unipoly())
            transcript.append_scalars_old(to_limbs(coeffs))
            round_polys.append([coeffs[0]] + coeffs[2:])  # UniPoly::compress: the linear term is dropped
            r_j = transcript.challenge_scalar_old(LABEL)
            rs.insert(0, from_limbs(r_j.reshape(1, 4))[0])  # fix_var_bot
            so.bind(r_j)
        fe_l = so.final_evals()[:num_i]
        transcript.append_scalars_old(fe_l)
        fe = from_limbs(fe_l)
        so.destroy()
        proofs.append((round_polys, fe))
        claim_point, claim_evs = rs, fe
    return (claim_point, claim_evs), proofs
