"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

Witness maps of the reference restated over python ints:
  Vec<F>::algfn_map / algfn_map_split            src/cleanup/polys/dense.rs:114-185
  vecvec_map / vecvec_map_split                   src/cleanup/polys/vecvec.rs:480-606
  vecvec_map_split_to_dense                       src/cleanup/polys/vecvec.rs:608-654
  AlgFnUtils::{map, map_split_hi}                 src/cleanup/utils/algfn.rs:49-90
  SplitIdx                                        src/cleanup/protocols/splits.rs:13-50
"""
from __future__ import annotations

from .field import P
from .sumcheck import VecVecPolynomial, log_2


def split_lo(var_idx, num_vars):
    """var_idx = ("LO", k) | ("HI", k) -> k in LO indexing (splits.rs:31-36)."""
    kind, k = var_idx
    return k if kind == "LO" else num_vars - k - 1


def interleave_bundles(l, r, bundle_size):
    """l.chunks(b).interleave(r.chunks(b)).flatten()  (dense.rs:137-138)."""
    out = []
    lc = [l[i:i + bundle_size] for i in range(0, len(l), bundle_size)]
    rc = [r[i:i + bundle_size] for i in range(0, len(r), bundle_size)]
    for i in range(max(len(lc), len(rc))):
        if i < len(lc):
            out += lc[i]
        if i < len(rc):
            out += rc[i]
    return out


def dense_map(polys, func):
    n = len(polys[0])
    outs = [[0] * n for _ in range(func.n_outs)]
    for idx in range(n):
        o = func.exec([p[idx] for p in polys])
        for k in range(func.n_outs):
            outs[k][idx] = o[k] % P
    return outs


def dense_map_split(polys, func, var_idx, bundle_size):
    n = len(polys[0])
    num_vars = log_2(n)
    seg = 1 << split_lo(var_idx, num_vars)
    outs = [[[] for _ in range(func.n_outs)] for _ in range(2)]
    for idx in range(n):
        o = func.exec([p[idx] for p in polys])
        side = (idx // seg) % 2
        for k in range(func.n_outs):
            outs[side][k].append(o[k] % P)
    return interleave_bundles(outs[0], outs[1], bundle_size)


def map_split_hi(polys, func):  # algfn.rs:82-89
    half = len(polys[0]) // 2
    return [dense_map([p[:half] for p in polys], func), dense_map([p[half:] for p in polys], func)]


def vecvec_map(polys, func):
    rl, cl = polys[0].row_logsize, polys[0].col_logsize
    rp = func.exec([p.row_pad for p in polys])
    cp = func.exec([p.col_pad for p in polys])
    datas = [[] for _ in range(func.n_outs)]
    for r in range(len(polys[0].data)):
        rows = [[] for _ in range(func.n_outs)]
        for i in range(len(polys[0].data[r])):
            o = func.exec([p.data[r][i] for p in polys])
            for k in range(func.n_outs):
                rows[k].append(o[k] % P)
        for k in range(func.n_outs):
            datas[k].append(rows[k])
    return [VecVecPolynomial(datas[k], rp[k], cp[k], rl, cl) for k in range(func.n_outs)]


def vecvec_map_split(polys, func, var_idx, bundle_size):
    rl, cl = polys[0].row_logsize, polys[0].col_logsize
    num_vars = rl + cl
    rp = func.exec([p.row_pad for p in polys])
    cp = func.exec([p.col_pad for p in polys])
    seg = 1 << split_lo(var_idx, num_vars)
    datas = [[[] for _ in range(func.n_outs)] for _ in range(2)]
    for r in range(len(polys[0].data)):
        rows = [[[] for _ in range(func.n_outs)] for _ in range(2)]
        for i in range(len(polys[0].data[r])):
            o = func.exec([p.data[r][i] for p in polys])
            for k in range(func.n_outs):
                rows[(i // seg) % 2][k].append(o[k] % P)
        if len(rows[0][0]) % 2 == 1:
            for s in range(2):
                for k in range(func.n_outs):
                    rows[s][k].append(rp[k])
        for s in range(2):
            for k in range(func.n_outs):
                datas[s][k].append(rows[s][k])
    l = [VecVecPolynomial(datas[0][k], rp[k], cp[k], rl - 1, cl, unchecked=True) for k in range(func.n_outs)]
    r_ = [VecVecPolynomial(datas[1][k], rp[k], cp[k], rl - 1, cl, unchecked=True) for k in range(func.n_outs)]
    return interleave_bundles(l, r_, bundle_size)


def vecvec_map_split_to_dense(polys, func, var_idx, bundle_size):
    rl, cl = polys[0].row_logsize, polys[0].col_logsize
    assert rl == 1
    num_vars = rl + cl
    rp = func.exec([p.row_pad for p in polys])
    cp = func.exec([p.col_pad for p in polys])
    seg = 1 << split_lo(var_idx, num_vars)
    outs = [[[] for _ in range(func.n_outs)] for _ in range(2)]
    for r in range(len(polys[0].data)):
        for i in range(len(polys[0].data[r])):
            o = func.exec([p.data[r][i] for p in polys])
            for k in range(func.n_outs):
                outs[(i // seg) % 2][k].append(o[k] % P)
        if len(outs[0][0]) < r + 1:
            for s in range(2):
                for k in range(func.n_outs):
                    outs[s][k].append(rp[k])
    n = 1 << cl
    l = [(k, outs[0][k]) for k in range(func.n_outs)]
    r_ = [(k, outs[1][k]) for k in range(func.n_outs)]
    res = []
    for k, data in interleave_bundles(l, r_, bundle_size):
        res.append(data + [cp[k]] * (n - len(data)))
    return res
