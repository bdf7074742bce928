#!/usr/bin/env python3
"""Secondary benchmark (BASELINE config[4]/[0]-shaped): witness generation + GKR proof of the EC-addition binary
tree over bucketed points (VecVecBintreeAdd, src/cleanup/protocols/gkrs/bintree_add.rs) on one B200.
Synthetic: M = y_size * 2^x point-digit incidences spread over y_size * 2^d bucket rows with uniform digits, random
field elements as coordinates (the arithmetic cost does not depend on the points being on the curve).
usage: python tools/bench_bintree.py --x 16 --d 8 --ysize 16 [--reps 3]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import gkr_msm_b200 as g  # noqa: E402
from gkr_msm_b200 import protocols as DP  # noqa: E402
from gkr_msm_b200.fieldutil import to_limb1  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--x", type=int, default=16)
    ap.add_argument("--d", type=int, default=8)
    ap.add_argument("--ysize", type=int, default=16)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    ylog = max(1, (a.ysize - 1).bit_length())
    ctx = g.Context(0)
    rng = np.random.default_rng(1)
    nrows = a.ysize << a.d
    # uniform digits: every y-row scatters its 2^x points over 2^d buckets
    lens = np.concatenate([np.bincount(rng.integers(0, 1 << a.d, size=1 << a.x), minlength=1 << a.d) for _ in range(a.ysize)]).astype(np.uint32)
    total = int(lens.sum())
    col_log = ylog + a.d
    polys = []
    for j, (rp, cp) in enumerate([(0, 0), (1, 1), (0, 0)]):
        flat = ctx.synth(100 + j, total).download() if j < 2 else np.tile(g.MONT_ONE, (total, 1))
        polys.append(ctx.upload_vecvec_flat(flat, lens, to_limb1(rp), to_limb1(cp), a.x, col_log))
    num_vars = a.x + col_log
    res = []
    for rep in range(a.reps + 1):
        ctx.sync()
        l0 = ctx.launches
        t0 = time.perf_counter()
        inputs = DP.GlueSplit.witness(ctx, polys)
        adv = DP.bintree_witness(ctx, ("vv", inputs), a.x, a.x, True)
        last = DP.bintree_last_step(ctx, adv[-1], a.x - 1)[1]
        ctx.sync()
        t1 = time.perf_counter()
        layers = DP.bintree_protocol(ctx, num_vars, a.x, a.x, True)
        tr = g.Transcript(b"fgstglsp")
        point = [int(v) for v in rng.integers(1, 1 << 62, size=col_log)]
        claims = (point, [1, 2, 3])  # proving time does not depend on the claim being true
        out = DP.simple_gkr_prove(layers, tr, claims, adv)
        ctx.sync()
        t2 = time.perf_counter()
        if rep:
            res.append((t1 - t0, t2 - t1, ctx.launches - l0, len(tr.proof())))
        del adv, inputs, last
    w = min(r[0] for r in res) * 1e3
    p = min(r[1] for r in res) * 1e3
    print(json.dumps({"workload": f"bintree GKR x={a.x} d={a.d} y_size={a.ysize}", "incidences": total, "bucket_rows": nrows,
                      "witness_ms": w, "prove_ms": p, "total_ms": w + p, "launches": res[-1][2], "proof_bytes": res[-1][3],
                      "leaves_per_s": total / ((w + p) * 1e-3)}))


if __name__ == "__main__":
    main()
