#!/usr/bin/env python3
"""Per-round latency of small sumchecks (the regime of ~90% of the rounds of a GKR-MSM proof)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gkr_msm_b200 as g  # noqa: E402
from gkr_msm_b200.fieldutil import to_limb1, to_limbs  # noqa: E402

ctx = g.Context(0)
reps = 20


def timeit(make_so, rounds, label):
    ts, cs = [], []
    for r in range(reps + 3):
        if r == 3:
            ctx.host_stats(True)
        ctx.sync()
        t0 = time.perf_counter()
        so = make_so()
        ctx.sync()
        t1 = time.perf_counter()
        tr = g.Transcript(b"x")
        g.sumcheck_prove(tr, so, rounds)
        t2 = time.perf_counter()
        so.destroy()
        if r >= 3:
            cs.append(t1 - t0)
            ts.append(t2 - t1)
    nl, nw, waits, _ = ctx.host_stats(True)
    print(f"{label:40s} launch-call {nl/1e3/max(waits,1):6.1f} us  wait {nw/1e3/max(waits,1):6.1f} us per round ({waits//reps} waits/prove)", flush=True)
    print(f"{label:40s} create {np.median(cs)*1e6:8.1f} us   prove {np.median(ts)*1e6:8.1f} us  = {np.median(ts)*1e6/rounds:6.1f} us/round", flush=True)


for nv in (4, 8, 12, 16):
    tabs = [ctx.synth(j, 1 << nv) for j in range(3)]
    timeit(lambda: ctx.dense_so(g.SO_PLAIN, g.GATE_PROD3, tabs, nv, to_limb1(0)), nv, f"dense prod3 2^{nv}")
for nv in (4, 8, 12):
    tabs = [ctx.synth(j, 1 << nv) for j in range(6)]
    pt = to_limbs(list(range(2, 2 + nv)))
    gp = to_limbs([1, 5, 25, 125])
    timeit(lambda: ctx.deg2_dense_so([(g.GATE_PRJ_L1, 1)], tabs, gp, to_limb1(0), pt), nv, f"deg2 dense prj_l1 2^{nv}")
for rowv, colv in ((4, 4), (8, 6), (10, 8)):
    nrows = 1 << colv
    lens = np.full(nrows, 1 << rowv, dtype=np.uint32)
    polys = [ctx.upload_vecvec_flat(ctx.synth(j, int(lens.sum())).download(), lens, to_limb1(0), to_limb1(0), rowv, colv) for j in range(6)]
    pt = to_limbs(list(range(2, 2 + rowv + colv)))
    gp = to_limbs([1, 5, 25, 125])
    timeit(lambda: ctx.deg2_vecvec_so(g.GATE_PRJ_L1, polys, gp, to_limb1(0), pt, colv), rowv + colv, f"deg2 vecvec prj_l1 rows 2^{colv} x 2^{rowv}")
