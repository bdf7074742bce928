#!/usr/bin/env python3
"""Commitment MSM on one B200: KzgProvingKey::commit (src/commitments/kzg.rs:123-126) over the device-generated mock SRS
(kzg.rs:84-97) with uniform 255-bit scalars.  Wall time per call (the call returns the affine result to the host).
usage: python tools/bench_msm.py [log_n ...]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gkr_msm_b200 as g  # noqa: E402
from gkr_msm_b200 import hostmath as H  # noqa: E402
from gkr_msm_b200.fieldutil import to_limb1  # noqa: E402

ctx = g.Context(0)
PRE = int(os.environ.get("MSM_PRE_C", "0"))  # fixed-base window table (gkr_srs_precompute) with this window, 0 = none
for log_n in [int(a) for a in sys.argv[1:]] or [16, 18, 20, 22]:
    n = 1 << log_n
    t0 = time.perf_counter()
    srs = g.Srs.mock_setup(ctx, to_limb1(0x1234567890ABCDEF1234567), H.g1_to_limbs(H.G1_GEN), n)
    ctx.sync()
    t_srs = time.perf_counter() - t0
    if PRE:
        srs.precompute(PRE)
    sc = ctx.synth(77, n)
    srs.msm(sc)  # warm-up
    ts = []
    for _ in range(5):
        ctx.sync()
        t0 = time.perf_counter()
        srs.msm(sc)
        ts.append(time.perf_counter() - t0)
    best = min(ts)
    print(json.dumps({"bench": "msm_g1", "log_n": log_n, "ms": best * 1e3, "points_per_s": n / best, "srs_setup_ms": t_srs * 1e3}), flush=True)
    srs.free()
    sc.free()
