#!/usr/bin/env python3
"""Headline secondary benchmark: `examples/pippenger` end to end on one B200 -- benchutils::run_pippenger
(src/cleanup/protocols/pippenger.rs:499-559: witness generation + phase-1 commitments + the whole proof), the span the
reference's criterion bench times (benches/pippenger.rs:40-45).  Inputs are synthetic: on-curve Bandersnatch points in
arithmetic progression, uniform `nbits`-bit scalars, the reference's own mock SRS (kzg.rs:84-97) generated on the device.

  python tools/bench_pippenger.py --x-logsize 16 --d-logsize 8 --nbits 128 --clm 0 [--reps 3]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def srs_tau(seed: int) -> int:
    """tau of the mock SRS: its own stream, so the worker ranks of a multi-GPU run derive it without generating the inputs"""
    from gkr_msm_b200.fieldutil import R_MOD
    return int.from_bytes(np.random.default_rng(seed ^ 0x7A5).bytes(32), "little") % R_MOD


def _points_chunk(job):
    """(k_start, step, count) -> Montgomery limbs of the x and y coordinates of (k_start + i step) G, i < count"""
    from gkr_msm_b200 import hostmath as H
    from gkr_msm_b200.fieldutil import to_limbs
    k_start, step, count = job
    pts = H.te_points_arithmetic_progression(k_start, step, count)
    return to_limbs([p[0] for p in pts]), to_limbs([p[1] for p in pts])


def gen_points(k0: int, step: int, n: int, procs: int):
    """the synthetic points (k0 + i step) G as a (2, n, 4) limb array; `procs` > 1 splits the progression over worker processes
    (pure-python big-integer arithmetic: 9 us per point and core)"""
    if procs <= 1 or n < (1 << 14):
        x, y = _points_chunk((k0, step, n))
        return np.stack([x, y])
    import multiprocessing as mp
    chunk = max(1 << 12, n // (procs * 4))
    jobs = [(k0 + i * step, step, min(chunk, n - i)) for i in range(0, n, chunk)]
    with mp.get_context("fork").Pool(procs) as pool:
        parts = pool.map(_points_chunk, jobs)
    return np.stack([np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])])


def team_worker(args, ctx=None):
    """rank > 0 of `--gpus N`: same SRS on cuda:rank, then serve slices of the leader's commitment MSMs (csrc/msm_team.cu)"""
    import gkr_msm_b200 as g
    from gkr_msm_b200 import hostmath as H
    from gkr_msm_b200 import pippenger as DPP

    own = ctx is None
    ctx = ctx or g.Context(args.team_worker)
    nv = args.x_logsize + args.clm
    kzg = DPP.KzgKey.mock_setup(ctx, int(args.team_tau, 16), H.G1_GEN, 2 * (1 << nv) - 1)
    team = g.MsmTeam(ctx, args.team_name, args.team_worker, args.gpus, 2 * (1 << nv), open_timeout_s=300.0)
    team.serve(kzg.srs, idle_timeout_s=300.0)
    team.close()
    if own:
        ctx.close()


def run(args, ctx=None, team_name=None):
    """team_name: under torchrun the worker ranks already exist (bench.py): lead the team of that name instead of spawning"""
    import gkr_msm_b200 as g
    from gkr_msm_b200 import hostmath as H
    from gkr_msm_b200 import pippenger as DPP
    from gkr_msm_b200 import profiling as PR
    from gkr_msm_b200.fieldutil import R_MOD, to_limbs

    rng = np.random.default_rng(args.seed)
    xl, dl, nbits, clm = args.x_logsize, args.d_logsize, args.nbits, args.clm
    cfg = DPP.pippenger_config(dl, xl, nbits, clm)
    n = 1 << xl
    t0 = time.perf_counter()
    if getattr(args, "points_file", ""):  # generated beforehand by `--gen-only` (bench.py: a process that holds a CUDA context does not fork)
        points_xy = np.load(args.points_file)
        assert points_xy.shape == (2, n, 4)
    else:
        points_xy = gen_points(0x1234567 + args.seed, 0x9E3779B97F4A7C15, n, getattr(args, "gen_procs", 1))
    raw = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    nbytes = nbits // 8  # from_le_bytes_mod_order(&bytes[..num_bits / 8]), pippenger.rs:465-467
    coefs = np.zeros((n, 4), np.uint64)
    b = raw.view(np.uint8).reshape(n, 32).copy()
    b[:, nbytes:] = 0
    coefs[:] = b.view(np.uint64).reshape(n, 4)
    r = [int.from_bytes(rng.bytes(32), "little") % R_MOD for _ in range(cfg["y_logsize"])]
    t_inputs = time.perf_counter() - t0

    ctx = ctx or g.Context(0)
    if getattr(args, "peer_pool", 0) > 1:  # the other GPUs of the box lend their HBM (gkr_ctx_peer_pool): instances beyond 180 GB
        ctx.peer_pool(args.peer_pool)
    t0 = time.perf_counter()
    nv = xl + clm
    tau = srs_tau(args.seed)
    pre_c = getattr(args, "precompute_c", -1)
    if pre_c < 0:
        pre_c = 20 if nv >= 20 else 0  # fixed-base window table of the SRS: measured to pay from 2^21-point commitments (16.4 -> 15.2 ms)
    kzg = DPP.KzgKey.mock_setup(ctx, tau, H.G1_GEN, 2 * (1 << nv) - 1, precompute_c=pre_c)
    key = DPP.KnucklesKey(ctx, kzg, nv, 2)
    ctx.sync()
    t_setup = time.perf_counter() - t0
    team, workers = None, []
    if getattr(args, "gpus", 1) > 1:  # commitment MSMs split by point range over N GPUs: this process leads, N - 1 workers serve
        import subprocess
        name = team_name or f"/gkr_msm_team_{os.getpid()}"
        team = g.MsmTeam(ctx, name, 0, args.gpus, 2 * (1 << nv))
        if team_name is None:
            for rk in range(1, args.gpus):
                workers.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--x-logsize", str(xl), "--clm", str(clm), "--gpus", str(args.gpus),
                                                 "--team-worker", str(rk), "--team-name", name, "--team-tau", "%x" % tau]))
        team.wait_ready(timeout_s=300.0)

    # peak device memory while proving (driver view, includes the stream-ordered pool): sampled every 5 ms
    mem_peak = [0]
    stop_sampling = [False]
    sampler = None
    if getattr(args, "mem", False):
        import threading
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(0)

        def sample():
            while not stop_sampling[0]:
                mem_peak[0] = max(mem_peak[0], pynvml.nvmlDeviceGetMemoryInfo(handle).used)
                time.sleep(0.005)

        sampler = threading.Thread(target=sample, daemon=True)
        sampler.start()
    times, proof_len, launches = [], 0, 0
    for rep in range(args.reps):
        tr = g.Transcript(b"fgstglsp")
        l0 = ctx.launches
        ctx.sync()
        ctx.host_stats(True)
        t0 = time.perf_counter()
        if args.python_host:
            dense_output, claims, pair = DPP.run_pippenger(ctx, tr, points_xy, coefs, cfg, r, key)
        else:
            native_out = g.run_pippenger_native(ctx, tr, kzg.srs, kzg.g0, key.dev, points_xy, coefs, dl, xl, nbits, clm, to_limbs(r))
            if getattr(args, "dump", "") and rep == args.reps - 1:  # everything tests/verify_dumped_proof.py needs besides the seed
                np.savez(args.dump, proof=np.frombuffer(tr.proof(), dtype=np.uint8), dense=np.ascontiguousarray(native_out[0]),
                         evs=np.ascontiguousarray(native_out[1]), pair=np.ascontiguousarray(native_out[2]), r=to_limbs(r),
                         meta=np.array([xl, dl, nbits, clm, args.seed], dtype=np.int64), tau=to_limbs([tau]))
        ctx.sync()
        times.append(time.perf_counter() - t0)
        launches = ctx.launches - l0
        proof_len = len(tr.proof())
        hs = ctx.host_stats(True)
    if args.profile:
        PR.PROFILE = {}
        ctx.host_stats(True)
        tr = g.Transcript(b"fgstglsp")
        t0 = time.perf_counter()
        DPP.run_pippenger(ctx, tr, points_xy, coefs, cfg, r, key)
        ctx.sync()
        tot = time.perf_counter() - t0
        for k, v in sorted(PR.PROFILE.items(), key=lambda kv: -kv[1]):
            print(f"  {v * 1e3:9.2f} ms  {k}", file=sys.stderr)
        print(f"  {tot * 1e3:9.2f} ms  total (with span syncs)", file=sys.stderr)
        nl, nw, waits, _ = ctx.host_stats(True)
        print(f"  round kernels: {waits} result waits, {nw / 1e6:.2f} ms spinning on results ({nw / 1e3 / max(waits, 1):.1f} us each), "
              f"{nl / 1e6:.2f} ms inside launch calls", file=sys.stderr)
        PR.PROFILE = None
    if team is not None:
        team.quit()
        for w in workers:
            w.wait(timeout=120)
        team.close()
    if sampler is not None:
        stop_sampling[0] = True
        sampler.join()
    best = min(times)
    return ({
        "peak_device_memory_gib": mem_peak[0] / 2**30 if mem_peak[0] else None,
        "bench": "run_pippenger (witness + commit + prove)", "host": "python" if args.python_host else "c++ (gkr_run_pippenger)", "x_logsize": xl, "d_logsize": dl, "nbits": nbits, "clm": clm,
        "y_size": cfg["y_size"], "incidences": cfg["y_size"] << xl, "prove_ms_best": best * 1e3, "prove_ms_all": [t * 1e3 for t in times],
        "proof_bytes": proof_len, "gpu_launches": launches, "input_generation_s": t_inputs, "srs_setup_s": t_setup,
        "srs_points": 2 * (1 << nv) - 1, "srs_fixed_base_window": pre_c, "n_gpus": getattr(args, "gpus", 1),
        "multi_gpu": "commitment MSMs of >= 2^18 points split by point range over the GPUs (csrc/msm_team.cu); everything else on GPU 0" if getattr(args, "gpus", 1) > 1 else None,
        "round_waits": hs[2], "round_wait_ms": hs[1] / 1e6, "round_launch_call_ms": hs[0] / 1e6,
        "peer_pool_gpus": getattr(args, "peer_pool", 0), "peer_pool_peak_gib": (ctx.peer_pool(0)[1] / 2**30) if getattr(args, "peer_pool", 0) > 1 else None})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--x-logsize", type=int, default=16)
    ap.add_argument("--d-logsize", type=int, default=8)
    ap.add_argument("--nbits", type=int, default=128)
    ap.add_argument("--clm", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--profile", action="store_true", help="print a per-phase breakdown of the last repetition (python host)")
    ap.add_argument("--python-host", action="store_true", help="time the python orchestration instead of gkr_run_pippenger (C++)")
    ap.add_argument("--precompute-c", type=int, default=-1, help="window of the fixed-base SRS table (0: none; default: 20 from x + clm >= 19)")
    ap.add_argument("--peer-pool", type=int, default=0, help="N > 1: GPUs 1..N-1 lend their HBM to the prover on GPU 0 (gkr_ctx_peer_pool)")
    ap.add_argument("--gen-procs", type=int, default=1, help="worker processes for the synthetic points (host-side python)")
    ap.add_argument("--dump", default="", help="write proof, outputs and pairing pair of the last repetition to this .npz (tests/verify_dumped_proof.py)")
    ap.add_argument("--gen-only", default="", help="only generate the synthetic points into this .npy file and exit (no GPU work)")
    ap.add_argument("--points-file", default="", help="points generated by --gen-only")
    ap.add_argument("--mem", action="store_true", help="sample the device memory in use while proving (pynvml) and report the peak")
    ap.add_argument("--gpus", type=int, default=1, help="N > 1: spawn N - 1 worker processes (cuda:1..N-1) that share the large commitment MSMs")
    ap.add_argument("--team-worker", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--team-name", default="", help=argparse.SUPPRESS)
    ap.add_argument("--team-tau", default="", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.team_worker > 0:
        team_worker(args)
        return
    if args.gen_only:
        np.save(args.gen_only, gen_points(0x1234567 + args.seed, 0x9E3779B97F4A7C15, 1 << args.x_logsize, args.gen_procs))
        return
    print(json.dumps(run(args)), flush=True)


if __name__ == "__main__":
    main()
