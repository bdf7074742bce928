#!/usr/bin/env python3
"""bench.py -- the reference's headline metric on B200.

Metric (BASELINE.json): "sumcheck evals/s vs HBM roofline" on config[1], the standalone dense sumcheck
prover (DenseSumcheckObjectSO + GenericSumcheckProtocol::prove, src/cleanup/protocols/sumcheck.rs:95-128,
241-347) with Prod3Fn (pushforward.rs:38-50) over P = 3 tables of 2^24 BLS12-381 Fr elements.
  unit: table-element-rounds per second = sum over rounds and tables of the table length at that round
         divided by the time of the whole proof (all rounds, Fiat-Shamir on the host).
A "step" is one complete sumcheck proof over one batch of synthetic tables.

  value : tables already resident in HBM when the timed region starts
  e2e   : same proof through the C ABI with HOST (pinned) buffers: H2D of the tables and D2H of the
          round messages / final evaluations inside the timed region

Launch:  python bench.py [--gpus N --steps K --warmup W] [--impl reference]
         (N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "sumcheck_evals_per_s"
UNIT = "table-element-rounds/s"
P_TABLES = 3
SEED = 20240


def elem_rounds(log_n: int, p: int = P_TABLES) -> int:
    return p * ((1 << (log_n + 1)) - 2)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 6:
                self.rows.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for nm, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_run(log_n: int, reps: int, min_seconds: float = 0.0):
    """The oracle's C port of DenseSumcheckObjectSO (OpenMP over all host threads) timed on `reps` full
    sumchecks over 2^log_n x 3 tables.  Returns (evals_per_s, seconds_per_step, threads)."""
    from oracle import coracle

    n = 1 << log_n
    tabs = [coracle.synth_table(SEED + j, n) for j in range(P_TABLES)]
    claim = coracle.gate_sum(0, 10, tabs)
    chals = coracle.synth_table(SEED + 100, log_n)
    chals[:, 2:] = 0  # 128-bit challenges like transcript.challenge(128)
    times = []
    t_all = time.perf_counter()
    while len(times) < reps or (time.perf_counter() - t_all) < min_seconds:
        t0 = time.perf_counter()
        coracle.dense_sumcheck(0, 10, tabs, log_n, claim, chals)
        times.append(time.perf_counter() - t0)
        if len(times) >= 64:
            break
    sec = statistics.median(times)
    return elem_rounds(log_n) / sec, sec, coracle.num_threads(), len(times)


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm must use all the host threads it can, so the
    variable is reset BEFORE the OpenMP runtime of the oracle library is loaded."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(n)
    return n


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is a Rust crate
    (nightly + un-vendored git deps) that cannot be built in this image, so this arm times the oracle's C
    port of the same algorithm (oracle/c/gkr_oracle.c) on all host threads -- rank 0 only."""
    if rank != 0:
        return
    use_all_host_threads()
    log_n = args.ref_log_n
    for _ in range(args.warmup):
        cpu_port_run(log_n, 1)
    val, sec, threads, reps = cpu_port_run(log_n, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u256 (BLS12-381 Fr, 4x64-bit Montgomery limbs)", "data": "synthetic",
        "config": {"workload": f"dense Prod3 sumcheck, P=3 tables x 2^{args.log_n} Fr (config[1])",
                   "sample": f"each step = one full sumcheck over 2^{log_n} x 3 tables (bounded sample of the 2^{args.log_n} workload)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"full Prod3 sumcheck, 2^{log_n} x 3 tables, median of {reps} runs, OpenMP {threads} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch

    import gkr_msm_b200 as g

    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    ctx = g.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    log_n = args.log_n
    n = 1 << log_n

    from gkr_msm_b200.sharded import ShardedProd3Sumcheck  # host-side driver (1 or N ranks)

    job = ShardedProd3Sumcheck(ctx, log_n_local=log_n, rank=rank, world=world, dist=dist, seed=SEED)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm -------------------------------------------------------------------
    for _ in range(args.warmup):
        job.prove_resident()
    ctx.timing_read()
    ctx.timing_enable(True)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = ctx.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        job.prove_resident()
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launches - l0
    ms_total = ev0.elapsed_time(ev1)
    launches_timed = ctx.timing_read()
    ctx.timing_enable(False)
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    total_units = elem_rounds(log_n + (world.bit_length() - 1))
    value = total_units / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel: the first fused fold+eval launch (2^log_n -> 2^(log_n-1)) -----
    peak, peak_src = load_peaks()
    dom = [ms for (kid, items, ms) in launches_timed if kid == 1 and items == (n >> 2)]
    alg_bytes = 48 * P_TABLES * n  # 32 B read + 16 B written per table element
    roofline = None
    if dom:
        avg_ms = sum(dom) / len(dom)
        ach = alg_bytes / (avg_ms * 1e-3) / 1e9
        kernel_ms = sum(ms for (_, _, ms) in launches_timed) / args.steps
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at 2^24 from the committed ncu --set full capture
        # (profiles/r01c_ncu_full_dense_round_prod3_fold_eval.csv): 1.610745 GB + 0.780195 GB per launch; other sizes: not captured
        traffic = 1610745000 + 780194560 if log_n == 24 else None
        roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "traffic_source": "profiles/r01c_ncu_full_dense_round_prod3_fold_eval.csv (ncu --set full, one launch)",
                    "kernel": "dense_round_kernel<SoProd3, fold+eval> (first fused round)", "avg_launch_ms": avg_ms,
                    "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                    "share_of_kernel_time": avg_ms / kernel_ms if kernel_ms else None,
                    "kernel_ms_per_step": kernel_ms, "modmul_per_launch": 12 * (n >> 2)}
        try:
            mm = ctx.bench_modmul(ilp=2, threads=128, blocks_per_sm=8, iters=1000)
            roofline["int_pipe"] = {"achieved_modmul_per_s": 12 * (n >> 2) / (avg_ms * 1e-3), "peak_modmul_per_s": mm,
                                    "frac": 12 * (n >> 2) / (avg_ms * 1e-3) / mm,
                                    "how": "peak = chains of dependent 8x32-bit Montgomery multiplications, ILP 2, 8 blocks x 128 thr per SM"}
        except Exception as e:  # pragma: no cover
            roofline["int_pipe"] = {"error": str(e)}

    # ---- end-to-end arm: host (pinned) tables in, round messages + final evals out -------------------
    job.prepare_host_inputs()
    for _ in range(min(args.warmup, 3)):
        job.prove_from_host()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        job.prove_from_host()
    e1.record(stream)
    barrier()
    wall = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(e0.elapsed_time(e1), 0.0)
    if dist is not None:
        t = torch.tensor([e2e_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_ms_step = e2e_ms / args.e2e_steps
    e2e = {"value": total_units / (e2e_ms_step * 1e-3), "unit": UNIT, "h2d_bytes_per_step": job.h2d_bytes * world,
           "d2h_bytes_per_step": job.d2h_bytes * world, "ms_per_step": e2e_ms_step, "wall_ms_per_step": wall / args.e2e_steps,
           "steps": args.e2e_steps}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        use_all_host_threads()
        val, sec, threads, reps = cpu_port_run(args.ref_log_n, 3, min_seconds=10.0)
        cpu = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"oracle C port (OpenMP), full Prod3 sumcheck over 2^{args.ref_log_n} x 3 tables, median of {reps} runs ({sec:.3f} s each)"}

    # ---- secondary: the whole `examples/pippenger` prover: BASELINE config[0] (x=16, d=8, 128 bit, clm 0) and the 2^20-point
    # shape of config[2] (x=20, d=10, 128 bit) on this one GPU -- "GKR-MSM prove ms @2^20 pts/128-bit" of BASELINE.json's metric
    pip, pip20 = None, None
    if world == 1 and not args.no_pippenger:
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("bench_pippenger", os.path.join(ROOT, "tools", "bench_pippenger.py"))
            bp = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(bp)

            def prove(x, d, reps):
                r = bp.run(argparse.Namespace(x_logsize=x, d_logsize=d, nbits=128, clm=0, reps=reps, seed=7, profile=False, python_host=False), ctx=ctx)
                return {"prove_ms": r["prove_ms_best"], "prove_ms_median": statistics.median(r["prove_ms_all"][1:]),
                        "what": "benchutils::run_pippenger (witness + phase-1 commitments + proof), wall clock, inputs on the host, SRS resident",
                        "config": f"x_logsize {x}, d_logsize {d}, nbits 128, clm 0 ({r['incidences']} point-digit incidences)",
                        "proof_bytes": r["proof_bytes"], "gpu_launches": r["gpu_launches"], "prove_ms_all": r["prove_ms_all"],
                        "round_waits": r["round_waits"], "round_wait_ms": r["round_wait_ms"]}

            pip = prove(16, 8, 4)
            if not args.no_pippenger_2e20:
                pip20 = prove(20, 10, 4)
        except Exception as e:  # pragma: no cover
            pip = pip or {"error": repr(e)}
            pip20 = pip20 or {"error": repr(e)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u256 (BLS12-381 Fr, 8x32-bit Montgomery limbs)", "data": "synthetic",
        "config": {"workload": f"dense Prod3 sumcheck (DenseSumcheckObjectSO + GenericSumcheckProtocol::prove), "
                               f"P=3 tables x 2^{log_n} Fr per GPU (BASELINE config[1])",
                   "log_n_per_gpu": log_n, "n_tables": P_TABLES, "gate": "Prod3Fn", "rounds": log_n + (world.bit_length() - 1),
                   "parallelism": f"hypercube sharded by top index bits over {world} GPU(s)",
                   "l2": "inputs (1.5 GiB per GPU) larger than the 126 MB L2", "transcript": "merlin on host, one challenge per round"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "pippenger_prove": pip, "pippenger_prove_2e20": pip20,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=24)
    ap.add_argument("--ref-log-n", type=int, default=20)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pippenger", action="store_true")
    ap.add_argument("--no-pippenger-2e20", action="store_true", help="skip the x=20 whole-prover leg (about 15 s of input generation)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
