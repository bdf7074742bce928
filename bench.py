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

    coracle.use_native()
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


def cpu_prover_run(x: int, d: int, nbits: int = 128, clm: int = 0, seed: int = 7):
    """The whole `examples/pippenger` prover on the host cores: oracle/c/pippenger_oracle.cpp (C++ / OpenMP restatement of
    benchutils::run_pippenger, pinned byte-for-byte to the python oracle and -- through tests/golden/pippenger_large.json --
    to the device prover), built with -march=native on this machine when gcc is here.  Same synthetic inputs as the GPU leg
    (tools/bench_pippenger.py): arithmetic-progression points, uniform nbits-bit scalars, mock SRS.  Key setup is NOT timed
    (neither is build_pippenger_data in the reference's bench, benches/pippenger.rs:40-45)."""
    from oracle import pippenger_oracle as PO
    from oracle.pyref import curves as CV
    from oracle.pyref.field import P as R_MOD
    from oracle.pyref.field import fq_vec_to_mont_u64, fr_vec_to_mont_u64

    native = True
    PO.lib(native)
    rng = np.random.default_rng(seed)
    n = 1 << x
    pts = PO.te_arithmetic_progression(0x1234567 + seed, 0x9E3779B97F4A7C15, n)
    raw = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    b = raw.view(np.uint8).reshape(n, 32).copy()
    b[:, nbits // 8:] = 0
    coefs = b.view(np.uint64).reshape(n, 4)
    y_size = (nbits + d - 1) // d
    yl = (y_size - 1).bit_length()
    r = [int.from_bytes(rng.bytes(32), "little") % R_MOD for _ in range(yl)]
    tau = int.from_bytes(rng.bytes(32), "little") % R_MOD
    t0 = time.perf_counter()
    key = PO.Key(fr_vec_to_mont_u64([tau])[0], fq_vec_to_mont_u64([CV.G1_GEN[0], CV.G1_GEN[1]]).reshape(12), x + clm, fr_vec_to_mont_u64([2])[0], native=native)
    t_setup = time.perf_counter() - t0
    out = PO.run_pippenger(key, pts, coefs, fr_vec_to_mont_u64(r), d, x, nbits, clm)
    key.close()
    threads = PO.num_threads(native)
    return {"value": out["seconds"]["total"] * 1e3, "unit": "ms", "cores": threads, "kind": "port",
            "sample": f"oracle/c/pippenger_oracle.cpp (C++/OpenMP port of benchutils::run_pippenger), one proof at x={x}, d={d}, {nbits} bit, clm {clm}, "
                      f"{threads} threads, build {os.path.basename(key.l._path)}; SRS setup {t_setup:.1f} s not included",
            "phases_s": out["seconds"], "proof_bytes": len(out["proof"]),
            "note": "the Rust reference's `--features parallel` build is not measurable in this image (no cargo, nightly + un-vendored git deps)"}


def cpu_prover_baselines(budget_s: float):
    """x = 16 (BASELINE config[0]) always; the 2^20-point shape only when its extrapolated time fits the budget"""
    res = {}
    try:
        res["pippenger_prove"] = cpu_prover_run(16, 8)
        est20 = res["pippenger_prove"]["value"] * 1e-3 * 14.0  # 13 x the incidences + a 16 x larger opening
        if est20 <= budget_s:
            res["pippenger_prove_2e20"] = cpu_prover_run(20, 10)
        else:
            res["pippenger_prove_2e20"] = {"skipped": f"extrapolated {est20:.0f} s of CPU time exceeds the {budget_s:.0f} s budget (--cpu-prover-budget)"}
    except Exception as e:  # pragma: no cover
        res["error"] = repr(e)
    return res


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is a Rust crate
    (nightly + un-vendored git deps) that cannot be built in this image, so this arm times the oracle's C / C++
    ports of the same algorithms on all host threads -- rank 0 only:
      metric line : oracle/c/gkr_oracle.c, the dense sumcheck object, on the SAME 2^log_n x 3 workload as the GPU arm
      extra keys  : oracle/c/pippenger_oracle.cpp, the whole prover (the first half of BASELINE.json's metric)"""
    if rank != 0:
        return
    use_all_host_threads()
    log_n = args.ref_log_n if args.ref_log_n > 0 else args.log_n
    for _ in range(min(args.warmup, 1)):
        cpu_port_run(log_n, 1)
    val, sec, threads, reps = cpu_port_run(log_n, min(args.steps, 5))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": reps,
        "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u256 (BLS12-381 Fr, 4x64-bit Montgomery limbs)", "data": "synthetic",
        "config": {"workload": f"dense Prod3 sumcheck, P=3 tables x 2^{args.log_n} Fr (config[1])", "same_size_as_gpu_arm": log_n == args.log_n,
                   "sample": f"each step = one full sumcheck over 2^{log_n} x 3 tables",
                   "what": "reference PORT (oracle/c/gkr_oracle.c, OpenMP): the Rust crate cannot be built in this image"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"full Prod3 sumcheck, 2^{log_n} x 3 tables, median of {reps} runs, OpenMP {threads} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_pippenger:
        line["pippenger_cpu"] = cpu_prover_baselines(args.cpu_prover_budget)
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

    # ---- roofline: every (kernel, size) group of the timed launches; the DOMINANT kernel is the longest one -------------
    peak, peak_src = load_peaks()
    groups = {}
    for (kid, items, ms) in launches_timed:
        groups.setdefault((kid, items), []).append(ms)
    KNAME = {0: "dense round kernel, eval only (round 0)", 1: "dense round kernel, fused fold(k) + eval(k+1)", 2: "dense gate sum", 3: "dense fold"}
    # wide (32x32+64 -> 64 bit) multiply-adds per item, from the generator's operation counts (tools/gen_field.py):
    #   fold by a 128-bit challenge 56, Montgomery product 112, unreduced multiply-accumulate 64
    #   eval only : 3 nodes x (112 + 64)                       = 528 per pair-triple
    #   fused     : 6 folds x 56 + 3 nodes x (112 + 64)         = 864 per quad-triple
    WIDE = {0: 528, 1: 864}
    roofline_all = []
    for (kid, items), mss in groups.items():
        if kid not in (0, 1):
            continue
        avg_ms = sum(mss) / len(mss)
        # algorithmic bytes: eval reads 2 x 32 B per pair and table; fused reads 4 x 32 B and writes 2 x 32 B per quad and table
        alg = (64 if kid == 0 else 192) * P_TABLES * items
        roofline_all.append({"kernel": KNAME[kid], "items": items, "launches": len(mss), "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": alg,
                             "achieved_gbs": alg / (avg_ms * 1e-3) / 1e9, "frac_hbm": alg / (avg_ms * 1e-3) / 1e9 / peak,
                             "wide_mads_per_launch": WIDE[kid] * items, "achieved_wide_mads_per_s": WIDE[kid] * items / (avg_ms * 1e-3)})
    roofline_all.sort(key=lambda e: -e["avg_launch_ms"])
    kernel_ms = sum(ms for (_, _, ms) in launches_timed) / args.steps
    roofline = None
    if roofline_all:
        dom = roofline_all[0]
        roofline = {"bound": "hbm", "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": dom["frac_hbm"], "traffic": None,
                    "kernel": dom["kernel"] + f", {dom['items']} items (the LONGEST launch of the step)", "avg_launch_ms": dom["avg_launch_ms"],
                    "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"], "peak_source": peak_src,
                    "share_of_kernel_time": dom["avg_launch_ms"] / kernel_ms if kernel_ms else None, "kernel_ms_per_step": kernel_ms,
                    "note": "both large dense kernels are bound by the 32-bit multiplier pipe before HBM (int_pipe below); "
                            "the fused fold+eval round is the HBM-relevant one (roofline_all[1])"}
        # whole step against HBM: 96 * P * 2^n algorithmic bytes over the step time
        roofline["whole_step"] = {"algorithmic_bytes": 96 * P_TABLES * n, "achieved_gbs": 96 * P_TABLES * n / (ms_step * 1e-3) / 1e9,
                                  "frac_hbm": 96 * P_TABLES * n / (ms_step * 1e-3) / 1e9 / peak}
        # measured DRAM traffic of the dominant kernels: only from an ncu capture of THIS build (profiles/r02_traffic.json records
        # the sha256 of the kernel sources it was taken from); stale captures are not reported
        try:
            import hashlib
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            h = hashlib.sha256()
            for f in tj["sources"]:
                h.update(open(os.path.join(ROOT, f), "rb").read())
            if h.hexdigest() == tj["sources_sha256"] and tj.get("log_n") == log_n:
                for e in roofline_all:
                    key = "eval" if e["kernel"].startswith("dense round kernel, eval") else "fused"
                    if e["items"] == tj[key]["items"]:
                        e["traffic"] = tj[key]["dram_bytes"]
                roofline["traffic"] = roofline_all[0].get("traffic")
                roofline["traffic_source"] = tj["source"]
            else:
                roofline["traffic_source"] = "no ncu capture of this exact kernel build (profiles/r02_traffic.json is older than the sources)"
        except Exception:
            roofline["traffic_source"] = "no ncu capture recorded for this build"
        # multiplier-pipe roofline: achieved wide multiply-adds per second over the measured peak of the same instruction
        try:
            from tools import lablib
            # the field routines are chains of mad.lo.cc / madc.hi.cc pairs = IMAD.WIDE.U32.X, which retires at HALF the rate of the
            # carry-less IMAD.WIDE.U32 on this part (both probes in csrc/lab/kernel_lab.cu): the chained rate is the peak that applies
            pk_x = max(lablib.imad_wide_x_peak(ctx, ilp=ilp, blocks_per_sm=8) for ilp in (2, 4))
            pk_plain = max(lablib.imad_wide_peak(ctx, ilp=ilp) for ilp in (8, 16))
            for e in roofline_all:
                e["frac_int_pipe"] = e["achieved_wide_mads_per_s"] / pk_x
            roofline["int_pipe"] = {"achieved_wide_mads_per_s": dom["achieved_wide_mads_per_s"], "peak_wide_mads_per_s": pk_x,
                                    "frac": dom["achieved_wide_mads_per_s"] / pk_x,
                                    "peak_carryless_wide_mads_per_s": pk_plain, "frac_of_carryless_peak": dom["achieved_wide_mads_per_s"] / pk_plain,
                                    "how": "peak = carry-chained multiply-add probe (mad.lo.cc / madc.hi.cc pairs = IMAD.WIDE.U32.X, the form the field "
                                           "routines use), 2-4 independent chains per thread, 8 x 256 threads per SM (csrc/lab/kernel_lab.cu); it is half "
                                           "the carry-less IMAD.WIDE.U32 rate (peak_carryless).  achieved = static multiply-add count of the field routines "
                                           "x items / measured launch time.  A carry-free 9 x 29-bit rewrite of the round kernel (full-rate IMAD.WIDE.U32, "
                                           "lab variant 20) was measured 2x slower: it needs 1.7x the multiply-adds and its 64-bit column carries load the "
                                           "ALU pipe (DESIGN.md 3.1)"}
        except Exception as e:  # pragma: no cover
            roofline["int_pipe"] = {"error": str(e)}
        roofline["roofline_all"] = roofline_all

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

    # ---- N > 1: strong-scaling leg (the SAME 2^log_n total table split over the ranks) next to the weak one above --------
    strong = None
    if world > 1 and log_n - (world.bit_length() - 1) >= 10:
        ln_local = log_n - (world.bit_length() - 1)
        sjob = ShardedProd3Sumcheck(ctx, log_n_local=ln_local, rank=rank, world=world, dist=dist, seed=SEED, exchange=job.exchange)
        for _ in range(3):
            sjob.prove_resident()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(args.steps):
            sjob.prove_resident()
        s1.record(stream)
        barrier()
        t = torch.tensor([s0.elapsed_time(s1)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s_ms = float(t.item()) / args.steps
        strong = {"total_log_n": log_n, "log_n_per_gpu": ln_local, "ms_per_step": s_ms, "value": elem_rounds(log_n) / (s_ms * 1e-3), "unit": UNIT,
                  "what": f"strong scaling: ONE 2^{log_n} x 3 sumcheck split by top index bits over {world} GPUs; compare with the 1-GPU ms_per_step of the "
                          "weak line (same total work).  Limiter: the ~20 small rounds and the per-round host exchange are latency, not bandwidth"}
        del sjob

    # ---- the ragged Deg2 sumcheck of one addition layer with the bucket rows split over the ranks (VecVec by rows) ----
    vecvec_rows = None
    if not args.no_vecvec:
        from gkr_msm_b200.sharded import ShardedVecVecSumcheck
        try:
            vjob = ShardedVecVecSumcheck(ctx, row_log=args.vecvec_row_log, col_local_log=args.vecvec_col_log, rank=rank, world=world,
                                         exchange=job.exchange if world > 1 else None)
            for _ in range(3):
                vjob.prove()
            barrier()
            v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            v0.record(stream)
            for _ in range(args.steps):
                vjob.prove()
            v1.record(stream)
            barrier()
            v_ms = v0.elapsed_time(v1)
            if dist is not None:
                t = torch.tensor([v_ms], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                v_ms = float(t.item())
            v_ms /= args.steps
            vecvec_rows = {"ms_per_step": v_ms, "table_elements_per_gpu": vjob.elements, "table_elements_per_s": vjob.elements * world / (v_ms * 1e-3),
                           "scaling": "weak", "num_vars": vjob.num_vars,
                           "what": f"VecVecDeg2Sumcheck::prove, gate twisted_edwards_add_l1 (6 polynomials), 2^{args.vecvec_col_log} rows of "
                                   f"2^{args.vecvec_row_log} elements per GPU, rows split by the top bits of the row index over {world} GPU(s); "
                                   "per sparse round two field elements per rank through the exchange, dense tail sharded like the headline"}
            del vjob
        except Exception as e:  # pragma: no cover
            vecvec_rows = {"error": repr(e)}

    # ---- N > 1: the whole prover with the commitment MSMs split by point range over the ranks (csrc/msm_team.cu) -----------
    pip20_multi = None
    if world > 1 and not args.no_pippenger and not args.no_pippenger_2e20:
        import importlib.util
        spec = importlib.util.spec_from_file_location("bench_pippenger", os.path.join(ROOT, "tools", "bench_pippenger.py"))
        bp = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bp)
        tname = f"/gkr_team_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}"
        pa = argparse.Namespace(x_logsize=20, d_logsize=10, nbits=128, clm=0, reps=4, seed=7, profile=False, python_host=False, gpus=world,
                                team_worker=rank, team_name=tname, team_tau="%x" % bp.srs_tau(7))
        try:
            if rank == 0:
                r = bp.run(pa, ctx=ctx, team_name=tname)
                pip20_multi = {"prove_ms": r["prove_ms_best"], "prove_ms_median": statistics.median(r["prove_ms_all"][1:]), "prove_ms_all": r["prove_ms_all"],
                               "config": f"x_logsize 20, d_logsize 10, nbits 128, clm 0 on {world} GPUs", "proof_bytes": r["proof_bytes"],
                               "multi_gpu": r["multi_gpu"], "gpu_launches_rank0": r["gpu_launches"]}
            else:
                bp.team_worker(pa, ctx=ctx)
        except Exception as e:  # pragma: no cover
            pip20_multi = {"error": repr(e)}
        barrier()

    # ---- N >= 4: BASELINE config[3] itself -- x = 24, 253-bit scalars, clm 2: 430 GiB of live tables.  Rank 0 proves with its tables
    # pooled over the HBM of all ranks' GPUs (gkr_ctx_peer_pool, NVLink peer access); the other ranks serve their share of the G1
    # work (msm_team.cu) from the same GPUs.  One proof is ~6 s on 8 GPUs, ~12 s on 4; inputs and SRS take another ~35 s.
    # OPT-IN (--config3, 8 GPUs): with only 4 GPUs the team workers' own memory leaves no peer with room for the largest single
    # slab (77 GiB) -- there the instance runs with the peer pool alone (tools/bench_pippenger.py --peer-pool 4, 11.7 s).
    config3 = None
    if world >= 8 and args.config3 and not args.no_pippenger:
        import importlib.util
        import subprocess
        spec = importlib.util.spec_from_file_location("bench_pippenger", os.path.join(ROOT, "tools", "bench_pippenger.py"))
        bp = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bp)
        os.environ.setdefault("GKR_PEER_RESERVE_GIB", "20")  # what every GPU keeps for its own team worker
        os.environ.setdefault("GKR_TEAM_TIMEOUT_S", "180")
        tname = f"/gkr_c3_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}"
        pfile = f"/dev/shm/gkr_c3_points_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}.npy"
        pa = argparse.Namespace(x_logsize=24, d_logsize=8, nbits=253, clm=2, reps=2, seed=7, profile=False, python_host=False, gpus=world,
                                team_worker=rank, team_name=tname, team_tau="%x" % bp.srs_tau(7), peer_pool=world, precompute_c=0,
                                points_file=pfile, dump="", mem=False, gen_procs=1)
        try:
            if rank == 0:
                subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "bench_pippenger.py"), "--x-logsize", "24", "--gen-procs", "16",
                                       "--gen-only", pfile], timeout=600)
                r = bp.run(pa, ctx=ctx, team_name=tname)
                os.unlink(pfile)
                config3 = {"prove_ms": r["prove_ms_best"], "prove_ms_all": r["prove_ms_all"], "proof_bytes": r["proof_bytes"],
                           "config": f"BASELINE config[3]: x_logsize 24, d_logsize 8, nbits 253 (full width), commitment-log-multiplicity 2 on {world} GPUs",
                           "incidences": r["incidences"], "peer_pool_peak_gib": r["peer_pool_peak_gib"], "gpu_launches_rank0": r["gpu_launches"],
                           "how": "rank 0 runs gkr_run_pippenger; its tables beyond 180 GB live in the HBM of the other GPUs (NVLink peer access); the "
                                  "commitment MSMs and the c / d bucket sums are split over all GPUs (csrc/msm_team.cu).  Verification of the same "
                                  "proof: profiles/r02_config3_x23_x24_peer_pool.txt"}
            else:
                bp.team_worker(pa, ctx=ctx)
        except Exception as e:  # pragma: no cover
            config3 = {"error": repr(e)}
        barrier()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        use_all_host_threads()
        ref_ln = args.ref_log_n if args.ref_log_n > 0 else log_n
        val, sec, threads, reps = cpu_port_run(ref_ln, 3, min_seconds=5.0)
        cpu = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"oracle C port (oracle/c/gkr_oracle.c, OpenMP, -march=native when gcc is on the box), full Prod3 sumcheck over 2^{ref_ln} x 3 tables "
                         f"(the GPU arm's size: {ref_ln == log_n}), median of {reps} runs ({sec:.3f} s each)",
               "note": "the Rust reference's `--features parallel` rayon build is not measurable in this image (no cargo)"}

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
        if not args.no_cpu_baseline:  # the same prover on the host cores (oracle/c/pippenger_oracle.cpp)
            use_all_host_threads()
            cb = cpu_prover_baselines(args.cpu_prover_budget)
            if isinstance(pip, dict) and "pippenger_prove" in cb:
                pip["cpu_baseline"] = cb["pippenger_prove"]
            if isinstance(pip20, dict) and "pippenger_prove_2e20" in cb:
                pip20["cpu_baseline"] = cb["pippenger_prove_2e20"]
            if "error" in cb and isinstance(pip, dict):
                pip["cpu_baseline"] = {"error": cb["error"]}
    if world > 1:
        pip20 = pip20_multi

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u256 (BLS12-381 Fr, 8x32-bit Montgomery limbs)", "data": "synthetic",
        "config": {"workload": f"dense Prod3 sumcheck (DenseSumcheckObjectSO + GenericSumcheckProtocol::prove), "
                               f"P=3 tables x 2^{log_n} Fr per GPU (BASELINE config[1])",
                   "log_n_per_gpu": log_n, "n_tables": P_TABLES, "gate": "Prod3Fn", "rounds": log_n + (world.bit_length() - 1),
                   "parallelism": f"hypercube sharded by top index bits over {world} GPU(s)",
                   "collective": ("none (one GPU)" if world == 1 else
                                  "per-round partial sums (deg x 32 B per rank) all-gathered through a POSIX shared-memory segment between the ranks of the "
                                  "box and reduced on every host (they must reach the host transcript anyway); NCCL only for rendezvous, barriers and "
                                  "the max-over-ranks timing all-reduce; no device collective on the data path"),
                   "l2": "inputs (1.5 GiB per GPU) larger than the 126 MB L2", "transcript": "merlin on host, one challenge per round"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "pippenger_prove": pip, "pippenger_prove_2e20": pip20, "strong_scaling": strong, "vecvec_rows": vecvec_rows, "pippenger_config3_x24": config3,
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
    ap.add_argument("--ref-log-n", type=int, default=0, help="size of the CPU arm's sumcheck (0: the GPU arm's --log-n)")
    ap.add_argument("--cpu-prover-budget", type=float, default=150.0, help="seconds of CPU time allowed for the x=20 whole-prover baseline")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pippenger", action="store_true")
    ap.add_argument("--config3", action="store_true", help="N >= 8: also prove BASELINE config[3] at x = 24 (about a minute of inputs / SRS, 6 s per proof)")
    ap.add_argument("--no-vecvec", action="store_true", help="skip the row-sharded ragged sumcheck leg")
    ap.add_argument("--vecvec-row-log", type=int, default=11)
    ap.add_argument("--vecvec-col-log", type=int, default=12, help="log2 of the rows per GPU of the row-sharded ragged sumcheck leg")
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
