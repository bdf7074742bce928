#!/usr/bin/env python3
"""Turn the raw ncu output of tools/profile_r02.sh (gpurun_out/, scratch) into the committed summaries under profiles/.

    python tools/profile_r02_summarise.py

Writes
  profiles/r02_launches_*.csv                  the launch lists (gpu__time_duration.sum per launch), unchanged
  profiles/r02_launch_share_*.txt              kernel share of each run, aggregated by kernel name
  profiles/r02_ncu_full_<kernel>.csv           the counters DESIGN.md cites from each `--set full` capture
  profiles/r02_traffic.json                    dram bytes of the two dominant dense kernels + sha256 of the kernel sources the
                                               capture was taken from (bench.py reports `traffic` only while that hash matches)
  profiles/r02_deg2_rounds_x20_summary.txt     every Deg2 round of an x = 20 proof: the ~444-block round of the round-1 verdict
                                               and the largest ragged rounds with their DRAM / multiplier-pipe fractions
"""
import collections
import csv
import hashlib
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")
KERNEL_SOURCES = ["gkr-msm_b200/csrc/dense_kernel.cuh", "gkr-msm_b200/csrc/field_gen.cuh", "gkr-msm_b200/csrc/field.cuh", "gkr-msm_b200/csrc/gates.cuh"]

KEEP = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum",
    "sm__inst_executed_pipe_lsu.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def read_launch_list(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    H = rows[hi]
    return H, [r for r in rows[hi + 1:] if len(r) == len(H)]


def share(path, out):
    H, rows = read_launch_list(path)
    kn, mv = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        name = re.sub(r"\(.*", "", r[kn])[:90]
        if "peak_kernel" in name or "modmul_bench" in name:  # measurement probes of bench.py's int_pipe leg, not part of a step
            continue
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# {os.path.basename(path)}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms of kernel time (ncu: cold, serialised)\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1] / 1e6:10.3f} ms  {100 * v[1] / tot:5.1f} %  {v[0]:6d} launches  {k}\n")


def full(raw_csv, out):
    rows = list(csv.reader(open(raw_csv, errors="ignore")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hi], rows[hi + 1]
    vals = {}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["launch", "metric", "unit", "value"])
        for li, r in enumerate(rows[hi + 2:]):
            for h, u, v in zip(hdr, units, r):
                if h in KEEP:
                    w.writerow([li, h, u, v])
                    if li == 0:
                        vals[h] = (v, u)
    return vals


def num(v):
    return float(v.replace(",", ""))


def to_bytes(v, u):
    x = num(v)
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def main():
    os.makedirs(DST, exist_ok=True)
    for name in ("r02_launches_bench_prod3_2e24.csv", "r02_launches_pippenger_x16.csv", "r02_launches_pippenger_x20.csv"):
        p = os.path.join(SRC, name)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(DST, name))
            share(p, os.path.join(DST, name.replace("launches", "launch_share").replace(".csv", ".txt")))
    traffic = {}
    for rep, key, out in (("r02_dense_eval", "eval", "r02_ncu_full_dense_round_eval_2e24.csv"), ("r02_dense_fused", "fused", "r02_ncu_full_dense_round_staged_fold_eval_2e24.csv"),
                          ("r02_msm_light", None, "r02_ncu_full_msm_accumulate_light_2e20.csv")):
        p = os.path.join(SRC, rep + ".raw.csv")
        if not os.path.exists(p):
            continue
        v = full(p, os.path.join(DST, out))
        if key:
            rd, wr = to_bytes(*v["dram__bytes_read.sum"]), to_bytes(*v["dram__bytes_write.sum"])
            grid = int(num(v["launch__grid_size"][0]))
            traffic[key] = {"dram_bytes": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "kernel": v["Kernel Name"][0],
                            "grid": grid, "items": (1 << 23) if key == "eval" else (1 << 22)}
    if traffic:
        h = hashlib.sha256()
        for f in KERNEL_SOURCES:
            h.update(open(os.path.join(ROOT, f), "rb").read())
        traffic.update({"log_n": 24, "sources": KERNEL_SOURCES, "sources_sha256": h.hexdigest(),
                        "source": "profiles/r02_ncu_full_dense_round_*_2e24.csv (ncu --set full --clock-control none, first 2^24-sized launch of each kernel)"})
        json.dump(traffic, open(os.path.join(DST, "r02_traffic.json"), "w"), indent=1)
    p = os.path.join(SRC, "r02_deg2_rounds_x20.csv")
    if os.path.exists(p):
        H, rows = read_launch_list(p)
        idx = {n: H.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
        per = collections.OrderedDict()
        for r in rows:
            d = per.setdefault(r[idx["ID"]], {"kernel": re.sub(r"\(.*", "", r[idx["Kernel Name"]])})
            d[r[idx["Metric Name"]]] = (r[idx["Metric Value"]], r[idx["Metric Unit"]])
        L = []
        for lid, d in per.items():
            try:
                t = num(d["gpu__time_duration.sum"][0]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(d["gpu__time_duration.sum"][1], 1e-3)
                L.append((int(lid), d["kernel"], int(num(d["launch__grid_size"][0])), t, d))
            except Exception:
                continue
        with open(os.path.join(DST, "r02_deg2_rounds_x20_summary.txt"), "w") as f:
            f.write(f"# {len(L)} Deg2 round launches of one x = 20 proof (ncu, one pass of cheap counters; times cold and serialised)\n")
            f.write(f"# total {sum(x[3] for x in L) / 1e3:.2f} ms\n")
            cols = ["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
                    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
                    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
                    "launch__registers_per_thread"]
            f.write("# id  grid  time_us  dram%  fmaheavy%  issue%  no_instr  long_sb  warps%  regs  kernel\n")

            def line(x):
                d = x[4]
                return f"{x[0]:6d} {x[2]:6d} {x[3]:9.1f}  " + "  ".join(f"{num(d[c][0]):7.2f}" if c in d else "    n/a" for c in cols) + f"  {x[1]}\n"

            f.write("## mid-size rounds (300..700 blocks) -- the round-1 verdict measured one of these at 77.7 us, no_instruction 4.6\n")
            for x in [x for x in L if 300 <= x[2] <= 700][:12]:
                f.write(line(x))
            f.write("## the 12 longest rounds\n")
            for x in sorted(L, key=lambda x: -x[3])[:12]:
                f.write(line(x))
    print(sorted(os.listdir(DST)))


if __name__ == "__main__":
    sys.exit(main())
