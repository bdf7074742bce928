#!/usr/bin/env python3
"""MEASUREMENT ONLY: times the variants of the dense round kernel compiled into csrc/lab/kernel_lab.cu (B200 only)
and prints the multiplier-pipe peak next to them.   python tools/kernel_lab.py [log_n] [variants...]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gkr_msm_b200 as g  # noqa: E402
from tools import lablib  # noqa: E402

ctx = g.Context(0)
lib = lablib.lab()
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
only = [int(v) for v in sys.argv[2:]]
n = 1 << log_n
tabs = [ctx.synth(j, n) for j in range(3)]
outs = [ctx.alloc(n // 2) for _ in range(3)]
vp = C.c_void_p
tin = (vp * 3)(*[t.h for t in tabs])
tout = (vp * 3)(*[t.h for t in outs])
names = {0: "regs, >=3 blocks/SM", 1: "regs, >=4 blocks/SM (spills)", 2: "smem accs, >=4 blocks/SM", 3: "smem accs, >=5 blocks/SM",
         4: "regs, >=2 blocks/SM", 5: "regs, 3 blocks, L2 prefetch +1", 6: "regs, 3 blocks, L2 prefetch +2", 7: "regs, 3 blocks, L2 prefetch +4",
         8: "cp.async staged, >=3 blocks/SM", 9: "cp.async staged, >=4 blocks/SM (spills)",
         10: "node-split, >=4 blocks of 96", 11: "node-split, >=5", 12: "node-split, >=6", 13: "node-split, >=7",
         14: "node-split, >=8", 15: "node-split, >=10",
         20: "reduced radix 9x29, >=3 blocks/SM", 21: "reduced radix 9x29, >=2 blocks/SM", 22: "reduced radix 9x29, >=4 blocks/SM (spills)"}
if not os.environ.get("LAB_SKIP_PEAKS"):
    for ilp in (4, 8, 16):
        print(f"IMAD.WIDE peak, {ilp} independent chains/thread: {lablib.imad_wide_peak(ctx, ilp=ilp):.4g} wide multiply-adds/s", flush=True)
    for bps in (2, 4, 8):
        print(f"IMAD.WIDE peak, 8 x 8 product pattern (distinct operand registers), {bps} x 256 threads/SM: {lablib.imad_wide_rot_peak(ctx, blocks_per_sm=bps):.4g} wide multiply-adds/s", flush=True)
    for ilp in (1, 2, 4):
        for bps in (2, 4, 8):
            print(f"IMAD.WIDE.X (carry-chained) peak, {ilp} chains/thread, {bps} x 256 threads/SM: {lablib.imad_wide_x_peak(ctx, ilp=ilp, blocks_per_sm=bps):.4g} wide multiply-adds/s", flush=True)
for mode in (0, 1):
    bytes_ = (32 * 3 * n) if mode == 0 else (48 * 3 * n)
    for variant in sorted(names):
        if only and variant not in only:
            continue
        for gm in (1, 2):
            ms = C.c_float(0)
            bps = C.c_int(0)
            rc = lib.gkr_lab_dense_prod3(ctx.h, variant, mode, tin, tout, C.c_uint64(n), 10, gm, C.byref(ms), C.byref(bps))
            if rc:
                print("variant", variant, "failed", rc, ctx.last_error() if hasattr(ctx, "last_error") else "")
                continue
            print(f"mode {mode} ({'eval' if mode == 0 else 'fast fold+eval'}) 2^{log_n} variant {variant} [{names[variant]}] grid x{gm}: "
                  f"{ms.value:.4f} ms  {bytes_ / ms.value / 1e6:.0f} GB/s  ({bps.value} blocks/SM)", flush=True)
