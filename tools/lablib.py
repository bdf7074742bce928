"""MEASUREMENT ONLY: ctypes loader of lib/libgkr_lab.so (gkr-msm_b200/csrc/lab/, `make -C gkr-msm_b200 lab`) -- the
kernel-variant lab and the pipe probes.  Used by tools/kernel_lab.py and bench.py's int_pipe leg; the product library does
not depend on it."""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAB = os.path.join(ROOT, "gkr-msm_b200", "lib", "libgkr_lab.so")
_lab = None


def lab():
    global _lab
    if _lab is None:
        if not os.path.exists(LAB):
            raise RuntimeError("libgkr_lab.so is not built: run `make -C gkr-msm_b200 lab`")
        _lab = C.CDLL(LAB)
        vp = C.c_void_p
        _lab.gkr_lab_dense_prod3.restype = C.c_int
        _lab.gkr_lab_dense_prod3.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        _lab.gkr_lab_imad_wide_peak.restype = C.c_int
        _lab.gkr_lab_imad_wide_peak.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        _lab.gkr_lab_imad_wide_rot_peak.restype = C.c_int
        _lab.gkr_lab_imad_wide_rot_peak.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        _lab.gkr_lab_imad_wide_x_peak.restype = C.c_int
        _lab.gkr_lab_imad_wide_x_peak.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        _lab.gkr_bench_modmul.restype = C.c_int
        _lab.gkr_bench_modmul.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    return _lab


def imad_wide_peak(ctx, ilp=8, threads=256, blocks_per_sm=8, iters=4000) -> float:
    """wide (32x32+64->64) multiply-adds per second with `ilp` independent chains per thread"""
    out = C.c_double(0)
    rc = lab().gkr_lab_imad_wide_peak(ctx.h, ilp, threads, blocks_per_sm, iters, C.byref(out))
    if rc:
        raise RuntimeError(f"gkr_lab_imad_wide_peak failed: {rc}")
    return out.value


def imad_wide_rot_peak(ctx, threads=256, blocks_per_sm=4, iters=500) -> float:
    """wide multiply-adds per second of an 8 x 8 limb product pattern (64 instructions over 8 + 8 distinct multiplier registers and
    15 accumulator columns): no operand-reuse-cache help, the rate a multi-precision product can reach"""
    out = C.c_double(0)
    rc = lab().gkr_lab_imad_wide_rot_peak(ctx.h, threads, blocks_per_sm, iters, C.byref(out))
    if rc:
        raise RuntimeError(f"gkr_lab_imad_wide_rot_peak failed: {rc}")
    return out.value


def imad_wide_x_peak(ctx, ilp=2, threads=256, blocks_per_sm=4, iters=4000) -> float:
    """carry-chained wide multiply-adds per second (mad.lo.cc / madc.hi.cc pairs = IMAD.WIDE.U32.X), `ilp` independent 8-pair chains"""
    out = C.c_double(0)
    rc = lab().gkr_lab_imad_wide_x_peak(ctx.h, ilp, threads, blocks_per_sm, iters, C.byref(out))
    if rc:
        raise RuntimeError(f"gkr_lab_imad_wide_x_peak failed: {rc}")
    return out.value


def modmul_chain_rate(ctx, ilp=2, threads=128, blocks_per_sm=8, iters=2000) -> float:
    """dependent Montgomery products per second (latency-shaped; NOT a peak)"""
    out = C.c_double(0)
    rc = lab().gkr_bench_modmul(ctx.h, ilp, threads, blocks_per_sm, iters, C.byref(out))
    if rc:
        raise RuntimeError(f"gkr_bench_modmul failed: {rc}")
    return out.value
