#!/usr/bin/env python3
"""Generator + CPU simulator for the multi-precision Montgomery arithmetic used by the CUDA kernels.

There is no GPU in the build container, so every carry chain is generated from a tiny PTX-subset IR
that is (a) printed as ONE inline-asm block per field operation (carry flag never crosses an asm
boundary) and (b) executed by the simulator below against python big-int arithmetic on random and
extreme operands before the header is written.  `python tools/gen_field.py --check` re-runs the
simulation only; `python tools/gen_field.py` rewrites gkr-msm_b200/csrc/field_gen.cuh.

Multiplication: word-serial Montgomery (CIOS) with the even/odd column split, so that each
`mad.lo.cc` / `madc.hi.cc` pair is fused by ptxas into one IMAD.WIDE.U32(.X) (64-bit multiply-add with
predicate carry) -- 2*N*N/2 wide multiply-adds for an N-limb field instead of 4*N*N/2 narrow ones.
"""
from __future__ import annotations

import argparse
import os
import random

M32 = 0xFFFFFFFF


class Prog:
    """A straight-line PTX-subset program over named 32-bit registers."""

    def __init__(self):
        self.ins = []          # (op, dst, [srcs])  srcs are reg names or int immediates
        self.tmp = []
        self.ntmp = 0

    def t(self):
        name = f"t{self.ntmp}"
        self.ntmp += 1
        self.tmp.append(name)
        return name

    def emit(self, op, dst, *srcs):
        self.ins.append((op, dst, list(srcs)))


def simulate(prog: Prog, regs: dict) -> dict:
    regs = dict(regs)
    cc = 0

    def val(s):
        return s & M32 if isinstance(s, int) else regs[s]

    for op, dst, srcs in prog.ins:
        v = [val(s) for s in srcs]
        if op == "mov":
            r = v[0]
        elif op == "mul.lo":
            r = (v[0] * v[1]) & M32
        elif op == "mul.hi":
            r = (v[0] * v[1]) >> 32
        elif op in ("mad.lo", "mad.lo.cc", "madc.lo", "madc.lo.cc", "mad.hi", "mad.hi.cc", "madc.hi", "madc.hi.cc"):
            prod = v[0] * v[1]
            part = (prod & M32) if ".lo" in op else (prod >> 32)
            s = part + v[2] + (cc if op.startswith("madc") else 0)
            r = s & M32
            if op.endswith(".cc"):
                cc = s >> 32
        elif op in ("add", "add.cc", "addc", "addc.cc"):
            s = v[0] + v[1] + (cc if op.startswith("addc") else 0)
            r = s & M32
            if op.endswith(".cc"):
                cc = s >> 32
        elif op in ("sub", "sub.cc", "subc", "subc.cc"):
            # PTX: sub.cc sets CF = borrow; subc subtracts borrow.
            s = v[0] - v[1] - (cc if op.startswith("subc") else 0)
            r = s & M32
            if op.endswith(".cc"):
                cc = 1 if s < 0 else 0
        elif op == "and":
            r = v[0] & v[1]
        elif op == "selp.ne0":      # dst = (v2 != 0) ? v0 : v1
            r = v[0] if v[2] != 0 else v[1]
        else:
            raise ValueError(op)
        regs[dst] = r
    return regs


PTX_OP = {
    "mov": "mov.b32", "mul.lo": "mul.lo.u32", "mul.hi": "mul.hi.u32",
    "mad.lo": "mad.lo.u32", "mad.lo.cc": "mad.lo.cc.u32", "madc.lo": "madc.lo.u32", "madc.lo.cc": "madc.lo.cc.u32",
    "mad.hi": "mad.hi.u32", "mad.hi.cc": "mad.hi.cc.u32", "madc.hi": "madc.hi.u32", "madc.hi.cc": "madc.hi.cc.u32",
    "add": "add.u32", "add.cc": "add.cc.u32", "addc": "addc.u32", "addc.cc": "addc.cc.u32",
    "sub": "sub.u32", "sub.cc": "sub.cc.u32", "subc": "subc.u32", "subc.cc": "subc.cc.u32",
    "and": "and.b32",
}


def to_cuda(prog: Prog, name: str, outs: list, ins: list, sig: str, inout: list | None = None) -> str:
    """Print prog as `__device__ __forceinline__ void name(sig)` holding a single asm block.
    outs / ins: lists of (regname, c_expr).  inout: regs that are both read and written ("+r")."""
    inout = inout or []
    opmap = {}
    cons = []
    for reg, expr in inout:
        opmap[reg] = f"%{len(opmap)}"
        cons.append(("+r", expr))
    for reg, expr in outs:
        opmap[reg] = f"%{len(opmap)}"
        cons.append(("=r", expr))
    n_out = len(opmap)
    for reg, expr in ins:
        opmap[reg] = f"%{len(opmap)}"
        cons.append(("r", expr))

    def o(s):
        if isinstance(s, int):
            return f"0x{s & M32:08x}"
        return opmap.get(s, s)

    lines = ["{"]
    if prog.tmp:
        lines.append(".reg .u32 " + ", ".join(prog.tmp) + ";")
    for op, dst, srcs in prog.ins:
        if op == "selp.ne0":
            lines.append("{ .reg .pred p; setp.ne.u32 p, %s, 0; selp.b32 %s, %s, %s, p; }" % (o(srcs[2]), o(dst), o(srcs[0]), o(srcs[1])))
        else:
            lines.append(f"{PTX_OP[op]} {o(dst)}, " + ", ".join(o(s) for s in srcs) + ";")
    lines.append("}")
    body = "\n".join('        "%s\\n\\t"' % l for l in lines)
    out_c = ", ".join(f'"{c}"({e})' for c, e in cons[:n_out])
    in_c = ", ".join(f'"{c}"({e})' for c, e in cons[n_out:])
    return (f"__device__ __forceinline__ void {name}({sig}) {{\n    asm(\n{body}\n        : {out_c}\n        : {in_c});\n}}\n")


def limbs(x, n):
    return [(x >> (32 * i)) & M32 for i in range(n)]


def from_limbs(l):
    return sum(v << (32 * i) for i, v in enumerate(l))


# ----------------------------------------------------------------------------------------------
# building blocks (operate on lists of register names)
def cmad_n(pr, acc, a, a_off, bi, n, carry_in=False):
    """acc[j], acc[j+1] += a[a_off+j] * bi for j = 0, 2, .., n-2 as one carry chain (carry left in CC)."""
    for j in range(0, n, 2):
        first = (j == 0 and not carry_in)
        pr.emit("mad.lo.cc" if first else "madc.lo.cc", acc[j], a[a_off + j], bi, acc[j])
        pr.emit("madc.hi.cc", acc[j + 1], a[a_off + j], bi, acc[j + 1])


def mul_n(pr, acc, a, a_off, bi, n):
    for j in range(0, n, 2):
        pr.emit("mul.lo", acc[j], a[a_off + j], bi)
        pr.emit("mul.hi", acc[j + 1], a[a_off + j], bi)


def madc_n_rshift(pr, odd, a, a_off, bi, n):
    """odd[j], odd[j+1] = a[a_off+j]*bi + odd[j+2], odd[j+3] (+carry in CC); top pair adds zero."""
    for j in range(0, n - 2, 2):
        pr.emit("madc.lo.cc", odd[j], a[a_off + j], bi, odd[j + 2])
        pr.emit("madc.hi.cc", odd[j + 1], a[a_off + j], bi, odd[j + 3])
    pr.emit("madc.lo.cc", odd[n - 2], a[a_off + n - 2], bi, 0)
    pr.emit("madc.hi", odd[n - 1], a[a_off + n - 2], bi, 0)


def final_sub(pr, x, n, mod, out):
    """out = x - p if x >= p else x   (x < 2p, n limbs)."""
    d = [pr.t() for _ in range(n)]
    brw = pr.t()
    for i in range(n):
        pr.emit("sub.cc" if i == 0 else "subc.cc", d[i], x[i], mod[i])
    pr.emit("subc", brw, 0, 0)          # 0xffffffff if x < p (borrow) else 0
    for i in range(n):
        pr.emit("selp.ne0", out[i], x[i], d[i], brw)


def gen_mont_mul(n, p, sqr=False):
    """out = a*b*R^-1 mod p, canonical.  p < 2^(32n-1), a, b < p."""
    pr = Prog()
    a = [f"a{i}" for i in range(n)]
    b = a if sqr else [f"b{i}" for i in range(n)]
    out = [f"r{i}" for i in range(n)]
    mod = limbs(p, n)
    inv = (-pow(p, -1, 1 << 32)) & M32
    even = [pr.t() for _ in range(n)]
    odd = [pr.t() for _ in range(n)]
    mi = pr.t()
    E, O = even, odd
    for i in range(n):
        bi = b[i]
        if i == 0:
            mul_n(pr, O, a, 1, bi, n)
            mul_n(pr, E, a, 0, bi, n)
        else:
            pr.emit("add.cc", E[0], E[0], O[1])
            madc_n_rshift(pr, O, a, 1, bi, n)
            cmad_n(pr, E, a, 0, bi, n)
            pr.emit("addc", O[n - 1], O[n - 1], 0)
        if inv == M32:
            pr.emit("sub", mi, 0, E[0])
        else:
            pr.emit("mul.lo", mi, E[0], inv)
        cmad_n(pr, O, mod, 1, mi, n)
        cmad_n(pr, E, mod, 0, mi, n)
        pr.emit("addc", O[n - 1], O[n - 1], 0)
        E, O = O, E
    # merge: value = E + (O >> 32)  (O[0] is zero by construction)
    pr.emit("add.cc", E[0], E[0], O[1])
    for i in range(1, n - 1):
        pr.emit("addc.cc", E[i], E[i], O[i + 1])
    pr.emit("addc", E[n - 1], E[n - 1], 0)
    final_sub(pr, E, n, mod, out)
    return pr, a, b, out


def gen_add(n, p):
    pr = Prog()
    a = [f"a{i}" for i in range(n)]
    b = [f"b{i}" for i in range(n)]
    out = [f"r{i}" for i in range(n)]
    mod = limbs(p, n)
    s = [pr.t() for _ in range(n)]
    for i in range(n):
        pr.emit("add.cc" if i == 0 else ("addc.cc" if i < n - 1 else "addc"), s[i], a[i], b[i])
    final_sub(pr, s, n, mod, out)       # p < 2^(32n-1) so a+b never carries out
    return pr, a, b, out


def gen_sub(n, p):
    pr = Prog()
    a = [f"a{i}" for i in range(n)]
    b = [f"b{i}" for i in range(n)]
    out = [f"r{i}" for i in range(n)]
    mod = limbs(p, n)
    d = [pr.t() for _ in range(n)]
    brw = pr.t()
    m = [pr.t() for _ in range(n)]
    for i in range(n):
        pr.emit("sub.cc" if i == 0 else "subc.cc", d[i], a[i], b[i])
    pr.emit("subc", brw, 0, 0)
    for i in range(n):
        pr.emit("and", m[i], brw, mod[i])
    for i in range(n):
        pr.emit("add.cc" if i == 0 else ("addc.cc" if i < n - 1 else "addc"), out[i], d[i], m[i])
    return pr, a, b, out


def check(n, p, iters=3000, seed=7):
    rng = random.Random(seed)
    R = 1 << (32 * n)
    Rinv = pow(R, -1, p)
    specials = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, R % p, (R * R) % p, (1 << (32 * n - 1)) % p]
    pr_mul, a, b, out = gen_mont_mul(n, p)
    pr_sqr, sa, _, sout = gen_mont_mul(n, p, sqr=True)
    pr_add, aa, ab, aout = gen_add(n, p)
    pr_sub, ua, ub, uout = gen_sub(n, p)
    cases = [(x, y) for x in specials for y in specials]
    cases += [(rng.randrange(p), rng.randrange(p)) for _ in range(iters)]
    for x, y in cases:
        regs = {a[i]: limbs(x, n)[i] for i in range(n)}
        regs.update({b[i]: limbs(y, n)[i] for i in range(n)})
        res = simulate(pr_mul, regs)
        got = from_limbs([res[o] for o in out])
        assert got == x * y * Rinv % p, ("mul", hex(x), hex(y))
        res = simulate(pr_sqr, {sa[i]: limbs(x, n)[i] for i in range(n)})
        assert from_limbs([res[o] for o in sout]) == x * x * Rinv % p, ("sqr", hex(x))
        regs = {aa[i]: limbs(x, n)[i] for i in range(n)}
        regs.update({ab[i]: limbs(y, n)[i] for i in range(n)})
        res = simulate(pr_add, regs)
        assert from_limbs([res[o] for o in aout]) == (x + y) % p, ("add", hex(x), hex(y))
        regs = {ua[i]: limbs(x, n)[i] for i in range(n)}
        regs.update({ub[i]: limbs(y, n)[i] for i in range(n)})
        res = simulate(pr_sub, regs)
        assert from_limbs([res[o] for o in uout]) == (x - y) % p, ("sub", hex(x), hex(y))
    return len(cases)


FR_P = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
FQ_P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB


def emit_field(prefix, n, p):
    s = []
    pr, a, b, out = gen_mont_mul(n, p)
    s.append(to_cuda(pr, f"{prefix}_mul_asm", [(out[i], f"r[{i}]") for i in range(n)],
                     [(a[i], f"a[{i}]") for i in range(n)] + [(b[i], f"b[{i}]") for i in range(n)],
                     f"uint32_t* __restrict__ r, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b"))
    pr, a, _, out = gen_mont_mul(n, p, sqr=True)
    s.append(to_cuda(pr, f"{prefix}_sqr_asm", [(out[i], f"r[{i}]") for i in range(n)],
                     [(a[i], f"a[{i}]") for i in range(n)],
                     f"uint32_t* __restrict__ r, const uint32_t* __restrict__ a"))
    pr, a, b, out = gen_add(n, p)
    s.append(to_cuda(pr, f"{prefix}_add_asm", [(out[i], f"r[{i}]") for i in range(n)],
                     [(a[i], f"a[{i}]") for i in range(n)] + [(b[i], f"b[{i}]") for i in range(n)],
                     f"uint32_t* __restrict__ r, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b"))
    pr, a, b, out = gen_sub(n, p)
    s.append(to_cuda(pr, f"{prefix}_sub_asm", [(out[i], f"r[{i}]") for i in range(n)],
                     [(a[i], f"a[{i}]") for i in range(n)] + [(b[i], f"b[{i}]") for i in range(n)],
                     f"uint32_t* __restrict__ r, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b"))
    return "\n".join(s)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--iters", type=int, default=3000)
    args = ap.parse_args()
    n1 = check(8, FR_P, args.iters)
    n2 = check(12, FQ_P, max(200, args.iters // 4))
    print(f"simulated {n1} Fr and {n2} Fq operand pairs: mul/sqr/add/sub all match big-int arithmetic")
    if args.check:
        return
    here = os.path.dirname(os.path.abspath(__file__))
    dst = os.path.join(here, "..", "gkr-msm_b200", "csrc", "field_gen.cuh")
    with open(dst, "w") as f:
        f.write("// GENERATED by tools/gen_field.py -- do not edit.  Each function is ONE inline-asm block; the carry\n"
                "// chains were executed by the generator's PTX-subset simulator against big-int arithmetic.\n"
                "#pragma once\n#include <cstdint>\n\n")
        f.write("// ---- BLS12-381 Fr: 8 x u32 limbs, R = 2^256 ----\n")
        f.write(emit_field("fr", 8, FR_P))
        f.write("\n// ---- BLS12-381 Fq: 12 x u32 limbs, R = 2^384 ----\n")
        f.write(emit_field("fq", 12, FQ_P))
    print("wrote", os.path.normpath(dst))


if __name__ == "__main__":
    main()
