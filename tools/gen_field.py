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


def special_prime(p):
    """p = 1 mod 2^32 with second limb 0xffffffff (BLS12-381 Fr): the two lowest limbs of m*p need no multiplier."""
    return (p & M32) == 1 and ((p >> 32) & M32) == M32


def reduce_round(pr, E, O, mod, n, inv, mi, special):
    """One Montgomery round on the even/odd accumulators: adds mi*p so that E[0] becomes 0 (mod 2^32)."""
    if inv == M32:
        pr.emit("sub", mi, 0, E[0])
    else:
        pr.emit("mul.lo", mi, E[0], inv)
    if not special:
        cmad_n(pr, O, mod, 1, mi, n)
        cmad_n(pr, E, mod, 0, mi, n)
        pr.emit("addc", O[n - 1], O[n - 1], 0)
        return
    # p0 = 1:          (E1:E0) += mi       -> E0 = 0, carry c = (E0 != 0) into E1
    # p1 = 2^32 - 1:   (O1:O0) += mi*2^32 - mi = (mi - c)*2^32 + E0      (E0 = -mi mod 2^32)
    dummy, c, hm = pr.t(), pr.t(), pr.t()
    pr.emit("add.cc", dummy, E[0], M32)
    pr.emit("addc", c, 0, 0)
    pr.emit("sub", hm, mi, c)
    pr.emit("add.cc", O[0], O[0], E[0])
    pr.emit("addc.cc", O[1], O[1], hm)
    for j in range(2, n, 2):
        pr.emit("madc.lo.cc", O[j], mod[1 + j], mi, O[j])
        pr.emit("madc.hi.cc", O[j + 1], mod[1 + j], mi, O[j + 1])
    pr.emit("mov", E[0], 0)
    pr.emit("add.cc", E[1], E[1], c)
    for j in range(2, n, 2):
        pr.emit("madc.lo.cc", E[j], mod[j], mi, E[j])
        pr.emit("madc.hi.cc", E[j + 1], mod[j], mi, E[j + 1])
    pr.emit("addc", O[n - 1], O[n - 1], 0)


def gen_mont_mul(n, p, sqr=False, rounds=None, init=False, special=None):
    """out = (c + a*b) * 2^(-32*rounds) mod p, canonical.  p < 2^(32n-1); a < p; b has `rounds` limbs (default n:
    the plain Montgomery product a*b*R^-1); c (only when init) is an n-limb addend < p -- with rounds = 4 this is the
    sumcheck fold by a 128-bit challenge, e0 + t*(e1 - e0), leaving a known 2^-128 scale that the host tracks."""
    pr = Prog()
    rounds = n if rounds is None else rounds
    special = special_prime(p) if special is None else special
    a = [f"a{i}" for i in range(n)]
    b = a if sqr else [f"b{i}" for i in range(rounds)]
    cin = [f"c{i}" for i in range(n)] if init else None
    out = [f"r{i}" for i in range(n)]
    mod = limbs(p, n)
    inv = (-pow(p, -1, 1 << 32)) & M32
    even = [pr.t() for _ in range(n)]
    odd = [pr.t() for _ in range(n)]
    mi = pr.t()
    E, O = even, odd
    if init:
        # state convention: value limb k = E[k] + O[k+1], O[0] = 0
        for k in range(n - 1):
            pr.emit("mov", E[k], 0)
            pr.emit("mov", O[k + 1], cin[k])
        pr.emit("mov", E[n - 1], cin[n - 1])
        pr.emit("mov", O[0], 0)
    for i in range(rounds):
        bi = b[i]
        if i == 0 and not init:
            mul_n(pr, O, a, 1, bi, n)
            mul_n(pr, E, a, 0, bi, n)
        else:
            pr.emit("add.cc", E[0], E[0], O[1])
            madc_n_rshift(pr, O, a, 1, bi, n)
            cmad_n(pr, E, a, 0, bi, n)
            pr.emit("addc", O[n - 1], O[n - 1], 0)
        reduce_round(pr, E, O, mod, n, inv, mi, special)
        E, O = O, E
    # merge: value = E + (O >> 32)  (O[0] is zero by construction)
    pr.emit("add.cc", E[0], E[0], O[1])
    for i in range(1, n - 1):
        pr.emit("addc.cc", E[i], E[i], O[i + 1])
    pr.emit("addc", E[n - 1], E[n - 1], 0)
    final_sub(pr, E, n, mod, out)
    if init:
        return pr, a, b, cin, out
    return pr, a, b, out


def gen_mul_wide_acc(n, extra=1):
    """acc (2n + extra limbs, in/out) += a * b  for arbitrary n-limb a, b -- NO reduction.  The partial products
    go to two temporaries (even- and odd-aligned 64-bit columns, so every mad.lo/madc.hi pair is one IMAD.WIDE)
    which are then added to the accumulator with two carry chains.  Montgomery reduction is linear, so a thread
    sums many products this way and reduces once (gen_* callers: the last multiplication of a sumcheck gate)."""
    pr = Prog()
    a = [f"a{i}" for i in range(n)]
    b = [f"b{i}" for i in range(n)]
    acc = [f"s{i}" for i in range(2 * n + extra)]
    P = [pr.t() for _ in range(2 * n)]        # P[k] <-> position k     (pairs (0,1),(2,3),..)
    Q = [pr.t() for _ in range(2 * n)]        # Q[k] <-> position k + 1 (pairs (1,2),(3,4),..)
    touched_P, touched_Q = set(), set()

    def chain(arr, touched, base, js, bi):
        """arr[base + j - js[0] ...] pairs += a[j] * bi for j in js (step 2), one carry chain."""
        first = True
        k = base
        for j in js:
            for half, op in ((0, "lo"), (1, "hi")):
                idx = k + half
                fresh = idx not in touched
                src = 0 if fresh else arr[idx]
                if first:
                    pr.emit(f"mad.{op}.cc", arr[idx], a[j], bi, src)
                    first = False
                else:
                    pr.emit(f"madc.{op}.cc", arr[idx], a[j], bi, src)
                touched.add(idx)
            k += 2
        if k < 2 * n:
            fresh = k not in touched
            pr.emit("addc", arr[k], 0 if fresh else arr[k], 0)
            touched.add(k)

    for i in range(n):
        ev = list(range(0, n, 2))
        od = list(range(1, n, 2))
        if i % 2 == 0:
            chain(P, touched_P, i, ev, b[i])            # positions i + j, even
            chain(Q, touched_Q, i, od, b[i])            # positions i + j (odd) -> Q index i + j - 1 = i + (j-1)
        else:
            chain(Q, touched_Q, i - 1, ev, b[i])        # positions i + j (odd) -> Q index i + j - 1
            chain(P, touched_P, i + 1, od, b[i])        # positions i + j, even
    for k in range(2 * n):
        if k not in touched_P:
            pr.emit("mov", P[k], 0)
        if k not in touched_Q:
            pr.emit("mov", Q[k], 0)
    top = 2 * n + extra
    for k in range(2 * n):
        pr.emit("add.cc" if k == 0 else "addc.cc", acc[k], acc[k], P[k])
    for k in range(2 * n, top):
        pr.emit("addc.cc" if k < top - 1 else "addc", acc[k], acc[k], 0)
    for k in range(1, 2 * n):
        pr.emit("add.cc" if k == 1 else "addc.cc", acc[k], acc[k], Q[k - 1])
    for k in range(2 * n, top):
        pr.emit("addc.cc" if k < top - 1 else "addc", acc[k], acc[k], Q[2 * n - 1] if k == 2 * n else 0)
    return pr, a, b, acc


def gen_reduce_once(n, p, times=2):
    """out = x mod p for an n-limb x < (times + 1) * p: `times` conditional subtractions."""
    pr = Prog()
    x = [f"a{i}" for i in range(n)]
    out = [f"r{i}" for i in range(n)]
    mod = limbs(p, n)
    cur = x
    for t in range(times):
        nxt = out if t == times - 1 else [pr.t() for _ in range(n)]
        final_sub(pr, cur, n, mod, nxt)
        cur = nxt
    return pr, x, out


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
    # operands with long runs of zero / all-one limbs exercise the carry paths
    for _ in range(iters // 4):
        x = from_limbs([rng.choice([0, M32, 1, rng.getrandbits(32)]) for _ in range(n)]) % p
        y = from_limbs([rng.choice([0, M32, 1, rng.getrandbits(32)]) for _ in range(n)]) % p
        cases.append((x, y))
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


def check_wide(n, p, iters=2000, seed=11, fold_limbs=4, extra=1):
    """fold by a short challenge, unreduced multiply-accumulate, conditional-subtraction reduce."""
    rng = random.Random(seed)
    full = (1 << (32 * n)) - 1
    pr_fold, fa, fb, fc, fout = gen_mont_mul(n, p, rounds=fold_limbs, init=True)
    inv_s = pow(1 << (32 * fold_limbs), -1, p)
    tmax = (1 << (32 * fold_limbs)) - 1
    edge = [0, 1, p - 1, p - 2, (p - 1) // 2]
    cases = [(c, d, t) for c in edge for d in edge for t in (0, 1, tmax, tmax - 1, 1 << (32 * fold_limbs - 1))]
    cases += [(rng.randrange(p), rng.randrange(p), rng.getrandbits(32 * fold_limbs)) for _ in range(iters)]
    for c, d, t in cases:
        regs = {fa[i]: limbs(d, n)[i] for i in range(n)}
        regs.update({fb[i]: limbs(t, fold_limbs)[i] for i in range(fold_limbs)})
        regs.update({fc[i]: limbs(c, n)[i] for i in range(n)})
        res = simulate(pr_fold, regs)
        assert from_limbs([res[o] for o in fout]) == (c + t * d) * inv_s % p, ("fold", hex(c), hex(d), hex(t))
    pr_mac, ma, mb, macc = gen_mul_wide_acc(n, extra)
    nacc = 2 * n + extra
    ops = [0, 1, full, full - 1, p - 1, 1 << (32 * n - 1)]
    mcases = [(x, y, 0) for x in ops for y in ops] + [(full, full, (1 << (32 * nacc)) - 1 - full * full)]
    mcases += [(rng.getrandbits(32 * n), rng.getrandbits(32 * n), rng.getrandbits(32 * nacc - 2)) for _ in range(iters)]
    for _ in range(iters // 4):
        x = from_limbs([rng.choice([0, M32, 1, rng.getrandbits(32)]) for _ in range(n)])
        y = from_limbs([rng.choice([0, M32, 1, rng.getrandbits(32)]) for _ in range(n)])
        mcases.append((x, y, rng.getrandbits(32 * nacc - 2)))
    for x, y, s0 in mcases:
        regs = {ma[i]: limbs(x, n)[i] for i in range(n)}
        regs.update({mb[i]: limbs(y, n)[i] for i in range(n)})
        regs.update({macc[i]: limbs(s0, nacc)[i] for i in range(nacc)})
        res = simulate(pr_mac, regs)
        assert from_limbs([res[o] for o in macc]) == s0 + x * y, ("mac", hex(x), hex(y), hex(s0))
    pr_red, rx, rout = gen_reduce_once(n, p, times=2)
    for x in [0, 1, p - 1, p, p + 1, 2 * p - 1, 2 * p, 2 * p + 1, min(full, 3 * p - 1)] + [rng.getrandbits(32 * n) for _ in range(iters)]:
        if x >= 3 * p:
            continue
        res = simulate(pr_red, {rx[i]: limbs(x, n)[i] for i in range(n)})
        assert from_limbs([res[o] for o in rout]) == x % p, ("reduce", hex(x))
    return len(cases) + len(mcases)


FR_P = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
FQ_P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB


def emit_field(prefix, n, p, wide=False):
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
    if wide:
        # r = (c + a * t) * 2^-128 mod p, t = 4 limbs: the sumcheck fold by a 128-bit challenge
        pr, a, b, c, out = gen_mont_mul(n, p, rounds=4, init=True)
        s.append(to_cuda(pr, f"{prefix}_fold128_asm", [(out[i], f"r[{i}]") for i in range(n)],
                         [(c[i], f"c[{i}]") for i in range(n)] + [(a[i], f"a[{i}]") for i in range(n)] + [(b[i], f"t[{i}]") for i in range(4)],
                         f"uint32_t* __restrict__ r, const uint32_t* __restrict__ c, const uint32_t* __restrict__ a, const uint32_t* __restrict__ t"))
        # s (2n+1 limbs) += a * b, unreduced
        pr, a, b, acc = gen_mul_wide_acc(n, 1)
        s.append(to_cuda(pr, f"{prefix}_mac_wide_asm", [],
                         [(a[i], f"a[{i}]") for i in range(n)] + [(b[i], f"b[{i}]") for i in range(n)],
                         f"uint32_t* __restrict__ s, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b",
                         inout=[(acc[i], f"s[{i}]") for i in range(2 * n + 1)]))
        pr, x, out = gen_reduce_once(n, p, times=2)
        s.append(to_cuda(pr, f"{prefix}_reduce2_asm", [(out[i], f"r[{i}]") for i in range(n)],
                         [(x[i], f"a[{i}]") for i in range(n)],
                         f"uint32_t* __restrict__ r, const uint32_t* __restrict__ a"))
    return "\n".join(s)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--iters", type=int, default=3000)
    args = ap.parse_args()
    n1 = check(8, FR_P, args.iters)
    n2 = check(12, FQ_P, max(200, args.iters // 4))
    n3 = check_wide(8, FR_P, args.iters)
    print(f"simulated {n1} Fr and {n2} Fq operand pairs: mul/sqr/add/sub all match big-int arithmetic")
    print(f"simulated {n3} Fr fold128 / wide multiply-accumulate / reduce cases")
    if args.check:
        return
    here = os.path.dirname(os.path.abspath(__file__))
    dst = os.path.join(here, "..", "gkr-msm_b200", "csrc", "field_gen.cuh")
    with open(dst, "w") as f:
        f.write("// GENERATED by tools/gen_field.py -- do not edit.  Each function is ONE inline-asm block; the carry\n"
                "// chains were executed by the generator's PTX-subset simulator against big-int arithmetic.\n"
                "#pragma once\n#include <cstdint>\n\n")
        f.write("// ---- BLS12-381 Fr: 8 x u32 limbs, R = 2^256 ----\n")
        f.write(emit_field("fr", 8, FR_P, wide=True))
        f.write("\n// ---- BLS12-381 Fq: 12 x u32 limbs, R = 2^384 ----\n")
        f.write(emit_field("fq", 12, FQ_P))
    print("wrote", os.path.normpath(dst))


if __name__ == "__main__":
    main()
