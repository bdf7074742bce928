"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

The closed set of polynomial gates (`AlgFn` / `AlgFnSO`) that the reference instantiates on the hot
path, restated over python ints mod r.  Every gate is a small object with
`n_ins, n_outs, deg, exec(args) -> list`.

Reference:
  src/cleanup/utils/algfn.rs:11-34        trait AlgFnSO / AlgFn
  src/cleanup/utils/algfn.rs:130-291      IdAlgFn, RepeatedAlgFn, StackedAlgFn, BitCheckFn
  src/cleanup/utils/twisted_edwards_ops.rs:10-81   the 7 twisted-Edwards addition gates
  src/utils.rs:32-49                      mul_by_a (a = -5), mul_by_d
  src/cleanup/protocols/pushforward/pushforward.rs:28-50, 255-281   Prod3Fn, AddInversesFn
  src/cleanup/protocols/pushforward/logup_mainphase.rs:32-61        LogupLayerFn
  src/cleanup/protocols/multiopen_reduction.rs:13-41                FoldedProdAlgFn
  src/cleanup/protocols/sumcheck.rs:706-741, 802-829                GammaWrapper, EqWrapper
"""
from __future__ import annotations

from .field import P, TE_D

# numeric ids shared with include/gkr_msm_b200.h (enum gkr_gate_id)
GATE_AFF_L1 = 0
GATE_AFF_L2 = 1
GATE_AFF_L3 = 2
GATE_PRJ_L1 = 3
GATE_PRJ_L2 = 4
GATE_PRJ_L3 = 5
GATE_TRI_L1 = 6
GATE_BITCHECK = 7
GATE_LOGUP_LAYER = 8
GATE_ADD_INVERSES = 9
GATE_PROD3 = 10
GATE_FOLDED_PROD = 11
GATE_ID = 12
GATE_AFF_L1_BITCHECK2 = 13  # Stacked(affine_l1, Repeated(BitCheck, 2))  (bintree_add.rs:259-273)


def mul_by_a(x):  # src/utils.rs:40-43 : -(4x + x)
    return (-(5 * x)) % P


def mul_by_d(x):  # src/utils.rs:46-48
    return (x * TE_D) % P


class Gate:
    gate_id = -1
    n_ins = 0
    n_outs = 0
    deg = 0

    def exec(self, a):
        raise NotImplementedError

    def __call__(self, a):
        return self.exec(a)


class AffL1(Gate):  # twisted_edwards_ops.rs:10-14
    gate_id, n_ins, n_outs, deg = GATE_AFF_L1, 4, 3, 2

    def exec(self, a):
        x1, y1, x2, y2 = a[0], a[1], a[2], a[3]
        return [x1 * y2 % P, x2 * y1 % P, (y1 * y2 - mul_by_a(x1 * x2 % P)) % P]


class AffL2(Gate):  # :16-20
    gate_id, n_ins, n_outs, deg = GATE_AFF_L2, 3, 3, 2

    def exec(self, a):
        return [(a[0] + a[1]) % P, a[2] % P, a[0] * a[1] % P]


class AffL3(Gate):  # :22-29
    gate_id, n_ins, n_outs, deg = GATE_AFF_L3, 3, 3, 2

    def exec(self, a):
        x, y, xy = a[0], a[1], a[2]
        dxy = mul_by_d(xy)
        m, p = (1 - dxy) % P, (1 + dxy) % P
        return [m * x % P, p * y % P, m * p % P]


class PrjL1(Gate):  # :31-40
    gate_id, n_ins, n_outs, deg = GATE_PRJ_L1, 6, 4, 2

    def exec(self, a):
        x1, y1, z1, x2, y2, z2 = a[0], a[1], a[2], a[3], a[4], a[5]
        return [x1 * y2 % P, x2 * y1 % P, (y1 * y2 - mul_by_a(x1 * x2 % P)) % P, z1 * z2 % P]


class PrjL2(Gate):  # :43-52
    gate_id, n_ins, n_outs, deg = GATE_PRJ_L2, 4, 4, 2

    def exec(self, a):
        x1y2, x2y1, c, zz = a[0], a[1], a[2], a[3]
        return [(x1y2 + x2y1) * zz % P, c * zz % P, zz * zz % P, x1y2 * x2y1 % P]


class PrjL3(Gate):  # :54-65
    gate_id, n_ins, n_outs, deg = GATE_PRJ_L3, 4, 3, 2

    def exec(self, a):
        x, y, z2, xy = a[0], a[1], a[2], a[3]
        dxy = mul_by_d(xy)
        m, p = (z2 - dxy) % P, (z2 + dxy) % P
        return [m * x % P, p * y % P, m * p % P]


class TriL1(Gate):  # :67-80  three projective L1 on (a,c), (b,d), (c,d)
    gate_id, n_ins, n_outs, deg = GATE_TRI_L1, 12, 12, 2

    def exec(self, p):
        a, b, c, d = p[0:3], p[3:6], p[6:9], p[9:12]
        l1 = PrjL1()
        return l1.exec(list(a) + list(c)) + l1.exec(list(b) + list(d)) + l1.exec(list(c) + list(d))


class BitCheck(Gate):  # algfn.rs:262-291
    gate_id, n_ins, n_outs, deg = GATE_BITCHECK, 1, 1, 2

    def exec(self, a):
        return [(a[0] * a[0] - a[0]) % P]


class Id(Gate):  # algfn.rs:130-164
    gate_id = GATE_ID

    def __init__(self, n):
        self.n_ins = self.n_outs = n
        self.deg = 1

    def exec(self, a):
        return [a[i] % P for i in range(self.n_ins)]


class Repeated(Gate):  # algfn.rs:187-225
    def __init__(self, f, count):
        self.f, self.count = f, count
        self.n_ins, self.n_outs, self.deg = f.n_ins * count, f.n_outs * count, f.deg

    def exec(self, a):
        out = []
        for i in range(self.count):
            out += self.f.exec(a[i * self.f.n_ins:(i + 1) * self.f.n_ins])
        return out


class Stacked(Gate):  # algfn.rs:227-259
    def __init__(self, f1, f2):
        self.f1, self.f2 = f1, f2
        self.n_ins, self.n_outs = f1.n_ins + f2.n_ins, f1.n_outs + f2.n_outs
        self.deg = max(f1.deg, f2.deg)

    def exec(self, a):
        return self.f1.exec(a[:self.f1.n_ins]) + self.f2.exec(a[self.f1.n_ins:self.n_ins])


class AffL1BitCheck2(Stacked):
    gate_id = GATE_AFF_L1_BITCHECK2

    def __init__(self):
        super().__init__(AffL1(), Repeated(BitCheck(), 2))


class LogupLayer(Gate):  # logup_mainphase.rs:42-61
    gate_id, n_ins, n_outs, deg = GATE_LOGUP_LAYER, 4, 2, 2

    def exec(self, a):
        return [(a[0] * a[3] + a[1] * a[2]) % P, a[1] * a[3] % P]


class AddInverses(Gate):  # pushforward.rs:266-281
    gate_id, n_ins, n_outs, deg = GATE_ADD_INVERSES, 2, 2, 2

    def exec(self, a):
        return [(a[0] + a[1]) % P, a[0] * a[1] % P]


# ---- single-output gates (AlgFnSO) ---------------------------------------------------------
class GateSO:
    n_ins = 0
    deg = 0

    def exec(self, a):
        raise NotImplementedError


class Prod3(GateSO):  # pushforward.rs:38-50
    gate_id, n_ins, deg = GATE_PROD3, 3, 3

    def exec(self, a):
        return a[0] * a[1] * a[2] % P


class FoldedProd(GateSO):  # multiopen_reduction.rs:13-41
    gate_id = GATE_FOLDED_PROD

    def __init__(self, gamma, nargs):
        from .sumcheck import make_gamma_pows
        self.gammas = make_gamma_pows(gamma, nargs)
        self.nargs, self.n_ins, self.deg = nargs, 2 * nargs, 2

    def exec(self, a):
        return sum(a[i] * a[i + self.nargs] % P * self.gammas[i] for i in range(self.nargs)) % P


class GammaWrapper(GateSO):  # sumcheck.rs:706-741  (gamma_pows = [g, g^2, ...], out0 + sum out_i g^i)
    def __init__(self, f, gamma):
        assert f.n_outs > 1
        self.f = f
        self.gamma_pows = [gamma % P]
        for _ in range(f.n_outs - 2):
            self.gamma_pows.append(gamma * self.gamma_pows[-1] % P)
        self.n_ins, self.deg = f.n_ins, f.deg

    def exec(self, a):
        out = self.f.exec(a)
        ret = out[0]
        for o, g in zip(out[1:], self.gamma_pows):
            ret += o * g
        return ret % P


class EqWrapper(GateSO):  # sumcheck.rs:802-829
    def __init__(self, f):
        self.f = f
        self.n_ins, self.deg = f.n_ins + 1, f.deg + 1

    def exec(self, a):
        return self.f.exec(a) * a[self.f.n_ins] % P


MO_GATES = {
    GATE_AFF_L1: AffL1, GATE_AFF_L2: AffL2, GATE_AFF_L3: AffL3,
    GATE_PRJ_L1: PrjL1, GATE_PRJ_L2: PrjL2, GATE_PRJ_L3: PrjL3,
    GATE_TRI_L1: TriL1, GATE_BITCHECK: BitCheck, GATE_LOGUP_LAYER: LogupLayer,
    GATE_ADD_INVERSES: AddInverses, GATE_AFF_L1_BITCHECK2: AffL1BitCheck2,
}


def gate_by_id(gid: int, param: int = 0):
    if gid == GATE_ID:
        return Id(param)
    return MO_GATES[gid]()
