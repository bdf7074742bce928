"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/c/libgkr_oracle.so (plain-C restatement).
Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "c", "libgkr_oracle.so")
_lib = None
_vp = C.c_void_p


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB


def use_native() -> bool:
    """bench.py's CPU legs: switch to a -march=native build made on THIS machine (the reference's README builds with
    target-cpu=native); keeps the portable prebuilt library when gcc is missing or the library is already loaded."""
    global LIB
    if _lib is not None:
        return LIB.endswith(".native.so")
    native = os.path.join(_HERE, "c", "libgkr_oracle.native.so")
    try:
        if not os.path.exists(native):
            subprocess.check_call(["make", "-C", _HERE, "-s", "c/libgkr_oracle.native.so"])
        LIB = native
        return True
    except Exception:
        return False


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.oracle_dense_sumcheck.restype = C.c_int
        _lib.oracle_dense_sumcheck.argtypes = [C.c_int, C.c_int, C.c_uint32, _vp, C.c_int, C.c_uint32, C.POINTER(_vp), _vp, _vp,
                                               C.c_uint32, _vp, _vp]
        _lib.oracle_eq_table.argtypes = [_vp, C.c_uint32, _vp, _vp]
        _lib.oracle_synth_table.argtypes = [C.c_uint64, C.c_uint64, _vp]
        _lib.oracle_gate_sum.argtypes = [C.c_int, C.c_int, C.c_uint32, _vp, C.c_int, C.c_uint64, C.POINTER(_vp), _vp]
        _lib.oracle_fr_mul.argtypes = [_vp, _vp, _vp]
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(_vp)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def fr_mul(a, b):
    a = np.ascontiguousarray(a, np.uint64)
    b = np.ascontiguousarray(b, np.uint64)
    out = np.zeros(4, np.uint64)
    lib().oracle_fr_mul(_p(a), _p(b), _p(out))
    return out


def synth_table(seed: int, n: int) -> np.ndarray:
    out = np.empty((n, 4), np.uint64)
    lib().oracle_synth_table(seed & 0xFFFFFFFFFFFFFFFF, n, _p(out))
    return out


def eq_table(point: np.ndarray, mult: np.ndarray) -> np.ndarray:
    point = np.ascontiguousarray(point, np.uint64).reshape(-1, 4)
    mult = np.ascontiguousarray(mult, np.uint64).reshape(4)
    n = point.shape[0]
    out = np.empty((1 << n, 4), np.uint64)
    lib().oracle_eq_table(_p(point), n, _p(mult), _p(out))
    return out


def gate_sum(so_kind, gate, tables, param=0, consts=None) -> np.ndarray:
    tabs = [np.ascontiguousarray(t, np.uint64).reshape(-1, 4) for t in tables]
    c = np.ascontiguousarray(consts, np.uint64).reshape(-1, 4) if consts is not None else np.zeros((16, 4), np.uint64)
    arr = (_vp * len(tabs))(*[t.ctypes.data for t in tabs])
    out = np.zeros(4, np.uint64)
    rc = lib().oracle_gate_sum(so_kind, gate, param, _p(c), len(tabs), tabs[0].shape[0], arr, _p(out))
    assert rc == 0, rc
    return out


def dense_sumcheck(so_kind, gate, tables, nv, claim, challenges, param=0, consts=None, rounds=None):
    """returns (evals [rounds, deg+1, 4], final_evals [P, 4] or None)"""
    tabs = [np.ascontiguousarray(t, np.uint64).reshape(-1, 4) for t in tables]
    c = np.ascontiguousarray(consts, np.uint64).reshape(-1, 4) if consts is not None else np.zeros((16, 4), np.uint64)
    if c.shape[0] < 16:
        c = np.concatenate([c, np.zeros((16 - c.shape[0], 4), np.uint64)])
    ch = np.ascontiguousarray(challenges, np.uint64).reshape(-1, 4)
    cl = np.ascontiguousarray(claim, np.uint64).reshape(4)
    rounds = nv if rounds is None else rounds
    deg = 2 if (so_kind == 0 and gate == 11) else 3
    ev = np.zeros((max(rounds, 1), deg + 1, 4), np.uint64)
    fe = np.zeros((len(tabs), 4), np.uint64)
    arr = (_vp * len(tabs))(*[t.ctypes.data for t in tabs])
    rc = lib().oracle_dense_sumcheck(so_kind, gate, param, _p(c), len(tabs), nv, arr, _p(cl), _p(ch), rounds, _p(ev), _p(fe))
    assert rc == 0, rc
    return ev[:rounds], (fe if rounds >= nv else None)
