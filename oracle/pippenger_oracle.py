"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/c/libpippenger_oracle.so, the C++ / OpenMP restatement of the whole
`examples/pippenger` prover (benchutils::run_pippenger, src/cleanup/protocols/pippenger.rs:499-559).
Importable only from tests/, tests/golden/*.py, __graft_entry__.smoke() and bench.py's CPU legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "c", "libpippenger_oracle.so")
LIB_NATIVE = os.path.join(_HERE, "c", "libpippenger_oracle.native.so")
_libs = {}
_vp = C.c_void_p


def build(native: bool = False) -> str:
    subprocess.check_call(["make", "-C", _HERE, "-s"] + (["native"] if native else []))
    return LIB_NATIVE if native else LIB


def lib(native: bool = False):
    """native=True: the -march=native build made on THIS machine (falls back to the portable prebuilt one)"""
    path = LIB
    if native:
        try:
            if not os.path.exists(LIB_NATIVE):
                build(native=True)
            path = LIB_NATIVE
        except Exception:
            path = LIB
    if path not in _libs:
        if not os.path.exists(path):
            build()
        l = C.CDLL(path)
        l.po_last_error.restype = C.c_char_p
        l.po_num_threads.restype = C.c_int
        l.po_key_create.restype = _vp
        l.po_key_create.argtypes = [_vp, _vp, C.c_uint32, _vp]
        l.po_key_destroy.argtypes = [_vp]
        l.po_key_point.argtypes = [_vp, C.c_uint64, _vp]
        l.po_key_commit.argtypes = [_vp, _vp, C.c_uint64, _vp]
        l.po_te_arithmetic_progression.restype = C.c_int
        l.po_te_arithmetic_progression.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, _vp]
        l.po_run_pippenger.restype = C.c_int
        l.po_run_pippenger.argtypes = [_vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _vp, C.c_uint64, _vp, _vp, C.c_uint64, _vp,
                                       _vp, _vp, _vp]
        l._path = path
        _libs[path] = l
    return _libs[path]


def _p(a):
    return a.ctypes.data_as(_vp)


class OracleError(RuntimeError):
    pass


class Key:
    """KzgProvingKey::mock_setup(tau, g0, _, 2 * 2^num_vars - 1) + KnucklesProvingKey::new(.., num_vars, k)
    (kzg.rs:84-97, knuckles.rs:65-81).  tau_limbs / k_limbs: 4 Montgomery u64 limbs; g0_limbs: 12 (affine x, y)."""

    def __init__(self, tau_limbs, g0_limbs, num_vars: int, k_limbs, native: bool = False):
        self.l = lib(native)
        tau = np.ascontiguousarray(tau_limbs, np.uint64).reshape(4)
        g0 = np.ascontiguousarray(g0_limbs, np.uint64).reshape(12)
        k = np.ascontiguousarray(k_limbs, np.uint64).reshape(4)
        self.h = self.l.po_key_create(_p(tau), _p(g0), num_vars, _p(k))
        if not self.h:
            raise OracleError(self.l.po_last_error().decode())
        self.num_vars = num_vars

    def point(self, i: int) -> np.ndarray:
        out = np.zeros(12, np.uint64)
        if self.l.po_key_point(self.h, i, _p(out)):
            raise OracleError("po_key_point failed")
        return out

    def commit_bytes(self, poly_limbs) -> bytes:
        a = np.ascontiguousarray(poly_limbs, np.uint64).reshape(-1, 4)
        out = np.zeros(48, np.uint8)
        if self.l.po_key_commit(self.h, _p(a), a.shape[0], _p(out)):
            raise OracleError(self.l.po_last_error().decode())
        return out.tobytes()

    def close(self):
        if self.h:
            self.l.po_key_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_pippenger(key: Key, points_xy, coefs_u64, r_limbs, d_logsize: int, x_logsize: int, num_bits: int, clm: int):
    """-> dict(proof=bytes, dense_output=(n_tables, 2^y_logsize, 4) u64, claim_evs=(n_tables, 4) u64, pair=(2, 12) u64,
    seconds={phase: s}).  points_xy: (2, n, 4) Montgomery limbs; coefs_u64: (n, 4) plain integers; r_limbs: (y_logsize, 4)."""
    n = 1 << x_logsize
    pts = np.ascontiguousarray(points_xy, np.uint64).reshape(2, n, 4)
    cf = np.ascontiguousarray(coefs_u64, np.uint64).reshape(n, 4)
    y_size = (num_bits + d_logsize - 1) // d_logsize
    y_logsize = max(0, (y_size - 1).bit_length())
    r = np.ascontiguousarray(r_limbs, np.uint64).reshape(max(y_logsize, 0), 4) if y_logsize else np.zeros((1, 4), np.uint64)
    cap = 4 << 20
    proof = np.zeros(cap, np.uint8)
    plen = C.c_uint64(0)
    n_tables = 3 * (d_logsize + 1)
    dense = np.zeros((n_tables, 1 << y_logsize, 4), np.uint64)
    nt = C.c_uint64(0)
    evs = np.zeros((n_tables, 4), np.uint64)
    pair = np.zeros((2, 12), np.uint64)
    secs = np.zeros(6, np.float64)
    rc = key.l.po_run_pippenger(key.h, _p(pts), _p(cf), _p(r), d_logsize, x_logsize, num_bits, clm, _p(proof), cap, C.byref(plen), _p(dense),
                                dense.shape[0] * dense.shape[1], C.byref(nt), _p(evs), _p(pair), _p(secs))
    if rc:
        raise OracleError(key.l.po_last_error().decode())
    assert nt.value == n_tables
    names = ["witness_and_phase1_commit", "ending_gkr", "second_phase", "pushforward", "open", "total"]
    return dict(proof=proof[:plen.value].tobytes(), dense_output=dense, claim_evs=evs, pair=pair, seconds=dict(zip(names, secs.tolist())))


def te_arithmetic_progression(k0: int, step: int, n: int) -> np.ndarray:
    """(2, n, 4) Montgomery limbs of the Bandersnatch points (k0 + i * step) G -- the synthetic inputs of the large workloads"""
    out = np.zeros((2, n, 4), np.uint64)
    if lib().po_te_arithmetic_progression(k0, step, n, _p(out)):
        raise OracleError(lib().po_last_error().decode())
    return out


def num_threads(native: bool = False) -> int:
    return int(lib(native).po_num_threads())
