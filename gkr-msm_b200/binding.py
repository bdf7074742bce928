"""ctypes binding of the C ABI (include/gkr_msm_b200.h) plus thin python mirrors of the reference's
host-side objects.  There is NO CPU fallback here: if the shared library is missing or no CUDA device
is present, construction fails loudly.

Field elements at this level are numpy uint64 arrays of shape (..., 4): canonical Montgomery limbs,
the reference's own `Vec<Fr>` memory layout.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgkr_msm_b200.so")

GATE_AFF_L1, GATE_AFF_L2, GATE_AFF_L3 = 0, 1, 2
GATE_PRJ_L1, GATE_PRJ_L2, GATE_PRJ_L3 = 3, 4, 5
GATE_TRI_L1, GATE_BITCHECK, GATE_LOGUP_LAYER, GATE_ADD_INVERSES = 6, 7, 8, 9
GATE_PROD3, GATE_FOLDED_PROD, GATE_ID, GATE_AFF_L1_BITCHECK2 = 10, 11, 12, 13
SO_PLAIN, SO_EQ_GAMMA = 0, 1

GKR_OK, GKR_ERR_CUDA, GKR_ERR_ARG, GKR_ERR_PROTOCOL, GKR_ERR_UNSUPPORTED = 0, -1, -2, -3, -4


class GkrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gkr_msm_b200 status {code}: {msg}")
        self.code = code


def build_library(force: bool = False) -> str:
    """Compile every CUDA source for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    args = ["make", "-C", _HERE, "-j8"]
    if force:
        subprocess.check_call(["make", "-C", _HERE, "clean"])
    subprocess.check_call(args)
    return LIB_PATH


_lib = None

_u64p = C.POINTER(C.c_uint64)
_u8p = C.POINTER(C.c_uint8)
_vp = C.c_void_p


def _sig(lib):
    def f(name, res, *args):
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = list(args)
    f("gkr_version", C.c_int)
    f("gkr_ctx_create", C.c_int, C.c_int, C.POINTER(_vp))
    f("gkr_ctx_destroy", None, _vp)
    f("gkr_last_error", C.c_char_p, _vp)
    f("gkr_ctx_sync", C.c_int, _vp)
    f("gkr_ctx_launch_count", C.c_uint64, _vp)
    f("gkr_ctx_stream", _vp, _vp)
    f("gkr_ctx_timing_enable", C.c_int, _vp, C.c_int)
    f("gkr_ctx_set_fast_fold", C.c_int, _vp, C.c_int)
    f("gkr_ctx_set_tuning", C.c_int, _vp, C.c_char_p, C.c_longlong)
    f("gkr_ctx_host_stats", C.c_int, _vp, _vp, C.c_int)
    f("gkr_ctx_timing_read", C.c_int, _vp, _vp, _vp, _vp, C.c_int)
    f("gkr_table_upload", C.c_int, _vp, _vp, C.c_uint64, C.POINTER(_vp))
    f("gkr_table_download", C.c_int, _vp, _vp, _vp)
    f("gkr_table_alloc", C.c_int, _vp, C.c_uint64, C.POINTER(_vp))
    f("gkr_table_synth", C.c_int, _vp, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(_vp))
    f("gkr_table_len", C.c_uint64, _vp)
    f("gkr_table_device_ptr", _vp, _vp)
    f("gkr_table_free", None, _vp)
    f("gkr_eq_table", C.c_int, _vp, _vp, C.c_uint32, _vp, C.POINTER(_vp))
    f("gkr_dense_gate_sum", C.c_int, _vp, C.c_int, C.c_int, C.c_uint32, _vp, C.c_uint32, C.POINTER(_vp), C.c_uint32, _vp)
    f("gkr_so_create_dense", C.c_int, _vp, C.c_int, C.c_int, C.c_uint32, _vp, C.c_uint32, C.POINTER(_vp), C.c_uint32,
      C.c_uint32, _vp, C.POINTER(_vp))
    f("gkr_so_unipoly", C.c_int, _vp, _vp, C.POINTER(C.c_uint32))
    f("gkr_so_bind", C.c_int, _vp, _vp)
    f("gkr_so_final_evals", C.c_int, _vp, _vp)
    f("gkr_so_claim", C.c_int, _vp, _vp)
    f("gkr_so_degree", C.c_uint32, _vp)
    f("gkr_so_num_polys", C.c_uint32, _vp)
    f("gkr_so_round", C.c_uint32, _vp)
    f("gkr_so_destroy", None, _vp)
    f("gkr_vecvec_upload", C.c_int, _vp, _vp, _vp, C.c_uint32, _vp, _vp, C.c_uint32, C.c_uint32, C.POINTER(_vp))
    f("gkr_vecvec_num_rows", C.c_uint32, _vp)
    f("gkr_vecvec_total_len", C.c_uint64, _vp)
    f("gkr_vecvec_download", C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32))
    f("gkr_vecvec_free", None, _vp)
    f("gkr_so_create_deg2_dense", C.c_int, _vp, _vp, _vp, C.c_uint32, C.POINTER(_vp), C.c_uint32, _vp, _vp, _vp, C.c_uint32, C.POINTER(_vp))
    f("gkr_so_create_deg2_vecvec", C.c_int, _vp, C.c_int, C.POINTER(_vp), C.c_uint32, _vp, _vp, _vp, C.c_uint32, C.c_uint32, C.POINTER(_vp))
    f("gkr_transcript_new", C.c_int, _vp, C.c_size_t, C.POINTER(_vp))
    f("gkr_transcript_free", None, _vp)
    f("gkr_transcript_write_scalars", C.c_int, _vp, _vp, C.c_uint32)
    f("gkr_transcript_write_raw", C.c_int, _vp, _vp, C.c_size_t)
    f("gkr_transcript_challenge", C.c_int, _vp, C.c_uint32, _vp)
    f("gkr_transcript_raw_challenge", C.c_int, _vp, _vp, C.c_size_t)
    f("gkr_transcript_proof_len", C.c_size_t, _vp)
    f("gkr_transcript_proof", C.c_int, _vp, _vp)
    f("gkr_sumcheck_prove", C.c_int, _vp, _vp, C.c_uint32, _vp, _vp, _vp)
    f("gkr_exchange_open", C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(_vp))
    f("gkr_exchange_close", None, _vp)
    f("gkr_exchange_allgather", C.c_int, _vp, _vp, C.c_uint32, _vp)
    f("gkr_sumcheck_prove_sharded", C.c_int, _vp, _vp, _vp, C.c_uint32, C.c_int, C.c_int, C.c_uint32, _vp, C.c_uint32, _vp, _vp, _vp, _vp)


def load_library():
    """dlopen the in-tree shared library; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GkrError(GKR_ERR_CUDA, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                         "(there is no CPU fallback)")
        _lib = C.CDLL(os.environ.get("GKR_LIB", LIB_PATH))  # GKR_LIB: an experimental build of the same sources (tools/kernel_lab)
        _sig(_lib)
    return _lib


def _limbs(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_vp)


class Context:
    """gkr_ctx: one per GPU / per calling thread."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = _vp()
        rc = self.lib.gkr_ctx_create(device, C.byref(h))
        if rc != 0:
            raise GkrError(rc, "gkr_ctx_create failed: no usable CUDA device (this backend has no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.gkr_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise GkrError(rc, self.lib.gkr_last_error(self.h).decode())

    def sync(self):
        self.check(self.lib.gkr_ctx_sync(self.h))

    @property
    def launches(self) -> int:
        return int(self.lib.gkr_ctx_launch_count(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.gkr_ctx_stream(self.h) or 0)

    def host_stats(self, reset: bool = True):
        """(ns in round-kernel launch calls, ns waiting for round results, waits, kernels launched)"""
        out = np.zeros(4, np.uint64)
        self.check(self.lib.gkr_ctx_host_stats(self.h, _ptr(out), 1 if reset else 0))
        return tuple(int(x) for x in out)

    def set_fast_fold(self, on: bool = True):
        self.check(self.lib.gkr_ctx_set_fast_fold(self.h, 1 if on else 0))

    def peer_pool(self, n_devices: int = 0):
        """gkr_ctx_peer_pool: lend this context the HBM of devices [0, n_devices); returns (bytes on peers now, their peak)"""
        self.lib.gkr_ctx_peer_pool.restype = C.c_int
        self.lib.gkr_ctx_peer_pool.argtypes = [_vp, C.c_int, _vp]
        st = np.zeros(2, np.uint64)
        self.check(self.lib.gkr_ctx_peer_pool(self.h, int(n_devices), _ptr(st)))
        return int(st[0]), int(st[1])

    def set_tuning(self, key: str, value: int):
        """kernel-selection knob by name (gkr_ctx_set_tuning): every setting is bit-exact"""
        self.check(self.lib.gkr_ctx_set_tuning(self.h, key.encode(), int(value)))

    def timing_enable(self, on: bool = True):
        self.check(self.lib.gkr_ctx_timing_enable(self.h, 1 if on else 0))

    def timing_read(self, max_n: int = 4096):
        """[(kernel_id, n_items, ms)] of every launch recorded since the last read."""
        kid = np.zeros(max_n, np.int32)
        items = np.zeros(max_n, np.uint64)
        ms = np.zeros(max_n, np.float32)
        n = self.lib.gkr_ctx_timing_read(self.h, _ptr(kid), _ptr(items), _ptr(ms), max_n)
        return [(int(kid[i]), int(items[i]), float(ms[i])) for i in range(n)]

    # -- tables ------------------------------------------------------------------------------
    def upload(self, limbs) -> "Table":
        a = _limbs(limbs).reshape(-1, 4)
        return self.upload_ptr(a.ctypes.data, a.shape[0], keep=a)

    def upload_ptr(self, host_ptr: int, n: int, keep=None) -> "Table":
        h = _vp()
        self.check(self.lib.gkr_table_upload(self.h, _vp(host_ptr), n, C.byref(h)))
        t = Table(self, h)
        t._keep = keep
        return t

    def alloc(self, n: int) -> "Table":
        h = _vp()
        self.check(self.lib.gkr_table_alloc(self.h, n, C.byref(h)))
        return Table(self, h)

    def synth(self, seed: int, n: int, first_index: int = 0) -> "Table":
        h = _vp()
        self.check(self.lib.gkr_table_synth(self.h, seed & 0xFFFFFFFFFFFFFFFF, first_index, n, C.byref(h)))
        return Table(self, h)

    def eq_table(self, point, mult=None) -> "Table":
        p = _limbs(point).reshape(-1, 4)
        if mult is None:
            mult = MONT_ONE
        m = _limbs(mult).reshape(4)
        h = _vp()
        self.check(self.lib.gkr_eq_table(self.h, _ptr(p), p.shape[0], _ptr(m), C.byref(h)))
        return Table(self, h)

    def gate_sum(self, so_kind, gate, tables, gate_param=0, consts=None) -> np.ndarray:
        c = _limbs(consts).reshape(-1, 4) if consts is not None else np.zeros((0, 4), np.uint64)
        arr = (_vp * len(tables))(*[t.h for t in tables])
        out = np.zeros(4, np.uint64)
        self.check(self.lib.gkr_dense_gate_sum(self.h, so_kind, gate, gate_param, _ptr(c), c.shape[0], arr, len(tables), _ptr(out)))
        return out

    def dense_so(self, so_kind, gate, tables, num_vars, claim, gate_param=0, consts=None) -> "SumcheckObject":
        c = _limbs(consts).reshape(-1, 4) if consts is not None else np.zeros((0, 4), np.uint64)
        arr = (_vp * len(tables))(*[t.h for t in tables])
        cl = _limbs(claim).reshape(4)
        h = _vp()
        self.check(self.lib.gkr_so_create_dense(self.h, so_kind, gate, gate_param, _ptr(c), c.shape[0], arr, len(tables),
                                                num_vars, _ptr(cl), C.byref(h)))
        return SumcheckObject(self, h, list(tables))


    # appended methods are attached below (deg2 objects, vecvec upload)


def _ctx_upload_vecvec(self, rows, row_pad, col_pad, row_logsize, col_logsize) -> "VecVec":
    """VecVecPolynomial::new: rows = list of (len_r, 4) uint64 arrays (Montgomery limbs)."""
    lens = np.array([len(r) for r in rows], dtype=np.uint32)
    flat = np.concatenate([_limbs(r).reshape(-1, 4) for r in rows] + [np.zeros((0, 4), np.uint64)]) if len(rows) else np.zeros((0, 4), np.uint64)
    flat = np.ascontiguousarray(flat)
    rp, cp = _limbs(row_pad).reshape(4), _limbs(col_pad).reshape(4)
    h = _vp()
    self.check(self.lib.gkr_vecvec_upload(self.h, _ptr(flat), _ptr(lens), len(rows), _ptr(rp), _ptr(cp), row_logsize, col_logsize, C.byref(h)))
    return VecVec(self, h)


def _ctx_deg2_dense_so(self, parts, tables, gamma_pows, claim, point) -> "SumcheckObject":
    """parts: list of (gate_id, repeat)."""
    pg = np.array([p[0] for p in parts], dtype=np.int32)
    pr = np.array([p[1] for p in parts], dtype=np.uint32)
    arr = (_vp * len(tables))(*[t.h for t in tables])
    gp, cl, pt = _limbs(gamma_pows).reshape(-1, 4), _limbs(claim).reshape(4), _limbs(point).reshape(-1, 4)
    h = _vp()
    self.check(self.lib.gkr_so_create_deg2_dense(self.h, _ptr(pg), _ptr(pr), len(parts), arr, len(tables), _ptr(gp), _ptr(cl), _ptr(pt),
                                                 pt.shape[0], C.byref(h)))
    return SumcheckObject(self, h, list(tables))


def _ctx_deg2_vecvec_so(self, gate, polys, gamma_pows, claim, point, col_logsize) -> "SumcheckObject":
    arr = (_vp * len(polys))(*[p.h for p in polys])
    gp, cl, pt = _limbs(gamma_pows).reshape(-1, 4), _limbs(claim).reshape(4), _limbs(point).reshape(-1, 4)
    h = _vp()
    self.check(self.lib.gkr_so_create_deg2_vecvec(self.h, gate, arr, len(polys), _ptr(gp), _ptr(cl), _ptr(pt), pt.shape[0], col_logsize,
                                                  C.byref(h)))
    return SumcheckObject(self, h, list(polys))


def _ctx_deg2_vecvec_shard_so(self, gate, polys, gamma_pows, point, col_logsize, shard, n_shards) -> "SumcheckObject":
    """row shard `shard` of `n_shards` of a VecVecDeg2SumcheckObjectSO (gkr_so_create_deg2_vecvec_shard): polys hold this shard's
    rows (col_logsize - log2(n_shards) column variables); point / col_logsize describe the whole object"""
    lib = self.lib
    lib.gkr_so_create_deg2_vecvec_shard.restype = C.c_int
    lib.gkr_so_create_deg2_vecvec_shard.argtypes = [_vp, C.c_int, C.POINTER(_vp), C.c_uint32, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32,
                                                    C.c_uint32, C.POINTER(_vp)]
    arr = (_vp * len(polys))(*[p.h for p in polys])
    gp, pt = _limbs(gamma_pows).reshape(-1, 4), _limbs(point).reshape(-1, 4)
    h = _vp()
    self.check(lib.gkr_so_create_deg2_vecvec_shard(self.h, gate, arr, len(polys), _ptr(gp), _ptr(pt), pt.shape[0], col_logsize, shard, n_shards,
                                                   C.byref(h)))
    return SumcheckObject(self, h, list(polys))


Context.upload_vecvec = _ctx_upload_vecvec
Context.deg2_dense_so = _ctx_deg2_dense_so
Context.deg2_vecvec_so = _ctx_deg2_vecvec_so
Context.deg2_vecvec_shard_so = _ctx_deg2_vecvec_shard_so


class VecVec:
    """gkr_vecvec: VecVecPolynomial<F> resident in HBM."""

    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h

    @property
    def num_rows(self) -> int:
        return int(self.ctx.lib.gkr_vecvec_num_rows(self.h))

    @property
    def total_len(self) -> int:
        return int(self.ctx.lib.gkr_vecvec_total_len(self.h))

    def download(self):
        """(rows: list of (len, 4) arrays, row_pad, col_pad, row_logsize, col_logsize)"""
        flat = np.zeros((max(self.total_len, 1), 4), np.uint64)
        lens = np.zeros(max(self.num_rows, 1), np.uint32)
        rp, cp = np.zeros(4, np.uint64), np.zeros(4, np.uint64)
        rl, cl = C.c_uint32(0), C.c_uint32(0)
        self.ctx.check(self.ctx.lib.gkr_vecvec_download(self.ctx.h, self.h, _ptr(flat), _ptr(lens), _ptr(rp), _ptr(cp), C.byref(rl), C.byref(cl)))
        rows, off = [], 0
        for r in range(self.num_rows):
            rows.append(flat[off:off + int(lens[r])].copy())
            off += int(lens[r])
        return rows, rp, cp, rl.value, cl.value

    def free(self):
        if self.h:
            self.ctx.lib.gkr_vecvec_free(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.free()
        except Exception:
            pass


MONT_ONE = np.array([0x00000001FFFFFFFE, 0x5884B7FA00034802, 0x998C4FEFECBC4FF5, 0x1824B159ACC5056F], dtype=np.uint64)


class Table:
    """gkr_table: a `Vec<Fr>` resident in HBM."""

    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h
        self._keep = None

    def __len__(self):
        return int(self.ctx.lib.gkr_table_len(self.h))

    @property
    def device_ptr(self) -> int:
        return int(self.ctx.lib.gkr_table_device_ptr(self.h) or 0)

    def download(self) -> np.ndarray:
        out = np.empty((len(self), 4), np.uint64)
        self.ctx.check(self.ctx.lib.gkr_table_download(self.ctx.h, self.h, _ptr(out)))
        return out

    def free(self):
        if self.h:
            self.ctx.lib.gkr_table_free(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.free()
        except Exception:
            pass


class SumcheckObject:
    """gkr_so: `trait Sumcheckable` (src/cleanup/protocols/sumchecks/vecvec_eq.rs:218-225)."""

    def __init__(self, ctx: Context, h, keep):
        self.ctx, self.h, self._keep = ctx, h, keep

    def unipoly(self) -> np.ndarray:
        """evaluations of the round polynomial at 0..deg, shape (deg+1, 4)."""
        out = np.zeros((8, 4), np.uint64)
        n = C.c_uint32(0)
        self.ctx.check(self.ctx.lib.gkr_so_unipoly(self.h, _ptr(out), C.byref(n)))
        return out[: n.value].copy()

    def bind(self, t):
        tt = _limbs(t).reshape(4)
        self.ctx.check(self.ctx.lib.gkr_so_bind(self.h, _ptr(tt)))

    def final_evals(self) -> np.ndarray:
        out = np.zeros((self.num_polys, 4), np.uint64)
        self.ctx.check(self.ctx.lib.gkr_so_final_evals(self.h, _ptr(out)))
        return out

    @property
    def claim(self) -> np.ndarray:
        out = np.zeros(4, np.uint64)
        self.ctx.lib.gkr_so_claim(self.h, _ptr(out))
        return out

    @property
    def degree(self) -> int:
        return int(self.ctx.lib.gkr_so_degree(self.h))

    @property
    def num_polys(self) -> int:
        return int(self.ctx.lib.gkr_so_num_polys(self.h))

    @property
    def round(self) -> int:
        return int(self.ctx.lib.gkr_so_round(self.h))

    def set_prelaunch(self, on: bool = True):
        """gkr_so_set_prelaunch: promise the strict unipoly -> bind alternation with no other device work in between, so that small
        rounds are enqueued one round ahead (the challenge travels through a mailbox in mapped host memory)"""
        self.ctx.lib.gkr_so_set_prelaunch.restype = C.c_int
        self.ctx.lib.gkr_so_set_prelaunch.argtypes = [_vp, C.c_int]
        self.ctx.check(self.ctx.lib.gkr_so_set_prelaunch(self.h, 1 if on else 0))

    def destroy(self):
        if self.h:
            self.ctx.lib.gkr_so_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.destroy()
        except Exception:
            pass


class Transcript:
    """gkr_transcript: ProofTranscript2 in prover mode (src/cleanup/proof_transcript.rs:76-147)."""

    def __init__(self, label: bytes):
        self.lib = load_library()
        h = _vp()
        buf = (C.c_uint8 * len(label)).from_buffer_copy(label) if label else None
        rc = self.lib.gkr_transcript_new(buf, len(label), C.byref(h))
        if rc:
            raise GkrError(rc, "gkr_transcript_new")
        self.h = h

    def write_scalars(self, limbs):
        a = _limbs(limbs).reshape(-1, 4)
        rc = self.lib.gkr_transcript_write_scalars(self.h, _ptr(a), a.shape[0])
        if rc:
            raise GkrError(rc, "write_scalars: non-canonical element")

    def write_raw(self, msg: bytes):
        buf = (C.c_uint8 * len(msg)).from_buffer_copy(msg) if msg else None
        rc = self.lib.gkr_transcript_write_raw(self.h, buf, len(msg))
        if rc:
            raise GkrError(rc, "write_raw")

    def challenge(self, bitsize: int = 128) -> np.ndarray:
        out = np.zeros(4, np.uint64)
        rc = self.lib.gkr_transcript_challenge(self.h, bitsize, _ptr(out))
        if rc:
            raise GkrError(rc, "challenge")
        return out

    def raw_challenge(self, n: int) -> bytes:
        buf = (C.c_uint8 * n)()
        rc = self.lib.gkr_transcript_raw_challenge(self.h, buf, n)
        if rc:
            raise GkrError(rc, "raw_challenge")
        return bytes(buf)

    # -- old API (src/transcript.rs:78-101) --
    def append_scalars_old(self, limbs):
        a = _limbs(limbs).reshape(-1, 4)
        self.lib.gkr_transcript_append_scalars_old.argtypes = [_vp, _vp, C.c_uint32]
        rc = self.lib.gkr_transcript_append_scalars_old(self.h, _ptr(a), a.shape[0])
        if rc:
            raise GkrError(rc, "append_scalars_old: non-canonical element")

    def challenge_scalar_old(self, label: bytes) -> np.ndarray:
        out = np.zeros(4, np.uint64)
        buf = (C.c_uint8 * len(label)).from_buffer_copy(label) if label else None
        self.lib.gkr_transcript_challenge_scalar_old.argtypes = [_vp, _vp, C.c_size_t, _vp]
        rc = self.lib.gkr_transcript_challenge_scalar_old(self.h, buf, len(label), _ptr(out))
        if rc:
            raise GkrError(rc, "challenge_scalar_old")
        return out

    def proof(self) -> bytes:
        n = int(self.lib.gkr_transcript_proof_len(self.h))
        buf = (C.c_uint8 * max(n, 1))()
        self.lib.gkr_transcript_proof(self.h, buf)
        return bytes(buf[:n])

    def free(self):
        if self.h:
            self.lib.gkr_transcript_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def sumcheck_prove(transcript: Transcript, so: SumcheckObject, num_rounds: int):
    """GenericSumcheckProtocol::prove (src/cleanup/protocols/sumcheck.rs:101-123).
    Returns (claim, point[num_rounds] (reversed like the reference), final_evals)."""
    claim = np.zeros(4, np.uint64)
    point = np.zeros((max(num_rounds, 1), 4), np.uint64)
    fe = np.zeros((so.num_polys, 4), np.uint64)
    rc = so.ctx.lib.gkr_sumcheck_prove(transcript.h, so.h, num_rounds, _ptr(claim), _ptr(point), _ptr(fe))
    so.ctx.check(rc)
    return claim, point[:num_rounds], fe


class Exchange:
    """gkr_exchange: shared-memory all-gather between the ranks (one process per GPU) of one box."""

    def __init__(self, name: str, rank: int, world: int, create: bool):
        self.lib = load_library()
        h = _vp()
        rc = self.lib.gkr_exchange_open(name.encode(), rank, world, 1 if create else 0, C.byref(h))
        if rc:
            raise GkrError(rc, f"gkr_exchange_open({name})")
        self.h, self.rank, self.world = h, rank, world

    def allgather(self, mine) -> np.ndarray:
        a = _limbs(mine).reshape(-1, 4)
        out = np.zeros((self.world, a.shape[0], 4), np.uint64)
        rc = self.lib.gkr_exchange_allgather(self.h, _ptr(a), a.shape[0], _ptr(out))
        if rc:
            raise GkrError(rc, "gkr_exchange_allgather")
        return out

    def close(self):
        if self.h:
            self.lib.gkr_exchange_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sumcheck_prove_sharded(transcript: Transcript, so: SumcheckObject, exchange, local_rounds: int, so_kind: int, gate: int,
                           global_claim, gate_param: int = 0, consts=None):
    """GenericSumcheckProtocol::prove over a hypercube sharded by its top index bits (exchange=None: 1 GPU).
    Returns (claim, point, final_evals) -- identical on every rank."""
    world = exchange.world if exchange is not None else 1
    g_ = world.bit_length() - 1
    c = _limbs(consts).reshape(-1, 4) if consts is not None else np.zeros((0, 4), np.uint64)
    claim = np.zeros(4, np.uint64)
    point = np.zeros((max(local_rounds + g_, 1), 4), np.uint64)
    fe = np.zeros((so.num_polys, 4), np.uint64)
    gc = _limbs(global_claim).reshape(4)
    rc = so.ctx.lib.gkr_sumcheck_prove_sharded(transcript.h, so.h, exchange.h if exchange is not None else None, local_rounds,
                                               so_kind, gate, gate_param, _ptr(c), c.shape[0], _ptr(gc), _ptr(claim), _ptr(point), _ptr(fe))
    so.ctx.check(rc)
    return claim, point[: local_rounds + g_], fe


def sumcheck_prove_sharded_vecvec(transcript: Transcript, so: SumcheckObject, exchange, num_vars: int, global_claim):
    """VecVecDeg2Sumcheck::prove with one row shard per rank (gkr_sumcheck_prove_sharded_vecvec; exchange=None: one shard).
    num_vars: variables of the WHOLE object.  Returns (claim, point, final_evals) -- identical on every rank."""
    lib = so.ctx.lib
    lib.gkr_sumcheck_prove_sharded_vecvec.restype = C.c_int
    lib.gkr_sumcheck_prove_sharded_vecvec.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp]
    claim = np.zeros(4, np.uint64)
    point = np.zeros((max(num_vars, 1), 4), np.uint64)
    fe = np.zeros((so.num_polys, 4), np.uint64)
    gc = _limbs(global_claim).reshape(4)
    so.ctx.check(lib.gkr_sumcheck_prove_sharded_vecvec(transcript.h, so.h, exchange.h if exchange is not None else None, _ptr(gc), _ptr(claim),
                                                       _ptr(point), _ptr(fe)))
    return claim, point[:num_vars], fe


# ---- witness maps (trait MapSplit) -------------------------------------------------------------------------
def _parts_arrays(parts):
    pg = np.array([p[0] for p in parts], dtype=np.int32)
    pr = np.array([p[1] for p in parts], dtype=np.uint32)
    return pg, pr


def _ctx_map_dense(self, parts, tables, split=None, bundle_size=1):
    """Vec::algfn_map (split=None) / Vec::algfn_map_split (split=("LO"|"HI", var_idx)).  Returns a list of Tables."""
    lib = self.lib
    if not hasattr(lib.gkr_map_dense, "_sig"):
        lib.gkr_map_dense.restype = C.c_int
        lib.gkr_map_dense.argtypes = [_vp, _vp, _vp, C.c_uint32, C.POINTER(_vp), C.c_uint32, C.c_int, C.c_uint32, C.c_uint32,
                                      C.POINTER(_vp), C.POINTER(C.c_uint32)]
        lib.gkr_map_dense._sig = True
    pg, pr = _parts_arrays(parts)
    arr = (_vp * len(tables))(*[t.h for t in tables])
    out = (_vp * 256)()
    n = C.c_uint32(0)
    kind, var = (-1, 0) if split is None else ({"LO": 0, "HI": 1}[split[0]], split[1])
    self.check(lib.gkr_map_dense(self.h, _ptr(pg), _ptr(pr), len(parts), arr, len(tables), kind, var, bundle_size, out, C.byref(n)))
    return [Table(self, _vp(out[i])) for i in range(n.value)]


def _ctx_map_vecvec(self, parts, polys, mode=0, bundle_size=1):
    """mode 0: vecvec_map, 1: vecvec_map_split (LO(0)), 2: vecvec_map_split_to_dense (returns Tables)."""
    lib = self.lib
    if not hasattr(lib.gkr_map_vecvec, "_sig"):
        lib.gkr_map_vecvec.restype = C.c_int
        lib.gkr_map_vecvec.argtypes = [_vp, _vp, _vp, C.c_uint32, C.POINTER(_vp), C.c_uint32, C.c_int, C.c_uint32, C.POINTER(_vp),
                                       C.POINTER(C.c_uint32)]
        lib.gkr_map_vecvec._sig = True
    pg, pr = _parts_arrays(parts)
    arr = (_vp * len(polys))(*[p.h for p in polys])
    out = (_vp * 256)()
    n = C.c_uint32(0)
    self.check(lib.gkr_map_vecvec(self.h, _ptr(pg), _ptr(pr), len(parts), arr, len(polys), mode, bundle_size, out, C.byref(n)))
    if mode == 2:
        return [Table(self, _vp(out[i])) for i in range(n.value)]
    return [VecVec(self, _vp(out[i])) for i in range(n.value)]


Context.map_dense = _ctx_map_dense
Context.map_vecvec = _ctx_map_vecvec


# ---- commitments ------------------------------------------------------------------------------------------
def _srs_sigs(lib):
    if hasattr(lib.gkr_srs_upload, "_sig"):
        return
    lib.gkr_srs_upload.restype = C.c_int
    lib.gkr_srs_upload.argtypes = [_vp, _vp, C.c_uint64, C.c_int, C.POINTER(_vp)]
    lib.gkr_srs_len.restype = C.c_uint64
    lib.gkr_srs_len.argtypes = [_vp]
    lib.gkr_srs_free.restype = None
    lib.gkr_srs_free.argtypes = [_vp]
    lib.gkr_msm_g1.restype = C.c_int
    lib.gkr_msm_g1.argtypes = [_vp, _vp, C.c_uint64, _vp, C.c_uint64, _vp]
    lib.gkr_srs_mock_setup.restype = C.c_int
    lib.gkr_srs_mock_setup.argtypes = [_vp, _vp, _vp, C.c_uint64, C.POINTER(_vp)]
    lib.gkr_msm_g1_batch.restype = C.c_int
    lib.gkr_msm_g1_batch.argtypes = [_vp, _vp, C.c_uint64, C.c_uint64, C.c_uint32, _vp, C.c_uint64, _vp]
    lib.gkr_g1_weighted_bucket_sums.restype = C.c_int
    lib.gkr_g1_weighted_bucket_sums.argtypes = [_vp, _vp, C.c_uint64, C.c_uint32, C.c_uint32, _vp]
    lib.gkr_srs_upload._sig = True


class Srs:
    """gkr_srs: G1 bases resident in HBM (affine (n, 12) or Jacobian (n, 18) uint64 Montgomery limbs)."""

    def __init__(self, ctx: Context, points, projective: bool = False):
        lib = ctx.lib
        _srs_sigs(lib)
        a = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 18 if projective else 12)
        h = _vp()
        ctx.check(lib.gkr_srs_upload(ctx.h, _ptr(a), a.shape[0], 1 if projective else 0, C.byref(h)))
        self.ctx, self.h, self.n = ctx, h, a.shape[0]

    def msm(self, scalars: "Table", n: int = None, first: int = 0) -> np.ndarray:
        """KzgProvingKey::commit: returns the affine result as (12,) uint64 (x | y), zeros for infinity."""
        n = len(scalars) if n is None else n
        out = np.zeros(12, np.uint64)
        self.ctx.check(self.ctx.lib.gkr_msm_g1(self.ctx.h, self.h, first, scalars.h, n, _ptr(out)))
        return out

    def precompute(self, c: int = 20):
        """fixed-base window table T[k][i] = 2^(c k) P_i (proving-key preprocessing, see include/gkr_msm_b200.h)"""
        self.ctx.lib.gkr_srs_precompute.restype = C.c_int
        self.ctx.lib.gkr_srs_precompute.argtypes = [_vp, _vp, C.c_int]
        self.ctx.check(self.ctx.lib.gkr_srs_precompute(self.ctx.h, self.h, int(c)))

    def msm_batch(self, scalars: "Table", n: int, first: int, stride: int, count: int) -> np.ndarray:
        """`count` MSMs with the same scalars over the base ranges [first + p * stride, + n): (count, 12) affine results."""
        out = np.zeros((count, 12), np.uint64)
        self.ctx.check(self.ctx.lib.gkr_msm_g1_batch(self.ctx.h, self.h, first, stride, count, scalars.h, n, _ptr(out)))
        return out

    def weighted_sums(self, group_log: int, count: int, first: int = 0) -> np.ndarray:
        """running-sum commitments sum_i i * B[first + (k << group_log) + i] of `count` bucket groups: (count, 12)."""
        out = np.zeros((count, 12), np.uint64)
        self.ctx.check(self.ctx.lib.gkr_g1_weighted_bucket_sums(self.ctx.h, self.h, first, group_log, count, _ptr(out)))
        return out

    def free(self):
        if self.h:
            self.ctx.lib.gkr_srs_free(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.free()
        except Exception:
            pass


def _srs_bucket_sums(self, point_idx, bucket_idx, n_buckets) -> "Srs":
    """PushForwardState::new bucket accumulation (pushforward.rs:398-429): returns the bucket sums as a resident Srs."""
    lib = self.ctx.lib
    if not hasattr(lib.gkr_g1_bucket_sums, "_sig"):
        lib.gkr_g1_bucket_sums.restype = C.c_int
        lib.gkr_g1_bucket_sums.argtypes = [_vp, _vp, _vp, _vp, C.c_uint64, C.c_uint32, C.POINTER(_vp)]
        lib.gkr_g1_weighted_bucket_sum.restype = C.c_int
        lib.gkr_g1_weighted_bucket_sum.argtypes = [_vp, _vp, _vp]
        lib.gkr_g1_download_affine.restype = C.c_int
        lib.gkr_g1_download_affine.argtypes = [_vp, _vp, _vp]
        lib.gkr_g1_bucket_sums._sig = True
    p = np.ascontiguousarray(point_idx, dtype=np.uint32)
    b = np.ascontiguousarray(bucket_idx, dtype=np.uint32)
    h = _vp()
    self.ctx.check(lib.gkr_g1_bucket_sums(self.ctx.h, self.h, _ptr(p), _ptr(b), p.shape[0], n_buckets, C.byref(h)))
    out = Srs.__new__(Srs)
    out.ctx, out.h, out.n = self.ctx, h, n_buckets
    return out


def _srs_weighted_sum(self) -> np.ndarray:
    out = np.zeros(12, np.uint64)
    self.ctx.check(self.ctx.lib.gkr_g1_weighted_bucket_sum(self.ctx.h, self.h, _ptr(out)))
    return out


def _srs_download_affine(self) -> np.ndarray:
    out = np.zeros((self.n, 12), np.uint64)
    self.ctx.check(self.ctx.lib.gkr_g1_download_affine(self.ctx.h, self.h, _ptr(out)))
    return out


def _srs_mock_setup(ctx: Context, tau, g0_xy, n: int) -> "Srs":
    """KzgProvingKey::mock_setup(tau, g0, _, n).ptau_1 (kzg.rs:84-97), generated on the device."""
    lib = ctx.lib
    _srs_sigs(lib)
    t, g0 = _limbs(tau).reshape(4), _limbs(g0_xy).reshape(12)
    h = _vp()
    ctx.check(lib.gkr_srs_mock_setup(ctx.h, _ptr(t), _ptr(g0), n, C.byref(h)))
    out = Srs.__new__(Srs)
    out.ctx, out.h, out.n = ctx, h, n
    return out


def _srs_bucket_sums_rows(self, idx: "U32Buf", x_logsize: int, clm: int, group_log: int) -> "Srs":
    """bucket sums of a resident digit / counter matrix (rows of 2^x_logsize entries): incidence (y, x) adds
    bases[x + 2^x_logsize * (y mod 2^clm)] to bucket ((y >> clm) << group_log) | idx[y][x]  (pushforward.rs:401-456)."""
    lib = self.ctx.lib
    if not hasattr(lib.gkr_g1_bucket_sums_rows, "_sig"):
        lib.gkr_g1_bucket_sums_rows.restype = C.c_int
        lib.gkr_g1_bucket_sums_rows.argtypes = [_vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(_vp)]
        lib.gkr_g1_bucket_sums_rows._sig = True
    h = _vp()
    self.ctx.check(lib.gkr_g1_bucket_sums_rows(self.ctx.h, self.h, idx.h, x_logsize, clm, group_log, C.byref(h)))
    rows = idx.n >> x_logsize
    out = Srs.__new__(Srs)
    out.ctx, out.h, out.n = self.ctx, h, (-(-rows // (1 << clm))) << group_log
    return out


Srs.mock_setup = staticmethod(_srs_mock_setup)
Srs.bucket_sums_rows = _srs_bucket_sums_rows
Srs.bucket_sums = _srs_bucket_sums
Srs.weighted_sum = _srs_weighted_sum
Srs.download_affine = _srs_download_affine


def _poly_sigs(lib):
    if hasattr(lib.gkr_u32_upload, "_sig"):
        return
    def f(name, res, *args):
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = list(args)
    f("gkr_u32_upload", C.c_int, _vp, _vp, C.c_uint64, C.POINTER(_vp))
    f("gkr_u32_free", None, _vp)
    f("gkr_table_from_u32", C.c_int, _vp, _vp, C.c_int, C.POINTER(_vp))
    f("gkr_table_gather", C.c_int, _vp, _vp, _vp, C.POINTER(_vp))
    f("gkr_table_lincomb", C.c_int, _vp, C.c_uint32, C.POINTER(_vp), _vp, _vp, _vp, _vp, C.c_uint64, C.POINTER(_vp))
    f("gkr_poly_eval", C.c_int, _vp, _vp, _vp, _vp)
    f("gkr_poly_div_by_linear", C.c_int, _vp, _vp, _vp, C.POINTER(_vp), _vp)
    f("gkr_knuckles_create", C.c_int, _vp, C.c_uint32, _vp, C.POINTER(_vp))
    f("gkr_knuckles_free", None, _vp)
    f("gkr_knuckles_compute_t", C.c_int, _vp, _vp, _vp, _vp, C.c_uint32, C.POINTER(_vp), _vp)
    lib.gkr_u32_upload._sig = True


class U32Buf:
    """gkr_u32buf: digit / counter arrays resident on the device."""

    def __init__(self, ctx: Context, vals):
        _poly_sigs(ctx.lib)
        a = np.ascontiguousarray(vals, dtype=np.uint32).reshape(-1)
        h = _vp()
        ctx.check(ctx.lib.gkr_u32_upload(ctx.h, _ptr(a), a.shape[0], C.byref(h)))
        self.ctx, self.h, self.n = ctx, h, a.shape[0]

    def to_field(self, negate=False) -> Table:
        h = _vp()
        self.ctx.check(self.ctx.lib.gkr_table_from_u32(self.ctx.h, self.h, 1 if negate else 0, C.byref(h)))
        return Table(self.ctx, h)

    def free(self):
        if self.h:
            self.ctx.lib.gkr_u32_free(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.free()
        except Exception:
            pass


def _ctx_gather(self, src: Table, idx: U32Buf) -> Table:
    _poly_sigs(self.lib)
    h = _vp()
    self.check(self.lib.gkr_table_gather(self.h, src.h, idx.h, C.byref(h)))
    return Table(self, h)


def _ctx_lincomb(self, terms, out_len) -> Table:
    """terms: list of (table, coef_limbs, src_off, dst_off, length); out[dst_off+i] += coef*table[src_off+i]."""
    _poly_sigs(self.lib)
    k = len(terms)
    arr = (_vp * max(k, 1))(*[(t[0].h if t[0] is not None else None) for t in terms])  # None = the all-ones table
    coefs = np.ascontiguousarray(np.stack([_limbs(t[1]).reshape(4) for t in terms]) if k else np.zeros((0, 4), np.uint64))
    so = np.array([t[2] for t in terms], dtype=np.uint64)
    do = np.array([t[3] for t in terms], dtype=np.uint64)
    ln = np.array([t[4] for t in terms], dtype=np.uint64)
    h = _vp()
    self.check(self.lib.gkr_table_lincomb(self.h, k, arr, _ptr(coefs), _ptr(so), _ptr(do), _ptr(ln), out_len, C.byref(h)))
    return Table(self, h)


def _ctx_poly_eval(self, poly: Table, x) -> np.ndarray:
    _poly_sigs(self.lib)
    xx, out = _limbs(x).reshape(4), np.zeros(4, np.uint64)
    self.check(self.lib.gkr_poly_eval(self.h, poly.h, _ptr(xx), _ptr(out)))
    return out


def _ctx_div_by_linear(self, poly: Table, pt):
    _poly_sigs(self.lib)
    xx, rem = _limbs(pt).reshape(4), np.zeros(4, np.uint64)
    h = _vp()
    self.check(self.lib.gkr_poly_div_by_linear(self.h, poly.h, _ptr(xx), C.byref(h), _ptr(rem)))
    return Table(self, h), rem


Context.gather = _ctx_gather
Context.lincomb = _ctx_lincomb
Context.poly_eval = _ctx_poly_eval
Context.div_by_linear = _ctx_div_by_linear


class Knuckles:
    """gkr_knuckles: KnucklesProvingKey's field part (inverses table) + compute_t (knuckles.rs:65-81, 111-154)."""

    def __init__(self, ctx: Context, num_vars: int, k):
        _poly_sigs(ctx.lib)
        kk = _limbs(k).reshape(4)
        h = _vp()
        ctx.check(ctx.lib.gkr_knuckles_create(ctx.h, num_vars, _ptr(kk), C.byref(h)))
        self.ctx, self.h, self.num_vars = ctx, h, num_vars

    def compute_t(self, poly: Table, point):
        pt = _limbs(point).reshape(-1, 4)
        opening = np.zeros(4, np.uint64)
        h = _vp()
        self.ctx.check(self.ctx.lib.gkr_knuckles_compute_t(self.ctx.h, self.h, poly.h, _ptr(pt), pt.shape[0], C.byref(h), _ptr(opening)))
        return Table(self.ctx, h), opening

    def free(self):
        if self.h:
            self.ctx.lib.gkr_knuckles_free(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.free()
        except Exception:
            pass


def _ctx_upload_vecvec_flat(self, flat, lens, row_pad, col_pad, row_logsize, col_logsize) -> "VecVec":
    """VecVecPolynomial::new from rows stored back to back (`flat`: (sum lens, 4) limbs, `lens`: row lengths)."""
    flat = np.ascontiguousarray(flat, dtype=np.uint64).reshape(-1, 4)
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    rp, cp = _limbs(row_pad).reshape(4), _limbs(col_pad).reshape(4)
    h = _vp()
    self.check(self.lib.gkr_vecvec_upload(self.h, _ptr(flat), _ptr(lens), lens.shape[0], _ptr(rp), _ptr(cp), row_logsize, col_logsize, C.byref(h)))
    return VecVec(self, h)


def _ctx_vecvec_gather(self, src, idx, lens, row_pad, col_pad, row_logsize, col_logsize) -> "VecVec":
    """VecVecPolynomial::new over rows gathered on the device: row r = src[idx[..]] (src None: all ones)."""
    lib = self.lib
    if not hasattr(lib.gkr_vecvec_gather, "_sig"):
        lib.gkr_vecvec_gather.restype = C.c_int
        lib.gkr_vecvec_gather.argtypes = [_vp, _vp, _vp, _vp, C.c_uint32, _vp, _vp, C.c_uint32, C.c_uint32, C.POINTER(_vp)]
        lib.gkr_vecvec_gather._sig = True
    idx = np.ascontiguousarray(idx, dtype=np.uint32)
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    rp, cp = _limbs(row_pad).reshape(4), _limbs(col_pad).reshape(4)
    h = _vp()
    self.check(lib.gkr_vecvec_gather(self.h, src.h if src is not None else None, _ptr(idx), _ptr(lens), lens.shape[0], _ptr(rp), _ptr(cp),
                                     row_logsize, col_logsize, C.byref(h)))
    return VecVec(self, h)


def _ctx_vecvec_gather_multi(self, srcs, idx, lens, row_pads, col_pads, row_logsize, col_logsize):
    """several polynomials over the same gathered rows (one index upload): srcs[k] None = all ones."""
    lib = self.lib
    if not hasattr(lib.gkr_vecvec_gather_multi, "_sig"):
        lib.gkr_vecvec_gather_multi.restype = C.c_int
        lib.gkr_vecvec_gather_multi.argtypes = [_vp, C.POINTER(_vp), C.c_uint32, _vp, _vp, C.c_uint32, _vp, _vp, C.c_uint32, C.c_uint32, C.POINTER(_vp)]
        lib.gkr_vecvec_gather_multi._sig = True
    k = len(srcs)
    idx = np.ascontiguousarray(idx, dtype=np.uint32)
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    rp = np.ascontiguousarray(np.stack([_limbs(p).reshape(4) for p in row_pads]))
    cp = np.ascontiguousarray(np.stack([_limbs(p).reshape(4) for p in col_pads]))
    arr = (_vp * k)(*[(t.h if t is not None else None) for t in srcs])
    out = (_vp * k)()
    self.check(lib.gkr_vecvec_gather_multi(self.h, arr, k, _ptr(idx), _ptr(lens), lens.shape[0], _ptr(rp), _ptr(cp), row_logsize, col_logsize, out))
    return [VecVec(self, _vp(out[i])) for i in range(k)]


Context.upload_vecvec_flat = _ctx_upload_vecvec_flat
Context.vecvec_gather = _ctx_vecvec_gather
Context.vecvec_gather_multi = _ctx_vecvec_gather_multi


def pushforward_bucketize(coefs_u64, y_size: int, d_logsize: int):
    """PushForwardState::new index bookkeeping (pushforward.rs:351-396) in the host library: returns
    (digits[y][n], counter[y][n], order[y][n], lens[y][2^d]) as uint32 arrays."""
    lib = load_library()
    lib.gkr_pushforward_bucketize.restype = C.c_int
    lib.gkr_pushforward_bucketize.argtypes = [_vp, C.c_uint64, C.c_uint32, C.c_uint32, _vp, _vp, _vp, _vp]
    co = np.ascontiguousarray(coefs_u64, dtype=np.uint64).reshape(-1, 4)
    n = co.shape[0]
    digits, counter, order = (np.empty((y_size, n), np.uint32) for _ in range(3))
    lens = np.empty((y_size, 1 << d_logsize), np.uint32)
    rc = lib.gkr_pushforward_bucketize(_ptr(co), n, y_size, d_logsize, _ptr(digits), _ptr(counter), _ptr(order), _ptr(lens))
    if rc:
        raise GkrError(rc, "gkr_pushforward_bucketize: bad arguments")
    return digits, counter, order, lens


def pushforward_bucketize_dev(ctx: "Context", coefs_u64, y_size: int, d_logsize: int):
    """The same bookkeeping by a stable counting sort on the device (csrc/bucketize.cu).  Returns host copies of
    (digits[y][n], counter[y][n], padded_order (bucket contents, even-padded with 0xffffffff), lens[y][2^d])."""
    lib = ctx.lib
    lib.gkr_pushforward_bucketize_dev.restype = C.c_int
    lib.gkr_pushforward_bucketize_dev.argtypes = [_vp, _vp, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _vp]
    lib.gkr_u32_download.restype = C.c_int
    lib.gkr_u32_download.argtypes = [_vp, _vp, _vp]
    lib.gkr_u32_len.restype = C.c_uint64
    lib.gkr_u32_len.argtypes = [_vp]
    _poly_sigs(lib)
    co = np.ascontiguousarray(coefs_u64, dtype=np.uint64).reshape(-1, 4)
    n = co.shape[0]
    lens = np.empty((y_size, 1 << d_logsize), np.uint32)
    hd, hc, ho = _vp(), _vp(), _vp()
    ctx.check(lib.gkr_pushforward_bucketize_dev(ctx.h, _ptr(co), n, y_size, d_logsize, C.byref(hd), C.byref(hc), C.byref(ho), _ptr(lens)))
    outs = []
    for h in (hd, hc, ho):
        a = np.empty(int(lib.gkr_u32_len(h)), np.uint32)
        ctx.check(lib.gkr_u32_download(ctx.h, h, _ptr(a)))
        lib.gkr_u32_free(h)
        outs.append(a)
    return outs[0].reshape(y_size, n), outs[1].reshape(y_size, n), outs[2], lens


FQ_MONT_ONE = np.array([0x760900000002FFFD, 0xEBF4000BC40C0002, 0x5F48985753C758BA, 0x77CE585370525745, 0x5C071A97A256EC6D,
                        0x15F65EC3FA80E493], dtype=np.uint64)


def g1_sum(points_xy) -> np.ndarray:
    """Sum of a few affine G1 points ((n, 12) Montgomery limbs, all-zero = infinity) on the host (csrc/host_g1.hpp): the
    combine step of an MSM split by point range over several GPUs -- G points, one inversion."""
    lib = load_library()
    lib.gkr_host_g1_horner.restype = C.c_int
    lib.gkr_host_g1_horner.argtypes = [_vp, C.c_int, C.c_int, _vp]
    pts = np.ascontiguousarray(points_xy, dtype=np.uint64).reshape(-1, 12)
    ws = np.zeros((pts.shape[0], 24), np.uint64)
    for i, p in enumerate(pts):
        if p.any():
            ws[i, :12] = p
            ws[i, 12:18] = FQ_MONT_ONE
            ws[i, 18:24] = FQ_MONT_ONE
    out = np.zeros(12, np.uint64)
    rc = lib.gkr_host_g1_horner(_ptr(ws), 0, pts.shape[0], _ptr(out))
    if rc:
        raise GkrError(rc, "gkr_host_g1_horner")
    return out


def binary_msm_prepare(ctx: Context, bases: "Srs", gamma: int) -> "Srs":
    """prepare_bases (src/binary_msm.rs:44-50) on the device: per chunk of `gamma` bases its 2^gamma - 1 subset sums, affine."""
    ctx.lib.gkr_binary_msm_prepare.restype = C.c_int
    ctx.lib.gkr_binary_msm_prepare.argtypes = [_vp, _vp, C.c_uint32, C.POINTER(_vp)]
    h = _vp()
    ctx.check(ctx.lib.gkr_binary_msm_prepare(ctx.h, bases.h, gamma, C.byref(h)))
    out = Srs.__new__(Srs)
    out.ctx, out.h, out.n = ctx, h, int(ctx.lib.gkr_srs_len(h))
    return out


def binary_msm(ctx: Context, prepared: "Srs", gamma: int, coefs) -> np.ndarray:
    """binary_msm (src/binary_msm.rs:19-29): coefs = prepare_coefs' bytes, one per chunk; affine result (12,) uint64."""
    ctx.lib.gkr_binary_msm.restype = C.c_int
    ctx.lib.gkr_binary_msm.argtypes = [_vp, _vp, C.c_uint32, _vp, C.c_uint64, _vp]
    co = np.ascontiguousarray(coefs, dtype=np.uint8).reshape(-1)
    out = np.zeros(12, np.uint64)
    ctx.check(ctx.lib.gkr_binary_msm(ctx.h, prepared.h, gamma, _ptr(co), co.shape[0], _ptr(out)))
    return out


class MsmTeam:
    """Commitment MSMs split by point range over the GPUs of one box (csrc/msm_team.cu).  rank 0 = leader (creates the
    shared segment and attaches the team to its context), ranks > 0 call serve(srs) and answer until the leader quits."""

    def __init__(self, ctx: Context, name: str, rank: int, world: int, max_n: int, open_timeout_s: float = 60.0):
        lib = ctx.lib
        lib.gkr_msm_team_open.restype = C.c_int
        lib.gkr_msm_team_open.argtypes = [_vp, C.c_char_p, C.c_int, C.c_int, C.c_uint64, C.c_int, C.POINTER(_vp)]
        lib.gkr_msm_team_serve.restype = C.c_int
        lib.gkr_msm_team_serve.argtypes = [_vp, _vp, _vp, C.c_double]
        lib.gkr_msm_team_wait_ready.restype = C.c_int
        lib.gkr_msm_team_wait_ready.argtypes = [_vp, _vp, C.c_double]
        lib.gkr_msm_team_quit.restype = None
        lib.gkr_msm_team_quit.argtypes = [_vp]
        lib.gkr_msm_team_close.restype = None
        lib.gkr_msm_team_close.argtypes = [_vp, _vp]
        self.ctx, self.rank, self.world, self.h = ctx, rank, world, _vp()
        import time as _time
        t0 = _time.time()
        while True:  # workers retry until the leader has created the segment
            rc = lib.gkr_msm_team_open(ctx.h, name.encode(), rank, world, max_n, 1 if rank == 0 else 0, C.byref(self.h))
            if rc == 0:
                break
            if rank == 0 or _time.time() - t0 > open_timeout_s:
                ctx.check(rc)
            _time.sleep(0.05)

    def set_min_n(self, n: int):
        self.ctx.lib.gkr_msm_team_set_min_n.restype = None
        self.ctx.lib.gkr_msm_team_set_min_n.argtypes = [_vp, C.c_uint64]
        self.ctx.lib.gkr_msm_team_set_min_n(self.ctx.h, int(n))

    def serve(self, srs: "Srs", idle_timeout_s: float = 120.0):
        self.ctx.check(self.ctx.lib.gkr_msm_team_serve(self.ctx.h, self.h, srs.h, float(idle_timeout_s)))

    def wait_ready(self, timeout_s: float = 120.0):
        self.ctx.check(self.ctx.lib.gkr_msm_team_wait_ready(self.ctx.h, self.h, float(timeout_s)))

    def quit(self):
        if self.h:
            self.ctx.lib.gkr_msm_team_quit(self.h)

    def close(self):
        if self.h:
            self.ctx.lib.gkr_msm_team_close(self.ctx.h, self.h)
            self.h = None


def run_pippenger_native(ctx: Context, transcript: Transcript, srs: "Srs", g0_xy, knuckles: "Knuckles", points_xy, coefs_u64, d_logsize: int,
                         x_logsize: int, num_bits: int, clm: int, r_limbs):
    """benchutils::run_pippenger (pippenger.rs:499-559) with the host orchestration in C++ (csrc/protocol.cu).
    Returns (dense_output (3(d+1), 2^y_logsize, 4), claim_evs (3(d+1), 4), pair (2, 12))."""
    lib = ctx.lib
    if not hasattr(lib.gkr_run_pippenger, "_sig"):
        lib.gkr_run_pippenger.restype = C.c_int
        lib.gkr_run_pippenger.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _vp, _vp, _vp, _vp]
        lib.gkr_run_pippenger._sig = True
    y_size = (num_bits + d_logsize - 1) // d_logsize
    yl = (y_size - 1).bit_length()
    n_out = 3 * (d_logsize + 1)
    px = np.ascontiguousarray(points_xy[0], dtype=np.uint64)
    py = np.ascontiguousarray(points_xy[1], dtype=np.uint64)
    co = np.ascontiguousarray(coefs_u64, dtype=np.uint64)
    g0 = _limbs(g0_xy).reshape(12)
    rr = _limbs(r_limbs).reshape(-1, 4)
    assert rr.shape[0] == yl and px.shape[0] == 1 << x_logsize
    dense = np.zeros((n_out, 1 << yl, 4), np.uint64)
    evs = np.zeros((n_out, 4), np.uint64)
    pair = np.zeros((2, 12), np.uint64)
    ctx.check(lib.gkr_run_pippenger(ctx.h, transcript.h, srs.h, _ptr(g0), knuckles.h, _ptr(px), _ptr(py), _ptr(co), d_logsize, x_logsize, num_bits,
                                    clm, _ptr(rr), _ptr(dense), _ptr(evs), _ptr(pair)))
    return dense, evs, pair
