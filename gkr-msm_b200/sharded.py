"""Host-side driver for the standalone dense sumcheck (BASELINE config[1]) on 1..N GPUs, one process per GPU.

Global instance: P = 3 tables of 2^(log_n_local + log2 world) elements, table j = SplitMix64 stream
`seed + j`; rank r owns the contiguous slice [r * 2^log_n_local, (r+1) * 2^log_n_local) -- the hypercube is
split by its TOP index bits, so the local rounds need no table exchange at all.
The per-round partial sums go through `Exchange` (shared memory between the ranks of one box); the
transcript is replicated, so every rank derives the same challenges and ends with the same proof.

Mirrors: GenericSumcheckProtocol::prove + DenseSumcheckObjectSO (src/cleanup/protocols/sumcheck.rs:95-128,
241-347) with Prod3Fn (src/cleanup/protocols/pushforward/pushforward.rs:38-50).
"""
from __future__ import annotations

import os

import numpy as np

from . import binding as g

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def _limbs_to_int(a) -> int:
    return int(a[0]) | (int(a[1]) << 64) | (int(a[2]) << 128) | (int(a[3]) << 192)


def _int_to_limbs(v: int) -> np.ndarray:
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def mont_add_many(rows) -> np.ndarray:
    """sum of Montgomery-form elements (addition commutes with the Montgomery map)."""
    s = 0
    for r in rows:
        s += _limbs_to_int(r)
    return _int_to_limbs(s % R_MOD)


class ShardedProd3Sumcheck:
    P = 3

    def __init__(self, ctx: g.Context, log_n_local: int, rank: int = 0, world: int = 1, dist=None, seed: int = 1, exchange=None):
        assert world & (world - 1) == 0, "world size must be a power of two"
        self.ctx, self.log_n, self.rank, self.world, self.dist, self.seed = ctx, log_n_local, rank, world, dist, seed
        n = 1 << log_n_local
        self.n = n
        self.tables = [ctx.synth(seed + j, n, first_index=rank * n) for j in range(self.P)]
        self.exchange = exchange
        if world > 1 and exchange is None:
            name = f"/gkr_msm_b200_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}"
            if rank == 0:
                self.exchange = g.Exchange(name, rank, world, create=True)
            dist.barrier()
            if rank != 0:
                self.exchange = g.Exchange(name, rank, world, create=False)
            dist.barrier()
        local = ctx.gate_sum(g.SO_PLAIN, g.GATE_PROD3, self.tables)
        if world > 1:
            allc = self.exchange.allgather(local.reshape(1, 4))
            self.claim = mont_add_many([allc[r, 0] for r in range(world)])
        else:
            self.claim = local
        self.host_tables = None
        self.h2d_bytes = self.P * n * 32
        deg = 3
        self.d2h_bytes = log_n_local * deg * 32 + self.P * 32
        self.last = None

    def _prove(self, tables):
        so = self.ctx.dense_so(g.SO_PLAIN, g.GATE_PROD3, tables, self.log_n, self.claim)
        tr = g.Transcript(b"fgstglsp")
        out = g.sumcheck_prove_sharded(tr, so, self.exchange, self.log_n, g.SO_PLAIN, g.GATE_PROD3, self.claim)
        tr.write_scalars(out[2])
        self.last = (out, tr.proof())
        so.destroy()
        return out

    def prove_resident(self):
        """tables already in HBM (what the surrounding protocol gives the sumcheck: witness maps and eq tables
        are produced on the device)."""
        return self._prove(self.tables)

    def prepare_host_inputs(self):
        import torch

        self.host_tables = []
        for t in self.tables:
            buf = torch.empty((self.n, 4), dtype=torch.int64).pin_memory()
            arr = buf.numpy().view(np.uint64)
            self.ctx.check(self.ctx.lib.gkr_table_download(self.ctx.h, t.h, arr.ctypes.data_as(g._vp)))
            self.host_tables.append((buf, arr))

    def prove_from_host(self):
        """end-to-end: tables start in (pinned) HOST memory -- upload, prove, results back on the host."""
        tabs = [self.ctx.upload_ptr(arr.ctypes.data, self.n, keep=buf) for (buf, arr) in self.host_tables]
        out = self._prove(tabs)
        for t in tabs:
            t.free()
        return out


class ShardedVecVecSumcheck:
    """VecVecDeg2Sumcheck::prove (src/cleanup/protocols/sumchecks/vecvec_eq.rs:447-500) with the bucket rows split over the ranks by the
    top bits of the row index: every rank holds 2^col_local_log full rows of 2^row_log elements of the 6 coordinate
    polynomials of a projective addition layer (gate twisted_edwards_add_l1, the reference's point padding), built on its GPU.
    Weak scaling: the rows per GPU are fixed.  The claim is synthetic (this is a timing job; parity of the sharded path with the
    single-GPU object and the oracle is in tests/test_gpu_sharded.py)."""
    P = 6

    def __init__(self, ctx: g.Context, row_log: int, col_local_log: int, rank: int = 0, world: int = 1, exchange=None, seed: int = 11):
        assert world & (world - 1) == 0, "world size must be a power of two"
        from .fieldutil import to_limb1, to_limbs

        self.ctx, self.rank, self.world, self.exchange = ctx, rank, world, exchange
        self.row_log, self.col_log = row_log, col_local_log + (world.bit_length() - 1)
        nrows, rlen = 1 << col_local_log, 1 << row_log
        n = nrows * rlen
        idx = np.arange(n, dtype=np.uint32)
        lens = np.full(nrows, rlen, dtype=np.uint32)
        pads = [(0, 0), (1, 1), (1, 1)] * 2
        self.polys = []
        for j in range(self.P):
            src = ctx.synth(seed + j, n, first_index=rank * n)
            self.polys.append(ctx.vecvec_gather(src, idx, lens, to_limb1(pads[j][0]), to_limb1(pads[j][1]), row_log, col_local_log))
            src.free()
        rng = np.random.default_rng(seed)
        from .fieldutil import R_MOD
        nv = self.row_log + self.col_log
        self.num_vars = nv
        self.point = to_limbs([int.from_bytes(rng.bytes(32), "little") % R_MOD for _ in range(nv)])
        gamma = int.from_bytes(rng.bytes(16), "little")
        self.gp = to_limbs([pow(gamma, i, R_MOD) for i in range(4)])
        self.claim = to_limb1(12345)
        self.elements = self.P * n  # table elements per GPU
        self.last = None

    def prove(self):
        so = self.ctx.deg2_vecvec_shard_so(g.GATE_PRJ_L1, self.polys, self.gp, self.point, self.col_log, self.rank, self.world)
        tr = g.Transcript(b"fgstglsp")
        out = g.sumcheck_prove_sharded_vecvec(tr, so, self.exchange, self.num_vars, self.claim)
        self.last = (out, tr.proof())
        so.destroy()
        return out


def sharded_commit(srs_slice: "g.Srs", scalars_slice: "g.Table", exchange) -> np.ndarray:
    """KzgProvingKey::commit (src/commitments/kzg.rs:123-126) split by POINT RANGE over the ranks of one box:
    every rank commits its own slice of the SRS / coefficient vector on its GPU, the G affine partial results (96 bytes
    each) are all-gathered through the shared-memory exchange and added on the host.  The group is commutative, so the
    result is the same point -- and the same canonical limbs -- as the single-GPU commitment.  exchange None: one rank."""
    local = srs_slice.msm(scalars_slice)
    if exchange is None:
        return local
    allp = exchange.allgather(local.reshape(3, 4))
    return g.g1_sum(allp.reshape(exchange.world, 12))
