#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (CPU only, imports oracle/): verify a proof written by `tools/bench_pippenger.py --dump FILE.npz` on a GPU box.

    python tests/verify_dumped_proof.py gpurun_out/cfg3_x24.npz

Regenerates the scalars from the seed (the bench's recipe), computes the true MSM in closed form -- the synthetic points are
(k0 + i step) G, so sum_i c_i P_i = (sum_i c_i (k0 + i step)) G -- and runs the oracle VERIFIER of the whole protocol
(oracle/pyref/pippenger.py::verify_pippenger: every sumcheck round, every claim reduction, the opening equation, and the pairing
check A == tau B of the mock setup) on the dumped proof bytes, output tables and claims.  Used for the instances that are too
large to verify while a multi-GPU box is held (x = 23 / 24 of BASELINE config[3]): minutes of python per million points."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyref import curves as CV  # noqa: E402
from oracle.pyref import pippenger as PP  # noqa: E402
from oracle.pyref.field import P, fq_vec_from_mont_u64, fr_vec_from_mont_u64  # noqa: E402
from oracle.pyref.transcript import ProofTranscript2  # noqa: E402


def main():
    d = np.load(sys.argv[1])
    x, dl, nbits, clm, seed = [int(v) for v in d["meta"]]
    n = 1 << x
    t0 = time.time()
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    b = raw.view(np.uint8).reshape(n, 32).copy()
    b[:, nbits // 8:] = 0
    # sum_i c_i (k0 + i step) = k0 * sum_i c_i + step * sum_i i c_i, exactly, with numpy: the scalars as 16-bit digits (16 per scalar),
    # the index range in blocks of 2^20 so that every partial sum stays below 2^63
    k0, step = 0x1234567 + seed, 0x9E3779B97F4A7C15
    q = b.view(np.uint16).reshape(n, 16).astype(np.uint64)
    s0, s1 = 0, 0
    blk = 1 << 20
    for lo in range(0, n, blk):
        hi = min(n, lo + blk)
        part = q[lo:hi]
        idx = np.arange(hi - lo, dtype=np.uint64)[:, None]
        col = part.sum(axis=0)             # < 2^16 * 2^20
        icol = (part * idx).sum(axis=0)    # < 2^16 * 2^20 * 2^20
        for j in range(16):
            cj, ij = int(col[j]) << (16 * j), int(icol[j]) << (16 * j)
            s0 += cj
            s1 += ij + lo * cj
    total = k0 * s0 + step * s1
    expected = CV.te_mul(total % CV.TE_SUBGROUP_ORDER, CV.TE_GEN)
    print(f"closed-form MSM over {n} scalars: {time.time() - t0:.1f} s", flush=True)
    cfg = PP.pippenger_config(dl, x, nbits, clm)
    r = fr_vec_from_mont_u64(d["r"])
    tau = fr_vec_from_mont_u64(d["tau"])[0]
    nv = x + clm
    # the VERIFIER's key: KnucklesProvingKey::new also tabulates 2^(nv+1) inverses for compute_t (prover side only; hours of python
    # at nv = 26), so the object is assembled without them
    okey = PP.KnucklesKey.__new__(PP.KnucklesKey)
    okey.kzg, okey.num_vars, okey.k, okey.inverses = PP.KzgKey(tau, CV.G1_GEN, 2 * (1 << nv) - 1), nv, 2, None
    dense_output = [fr_vec_from_mont_u64(t) for t in d["dense"]]
    evs = fr_vec_from_mont_u64(d["evs"])
    proof = d["proof"].tobytes()
    t0 = time.time()
    tv = ProofTranscript2.start_verifier(b"fgstglsp", proof)
    got = PP.verify_pippenger(tv, cfg, dense_output, (list(r), list(evs)), okey, expected)
    assert tv.ctr == len(proof), "the verifier did not consume the whole proof"
    assert got == expected, "the proved result differs from the true MSM"

    def g1(xy):
        a, bb = fq_vec_from_mont_u64(np.asarray(xy).reshape(2, 6))
        return None if (a, bb) == (0, 0) else (a, bb)

    okey.kzg.verify_pair((g1(d["pair"][0]), g1(d["pair"][1])))
    print(f"x_logsize {x}, d_logsize {dl}, {nbits}-bit scalars, clm {clm}: proof of {len(proof)} bytes ACCEPTED by the oracle verifier, "
          f"result == closed-form MSM, pairing pair A == tau B  ({time.time() - t0:.1f} s)", flush=True)
    bad = bytearray(proof)
    bad[len(bad) // 2] ^= 1
    try:
        PP.verify_pippenger(ProofTranscript2.start_verifier(b"fgstglsp", bytes(bad)), cfg, dense_output, (list(r), list(evs)), okey, expected)
    except AssertionError:
        print("a flipped proof bit is rejected", flush=True)
    else:
        raise SystemExit("a corrupted proof was accepted")


if __name__ == "__main__":
    main()
