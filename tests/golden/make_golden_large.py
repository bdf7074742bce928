#!/usr/bin/env python3
"""Mint proof digests at the BASELINE sizes from the independent C++ prover (TEST INFRASTRUCTURE).

    python tests/golden/make_golden_large.py [--only NAME]      # rewrites tests/golden/pippenger_large.json

The python oracle (oracle/pyref) cannot reach x_logsize >= 12 in reasonable time, so the byte-level targets for BASELINE
config[0] (x=16, d=8, 128 bit), the 2^20-point shape of config[2] (x=20, d=10) and two mid sizes come from
oracle/c/pippenger_oracle.cpp -- a restatement of benchutils::run_pippenger (src/cleanup/protocols/pippenger.rs:499-559)
written from the reference sources, pinned bit-for-bit to oracle/pyref on the small golden proofs
(tests/test_pippenger_oracle.py).  This script never imports the product package (gkr-msm_b200): the digests are what an
independent CPU prover produces from the seeded inputs below, and tests/test_gpu_pippenger.py asserts that the DEVICE
prover reproduces them.  They are NOT outputs of the Rust reference (which cannot be built in this image).

Input recipe (shared with tests/test_gpu_pippenger.py::large_instance):
  rng    = numpy.random.default_rng(1000 * d + x)
  points = (k0 + i * step) G on Bandersnatch, k0 = 0x1234567 + x, step = 0x9E3779B97F4A7C15
  coefs  = rng.bytes(32 n) cut to num_bits / 8 bytes each (pippenger.rs:464-466)
  r      = y_logsize draws of rng.bytes(32) mod r;  tau = rng.bytes(32) mod r;  g0 = the G1 generator;  k = 2
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pippenger_oracle as PO  # noqa: E402
from oracle.pyref import curves as CV  # noqa: E402
from oracle.pyref.field import P, fq_vec_to_mont_u64, fr_vec_to_mont_u64  # noqa: E402

CONFIGS = [(6, 12, 128, 0), (5, 10, 64, 2), (8, 16, 128, 0), (10, 20, 128, 0), (8, 14, 253, 2),  # (8, 14, 253, 2): the shape of BASELINE config[3]
           (8, 18, 253, 2), (10, 22, 128, 0)]  # the two largest: minutes of CPU time each, compared on the device under GKR_TEST_LARGE_GOLDEN=1
STEP = 0x9E3779B97F4A7C15


def large_instance(d, x, nbits, clm):
    """seeded inputs as boundary arrays: points_xy (2, n, 4), coefs (n, 4), r limbs, tau (int), k0"""
    rng = np.random.default_rng(1000 * d + x)
    n = 1 << x
    k0 = 0x1234567 + x
    y_size = (nbits + d - 1) // d
    y_logsize = (y_size - 1).bit_length()
    raw = np.frombuffer(rng.bytes(32 * n), dtype=np.uint8).reshape(n, 32).copy()
    raw[:, nbits // 8:] = 0
    coefs = raw.view(np.uint64).reshape(n, 4)
    r = [int.from_bytes(rng.bytes(32), "little") % P for _ in range(y_logsize)]
    tau = int.from_bytes(rng.bytes(32), "little") % P
    return dict(n=n, k0=k0, coefs=coefs, r=r, tau=tau, y_logsize=y_logsize)


def name_of(d, x, nbits, clm):
    return f"pippenger_d{d}_x{x}_n{nbits}_c{clm}"


def mint(d, x, nbits, clm):
    inst = large_instance(d, x, nbits, clm)
    t0 = time.time()
    pts = PO.te_arithmetic_progression(inst["k0"], STEP, inst["n"])
    g0 = fq_vec_to_mont_u64([CV.G1_GEN[0], CV.G1_GEN[1]]).reshape(12)
    key = PO.Key(fr_vec_to_mont_u64([inst["tau"]])[0], g0, x + clm, fr_vec_to_mont_u64([2])[0])
    t1 = time.time()
    out = PO.run_pippenger(key, pts, inst["coefs"], fr_vec_to_mont_u64(inst["r"]), d, x, nbits, clm)
    t2 = time.time()
    key.close()
    print(f"{name_of(d, x, nbits, clm)}: setup {t1 - t0:.1f} s, prove {t2 - t1:.1f} s on {PO.num_threads()} threads, proof {len(out['proof'])} bytes", flush=True)
    return {
        "d_logsize": d, "x_logsize": x, "num_bits": nbits, "commitment_log_multiplicity": clm,
        "rng": "numpy.random.default_rng(1000 * d + x)", "k0": inst["k0"], "step": STEP,
        "proof_len": len(out["proof"]), "proof_sha256": hashlib.sha256(out["proof"]).hexdigest(),
        "dense_output_sha256": hashlib.sha256(out["dense_output"].tobytes()).hexdigest(),
        "claim_evs_sha256": hashlib.sha256(out["claim_evs"].tobytes()).hexdigest(),
        "pair_sha256": hashlib.sha256(out["pair"].tobytes()).hexdigest(),
        "minted_by": "oracle/c/pippenger_oracle.cpp (independent C++ prover), NOT the Rust reference",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    path = os.path.join(HERE, "pippenger_large.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    for cfg in CONFIGS:
        nm = name_of(*cfg)
        if args.only and args.only != nm:
            continue
        data[nm] = mint(*cfg)
        with open(path, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
            f.write("\n")


if __name__ == "__main__":
    main()
