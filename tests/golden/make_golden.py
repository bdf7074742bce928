#!/usr/bin/env python3
"""Mint the golden fixtures of tests/golden/ from the oracle (TEST INFRASTRUCTURE).

The reference holds no golden vectors or known-answer tests for this path and cannot be built in
this image (nightly Rust + un-vendored git dependencies), so these fixtures are minted from the python restatement
(oracle/pyref) -- they pin the ORACLE against regressions and give the GPU parity tests a byte-level target that does not
depend on re-running the slow python prover.  They are NOT outputs of the Rust reference: byte-level parity against a
real reference run stays "parity unpinned" (DESIGN.md section 4).

    python tests/golden/make_golden.py            # rewrites tests/golden/*.json, *.proof

Fixtures:
  pippenger_d{d}_x{x}_n{nbits}_c{clm}.proof   serialized proof bytes of examples/pippenger's prover
                                              (src/cleanup/protocols/pippenger.rs:122-294) on seeded inputs
  pippenger.json                              per config: seed recipe, sha256 of the proof, output tables, claims, the MSM
  dense_sumcheck.json                         DenseSumcheckObjectSO round messages (protocols/sumcheck.rs:277-332) for
                                              Prod3 on generator-seeded tables, with 128-bit challenges
  msm_g1.json                                 KzgProvingKey::commit (commitments/kzg.rs:123-126) of seeded scalars over a
                                              mock SRS, affine result
  eq_table.json                               eq_poly_sequence_last (src/utils.rs:252-262) on a seeded point
"""
import hashlib
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.pyref import curves as CV  # noqa: E402
from oracle.pyref import pippenger as PP  # noqa: E402
from oracle.pyref import sumcheck as S  # noqa: E402
from oracle.pyref import gates as G  # noqa: E402
from oracle.pyref.field import P  # noqa: E402
from oracle.pyref.transcript import ProofTranscript2  # noqa: E402

PIPPENGER_CONFIGS = [(2, 3, 6, 0), (2, 3, 8, 1), (3, 4, 7, 0), (2, 2, 8, 2)]


def hx(v):
    return "%064x" % v


def pippenger_instance(d, x, nbits, clm):
    """the seeded instance of tests/test_oracle_pippenger.make_instance (python `random.Random`, fixed consumption order)"""
    rng = random.Random(1000 * d + 100 * x + 10 * nbits + clm)
    cfg = PP.pippenger_config(d, x, nbits, clm)
    points = [CV.te_random_point(rng) for _ in range(1 << x)]
    coefs = [rng.randrange(1 << nbits) for _ in range(1 << x)]
    r = [rng.randrange(P) for _ in range(cfg["y_logsize"])]
    tau = rng.randrange(1, P)
    g0 = CV.g1_mul(rng.randrange(1, P), CV.G1_GEN)
    key = PP.KnucklesKey(PP.KzgKey(tau, g0, 2 * (1 << (x + clm)) - 1), x + clm, 2)
    return cfg, points, coefs, r, key


def mint_pippenger():
    out = {}
    for d, x, nbits, clm in PIPPENGER_CONFIGS:
        cfg, points, coefs, r, key = pippenger_instance(d, x, nbits, clm)
        tp = ProofTranscript2.start_prover(b"fgstglsp")
        dense, claims = PP.run_pippenger(tp, points, coefs, cfg, r, key)
        proof = tp.end()
        name = f"pippenger_d{d}_x{x}_n{nbits}_c{clm}"
        with open(os.path.join(HERE, name + ".proof"), "wb") as f:
            f.write(proof)
        res = CV.te_msm(points, coefs)
        out[name] = {
            "d_logsize": d, "x_logsize": x, "num_bits": nbits, "commitment_log_multiplicity": clm,
            "seed": 1000 * d + 100 * x + 10 * nbits + clm,
            "points": [[hx(p[0]), hx(p[1])] for p in points], "coefs": [hx(c) for c in coefs], "r": [hx(v) for v in r],
            "tau": hx(key.kzg.tau), "g0": ["%096x" % key.kzg.g0[0], "%096x" % key.kzg.g0[1]],
            "proof_len": len(proof), "proof_sha256": hashlib.sha256(proof).hexdigest(),
            "dense_output": [[hx(v) for v in t] for t in dense],
            "claims_point": [hx(v) for v in claims[0]], "claims_evs": [hx(v) for v in claims[1]],
            "msm_result": [hx(res[0]), hx(res[1])],
        }
    with open(os.path.join(HERE, "pippenger.json"), "w") as f:
        json.dump(out, f, indent=1)


def synth_table(seed, n):
    """the counter-based table generator shared by the device (gkr_table_synth), the C oracle (oracle_synth_table) and this
    file: SplitMix64 stream -> 4 limbs -> top limb masked to 62 bits... taken from the C oracle so the three agree"""
    from oracle import coracle
    from oracle.pyref.field import fr_vec_from_mont_u64
    return fr_vec_from_mont_u64(coracle.synth_table(seed, n))


def mint_dense():
    out = {}
    rng = random.Random(77)
    for name, nv, gate_name in [("prod3_n6", 6, "prod3"), ("prod3_n9", 9, "prod3")]:
        tabs = [synth_table(100 + j, 1 << nv) for j in range(3)]
        claim = sum(a * b % P * c for a, b, c in zip(*tabs)) % P
        so = S.DenseSumcheckObjectSO([list(t) for t in tabs], G.Prod3(), nv, claim)
        chals, msgs = [], []
        for _ in range(nv):
            u = so.unipoly()
            msgs.append({"coeffs": [hx(c) for c in u], "evals_0_to_deg": [hx(c) for c in so.last_evals]})
            t = rng.randrange(1 << 128)
            chals.append(hx(t))
            so.bind(t)
        out[name] = {"gate": gate_name, "num_vars": nv, "table_seeds": [100, 101, 102], "claim": hx(claim), "challenges": chals,
                     "rounds": msgs, "final_evals": [hx(v) for v in so.final_evals()]}
    with open(os.path.join(HERE, "dense_sumcheck.json"), "w") as f:
        json.dump(out, f, indent=1)


def mint_msm_eq():
    rng = random.Random(4242)
    tau = rng.randrange(1, P)
    g0 = CV.g1_mul(rng.randrange(1, P), CV.G1_GEN)
    n = 48
    key = PP.KzgKey(tau, g0, n)
    poly = [rng.randrange(P) for _ in range(n - 3)] + [0, 1, P - 1]
    c1 = key.commit_literal(poly)
    small = [rng.randrange(1 << 16) for _ in range(n)]
    c2 = key.commit_literal(small)
    with open(os.path.join(HERE, "msm_g1.json"), "w") as f:
        json.dump({"tau": hx(tau), "g0": ["%096x" % g0[0], "%096x" % g0[1]], "n": n,
                   "full_scalars": [hx(v) for v in poly], "full_commit": ["%096x" % c1[0], "%096x" % c1[1]],
                   "full_commit_bytes": PP.g1_serialize(c1).hex(),
                   "small_scalars": [hx(v) for v in small], "small_commit": ["%096x" % c2[0], "%096x" % c2[1]],
                   "small_commit_bytes": PP.g1_serialize(c2).hex()}, f, indent=1)
    pt = [rng.randrange(P) for _ in range(7)]
    eq = S.eq_poly_sequence_last(pt)
    with open(os.path.join(HERE, "eq_table.json"), "w") as f:
        json.dump({"point": [hx(v) for v in pt], "sha256_of_be_hex_lines": hashlib.sha256("\n".join(hx(v) for v in eq).encode()).hexdigest(),
                   "first8": [hx(v) for v in eq[:8]], "last": hx(eq[-1]), "sum": hx(sum(eq) % P)}, f, indent=1)


if __name__ == "__main__":
    mint_pippenger()
    mint_dense()
    mint_msm_eq()
    print("golden fixtures written to", HERE)
