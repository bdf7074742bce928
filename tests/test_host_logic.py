"""CPU tests of the host side of the product: the C-ABI library loads and exports every symbol the
header declares, the C++ merlin transcript matches the oracle / published vector, and a context
refuses to be created without a GPU (no CPU fallback)."""
import os
import random
import re

import numpy as np
import pytest

import gkr_msm_b200 as g
from oracle.pyref.field import P
from oracle.pyref.transcript import ProofTranscript2
from tests.util import from_limbs, to_limbs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = g.load_library()
    hdr = open(os.path.join(ROOT, "include", "gkr_msm_b200.h")).read()
    names = set(re.findall(r"\b(gkr_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    for n in sorted(names):
        assert hasattr(lib, n), f"symbol {n} declared in include/gkr_msm_b200.h is not exported"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(g.GkrError) as e:
        g.Context(0)
    assert e.value.code == g.GKR_ERR_CUDA


def test_cpp_transcript_matches_oracle_and_merlin_vector():
    rng = random.Random(1)
    t = g.Transcript(b"fgstglsp")
    o = ProofTranscript2.start_prover(b"fgstglsp")
    for step in range(6):
        xs = [rng.randrange(P) for _ in range(1 + step % 3)]
        t.write_scalars(to_limbs(xs))
        o.write_scalars(xs)
        bits = (128, 512, 8, 255)[step % 4]
        assert from_limbs(t.challenge(bits).reshape(1, 4))[0] == o.challenge(bits)
    t.write_raw(b"\x01\x02\x03" * 100)  # crosses the 166-byte STROBE rate
    o.write_raw_msg(b"\x01\x02\x03" * 100)
    assert t.raw_challenge(200) == o.raw_challenge(200)
    assert t.proof() == o.end()


def test_cpp_transcript_rejects_non_canonical_scalar():
    t = g.Transcript(b"x")
    bad = np.array([[0xFFFFFFFFFFFFFFFF] * 4], dtype=np.uint64)
    with pytest.raises(g.GkrError):
        t.write_scalars(bad)
