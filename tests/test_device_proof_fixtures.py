"""Proofs MADE ON THE B200 (tools/bench_pippenger.py --dump under gpurun; profiles/r02_config3_x23_x24_peer_pool.txt), committed as
fixtures and checked here on the CPU by the oracle VERIFIER of the whole protocol (oracle/pyref/pippenger.py::verify_pippenger):
every sumcheck round, every claim reduction, the opening equation and the pairing pair of the mock setup -- and the proved result
must be the true MSM, known in closed form for the synthetic points (tests/verify_dumped_proof.py).  The x = 24 fixture is BASELINE
config[3] itself: 2^24 points, 253-bit scalars, d_logsize 8, commitment-log-multiplicity 2 (the proof was made by one B200 with its
tables pooled over four GPUs; the 8-GPU run wrote the same bytes).  No GPU is involved in this test."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.mark.parametrize("name", ["config0_x16_d8_n128_c0", "config3_x23_d8_n253_c2", "config3_x24_d8_n253_c2"])
def test_device_made_proof_is_accepted_by_the_oracle_verifier(name):
    res = subprocess.run([sys.executable, os.path.join(HERE, "verify_dumped_proof.py"), os.path.join(HERE, "golden", "device_proofs", name + ".npz")],
                         capture_output=True, text=True, timeout=1200)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "ACCEPTED by the oracle verifier" in res.stdout and "a flipped proof bit is rejected" in res.stdout
