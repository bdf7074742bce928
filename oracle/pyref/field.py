"""TEST INFRASTRUCTURE ONLY -- CPU oracle (big-int restatement), never imported by the product path.

Field constants and boundary encodings for the two fields on the GKR-MSM hot path.

* Fr  = BLS12-381 scalar field = Bandersnatch base field (sumcheck / GKR tables).
  Reference: `ark_bls12_381::Fr` (ark-bls12-381 0.4.0, not vendored; Cargo.lock:112-248),
  used as `F` everywhere under src/cleanup (e.g. src/cleanup/protocols/pippenger.rs:519).
* Fq  = BLS12-381 base field (G1 commitments, src/commitments/kzg.rs).

Boundary layout: an element is 4 (Fr) / 6 (Fq) little-endian u64 limbs of the
value in Montgomery form x*R mod p, R = 2^256 (Fr) / 2^384 (Fq) -- `Fp<MontBackend<_, N>, N>` of
ark-ff 0.4.2.  Inside the oracle elements are plain python ints in [0, p) (standard form); the
Montgomery map is applied only when limbs are produced/consumed.

The only literal field constant inside the reference tree is `COEFF_D` (src/utils.rs:34-37); it is
checked against the Bandersnatch curve parameter in tests/test_oracle_pins.py.
"""
from __future__ import annotations

import numpy as np

# --- BLS12-381 scalar field ---------------------------------------------------------------
FR_MODULUS = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
FR_R = (1 << 256) % FR_MODULUS
FR_R_INV = pow(1 << 256, -1, FR_MODULUS)
FR_LIMBS64 = 4

# --- BLS12-381 base field -----------------------------------------------------------------
FQ_MODULUS = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
FQ_R = (1 << 384) % FQ_MODULUS
FQ_R_INV = pow(1 << 384, -1, FQ_MODULUS)
FQ_LIMBS64 = 6

# src/utils.rs:34-37 -- Montgomery limbs of the twisted Edwards `d` of Bandersnatch.
REF_COEFF_D_MONT_LIMBS = (12167860994669987632, 4043113551995129031, 6052647550941614584, 3904213385886034240)
# Bandersnatch: a = -5, d = 138827208126141220649022263972958607803 / 171449701953573178309673572579671231137  (mod r)
TE_A = FR_MODULUS - 5
TE_D = 0x6389C12633C267CBC66E3BF86BE3B6D8CB66677177E54F92B369F2F5188D58E7

P = FR_MODULUS  # short alias used all over the oracle


def limbs_to_int(limbs) -> int:
    v = 0
    for i, l in enumerate(limbs):
        v |= int(l) << (64 * i)
    return v


def int_to_limbs(v: int, n: int):
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def fr_to_mont(x: int) -> int:
    return (x * FR_R) % FR_MODULUS


def fr_from_mont(m: int) -> int:
    return (m * FR_R_INV) % FR_MODULUS


def fq_to_mont(x: int) -> int:
    return (x * FQ_R) % FQ_MODULUS


def fq_from_mont(m: int) -> int:
    return (m * FQ_R_INV) % FQ_MODULUS


def fr_vec_to_mont_u64(vals) -> np.ndarray:
    """list of standard-form ints -> (n, 4) uint64 array of canonical Montgomery limbs (boundary layout)."""
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, x in enumerate(vals):
        m = fr_to_mont(x % FR_MODULUS)
        out[i, 0] = m & 0xFFFFFFFFFFFFFFFF
        out[i, 1] = (m >> 64) & 0xFFFFFFFFFFFFFFFF
        out[i, 2] = (m >> 128) & 0xFFFFFFFFFFFFFFFF
        out[i, 3] = (m >> 192) & 0xFFFFFFFFFFFFFFFF
    return out


def fr_vec_from_mont_u64(arr) -> list:
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    out = []
    for row in arr:
        m = int(row[0]) | (int(row[1]) << 64) | (int(row[2]) << 128) | (int(row[3]) << 192)
        assert m < FR_MODULUS, "non-canonical Montgomery limbs"
        out.append(fr_from_mont(m))
    return out


def fq_vec_to_mont_u64(vals) -> np.ndarray:
    out = np.empty((len(vals), 6), dtype=np.uint64)
    for i, x in enumerate(vals):
        m = fq_to_mont(x % FQ_MODULUS)
        for j in range(6):
            out[i, j] = (m >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def fq_vec_from_mont_u64(arr) -> list:
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 6)
    out = []
    for row in arr:
        m = limbs_to_int(row)
        assert m < FQ_MODULUS, "non-canonical Montgomery limbs"
        out.append(fq_from_mont(m))
    return out


def from_le_bytes_mod_order(b: bytes, p: int = FR_MODULUS) -> int:
    """ark_ff::PrimeField::from_le_bytes_mod_order (used by TProofTranscript2::challenge,
    src/cleanup/proof_transcript.rs:33-41)."""
    return int.from_bytes(b, "little") % p


def fr_serialize(x: int) -> bytes:
    """ark-serialize compressed Fr: 32 bytes little-endian of the canonical (non-Montgomery) value
    (src/cleanup/proof_transcript.rs:52-57 `write_scalars`)."""
    return (x % FR_MODULUS).to_bytes(32, "little")


def fr_deserialize(b: bytes) -> int:
    v = int.from_bytes(b, "little")
    assert v < FR_MODULUS
    return v


class SplitMix64:
    """Counter-based generator used for ALL synthetic inputs in this repo (tests, bench, goldens).
    The CUDA side (csrc) and the C oracle implement the same stream so inputs never cross PCIe in the
    device-resident benchmark.  value(i) = 4 successive outputs -> 256-bit LE integer -> mod r."""

    MASK = 0xFFFFFFFFFFFFFFFF

    def __init__(self, seed: int):
        self.state = seed & self.MASK

    def next(self) -> int:
        self.state = (self.state + 0x9E3779B97F4A7C15) & self.MASK
        z = self.state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & self.MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & self.MASK
        return z ^ (z >> 31)

    def fr(self) -> int:
        v = 0
        for j in range(4):
            v |= self.next() << (64 * j)
        return v % FR_MODULUS

    def below(self, n: int) -> int:
        return self.next() % n


def synth_fr(seed: int, index: int) -> int:
    """Random-access form: element `index` of stream `seed` (state jumps by 4 outputs per element)."""
    g = SplitMix64((seed + 4 * index * 0x9E3779B97F4A7C15) & SplitMix64.MASK)
    return g.fr()


def synth_fr_mont_limbs(seed: int, index: int) -> int:
    """What the device generator stores: the 256-bit draw reduced mod r is taken AS the Montgomery
    representation (saves a multiplication; any canonical residue is a valid table entry)."""
    return synth_fr(seed, index)
