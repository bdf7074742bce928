"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

Fiat-Shamir transcript of the reference, restated:

  src/cleanup/proof_transcript.rs:17-147  TProofTranscript2 / ProofTranscript2
      write_raw_msg  = merlin append_message(b"", msg) + proof.extend(msg)
      raw_challenge  = merlin challenge_bytes(b"", buf)
      challenge(bits)= F::from_le_bytes_mod_order(raw_challenge((bits+7)/8))
      write_scalars  = ark-serialize compressed Fr = 32 B little-endian canonical value each

merlin 3.0.0 / keccak 0.1.5 are third-party crates absent from /root/reference (Cargo.lock);
their published construction (STROBE-128 over Keccak-f[1600], "Merlin v1.0" domain separator) is
restated here.  Pin: the test vector from merlin's own documentation ("test protocol" / "some label"
/ "some data" / "challenge" -> d5a21972...0615) is asserted in tests/test_oracle_pins.py, and the
Keccak-f[1600] permutation is pinned against hashlib.sha3_256.
"""
from __future__ import annotations

from .field import FR_MODULUS, fr_deserialize, fr_serialize, from_le_bytes_mod_order

_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
_ROT = [
    [0, 36, 3, 41, 18],
    [1, 44, 10, 45, 2],
    [62, 6, 43, 15, 61],
    [28, 55, 25, 21, 56],
    [27, 20, 39, 8, 14],
]
_M = 0xFFFFFFFFFFFFFFFF


def _rol(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & _M if n else x


def keccak_f1600(state: bytearray) -> None:
    a = [[int.from_bytes(state[8 * (x + 5 * y):8 * (x + 5 * y) + 8], "little") for y in range(5)] for x in range(5)]
    for rnd in range(24):
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= _RC[rnd]
    for x in range(5):
        for y in range(5):
            state[8 * (x + 5 * y):8 * (x + 5 * y) + 8] = (a[x][y] & _M).to_bytes(8, "little")


STROBE_R = 166
FLAG_I, FLAG_A, FLAG_C, FLAG_T, FLAG_M, FLAG_K = 1, 2, 4, 8, 16, 32


class Strobe128:
    def __init__(self, protocol_label: bytes):
        st = bytearray(200)
        st[0:6] = bytes([1, STROBE_R + 2, 1, 0, 1, 96])
        st[6:18] = b"STROBEv1.0.2"
        keccak_f1600(st)
        self.state, self.pos, self.pos_begin, self.cur_flags = st, 0, 0, 0
        self.meta_ad(protocol_label, False)

    def _run_f(self):
        self.state[self.pos] ^= self.pos_begin
        self.state[self.pos + 1] ^= 0x04
        self.state[STROBE_R + 1] ^= 0x80
        keccak_f1600(self.state)
        self.pos = 0
        self.pos_begin = 0

    def _absorb(self, data: bytes):
        for b in data:
            self.state[self.pos] ^= b
            self.pos += 1
            if self.pos == STROBE_R:
                self._run_f()

    def _squeeze(self, n: int) -> bytes:
        out = bytearray(n)
        for i in range(n):
            out[i] = self.state[self.pos]
            self.state[self.pos] = 0
            self.pos += 1
            if self.pos == STROBE_R:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags: int, more: bool):
        if more:
            assert self.cur_flags == flags
            return
        assert flags & FLAG_T == 0
        old_begin = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old_begin, flags]))
        force_f = (flags & (FLAG_C | FLAG_K)) != 0
        if force_f and self.pos != 0:
            self._run_f()

    def meta_ad(self, data: bytes, more: bool):
        self._begin_op(FLAG_M | FLAG_A, more)
        self._absorb(data)

    def ad(self, data: bytes, more: bool):
        self._begin_op(FLAG_A, more)
        self._absorb(data)

    def prf(self, n: int, more: bool) -> bytes:
        self._begin_op(FLAG_I | FLAG_A | FLAG_C, more)
        return self._squeeze(n)


class MerlinTranscript:
    def __init__(self, label: bytes):
        self.strobe = Strobe128(b"Merlin v1.0")
        self.append_message(b"dom-sep", label)

    def append_message(self, label: bytes, message: bytes):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(len(message).to_bytes(4, "little"), True)
        self.strobe.ad(message, False)

    def challenge_bytes(self, label: bytes, n: int) -> bytes:
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(n.to_bytes(4, "little"), True)
        return self.strobe.prf(n, False)


class ProofTranscript2:
    """proof_transcript.rs:76-147."""

    def __init__(self, pparam: bytes, proof: bytes | None = None):
        self.merlin = MerlinTranscript(pparam)
        self.prover = proof is None
        self.proof = bytearray() if proof is None else bytes(proof)
        self.ctr = 0

    @classmethod
    def start_prover(cls, pparam: bytes):
        return cls(pparam)

    @classmethod
    def start_verifier(cls, pparam: bytes, proof: bytes):
        return cls(pparam, proof)

    def end(self) -> bytes:
        assert self.prover
        return bytes(self.proof)

    def raw_challenge(self, bytesize: int) -> bytes:
        return self.merlin.challenge_bytes(b"", bytesize)

    def write_raw_msg(self, msg: bytes):
        assert self.prover
        self.merlin.append_message(b"", msg)
        self.proof += msg

    def read_raw_msg(self, bytesize: int) -> bytes:
        assert not self.prover
        assert self.ctr + bytesize <= len(self.proof), "Out of bounds"
        msg = self.proof[self.ctr:self.ctr + bytesize]
        self.ctr += bytesize
        self.merlin.append_message(b"", msg)
        return msg

    def challenge(self, bitsize: int, p: int = FR_MODULUS) -> int:
        return from_le_bytes_mod_order(self.raw_challenge((bitsize + 7) // 8), p)

    def challenge_vec(self, n: int, bitsize: int):
        bs = (bitsize + 7) // 8
        raw = self.raw_challenge(n * bs)
        return [from_le_bytes_mod_order(raw[i * bs:(i + 1) * bs]) for i in range(n)]

    def write_scalars(self, vals):
        self.write_raw_msg(b"".join(fr_serialize(v) for v in vals))

    def read_scalars(self, n: int):
        raw = self.read_raw_msg(32 * n)
        return [fr_deserialize(raw[32 * i:32 * i + 32]) for i in range(n)]
