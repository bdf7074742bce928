"""Tiny host-side helpers for O(1) protocol glue (claims, gamma powers, split challenges): python ints in
standard form <-> the boundary's canonical Montgomery limbs.  Nothing table-sized goes through here."""
from __future__ import annotations

import numpy as np

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
_R = (1 << 256) % R_MOD
_RINV = pow(1 << 256, -1, R_MOD)
_M64 = 0xFFFFFFFFFFFFFFFF


def to_limbs(vals) -> np.ndarray:
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        m = (v % R_MOD) * _R % R_MOD
        out[i] = [m & _M64, (m >> 64) & _M64, (m >> 128) & _M64, (m >> 192) & _M64]
    return out


def to_limb1(v) -> np.ndarray:
    return to_limbs([v])[0]


def from_limbs(arr) -> list:
    a = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    return [((int(r[0]) | (int(r[1]) << 64) | (int(r[2]) << 128) | (int(r[3]) << 192)) * _RINV) % R_MOD for r in a]


def make_gamma_pows(gamma: int, count: int) -> list:
    """src/utils.rs:126-135 (always at least [1, gamma])"""
    g = [1, gamma % R_MOD]
    for i in range(2, count):
        g.append(g[i - 1] * gamma % R_MOD)
    return g
