"""Shared helpers for the parity tests: oracle ints <-> boundary limbs."""
import random

import numpy as np

from oracle.pyref.field import P, fr_vec_from_mont_u64, fr_vec_to_mont_u64


def to_limbs(vals):
    return fr_vec_to_mont_u64(list(vals))


def to_limb1(v):
    return fr_vec_to_mont_u64([v])[0]


def from_limbs(arr):
    return fr_vec_from_mont_u64(arr)


def rand_table(rng: random.Random, n: int):
    return [rng.randrange(P) for _ in range(n)]


def rand_chal128(rng: random.Random):
    return rng.randrange(1 << 128)
