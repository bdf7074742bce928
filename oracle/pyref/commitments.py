"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

Commitment-side algebra of the reference restated over python ints:
  div_by_linear, ev                      src/commitments/kzg.rs:73-81, 142-150
  KnucklesProvingKey::new / compute_t    src/commitments/knuckles.rs:65-81, 111-154
  bucket accumulation + running sums     src/cleanup/protocols/pushforward/pushforward.rs:398-429, 504-524
  bucketed_msm                           src/pullback.rs:28-59
"""
from __future__ import annotations

from . import curves as CV
from .field import P


def div_by_linear(poly, pt):
    quotient = [0] * (len(poly) - 1)
    rem = poly[-1]
    for i in range(len(quotient) - 1, -1, -1):
        quotient[i] = rem
        rem = (poly[i] + rem * pt) % P
    return quotient, rem


def ev(poly, x):
    power, acc = 1, 0
    for c in poly:
        acc = (acc + c * power) % P
        power = power * x % P
    return acc


def knuckles_inverses(num_vars, k):
    n = 1 << num_vars
    k_pows, power = [], 1
    for _ in range(2 * n - 1):
        k_pows.append(power)
        power = power * k % P
    k_n = k_pows[n - 1]
    k_pows = [(x - k_n) % P for x in k_pows]
    k_pows[n - 1] = (k_pows[n - 1] + 1) % P
    return [pow(x, -1, P) if x else 0 for x in k_pows]


def compute_t(num_vars, inverses, poly, point):
    assert len(point) == num_vars
    pt = list(reversed(point))
    n = 1 << num_vars
    assert len(poly) <= n
    t = list(poly) + [0] * (2 * n - 1 - len(poly))
    t_scaled = [0] * (2 * n - 1)
    pt_rev = [(1 - x) % P for x in pt]
    curr = n
    for i in range(num_vars):
        for idx in range(curr):
            t_scaled[idx] = t[idx] * pt_rev[i] % P
        offset = 1 << i
        curr += offset
        for idx in range(curr):
            if idx < offset:
                t[idx] = (t[idx] - t_scaled[idx]) % P
            else:
                t[idx] = (t[idx] - t_scaled[idx] + t_scaled[idx - offset]) % P
    opening = t[n - 1]
    t[n - 1] = 0
    return [x * inv % P for x, inv in zip(t, inverses)], opening


def bucket_sums(bases, point_idx, bucket_idx, n_buckets):
    out = [None] * n_buckets
    for p, b in zip(point_idx, bucket_idx):
        out[b] = CV.g1_add(out[b], bases[p])
    return out


def running_sum_commit(buckets):
    """pushforward.rs:504-524: acc = sum_{i=0}^{len-2} running_sum_i, i.e. sum_i i * B_i."""
    acc, running = None, None
    ln = len(buckets)
    for i in range(ln - 1):
        running = CV.g1_add(running, buckets[ln - i - 1])
        acc = CV.g1_add(acc, running)
    return acc


# ---- old API: binary_msm (src/binary_msm.rs:13-54) -------------------------------------------------------------------
def into_u8(bits):  # binary_msm.rs:13-17: the first bit of the chunk is the most significant
    s = 0
    for b in list(bits)[:8]:
        s = (s << 1) + (1 if b else 0)
    return s


def prepare_coefs(bits, gamma):  # binary_msm.rs:52-54
    bits = list(bits)
    return [into_u8(bits[i:i + gamma]) for i in range(0, len(bits), gamma)]


def prepare_chunk(chunk, gamma):  # binary_msm.rs:32-43: entry i - 1 = sum of chunk[len - 1 - idx] over the set bits idx of i
    out = []
    for i in range(1, 1 << gamma):
        acc = None
        for idx, b in zip(range(gamma), reversed(chunk)):
            if (1 << idx) & i:
                acc = CV.g1_add(acc, b)
        out.append(acc)
    return out


def prepare_bases(bases, gamma):  # binary_msm.rs:44-50
    return [prepare_chunk(bases[i:i + gamma], gamma) for i in range(0, len(bases), gamma)]


def binary_msm(coefs, pbases):  # binary_msm.rs:19-29
    assert len(coefs) == len(pbases)
    acc = None
    for base, idx in zip(pbases, coefs):
        if idx != 0:
            acc = CV.g1_add(acc, base[idx - 1])
    return acc
