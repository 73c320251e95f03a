"""Philox4x32 counter-based generator (10 rounds; 7 for the field stream) in numpy -- TEST INFRASTRUCTURE ONLY.

The reference draws from unseeded global generators (numpy MT19937:
PPOV2.1/environment.py:44,58,60,108; torch: train_ppo2.0.py:43,162), which cannot be
reproduced on a device and are not part of its contract.  The CUDA path instead uses
Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; the
Random123 algorithm, also what cuRAND/torch-CUDA use) keyed by a 64-bit seed and counted
by (index, episode, global env id, stream tag), so results do not depend on how envs are
sharded over GPUs.  This file restates the published algorithm so that the integer side
of the device generator can be checked bit-exactly (known-answer vectors of Random123's
``kat_vectors`` are in ``tests/test_philox.py``).

Stream layout (must match ``csrc/plume_core.h``):

=========  =====================================================================
tag        counter = (c0, c1, c2, c3)
=========  =====================================================================
TAG_SRC=1  (0, episode, env, 1): words 0,1 -> u_x (53-bit double), words 2,3 -> u_y
TAG_FIELD=2 (cell>>2, episode, env, 2), Philox4x32-7 (FIELD_ROUNDS): one call per FOUR cells, cell = x*G + y;
           word 0: [31..12] radius uniform (m+1) 2^-20 of cells 4q, 4q+1, [11..0] u = m 2^-12 of cell 4q
           word 1: [31..12] radius uniform of cells 4q+2, 4q+3,          [11..0] u of cell 4q+1
           word 2: [15..0] angle uniform m 2^-16 of the first pair, [31..16] of the second pair
           word 3: [11..0] u of cell 4q+2, [23..12] u of cell 4q+3
           Box-Muller per pair: z_even = r cos(2 pi a), z_odd = r sin(2 pi a)
TAG_STEP=3 (step, episode, env, 3): Box-Muller(words 0,1) -> the two randn of one step;
           step = step_count BEFORE the step (0 for the first step of an episode)
TAG_ACT=4  (step, episode, env, 4): word 0 -> uniform for the inverse-CDF action draw (same step index)
TAG_WIND=5 (0, episode, env, 5): words 0,1 -> wind direction phi = 2 pi u53, words 2,3 -> speed 1 + 4 u53
           (README dispersion plume only)
=========  =====================================================================
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

TAG_SRC, TAG_FIELD, TAG_STEP, TAG_ACT, TAG_WIND = 1, 2, 3, 4, 5


FIELD_ROUNDS = 7      # the field stream's reduced-round variant (Random123's Crush-resistant minimum); others: 10


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    return philox4x32(c0, c1, c2, c3, k0, k1, 10)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds: int = 10):
    """Vectorised over equally shaped integer arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def seed_key(seed: int):
    return seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF


def uniform24(r):
    """float32 in [0,1): top 24 bits."""
    return ((np.asarray(r, dtype=np.uint32) >> np.uint32(8)).astype(np.float32)) * np.float32(2.0 ** -24)


def uniform24_open0(r):
    """float32 in (0,1]: (top 24 bits + 1) * 2^-24, the radius argument of Box-Muller."""
    return ((np.asarray(r, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) \
        * np.float32(2.0 ** -24)


def uniform53(r_hi, r_lo):
    """double in [0,1) from two words (27 + 26 bits, the numpy/MT 'genrand_res53' recipe)."""
    a = (np.asarray(r_hi, dtype=np.uint32) >> np.uint32(5)).astype(np.float64)
    b = (np.asarray(r_lo, dtype=np.uint32) >> np.uint32(6)).astype(np.float64)
    return (a * 67108864.0 + b) / 9007199254740992.0


def box_muller64(r0, r1):
    """Reference-precision Box-Muller of the two words (the device evaluates the same
    formula in float32 with fast intrinsics; compare with an absolute tolerance)."""
    u1 = uniform24_open0(r0).astype(np.float64)
    u2 = uniform24(r1).astype(np.float64)
    rad = np.sqrt(-2.0 * np.log(u1))
    return rad * np.cos(2 * np.pi * u2), rad * np.sin(2 * np.pi * u2)


def source_uniforms(seed: int, env, episode):
    k0, k1 = seed_key(seed)
    r = philox4x32_10(0, episode, env, TAG_SRC, k0, k1)
    return uniform53(r[0], r[1]), uniform53(r[2], r[3])


def field_words(seed: int, env, episode, cell):
    """The four words of the quad a cell belongs to."""
    k0, k1 = seed_key(seed)
    return philox4x32(np.asarray(cell) >> 2, episode, env, TAG_FIELD, k0, k1, FIELD_ROUNDS)


def field_noise64(seed: int, env, episode, cell):
    """(z, u) of a cell in float64 Box-Muller precision; u is exact."""
    cell = np.asarray(cell)
    r = [np.asarray(w, dtype=np.uint32) for w in field_words(seed, env, episode, cell)]
    k = cell & 3
    second = k >= 2
    radius = np.where(second, r[1] >> np.uint32(12), r[0] >> np.uint32(12)).astype(np.float64)
    angle = np.where(second, r[2] >> np.uint32(16), r[2] & np.uint32(0xFFFF)).astype(np.float64)
    u1 = (radius + 1.0) * 2.0 ** -20
    u2 = angle * 2.0 ** -16
    rad = np.sqrt(-2.0 * np.log(u1))
    z = np.where((k & 1).astype(bool), rad * np.sin(2 * np.pi * u2), rad * np.cos(2 * np.pi * u2))
    m = np.select([k == 0, k == 1, k == 2], [r[0], r[1], r[3]], r[3] >> np.uint32(12)) & np.uint32(0xFFF)
    return z, (m.astype(np.float32) * np.float32(2.0 ** -12))


def step_noise64(seed: int, env, episode, step):
    k0, k1 = seed_key(seed)
    r = philox4x32_10(step, episode, env, TAG_STEP, k0, k1)
    return box_muller64(r[0], r[1])


def action_uniform(seed: int, env, episode, step):
    k0, k1 = seed_key(seed)
    r = philox4x32_10(step, episode, env, TAG_ACT, k0, k1)
    return uniform24(r[0])


def wind(seed: int, env, episode):
    """(cos phi, sin phi, speed) of the README dispersion plume for (env, episode)."""
    k0, k1 = seed_key(seed)
    r = philox4x32_10(0, episode, env, TAG_WIND, k0, k1)
    phi = 6.283185307179586 * uniform53(r[0], r[1])
    return np.cos(phi), np.sin(phi), 1.0 + 4.0 * uniform53(r[2], r[3])
