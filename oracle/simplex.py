"""TEST INFRASTRUCTURE (checker only; see oracle/__init__.py): CPU restatement of the simplex-noise path of the
reconstruction loop, SURVEY.md §8 f-2.

Reference: src/utils/simplex_noise.py - `generate_simplex_noise` (:15-79) draws one seed per (channel, image), builds the
permutation tables (`_init`, :559-577) and fills noise[j, i] with `rand_3d_fixed_T_octaves` (:141-159): 6 octaves of 3-D
OpenSimplex noise (`_noise3`, :704-1271; K. Spencer's public-domain OpenSimplex, via lmas/opensimplex and AnoDDPM) on
the plane z = t / frequency. Called from src/trainers/reconstruct.py:133-139 when --simplex_noise=1.

Restated, not transcribed: the reference spells out every lattice case; here the eight candidate lattice points of a
cell are produced by a small selection rule (three unit axes, "near" and "far" vertices, mirror symmetry between the two
tetrahedra) and one uniform contribution loop evaluates them. Every floating-point expression keeps the reference's
operand order, so the result is bit-identical in fp64.

PINNED: tests/test_simplex.py runs the reference's own `_noise3` / `_init` / `rand_3d_fixed_T_octaves` here (numba
stubbed out so the same Python executes un-jitted) against this file when /root/reference is present, and the committed
golden vectors tests/golden/simplex_*.npz (made by tests/golden/make_simplex_golden.py from the reference) everywhere."""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

STRETCH = -1.0 / 6  # (1 / sqrt(3 + 1) - 1) / 3
SQUISH = 1.0 / 3    # (sqrt(3 + 1) - 1) / 3
NORM = 103
_MUL, _INC = 6364136223846793005, 1442695040888963407
_M64 = (1 << 64) - 1


def _wrap(v: int) -> int:
    """two's-complement int64 of a Python int (the reference's c_int64(x).value)"""
    v &= _M64
    return v - (1 << 64) if v >> 63 else v


def tables(seed: int) -> Tuple[np.ndarray, np.ndarray]:
    """(perm, gradient id) of `_init`: an LCG-driven draw without replacement from 0..255; gradient id = perm % 24."""
    perm = np.zeros(256, dtype=np.int64)
    source = list(range(256))
    for _ in range(3):
        seed = _wrap(seed * _MUL + _INC)
    for i in range(255, -1, -1):
        seed = _wrap(seed * _MUL + _INC)
        r = (seed + 31) % (i + 1)  # Python's floor-mod: never negative
        perm[i] = source[r]
        source[r] = source[i]
    return perm, perm % 24


def gradient(g: int) -> Tuple[int, int, int]:
    """Gradient g of the 24 (the reference's GRADIENTS3 table, :191-266): 11 on axis g % 3, 4 on the others; signs from
    the bits of g // 3 (x positive iff bit 0, y negative iff bit 1, z negative iff bit 2)."""
    blk, axis = divmod(g, 3)
    sign = (1 if blk & 1 else -1, -1 if blk & 2 else 1, -1 if blk & 4 else 1)
    return tuple(sign[a] * (11 if a == axis else 4) for a in range(3))


def _unit(axis: int, scale: int = 1) -> List[int]:
    v = [0, 0, 0]
    v[axis] = scale
    return v


def candidates(xins: float, yins: float, zins: float):
    """The (up to) eight lattice points of the cell that may contribute, in the reference's summation order, as
    (i, j, k, squish_first). squish_first marks the one candidate whose offset the reference subtracts after the squish
    term (a rounding-order detail)."""
    s = (xins, yins, zins)
    in_sum = xins + yins + zins
    if in_sum <= 1:  # tetrahedron at the origin (:734-860)
        a, b = 0, 1
        if s[a] >= s[b] and zins > s[b]:
            b = 2
        elif s[a] < s[b] and zins > s[a]:
            a = 2
        w = 1 - in_sum
        if w > s[a] or w > s[b]:  # the origin is one of the two closest vertices
            c = b if s[b] > s[a] else a
            lo, hi = [ax for ax in range(3) if ax != c]
            e0, e1 = _unit(c), _unit(c)
            e0[lo] -= 1
            e1[hi] -= 1
        else:
            k = 3 - a - b
            e0 = [1, 1, 1]
            e0[k] = 0
            e1 = list(e0)
            e1[k] = -1
        base = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)]
        return [(*v, False) for v in base] + [(*e0, False), (*e1, False)]
    if in_sum >= 2:  # tetrahedron at (1,1,1): the mirror image (:861-1009)
        a, b = 0, 1
        if s[a] <= s[b] and zins < s[b]:
            b = 2
        elif s[a] > s[b] and zins < s[a]:
            a = 2
        w = 3 - in_sum
        if w < s[a] or w < s[b]:  # (1,1,1) is one of the two closest vertices
            c = b if s[b] < s[a] else a
            lo, hi = [ax for ax in range(3) if ax != c]
            e0, e1 = [1, 1, 1], [1, 1, 1]
            e0[c] = e1[c] = 0
            e0[lo] += 1
            e1[hi] += 1
        else:
            k = 3 - a - b
            e0, e1 = _unit(k), _unit(k, 2)
        base = [(1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1)]
        return [(*v, False) for v in base] + [(*e0, False), (*e1, False)]
    # octahedron in between (:1010-1246): per pair of opposite vertices keep the nearer one, then the best two
    def pick(p, far_pt, near_pt):
        return (p - 1, far_pt, True) if p > 1 else (1 - p, near_pt, False)

    a_score, a_pt, a_far = pick(xins + yins, (1, 1, 0), (0, 0, 1))
    b_score, b_pt, b_far = pick(xins + zins, (1, 0, 1), (0, 1, 0))
    score, c_pt, c_far = pick(yins + zins, (0, 1, 1), (1, 0, 0))
    if a_score <= b_score and a_score < score:
        a_pt, a_far = c_pt, c_far
    elif a_score > b_score and b_score < score:
        b_pt, b_far = c_pt, c_far
    first_set = lambda v: 0 if v[0] else (1 if v[1] else 2)      # noqa: E731
    first_clear = lambda v: 0 if not v[0] else (1 if not v[1] else 2)  # noqa: E731
    flip = lambda k: [-1 if ax == k else 1 for ax in range(3)]   # noqa: E731  a permutation of (-1, 1, 1)
    if a_far == b_far:
        if a_far:
            e0 = [1, 1, 1]
            e1 = _unit(first_set([x & y for x, y in zip(a_pt, b_pt)]), 2)
        else:
            e0 = [0, 0, 0]
            e1 = flip(first_clear([x | y for x, y in zip(a_pt, b_pt)]))
        e1_flag = False
    else:
        far_pt, near_pt = (a_pt, b_pt) if a_far else (b_pt, a_pt)
        e0 = flip(first_clear(far_pt))
        e1 = _unit(first_set(near_pt), 2)
        e1_flag = True
    base = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1)]
    return [(*v, False) for v in base] + [(*e0, False), (*e1, e1_flag)]


def noise3(x: float, y: float, z: float, perm: Sequence[int], grad_id: Sequence[int]) -> float:
    """3-D OpenSimplex value at (x, y, z): sum over the candidates of max(0, 2 - |d|^2)^4 * (gradient . d), / 103."""
    stretch = (x + y + z) * STRETCH
    xs, ys, zs = x + stretch, y + stretch, z + stretch
    xsb, ysb, zsb = math.floor(xs), math.floor(ys), math.floor(zs)
    squish = (xsb + ysb + zsb) * SQUISH
    dx0, dy0, dz0 = x - (xsb + squish), y - (ysb + squish), z - (zsb + squish)
    value = 0.0
    for i, j, k, squish_first in candidates(xs - xsb, ys - ysb, zs - zsb):
        sq = (i + j + k) * SQUISH
        if squish_first:
            dx, dy, dz = dx0 - sq - i, dy0 - sq - j, dz0 - sq - k
        else:
            dx, dy, dz = dx0 - i - sq, dy0 - j - sq, dz0 - k - sq
        attn = 2 - dx * dx - dy * dy - dz * dz
        if attn > 0:
            g = gradient(int(grad_id[(perm[(perm[(xsb + i) & 0xFF] + ysb + j) & 0xFF] + zsb + k) & 0xFF]))
            attn *= attn
            value += attn * attn * (g[0] * dx + g[1] * dy + g[2] * dz)
    return value / NORM


def fractal_fixed_t(shape: Tuple[int, int], t: float, perm, grad_id, octaves: int = 6, persistence: float = 0.8,
                    frequency: float = 64) -> np.ndarray:
    """`rand_3d_fixed_T_octaves` (:141-159): fp64 [H, W]; pixel (y, x) samples (x / f, y / f, t / f) per octave."""
    h, w = shape
    out = np.zeros((h, w))
    amp = 1
    for _ in range(octaves):
        for yy in range(h):
            for xx in range(w):
                out[yy, xx] += amp * noise3(xx / frequency, yy / frequency, t / frequency, perm, grad_id)
        frequency /= 2
        amp *= persistence
    return out


def simplex_noise(seeds: np.ndarray, t: Sequence[int], shape: Tuple[int, int], octaves: int = 6,
                  persistence: float = 0.8, frequency: float = 64) -> np.ndarray:
    """`generate_simplex_noise` for given seeds [C, B] (drawn channel-outer, image-inner like the reference): fp32
    [B, C, H, W]."""
    c, b = seeds.shape
    out = np.empty((b, c) + tuple(shape), dtype=np.float32)
    for i in range(c):
        for j in range(b):
            perm, gid = tables(int(seeds[i, j]))
            out[j, i] = fractal_fixed_t(shape, int(t[j]), perm, gid, octaves, persistence, frequency).astype(np.float32)
    return out
