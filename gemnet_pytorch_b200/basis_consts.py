"""Host-side constants of the basis kernels (product code: no scipy / sympy / oracle imports).

The reference bakes these numbers into sympy-generated formulas:
  * zeros of the spherical Bessel functions, found by bracketing + root finding and ROUNDED TO FLOAT32
    (gemnet/model/layers/basis_utils.py:14-29),
  * normalisers 1/sqrt(0.5 j_{l+1}(z_ln)^2) evaluated at the float32-rounded zero (basis_utils.py:57-66),
  * spherical-harmonic prefactors sqrt((2l+1)/(4pi) (l-|m|)!/(l+|m|)!) (basis_utils.py:83-104) with the
    sqrt(2) (-1)^m factor of the real harmonics (basis_utils.py:223-243).
They are computed here in float64 with a closed-form j_l and plain bisection.
"""
import math
from functools import lru_cache

import numpy as np


def _sph_jn(l, t):
    """j_l(t), t >= ~3 (upward recurrence is stable there)."""
    s, c = math.sin(t), math.cos(t)
    j0 = s / t
    if l == 0:
        return j0
    j1 = (s / t - c) / t
    for k in range(1, l):
        j0, j1 = j1, (2 * k + 1) / t * j1 - j0
    return j1


def _bisect_root(f, a, b):
    fa, fb = f(a), f(b)
    if fa == 0.0:
        return a
    if fb == 0.0:
        return b
    assert fa * fb < 0, "root not bracketed"
    for _ in range(200):
        m = 0.5 * (a + b)
        fm = f(m)
        if fm == 0.0 or (b - a) < 1e-15 * max(1.0, abs(m)):
            return m
        if fa * fm < 0:
            b, fb = m, fm
        else:
            a, fa = m, fm
    return 0.5 * (a + b)


@lru_cache(maxsize=None)
def bessel_zeros(num_spherical, num_radial):
    """(num_spherical, num_radial) float32: zeros of j_l interlace those of j_{l-1} (basis_utils.py:14-29)."""
    n, k = num_spherical, num_radial
    zeros = np.zeros((n, k), dtype=np.float32)
    zeros[0] = np.arange(1, k + 1) * np.pi
    points = np.arange(1, k + n) * np.pi
    roots = np.zeros(k + n - 1, dtype=np.float32)
    for l in range(1, n):
        for j in range(k + n - 1 - l):
            roots[j] = _bisect_root(lambda t: _sph_jn(l, t), float(points[j]), float(points[j + 1]))
        points = roots  # float32-rounded brackets for the next order, as the reference
        zeros[l, :k] = roots[:k]
    return zeros


@lru_cache(maxsize=None)
def bessel_tables(num_spherical, num_radial):
    """(zeros, normalisers) as float64 arrays of shape (num_spherical, num_radial)."""
    z = bessel_zeros(num_spherical, num_radial).astype(np.float64)
    norm = np.zeros_like(z)
    for l in range(num_spherical):
        for i in range(num_radial):
            norm[l, i] = 1.0 / math.sqrt(0.5 * _sph_jn(l + 1, z[l, i]) ** 2)
    return z, norm


def _pref(l, m):
    return math.sqrt((2 * l + 1) / (4 * math.pi) * math.factorial(l - abs(m)) / math.factorial(l + abs(m)))


@lru_cache(maxsize=None)
def legendre_prefactors(num_spherical):
    """float32[num_spherical]: Y_l0 = pref[l] * P_l(cos theta)."""
    return np.array([_pref(l, 0) for l in range(num_spherical)], dtype=np.float32)


@lru_cache(maxsize=None)
def real_sph_harm_prefactors(num_spherical):
    """float32[num_spherical^2] in the reference's list order (l ascending; m = 0, +1..+l, -l..-1)."""
    out = []
    for l in range(num_spherical):
        pos = [_pref(l, 0)] + [math.sqrt(2.0) * (-1) ** m * _pref(l, m) for m in range(1, l + 1)]
        neg = [math.sqrt(2.0) * (-1) ** m * _pref(l, m) for m in range(l, 0, -1)]
        out.extend(pos + neg)
    return np.array(out, dtype=np.float32)


@lru_cache(maxsize=None)
def l_of_s(num_spherical, tensor_basis):
    """int32 map s -> l: identity for the 2-D basis, repeat_interleave(2l+1) for the 3-D one (basis_layers.py:251-253)."""
    if not tensor_basis:
        return np.arange(num_spherical, dtype=np.int32)
    return np.repeat(np.arange(num_spherical), 2 * np.arange(num_spherical) + 1).astype(np.int32)
