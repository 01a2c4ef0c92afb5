"""TEST INFRASTRUCTURE — CPU restatement of the reference basis functions (rows H3, H4, H5 of SURVEY §8a).

Follows
  gemnet/model/layers/envelope.py:14-29           (polynomial envelope)
  gemnet/model/layers/basis_layers.py:45-49       (BesselBasisLayer.forward)
  gemnet/model/layers/basis_layers.py:119-162     (SphericalBasisLayer.forward)
  gemnet/model/layers/basis_layers.py:239-295     (TensorBasisLayer.forward)
  gemnet/model/layers/basis_utils.py:14-29        (Jn_zeros: float32-rounded zeros via brentq)
  gemnet/model/layers/basis_utils.py:47-80        (bessel_basis: normaliser 1/sqrt(0.5 j_{l+1}(z_ln)^2))
  gemnet/model/layers/basis_utils.py:83-104       (sph_harm_prefactor)
  gemnet/model/layers/basis_utils.py:107-171      (associated Legendre recurrences, incl. Condon-Shortley sign)
  gemnet/model/layers/basis_utils.py:174-253      (real_sph_harm; list order m = 0, +1..+l, -l..-1)

The reference generates closed-form sympy expressions and evaluates them with torch ops.  This
restatement evaluates the same mathematical functions with recurrences written directly in torch
(so it works in fp32 and fp64 and is differentiable).  In fp64 the two agree to ~1e-12; in fp32
both carry cancellation noise for small arguments / high l (SURVEY §7 hard part 2), so fp32
comparisons against the reference use tolerances scaled by that noise.
"""
import math
from functools import lru_cache

import numpy as np
import torch


# ----------------------------------------------------------------------------- constants
@lru_cache(maxsize=None)
def bessel_zeros(num_spherical: int, num_radial: int) -> np.ndarray:
    """First `num_radial` zeros of j_l, l < num_spherical, rounded to float32 (basis_utils.py:14-29)."""
    from scipy import special as sp
    from scipy.optimize import brentq

    n, k = num_spherical, num_radial
    zerosj = np.zeros((n, k), dtype="float32")
    zerosj[0] = np.arange(1, k + 1) * np.pi
    points = np.arange(1, k + n) * np.pi
    racines = np.zeros(k + n - 1, dtype="float32")
    for i in range(1, n):
        for j in range(k + n - 1 - i):
            racines[j] = brentq(lambda r: sp.spherical_jn(i, r), points[j], points[j + 1])
        points = racines
        zerosj[i][:k] = racines[:k]
    return zerosj


@lru_cache(maxsize=None)
def bessel_normalizers(num_spherical: int, num_radial: int) -> np.ndarray:
    """N_ln = (0.5 * j_{l+1}(z_ln)^2)^-1/2 in float64, z_ln the float32-rounded zero (basis_utils.py:57-66)."""
    from scipy import special as sp

    z = bessel_zeros(num_spherical, num_radial)
    out = np.zeros(z.shape, dtype=np.float64)
    for l in range(num_spherical):
        for i in range(num_radial):
            out[l, i] = 1.0 / math.sqrt(0.5 * float(sp.spherical_jn(l + 1, z[l, i])) ** 2)
    return out


def sph_harm_prefactor(l: int, m: int) -> float:
    """sqrt((2l+1)/(4 pi) (l-|m|)!/(l+|m|)!)  (basis_utils.py:83-104)."""
    return math.sqrt((2 * l + 1) / (4 * math.pi) * math.factorial(l - abs(m)) / math.factorial(l + abs(m)))


# ----------------------------------------------------------------------------- envelope / radial
def envelope(x: torch.Tensor, p: int) -> torch.Tensor:
    """envelope.py:14-29."""
    a = -(p + 1) * (p + 2) / 2
    b = p * (p + 2)
    c = -p * (p + 1) / 2
    val = 1 + a * x ** p + b * x ** (p + 1) + c * x ** (p + 2)
    return torch.where(x < 1, val, torch.zeros_like(x))


def bessel_rbf(d: torch.Tensor, frequencies: torch.Tensor, cutoff: float, p: int) -> torch.Tensor:
    """BesselBasisLayer.forward (basis_layers.py:45-49): (E,) -> (E, num_radial)."""
    inv_cutoff = 1 / cutoff
    norm_const = (2 * inv_cutoff) ** 0.5
    d = d[:, None]
    d_scaled = d * inv_cutoff
    return envelope(d_scaled, p) * norm_const * torch.sin(frequencies * d_scaled) / d


def spherical_jn_upward(lmax: int, x: torch.Tensor):
    """[j_0(x), ..., j_lmax(x)] by the closed forms the reference's sympy code generates
    (basis_utils.py:32-44): j_0 = sin/x, j_1 = sin/x^2 - cos/x, j_{l+1} = (2l+1)/x j_l - j_{l-1}."""
    s, c = torch.sin(x), torch.cos(x)
    js = [s / x]
    if lmax >= 1:
        js.append(s / (x * x) - c / x)
    for l in range(1, lmax):
        js.append((2 * l + 1) / x * js[l] - js[l - 1])
    return js


def radial_sph_bessel(d: torch.Tensor, num_spherical: int, num_radial: int, cutoff: float, p: int) -> torch.Tensor:
    """rbf_env of SphericalBasisLayer/TensorBasisLayer (basis_layers.py:121-128 / 241-250):
    (E,) -> (E, num_spherical, num_radial), value env(x) * cutoff^-1.5 * N_ln * j_l(z_ln x), x = d/cutoff."""
    inv_cutoff = 1 / cutoff
    x = d * inv_cutoff
    z = torch.as_tensor(bessel_zeros(num_spherical, num_radial).astype(np.float64), dtype=d.dtype)
    nrm = torch.as_tensor(bessel_normalizers(num_spherical, num_radial), dtype=d.dtype)
    arg = z[None, :, :] * x[:, None, None]  # (E, S, R)
    cols = []
    for l in range(num_spherical):
        cols.append(spherical_jn_upward(l, arg[:, l, :])[l])
    j = torch.stack(cols, dim=1)  # (E, S, R)
    out = j * nrm[None] * (inv_cutoff ** 1.5)
    return envelope(x, p)[:, None, None] * out


# ----------------------------------------------------------------------------- angular
def legendre_y_l0(angle: torch.Tensor, num_spherical: int) -> torch.Tensor:
    """Y_l0(angle), l < num_spherical: (T,) -> (T, num_spherical)  (basis_layers.py:130-131, basis_utils.py:131-138,214-215)."""
    z = torch.cos(angle)
    P = [torch.ones_like(z)]
    if num_spherical > 1:
        P.append(z)
    for l in range(2, num_spherical):
        P.append(((2 * l - 1) * z * P[l - 1] - (l - 1) * P[l - 2]) / l)
    return torch.stack([sph_harm_prefactor(l, 0) * P[l] for l in range(num_spherical)], dim=1)


def real_sph_harm_all(theta: torch.Tensor, phi: torch.Tensor, num_spherical: int) -> torch.Tensor:
    """Real spherical harmonics Y_lm(theta, phi), l < num_spherical, in the reference's list order
    (l ascending; within l: m = 0, +1..+l, -l..-1): (Q,), (Q,) -> (Q, num_spherical^2).
    (basis_layers.py:216-229,269-270; basis_utils.py:139-171,217-243)."""
    L = num_spherical
    z = torch.cos(theta)
    # (1 - z^2)^0.5 with z = cos(theta) is |sin(theta)|  (basis_utils.py:143)
    st = torch.sqrt(torch.clamp(1 - z * z, min=0))
    P = [[None] * (l + 1) for l in range(L)]
    P[0][0] = torch.ones_like(z)
    for l in range(1, L):
        P[l][l] = (1 - 2 * l) * st * P[l - 1][l - 1]
    for m in range(0, L - 1):
        P[m + 1][m] = (2 * m + 1) * z * P[m][m]
    for l in range(2, L):
        for m in range(l - 1):
            P[l][m] = ((2 * l - 1) * z * P[l - 1][m] - (l + m - 1) * P[l - 2][m]) / (l - m)
    out = []
    for l in range(L):
        pos = [sph_harm_prefactor(l, 0) * P[l][0]]
        neg = []
        for m in range(1, l + 1):
            pref = (2 ** 0.5) * ((-1) ** m) * sph_harm_prefactor(l, m)
            pos.append(pref * P[l][m] * torch.cos(m * phi))
            neg.append(pref * P[l][m] * torch.sin(m * phi))
        # python list index -m  <=>  position 2l+1-m: after the +m entries come m = -l, ..., -1
        out.extend(pos + neg[::-1])
    return torch.stack(out, dim=1)
