"""Real spherical-harmonic transform tables (fp64 on the host), torch-harmonics conventions
(the reference reaches them through earth2mip's fcnv2_sm,
/root/reference/skyrim/core/models/fourcastnet_v2.py:36-37):

    forward:  F[k, m] = (2 pi / nlon) * sum_j x[k, j] exp(-i m phi_j)          (rfft, norm="forward", * 2 pi)
              X[l, m] = sum_k  w_k  Pbar_l^m(cos theta_k)  F[k, m]
    inverse:  F[k, m] = sum_l  Pbar_l^m(cos theta_k) X[l, m]
              x[k, j] = irfft(F[k, :], n=nlon, norm="forward")

Pbar: orthonormal associated Legendre functions with Condon-Shortley phase.  Grids:
"equiangular" (nlat points pole to pole, Clenshaw-Curtis weights) and "legendre-gauss".
Both the CPU oracle and the CUDA engine consume these tables (the engine receives them as
extra entries of the weight arena), so the C++ side contains no Legendre code.
"""
from __future__ import annotations

import numpy as np


def clenshaw_curtis(n: int):
    """nodes cos(theta) (theta = pi k/(n-1), north to south) and weights on [-1, 1]."""
    N = n - 1
    k = np.arange(n)
    theta = np.pi * k / N
    w = np.zeros(n)
    j = np.arange(1, N // 2 + 1)
    b = np.where(2 * j == N, 1.0, 2.0)
    for kk in range(n):
        s = 1.0 - np.sum(b / (4.0 * j * j - 1.0) * np.cos(2.0 * j * kk * np.pi / N))
        w[kk] = (1.0 if kk in (0, N) else 2.0) / N * s
    return np.cos(theta), w


def legendre_gauss(n: int):
    x, w = np.polynomial.legendre.leggauss(n)
    return x[::-1].copy(), w[::-1].copy()  # north (cos = +1) first


def legpoly(mmax: int, lmax: int, cost: np.ndarray) -> np.ndarray:
    """Pbar[m, l, k], orthonormal (int over the sphere of |Y_lm|^2 = 1), Condon-Shortley phase."""
    nmax = max(mmax, lmax)
    t = np.asarray(cost, dtype=np.float64)
    v = np.zeros((nmax, nmax, t.size))
    v[0, 0] = 1.0 / np.sqrt(4.0 * np.pi)
    for l in range(1, nmax):
        v[l - 1, l] = np.sqrt(2 * l + 1) * t * v[l - 1, l - 1]
        v[l, l] = np.sqrt((2 * l + 1) * (1 + t) * (1 - t) / (2 * l)) * v[l - 1, l - 1]
    for l in range(2, nmax):
        for m in range(0, l - 1):
            v[m, l] = (t * np.sqrt((2 * l - 1) / (l - m) * (2 * l + 1) / (l + m)) * v[m, l - 1]
                       - np.sqrt((l + m - 1) / (l - m) * (2 * l + 1) / (2 * l - 3) * (l - m - 1) / (l + m)) * v[m, l - 2])
    v = v[:mmax, :lmax]
    v[1::2] *= -1.0
    return v


def grid_nodes(nlat: int, grid: str):
    if grid == "equiangular":
        return clenshaw_curtis(nlat)
    if grid == "legendre-gauss":
        return legendre_gauss(nlat)
    raise ValueError(grid)


def sht_tables(nlat: int, lmax: int, mmax: int, grid: str):
    """(fwd[m, l, k] = w_k Pbar, inv[m, k, l] = Pbar)"""
    cost, w = grid_nodes(nlat, grid)
    p = legpoly(mmax, lmax, cost)
    fwd = p * w[None, None, :]
    inv = np.ascontiguousarray(p.transpose(0, 2, 1))
    return fwd, inv


def dft_matrices(nlon: int, mmax: int):
    """fwd[(m, re/im), j] and inv[j, (m, re/im)] of the truncated real DFT pair above."""
    j = np.arange(nlon)
    m = np.arange(mmax)
    ang = 2.0 * np.pi * np.outer(m, j) / nlon
    fwd = np.empty((mmax, 2, nlon))
    fwd[:, 0] = (2.0 * np.pi / nlon) * np.cos(ang)
    fwd[:, 1] = -(2.0 * np.pi / nlon) * np.sin(ang)
    c = np.full(mmax, 2.0)
    c[0] = 1.0
    if nlon % 2 == 0 and mmax - 1 >= nlon // 2:
        c[nlon // 2] = 1.0
    inv = np.empty((nlon, mmax, 2))
    inv[:, :, 0] = (c[:, None] * np.cos(ang)).T
    inv[:, :, 1] = (-c[:, None] * np.sin(ang)).T
    inv[:, 0, 1] = 0.0                      # irfft ignores Im of the DC (and Nyquist) mode
    if nlon % 2 == 0 and mmax - 1 >= nlon // 2:
        inv[:, nlon // 2, 1] = 0.0
    return fwd.reshape(2 * mmax, nlon), inv.reshape(nlon, 2 * mmax)


class RealSHT:
    """numpy reference transform pair (used by the oracle)."""

    def __init__(self, nlat, nlon, lmax, mmax, grid):
        self.nlat, self.nlon, self.lmax, self.mmax = nlat, nlon, lmax, mmax
        self.fwd, self.inv = sht_tables(nlat, lmax, mmax, grid)

    def forward(self, x):  # (..., nlat, nlon) -> complex (..., lmax, mmax)
        f = 2.0 * np.pi * np.fft.rfft(x, axis=-1, norm="forward")[..., : self.mmax]
        return np.einsum("mlk,...km->...lm", self.fwd, f)

    def inverse(self, X):  # complex (..., lmax, mmax) -> (..., nlat, nlon)
        f = np.einsum("mkl,...lm->...km", self.inv, X)
        return np.fft.irfft(f, n=self.nlon, axis=-1, norm="forward")
