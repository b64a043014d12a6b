#!/usr/bin/env python
"""Coefficients of the two polynomial approximations used by the kernels, with their error.

  * q(x) = -log2(erfc(x / sqrt 2)) on [0, 4.95], degree 6  ->  exact-erf GELU as x*Phi(x), Phi(-|x|) = 0.5 * 2^-q(|x|)
    (csrc/gemm.cuh gelu_erf, csrc/mlp.cuh gelu_erf_x2)
  * 2^f on [-0.5, 0.5], degree 3 (relative-error weighted)  ->  exp2 on the FMA pipes (csrc/ptx.cuh exp2_poly_x2)
"""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P
from scipy.special import erf, erfc


def fit_gelu(zmax=3.5, deg=6):
    xmax = zmax * np.sqrt(2)
    u = np.cos(np.pi * (np.arange(8000) + 0.5) / 8000)
    x = (u + 1) / 2 * xmax
    q = -np.log2(erfc(x / np.sqrt(2)))
    px = P.Polynomial(C.cheb2poly(C.chebfit(u, q, deg)))(P.Polynomial([-1, 2 / xmax]))
    coef = px.coef
    xx = np.linspace(-8, 8, 400001).astype(np.float32)
    ax = np.minimum(np.abs(xx), np.float32(xmax))
    acc = np.full_like(ax, np.float32(coef[-1]))
    for k in range(deg - 1, -1, -1):
        acc = (acc * ax + np.float32(coef[k])).astype(np.float32)
    w = np.float32(0.5) * np.exp2(-acc.astype(np.float64)).astype(np.float32)
    g = np.where(xx >= 0, xx - xx * w, xx * w).astype(np.float32)
    ref = 0.5 * xx.astype(np.float64) * (1 + erf(xx.astype(np.float64) / np.sqrt(2)))
    print("gelu: xmax", xmax, "coefficients x^0..x^6:", [float(np.float32(v)) for v in coef])
    print("      max |gelu - exact| in fp32:", float(np.abs(g - ref).max()))


def fit_exp2(deg=3):
    u = np.cos(np.pi * (np.arange(4000) + 0.5) / 4000)
    y = np.exp2(u * 0.5)
    px = P.Polynomial(C.cheb2poly(C.chebfit(u, y, deg, w=1 / y)))(P.Polynomial([0, 2.0]))
    coef = px.coef
    f = np.linspace(-0.5, 0.5, 200001).astype(np.float32)
    acc = np.full_like(f, np.float32(coef[-1]))
    for k in range(deg - 1, -1, -1):
        acc = (acc * f + np.float32(coef[k])).astype(np.float32)
    print("exp2: coefficients f^0..f^3:", [float(np.float32(v)) for v in coef])
    print("      max relative error in fp32:", float(np.abs(acc / np.exp2(f.astype(np.float64)) - 1).max()))


if __name__ == "__main__":
    fit_gelu()
    fit_exp2()
