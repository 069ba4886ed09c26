#!/usr/bin/env python
"""Reproduces the coefficients and the error bound of gelu_erf_fast (csrc/common.cuh): erfc(z) = 2^(-Q(z)),
Q a degree-6 polynomial through the origin, erfc-weighted least-squares fit of -log2(erfc(z)) on [0, 4.2],
expressed in |x| = sqrt(2) z; evaluated in float32 like the kernel."""
import numpy as np
from scipy.special import erf, erfc

ZMAX, N = 4.2, 6
z = np.cos(np.pi * (np.arange(4000) + 0.5) / 4000) * ZMAX / 2 + ZMAX / 2
q, w = -np.log2(erfc(z)), erfc(z)
a = np.vander(z, N + 1, increasing=True)[:, 1:]
co = np.linalg.lstsq(a * w[:, None], q * w, rcond=None)[0]
cx = np.array([co[k - 1] / 2 ** (k / 2) for k in range(1, N + 1)])
print("|x| clamp:", ZMAX * np.sqrt(2))
print("coefficients (x^1 .. x^6 of Q):", [f"{c:.9e}" for c in cx])
x = np.linspace(-9, 9, 1800001).astype(np.float32)
ax = np.minimum(np.abs(x), np.float32(ZMAX * np.sqrt(2)))
c32 = cx.astype(np.float32)
acc = np.zeros_like(ax) + c32[5]
for k in range(4, -1, -1):
    acc = (acc * ax + c32[k]).astype(np.float32)
e = np.exp2(-(acc * ax).astype(np.float32)).astype(np.float32)
hx = np.float32(0.5) * x
g = (hx + np.abs(hx) * (np.float32(1) - e)).astype(np.float32)
ref = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
print("GELU max abs error:", np.abs(g - ref).max())
