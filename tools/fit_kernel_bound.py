#!/usr/bin/env python3
"""Fits the cubic lower bound g(s) <= f(sqrt(s)) of the cubic-spline kernel shape used by the certification pass of
k_levelset (csrc/ss_kernels.cuh, SS_G0..SS_G3): a linear program over a fine grid in s = (r/h)^2 that maximises the captured
weight (3-D shell measure) subject to g <= f everywhere."""
import numpy as np
from scipy.optimize import linprog


def f(q):
    q = np.asarray(q)
    return np.where(q <= 0.5, 1 - 6 * q ** 2 + 6 * q ** 3, np.where(q < 1, 2 * (1 - q) ** 3, 0.0))


deg = 3
s = np.concatenate([np.linspace(0, 1.0, 4001), np.linspace(1.0, 4.0, 1200)])
A = np.stack([s ** k for k in range(deg + 1)], 1)
b = f(np.sqrt(s))
so = np.linspace(0, 0.36, 1500)
c = -(np.stack([so ** k for k in range(deg + 1)], 1) * np.sqrt(so)[:, None]).sum(0)
res = linprog(c, A_ub=A, b_ub=b - 1e-7, bounds=[(None, None)] * (deg + 1))
a = res.x
q = np.linspace(0, 1.2, 240001)
g = np.maximum(0, sum(a[k] * (q * q) ** k for k in range(deg + 1)))
assert (g <= f(q) + 1e-6).all()
print("coefficients (G0..G3):", a, " captured weight:", 32 * np.sum(g * q * q) * (q[1] - q[0]))
