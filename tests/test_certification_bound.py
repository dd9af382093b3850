"""CPU tests of the invariant behind the level-set certification pass: the cubic g(s) used by
`ss_certify_box` never exceeds the kernel shape, and partial sums built from it never exceed the reference's exact value."""
import os
import re

import numpy as np

from conftest import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _coeffs():
    src = open(os.path.join(ROOT, "splashsurf_b200", "csrc", "ss_kernels.cuh")).read()
    return [float(re.search(rf"#define SS_G{k} (-?[0-9.]+)f", src).group(1)) for k in range(4)]


def _f(q):
    q = np.asarray(q, dtype=np.float64)
    return np.where(q <= 0.5, 1 - 6 * q ** 2 + 6 * q ** 3, np.where(q < 1, 2 * (1 - q) ** 3, 0.0))


def test_cubic_is_a_lower_bound_of_the_kernel_shape():
    g0, g1, g2, g3 = _coeffs()
    s = np.concatenate([np.linspace(0, 1.0, 2_000_001), np.linspace(1.0, 9.0, 200_001)])
    g = np.maximum(0.0, g0 + s * (g1 + s * (g2 + s * g3)))
    assert (g <= _f(np.sqrt(s)) + 1e-12).all()
    # float32 evaluation (as on the device, Horner with FMAs) stays below as well: the -1e-5 offset in G0 covers rounding
    s32 = s.astype(np.float32)
    g32 = np.maximum(np.float32(0), np.float32(g0) + s32 * (np.float32(g1) + s32 * (np.float32(g2) + s32 * np.float32(g3))))
    assert (g32.astype(np.float64) <= _f(np.sqrt(s32.astype(np.float64))) + 2e-6).all()
    q = np.linspace(0, 1.2, 240001)
    cap = 32 * np.sum(np.maximum(0.0, g0 + q * q * (g1 + q * q * (g2 + q * q * g3))) * q * q) * (q[1] - q[0])
    assert 0.93 < cap < 1.0          # captures ~94 % of the kernel's volume integral


def test_partial_sums_never_exceed_the_exact_level_set(oracle_mod):
    """On the reference's grid-loop fixture: for random grid points, the bound summed over ALL particles within 0.8 h (the
    largest set a warp can use) is <= the exact tile value (AVX and scalar arithmetic) times (1 + 1e-4)."""
    g = load_golden("grid_loop_subdomain_33")
    g0, g1, g2, g3 = _coeffs()
    h, c, m = float(g["h"]), float(g["cell_size"]), float(g["rest_mass"])
    common = dict(global_min=g["global_min"], cube_size=g["cell_size"], subdomain_ijk=g["subdomain_ijk"], subdomain_cubes=64,
                  subdomain_min=g["subdomain_min"], h=g["h"], rest_mass=g["rest_mass"])
    p, rho = g["particles"].astype(np.float64), g["densities"].astype(np.float64)
    vol = m / rho
    sigma = 8.0 / (np.pi * h ** 3)
    rng = np.random.default_rng(3)
    pts = rng.integers(0, 65, size=(400, 3))
    for mode in (0, 1):
        tile = oracle_mod.levelset_tile(g["particles"], g["densities"], mode=mode, **common).astype(np.float64)
        for i, j, k in pts:
            x = g["subdomain_min"].astype(np.float64) + np.array([i, j, k]) * c
            s = ((p - x[None]) ** 2).sum(1) / h ** 2
            sel = s < 0.64
            lb = sigma * np.sum(vol[sel] * np.maximum(0.0, g0 + s[sel] * (g1 + s[sel] * (g2 + s[sel] * g3))))
            assert lb <= tile[i, j, k] * (1 + 1e-4) + 1e-12, (mode, i, j, k, lb, tile[i, j, k])
