#!/usr/bin/env python3
"""Differential fuzzing of the device path against the pinned oracle WITHOUT a GPU: random small particle clouds and random
parameters go through the CPU execution of the CUDA sources (tests/emul/cuda_emul.h) and through oracle.reconstruct; any
difference (densities, subdomain lists, connectivity, a single vertex bit) is reported with the seed that reproduces it.

    python tools/fuzz_emulated.py --cases 200 --seed 0
"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def random_case(rng):
    r = float(rng.choice([0.01, 0.025, 0.05]))
    d = 2 * r
    kind = rng.choice(["lattice", "blobs", "sheet", "uniform"])
    if kind == "lattice":
        n = rng.integers(3, 11, size=3)
        g = np.stack(np.meshgrid(*[np.arange(k) for k in n], indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * d
        x = g + rng.uniform(-0.3 * d, 0.3 * d, g.shape)
    elif kind == "blobs":
        parts = []
        for _ in range(int(rng.integers(1, 5))):
            c = rng.uniform(-10 * d, 10 * d, 3)
            parts.append(c + rng.normal(0, rng.uniform(0.5, 3.0) * d, (int(rng.integers(5, 400)), 3)))
        x = np.concatenate(parts)
    elif kind == "sheet":
        n = rng.integers(4, 16, size=2)
        g = np.stack(np.meshgrid(np.arange(n[0]), np.arange(int(rng.integers(1, 3))), np.arange(n[1]), indexing="ij"), -1).reshape(-1, 3) * d
        x = g + rng.uniform(-0.2 * d, 0.2 * d, g.shape)
    else:
        x = rng.uniform(0, rng.uniform(4, 14) * d, (int(rng.integers(1, 900)), 3))
    x = (x + rng.uniform(-50 * d, 50 * d, 3)).astype(np.float32)
    kw = dict(particle_radius=r, smoothing_length=float(rng.choice([1.2, 1.5, 2.0, 2.2, 2.5])), cube_size=float(rng.choice([0.3, 0.5, 0.75, 1.0, 1.1, 1.5])),
              iso_surface_threshold=float(rng.choice([0.3, 0.6, 0.6, 0.8])), rest_density=float(rng.choice([1000.0, 850.0])),
              simd=bool(rng.integers(0, 2)))
    mode = rng.choice(["subdomain", "subdomain", "global", "auto"])
    if mode == "subdomain":
        kw.update(subdomain_num_cubes_per_dim=int(rng.choice([8, 12, 16, 20, 24, 32, 40, 64])), subdomain_grid_auto_disable=False)
    elif mode == "global":
        kw.update(subdomain_grid=False)
    if rng.integers(0, 5) == 0:
        lo, hi = x.min(0), x.max(0)
        kw.update(aabb_min=[float(v) for v in lo + rng.uniform(0, 0.3, 3) * (hi - lo)], aabb_max=[float(v) for v in hi - rng.uniform(0, 0.3, 3) * (hi - lo)])
    opts = dict(exact=bool(rng.integers(0, 4) == 0), batch=int(rng.choice([0, 0, 1, 3])))
    return x, kw, opts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--variant", type=int, default=0, help="level-set launch structure (ss_context_set_levelset_variant)")
    ap.add_argument("--max-points", type=float, default=6e6, help="skip cases whose grid has more points than this (emulation is slow)")
    a = ap.parse_args()
    import oracle
    import splashsurf_b200 as ss
    from test_emulated_pipeline import build_emulated_library
    ss._LIB = ss._bind(C.CDLL(build_emulated_library()))
    bad = 0
    t0 = time.time()
    for i in range(a.cases):
        seed = a.seed * 1_000_003 + i
        x, kw, opts = random_case(np.random.default_rng(seed))
        o = oracle.reconstruct(x, **kw)
        if o["rc"] != 0:
            try:
                ss.reconstruct_surface(x, **kw)
                print(f"[{seed}] oracle failed with rc={o['rc']} but the device path succeeded: {kw}"); bad += 1
            except ss.SplashsurfError as e:
                if e.code != o["rc"]:
                    print(f"[{seed}] error codes differ: oracle {o['rc']} device {e.code}: {kw}"); bad += 1
            continue
        if float(np.prod(o["grid"]["npoints"].astype(np.float64))) > a.max_points:
            continue
        ctx = ss.Context()
        ctx.set_levelset_exact_everywhere(opts["exact"])
        ctx.set_levelset_variant(a.variant)
        if opts["batch"]:
            ctx.set_tile_batch(opts["batch"])
        try:
            g = ss.reconstruct_surface(x, with_debug=True, context=ctx, **kw)
        except ss.SplashsurfError as e:
            print(f"[{seed}] device path failed ({e.code}: {e}) where the oracle succeeded: n={len(x)} {kw} {opts}"); bad += 1
            continue
        finally:
            ctx.close()
        ok = np.array_equal(g.particle_densities, o["particle_densities"])
        if o["used_decomposition"]:
            ok = ok and np.array_equal(g.subdomains["flat"], o["subdomain_flat"]) and np.array_equal(g.subdomains["sparse"], o["subdomain_sparse"])
        m = oracle.mesh_parity(g.mesh.vertices, g.mesh.triangles, g.vertex_edge_keys, o["vertices"], o["triangles"], o["vertex_keys"],
                               kw.get("subdomain_num_cubes_per_dim", 64))
        ok = ok and m["keys_equal"] and m["triangles_equal"] and m["n_not_bitexact"] == 0
        if not ok:
            bad += 1
            print(f"[{seed}] MISMATCH n={len(x)} {kw} {opts} {m}")
        elif i % 10 == 0:
            print(f"[{seed}] ok n={len(x)} nv={g.mesh.nvertices} decomposition={o['used_decomposition']} ({time.time() - t0:.0f}s)", flush=True)
    print(f"{a.cases} cases, {bad} mismatches, {time.time() - t0:.0f}s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
