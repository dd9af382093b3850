#!/usr/bin/env python3
"""Verbose GPU-vs-oracle parity run (debugging aid; the pytest -m gpu suite is the real gate)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
import splashsurf_b200 as ss
from splashsurf_b200 import synthetic as syn


def check(name, p, **kw):
    t = time.time(); o = oracle.reconstruct(p, **kw); to = time.time() - t
    t = time.time(); g = ss.reconstruct_surface(p, with_debug=True, **kw); tg = time.time() - t
    dens_ok = np.array_equal(g.particle_densities, o["particle_densities"])
    dd = np.abs(g.particle_densities - o["particle_densities"])
    sub_ok = (np.array_equal(g.subdomains["flat"], o["subdomain_flat"]) and np.array_equal(g.subdomains["count"], o["subdomain_count"])
              and np.array_equal(g.subdomains["sparse"], o["subdomain_sparse"]))
    m = oracle.mesh_parity(g.mesh.vertices, g.mesh.triangles, g.vertex_edge_keys, o["vertices"], o["triangles"], o["vertex_keys"],
                           kw.get("subdomain_num_cubes_per_dim", 64))
    print(f"[{name}] n={len(p)} oracle {to:.2f}s gpu {tg:.2f}s | dens exact={dens_ok} maxdiff={dd.max() if len(dd) else 0:.3g} "
          f"| subdomains ok={sub_ok} ({len(o['subdomain_flat'])}, sparse {int(o['subdomain_sparse'].sum())}) | {m}")
    print("   timings", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in g.timings.items()})
    return dens_ok and sub_ok and m["keys_equal"] and m["triangles_equal"] and m.get("n_not_bitexact", 1) == 0


def main():
    ok = True
    anchor = os.path.join(ROOT, "tests", "golden", "cfg1_particles.npy")
    if os.path.exists(anchor):
        ok &= check("cfg1", np.load(anchor), particle_radius=0.025, smoothing_length=2.2, cube_size=1.1)
    ok &= check("cube20", syn.jittered_cube(20, 0.025, 11), particle_radius=0.025, smoothing_length=2.0, cube_size=0.5)
    ok &= check("cube40", syn.jittered_cube(40, 0.025, 1234), particle_radius=0.025, smoothing_length=2.0, cube_size=0.5)
    ok &= check("cube40-scalar", syn.jittered_cube(40, 0.025, 5), particle_radius=0.025, smoothing_length=2.0, cube_size=0.5, simd=False)
    ok &= check("cube30-S32-c075", syn.jittered_cube(30, 0.025, 6), particle_radius=0.025, smoothing_length=2.0, cube_size=0.75, subdomain_num_cubes_per_dim=32)
    ok &= check("splash", syn.splash((30, 32, 30), 8, 0.025, 4), particle_radius=0.025, smoothing_length=2.0, cube_size=0.5)
    ok &= check("splash-c045", syn.splash((24, 24, 24), 6, 0.025, 9), particle_radius=0.025, smoothing_length=2.0, cube_size=0.45)
    ok &= check("splash-aabb", syn.splash((30, 32, 30), 8, 0.025, 4), particle_radius=0.025, smoothing_length=2.0, cube_size=0.5,
                aabb_min=[-0.1, -0.1, -0.1], aabb_max=[1.0, 2.5, 1.0])
    if "--big" in sys.argv:
        ok &= check("cube100", syn.jittered_cube(100, 0.025, 1234), particle_radius=0.025, smoothing_length=2.0, cube_size=0.5)
    print("ALL OK" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
