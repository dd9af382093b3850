#!/usr/bin/env python3
"""Work statistics of the certification sweep (ss_certify_box) on a dense dam-break sample, taken on the CPU executor built
with -DSS_EMUL_STATS: words of 32 candidates visited and candidates evaluated per warp box and ring, share of the boxes
certified in ring 0 / ring 1 / not at all.  Guides the kernel work of round 2."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import splashsurf_b200 as ss
    from splashsurf_b200 import synthetic as syn
    emul = os.path.join(ROOT, "tests", "emul")
    so = os.path.join(emul, "libsplashsurf_emul_stats.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-w", "-x", "c++", "-DSS_HOST_EMUL", "-DSS_EMUL_STATS",
                           "-I/usr/local/cuda/include", "-include", os.path.join(emul, "cuda_emul.h"), "-shared", "-fPIC", "-pthread", "-o", so,
                           os.path.join(ROOT, "splashsurf_b200", "csrc", "ss_pipeline.cu")])
    L = C.CDLL(so)
    ss._LIB = ss._bind(L)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000
    x = syn.jittered_cube(int(round(n ** (1.0 / 3.0))), 0.01, 3) if os.environ.get("SS_STATS_CUBE") else syn.dam_break_scaled(n, 0.01, 3)
    g = ss.reconstruct_surface(x, particle_radius=0.01, smoothing_length=2.0, cube_size=0.5)
    st = (C.c_ulonglong * 15)()
    L.ss_emul_stats_read(st, 1)
    v = np.array([int(q) for q in st], dtype=np.float64).reshape(3, 5)
    boxes = v[:, 0].sum()
    print(f"{len(x)} particles, {g.timings['bricks_levelset']} bricks, {int(boxes)} warp boxes, avg candidates per brick C: see bricks")
    for o, name in enumerate(("certified in ring 0", "certified in ring 1", "not certified")):
        b = max(v[o, 0], 1.0)
        print(f"{name:20s} {v[o, 0] / boxes * 100:5.1f} % of the boxes | ring 0: {v[o, 1] / b:5.2f} words, {v[o, 2] / b:5.2f} candidates | "
              f"ring 1: {v[o, 3] / b:5.2f} words, {v[o, 4] / b:5.2f} candidates  (per box)")


if __name__ == "__main__":
    main()
