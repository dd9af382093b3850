#!/usr/bin/env python3
"""Differential fuzzing of the device post-processing entries (ss_post.cuh) on the CPU executor against oracle/postprocess.py:
random clouds, random reconstruction parameters, random combinations of the pipeline's post-processing switches.

    python tools/fuzz_postprocess.py --cases 100 --seed 0
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
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=50)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    import oracle
    import splashsurf_b200 as ss
    from fuzz_emulated import random_case
    from test_emulated_pipeline import build_emulated_library
    import test_zz_gpu_postprocess as T
    ss._LIB = ss._bind(C.CDLL(build_emulated_library()))
    bad = done = 0
    t0 = time.time()
    for i in range(a.cases):
        seed = a.seed * 1_000_003 + i
        rng = np.random.default_rng(seed)
        x, kw, _ = random_case(rng)
        kw.pop("rest_density", None)
        post = dict(mesh_smoothing_weights=bool(rng.integers(0, 2)), mesh_smoothing_weights_normalization=float(rng.choice([5.0, 13.0, 20.0])),
                    mesh_smoothing_iters=[None, 1, 2, 5][int(rng.integers(0, 4))], compute_normals=bool(rng.integers(0, 2)),
                    sph_normals=bool(rng.integers(0, 2)), normals_smoothing_iters=[None, 1, 3][int(rng.integers(0, 3))])
        attributes = {"a": rng.normal(size=len(x)).astype(np.float32), "v": rng.normal(size=x.shape).astype(np.float32)}
        o = oracle.reconstruct(x, **kw)
        if o["rc"] != 0 or len(o["vertices"]) == 0 or len(o["vertices"]) > 60000 or float(np.prod(o["grid"]["npoints"].astype(np.float64))) > 4e6:
            continue
        try:
            o, ref = T._oracle_pipeline(oracle, x, kw, post, attributes)
            m, rec = ss.reconstruction_pipeline(x, attributes_to_interpolate=attributes, **kw, **post, output_mesh_smoothing_weights=True,
                                                output_raw_normals=True, with_debug=True)
            got = dict(m.point_attributes); got["vertices"] = m.mesh.vertices
            want = {k: v for k, v in ref.items() if k in got}
            if post["compute_normals"] and not post["sph_normals"]:
                from oracle import postprocess as pp      # area normals: compare on the device's own vertices (conditioning)
                for key in ("normals", "raw_normals"):
                    want.pop(key, None)
                raw = pp.vertex_normals(m.mesh.vertices, m.mesh.triangles)
                mine = got.get("raw_normals", got["normals"])
                ok = np.isfinite(raw).all(axis=1) & np.isfinite(mine).all(axis=1)
                assert np.abs(mine[ok].astype(np.float64) - raw[ok]).max(initial=0.0) <= 1e-4, "area normals"
            # vertices without any particle in range give NaN in both (0 * inf), compare where the oracle is finite
            fin = {k: np.isfinite(np.asarray(v, dtype=np.float64)).reshape(len(v), -1).all(axis=1) for k, v in want.items()}
            oa, ob = T._key_order(rec.vertex_edge_keys), T._key_order(o["vertex_keys"])
            for k in want:
                aa = np.asarray(got[k], np.float64)[oa]; bb = np.asarray(want[k], np.float64)[ob]; f = fin[k][ob]
                assert np.isfinite(aa[f]).all(), k
                d = np.abs(aa[f] - bb[f]).reshape(int(f.sum()), -1).max(axis=1) if f.any() else np.zeros(1)
                scale = max(float(np.abs(bb[f]).max(initial=0.0)), 1.0)
                if "normals" in k:
                    # unit vectors g / |g|: where the SPH gradient nearly cancels, 1-ulp differences of the (smoothed) vertex
                    # positions are amplified by 1 / |g|; such vertices are rare, so the bulk must agree tightly and the
                    # worst vertex loosely
                    assert np.quantile(d, 0.99) <= 5e-4 and d.max() <= 5e-2, (k, float(np.quantile(d, 0.99)), float(d.max()))
                else:
                    assert d.max() / scale <= 5e-5, (k, float(d.max() / scale))
            done += 1
        except AssertionError as e:
            bad += 1
            print(f"[{seed}] MISMATCH {e} n={len(x)} {kw} {post}", flush=True)
        if i % 10 == 0:
            print(f"[{seed}] {done} compared, {bad} mismatches ({time.time() - t0:.0f}s)", flush=True)
    print(f"{a.cases} cases, {done} compared, {bad} mismatches, {time.time() - t0:.0f}s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
