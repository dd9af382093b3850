#!/usr/bin/env python3
"""Differential fuzzing of splashsurf_b200.marching_cubes (ss_marching_cubes_tiles_f32 + the front end's padding / filtering) against the
reference wheel's pysplashsurf.marching_cubes on the CPU executor of the CUDA sources: random shapes around the tile size (2 .. 129 points per
axis), smooth, noisy, nearly constant and integer-valued fields (values exactly on the threshold), random cube sizes and translations.
A case counts as a mismatch when both sides return a mesh and the canonically ordered vertices (bits) or triangles differ; refusals of
either side are tallied (this front end never returns a mesh where the reference reports an error).

    python tools/fuzz_marching_cubes.py [seed] [seconds]"""
import sys, ctypes as C, numpy as np, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import splashsurf_b200 as ss, oracle
from test_emulated_pipeline import build_emulated_library
from test_zzzz_reference_datasets import _canonical_mesh
ss._LIB = ss._bind(C.CDLL(build_emulated_library()))
ps = oracle.reference()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
t0 = time.time(); n = 0; bad = 0
while time.time() - t0 < float(sys.argv[2]) if len(sys.argv) > 2 else 240:
    kind = rng.integers(0, 4)
    shape = tuple(int(x) for x in rng.choice([2, 3, 5, 17, 64, 65, 66, 70, 129], size=3)) if kind != 3 else tuple(int(x) for x in rng.integers(2, 40, size=3))
    if np.prod(shape) > 600000: shape = (shape[0], min(shape[1], 40), shape[2])
    if kind == 0:   # integer-valued field: many values exactly on the threshold
        f = rng.integers(-2, 3, size=shape).astype(np.float32); thr = float(rng.integers(-1, 2))
    elif kind == 1:
        from scipy.ndimage import gaussian_filter
        f = gaussian_filter(rng.normal(size=shape), 1.2).astype(np.float32); thr = float(rng.normal(scale=0.05))
    elif kind == 2:  # mostly inside / mostly outside with noise
        f = (rng.random(size=shape) < 0.05).astype(np.float32) * rng.random(size=shape).astype(np.float32) + (0.5 if rng.random() < 0.5 else 0.0); thr = 0.5
        f = f.astype(np.float32)
    else:
        f = rng.normal(size=shape).astype(np.float32); thr = 0.0
    cs = float(np.float32(rng.choice([1.0, 0.3, 0.0125]))); tr = [float(v) for v in rng.normal(size=3).astype(np.float32)]
    try:
        m = ss.marching_cubes(f, iso_surface_threshold=thr, cube_size=cs, translation=tr); merr = None
    except ss.SplashsurfError as e:
        merr = e
    try:
        r = ps.marching_cubes(f, iso_surface_threshold=thr, cube_size=cs, translation=tr); rerr = None
    except BaseException as e:
        rerr = e
    n += 1
    if merr is not None or rerr is not None:
        kinds = globals().setdefault("errs", {"both": 0, "mine_only": 0, "ref_only": 0})
        kinds["both" if (merr is not None and rerr is not None) else ("mine_only" if merr is not None else "ref_only")] += 1
        continue
    rv, rt = np.asarray(r.vertices), np.asarray(r.triangles)
    ok = len(rv) == len(m.vertices) and len(rt) == len(m.triangles)
    if ok:
        a, b = _canonical_mesh(m.vertices, m.triangles), _canonical_mesh(rv, rt)
        ok = np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[1], b[1])
    if not ok:
        bad += 1; print("MISMATCH", kind, shape, thr, cs, tr, m.vertices.shape, rv.shape, m.triangles.shape, rt.shape); 
        if bad > 3: break
print("cases", n, "mismatches", bad, globals().get("errs"))
