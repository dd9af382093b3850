#!/usr/bin/env python3
"""A seconds-long GPU session for the entries added after the last full GPU run of round 2: runs the GPU-marked checks of
tests/test_zzzz_reference_datasets.py / tests/test_zzzzz_pysplashsurf_tests.py directly (no pytest start-up, no rebuild), cheapest first, and
prints one line per check as it finishes -- so that a call that is cut short still reports what it reached.

    python tools/gpu_quick_new_entries.py > gpurun_out/new_entries_gpu.log 2>&1"""
import os
import sys
import tempfile
import time
import pathlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
T0 = time.time()


def say(*a):
    print(f"[{time.time() - T0:6.1f} s]", *a, flush=True)


import splashsurf_b200 as ss  # noqa: E402
ss.load_library()
import oracle  # noqa: E402
say("library + oracle loaded; reference wheel available:", oracle.reference_available())
import test_zzzz_reference_datasets as D  # noqa: E402
import test_zzzzz_pysplashsurf_tests as P  # noqa: E402

tmp = pathlib.Path(tempfile.mkdtemp())
(tmp / "cli").mkdir()
(tmp / "py").mkdir()
CHECKS = [
    ("test_simple.rs (both strategies)", lambda: D.check_test_simple(ss.reconstruct_surface, ss)),
    ("neighbourhood hand cases through reconstruct_surface", lambda: D._check_ns(lambda p, **kw: ss.reconstruct_surface(p, global_neighborhood_list=True, **kw),
                                                                                 lambda g, n: [g.particle_neighbors[i] for i in range(n)])),
    ("stand-alone SphInterpolator", lambda: D.check_sph_interpolator(ss, oracle)),
    ("stand-alone neighbourhood search", lambda: D.check_neighborhood_search(ss, oracle)),
    ("command-line frame sequence", lambda: D.check_cli_sequence(ss, tmp / "cli")),
    ("the reference's Python tests", lambda: P.run_all(ss, tmp / "py", oracle)),
    ("marching_cubes on dense arrays", lambda: D.check_marching_cubes(ss, oracle)),
]
failed = 0
for name, fn in CHECKS:
    t = time.time()
    try:
        fn()
        say(f"PASS  {name}  ({time.time() - t:.1f} s)")
    except BaseException as e:                       # noqa: BLE001 - report and go on
        failed += 1
        say(f"FAIL  {name}: {type(e).__name__}: {str(e)[:300]}")
say("done,", failed, "failed")
sys.exit(1 if failed else 0)
