#!/usr/bin/env python3
"""tests/golden/pysplashsurf_tests.npz: the particle files of the reference's own Python tests (pysplashsurf/tests/ParticleData_Random_1000.vtk,
ParticleData_Fluid_5.vtk with its `velocity` / `id` point data, ParticleData_Fluid_50.bgeo) read with this package's readers, so that
tests/test_zzzzz_pysplashsurf_tests.py -- those tests, re-run against `import splashsurf_b200 as pysplashsurf` -- also runs on the GPU box."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from splashsurf_b200 import particle_formats as pf  # noqa: E402

D = "/root/reference/pysplashsurf/tests"
rnd = pf.read_vtk(os.path.join(D, "ParticleData_Random_1000.vtk"))
fl = pf.read_vtk(os.path.join(D, "ParticleData_Fluid_5.vtk"))
bg, battr = pf.read_bgeo(os.path.join(D, "ParticleData_Fluid_50.bgeo"))
assert len(bg) == 4732                                                   # test_bgeo.py
out = os.path.join(ROOT, "tests", "golden", "pysplashsurf_tests.npz")
np.savez_compressed(out, random_1000=rnd.points.astype(np.float32), fluid_5=fl.points.astype(np.float32), fluid_5_velocity=fl.point_data["velocity"],
                    fluid_5_id=fl.point_data["id"], fluid_50_bgeo=bg)
print(out, os.path.getsize(out) // 1024, "KiB")
