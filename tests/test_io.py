"""CPU tests of the harness I/O helpers against the reference binary's own reader / writer."""
import os

import numpy as np
import pytest


def test_xyz_roundtrip_and_obj_roundtrip(tmp_path):
    from splashsurf_b200 import io
    p = np.random.default_rng(0).normal(size=(257, 3)).astype(np.float32)
    io.write_xyz(str(tmp_path / "p.xyz"), p)
    assert np.array_equal(io.read_xyz(str(tmp_path / "p.xyz")), p)
    v = np.array([[0, 0, 0], [1, 0, 0.5], [0, 1e-7, 2.25]], np.float32)
    t = np.array([[0, 1, 2]])
    io.write_obj(str(tmp_path / "m.obj"), v, t)
    v2, t2 = io.read_obj(str(tmp_path / "m.obj"))
    assert np.array_equal(v2, v) and np.array_equal(t2, t)          # shortest round-trip formatting is lossless


@pytest.mark.skipif(not os.path.exists("/root/reference"), reason="reference tree only exists in the build container")
def test_obj_and_xyz_interoperate_with_reference_cli(tmp_path, oracle_mod):
    """The reference CLI reads our .xyz and its OBJ output parses with our reader to the mesh its API returns."""
    if not oracle_mod.reference_available():
        pytest.skip("oracle/_ref not unpacked")
    from splashsurf_b200 import io, synthetic as syn
    ps = oracle_mod.reference()
    p = syn.jittered_cube(8, 0.025, 5)
    xyz, obj = str(tmp_path / "p.xyz"), str(tmp_path / "out.obj")
    io.write_xyz(xyz, p)
    ps.run_splashsurf(["splashsurf", "reconstruct", xyz, "-r=0.025", "-l=2.0", "-c=0.75", "-o", obj, "-q"])
    v, t = io.read_obj(obj)
    r = ps.reconstruct_surface(p, particle_radius=0.025, smoothing_length=2.0, cube_size=0.75, subdomain_grid_auto_disable=False)
    assert len(v) == len(r.mesh.vertices) and len(t) == len(r.mesh.triangles)
    ours = str(tmp_path / "ours.obj")
    io.write_obj(ours, np.asarray(r.mesh.vertices), np.asarray(r.mesh.triangles))
    v2, t2 = io.read_obj(ours)
    assert np.array_equal(v2, np.asarray(r.mesh.vertices)) and np.array_equal(t2, np.asarray(r.mesh.triangles).astype(np.int64))
