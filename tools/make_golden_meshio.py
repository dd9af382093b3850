#!/usr/bin/env python3
"""Golden files for the mesh writers: the REFERENCE CLI (oracle/_ref, `run_splashsurf`) reconstructs a tiny cloud and
writes the mesh as .vtk, .ply and .obj -- once with smoothing weights + normals as attributes, once as a mixed triangle / quad mesh with
normals.  tests/test_io.py reads the .ply (it holds every value), writes the three formats with the library and compares whole files.

    python tools/make_golden_meshio.py          # -> tests/golden/meshio_{attr,quad}.{vtk,ply,obj}, meshio_particles.xyz
"""
import os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
BASE = ["-r=0.025", "-l=2.0", "-c=1.0"]
CASES = {
    "attr": ["--mesh-smoothing-iters=2", "--mesh-smoothing-weights=on", "--output-smoothing-weights=on", "--normals=on"],
    "quad": ["--generate-quads=on", "--normals=on"],
}


def run_reference_cli(args):
    """One CLI run per process: the reference installs its logger once."""
    code = "import sys; sys.path.insert(0, %r); import oracle; oracle.reference().run_splashsurf(['splashsurf'] + sys.argv[1:])" % ROOT
    subprocess.check_call([sys.executable, "-c", code] + list(args))


def main():
    from splashsurf_b200 import io, synthetic as syn
    xyz = os.path.join(GOLD, "meshio_particles.xyz")
    io.write_xyz(xyz, syn.jittered_cube(4, 0.025, 3))
    for name, flags in CASES.items():
        for ext in ("vtk", "ply", "obj"):
            out = os.path.join(GOLD, f"meshio_{name}.{ext}")
            run_reference_cli(["reconstruct", xyz, *BASE, *flags, "-o", out, "-q"])
            print(out, os.path.getsize(out))


if __name__ == "__main__":
    main()
