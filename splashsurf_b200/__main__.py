"""`python -m splashsurf_b200 reconstruct <particles> -r <radius> -l <smoothing length> -c <cube size> [-o out.obj]`

A thin stand-in for `splashsurf reconstruct` (splashsurf/src/reconstruct.rs:36-380) limited to the hot path: the relative
`-l` / `-c` values are multiplied by the particle radius like the reference CLI does (reconstruct.rs:628-629)."""
import argparse
import sys
import time

import numpy as np


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m splashsurf_b200")
    sub = ap.add_subparsers(dest="cmd", required=True)
    r = sub.add_parser("reconstruct")
    r.add_argument("input")
    r.add_argument("-r", "--particle-radius", type=float, required=True)
    r.add_argument("-l", "--smoothing-length", type=float, required=True)
    r.add_argument("-c", "--cube-size", type=float, required=True)
    r.add_argument("-t", "--surface-threshold", type=float, default=0.6)
    r.add_argument("--rest-density", type=float, default=1000.0)
    r.add_argument("--subdomain-cubes", type=int, default=64)
    r.add_argument("--subdomain-grid", choices=["on", "off"], default="on")
    r.add_argument("--simd", choices=["on", "off"], default="on")
    r.add_argument("--sph-normals", choices=["on", "off"], default="off")
    r.add_argument("-o", "--output-file", default=None)
    a = ap.parse_args(argv)
    from . import io, reconstruct_surface
    p = io.read_particles(a.input)
    t = time.perf_counter()
    # the reference CLI always uses the subdomain grid when it is on (auto_disable is inverted there)
    res = reconstruct_surface(p, particle_radius=a.particle_radius, rest_density=a.rest_density, smoothing_length=a.smoothing_length,
                              cube_size=a.cube_size, iso_surface_threshold=a.surface_threshold, simd=a.simd == "on",
                              subdomain_grid=a.subdomain_grid == "on", subdomain_grid_auto_disable=False,
                              subdomain_num_cubes_per_dim=a.subdomain_cubes, sph_normals=a.sph_normals == "on")
    dt = time.perf_counter() - t
    print(f"{len(p)} particles -> {res.mesh.nvertices} vertices, {res.mesh.ncells} triangles in {dt:.3f} s", file=sys.stderr)
    if a.output_file:
        if a.output_file.endswith(".obj"):
            io.write_obj(a.output_file, res.mesh.vertices, res.mesh.triangles, res.normals)
        elif a.output_file.endswith(".vtk"):
            io.write_vtk_mesh(a.output_file, res.mesh.vertices, res.mesh.triangles)
        else:
            np.savez(a.output_file, vertices=res.mesh.vertices, triangles=res.mesh.triangles)
    return 0


if __name__ == "__main__":
    sys.exit(main())
