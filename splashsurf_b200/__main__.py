"""`python -m splashsurf_b200 reconstruct <particles> -r <radius> -l <smoothing length> -c <cube size> [-o out.obj]`

and `python -m splashsurf_b200 convert (--particles <file> | --mesh <file>) -o <file> [--overwrite] [--domain-min x y z --domain-max x y z]`
(splashsurf/src/convert.rs).

A thin stand-in for `splashsurf reconstruct` (splashsurf/src/reconstruct.rs:36-380): the relative `-l` / `-c` values are multiplied
by the particle radius like the reference CLI does (reconstruct.rs:628-629); the post-processing switches (`--mesh-cleanup`,
`--decimate-barnacles`, `--mesh-smoothing-iters`, `--normals`, `--sph-normals`, `--generate-quads`, `--mesh-aabb-min/-max`,
`--check-mesh`, `-a <attribute>` from a VTK / VTU / BGEO particle file) go through `reconstruction_pipeline` with the reference's option names and defaults, and the output file (`.vtk`,
`.ply`, `.obj` with the attributes the reference writes) comes from the library's writer, byte for byte the reference CLI's file."""
import argparse
import sys
import time

import numpy as np


def convert(a) -> int:
    """`convert_subcommand` (splashsurf/src/convert.rs:58-152): particles .vtk / .vtu / .bgeo / .ply / .xyz / .json -> .vtk / .bgeo / .json
    (optionally filtered by a half-open domain box, aabb.rs:220-222), or a triangle mesh .vtk / .ply -> .obj / .vtk / .ply."""
    import os
    from . import io, particle_formats as pf
    if (a.input_particles is None) == (a.input_mesh is None):
        print("Aborting: " + ("No input file specified, either a particle or mesh input file has to be specified." if a.input_particles is None
                              else "the argument '--particles' cannot be used with '--mesh'"), file=sys.stderr)
        return 1
    if (a.domain_min is None) != (a.domain_max is None):
        print("Aborting: --domain-min and --domain-max have to be specified together", file=sys.stderr)
        return 1
    if not a.overwrite and os.path.exists(a.output_file):
        print(f'Aborting: Output file "{a.output_file}" already exists. Use overwrite flag to ignore this.', file=sys.stderr)
        return 1
    try:
        if a.input_particles is not None:
            p = pf.particles_from_file(a.input_particles)
            if a.domain_min is not None:
                lo, hi = np.asarray(a.domain_min, np.float64).astype(np.float32), np.asarray(a.domain_max, np.float64).astype(np.float32)
                p = p[np.all(p >= lo, axis=1) & np.all(p < hi, axis=1)]
            pf.write_particle_positions(a.output_file, p)
        else:
            ext = os.path.splitext(a.input_mesh)[1].lower()
            if ext == ".vtk":
                v, t = pf.read_vtk_surface_mesh(a.input_mesh)
                attrs = {}
            elif ext == ".ply":
                v, t, attrs = pf.read_ply_surface_mesh(a.input_mesh)
            elif not ext:
                raise ValueError("Unable to detect file format of mesh input file (file name has to end with supported extension)")
            else:
                raise ValueError(f'Unsupported file format extension "{ext[1:]}" for reading surface meshes')
            io.write_mesh(a.output_file, (v, t), point_attributes=attrs or None)
    except (ValueError, OSError) as e:
        print(f"Error occurred: {e}", file=sys.stderr)
        return 1
    return 0


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
    r.add_argument("-a", "--interpolate_attribute", dest="interpolate_attributes", action="append", default=[], metavar="ATTRIBUTE_NAME")
    # post-processing, names and defaults of the reference CLI (reconstruct.rs:150-300)
    r.add_argument("--normals", choices=["on", "off"], default="off")
    r.add_argument("--normals-smoothing-iters", type=int, default=None)
    r.add_argument("--mesh-cleanup", choices=["on", "off"], default=None)        # on by default when smoothing runs (reconstruct.rs:200-213)
    r.add_argument("--mesh-cleanup-snap-dist", type=float, default=None)
    r.add_argument("--decimate-barnacles", choices=["on", "off"], default="off")
    r.add_argument("--keep-verts", choices=["on", "off"], default="off")
    r.add_argument("--mesh-smoothing-iters", type=int, default=None)
    r.add_argument("--mesh-smoothing-weights", choices=["on", "off"], default="off")
    r.add_argument("--mesh-smoothing-weights-normalization", type=float, default=13.0)
    r.add_argument("--output-smoothing-weights", choices=["on", "off"], default="off")
    r.add_argument("--output-raw-normals", choices=["on", "off"], default="off")
    r.add_argument("--generate-quads", choices=["on", "off"], default="off")
    r.add_argument("--quad-max-edge-diag-ratio", type=float, default=1.75)
    r.add_argument("--quad-max-normal-angle", type=float, default=10.0)
    r.add_argument("--quad-max-interior-angle", type=float, default=135.0)
    r.add_argument("--mesh-aabb-min", type=float, nargs=3, default=None)
    r.add_argument("--mesh-aabb-max", type=float, nargs=3, default=None)
    r.add_argument("--mesh-aabb-clamp-verts", choices=["on", "off"], default="off")
    r.add_argument("--check-mesh", choices=["on", "off"], default="off")
    r.add_argument("-o", "--output-file", default=None)
    # `splashsurf convert` (splashsurf/src/convert.rs:13-56)
    cv = sub.add_parser("convert")
    cv.add_argument("--particles", dest="input_particles", default=None)
    cv.add_argument("--mesh", dest="input_mesh", default=None)
    cv.add_argument("-o", dest="output_file", required=True)
    cv.add_argument("--overwrite", action="store_true")
    cv.add_argument("--domain-min", type=float, nargs=3, default=None, metavar=("X_MIN", "Y_MIN", "Z_MIN"))
    cv.add_argument("--domain-max", type=float, nargs=3, default=None, metavar=("X_MAX", "Y_MAX", "Z_MAX"))
    a = ap.parse_args(argv)
    if a.cmd == "convert":
        return convert(a)
    from . import MeshWithData, io, reconstruct_surface, reconstruction_pipeline
    p = io.read_particles(a.input)
    t = time.perf_counter()
    on = lambda v: v == "on"         # noqa: E731
    # the reference CLI always uses the subdomain grid when it is on (auto_disable is inverted there)
    base = dict(particle_radius=a.particle_radius, rest_density=a.rest_density, smoothing_length=a.smoothing_length, cube_size=a.cube_size,
                iso_surface_threshold=a.surface_threshold, simd=on(a.simd), subdomain_grid=on(a.subdomain_grid), subdomain_grid_auto_disable=False,
                subdomain_num_cubes_per_dim=a.subdomain_cubes)
    if a.mesh_cleanup is None:
        a.mesh_cleanup = "on" if a.mesh_smoothing_iters not in (None, 0) else "off"
    attrs = io.read_particle_attributes(a.input, a.interpolate_attributes)
    for name, arr in attrs.items():
        if arr.dtype.kind != "f":                      # BGEO Int attributes load as u64; the reference cannot interpolate them either (reconstruct.rs:1387)
            raise ValueError(f'Interpolation of this attribute type not implemented (attribute "{name}")')
    post = any([bool(attrs), on(a.normals), on(a.mesh_cleanup), on(a.decimate_barnacles), a.mesh_smoothing_iters is not None, on(a.generate_quads),
                a.mesh_aabb_min is not None, on(a.check_mesh)])
    if post:
        # --sph-normals only selects how --normals are computed (reconstruct.rs:1094-1149)
        out, _ = reconstruction_pipeline(p, attributes_to_interpolate=attrs, compute_normals=on(a.normals), sph_normals=on(a.sph_normals),
                                         normals_smoothing_iters=a.normals_smoothing_iters, mesh_smoothing_iters=a.mesh_smoothing_iters,
                                         mesh_smoothing_weights=on(a.mesh_smoothing_weights),
                                         mesh_smoothing_weights_normalization=a.mesh_smoothing_weights_normalization,
                                         output_mesh_smoothing_weights=on(a.output_smoothing_weights), output_raw_normals=on(a.output_raw_normals),
                                         mesh_cleanup=on(a.mesh_cleanup),
                                         mesh_cleanup_snap_dist=a.mesh_cleanup_snap_dist, decimate_barnacles=on(a.decimate_barnacles),
                                         keep_vertices=on(a.keep_verts), generate_quads=on(a.generate_quads),
                                         quad_max_edge_diag_ratio=a.quad_max_edge_diag_ratio, quad_max_normal_angle=a.quad_max_normal_angle,
                                         quad_max_interior_angle=a.quad_max_interior_angle, mesh_aabb_min=a.mesh_aabb_min, mesh_aabb_max=a.mesh_aabb_max,
                                         mesh_aabb_clamp_vertices=on(a.mesh_aabb_clamp_verts), check_mesh_closed=on(a.check_mesh),
                                         check_mesh_manifold=on(a.check_mesh), check_mesh_orientation=on(a.check_mesh), **base)
    else:
        out = MeshWithData(reconstruct_surface(p, **base).mesh, {}, {})
    dt = time.perf_counter() - t
    quads = out.mesh.get_quads() if on(a.generate_quads) else None
    tris = out.mesh.get_triangles() if on(a.generate_quads) else out.mesh.triangles
    print(f"{len(p)} particles -> {out.nvertices} vertices, {len(tris)} triangles" + (f", {len(quads)} quads" if quads is not None else "") + f" in {dt:.3f} s",
          file=sys.stderr)
    if a.output_file:
        if a.output_file.endswith(".npz"):
            np.savez(a.output_file, vertices=out.mesh.vertices, triangles=tris, **({"quads": quads} if quads is not None else {}), **out.point_attributes)
        else:
            t = time.perf_counter()
            io.write_mesh(a.output_file, out)              # .vtk / .ply / .obj as the reference CLI writes them (io.rs:276-316)
            print(f"wrote {a.output_file} in {time.perf_counter() - t:.3f} s", file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
