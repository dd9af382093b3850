#!/usr/bin/env python3
"""Mesh-writer timing beside the reference's own ("the mesh write is 24 % of CLI wall time on fine grids").  The REFERENCE CLI
(oracle/_ref) reconstructs a jittered cube and writes the mesh with normals as .obj / .vtk / .ply; its profiling tree gives the time of
`write surface mesh to file`.  The library's writer (ss_write_mesh_f32, host threads) then writes the SAME mesh -- read back from the
reference's .ply -- and the files are compared byte for byte (the reference runs with -n=1: with several threads the vertex order of its mesh changes from run to run, and
the three files have to hold the same mesh).  No GPU involved; runs on whatever host cores the box has.

    python tools/bench_meshio.py --side 100 --cube 0.5 [--json out.json]
"""
import argparse, json, os, re, subprocess, sys, tempfile, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=100, help="particles per axis of the jittered cube")
    ap.add_argument("--cube", type=float, default=0.5, help="cube size in particle radii")
    ap.add_argument("--repeat", type=int, default=5)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    import splashsurf_b200 as ss
    from splashsurf_b200 import io, synthetic as syn
    code = "import sys; sys.path.insert(0, %r); import oracle; oracle.reference().run_splashsurf(['splashsurf'] + sys.argv[1:])" % ROOT
    out = {"host_threads": os.cpu_count(), "particles": a.side ** 3, "formats": {}}
    with tempfile.TemporaryDirectory() as d:
        xyz = os.path.join(d, "p.xyz")
        io.write_xyz(xyz, syn.jittered_cube(a.side, 0.025, 1))
        for ext in ("ply", "obj", "vtk"):
            ref = os.path.join(d, f"ref.{ext}")
            best = None
            for _ in range(a.repeat):
                log = subprocess.run([sys.executable, "-c", code, "reconstruct", xyz, "-r=0.025", "-l=2.0", f"-c={a.cube}", "--normals=on", "-n=1", "-o", ref],
                                     check=True, capture_output=True, text=True)
                m = re.search(r"write surface mesh to file: [\d.]+%, ([\d.]+)ms", log.stdout + log.stderr)
                best = float(m.group(1)) if best is None else min(best, float(m.group(1)))
            out["formats"][ext] = {"reference_ms": best, "bytes": os.path.getsize(ref)}
        v, t, q, attrs = io.read_ply_mesh(os.path.join(d, "ref.ply"))
        out["vertices"], out["triangles"] = len(v), len(t)
        mwd = ss.MeshWithData(ss.TriMesh3d(v, t), attrs, {})
        for ext in ("ply", "obj", "vtk"):
            ours = os.path.join(d, f"ours.{ext}")
            for threads in (1, 0):
                best = None
                for _ in range(a.repeat):
                    t0 = time.perf_counter()
                    ss.write_mesh(ours, mwd, threads=threads)
                    dt = (time.perf_counter() - t0) * 1e3
                    best = dt if best is None else min(best, dt)
                out["formats"][ext]["library_ms_1_thread" if threads == 1 else "library_ms"] = round(best, 2)
            same = open(ours, "rb").read() == open(os.path.join(d, f"ref.{ext}"), "rb").read()
            out["formats"][ext]["identical"] = same
            f = out["formats"][ext]
            f["speedup"] = round(f["reference_ms"] / f["library_ms"], 2)
            f["library_GBps"] = round(f["bytes"] / f["library_ms"] / 1e6, 2)
    print(json.dumps(out))
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)
    return 0 if all(f["identical"] for f in out["formats"].values()) else 1


if __name__ == "__main__":
    sys.exit(main())
