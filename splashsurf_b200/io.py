"""Particle / mesh file helpers for the harness: the formats the reference's CLI reads and writes around
the hot path.  Plain numpy; nothing here is on the timed path.

* `.xyz`   raw native-endian f32 triples                      (splashsurf_lib/src/io/xyz_format.rs:10-37)
* `.vtk`   legacy VTK, BINARY big-endian float POINTS         (fixtures under data/)
* `.obj`   `v x y z` / optional `vn` / `f a b c` (1-based; `a//a` with normals)   (splashsurf_lib/src/io/obj_format.rs:17-71)
"""
from __future__ import annotations

import numpy as np


def read_xyz(path: str) -> np.ndarray:
    raw = np.fromfile(path, dtype=np.float32)
    return np.ascontiguousarray(raw[: (len(raw) // 3) * 3].reshape(-1, 3))


def write_xyz(path: str, particles: np.ndarray) -> None:
    np.ascontiguousarray(particles, dtype=np.float32).tofile(path)


def read_vtk_points(path: str) -> np.ndarray:
    b = open(path, "rb").read()
    k = b.index(b"POINTS")
    e = b.index(b"\n", k)
    n = int(b[k:e].split()[1])
    return np.frombuffer(b[e + 1:e + 1 + 12 * n], dtype=">f4").reshape(n, 3).astype("<f4")


def read_particles(path: str) -> np.ndarray:
    if path.endswith(".xyz"):
        return read_xyz(path)
    if path.endswith(".vtk"):
        return read_vtk_points(path)
    if path.endswith(".npy"):
        return np.ascontiguousarray(np.load(path), dtype=np.float32).reshape(-1, 3)
    raise ValueError(f"unsupported particle file: {path}")


def write_obj(path: str, vertices: np.ndarray, triangles: np.ndarray, normals: np.ndarray | None = None) -> None:
    v = np.asarray(vertices, dtype=np.float32)
    t = np.asarray(triangles).astype(np.int64) + 1
    with open(path, "w") as f:
        # Rust's `{}` prints the shortest representation that round-trips; repr(float(f32)) of the widened value would
        # print f64 digits, so go through numpy's shortest f32 formatting
        for row in v:
            f.write("v %s %s %s\n" % tuple(np.format_float_positional(x, unique=True, trim="0") if abs(x) < 1e16 and (abs(x) >= 1e-5 or x == 0)
                                           else np.format_float_scientific(x, unique=True) for x in row))
        if normals is not None:
            for row in np.asarray(normals, dtype=np.float32):
                f.write("vn %s %s %s\n" % tuple(np.format_float_positional(x, unique=True, trim="0") for x in row))
            for a, b, c in t:
                f.write(f"f {a}//{a} {b}//{b} {c}//{c}\n")
        else:
            for a, b, c in t:
                f.write(f"f {a} {b} {c}\n")


def read_obj(path: str):
    verts, tris = [], []
    for line in open(path):
        if line.startswith("v "):
            verts.append([float(x) for x in line.split()[1:4]])
        elif line.startswith("f "):
            tris.append([int(tok.split("/")[0]) - 1 for tok in line.split()[1:4]])
    return np.asarray(verts, dtype=np.float32).reshape(-1, 3), np.asarray(tris, dtype=np.int64).reshape(-1, 3)


def write_vtk_mesh(path: str, vertices: np.ndarray, triangles: np.ndarray) -> None:
    """Legacy VTK unstructured grid (binary, big-endian), triangles as cell type 5."""
    v = np.asarray(vertices, dtype=">f4")
    t = np.asarray(triangles).astype(">i4")
    with open(path, "wb") as f:
        f.write(b"# vtk DataFile Version 4.2\nmesh\nBINARY\nDATASET UNSTRUCTURED_GRID\n")
        f.write(f"POINTS {len(v)} float\n".encode())
        f.write(v.tobytes())
        cells = np.concatenate([np.full((len(t), 1), 3, dtype=">i4"), t], axis=1)
        f.write(f"\nCELLS {len(t)} {len(t) * 4}\n".encode())
        f.write(cells.tobytes())
        f.write(f"\nCELL_TYPES {len(t)}\n".encode())
        f.write(np.full(len(t), 5, dtype=">i4").tobytes())
        f.write(b"\n")
