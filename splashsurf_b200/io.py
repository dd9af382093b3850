"""Particle / mesh file helpers for the harness: the formats the reference's CLI reads and writes around
the hot path.  Nothing here is on the timed path.

* `.xyz`   raw native-endian f32 triples                      (splashsurf_lib/src/io/xyz_format.rs:10-37)
* `.vtk` / `.vtu` / `.ply` / `.bgeo` / `.json` particle files: particle_formats.py (every format the reference reads / writes)
* `.obj`   `v x y z` / optional `vn` / `f a b c` (1-based; `a//a` with normals)   (splashsurf_lib/src/io/obj_format.rs:17-71)
* `.ply`   binary little endian, attributes inside the vertex records             (splashsurf_lib/src/io/ply_format.rs:190-267)

The mesh writers are native (csrc/ss_meshio.inc behind ss_write_mesh_f32): multi-threaded, and byte for byte the files the
reference CLI writes for the same mesh (tests/test_io.py).  The readers are numpy.
"""
from __future__ import annotations

import numpy as np


def read_xyz(path: str) -> np.ndarray:
    raw = np.fromfile(path, dtype=np.float32)
    return np.ascontiguousarray(raw[: (len(raw) // 3) * 3].reshape(-1, 3))


def write_xyz(path: str, particles: np.ndarray) -> None:
    np.ascontiguousarray(particles, dtype=np.float32).tofile(path)


def read_vtk_points(path: str) -> np.ndarray:
    """POINTS of a legacy (ASCII / BINARY, float / double) or XML VTK file as float32 (vtk_format.rs:141-155)."""
    from .particle_formats import read_vtk_particles
    return read_vtk_particles(path)


def read_particles(path: str) -> np.ndarray:
    """`particles_from_file` (splashsurf_lib/src/io.rs:17-43): .vtk / .vtu / .xyz / .ply / .bgeo / .json by extension (+ .npy)."""
    if str(path).lower().endswith(".npy"):
        return np.ascontiguousarray(np.load(path), dtype=np.float32).reshape(-1, 3)
    from .particle_formats import particles_from_file
    return particles_from_file(path)


def write_particles(path: str, particles: np.ndarray, enable_compression: bool = True) -> None:
    """`write_particle_positions` (splashsurf/src/io.rs:195-235): .vtk / .bgeo / .json by extension (+ .xyz, .npy for the harness)."""
    low = str(path).lower()
    if low.endswith(".xyz"):
        return write_xyz(path, particles)
    if low.endswith(".npy"):
        return np.save(path, np.ascontiguousarray(particles, dtype=np.float32))
    from .particle_formats import write_particle_positions
    write_particle_positions(path, particles, enable_compression)


def write_mesh(path, mesh, **kw) -> None:
    """`splashsurf::io::write_mesh` (splashsurf/src/io.rs:276-316): .vtk / .ply / .obj by extension, through the library's
    multi-threaded writer (ss_write_mesh_f32); see `splashsurf_b200.write_mesh`."""
    from . import write_mesh as _w
    _w(path, mesh, **kw)


def write_obj(path: str, vertices: np.ndarray, triangles: np.ndarray, normals: np.ndarray | None = None) -> None:
    """obj_format.rs:17-71: numbers as Rust's `{}` prints them (shortest round trip, no exponent, no trailing ".0")."""
    write_mesh(path, (vertices, triangles), point_attributes=({"normals": normals} if normals is not None else None), file_format="obj")


def read_obj(path: str):
    verts, tris = [], []
    for line in open(path):
        if line.startswith("v "):
            verts.append([float(x) for x in line.split()[1:4]])
        elif line.startswith("f "):
            tris.append([int(tok.split("/")[0]) - 1 for tok in line.split()[1:4]])
    return np.asarray(verts, dtype=np.float32).reshape(-1, 3), np.asarray(tris, dtype=np.int64).reshape(-1, 3)


def write_vtk_mesh(path: str, vertices: np.ndarray, triangles: np.ndarray, point_attributes: dict | None = None) -> None:
    """Legacy VTK unstructured grid (binary, big-endian), triangles as cell type 5, laid out as the reference's writer does
    (vtk_format.rs:188-211 through vtkio)."""
    write_mesh(path, (vertices, triangles), point_attributes=point_attributes, file_format="vtk")


def write_ply(path: str, vertices: np.ndarray, triangles: np.ndarray, point_attributes: dict | None = None) -> None:
    """ply_format.rs:190-267: binary little endian, point attributes inside the vertex records."""
    write_mesh(path, (vertices, triangles), point_attributes=point_attributes, file_format="ply")


def read_ply_mesh(path: str):
    """Reads a binary little-endian PLY as `write_ply` / the reference's `mesh_to_ply` produce it.  Returns
    ``(vertices, triangles, quads, point_attributes)``; nx / ny / nz come back as "normals", <name>_x/_y/_z as one (n, 3) array."""
    b = open(path, "rb").read()
    k = b.index(b"end_header\n") + 11
    lines = b[:k].decode().split("\n")
    if lines[0] != "ply" or lines[1] != "format binary_little_endian 1.0":
        raise ValueError("only binary little-endian PLY files are read")
    nv = nf = 0
    props, section = [], None
    for ln in lines[2:]:
        tok = ln.split()
        if tok[:1] == ["element"]:
            section = tok[1]
            if section == "vertex":
                nv = int(tok[2])
            elif section == "face":
                nf = int(tok[2])
        elif tok[:1] == ["property"] and section == "vertex":
            if tok[1] not in ("float", "uint"):
                raise ValueError(f"unsupported vertex property type {tok[1]}")
            props.append((tok[2], "<f4" if tok[1] == "float" else "<u4"))
    rec = np.frombuffer(b, dtype=np.dtype(props), count=nv, offset=k)
    verts = np.stack([rec["x"], rec["y"], rec["z"]], axis=1).astype(np.float32) if nv else np.zeros((0, 3), np.float32)
    attrs, names, i = {}, [p[0] for p in props[3:]], 0
    while i < len(names):
        n = names[i]
        if n == "nx" and names[i:i + 3] == ["nx", "ny", "nz"]:
            attrs["normals"] = np.stack([rec["nx"], rec["ny"], rec["nz"]], axis=1)
            i += 3
        elif n.endswith("_x") and names[i:i + 3] == [n[:-2] + s for s in ("_x", "_y", "_z")]:
            attrs[n[:-2]] = np.stack([rec[n], rec[names[i + 1]], rec[names[i + 2]]], axis=1)
            i += 3
        else:
            attrs[n] = rec[n].astype(np.uint64) if rec.dtype[n].kind == "u" else rec[n].copy()
            i += 1
    o = k + nv * rec.dtype.itemsize
    tris, quads = [], []
    body = np.frombuffer(b, dtype=np.uint8, offset=o)
    if nf and len(body) == 13 * nf and np.all(body[::13] == 3):                 # the common case: triangles only
        tris = np.frombuffer(b, dtype=np.dtype([("n", "u1"), ("i", "<u4", 3)]), count=nf, offset=o)["i"]
    else:
        p = 0
        for _ in range(nf):
            n = int(body[p])
            idx = np.frombuffer(body[p + 1:p + 1 + 4 * n].tobytes(), "<u4")
            (tris if n == 3 else quads).append(idx)
            p += 1 + 4 * n
    return (verts, np.asarray(tris, dtype=np.uint32).reshape(-1, 3), np.asarray(quads, dtype=np.uint32).reshape(-1, 4), attrs)


def read_vtk_mesh(path: str):
    """Reads a legacy binary VTK unstructured grid as `write_vtk_mesh` / the reference's `write_vtk` produce it.  Returns
    ``(vertices, triangles, quads, point_attributes, cell_attributes)``."""
    b = open(path, "rb").read()

    def block(key, start):
        k = b.index(key, start)
        e = b.index(b"\n", k)
        return b[k:e].split(), e + 1
    tok, o = block(b"POINTS", 0)
    nv = int(tok[1])
    verts = np.frombuffer(b, ">f4", 3 * nv, o).reshape(nv, 3).astype("<f4")
    tok, o = block(b"CELLS", o + 12 * nv)
    nc, size = int(tok[1]), int(tok[2])
    cells = np.frombuffer(b, ">u4", size, o).astype(np.uint32)
    tok, o = block(b"CELL_TYPES", o + 4 * size)
    types = np.frombuffer(b, ">i4", nc, o)
    nt = int(np.count_nonzero(types == 5))
    if nt + int(np.count_nonzero(types == 9)) != nc:
        raise ValueError("only triangle and quad cells are read")
    tris = cells[:4 * nt].reshape(-1, 4)[:, 1:]
    quads = cells[4 * nt:].reshape(-1, 5)[:, 1:]
    o += 4 * nc
    out = []
    for key, n in ((b"POINT_DATA", nv), (b"CELL_DATA", nc)):
        tok, o = block(key, o)
        attrs = {}
        while b[o:o + 9] == b"\nSCALARS " or b[o:o + 8] == b"SCALARS ":
            tok, o = block(b"SCALARS", o)
            name, typ, comps = tok[1].decode(), tok[2], int(tok[3])
            o = b.index(b"\n", o) + 1                                             # LOOKUP_TABLE default
            dt = {b"float": ">f4", b"unsigned_long": ">u8"}[typ]
            a = np.frombuffer(b, dt, n * comps, o)
            o += a.nbytes + 1
            attrs[name] = a.astype(a.dtype.newbyteorder("<")).reshape((n, comps) if comps > 1 else (n,))
        out.append(attrs)
    return verts, np.ascontiguousarray(tris), np.ascontiguousarray(quads), out[0], out[1]


def read_vtk_point_data(path: str) -> dict:
    """POINT_DATA arrays of a legacy (ASCII / BINARY: SCALARS + lookup table, VECTORS / NORMALS, FIELD arrays) or XML VTK file:
    name -> array of shape (n,) or (n, components) in the file's type."""
    from .particle_formats import read_vtk
    return dict(read_vtk(path).point_data.items())


def read_particle_attributes(path: str, names) -> dict:
    """The `-a / --interpolate_attribute` inputs of the reference CLI (reconstruct.rs:196-198; splashsurf/src/io.rs:66-190): named point
    attributes of a VTK / VTU (vtk_format.rs:96-140, :318-346) or BGEO (bgeo_format.rs:45-66) particle file, converted like the reference
    converts them -- scalars stored as u32 / f32 / f64 and 3-vectors stored as f32 / f64 become float32; anything else is an error."""
    from .particle_formats import particle_attributes_from_file
    return particle_attributes_from_file(path, names)
