"""Particle / mesh file helpers for the harness: the formats the reference's CLI reads and writes around
the hot path.  Nothing here is on the timed path.

* `.xyz`   raw native-endian f32 triples                      (splashsurf_lib/src/io/xyz_format.rs:10-37)
* `.vtk`   legacy VTK, BINARY big-endian float POINTS         (fixtures under data/)
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


_VTK_TYPES = {b"float": ">f4", b"double": ">f8", b"int": ">i4", b"unsigned_int": ">u4", b"long": ">i8", b"unsigned_long": ">u8",
              b"short": ">i2", b"unsigned_short": ">u2", b"char": ">i1", b"unsigned_char": ">u1"}


def read_vtk_point_data(path: str) -> dict:
    """POINT_DATA arrays of a legacy BINARY VTK file as written by SPlisHSPlasH and by the reference (SCALARS + lookup table, VECTORS / NORMALS,
    FIELD arrays): name -> array of shape (n,) or (n, components) in the file's type."""
    b = open(path, "rb").read()
    if b.split(b"\n", 3)[2].strip().upper() != b"BINARY":
        raise ValueError("only BINARY legacy VTK files are read")
    k = b.find(b"POINT_DATA")
    if k < 0:
        return {}

    def line(o):
        while o < len(b) and b[o:o + 1] in b" \n\r\t":
            o += 1
        e = b.find(b"\n", o)
        e = len(b) if e < 0 else e
        return b[o:e].split(), e + 1

    def array(o, typ, n):
        if typ not in _VTK_TYPES:
            raise ValueError(f"unsupported VTK data type {typ.decode()}")
        a = np.frombuffer(b, _VTK_TYPES[typ], n, o)
        return a.astype(a.dtype.newbyteorder("<")), o + a.nbytes
    tok, o = line(k)
    n = int(tok[1])
    out = {}
    while True:
        tok, o2 = line(o)
        if not tok or tok[0] in (b"CELL_DATA", b"POINT_DATA"):
            break
        kind = tok[0]
        if kind == b"SCALARS":
            comps = int(tok[3]) if len(tok) > 3 else 1
            tok2, o3 = line(o2)
            if tok2[:1] == [b"LOOKUP_TABLE"]:
                o2 = o3
            a, o = array(o2, tok[2], n * comps)
            out[tok[1].decode()] = a.reshape(n, comps) if comps > 1 else a
        elif kind in (b"VECTORS", b"NORMALS"):
            a, o = array(o2, tok[2], 3 * n)
            out[tok[1].decode()] = a.reshape(n, 3)
        elif kind == b"FIELD":
            o = o2
            for _ in range(int(tok[2])):
                t2, o = line(o)
                comps, tuples = int(t2[1]), int(t2[2])
                a, o = array(o, t2[3], comps * tuples)
                out[t2[0].decode()] = a.reshape(tuples, comps) if comps > 1 else a
        else:
            raise ValueError(f"unsupported POINT_DATA entry {kind.decode()}")
    return out


def read_particle_attributes(path: str, names) -> dict:
    """The `-a / --interpolate_attribute` inputs of the reference CLI (reconstruct.rs:196-198; vtk_format.rs:96-140, :318-346): named point
    attributes of a VTK particle file as float32 -- scalars stored as u32 / f32 / f64, 3-vectors stored as f32 / f64; anything else is an
    error, like in the reference."""
    if not names:
        return {}
    if not path.endswith(".vtk"):
        raise ValueError("attributes are read from legacy .vtk particle files only")
    data = read_vtk_point_data(path)
    out = {}
    for name in names:
        if name not in data:
            raise ValueError(f"Attribute {name} not found in VTK file")
        a = data[name]
        if a.ndim == 1 and (a.dtype.kind == "f" or a.dtype == np.uint32):
            out[name] = np.ascontiguousarray(a, dtype=np.float32)
        elif a.ndim == 2 and a.shape[1] == 3 and a.dtype.kind == "f":
            out[name] = np.ascontiguousarray(a, dtype=np.float32)
        elif a.ndim == 1 or a.shape[1] == 3:
            raise ValueError(f'Attribute "{name}": unsupported VTK data type for {"scalars" if a.ndim == 1 else "vectors"}')
        else:
            raise ValueError(f'Attribute "{name}": unsupported number of components ({a.shape[1]}) in VTK IO buffer')
    return out
