"""Particle file formats either side of the hot path: everything `splashsurf_lib::io::particles_from_file`
reads (splashsurf_lib/src/io.rs:17-43: .vtk, .vtu, .xyz, .ply, .bgeo, .json) and everything `write_particle_positions` writes
(splashsurf/src/io.rs:195-235: .vtk, .bgeo, .json), plus the point attributes of VTK / VTU / BGEO files behind `-a <attribute>`
(splashsurf/src/io.rs:66-190).  Host numpy, nothing here is on the timed path.

* `.bgeo`  Houdini's classic binary format, version 5, big endian, optionally gzip-compressed (io/bgeo_format.rs)
* `.json`  `[[x, y, z], ...]` as f64; numbers printed as serde_json / Ryu print them (io/json_format.rs)
* `.ply`   ascii / binary_little_endian / binary_big_endian, float `x y z` of the `vertex` element (io/ply_format.rs:19-71)
* `.vtk`   legacy, ASCII or BINARY, float or double POINTS, UNSTRUCTURED_GRID or POLYDATA (io/vtk_format.rs through vtkio 0.6.3)
* `.vtu`   XML unstructured grid / poly data: ascii, inline base64 and appended (raw / base64) arrays, zlib compression,
           UInt32 / UInt64 headers, either byte order

The writers reproduce the reference's files byte for byte (the gzip stream of a compressed .bgeo is implementation specific: its
payload is identical); tests/test_particle_formats.py runs them beside the reference CLI's `convert`.
"""
from __future__ import annotations

import base64
import gzip
import json
import math
import re
import struct
import zlib

import numpy as np


# ------------------------------------------------------------------------------------------------------------------ BGEO
_BGEO_FLOAT, _BGEO_INT, _BGEO_STRING, _BGEO_INDEXED_STRING, _BGEO_VECTOR = 0, 1, 2, 4, 5


def read_bgeo(path: str):
    """`load_bgeo_file` (bgeo_format.rs:68-100, parser :299-560): returns ``(positions (n, 3) f32, attributes)`` with the named point
    attributes in file order -- Int -> int32 (n,), Float -> float32 (n,) or (n, size), Vector -> float32 (n, size)."""
    raw = open(path, "rb").read()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    if raw[:4] == b"\x7fNSJ":
        raise ValueError("Error while parsing the BGEO file contents: UnsupportedFormatVersion (the new JSON-based bgeo format is not read)")
    if raw[:4] != b"Bgeo":
        raise ValueError("Error while parsing the BGEO file contents: MagicBytesNotFound")
    if len(raw) < 41:
        raise ValueError("Error while parsing the BGEO file contents: unexpected end of file in the header")
    _version_char, version = struct.unpack(">Bi", raw[4:9])
    if version != 5:
        raise ValueError("Error while parsing the BGEO file contents: UnsupportedFormatVersion")
    (num_points, _num_prims, _npg, _nprg, num_point_attrib, _nva, _npa, _na) = struct.unpack(">8i", raw[9:41])
    o = 41
    fields = [("position", ">f4", (3,)), ("unknown", ">f4", ())]
    names = []
    for _ in range(num_point_attrib):
        if len(raw) < o + 2 or len(raw) < o + 8 + struct.unpack(">H", raw[o:o + 2])[0]:
            raise ValueError("Error while parsing the BGEO file contents: unexpected end of file in the attribute definitions")
        (ln,) = struct.unpack(">H", raw[o:o + 2])
        try:
            name = raw[o + 2:o + 2 + ln].decode("utf-8")
        except UnicodeDecodeError:
            raise ValueError("Error while parsing the BGEO file contents: InvalidAttributeName") from None
        o += 2 + ln
        size, typ = struct.unpack(">Hi", raw[o:o + 6])
        o += 6
        if typ not in (_BGEO_FLOAT, _BGEO_INT, _BGEO_STRING, _BGEO_INDEXED_STRING, _BGEO_VECTOR):
            raise ValueError("Error while parsing the BGEO file contents: UnknownAttributeType")
        if typ in (_BGEO_STRING, _BGEO_INDEXED_STRING):
            raise ValueError(f'Error while parsing the BGEO file contents: UnsupportedAttributeType({"String" if typ == _BGEO_STRING else "IndexedString"})')
        o += 4 * size                                                                   # default values
        # the reference reads ONE value per point for Int / Float attributes whatever `size` says and then asserts the count
        # (bgeo_format.rs:520-538, :470-481): sizes other than 1 are a panic there, an error here
        if typ != _BGEO_VECTOR and size != 1:
            raise ValueError(f'failed to read attribute "{name}": {"Int" if typ == _BGEO_INT else "Float"} attributes with {size} components are not supported')
        # a repeated name gets a private field; lookups by name find the first one, like the reference's loop
        names.append((name, f"a{len(names)}"))
        fields.append((names[-1][1], ">i4" if typ == _BGEO_INT else ">f4", (size,) if typ == _BGEO_VECTOR else ()))
    dt = np.dtype(fields)
    if len(raw) < o + num_points * dt.itemsize:
        raise ValueError("Error while parsing the BGEO file contents: unexpected end of file in the point data")
    rec = np.frombuffer(raw, dtype=dt, count=max(num_points, 0), offset=o)
    positions = np.ascontiguousarray(rec["position"], dtype=np.float32).reshape(-1, 3)
    attrs = {}
    for name, field in names:
        if name not in attrs:
            a = rec[field]
            attrs[name] = np.ascontiguousarray(a, dtype=np.int32 if a.dtype.kind == "i" else np.float32)
    return positions, attrs


def _bgeo_payload(particles: np.ndarray) -> bytes:
    p = np.ascontiguousarray(particles, dtype=np.float32).reshape(-1, 3)
    n = len(p)
    if n > 2**31 - 1:
        raise ValueError(f"number of particles ({n}) is too large for bgeo format (max {2**31 - 1})")
    head = b"Bgeo" + struct.pack(">Bi8i", 86, 5, n, 0, 0, 0, 0, 0, 0, 0)
    body = np.empty((n, 4), dtype=">f4")
    body[:, :3] = p
    body[:, 3] = 1.0                                                                    # the weight column (bgeo_format.rs:161)
    return head + body.tobytes() + b"\x00\xff"


def write_bgeo(path: str, particles: np.ndarray, enable_compression: bool = True) -> None:
    """`particles_to_bgeo` (bgeo_format.rs:102-257): version-5 header without attributes, per point x y z and a weight of 1,
    the 0x00 0xff trailer; gzip "fast" when compression is on (the CLI's default, splashsurf/src/io.rs:36-41)."""
    payload = _bgeo_payload(particles)
    if enable_compression:
        # header as flate2 writes it: no name, no mtime, XFL = 4 (fastest), OS = 255 (unknown)
        co = zlib.compressobj(1, zlib.DEFLATED, -15)
        payload = (b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x04\xff" + co.compress(payload) + co.flush()
                   + struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload) & 0xFFFFFFFF))
    with open(path, "wb") as f:
        f.write(payload)


# ------------------------------------------------------------------------------------------------------------------ JSON
def format_f64_json(x: float) -> str:
    """A finite f64 the way serde_json prints it (Ryu's `format64`): shortest round-trip digits; positional with at least one
    fractional digit for decimal exponents -5 < e10 <= 16, otherwise `d.ddde-7` / `1e16` without '+' or padding."""
    if x != x or x in (math.inf, -math.inf):
        return "null"                                                                   # serde_json's rendering of non-finite floats
    if x == 0.0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    r = repr(float(x))
    sign = ""
    if r[0] == "-":
        sign, r = "-", r[1:]
    mant, _, e = r.partition("e")
    exp10 = int(e) if e else 0
    ip, _, fp = mant.partition(".")
    digits = (ip + fp).lstrip("0")
    k = exp10 - len(fp)                                                                 # value = digits * 10^k
    stripped = digits.rstrip("0")
    k += len(digits) - len(stripped)
    digits = stripped
    n = len(digits)
    kk = n + k                                                                          # 10^(kk-1) <= value < 10^kk
    if 0 <= k and kk <= 16:
        return sign + digits + "0" * k + ".0"
    if 0 < kk <= 16:
        return sign + digits[:kk] + "." + digits[kk:]
    if -5 < kk <= 0:
        return sign + "0." + "0" * (-kk) + digits
    if n == 1:
        return f"{sign}{digits}e{kk - 1}"
    return f"{sign}{digits[0]}.{digits[1:]}e{kk - 1}"


def write_json(path: str, particles: np.ndarray) -> None:
    """`particles_to_json` (json_format.rs:55-94): one compact array of `[x,y,z]` triples, f32 widened to f64."""
    p = np.ascontiguousarray(particles, dtype=np.float32).reshape(-1, 3).astype(np.float64)
    fmt = format_f64_json
    with open(path, "w") as f:
        f.write("[")
        f.write(",".join("[" + fmt(x) + "," + fmt(y) + "," + fmt(z) + "]" for x, y, z in p.tolist()))
        f.write("]")


def read_json(path: str) -> np.ndarray:
    """`particles_from_json` (json_format.rs:15-53): an array of three-number arrays, narrowed from f64."""
    try:
        with open(path, "r") as f:
            data = json.load(f)
    except (json.JSONDecodeError, UnicodeDecodeError) as e:
        raise ValueError(f"Reading of file to JSON structure failed. Not a valid JSON file. ({e})") from None
    ok = isinstance(data, list) and all(isinstance(q, list) and len(q) == 3 and all(isinstance(v, (int, float)) and not isinstance(v, bool) for v in q)
                                        for q in data)
    if not ok:
        raise ValueError("Parsing of JSON structure as particle positions failed. Expected JSON file containing particle positions like e.g. "
                         "'[[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]'.")
    with np.errstate(over="ignore"):
        return np.asarray(data, dtype=np.float64).reshape(-1, 3).astype(np.float32)


# ------------------------------------------------------------------------------------------------------------------- PLY
_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
              "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def parse_ply(path: str):
    """A PLY file as ply-rs' default parser sees it: ``(elements, payload)`` with ``elements`` = ordered ``(name, count, properties)``,
    properties = ``(name, scalar type | None, (count type, item type) | None)`` and ``payload[name]`` = dict property -> array (scalars)
    or list of arrays (lists).  ascii, binary_little_endian and binary_big_endian."""
    b = open(path, "rb").read()
    m = re.search(rb"end_header[ \t]*\r?\n", b)
    if m is None or not b.startswith(b"ply"):
        raise ValueError("Failed to parse PLY file: missing 'ply' magic or 'end_header'")
    fmt, elements = None, []
    for ln in b[:m.start()].decode("ascii", "replace").splitlines()[1:]:
        tok = ln.split()
        if not tok or tok[0] in ("comment", "obj_info"):
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            elements.append((tok[1], int(tok[2]), []))
        elif tok[0] == "property" and elements:
            if tok[1] == "list":
                if tok[2] not in _PLY_TYPES or tok[3] not in _PLY_TYPES:
                    raise ValueError(f"Failed to parse PLY file: unknown property type in '{ln}'")
                elements[-1][2].append((tok[4], None, (_PLY_TYPES[tok[2]], _PLY_TYPES[tok[3]])))
            else:
                if tok[1] not in _PLY_TYPES:
                    raise ValueError(f"Failed to parse PLY file: unknown property type in '{ln}'")
                elements[-1][2].append((tok[2], _PLY_TYPES[tok[1]], None))
        else:
            raise ValueError(f"Failed to parse PLY file: unexpected header line '{ln}'")
    if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
        raise ValueError(f"Failed to parse PLY file: unsupported format '{fmt}'")
    payload, o = {}, m.end()
    if fmt == "ascii":
        lines = b[o:].decode("ascii", "replace").splitlines()
        lines = [ln for ln in lines if ln.strip()]
        li = 0
        for name, count, props in elements:
            rows = lines[li:li + count]
            if len(rows) < count:
                raise ValueError(f"Failed to parse PLY file: element '{name}' is truncated")
            li += count
            cols = {p[0]: [] for p in props}
            if all(p[2] is None for p in props):
                tab = np.array([r.split()[:len(props)] for r in rows], dtype=object).reshape(count, len(props))
                for j, (pn, st, _) in enumerate(props):
                    cols[pn] = tab[:, j].astype(np.float64).astype(st) if count else np.zeros(0, st)
            else:
                for r in rows:
                    tok, t = r.split(), 0
                    for pn, st, lt in props:
                        if lt is None:
                            cols[pn].append(np.float64(tok[t]).astype(st))
                            t += 1
                        else:
                            n = int(tok[t])
                            cols[pn].append(np.asarray(tok[t + 1:t + 1 + n], dtype=np.float64).astype(lt[1]))
                            t += 1 + n
                for pn, st, lt in props:
                    if lt is None:
                        cols[pn] = np.asarray(cols[pn], dtype=st)
            payload[name] = cols
        return elements, payload
    bo = "<" if fmt == "binary_little_endian" else ">"
    for name, count, props in elements:
        cols = {}
        if all(p[2] is None for p in props):
            dt = np.dtype([(f"f{j}", bo + p[1]) for j, p in enumerate(props)])          # packed records, fields by position
            if len(b) < o + count * dt.itemsize:
                raise ValueError(f"Failed to parse PLY file: element '{name}' is truncated")
            rec = np.frombuffer(b, dtype=dt, count=count, offset=o)
            o += count * dt.itemsize
            for j, p in enumerate(props):
                cols[p[0]] = rec[rec.dtype.names[j]].astype(p[1])
        elif len(props) == 1 and count:
            # one list per record (faces): fast path when every list has the same length
            (pn, _, (ct, it)) = props[0]
            cs, isz = np.dtype(ct).itemsize, np.dtype(it).itemsize
            n0 = int(np.frombuffer(b, bo + ct, 1, o)[0])
            stride = cs + n0 * isz
            uniform = len(b) >= o + count * stride and bool(np.all(np.frombuffer(b, np.dtype({"names": ["n"], "formats": [bo + ct], "itemsize": stride}), count, o)["n"] == n0))
            if uniform:
                rec = np.frombuffer(b, np.dtype({"names": ["n", "i"], "formats": [bo + ct, (bo + it, (n0,))], "offsets": [0, cs], "itemsize": stride}), count, o)
                cols[pn] = list(rec["i"].astype(it))
                o += count * stride
            else:
                out = []
                for _ in range(count):
                    n = int(np.frombuffer(b, bo + ct, 1, o)[0])
                    out.append(np.frombuffer(b, bo + it, n, o + cs).astype(it))
                    o += cs + n * isz
                cols[pn] = out
        else:
            for p in props:
                cols[p[0]] = []
            for _ in range(count):
                for pn, st, lt in props:
                    if lt is None:
                        cols[pn].append(np.frombuffer(b, bo + st, 1, o)[0])
                        o += np.dtype(st).itemsize
                    else:
                        n = int(np.frombuffer(b, bo + lt[0], 1, o)[0])
                        o += np.dtype(lt[0]).itemsize
                        cols[pn].append(np.frombuffer(b, bo + lt[1], n, o).astype(lt[1]))
                        o += n * np.dtype(lt[1]).itemsize
            for pn, st, lt in props:
                if lt is None:
                    cols[pn] = np.asarray(cols[pn], dtype=st)
        payload[name] = cols
    return elements, payload


def _ply_vec3(elements, payload, names):
    props = {p[0]: p for e in elements if e[0] == "vertex" for p in e[2]}
    for n in names:
        if n not in props:
            raise ValueError(f"PLY vertex element has no '{n}' property")
        if props[n][1] != "f4":
            raise ValueError("Vertex properties have wrong PLY data type (expected float)")
    v = payload["vertex"]
    return np.stack([v[names[0]], v[names[1]], v[names[2]]], axis=1).astype(np.float32) if len(v[names[0]]) else np.zeros((0, 3), np.float32)


def read_ply_particles(path: str) -> np.ndarray:
    """`particles_from_ply` (ply_format.rs:19-71): the float x / y / z of the 'vertex' element."""
    elements, payload = parse_ply(path)
    if "vertex" not in payload:
        raise ValueError("PLY file is missing a 'vertex' element")
    return _ply_vec3(elements, payload, ("x", "y", "z"))


def read_ply_surface_mesh(path: str):
    """`surface_mesh_from_ply` (ply_format.rs:28-36, :73-188): ``(vertices, triangles, point_attributes)`` -- float x / y / z, faces as
    `list <any> uint vertex_indices` with exactly three indices, nx / ny / nz (when all three are present) as "normals"."""
    elements, payload = parse_ply(path)
    if "vertex" not in payload:
        raise ValueError("PLY file is missing a 'vertex' element")
    verts = _ply_vec3(elements, payload, ("x", "y", "z"))
    if "face" not in payload:
        raise ValueError("PLY file is missing a 'face' element")
    fprops = {p[0]: p for e in elements if e[0] == "face" for p in e[2]}
    if "vertex_indices" not in fprops:
        if [e for e in elements if e[0] == "face"][0][1] == 0:
            tris = np.zeros((0, 3), np.uint64)
        else:
            raise ValueError("A face is missing a 'vertex_indices' element")
    else:
        p = fprops["vertex_indices"]
        if p[2] is None or p[2][1] != "u4":
            raise ValueError("Index properties have wrong PLY data type (expected uint)")
        faces = payload["face"]["vertex_indices"]
        for f in faces:
            if len(f) != 3:
                raise ValueError(f"Invalid number of vertex indices per face: {len(f)} (expected 3)")
        tris = np.asarray(faces, dtype=np.uint64).reshape(-1, 3)
    attrs = {}
    vprops = {p[0] for e in elements if e[0] == "vertex" for p in e[2]}
    if {"nx", "ny", "nz"} <= vprops:
        attrs["normals"] = _ply_vec3(elements, payload, ("nx", "ny", "nz"))
    return verts, tris, attrs


# ------------------------------------------------------------------------------------------------------------ legacy VTK
_VTK_LEGACY = {"bit": None, "unsigned_char": "u1", "char": "i1", "unsigned_short": "u2", "short": "i2", "unsigned_int": "u4", "int": "i4",
               "unsigned_long": "u8", "long": "i8", "float": "f4", "double": "f8", "vtktypeint64": "i8", "vtktypeint32": "i4",
               "vtktypeuint64": "u8", "vtktypeuint32": "u4", "vtktypeuint8": "u1", "vtkidtype": "i8"}


class _LegacyCursor:
    """Token / raw-block reader over a legacy VTK file (keywords are case-insensitive, binary blocks big endian)."""

    def __init__(self, b: bytes, o: int, binary: bool):
        self.b, self.o, self.binary = b, o, binary

    def skip_ws(self):
        b, o = self.b, self.o
        while o < len(b) and b[o] in b" \t\r\n":
            o += 1
        self.o = o

    def token(self):
        self.skip_ws()
        b, o = self.b, self.o
        e = o
        while e < len(b) and b[e] not in b" \t\r\n":
            e += 1
        self.o = e
        return b[o:e].decode("latin-1") if e > o else None

    def peek(self):
        o = self.o
        t = self.token()
        self.o = o
        return t

    def rest_of_line(self):
        e = self.b.find(b"\n", self.o)
        e = len(self.b) if e < 0 else e
        s = self.b[self.o:e].decode("latin-1").strip()
        self.o = min(e + 1, len(self.b))
        return s

    def array(self, typ: str, n: int) -> np.ndarray:
        t = _VTK_LEGACY.get(typ.lower())
        if t is None:
            raise ValueError(f"unsupported VTK data type '{typ}'")
        if self.binary:
            # exactly one line break separates the header line from the block
            if self.b[self.o:self.o + 2] == b"\r\n":
                self.o += 2
            elif self.b[self.o:self.o + 1] == b"\n":
                self.o += 1
            a = np.frombuffer(self.b, ">" + t, n, self.o)
            self.o += a.nbytes
            return a.astype(t)
        out = np.empty(n, dtype=np.float64 if t[0] == "f" else np.int64 if t[0] == "i" else np.uint64)
        for i in range(n):
            tok = self.token()
            if tok is None:
                raise ValueError("unexpected end of the VTK file inside a data array")
            out[i] = float(tok) if t[0] == "f" else int(tok)
        return out.astype(t)


def _read_vtk_legacy(b: bytes):
    lines = b.split(b"\n", 3)
    if len(lines) < 4 or not lines[0].lower().startswith(b"# vtk datafile"):
        raise ValueError("not a legacy VTK file (missing '# vtk DataFile Version' line)")
    kind = lines[2].strip().upper()
    if kind not in (b"ASCII", b"BINARY"):
        raise ValueError("legacy VTK file is neither ASCII nor BINARY")
    c = _LegacyCursor(b, len(lines[0]) + len(lines[1]) + len(lines[2]) + 3, kind == b"BINARY")
    if (c.token() or "").upper() != "DATASET":
        raise ValueError("legacy VTK file has no DATASET line")
    dataset = (c.token() or "").upper()
    if dataset not in ("UNSTRUCTURED_GRID", "POLYDATA"):
        raise ValueError("VTK file does not contain supported data set pieces")
    points, point_data, section, n_section, cell_verts = None, {}, None, 0, None
    while True:
        kw = c.token()
        if kw is None:
            break
        K = kw.upper()
        if K == "POINTS":
            n, typ = int(c.token()), c.token()
            points = c.array(typ, 3 * n).reshape(n, 3)
        elif K in ("CELLS", "VERTICES", "LINES", "POLYGONS", "TRIANGLE_STRIPS"):
            _nc, size = int(c.token()), int(c.token())
            nxt = c.peek()
            if nxt is not None and nxt.upper() == "OFFSETS":                            # VTK 5.x layout: OFFSETS <type> / CONNECTIVITY <type>
                c.token(); offs = c.array(c.token(), _nc).astype(np.int64)
                c.token(); conn = c.array(c.token(), size).astype(np.int64)
                cv = _xml_cells_to_legacy(conn, offs[1:])
            else:
                cv = c.array("int", size).astype(np.int64)
            if K == "CELLS":
                cell_verts = cv
        elif K == "CELL_TYPES":
            c.array("int", int(c.token()))
        elif K in ("POINT_DATA", "CELL_DATA"):
            section, n_section = K, int(c.token())
        elif K == "SCALARS":
            name, typ = c.token(), c.token()
            rest = c.rest_of_line()
            comps = int(rest) if rest else 1
            if (c.peek() or "").upper() == "LOOKUP_TABLE":
                c.token(); c.token()
                c.rest_of_line()
            a = c.array(typ, n_section * comps)
            if section == "POINT_DATA":
                point_data.setdefault(name, a.reshape(n_section, comps) if comps > 1 else a)
        elif K in ("VECTORS", "NORMALS", "COLOR_SCALARS", "TEXTURE_COORDINATES", "TENSORS"):
            name = c.token()
            if K == "COLOR_SCALARS":
                comps = int(c.token())
                a = c.array("unsigned_char" if c.binary else "float", n_section * comps)
            elif K == "TEXTURE_COORDINATES":
                comps = int(c.token())
                a = c.array(c.token(), n_section * comps)
            else:
                comps = 9 if K == "TENSORS" else 3
                a = c.array(c.token(), n_section * comps)
            if section == "POINT_DATA":
                point_data.setdefault(name, a.reshape(n_section, comps) if comps > 1 else a)
        elif K == "LOOKUP_TABLE":
            c.token()
            size = int(c.token())
            c.array("unsigned_char" if c.binary else "float", 4 * size)
        elif K == "FIELD":
            c.token()
            for _ in range(int(c.token())):
                name, comps, tuples, typ = c.token(), int(c.token()), int(c.token()), c.token()
                a = c.array(typ, comps * tuples)
                if section == "POINT_DATA":
                    point_data.setdefault(name, a.reshape(tuples, comps) if comps > 1 else a)
        elif K == "METADATA":
            # INFORMATION / COMPONENT_NAMES blocks end at an empty line
            while c.rest_of_line():
                pass
        else:
            raise ValueError(f"unsupported legacy VTK section '{kw}'")
    if points is None:
        raise ValueError("No supported pieces in VTK file")
    return _VtkPiece(points, point_data, dataset == "UNSTRUCTURED_GRID", cell_verts)


class _VtkPiece:
    """First piece of a VTK file: points, point data, and (unstructured grids) the cells in the legacy `n i0 .. i(n-1)` layout."""

    def __init__(self, points, point_data, unstructured, cell_verts):
        self.points, self.point_data, self.unstructured, self._cells = points, point_data, unstructured, cell_verts

    @property
    def cell_verts(self):
        return self._cells() if callable(self._cells) else self._cells

    def __iter__(self):                                  # (points, point_data) = read_vtk(path)
        return iter((self.points, self.point_data))


def _xml_cells_to_legacy(connectivity: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    """`VertexNumbers::into_legacy` (vtkio): connectivity + END offsets -> `n i0 .. i(n-1)` per cell."""
    offsets = np.asarray(offsets, dtype=np.int64)
    starts = np.concatenate([[0], offsets[:-1]]) if len(offsets) else offsets
    counts = offsets - starts
    out = np.empty(len(connectivity) + len(offsets), dtype=np.int64)
    pos = starts + np.arange(len(offsets))
    out[pos] = counts
    mask = np.ones(len(out), dtype=bool)
    mask[pos] = False
    out[mask] = connectivity
    return out


# --------------------------------------------------------------------------------------------------------------- XML VTK
_VTK_XML = {"Int8": "i1", "UInt8": "u1", "Int16": "i2", "UInt16": "u2", "Int32": "i4", "UInt32": "u4", "Int64": "i8", "UInt64": "u8",
            "Float32": "f4", "Float64": "f8"}


def _b64_stream(text: bytes) -> bytes:
    """Decodes back-to-back base64 blocks (VTK encodes the block header and the data separately: padding may occur mid-stream)."""
    t = re.sub(rb"\s+", b"", text)
    out, o = [], 0
    while o < len(t):
        e = t.find(b"=", o)
        if e < 0:
            out.append(base64.b64decode(t[o:len(t) - (len(t) - o) % 4]))
            break
        e = o + ((e - o) // 4 + 1) * 4
        out.append(base64.b64decode(t[o:e]))
        o = e
    return b"".join(out)


def _b64_take(t: bytes, o: int, nbytes: int):
    """Decodes one base64 block of `nbytes` payload bytes starting at character offset `o`; returns (bytes, next offset)."""
    nchar = (nbytes + 2) // 3 * 4
    return base64.b64decode(t[o:o + nchar])[:nbytes], o + nchar


def _vtk_xml_block(get, bo: str, hdr: str, compressed: bool, joint_header_ok: bool) -> bytes:
    """One data block: `get(o, n, joint)` returns n payload bytes at byte offset o of the block's (decoded) stream."""
    hs = np.dtype(hdr).itemsize
    if not compressed:
        n = int(np.frombuffer(get(0, hs, False), bo + hdr)[0])
        return get(hs, n, joint_header_ok)
    nb, _bs, _last = (int(v) for v in np.frombuffer(get(0, 3 * hs, False), bo + hdr))
    if nb == 0:
        return b""
    sizes = [int(v) for v in np.frombuffer(get(0, (3 + nb) * hs, False), bo + hdr)[3:]]
    out, o = [], (3 + nb) * hs
    for s in sizes:
        out.append(zlib.decompress(get(o, s, False)))
        o += s
    return b"".join(out)


def _read_vtk_xml(b: bytes):
    k = b.find(b"<AppendedData")
    app, app_enc = None, None
    head = b
    if k >= 0:
        e = b.index(b">", k)
        m = re.search(rb'encoding\s*=\s*"([^"]*)"', b[k:e])
        app_enc = m.group(1).decode().lower() if m else "base64"
        u = b.index(b"_", e) + 1
        end = b.rfind(b"</AppendedData>")
        app = b[u:end if end >= 0 else len(b)]
        head = b[:k] + b"</VTKFile>"
    import xml.etree.ElementTree as ET
    try:
        root = ET.fromstring(head)
    except ET.ParseError as e:
        raise ValueError(f"Failed to load VTK file: XML error ({e})") from None
    if root.tag != "VTKFile":
        raise ValueError("Failed to load VTK file: no VTKFile element")
    bo = "<" if root.get("byte_order", "LittleEndian") == "LittleEndian" else ">"
    hdr = _VTK_XML.get(root.get("header_type", "UInt32"), "u4")
    comp = root.get("compressor")
    if comp not in (None, "", "vtkZLibDataCompressor"):
        raise ValueError(f"Failed to load VTK file: unsupported compressor {comp}")
    compressed = bool(comp)
    grid = root.find("UnstructuredGrid")
    if grid is None:
        grid = root.find("PolyData")
    if grid is None:
        raise ValueError("VTK file does not contain supported data set pieces")
    piece = grid.find("Piece")
    if piece is None:
        raise ValueError("No supported pieces in VTK file")
    npoints = int(piece.get("NumberOfPoints", "0"))
    app_text = re.sub(rb"\s+", b"", app) if (app is not None and app_enc == "base64") else None

    def decode(da):
        t = _VTK_XML.get(da.get("type", ""))
        if t is None:
            raise ValueError(f"Failed to load VTK file: unsupported DataArray type {da.get('type')}")
        comps = int(da.get("NumberOfComponents", "1") or "1")
        fmt = da.get("format", "ascii").lower()
        if fmt == "ascii":
            tok = (da.text or "").split()
            a = np.asarray(tok, dtype=np.float64 if t[0] == "f" else np.int64 if t[0] == "i" else np.uint64).astype(t)
        else:
            if fmt == "binary":
                stream = _b64_stream((da.text or "").encode("ascii"))
                raw = _vtk_xml_block(lambda o, n, joint: stream[o:o + n], bo, hdr, compressed, True)
            elif fmt == "appended":
                if app is None:
                    raise ValueError("Failed to load VTK file: appended DataArray without an AppendedData section")
                off = int(da.get("offset", "0"))
                if app_enc == "raw":
                    raw = _vtk_xml_block(lambda o, n, joint: app[off + o:off + o + n], bo, hdr, compressed, True)
                else:
                    hs = np.dtype(hdr).itemsize
                    if not compressed:
                        # VTK encodes the length prefix and the data as two base64 blocks; some writers (and vtkio) encode them as one
                        h, o2 = _b64_take(app_text, off, hs)
                        n = int(np.frombuffer(h, bo + hdr)[0])
                        sep, _ = _b64_take(app_text, o2, n)
                        joint = _b64_take(app_text, off, hs + n)[0][hs:]
                        pad = app_text[off:o2].endswith(b"=")
                        raw = sep if (pad or hs % 3 == 0) else joint
                    else:
                        h3, _ = _b64_take(app_text, off, 3 * hs)
                        nb = int(np.frombuffer(h3, bo + hdr)[0])
                        hfull, o2 = _b64_take(app_text, off, (3 + nb) * hs)
                        sizes = [int(v) for v in np.frombuffer(hfull, bo + hdr)[3:]]
                        blob, _ = _b64_take(app_text, o2, sum(sizes))
                        out, o3 = [], 0
                        for s in sizes:
                            out.append(zlib.decompress(blob[o3:o3 + s]))
                            o3 += s
                        raw = b"".join(out)
            else:
                raise ValueError(f"Failed to load VTK file: unsupported DataArray format {fmt}")
            a = np.frombuffer(raw, bo + t, len(raw) // np.dtype(t).itemsize).astype(t)
        return a.reshape(-1, comps) if comps > 1 else a
    pts_el = piece.find("Points")
    da = pts_el.find("DataArray") if pts_el is not None else None
    if da is None:
        points = np.zeros((0, 3), np.float32)
    else:
        points = decode(da).reshape(-1, 3)
    if len(points) != npoints:
        raise ValueError(f"Failed to load VTK file: Points array holds {len(points)} points, the piece declares {npoints}")
    point_data = {}
    pd = piece.find("PointData")
    if pd is not None:
        for da in pd.findall("DataArray"):
            name = da.get("Name", "")
            if name not in point_data:
                point_data[name] = (da, None)
    def cells():
        ce = piece.find("Cells")
        if ce is None:
            return None
        arr = {da.get("Name", "").lower(): da for da in ce.findall("DataArray")}
        if "connectivity" not in arr or "offsets" not in arr:
            return None
        return _xml_cells_to_legacy(decode(arr["connectivity"]).astype(np.int64).reshape(-1), decode(arr["offsets"]).astype(np.int64).reshape(-1))
    # attributes and cells are decoded lazily (an unsupported array that nobody asks for is not an error)
    return _VtkPiece(points, _LazyArrays(point_data, decode), grid.tag == "UnstructuredGrid", cells)


class _LazyArrays(dict):
    def __init__(self, entries, decode):
        super().__init__(entries)
        self._decode = decode

    def __getitem__(self, k):
        da, a = super().__getitem__(k)
        if a is None:
            a = self._decode(da)
            super().__setitem__(k, (da, a))
        return a

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]


def read_vtk(path: str):
    """`VtkFile::load_file` + first piece (vtk_format.rs:34-75, :141-155): ``(points, point_data)`` of a legacy `.vtk` or an XML
    `.vtu` / `.vtp` file; `points` in the file's float type, `point_data` name -> array in the file's type."""
    b = open(path, "rb").read()
    if b.lstrip()[:1] == b"<":
        return _read_vtk_xml(b)
    return _read_vtk_legacy(b)


def read_vtk_surface_mesh(path: str):
    """`surface_mesh_from_vtk` (vtk_format.rs:168-184, :250-316): ``(vertices f32, triangles u64)`` of an unstructured grid of triangle
    cells (an empty first cell is skipped, anything else is an error); attributes are not loaded, like in the reference."""
    piece = read_vtk(path)
    if not piece.unstructured:
        raise ValueError("Unsupported piece type for loading surface mesh")
    if piece.points.dtype.kind != "f":
        raise ValueError("Point coordinate IOBuffer does not contain f32 or f64 values")
    cv = piece.cell_verts
    cv = np.zeros(0, np.int64) if cv is None else cv
    if len(cv) and cv[0] == 0:
        cv = cv[1:]
    if len(cv) % 4 != 0:
        raise ValueError("Length of cell vertex array is invalid. Expected 4 values per cell (3 for each triangle vertex index + 1 for vertex count). "
                         f"There are {len(cv)} values.")
    cells = cv.reshape(-1, 4)
    bad = np.nonzero(cells[:, 0] != 3)[0]
    if len(bad):
        raise ValueError(f"Expected only triangle cells. Invalid number of vertex indices ({int(cells[bad[0], 0])}) of cell {int(bad[0])}")
    with np.errstate(over="ignore"):
        return np.ascontiguousarray(piece.points, dtype=np.float32).reshape(-1, 3), np.ascontiguousarray(cells[:, 1:], dtype=np.uint64)


def read_vtk_particles(path: str) -> np.ndarray:
    """`particles_from_vtk` (vtk_format.rs:141-155, :60-75): f32 or f64 coordinates, narrowed to f32."""
    points, _ = read_vtk(path)
    if points.dtype.kind != "f":
        raise ValueError("Point coordinate IOBuffer does not contain f32 or f64 values")
    with np.errstate(over="ignore"):
        return np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)


def write_vtk_particles(path: str, particles: np.ndarray) -> None:
    """`particles_to_vtk` (vtk_format.rs:157-166; `Particles` -> UnstructuredGridPiece, mesh.rs): legacy BINARY file titled "particles",
    float POINTS, one VERTEX cell (type 1) per particle, empty POINT_DATA / CELL_DATA sections -- vtkio 0.6.3's layout."""
    p = np.ascontiguousarray(particles, dtype=np.float32).reshape(-1, 3)
    n = len(p)
    cells = np.empty((n, 2), dtype=">u4")
    cells[:, 0] = 1
    cells[:, 1] = np.arange(n, dtype=np.uint32)
    with open(path, "wb") as f:
        f.write(b"# vtk DataFile Version 4.2\nparticles\nBINARY\n\nDATASET UNSTRUCTURED_GRID\n")
        f.write(f"POINTS {n} float\n".encode())
        f.write(p.astype(">f4").tobytes())
        f.write(f"\n\nCELLS {n} {2 * n}\n".encode())
        f.write(cells.tobytes())
        f.write(f"\n\nCELL_TYPES {n}\n".encode())
        f.write(np.full(n, 1, dtype=">i4").tobytes())
        f.write(f"\n\nPOINT_DATA {n}\n\nCELL_DATA {n}\n\n".encode())


# ------------------------------------------------------------------------------------------------------------ dispatchers
def _extension(path: str) -> str:
    import os
    ext = os.path.splitext(str(path))[1]
    if not ext:
        raise ValueError("Unable to detect file format of particle input file (file name has to end with supported extension)")
    return ext[1:].lower()


def particles_from_file(path: str) -> np.ndarray:
    """`splashsurf_lib::io::particles_from_file` (splashsurf_lib/src/io.rs:17-43): format by (case-insensitive) extension."""
    ext = _extension(path)
    if ext in ("vtk", "vtu"):
        return read_vtk_particles(path)
    if ext == "xyz":
        raw = np.fromfile(path, dtype=np.float32)
        return np.ascontiguousarray(raw[: (len(raw) // 3) * 3].reshape(-1, 3))
    if ext == "ply":
        return read_ply_particles(path)
    if ext == "bgeo":
        return read_bgeo(path)[0]
    if ext == "json":
        return read_json(path)
    raise ValueError(f'Unsupported file format extension "{ext}" for reading particles')


def write_particle_positions(path: str, particles: np.ndarray, enable_compression: bool = True) -> None:
    """`splashsurf::io::write_particle_positions` (splashsurf/src/io.rs:195-235): .vtk, .bgeo (gzip by default), .json."""
    import os
    ext = os.path.splitext(str(path))[1]
    if not ext:
        raise ValueError("Unable to detect file format of particle output file (file name has to end with supported extension)")
    ext = ext[1:].lower()
    if ext == "vtk":
        write_vtk_particles(path, particles)
    elif ext == "bgeo":
        write_bgeo(path, particles, enable_compression)
    elif ext == "json":
        write_json(path, particles)
    else:
        raise ValueError(f'Unsupported file format extension "{ext}" for writing particles')


def particle_attributes_from_file(path: str, names) -> dict:
    """`read_particle_positions_with_attributes` (splashsurf/src/io.rs:66-190): the named point attributes as the reference converts
    them for interpolation -- VTK / VTU (vtk_format.rs:96-140, :318-346): scalars stored as u32 / f32 / f64 and 3-vectors stored as
    f32 / f64 become float32; BGEO (bgeo_format.rs:45-66, :273-296): Float -> float32, Vector of size 3 -> float32 (n, 3),
    Int -> uint64 (negative values are an error).  Every requested name has to exist; other formats carry no attributes."""
    names = list(names or [])
    if not names:
        return {}
    ext = _extension(path)
    out = {}
    if ext in ("vtk", "vtu"):
        _, data = read_vtk(path)
        missing = [n for n in names if n not in data]
        if missing:
            raise ValueError('Missing attribute(s) "' + '", "'.join(missing) + '" in input file')
        for name in names:
            a = data[name]
            comps = 1 if a.ndim == 1 else a.shape[1]
            if comps == 1:
                a = a.reshape(-1)
                if not (a.dtype.kind == "f" and a.dtype.itemsize in (4, 8)) and a.dtype != np.uint32:
                    raise ValueError(f'Attribute "{name}": Unsupported IOBuffer scalar data type')
            elif comps == 3:
                if not (a.dtype.kind == "f" and a.dtype.itemsize in (4, 8)):
                    raise ValueError(f'Attribute "{name}": Unsupported IOBuffer vector data type')
            else:
                raise ValueError(f'Attribute "{name}": Unsupported number of components ({comps}) in VTK IO buffer')
            with np.errstate(over="ignore"):
                out[name] = np.ascontiguousarray(a, dtype=np.float32)
        return out
    if ext == "bgeo":
        _, data = read_bgeo(path)
        missing = [n for n in names if n not in data]
        if missing:
            raise ValueError('Missing attribute(s) "' + '", "'.join(missing) + '" in input file')
        for name in names:
            a = data[name]
            if a.dtype.kind == "i":
                if a.ndim != 1 or (a < 0).any():
                    raise ValueError(f'Failed to convert attribute "{name}": failed to convert integer attribute')
                out[name] = a.astype(np.uint64)
            elif a.ndim == 1:
                out[name] = a
            elif a.shape[1] == 3:
                out[name] = a
            else:
                raise ValueError(f'Failed to convert attribute "{name}": unsupported vector attribute size: {a.shape[1]}')
        return out
    raise ValueError(f'Unsupported file format extension "{ext}" for reading particles and attributes')
