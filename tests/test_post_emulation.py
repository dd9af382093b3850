"""CPU tests of the DEVICE post-processing kernels (splashsurf_b200/csrc/ss_post.cuh) without a GPU.

tests/emul/post_emul.cpp compiles the kernel sources with g++ on the CPU executor (tests/emul/cuda_emul.h) and launches
single kernels; the sorts / scans between them are done by numpy.  The kernels are per-thread code (no warp or block collectives), so this executes the
same statements the GPU does and checks their indexing (splat-bin queries, CSR construction, iteration buffers) and
arithmetic against oracle/postprocess.py.  The product never uses this path (it is not a CPU fallback: the harness lives
under tests/ and is not part of the library)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

REL = 2e-5
EMUL_DIR = os.path.join(ROOT, "tests", "emul")


def _close(a, b, tol=REL):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert not np.isnan(a).any() and not np.isnan(b).any()
    err = float(np.abs(a - b).max())
    assert err <= tol * max(float(np.abs(b).max()), 1.0), err


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(EMUL_DIR, "libpost_emul.so")
    deps = [os.path.join(EMUL_DIR, "post_emul.cpp"), os.path.join(EMUL_DIR, "cuda_emul.h")] + \
        [os.path.join(ROOT, "splashsurf_b200", "csrc", f) for f in ("ss_common.cuh", "ss_kernels.cuh", "ss_post.cuh")]
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-w", "-I" + cuda_inc, "-I" + EMUL_DIR, "-shared", "-fPIC",
                               "-pthread", "-Wl,-Bsymbolic", "-o", so, deps[0]])
    L = C.CDLL(so)
    L.emul_sizeof_dev.restype = C.c_uint
    L.emul_dev_nbin_sub.restype = C.c_int
    L.emul_csr.restype = C.c_uint32
    return L


def _p(a):
    return C.c_void_p(a.ctypes.data)


class Scene:
    """A small splash: oracle reconstruction + the splat bins built by the emulated binning kernels."""

    def __init__(self, L, oracle_mod, S=16):
        from splashsurf_b200 import synthetic
        self.L = L
        r, l, cube = 0.025, 2.0, 0.75
        self.x = synthetic.splash((12, 11, 10), 3, r, 91)
        self.o = oracle_mod.reconstruct(self.x, particle_radius=r, smoothing_length=l, cube_size=cube, subdomain_num_cubes_per_dim=S,
                                        subdomain_grid_auto_disable=False)
        assert self.o["rc"] == 0
        _, h, c = oracle_mod.absolute_params(r, l, cube)
        self.h, self.c, self.r, self.S = float(h), float(c), r, S
        self.rho = np.ascontiguousarray(self.o["particle_densities"], dtype=np.float32)
        self.verts = np.ascontiguousarray(self.o["vertices"], dtype=np.float32)
        self.tris = np.ascontiguousarray(self.o["triangles"], dtype=np.uint32)
        g = self.o["grid"]
        gmin = np.asarray(g["aabb_min"], dtype=np.float32)
        nsd = (np.asarray(g["ncells"], dtype=np.int64) + S - 1) // S
        margin = np.float32(np.float32(np.ceil(np.float32(h) / np.float32(c))) * np.float32(c)) * np.float32(1.01)
        self.D = (C.c_uint8 * L.emul_sizeof_dev())()
        L.emul_make_dev(self.D, (C.c_float * 3)(*gmin), C.c_float(c), C.c_float(h), C.c_int(S), (C.c_int * 3)(*[int(v) for v in nsd]),
                        C.c_float(float(margin)))
        # memberships: every particle within the ghost margin of a tile, tiles in ascending flat order, particles ascending
        sub = np.float32(c) * np.float32(S)
        cids, pidx, flats = [], [], []
        for ti in range(nsd[0]):
            for tj in range(nsd[1]):
                for tk in range(nsd[2]):
                    lo = gmin + np.array([ti, tj, tk], np.float32) * sub
                    inside = np.all((self.x > lo - margin) & (self.x < lo + sub + margin), axis=1)
                    ids = np.nonzero(inside)[0]
                    if len(ids):
                        cids.append(np.full(len(ids), len(flats), np.uint32)); pidx.append(ids.astype(np.uint32))
                        flats.append((ti * nsd[1] + tj) * nsd[2] + tk)
        assert len(flats) > 1                                         # several tiles: the tile lookup is exercised
        self.sub_flat = np.asarray(flats, dtype=np.uint32)
        cid = np.concatenate(cids); pid = np.concatenate(pidx)
        self.m = len(cid)
        key = np.empty(self.m, np.uint32)
        L.emul_bin_keys(self.D, _p(self.x), C.c_uint32(self.m), _p(cid), _p(self.sub_flat), _p(pid), _p(key))
        order = np.argsort(key, kind="stable")
        self.key = np.ascontiguousarray(key[order]); self.pidx = np.ascontiguousarray(pid[order])
        nkeys = len(flats) * L.emul_dev_nbin_sub(self.D)
        self.bin_start = np.empty(nkeys, np.uint32); self.bin_end = np.zeros(nkeys, np.uint32)
        L.emul_bin_tables(_p(self.key), C.c_uint32(self.m), _p(self.bin_start), _p(self.bin_end), C.c_uint32(nkeys))
        self.rec = np.zeros((self.m, 4), np.float32); ks = np.zeros(self.m, np.int32)
        L.emul_records(self.D, _p(self.x), _p(self.rho), C.c_uint32(self.m), _p(self.key), _p(self.pidx), _p(self.sub_flat), _p(self.rec), _p(ks))
        self.sphere_mass = float(oracle_mod.sph_rest_mass(r))

    def query(self):
        return (_p(self.sub_flat), C.c_uint32(len(self.sub_flat)), _p(self.bin_start), _p(self.bin_end), _p(self.pidx), _p(self.rec),
                _p(self.rho), C.c_float(self.sphere_mass))

    def adjacency(self):
        keys = np.empty(6 * len(self.tris), np.uint64)
        self.L.emul_edge_keys(_p(self.tris), C.c_uint32(len(self.tris)), _p(keys))
        keys.sort()
        row = np.empty(len(self.verts) + 1, np.uint32); idx = np.empty(len(keys), np.uint32)
        total = self.L.emul_csr(_p(keys), C.c_uint32(len(keys)), C.c_uint32(len(self.verts)), _p(row), _p(idx))
        return row, np.ascontiguousarray(idx[:total])


@pytest.fixture(scope="module")
def scene(emul, oracle_mod):
    return Scene(emul, oracle_mod)


def test_weighted_neighbor_counts(scene):
    from oracle import postprocess as pp
    wnc = np.full(len(scene.x), -1.0, np.float32)
    scene.L.emul_weighted_ncount(scene.D, *scene.query(), _p(scene.key), C.c_uint32(scene.m), _p(wnc))
    assert (wnc >= 0).all()                                           # every particle written exactly by its containing tile
    _close(wnc, pp.weighted_neighbor_counts(scene.x, scene.h))


@pytest.mark.parametrize("dim", [1, 3])
def test_interpolate_quantity(scene, dim):
    from oracle import postprocess as pp
    rng = np.random.default_rng(5)
    vals = rng.normal(size=(len(scene.x), dim)).astype(np.float32) if dim == 3 else (scene.x[:, 1] * 2 + 1).astype(np.float32)
    # points: the vertices, moved by up to one cell as smoothing does (also across tile faces)
    pts = (scene.verts + rng.uniform(-scene.c, scene.c, size=scene.verts.shape).astype(np.float32)).astype(np.float32)
    out = np.empty((len(pts), dim), np.float32) if dim == 3 else np.empty(len(pts), np.float32)
    for corr in (1, 0):
        scene.L.emul_interpolate(scene.D, *scene.query(), _p(pts), C.c_uint32(len(pts)), _p(np.ascontiguousarray(vals)), C.c_int(dim),
                                 C.c_int(corr), _p(out))
        _close(out, pp.interpolate_quantity(scene.x, scene.rho, scene.sphere_mass, scene.h, vals, pts, bool(corr)))


def test_sph_normals_at_moved_vertices(scene, oracle_mod):
    rng = np.random.default_rng(6)
    pts = (scene.verts + rng.uniform(-0.5 * scene.c, 0.5 * scene.c, size=scene.verts.shape).astype(np.float32)).astype(np.float32)
    out = np.empty_like(pts)
    scene.L.emul_sph_normals(scene.D, *scene.query(), _p(pts), C.c_uint32(len(pts)), _p(out))
    ref = oracle_mod.sph_normals(scene.x, scene.rho, pts, compact_support_radius=scene.h, particle_rest_mass=scene.sphere_mass)
    _close(out, ref, 5e-5)


def test_vertex_connectivity_csr(scene):
    from oracle import postprocess as pp
    row, idx = scene.adjacency()
    off, adj = pp.vertex_vertex_connectivity(scene.tris, len(scene.verts))
    assert np.array_equal(row.astype(np.int64), off)
    for i in range(0, len(scene.verts), 7):
        assert np.array_equal(idx[row[i]:row[i + 1]], np.sort(adj[off[i]:off[i + 1]]))
    assert all(np.all(np.diff(idx[row[i]:row[i + 1]].astype(np.int64)) > 0) for i in range(len(scene.verts)))


@pytest.mark.parametrize("iterations", [1, 4, 5])
def test_weighted_laplacian_smoothing(scene, iterations):
    from oracle import postprocess as pp
    row, idx = scene.adjacency()
    off, adj = pp.vertex_vertex_connectivity(scene.tris, len(scene.verts))
    w = np.random.default_rng(7).uniform(0, 1, len(scene.verts)).astype(np.float32)
    v = scene.verts.copy()
    scene.L.emul_laplacian(C.c_uint32(len(v)), _p(v), _p(row), _p(idx), _p(w), C.c_float(0.9), C.c_uint32(iterations))
    _close(v, pp.laplacian_smoothing(scene.verts, off, adj, iterations, 0.9, w))
    v1 = scene.verts.copy()
    scene.L.emul_laplacian(C.c_uint32(len(v1)), _p(v1), _p(row), _p(idx), None, C.c_float(1.0), C.c_uint32(iterations))
    _close(v1, pp.laplacian_smoothing(scene.verts, off, adj, iterations, 1.0, np.ones(len(v1), np.float32)))


def test_normals_area_weighted_and_smoothed(scene):
    from oracle import postprocess as pp
    nt, nv = len(scene.tris), len(scene.verts)
    keys = np.empty(3 * nt, np.uint64)
    scene.L.emul_corner_keys(_p(scene.tris), C.c_uint32(nt), _p(keys))
    keys.sort()
    row = np.empty(nv + 1, np.uint32); inc = np.empty(len(keys), np.uint32)
    assert scene.L.emul_csr(_p(keys), C.c_uint32(len(keys)), C.c_uint32(nv), _p(row), _p(inc)) == 3 * nt
    n = np.empty_like(scene.verts)
    scene.L.emul_area_normals(C.c_uint32(nv), _p(scene.verts), _p(scene.tris), _p(row), _p(inc), _p(n))
    _close(n, pp.vertex_normals(scene.verts, scene.tris))
    arow, aidx = scene.adjacency()
    off, adj = pp.vertex_vertex_connectivity(scene.tris, nv)
    for it in (1, 2, 3):
        s = n.copy()
        scene.L.emul_smooth_normals(C.c_uint32(nv), _p(s), _p(arow), _p(aidx), C.c_uint32(it))
        _close(s, pp.laplacian_smoothing_normals(n, off, adj, it))


def test_smoothstep(scene):
    from oracle import postprocess as pp
    wnn = np.linspace(-2, 30, 257).astype(np.float32)
    out = np.empty_like(wnn)
    scene.L.emul_smoothstep(C.c_uint32(len(wnn)), _p(wnn), C.c_float(13.0), _p(out))
    _close(out, pp.smoothing_weights(wnn, 13.0), 1e-6)
