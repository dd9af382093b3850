"""Multi-GPU reconstruct: one process per GPU, `torch.distributed` (NCCL over NVLink on GPUs, gloo in the CPU tests)
for the plumbing, the C ABI for every kernel.

The reference parallelises over subdomains (dense_subdomains.rs:521-526, :1581-1598); subdomains are independent once
each one has its ghost particles.  Across ranks the subdomain grid is cut into slabs along its longest axis, balanced
by particle count.  One exchange step (variable all-to-all) gives every rank the particles that are members of its
slab's subdomains plus one halo layer of subdomains that is processed for particle densities only -- so ghost
particles get bit-identical densities without a second exchange.  Global particle order (rank r holds the global
indices [offset_r, offset_{r+1})) is preserved by the exchange, because the reference's per-grid-point summation
order is ascending particle index.  There is no data-path collective besides that halo exchange, two tiny
reductions (bounding box, slab histogram + max subdomain population) and the optional mesh gather to rank 0.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist


# fixed cost of one occupied subdomain tile in the slab balance, in "particle equivalents" (measured: a 64^3-cell tile of
# bulk fluid holds 4096 particles; a nearly empty tile still costs about as much level-set + marching-cubes time)
TILE_COST_PARTICLES = 4096.0

# ---------------------------------------------------------------------------- partition plan (pure host logic) ----
@dataclass
class SlabPlan:
    axis: int                  # partition axis of the subdomain grid
    nsub_axis: int             # subdomains along that axis
    cuts: List[int]            # world+1 cut positions: rank r owns subdomain indices [cuts[r], cuts[r+1])
    srad: int                  # ghost reach in subdomains (ceil(margin / subdomain size))
    halo: int                  # density-only subdomain layers kept around a slab (= srad)
    margin: float = 0.0        # ghost particle margin (dense_subdomains.rs:120-121)
    sub_size: float = 0.0      # edge length of a subdomain
    gmin_axis: Optional[float] = None   # subdomain grid aabb.min along `axis` (set by plan_partition)

    def own(self, r: int):
        return self.cuts[r], self.cuts[r + 1]

    def recv_range(self, r: int):
        """Owner-subdomain index range (along the axis) of the particles rank r needs: members of the subdomains in
        [own_lo - halo, own_hi + halo) are owned by subdomains at most srad further out (+1 for rounding slack)."""
        lo, hi = self.own(r)
        if hi <= lo:
            return 0, 0
        reach = self.halo + self.srad + 1
        return lo - reach, hi + reach

    def recv_interval(self, r: int):
        """Coordinate interval (along the axis) of the particles rank r needs: the members -- owned or ghost -- of the
        subdomain layers [own_lo - halo, own_hi + halo) lie within the ghost margin of that range of layers.  A little slack
        (0.2 % of the margin plus a few float32 ulps of the coordinate) makes it a superset of what the library's f32
        classifier keeps; the library filters exactly.  Far tighter than `recv_range`'s whole layers."""
        lo, hi = self.own(r)
        if hi <= lo or self.gmin_axis is None:
            return 0.0, 0.0
        a = self.gmin_axis + (lo - self.halo) * self.sub_size
        b = self.gmin_axis + (hi + self.halo) * self.sub_size
        slack = self.margin * 1.002 + 1e-5 * (abs(a) + abs(b) + self.sub_size)
        return a - slack, b + slack


def balanced_cuts(hist: np.ndarray, world: int) -> List[int]:
    """Cut positions so that every rank owns a contiguous run of subdomain layers with ~equal particle counts."""
    n = len(hist)
    csum = np.concatenate([[0], np.cumsum(hist.astype(np.float64))])
    total = csum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(csum, target, side="left"))
        if k > 0 and abs(csum[k - 1] - target) <= abs(csum[min(k, n)] - target):
            k -= 1
        k = min(max(k, cuts[-1]), n)
        cuts.append(k)
    cuts.append(n)
    return cuts


def make_plan(grid_ncells, subdomain_cubes: int, cube_size: float, compact_support: float, hist_axis=None, world: int = 1,
              axis: Optional[int] = None) -> SlabPlan:
    S = int(subdomain_cubes)
    nsd = [(int(n) + S - 1) // S for n in grid_ncells]
    ax = int(np.argmax(nsd)) if axis is None else int(axis)
    # ghost margin = ceil(h / c) * c * 1.01 (dense_subdomains.rs:120-121); reach in subdomains (:1827)
    c32, h32 = np.float32(cube_size), np.float32(compact_support)
    margin = np.float32(np.float32(np.ceil(h32 / c32) * c32) * np.float32(1.01))
    srad = int(np.ceil(margin / np.float32(c32 * np.float32(S))))
    hist = np.zeros(nsd[ax]) if hist_axis is None else np.asarray(hist_axis)
    return SlabPlan(ax, nsd[ax], balanced_cuts(hist, world), srad, srad, float(margin), float(np.float32(c32 * np.float32(S))))


def plan_partition(x: torch.Tensor, grid_min, grid_ncells, S: int, cube_size: float, compact_support: float, world: int, group=None):
    """Collective: every rank passes its local particles (any device) and gets the same SlabPlan plus its particles' owner
    layers.  Work model per layer of subdomains: particles + a fixed cost per occupied subdomain tile (every tile pays for
    its bricks and 65^3 grid points even when it only holds a thin sheet of fluid)."""
    plan0 = make_plan(grid_ncells, S, cube_size, compact_support, None, world)
    ax = plan0.axis
    sub_size = float(np.float32(np.float32(cube_size) * np.float32(S)))
    nsd = [(int(nc) + S - 1) // S for nc in grid_ncells]
    o = [owner_layer(x[:, d], float(grid_min[d]), sub_size).clamp(0, nsd[d] - 1) for d in range(3)]
    layer = owner_layer(x[:, ax], float(grid_min[ax]), sub_size)
    hist = torch.bincount(o[ax], minlength=nsd[ax]).to(torch.float64)
    dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    occ = torch.zeros(nsd[0] * nsd[1] * nsd[2], dtype=torch.int32, device=x.device)
    occ[(o[0] * nsd[1] + o[1]) * nsd[2] + o[2]] = 1
    dist.all_reduce(occ, op=dist.ReduceOp.MAX, group=group)
    other = tuple(d for d in range(3) if d != ax)
    tiles = occ.view(nsd[0], nsd[1], nsd[2]).sum(dim=other).to(torch.float64)
    work = hist + TILE_COST_PARTICLES * tiles
    plan = make_plan(grid_ncells, S, cube_size, compact_support, work.cpu().numpy(), world, axis=ax)
    plan.gmin_axis = float(grid_min[ax])
    return plan, layer


def owner_layer(x_axis: torch.Tensor, gmin: float, sub_size: float) -> torch.Tensor:
    """Subdomain layer index of each particle along the partition axis (float64 on purpose: only used to route
    particles conservatively, +-1 slack is added by SlabPlan.recv_range)."""
    return torch.floor((x_axis.to(torch.float64) - float(gmin)) / float(sub_size)).to(torch.int64)


def exchange_particles(x_local: torch.Tensor, layer: torch.Tensor, plan: SlabPlan, world: int, group=None, payload: Optional[torch.Tensor] = None):
    """Variable all-to-all: returns (particles for this rank in ascending global order, counts received per source).
    A particle goes to every rank whose kept layers it can be a member of: by coordinate interval (`recv_interval`) when the
    plan knows the grid origin, else by whole owner layers (`recv_range`).  `payload` ((n, 3) rows, default: the particles
    themselves) is what travels; the routing always comes from `x_local`."""
    if plan.gmin_axis is not None and x_local.shape[1] >= 3:
        coord = x_local[:, plan.axis].to(torch.float64)
        los = torch.tensor([plan.recv_interval(r)[0] for r in range(world)], dtype=torch.float64, device=x_local.device)
        his = torch.tensor([plan.recv_interval(r)[1] for r in range(world)], dtype=torch.float64, device=x_local.device)
        masks = (coord[None, :] >= los[:, None]) & (coord[None, :] < his[:, None])      # (world, n)
    else:
        los = torch.tensor([plan.recv_range(r)[0] for r in range(world)], dtype=torch.int64, device=x_local.device)
        his = torch.tensor([plan.recv_range(r)[1] for r in range(world)], dtype=torch.int64, device=x_local.device)
        masks = (layer[None, :] >= los[:, None]) & (layer[None, :] < his[:, None])      # (world, n)
    idx = torch.nonzero(masks)                                                           # row-major: by rank, then ascending index
    data = x_local if payload is None else payload
    send = data[idx[:, 1]].contiguous()
    counts = [int(v) for v in masks.sum(dim=1).tolist()]
    cnt_in = torch.tensor(counts, dtype=torch.int64, device=x_local.device)
    cnt_out = torch.empty(world, dtype=torch.int64, device=x_local.device)
    dist.all_to_all_single(cnt_out, cnt_in, group=group)
    out_counts = [int(v) for v in cnt_out.tolist()]
    recv = torch.empty((sum(out_counts), 3), dtype=data.dtype, device=x_local.device)
    dist.all_to_all_single(recv.view(-1), send.view(-1), output_split_sizes=[3 * v for v in out_counts],
                           input_split_sizes=[3 * v for v in counts], group=group)
    return recv, out_counts


# ---------------------------------------------------------------------------- device helpers ----
class _CudaView:
    """Zero-copy torch view of device memory owned by the C library (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _view(ptr, shape, typestr, device):
    if not ptr or int(np.prod(shape)) == 0:
        dt = {"<f4": torch.float32, "<u4": torch.int32, "<i4": torch.int32, "<i8": torch.int64, "<u8": torch.int64}[typestr]
        return torch.empty(tuple(shape), dtype=dt, device=device)
    typestr = typestr.replace("<u4", "<i4").replace("<u8", "<i8")
    if device.type != "cuda":        # host memory: only when the runner is driven by the tests' CPU executor of the CUDA sources
        dt = np.dtype(typestr)
        buf = (C.c_char * (int(np.prod(shape)) * dt.itemsize)).from_address(int(ptr))
        return torch.from_numpy(np.frombuffer(buf, dtype=dt).reshape(tuple(shape)))
    return torch.as_tensor(_CudaView(ptr, shape, typestr), device=device)


class _HostEvent:
    """Stands in for torch.cuda.Event when the runner works on host memory (tests only)."""

    def record(self):
        import time
        self.t = time.perf_counter()

    def elapsed_time(self, other) -> float:
        return (other.t - self.t) * 1e3


def _event(device):
    return torch.cuda.Event(enable_timing=True) if device.type == "cuda" else _HostEvent()


def _sync(device):
    if device.type == "cuda":
        torch.cuda.synchronize()


def _pinned(t: torch.Tensor, device) -> torch.Tensor:
    return t.pin_memory() if device.type == "cuda" else t


class Runner:
    """Drives one reconstruct step on this rank's GPU (world == 1: plain C-ABI call)."""

    def __init__(self, ctx, params, world: int, rank: int, local_rank: int, group=None, device=None, protocol: str = "two_call"):
        self.ctx, self.params, self.world, self.rank, self.local_rank, self.group = ctx, params, world, rank, local_rank, group
        # how the global maximum subdomain population (sparse rule) reaches the library: "two_call" = decomposition pre-pass,
        # all-reduce, full call (verified on 2/4/8 GPUs); "callback" = one call, the library calls back for the all-reduce
        # after its decomposition (ss_reconstruct_partition_cb_f32; verified over gloo on the CPU executor only)
        if protocol not in ("two_call", "callback"):
            raise ValueError("protocol must be 'two_call' or 'callback'")
        self.protocol = protocol
        self._out_v = self._out_t = None
        # `device` is only overridden by the tests that drive the runner over gloo with the CPU executor of the CUDA sources
        self.device = torch.device("cuda", local_rank) if device is None else torch.device(device)
        self.last_plan: Optional[SlabPlan] = None

    # -- input sharding used by the bench: rank r holds a contiguous range of global particle indices
    def take_local(self, particles: np.ndarray) -> np.ndarray:
        if self.world == 1:
            return particles
        n = len(particles)
        lo, hi = (n * self.rank) // self.world, (n * (self.rank + 1)) // self.world
        return np.ascontiguousarray(particles[lo:hi])

    # -- single GPU
    def _step_single(self, xyz_ptr: int, n: int, copy_out: bool) -> dict:
        L = self.ctx._L
        s = self.ctx.reconstruct_raw(xyz_ptr, n, self.params)
        try:
            return self._collect(s, copy_out, n_local=n)
        finally:
            self.ctx.free_surface(s)

    def _collect(self, s, copy_out: bool, n_local: int, extra_ms: float = 0.0) -> dict:
        L = self.ctx._L
        tm = self.ctx.timings(s)
        nv, nt = L.ss_surface_num_vertices(s), L.ss_surface_num_triangles(s)
        out = {"timings": tm, "device_ms": tm["upload"] + tm["total_device"] + extra_ms, "launches": int(tm["kernel_launches"]), "nv": nv,
               "nt": nt, "nsub": L.ss_surface_num_subdomains(s), "d2h_bytes": 0}
        cnt = np.zeros(out["nsub"], np.uint64)
        owned = np.ones(out["nsub"], np.uint8)
        if out["nsub"]:
            L.ss_surface_copy_subdomains(s, None, cnt.ctypes.data, None)
            L.ss_surface_copy_subdomain_owned(s, owned.ctypes.data)
        out["memberships"] = float(cnt[owned.astype(bool)].sum())
        out["nsub_owned"] = int(owned.sum())
        if copy_out and self.world == 1:
            if self._out_v is None or self._out_v.numel() < nv * 3:
                self._out_v = _pinned(torch.empty(max(nv * 3, 1), dtype=torch.float32), self.device)
            if self._out_t is None or self._out_t.numel() < nt * 3:
                self._out_t = _pinned(torch.empty(max(nt * 3, 1), dtype=torch.int32), self.device)
            rc = L.ss_surface_copy_vertices(s, C.c_void_p(self._out_v.data_ptr()))
            rc |= L.ss_surface_copy_triangles_u32(s, C.c_void_p(self._out_t.data_ptr()))
            if rc:
                raise RuntimeError("mesh copy-out failed")
            out["d2h_bytes"] = nv * 12 + nt * 12
        return out

    # -- public step: `x` is this rank's (n, 3) float32 tensor (cuda, or pinned host for the end-to-end path)
    def step(self, x, n: Optional[int] = None, copy_out: bool = False) -> dict:
        if self.world == 1:
            if isinstance(x, int):
                return self._step_single(x, int(n), copy_out)
            return self._step_single(x.data_ptr(), x.shape[0], copy_out)
        return self._step_multi(x, copy_out)

    def _step_multi(self, x: torch.Tensor, copy_out: bool) -> dict:
        L, p, world, rank = self.ctx._L, self.params, self.world, self.rank
        t_ev = [_event(self.device) for _ in range(3)]
        t_ev[0].record()
        xd = x.to(self.device, non_blocking=True) if x.device.type != self.device.type else x
        # 1. global bounding box -> the grid of ALL particles (lib.rs:476-516), identical on every rank
        if xd.shape[0]:
            mn, mx = xd.min(dim=0).values, xd.max(dim=0).values
        else:
            mn = torch.full((3,), float("inf"), device=self.device); mx = -mn
        dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=self.group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
        corners = torch.stack([mn, mx]).cpu().numpy().astype(np.float32)
        from . import _Grid
        grid = _Grid()
        rc = L.ss_grid_for_reconstruction_f32(self.ctx._h, C.c_void_p(corners.ctypes.data), C.c_uint64(2), C.byref(p), C.byref(grid))
        if rc:
            raise RuntimeError((L.ss_last_error() or b"").decode())
        # 2. slab plan balanced by a work model (particles + occupied tiles per layer)
        plan, layer = plan_partition(xd, [float(v) for v in grid.aabb_min], [int(v) for v in grid.cells_per_dim], int(p.subdomain_num_cubes_per_dim),
                                     float(p.cube_size), float(p.compact_support_radius), world, self.group)
        ax = plan.axis
        self.last_plan = plan
        # 3. halo exchange (variable all-to-all over NCCL); keeps ascending global particle order
        recv, counts = exchange_particles(xd, layer, plan, world, self.group)
        t_ev[1].record()
        own_lo, own_hi = plan.own(rank)
        # 4. local maximum subdomain population -> global maximum (sparse rule, dense_subdomains.rs:1242-1251).  Every rank
        #    makes the same two library calls and the same all-reduce, also ranks that received no particles.
        _sync(self.device)
        s = C.c_void_p()
        pre_launches = 0
        if self.protocol == "callback":
            failure = []

            def _reduce(local_max, _user):
                try:
                    t = torch.tensor([int(local_max)], dtype=torch.int64, device=self.device)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
                    return int(t.item())
                except BaseException as exc:                      # ctypes would swallow it: remember and re-raise after the call
                    failure.append(exc)
                    return int(local_max)
            cb = C.CFUNCTYPE(C.c_uint64, C.c_uint64, C.c_void_p)(_reduce)
            rc = L.ss_reconstruct_partition_cb_f32(self.ctx._h, C.c_void_p(recv.data_ptr()), C.c_uint64(recv.shape[0]), C.byref(p),
                                                   C.byref(grid), ax, own_lo, own_hi, plan.halo, cb, None, C.byref(s))
            if failure:
                raise failure[0]
            if rc:
                raise RuntimeError((L.ss_last_error() or b"").decode())
        else:
            rc = L.ss_reconstruct_partition_f32(self.ctx._h, C.c_void_p(recv.data_ptr()), C.c_uint64(recv.shape[0]), C.byref(p),
                                                C.byref(grid), ax, own_lo, own_hi, plan.halo, C.c_uint64(0), 1, C.byref(s))
            if rc:
                raise RuntimeError((L.ss_last_error() or b"").decode())
            gmax = torch.tensor([L.ss_surface_max_subdomain_particles(s)], dtype=torch.int64, device=self.device)
            pre_launches = int(self.ctx.timings(s)["kernel_launches"])
            self.ctx.free_surface(s)
            dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=self.group)
            # 5. this rank's slab
            s = C.c_void_p()
            rc = L.ss_reconstruct_partition_f32(self.ctx._h, C.c_void_p(recv.data_ptr()), C.c_uint64(recv.shape[0]), C.byref(p),
                                                C.byref(grid), ax, own_lo, own_hi, plan.halo, C.c_uint64(int(gmax.item())), 0, C.byref(s))
            if rc:
                raise RuntimeError((L.ss_last_error() or b"").decode())
        try:
            t_ev[2].record()
            _sync(self.device)
            out = self._collect(s, False, n_local=x.shape[0])
            # events on torch's stream bracket the whole step: exchange (NCCL) + both host-synchronous library calls
            out["device_ms"] = t_ev[0].elapsed_time(t_ev[2])
            out["launches"] += pre_launches
            out["recv_particles"] = int(recv.shape[0])
            out["exchange_ms"] = t_ev[0].elapsed_time(t_ev[1])
            out["plan"] = plan
            if copy_out:
                out.update(self._gather_mesh(s, plan))
            return out
        finally:
            self.ctx.free_surface(s)

    # -- mesh assembly on rank 0: concatenate per-rank meshes, weld the vertices on inter-slab faces, copy to host
    def _gather_mesh(self, s, plan: SlabPlan) -> dict:
        L, world, rank = self.ctx._L, self.world, self.rank
        nv, nt = L.ss_surface_num_vertices(s), L.ss_surface_num_triangles(s)
        v = _view(L.ss_surface_device_vertices(s), (nv, 3), "<f4", self.device)
        t = _view(L.ss_surface_device_triangles(s), (nt, 3), "<u4", self.device)
        k = _view(L.ss_surface_device_vertex_keys(s), (nv,), "<u8", self.device)
        sizes = torch.tensor([nv, nt], dtype=torch.int64, device=self.device)
        all_sizes = [torch.empty_like(sizes) for _ in range(world)]
        dist.all_gather(all_sizes, sizes, group=self.group)
        all_sizes = [tuple(int(a) for a in z.tolist()) for z in all_sizes]
        if rank != 0:
            if nv:
                dist.send(v.contiguous(), 0, group=self.group); dist.send(k.contiguous(), 0, group=self.group)
            if nt:
                dist.send(t.contiguous(), 0, group=self.group)
            return {"d2h_bytes": 0, "nv_global": None, "nt_global": None}
        vs, ks, ts, off = [v.clone()], [k.clone()], [t.clone()], nv
        for r in range(1, world):
            rv, rt = all_sizes[r]
            bv = torch.empty((rv, 3), dtype=torch.float32, device=self.device)
            bk = torch.empty((rv,), dtype=torch.int64, device=self.device)
            bt = torch.empty((rt, 3), dtype=torch.int32, device=self.device)
            if rv:
                dist.recv(bv, r, group=self.group); dist.recv(bk, r, group=self.group)
            if rt:
                dist.recv(bt, r, group=self.group)
            vs.append(bv); ks.append(bk); ts.append(bt + off)
            off += rv
        V, K, T = torch.cat(vs).contiguous(), torch.cat(ks).contiguous(), torch.cat(ts).contiguous()
        # candidates: vertices on an inter-slab face = edge not along the partition axis whose point index along the
        # axis is a multiple of S at a cut (key layout: i << 42 | j << 22 | k << 2 | axis)
        S = int(self.params.subdomain_num_cubes_per_dim)
        shift = (42, 22, 2)[plan.axis]
        coord = (K >> shift) & 0xFFFFF
        eaxis = K & 3
        cut_pts = torch.tensor([c * S for c in plan.cuts[1:-1]], dtype=torch.int64, device=self.device)
        cand = torch.nonzero((eaxis != plan.axis) & torch.isin(coord, cut_pts)).view(-1).to(torch.int32).contiguous()
        nv_out = C.c_uint64(V.shape[0])
        _sync(self.device)
        rc = L.ss_weld_meshes(self.ctx._h, C.c_void_p(V.data_ptr()), C.c_void_p(K.data_ptr()), C.c_uint64(V.shape[0]), C.c_void_p(T.data_ptr()),
                              C.c_uint64(T.shape[0]), C.c_void_p(cand.data_ptr()), C.c_uint64(cand.shape[0]), C.byref(nv_out))
        if rc:
            raise RuntimeError((L.ss_last_error() or b"").decode())
        nvg, ntg = int(nv_out.value), int(T.shape[0])
        if self._out_v is None or self._out_v.numel() < nvg * 3:
            self._out_v = _pinned(torch.empty(max(nvg * 3, 1), dtype=torch.float32), self.device)
        if self._out_t is None or self._out_t.numel() < ntg * 3:
            self._out_t = _pinned(torch.empty(max(ntg * 3, 1), dtype=torch.int32), self.device)
        self._out_v[:nvg * 3].copy_(V[:nvg].view(-1), non_blocking=True)
        self._out_t[:ntg * 3].copy_(T.view(-1), non_blocking=True)
        _sync(self.device)
        return {"d2h_bytes": nvg * 12 + ntg * 12, "nv_global": nvg, "nt_global": ntg, "keys_global": K[:nvg]}

    def gathered_mesh(self, nv: int, nt: int):
        """Host copies of the last gathered mesh (rank 0)."""
        return self._out_v[:nv * 3].view(-1, 3).numpy().copy(), self._out_t[:nt * 3].view(-1, 3).numpy().copy()
