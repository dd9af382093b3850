"""Per-rank driver of the reconstruct step: one process per GPU (torch.distributed / NCCL for the plumbing).

world == 1: a plain call into the C ABI.  world > 1: see Runner.step (slab partition of the subdomain grid with a
ghost-particle halo exchange) -- docs in DESIGN.md (row e)."""
from __future__ import annotations

import ctypes as C

import numpy as np


class Runner:
    def __init__(self, ctx, params, world: int, rank: int, local_rank: int):
        self.ctx, self.params, self.world, self.rank, self.local_rank = ctx, params, world, rank, local_rank
        self._out_v = self._out_t = None
        if world > 1:
            raise NotImplementedError("multi-GPU reconstruct is not built yet (DESIGN.md row e)")

    def take_local(self, particles: np.ndarray) -> np.ndarray:
        return particles

    def step(self, xyz_ptr: int, n: int, copy_out: bool) -> dict:
        import torch
        L = self.ctx._L
        s = self.ctx.reconstruct_raw(xyz_ptr, n, self.params)
        try:
            tm = self.ctx.timings(s)
            nv, nt = L.ss_surface_num_vertices(s), L.ss_surface_num_triangles(s)
            out = {"timings": tm, "device_ms": tm["upload"] + tm["total_device"], "launches": int(tm["kernel_launches"]), "nv": nv, "nt": nt,
                   "nsub": L.ss_surface_num_subdomains(s), "d2h_bytes": 0}
            flat_cnt = np.empty(out["nsub"], np.uint64)
            L.ss_surface_copy_subdomains(s, None, flat_cnt.ctypes.data, None)
            out["memberships"] = float(flat_cnt.sum())
            if copy_out:
                if self._out_v is None or self._out_v.numel() < nv * 3:
                    self._out_v = torch.empty(max(nv * 3, 1), dtype=torch.float32).pin_memory()
                if self._out_t is None or self._out_t.numel() < nt * 3:
                    self._out_t = torch.empty(max(nt * 3, 1), dtype=torch.int32).pin_memory()
                rc = L.ss_surface_copy_vertices(s, C.c_void_p(self._out_v.data_ptr()))
                rc |= L.ss_surface_copy_triangles_u32(s, C.c_void_p(self._out_t.data_ptr()))
                if rc:
                    raise RuntimeError("mesh copy-out failed")
                out["d2h_bytes"] = nv * 12 + nt * 12
            return out
        finally:
            self.ctx.free_surface(s)
