"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: slab plan, ordering-preserving halo exchange."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import free_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["SS_ROOT"])
from splashsurf_b200 import distributed as ssd, synthetic as syn

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
r, h, c, S = 0.025, 0.1, 0.0125, 64
p_all = syn.splash((40, 14, 14), 5, r, seed=5)
p_all = p_all[np.random.default_rng(0).permutation(len(p_all))]            # arbitrary global order
n = len(p_all)
lo, hi = n * rank // world, n * (rank + 1) // world
x = torch.from_numpy(p_all[lo:hi].copy())
gid = torch.arange(lo, hi, dtype=torch.float32)                            # carried as a 4th component for the check

# the grid every rank derives (oracle arithmetic is what the C ABI mirrors; here only ncells/min matter)
import oracle
g = oracle.reconstruct(p_all, particle_radius=r, smoothing_length=2.0, cube_size=0.5)
gmin, ncells = g["subdomain_grid"]["aabb_min"], g["grid"]["ncells"]
plan, layer = ssd.plan_partition(x, [float(v) for v in gmin], [int(v) for v in ncells], S, c, h, world)
ax = plan.axis
hist = torch.tensor([float(n)])
recv, counts = ssd.exchange_particles(x, layer, plan, world)
# same exchange on the ids to learn which global particles arrived, in which order
ids3 = torch.stack([gid, gid, gid], dim=1)
recv_ids, _ = ssd.exchange_particles(x, layer, plan, world, payload=ids3)
ids = recv_ids[:, 0].to(torch.int64).numpy()
# an empty rank (no particles at all) must go through the same collectives without hanging
empty = x[:0]
plan_e, layer_e = ssd.plan_partition(x if rank == 0 else empty, [float(v) for v in gmin], [int(v) for v in ncells], S, c, h, world)
recv_e, _ = ssd.exchange_particles(x if rank == 0 else empty, layer if rank == 0 else layer_e, plan_e, world)
res = {"rank": rank, "axis": ax, "cuts": plan.cuts, "srad": plan.srad, "ids": ids.tolist(), "n": n,
       "match": bool(np.array_equal(recv.numpy(), p_all[ids])), "hist_total": float(hist.sum()), "empty_case_recv": int(recv_e.shape[0])}
json.dump(res, open(os.path.join(os.environ["SS_OUT"], f"rank{rank}.json"), "w"))
dist.destroy_process_group()
'''


def test_slab_plan_and_exchange_gloo(tmp_path, oracle_mod):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SS_ROOT=ROOT, SS_OUT=str(tmp_path), OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    from splashsurf_b200 import synthetic as syn
    res = [json.load(open(tmp_path / f"rank{k}.json")) for k in range(2)]
    p_all = syn.splash((40, 14, 14), 5, 0.025, seed=5)
    p_all = p_all[np.random.default_rng(0).permutation(len(p_all))]
    g = oracle_mod.reconstruct(p_all, particle_radius=0.025, smoothing_length=2.0, cube_size=0.5)
    sg = g["subdomain_grid"]
    ax = res[0]["axis"]
    assert ax == int(np.argmax(sg["ncells"])) and res[0]["cuts"] == res[1]["cuts"] and res[0]["hist_total"] == len(p_all)
    cuts, srad = res[0]["cuts"], res[0]["srad"]
    assert cuts[0] == 0 and cuts[-1] == int(sg["ncells"][ax]) and cuts[1] > 0
    # owner layer + ghost reach along the axis in the reference's f32 arithmetic (dense_subdomains.rs:1817-1856)
    dx = np.float32(sg["cell_size"]); gmin = np.float32(sg["aabb_min"][ax])
    xa = p_all[:, ax].astype(np.float32)
    own = np.floor((xa - gmin) / dx).astype(np.int64)
    margin = np.float32(np.float32(np.ceil(np.float32(0.1) / np.float32(0.0125)) * np.float32(0.0125)) * np.float32(1.01))
    for k in range(2):
        ids = np.asarray(res[k]["ids"], dtype=np.int64)
        assert res[k]["match"]
        assert (np.diff(ids) > 0).all(), "exchange must keep ascending global particle order"
        lo, hi = cuts[k] - srad, cuts[k + 1] + srad            # kept subdomain layers (owned + density halo)
        # every particle that is a member (owner or ghost) of a kept layer: within the ghost margin of the kept range, in
        # the reference's f32 arithmetic (dense_subdomains.rs:1846-1851: distance to the face < margin)
        lo_face = gmin + np.float32(lo) * dx
        hi_face = gmin + np.float32(hi) * dx
        need = ((lo_face - xa) < margin) & ((xa - hi_face) < margin)
        got = set(ids.tolist())
        assert set(np.nonzero(need)[0].tolist()) <= got
        # and the routing is tight: nothing farther than 1.05 margins from the kept range travels
        far = ((lo_face - xa) > np.float32(1.05) * margin) | ((xa - hi_face) > np.float32(1.05) * margin)
        assert not (set(np.nonzero(far)[0].tolist()) & got)
        minc = gmin + own.astype(np.float32) * dx
        near_lo = (xa - minc) < margin                        # really a ghost of the layer below
        assert near_lo.any()
    # balance: both ranks get a comparable share
    sizes = [len(res[k]["ids"]) for k in range(2)]
    assert max(sizes) < 0.8 * len(p_all) + 2000


def test_balanced_cuts():
    from splashsurf_b200.distributed import balanced_cuts
    assert balanced_cuts(np.array([10, 10, 10, 10]), 2) == [0, 2, 4]
    assert balanced_cuts(np.array([100, 1, 1, 1, 1]), 2)[1] in (1,)
    c = balanced_cuts(np.array([5, 5]), 4)
    assert c[0] == 0 and c[-1] == 2 and all(b >= a for a, b in zip(c, c[1:]))
    assert balanced_cuts(np.zeros(3), 2)[-1] == 3
