#!/usr/bin/env python3
"""Multi-GPU parity check (launch with torchrun, one rank per GPU): the gathered + welded mesh of the partitioned run
must equal the oracle's mesh bit for bit.  Exit code 0 on success."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import splashsurf_b200 as ss
from splashsurf_b200 import distributed as ssd, synthetic as syn


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    import datetime
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=int(os.environ.get("SS_MGPU_TIMEOUT", "120"))))
    ok = True
    cases = [
        ("splash", syn.splash((60, 20, 20), 8, 0.025, 31), dict(particle_radius=0.025, smoothing_length=2.0, cube_size=0.5)),
        ("dam", syn.dam_break_scaled(400_000, 0.01, 32), dict(particle_radius=0.01, smoothing_length=2.0, cube_size=0.5)),
        ("dam_S32", syn.dam_break_scaled(150_000, 0.01, 33), dict(particle_radius=0.01, smoothing_length=2.0, cube_size=0.75, subdomain_num_cubes_per_dim=32)),
    ]
    for name, p_all, kw in cases:
        ctx = ss.Context(local)
        params = ss.make_params(**kw)
        runner = ssd.Runner(ctx, params, world, rank, local)
        runner.want_keys = True
        x = torch.from_numpy(runner.take_local(p_all)).cuda()
        res = runner.step(x, copy_out=True)
        if rank == 0:
            import oracle
            o = oracle.reconstruct(p_all, **kw)
            v, t = runner.gathered_mesh(res["nv_global"], res["nt_global"])
            K = res["keys_global"].cpu().numpy().astype(np.uint64)
            keys = np.stack([(K >> 42) & 0xFFFFF, (K >> 22) & 0xFFFFF, (K >> 2) & 0xFFFFF, K & 3], axis=1).astype(np.int64)
            m = oracle.mesh_parity(v, t.astype(np.int64), keys, o["vertices"], o["triangles"], o["vertex_keys"], kw.get("subdomain_num_cubes_per_dim", 64))
            good = m["keys_equal"] and m["triangles_equal"] and m["n_not_bitexact"] == 0
            print(f"[{name}] world={world} cuts={res['plan'].cuts} axis={res['plan'].axis} recv={res['recv_particles']} of {len(p_all)} -> {m}", flush=True)
            ok &= bool(good)
        runner.close()
        ctx.close()
        dist.barrier()
        # the user-facing call (DistributedReconstructor): same mesh as the single-device call of rank 0, with a particle AABB and SPH normals
        box = dict(aabb_min=(p_all.min(axis=0) - 0.05).tolist(), aabb_max=(np.quantile(p_all, 0.8, axis=0)).tolist())
        rec = ssd.DistributedReconstructor(sph_normals=True, **kw, **box)
        n = len(p_all)
        mwd = rec(p_all[(n * rank) // world:(n * (rank + 1)) // world])
        if rank == 0:
            one = ss.reconstruct_surface(p_all, sph_normals=True, context=rec.ctx, **kw, **box)

            def canon(v, t, nrm):
                o = np.lexsort(v.T[::-1]); r = np.empty(len(v), np.int64)
                vs = v[o].view(np.uint32).reshape(-1, 3)
                r[o] = np.cumsum(np.r_[True, (vs[1:] != vs[:-1]).any(axis=1)]) - 1        # coincident vertices share a number
                t = r[np.asarray(t).astype(np.int64)]
                t = np.stack([np.roll(row, -s) for row, s in zip(t, np.argmin(t, axis=1))]) if len(t) else t
                return v[o], t[np.lexsort(t.T[::-1])] if len(t) else t, nrm[o]
            has_n = "normals" in mwd.point_attributes          # (the CPU executor's host path gathers no normals)
            a = canon(mwd.mesh.vertices, mwd.mesh.triangles, mwd.point_attributes["normals"] if has_n else one.normals * 0)
            b = canon(one.mesh.vertices, one.mesh.triangles, one.normals if has_n else one.normals * 0)
            good = np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[1], b[1]) and np.abs(a[2] - b[2]).max() <= 2e-5
            print(f"[{name}] DistributedReconstructor world={world}: {len(a[0])} vertices, {len(a[1])} triangles, equal to the single-device call: {good}", flush=True)
            ok &= bool(good)
        rec.close()
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    return 0 if int(flag.item()) else 1


if __name__ == "__main__":
    sys.exit(main())
