#!/usr/bin/env python3
"""Full-size parity on the GPU box: the reference's own binary (pysplashsurf 0.14.0 wheel, oracle/_ref) against the CUDA path
on the BASELINE workloads (cfg-3: 10 M, cfg-4: 50 M particles), and the N-GPU result against the 1-GPU result.

  python tools/parity_full.py --workload cfg4 --out gpurun_out/parity_cfg4_1gpu.json            # 1 GPU: wheel vs CUDA
  torchrun --nproc-per-node 8 tools/parity_full.py --workload cfg4 --against profiles/parity_cfg4_1gpu.json \
           --out gpurun_out/parity_cfg4_8gpu.json                                               # N GPUs: digests vs the 1-GPU run

What is compared:
  * particle densities, bit for bit (1 GPU);
  * the vertex key sets: every reference vertex is matched to the CUDA vertex on the same marching-cubes grid edge.  The
    reference's keys are derived from its vertex POSITIONS ALONE (no oracle involved): q = (v - grid.min) / cell, the one
    non-integer coordinate is the edge axis.  Vertices within 1e-3 cells of a lattice point (interpolation weight ~0 or ~1, where
    the position cannot tell which of the six incident edges carries the vertex) are matched to the nearest CUDA vertex on
    an edge incident to that lattice point; the triangle comparison below then validates that choice;
  * the triangle arrays after canonical ordering (vertices sorted by edge key, triangles rotated to their smallest index and
    sorted): must be identical;
  * vertex positions: bit-exact for every vertex that does not lie on a subdomain face (there the reference keeps whichever
    subdomain's copy its hash map saw first, a machine-dependent <= 2e-6 difference), max abs / rel difference over all.
Digests (order-sensitive 64-bit sums over the canonical arrays) make results comparable across runs and boxes.
All heavy array work runs in torch on the GPU (sorts over 60 M triangles).  Test/measurement infrastructure, not product.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

RECON_KW = dict(particle_radius=0.01, smoothing_length=2.0, cube_size=0.5, iso_surface_threshold=0.6)
TOL_CELLS = 1.0e-3


def revision():
    try:
        return open(os.path.join(ROOT, ".revision")).read().strip()
    except OSError:
        return "unknown"


def make_cloud(workload, particles=None):
    from splashsurf_b200 import synthetic as syn
    if workload == "cfg3":
        return syn.dam_break_10m(), dict(RECON_KW)
    if workload == "cfg4":
        return syn.dam_break_50m(), dict(RECON_KW)
    if workload == "cfg5":
        kw = dict(RECON_KW); kw.update(particle_radius=0.005, cube_size=0.45)
        return syn.splash_200m(), kw
    if workload == "scaled":
        return syn.dam_break_scaled(int(particles), 0.01, 3), dict(RECON_KW)
    raise SystemExit("unknown workload")


# ------------------------------------------------------------------ torch helpers ----
def pack(k4):
    return (k4[:, 0] << 42) | (k4[:, 1] << 22) | (k4[:, 2] << 2) | k4[:, 3]


def digest(t):
    """Order-sensitive 64-bit digest of an integer tensor (wrapping int64 arithmetic)."""
    import torch
    x = t.reshape(-1).to(torch.int64)
    idx = torch.arange(x.numel(), dtype=torch.int64, device=x.device)
    w = (idx * -7046029254386353131 + 1442695040888963407) | 1
    return format(int(((x + 0x5851F42D) * w).sum().item()) & 0xFFFFFFFFFFFFFFFF, "016x")


def canonicalize(v, t, k):
    """v (n,3) f32, t (m,3) int64, k (n,) packed int64 unique keys -> canonical (v, t, k)."""
    import torch
    order = torch.argsort(k)
    inv = torch.empty_like(order)
    inv[order] = torch.arange(len(order), device=order.device)
    v, k = v[order], k[order]
    t = inv[t]
    m = t.argmin(1)
    t = torch.stack([t.gather(1, ((m + s) % 3)[:, None])[:, 0] for s in range(3)], dim=1)
    for col in (2, 1, 0):
        t = t[torch.sort(t[:, col], stable=True).indices]
    return v, t, k


def keys_from_positions(v, gmin, cell):
    """-> (packed key guess, ambiguous mask, nearest lattice point (n,3)); float64 on the f32 positions."""
    import torch
    q = (v.double() - gmin[None]) / cell
    rq = torch.round(q)
    frac = (q - rq).abs()
    axis = frac.argmax(1)
    ijk = rq.to(torch.int64)
    rows = torch.arange(len(v), device=v.device)
    ijk[rows, axis] = torch.floor(q[rows, axis]).to(torch.int64)
    amb = frac.max(1).values < TOL_CELLS
    return pack(torch.cat([ijk, axis[:, None]], 1)), amb, rq.to(torch.int64)


def match_reference_keys(rv, rt, gv, gt, gk, gmin, cell):
    """Assign to every reference vertex the index of the CUDA vertex on the same grid edge.  Returns (assign, stats).

    Unambiguous vertices: by the key derived from the position.  Ambiguous vertices (within TOL_CELLS of a lattice point P):
    candidates are the CUDA vertices on the six edges incident to P; the candidate that shares the most mesh neighbours with
    the reference vertex wins (neighbours = unambiguous vertices of its triangles, already matched), ties by distance."""
    import torch
    dev = rv.device
    kr, amb, P = keys_from_positions(rv, gmin, cell)
    sk, perm = torch.sort(gk)
    n = len(sk)
    nvg = len(gv)

    def lookup(keys):
        pos = torch.searchsorted(sk, keys).clamp(max=max(n - 1, 0))
        hit = sk[pos] == keys
        return torch.where(hit, perm[pos], torch.full_like(pos, -1))

    assign = lookup(kr)
    assign[amb] = -1
    stats = {"n_ambiguous": int(amb.sum()), "n_unambiguous_missing": int(((assign < 0) & ~amb).sum())}
    ia = torch.nonzero(amb)[:, 0]
    if len(ia):
        Pa = P[ia]
        cands, dists = [], []
        for ax in range(3):
            for off in (0, -1):
                pk = Pa.clone()
                pk[:, ax] += off
                c = lookup(pack(torch.cat([pk, torch.full((len(pk), 1), ax, dtype=torch.int64, device=dev)], 1)))
                d = (gv[c.clamp(min=0)].double() - rv[ia].double()).abs().max(1).values
                d = torch.where(c >= 0, d, torch.full_like(d, 1.0e3))
                cands.append(c); dists.append(d)
        cands, dists = torch.stack(cands, 1), torch.stack(dists, 1)           # (nA, 6)
        # -- shared-neighbour score
        row_of = torch.full((len(rv),), -1, dtype=torch.int64, device=dev)
        row_of[ia] = torch.arange(len(ia), device=dev)
        tsel = rt[amb[rt].any(1)]                                              # reference triangles touching an ambiguous vertex
        src = torch.cat([tsel[:, [0, 0, 1, 1, 2, 2]].reshape(-1)])
        dst = torch.cat([tsel[:, [1, 2, 0, 2, 0, 1]].reshape(-1)])
        keep = amb[src] & ~amb[dst] & (assign[dst] >= 0)
        src_row, nb = row_of[src[keep]], assign[dst[keep]]                      # (E,), neighbours as CUDA vertex ids
        is_cand = torch.zeros(nvg, dtype=torch.bool, device=dev)
        is_cand[cands[cands >= 0]] = True
        gsel = gt[is_cand[gt].any(1)]
        gs = gsel[:, [0, 0, 1, 1, 2, 2]].reshape(-1)
        gd = gsel[:, [1, 2, 0, 2, 0, 1]].reshape(-1)
        gkeep = is_cand[gs]
        gedges = torch.unique(gs[gkeep] * nvg + gd[gkeep])                      # sorted
        score = torch.zeros(cands.shape, dtype=torch.int64, device=dev)
        for kcol in range(6):
            ck = cands[src_row, kcol]
            q = ck.clamp(min=0) * nvg + nb
            pos = torch.searchsorted(gedges, q).clamp(max=max(len(gedges) - 1, 0))
            hit = (gedges[pos] == q) & (ck >= 0) if len(gedges) else torch.zeros_like(ck, dtype=torch.bool)
            score[:, kcol].index_add_(0, src_row, hit.to(torch.int64))
        rank = -score.double() * 1.0e6 + dists                                 # more shared neighbours first, then nearer
        srt = torch.argsort(rank, dim=1)
        cands, dists, score = cands.gather(1, srt), dists.gather(1, srt), score.gather(1, srt)
        assign[ia] = cands[:, 0]
        taken = torch.bincount(assign[assign >= 0], minlength=nvg)
        conflicted = ia[(assign[ia] >= 0) & (taken[assign[ia].clamp(min=0)] > 1)]
        stats["n_ambiguous_conflicts"] = int(len(conflicted))
        if 0 < len(conflicted) <= 500000:
            # the best-ranked claimant keeps a contested candidate, the others move on to their next free candidate
            a_cpu, c_cpu, r_cpu = assign.cpu().numpy(), cands.cpu().numpy(), rank.gather(1, srt).cpu().numpy()
            conf = conflicted.cpu().numpy()
            rows = row_of[conflicted].cpu().numpy()
            order = np.argsort(r_cpu[rows, 0], kind="stable")
            all_taken = np.zeros(nvg, dtype=bool)
            all_taken[a_cpu[a_cpu >= 0]] = True
            all_taken[a_cpu[conf]] = False
            for oi in order:
                a, r = int(conf[oi]), int(rows[oi])
                a_cpu[a] = -1
                for c in c_cpu[r]:
                    if c >= 0 and not all_taken[c]:
                        a_cpu[a] = int(c); all_taken[c] = True; break
            assign = torch.from_numpy(a_cpu).to(dev)
    stats["n_unmatched"] = int((assign < 0).sum())
    ok = assign >= 0
    stats["bijective"] = bool(stats["n_unmatched"] == 0 and len(rv) == len(gv) and int(torch.bincount(assign[ok], minlength=nvg).max()) == 1)
    return assign, stats


def compare(rv, rt, gv, gt, gk4, gmin, cell, S):
    """All inputs torch tensors on one device; gk4 (n,4) int64 true CUDA keys."""
    import torch
    out = {"nv": [int(len(rv)), int(len(gv))], "nt": [int(len(rt)), int(len(gt))]}
    gk = pack(gk4)
    assign, st = match_reference_keys(rv, rt, gv, gt, gk, gmin, cell)
    out.update(st)
    out["keys_equal"] = bool(st["bijective"])
    gvc, gtc, gkc = canonicalize(gv, gt, gk)
    out["digest_cuda"] = {"keys": digest(gkc), "triangles": digest(gtc), "positions": digest(gvc.view(torch.int32))}
    if not out["keys_equal"]:
        out["triangles_equal"] = False
        return out
    rk = gk[assign]
    rvc, rtc, rkc = canonicalize(rv, rt, rk)
    out["triangles_equal"] = bool(rtc.shape == gtc.shape and torch.equal(rtc, gtc))
    diff = (rvc.double() - gvc.double()).abs()
    scale = torch.maximum(rvc.abs(), gvc.abs()).double().clamp(min=1e-30)
    out["max_abs"] = float(diff.max()); out["max_rel"] = float((diff / scale).max())
    neq = (rvc.view(torch.int32) != gvc.view(torch.int32)).any(1)
    out["n_not_bitexact"] = int(neq.sum())
    ax = gkc & 3
    comp = [(gkc >> 42) & 0xFFFFF, (gkc >> 22) & 0xFFFFF, (gkc >> 2) & 0xFFFFF]
    onface = torch.zeros_like(neq)
    for d in range(3):
        onface |= (ax != d) & (comp[d] % S == 0)
    out["n_on_subdomain_faces"] = int(onface.sum())
    out["n_interior_not_bitexact"] = int((neq & ~onface).sum())
    out["digest_reference"] = {"triangles": digest(rtc), "positions_interior": digest(rvc.view(torch.int32)[~onface])}
    out["digest_cuda"]["positions_interior"] = digest(gvc.view(torch.int32)[~onface])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg4", "cfg5", "scaled"])
    ap.add_argument("--particles", type=int, default=400000)
    ap.add_argument("--out", default=None)
    ap.add_argument("--against", default=None, help="N > 1: JSON of the 1-GPU run whose CUDA digests the gathered mesh must reproduce")
    ap.add_argument("--levelset-variant", type=int, default=None)
    ap.add_argument("--no-reference", action="store_true", help="1 GPU: skip the wheel (digests only)")
    args = ap.parse_args()
    import torch
    import splashsurf_b200 as ss
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    p, kw = make_cloud(args.workload, args.particles)
    S = 64
    res = {"tool": "tools/parity_full.py", "revision": revision(), "workload": args.workload, "particles": int(len(p)), "n_gpus": world,
           "kwargs": kw, "tolerance": "bit-exact (densities, keys, triangles, interior vertex positions); subdomain-face vertices <= 2e-6 abs"}

    if world > 1:
        import datetime
        import torch.distributed as dist
        from splashsurf_b200 import distributed as ssd
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=int(os.environ.get("SS_MGPU_TIMEOUT", "180"))))
        ctx = ss.Context(local)
        if args.levelset_variant is not None:
            ctx.set_levelset_variant(args.levelset_variant)
        runner = ssd.Runner(ctx, ss.make_params(**kw), world, rank, local)
        runner.want_keys = True
        x = torch.from_numpy(runner.take_local(p)).cuda()
        del p
        t0 = time.perf_counter()
        r = runner.step(x, copy_out=True)
        torch.cuda.synchronize()
        res["cuda_seconds_first_call"] = time.perf_counter() - t0
        ok = True
        if rank == 0:
            v, t = runner.gathered_mesh(r["nv_global"], r["nt_global"])
            K = r["keys_global"].to(dev)
            runner.close()
            ctx.close()
            gv, gt = torch.from_numpy(v).to(dev), torch.from_numpy(t.astype(np.int64)).to(dev)
            gvc, gtc, gkc = canonicalize(gv, gt, K)
            res["nv"], res["nt"] = int(len(gv)), int(len(gt))
            res["digest_cuda"] = {"keys": digest(gkc), "triangles": digest(gtc), "positions": digest(gvc.view(torch.int32))}
            res["plan"] = {"axis": r["plan"].axis, "cuts": r["plan"].cuts}
            if args.against and os.path.exists(args.against):
                ref = json.load(open(args.against))
                res["against"] = {"file": args.against, "revision": ref.get("revision"), "digest_cuda": ref["mesh"]["digest_cuda"]}
                res["equal_to_1gpu"] = {k: res["digest_cuda"][k] == ref["mesh"]["digest_cuda"].get(k) for k in ("keys", "triangles", "positions")}
                ok = all(res["equal_to_1gpu"].values())
            else:
                res["against"] = None
            res["pass"] = bool(ok)
            print(json.dumps(res))
            if args.out:
                json.dump(res, open(args.out, "w"), indent=1)
        else:
            runner.close()
            ctx.close()
        dist.barrier()
        dist.destroy_process_group()
        return 0

    # ---- one GPU: CUDA path
    ctx = ss.Context(local)
    if args.levelset_variant is not None:
        ctx.set_levelset_variant(args.levelset_variant)
    t0 = time.perf_counter()
    g = ss.reconstruct_surface(p, context=ctx, with_debug=True, **kw)
    res["cuda_seconds_first_call"] = time.perf_counter() - t0
    res["cuda_timings_ms"] = {k: round(float(v), 3) for k, v in g.timings.items()}
    ctx.close()
    gmin = torch.tensor(np.asarray(g.grid.aabb.min, dtype=np.float64), device=dev)
    cell = float(np.float32(g.grid.cell_size))
    gv = torch.from_numpy(g.mesh.vertices).to(dev)
    gt = torch.from_numpy(g.mesh.triangles.astype(np.int64)).to(dev)
    gk4 = torch.from_numpy(g.vertex_edge_keys).to(dev)
    if args.no_reference:
        gvc, gtc, gkc = canonicalize(gv, gt, pack(gk4))
        res["mesh"] = {"nv": int(len(gv)), "nt": int(len(gt)),
                       "digest_cuda": {"keys": digest(gkc), "triangles": digest(gtc), "positions": digest(gvc.view(torch.int32))}}
        res["pass"] = None
    else:
        import oracle
        ps = oracle.reference()
        t0 = time.perf_counter()
        r = ps.reconstruct_surface(p, multi_threading=True, simd=True, subdomain_grid=True, subdomain_num_cubes_per_dim=S, **kw)
        res["reference_seconds"] = time.perf_counter() - t0
        res["reference_threads"] = os.cpu_count()
        rho_r = np.asarray(r.particle_densities)
        res["densities_bitexact"] = bool(np.array_equal(rho_r.view(np.uint32), g.particle_densities.view(np.uint32)))
        res["grid_equal"] = bool(np.array_equal(np.asarray(r.grid.aabb.min, np.float32), np.asarray(g.grid.aabb.min, np.float32))
                                 and list(r.grid.ncells_per_dim) == list(g.grid.ncells_per_dim))
        rv = torch.from_numpy(np.ascontiguousarray(np.asarray(r.mesh.vertices, dtype=np.float32))).to(dev)
        rt = torch.from_numpy(np.asarray(r.mesh.triangles).astype(np.int64)).to(dev)
        del r
        res["mesh"] = compare(rv, rt, gv, gt, gk4, gmin, cell, S)
        m = res["mesh"]
        res["pass"] = bool(res["densities_bitexact"] and res["grid_equal"] and m["keys_equal"] and m["triangles_equal"]
                           and m.get("n_interior_not_bitexact", 1) == 0 and m.get("max_abs", 1.0) <= 2.0e-6 * 4)
    print(json.dumps(res))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)
    return 0 if res["pass"] in (True, None) else 1


if __name__ == "__main__":
    sys.exit(main())
