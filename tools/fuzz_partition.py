#!/usr/bin/env python3
"""Differential fuzzing of the slab-partition (multi-GPU) entries WITHOUT a GPU: random clouds and parameters, 2-4 virtual ranks with
random slab cuts (idle ranks included) run one after another on the CPU executor of the CUDA sources, their meshes are welded like
rank 0 does, and the result has to equal the oracle's single-device mesh bit for bit.  Every case also checks that the plan statistics
of the "stats" protocol (ss_partition_members_f32 on rank-local particles, summed) give the same global maximum subdomain population
as the decomposition pre-pass of the "two_call" protocol.

    python tools/fuzz_partition.py --cases 200 --seed 0
"""
import argparse, ctypes as C, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--max-points", type=float, default=4e6)
    a = ap.parse_args()
    import oracle
    import splashsurf_b200 as ss
    from test_emulated_pipeline import build_emulated_library, _virtual_ranks
    from fuzz_emulated import random_case
    ss.load_library()
    ss._LIB = ss._bind(C.CDLL(build_emulated_library()))
    bad = done = 0
    t0 = time.time()
    for i in range(a.cases):
        seed = a.seed * 1_000_003 + i
        rng = np.random.default_rng(seed)
        x, kw, _ = random_case(rng)
        kw.pop("aabb_min", None); kw.pop("aabb_max", None)                    # the partitioned entry takes pre-filtered particles
        kw.pop("subdomain_grid", None)
        kw.update(subdomain_num_cubes_per_dim=int(rng.choice([8, 12, 16, 20, 24, 32])), subdomain_grid_auto_disable=False)
        o = oracle.reconstruct(x, **kw)
        if o["rc"] != 0 or float(np.prod(o["grid"]["npoints"].astype(np.float64))) > a.max_points:
            continue
        S = kw["subdomain_num_cubes_per_dim"]
        nl = int(max((int(c) + S - 1) // S for c in o["grid"]["ncells"]))
        world = int(rng.integers(2, 5))
        cuts = None
        if rng.integers(0, 2):
            cuts = [0] + sorted(int(c) for c in rng.integers(0, nl + 1, size=world - 1)) + [nl]
        try:
            v, t, keys, plan, nrecv = _virtual_ranks(ss, oracle, x, kw, world, bool(rng.integers(0, 2)), cuts)
        except AssertionError as e:
            bad += 1
            print(f"[{seed}] FAILED n={len(x)} world={world} cuts={cuts} {kw}: {e}")
            continue
        m = oracle.mesh_parity(v, t, keys, o["vertices"], o["triangles"], o["vertex_keys"], S)
        done += 1
        if not (m["keys_equal"] and m["triangles_equal"] and m["n_not_bitexact"] == 0):
            bad += 1
            print(f"[{seed}] MISMATCH n={len(x)} world={world} cuts={plan.cuts} {kw} {m}")
        elif done % 10 == 0:
            print(f"[{seed}] ok n={len(x)} nv={len(v)} world={world} cuts={plan.cuts} ({time.time() - t0:.0f}s)", flush=True)
    print(f"{a.cases} cases, {done} compared, {bad} mismatches, {time.time() - t0:.0f}s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
