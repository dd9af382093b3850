#!/bin/bash
# One GPU-box session that answers the open questions of round 1 (run under gpurun, ONE GPU):
#
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh'
#
# Everything is wrapped in `timeout`; outputs go to gpurun_out/ (merged back by gpurun).  Nothing printed under ncu is a
# bench value.  Order: correctness first, then the two level-set variants, then profiles.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"

echo "== gpu tests"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== parity of level-set variant 1 on the GPU (same assertions as the seeded parity tests)"
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/variant1_parity.log
import sys; sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, oracle, splashsurf_b200 as ss
from splashsurf_b200 import synthetic as syn
import test_gpu_parity as G
bad = 0
for name, gen, kw in G.SEEDED:
    p = gen(syn); o = oracle.reconstruct(p, **kw)
    ctx = ss.Context(); ctx.set_levelset_variant(1)
    g = ss.reconstruct_surface(p, with_debug=True, context=ctx, **kw); ctx.close()
    m = oracle.mesh_parity(g.mesh.vertices, g.mesh.triangles, g.vertex_edge_keys, o["vertices"], o["triangles"], o["vertex_keys"], kw.get("subdomain_num_cubes_per_dim", 64))
    ok = np.array_equal(g.particle_densities, o["particle_densities"]) and m["keys_equal"] and m["triangles_equal"] and m["n_not_bitexact"] == 0
    bad += not ok
    print(name, "ok" if ok else m, g.timings["levelset_launches"])
print("variant 1:", "ALL BIT-EXACT" if not bad else f"{bad} MISMATCHES")
PY
for v in 0 1; do
  echo "== bench, level-set variant $v"
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --levelset-variant $v > gpurun_out/bench_variant$v.json 2> gpurun_out/bench_variant$v.err
  tail -c 600 gpurun_out/bench_variant$v.json
done
echo "== post-processing entries"; timeout 300 python tools/bench_postprocess.py --particles 10000000 > gpurun_out/bench_postprocess.json 2>&1; tail -c 800 gpurun_out/bench_postprocess.json
echo "== launch list (default bench command, 2 steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
for v in 0 1; do
  k=$([ $v = 0 ] && echo k_levelset || echo k_certify)
  echo "== ncu --set full of $k (4 M particles)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/prof_${k}_r2 -f \
      python bench.py --particles 4000000 --steps 1 --warmup 1 --no-cpu-baseline --levelset-variant $v > gpurun_out/ncu_$k.log 2>&1
done
echo "== memcheck on a small case"; timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/memcheck.log
