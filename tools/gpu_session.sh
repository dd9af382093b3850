#!/bin/bash
# One-GPU session (run under gpurun):   git rev-parse HEAD > .revision; gpurun --timeout 2400 -- 'bash tools/gpu_session.sh [stages]'
# Stages (default: all):  tests variant parity bench ref launches ncu post memcheck
# Everything is wrapped in `timeout`; outputs go to gpurun_out/ (merged back by gpurun).  Nothing printed under ncu is a bench value.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
STAGES="${*:-tests variant parity bench ref launches ncu post memcheck}"
has() { [[ " $STAGES " == *" $1 "* ]]; }
TAG=${SS_TAG:-r2}

if has tests; then
  echo "== gpu tests"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_$TAG.log
fi
if has variant; then
  echo "== parity of level-set variant 1 on the GPU (same assertions as the seeded parity tests)"
  timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/variant1_parity_$TAG.log
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
fi
if has parity; then
  for wl in cfg3 cfg4; do
    echo "== full-size parity vs the reference wheel, $wl"
    timeout 900 python tools/parity_full.py --workload $wl --out gpurun_out/parity_${wl}_1gpu.json 2> gpurun_out/parity_${wl}_1gpu.err | cut -c1-1500
    tail -3 gpurun_out/parity_${wl}_1gpu.err
  done
fi
if has bench; then
  for v in ${SS_VARIANTS:-0 1}; do
    echo "== bench, level-set variant $v"
    extra=$([ $v = 0 ] && echo "" || echo "--no-cpu-baseline")
    timeout 900 python bench.py --steps 5 --warmup 3 $extra --levelset-variant $v > gpurun_out/bench_variant${v}_$TAG.json 2> gpurun_out/bench_variant${v}_$TAG.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_variant${v}_$TAG.json"))
    print({k: d.get(k) for k in ("value", "ms_per_step", "e2e", "stage_ms_last_step", "cpu_baseline", "cpu_stage_ms", "gpu_stage_ms_same_sample", "mesh")})
    print(d["roofline"])
except Exception as e:
    print("bench failed:", e); print(open("gpurun_out/bench_variant${v}_$TAG.err").read()[-1500:])
PY
  done
  echo "== bench cfg3"
  timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_$TAG.json 2> gpurun_out/bench_cfg3_$TAG.err; cut -c1-700 gpurun_out/bench_cfg3_$TAG.json
fi
if has ref; then
  echo "== reference arm on the full workload"
  timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err; cut -c1-900 gpurun_out/bench_reference_$TAG.json
fi
if has launches; then
  echo "== launch list (default bench command, 2 steps)"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
  wc -l gpurun_out/launches_$TAG.csv
fi
if has ncu; then
  echo "== ncu --set full: k_levelset (both launches) + k_density on cfg3 (10 M particles)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_levelset|k_density' -c 3 -o gpurun_out/prof_levelset_density_$TAG -f \
      python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
  echo "== ncu --set full: k_certify (variant 1) on cfg3"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_certify' -c 1 -o gpurun_out/prof_certify_$TAG -f \
      python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu-baseline --levelset-variant 1 > gpurun_out/ncu_full_v1_$TAG.log 2>&1
  ls -la gpurun_out/*.ncu-rep
fi
if has ncu2; then
  echo "== ncu --set full, level-set variant 2 on cfg3: k_density, k_certify_warp, k_levelset (exact pass, fix-up pass)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_certify_warp|k_exact_warp|k_density<' -c 4 -o gpurun_out/prof_variant2_$TAG -f \
      python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu-baseline --levelset-variant 2 > gpurun_out/ncu_full_v2_$TAG.log 2>&1
  ls -la gpurun_out/prof_variant2_$TAG.ncu-rep
fi
if has ab; then
  echo "== A/B against the defaults (level-set variant 2, density variant 2, mc variant 1)"
  for ab in ${SS_AB:-"--density-variant=1" "--density-variant=0" "--mc-variant=0"}; do
    timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $ab > gpurun_out/bench_ab_$TAG.json 2> gpurun_out/bench_ab_$TAG.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_ab_$TAG.json")); print("$ab", {k: d.get(k) for k in ("value", "ms_per_step", "stage_ms_last_step")})
except Exception as e:
    print("bench failed:", e); print(open("gpurun_out/bench_ab_$TAG.err").read()[-1500:])
PY
  done
fi
if has ncu4; then
  echo "== ncu --set full on the bench workload (cfg-4, 50 M): density, certification, exact pass (+ fix-up pass), fix-up sweep, marching cubes"
  python -c "from splashsurf_b200 import build; print(build.source_hash(), build.levelset_source_hash())" > gpurun_out/source_sha_$TAG.txt
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_certify_warp|k_exact_warp|k_density_cells|k_mc_count_warp|k_mc_emit_warp|k_fixup_flags_warp' -c 7 \
      -o gpurun_out/prof_cfg4_$TAG -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_cfg4_$TAG.log 2>&1
  ls -la gpurun_out/prof_cfg4_$TAG.ncu-rep; cat gpurun_out/source_sha_$TAG.txt
fi
if has launches5; then
  echo "== launch list of one cfg-5 step (200 M splash, SPH normals)"
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_cfg5_$TAG.csv \
      python bench.py --workload cfg5 --sph-normals --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench_cfg5_$TAG.log 2>&1
  wc -l gpurun_out/launches_cfg5_$TAG.csv
fi
if has cfg5; then
  echo "== cfg-5 (200 M splash, c = 0.45 r, SPH normals) on one GPU"
  timeout 1200 python bench.py --workload cfg5 --sph-normals --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cfg5_1gpu_$TAG.json 2> gpurun_out/bench_cfg5_1gpu_$TAG.err
  cut -c1-1800 gpurun_out/bench_cfg5_1gpu_$TAG.json; tail -3 gpurun_out/bench_cfg5_1gpu_$TAG.err | cut -c1-300
fi
if has ncu5; then
  echo "== ncu --set full on cfg-5: k_certify_warp, k_exact_warp, k_exact_warp_big, k_sph_normals"
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_certify_warp|k_exact_warp|k_sph_normals' -c 5 -o gpurun_out/prof_cfg5_$TAG -f \
      python bench.py --workload cfg5 --sph-normals --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_cfg5_$TAG.log 2>&1
  ls -la gpurun_out/prof_cfg5_$TAG.ncu-rep
fi
if has post; then
  echo "== post-processing entries"; timeout 300 python tools/bench_postprocess.py --particles 10000000 > gpurun_out/bench_postprocess_$TAG.json 2>&1; tail -c 800 gpurun_out/bench_postprocess_$TAG.json
fi
if has memcheck; then
  echo "== memcheck on a small case"; timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/memcheck_$TAG.log
fi
echo "== session done"
