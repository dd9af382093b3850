#!/bin/bash
# Multi-GPU session (run under `gpurun --gpus N`, N = 2, 4 or 8):   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_session_multi.sh 2 [stages]'
# Stages (default: check parity bench): check = small-case parity at every rank count <= N, parity = cfg-4 digests vs the 1-GPU run,
# bench = bench.py at N ranks (both runner protocols), cfg5 = the 200 M splash with SPH normals.
# Every multi-rank command sits under a short `timeout`: a mismatched collective must cost seconds, not the GPU budget.
set -u
N=${1:-2}
shift || true
STAGES="${*:-check parity bench}"
has() { [[ " $STAGES " == *" $1 "* ]]; }
TAG=${SS_TAG:-r2}
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
run() { local n=$1 t=$2 port=$3; shift 3; SS_MGPU_TIMEOUT=90 timeout "$t" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port "$port" "$@"; }
if has check; then
  for n in 2 4 8; do
    [ "$n" -le "$N" ] || continue
    echo "== parity of the assembled mesh, small cases ($n ranks)"
    run $n 240 $((29610 + n)) tools/mgpu_check.py 2>&1 | grep "^\[" | cut -c1-260 | tee gpurun_out/mgpu_check_${n}_$TAG.log
  done
fi
if has parity; then
  echo "== cfg-4 at $N ranks: digests of the assembled mesh vs the 1-GPU run"
  run $N 600 29631 tools/parity_full.py --workload cfg4 --against profiles/parity_cfg4_1gpu.json --out gpurun_out/parity_cfg4_${N}gpu.json 2> gpurun_out/parity_cfg4_${N}gpu.err | cut -c1-900
  tail -2 gpurun_out/parity_cfg4_${N}gpu.err | cut -c1-300
fi
if has bench; then
  for proto in ${SS_PROTOCOLS:-stats}; do
    echo "== bench, $N ranks, runner protocol $proto"
    run $N 900 29641 bench.py --gpus "$N" --steps ${SS_STEPS:-20} --warmup 5 --runner-protocol $proto > gpurun_out/bench_${N}gpu_${proto}_$TAG.json 2> gpurun_out/bench_${N}gpu_${proto}_$TAG.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${N}gpu_${proto}_$TAG.json"))
    print({k: d.get(k) for k in ("value", "ms_per_step", "e2e", "step_ms_rank0", "stage_ms_last_step", "mesh")})
except Exception as e:
    print("bench failed:", e); print(open("gpurun_out/bench_${N}gpu_${proto}_$TAG.err").read()[-2500:])
PY
    grep "device ms/step\|runner phases" gpurun_out/bench_${N}gpu_${proto}_$TAG.err | cut -c1-260
  done
fi
if has cfg5; then
  echo "== cfg-5 (200 M splash, c = 0.45 r, SPH normals) at $N ranks"
  run $N 1500 29651 bench.py --gpus "$N" --workload cfg5 --steps ${SS_CFG5_STEPS:-6} --warmup ${SS_CFG5_WARMUP:-6} --sph-normals > gpurun_out/bench_cfg5_${N}gpu_$TAG.json 2> gpurun_out/bench_cfg5_${N}gpu_$TAG.err
  cut -c1-1500 gpurun_out/bench_cfg5_${N}gpu_$TAG.json; grep "device ms/step\|runner phases" gpurun_out/bench_cfg5_${N}gpu_$TAG.err | cut -c1-260
fi
echo "== session done"
