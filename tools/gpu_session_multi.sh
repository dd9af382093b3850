#!/bin/bash
# Multi-GPU session (run under `gpurun --gpus N`, N = 2, 4 or 8):   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_session_multi.sh 2'
# Every multi-rank command sits under a short `timeout`: a mismatched collective must cost seconds, not the GPU budget
# (round 1 lost its remaining budget to one such hang).
set -u
N=${1:-2}
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
run() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
echo "== parity of the gathered + welded mesh ($N ranks)"
run 300 29611 tools/mgpu_check.py 2>&1 | grep "^\[" | cut -c1-220 | tee gpurun_out/mgpu_check_$N.log
for proto in two_call callback; do
  echo "== bench, $N ranks, runner protocol $proto"
  run 600 29612 bench.py --gpus "$N" --steps 3 --warmup 3 --runner-protocol $proto > gpurun_out/bench_${N}gpu_$proto.json 2> gpurun_out/bench_${N}gpu_$proto.err
  tail -c 500 gpurun_out/bench_${N}gpu_$proto.json; grep "device ms/step" gpurun_out/bench_${N}gpu_$proto.err | cut -c1-120
done
