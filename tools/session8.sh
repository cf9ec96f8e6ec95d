#!/bin/bash
# The 8-GPU measurement session (BASELINE.json configs[1], [2], [4], the 20M-point scene and a corner of the configs[3] sweep).
# usage: tools/session8.sh [full|cfg2]
set -x
MODE=${1:-full}
tools/scale_run.sh "8" cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-local
if [ "$MODE" = "full" ]; then
  tools/scale_run.sh "8" cfg3 --T 1800 --M 5000000 --scene surface --steps 3 --warmup 3 --no-cpu-baseline --no-local
  tools/scale_run.sh "8" cfg2_20M --M 20000000 --steps 3 --warmup 3 --no-cpu-baseline --no-local
  tools/scale_run.sh "8" cfg5 --T 4800 --clips 16 --M 20000000 --scene surface --steps 3 --warmup 3 --no-cpu-baseline --no-local
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 tools/nn_microbench.py --queries 1000000,3000000 --points 1000000,20000000 --reps 3 2>gpurun_out/microbench_r02_8gpu.err | grep "^{" > gpurun_out/microbench_r02_8gpu.jsonl
cat gpurun_out/microbench_r02_8gpu.jsonl; tail -3 gpurun_out/microbench_r02_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29578 tools/phase_times.py 2>/dev/null | grep "^{" > gpurun_out/phase_times_r02_8gpu.jsonl
cat gpurun_out/phase_times_r02_8gpu.jsonl
