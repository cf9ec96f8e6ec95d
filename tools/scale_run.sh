#!/bin/bash
# Strong-scaling runs of bench.py on one box: tools/scale_run.sh "<N list>" <tag> [extra bench args...]
# e.g.  tools/scale_run.sh "8 4 2" cfg2 --steps 10 --warmup 3 --no-cpu-baseline --no-local
NS="$1"; TAG="$2"; shift 2
mkdir -p gpurun_out
for N in $NS; do
  OUT=gpurun_out/bench_r02_${TAG}_${N}gpu.json
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 "$@" > $OUT 2> gpurun_out/bench_r02_${TAG}_${N}gpu.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N "$@" > $OUT 2> gpurun_out/bench_r02_${TAG}_${N}gpu.err
  fi
  echo "N=$N rc=$?"; tail -c 300 gpurun_out/bench_r02_${TAG}_${N}gpu.err | tail -2
  python - "$OUT" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(" value", round(d["value"], 2), "steps/s  ms", round(d["ms_per_step"], 3), d["launch"], "| e2e", round(d["e2e"]["value"], 2),
          "|", {k.split(" ")[0]: round(v["ms_mean"], 3) for k, v in d["kernels"].items()}, "| cond", {k: (round(v, 2) if v else v) for k, v in d["ms_per_step_by_condition"].items()})
except Exception as ex:
    print(" no line:", ex)
PY
done
