#!/bin/bash
# Per-kernel durations of one fit step for a kernel-name regex (run on the GPU box):  tools/step_kernels.sh <regex> <out.csv> [env...]
re=$1; out=$2; shift 2
env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -k regex:$re --csv --log-file $out \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - "$out" <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]; ik = h.index("Kernel Name"); iv = h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try: t = float(r[iv].replace(",", ""))
    except ValueError: continue
    agg.setdefault(r[ik][:60], []).append(t / 1e6)
for k, v in agg.items():
    print(f"{k:60s} n={len(v):3d} last={v[-1]:8.3f} ms  min={min(v):8.3f}")
PY
