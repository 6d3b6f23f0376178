#!/bin/bash
# A/B of the low-tiles-first ordering (F16_PRIORITY) on the GPU box: tools/prio_ab.sh <tag>
tag=$1
for p in 0 1; do
  for cfg in "fp32 1500" "fp64 400"; do
    set -- $cfg
    F16_PRIORITY=$p python bench.py --mode $1 --ground on --steps $2 --warmup 3 --e2e-steps 0 --e2e-warmup 3 --no-cpu-baseline --no-e2e-variants > gpurun_out/${tag}_p${p}_$1.json 2> gpurun_out/${tag}_p${p}_$1.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_p${p}_$1.json").read().strip().splitlines()[-1])
print("priority=$p $1 ground on: ms %.4f value %.4e | off %.4f | carry-over %.4f" % (d["ms_per_step"], d["value"], d["config"]["ground_reactions"]["other_setting_ms_per_step"], d["config"]["reset"]["carryover_with_ground_reactions_ms_per_step"]))
PY
  done
done
