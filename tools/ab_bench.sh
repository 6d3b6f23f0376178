#!/bin/bash
# A/B of library builds on the GPU box: tools/ab_bench.sh <tag> lib1 lib2 ...  ("main" = the in-tree library)
tag=$1; shift
for lib in "$@"; do
  for mode in fp32 fp64; do
    steps=1500; [ $mode = fp64 ] && steps=400
    name=$(basename $lib .so)
    if [ "$lib" = main ]; then unset F16_B200_LIB; else export F16_B200_LIB=$PWD/$lib; fi
    python bench.py --mode $mode --steps $steps --warmup 3 --e2e-steps 0 --e2e-warmup 3 --no-cpu-baseline --no-e2e-variants > gpurun_out/${tag}_${name}_${mode}.json 2> gpurun_out/${tag}_${name}_${mode}.err
    python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_${name}_${mode}.json"))
print("$name $mode ms/step %.4f value %.3e ground_other %.4f carry %.4f" % (d["ms_per_step"], d["value"], d["config"]["ground_reactions"]["other_setting_ms_per_step"], d["config"]["reset"]["carryover_with_ground_reactions_ms_per_step"]))
PY
  done
done
