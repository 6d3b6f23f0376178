#!/bin/bash
# Steady-state ncu evidence (run on the GPU box under gpurun): tools/profile_r2.sh <tag> [fp32|fp64|both] [layout]
# Every profiled command is first run plain (exit 0). The capture is launch 650 of the step kernel: 600 pre-roll steps
# after reset, so episodes are ending and auto-resetting in the captured launch.
set -u
tag=${1:-r2x}; which=${2:-both}; layout=${3:-ring}
out=gpurun_out
for mode in fp32 fp64; do
  [ "$which" != both ] && [ "$which" != $mode ] && continue
  args="--mode $mode --layout $layout --steps 100"
  python tools/ab_step.py $args > $out/${tag}_${mode}_plain.txt 2>&1 || { echo "plain run failed"; tail -5 $out/${tag}_${mode}_plain.txt; exit 1; }
  cat $out/${tag}_${mode}_plain.txt
  ncu --set full --clock-control none --import-source on -k regex:f16_step_kernel --launch-skip 650 --launch-count 1 \
      -o $out/${tag}_step_${mode}_${layout} -f python tools/ab_step.py $args > $out/${tag}_${mode}_ncu.log 2>&1
  echo "ncu $mode rc=$?"
  ncu -i $out/${tag}_step_${mode}_${layout}.ncu-rep --page raw --csv > $out/${tag}_step_${mode}_${layout}_raw.csv 2>/dev/null
done
ls -la $out/${tag}_*
