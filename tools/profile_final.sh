#!/bin/bash
# Round-end ncu evidence for profiles/ (run on the GPU box under gpurun): tools/profile_final.sh <tag>
# Every profiled command is first run plain (exit 0).
set -u
tag=${1:-r2z}
out=gpurun_out
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err || { echo "bench failed"; tail -5 $out/${tag}_bench.err; exit 1; }
echo "bench rc=0"; cut -c1-300 $out/${tag}_bench.json
python bench.py --mode fp64 --steps 300 --no-legs --no-cpu-baseline --e2e-steps 20 --e2e-warmup 100 --no-e2e-variants > $out/${tag}_bench_fp64.json 2> $out/${tag}_bench_fp64.err; echo "bench fp64 rc=$?"
# the driver's command shape, with a short pre-roll so that the launch list stays small
short="--steps 20 --warmup 5 --preroll 40 --leg-steps 10 --e2e-steps 3 --e2e-warmup 3 --no-cpu-baseline --no-e2e-variants --no-amppo"
python bench.py $short > $out/${tag}_short.json 2> $out/${tag}_short.err || { echo "short bench failed"; tail -5 $out/${tag}_short.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/${tag}_launches.csv python bench.py $short > $out/${tag}_ncu_launches.log 2>&1
echo "launch list rc=$?"
cap() {   # name, ab_step args
  name=$1; shift
  python tools/ab_step.py "$@" --steps 100 > $out/${tag}_${name}_plain.txt 2>&1 || { echo "plain $name failed"; return; }
  cat $out/${tag}_${name}_plain.txt
  ncu --set full --clock-control none --import-source on -k regex:f16_step_kernel --launch-skip 650 --launch-count 1 -o $out/${tag}_step_${name} -f \
      python tools/ab_step.py "$@" --steps 100 > $out/${tag}_${name}_ncu.log 2>&1
  echo "ncu $name rc=$?"
  ncu -i $out/${tag}_step_${name}.ncu-rep --page raw --csv > $out/${tag}_step_${name}_raw.csv 2>/dev/null
}
cap fp32_ring --mode fp32 --layout ring
cap fp32_stacked --mode fp32 --layout stacked
cap fp32_frame --mode fp32 --layout frame
cap fp32_ring_all_details --mode fp32 --layout ring --ground on --reset carryover
cap fp64_ring --mode fp64 --layout ring
ls -la $out/${tag}_* | awk '{print $5, $9}'
