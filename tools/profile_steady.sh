#!/bin/bash
# Steady-state ncu evidence for profiles/ (run on the GPU box under gpurun): the bench line, the launch list of a
# short bench, and one full capture each of the FP32 and FP64 step kernels at a launch where episodes are ending.
# Every profiled command is first run plain (exit 0). usage: tools/profile_steady.sh <tag>
set -u
tag=${1:-rX}
out=gpurun_out
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err || { echo "bench failed"; tail -5 $out/${tag}_bench.err; exit 1; }
echo "bench rc=0"; cat $out/${tag}_bench.json | cut -c1-400
short="--steps 30 --warmup 3 --e2e-steps 0 --e2e-warmup 3 --no-cpu-baseline --no-e2e-variants"
python bench.py $short > $out/${tag}_short.json 2> $out/${tag}_short.err || { echo "short bench failed"; tail -5 $out/${tag}_short.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv python bench.py $short > $out/${tag}_ncu1.log 2>&1
echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:f16_step_kernel --launch-skip 1200 --launch-count 1 -o $out/${tag}_step_fp32 -f \
    python bench.py --steps 1300 --warmup 3 --e2e-steps 0 --e2e-warmup 3 --no-cpu-baseline --no-e2e-variants > $out/${tag}_ncu2.log 2>&1
echo "ncu2 rc=$?"
ncu -i $out/${tag}_step_fp32.ncu-rep --page raw --csv > $out/${tag}_step_fp32_raw.csv 2>/dev/null
ncu -i $out/${tag}_step_fp32.ncu-rep --page source --csv > $out/${tag}_step_fp32_source.csv 2>/dev/null
ls -la $out/${tag}_*
