#!/bin/bash
# ncu evidence for profiles/: run on the GPU box (gpurun). Each profiled command is first run plain (exit 0).
# usage: tools/profile_round.sh <tag>
set -u
tag=${1:-rX}
out=gpurun_out
small="--steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 3 --e2e-warmup 3 --no-e2e-variants"
python bench.py $small > $out/${tag}_plain.json 2> $out/${tag}_plain.err || { echo "plain bench failed"; tail -5 $out/${tag}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv python bench.py $small > $out/${tag}_ncu1.log 2>&1
# device-resident step kernel (stacked layout): the 3 warm-up + 4 timed launches come first
ncu --set full --clock-control none --import-source on -k regex:f16_step_kernel -s 4 -c 1 -o $out/${tag}_step_stacked -f python bench.py $small > $out/${tag}_ncu2.log 2>&1
# frame-layout step kernel pieces of the end-to-end path (launches 8.. of the step kernel: one piece = N/4 envs)
ncu --set full --clock-control none --import-source on -k regex:f16_step_kernel -s 12 -c 1 -o $out/${tag}_step_frame -f python bench.py $small > $out/${tag}_ncu3.log 2>&1
ls -la $out/${tag}_*
