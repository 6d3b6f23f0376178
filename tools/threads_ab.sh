#!/bin/bash
# N-GPU A/B of the host-window worker-thread count: tools/threads_ab.sh <tag> <ngpus> <threads...>  ("auto" = heuristic)
tag=$1; n=$2; shift; shift
out=gpurun_out
args="--gpus $n --steps 200 --warmup 3 --e2e-steps 80 --e2e-warmup 350 --no-cpu-baseline --no-e2e-variants"
port=29600
for t in "$@"; do
  port=$((port + 1))
  if [ "$t" = auto ]; then unset F16_HOSTWIN_THREADS; else export F16_HOSTWIN_THREADS=$t; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py $args > $out/${tag}_t${t}.json 2> $out/${tag}_t${t}.err
  echo "threads=$t rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("$out/${tag}_t${t}.json").read().strip().splitlines()[-1])
    print("threads=$t value %.3e e2e %.3e phases %s" % (d["value"], d["e2e"]["value"], d["e2e"]["host_ms_per_step_by_phase"]))
except Exception as e:
    print("threads=$t failed", e)
PY
done
