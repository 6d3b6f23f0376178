#!/bin/bash
# 8-GPU (or N-GPU) A/B of the host-window NUMA placement: tools/numa_ab.sh <tag> <ngpus>
tag=$1; n=${2:-8}
out=gpurun_out
{ nproc; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; nvidia-smi topo -m; for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302" $d/class 2>/dev/null; then echo "$d numa $(cat $d/numa_node)"; fi; done; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; free -g | head -2; } > $out/${tag}_topology.txt 2>&1
args="--gpus $n --steps 200 --warmup 3 --e2e-steps 80 --e2e-warmup 350 --no-cpu-baseline --no-e2e-variants"
for numa in 0 1 2; do
  F16_HOSTWIN_NUMA=$numa python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + numa)) bench.py $args > $out/${tag}_numa${numa}.json 2> $out/${tag}_numa${numa}.err
  echo "numa=$numa rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("$out/${tag}_numa${numa}.json").read().strip().splitlines()[-1])
    print("numa=$numa value %.3e e2e %.3e node %s phases %s" % (d["value"], d["e2e"]["value"], d["e2e"].get("numa_node_rank0"), d["e2e"]["host_ms_per_step_by_phase"]))
except Exception as e:
    print("numa=$numa failed", e)
PY
done
