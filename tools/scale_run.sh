#!/bin/bash
# 1/2/4/8-GPU scaling evidence on ONE box (run under `gpurun --gpus 8`): tools/scale_run.sh <tag>
# bench.py at N = 1, 2, 4, 8 (weak scaling at 1M envs per GPU + the strong-scaling point of configs[3]: 1M envs in total),
# the host-memory write bandwidth of the box in the patterns the host-window path uses, and the box's topology.
tag=${1:-r2}
out=gpurun_out
{ nproc; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"; free -g | head -2; nvidia-smi topo -m | head -12; } > $out/${tag}_8gpu_box.txt 2>&1
nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/host_write_bw tools/host_write_bw.cu > /dev/null 2>&1 && /tmp/host_write_bw > $out/${tag}_host_write_bw_8gpu.json
cat $out/${tag}_host_write_bw_8gpu.json
for n in 1 2 4 8; do
  extra="--no-e2e-variants"; [ $n = 8 ] && extra=""
  if [ $n = 1 ]; then launcher="python"; else launcher="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n"; fi
  $launcher bench.py --gpus $n --steps 300 --warmup 5 --e2e-steps 60 --no-cpu-baseline --no-amppo $extra > $out/${tag}_scale_n$n.json 2> $out/${tag}_scale_n$n.err
  echo "N=$n rc=$?"; cut -c1-160 $out/${tag}_scale_n$n.json
done
