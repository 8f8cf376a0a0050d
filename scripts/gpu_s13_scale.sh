#!/bin/bash
# Multi-GPU evidence: weak scaling (images sharded, the driver's SCALE mode) and strong scaling (t-start grid of ONE global
# batch sharded, plms_state=reset) at N GPUs. usage: gpu_s13_scale.sh N
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 --no_cpu_baseline > gpurun_out/s13_weak_n$N.json 2> gpurun_out/s13_weak_n$N.err; echo "weak rc=$?"; cut -c1-140 gpurun_out/s13_weak_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 3 --no_cpu_baseline --shard t_starts > gpurun_out/s13_strong_n$N.json 2> gpurun_out/s13_strong_n$N.err; echo "strong rc=$?"; cut -c1-140 gpurun_out/s13_strong_n$N.json
if [ "$N" = "2" ]; then
  python bench.py --steps 2 --warmup 3 --no_cpu_baseline --plms_state reset > gpurun_out/s13_reset_n1.json 2> gpurun_out/s13_reset_n1.err; echo "n1 reset rc=$?"; cut -c1-140 gpurun_out/s13_reset_n1.json
fi
