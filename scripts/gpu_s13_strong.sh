#!/bin/bash
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 3 --no_cpu_baseline --shard t_starts > gpurun_out/s13_strong_n$N.json 2> gpurun_out/s13_strong_n$N.err; echo "strong rc=$?"; cut -c1-140 gpurun_out/s13_strong_n$N.json
