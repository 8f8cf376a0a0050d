#!/bin/bash
# Where the time of a small-batch halo launch goes: per-role cycle counters at batch 8 / 256, launch list at batch 8.
mkdir -p gpurun_out
for b in 8 256; do
  echo "== batch $b" >> gpurun_out/s24_cycles.log
  DDPM_HALO_CYCLES=1 DDPM_HALO_CYCLES_PRINT=1 timeout 300 python scripts/bench_conv.py --batch $b --impls 3 --gn --iters 1 >> gpurun_out/s24_cycles.log 2>&1
done
timeout 300 python scripts/bench_conv.py --batch 8 --impls 3 --gn --iters 50 > gpurun_out/s24_conv_b8.log 2>&1
timeout 300 python scripts/bench_conv.py --batch 256 --impls 3 --gn --iters 20 > gpurun_out/s24_conv_b256.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 120 --csv --log-file gpurun_out/s24_launches_b8.csv python bench.py --batch 8 --steps 1 --warmup 1 --no_cpu_baseline > gpurun_out/s24_ncu.log 2>&1
tail -30 gpurun_out/s24_cycles.log; cat gpurun_out/s24_conv_b8.log
