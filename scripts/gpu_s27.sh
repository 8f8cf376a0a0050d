#!/bin/bash
# K-split accumulators for the narrow small-batch tiles: parity, batch-8 / batch-32 bench, in-kernel timeline.
# (historical record: the K-split accumulator code this measured was removed afterwards - profiles/r02_halo_small_batch_s26.md)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gemm_gpu.py -q -x > gpurun_out/s27_conv.log 2>&1; echo "conv rc=$?"; tail -3 gpurun_out/s27_conv.log
for f in 0 1 2; do
  DDPM_HALO_FINE=$f timeout 300 python bench.py --config fmnist_b8 --steps 5 --warmup 3 --no_cpu_baseline > gpurun_out/s27_b8_fine$f.json 2> gpurun_out/s27_b8_fine$f.err
  python -c "import json;d=json.load(open('gpurun_out/s27_b8_fine$f.json'));print('fine$f b8', d['value'], d['unet_fwd_ms'])"
done
for b in 32 64; do
for f in 0 2; do
  DDPM_HALO_FINE=$f timeout 300 python bench.py --batch $b --steps 3 --warmup 3 --no_cpu_baseline --no_secondary > gpurun_out/s27_b${b}_fine$f.json 2> gpurun_out/s27_b${b}_fine$f.err
  python -c "import json;d=json.load(open('gpurun_out/s27_b${b}_fine$f.json'));print('fine$f b$b', d['value'], d['unet_fwd_ms'])"
done
done
DDPM_HALO_CYCLES=1 DDPM_HALO_CYCLES_PRINT=1 timeout 300 python scripts/bench_conv.py --batch 8 --impls 3 --gn --iters 1 > gpurun_out/s27_timeline_b8.log 2>&1
awk '/timeline/{last=$0} /GF/{print last; print}' gpurun_out/s27_timeline_b8.log | cut -c1-330
