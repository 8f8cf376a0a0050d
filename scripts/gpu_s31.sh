#!/bin/bash
# K-split accumulators on the narrow tiles, now that the issue stream is short: parity, A/B against the no-split variant.
# (historical record: the K-split accumulator code this measured was removed afterwards - profiles/r02_halo_small_batch_s26.md)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gemm_gpu.py -q -x > gpurun_out/s31_conv.log 2>&1; echo "conv rc=$?"; tail -3 gpurun_out/s31_conv.log
V=ddpm_ood_b200/csrc/experiments/variants/lib_nosplit.so
for b in 8 32; do
for v in split nosplit split nosplit; do
  if [ $v = nosplit ]; then export DDPM_LIB_VARIANT=$V; else unset DDPM_LIB_VARIANT; fi
  if [ $b = 8 ]; then A="--config fmnist_b8 --steps 5"; else A="--batch $b --steps 3 --no_secondary"; fi
  timeout 300 python bench.py $A --warmup 3 --no_cpu_baseline > gpurun_out/s31_b${b}_$v.json 2> gpurun_out/s31_b${b}_$v.err
  python -c "import json;d=json.load(open('gpurun_out/s31_b${b}_$v.json'));print('$v b$b', d['value'], d['unet_fwd_ms'])"
done
done
unset DDPM_LIB_VARIANT
DDPM_HALO_CYCLES=1 DDPM_HALO_CYCLES_PRINT=1 timeout 300 python scripts/bench_conv.py --batch 8 --impls 3 --gn --iters 1 --first 8 --shapes 10 > gpurun_out/s31_timeline_b8.log 2>&1
awk '/timeline/{last=$0} /GF/{print last; print}' gpurun_out/s31_timeline_b8.log | cut -c1-330
