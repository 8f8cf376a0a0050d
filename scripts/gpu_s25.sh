#!/bin/bash
# 64-wide N tiles for small batches: parity on every tiling, then batch-8 A/B (DDPM_HALO_FINE=1 is the previous state).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gemm_gpu.py -q -x > gpurun_out/s25_conv.log 2>&1; echo "conv rc=$?"; tail -3 gpurun_out/s25_conv.log
for f in 1 2; do
  DDPM_HALO_FINE=$f timeout 300 python bench.py --config fmnist_b8 --steps 5 --warmup 3 --no_cpu_baseline > gpurun_out/s25_b8_fine$f.json 2> gpurun_out/s25_b8_fine$f.err
  python -c "import json;d=json.load(open('gpurun_out/s25_b8_fine$f.json'));print('fine$f b8', d['value'], d['unet_fwd_ms'])"
done
for f in 1 2; do
  DDPM_HALO_FINE=$f timeout 300 python bench.py --batch 32 --steps 3 --warmup 3 --no_cpu_baseline --no_secondary > gpurun_out/s25_b32_fine$f.json 2> gpurun_out/s25_b32_fine$f.err
  python -c "import json;d=json.load(open('gpurun_out/s25_b32_fine$f.json'));print('fine$f b32', d['value'], d['unet_fwd_ms'])"
done
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_conv_gemm_gpu.py > gpurun_out/s25_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s25_pytest.log
