#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gemm_gpu.py -m gpu -q -x -k "halo" > gpurun_out/s8_halo.log 2>&1; echo "halo rc=$?"; tail -25 gpurun_out/s8_halo.log
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_golden_configs_gpu.py tests/test_vqvae_gpu.py -m gpu -q -x > gpurun_out/s8_unet.log 2>&1; echo "unet rc=$?"; tail -15 gpurun_out/s8_unet.log
timeout 300 python bench.py --config brats_latent --steps 1 --warmup 3 --no_cpu_baseline > gpurun_out/s8_bench_brats.json 2> gpurun_out/s8_bench_brats.err; echo "brats rc=$?"; cut -c1-200 gpurun_out/s8_bench_brats.json
