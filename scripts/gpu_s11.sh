#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gemm_gpu.py tests/test_unet_gpu.py -m gpu -q -x > gpurun_out/s11_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s11_pytest.log
{
echo "== main (two epilogue warp groups, 640 threads)"
timeout 200 python scripts/bench_conv.py --gn --impls 3 --shapes 11 --batch 592 --iters 10
echo "== variant epi1 (one group, 512 threads)"
DDPM_LIB_VARIANT=ddpm_ood_b200/csrc/experiments/variants/lib_epi1.so timeout 200 python scripts/bench_conv.py --gn --impls 3 --shapes 11 --batch 592 --iters 10
} > gpurun_out/s11_conv.log 2>&1
cat gpurun_out/s11_conv.log
timeout 400 python bench.py --steps 2 --warmup 3 --no_cpu_baseline > gpurun_out/s11_bench.json 2> gpurun_out/s11_bench.err; echo "bench rc=$?"; cut -c1-120 gpurun_out/s11_bench.json
DDPM_LIB_VARIANT=ddpm_ood_b200/csrc/experiments/variants/lib_epi1.so timeout 400 python bench.py --steps 2 --warmup 3 --no_cpu_baseline > gpurun_out/s11_bench_epi1.json 2> gpurun_out/s11_bench_epi1.err; echo "bench epi1 rc=$?"; cut -c1-120 gpurun_out/s11_bench_epi1.json
