#!/bin/bash
# Round-2 GPU session 1: full GPU test suite, secondary bench lines, compute-sanitizer memcheck.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s1_pytest.log
tail -5 gpurun_out/s1_pytest.log
for cfg in "fmnist_b8" "celeba64" "brats_latent" "cifar"; do
  timeout 600 python bench.py --config $cfg --steps 1 --warmup 3 > gpurun_out/s1_bench_$cfg.json 2> gpurun_out/s1_bench_$cfg.err; echo "$cfg rc=$?"
done
DDPM_CHAIN_GRAPH=0 timeout 300 python bench.py --config fmnist_b8 --steps 1 --warmup 3 --no_cpu_baseline > gpurun_out/s1_bench_fmnist_b8_nograph.json 2> gpurun_out/s1_bench_fmnist_b8_nograph.err
timeout 300 python bench.py --batch 256 --steps 2 --warmup 3 --no_cpu_baseline > gpurun_out/s1_bench_b256.json 2> gpurun_out/s1_bench_b256.err
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/s1_memcheck.log python -m pytest tests/test_conv_gemm_gpu.py tests/test_attention_gpu.py tests/test_groupnorm_gpu.py -m gpu -q -x > gpurun_out/s1_memcheck_pytest.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/s1_memcheck_pytest.log; tail -3 gpurun_out/s1_memcheck.log
