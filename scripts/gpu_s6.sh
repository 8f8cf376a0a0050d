#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gemm_gpu.py tests/test_unet_gpu.py tests/test_attention_gpu.py -m gpu -q -x > gpurun_out/s6_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s6_pytest.log
{
echo "== main (dynamic A ring, A256=4 B256=6 B128=7)"
timeout 200 python scripts/bench_conv.py --gn --impls 3 --shapes 11 --batch 592 --iters 10
for v in b5 a5b5; do
  echo "== variant $v"
  DDPM_LIB_VARIANT=ddpm_ood_b200/csrc/experiments/variants/lib_$v.so timeout 200 python scripts/bench_conv.py --gn --impls 3 --shapes 11 --batch 592 --iters 10
done
} > gpurun_out/s6_conv.log 2>&1
cat gpurun_out/s6_conv.log
timeout 400 python bench.py --steps 2 --warmup 3 --no_cpu_baseline > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err; echo "bench rc=$?"; cat gpurun_out/s6_bench.json | cut -c1-300
