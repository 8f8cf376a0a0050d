#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gemm_gpu.py tests/test_unet_gpu.py -m gpu -q -x > gpurun_out/s5_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s5_pytest.log
{
echo "== main (256-bit stores, addend prefetch, shared scale/shift)"
timeout 200 python scripts/bench_conv.py --gn --impls 3 --shapes 11 --batch 592 --iters 10
for v in a5b5 a4b6; do
  echo "== variant $v"
  DDPM_LIB_VARIANT=ddpm_ood_b200/csrc/experiments/variants/lib_$v.so timeout 200 python scripts/bench_conv.py --gn --impls 3 --shapes 11 --first 4 --batch 592 --iters 10
done
} > gpurun_out/s5_conv.log 2>&1
cat gpurun_out/s5_conv.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/s5_launches_brats.csv python bench.py --config brats_latent --steps 1 --warmup 1 --no_cpu_baseline --profile_every 0 > gpurun_out/s5_ncu_brats.log 2>&1; echo "ncu rc=$?"
