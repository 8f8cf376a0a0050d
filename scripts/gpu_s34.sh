#!/bin/bash
# A/B on one box: nine-tap loop of the wide tilings unrolled with constant row offsets (variant) vs the generic loop.
# (historical record: the unrolled nine-tap variant this measured was not adopted - profiles/r02_halo_small_batch_s26.md)
mkdir -p gpurun_out
V=ddpm_ood_b200/csrc/experiments/variants/lib_unroll9.so
DDPM_LIB_VARIANT=$V timeout 600 python -m pytest tests/test_conv_gemm_gpu.py -q -x -k "halo" > gpurun_out/s34_conv.log 2>&1; echo "conv(variant) rc=$?"; tail -2 gpurun_out/s34_conv.log
for v in base unroll9 base unroll9; do
  if [ $v = unroll9 ]; then export DDPM_LIB_VARIANT=$V; else unset DDPM_LIB_VARIANT; fi
  timeout 300 python bench.py --steps 2 --warmup 2 --no_cpu_baseline --no_secondary > gpurun_out/s34_b1184_$v.json 2> gpurun_out/s34_b1184_$v.err
  python -c "import json;d=json.load(open('gpurun_out/s34_b1184_$v.json'));print('$v b1184', d['value'], d['unet_fwd_ms'], d['roofline']['frac'], d['clocks']['sm_mhz'])"
done
for v in base unroll9; do
  if [ $v = unroll9 ]; then export DDPM_LIB_VARIANT=$V; else unset DDPM_LIB_VARIANT; fi
  timeout 300 python bench.py --batch 256 --steps 3 --warmup 3 --no_cpu_baseline --no_secondary > gpurun_out/s34_b256_$v.json 2> gpurun_out/s34_b256_$v.err
  python -c "import json;d=json.load(open('gpurun_out/s34_b256_$v.json'));print('$v b256', d['value'], d['unet_fwd_ms'], d['roofline']['frac'])"
done
