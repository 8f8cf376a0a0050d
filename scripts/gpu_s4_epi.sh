#!/bin/bash
# Epilogue composition of the halo kernel: bench_conv with the timing-only debug bits (8 = no statistics, 16 = no output
# stores, 2 = no epilogue at all, 4 = no MMAs).
mkdir -p gpurun_out
for dbg in 0 8 16 24 2 4; do
  echo "== DDPM_HALO_DBG=$dbg"
  DDPM_HALO_DBG=$dbg timeout 200 python scripts/bench_conv.py --gn --impls 3 --shapes 11 --batch 592 --iters 10
done > gpurun_out/s4_epi.log 2>&1
cat gpurun_out/s4_epi.log
