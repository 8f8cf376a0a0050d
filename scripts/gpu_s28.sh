#!/bin/bash
# What paces the K loop of a narrow tile: timeline with the MMAs skipped (dbg 4), the transform skipped (dbg 1), both (5).
mkdir -p gpurun_out
for d in 0 4 1 5; do
  echo "== DDPM_HALO_DBG=$d" >> gpurun_out/s28_timeline.log
  DDPM_HALO_DBG=$d DDPM_HALO_CYCLES=1 DDPM_HALO_CYCLES_PRINT=1 timeout 300 python scripts/bench_conv.py --batch 8 --impls 3 --gn --iters 1 --first 8 --shapes 10 >> gpurun_out/s28_timeline.log 2>&1
done
awk '/== DDPM/{print} /halo cycles/{c=$0} /timeline/{last=$0} /GF/{print c; print last; print}' gpurun_out/s28_timeline.log | cut -c1-330
