#!/bin/bash
# Wall-clock timeline of one halo launch at batch 8 (stamps inside the kernel), per layer shape.
mkdir -p gpurun_out
DDPM_HALO_CYCLES=1 DDPM_HALO_CYCLES_PRINT=1 timeout 300 python scripts/bench_conv.py --batch 8 --impls 3 --gn --iters 1 > gpurun_out/s26_timeline_b8.log 2>&1
grep -B0 -A0 "timeline\|GF" gpurun_out/s26_timeline_b8.log | awk '/timeline/{last=$0} /GF/{print last; print}'
