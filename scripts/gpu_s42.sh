#!/bin/bash
# ncu --set full of eight halo-kernel launches of one forward at the bench batch on the FINAL library: the up1 / up2 levels
# (up1.res0.conv2 .. up2.res1.conv2; 24 halo launches per forward, skip = 42 forwards + 16).
mkdir -p gpurun_out
timeout 420 ncu --set full --import-source on --clock-control none -k regex:conv_halo_kernel --launch-skip 1024 -c 8 -o gpurun_out/s42_halo_full python bench.py --steps 1 --warmup 1 --no_cpu_baseline --no_secondary --profile_every 0 > gpurun_out/s42_ncu.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/s42_halo_full.ncu-rep
