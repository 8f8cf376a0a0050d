#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/s9_launches_brats.csv python bench.py --config brats_latent --steps 1 --warmup 1 --no_cpu_baseline --profile_every 0 > gpurun_out/s9_ncu_brats.log 2>&1; echo "ncu rc=$?"
