#!/bin/bash
# Launch list of a forward at batch 8 (BASELINE configs[0]) on the final small-batch tilings.
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 120 --csv --log-file gpurun_out/s41_launches_b8.csv python bench.py --config fmnist_b8 --steps 1 --warmup 1 --no_cpu_baseline > gpurun_out/s41_ncu.log 2>&1; echo "rc=$?"
grep -c conv_halo gpurun_out/s41_launches_b8.csv
