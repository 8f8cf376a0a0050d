#!/bin/bash
# Final-state evidence: whole GPU suite, smoke(), bench lines, launch list, DRAM traffic capture.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s14_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s14_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s14_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/s14_smoke.log
timeout 400 python bench.py --steps 2 --warmup 3 > gpurun_out/s14_bench.json 2> gpurun_out/s14_bench.err; echo "bench rc=$?"; cut -c1-140 gpurun_out/s14_bench.json
timeout 300 python bench.py --batch 256 --steps 2 --warmup 3 --no_cpu_baseline > gpurun_out/s14_bench_b256.json 2> gpurun_out/s14_bench_b256.err; echo "b256 rc=$?"; cut -c1-140 gpurun_out/s14_bench_b256.json
timeout 300 python bench.py --config fmnist_b8 --steps 2 --warmup 3 --no_cpu_baseline > gpurun_out/s14_bench_b8.json 2> gpurun_out/s14_bench_b8.err; echo "b8 rc=$?"; cut -c1-140 gpurun_out/s14_bench_b8.json
for cfg in celeba64 brats_latent; do
  timeout 400 python bench.py --config $cfg --steps 1 --warmup 3 --no_cpu_baseline > gpurun_out/s14_bench_$cfg.json 2> gpurun_out/s14_bench_$cfg.err; echo "$cfg rc=$?"; cut -c1-140 gpurun_out/s14_bench_$cfg.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 300 --csv --log-file gpurun_out/s14_launches.csv python bench.py --steps 1 --warmup 1 --no_cpu_baseline --profile_every 0 > gpurun_out/s14_ncu1.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct --clock-control none -s 2000 -c 70 --csv --log-file gpurun_out/s14_traffic.csv python bench.py --steps 1 --warmup 1 --no_cpu_baseline --profile_every 0 > gpurun_out/s14_ncu2.log 2>&1; echo "ncu traffic rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/s14_launches_brats.csv python bench.py --config brats_latent --steps 1 --warmup 1 --no_cpu_baseline --profile_every 0 > gpurun_out/s14_ncu3.log 2>&1; echo "ncu brats rc=$?"
