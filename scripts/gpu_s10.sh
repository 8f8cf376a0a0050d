#!/bin/bash
# Round-2 evidence session: whole GPU suite, bench lines, launch list, DRAM traffic + tensor-pipe capture, ncu --set full of the
# halo kernels at the bench batch, racecheck.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s10_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s10_pytest.log
timeout 400 python bench.py --steps 2 --warmup 3 > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err; echo "bench rc=$?"; cut -c1-160 gpurun_out/s10_bench.json
timeout 300 python bench.py --batch 256 --steps 2 --warmup 3 --no_cpu_baseline > gpurun_out/s10_bench_b256.json 2> gpurun_out/s10_bench_b256.err; echo "b256 rc=$?"; cut -c1-160 gpurun_out/s10_bench_b256.json
timeout 300 python bench.py --config fmnist_b8 --steps 2 --warmup 3 > gpurun_out/s10_bench_b8.json 2> gpurun_out/s10_bench_b8.err; echo "b8 rc=$?"; cut -c1-160 gpurun_out/s10_bench_b8.json
for cfg in celeba64 brats_latent; do
  timeout 400 python bench.py --config $cfg --steps 1 --warmup 3 --no_cpu_baseline > gpurun_out/s10_bench_$cfg.json 2> gpurun_out/s10_bench_$cfg.err; echo "$cfg rc=$?"; cut -c1-160 gpurun_out/s10_bench_$cfg.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 300 --csv --log-file gpurun_out/s10_launches.csv python bench.py --steps 1 --warmup 1 --no_cpu_baseline --profile_every 0 > gpurun_out/s10_ncu1.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct --clock-control none -s 2000 -c 70 --csv --log-file gpurun_out/s10_traffic.csv python bench.py --steps 1 --warmup 1 --no_cpu_baseline --profile_every 0 > gpurun_out/s10_ncu2.log 2>&1; echo "ncu traffic rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_halo_kernel --launch-skip 1022 -c 8 -o gpurun_out/s10_halo_full python bench.py --steps 1 --warmup 1 --no_cpu_baseline --profile_every 0 > gpurun_out/s10_ncu3.log 2>&1; echo "ncu full rc=$?"
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/s10_racecheck.log python -m pytest tests/test_conv_gemm_gpu.py tests/test_attention_gpu.py -m gpu -q -x -k "halo3d_8vox or halo_pair_8px_256 or halo_32px or halo_conv2_plus or fused or halo_16px" > gpurun_out/s10_racecheck_pytest.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/s10_racecheck_pytest.log; tail -3 gpurun_out/s10_racecheck.log
ls -la gpurun_out/s10_halo_full.ncu-rep
