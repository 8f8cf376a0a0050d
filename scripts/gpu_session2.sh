#!/bin/bash
# Round-2 GPU session 2: fused AttentionBlock kernel parity, whole suite, bench A/B fused vs four-launch attention.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attention_gpu.py -m gpu -q > gpurun_out/s2_attn.log 2>&1; echo "attn rc=$?"; tail -15 gpurun_out/s2_attn.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/s2_pytest.log
timeout 400 python bench.py --steps 2 --warmup 3 --no_cpu_baseline > gpurun_out/s2_bench_fused.json 2> gpurun_out/s2_bench_fused.err; echo "bench fused rc=$?"
DDPM_ATTN_FUSED=0 timeout 400 python bench.py --steps 2 --warmup 3 --no_cpu_baseline > gpurun_out/s2_bench_unfused.json 2> gpurun_out/s2_bench_unfused.err; echo "bench unfused rc=$?"
timeout 300 python bench.py --config fmnist_b8 --steps 2 --warmup 3 > gpurun_out/s2_bench_b8.json 2> gpurun_out/s2_bench_b8.err; echo "b8 rc=$?"
DDPM_CHAIN_GRAPH=0 timeout 300 python bench.py --config fmnist_b8 --steps 2 --warmup 3 --no_cpu_baseline > gpurun_out/s2_bench_b8_nograph.json 2> gpurun_out/s2_bench_b8_nograph.err
timeout 300 python bench.py --config brats_latent --steps 1 --warmup 3 > gpurun_out/s2_bench_brats.json 2> gpurun_out/s2_bench_brats.err; echo "brats rc=$?"
# launch list of one forward chain at the bench batch (shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 300 --csv --log-file gpurun_out/s2_launches.csv python bench.py --steps 1 --warmup 1 --no_cpu_baseline --profile_every 0 > gpurun_out/s2_ncu_bench.log 2>&1; echo "ncu rc=$?"
