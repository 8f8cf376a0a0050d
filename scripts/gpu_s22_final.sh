#!/bin/bash
# End-of-round check on a fresh box: the whole GPU suite, smoke(), the default bench line.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s22_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s22_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s22_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s22_smoke.log
timeout 600 python bench.py > gpurun_out/s22_bench.json 2> gpurun_out/s22_bench.err; echo "bench rc=$?"; cut -c1-140 gpurun_out/s22_bench.json
timeout 300 python bench.py --config brats_latent --steps 1 --warmup 3 --no_cpu_baseline > gpurun_out/s22_bench_brats_latent.json 2>/dev/null; cut -c1-140 gpurun_out/s22_bench_brats_latent.json
timeout 300 python bench.py --config celeba64 --steps 1 --warmup 3 --no_cpu_baseline > gpurun_out/s22_bench_celeba64.json 2>/dev/null; cut -c1-140 gpurun_out/s22_bench_celeba64.json
