#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gemm_gpu.py -m gpu -q -x -k halo > gpurun_out/s12_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s12_pytest.log
{
echo "== main (two epilogue warp groups, 640 threads, setmaxnreg 40 / 120)"
timeout 200 python scripts/bench_conv.py --gn --impls 3 --shapes 11 --batch 592 --iters 10
} > gpurun_out/s12_conv.log 2>&1
cat gpurun_out/s12_conv.log
timeout 400 python bench.py --steps 2 --warmup 3 --no_cpu_baseline > gpurun_out/s12_bench.json 2> gpurun_out/s12_bench.err; echo "bench rc=$?"; cut -c1-120 gpurun_out/s12_bench.json
