#!/bin/bash
# compute-sanitizer memcheck over the small-batch halo instantiations added this round (<128,1>, <64,1>, tap-row stages)
# and the engine fallback-switch tests.
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --log-file gpurun_out/s40_memcheck.log python -m pytest tests/test_conv_gemm_gpu.py -m gpu -q -x -k "small_batch or upsample or halo_pair or halo_32px" > gpurun_out/s40_pytest.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/s40_memcheck.log; tail -3 gpurun_out/s40_pytest.log
