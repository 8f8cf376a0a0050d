#!/bin/bash
# Engine fallback switches under test (read at plan time now); graph replay == per-kernel launches.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_vqvae_gpu.py tests/test_attention_gpu.py -q -x > gpurun_out/s33_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/s33_pytest.log
