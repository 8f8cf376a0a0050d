#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_recon_gpu.py -q -x > gpurun_out/s36_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/s36_pytest.log
