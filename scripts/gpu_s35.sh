#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_golden_configs_gpu.py -q -x -k "replication" --durations=5 > gpurun_out/s35_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/s35_pytest.log
