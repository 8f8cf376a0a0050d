#!/bin/bash
# Last look at HEAD: the whole GPU suite and smoke().
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -x > gpurun_out/s43_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/s43_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s43_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s43_smoke.log
