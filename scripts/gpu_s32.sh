#!/bin/bash
# State check after the small-batch work: whole GPU suite, smoke(), the default bench line (with its secondary lines).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s32_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s32_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s32_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s32_smoke.log
timeout 900 python bench.py > gpurun_out/s32_bench.json 2> gpurun_out/s32_bench.err; echo "bench rc=$?"; cut -c1-140 gpurun_out/s32_bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/s32_bench.json'))
print('headline', d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])
for s in d.get('secondary', []):
    print(s['workload'][-60:], s['value'], s['e2e']['value'], s.get('roofline',{}).get('frac'), s.get('cpu_baseline',{}).get('value'))
PY
