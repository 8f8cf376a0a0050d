#!/bin/bash
# End-of-round check on a fresh box (2 GPUs): whole GPU suite, smoke(), the default bench line (N = 1, with its secondary
# lines), and the driver's N = 2 launch form (weak scaling over images; strong scaling over t-starts).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s39_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s39_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s39_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s39_smoke.log
timeout 900 python bench.py > gpurun_out/s39_bench.json 2> gpurun_out/s39_bench.err; echo "bench rc=$?"; cut -c1-140 gpurun_out/s39_bench.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/s39_bench_n2.json 2> gpurun_out/s39_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-140 gpurun_out/s39_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 2 --warmup 3 --shard t_starts > gpurun_out/s39_bench_n2_strong.json 2> gpurun_out/s39_bench_n2_strong.err; echo "bench n2 strong rc=$?"; cut -c1-140 gpurun_out/s39_bench_n2_strong.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/s39_bench.json'))
print('headline', d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])
for s in d.get('secondary', []):
    print(s['workload'][-60:], s['value'], s['e2e']['value'], s.get('roofline',{}).get('frac'), s.get('cpu_baseline',{}).get('value'))
PY
