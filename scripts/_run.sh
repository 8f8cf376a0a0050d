cd /root/repo
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
 timeout 600 python bench.py > gpurun_out/bench_s6.json 2> gpurun_out/bench_s6.err; tail -c 600 gpurun_out/bench_s6.json
 timeout 600 python bench.py --batch 1184 --no_cpu_baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batch1184', d['value'], d['e2e']['value'], d['unet_fwd_ms'], d['roofline']['frac'], d['unet_frac_of_tensor_peak_whole_step'], d['clocks'])"
 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 300 --csv --log-file gpurun_out/launches_s7.csv python bench.py --steps 1 --warmup 1 --skip 64 --profile_every 0 --no_cpu_baseline --batch 256 > /dev/null 2>&1
 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 150 --csv --log-file gpurun_out/launches_s7_b592.csv python bench.py --steps 1 --warmup 1 --skip 64 --profile_every 0 --no_cpu_baseline --batch 592 > /dev/null 2>&1
 ls -la gpurun_out/launches_s7*
) > gpurun_out/run59.log 2>&1
