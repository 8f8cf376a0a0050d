cd /root/repo
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
 timeout 600 python bench.py > gpurun_out/bench_s5.json 2> gpurun_out/bench_s5.err; tail -c 3000 gpurun_out/bench_s5.json
) > gpurun_out/run54.log 2>&1
