cd /root/repo
(timeout 900 python bench.py > gpurun_out/bench_s7.json 2> gpurun_out/bench_s7.err; tail -c 400 gpurun_out/bench_s7.json; tail -3 gpurun_out/bench_s7.err
) > gpurun_out/run60.log 2>&1
