cd /root/repo
(timeout 500 python -m pytest tests/test_conv_gemm_gpu.py tests/test_groupnorm_gpu.py tests/test_unet_gpu.py tests/test_recon_gpu.py -q -x 2>&1 | tail -4
 for v in new old new old; do
 cp scripts/_$v.so ddpm_ood_b200/csrc/libddpm_ood_b200.so
 timeout 200 python bench.py --batch 592 --steps 2 --warmup 3 --no_cpu_baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['value'], d['unet_fwd_ms'], d['roofline']['frac'], d['clocks']['sm_mhz'])"
 done
) > gpurun_out/run58.log 2>&1
