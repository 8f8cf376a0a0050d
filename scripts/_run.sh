cd /root/repo
(timeout 400 python -m pytest tests/test_unet_gpu.py tests/test_recon_gpu.py -q -x 2>&1 | tail -4
 for v in 1 0; do
 DDPM_ATTN_ID_RESIDUAL_MMA=$v timeout 200 python bench.py --batch 592 --steps 2 --warmup 3 --no_cpu_baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('attn_id_residual_mma=$v', d['value'], d['unet_fwd_ms'], d['roofline']['frac'], d['clocks']['sm_mhz'])"
 done
) > gpurun_out/run57.log 2>&1
