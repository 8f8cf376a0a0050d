"""Benchmark of the reconstruction hot path (BASELINE.json metric: reconstructions/sec, one reconstruction = one
(image, t-start) pair carried through forward noising, the PLMS chain of UNet evaluations, clamp, LPIPS and MSE).

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
  python bench.py --impl reference --steps K --warmup W     # the reference algorithm on the host cores (oracle port)

Workload (configs[1] of BASELINE.json): FashionMNIST-shaped synthetic images 1x32x32 in [0,1), `small`
DiffusionModelUNet with random non-zero weights, scaled_linear_beta 0.0015->0.0195, 100 inference steps,
inference_skip_factor=4 -> 25 t-starts, 1250 UNet evaluations per batch. A "step" is one batch of `--batch` images x 25
t-starts. The batch size is the reference CLI's free `--batch_size` knob (default 256, reconstruct.py:89; BASELINE.json
does not fix it): the default here is 1184 = 16 images per CTA pair of the 148-SM part, which makes every UNet level a
whole number of waves and amortises each kernel's prologue over two work items at the 8-pixel level (measured: 2542
reconstructions/s, 2490 at 592, 2250 at 256); `--batch 256` reproduces the reference default.
Under torchrun every rank processes its own batch (images are what the reference shards; weak scaling);
the only collective is the gather of the [25, B, 2] score tensor.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "reconstructions/sec (image x t-start)"
UNIT = "reconstructions/s"
CHANNELS, SIZE = 1, 32
NUM_CHANNELS, ATTN, NRES = (128, 256, 256), (False, False, True), (1, 1, 1)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1184,
                    help="images per rank per step (16 per CTA pair on 148 SMs; the reference CLI default is 256)")
    ap.add_argument("--skip", type=int, default=4, help="inference_skip_factor")
    ap.add_argument("--plms_state", default="carry", choices=["carry", "reset"])
    ap.add_argument("--profile_every", type=int, default=50, help="event-profile every n-th UNet forward (0 = off)")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    return ap.parse_args()


def workload_name(args) -> str:
    return (f"FashionMNIST-shaped 1x{SIZE}x{SIZE}, small UNet, 100 steps, skip_factor={args.skip} "
            f"({len(_chains(args.skip))} t-starts), batch={args.batch}/GPU, plms_state={args.plms_state}")


def _chains(skip):
    from ddpm_ood_b200.synthetic import chain_lengths

    return chain_lengths(100, skip)


# ------------------------------------------------------------------------------------------------ CPU arm (oracle)
def _cpu_sample(budget_s: float, n_runs: int):
    """Time the oracle loop (fp32 PyTorch on all host cores) on a bounded sample of the workload.
    Returns (recon_per_s_per_run list, cores, sample description). The sample keeps the workload's mean of 50 UNet
    evaluations per reconstruction: t-starts {10,330,650,970} (skip 32) or the single t-start 490."""
    import torch

    from oracle import unet as ou
    from oracle.lpips import PerceptualLoss as OraclePL
    from oracle.recon_loop import LoopConfig, reconstruct_batch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = ou.randomize_(ou.make_small(2, CHANNELS), seed=0).eval()
    pl = OraclePL(dimensions=2, include_pixel_loss=False, is_fake_3d=False, lpips_normalize=True, spatial=False)
    B = int(os.environ.get("DDPM_REF_SAMPLE_BATCH", "8"))
    x0 = torch.rand((B, CHANNELS, SIZE, SIZE), generator=torch.Generator().manual_seed(0))
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        ts = torch.full((B,), 500, dtype=torch.long)
        model(x0, ts)  # thread-pool / allocator warm-up
        t0 = time.perf_counter()
        for _ in range(3):
            model(x0, ts)
        fwd = (time.perf_counter() - t0) / 3
    if fwd * 200 * n_runs <= budget_s:
        starts, desc = [10, 330, 650, 970], "t-starts {10,330,650,970}"
    else:
        starts, desc = [490], "t-start {490}"
    cfg = LoopConfig(inference_skip_factor=1)
    rates = []
    for _ in range(n_runs):
        noise = [torch.randn(x0.shape, generator=g) for _ in starts]
        t0 = time.perf_counter()
        reconstruct_batch(model, pl, x0, lambda i, t: noise[i], cfg, t_starts=starts)
        dt = time.perf_counter() - t0
        rates.append(B * len(starts) / dt)
    sample = (f"oracle fp32 loop, batch {B}, {desc} of the 100-step grid ({sum(c for c in _chain_for(starts))} UNet "
              f"evaluations, mean 50 per reconstruction), torch.set_num_threads({cores})")
    return rates, cores, sample


def _chain_for(starts):
    return [t // 10 + 1 for t in starts]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_runs = args.warmup + args.steps
    rates, cores, sample = _cpu_sample(budget_s=240.0, n_runs=n_runs)
    timed = rates[args.warmup:]
    value = len(timed) / sum(1.0 / r for r in timed)  # total reconstructions / total time over the K timed steps
    B = int(os.environ.get("DDPM_REF_SAMPLE_BATCH", "8"))
    per_step = B * (4 if "330" in sample else 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * per_step / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "note": "each step is a bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from ddpm_ood_b200 import _lib
    from ddpm_ood_b200.losses import PerceptualLoss
    from ddpm_ood_b200.networks import DiffusionModelUNet
    from ddpm_ood_b200.reconstruction import BatchReconstructor, ReconConfig
    from ddpm_ood_b200.synthetic import randomize_, unet_flops_per_image

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", init_method="env://", device_id=dev)
    _lib.lib()  # fail loudly if the extension is missing

    model = DiffusionModelUNet(spatial_dims=2, in_channels=CHANNELS, out_channels=CHANNELS, num_channels=NUM_CHANNELS,
                               attention_levels=ATTN, num_res_blocks=1, num_head_channels=256)
    randomize_(model, seed=0)
    model = model.to(dev).eval()
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pl = PerceptualLoss(dimensions=2, include_pixel_loss=False, is_fake_3d=False, lpips_normalize=True,
                            spatial=False, allow_synthetic_weights=True).to(dev)
    cfg = ReconConfig(beta_schedule="scaled_linear_beta", beta_start=0.0015, beta_end=0.0195,
                      plms_state=args.plms_state)
    eng = BatchReconstructor(model, pl, cfg, dev)
    B = args.batch
    chains = _chains(args.skip)
    n_t = len(chains)
    recon_per_step = B * n_t
    g = torch.Generator().manual_seed(1234 + rank)
    host_images = torch.rand((B, CHANNELS, SIZE, SIZE), generator=g).pin_memory()
    dev_images = host_images.to(dev)
    gathered = torch.empty((world, n_t, B, 2), dtype=torch.float32, device=dev) if world > 1 else None

    def step(images):
        res = eng.score_batch(images, args.skip)
        scores = torch.stack([res["perceptual_difference"], res["mse"]], dim=-1)  # [n_t, B, 2] on device
        if world > 1:
            dist.all_gather_into_tensor(gathered, scores)  # the one collective of the path (reference :238-242)
            return gathered
        return scores

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(dev_images)
    fence()

    def launches():
        return model.launch_count() + pl.perceptual_function.launch_count()

    # ---- timed region 1: inputs resident in HBM
    if args.profile_every > 0:
        model.set_profile(args.profile_every)
        model.read_profile(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    e0.record()
    for _ in range(args.steps):
        step(dev_images)
    e1.record()
    fence()
    ms = e0.elapsed_time(e1)
    n_launch = launches() - l0 + args.steps * n_t * 2  # + add_noise and clamp_mse per t-start
    prof = model.read_profile(reset=True) if args.profile_every > 0 else {}
    model.set_profile(0)
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed region 2: end to end through the public API with HOST buffers (H2D of the images and D2H of the
    # scores inside the timed region, every step)
    fence()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    host_scores = None
    for _ in range(args.steps):
        host_scores = step(host_images).cpu()  # score_batch does the (pinned, async) H2D itself
    e3.record()
    fence()
    ms_e2e = e2.elapsed_time(e3)
    assert host_scores is not None and bool(torch.isfinite(host_scores).all())

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total = recon_per_step * world * args.steps
    value = total / (ms / 1000.0)
    e2e_value = total / (ms_e2e / 1000.0)

    # ---- roofline of the dominant kernel family: the tcgen05 convolutions (conv_halo_kernel for the ResnetBlock 3x3 convs
    # with GroupNorm+SiLU applied in shared memory, conv_gemm_2cta_kernel for stride-2 / upsample convs and attention Linears)
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peaks = json.loads(peaks_path.read_text())
        peak_tf, peak_src = float(peaks["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
        peak_hbm = float(peaks["hbm_gbs"])
    else:
        peak_tf, peak_src, peak_hbm = 1400.0, "fallback (B200_PROFILING.md sustained)", 6650.0
    roofline = None
    breakdown = {}
    fwd_ms = None
    if prof and "conv_gemm" in prof:
        cg = prof["conv_gemm"]
        achieved = cg["flops"] / (cg["ms"] / 1000.0) / 1e12
        roofline = {"kernel": "conv_halo_kernel + conv_gemm_2cta_kernel (tcgen05 implicit-GEMM convolutions)",
                    "bound": "tensor", "achieved": achieved, "peak": peak_tf,
                    "unit": "TFLOP/s", "frac": achieved / peak_tf,
                    # dram__bytes_read.sum + dram__bytes_write.sum per conv_halo launch, mean of the 22 launches of one
                    # forward at batch 256 (profiles/r01_halo_ncu_full_s3.md)
                    "traffic": 83.4e6 * args.batch / 256.0, "traffic_unit": "bytes/launch (ncu at batch 256, scaled by batch; conv_halo)",
                    "peak_source": peak_src,
                    "launches_timed": cg["launches"],
                    "avg_launch_us": 1000.0 * cg["ms"] / cg["launches"],
                    "flop_per_launch_avg": cg["flops"] / cg["launches"]}
        tot = prof["_total"]
        fwd_ms = tot["ms"] / max(tot["forwards"], 1)
        for k, v in prof.items():
            if k == "_total":
                continue
            d = {"share": v["ms"] / tot["ms"], "ms_per_fwd": v["ms"] / tot["forwards"]}
            if k == "groupnorm_silu" or k == "upsample":
                d["GB/s"] = v["bytes"] / (v["ms"] / 1000.0) / 1e9
                d["frac_hbm_peak"] = d["GB/s"] / peak_hbm
            breakdown[k] = d
    flops_img = unet_flops_per_image(NUM_CHANNELS, ATTN, NRES, CHANNELS, CHANNELS, (SIZE, SIZE))
    evals = sum(chains)
    unet_tflops = flops_img * B * evals * world * args.steps / (ms / 1000.0) / 1e12

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16 operands, fp32 accumulate", "data": "synthetic",
        "config": {"workload": workload_name(args), "global_batch": B * world, "t_starts": n_t,
                   "unet_evals_per_step_per_gpu": evals, "parallelism": f"images sharded over {world} rank(s)",
                   "l2": f"per-step working set (weights 35 MB + ~{6 * B / 1024:.1f} GB activations at batch {B}) exceeds the "
                         "126 MB L2; no flush between steps"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": host_images.numel() * 4 * world,
                "d2h_bytes_per_step": n_t * B * 2 * 4 * world, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(n_launch),
        "roofline": roofline,
        "unet_fwd_ms": fwd_ms,
        "unet_tflops_whole_step": unet_tflops,
        "unet_frac_of_tensor_peak_whole_step": unet_tflops / (peak_tf * world),
        "breakdown": breakdown,
    }
    if world == 1 and not args.no_cpu_baseline:
        rates, cores, sample = _cpu_sample(budget_s=25.0, n_runs=1)
        line["cpu_baseline"] = {"value": rates[0], "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
