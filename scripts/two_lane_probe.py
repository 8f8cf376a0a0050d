"""Does running two half-batches on two streams (alternating one UNet + PLMS step each) beat one full batch on one
stream? Kernels of the two lanes fill each other's tails and latency-bound phases; the part is power-capped, so the gain
is whatever idle time costs in energy."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ddpm_ood_b200 import _lib  # noqa: E402
from ddpm_ood_b200.networks import DiffusionModelUNet  # noqa: E402
from ddpm_ood_b200.schedulers import PNDMScheduler  # noqa: E402


def main():
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 592
    n_steps = 40
    model = DiffusionModelUNet(spatial_dims=2, in_channels=1, out_channels=1, num_channels=(128, 256, 256),
                               attention_levels=(False, False, True), num_res_blocks=1, num_head_channels=256,
                               with_conditioning=False).cuda().eval()

    def lanes_run(n_lanes):
        b = total // n_lanes
        streams = [torch.cuda.Stream() for _ in range(n_lanes)]
        scheds, xs, bufs = [], [], []
        for _ in range(n_lanes):
            s = PNDMScheduler(num_train_timesteps=1000, skip_prk_steps=True, schedule="scaled_linear_beta",
                              beta_start=0.0015, beta_end=0.0195)
            s.set_timesteps(100)
            x = torch.randn((b, 1, 32, 32), device="cuda")
            scheds.append(s)
            xs.append(x)
            bufs.append(s._buffers(x))
        ts = [int(t) for t in scheds[0].timesteps[:n_steps]]

        def run():
            for t in ts:
                for lane in range(n_lanes):
                    with torch.cuda.stream(streams[lane]):
                        model.workspace_lane = lane
                        step = (_lib.PlmsStep * 1)()
                        step[0] = scheds[lane]._plan_step(t)
                        model.run_chain(xs[lane], [t], step, *bufs[lane])
            model.workspace_lane = 0

        run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in streams:
            st.wait_event(e0)
        run()
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"lanes={n_lanes} batch/lane={b}: {ms / n_steps:.3f} ms per step of {total} images "
              f"({total * n_steps / ms * 1e3:.0f} image-forwards/s), host {1e3 * (time.perf_counter() - t0):.0f} ms")

    for n_lanes in (1, 2, 1, 2):
        lanes_run(n_lanes)


if __name__ == "__main__":
    main()
