"""The JSON line `bench.py` prints is a contract with the driver. The committed line of the final tree
(profiles/r02_bench_s39.json, produced on a B200 by scripts/gpu_s39_final.sh) must carry every key of that contract with
consistent values; and `--impl reference` must print the same shape (checked on the parser level here: running it takes
minutes of CPU)."""
import json
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _line(name):
    p = ROOT / "profiles" / name
    if not p.exists():
        pytest.skip(f"{name} not committed")
    return json.loads(p.read_text().strip().splitlines()[-1])


def test_default_line_carries_the_whole_contract():
    d = _line("r02_bench_s39.json")
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["metric"].startswith("reconstructions/sec") and d["unit"] == "reconstructions/s"
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    # value = reconstructions of K steps / device time
    recon = d["config"]["global_batch"] * d["config"]["t_starts"]
    assert abs(d["value"] - recon / (d["ms_per_step"] / 1000.0)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["h2d_bytes_per_step"] == d["config"]["global_batch"] * 32 * 32 * 4          # fp32 1x32x32 images
    assert e["d2h_bytes_per_step"] == d["config"]["t_starts"] * d["config"]["global_batch"] * 2 * 4
    assert 0.9 < e["value"] / d["value"] < 1.1                                            # measured, not copied
    assert e["value"] != d["value"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.5 < r["frac"] < 1.0
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["unit"] == d["unit"] and c["value"] > 0 and c["sample"]
    assert d["gpu_launches"] > 1000
    k = d["clocks"]
    assert k["sm_mhz"] and k["sm_max_mhz"] and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown",
                                                                       "sw_thermal_slowdown"}
    sec = d["secondary"]
    assert [s["workload"].split("batch=")[1].split(",")[0] for s in sec] == ["256", "8"]
    assert "cpu_baseline" in sec[1] and sec[1]["cpu_baseline"]["sample"].startswith("oracle fp32 loop, batch 8")


@pytest.mark.parametrize("name,n,scaling", [("r02_bench_s39_n2_weak.json", 2, "weak"),
                                            ("r02_bench_s39_n2_strong.json", 2, "strong")])
def test_multi_gpu_lines(name, n, scaling):
    d = _line(name)
    assert d["n_gpus"] == n and d["scaling"] == scaling
    assert "secondary" not in d and "cpu_baseline" not in d  # rank 0 at N = 1 only
    per_gpu = d["value"] / n
    assert 2000 < per_gpu < 3500


def test_bench_flags_the_driver_uses(monkeypatch):
    import importlib
    import sys

    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", "4", "--steps", "7", "--warmup", "3", "--impl", "reference"])
    sys.path.insert(0, str(ROOT))
    try:
        bench = importlib.import_module("bench")
        a = bench.parse()
    finally:
        sys.path.pop(0)
    assert (a.gpus, a.steps, a.warmup, a.impl, a.config, a.batch, a.skip, a.plms_state) == (4, 7, 3, "reference", "fmnist",
                                                                                            1184, 4, "carry")
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    d = bench.parse()
    assert (d.gpus, d.steps, d.warmup, d.impl) == (1, 3, 3, "ours")
