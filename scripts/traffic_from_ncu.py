"""profiles/traffic.json entry from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,... --csv` capture of
bench.py: mean DRAM bytes (read + write) per launch over the tcgen05 conv launches (conv_halo_kernel, conv_gemm*).
usage: python scripts/traffic_from_ncu.py capture.csv <config> <batch> <source-name-under-profiles/>"""
import csv
import json
import sys
from collections import defaultdict
from pathlib import Path

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    path, config, batch, source = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    lines = [ln for ln in open(path, newline="") if not ln.startswith("==")]
    per = defaultdict(dict)
    for r in csv.DictReader(lines):
        name = r["Metric Name"]
        if name.startswith("dram__bytes"):
            per[r["ID"]][name] = float(r["Metric Value"].replace(",", "")) * UNIT[r["Metric Unit"]]
            per[r["ID"]]["kernel"] = r["Kernel Name"]
    conv = [v for v in per.values() if "conv_halo_kernel" in v["kernel"] or "conv_gemm" in v["kernel"]]
    total = sum(v.get("dram__bytes_read.sum", 0.0) + v.get("dram__bytes_write.sum", 0.0) for v in conv)
    rec = {"config": config, "batch": batch, "bytes_per_launch": total / len(conv), "conv_launches": len(conv),
           "source": source}
    out = Path(__file__).resolve().parents[1] / "profiles" / "traffic.json"
    recs = json.loads(out.read_text()) if out.exists() else []
    recs = [r for r in recs if not (r["config"] == config and r["batch"] == batch)] + [rec]
    out.write_text(json.dumps(recs, indent=1) + "\n")
    print(rec)


if __name__ == "__main__":
    main()
