"""Turn an `ncu --csv --log-file` launch list (metrics gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum) of `bench.py --no-cpu --no-e2e --steps K --warmup W` into the committed summaries:
  profiles/<tag>_launches_step.csv  one row per launch of the LAST timed step, with its share of the step
  profiles/<tag>_traffic.json       DRAM bytes per launch per kernel (read + write), what bench.py's roofline.traffic cites
Usage: python tools/summarize_ncu.py gpurun_out/launches.csv r1"""
import csv
import io
import json
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}


def short(name):
    name = name.replace("void ", "")
    return name.split("(")[0]


def main(path, tag):
    lines = Path(path).read_text().splitlines()
    hi = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[hi:]))))
    L = OrderedDict()
    for r in rows:
        d = L.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["ms"] = v * TIME[r["Metric Unit"]]
        elif r["Metric Name"] == "dram__bytes_read.sum":
            d["rd"] = v * UNIT[r["Metric Unit"]]
        elif r["Metric Name"] == "dram__bytes_write.sum":
            d["wr"] = v * UNIT[r["Metric Unit"]]
    ids = list(L)
    start = [i for i in ids if "kp_estimate" in L[i]["name"] or "k_estimate" in L[i]["name"]][-1]
    step = [L[i] for i in ids if i >= start]
    tot = sum(d["ms"] for d in step)
    out = ROOT / "profiles" / f"{tag}_launches_step.csv"
    with out.open("w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["kernel", "grid", "block", "ms", "share_of_step", "dram_read_GB", "dram_write_GB", "dram_GBps"])
        for d in step:
            by = d.get("rd", 0) + d.get("wr", 0)
            w.writerow([short(d["name"]), d["grid"], d["block"], f'{d["ms"]:.4f}', f'{d["ms"] / tot:.4f}', f'{d.get("rd", 0) / 1e9:.4f}',
                        f'{d.get("wr", 0) / 1e9:.4f}', f'{by / 1e9 / (d["ms"] / 1e3):.0f}' if d["ms"] > 0 else ""])
        w.writerow(["TOTAL", "", "", f"{tot:.4f}", "1.0", "", "", ""])
    traffic = OrderedDict()
    for d in step:
        k = short(d["name"])                                       # largest launch of each kernel in the step (gated launches return at once)
        traffic[k] = max(traffic.get(k, 0.0), d.get("rd", 0) + d.get("wr", 0))
    (ROOT / "profiles" / f"{tag}_traffic.json").write_text(json.dumps({
        "config": "bench.py default (400^3 cells per GPU; step = E, M+B on M0; F; U on M1)",
        "source": f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none ({out.name}); "
                  "per launch, cold cache",
        "traffic_bytes_per_launch": traffic}, indent=1) + "\n")
    print(f"{len(step)} launches, {tot:.4f} ms ->", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "r1")
