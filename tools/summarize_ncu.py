"""Turn an `ncu --csv --log-file` launch list (metrics gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum) of `bench.py --no-cpu --no-e2e --no-parity --steps K --warmup W` into the committed summaries:
  profiles/<tag>_launches_step.csv  one row per launch of the LAST timed step, with its stage and its share of the step
  profiles/<tag>_traffic.json       measured DRAM bytes (read + write) per stage, per step and per launch of each kernel:
                                    what bench.py's roofline.traffic / stages[*].dram_bytes / roofline.step_frac cite
Usage: python tools/summarize_ncu.py gpurun_out/launches.csv r2 [workload base shift_cells]"""
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


def main(path, tag, workload="S2", base=176, shift=4.0):
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
    # stages of bench.py's step: E | M+B | F (first fused map) | U | F2 (second fused map)
    stage, nmap = "E", 0
    for d in step:
        k = short(d["name"])
        if "estimate" in k:
            stage = "E"
        elif "k_map_fields_fused" in k:
            nmap += 1
            stage = "F" if nmap == 1 else "F2"
        elif stage == "E":
            stage = "M+B"
        elif stage == "F":
            stage = "U"
        d["stage"] = stage
    step = [d for d in step if not (d["stage"] == "F2" and "k_map_fields_fused" not in d["name"])]   # nothing of the step follows F2
    tot = sum(d["ms"] for d in step)
    out = ROOT / "profiles" / f"{tag}_launches_step.csv"
    with out.open("w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["stage", "kernel", "grid", "block", "ms", "share_of_step", "dram_read_GB", "dram_write_GB", "dram_GBps"])
        for d in step:
            by = d.get("rd", 0) + d.get("wr", 0)
            w.writerow([d["stage"], short(d["name"]), d["grid"], d["block"], f'{d["ms"]:.4f}', f'{d["ms"] / tot:.4f}', f'{d.get("rd", 0) / 1e9:.4f}',
                        f'{d.get("wr", 0) / 1e9:.4f}', f'{by / 1e9 / (d["ms"] / 1e3):.0f}' if d["ms"] > 0 else ""])
        w.writerow(["", "TOTAL", "", "", f"{tot:.4f}", "1.0", "", "", ""])
    per_kernel = OrderedDict()
    stage_bytes, stage_ms = OrderedDict(), OrderedDict()
    for d in step:
        k = short(d["name"])                                       # largest launch of each kernel in the step
        by = d.get("rd", 0) + d.get("wr", 0)
        per_kernel[k] = max(per_kernel.get(k, 0.0), by)
        stage_bytes[d["stage"]] = stage_bytes.get(d["stage"], 0.0) + by
        stage_ms[d["stage"]] = stage_ms.get(d["stage"], 0.0) + d["ms"]
    (ROOT / "profiles" / f"{tag}_traffic.json").write_text(json.dumps({
        "workload": workload, "base": int(base), "shift_cells": float(shift),
        "source": f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none ({out.name}); "
                  "per launch, cold cache, serialised",
        "launches_in_step": len(step), "kernel_ms_in_step_serialised": tot,
        "stage_dram_bytes": stage_bytes, "stage_kernel_ms_serialised": stage_ms, "step_dram_bytes": sum(stage_bytes.values()),
        "kernel_dram_bytes_per_launch": per_kernel}, indent=1) + "\n")
    print(f"{len(step)} launches, {tot:.4f} ms ->", out)


if __name__ == "__main__":
    a = sys.argv
    main(a[1], a[2] if len(a) > 2 else "r2", *(a[3:6]))
