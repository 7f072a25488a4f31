"""bench.py contract, the parts that run without a GPU: the reference arm (`--impl reference` = the oracle on the host
cores) prints ONE JSON line with the keys the driver reads, and the committed profile of the GPU arm carries the
roofline / cpu_baseline / e2e / clocks objects."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config"}


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--base", "12"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "Mcells/s" and d["higher_is_better"] is True and d["value"] > 0
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]
    # the moved front makes refinement, several 2:1 sweeps and unrefinement all fire
    assert d["config"]["n_refine"] > 0 and d["config"]["n_unrefine"] > 0 and d["config"]["sweeps_refine_gauss_seidel"] >= 2
    assert d["scaling"] == "strong"


def test_committed_gpu_profile_has_the_contract_objects():
    d = json.loads((ROOT / "profiles" / "r2_bench_n1.json").read_text().strip().splitlines()[-1])
    base = json.loads((ROOT / "BASELINE.json").read_text())
    assert BASE_KEYS <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] >= 1 and d["warmup"] >= 3 and d["dtype"] == "f64" and d["data"] == "synthetic"
    if isinstance(base.get("metric"), str):
        assert d["metric"].split(" (")[0] in base["metric"] or base["metric"].split(";")[0].split(",")[0][:12] in d["metric"]
    roof = d["roofline"]
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    assert roof["traffic"] is None or roof["traffic"] > 0
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
    assert d["gpu_launches"] > 0
    # round 2: the configuration the north star names (graded mesh, moved front) with checked decisions, bandwidth figures
    # from measured DRAM bytes only, and the topology hand-over next to the step
    assert d["scaling"] == "strong" and d["config"]["workload"].startswith("S2")
    assert d["config"]["sweeps_refine"] >= 2 and d["config"]["n_unrefine"] > 0
    assert d["parity"]["ok"] is True and d["decisions"]["timed_lists_equal_first_pass"] is True
    assert all(("GBps" not in st) or st["GBps"] <= 1.2 * roof["peak"] for st in d["stages"].values())
    assert "step_frac" not in roof or roof["step_frac"] <= 1.2
    assert d["handover"]["decisions_equal_full_hand_over"] is True and d["handover"]["incremental_bytes"] < d["handover"]["full_bytes"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
