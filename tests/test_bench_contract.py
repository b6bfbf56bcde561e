"""The JSON line `bench.py` prints is a contract with the driver (task statement, "Measurement"): these CPU tests pin its shape
on the lines committed under profiles/ (written by real B200 runs of the same script) and the parts of bench.py that need no GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        rows = [json.loads(l) for l in f if l.startswith("{")]
    assert rows, name
    return rows[-1]


@pytest.mark.parametrize("name", ["r02h_bench_dambreak2d_1m.json", "r02h_bench_n2.json", "r02g_bench_n8.json"])
def test_our_arm_line_has_every_contract_key(name):
    d = _line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in d, k
    assert d["metric"] == "particle-steps/sec" and d["unit"] == "particle-steps/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None      # BASELINE.md holds no published number
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["scaling"] in ("weak", "strong")
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] < d["value"]                      # host copies inside the timed region: never faster than the device-timed value
    assert d["gpu_launches"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"] and isinstance(c["reasons"], list)
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # value is the whole-job aggregate: particles x steps / time
    assert abs(d["value"] - d["config"]["particles"] * 1e3 / d["ms_per_step"]) <= 1e-6 * d["value"]
    if d["n_gpus"] == 1:
        b = d["cpu_baseline"]
        assert b["kind"] in ("reference", "port") and b["cores"] >= 1 and b["value"] > 0 and b["unit"] == d["unit"] and b["sample"]
        assert r["traffic"] and r["dram_frac"] and 0 < r["dram_frac"] < 1.2
    large = d["large"]
    assert large["workload"] == "dambreak3d_10m" and large["scaling"] == "strong" and large["ms_per_step"] > 0


def test_reference_arm_line():
    d = _line("r02h_bench_reference_arm.json")
    ours = _line("r02h_bench_dambreak2d_1m.json")
    assert d["impl"] == "reference"
    assert d["metric"] == ours["metric"] and d["unit"] == ours["unit"] and d["higher_is_better"] == ours["higher_is_better"]
    assert d["config"] == ours["config"], "both arms must print the same config for the same command line"
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    b = d["cpu_baseline"]
    assert b["kind"] == "reference" and b["cores"] >= 1 and b["value"] == d["value"] and "Boost" in b["sample"]


def test_default_workloads_and_openmp_policy():
    import bench
    assert bench.default_workload(1) == "dambreak2d_1m"            # BASELINE.json configs[1]
    assert [bench.default_workload(n) for n in (2, 4, 8)] == ["dambreak2d_2m", "dambreak2d_4m", "dambreak2d_8m"]
    assert bench.LARGE_WORKLOAD == "dambreak3d_10m"
    for name in bench.WORKLOADS:
        assert callable(bench.WORKLOADS[name][0]) and bench.WORKLOADS[name][1]
    # several ranks: OpenMP settings of the launcher are left alone (test_multi_gpu.py has the full story)
    env = {"WORLD_SIZE": "4"}
    is_ref, world = bench.configure_openmp(["bench.py", "--gpus", "4"], env)
    assert not is_ref and world == 4 and "OMP_PROC_BIND" not in env


def test_help_runs_without_a_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "--impl" in r.stdout and "--gpus" in r.stdout
