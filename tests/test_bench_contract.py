"""bench.py prints exactly one JSON line on stdout with the keys the driver reads. The reference arm runs on the CPU
(here); the product arm needs a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
             "e2e", "cpu_baseline"}


def run_bench(*args, timeout=600):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry the JSON line only: %r" % lines[:3]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["scaling"] == "strong"
    assert d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None and d["data"] == "synthetic" and d["config"]["spp"] == 256
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and abs(cb["value"] - d["value"]) < 1e-9
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_product_arm_line():
    """Quick variant of the default run (64 spp instead of 256): the headline line with its blocks."""
    d = run_bench("--steps", "3", "--warmup", "3", "--no-cpu", "--spp", "64")
    assert (BASE_KEYS | {"roofline", "clocks", "gpu_launches", "raycast", "spp_per_s"}) <= set(d) and "impl" not in d
    assert d["metric"] == "Mrays/s" and d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] >= 3 and d["scaling"] == "strong"
    assert d["value"] > 500 and d["spp_per_s"] > 50 and d["gpu_launches"] > 0 and d["config"]["spp"] == 64 and d["config"]["triangles"] > 1000000
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["achieved"] > 0
    assert set(r["kernels"]) == {"raygen", "trace", "shade", "shadow", "accumulate"} and 2 < r["per_ray"]["closest_interior_nodes"] < 200
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 1920 * 1080 * 12 + 48 and e["d2h_bytes_per_step"] == 1920 * 1080 * 12 and e["steps"] == 3
    assert e["value"] < d["value"] * 1.02
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    c2 = d["raycast"]
    assert c2["metric"] == "Mrays/s" and c2["value"] > 500 and c2["gpu_launches"] == 4 * 3 and c2["config"]["triangles"] == 1002528
    assert c2["roofline"]["bound"] == "hbm" and c2["roofline"]["achieved"] > 0
    ce = c2["e2e"]
    assert ce["h2d_bytes_per_step"] == 4 * 4194304 * 24 and ce["d2h_bytes_per_step"] == 2 * 4194304 * 48 + 2 * 4194304 and ce["value"] < c2["value"] and ce["steps"] == 3
    assert set(ce["host_link"]) >= {"h2d_gbs", "d2h_gbs", "duplex_gbs_each_way"}


@pytest.mark.gpu
def test_raycast_block_checks_parity_against_the_reference_in_the_same_run():
    """--workload raycast with the CPU leg: the records the GPU produced in the timed region are compared, on the sampled
    rows, with what the reference returns for the same rays."""
    d = run_bench("--workload", "raycast", "--steps", "2", "--warmup", "3")
    p = d["parity"]
    assert p["rays"] == 4 * 262144 and p["id_mismatch"] == 0 and p["max_rel_dist"] == 0.0
    for k in ("closest_coherent", "closest_incoherent"):
        assert p[k]["hit_flag_mismatch"] == 0 and p[k]["dist_bits_mismatch"] == 0
    assert p["any_coherent"]["flag_mismatch"] == 0 and p["any_incoherent"]["flag_mismatch"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] > 0
