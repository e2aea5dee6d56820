"""The benchmark contract, checked without a GPU: the committed JSON lines of `bench.py` (profiles/r2_bench_n*.json, taken on
B200s by the final kernels of the round) carry every key the driver reads, their numbers are consistent with one another, and
the device arm refuses to run when there is no CUDA device (no CPU fallback on the product path)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(n):
    path = os.path.join(ROOT, "profiles", f"r2_bench_n{n}.json")
    if not os.path.exists(path):
        pytest.skip(f"{path} not committed")
    with open(path) as f:
        return json.load(f)


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_committed_bench_lines_follow_the_contract(n):
    d = _line(n)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "gpu_launches", "roofline"):
        assert key in d, key
    assert d["metric"] == "sgd_rating_updates_per_sec" and d["unit"] == "rating-updates/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == n and d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["vs_baseline"] is None  # BASELINE.md publishes no number
    assert "workload" in d["config"] and "model" not in d["config"]
    # the same matrix for every arm and every N
    assert d["config"]["matrix_crc"] == "82cab447" and d["config"]["same_matrix_on_all_ranks"] is True
    # no thermal / hardware slowdown in the timed region
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["clocks"]["sm_mhz"] >= 0.9 * d["clocks"]["sm_max_mhz"]
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["l2"]["bound"] == "l2" and abs(r["l2"]["frac"] - r["l2"]["achieved"] / r["l2"]["peak"]) < 1e-9
    # value = ratings visited / device time
    if n == 1:
        assert abs(d["value"] - d["config"]["train_nnz"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
        # algorithmic bytes per update x updates / kernel time (16 r + 12)
        assert r["algorithmic_bytes_per_update"] == 16 * d["config"]["rank"] + 12
        assert abs(r["achieved"] - r["algorithmic_bytes_per_update"] * d["value"] / 1e9) < 1e-6 * r["achieved"]
        assert r["traffic"] and r["traffic"] > 0
        cb = d["cpu_baseline"]
        assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
        e = d["e2e"]
        assert e["h2d_bytes_per_step"] > 8 * d["config"]["train_nnz"] and e["d2h_bytes_per_step"] > 0
        assert 0 < e["value"] < d["value"]  # host buffers and copies inside the timed region
    else:
        assert d["dsgd"]["plan"] == "reference"  # the reference's partitions and update sequences
        assert abs(d["value"] - d["dsgd"]["ratings_visited"] / d["steps"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
        assert len(d["dsgd"]["ratings_visited_per_rank"]) == n
        oc = d["oracle_check"]
        assert len(oc["device_val_rmse"]) == len(oc["oracle_val_rmse"]) > 0
    # BASELINE.json's second half: ALS epoch seconds at rank 64 (and the rank-128 config) at every N
    for k in ("als_rank64", "als_rank128", "ccdpp_rank64"):
        assert k in d["solvers"], k
    assert d["solvers"]["als_rank64"]["epoch_sec"] > 0
    assert d["yahoo"]["matrix_crc"] == "f4b4f1dc"  # BASELINE.json configs[4]


def test_end_to_end_step_at_several_gpus_is_in_the_lines_taken_after_it_was_added():
    for n in (4,):
        e = _line(n)["e2e"]
        assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["steps_timed"] >= 2


def test_device_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout)
    assert not p.stdout.strip().startswith("{")  # no JSON line without a measurement
