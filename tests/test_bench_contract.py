"""The committed bench lines (profiles/bench_r1c_*.json, written by bench.py on a B200) carry every key the measurement
contract names; guards bench.py's output format without needing a GPU."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = sorted(glob.glob(os.path.join(ROOT, "profiles", "bench_r1c_*.json")) +
               glob.glob(os.path.join(ROOT, "profiles", "bench_r2_n*.json")))


@pytest.mark.parametrize("path", LINES, ids=[os.path.basename(p) for p in LINES])
def test_committed_bench_line_has_the_contract_keys(path):
    d = json.loads(open(path).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert k in d, k
    assert d["metric"] == "denoising steps/sec" and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["scaling"] == ("weak" if d["n_gpus"] == 1 else "strong")
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.02
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1.0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if d["n_gpus"] == 1 and "cpu_baseline" in d:
        c = d["cpu_baseline"]
        assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    if "bench_r2_" in path:
        # round 2: the line proves what it timed -- the benched model against the oracle, the sharded forward against
        # the un-sharded one, the hoisted fractions
        p = d["parity"]
        assert p["ok"] and p["rel_l2"] < p["tolerance"]["rel_l2"] and p["pearson"] > p["tolerance"]["pearson"]
        for k in ("gemm_frac", "attention_frac", "step_frac"):
            assert 0 < d[k] < 1.0, k
        assert d["flops_per_step"]["executed"] <= d["flops_per_step"]["as_reference"]
        if d["n_gpus"] > 1:
            assert d["cp_parity"]["ok"] and d["cp_parity"]["bit_exact_pinned_kernels"] is True
        if "vae" in d:
            assert d["vae_frames_per_s"] == d["vae"]["value"]
            if d["n_gpus"] > 1:
                assert d["vae"]["shard_parity"]["bit_exact"] is True
            else:
                assert d["vae"]["parity"]["ok"]
        if "fp8" in d:
            assert d["fp8"]["parity"]["ok"] and d["fp8"]["value"] > d["value"] and "fp8" in d["fp8"]["dtype"]


def test_there_are_committed_bench_lines():
    assert LINES, "profiles/bench_r1c_*.json missing"
