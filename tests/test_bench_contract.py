"""CPU: the one-JSON-line contract of bench.py.  (1) The committed line of the final code on a B200 (profiles/r2_bench_line.json, written
by tools/gpu/verify.sh) carries every key the contract names and its numbers are mutually consistent; (2) the reference arm
(`--impl reference`: the oracle port on the host cores, the only leg that needs no GPU) runs here and prints the same line shape."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
             "data", "config")


def test_committed_bench_line_honours_the_contract():
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_line.json")))
    for k in BASE_KEYS + ("roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches"):
        assert k in d, k
    assert d["unit"] == "shapes/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"] and d["data"] == "synthetic"
    assert "shapes/sec" in d["metric"] and "vox_res=128" in d["metric"]                                 # BASELINE.json's metric
    shapes = d["config"]["shapes_per_gpu"] * d["n_gpus"]
    assert abs(d["value"] - shapes / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1 / 3 + 1e-9          # fp16x3 executes 3x the algorithmic MMAs
    assert abs(r["achieved"] - r["algorithmic_flop_per_launch"] / (r["avg_launch_ms"] * 1e-3) / 1e12) < 1e-6 * r["achieved"]
    assert r["traffic"] is None or r["traffic"] > 3.4e7                                                # never below the algorithmic bytes
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != d["value"] and 0.5 * d["value"] < e["value"] < 1.05 * d["value"]              # its own measurement, same metric
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["sample"] and c["unit"] == d["unit"] and c["value"] > 0
    k = d["clocks"]
    assert k["sm_mhz"] <= k["sm_max_mhz"] and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["gpu_launches"] > 100 * d["steps"]
    # the rest of the measurement record travels in the same line at N = 1
    for key in ("eager_gpu_baseline", "chamfer_vs_ref", "e2e_reference_api", "shard_config4", "config1_cpu_encoder_forward", "config2_vox64",
                "config3_train_bf16_batch32", "config5_eval_256", "config5_eval_bruteforce_64"):
        assert key in d and "error" not in d[key], key
    assert all(v["bit_identical"] for v in d["chamfer_vs_ref"].values())


def test_reference_arm_runs_on_the_host_cores():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines                                                                      # ONE JSON line on stdout
    d = json.loads(lines[0])
    for k in BASE_KEYS + ("impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["steps"] == 1 and d["value"] > 0 and d["unit"] == "shapes/s"
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
