"""-m gpu: bench.py prints the contract line (one JSON object: metric, value, roofline, e2e through the C ABI with host
buffers, launch count = steps, clocks) and its self-check against the oracle passes."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_on_one_gpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "4", "--warmup", "3", "--no-extra", "--no-cpu-baseline"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["metric"] == "env-steps/sec" and d["unit"] == "env-steps/s" and d["n_gpus"] == 1 and d["steps"] == 4 and d["warmup"] == 3
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f32" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["value"] > 1e9 and abs(d["value"] - 4096 * 1000 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["gpu_launches"] == 4                                  # one fused launch of ours per bench step
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.05 < r["frac"] < 1.0
    assert "window_kernel" in r["kernel"] and r["bytes_per_env_step"] == pytest.approx(137.188)
    e = d["e2e"]
    assert 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] == 4096 * 1000 * 4 and e["d2h_bytes_per_step"] > 4096 * 1000 * 16
    assert d["self_check"]["env_steps_compared"] > 30000 and d["self_check"]["excluded_frac"] < 0.05
    assert d["config"]["envs_per_gpu"] == 4096 and d["config"]["rollout_steps"] == 1000 and "workload" in d["config"]
    assert d["clocks"] is None or "sm_mhz" in d["clocks"]
