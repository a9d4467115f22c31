"""GPU: bench.py prints one JSON line that honours the driver's contract (keys, units, live roofline, e2e bytes)."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def run_bench(*extra):
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "2", "--warmup", "3", *extra],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def check_common(d):
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["unit"] == "program-steps/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["n_gpus"] == 1 and d["steps"] == 2
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]  # host transfers are inside the e2e region
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1.3
    assert "sm_mhz" in d["clocks"] and "reasons" in d["clocks"]


def test_iqap_bench_line():
    d = run_bench("--batch", "256", "--cpu-sample", "4")
    check_common(d)
    assert d["dtype"] == "bf16"
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    assert d["value"] > 20 * c["value"]
    assert d["pipeline"]["depth"] == 2 and d["pipeline"]["serial_ms_per_step"] > 0


def test_fa_bench_line():
    d = run_bench("--workload", "fa", "--batch", "256", "--no-cpu-baseline")
    check_common(d)
    assert d["roofline"]["kernel"] in d["kernels"]
