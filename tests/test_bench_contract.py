"""bench.py keeps the driver's contract: one JSON line with the agreed keys, for the CPU reference arm
(runs anywhere) and for the B200 arm (GPU box)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args, timeout=600):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert res.returncode == 0, res.stdout + res.stderr
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, res.stdout
    return json.loads(lines[0])


def test_reference_arm_line(built):
    """--impl reference: the reference's own host pipeline (oracle/_ref) timed on the host cores."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libspfft_ref.so")):
        pytest.skip("oracle/_ref/libspfft_ref.so not built on this machine")
    d = _run(["--impl", "reference", "--size", "32", "--steps", "2", "--warmup", "3"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "backward+forward pairs/sec" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("32^3 C2C double")


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [[], ["--bands", "3"], ["--type", "r2c", "--precision", "single"]],
                         ids=["single-transform", "bands", "r2c-single"])
def test_b200_arm_line(built, extra):
    d = _run(["--size", "64", "--steps", "3", "--warmup", "3", "--no-cpu-baseline"] + extra)
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["n_gpus"] == 1
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
