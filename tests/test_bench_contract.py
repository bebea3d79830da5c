"""bench.py's reference arm (the contract's `--impl reference` line) on the CPU: schema, rank handling.  The GPU arm's line is
checked by the driver on the GPU box; here we make sure the CPU arm keeps printing what the driver parses."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ, TSG_BENCH_CPU_SECTORS="64", **env_extra)      # a 1/64 azimuth sector: seconds instead of half a minute
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args], capture_output=True, text=True,
                         env=env, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip()


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "oracle", "_ref")), reason="oracle/_ref (the compiled reference CPU backend) is not built")
def test_reference_arm_line():
    line = json.loads(_run({}, "--gpus", "1", "--steps", "1", "--warmup", "0").splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["vs_baseline"] is None and line["gpu_launches"] == 0
    assert line["unit"] == "scans/s" and line["value"] > 0 and "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and "sector" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2", "--steps", "1", "--warmup", "0") == ""
