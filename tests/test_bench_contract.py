"""bench.py's reference arm runs without a GPU: its JSON line must carry the contract keys the driver reads
(metric / value / unit / n_gpus / steps / warmup / ms_per_step / higher_is_better / scaling / vs_baseline / dtype /
data / config.workload / impl / cpu_baseline / e2e).  The repo arm's line is produced on the GPU box only; its
`config` has to be identical to the reference arm's (the driver compares them), which `bench.base_config` builds
for both arms (checked here: the reference arm's `config` equals `base_config(cfg)`)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*flags):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *flags],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


@pytest.mark.parametrize("cfg", ["cfg1", "f1"])
def test_reference_arm_line_has_the_contract_keys(cfg):
    line = _run("--config", cfg, "--steps", "2", "--warmup", "1")
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1
    assert line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert isinstance(line["config"].get("workload"), str) and "model" not in line["config"]
    if cfg == "cfg1":
        sys.path.insert(0, ROOT)
        import bench
        assert line["config"] == json.loads(json.dumps(bench.base_config(cfg)))
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == line["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert line.get("gpu_launches", 0) == 0
