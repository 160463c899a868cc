"""The output contract of bench.py that the driver parses: exactly one JSON line on stdout with the agreed keys.  Checked here on
the reference arm (the only one that runs without a GPU), on a tiny CPU sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0",
                        "--cpu-nel", "4"], capture_output=True, text=True, timeout=600, env=dict(os.environ, NCCL_DEBUG="VERSION"))
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "cpu_baseline", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "GDOF/s" and d["dtype"] == "f64" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["value"] > 0 and d["n_gpus"] == 1
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] == 0
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] == "port"
    assert "workload" in d["config"] and "model" not in d["config"]


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                        "--cpu-nel", "4"], capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""
