"""CPU tier: the reference arm of bench.py (the oracle on the host cores) prints one well-formed JSON line with the
contract's keys, and non-zero ranks of a torchrun launch stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--g", "24"],
                          capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    out = _run({})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "LM iterations/s" and d["unit"] == "iterations/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    # measured at the size it names, nothing scaled
    assert d["extrapolated"] is False and d["steps_measured"] >= 1 and d["cpu_baseline"]["measured"]["bodies"] == d["steps_measured"]
    assert abs(d["ms_per_step"] * d["value"] - 1000.0) < 1e-6
    assert d["e2e"] == {"value": d["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_silently():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""
