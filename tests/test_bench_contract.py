"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) runs the oracle port on the host and
prints one JSON line with the keys the driver reads; under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "auto-labeled objects/sec" and d["unit"] == "objects/s"
    assert d["higher_is_better"] is True and d["scaling"] == "strong" and d["data"] == "synthetic"
    assert d["steps"] == 1 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    # "reference" when the reference tree is mounted (this container), "port" (the oracle) on the GPU box
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "objects/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_falls_back_to_the_oracle_port_without_the_reference_tree():
    d = json.loads(_run({"AL3D_REFERENCE_ROOT": "/nonexistent"})[0])
    assert d["cpu_baseline"]["kind"] == "port"


def test_reference_arm_other_ranks_exit_silently():
    assert _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}) == []
