"""bench.py's reference arm (`--impl reference`): the CPU port of the reference's path, timed on the
host cores.  Contract checks that need no GPU: the JSON line's keys, all host threads in use even
when the launcher exports OMP_NUM_THREADS=1 (torchrun does), ranks other than 0 stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                           "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=600)


def test_reference_arm_line_and_threads():
    res = _run({"OMP_NUM_THREADS": "1"})
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "rays/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("geodesic rays/sec") and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and abs(d["e2e"]["value"] - d["value"]) < 1e-6 * d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and "sample" in cb
    assert cb["cores"] == len(os.sched_getaffinity(0))
    assert d["gpu_launches"] == 0 and "workload" in d["config"]


def test_reference_arm_other_ranks_are_silent():
    res = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0 and res.stdout.strip() == ""
