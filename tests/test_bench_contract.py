"""bench.py contract that can be checked without a GPU: the reference arm (`--impl reference`) falls back to the CPU oracle
port when the reference's CUDA build cannot run (no device here), prints ONE JSON line with the driver's keys, and under a
multi-rank launch only rank 0 prints.  The GPU arm itself is exercised on the B200 box (profiles/bench_r1_*.json)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, *args):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    env.update(extra_env or {})
    env["CUDA_VISIBLE_DEVICES"] = ""          # also on a GPU box this test is about the fallback
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C1", "--steps", "1",
                           "--warmup", "0", *args], env=env, capture_output=True, text=True, timeout=300, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "propagator steps/sec" and d["unit"] == "steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("C1")
    assert d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "steps" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_never_loads_this_library():
    """The reference arm must time the reference's library only: no module of parament_b200 (whose import dlopens
    parament_b200/lib/libparament.so) may be imported by it."""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    env["CUDA_VISIBLE_DEVICES"] = ""
    r = subprocess.run([sys.executable, "-X", "importtime", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C1",
                        "--steps", "1", "--warmup", "0", "--configs", "none"], env=env, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "parament_b200" not in r.stderr
    assert "workloads" in r.stderr        # the generator it does import lives outside the package


def test_reference_arm_other_ranks_stay_silent():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == ""
