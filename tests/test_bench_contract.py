"""CPU checks of the bench.py contract that do not need a GPU: the reference arm (`--impl reference`) runs the oracle port of the
reference's procedure (the real itm_eval for small galleries) on the host cores and prints ONE JSON line with the agreed keys; under
torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--impl", "reference", "--gallery", "3000", "--queries", "64", "--steps", "2", "--warmup", "1"]


def _run(extra_env=None, args=SMALL):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=env, timeout=300)


def test_reference_arm_prints_one_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "recall_at_k_queries_per_sec" and d["unit"] == "queries/s"
    assert d["higher_is_better"] is True and d["scaling"] == "strong" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("large-gallery Recall@K sweep") and d["config"]["gallery"] == 3000
    cb = d["cpu_baseline"]
    # small galleries run the REAL reference's itm_eval when its tree is there (build container, or the baseline/_ref copy on the box)
    sys.path.insert(0, ROOT)
    from oracle import reference_loader as RL
    assert cb["kind"] == ("reference" if RL.reference_available() else "port")
    assert cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and "queries per step" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2"}, SMALL + ["--gpus", "2"])
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_reports_the_requested_gpu_count_in_its_config():
    d = json.loads(_run(args=SMALL + ["--gpus", "4"]).stdout.strip().splitlines()[-1])
    assert d["n_gpus"] == 4 and d["config"]["parallelism"] == "gallery-sharded x4"
