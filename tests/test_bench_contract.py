"""bench.py's contract on a box without a GPU: the reference arm prints exactly one JSON line with
the keys the driver reads, and workloads that are not a BASELINE configuration say so."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [x for x in r.stdout.splitlines() if x.strip()]
    assert len(lines) == 1, lines                      # the reference's own chatter goes to stderr
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "k-mers counted/s at k=31" and d["unit"] == "kmers/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["config"]["workload"].startswith("configs[1]")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert set(cb["stage_seconds_last_step"]) >= {"sort_thread_s", "merge_wall_s"} or cb["kind"] == "port"
    assert d["e2e"] == {"value": d["value"], "unit": "kmers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_overridden_workloads_are_labelled():
    sys.path.insert(0, ROOT)
    import bench
    argv = sys.argv
    try:
        sys.argv = ["bench.py", "--k", "96"]
        a = bench.parse_args()
        assert a.cfg["k"] == 96 and "k=96" in a.cfg["name"] and "configs[" not in a.cfg["name"]
        assert bench.metric_for(96) == "k-mers counted/s at k=96" and bench.metric_for(31) == bench.METRIC
        sys.argv = ["bench.py", "--config", "c3"]
        a = bench.parse_args()
        assert a.cfg["name"].startswith("configs[2]") and a.cfg["k"] == 63 and a.cfg["reads"] == 200_000_000
    finally:
        sys.argv = argv
