"""bench.py contract pieces that do not need a GPU: the algorithmic byte / flop counts of SURVEY §8(d), the host-core probe,
and the reference arm's JSON line (run on a tiny sample)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_survey_byte_and_flop_counts():
    assert bench.bytes_inst(12, 4, 50, 4, False, False) == 3256    # C2: identical instances, cold, fp32
    assert bench.bytes_inst(12, 4, 50, 4, True, True) == 6440      # C3: per-instance references
    assert bench.bytes_inst(6, 3, 100, 8, True, True) == 14440     # C4: fp64, per-instance references
    assert bench.bytes_inst(4, 1, 10, 8, True, True) == 856        # C1
    assert bench.flops_iter(12, 4, 50) == 64560
    assert bench.flops_iter(4, 1, 10) == 1865
    assert bench.flops_iter(16, 8, 100) == 262904


def test_host_core_probe():
    c = bench.host_cores()
    assert 1 <= c["effective"] <= c["affinity"] <= c["cpu_count"]
    assert c["cgroup_quota_cores"] is None or c["cgroup_quota_cores"] > 0


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--no-extras",
                        "--cpu-per-thread", "16"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["config"] == bench.CONFIG
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["one_thread"] > 0 and cb["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
