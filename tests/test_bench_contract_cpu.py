"""The parts of bench.py's contract that can be checked without a GPU: the reference arm (the CPU port of the
reference's algorithm on a bounded sample of the headline workload) prints ONE JSON line with the keys the driver
reads, and the interval-union helper that turns overlapping launch intervals into kernel time behaves."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-rows", "200"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "BCA macro-F1@5 instances/sec per sweep"
    assert d["unit"] == "instances/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]


def test_union_of_overlapping_launch_intervals():
    sys.path.insert(0, ROOT)
    import bench
    # two launches overlapping by half + one disjoint: the union, not the sum
    assert abs(bench.union_ms([0.0, 0.5, 2.0], [1.0, 1.5, 2.5]) - 2.0) < 1e-12
    assert bench.union_ms([], []) == 0.0
