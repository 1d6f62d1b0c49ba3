"""The bench line's contract on the CPU side: the reference arm (the oracle port of GNN.py on the host cores) prints the
keys the driver reads, honours --steps / --warmup, and carries the SAME `config` object as this framework's arm; under a
multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_line_and_shared_config():
    lines = _run()
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["steps"] == 2 and line["warmup"] == 1 and line["gpu_launches"] == 0
    assert line["higher_is_better"] is True and line["unit"] == "reactions/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "reactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, ROOT)
    try:
        import bench
    finally:
        sys.path.remove(ROOT)
    cfg = dict(bench.CONFIGS["cfg2"])
    # both arms build `config` with the same function from the same arguments: the driver's same_config comparison
    assert line["config"] == bench.bench_config(bench.workload_string(cfg), cfg, cfg["batch"])
    assert line["metric"] == "reactions/sec (CGR-MPNN-3D d%d h%d fwd)" % (cfg["depth"], cfg["hidden"])


def test_reference_arm_other_ranks_print_nothing():
    assert _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}) == []
