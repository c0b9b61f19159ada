"""bench.py's CPU arm (--impl reference) prints exactly one JSON line with the contract's keys; it is the one mode that may
execute oracle/ (the reference's Cython engine compiled as-is when oracle/_ref is present, else the C port)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "movegen",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "legal_move_positions_per_sec" and d["unit"] == "positions/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["value"] > 1e4 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and "workload" in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": "positions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_selfplay_reference_arm_runs_the_unmodified_reference_and_prints_the_gpu_arms_config():
    """--impl reference for the headline workload = the reference's own parallel_self_play() CPU mode from
    baseline/_ref/training (mirror built by `make -C oracle baseline`), same `config` object as the GPU arm."""
    import pytest
    sys.path.insert(0, ROOT)
    import bench_reference
    import bench_selfplay
    if not bench_reference.available():
        pytest.skip("baseline/_ref/training not built (reference tree absent)")
    env = dict(os.environ, XQ_BENCH_SIMS="12")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mcts_sims_per_sec" and d["unit"] == "sims/s"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["workers"] == max(1, (os.cpu_count() or 1) - 1) and cb["value"] > 0
    assert "baseline/_ref/training/parallel_selfplay.py" in cb["sample"]
    assert d["config"] == bench_selfplay.workload_config(1, bench_selfplay.GAMES_PER_GPU, 12)
    assert d["e2e"]["value"] == d["value"] and d["gpu_launches"] == 0
