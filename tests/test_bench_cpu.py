"""CPU: the reference arm of bench.py (the oracle port timed on the host cores) prints the contract's JSON line."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--cpu-seconds", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "samples/s" and line["higher_is_better"] is True
    assert line["steps"] == 2 and line["value"] > 0 and line["vs_baseline"] is None
    assert line["metric"].startswith("audio samples/sec")
    assert line["config"]["workload"] == "cfg2"
    # both arms print the SAME config dict (the driver compares them)
    sys.path.insert(0, str(ROOT))
    import bench
    assert line["config"] == bench.base_config("cfg2", 1, 1)
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == dict(value=line["value"], unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_reference_arm_other_ranks_exit_quietly():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_host_affinity_slices_are_disjoint_and_cover():
    """Eight ranks on one socket: every rank gets its own slice of the GPU-local cores (bench.py pins with it)."""
    sys.path.insert(0, str(ROOT))
    from neural_audio_spring_reverb_b200 import hostaffinity as H
    cpus = list(range(32))
    parts = [H.slice_for_rank(cpus, r, 8, list(range(8))) for r in range(8)]
    assert all(len(p) == 4 for p in parts)
    assert sorted(c for p in parts for c in p) == cpus
    # fewer cores than ranks: still one core each
    assert all(len(H.slice_for_rank([0, 1], r, 8, list(range(8)))) == 1 for r in range(8))
    # two NUMA groups of four ranks
    assert H.slice_for_rank(list(range(16, 32)), 5, 8, [4, 5, 6, 7]) == [20, 21, 22, 23]
    info = H.pin_to_gpu(0, 1)
    assert isinstance(info, dict) and "pinned" in info
    H.unpin()
