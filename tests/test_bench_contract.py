"""CPU: the reference arm of bench.py prints one JSON line with the contract's keys; the GPU arm
refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line(oracle):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--c1-views", "2"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    lines = [l for l in p.stdout.decode().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "fdk"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "photon_histories_per_s" and d["unit"] == "histories/s"
    assert d["value"] > 1e5 and d["vs_baseline"] is None and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "histories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["fdk"]["unit"] == "GUPS" and d["fdk"]["value"] > 1e-3
    # the thread count the oracle really ran with is what `cores` states (a launcher may export OMP_NUM_THREADS=1)
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1) == d["cpu_baseline"]["host_cores"]
    assert d["c1"]["config"]["workload"].startswith("C1 ") and d["c1"]["mc"]["value"] > 1e5 and d["mc_c4"]["value"] > 1e4


def test_reference_arm_uses_all_cores_under_a_launcher_that_sets_omp_num_threads(oracle):
    """torch.distributed.run exports OMP_NUM_THREADS=1 (the N >= 2 reference numbers collapsed 25x)"""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--c1-views", "0", "--skip-shipped"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    d = json.loads([l for l in p.stdout.decode().splitlines() if l.startswith("{")][0])
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    import bench
    assert d["config"]["workload"] == bench.WORKLOAD_C2


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120, cwd=ROOT, env=env)
    assert p.returncode == 0 and not p.stdout.strip()


def test_gpu_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300, cwd=ROOT)
    assert p.returncode != 0 and b"no CUDA device" in (p.stderr + p.stdout)
