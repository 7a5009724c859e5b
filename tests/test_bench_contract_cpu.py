"""CPU test of bench.py's reference arm: `bench.py --impl reference` must run without a GPU (it times
the restated reference algorithm, oracle/pyfft_port.c, on the host cores) and print ONE JSON line with
the keys the driver reads; under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--workload", "cfg1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    return [l for l in res.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_json_contract():
    lines = _run({"OMP_NUM_THREADS": "1"})            # torchrun exports this; the arm must not obey it
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GFLOP/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("C2C FFT GFLOP/s") and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("cfg1")
    sys.path.insert(0, ROOT)
    import bench                                       # both arms print workload_config(): the driver's same-config check
    assert d["config"] == bench.workload_config("cfg1", bench.WORKLOADS["cfg1"][1])
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["sample"]
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count()
    assert cb["cores"] == avail
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_only_rank0_prints():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
