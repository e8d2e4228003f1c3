"""bench.py contract checks that need no GPU: the reference arm's JSON line (the CPU port of the reference's path,
a bounded sample per step) and the command-line surface the driver uses."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ, **(env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT, env=e)


def test_reference_arm_prints_the_contract_line():
    out = _run("--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0", "--cpu-seconds", "1")
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "stream-s/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("stream-seconds of 48 kHz stereo analysed/sec")
    assert line["value"] > 0 and line["steps"] == 1 and line["gpu_launches"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "cpu" in cb and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "stream-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    out = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-seconds", "1",
               env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_refuses_to_run_without_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            return
    except Exception:
        pass
    out = _run("--steps", "1", "--no-cpu", "--no-e2e")
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_numa_binding_helper_is_harmless_without_topology():
    sys.path.insert(0, ROOT)
    import bench
    before = os.sched_getaffinity(0)
    try:
        info = bench.bind_to_gpu_numa_node(0, 1)
        assert set(info) >= {"node", "cpus", "how"} and info["cpus"] >= 1
        assert bench._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    finally:
        os.sched_setaffinity(0, before)
