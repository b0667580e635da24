"""bench.py's reference arm runs without a GPU: check its JSON line against the driver's contract
(the GPU arm is exercised on the GPU box; its line is committed under profiles/)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=600, env=e, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip().splitlines()


def test_reference_arm_line():
    lines = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--points", "2048")
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "network evals/sec (c64, batched)" and j["unit"] == "evals/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["gpu_launches"] == 0
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"] == {"value": j["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["config"]["workload"].startswith("cfg2")


def test_reference_arm_other_ranks_print_nothing():
    # under torchrun only rank 0 runs the CPU arm; the other ranks exit 0 without work
    assert run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"}) == []


def test_committed_headline_line_keeps_the_contract():
    j = json.loads(open(os.path.join(ROOT, "profiles", "r01_bench_cfg2.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in j, k
    assert j["roofline"]["bound"] == "hbm" and 0.5 < j["roofline"]["frac"] < 1.2 and j["roofline"]["traffic"] > 2.7e8
    assert j["e2e"]["h2d_bytes_per_step"] == 4 * 64 * (1 << 20) and j["e2e"]["d2h_bytes_per_step"] == 16 * (1 << 20)
    assert j["gpu_launches"] == j["steps"] and j["vs_baseline"] is None and j["dtype"] == "f64"
