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


def test_committed_round2_lines_carry_every_config_and_the_collective():
    """profiles/r02_bench_n*.json are bench.py's own lines as the driver launches it (N = 1, 2, 4, 8): the headline
    contract plus the `workloads` block — every BASELINE config with its own roofline and kernel name, and at N > 1 the
    Monte-Carlo sum exchanged every step through the peer kernel and through NCCL, with the self-check green."""
    for n in (1, 2, 4, 8):
        j = json.loads(open(os.path.join(ROOT, "profiles", f"r02_bench_n{n}.json")).read().strip().splitlines()[-1])
        assert j["n_gpus"] == n and j["metric"] == "network evals/sec (c64, batched)" and j["config"]["workload"].startswith("cfg2")
        assert 0.9 < j["roofline"]["frac"] < 1.1 and j["roofline"]["kernel"].startswith("spn_fused_net_")
        assert j["gpu_launches"] == j["steps"]
        e = j["e2e"]
        assert e["h2d_bytes_per_step"] == 4 * 64 * (1 << 20) and e["d2h_bytes_per_step"] == 16 * (1 << 20)
        assert e["frac_of_link_ceiling"] > 0.9 and e["link_ceiling"]["value"] > 0
        w = j["workloads"]
        want = {"cfg3_point_major", "cfg3_component_major", "cfg5", "cfg5r", "mc_sum_exchange", "cfg4"}
        if n > 1:
            want |= {"cfg5_mc_sum_every_step_peer_kernel", "cfg5_mc_sum_every_step_nccl"}
        assert want <= set(w), want - set(w)
        for k in want - {"mc_sum_exchange"}:
            assert "error" not in w[k] and w[k]["value"] > 0 and w[k]["roofline"]["frac"] > 0.5, (n, k, w[k])
            assert w[k]["clocks"]["sm_mhz"] and not (set(w[k]["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"})
        assert w["cfg4"]["roofline"]["bound"] == "tensor" and w["cfg4"]["roofline"]["frac"] > 0.9
        assert w["mc_sum_exchange"]["selfcheck"]["passed"] and w["mc_sum_exchange"]["world"] == n
        if n > 1:
            ar = w["mc_sum_exchange"]["allreduce_us"]["2"]
            assert ar["peer_kernel"] < ar["nccl"]
            assert w["cfg5_mc_sum_every_step_peer_kernel"]["mc_sum"]["comm_status"] == 0
        else:
            assert j["cpu_baseline"]["value_1core"] > 0 and j["cpu_baseline"]["cores"] >= 1 and "march=native" in j["cpu_baseline"]["build"]
            assert j["sustained"]["seconds"] >= 1.0
    ref = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_reference_arm.json")).read().strip().splitlines()[-1])
    one = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_n1.json")).read().strip().splitlines()[-1])
    assert ref["config"] == one["config"]          # both arms describe the workload with one function
