#!/usr/bin/env python
"""Markdown tables of the committed round-2 bench lines (profiles/r02_bench_n*.json) — pasted into profiles/README.md."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(n):
    p = os.path.join(ROOT, "profiles", f"r02_bench_n{n}.json")
    if not os.path.exists(p):
        return None
    return json.loads(open(p).read().strip().splitlines()[-1])


def fmt(x, d=3):
    return "—" if x is None else f"{x:.{d}g}"


lines = {n: load(n) for n in (1, 2, 4, 8)}
print("| quantity | " + " | ".join(f"N = {n}" for n in lines if lines[n]) + " |")
print("|---|" + "---|" * sum(1 for n in lines if lines[n]))
rows = [
    ("cfg2 `value` (evals/s, HBM-resident)", lambda j: fmt(j["value"], 4)),
    ("cfg2 roofline frac (272 B/eval vs measured copy bandwidth), per GPU", lambda j: fmt(j["roofline"]["frac"], 3)),
    ("cfg2 `e2e` (evals/s, host buffers, async pipeline)", lambda j: fmt(j["e2e"]["value"], 4)),
    ("… with one synchronous call per step", lambda j: fmt(j["e2e"]["synchronous_calls"]["value"], 4)),
    ("link ceiling: plain concurrent copies of the same bytes (evals/s)", lambda j: fmt(j["e2e"]["link_ceiling"]["value"], 4)),
    ("e2e / link ceiling", lambda j: fmt(j["e2e"]["frac_of_link_ceiling"], 3)),
    ("H2D GB/s per GPU inside e2e / plain concurrent H2D alone", lambda j: f"{j['e2e']['h2d_gbs_in_e2e']:.1f} / {j['e2e']['link_ceiling']['h2d_gbs_per_gpu_without_d2h']:.1f}"),
]
for name, fn in rows:
    print(f"| {name} | " + " | ".join(fn(lines[n]) for n in lines if lines[n]) + " |")
wl_rows = ["cfg3_point_major", "cfg3_component_major", "cfg5", "cfg5_mc_sum_every_step_peer_kernel", "cfg5_mc_sum_every_step_nccl",
           "cfg5r", "cfg4"]
for w in wl_rows:
    def cell(j):
        v = j.get("workloads", {}).get(w)
        if not v or "value" not in v:
            return "—"
        r = v["roofline"]
        extra = f", {r['achieved']:.1f} TF/s" if r["unit"] == "TFLOP/s" else ""
        clk = v["clocks"]
        cap = " (power cap)" if "sw_power_cap" in clk.get("reasons", []) else ""
        return f"{v['value']:.4g} ({r['frac']:.3g}{extra}){cap}"
    print(f"| `{w}` evals/s (roofline frac) | " + " | ".join(cell(lines[n]) for n in lines if lines[n]) + " |")


def ex(j, k):
    e = j.get("workloads", {}).get("mc_sum_exchange", {}).get("allreduce_us", {}).get(k)
    if not e:
        return "—"
    return f"{e['peer_kernel']:.1f} / " + (f"{e['nccl']:.1f}" if e.get("nccl") else "—")


for k in ("2", "32", "512"):
    print(f"| all-reduce of {k} doubles: peer kernel / NCCL (µs) | " + " | ".join(ex(lines[n], k) for n in lines if lines[n]) + " |")
print("| peer-kernel self-check (identical bits on every rank, = NCCL to 1e-15) | "
      + " | ".join(str(lines[n]["workloads"]["mc_sum_exchange"]["selfcheck"]["passed"]) for n in lines if lines[n]) + " |")
