#!/usr/bin/env python
"""bench.py — network evaluations / second (c64, batched) of the hot path on N B200s.

Workload (default, N=1): BASELINE.json configs[1] — the e+e- -> mu+mu- tree spinor-chain network
evaluated over 2^20 synthetic phase-space points per GPU.  One "step" = one pass
of the network executor over one batch: ONE launch of the fused sm_100a network kernel, which reads
the four spinors of every point (256 B) and writes the amplitude (16 B).

  value         whole-job evals/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e           the same through the C ABI with HOST buffers: pinned H2D of the step's inputs,
                execute, D2H of the per-point results, all inside the timed region
  roofline      algorithmic bytes per launch / measured launch duration vs MEASURED_PEAKS.json HBM GB/s
  cpu_baseline  oracle (C++ restatement of the reference's CPU Contract/Network path) timed on the
                host cores on a bounded sample of the same workload
  --impl reference   the reference arm: the same oracle, all host threads, same config/metric/unit
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "network evals/sec (c64, batched)"
UNIT = "evals/s"
WORKLOAD_DESC = {
    "cfg2": "cfg2: e+e- -> mu+mu- tree spinor-chain network [vbar g^mu u][ubar g_mu v], Bispinor4 x Minkowski4, "
            "Complex<f64>, 2^20 synthetic phase-space points per GPU",
    "cfg3": "cfg3: SU(3) colour-factor network X.T + D.f (Dense per point x Sparse shared)",
    "cfg5": "cfg5: one-loop numerator network (gl_03 topology, massive), Monte-Carlo sum",
    "cfg5r": "cfg5r: cfg5 with the real loop momenta declared as f64 tensors (64 B per evaluation)",
    "cfg2m": "cfg2m: cfg2 with device-side leaf filling — inputs are the four massless momenta, the spinors are built "
             "inside the kernel (SURVEY 8f-1)",
    "cfg2mr": "cfg2mr: cfg2m with the momenta stored as f64 (32 B each): 144 B per evaluation",
    "cfg4": "cfg4: dense pairwise contraction of rank-6 c64 tensors, dim-32 indices, 3 shared (ZGEMM 32768^3, FP64 DMMA)",
}
DEFAULT_POINTS = {"cfg2": 1 << 20, "cfg2m": 1 << 20, "cfg2mr": 1 << 20, "cfg3": 1 << 20, "cfg5": 1 << 22, "cfg5r": 1 << 22, "cfg4": 1}
SCALE = {"cfg2": 1.0, "cfg2m": 1.0, "cfg2mr": 1.0, "cfg3": 1.0, "cfg5": 100.0, "cfg5r": 100.0}


def make_inputs(spec, wl, n, seed):
    from spenso_b200 import workloads

    if wl == "cfg2m":
        return workloads.massless_momenta(n, seed)
    if wl == "cfg2mr":
        return [np.ascontiguousarray(a.real) for a in workloads.massless_momenta(n, seed)]
    arrs = workloads.synthetic_inputs(spec, n, seed, scale=SCALE[wl], real=wl in ("cfg5", "cfg5r"))
    if wl == "cfg5r":
        arrs = [np.ascontiguousarray(a.real) for a in arrs]
    return arrs


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOAD_DESC))
    ap.add_argument("--points", type=int, default=0, help="points per GPU per step (default: the config's)")
    ap.add_argument("--layout", default="point", choices=["point", "component"])
    ap.add_argument("--exec", dest="exec_mode", default="fused", choices=["fused", "stepwise"],
                    help="stepwise = one pairwise kernel per contraction with intermediates in HBM (comparison only)")
    ap.add_argument("--mc-sum", default="final", choices=["final", "step"],
                    help="sum workloads at N>1: exchange the Monte-Carlo partial sums once after the K steps, or every step")
    ap.add_argument("--allreduce", default="peer", choices=["peer", "nccl"],
                    help="the exchange: peer = csrc/comm.cu (one kernel over NVLink peer memory), nccl = torch.distributed")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dim", type=int, default=32, help="cfg4: index dimension (32 = the BASELINE config)")
    ap.add_argument("--real", action="store_true", help="cfg4: f64 tensors (DGEMM) instead of Complex<f64> (ZGEMM)")
    return ap.parse_args()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---- CPU side: the oracle (only for cpu_baseline / --impl reference) ---------------------------------------
def oracle_rate(spec, arrays_fn, n_threads, target_seconds):
    """evals/s of the oracle port on a bounded sample; returns (rate, sample_points, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle

    probe_n = 2048 * n_threads
    arrs = arrays_fn(probe_n)
    t0 = time.perf_counter()
    _oracle.eval_spec(spec, arrs, n_threads=n_threads, want_out=False)
    dt = time.perf_counter() - t0
    rate = probe_n / dt
    n = int(min(max(rate * target_seconds, probe_n), 1 << 22))
    arrs = arrays_fn(n)
    passes, total = 0, 0.0
    while total < target_seconds and passes < 64:      # several passes over the sample until ~target_seconds of CPU work
        t0 = time.perf_counter()
        _oracle.eval_spec(spec, arrs, n_threads=n_threads, want_out=False)
        total += time.perf_counter() - t0
        passes += 1
    return n * passes / total, n * passes, total


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path.  The reference is Rust and cannot be
    built in this image (no cargo/rustc), so this is the oracle port of its Contract/Network::execute path,
    one network evaluation per point, all host threads."""
    from spenso_b200 import workloads

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "cfg4":
        print(json.dumps({"impl": "reference", "unavailable": "cfg4 (16 GiB operands) has no bounded CPU arm; "
                          "the reference arm is defined for the network workloads"}))
        return
    spec = workloads.WORKLOADS[args.workload]()
    n_threads = os.cpu_count() or 1
    real = args.workload == "cfg5"
    fn = lambda n: make_inputs(spec, args.workload, n, 1234)
    points = args.points or DEFAULT_POINTS[args.workload]
    # each step = a bounded sample of the workload; keep the whole run within a few minutes
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle

    probe = 1024 * n_threads
    arrs = fn(probe)
    t0 = time.perf_counter()
    _oracle.eval_spec(spec, arrs, n_threads=n_threads, want_out=False)
    rate = probe / (time.perf_counter() - t0)
    budget_s = 120.0
    sample = int(max(probe, min(points, rate * budget_s / max(args.steps + args.warmup, 1))))
    arrs = fn(sample)
    for _ in range(args.warmup):
        _oracle.eval_spec(spec, arrs, n_threads=n_threads, want_out=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _oracle.eval_spec(spec, arrs, n_threads=n_threads, want_out=False)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "points_per_step": sample,
                   "note": "bounded sample of the workload per step; CPU, no GPU involved"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_threads, "kind": "port",
                         "sample": f"{sample} points/step x {args.steps} steps, {n_threads} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- clocks ----------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, device_index: int):
        self.samples, self.reasons, self.stop = [], set(), False
        self.max_mhz = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def sample(self):
        """One sample now (also called from the main thread right after the timed work is enqueued, so that even a
        sub-millisecond timed region is observed while the GPU is busy)."""
        nv = self.nv
        if nv is None:
            return
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            for bit, nm in names.items():
                if r & bit:
                    self.reasons.add(nm)
        except Exception:
            pass

    def _run(self):
        while not self.stop:
            self.sample()
            time.sleep(0.001)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.nv:
            self.t.join(timeout=1.0)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def bind_to_gpu_numa_node(device_index: int):
    """Pin this rank to the CPUs next to its GPU so that its pinned host buffers are first-touched on the local
    NUMA node (N ranks share the host's memory controllers and PCIe root complexes in the e2e leg)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_words = (os.cpu_count() + 63) // 64 + 4
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


# ---- GPU arm --------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    import spenso_b200 as sp
    from spenso_b200 import workloads
    from spenso_b200.sharding import allreduce_partial_sums, connect_peers, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — spenso_b200 has no CPU execution path "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        bind_to_gpu_numa_node(local_rank)

    wl = args.workload
    spec = workloads.WORKLOADS[wl]()
    n = args.points or DEFAULT_POINTS[wl]
    layout = sp.POINT_MAJOR if args.layout == "point" else sp.COMPONENT_MAJOR
    stream = torch.cuda.current_stream()
    ctx = sp.Context(local_rank, stream=stream.cuda_stream)
    stepwise = args.exec_mode == "stepwise"
    net = workloads.build_engine(spec, ctx, sp.STEPWISE if stepwise else sp.FUSED)
    out_size = net.result_size
    leaf_slots = [x[1] for x in spec.nodes if x[0] == "batched"]
    want_sum = wl in ("cfg5", "cfg5r")        # Monte-Carlo estimate: no per-point store, only the sum over points

    # synthetic kinematics: ROT rotating input sets so that consecutive steps never re-read what L2 holds
    bytes_in = sum(sz * (16 if dt == sp.C64 else 8) for (_, sz), dt in zip(spec.inputs, spec.input_dtypes)) * n
    rot = max(2, min(8, int(np.ceil(2.5 * 126e6 / bytes_in)) + 1))
    host_sets, dev_sets = [], []
    first_point, n_mine = shard_range(n * world, rank, world)   # weak scaling: n points per rank per step
    assert n_mine == n
    for r in range(rot):
        arrs = make_inputs(spec, wl, n, 1234 + 1000 * rank + r)
        if r == 0:   # one pinned host copy feeds the e2e leg (its H2D overwrites the device inputs every step)
            host_sets.append([torch.from_numpy(a.view(np.float64).reshape(n, -1)).pin_memory() for a in arrs])
        tens = []
        for sl, a, dt in zip(leaf_slots, arrs, spec.input_dtypes):
            t = sp.BatchTensor.empty(ctx, sl, n, dt, layout)
            t.upload(a)
            tens.append(t)
        dev_sets.append(tens)
    out = None if want_sum else sp.BatchTensor.empty(ctx, net.result_slots, n, sp.C64, layout)
    sums = torch.zeros(2 * out_size, dtype=torch.float64, device=dev)
    sum_ptr = sums.data_ptr() if want_sum else 0
    # the one collective of the path (sum workloads, N > 1): Monte-Carlo partial sums over NVLink peer memory
    comm = connect_peers(ctx, max_f64=max(2 * out_size, 2)) if (want_sum and world > 1 and args.allreduce == "peer") else None
    per_step = want_sum and world > 1 and args.mc_sum == "step"
    if per_step and comm is not None:
        net.set_comm(comm)       # every step: network kernel + ONE fold/exchange kernel

    def step(i):
        net.execute(dev_sets[i % rot], out, sum_out_ptr=sum_ptr)
        if per_step and comm is None:
            allreduce_partial_sums(sums)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()

    # ---- timed region: exactly K steps ------------------------------------------------------------------------
    # The K launches are captured into ONE CUDA graph (the step is a single ~50 us kernel: eager launches from
    # Python would time the interpreter) and replayed between two events on the launching stream.  The
    # roofline uses (elapsed / K) as the kernel's launch duration: an upper bound that includes the
    # inter-kernel gap.  A second, eager pass with an event pair around every launch is reported beside it.
    K = args.steps
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    graph, cap_stream = None, torch.cuda.Stream(device=dev)
    try:
        if stepwise:
            raise RuntimeError("stepwise execution allocates intermediates: not capturable")
        ctx.set_stream(cap_stream.cuda_stream)
        cap_stream.wait_stream(stream)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=cap_stream):
            for i in range(K):
                step(i)
        graph = g
    except Exception as e:  # capture not possible: eager launches on the same stream
        print(f"bench.py: CUDA graph capture failed ({e}); timing eager launches", file=sys.stderr)
        graph = None
    ctx.set_stream(stream.cuda_stream)
    torch.cuda.synchronize()
    if graph is not None:
        graph.replay()          # one untimed replay: graph upload
        torch.cuda.synchronize()
    launches0 = ctx.launch_count
    with ClockSampler(local_rank) as clocks:
        barrier()
        t_begin.record(stream)
        if graph is not None:
            graph.replay()
        else:
            for i in range(K):
                step(i)
        if want_sum and not per_step:      # the final Monte-Carlo sum (no-op at N=1)
            if comm is not None:
                comm.allreduce_sum(sums.data_ptr(), 2 * out_size)
            else:
                allreduce_partial_sums(sums)
        t_end.record(stream)
        clocks.sample()                    # the timed work is in flight right now
        barrier()
    launches = (ctx.launch_count - launches0) if graph is None else K * (2 if want_sum else 1)
    elapsed_ms = t_begin.elapsed_time(t_end)
    kernel_ms = elapsed_ms / K
    # eager pass: event pair around every launch (includes event/launch overhead of a lone launch)
    KE2 = min(K, 50)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(KE2)]
    for i in range(KE2):
        ev[i][0].record(stream)
        step(i)
        ev[i][1].record(stream)
    torch.cuda.synchronize()
    kernel_ms_eager = float(np.median([a.elapsed_time(b) for a, b in ev]))
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = n * K * world / (elapsed_ms * 1e-3)

    # ---- e2e: host buffers through the public API, copies inside the timed region --------------------------------
    # One call per step: spn_net_execute_host (Network::execute on host tensors) — pinned host inputs in, per-point
    # results (or the Monte-Carlo sum) back in host memory; the library pipelines H2D / kernel / D2H over chunks.
    host_in = [p.numpy().view(np.complex128 if dt == sp.C64 else np.float64).reshape(n, -1)
               for p, dt in zip(host_sets[0], spec.input_dtypes)]
    host_out_t = torch.empty(n * out_size * 2, dtype=torch.float64).pin_memory()
    host_out = None if want_sum else host_out_t.numpy().view(np.complex128).reshape(n, out_size)
    h2d = sum(int(p.numel()) * 8 for p in host_sets[0])
    d2h = 16 * out_size if want_sum else n * out_size * 16
    e2e_chunk = max(n // 8, 1 << 16)   # 128 Ki points for cfg2: measured sweet spot (tools/e2e_sweep.py)

    e2e_net = net if not stepwise else workloads.build_engine(spec, ctx, sp.FUSED)   # execute_host is the fused path

    def e2e_step(i):
        e2e_net.execute_host(host_in, out=host_out, want_sum=want_sum, chunk_points=e2e_chunk)

    for i in range(3):
        e2e_step(i)
    barrier()
    KE = max(1, args.e2e_steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(KE):
        e2e_step(i)
    e1.record(stream)
    barrier()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    # the pipeline runs on the library's private streams and the call is synchronous: the host clock around the
    # K calls is the end-to-end time (the events on the launching stream bracket the same region)
    e2e_ms = max(e0.elapsed_time(e1), e2e_wall_ms)
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = n * KE * world / (e2e_ms * 1e-3)
    # what the link can do: the same pinned input buffers copied to the device with plain async copies (no kernel, no
    # D2H), best of 3 — the ceiling of any host-fed number on this box
    link_gbs = None
    try:
        dsts = [torch.empty(p.numel(), dtype=torch.float64, device=dev) for p in host_sets[0]]
        best = 1e30
        for _ in range(3):
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            c0.record(stream)
            for d, p in zip(dsts, host_sets[0]):
                d.copy_(p.view(-1), non_blocking=True)
            c1.record(stream)
            torch.cuda.synchronize()
            best = min(best, c0.elapsed_time(c1))
        link_gbs = h2d / (best * 1e-3) / 1e9
        del dsts
    except Exception:
        link_gbs = None

    # ---- roofline of the dominant kernel (the fused network kernel) ----------------------------------------------------
    peak, peak_src = measured_peaks()
    alg_bytes = spec.algorithmic_bytes * n
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tfile = os.path.join(ROOT, "profiles", f"traffic_{wl}.json")
    if os.path.exists(tfile):
        try:
            traffic = json.load(open(tfile)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": "spn_fused_net_*",
                "kernel_ms": kernel_ms, "kernel_ms_eager_event_pair": kernel_ms_eager,
                "timing": "CUDA graph of K launches" if graph is not None else "eager launches",
                "algorithmic_bytes_per_eval": spec.algorithmic_bytes,
                "fp64_instr_per_eval": net.fp64_instr_per_point}

    line = None
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:   # cpu_baseline: rank 0 at N=1 only
            nthr = os.cpu_count() or 1
            fn = lambda m: make_inputs(spec, wl, m, 1234)
            rate, sample, secs = oracle_rate(spec, fn, nthr, args.cpu_seconds)
            cpu = {"value": rate, "unit": UNIT, "cores": nthr, "kind": "port",
                   "sample": f"{sample} point evaluations of the same workload (passes over a 2^22-point sample), {nthr} threads, {secs:.1f} s "
                             "(C++ restatement of spenso's CPU Contract/Network::execute path)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[wl], "points_per_gpu_per_step": n, "layout": args.layout + "-major",
                       "result": "sum over points" if want_sum else "per-point tensor", "exec": args.exec_mode,
                       "l2": f"{rot} rotating input sets of {bytes_in / 1e6:.0f} MB (> 126 MB L2 between reuses)",
                       "parallelism": f"points sharded, {world} rank(s), no data-path collective"
                                      + (f"; Monte-Carlo sum exchanged {'every step' if per_step else 'once after the K steps'} "
                                         f"via {'NVLink peer-memory kernel (csrc/comm.cu)' if comm is not None else 'NCCL all-reduce'}"
                                         if (want_sum and world > 1) else "")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": KE, "ms_per_step": e2e_ms / KE, "api": "spn_net_execute_host (chunked H2D/kernel/D2H pipeline)",
                    "chunk_points": e2e_chunk, "h2d_gbs_in_e2e": h2d * KE / (e2e_ms * 1e-3) / 1e9,
                    "h2d_gbs_plain_copy": link_gbs},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        if comm is not None:
            net.set_comm(None)
            comm.close(dist.barrier)
        dist.destroy_process_group()
    return line


def run_cfg4(args):
    """Config 4: ONE pairwise contraction per step through spn_tensor_contract (the Contract trait): rank-6
    tensors with dim-d indices, three shared => complex GEMM with M = N = K = d^3 on the FP64 tensor pipe.
    N > 1: the first free index of A (rows of A and C) is cut into N blocks, B is replicated, no collective
     => strong scaling.  Not the headline line; reported against the cuBLAS ZGEMM throughput
    measured in the same run."""
    import torch
    import torch.distributed as dist

    import spenso_b200 as sp
    from spenso_b200.sharding import shard_range

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    d = args.dim
    n = d ** 3
    first, d1 = shard_range(d, rank, world)        # this rank's block of the first index of A
    rows = d1 * d * d
    stream = torch.cuda.current_stream()
    ctx = sp.Context(local_rank, stream=stream.cuda_stream)
    g = torch.Generator(device=dev).manual_seed(1236)
    if args.real:
        mk = lambda r: torch.rand(r, n, generator=g, device=dev, dtype=torch.float64) * 2 - 1
    else:
        mk = lambda r: torch.complex(torch.rand(r, n, generator=g, device=dev, dtype=torch.float64) * 2 - 1,
                                     torch.rand(r, n, generator=g, device=dev, dtype=torch.float64) * 2 - 1)
    edt = sp.F64 if args.real else sp.C64
    B = mk(n)                                       # same seed on every rank: B is replicated
    A = mk(max(rows, 1))
    sa = [sp.Euclidean.new_slot(d1 if i == 1 else d, i) for i in (1, 2, 3, 4, 5, 6)]
    ta = sp.BatchTensor.wrap(ctx, sa, 1, A.data_ptr(), dtype=edt, keepalive=A)
    tb = sp.BatchTensor.wrap(ctx, [sp.Euclidean.new_slot(d, i) for i in (4, 5, 6, 7, 8, 9)], 1, B.data_ptr(), dtype=edt,
                             keepalive=B)
    K, W = min(args.steps, 3), 3 if d < 32 else 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        ta.contract(tb)
    barrier()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0.record(stream)
        for _ in range(K):
            c = ta.contract(tb)
        e1.record(stream)
        barrier()
    ms = e0.elapsed_time(e1) / K
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    fpm = 2.0 if args.real else 8.0      # flops per multiply-add of the element type
    flops = fpm * n * n * n
    del c
    # denominator: cuBLAS ZGEMM / DGEMM measured here (best of 4), 8192^3
    m = 8192
    X = torch.randn(m, m, dtype=torch.float64 if args.real else torch.complex128, device=dev)
    Y = torch.randn(m, m, dtype=torch.float64 if args.real else torch.complex128, device=dev)
    best = 1e30
    for _ in range(4):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        Z = X @ Y
        a1.record()
        torch.cuda.synchronize()
        best = min(best, a0.elapsed_time(a1))
    cublas_tf = fpm * m ** 3 / (best * 1e-3) / 1e12
    achieved = flops / world / (ms * 1e-3) / 1e12      # per GPU
    if rank == 0:
        line = {
            "metric": METRIC, "value": 1e3 / ms, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC["cfg4"], "dim": d, "M=N=K": n,
                       "element": "f64" if args.real else "Complex<f64>",
                       "l2": f"operands {2 * n * n * (8 if args.real else 16) / 2**30:.0f} GiB >> 126 MB L2",
                       "parallelism": f"rows of A/C cut into {world} block(s), B replicated, no collective"},
            "gpu_launches": int(ctx.launch_count - launches0),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": cublas_tf, "unit": "TFLOP/s",
                         "frac": achieved / cublas_tf, "traffic": None, "kernel": "contract_dgemm_kernel" if args.real else "contract_zgemm_kernel",
                         "peak_source": f"cuBLAS {'DGEMM' if args.real else 'ZGEMM'} {m}^3 measured in this run (FP64 tensor pipe; tcgen05 has no f64 kind)",
                         "flops_per_eval": flops, "per_gpu": True},
            "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "cfg4":
        run_cfg4(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
