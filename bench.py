#!/usr/bin/env python
"""bench.py — network evaluations / second (c64, batched) of the hot path on N B200s.

Headline workload (N=1 and weak-scaled at N>1): BASELINE.json configs[1] — the e+e- -> mu+mu- tree spinor-chain
network evaluated over 2^20 synthetic phase-space points per GPU.  One "step" = one pass of
the network executor over one batch: ONE launch of the fused sm_100a network kernel, which reads the four spinors of
every point (256 B) and writes the amplitude (16 B).

  value         whole-job evals/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e           the same through the C ABI with HOST buffers: pinned H2D of the step's inputs, execute, D2H of the
                per-point results, all inside the timed region; beside it the ceiling the host link gives when all
                ranks copy the same bytes at the same time
  roofline      algorithmic bytes per launch / measured launch duration vs MEASURED_PEAKS.json HBM GB/s
  cpu_baseline  oracle (C++ restatement of the reference's CPU Contract/Network path) timed on the host cores on a
                bounded sample of the same workload: all cores and one core
  workloads     the other BASELINE configs measured in the same run, each with its own roofline and clocks:
                cfg3 (both layouts, 2^22 points), cfg5 / cfg5r (1e8 points over the N GPUs, Monte-Carlo sum; at N>1
                the sum exchanged every step through the NVLink peer-memory kernel and through NCCL), cfg4 (one
                rank-6 dim-32 contraction = ZGEMM 32768^3 on the FP64 tensor pipe), the exchange kernel alone, and a
                bit-identity self-check of it against NCCL
  --impl reference   the reference arm: the same oracle, all host threads, same config/metric/unit
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
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
    "cfg3": "cfg3: SU(3) colour-factor network X.T + D.f (Dense per point x Sparse shared), 2^22 points",
    "cfg5": "cfg5 (massive variant): one-loop numerator network of the gl_03 topology with the second fermion propagator "
            "numerator made massive (the literal gl_03 numerator is identically zero), Monte-Carlo sum over 1e8 points",
    "cfg5r": "cfg5r: cfg5 (massive variant) with the real loop momenta declared as f64 tensors (64 B per evaluation)",
    "cfg2m": "cfg2m: cfg2 with device-side leaf filling — inputs are the four massless momenta, the spinors are built "
             "inside the kernel (SURVEY 8f-1)",
    "cfg2mr": "cfg2mr: cfg2m with the momenta stored as f64 (32 B each): 144 B per evaluation",
    "cfg4": "cfg4: dense pairwise contraction of rank-6 c64 tensors, dim-32 indices, 3 shared (ZGEMM 32768^3, FP64 DMMA)",
}
# points per GPU per step when a workload is the headline (--workload X)
DEFAULT_POINTS = {"cfg2": 1 << 20, "cfg2m": 1 << 20, "cfg2mr": 1 << 20, "cfg3": 1 << 20, "cfg5": 1 << 22, "cfg5r": 1 << 22, "cfg4": 1}
SCALE = {"cfg2": 1.0, "cfg2m": 1.0, "cfg2mr": 1.0, "cfg3": 1.0, "cfg5": 100.0, "cfg5r": 100.0}
SUM_WORKLOADS = ("cfg5", "cfg5r")
FP64_PIPE_INSTR_PER_S = 148 * 64 * 1.965e9      # 148 SMs x 64 DFMA/clk x 1.965 GHz (nominal, non-tensor FP64 pipe)


def make_inputs(spec, wl, n, seed):
    from spenso_b200 import workloads

    if wl == "cfg2m":
        return workloads.massless_momenta(n, seed)
    if wl == "cfg2mr":
        return [np.ascontiguousarray(a.real) for a in workloads.massless_momenta(n, seed)]
    arrs = workloads.synthetic_inputs(spec, n, seed, scale=SCALE[wl], real=wl in SUM_WORKLOADS)
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
    ap.add_argument("--suite", default=None, choices=["full", "headline"],
                    help="full (default for the default workload): the headline plus the `workloads` block")
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
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sustained-seconds", type=float, default=1.0, help="length of the sustained re-measurement of the headline")
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


def headline_config(args, wl, n, world, rot=None, bytes_in=None, want_sum=None, per_step=False, comm_kind=None):
    """The `config` object — built by ONE function for both arms so that they carry identical keys and values."""
    from spenso_b200 import workloads

    spec = workloads.WORKLOADS[wl]()
    if bytes_in is None:
        import spenso_b200 as sp
        bytes_in = sum(sz * (16 if dt == sp.C64 else 8) for (_, sz), dt in zip(spec.inputs, spec.input_dtypes)) * n
    if rot is None:
        rot = rotation_sets(bytes_in)
    if want_sum is None:
        want_sum = wl in SUM_WORKLOADS
    par = f"points sharded, {world} rank(s), no data-path collective"
    if want_sum and world > 1:
        par += (f"; Monte-Carlo sum exchanged {'every step' if per_step else 'once after the K steps'} via "
                f"{'NVLink peer-memory kernel (csrc/comm.cu)' if comm_kind == 'peer' else 'NCCL all-reduce'}")
    return {"workload": WORKLOAD_DESC[wl], "points_per_gpu_per_step": n, "layout": args.layout + "-major",
            "result": "sum over points" if want_sum else "per-point tensor", "exec": args.exec_mode,
            "l2": f"{rot} rotating input sets of {bytes_in / 1e6:.0f} MB (> 126 MB L2 between reuses)",
            "parallelism": par}


def rotation_sets(bytes_in):
    return max(2, min(8, int(np.ceil(2.5 * 126e6 / bytes_in)) + 1)) if bytes_in < 2.5 * 126e6 else 1


# ---- CPU side: the oracle (only for cpu_baseline / --impl reference) ---------------------------------------
def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def native_oracle():
    """Build the oracle for THIS host (-O3 -march=native) and make tests/_oracle.py load it.
    The committed liboracle.so is built without -march because it travels from the build container to the GPU box.
    Returns the flags string recorded in the JSON line."""
    flags = "-O3 -ffp-contract=off (portable build)"
    try:
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "native"], timeout=240,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        path = os.path.join(ROOT, "oracle", "_native", "liboracle.so")
        if os.path.exists(path):
            os.environ["SPN_ORACLE_LIB"] = path
            flags = "-O3 -march=native -ffp-contract=off (built on this host)"
    except Exception:
        pass
    return flags


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
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    if args.workload == "cfg4":
        print(json.dumps({"impl": "reference", "unavailable": "cfg4 (16 GiB operands) has no bounded CPU arm; "
                          "the reference arm is defined for the network workloads"}))
        return
    build = native_oracle()
    spec = workloads.WORKLOADS[args.workload]()
    n_threads = os.cpu_count() or 1
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
        "config": headline_config(args, args.workload, points, max(world, args.gpus), per_step=args.mc_sum == "step",
                                  comm_kind=args.allreduce),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_threads, "kind": "port", "cpu_model": cpu_model(),
                         "build": build,
                         "sample": f"{sample} points/step x {args.steps} steps, {n_threads} threads (bounded sample of the "
                                   "workload per step; CPU only, no GPU involved; C++ restatement of spenso's CPU "
                                   "Contract/Network::execute path)"},
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
class Job:
    """Process-wide plumbing of one rank: device, stream, context, process group."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        import spenso_b200 as sp

        self.torch, self.dist, self.sp = torch, dist, sp
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — spenso_b200 has no CPU execution path "
                             "(use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device(f"cuda:{self.local_rank}")
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            bind_to_gpu_numa_node(self.local_rank)
        self.stream = torch.cuda.current_stream()
        self.ctx = sp.Context(self.local_rank, stream=self.stream.cuda_stream)
        self._comm = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def comm(self, max_f64=512):
        """The NVLink peer-memory communicator of this job (created once)."""
        if self._comm is None:
            from spenso_b200.sharding import connect_peers

            self._comm = connect_peers(self.ctx, max_f64=max_f64)
        return self._comm

    def free_gib(self):
        """free device memory, the minimum over the ranks (every rank must take the same decisions)"""
        free, _ = self.torch.cuda.mem_get_info(self.dev)
        return -self.max_over_ranks(-free / 2 ** 30)

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            if self._comm is not None:
                self._comm.close(self.dist.barrier)
            self.dist.destroy_process_group()
        elif self._comm is not None:
            self._comm.close()


def device_inputs(job, spec, wl, n, layout, seed):
    """Synthetic kinematics generated ON the device (workloads whose inputs never need a host copy): i.i.d. U(-1,1)
    * scale per real component, imaginary parts zero for the real-momentum workloads.  Returns BatchTensors that wrap
    torch storage."""
    torch, sp = job.torch, job.sp
    g = torch.Generator(device=job.dev).manual_seed(seed)
    leaf_slots = [x[1] for x in spec.nodes if x[0] == "batched"]
    tens = []
    for sl, (_, size), dt in zip(leaf_slots, spec.inputs, spec.input_dtypes):
        w = 2 if dt == sp.C64 else 1
        shape = (n, size, w) if layout == sp.POINT_MAJOR else (size, n, w)
        x = torch.empty(shape, dtype=torch.float64, device=job.dev)
        # chunked fill: torch.rand of the whole tensor would need a second full-size temporary
        flat = x.view(-1)
        step = 1 << 28
        for lo in range(0, flat.numel(), step):
            seg = flat[lo:lo + step]
            seg.uniform_(-SCALE[wl], SCALE[wl], generator=g)
        if wl in SUM_WORKLOADS and dt == sp.C64:
            x[..., 1] = 0.0
        tens.append(sp.BatchTensor.wrap(job.ctx, sl, n, x.data_ptr(), dtype=dt, layout=layout, keepalive=x))
    return tens


def kernel_name_of(net, key):
    try:
        m = re.search(r"\b(spn_fused_net_[0-9a-f]{8}_[a-z]+)\s*\(", net.prepare_variant(key))
        return m.group(1) if m else None
    except Exception:
        return None


def traffic_of(wl, layout_name, kernel):
    """dram bytes per launch from the committed ncu --set full capture — only when it was taken on the kernel that is
    launched now (the file records the kernel name)."""
    for name in (f"traffic_{wl}_{layout_name}.json", f"traffic_{wl}.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            try:
                d = json.load(open(path))
            except Exception:
                continue
            if kernel and d.get("kernel") == kernel:
                return d.get("dram_bytes_per_launch"), {"traffic_source": d.get("source"), "traffic_points": d.get("points")}
            return None, {"traffic_note": f"profiles/{name} was captured on {d.get('kernel')}, not on the kernel launched now"}
    return None, {}


def measure_network(job, args, wl, n, layout_name="point", K=20, W=3, exec_mode="fused", mc_sum="final", allreduce="peer",
                    on_device_inputs=True, e2e=False, cpu=False, sustained_s=0.0, total_points_note=None, min_ms=0.0):
    """One workload on this job: returns the entry (rank 0) with value / roofline / clocks (+ e2e, cpu_baseline)."""
    torch, dist, sp = job.torch, job.dist, job.sp
    from spenso_b200 import workloads
    from spenso_b200.sharding import allreduce_partial_sums

    ctx, dev, stream, world, rank = job.ctx, job.dev, job.stream, job.world, job.rank
    spec = workloads.WORKLOADS[wl]()
    layout = sp.POINT_MAJOR if layout_name == "point" else sp.COMPONENT_MAJOR
    stepwise = exec_mode == "stepwise"
    net = workloads.build_engine(spec, ctx, sp.STEPWISE if stepwise else sp.FUSED)
    out_size = net.result_size
    leaf_slots = [x[1] for x in spec.nodes if x[0] == "batched"]
    want_sum = wl in SUM_WORKLOADS        # Monte-Carlo estimate: no per-point store, only the sum over points

    # synthetic kinematics: ROT rotating input sets so that consecutive steps never re-read what L2 holds
    bytes_in = sum(sz * (16 if dt == sp.C64 else 8) for (_, sz), dt in zip(spec.inputs, spec.input_dtypes)) * n
    rot = rotation_sets(bytes_in)
    host_sets, dev_sets = [], []
    for r in range(rot):
        if on_device_inputs and not e2e:
            dev_sets.append(device_inputs(job, spec, wl, n, layout, 1234 + 1000 * rank + r))
            continue
        arrs = make_inputs(spec, wl, n, 1234 + 1000 * rank + r)
        if r == 0:   # one pinned host copy feeds the e2e leg (its H2D overwrites nothing the device-resident leg reads)
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
    comm = job.comm() if (want_sum and world > 1 and allreduce == "peer") else None
    per_step = want_sum and world > 1 and mc_sum == "step"
    if per_step and comm is not None:
        net.set_comm(comm)       # every step: network kernel + ONE fold/exchange kernel

    def step(i):
        net.execute(dev_sets[i % rot], out, sum_out_ptr=sum_ptr)
        if per_step and comm is None:
            allreduce_partial_sums(sums)

    W = max(W, 3)
    l0 = ctx.launch_count
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step(0)
    w0.record(stream)
    for i in range(1, W):
        step(i)
    w1.record(stream)
    launches_per_step = (ctx.launch_count - l0) / W
    job.barrier()
    if min_ms > 0:
        # short steps (a few hundred microseconds at N = 8): enough of them that the skew between the ranks' start times
        # (a fraction of a millisecond after a barrier) does not weigh on the timed region; every rank takes the same K
        step_ms = job.max_over_ranks(w0.elapsed_time(w1) / (W - 1))
        K = int(min(400, max(K, np.ceil(min_ms / max(step_ms, 1e-3)))))

    # ---- timed region: exactly K steps ------------------------------------------------------------------------
    # The K launches are captured into ONE CUDA graph (a step is a single 50 us .. few ms kernel: eager launches from
    # Python would time the interpreter) and replayed between two events on the launching stream.  The roofline uses
    # (elapsed / K) as the kernel's launch duration: an upper bound that includes the inter-kernel gap.
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

    def timed_region():
        with ClockSampler(job.local_rank) as clk:
            job.barrier()
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
            clk.sample()                       # the timed work is in flight right now
            job.barrier()
        return t_begin.elapsed_time(t_end), clk

    elapsed_ms_local, clocks = timed_region()
    launches = (ctx.launch_count - launches0) if graph is None else int(round(K * launches_per_step))
    elapsed_ms = job.max_over_ranks(elapsed_ms_local)

    # eager pass: event pair around every launch (includes event/launch overhead of a lone launch)
    KE2 = min(K, 50)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(KE2)]
    for i in range(KE2):
        ev[i][0].record(stream)
        step(i)
        ev[i][1].record(stream)
    torch.cuda.synchronize()
    kernel_ms_eager = float(np.median([a.elapsed_time(b) for a, b in ev]))

    # One re-measurement, decided identically on every rank, when the timed region cannot be trusted: a slowdown reason
    # other than the power cap was sampled, or the K graph-replayed steps took more than 1.3x K times the per-launch
    # event time just measured on the slowest rank (seen once at N = 4: a 30 ms region lasting 95 ms).  Both
    # measurements are reported.
    remeasured = None
    if graph is not None:
        bad_clock = any(r in clocks.reasons for r in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"))
        eager_max = job.max_over_ranks(kernel_ms_eager)
        inconsistent = elapsed_ms / K > 1.3 * eager_max + 0.02
        if job.max_over_ranks(1.0 if (bad_clock or inconsistent) else 0.0) > 0:
            first = {"ms_per_step": elapsed_ms / K, "clocks": clocks.summary(),
                     "why": "slowdown reason sampled" if bad_clock else
                            f"timed region {elapsed_ms / K / max(eager_max, 1e-9):.2f}x the per-launch event time"}
            elapsed_ms_local, clocks = timed_region()
            elapsed_ms = job.max_over_ranks(elapsed_ms_local)
            remeasured = {"rejected_first_measurement": first}
    kernel_ms = elapsed_ms_local / K
    value = n * K * world / (elapsed_ms * 1e-3)
    comm_status = comm.status if comm is not None else None
    sum_value = [float(x) for x in sums[:2].tolist()] if want_sum else None

    # sustained: the same graph replayed back to back for >= sustained_s seconds (power/clock steady state)
    sustained = None
    if sustained_s > 0 and graph is not None:
        reps = max(1, int(np.ceil(1.2 * sustained_s * 1e3 / max(elapsed_ms_local, 1e-3))))   # margin: at least sustained_s
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(job.local_rank) as sclocks:
            job.barrier()
            s0.record(stream)
            for _ in range(reps):
                graph.replay()
            s1.record(stream)
            sclocks.sample()
            job.barrier()
        s_ms = job.max_over_ranks(s0.elapsed_time(s1))
        sustained = {"seconds": s_ms * 1e-3, "steps": reps * K, "ms_per_step": s_ms / (reps * K),
                     "value": n * K * reps * world / (s_ms * 1e-3), "clocks": sclocks.summary()}

    # ---- e2e: host buffers through the public API, copies inside the timed region --------------------------------
    e2e_obj = None
    if e2e and not args.no_e2e:
        e2e_obj = measure_e2e(job, args, spec, net if not stepwise else workloads.build_engine(spec, ctx, sp.FUSED),
                              host_sets[0], n, out_size, want_sum)

    # ---- roofline of the dominant kernel (the fused network kernel) ----------------------------------------------------
    peak, peak_src = measured_peaks()
    alg_bytes = spec.algorithmic_bytes * n
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    lay = "p" if layout == sp.POINT_MAJOR else "c"
    key = lay * len(spec.inputs) + ("n" if want_sum else lay) + ("s" if want_sum else "")
    kname = kernel_name_of(net, key) if not stepwise else None
    traffic, tnote = traffic_of(wl, layout_name, kname)
    if traffic is not None and tnote.get("traffic_points") and tnote["traffic_points"] != n:
        traffic = traffic * n / tnote["traffic_points"]     # per launch of THIS size (traffic is linear in the points)
        tnote["traffic_note"] = f"scaled from a capture at {tnote['traffic_points']} points per launch"
    fp = (net.variant_fp64_instr(key) or net.fp64_instr_per_point) if not stepwise else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": kname or "pairwise kernels (stepwise)",
                "kernel_ms": kernel_ms, "kernel_ms_eager_event_pair": kernel_ms_eager,
                "timing": "CUDA graph of K launches" if graph is not None else "eager launches",
                "algorithmic_bytes_per_eval": spec.algorithmic_bytes,
                "bytes_read_per_eval": net.input_bytes_per_point if not stepwise else None,
                "fp64_instr_per_eval": fp}
    roofline.update(tnote)
    if not stepwise and net.input_bytes_per_point != spec.algorithmic_bytes - (0 if want_sum else 16 * out_size):
        # components that only meet structural zeros of the shared tensors are never read (cfg3)
        touched = net.input_bytes_per_point + (0 if want_sum else 16 * out_size)
        roofline["achieved_bytes_touched"] = touched * n / (kernel_ms * 1e-3) / 1e9
        roofline["frac_bytes_touched"] = roofline["achieved_bytes_touched"] / peak
    if fp:
        roofline["fp64_pipe_frac"] = fp * n / (kernel_ms * 1e-3) / FP64_PIPE_INSTR_PER_S

    if per_step and comm is not None:
        net.set_comm(None)
    entry = None
    if rank == 0:
        cfg = headline_config(args, wl, n, world, rot, bytes_in, want_sum, per_step, "peer" if comm is not None else "nccl")
        cfg["layout"] = layout_name + "-major"
        cfg["exec"] = exec_mode
        if total_points_note:
            cfg["points"] = total_points_note
        entry = {"value": value, "unit": UNIT, "steps": K, "warmup": W, "ms_per_step": elapsed_ms / K, "config": cfg,
                 "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks.summary()}
        if sustained:
            entry["sustained"] = sustained
        if remeasured:
            entry["remeasured"] = remeasured
        if e2e_obj:
            entry["e2e"] = e2e_obj
        if want_sum:
            entry["mc_sum"] = {"re_im": sum_value, "comm_status": comm_status}
        if cpu and not args.no_cpu_baseline and world == 1:   # cpu_baseline: rank 0 at N=1 only
            build = native_oracle()
            nthr = os.cpu_count() or 1
            fn = lambda m: make_inputs(spec, wl, m, 1234)
            rate, sample, secs = oracle_rate(spec, fn, nthr, args.cpu_seconds)
            rate1, sample1, secs1 = oracle_rate(spec, fn, 1, min(args.cpu_seconds / 3, 4.0))
            entry["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": nthr, "kind": "port", "cpu_model": cpu_model(), "build": build,
                "value_1core": rate1,
                "sample": f"{sample} point evaluations of the same workload (passes over a <= 2^22-point sample), {nthr} threads, "
                          f"{secs:.1f} s; 1 core: {sample1} evaluations, {secs1:.1f} s (C++ restatement of spenso's CPU "
                          "Contract/Network::execute path)"}
    del dev_sets, out
    return entry


def measure_e2e(job, args, spec, net, host_set, n, out_size, want_sum):
    """Network::execute on HOST data through spn_net_execute_host, copies inside the timed region — and, beside it,
    what the host link gives when every rank moves the same bytes with plain copies at the same time."""
    torch, sp = job.torch, job.sp
    dev, stream, world = job.dev, job.stream, job.world
    host_in = [p.numpy().view(np.complex128 if dt == sp.C64 else np.float64).reshape(n, -1)
               for p, dt in zip(host_set, spec.input_dtypes)]
    host_out_t = torch.empty(n * out_size * 2, dtype=torch.float64).pin_memory()
    host_out = None if want_sum else host_out_t.numpy().view(np.complex128).reshape(n, out_size)
    h2d = sum(int(p.numel()) * 8 for p in host_set)
    d2h = 16 * out_size if want_sum else n * out_size * 16
    e2e_chunk = int(os.environ.get("SPN_E2E_CHUNK", "0")) or max(n // 8, 1 << 16)

    sum_arr = np.zeros(out_size, dtype=np.complex128) if want_sum else None

    def e2e_step(i):
        # one call per step; the calls are asynchronous and the library keeps its H2D / kernel / D2H pipeline full
        # across them (a Monte-Carlo driver submits batch after batch); the results are awaited once, after the K steps
        net.execute_host_async(host_in, out=host_out, sum_out=sum_arr, chunk_points=e2e_chunk)

    for i in range(3):
        e2e_step(i)
    net.host_wait()
    job.barrier()
    KE = max(1, args.e2e_steps)
    t0 = time.perf_counter()
    for i in range(KE):
        e2e_step(i)
    net.host_wait()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    job.barrier()
    # the pipeline runs on the library's private streams: the host clock from the first submission to the completion of
    # the last result copy is the end-to-end time
    e2e_ms = job.max_over_ranks(e2e_wall_ms)
    e2e_value = n * KE * world / (e2e_ms * 1e-3)
    # the same with one synchronous call per step (spn_net_execute_host: the pipeline drains at every call)
    t0 = time.perf_counter()
    for i in range(KE):
        net.execute_host(host_in, out=host_out, want_sum=want_sum, chunk_points=e2e_chunk)
    sync_ms = job.max_over_ranks((time.perf_counter() - t0) * 1e3)
    job.barrier()

    # The ceiling of ANY host-fed number on this box at this N: every rank copies the same pinned input buffers to its
    # device (and the same number of result bytes back, on a second stream) with plain async copies — no kernel — all
    # ranks at the same time, best of 3, max over ranks.
    link_ms, link_h2d_alone = None, None
    try:
        dsts = [torch.empty(p.numel(), dtype=torch.float64, device=dev) for p in host_set]
        back = torch.empty(max(d2h // 8, 2), dtype=torch.float64, device=dev)
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        best, best_h = 1e30, 1e30
        for both in (True, True, True, False, False, False):
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            job.barrier()
            c0.record(stream)
            s_in.wait_stream(stream)
            s_out.wait_stream(stream)
            REPS = 4          # back to back, like the steps of the e2e leg
            with torch.cuda.stream(s_in):
                for _ in range(REPS):
                    for d, p in zip(dsts, host_set):
                        d.copy_(p.view(-1), non_blocking=True)
            if both and not want_sum:
                with torch.cuda.stream(s_out):
                    for _ in range(REPS):
                        host_out_t[:back.numel()].copy_(back, non_blocking=True)
            stream.wait_stream(s_in)
            stream.wait_stream(s_out)
            c1.record(stream)
            torch.cuda.synchronize()
            ms = job.max_over_ranks(c0.elapsed_time(c1)) / REPS
            if both:
                best = min(best, ms)
            else:
                best_h = min(best_h, ms)
        link_ms, link_h2d_alone = best, best_h
        del dsts, back
    except Exception as e:  # noqa: BLE001
        print(f"bench.py: link ceiling measurement failed ({e})", file=sys.stderr)
    ceiling = n * world / (link_ms * 1e-3) if link_ms else None
    return {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "steps": KE, "ms_per_step": e2e_ms / KE,
            "api": "spn_net_execute_host_async per step + spn_net_host_wait after the K steps (chunked H2D/kernel/D2H "
                   "pipeline that stays full across calls)",
            "synchronous_calls": {"value": n * KE * world / (sync_ms * 1e-3), "ms_per_step": sync_ms / KE,
                                  "api": "spn_net_execute_host (drains at every call)"},
            "chunk_points": e2e_chunk, "h2d_gbs_in_e2e": h2d * KE / (e2e_ms * 1e-3) / 1e9,
            "link_ceiling": {"value": ceiling, "unit": UNIT, "ms_per_step": link_ms,
                             "h2d_gbs_per_gpu": (h2d / (link_ms * 1e-3) / 1e9) if link_ms else None,
                             "h2d_gbs_per_gpu_without_d2h": (h2d / (link_h2d_alone * 1e-3) / 1e9) if link_h2d_alone else None,
                             "how": f"plain pinned cudaMemcpyAsync of the same {h2d} B in + {d2h} B out per step on every one of the "
                                    f"{world} rank(s) at the same time, 4 steps back to back, no kernel; best of 3, max over ranks"},
            "frac_of_link_ceiling": (e2e_value / ceiling) if ceiling else None}


def measure_exchange(job):
    """The collective alone: all-reduce (SUM) of n doubles through mc_sum_exchange_kernel and through NCCL, launched
    back to back inside one CUDA graph; plus a bit-identity self-check of the peer kernel (every rank must hold the
    same bits; against NCCL to 1e-15, bit-equal at N = 2)."""
    torch, dist = job.torch, job.dist
    dev, world, rank, ctx = job.dev, job.world, job.rank, job.ctx
    comm = job.comm()
    stream = torch.cuda.Stream(device=dev)

    def timed_graph(body, iters, reps=5):
        ctx.set_stream(stream.cuda_stream)
        try:
            with torch.cuda.stream(stream):
                for _ in range(3):
                    body()
            job.barrier()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                for _ in range(iters):
                    body()
            job.barrier()
            with torch.cuda.stream(stream):
                g.replay()
            job.barrier()
            ts = []
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                job.barrier()
                with torch.cuda.stream(stream):
                    e0.record(stream)
                    g.replay()
                    e1.record(stream)
                job.barrier()
                ts.append(e0.elapsed_time(e1) * 1e3 / iters)
            return job.max_over_ranks(float(np.median(ts)))
        finally:
            ctx.set_stream(job.stream.cuda_stream)

    res = {"world": world, "kernel": "mc_sum_exchange_kernel (csrc/comm.cu)", "allreduce_us": {}}
    launches0 = ctx.launch_count
    for n in (2, 32, 512):
        x = torch.ones(n, dtype=torch.float64, device=dev) * (rank + 1) * 1e-3
        peer = timed_graph(lambda: comm.allreduce_sum(x.data_ptr(), n), 200)
        nccl = None
        if world > 1:
            x.fill_((rank + 1) * 1e-3)

            def nccl_body():
                dist.all_reduce(x)
                x.mul_(0.5)       # keeps the values bounded; one tiny elementwise launch rides along

            def mul_body():
                x.mul_(0.5)

            try:
                nccl = timed_graph(nccl_body, 200) - timed_graph(mul_body, 200)
            except Exception as e:  # noqa: BLE001
                print(f"bench.py: NCCL all-reduce under graph capture failed ({e})", file=sys.stderr)
        res["allreduce_us"][str(n)] = {"peer_kernel": peer, "nccl": nccl}
    # self-check (the body of tests/test_peer_comm.py at THIS world size)
    ok, n_ex = True, 0
    for it in range(30):
        n = (2, 32, 512)[it % 3]
        g = torch.Generator(device=dev).manual_seed(100 * it + rank)
        x = torch.rand(n, generator=g, device=dev, dtype=torch.float64) - 0.5
        y = x.clone()
        comm.allreduce_sum(x.data_ptr(), n)
        torch.cuda.synchronize()
        n_ex += 1
        if world > 1:
            dist.all_reduce(y)
            gathered = [torch.empty_like(x) for _ in range(world)]
            dist.all_gather(gathered, x)
            ok &= all(torch.equal(gathered[0], t) for t in gathered)          # identical bits on every rank
            ok &= bool(torch.allclose(x, y, rtol=1e-15, atol=1e-15))
            if world == 2:
                ok &= bool(torch.equal(x, y))
        else:
            ok &= bool(torch.equal(x, y))
    flag = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res["selfcheck"] = {"passed": bool(flag.item() == 1.0), "exchanges": n_ex, "comm_status": comm.status,
                        "what": "identical bits on every rank; equal to NCCL all-reduce to 1e-15 (bit-equal at N = 2)"}
    res["gpu_launches"] = int(ctx.launch_count - launches0)
    return res


def measure_cfg4(job, args, K=1, W=1, standalone=False):
    """Config 4: ONE pairwise contraction per step through spn_tensor_contract (the Contract trait): rank-6
    tensors with dim-d indices, three shared => complex GEMM with M = N = K = d^3 on the FP64 tensor pipe.
    N > 1: the first free index of A (rows of A and C) is cut into N blocks, B is replicated, no collective
     => strong scaling.  Reported against the cuBLAS ZGEMM throughput measured in the same run."""
    torch, sp = job.torch, job.sp
    from spenso_b200.sharding import shard_range

    dev, world, rank, ctx, stream = job.dev, job.world, job.rank, job.ctx, job.stream
    d = args.dim
    n = d ** 3
    first, d1 = shard_range(d, rank, world)        # this rank's block of the first index of A
    rows = d1 * d * d
    g = torch.Generator(device=dev).manual_seed(1236)
    w = 1 if args.real else 2

    def mk(r):
        x = torch.empty(r, n, w, dtype=torch.float64, device=dev)
        flat = x.view(-1)
        for lo in range(0, flat.numel(), 1 << 28):
            flat[lo:lo + (1 << 28)].uniform_(-1.0, 1.0, generator=g)
        return x

    edt = sp.F64 if args.real else sp.C64
    B = mk(n)                                       # same seed on every rank: B is replicated
    A = mk(max(rows, 1))
    sa = [sp.Euclidean.new_slot(d1 if i == 1 else d, i) for i in (1, 2, 3, 4, 5, 6)]
    ta = sp.BatchTensor.wrap(ctx, sa, 1, A.data_ptr(), dtype=edt, keepalive=A)
    tb = sp.BatchTensor.wrap(ctx, [sp.Euclidean.new_slot(d, i) for i in (4, 5, 6, 7, 8, 9)], 1, B.data_ptr(), dtype=edt,
                             keepalive=B)
    for _ in range(W):
        c = ta.contract(tb)
        del c
    job.barrier()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(job.local_rank) as clocks:
        job.barrier()
        e0.record(stream)
        for _ in range(K):
            c = ta.contract(tb)
        e1.record(stream)
        clocks.sample()
        job.barrier()
    ms = job.max_over_ranks(e0.elapsed_time(e1) / K)
    launches = int(ctx.launch_count - launches0)
    fpm = 2.0 if args.real else 8.0      # flops per multiply-add of the element type
    flops = fpm * n * n * n
    del c, ta, tb, A, B
    ctx.release_cached_memory()
    torch.cuda.empty_cache()
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
    del X, Y, Z
    cublas_tf = fpm * m ** 3 / (best * 1e-3) / 1e12
    achieved = flops / world / (ms * 1e-3) / 1e12      # per GPU
    if rank != 0:
        return None
    return {
        "value": 1e3 / ms, "unit": UNIT, "steps": K, "warmup": W, "ms_per_step": ms, "scaling": "strong",
        "config": {"workload": WORKLOAD_DESC["cfg4"], "dim": d, "M=N=K": n,
                   "element": "f64" if args.real else "Complex<f64>",
                   "l2": f"operands {2 * n * n * (8 if args.real else 16) / 2**30:.0f} GiB >> 126 MB L2",
                   "parallelism": f"rows of A/C cut into {world} block(s), B replicated, no collective"},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": cublas_tf, "unit": "TFLOP/s",
                     "frac": achieved / cublas_tf, "traffic": None,
                     "kernel": "contract_dgemm_kernel" if args.real else "contract_zgemm_kernel",
                     "peak_source": f"cuBLAS {'DGEMM' if args.real else 'ZGEMM'} {m}^3 measured in this run (FP64 tensor pipe; "
                                    "tcgen05 has no f64 kind)",
                     "flops_per_eval": flops, "per_gpu": True},
        "clocks": clocks.summary(),
    }


def run_suite(job, args):
    """The `workloads` block: every other BASELINE config in the same run (rank 0 returns the dict)."""
    out = {}
    K = max(3, min(args.steps, 50))
    world = job.world

    def guarded(name, fn):
        try:
            r = fn()
        except Exception as e:  # noqa: BLE001 — one failing workload must not take the headline line down
            import traceback

            traceback.print_exc(file=sys.stderr)
            r = {"error": f"{type(e).__name__}: {e}"}
        job.torch.cuda.empty_cache()
        try:
            job.ctx.release_cached_memory()
        except Exception:
            pass
        if job.rank == 0:
            out[name] = r

    # cfg3: 2^22 points per GPU (39 GB of dense per-point operands), both layouts
    n3 = int(os.environ.get("SPN_BENCH_CFG3_POINTS", 1 << 22))
    if job.free_gib() > (n3 * 9344 / 2 ** 30) * 1.1 + 4:
        guarded("cfg3_point_major", lambda: measure_network(job, args, "cfg3", n3, "point", K=K, min_ms=30.0))
        guarded("cfg3_component_major", lambda: measure_network(job, args, "cfg3", n3, "component", K=K, min_ms=30.0))
    elif job.rank == 0:
        out["cfg3_point_major"] = {"skipped": f"needs {n3 * 9344 / 2**30:.0f} GiB free, have {job.free_gib():.0f}"}
    # cfg5 / cfg5r: 1e8 points in total, sharded over the ranks (BASELINE configs[4]) => strong scaling
    total5 = int(os.environ.get("SPN_BENCH_CFG5_POINTS", 100_000_000))
    n5 = -(-total5 // world)
    note5 = f"{n5 * world} points in total = {n5} per GPU x {world} GPU(s) (1e8 sharded; strong scaling)"
    guarded("cfg5", lambda: measure_network(job, args, "cfg5", n5, "point", K=K, total_points_note=note5, min_ms=30.0))
    guarded("cfg5r", lambda: measure_network(job, args, "cfg5r", n5, "point", K=K, total_points_note=note5, min_ms=30.0))
    if world > 1:
        guarded("cfg5_mc_sum_every_step_peer_kernel",
                lambda: measure_network(job, args, "cfg5", n5, "point", K=K, mc_sum="step", allreduce="peer", total_points_note=note5,
                                        min_ms=30.0))
        guarded("cfg5_mc_sum_every_step_nccl",
                lambda: measure_network(job, args, "cfg5", n5, "point", K=K, mc_sum="step", allreduce="nccl", total_points_note=note5,
                                        min_ms=30.0))
    guarded("mc_sum_exchange", lambda: measure_exchange(job))
    # cfg4: one ZGEMM-shaped contraction (48 GiB of operands at N = 1)
    need4 = (2 + 1.0 / world) * 16.0 + 6
    if job.free_gib() > need4:
        guarded("cfg4", lambda: measure_cfg4(job, args, K=1, W=1))
    elif job.rank == 0:
        out["cfg4"] = {"skipped": f"needs {need4:.0f} GiB free, have {job.free_gib():.0f}"}
    return out


def run_b200(args):
    job = Job()
    wl = args.workload
    suite = args.suite or ("full" if (wl == "cfg2" and args.exec_mode == "fused" and args.layout == "point" and not args.points) else "headline")
    n = args.points or DEFAULT_POINTS[wl]
    head = measure_network(job, args, wl, n, args.layout, K=args.steps, W=args.warmup, exec_mode=args.exec_mode,
                           mc_sum=args.mc_sum, allreduce=args.allreduce, on_device_inputs=args.no_e2e, e2e=not args.no_e2e, cpu=True,
                           sustained_s=args.sustained_seconds)
    extra = run_suite(job, args) if suite == "full" else None
    if job.rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": job.world, "steps": head["steps"],
                "warmup": head["warmup"], "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": head["config"]}
        for k in ("e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks", "sustained", "mc_sum"):
            if k in head:
                line[k] = head[k]
        if "cpu_baseline" not in line:
            line["cpu_baseline"] = None
        if extra is not None:
            line["workloads"] = extra
            line["gpu_launches_all_workloads"] = int(head["gpu_launches"] + sum(
                v.get("gpu_launches", 0) for v in extra.values() if isinstance(v, dict)))
        print(json.dumps(line), flush=True)
    job.close()


def run_cfg4(args):
    job = Job()
    e = measure_cfg4(job, args, K=min(args.steps, 3), W=3 if args.dim < 32 else 1, standalone=True)
    if job.rank == 0:
        line = {"metric": METRIC, "value": e["value"], "unit": UNIT, "n_gpus": job.world, "steps": e["steps"],
                "warmup": e["warmup"], "ms_per_step": e["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": e["config"],
                "gpu_launches": e["gpu_launches"], "roofline": e["roofline"], "clocks": e["clocks"]}
        print(json.dumps(line), flush=True)
    job.close()


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
