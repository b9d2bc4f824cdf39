#!/usr/bin/env python
"""bench.py -- agent-steps/sec of the DecentralPlannerNet hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one addGSO + forward of one batch of synthetic rollout-like input at the
configuration BASELINE.json quotes the metric on (configs[1]): DCP K=3, 10 agents, 20x20 map,
batch 64 inference per GPU (640 agent-steps per step per GPU, weak scaling).

value   device-resident: inputs already in HBM, K back-to-back steps timed with CUDA events
        on the launching stream, max over ranks.  Consecutive steps read different input
        batches from a pool larger than L2 (no L2-resident inputs between iterations).
e2e     the same metric through the public host-buffer call (DecentralPlannerNet.infer_host
        -> gpp_planner_forward_host): every step copies that step's x and S from pinned host
        memory to the device and the logits back, inside the timed region.
roofline / roofline_graph_filter   algorithmic flops / bytes per launch divided by
        the kernel's mean duration measured with CUDA events recorded around each kernel on
        the launching stream in a second timed pass of the same steps.
cpu_baseline   the oracle port (oracle/planner_oracle.py: the reference's own PyTorch-CPU op
        sequence) timed on this box's host cores on a bounded number of steps of the SAME
        workload (rank 0, N=1 only).
--impl reference   times that same CPU path as the reference arm (the reference is pure
        Python/PyTorch and /root/reference does not travel to the GPU box).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "agent-steps/sec (DCP K=3, 10 agents, 20x20 map)"
UNIT = "agent-steps/s"
K_TAPS, N_AGENTS, MAP_W, BATCH = 3, 10, 20, 64
WORKLOAD = "DCP K=3, 10 agents, 20x20 map, batch=64 inference per GPU (BASELINE.json configs[1])"

# SURVEY.md section 8d algorithmic figures
GF_BYTES_PER_AGENT_STEP = 4 * (128 * N_AGENTS + N_AGENTS * N_AGENTS + 128 * N_AGENTS) // N_AGENTS   # 1064
GF_FLOPS_PER_AGENT_STEP = (2 * 128 * N_AGENTS * N_AGENTS * (K_TAPS - 1) + 2 * N_AGENTS * K_TAPS * 128 * 128) // N_AGENTS
FE_FLOPS_PER_AGENT_STEP = 2476224 + 32768          # 5 conv layers + compress MLP


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured"
    except Exception:
        return 6650.0, 1400.0, "fallback"       # B200_PROFILING.md fallback figures


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while a timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.01)

    def __enter__(self):
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_state_dict():
    from oracle import planner_oracle as po      # parameter construction only (not timed, not compute)
    sd = po.init_state_dict(K_TAPS, seed=1337)
    po.randomize_bn_stats(sd)
    return sd


def make_inputs(unique, seed):
    from gnn_pathplanning_b200 import synthetic
    xs, Ss = [], []
    for i in range(unique):
        x, S = synthetic.make_batch(BATCH, N_AGENTS, MAP_W, seed=seed + i)
        xs.append(torch.from_numpy(x))
        Ss.append(torch.from_numpy(S))
    return xs, Ss


def cpu_forward_factory(sd, xs, Ss):
    from oracle import planner_oracle as po

    def step(i):
        with torch.no_grad():
            return po.planner_forward(sd, Ss[i % len(Ss)], xs[i % len(xs)])
    return step


def usable_cpus():
    """CPUs this process may really use: affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            q, per = f.read().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per) + 0.5)))
    except Exception:
        pass
    return n


def pick_cpu_threads(step, budget_s=6.0):
    """The reference's PyTorch-CPU path is N sequential tiny convolutions; on a many-core host
    it gets SLOWER with more intra-op threads (measured: 19 s/step at 128 threads vs ~20 ms at
    8).  The CPU arm is given its best case: the thread count with the highest measured
    throughput among {all usable, 64, 32, 16, 8, 4, 2, 1}."""
    cap = usable_cpus()
    cands = sorted({c for c in (cap, 64, 32, 16, 8, 4, 2, 1) if c <= cap}, reverse=True)
    best, best_t = 1, float("inf")
    t_start = time.perf_counter()
    for c in reversed(cands):                       # small counts first: they are cheap to probe
        torch.set_num_threads(c)
        step(0)
        t0 = time.perf_counter()
        step(1)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
        if time.perf_counter() - t_start > budget_s or dt > 4 * best_t:
            break
    torch.set_num_threads(best)
    return best, cap


def time_cpu(step, budget_s, min_steps=5, warmup=3):
    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    n = 0
    while True:
        step(n)
        n += 1
        el = time.perf_counter() - t0
        if n >= min_steps and el >= budget_s:
            break
    return n, el


def saturated_filter_roofline(gp, dev, batch, hbm_peak, peak_src):
    """The fused graph filter alone (node-major in/out, + ReLU) on a batch large enough to fill the
    machine: the headline configuration moves only 0.68 MB per launch and is launch-latency bound, so
    this is where the kernel's bandwidth fraction is visible.  Inputs (batch x 10 x 128 floats and
    the GSOs) exceed L2; CUDA events around back-to-back launches; the ~3 us tap re-layout launch
    that precedes every standalone call is inside the timed region."""
    g = torch.Generator().manual_seed(0)
    w = ((torch.rand(128, 1, K_TAPS, 128, generator=g) - 0.5) * 0.2).to(dev)
    b = (torch.rand(128, 1, generator=g) - 0.5).to(dev)
    x = torch.randn(batch, N_AGENTS, 128, device=dev)
    # rollout-like GSOs: normalised adjacency of random agent positions on the 20x20 map, radius 6
    pos = torch.randint(0, MAP_W, (batch, N_AGENTS, 2), device=dev).float()
    A = ((pos[:, :, None, :] - pos[:, None, :, :]).norm(dim=-1) < 6.0).float() * (1.0 - torch.eye(N_AGENTS, device=dev))
    dinv = A.sum(-1).clamp(min=1.0).rsqrt() * (A.sum(-1) > 0)
    S = (dinv[:, :, None] * A * dinv[:, None, :]).contiguous()
    for _ in range(3):
        gp.graph_filter(x, S, w, b, True, gp.NODE_MAJOR, gp.NODE_MAJOR)
    torch.cuda.synchronize()
    iters = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        gp.graph_filter(x, S, w, b, True, gp.NODE_MAJOR, gp.NODE_MAJOR)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / iters
    nbytes = GF_BYTES_PER_AGENT_STEP * batch * N_AGENTS
    gbs = nbytes / sec / 1e9
    return {"kernel": "gf_fwd_tc_kernel (tcgen05 3xTF32, TMEM accumulators) via gpp_graph_filter_forward",
            "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
            "traffic": None, "peak_source": peak_src, "mean_launch_us": sec * 1e6,
            "workload": "K=3, 10 agents, %d episodes (%.0f MB of node signals + GSOs, > L2)" % (batch, nbytes / 2e6),
            "agent_steps_per_s": batch * N_AGENTS / sec, "algorithmic_bytes_per_launch": nbytes}


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation of the path = its PyTorch-CPU op
    sequence, restated in oracle/planner_oracle.py (the Python reference cannot travel to the
    GPU box).  All host threads, same config/metric/unit as the CUDA arm."""
    if rank != 0:
        return
    sd = make_state_dict()
    xs, Ss = make_inputs(4, 1337)
    step = cpu_forward_factory(sd, xs, Ss)
    threads, cap = pick_cpu_threads(step)
    # bounded sample: if K full-batch steps would not end within ~4 minutes on this host, each
    # step processes the first `sub` episodes of the batch instead (same per-episode work)
    t0 = time.perf_counter()
    step(0)
    t_step = time.perf_counter() - t0
    sub = BATCH
    if t_step * (args.steps + args.warmup) > 240.0:
        sub = max(1, int(BATCH * 240.0 / (t_step * (args.steps + args.warmup))))
        xs, Ss = [x[:sub].contiguous() for x in xs], [s[:sub].contiguous() for s in Ss]
        step = cpu_forward_factory(sd, xs, Ss)
    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    el = time.perf_counter() - t0
    val = sub * N_AGENTS * args.steps / el
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "K": K_TAPS, "num_agents": N_AGENTS, "map_w": MAP_W,
                   "batch_per_gpu": BATCH, "note": "CPU arm runs one batch per step on rank 0 only"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "usable_cpus": cap,
                         "sample": "%d steps of %d episodes x 10 agents of the same workload; thread count "
                                   "auto-picked for best CPU throughput" % (args.steps, sub)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--pool-mb", type=float, default=160.0, help="device input pool size (> L2)")
    ap.add_argument("--gf-mode", default="auto", choices=["auto", "cuda", "tc"],
                    help="graph-filter kernel of the planner: auto / CUDA-core / tcgen05")
    ap.add_argument("--fe-mode", default="auto", choices=["auto", "cuda", "tc"],
                    help="feature-extractor (CNN) kernel of the planner: auto / CUDA-core / tcgen05")
    ap.add_argument("--sat-batch", type=int, default=32768,
                    help="episodes in the saturated graph-filter roofline measurement (0 = skip)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 200
        args.warmup = args.warmup if args.warmup is not None else 3
        run_reference(args, rank, world)
        return
    args.steps = args.steps if args.steps is not None else 4000
    args.warmup = max(3, args.warmup if args.warmup is not None else 50)

    import gnn_pathplanning_b200 as gp
    from gnn_pathplanning_b200 import _lib
    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    class Cfg:
        num_agents, nGraphFilterTaps, device = N_AGENTS, K_TAPS, dev

    sd = make_state_dict()
    model = gp.DecentralPlannerNet(Cfg())
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    model.set_graph_filter_mode(args.gf_mode)
    model.set_feature_mode(args.fe_mode)

    # ---- inputs: a few unique synthetic batches, replicated (sample-permuted) into a pool > L2
    unique = 8
    xs_h, Ss_h = make_inputs(unique, 1337 + 1000 * rank)
    bytes_per_batch = xs_h[0].numel() * 4 + Ss_h[0].numel() * 4
    pool_n = max(unique, int(args.pool_mb * 1e6 / bytes_per_batch) + 1)
    g = torch.Generator().manual_seed(rank)
    pool_x = torch.empty((pool_n,) + tuple(xs_h[0].shape), device=dev)
    pool_S = torch.empty((pool_n,) + tuple(Ss_h[0].shape), device=dev)
    for i in range(pool_n):
        perm = torch.randperm(BATCH, generator=g)
        pool_x[i].copy_(xs_h[i % unique][perm])
        pool_S[i].copy_(Ss_h[i % unique][perm])
    logits = None

    def dev_step(i):
        model.addGSO(pool_S[i % pool_n])
        return model(pool_x[i % pool_n])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    with torch.no_grad():
        # ---- value: device-resident ------------------------------------------------------
        for i in range(args.warmup):
            dev_step(i)
        barrier()
        _lib.reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with sampler:
            e0.record()
            for i in range(args.steps):
                logits = dev_step(i)
            e1.record()
            torch.cuda.synchronize()
        launches = _lib.launch_count()
        barrier()
        ms = e0.elapsed_time(e1)

        # ---- per-kernel durations (same steps, events recorded around each kernel) -------
        nat = model._native_for(dev)
        prof_steps = min(args.steps, 2000)
        _lib.check(nat.lib.gpp_planner_set_profiling(nat.handle, 1))
        for i in range(prof_steps):
            dev_step(i)
        import ctypes as C
        fe_ms, gf_ms, nst = C.c_double(), C.c_double(), C.c_int()
        _lib.check(nat.lib.gpp_planner_get_profile(nat.handle, C.byref(fe_ms), C.byref(gf_ms), C.byref(nst)))
        _lib.check(nat.lib.gpp_planner_set_profiling(nat.handle, 0))
        barrier()

        # ---- e2e: host buffers through the public host call ------------------------------
        hx = [t.pin_memory() for t in xs_h]
        hS = [t.pin_memory() for t in Ss_h]
        hout = torch.empty(N_AGENTS, BATCH, 5).pin_memory()
        e2e_steps = max(2, args.steps // 4)
        for i in range(max(3, args.warmup // 4)):
            model.infer_host(hx[i % unique], hS[i % unique], hout)
        barrier()
        # (a) synchronous: one blocking host call per step
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            model.infer_host(hx[i % unique], hS[i % unique], hout)
        torch.cuda.synchronize()
        e2e_sync_s = time.perf_counter() - t0
        barrier()
        # (b) pipelined over independent episode batches (depth 3): step i is enqueued before step i-2
        # is waited for, as a rollout driver advancing three batches of episodes in turn would do;
        # every step still reads its x / S from pinned host memory and writes its logits back to it
        DEPTH = 3
        houts = [torch.empty(N_AGENTS, BATCH, 5).pin_memory() for _ in range(DEPTH)]
        checksum = 0.0
        for i in range(max(3, args.warmup // 4)):          # untimed: staging slots, copy stream and events are created
            model.wait(model.infer_host_async(hx[i % unique], hS[i % unique], houts[i % DEPTH]))
        barrier()
        with sampler:
            t0 = time.perf_counter()
            inflight = []
            for i in range(e2e_steps):
                inflight.append((model.infer_host_async(hx[i % unique], hS[i % unique], houts[i % DEPTH]), i % DEPTH))
                if len(inflight) >= DEPTH:
                    tk, slot = inflight.pop(0)
                    model.wait(tk)
                    checksum += float(houts[slot][0, 0, 0])      # the step's result is read on the host
            for tk, slot in inflight:
                model.wait(tk)
                checksum += float(houts[slot][0, 0, 0])
            torch.cuda.synchronize()
            e2e_s = time.perf_counter() - t0
        barrier()

    # ---- max over ranks --------------------------------------------------------------------
    if dist is not None:
        t = torch.tensor([ms, e2e_s, e2e_sync_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s, e2e_sync_s = float(t[0]), float(t[1]), float(t[2])
    agent_steps = BATCH * N_AGENTS
    value = world * agent_steps * args.steps / (ms * 1e-3)
    e2e_value = world * agent_steps * e2e_steps / e2e_s

    if rank == 0:
        hbm_peak, tf_peak, peak_src = load_peaks()
        n = max(1, nst.value)
        fe_s, gf_s = fe_ms.value * 1e-3 / n, gf_ms.value * 1e-3 / n
        fe_tflops = FE_FLOPS_PER_AGENT_STEP * agent_steps / fe_s / 1e12
        gf_gbs = GF_BYTES_PER_AGENT_STEP * agent_steps / gf_s / 1e9
        clocks = sampler.summary()
        sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
        fp32_peak = 148 * 128 * 2 * float(sm_mhz) * 1e6 / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "K": K_TAPS, "num_agents": N_AGENTS, "map_w": MAP_W,
                       "batch_per_gpu": BATCH, "global_batch": BATCH * world,
                       "parallelism": "episode-sharded replicas x%d, no collective (inference)" % world,
                       "l2_policy": "inputs rotate through a %.0f MB device pool (> 126 MB L2), %d batches"
                                    % (pool_n * bytes_per_batch / 1e6, pool_n)},
            "e2e": {"value": e2e_value, "unit": UNIT, "steps": e2e_steps,
                    "h2d_bytes_per_step": bytes_per_batch, "d2h_bytes_per_step": N_AGENTS * BATCH * 5 * 4,
                    "api": "DecentralPlannerNet.infer_host_async/wait -> gpp_planner_forward_host_async (pinned host "
                           "buffers; inputs staged by a small copy kernel on a second stream while the previous step's "
                           "kernels run, logits written straight to host), 3 independent episode batches in flight",
                    "sync_value": world * agent_steps * e2e_steps / e2e_sync_s,
                    "sync_api": "DecentralPlannerNet.infer_host -> gpp_planner_forward_host, one blocking call per step"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "feature_kernel (CNN + compress MLP, fp32 FMA)", "bound": "tensor",
                         "achieved": fe_tflops, "peak": tf_peak, "unit": "TFLOP/s", "frac": fe_tflops / tf_peak,
                         "traffic": None, "peak_source": peak_src + " bf16 dense (sustained)",
                         "mean_launch_us": fe_s * 1e6,
                         "algorithmic_flops_per_launch": FE_FLOPS_PER_AGENT_STEP * agent_steps,
                         # the kernel computes in fp32 on the CUDA cores (parity bar 1e-5): its own ceiling is the
                         # fp32 FMA rate, 148 SMs x 128 lanes x 2 flop x SM clock
                         "fp32_fma_peak": fp32_peak, "frac_of_fp32_fma_peak": fe_tflops / fp32_peak},
            "roofline_graph_filter": {"kernel": "gf_fwd_kernel (K-tap filter + ReLU + action MLP)", "bound": "hbm",
                                      "achieved": gf_gbs, "peak": hbm_peak, "unit": "GB/s",
                                      "frac": gf_gbs / hbm_peak, "traffic": None, "peak_source": peak_src,
                                      "mean_launch_us": gf_s * 1e6,
                                      "algorithmic_bytes_per_launch": GF_BYTES_PER_AGENT_STEP * agent_steps,
                                      "algorithmic_flops_per_launch": GF_FLOPS_PER_AGENT_STEP * agent_steps},
        }
        if args.sat_batch > 0:
            line["roofline_graph_filter_saturated"] = saturated_filter_roofline(gp, dev, args.sat_batch, hbm_peak, peak_src)
        if world == 1:
            step = cpu_forward_factory(sd, xs_h, Ss_h)
            threads, cap = pick_cpu_threads(step, budget_s=min(6.0, max(0.5, args.cpu_seconds)))
            n_cpu, el = time_cpu(step, args.cpu_seconds)
            line["cpu_baseline"] = {"value": agent_steps * n_cpu / el, "unit": UNIT,
                                    "cores": threads, "kind": "port", "usable_cpus": cap,
                                    "sample": "%d full steps (batch 64 x 10 agents) of the same workload, %.1f s"
                                              % (n_cpu, el)}
            # parity spot-check of the timed configuration (not timed)
            ref = torch.stack(step(0)).numpy()
            model.addGSO(Ss_h[0].to(dev))
            with torch.no_grad():
                got = torch.stack(model(xs_h[0].to(dev))).cpu().numpy()
            line["parity_max_rel"] = float(np.abs(got - ref).max() / np.abs(ref).max())
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
