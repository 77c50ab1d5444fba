#!/usr/bin/env python
"""Benchmark of the VFM training step (BASELINE.json metric: fwd+bwd(+Adam) samples/s and
achieved HBM GB/s vs the measured peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload ml20m|sideinfo|big100m|ml100k]

A "step" is one pass of the hot path -- batch plan, forward, backward, Adam on the touched
rows -- over one batch of a synthetic dataset of the named shape.  Batches
are consecutive, never shuffled, as in the reference loaders (vfm-torch.py:121-122).

One JSON line is printed by rank 0; see DESIGN.md ("Measurement") for every key.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from vae_b200 import synth                                   # noqa: E402

METRIC = "vfm_train_step_samples_per_s"
UNIT = "samples/s"


# ------------------------------------------------------------------------------------------
def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(B, F, d, U):
    """ids, target, prediction + per touched row parameters, Adam m and v read and
    written once, plus its int64 train count."""
    return B * (8 * F + 4 + 4) + U * ((2 * d + 2) * 4 * 6 + 8)


def rows_kernel_bytes(B, F, d, U):
    """Compulsory traffic of the dominant kernel (k_adam_rows): parameters, Adam m and v of every
    touched row read once and written once ("Kernels").  The gradient / noise rows it
    also reads are L2-resident scratch and are not credited."""
    return U * (2 * d + 2) * 4 * 6


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs.

    The thread is started BEFORE the warm-up (its first NVML calls are slow and, on a timed region of a
    few milliseconds, landed inside it: 115 vs 103 us per step at --steps 20) and samples every 50 ms; the
    samples that count are those between ``mark()`` and ``stop()``, plus one taken by the caller with
    ``sample()`` right after the last step is enqueued, i.e. while the GPU is inside the timed region."""

    def __init__(self, index: int, period_s: float = 0.05):
        super().__init__(daemon=True)
        self.period, self.samples, self.reasons, self.max_mhz = period_s, [], set(), None
        self._stop_evt = threading.Event()
        self._t_mark = None
        self._lock = threading.Lock()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.names = {pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                          pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                          pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                          pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown"}
        except Exception:                                     # pragma: no cover
            self.nv = None

    def sample(self):
        if self.nv is None:
            return
        try:
            mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
            mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        except Exception:                                     # pragma: no cover
            return
        with self._lock:
            if self._t_mark is None:
                return                                        # before the timed region: warms NVML up only
            self.samples.append(mhz)
            for bit, name in self.names.items():
                if mask & bit:
                    self.reasons.add(name)

    def run(self):
        while self.nv is not None and not self._stop_evt.is_set():
            self.sample()
            time.sleep(self.period)

    def mark(self):
        """Start of the timed region: samples count from here."""
        with self._lock:
            self._t_mark = time.time()

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def learning_rate(w: synth.Workload) -> float:
    """vfm-torch.py:92 (1 / (1 + N_train // batch)); vfm-tomasrch.py:518 (0.1)."""
    return 0.1 if w.variant == "closed" else 1.0 / (1 + w.n_train // w.batch)


def alpha_0(w: synth.Workload) -> float:
    """vfm-tomasrch.py:742-743: alpha_0 = 0.5 * number of batches per epoch."""
    return 0.5 * -(-w.n_train // w.batch)


def make_model(w: synth.Workload, device, train_counts, lr, max_batch=None):
    torch.manual_seed(synth.PARAM_SEED)
    max_batch = max_batch or w.batch
    if w.variant == "closed":                                 # vfm-tomasrch.py path (BASELINE config 2)
        from vae_b200.vfm_tomasrch import CF as ClosedCF
        return ClosedCF(w.d, w.n_fields, w.field_sizes, 1, alpha_0(w), "reg", train_counts=torch.from_numpy(train_counts),
                        n_train=w.n_train, max_batch=max_batch, lr=lr, device=device)
    from vae_b200.vfm_torch import CF
    kl = "torch" if w.n_fields == 2 else "group"
    return CF(w.d, output=w.output, n_users=w.field_sizes[0], n_items=w.field_sizes[1],
              train_counts=torch.from_numpy(train_counts), field_sizes=w.field_sizes, kl_weighting=kl,
              n_train=w.n_train, max_batch=max_batch, seed=synth.NOISE_SEED, lr=lr, device=device)


def make_port(w: synth.Workload, train_counts, lr):
    """The oracle's torch restatement of the reference model + its optimizer + a step function
    (CPU leg only: cpu_baseline / parity_vs_port / --impl reference)."""
    from oracle import vfm_port
    tc = torch.from_numpy(train_counts)
    torch.manual_seed(synth.PARAM_SEED)
    if w.variant == "closed":
        port = vfm_port.ClosedPort(w.d, w.field_sizes, alpha_0=alpha_0(w))
        opt = torch.optim.Adam(port.parameters(), lr=lr)
        tcf = tc.to(torch.float32)
        step = lambda x, y, noise=None: vfm_port.closed_port_step(port, opt, x, y, w.n_train, tcf)
        return port, opt, step
    kl = "torch" if w.n_fields == 2 else "group"
    port = vfm_port.SampledPort(w.field_sizes[0], w.field_sizes[1], w.d, tc, output=w.output,
                                field_sizes=w.field_sizes, kl_weighting=kl,
                                interaction="prod" if w.n_fields == 2 else "pairwise")
    with torch.no_grad():                                    # N(0,1) init of 10^7..10^8 values holds a few exact zeros: a
        for prm in port.parameters():                        # raw scale of 0 is sigma = 0 under the abs link, which torch's
            prm[prm == 0] = 1e-6                             # Normal rejects (the CUDA path floors sigma, common.cuh)
    opt = torch.optim.Adam(port.parameters(), lr=lr)
    step = lambda x, y, noise=None: vfm_port.sampled_port_step(port, opt, x, y, w.n_train, noise)
    return port, opt, step


def apply_tuning(args):
    from vae_b200 import _lib as L
    for kv in args.tune:
        k, v = kv.split("=")
        L.check(L.lib().vfmb_set_tuning(k.encode(), int(v)), "vfmb_set_tuning")


def run_ours(args):
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    apply_tuning(args)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    w = synth.make_workload(args.workload, n_rows=args.rows)
    B, F, d = w.batch, w.n_fields, w.d
    n_batches = w.n_train // B
    assert n_batches >= 1
    tc = w.train_counts()
    tc[tc == 0] = 1                                           # rows never seen in training: weight 1
    lr = learning_rate(w)
    model = make_model(w, device, tc, lr)
    # weak scaling over independent replicas is NOT the multi-GPU mode of this path; see
    # vae_b200/dist.py.  For N>1 each rank takes every world-th batch of the global stream.
    x_all = torch.from_numpy(w.x[: n_batches * B]).to(device)
    y_all = torch.from_numpy(w.y[: n_batches * B]).to(device)

    def batch(i):
        j = (i * world + rank) % n_batches
        return x_all[j * B:(j + 1) * B], y_all[j * B:(j + 1) * B]

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    W, K = args.warmup, args.steps

    dp = None
    if world > 1:                                            # mode A: replicated tables, dense all-reduce
        assert w.variant == "sampled", "multi-GPU modes exist for the sampled step"
        from vae_b200.dist import DataParallelSampled
        dp = DataParallelSampled(model, world, dense_adam=False)

    def run_steps(lo, hi):
        """Software-pipelined loop: the plan of batch i+1 is built on a side stream while the
        kernels of batch i run (every batch still gets its own plan, built on the GPU)."""
        out = None
        if dp is not None:
            for i in range(lo, hi):
                out = dp.step(*batch(i))
            return out
        if args.plan == "graph":                             # CUDA-graph replay of the same pipelined loop:
            for i in range(lo, hi):                          # run batch i (the plan of batch i+1 is built on the
                out = loop.step(*batch(i + 2))               # side branch of the same graph), stage batch i+2
            return out
        if args.plan == "prefetch":
            model.prefetch_plan(batch(lo)[0])
        for i in range(lo, hi):
            xb, yb = batch(i)
            if args.plan == "prefetch" and i + 1 < hi:
                model.prefetch_plan(batch(i + 1)[0])
            if args.plan == "cached":
                out = model.fused_step(xb, yb, plan=static_plans[(i * world + rank) % n_batches])
            else:
                out = model.fused_step(xb, yb)
        return out

    from vae_b200 import _lib as L
    # three staging slots: the copy of batch i+2 into its slot runs on a copy stream, off the step's stream
    gkw = dict(reserve=args.reserve, side_priority=args.side_priority)
    loop = model.graphed_loop(B, depth=3, **gkw) if (args.plan == "graph" and dp is None) else None
    if loop is not None:
        loop.start(*batch(0))                                # the pipeline runs on from here: warm-up, then timed
        loop.stage(*batch(1))
    static_plans = {}
    if args.plan == "cached":                               # never-shuffled loader: plans recur every epoch
        for i in range(W, W + K):
            j = (i * world + rank) % n_batches
            if j not in static_plans:
                static_plans[j] = model.static_plan(batch(i)[0])
    sampler = ClockSampler(local)
    sampler.start()
    run_steps(0, W)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    n_launch0 = int(L.lib().vfmb_launch_count())
    sampler.mark()
    ev0.record()
    last = run_steps(W, W + K)
    ev1.record()
    sampler.sample()                                         # the GPU is still inside the timed region here
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    loss = float(last["loss"].item())
    # own kernels launched inside the timed region: counted by the library at every launch site
    # (eager launches), plus the kernel nodes of the replayed graphs (counted when they were captured)
    gpu_launches = int(L.lib().vfmb_launch_count()) - n_launch0
    if loop is not None:
        gpu_launches += loop.launches_per_graph * K

    # dominant kernel (k_adam_rows), timed per launch with CUDA events recorded by the library on the
    # launch stream immediately around it (vfmb_profile_events), over live training steps
    kr = []
    if dp is None:
        for i in range(W + K, W + K + min(K, 200)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(), e1.record()                          # materialise the cudaEvent_t handles
            L.lib().vfmb_profile_events(e0.cuda_event, e1.cuda_event)
            model.fused_step(*batch(i))
            kr.append((e0, e1))
        L.lib().vfmb_profile_events(None, None)
    torch.cuda.synchronize()
    rows_ms = float(np.mean([a.elapsed_time(b) for a, b in kr])) if kr else float("nan")

    # measured U of the timed batches (outside the timed region)
    us = [int(torch.unique(batch(i)[0]).numel()) for i in range(W, W + min(K, 64))]
    U = float(np.mean(us))

    # end to end through the public API with HOST inputs: pinned x/y -> device every step,
    # loss/KL scalars back to pinned host memory every step
    e2e = measure_e2e(model, w, rank, world, n_batches, W, min(K, 500), device, barrier, dp, loop, gkw)

    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    dp_par = None
    if world > 1 and dp is not None:                          # real-rank parity of mode A (every rank takes part)
        try:
            dp_par = dp_parity(args, w, device, rank, world)
        except Exception as exc:                              # the throughput line stands on its own
            dp_par = {"error": f"{type(exc).__name__}: {exc}"}
    if rank != 0:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    value = K * B * world / (ms * 1e-3)
    step_bytes = algorithmic_bytes(B, F, d, U)
    rk_bytes = rows_kernel_bytes(B, F, d, U)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get(args.workload, {}).get("k_adam_rows_dram_bytes")
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{w.name}: {'+'.join(map(str, w.field_sizes))} rows, d={d}, "
                               f"{w.n_train} ratings, batch {B}, {w.variant} ELBO, {w.output}",
                   "fields": F, "unique_rows_per_step": U, "adam": "touched rows (lazy)",
                   "noise": "none (closed form)" if w.variant == "closed" else "Philox4x32-10 in-kernel", "plan": {"inline": "built every step on the step's stream",
                            "prefetch": "built every step, one batch ahead on a side stream",
                            "graph": "built every step, one batch ahead on a side stream; step + plan "
                                     "replayed as one CUDA graph",
                            "cached": "precomputed per batch (never-shuffled loader)"}[args.plan]
                           + " (own tiled LSD radix sort, no library kernels)",
                   "l2": "params+Adam state 258 MB > 126 MB L2; consecutive distinct batches, no flush"
                   if args.workload == "ml20m" else
                   ("state fits L2: the step is launch/latency-bound, no roofline claim (BASELINE.md section 3)"
                    if args.workload in ("ml100k", "fraction") else "consecutive distinct batches, no flush"),
                   "parallelism": (f"dp{world}: replicated tables, batch {B}/GPU, one NCCL all-reduce of the "
                                   f"dense gradient ({(w.rows * (2 * d + 3) + 16) * 4 / 1e6:.1f} MB) per step")
                   if world > 1 else "single"},
        "roofline": {"bound": "hbm", "kernel": ("k_cadam" if w.variant == "closed" else "k_adam_rows")
                                                + " (chain rule + Adam on the touched rows)",
                     "achieved": (rk_bytes / (rows_ms * 1e-3) / 1e9) if rows_ms == rows_ms else None,
                     "peak": peak, "unit": "GB/s",
                     "frac": (rk_bytes / (rows_ms * 1e-3) / 1e9 / peak) if rows_ms == rows_ms else None,
                     "traffic": traffic, "peak_source": peak_src,
                     "kernel_ms": rows_ms if rows_ms == rows_ms else None, "algorithmic_bytes": rk_bytes},
        "roofline_step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms / K * 1e-3) / 1e9,
                          "frac": step_bytes / (ms / K * 1e-3) / 1e9 / peak, "unit": "GB/s"},
        "e2e": e2e, "gpu_launches": gpu_launches * world,
        "gpu_launches_per_step": gpu_launches / K,
        "library_launches_per_step": "none (2 cudaMemsetAsync nodes in the plan)" if args.plan != "cached" else "none",
        "clocks": clocks, "final_loss": loss,
    }
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(args, budget_s=args.cpu_budget)
        out["parity"] = parity_vs_port(args, device)
    if dp_par is not None:
        out["parity"] = dp_par
    print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def measure_e2e(model, w, rank, world, n_batches, W, K, device, barrier, dp=None, loop=None, gkw=None):
    B, F = w.batch, w.n_fields
    xh = torch.from_numpy(w.x[: n_batches * B]).pin_memory()
    yh = torch.from_numpy(w.y[: n_batches * B]).pin_memory()
    xd = [torch.empty((B, F), dtype=torch.int64, device=device) for _ in range(3)]
    yd = [torch.empty(B, dtype=torch.float32, device=device) for _ in range(3)]
    res = torch.empty((K + W, 16), dtype=torch.float32).pin_memory()
    # the reference loop pulls the predictions of every batch to the host (vfm-torch.py:363)
    predh = torch.empty((4, B), dtype=torch.float32).pin_memory()

    copied = [torch.cuda.Event() for _ in range(3)]

    def stage(i):
        """pinned host -> device copy of batch i's ids and targets, then its plan on the side stream"""
        j = (i * world + rank) % n_batches
        s = i % 3
        xd[s].copy_(xh[j * B:(j + 1) * B], non_blocking=True)
        yd[s].copy_(yh[j * B:(j + 1) * B], non_blocking=True)
        copied[s].record()
        if dp is None:
            model.prefetch_plan(xd[s], after=copied[s])

    def host_batch(i):
        j = (i * world + rank) % n_batches
        return xh[j * B:(j + 1) * B], yh[j * B:(j + 1) * B]

    hloop = model.graphed_loop(B, depth=3, **(gkw or {})) if loop is not None else None

    started = []
    d2h = torch.cuda.Stream(device=device)
    d2h_done = [torch.cuda.Event() for _ in range(3)]

    def run_graph(lo, hi):
        """graphed loop fed straight from pinned host memory: the staging copies ARE the H2D copies
        (batch i+2 is copied on a copy stream while step i runs and the plan of i+1 is built); the
        outputs of step i (its own slot) go back to the host on a third stream while step i+1 runs;
        the pipeline is started once and runs on through warm-up and the timed steps"""
        cur = torch.cuda.current_stream(device)
        if not started:
            hloop.start(*host_batch(lo))
            hloop.stage(*host_batch(lo + 1))
            started.append(True)
        for i in range(lo, hi):
            hloop.stage(*host_batch(i + 2))
            s = hloop.head
            cur.wait_event(d2h_done[s])                       # this replay rewrites the outputs of slot s
            out = hloop.step()
            d2h.wait_event(hloop.done[s])
            with torch.cuda.stream(d2h):
                res[i].copy_(out["stats"][:16], non_blocking=True)
                predh[i % 4].copy_(out["pred"], non_blocking=True)
                d2h_done[s].record(d2h)

    def run(lo, hi):
        if loop is not None:
            return run_graph(lo, hi)
        stage(lo)
        for i in range(lo, hi):
            if i + 1 < hi:
                stage(i + 1)
            s = i % 3
            out = model.fused_step(xd[s], yd[s]) if dp is None else dp.step(xd[s], yd[s])
            res[i].copy_(out["stats"][:16], non_blocking=True)      # loss / KL / NLL back to the host
            predh[i % 4].copy_(out["pred"], non_blocking=True)      # and the batch's predictions

    run(0, W)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(W, W + K)
    torch.cuda.current_stream(device).wait_stream(d2h)         # the last outputs have reached the host
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    assert np.isfinite(res[W:W + K, 0].numpy()).all(), "non-finite loss in the e2e run"
    return {"value": K * B * world / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * F * 8 + B * 4,
            "d2h_bytes_per_step": 16 * 4 + B * 4, "steps": K, "ms_per_step": ms / K}


# ------------------------------------------------------------------------------------------
def run_sharded(args):
    """N > 1: mode B -- tables row-sharded over the ranks (owner = row mod P), every rank feeds its
    own batch of `batch` samples (weak scaling); per step 3 exchanges (ids, sampled rows, row
    gradients), written by the pack kernels straight into the peers' buffers over NVLink
    (`--exchange peer`, default) or run as NCCL all-to-alls (`--exchange nccl`), with the plans and
    the id exchange of the next batch under the current step (ShardedPipeline).  The step on the
    global batch equals the single-process step on that batch (tests/test_gpu_dp.py)."""
    import torch.distributed as dist
    from vae_b200.dist import ShardedSampled
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    w = synth.make_workload(args.workload, n_rows=args.rows)
    B, F, d = w.batch, w.n_fields, w.d
    n_batches = w.n_train // B
    tc = w.train_counts()
    tc[tc == 0] = 1                                            # rows never seen in training: weight 1
    lr = 1.0 / (1 + w.n_train // B)
    kl = "torch" if F == 2 else "group"
    exchange = args.exchange
    if exchange == "peer":                                     # NVLink peer memory; all ranks must agree
        ok = 1
        try:
            import importlib
            importlib.import_module("torch.distributed._symmetric_memory")
        except Exception:
            ok = 0
        flag = torch.tensor([ok], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            exchange = "nccl"
    def build(ex):
        return ShardedSampled(d, w.field_sizes, torch.from_numpy(tc), w.n_train, B, world, rank,
                              output=w.output, kl_weighting=kl, seed=synth.NOISE_SEED, lr=lr, device=device,
                              slack=args.slack, exchange=ex)
    model = None
    if exchange == "peer":
        try:
            model = build("peer")
            ok = 1
        except Exception as exc:                              # no peer mapping on this box: NCCL all-to-alls
            print(f"[bench] rank {rank}: peer-memory exchange unavailable ({type(exc).__name__}: {exc}); using NCCL",
                  file=sys.stderr)
            ok = 0
        flag = torch.tensor([ok], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            exchange, model = "nccl", None
    if model is None:
        model = build(None)
    x_all = torch.from_numpy(w.x[: n_batches * B]).to(device)
    y_all = torch.from_numpy(w.y[: n_batches * B]).to(device)

    def batch(i):
        j = (i * world + rank) % n_batches
        return x_all[j * B:(j + 1) * B], y_all[j * B:(j + 1) * B]

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    W, K = args.warmup, args.steps
    sampler = ClockSampler(local)
    sampler.start()
    for i in range(W):
        out = model.step(*batch(i))
    barrier()
    step = model.step
    graphed = False
    pipe = None
    if args.plan == "graph":                                  # whole step incl. the collectives as one CUDA graph
        try:
            if exchange == "peer" and not args.no_pipeline:
                from vae_b200.dist import ShardedPipeline
                pipe = ShardedPipeline(model, reserve=args.reserve, side_priority=args.side_priority)
            else:
                step = model.graphed_step()
            graphed = True
        except Exception as exc:                              # capture of NCCL collectives unsupported: stay eager
            print(f"[bench] rank {rank}: graph capture failed ({type(exc).__name__}: {exc}); eager steps", file=sys.stderr)
            step = model.step
        flag = torch.tensor([1 if graphed else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)           # all ranks must agree
        if int(flag.item()) == 0:
            step, graphed, pipe = model.step, False, None
        if pipe is not None:
            nxt = {"i": 0}

            def step(xb, yb):                                 # same call shape as model.step: the batch passed
                # is the one staged for the FOLLOWING replay; batches are consumed in order
                return pipe.step(xb, yb)
            pipe.start(*batch(0), *batch(1))
            for i in range(W):
                out = pipe.step(*batch(i + 2))
        else:
            for i in range(W):
                out = step(*batch(i))
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    ev0.record()
    if pipe is not None:
        for i in range(W, W + K):
            out = pipe.step(*batch(i + 2))                    # runs batch i, stages batch i+2
    else:
        for i in range(W, W + K):
            out = step(*batch(i))
    ev1.record()
    sampler.sample()                                         # the GPU is still inside the timed region here
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    model.check_overflow()
    loss = float(out["loss"].item())

    # end to end: pinned host ids/targets -> device every step (copy stream, one batch ahead), the step's
    # predictions and loss terms back to pinned host memory every step (a third stream; the outputs of a step
    # live in their own slot, the next step writes another one)
    xh = torch.from_numpy(w.x[: n_batches * B]).pin_memory()
    yh = torch.from_numpy(w.y[: n_batches * B]).pin_memory()
    Ke = min(K, 200)
    xd = [torch.empty((B, F), dtype=torch.int64, device=device) for _ in range(2)]
    yd = [torch.empty(B, dtype=torch.float32, device=device) for _ in range(2)]
    res = torch.empty((Ke, 16), dtype=torch.float32).pin_memory()
    predh = torch.empty((4, B), dtype=torch.float32).pin_memory()
    h2d, d2h = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
    up = [torch.cuda.Event() for _ in range(2)]               # batch uploaded into xd/yd[k]
    used = [torch.cuda.Event() for _ in range(2)]             # xd/yd[k] consumed by the step's staging copy
    outs_read = [torch.cuda.Event() for _ in range(4)]
    cur = torch.cuda.current_stream(device)

    def upload(i):
        j = ((W + K + i) * world + rank) % n_batches
        k = i % 2
        h2d.wait_event(used[k])
        with torch.cuda.stream(h2d):
            xd[k].copy_(xh[j * B:(j + 1) * B], non_blocking=True)
            yd[k].copy_(yh[j * B:(j + 1) * B], non_blocking=True)
            up[k].record(h2d)

    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    upload(0)
    for i in range(Ke):
        k = i % 2
        if i + 1 < Ke:
            upload(i + 1)
        cur.wait_event(up[k])
        cur.wait_event(outs_read[i % 4])                      # (generous: a slot's outputs are rewritten 3 steps later)
        if pipe is None:
            cur.wait_stream(d2h)                              # unpipelined steps share one output slot
        o = pipe.step(xd[k], yd[k]) if pipe is not None else step(xd[k], yd[k])
        used[k].record(cur)
        done = torch.cuda.Event()
        done.record(cur)
        d2h.wait_event(done)
        with torch.cuda.stream(d2h):
            res[i].copy_(o["stats"][:16], non_blocking=True)
            predh[i % 4].copy_(o["pred"], non_blocking=True)
            outs_read[(i + 3) % 4].record(d2h)
    cur.wait_stream(d2h)
    e1.record()
    barrier()
    ms_e = e0.elapsed_time(e1)

    # where the step time goes: CUDA events between the phases, over a few extra steps
    from vae_b200 import _lib as L
    model.enable_timing(True)
    n0 = int(L.lib().vfmb_launch_count())
    for i in range(W + K, W + K + 20):
        model.step(*batch(i))
    launches_per_step = (int(L.lib().vfmb_launch_count()) - n0) / 20
    phases = {k: round(v, 4) for k, v in model.phase_times().items()}
    model.enable_timing(False)

    # unique rows of the GLOBAL batches (outside the timed region) for the algorithmic bytes
    us = []
    for i in range(W, W + min(K, 16)):
        xb = batch(i)[0].reshape(-1)
        allx = [torch.empty_like(xb) for _ in range(world)]
        dist.all_gather(allx, xb)
        us.append(int(torch.unique(torch.cat(allx)).numel()))
    U = float(np.mean(us))
    t = torch.tensor([ms, ms_e], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e = float(t[0].item()), float(t[1].item())
    try:
        parity = sharded_parity(args, w, tc, lr, kl, exchange, device, rank, world)
    except Exception as exc:                                   # never lose the timing line to the check
        parity = {"error": f"{type(exc).__name__}: {exc}"}
    if rank == 0:
        peak, peak_src = peaks()
        # per step: every rank reads its ids/targets and writes predictions; every touched row of the
        # global batch is read and written once by its owner (parameters + both moments)
        step_bytes = world * B * (8 * F + 8) + U * ((2 * d + 2) * 4 * 6 + 8)
        a2a = model.M * (2 * 4 + 2 * (d + 1) * 4)
        out_json = {
            "metric": METRIC, "value": K * B * world / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{w.name}: {'+'.join(map(str, w.field_sizes))} rows, d={d}, batch {B} per GPU, "
                                   f"{w.variant} ELBO, {w.output}",
                       "fields": F, "unique_rows_per_global_step": U, "adam": "touched rows (lazy), on the owner",
                       "noise": "Philox4x32-10 in-kernel, keyed by the global row id (rank-invariant)",
                       "plan": "built every step on the step's stream (requester and owner side)"
                               + ("; plans and id exchange of batch i+1 run under step i (side branch of the step's "
                                  "CUDA graph)" if pipe is not None else
                                  "; whole step incl. the collectives replayed as one CUDA graph" if graphed else ""),
                       "l2": "consecutive distinct batches, no flush",
                       "parallelism": (f"sharded{world}: rows r mod {world}; per step 3 exchanges (ids, sampled rows, row "
                                       f"gradients; {a2a / 1e6:.1f} MB of slots per rank) "
                                       + ("written by the step kernels themselves straight into the peers' buffers over "
                                          "NVLink (symmetric memory) and read in place, 3 signal-pad barriers, no collective calls"
                                          if exchange == "peer" else "as NCCL all-to-alls + 2 small all-reduces")
                                       + f"; global batch {B * world}")},
            "roofline": {"bound": "hbm", "kernel": "whole step (all ranks)", "achieved": step_bytes / (ms / K * 1e-3) / 1e9,
                         "peak": peak * world, "unit": "GB/s", "frac": step_bytes / (ms / K * 1e-3) / 1e9 / (peak * world),
                         "traffic": None, "peak_source": peak_src + f" x {world} GPUs", "algorithmic_bytes": step_bytes},
            "e2e": {"value": Ke * B * world / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * F * 8 + B * 4,
                    "d2h_bytes_per_step": 16 * 4 + B * 4, "steps": Ke, "ms_per_step": ms_e / Ke},
            # own kernels per rank and step, counted by the library over a few eager steps (the barriers are
            # torch symmetric-memory kernels / the collectives NCCL and are not counted)
            "gpu_launches": launches_per_step * K * world, "gpu_launches_per_step_per_rank": launches_per_step,
            "clocks": clocks, "final_loss": loss, "phase_ms": phases, "parity": parity,
        }
        print(json.dumps(out_json), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    if graphed:
        # tearing down a communicator whose collectives live in a captured graph hung at exit
        # (observed once at N=2): results are printed, leave without the destructor chain
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)
    dist.destroy_process_group()


def dp_parity(args, w, device, rank, world):
    """Real-rank parity of mode A (outside the timed region): one data-parallel step (replicated tables, every
    rank its slice of the global batch, dense all-reduce, touched-rows Adam) against the single-process fused
    step on the concatenated batch from the same seeded parameters (noise is keyed by the global row id)."""
    import torch.distributed as dist
    from vae_b200.dist import DataParallelSampled
    B = w.batch
    tc = w.train_counts()
    tc[tc == 0] = 1
    lr = learning_rate(w)
    models = []
    for max_batch in (B, B * world):
        torch.manual_seed(synth.PARAM_SEED)
        models.append(make_model(w, device, tc, lr, max_batch=max_batch))
    rep, ref = models
    dpm = DataParallelSampled(rep, world, dense_adam=False)
    out = dpm.step(torch.from_numpy(w.x[rank * B:(rank + 1) * B]).to(device), torch.from_numpy(w.y[rank * B:(rank + 1) * B]).to(device))
    torch.cuda.synchronize()
    res = None
    if rank == 0:
        r = ref.fused_step(torch.from_numpy(w.x[: world * B]).to(device), torch.from_numpy(w.y[: world * B]).to(device))
        torch.cuda.synchronize()
        want, got = ref.entity_params.weight.detach(), rep.entity_params.weight.detach()
        err = (got - want).abs()
        tol = 1e-5 * want.abs() + 1e-4 * lr
        res = {"what": f"one mode-A step on {world} ranks vs the single-process fused step on the global batch "
                       f"({world * B} samples), same seeded parameters, in-kernel Philox noise",
               "loss_rel_err": abs(out["loss"].item() - r["loss"].item()) / abs(r["loss"].item()),
               "params_max_err_over_lr": float((err.max() / lr).item()),
               "params_fraction_outside_1e-5rel+1e-4lr": float((err > tol).float().mean().item()),
               "bias_max_err_over_lr": float(((rep.bias_params.weight.detach() - ref.bias_params.weight.detach()).abs().max() / lr).item())}
    dist.barrier()
    return res


def sharded_parity(args, w, tc, lr, kl, exchange, device, rank, world):
    """Real-rank parity of mode B (outside the timed region): one sharded step on the global batch (the
    concatenation of the ranks' first batches, in-kernel Philox noise) against the single-process fused step
    on that batch from the same seeded parameters.  Noise is keyed by the global row id, so both draw the
    same values.  Tables that do not fit one GPU are skipped."""
    import torch.distributed as dist
    from vae_b200.dist import ShardedSampled
    if w.rows > 5_000_000:
        return {"skipped": "table does not fit one GPU for the single-process reference step"}
    B, d = w.batch, w.d
    torch.manual_seed(synth.PARAM_SEED)
    from vae_b200.vfm_torch import CF
    ref = CF(d, output=w.output, n_users=w.field_sizes[0], n_items=w.field_sizes[1], train_counts=torch.from_numpy(tc),
             field_sizes=w.field_sizes, kl_weighting=kl, n_train=w.n_train, max_batch=B * world, seed=synth.NOISE_SEED,
             lr=lr, device=device)                              # same seed on every rank: same initial parameters
    init = {"bias": ref.bias_params.weight.detach().cpu(), "entity": ref.entity_params.weight.detach().cpu(),
            "alpha": float(ref.alpha.item()), "global_bias_mean": float(ref.global_bias_mean.item()),
            "global_bias_scale": float(ref.global_bias_scale.item())}
    sh = ShardedSampled(d, w.field_sizes, torch.from_numpy(tc), w.n_train, B, world, rank, output=w.output,
                        kl_weighting=kl, seed=synth.NOISE_SEED, lr=lr, device=device, slack=args.slack,
                        exchange="peer" if exchange == "peer" else None, init=init)
    xl = torch.from_numpy(w.x[rank * B:(rank + 1) * B]).to(device)
    yl = torch.from_numpy(w.y[rank * B:(rank + 1) * B]).to(device)
    out = sh.step(xl, yl)
    torch.cuda.synchronize()
    sh.check_overflow()
    gid, bias_l, ent_l = sh.gather_tables()
    n_max = (w.rows + world - 1) // world
    pad = lambda t: torch.cat((t, torch.zeros((n_max - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=device)))
    ents = [torch.empty((n_max, 2 * d), device=device) for _ in range(world)]
    biases = [torch.empty((n_max, 2), device=device) for _ in range(world)]
    dist.all_gather(ents, pad(ent_l))
    dist.all_gather(biases, pad(bias_l))
    res = None
    if rank == 0:
        xg = torch.from_numpy(w.x[: world * B]).to(device)
        yg = torch.from_numpy(w.y[: world * B]).to(device)
        r = ref.fused_step(xg, yg)
        torch.cuda.synchronize()
        ent = torch.empty((w.rows, 2 * d), device=device)
        bia = torch.empty((w.rows, 2), device=device)
        for q in range(world):
            n = len(range(q, w.rows, world))
            ent[q::world], bia[q::world] = ents[q][:n], biases[q][:n]
        want = ref.entity_params.weight.detach()
        err = (ent - want).abs()
        tol = 1e-5 * want.abs() + 1e-4 * lr
        res = {"what": f"one mode-B step on {world} ranks vs the single-process fused step on the global batch "
                       f"({world * B} samples), same seeded parameters, in-kernel Philox noise",
               "loss_rel_err": abs(out["loss"].item() - r["loss"].item()) / abs(r["loss"].item()),
               "params_max_err_over_lr": float((err.max() / lr).item()),
               "params_fraction_outside_1e-5rel+1e-4lr": float((err > tol).float().mean().item()),
               "bias_max_err_over_lr": float(((bia - ref.bias_params.weight.detach()).abs().max() / lr).item()),
               "scalars_max_abs_err": float((sh.scalars[:3] - ref._scalars[:3]).abs().max().item())}
    dist.barrier()
    del sh, ref
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------
def cpu_baseline(args, budget_s=20.0):
    """The oracle port (torch CPU, same ATen ops as the reference: unique, embedding / indexing,
    distributions, autograd, dense Adam) on this box's host cores, on a bounded number of full-size steps."""
    w = synth.make_workload(args.workload, n_rows=args.rows)
    if w.rows > 20_000_000:                                   # dense grads + dense Adam cannot hold 100M rows
        return {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                "sample": "skipped: dense reference cannot allocate this table"}
    B = w.batch
    tc = w.train_counts()
    tc[tc == 0] = 1
    _, _, step = make_port(w, tc, learning_rate(w))
    x, y = torch.from_numpy(w.x), torch.from_numpy(w.y)
    n_batches = w.n_train // B
    times = []
    t_start = time.time()
    i = 0
    while True:
        j = i % n_batches
        t0 = time.perf_counter()
        step(x[j * B:(j + 1) * B], y[j * B:(j + 1) * B])
        dt = time.perf_counter() - t0
        if i >= 2:
            times.append(dt)
        i += 1
        if len(times) >= 5 and time.time() - t_start > budget_s:
            break
        if len(times) >= 200:
            break
    med = float(np.median(times))
    return {"value": B / med, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "host_cpus": os.cpu_count(), "torch": torch.__version__,
            "sample": f"{len(times)} full-size steps (batch {B}) after 2 warm-ups, median step {med * 1e3:.1f} ms"}


def parity_vs_port(args, device):
    """Part of the CPU leg (the only place bench.py runs the oracle): one training step of the timed
    configuration -- same code path as the timed steps (fused step; in-kernel Philox noise for the sampled
    variant), first batch, fresh parameters -- against the torch port of the reference (fed the exported
    Philox draws)."""
    w = synth.make_workload(args.workload, n_rows=args.rows)
    if w.rows > 20_000_000:
        return {"skipped": "dense reference cannot allocate this table"}
    B = w.batch
    tc = w.train_counts()
    tc[tc == 0] = 1
    lr = learning_rate(w)
    model = make_model(w, device, tc, lr)
    port, _, step = make_port(w, tc, lr)
    own = port.state_dict()
    port.load_state_dict({k: v.detach().cpu().reshape(own[k].shape) for k, v in model.state_dict().items() if k in own},
                         strict=False)
    x, y = torch.from_numpy(w.x[:B]), torch.from_numpy(w.y[:B])
    uniq = torch.unique(x)
    noise = [n.cpu() for n in model.philox_noise(uniq)] if w.variant == "sampled" else None
    res = model.fused_step(x.to(device), y.to(device))
    torch.cuda.synchronize()
    po = step(x, y, noise)
    pred, want = res["pred"].cpu().numpy().astype(np.float64), po["pred"].numpy().astype(np.float64)
    out = {"oracle": "oracle/vfm_port.py (fp32 torch restatement of the reference step)"
                     + (", same Philox draws injected" if noise is not None else ""),
           "pred_max_err_over_rms": float(np.max(np.abs(pred - want)) / np.sqrt(np.mean(want ** 2))),
           "loss_rel_err": float(abs(res["loss"].item() - po["loss"].item()) / abs(po["loss"].item()))}
    u = uniq.numpy()
    ent = "entity_params" if w.variant == "closed" else "entity_params.weight"
    got = dict(model.named_parameters())[ent].detach().cpu().numpy()[u].astype(np.float64)
    ref = dict(port.named_parameters())[ent].detach().numpy()[u].astype(np.float64)
    err = np.abs(got - ref)
    outside = err > 1e-5 * np.abs(ref) + 1e-4 * lr
    out.update({"params_rows_checked": int(len(u)), "params_max_err_over_lr": float(err.max() / lr),
                "params_fraction_outside_1e-5rel+1e-4lr": float(outside.mean()),
                "note": "first Adam step from zero moments moves every element by ~lr*sign(g): elements outside the "
                        "bound are those whose gradient is ~0 in fp32 (tests/test_gpu_bench_shapes.py proves it per element)"})
    return out


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path.  The reference is pure
    Python/torch and is not present on the GPU box, so this times the oracle port
    (oracle/vfm_port.py: same ATen op sequence), with all host threads torch will use."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)               # torchrun exports OMP_NUM_THREADS=1: use every host core
    w = synth.make_workload(args.workload, n_rows=args.rows)
    t0 = time.time()
    cb = cpu_baseline(args, budget_s=max(20.0, min(150.0, 0.05 * args.steps)))
    wall = time.time() - t0
    v = cb["value"]
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": (w.batch / v * 1e3) if v else None, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{w.name}: {'+'.join(map(str, w.field_sizes))} rows, d={w.d}, "
                                  f"{w.n_train} ratings, batch {w.batch}, {w.variant} ELBO, {w.output}",
                      "fields": w.n_fields, "adam": "dense (torch.optim.Adam)", "wall_s": wall,
                      "implementation": "oracle/vfm_port.py: the reference's step restated with the same ATen calls "
                                        "(== the AST-sliced reference to 1e-12, tests/test_oracle_vs_reference.py); "
                                        "the reference scripts themselves cannot be imported (they train at import time)"},
           "cpu_baseline": cb,
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ml20m")
    ap.add_argument("--rows", type=int, default=None, help="override the synthetic dataset size")
    ap.add_argument("--plan", default="graph", choices=["inline", "prefetch", "graph", "cached"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--reserve", type=int, default=1, help="graph mode: block slots per SM the step kernels leave to the plan")
    ap.add_argument("--side-priority", type=int, default=-1, help="graph mode: stream priority of the plan branch")
    ap.add_argument("--tune", action="append", default=[], metavar="KEY=VALUE",
                    help="library launch knob (vfmb_set_tuning), e.g. prefetch_mv=3; repeatable")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--parallel", default="auto", choices=["auto", "dp", "sharded"],
                    help="N>1: dp = replicated tables + dense all-reduce (mode A), sharded = row-sharded tables + "
                         "all-to-all (mode B); auto = sharded unless the dense gradient is under 8 MB")
    ap.add_argument("--slack", type=float, default=0.75, help="mode B: slot capacity as a fraction of B*F/P")
    ap.add_argument("--no-pipeline", action="store_true", help="mode B peer: serial graphed step instead of the pipeline")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="mode B data path: NVLink peer memory written by the pack kernels, or NCCL all-to-alls")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args)
    elif world > 1 and (args.parallel == "sharded" or (args.parallel == "auto" and args.workload != "ml100k")):
        run_sharded(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
