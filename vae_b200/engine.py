"""Device buffers and call plumbing shared by the two model variants.

Everything here is host-side glue around the C ABI: PyTorch owns the device
memory and the streams, the kernels in ``csrc/`` do the work.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib as L


def _i32(n, device):
    return torch.empty(int(n), dtype=torch.int32, device=device)


def _f32(n, device):
    return torch.empty(int(n), dtype=torch.float32, device=device)


class BatchPlan:
    """Device arrays of one batch plan (the ``torch.unique`` outputs of
    vfm-torch.py:190-192 / vfm-tomasrch.py:537-545 plus the sorted occurrence
    segments the backward walks).  Integer contents are bit-identical to
    ``torch.unique(sorted=True, return_inverse=True, return_counts=True)``."""

    def __init__(self, B: int, F: int, R: int, device):
        cap = L.PlanCapacity()
        L.check(L.lib().vfmb_plan_capacity(B, F, R, C.byref(cap)), "vfmb_plan_capacity")
        self.B_cap, self.F, self.R = B, F, R
        self.u_cap, self.n_tiles, self.tile = int(cap.u_cap), int(cap.n_tiles), int(cap.tile)
        self.uniq = _i32(self.u_cap, device)
        self.inverse = _i32(B * F, device)
        self.seg_off = _i32(self.u_cap + 1, device)
        self.occ = _i32(B * F, device)
        self.pos_of = _i32(B * F, device)
        self.pos_rank = _i32(B * F, device)
        self.partner = _i32(B * F, device)
        self.urec = _i32(4 * self.u_cap, device)
        self.class_off = torch.zeros(L.MAX_FIELDS + 1, dtype=torch.int32, device=device)
        self.z = torch.zeros(L.MAX_FIELDS, dtype=torch.float32, device=device)
        self.meta = torch.zeros(8, dtype=torch.int32, device=device)
        self.hot = torch.zeros(int(cap.cut_rows_cap), dtype=torch.int32, device=device)
        self.workspace = torch.empty(int(cap.workspace_bytes), dtype=torch.uint8, device=device)
        self.B = 0
        self.struct = L.Plan(L.ptr(self.uniq), L.ptr(self.inverse), L.ptr(self.seg_off),
                             L.ptr(self.occ), L.ptr(self.pos_of), L.ptr(self.pos_rank), L.ptr(self.partner),
                             L.ptr(self.urec), L.ptr(self.class_off), L.ptr(self.z), L.ptr(self.meta),
                             L.ptr(self.hot))

    def build(self, cfg: L.Config, x: torch.Tensor, train_counts: torch.Tensor) -> "BatchPlan":
        assert x.dtype == torch.int64 and x.is_cuda and x.is_contiguous(), "x: contiguous CUDA int64 [B,F]"
        assert x.shape[0] <= self.B_cap and x.shape[1] == self.F
        self.B = int(x.shape[0])
        stream = torch.cuda.current_stream(x.device).cuda_stream
        L.check(L.lib().vfmb_plan_build(C.byref(cfg), x.data_ptr(), train_counts.data_ptr(),
                                        C.byref(self.struct), self.workspace.data_ptr(),
                                        self.workspace.numel(), stream), "vfmb_plan_build")
        return self

    # -- host views (synchronise; for tests and the drop-in API only)
    def num_unique(self) -> int:
        return int(self.meta[0].item())

    def check_ids(self) -> None:
        if int(self.meta[2].item()) != 0:
            raise IndexError("row id out of range in batch")

    def as_unique(self):
        """(uniq int64 [U], inverse int64 [B,F], counts int64 [U]) like torch.unique."""
        U = self.num_unique()
        uniq = self.uniq[:U].long()
        inverse = self.inverse[: self.B * self.F].long().reshape(self.B, self.F)
        counts = (self.seg_off[1:U + 1] - self.seg_off[:U]).long()
        return uniq, inverse, counts


class PlanPipeline:
    """Batch plans for a training loop.

    The plan of a batch depends only on its ids, never on the parameters, so it can be
    built ahead of the step that consumes it: ``prefetch(x_next)`` enqueues the plan
    kernels on a side stream, where they overlap the (HBM-bound) kernels of the current
    step; ``acquire(x)`` returns that plan (making the current stream wait for it) or
    builds one in line.  Plans of a never-shuffled loader (vfm-torch.py:121-122) can be
    kept: ``build_static`` returns a plan object the caller may pass back every epoch."""

    def __init__(self, B_cap: int, F: int, R: int, device, depth: int = 2):
        self.B_cap, self.F, self.R, self.device = B_cap, F, R, device
        self.ring = [BatchPlan(B_cap, F, R, device) for _ in range(depth)]
        self.free_evt = [torch.cuda.Event() for _ in range(depth)]
        self.ready_evt = [torch.cuda.Event() for _ in range(depth)]
        self.slot_of = {}                                   # id(plan) -> ring slot
        for i, p in enumerate(self.ring):
            self.slot_of[id(p)] = i
        self.next = 0
        self.pending = {}                                   # (data_ptr, B) -> slot
        self.stream = torch.cuda.Stream(device=device, priority=-1)   # tiny latency-bound kernels first
        self.u_cap, self.n_tiles = self.ring[0].u_cap, self.ring[0].n_tiles

    def _take_slot(self) -> int:
        slot = self.next
        self.next = (self.next + 1) % len(self.ring)
        for k in [k for k, v in self.pending.items() if v == slot]:
            del self.pending[k]                             # an unused prefetch is overwritten
        return slot

    def prefetch(self, cfg: L.Config, x: torch.Tensor, train_counts: torch.Tensor, after=None) -> None:
        """``after``: optional event the ids become valid at (e.g. their host-to-device copy)."""
        key = (x.data_ptr(), int(x.shape[0]))
        if key in self.pending:
            return
        slot = self._take_slot()
        self.stream.wait_event(self.free_evt[slot])          # last consumer of this buffer is done
        if after is not None:
            self.stream.wait_event(after)
        with torch.cuda.stream(self.stream):
            self.ring[slot].build(cfg, x, train_counts)
            self.ready_evt[slot].record(self.stream)
        self.pending[key] = slot

    def acquire(self, cfg: L.Config, x: torch.Tensor, train_counts: torch.Tensor) -> BatchPlan:
        key = (x.data_ptr(), int(x.shape[0]))
        cur = torch.cuda.current_stream(self.device)
        slot = self.pending.pop(key, None)
        if slot is not None:
            cur.wait_event(self.ready_evt[slot])
            self.ring[slot].B = int(x.shape[0])
            return self.ring[slot]
        slot = self._take_slot()
        cur.wait_event(self.free_evt[slot])
        return self.ring[slot].build(cfg, x, train_counts)

    def release(self, plan: BatchPlan) -> None:
        """Call after the last kernel that reads ``plan`` was enqueued on the current stream."""
        slot = self.slot_of.get(id(plan))
        if slot is not None:
            self.free_evt[slot].record(torch.cuda.current_stream(self.device))

    def build_static(self, cfg: L.Config, x: torch.Tensor, train_counts: torch.Tensor) -> BatchPlan:
        """A plan outside the ring (kept by the caller, e.g. one per batch of a fixed dataset)."""
        plan = BatchPlan(int(x.shape[0]), self.F, self.R, self.device)
        plan.u_cap_shared = self.u_cap
        return plan.build(cfg, x, train_counts)


class GraphedLoop:
    """CUDA-graph replay of the software-pipelined training loop (fixed batch size).

    Launching the ~16 small kernels of a step from Python costs more host time than the GPU
    needs to run them; the step is launch-bound.  All per-step state (Adam step counter, Philox
    counter, reduction counters) lives on the device and the kernels take no per-step host
    scalars, so a step can be captured once and replayed.  One graph per staging slot s holds:

        main stream   forward + backward + Adam on plan[s], targets ys[s]
        side stream   plan of the NEXT batch from xs[s+1] into plan[s+1]   (forked / joined)

    Batches are consumed in the order they are staged.  ``depth`` = number of staging slots:
    2 when the batches are already on the device (staging copies run on the main stream);
    3 when they come from pinned host memory -- the host-to-device copy of batch i+2 then runs
    on a copy stream while step i executes.  Usage (depth 2)::

        loop = model.graphed_loop(B)
        loop.start(x0, y0)
        for x_next, y_next in batches[1:]:
            res = loop.step(x_next, y_next)      # runs the step on the oldest staged batch
        res = loop.step()                         # last staged batch
    """

    def __init__(self, model, B: int, step_fn, depth: int = 2, reserve: int = 1, side_priority: int = -1,
                 plan_in_graph: bool = True):
        # plan_in_graph=False (measurement only, scripts/graph_overheads.py): the graphs hold the step alone and
        # replay whatever plans the staging slots were left with
        assert depth in (2, 3)
        self.model, self.B, self.device, self.depth = model, int(B), model.device, depth
        F = model.F if hasattr(model, "F") else model.G
        D = depth
        self.xs = [torch.zeros((B, F), dtype=torch.int64, device=self.device) for _ in range(D)]
        self.ys = [torch.zeros(B, dtype=torch.float32, device=self.device) for _ in range(D)]
        self.plans = [BatchPlan(B, F, model.R, self.device) for _ in range(D)]
        # outputs per staging slot (predictions, logits, loss terms): step i+1 writes another slot, so a
        # caller can copy step i's outputs to the host on its own stream while step i+1 runs
        S = int(getattr(model, "S", 1))
        self.outs = [SlotOutputs(S, B, self.device) for _ in range(D)]
        # high priority: the plan kernels are tiny and latency-bound; they should slip in as soon
        # as blocks of the (machine-filling) step kernels retire
        self.side = torch.cuda.Stream(device=self.device, priority=side_priority)
        self.copy = torch.cuda.Stream(device=self.device) if depth == 3 else None
        self.copied = [torch.cuda.Event() for _ in range(D)]
        self.done = [torch.cuda.Event() for _ in range(D)]
        self.cfg = model._config(B)
        self.graphs = []
        self.head = 0                                       # slot of the next batch to run
        self.tail = 0                                       # slot the next staged batch goes to
        self.n_staged = 0
        # warm-up on a side stream (lazy allocations, module loading), then capture
        warm = torch.cuda.Stream(device=self.device)
        warm.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(warm):
            for s in range(D):
                self.plans[s].build(self.cfg, self.xs[s], model.train_counts)
        torch.cuda.current_stream(self.device).wait_stream(warm)
        torch.cuda.synchronize(self.device)
        # the step kernels are captured one block slot per SM short of a full wave, so that the
        # plan's blocks (side branch) are placed immediately instead of displacing step blocks
        L.check(L.lib().vfmb_set_grid_reserve(reserve), "vfmb_set_grid_reserve")
        n0 = int(L.lib().vfmb_launch_count())
        try:
            for s in range(D):
                nxt = (s + 1) % D
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    cur = torch.cuda.current_stream(self.device)
                    if plan_in_graph:
                        self.side.wait_stream(cur)               # fork
                        with torch.cuda.stream(self.side):
                            self.plans[nxt].build(self.cfg, self.xs[nxt], model.train_counts)
                    step_fn(self.plans[s], self.ys[s], self.outs[s])   # main branch
                    if plan_in_graph:
                        cur.wait_stream(self.side)               # join
                self.graphs.append(g)
        finally:
            L.lib().vfmb_set_grid_reserve(0)
        # own kernel nodes per captured graph (the library counts every launch it makes)
        self.launches_per_graph = (int(L.lib().vfmb_launch_count()) - n0) // D

    def stage(self, x: torch.Tensor, y: torch.Tensor) -> None:
        """Copy a batch (device or pinned host) into the next free staging slot."""
        assert self.n_staged < self.depth, "all staging slots are in use; call step() first"
        t = self.tail
        if self.copy is None:
            self.xs[t].copy_(x, non_blocking=True)
            self.ys[t].copy_(y, non_blocking=True)
        else:
            self.copy.wait_event(self.done[t])               # the last step that read this slot
            with torch.cuda.stream(self.copy):
                self.xs[t].copy_(x, non_blocking=True)
                self.ys[t].copy_(y, non_blocking=True)
                self.copied[t].record(self.copy)
        self.tail = (t + 1) % self.depth
        self.n_staged += 1

    def start(self, x: torch.Tensor, y: torch.Tensor) -> None:
        """Stage the first batch and build its plan (outside the graphs)."""
        self.head = self.tail = self.n_staged = 0
        self.stage(x, y)
        cur = torch.cuda.current_stream(self.device)
        if self.copy is not None:
            cur.wait_event(self.copied[0])
        self.plans[0].build(self.cfg, self.xs[0], self.model.train_counts)

    def step(self, x_next: Optional[torch.Tensor] = None, y_next: Optional[torch.Tensor] = None):
        """Run the step on the oldest staged batch; the plan of the one staged after it is built
        concurrently.  ``x_next, y_next`` (optional) are staged first.  Returns a ``StepResult``."""
        if x_next is not None:
            self.stage(x_next, y_next)
        assert self.n_staged > 0, "nothing staged: call start() / stage() first"
        s = self.head
        cur = torch.cuda.current_stream(self.device)
        if self.copy is not None:                            # the side branch reads slot s+1
            cur.wait_event(self.copied[(s + 1) % self.depth])
        self.model._sync_scalars()
        self.graphs[s].replay()
        self.done[s].record(cur)
        self.head = (s + 1) % self.depth
        self.n_staged -= 1
        self.last_slot = s
        return StepResult(self.outs[s], self.B)


class SlotOutputs:
    """Outputs of the step captured in one graph of a GraphedLoop (duck-types StepBuffers for StepResult)."""

    def __init__(self, S: int, B: int, device):
        self.S = S
        self.pred = _f32(S * B, device)
        self.mean = _f32(S * B, device)
        self.stats = torch.zeros(L.STATS, dtype=torch.float32, device=device)


class StepBuffers:
    """Scratch and outputs of one step for batches up to ``B`` samples."""

    def __init__(self, cfg: L.Config, plan: BatchPlan, device, n_scalars: int, need_msg: bool,
                 closed: bool = False):
        B, d = plan.B_cap, cfg.d
        wv, wg = (2, 3) if closed else (1, 1)              # closed form: vs=[mu|rho^2], grow=[A|B|C]
        S = 1 if closed else max(1, int(cfg.S))            # S > 1: [S][u_cap] sampled rows / noise / gradients
        self.S = S
        self.vs = _f32(S * plan.u_cap * d * wv, device)
        self.ws = _f32(S * plan.u_cap, device)
        self.es = _f32(S * plan.u_cap * d, device)
        self.ebs = _f32(S * plan.u_cap, device)
        self.cq = _f32(plan.u_cap, device)
        self.grow = _f32(S * plan.u_cap * d * wg, device)
        self.gws = _f32(S * plan.u_cap, device)
        self.msg = _f32(S * B * d * wg, device) if need_msg else None
        self.pred = _f32(S * B, device)                    # [S, B] (likelihood batch shape, vfm-torch.py:265)
        self.mean = _f32(S * B, device)
        self.resid = _f32(B, device)
        self.rsorted = _f32(B * plan.F, device)
        n_part = int(L.lib().vfmb_partials_doubles(C.byref(cfg)))
        self.partials = torch.zeros(n_part, dtype=torch.float64, device=device)
        self.counters = torch.zeros(8, dtype=torch.int32, device=device)
        self.stats = torch.zeros(L.STATS, dtype=torch.float32, device=device)
        self.grad_scalars = torch.zeros(n_scalars, dtype=torch.float32, device=device)

    def io(self, y=None, noise: Optional[Sequence[torch.Tensor]] = None, grad_bias=None,
           grad_entity=None, resid=None, kl_bias_out=None, kl_entity_out=None) -> L.StepIO:
        e0 = eb = ee = None
        if noise is not None:
            e0, eb, ee = noise
            for t in (e0, eb, ee):
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        r = self.resid if resid is None else resid
        return L.StepIO(L.ptr(y), L.ptr(e0), L.ptr(eb), L.ptr(ee), L.ptr(self.vs), L.ptr(self.ws),
                        L.ptr(self.es), L.ptr(self.ebs), L.ptr(self.cq), L.ptr(self.grow), L.ptr(self.gws),
                        L.ptr(self.msg), L.ptr(self.pred), L.ptr(self.mean), L.ptr(r),
                        L.ptr(self.rsorted), L.ptr(self.partials), L.ptr(self.counters), L.ptr(self.stats),
                        L.ptr(kl_bias_out), L.ptr(kl_entity_out),
                        L.ptr(grad_bias), L.ptr(grad_entity), L.ptr(self.grad_scalars))


class StepResult:
    """Outputs of a fused step: lazy views of the (reused) device buffers, so that the hot loop
    does not pay for tensor slicing it never looks at.  ``res["loss"]`` etc. work like a dict."""
    __slots__ = ("_buf", "_B")
    _STAT = {"loss": L.ST_LOSS, "kl": L.ST_KL, "nll_mean": L.ST_NLL_MEAN}

    def __init__(self, buf: "StepBuffers", B: int):
        self._buf, self._B = buf, B

    def __getitem__(self, key: str):
        if key in self._STAT:
            return self._buf.stats[self._STAT[key]]
        if key == "stats":
            return self._buf.stats
        if key == "pred":
            return self._buf.mean[: self._B] if self._buf.S == 1 else self._buf.mean[: self._buf.S * self._B].view(self._buf.S, self._B)
        if key == "logits":
            return self._buf.pred[: self._B] if self._buf.S == 1 else self._buf.pred[: self._buf.S * self._B].view(self._buf.S, self._B)
        raise KeyError(key)

    def keys(self):
        return ["loss", "kl", "nll_mean", "pred", "logits", "stats"]


def make_config(B, F, d, R, S, likelihood, link, class_bounds, class_sizes, n_train, seed,
                interaction: str = "prod") -> L.Config:
    cfg = L.Config()
    cfg.interaction = {"prod": L.INTER_PROD, "pairwise": L.INTER_PAIRWISE}[interaction]
    cfg.B, cfg.F, cfg.d, cfg.R, cfg.S = int(B), int(F), int(d), int(R), int(S)
    cfg.likelihood = L.GAUSSIAN if likelihood == "reg" else L.BERNOULLI
    cfg.link = {"abs": L.LINK_ABS, "softplus": L.LINK_SOFTPLUS}[link]
    cfg.n_classes = len(class_sizes)
    for i, b in enumerate(class_bounds):
        cfg.class_bound[i] = int(b)
    for i, s in enumerate(class_sizes):
        cfg.class_size[i] = float(s)
    cfg.n_train = float(n_train)
    cfg.seed = int(seed)
    return cfg


def current_stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(device) -> torch.device:
    device = torch.device(device)
    if device.type != "cuda" or not torch.cuda.is_available():
        raise RuntimeError("vae_b200 runs on CUDA devices only (sm_100a kernels, no CPU fallback)")
    return device
