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
        self.z = torch.zeros(L.MAX_FIELDS, dtype=torch.float32, device=device)
        self.meta = torch.zeros(8, dtype=torch.int32, device=device)
        self.workspace = torch.empty(int(cap.workspace_bytes), dtype=torch.uint8, device=device)
        self.B = 0
        self.struct = L.Plan(L.ptr(self.uniq), L.ptr(self.inverse), L.ptr(self.seg_off),
                             L.ptr(self.occ), L.ptr(self.pos_of), L.ptr(self.pos_rank), L.ptr(self.partner),
                             L.ptr(self.urec), L.ptr(self.z), L.ptr(self.meta))

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


class StepBuffers:
    """Scratch and outputs of one step for batches up to ``B`` samples."""

    def __init__(self, cfg: L.Config, plan: BatchPlan, device, n_scalars: int, need_msg: bool):
        B, d = plan.B_cap, cfg.d
        self.vs = _f32(plan.u_cap * d, device)
        self.ws = _f32(plan.u_cap, device)
        self.es = _f32(plan.u_cap * d, device)
        self.ebs = _f32(plan.u_cap, device)
        self.cq = _f32(plan.u_cap, device)
        self.grow = _f32(plan.u_cap * d, device)
        self.gws = _f32(plan.u_cap, device)
        self.msg = _f32(B * d, device) if need_msg else None
        self.pred = _f32(B, device)
        self.mean = _f32(B, device)
        self.resid = _f32(B, device)
        self.rsorted = _f32(B * plan.F, device)
        n_part = int(L.lib().vfmb_partials_doubles(C.byref(cfg)))
        self.partials = torch.zeros(n_part, dtype=torch.float64, device=device)
        self.counters = torch.zeros(8, dtype=torch.int32, device=device)
        self.stats = torch.zeros(L.STATS, dtype=torch.float32, device=device)
        self.grad_scalars = torch.zeros(n_scalars, dtype=torch.float32, device=device)

    def io(self, y=None, noise: Optional[Sequence[torch.Tensor]] = None, grad_bias=None,
           grad_entity=None, resid=None) -> L.StepIO:
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
                        L.ptr(grad_bias), L.ptr(grad_entity), L.ptr(self.grad_scalars))


def make_config(B, F, d, R, S, likelihood, link, class_bounds, class_sizes, n_train, seed) -> L.Config:
    cfg = L.Config()
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
