"""Multi-GPU execution of the VFM step (one process per GPU, ``torch.distributed`` / NCCL).

The reference is single-process, so these modes have no reference counterpart;
their contract is: the step on the GLOBAL batch (the concatenation of the ranks' local batches)
equals the single-process step on that batch -- same per-entity noise (Philox is keyed by row id
and step, never by rank or batch position), same global batch counts in the KL weights.

Mode A -- ``DataParallelSampled``: replicated tables, batch data parallel, ONE dense all-reduce
per step over ``[grad_entity | grad_bias | batch counts | tail]``.  Exact reference semantics
(including dense Adam over every row) at a cost of ``R*(2d+3)*4`` bytes per step, so it is the
mode for small tables (BASELINE configs 1-2).  Large tables need row sharding with an
all-to-all of touched rows (mode B), which is not in this round.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib as L
from .engine import current_stream, make_config

DP_TAIL = 16
T_NLL, T_RESID, T_SQERR = 8, 9, 10


@dataclass(frozen=True)
class FlatLayout:
    """Layout of the all-reduced fp32 buffer: dense entity gradient, dense bias gradient, dense
    per-row batch counts, then a 16-float tail {Z_f (8), sum nll, sum resid, sum sq. err}."""
    R: int
    d: int

    @property
    def n_entity(self) -> int:
        return self.R * 2 * self.d

    @property
    def off_bias(self) -> int:
        return self.n_entity

    @property
    def off_counts(self) -> int:
        return self.off_bias + 2 * self.R

    @property
    def off_tail(self) -> int:
        return self.off_counts + self.R

    @property
    def numel(self) -> int:
        return self.off_tail + DP_TAIL

    def views(self, flat: torch.Tensor):
        return (flat[: self.n_entity].view(self.R, 2 * self.d),
                flat[self.off_bias: self.off_counts].view(self.R, 2),
                flat[self.off_counts: self.off_tail],
                flat[self.off_tail:])


def allreduce_flat(flat: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the packed buffer over the ranks (NCCL on GPUs; any backend in tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def local_slice(n_global: int, rank: int, world: int):
    """Contiguous slice of the global batch owned by ``rank`` (batches are never shuffled)."""
    per = n_global // world
    assert per * world == n_global, "global batch must divide evenly over the ranks"
    return slice(rank * per, (rank + 1) * per)


class DataParallelSampled:
    """Mode A driver around a replicated ``vae_b200.vfm_torch.CF``."""

    def __init__(self, model, world: int, group=None, dense_adam: bool = True):
        self.model, self.world, self.group, self.dense_adam = model, int(world), group, bool(dense_adam)
        self.layout = FlatLayout(model.R, model.d)
        self.flat = torch.zeros(self.layout.numel, dtype=torch.float32, device=model.device)
        self._cfg_local, self._cfg_global = {}, {}

    def _configs(self, B_local: int):
        m = self.model
        if B_local not in self._cfg_local:
            mk = lambda B, n_train: make_config(B, m.F, m.d, m.R, m.S, m.output, m.link_name, m._class_bounds,
                                                m._class_sizes, n_train, m.seed)
            # local kernels scale residuals by n_train_eff / B_local = n_train / B_global
            self._cfg_local[B_local] = mk(B_local, m.n_train / self.world)
            self._cfg_global[B_local] = mk(B_local * self.world, m.n_train)
        return self._cfg_local[B_local], self._cfg_global[B_local]

    @torch.no_grad()
    def local_backward(self, x: torch.Tensor, y: torch.Tensor,
                       noise: Optional[Sequence[torch.Tensor]] = None) -> torch.Tensor:
        """Forward + data-term backward on the local slice; returns the packed buffer to reduce.
        ``noise`` (tests) is indexed by the LOCAL unique rank, as in ``CF.fused_step``."""
        m, lay = self.model, self.layout
        m._sync_scalars()
        x = x.to(m.device).contiguous()
        y = y.to(m.device, torch.float32).contiguous()
        B = int(x.shape[0])
        cfg_l, _ = self._configs(B)
        m._ensure(B)
        m._cfg = cfg_l
        m._plan = m._pipe.acquire(cfg_l, x, m.train_counts)
        noise = m._prep_noise(noise)
        self.flat.zero_()
        g_entity, g_bias, counts, tail = lay.views(self.flat)
        io = m._buf.io(y=y, noise=noise, grad_bias=g_bias, grad_entity=g_entity)
        io.grad_scalars = None
        tab, s, lib = m._tables(), current_stream(m.device), L.lib()
        L.check(lib.vfmb_sampled_forward(C.byref(cfg_l), C.byref(tab), C.byref(m._plan.struct), C.byref(io), s),
                "vfmb_sampled_forward")
        L.check(lib.vfmb_sampled_backward(C.byref(cfg_l), C.byref(tab), C.byref(m._plan.struct), C.byref(io),
                                          C.byref(m.adam), L.GRAD_ONLY, 0.0, s), "vfmb_sampled_backward")
        L.check(lib.vfmb_dp_scatter_counts(C.byref(cfg_l), C.byref(m._plan.struct), C.byref(io),
                                           counts.data_ptr(), tail.data_ptr(), s), "vfmb_dp_scatter_counts")
        m._pipe.release(m._plan)
        self._B_local = B
        self._eps0 = noise[0] if noise is not None else None
        return self.flat

    @torch.no_grad()
    def apply(self, flat: torch.Tensor) -> dict:
        """KL gradient with global counts + Adam + scalar update from the reduced buffer."""
        m, lay = self.model, self.layout
        _, cfg_g = self._configs(self._B_local)
        g_entity, g_bias, counts, tail = lay.views(flat)
        L.check(L.lib().vfmb_dp_apply_sampled(C.byref(cfg_g), C.byref(m._tables()), g_entity.data_ptr(),
                                              g_bias.data_ptr(), counts.data_ptr(), tail.data_ptr(),
                                              L.ptr(self._eps0), C.byref(m.adam), int(self.dense_adam),
                                              m._buf.partials.data_ptr(), m._buf.counters.data_ptr(),
                                              m._buf.stats.data_ptr(), current_stream(m.device)),
                "vfmb_dp_apply_sampled")
        st = m._buf.stats
        return {"loss": st[L.ST_LOSS], "kl": st[L.ST_KL], "nll_mean": st[L.ST_NLL_MEAN],
                "pred": m._buf.mean[: self._B_local], "stats": st}

    def step(self, x_local: torch.Tensor, y_local: torch.Tensor, noise=None) -> dict:
        flat = self.local_backward(x_local, y_local, noise)
        allreduce_flat(flat, self.group)
        return self.apply(flat)
