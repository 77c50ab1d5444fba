"""Closed-form Gaussian Variational Factorization Machine with learnable per-group priors --
the drop-in for ``class CF`` and the training step of the reference's ``vfm-tomasrch.py``.

Reference API (vfm-tomasrch.py:190-191, 323, 451-453):

    model = CF(embedding_size, n_groups, group_sizes, n_var_samples, alpha_0, output)
    likelihood, kls, partial_loss = model(inverse_group, group_present, closed_form_loss=True, target=y)
    loss = -N_train * partial_loss / B + kls[0] + sum_u c_u (kls[1] + kls[2].sum(1))      # :569-588

Same parameter names / shapes / ``[mean | raw scale]`` row layout and the same seeded
initialisation order (:194-260).  ``entity_count`` (``bincount(X_train)``, :182) and ``N_train``
are constructor arguments instead of module globals.

``forward`` evaluates the same quantities with the CUDA kernels.  Under autograd (grad mode on and
``closed_form_loss=True``) it is differentiable, so the reference's own loop (:548-594: ``model(...)``,
the loss expression, ``loss.backward()``, ``torch.optim.Adam.step()``) runs unchanged on this module:
``partial_loss`` and the per-row KL terms ``kls[1]``, ``kls[2]`` come from a ``torch.autograd.Function``
whose backward runs the CUDA backward with the upstream gradients; ``kls[0]`` is differentiated by
torch.  The fast path is ``fused_step(x, y)``: plan, forward, backward and Adam on the touched rows on
the device, including the gradients of the prior parameters.  ``gradients(x, y)`` returns what
``loss.backward()`` leaves in ``.grad``.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch
from torch import distributions, nn

from . import _lib as L
from .engine import (BatchPlan, GraphedLoop, PlanPipeline, StepBuffers, StepResult, current_stream, make_config,
                     require_cuda)


class _ClosedFn(torch.autograd.Function):
    """(pred, partial_loss, kls[1], kls[2]) = f(bias_params, entity_params, scalar parameters).

    Forward = vfmb_closed_forward; backward = vfmb_closed_backward_weighted in gradient-only mode with
    the upstream gradients: d loss / d partial_loss and the per-row weights of the KL terms.  The KL
    weights must be uniform along a row and shared by kls[1][u] and kls[2][u, :] -- the structure of
    both loss expressions of the script (:569-588: ``(kls[1] + kls[2].sum(axis=1)) * w``); anything else
    poisons the gradients with NaN rather than returning wrong numbers."""

    @staticmethod
    def forward(ctx, model, raw, y, n_unique, bias_params, entity_params, *scalar_params):
        klb = torch.empty(n_unique, dtype=torch.float32, device=model.device)
        kle = torch.empty(n_unique, model.d, dtype=torch.float32, device=model.device)
        model._run_forward(raw, y, kl_out=(klb, kle))
        B = model._cfg.B
        pred = model._buf.pred[:B].clone()
        partial = model._buf.stats[L.ST_NLL_MEAN].clone()
        ctx.model = model
        ctx.mark_non_differentiable(pred)
        return pred, partial, klb, kle

    @staticmethod
    def backward(ctx, g_pred, g_partial, g_klb, g_kle):
        model = ctx.model
        if (g_klb is None) != (g_kle is None):
            raise RuntimeError("vae_b200.vfm_tomasrch: the loss must weight kls[1] and kls[2] together "
                               "(vfm-tomasrch.py:569-588)")
        U = model._io_n_unique
        if g_klb is None:
            w = torch.zeros(U, dtype=torch.float32, device=model.device)
            bad = torch.zeros((), dtype=torch.bool, device=model.device)
        else:
            w = g_klb.to(torch.float32).contiguous()
            spread = (g_kle.to(torch.float32) - w[:, None]).abs().max()
            bad = spread > 1e-6 * w.abs().max().clamp_min(1e-30)
        scale = 0.0 if g_partial is None else -float(g_partial.item())
        g_bias = torch.zeros_like(model.bias_params.data)
        g_entity = torch.zeros_like(model.entity_params.data)
        model._io.grad_bias, model._io.grad_entity = L.ptr(g_bias), L.ptr(g_entity)
        L.check(L.lib().vfmb_closed_backward_weighted(
            C.byref(model._cfg), C.byref(model._tables()), C.byref(model._plan.struct), C.byref(model._io),
            C.byref(model.adam), L.GRAD_ONLY, w.data_ptr(), scale, 0.0, current_stream(model.device)),
            "vfmb_closed_backward_weighted")
        model._pipe.release(model._plan)
        poison = torch.where(bad, torch.full((), float("nan"), device=model.device), torch.ones((), device=model.device))
        gs = model._buf.grad_scalars
        outs = [(gs[off:off + p.numel()] * poison).reshape(p.shape) for p, off in model._scalar_params()]
        return (None, None, None, None, g_bias * poison, g_entity * poison, *outs)


class CF(nn.Module):
    """Recommender system: closed-form VFM (drop-in for vfm-tomasrch.py ``CF``)."""

    def __init__(self, embedding_size: int = 2, n_groups: int = 2, group_sizes: Optional[Sequence[int]] = None,
                 n_var_samples: int = 1, alpha_0: float = 300, output: str = "reg", *,
                 train_counts: torch.Tensor, n_train: int, max_batch: int = 8000, lr: float = 0.1,
                 betas=(0.9, 0.999), eps: float = 1e-8, device="cuda"):
        super().__init__()
        device = require_cuda(device)
        lib = L.lib()
        assert output == "reg", "the closed-form ELBO is Gaussian (vfm-tomasrch.py:738)"
        assert group_sizes is not None and n_groups == len(group_sizes)
        self.output, self.n_var_samples = output, n_var_samples
        self.embedding_size = self.d = int(embedding_size)
        self.n_groups = self.G = int(n_groups)
        self.group_sizes = list(map(int, group_sizes))
        self.R = int(sum(self.group_sizes))
        self.n_train, self.max_batch = float(n_train), int(max_batch)
        self.adam = L.Adam(lr, betas[0], betas[1], eps)
        G, d = self.G, self.d
        tc = torch.as_tensor(train_counts).reshape(-1).to(torch.float32)
        full = torch.zeros(self.R, dtype=torch.float32)
        full[: min(len(tc), self.R)] = tc[: self.R]
        self.register_buffer("train_counts", full.to(device), persistent=False)

        # ---- initial values in the reference's creation order (vfm-tomasrch.py:194-260), on the CPU
        start_scale = 0.2
        init = {"alpha": torch.Tensor([alpha_0]), "mean_global_bias_prior": torch.Tensor([0.]),
                "scale_global_bias_prior": torch.Tensor([1.]),
                "mean_global_bias": torch.normal(torch.zeros(1), torch.ones(1)),
                "scale_global_bias": torch.Tensor([start_scale])}
        bias0 = torch.cat([torch.cat((torch.normal(torch.zeros(n, 1), 1e-1 * torch.ones(n, 1)),
                                      start_scale * torch.ones(n, 1)), dim=1) for n in self.group_sizes])
        ent0 = torch.cat([torch.cat((torch.normal(torch.zeros(n, d), 1e-7 * torch.ones(n, d)),
                                     start_scale * torch.ones(n, d)), dim=1) for n in self.group_sizes])

        # ---- scalar block the kernels read; every scalar Parameter is a view into it
        self._off = {"pbm": lib.vfmb_closed_off_bias_prior_mean(G, d, 0),
                     "pbs": lib.vfmb_closed_off_bias_prior_scale(G, d, 0),
                     "pem": lib.vfmb_closed_off_entity_prior_mean(G, d, 0),
                     "pes": lib.vfmb_closed_off_entity_prior_scale(G, d, 0)}
        self.n_scalars = int(lib.vfmb_closed_scalar_count(G, d))
        blk = torch.zeros(self.n_scalars, dtype=torch.float32)
        blk[L.C_ALPHA], blk[L.C_GB_PRIOR_MEAN], blk[L.C_GB_PRIOR_SCALE] = init["alpha"][0], 0.0, 1.0
        blk[L.C_GB_MEAN], blk[L.C_GB_SCALE] = init["mean_global_bias"][0], start_scale
        blk[self._off["pbs"]: self._off["pbs"] + G] = 1.0
        blk[self._off["pes"]: self._off["pes"] + G * d] = 1.0
        self._scalars = blk.to(device)
        view = lambda o, n: nn.Parameter(self._scalars[o:o + n])
        self.alpha = view(L.C_ALPHA, 1)
        self.mean_global_bias_prior = view(L.C_GB_PRIOR_MEAN, 1)
        self.scale_global_bias_prior = view(L.C_GB_PRIOR_SCALE, 1)
        self.mean_global_bias = view(L.C_GB_MEAN, 1)
        self.scale_global_bias = view(L.C_GB_SCALE, 1)
        self.mean_group_bias_prior = nn.ParameterList([view(self._off["pbm"] + g, 1) for g in range(G)])
        self.scale_group_bias_prior = nn.ParameterList([view(self._off["pbs"] + g, 1) for g in range(G)])
        self.bias_params = nn.Parameter(bias0.to(device))
        self.mean_group_entity_prior = nn.ParameterList([view(self._off["pem"] + g * d, d) for g in range(G)])
        self.scale_group_entity_prior = nn.ParameterList([view(self._off["pes"] + g * d, d) for g in range(G)])
        self.entity_params = nn.Parameter(ent0.to(device))

        z = lambda t: torch.zeros_like(t)
        self.register_buffer("bias_m", z(self.bias_params.data), persistent=False)
        self.register_buffer("bias_v", z(self.bias_params.data), persistent=False)
        self.register_buffer("entity_m", z(self.entity_params.data), persistent=False)
        self.register_buffer("entity_v", z(self.entity_params.data), persistent=False)
        self._scalars_m, self._scalars_v = z(self._scalars), z(self._scalars)
        self.adam_step = torch.zeros(1, dtype=torch.int32, device=device)
        self._class_bounds = list(np.cumsum(self.group_sizes)[:-1])
        self._pipe: Optional[PlanPipeline] = None
        self._buf: Optional[StepBuffers] = None
        self._plan: Optional[BatchPlan] = None
        self._cfg_cache = {}

    # ------------------------------------------------------------------ plumbing
    @property
    def device(self):
        return self.entity_params.device

    def _scalar_params(self):
        yield self.alpha, L.C_ALPHA
        yield self.mean_global_bias_prior, L.C_GB_PRIOR_MEAN
        yield self.scale_global_bias_prior, L.C_GB_PRIOR_SCALE
        yield self.mean_global_bias, L.C_GB_MEAN
        yield self.scale_global_bias, L.C_GB_SCALE
        for g in range(self.G):
            yield self.mean_group_bias_prior[g], self._off["pbm"] + g
            yield self.scale_group_bias_prior[g], self._off["pbs"] + g
            yield self.mean_group_entity_prior[g], self._off["pem"] + g * self.d
            yield self.scale_group_entity_prior[g], self._off["pes"] + g * self.d

    def _sync_scalars(self):
        base = self._scalars.data_ptr()
        for p, off in self._scalar_params():
            if p.data_ptr() != base + 4 * off:               # re-bound (e.g. .to()): re-pack
                with torch.no_grad():
                    self._scalars[off:off + p.numel()] = p.detach().reshape(-1)
                p.data = self._scalars[off:off + p.numel()]

    def _config(self, B: int) -> L.Config:
        cfg = self._cfg_cache.get(B)
        if cfg is None:
            cfg = make_config(B, self.G, self.d, self.R, 1, "reg", "abs", self._class_bounds,
                              self.group_sizes, self.n_train, 0)
            self._cfg_cache[B] = cfg
        return cfg

    def _ensure(self, B: int):
        if self._pipe is None or B > self._pipe.B_cap:
            cap = max(B, self.max_batch if self._pipe is None else B)
            self._pipe = PlanPipeline(cap, self.G, self.R, self.device)
            self._buf = StepBuffers(self._config(cap), self._pipe.ring[0], self.device, self.n_scalars,
                                    need_msg=self.G > 2, closed=True)

    def _tables(self) -> L.Tables:
        return L.Tables(L.ptr(self.bias_params.data), L.ptr(self.bias_m), L.ptr(self.bias_v),
                        L.ptr(self.entity_params.data), L.ptr(self.entity_m), L.ptr(self.entity_v),
                        L.ptr(self.train_counts), L.ptr(self._scalars), L.ptr(self._scalars_m),
                        L.ptr(self._scalars_v), L.ptr(self.adam_step))

    def prefetch_plan(self, x: torch.Tensor, after=None) -> None:
        self._ensure(int(x.shape[0]))
        self._pipe.prefetch(self._config(int(x.shape[0])), x, self.train_counts, after)

    def static_plan(self, x: torch.Tensor) -> BatchPlan:
        x = x.to(self.device).contiguous()
        self._ensure(int(x.shape[0]))
        return self._pipe.build_static(self._config(int(x.shape[0])), x, self.train_counts)

    def _run_forward(self, x, y, plan=None, kl_out=None):
        self._sync_scalars()
        if plan is None:
            x = x.to(self.device).contiguous()
            B = int(x.shape[0])
            self._ensure(B)
            self._cfg = self._config(B)
            self._plan = self._pipe.acquire(self._cfg, x, self.train_counts)
        else:
            self._ensure(plan.B)
            self._cfg, self._plan = self._config(plan.B), plan
        if y is not None:
            y = y.to(self.device, torch.float32).contiguous()
        klb, kle = kl_out if kl_out is not None else (None, None)
        self._io = self._buf.io(y=y, kl_bias_out=klb, kl_entity_out=kle)
        self._y = y                                          # keep alive until the kernels ran
        L.check(L.lib().vfmb_closed_forward(C.byref(self._cfg), C.byref(self._tables()),
                                            C.byref(self._plan.struct), C.byref(self._io),
                                            current_stream(self.device)), "vfmb_closed_forward")

    def _run_backward(self, mode, grad_bias=None, grad_entity=None):
        self._io.grad_bias, self._io.grad_entity = L.ptr(grad_bias), L.ptr(grad_entity)
        L.check(L.lib().vfmb_closed_backward(C.byref(self._cfg), C.byref(self._tables()),
                                             C.byref(self._plan.struct), C.byref(self._io), C.byref(self.adam),
                                             mode, current_stream(self.device)), "vfmb_closed_backward")
        self._pipe.release(self._plan)

    # ------------------------------------------------------------------ fast path
    def configure_adam(self, lr: float, betas=(0.9, 0.999), eps: float = 1e-8):
        self.adam = L.Adam(lr, betas[0], betas[1], eps)

    @torch.no_grad()
    def fused_step(self, x: torch.Tensor, y: torch.Tensor, plan: Optional[BatchPlan] = None,
                   update: bool = True) -> StepResult:
        """One training step of vfm-tomasrch.py:535-594 on raw ids ``x`` (int64 ``[B,G]``, global row
        ids as in ``X_train``).  ``res["loss"]`` is the step's ELBO loss, ``res["pred"]`` the
        product-form prediction (``outputs.mean``, :342-348), ``res["nll_mean"]`` the partial_loss."""
        self._run_forward(x, y, plan)
        if update:
            self._run_backward(L.ADAM_TOUCHED)
        else:
            self._pipe.release(self._plan)
        return StepResult(self._buf, self._cfg.B)

    def graphed_loop(self, B: int, depth: int = 2, **kw) -> GraphedLoop:
        """CUDA-graph replay of the training loop for batches of exactly ``B`` samples (the step on
        batch i and the plan of batch i+1 captured once): the ML-100K-sized step is launch-bound, so the
        host launch overhead is what the eager loop measures.  See engine.GraphedLoop."""
        self._ensure(B)
        self._sync_scalars()

        def step_fn(plan, y, outs):
            self._cfg, self._plan = self._config(B), plan
            io = self._buf.io(y=y)
            io.pred, io.mean, io.stats = L.ptr(outs.pred), L.ptr(outs.mean), L.ptr(outs.stats)
            self._graph_io = getattr(self, "_graph_io", []) + [io]      # keep the structs alive
            tab, s = self._tables(), current_stream(self.device)
            L.check(L.lib().vfmb_closed_forward(C.byref(self._cfg), C.byref(tab), C.byref(plan.struct), C.byref(io), s),
                    "vfmb_closed_forward")
            L.check(L.lib().vfmb_closed_backward(C.byref(self._cfg), C.byref(tab), C.byref(plan.struct), C.byref(io),
                                                 C.byref(self.adam), L.ADAM_TOUCHED, s), "vfmb_closed_backward")
        return GraphedLoop(self, B, step_fn, depth, **kw)

    @torch.no_grad()
    def gradients(self, x: torch.Tensor, y: torch.Tensor) -> dict:
        """Dense gradients of the step loss, keyed like ``named_parameters()``."""
        self._run_forward(x, y)
        g_bias = torch.zeros_like(self.bias_params.data)
        g_entity = torch.zeros_like(self.entity_params.data)
        self._run_backward(L.GRAD_ONLY, g_bias, g_entity)
        gs, d, G, o = self._buf.grad_scalars, self.d, self.G, self._off
        out = {"bias_params": g_bias, "entity_params": g_entity,
               "alpha": gs[L.C_ALPHA:L.C_ALPHA + 1].clone(),
               "mean_global_bias": gs[L.C_GB_MEAN:L.C_GB_MEAN + 1].clone(),
               "scale_global_bias": gs[L.C_GB_SCALE:L.C_GB_SCALE + 1].clone(),
               "mean_global_bias_prior": gs[L.C_GB_PRIOR_MEAN:L.C_GB_PRIOR_MEAN + 1].clone(),
               "scale_global_bias_prior": gs[L.C_GB_PRIOR_SCALE:L.C_GB_PRIOR_SCALE + 1].clone(),
               "loss": self._buf.stats[L.ST_LOSS].clone(), "pred": self._buf.pred[: self._cfg.B].clone()}
        for g in range(G):
            out[f"mean_group_bias_prior.{g}"] = gs[o["pbm"] + g: o["pbm"] + g + 1].clone()
            out[f"scale_group_bias_prior.{g}"] = gs[o["pbs"] + g: o["pbs"] + g + 1].clone()
            out[f"mean_group_entity_prior.{g}"] = gs[o["pem"] + g * d: o["pem"] + (g + 1) * d].clone()
            out[f"scale_group_entity_prior.{g}"] = gs[o["pes"] + g * d: o["pes"] + (g + 1) * d].clone()
        return out

    # ------------------------------------------------------------------ reference API
    def forward(self, x: List[torch.Tensor], x_unique: List[torch.Tensor], closed_form_loss: bool = False,
                target=False):
        """``(likelihood, kls[, partial_loss])`` as vfm-tomasrch.py:323-453, from the per-group
        inverse indices ``x`` and unique lists ``x_unique`` the script's loop builds (:536-545).
        With grad mode on and ``closed_form_loss=True`` the outputs carry gradients (training loop,
        :548-594); otherwise values only (the evaluation path, :664-678)."""
        raw = torch.stack([x_unique[g].to(self.device)[x[g].to(self.device)] for g in range(self.G)], dim=1).contiguous()
        U = sum(int(len(u)) for u in x_unique)
        self._io_n_unique = U
        absl = torch.abs
        if torch.is_grad_enabled() and closed_form_loss:
            pred, partial, klb, kle = _ClosedFn.apply(self, raw, target, U, self.bias_params, self.entity_params,
                                                      *[p for p, _ in self._scalar_params()])
            likelihood = distributions.normal.Normal(pred, torch.sqrt(1 / absl(self.alpha)))
            q0 = distributions.normal.Normal(self.mean_global_bias, absl(self.scale_global_bias))
            p0 = distributions.normal.Normal(self.mean_global_bias_prior, absl(self.scale_global_bias_prior))
            return likelihood, [distributions.kl.kl_divergence(q0, p0), klb, kle], partial
        with torch.no_grad():
            klb = torch.empty(U, dtype=torch.float32, device=self.device)
            kle = torch.empty(U, self.d, dtype=torch.float32, device=self.device)
            y = target if closed_form_loss else None
            self._run_forward(raw, y, kl_out=(klb, kle))
            self._pipe.release(self._plan)
            B = self._cfg.B
            pred = self._buf.pred[:B].clone()
            likelihood = distributions.normal.Normal(pred, torch.sqrt(1 / absl(self.alpha.detach())))
            q0 = distributions.normal.Normal(self.mean_global_bias.detach(), absl(self.scale_global_bias.detach()))
            p0 = distributions.normal.Normal(self.mean_global_bias_prior.detach(),
                                             absl(self.scale_global_bias_prior.detach()))
            kls = [distributions.kl.kl_divergence(q0, p0), klb, kle]
            if closed_form_loss:
                return likelihood, kls, self._buf.stats[L.ST_NLL_MEAN].clone()
            return likelihood, kls

    @torch.no_grad()
    def predict(self, x: torch.Tensor, bounds=None) -> torch.Tensor:
        """Mean prediction on raw ids (global + bias means + product of factor means), optionally
        clipped to ``BOUNDS`` (vfm-tomasrch.py:35, 678)."""
        x = x.to(self.device).contiguous()
        B = int(x.shape[0])
        out = torch.empty(B, dtype=torch.float32, device=self.device)
        cfg = self._config(B)
        L.check(L.lib().vfmb_predict_mean(C.byref(cfg), L.ptr(self.bias_params.data), L.ptr(self.entity_params.data),
                                          float(self.mean_global_bias.item()), x.data_ptr(), out.data_ptr(),
                                          current_stream(self.device)), "vfmb_predict_mean")
        return out.clip(*bounds) if bounds is not None else out
