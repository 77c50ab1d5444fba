"""Sampled-ELBO Variational Factorization Machine -- the drop-in for
``class CF`` and the training step of the reference's ``vfm-torch.py``.

Same PyTorch-facing API as the reference (vfm-torch.py:129-324):

    model = CF(embedding_size, output='reg', n_users=N, n_items=M, train_counts=nb_occ)
    likelihood, last_logits, mean_logits, kl_term = model(x)            # x: int64 [B,2]
    loss = -likelihood.log_prob(y).mean() * n_train + kl_term           # vfm-torch.py:359
    loss.backward(); optimizer.step()                                    # works (dense grads)

with the same parameter names, shapes and ``[mean | raw scale]`` row layout
(``bias_params.weight [R,2]``, ``entity_params.weight [R,2d]``).  The values
the reference keeps in module globals (``N, M, nb_occ, N_VARIATIONAL_SAMPLES,
LINK``; vfm-torch.py:18-19, 87-89, 125-126) are constructor arguments here.

The fast path is ``fused_step(x, y)``: plan, forward, backward and Adam on the
touched rows in five kernel launches, nothing leaving the device.

All arithmetic runs in hand-written sm_100a kernels behind the C ABI
(``include/vfm_b200.h``); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch
from torch import distributions, nn

from . import _lib as L
from .engine import (BatchPlan, GraphedLoop, PlanPipeline, StepBuffers, StepResult, current_stream,
                     make_config, require_cuda)

_LINKS = {"abs": torch.abs, "softplus": nn.functional.softplus}


class _SampledFn(torch.autograd.Function):
    """pred, kl_rows = f(bias table, entity table, global mean, global raw scale).

    Forward = vfmb_sampled_forward; backward = vfmb_sampled_backward in
    gradient-only mode (dense gradient tables, as ``nn.Embedding(sparse=False)``
    produces in the reference)."""

    @staticmethod
    def forward(ctx, bias_w, entity_w, gb_mean, gb_scale, model, x, noise):
        out = model._forward_kernels(x, None, noise)
        ctx.model, ctx.noise = model, noise
        ctx.e0 = model._global_eps(out)
        ctx.save_for_backward(gb_scale)
        return out["pred"].reshape(model.S, -1).clone(), out["kl_rows"].reshape(1).clone()

    @staticmethod
    def backward(ctx, g_pred, g_kl):
        model = ctx.model
        (gb_scale,) = ctx.saved_tensors
        # dloss/dpred is [S, B]; every sampled copy of a row sees rho_n = mean_s dloss/dpred[s, n]
        g_sb = g_pred.reshape(model.S, -1).float()
        resid = g_sb.mean(dim=0).contiguous()
        g_bias = torch.zeros_like(model.bias_params.weight)
        g_entity = torch.zeros_like(model.entity_params.weight)
        kl_scale = float(g_kl.reshape(-1)[0].item()) if g_kl is not None else 0.0
        model._backward_kernels(ctx.noise, L.GRAD_ONLY, kl_scale, resid=resid, grad_bias=g_bias,
                                grad_entity=g_entity, want_scalars=False)
        total_s = g_sb.sum(dim=1)                            # [S]: w0_s enters pred[s, :] only
        link = model.link_name
        dlink = torch.sign(gb_scale) if link == "abs" else torch.sigmoid(gb_scale)
        return (g_bias, g_entity, total_s.sum().reshape(1), (dlink * (ctx.e0 * total_s).sum()).reshape(1),
                None, None, None)


class CF(nn.Module):
    """Recommender system: sampled-ELBO VFM (drop-in for vfm-torch.py ``CF``)."""

    def __init__(self, embedding_size: int, output: str = "reg", *, n_users: int, n_items: int,
                 train_counts: torch.Tensor, n_var_samples: int = 1, link: str = "abs",
                 field_sizes: Optional[Sequence[int]] = None, kl_weighting: str = "torch",
                 interaction: Optional[str] = None, n_train: Optional[int] = None, max_batch: int = 65536, seed: int = 7,
                 lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, device="cuda"):
        super().__init__()
        device = require_cuda(device)
        L.lib()                                             # fail loudly if the extension is absent
        self.output, self.link_name = output, link
        self.N, self.M, self.d, self.S = int(n_users), int(n_items), int(embedding_size), int(n_var_samples)
        self.field_sizes = list(field_sizes) if field_sizes is not None else [self.N, self.M]
        self.F = len(self.field_sizes)
        self.R = int(sum(self.field_sizes))
        self.kl_weighting = kl_weighting
        # factor interaction of the deterministic predictions (last_logits / mean_logits / predict_mean):
        # the scripts' `prod(axis)` for two fields; for F > 2 the sampled step optimises the pairwise FM
        # (SURVEY N6), so the mean prediction must evaluate that same model unless asked otherwise
        self.interaction = interaction if interaction is not None else ("prod" if self.F == 2 else "pairwise")
        assert self.interaction in ("prod", "pairwise")
        self.seed, self.max_batch = int(seed), int(max_batch)
        self.adam = L.Adam(lr, betas[0], betas[1], eps)
        if kl_weighting == "torch":                         # vfm-torch.py:316, `uniq <= N`
            assert self.F == 2
            self._class_bounds, self._class_sizes = [self.N + 1], [self.N, self.M]
        else:                                               # vfm-tomasrch.py:574-587
            self._class_bounds = list(np.cumsum(self.field_sizes)[:-1])
            self._class_sizes = list(self.field_sizes)
        tc = torch.as_tensor(train_counts).reshape(-1).to(torch.float32)
        full = torch.zeros(self.R, dtype=torch.float32)
        full[: min(len(tc), self.R)] = tc[: self.R]        # minlength=R
        self.register_buffer("train_counts", full.to(device), persistent=False)
        self.n_train = float(n_train) if n_train is not None else float(tc.sum().item() / self.F)

        # parameters: created on the CPU in the reference's order so that torch.manual_seed
        # gives the reference's initial values (vfm-torch.py:136-153), then moved
        alpha = torch.Tensor([1e9])
        nn.init.uniform_(alpha)
        bias = nn.Embedding(self.R, 2)
        entity = nn.Embedding(self.R, 2 * embedding_size)
        scal = torch.zeros(L.S_COUNT, dtype=torch.float32)
        scal[L.S_ALPHA], scal[L.S_GB_MEAN], scal[L.S_GB_SCALE] = alpha[0], 0.0, 1.0
        self._scalars = scal.to(device)
        self.alpha = nn.Parameter(self._scalars[L.S_ALPHA:L.S_ALPHA + 1])
        self.global_bias_mean = nn.Parameter(self._scalars[L.S_GB_MEAN:L.S_GB_MEAN + 1])
        self.global_bias_scale = nn.Parameter(self._scalars[L.S_GB_SCALE:L.S_GB_SCALE + 1])
        # never used by the reference either; kept for state_dict compatibility (:139-143)
        self.prec_global_bias_prior = nn.Parameter(torch.ones(1, device=device))
        self.prec_user_bias_prior = nn.Parameter(torch.ones(1, device=device))
        self.prec_item_bias_prior = nn.Parameter(torch.ones(1, device=device))
        self.prec_user_entity_prior = nn.Parameter(torch.ones(embedding_size, device=device))
        self.prec_item_entity_prior = nn.Parameter(torch.ones(embedding_size, device=device))
        with torch.no_grad():       # a raw scale of exactly 0 is sigma = 0: invalid in the reference too
            # (Normal(scale=0) raises); N(0,1) init of >= 10^8 values returns a few -- nudge them
            for t in (bias.weight[:, 1:], entity.weight[:, embedding_size:]):
                t[t == 0] = 1e-4
        self.bias_params = bias.to(device)
        self.entity_params = entity.to(device)

        z = lambda t: torch.zeros_like(t)
        self.register_buffer("bias_m", z(self.bias_params.weight), persistent=False)
        self.register_buffer("bias_v", z(self.bias_params.weight), persistent=False)
        self.register_buffer("entity_m", z(self.entity_params.weight), persistent=False)
        self.register_buffer("entity_v", z(self.entity_params.weight), persistent=False)
        self._scalars_m, self._scalars_v = z(self._scalars), z(self._scalars)
        self.adam_step = torch.zeros(1, dtype=torch.int32, device=device)
        # Philox step word: [0] = index the next forward draws with, [1] = index of the last forward.
        # Separate from adam_step: the drop-in path (model(x) -> loss.backward() -> torch.optim.Adam)
        # never advances adam_step, and evaluation forwards draw fresh noise too (vfm-torch.py:402-403)
        self.noise_step = torch.zeros(2, dtype=torch.int32, device=device)

        # posterior-mean snapshots (vfm-torch.py:155-160, 179-185)
        self.saved_global_biases, self.saved_mean_biases, self.saved_mean_entities = [], [], []
        self.mean_saved_global_biases = self.mean_saved_mean_biases = self.mean_saved_mean_entities = None
        self._plan: Optional[BatchPlan] = None            # plan of the last step
        self._pipe: Optional[PlanPipeline] = None
        self._buf: Optional[StepBuffers] = None
        self._cfg: Optional[L.Config] = None
        self._cfg_cache = {}
        self._tab_cache = None                            # (entity ptr, bias ptr, Tables)
        self._fast_io = None

    # ------------------------------------------------------------------ plumbing
    @property
    def device(self):
        return self.entity_params.weight.device

    def _ensure(self, B: int):
        if self._pipe is None or B > self._pipe.B_cap:
            cap = max(B, self.max_batch if self._pipe is None else B)
            cfg = self._config(cap)
            self._pipe = PlanPipeline(cap, self.F, self.R, self.device)
            self._buf = StepBuffers(cfg, self._pipe.ring[0], self.device, L.S_COUNT, need_msg=self.F > 2)

    def _config(self, B: int) -> L.Config:
        cfg = self._cfg_cache.get(B)
        if cfg is None:
            cfg = make_config(B, self.F, self.d, self.R, self.S, self.output, self.link_name,
                              self._class_bounds, self._class_sizes, self.n_train, self.seed, self.interaction)
            self._cfg_cache[B] = cfg
        return cfg

    def _sync_scalars(self):
        """The three scalar parameters live in one device block the kernels read;
        re-pack if someone re-bound the parameters (load_state_dict keeps them)."""
        base = self._scalars.data_ptr()
        ok = (self.alpha.data_ptr() == base + 4 * L.S_ALPHA
              and self.global_bias_mean.data_ptr() == base + 4 * L.S_GB_MEAN
              and self.global_bias_scale.data_ptr() == base + 4 * L.S_GB_SCALE)
        if not ok:
            with torch.no_grad():
                self._scalars[L.S_ALPHA] = self.alpha.detach().reshape(-1)[0]
                self._scalars[L.S_GB_MEAN] = self.global_bias_mean.detach().reshape(-1)[0]
                self._scalars[L.S_GB_SCALE] = self.global_bias_scale.detach().reshape(-1)[0]
            self.alpha.data = self._scalars[L.S_ALPHA:L.S_ALPHA + 1]
            self.global_bias_mean.data = self._scalars[L.S_GB_MEAN:L.S_GB_MEAN + 1]
            self.global_bias_scale.data = self._scalars[L.S_GB_SCALE:L.S_GB_SCALE + 1]

    def _tables(self) -> L.Tables:
        key = (self.entity_params.weight.data_ptr(), self.bias_params.weight.data_ptr(),
               self._scalars.data_ptr())
        if self._tab_cache is not None and self._tab_cache[0] == key:
            return self._tab_cache[1]
        tab = self._make_tables()
        self._tab_cache = (key, tab)
        return tab

    def _make_tables(self) -> L.Tables:
        return L.Tables(L.ptr(self.bias_params.weight), L.ptr(self.bias_m), L.ptr(self.bias_v),
                        L.ptr(self.entity_params.weight), L.ptr(self.entity_m), L.ptr(self.entity_v),
                        L.ptr(self.train_counts), L.ptr(self._scalars), L.ptr(self._scalars_m),
                        L.ptr(self._scalars_v), L.ptr(self.adam_step), L.ptr(self.noise_step))

    def _prep_noise(self, noise):
        if noise is None:
            return None
        return tuple(torch.as_tensor(t, dtype=torch.float32, device=self.device).contiguous()
                     for t in noise)

    def plan(self, x: torch.Tensor) -> BatchPlan:
        """The batch plan for ``x`` (int64 ``[B,F]``): the prefetched one if ``prefetch_plan(x)``
        was called, else built now on the current stream."""
        x = x.to(self.device, non_blocking=True).contiguous()
        self._ensure(int(x.shape[0]))
        self._cfg = self._config(int(x.shape[0]))
        self._plan = self._pipe.acquire(self._cfg, x, self.train_counts)
        return self._plan

    def prefetch_plan(self, x: torch.Tensor, after=None) -> None:
        """Start building the plan of an upcoming batch on a side stream (it depends only on
        the ids).  ``x`` must stay unchanged until the step that uses it; ``after`` is an
        optional CUDA event marking when ``x`` is valid (e.g. its host-to-device copy)."""
        assert x.is_cuda and x.is_contiguous() and x.dtype == torch.int64
        self._ensure(int(x.shape[0]))
        self._pipe.prefetch(self._config(int(x.shape[0])), x, self.train_counts, after)

    def static_plan(self, x: torch.Tensor) -> BatchPlan:
        """A reusable plan for a batch that recurs every epoch (the reference loaders never
        shuffle, vfm-torch.py:121-122); pass it as ``fused_step(x, y, plan=...)``."""
        x = x.to(self.device).contiguous()
        self._ensure(int(x.shape[0]))
        return self._pipe.build_static(self._config(int(x.shape[0])), x, self.train_counts)

    def _forward_kernels(self, x, y, noise, plan: Optional[BatchPlan] = None):
        self._sync_scalars()
        if plan is None:
            self.plan(x)
        else:
            self._ensure(plan.B)
            self._cfg, self._plan = self._config(plan.B), plan
        noise = self._prep_noise(noise)
        if y is not None:
            y = y.to(self.device, torch.float32, non_blocking=True).contiguous()
        tab = self._tables()
        io = self._buf.io(y=y, noise=noise)
        L.check(L.lib().vfmb_sampled_forward(C.byref(self._cfg), C.byref(tab), C.byref(self._plan.struct),
                                             C.byref(io), current_stream(self.device)),
                "vfmb_sampled_forward")
        B, S = self._cfg.B, self.S
        st = self._buf.stats
        shape = (B,) if S == 1 else (S, B)                  # likelihood batch shape [S, B] (vfm-torch.py:265)
        return {"pred": self._buf.pred[:S * B].view(shape), "mean": self._buf.mean[:S * B].view(shape), "stats": st,
                "kl_rows": st[L.ST_KL_ROWS], "noise": noise}

    def _global_eps(self, out):
        """The N(0,1) draws ``[S]`` behind the sampled global bias of the last forward."""
        S = self.S
        if out["noise"] is not None:
            return out["noise"][0].reshape(-1)[:S].detach()
        sig0 = _LINKS[self.link_name](self._scalars[L.S_GB_SCALE])
        w0 = out["stats"][L.ST_W0:L.ST_W0 + 1] if S == 1 else out["stats"][L.ST_W0_S:L.ST_W0_S + S]
        return ((w0 - self._scalars[L.S_GB_MEAN]) / sig0).detach()

    def _backward_kernels(self, noise, mode, kl_scale, resid=None, grad_bias=None, grad_entity=None,
                          want_scalars=True):
        noise = self._prep_noise(noise)
        tab = self._tables()
        io = self._buf.io(noise=noise, grad_bias=grad_bias, grad_entity=grad_entity, resid=resid)
        if mode == L.GRAD_ONLY and not want_scalars:
            io.grad_scalars = None
        L.check(L.lib().vfmb_sampled_backward(C.byref(self._cfg), C.byref(tab), C.byref(self._plan.struct),
                                              C.byref(io), C.byref(self.adam), mode, float(kl_scale),
                                              current_stream(self.device)), "vfmb_sampled_backward")
        self._pipe.release(self._plan)

    # ------------------------------------------------------------------ reference API
    def forward(self, x: torch.Tensor, noise: Optional[Sequence[torch.Tensor]] = None):
        """``(likelihood, last_logits, mean_logits, kl_term)`` as vfm-torch.py:189-324.

        ``noise`` optionally injects the N(0,1) draws ``(eps0 [S,1], eps_bias [S,U],
        eps_entity [S,U,d])`` indexed by unique rank; default is the Philox stream."""
        link = _LINKS[self.link_name]
        pred, kl_rows = _SampledFn.apply(self.bias_params.weight, self.entity_params.weight,
                                         self.global_bias_mean, self.global_bias_scale, self, x, noise)
        if self.mean_saved_mean_biases is not None:        # vfm-torch.py:248-259
            last_logits = self._mean_logits(x, self.saved_global_biases[-1], self.saved_mean_biases[-1],
                                            self.saved_mean_entities[-1])
            mean_logits = self._mean_logits(x, self.mean_saved_global_biases, self.mean_saved_mean_biases,
                                            self.mean_saved_mean_entities)
        else:
            last_logits = mean_logits = None
        if self.output == "reg":
            likelihood = distributions.normal.Normal(pred, torch.sqrt(1 / link(self.alpha)))
        else:
            likelihood = distributions.bernoulli.Bernoulli(logits=pred)
        q0 = distributions.normal.Normal(self.global_bias_mean, link(self.global_bias_scale))
        kl0 = distributions.kl.kl_divergence(q0, distributions.normal.Normal(0., 1.))
        return likelihood, last_logits, mean_logits, kl0 + kl_rows

    def save_weights(self):
        """Snapshot of the posterior means (vfm-torch.py:179-185), kept on the device."""
        d = self.d
        self.saved_global_biases.append(self.global_bias_mean.detach().clone())
        self.saved_mean_biases.append(self.bias_params.weight[:, 0].detach().clone())
        self.saved_mean_entities.append(self.entity_params.weight[:, :d].detach().clone())
        n = len(self.saved_global_biases)
        if n == 1:
            self.mean_saved_global_biases = self.saved_global_biases[0].clone()
            self.mean_saved_mean_biases = self.saved_mean_biases[0].clone()
            self.mean_saved_mean_entities = self.saved_mean_entities[0].clone()
        else:                                               # running mean == mean over the list
            self.mean_saved_global_biases += (self.saved_global_biases[-1] - self.mean_saved_global_biases) / n
            self.mean_saved_mean_biases += (self.saved_mean_biases[-1] - self.mean_saved_mean_biases) / n
            self.mean_saved_mean_entities += (self.saved_mean_entities[-1] - self.mean_saved_mean_entities) / n

    def _mean_logits(self, x, global_bias, mean_biases, mean_entities) -> torch.Tensor:
        """global + sum of bias means + <product of factor means> via vfmb_predict_mean."""
        x = x.to(self.device).contiguous()
        B = int(x.shape[0])
        bias2 = torch.stack((mean_biases, torch.zeros_like(mean_biases)), dim=1).contiguous()
        ent2 = torch.cat((mean_entities, torch.zeros_like(mean_entities)), dim=1).contiguous()
        out = torch.empty(B, dtype=torch.float32, device=self.device)
        cfg = self._config(B)
        L.check(L.lib().vfmb_predict_mean(C.byref(cfg), bias2.data_ptr(), ent2.data_ptr(),
                                          float(global_bias.reshape(-1)[0].item()), x.data_ptr(),
                                          out.data_ptr(), current_stream(self.device)), "vfmb_predict_mean")
        return out

    def predict_mean(self, x: torch.Tensor) -> torch.Tensor:
        """Deterministic prediction from the current posterior means (no plan, no RNG)."""
        x = x.to(self.device).contiguous()
        B = int(x.shape[0])
        out = torch.empty(B, dtype=torch.float32, device=self.device)
        cfg = self._config(B)
        L.check(L.lib().vfmb_predict_mean(C.byref(cfg), L.ptr(self.bias_params.weight),
                                          L.ptr(self.entity_params.weight),
                                          float(self.global_bias_mean.item()), x.data_ptr(),
                                          out.data_ptr(), current_stream(self.device)), "vfmb_predict_mean")
        return out

    @torch.no_grad()
    def predict_proba(self, x: torch.Tensor, n_samples: int = 64, per_occurrence: bool = False, return_logit_mean: bool = False):
        """``(proba_means, logit_variances)`` over ``n_samples`` variational samples of every row of
        ``x`` -- vfm.py:1047-1057 ``predict_proba`` (mean over the samples of the likelihood mean, and the
        population variance of the logits), the statistic behind the active-learning question selection
        (vfm.py:1024-1045).  One kernel launch for any ``n_samples``; consumes one noise index.
        ``per_occurrence``: independent draws per row of ``x`` (vfm.py:440-445) instead of one draw per
        entity (vfm-torch.py:238-245); the per-row statistics are the same in distribution."""
        self._sync_scalars()
        x = x.to(self.device).contiguous()
        B = int(x.shape[0])
        cfg = self._config(B)
        pm = torch.empty(B, dtype=torch.float32, device=self.device)
        lm = torch.empty(B, dtype=torch.float32, device=self.device)
        lv = torch.empty(B, dtype=torch.float32, device=self.device)
        s = current_stream(self.device)
        L.check(L.lib().vfmb_predict_sampled(C.byref(cfg), C.byref(self._tables()), x.data_ptr(), int(n_samples),
                                             int(bool(per_occurrence)), pm.data_ptr(), lm.data_ptr(), lv.data_ptr(), s),
                "vfmb_predict_sampled")
        L.check(L.lib().vfmb_adam_step_advance(self.noise_step.data_ptr(), s))       # next forward: fresh noise
        return (pm, lv, lm) if return_logit_mean else (pm, lv)

    # ------------------------------------------------------------------ fast path
    def configure_adam(self, lr: float, betas=(0.9, 0.999), eps: float = 1e-8):
        self.adam = L.Adam(lr, betas[0], betas[1], eps)

    @torch.no_grad()
    def fused_step(self, x: torch.Tensor, y: torch.Tensor,
                   noise: Optional[Sequence[torch.Tensor]] = None, update: bool = True,
                   plan: Optional[BatchPlan] = None) -> dict:
        """Plan + forward + backward + Adam on the touched rows (vfm-torch.py:351-370 in
        one go).  Returns device tensors (views of reused buffers -- clone to keep):
        ``loss [ ]``, ``kl [ ]``, ``pred [B]`` (likelihood mean), ``logits [B]``."""
        if (noise is None and update and x.is_cuda and y.is_cuda and y.dtype == torch.float32
                and x.is_contiguous() and y.is_contiguous()):
            return self._fused_step_fast(x, y, plan)
        out = self._forward_kernels(x, y, noise, plan)
        if update:
            self._backward_kernels(noise, L.ADAM_TOUCHED, 1.0)
        else:
            self._pipe.release(self._plan)
        st = out["stats"]
        return {"loss": st[L.ST_LOSS], "kl": st[L.ST_KL], "nll_mean": st[L.ST_NLL_MEAN],
                "pred": out["mean"], "logits": out["pred"], "stats": st}

    def graphed_loop(self, B: int, depth: int = 2, **kw) -> GraphedLoop:
        """CUDA-graph replay of the training loop for batches of exactly ``B`` samples: the step
        on the current batch and the plan of the next one are captured once and replayed, which
        removes the host launch overhead (the step is otherwise launch-bound).  See GraphedLoop."""
        self._ensure(B)
        self._sync_scalars()

        def step_fn(plan, y, outs):
            self._cfg, self._plan = self._config(B), plan
            io = self._buf.io(y=y)
            io.pred, io.mean, io.stats = L.ptr(outs.pred), L.ptr(outs.mean), L.ptr(outs.stats)
            self._graph_io = getattr(self, "_graph_io", []) + [io]      # keep the structs alive
            L.check(L.lib().vfmb_sampled_step(C.byref(self._cfg), C.byref(self._tables()), C.byref(plan.struct),
                                              C.byref(io), C.byref(self.adam), current_stream(self.device)),
                    "vfmb_sampled_step")
        return GraphedLoop(self, B, step_fn, depth, **kw)

    def _fused_step_fast(self, x, y, plan):
        """Hot loop: cached ctypes structures, one C call for forward + backward + Adam."""
        self._sync_scalars()
        B = int(x.shape[0])
        if plan is None:
            self._ensure(B)
            self._cfg = self._config(B)
            self._plan = self._pipe.acquire(self._cfg, x, self.train_counts)
        else:
            self._ensure(plan.B)
            self._cfg, self._plan = self._config(plan.B), plan
        io = self._fast_io
        if io is None or io[0] is not self._buf:
            io = (self._buf, self._buf.io())
            self._fast_io = io
        io[1].y = y.data_ptr()
        L.check(L.lib().vfmb_sampled_step(C.byref(self._cfg), C.byref(self._tables()),
                                          C.byref(self._plan.struct), C.byref(io[1]), C.byref(self.adam),
                                          current_stream(self.device)), "vfmb_sampled_step")
        self._pipe.release(self._plan)
        return StepResult(self._buf, B)

    @torch.no_grad()
    def gradients(self, x, y, noise=None) -> dict:
        """Dense gradients of the step loss (what ``loss.backward()`` leaves in ``.grad``
        in the reference) without touching the parameters; for parity tests and for
        the data-parallel all-reduce mode."""
        out = self._forward_kernels(x, y, noise)
        g_bias = torch.zeros_like(self.bias_params.weight)
        g_entity = torch.zeros_like(self.entity_params.weight)
        self._backward_kernels(noise, L.GRAD_ONLY, 1.0, grad_bias=g_bias, grad_entity=g_entity)
        gs = self._buf.grad_scalars
        return {"bias_params.weight": g_bias, "entity_params.weight": g_entity,
                "alpha": gs[L.S_ALPHA:L.S_ALPHA + 1].clone(),
                "global_bias_mean": gs[L.S_GB_MEAN:L.S_GB_MEAN + 1].clone(),
                "global_bias_scale": gs[L.S_GB_SCALE:L.S_GB_SCALE + 1].clone(),
                "loss": out["stats"][L.ST_LOSS].clone(), "pred": out["mean"].clone()}

    @torch.no_grad()
    def dense_adam_step(self, grads: dict):
        """Reference-exact optimiser semantics: dense torch.optim.Adam over every row
        (vfm-torch.py:339,370) from dense gradients, e.g. after a DP all-reduce."""
        s = current_stream(self.device)
        lib = L.lib()
        for p, m, v, g in ((self.bias_params.weight, self.bias_m, self.bias_v, grads["bias_params.weight"]),
                           (self.entity_params.weight, self.entity_m, self.entity_v, grads["entity_params.weight"])):
            L.check(lib.vfmb_adam_dense(p.data_ptr(), m.data_ptr(), v.data_ptr(), g.contiguous().data_ptr(),
                                        p.numel(), C.byref(self.adam), self.adam_step.data_ptr(), s))
        gs = torch.zeros_like(self._scalars)
        gs[L.S_GB_MEAN], gs[L.S_GB_SCALE] = grads["global_bias_mean"][0], grads["global_bias_scale"][0]
        n = L.S_COUNT
        if self.output == "reg":
            gs[L.S_ALPHA] = grads["alpha"][0]
            L.check(lib.vfmb_adam_dense(self._scalars.data_ptr(), self._scalars_m.data_ptr(),
                                        self._scalars_v.data_ptr(), gs.data_ptr(), n, C.byref(self.adam),
                                        self.adam_step.data_ptr(), s))
        else:                                               # alpha has no gradient
            off = 4 * L.S_GB_MEAN
            L.check(lib.vfmb_adam_dense(self._scalars.data_ptr() + off, self._scalars_m.data_ptr() + off,
                                        self._scalars_v.data_ptr() + off, gs.data_ptr() + off, 2,
                                        C.byref(self.adam), self.adam_step.data_ptr(), s))
        L.check(lib.vfmb_adam_step_advance(self.adam_step.data_ptr(), s))

    @torch.no_grad()
    def philox_noise(self, uniq: torch.Tensor, step: Optional[int] = None):
        """The N(0,1) draws the Philox path uses at noise index ``step`` for the given unique rows
        (default: the index the NEXT forward will draw with; ``noise_step[1]`` is the last one used)."""
        uniq = uniq.to(self.device, torch.int32).contiguous()
        U = int(uniq.numel())
        step = int(self.noise_step[0].item()) if step is None else int(step)
        S = self.S
        e0 = torch.empty(S, device=self.device)
        eb = torch.empty(S * U, device=self.device)
        ee = torch.empty(S * U * self.d, device=self.device)
        cfg = self._config(1)
        L.check(L.lib().vfmb_philox_normals(C.byref(cfg), uniq.data_ptr(), U, step, e0.data_ptr(),
                                            eb.data_ptr(), ee.data_ptr(), current_stream(self.device)))
        return e0.reshape(S, 1), eb.reshape(S, U), ee.reshape(S, U, self.d)
