"""CPU restatement ("port") of the reference VFM step in plain torch.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  This module is the
checker for the CUDA path and the timed CPU baseline; the product never
imports it.

It follows the reference op for op (same ATen calls, so its CPU cost is
representative of the reference's), but is written against explicit
arguments instead of module globals so that it can travel to the GPU box:

* ``SampledPort``  <- ``class CF`` + loop of ``vfm-torch.py:129-324, 351-370``
* ``ClosedPort``   <- ``class CF`` + loop of ``vfm-tomasrch.py:186-453, 535-594``

Both are validated against the AST-sliced reference classes in
``tests/test_oracle_vs_reference.py`` (this container) and against
``tests/golden/*.npz`` (everywhere).

Extensions beyond the reference, used only where the reference has no code
(BASELINE config 4): ``interaction="pairwise"`` for F>2 fields in the sampled
model (authority: the pairwise closed form ``vfm-tomasrch.py:379-393`` and the
TF original ``vfm.py:467-474``) and ``kl_weighting="group"`` (the per-group
normaliser of ``vfm-tomasrch.py:574-587``).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch
from torch import nn
from torch.distributions import Bernoulli, Normal, kl_divergence

_LINKS = {"abs": torch.abs, "softplus": nn.functional.softplus}


def _pairwise(v: torch.Tensor, field_axis: int) -> torch.Tensor:
    """sum_{i<j} v_i * v_j over ``field_axis`` = 0.5((sum v)^2 - sum v^2), per k."""
    s = v.sum(dim=field_axis)
    return 0.5 * (s * s - (v * v).sum(dim=field_axis))


class SampledPort(nn.Module):
    """Sampled-ELBO VFM (vfm-torch.py:129-324).

    Parameter names, shapes and initialisation order match the reference so
    that ``state_dict`` round-trips and ``torch.manual_seed`` gives the same
    initial values (vfm-torch.py:136-153)."""

    def __init__(self, n_users: int, n_items: int, embedding_size: int,
                 train_counts: torch.Tensor, output: str = "reg", n_var_samples: int = 1,
                 link: str = "abs", field_sizes: Optional[Sequence[int]] = None,
                 interaction: str = "prod", kl_weighting: str = "torch",
                 faithful_cost: bool = True):
        super().__init__()
        self.N, self.M, self.d = n_users, n_items, embedding_size
        self.S, self.output, self.link = n_var_samples, output, _LINKS[link]
        self.field_sizes = list(field_sizes) if field_sizes is not None else [n_users, n_items]
        self.interaction, self.kl_weighting = interaction, kl_weighting
        self.faithful_cost = faithful_cost
        self.register_buffer("train_counts", train_counts.clone())
        rows = sum(self.field_sizes)
        # vfm-torch.py:136-145 (five prec_* parameters exist but are never used)
        self.alpha = nn.Parameter(torch.Tensor([1e9]))
        self.global_bias_mean = nn.Parameter(torch.Tensor([0.]))
        self.global_bias_scale = nn.Parameter(torch.Tensor([1.]))
        self.prec_global_bias_prior = nn.Parameter(torch.Tensor([1.]))
        self.prec_user_bias_prior = nn.Parameter(torch.Tensor([1.]))
        self.prec_item_bias_prior = nn.Parameter(torch.Tensor([1.]))
        self.prec_user_entity_prior = nn.Parameter(torch.ones(embedding_size))
        self.prec_item_entity_prior = nn.Parameter(torch.ones(embedding_size))
        nn.init.uniform_(self.alpha)
        self.bias_params = nn.Embedding(rows, 2)                    # [mean, raw scale]
        self.entity_params = nn.Embedding(rows, 2 * embedding_size)  # [mean(d) | raw scale(d)]

    def forward(self, x: torch.Tensor, noise: Optional[Sequence[torch.Tensor]] = None):
        d, S, link = self.d, self.S, self.link
        F = x.shape[1]
        # plan: vfm-torch.py:190-192
        uniq, inverse, cnt = torch.unique(x, return_inverse=True, return_counts=True)
        col = [torch.unique(x[:, f], return_counts=True) for f in range(F)]
        if self.faithful_cost:          # dead gathers of vfm-torch.py:205-206
            self.bias_params(x), self.entity_params(x)
        b_rows = self.bias_params(uniq)
        e_rows = self.entity_params(uniq)
        q0 = Normal(self.global_bias_mean, link(self.global_bias_scale))
        qb = Normal(b_rows[:, 0], link(b_rows[:, 1]))
        qe = Normal(e_rows[:, :d], link(e_rows[:, d:]))
        if noise is None:               # vfm-torch.py:238-241, draw order preserved
            w0, w, v = q0.rsample((S,)), qb.rsample((S,)), qe.rsample((S,))
        else:
            e0, eb, ee = noise
            w0 = q0.loc + e0 * q0.scale
            w = qb.loc + eb * qb.scale
            v = qe.loc + ee * qe.scale
        bias_term = w[:, inverse].sum(dim=2).mean(dim=0)
        vg = v[:, inverse]                                     # [S,B,F,d]
        if self.interaction == "prod":  # vfm-torch.py:245
            fm = vg.prod(dim=2).sum(dim=2).mean(dim=0)
        else:
            fm = _pairwise(vg, 2).sum(dim=2).mean(dim=0)
        pred = w0 + bias_term + fm                             # [S,B] (vfm-torch.py:265)
        if self.output == "reg":
            lik = Normal(pred, torch.sqrt(1 / link(self.alpha)))
        else:
            lik = Bernoulli(logits=pred)
        prior = Normal(0., 1.)
        kl_rows = kl_divergence(qb, prior) + kl_divergence(qe, prior).sum(dim=1)
        tc = self.train_counts
        ratio = cnt / tc[uniq]
        z = [(c / tc[u]).sum(dim=0) for (u, c) in col]
        if self.kl_weighting == "torch":   # vfm-torch.py:314-317 incl. the `<= N` quirk
            factor = (uniq <= self.N) * self.N / z[0] + (uniq > self.N) * self.M / z[1]
        else:                              # per-group, vfm-tomasrch.py:574-587
            bounds = torch.tensor([0] + list(self.field_sizes)).cumsum(0)
            grp = torch.bucketize(uniq, bounds[1:], right=True)
            factor = torch.stack([self.field_sizes[g] / z[g] for g in range(F)])[grp]
        kl = kl_divergence(q0, prior) + (kl_rows * ratio * factor).sum(dim=0)
        return lik, kl, {"uniq": uniq, "inverse": inverse, "counts": cnt,
                         "col_uniq": [u for u, _ in col], "col_counts": [c for _, c in col],
                         "pred": pred}


def sampled_loss(lik, kl, y: torch.Tensor, n_train: int) -> torch.Tensor:
    """vfm-torch.py:359."""
    return -lik.log_prob(y.float()).mean() * n_train + kl


def sampled_port_step(model: SampledPort, optimizer, x, y, n_train: int, noise=None,
                      update: bool = True) -> dict:
    lik, kl, aux = model(x, noise)
    loss = sampled_loss(lik, kl, y, n_train)
    out = {"loss": loss.detach().clone(), "kl": kl.detach().clone(),
           "pred": lik.mean.squeeze().detach().clone(), **{k: v for k, v in aux.items() if k != "pred"}}
    if update:
        optimizer.zero_grad()
        loss.backward()
        out["grads"] = {k: (p.grad.detach().clone() if p.grad is not None else None)
                        for k, p in model.named_parameters()}
        optimizer.step()
    return out


class ClosedPort(nn.Module):
    """Closed-form Gaussian VFM with learnable per-group priors
    (vfm-tomasrch.py:186-453).  Initialisation order follows :194-260."""

    def __init__(self, embedding_size: int, group_sizes: Sequence[int], alpha_0: float = 300.,
                 output: str = "reg", start_scale: float = 0.2):
        super().__init__()
        self.d, self.G, self.group_sizes = embedding_size, len(group_sizes), list(group_sizes)
        self.output = output
        d = embedding_size
        self.alpha = nn.Parameter(torch.Tensor([alpha_0]))
        self.mean_global_bias_prior = nn.Parameter(torch.Tensor([0.]))
        self.scale_global_bias_prior = nn.Parameter(torch.Tensor([1.]))
        self.mean_global_bias = nn.Parameter(torch.normal(torch.zeros(1), torch.ones(1)))
        self.scale_global_bias = nn.Parameter(torch.Tensor([start_scale]))
        self.mean_group_bias_prior = nn.ParameterList(
            [nn.Parameter(torch.zeros(1)) for _ in group_sizes])
        self.scale_group_bias_prior = nn.ParameterList(
            [nn.Parameter(torch.ones(1)) for _ in group_sizes])
        self.bias_params = nn.Parameter(torch.cat([
            torch.cat((torch.normal(torch.zeros(n, 1), 1e-1 * torch.ones(n, 1)),
                       start_scale * torch.ones(n, 1)), dim=1) for n in group_sizes]))
        self.mean_group_entity_prior = nn.ParameterList(
            [nn.Parameter(torch.zeros(d)) for _ in group_sizes])
        self.scale_group_entity_prior = nn.ParameterList(
            [nn.Parameter(torch.ones(d)) for _ in group_sizes])
        self.entity_params = nn.Parameter(torch.cat([
            torch.cat((torch.normal(torch.zeros(n, d), 1e-7 * torch.ones(n, d)),
                       start_scale * torch.ones(n, d)), dim=1) for n in group_sizes]))

    def forward(self, x: torch.Tensor, y: Optional[torch.Tensor] = None):
        """``x`` is the raw ``[B,G]`` id batch; the per-column plan of the
        script's loop (vfm-tomasrch.py:536-545) is built here."""
        d, G = self.d, self.G
        plan = [torch.unique(x[:, g], return_inverse=True, return_counts=True) for g in range(G)]
        present = [p for p, _, _ in plan]
        sizes = [len(p) for p in present]
        rows = torch.cat(present)
        # priors broadcast per group (update_priors, :262-290)
        p0 = Normal(self.mean_global_bias_prior, torch.abs(self.scale_global_bias_prior))
        pb = Normal(torch.cat([self.mean_group_bias_prior[g].repeat(sizes[g]) for g in range(G)]),
                    torch.abs(torch.cat([self.scale_group_bias_prior[g].repeat(sizes[g])
                                         for g in range(G)])))
        pe = Normal(torch.cat([self.mean_group_entity_prior[g].repeat(sizes[g], 1)
                               for g in range(G)]),
                    torch.abs(torch.cat([self.scale_group_entity_prior[g].repeat(sizes[g], 1)
                                         for g in range(G)])))
        # posteriors (draw, :292-313) -- never sampled
        q0 = Normal(self.mean_global_bias, torch.abs(self.scale_global_bias))
        br, er = self.bias_params[rows], self.entity_params[rows]
        qb = Normal(br[:, 0], torch.abs(br[:, 1]))
        qe = Normal(er[:, :d], torch.abs(er[:, d:]))
        # prediction from posterior means: product over groups (:327-348)
        offs = [sum(sizes[:g]) for g in range(G)]
        pos = [plan[g][1] + offs[g] for g in range(G)]
        bias = torch.stack([qb.mean.index_select(0, p) for p in pos], dim=1)
        ent = torch.stack([qe.mean.index_select(0, p) for p in pos], dim=1)
        pred = q0.mean + bias.sum(dim=1) + ent.prod(dim=1).sum(dim=1)
        if self.output == "reg":
            lik = Normal(pred, torch.sqrt(1 / torch.abs(self.alpha)))
        else:
            lik = Bernoulli(logits=pred)
        kls = [kl_divergence(q0, p0), kl_divergence(qb, pb), kl_divergence(qe, pe)]
        aux = {"present": present, "inverse": [i for _, i, _ in plan],
               "counts": [c for _, _, c in plan]}
        if y is None:
            return lik, kls, None, aux
        # closed-form expected squared error (:369-449)
        # the reference re-gathers every [B,d] slice inside the pair loop; keep that cost
        mu = lambda g: self.entity_params[x[:, g], :d]
        rho2 = lambda g: self.entity_params[x[:, g], d:] ** 2
        y_bar = self.mean_global_bias + sum(self.bias_params[x[:, g], 0] for g in range(G))
        t_n = self.scale_global_bias ** 2 + sum(self.bias_params[x[:, g], 1] ** 2 for g in range(G))
        for i in range(G):
            for j in range(i + 1, G):
                y_bar = y_bar + torch.einsum("ab,ab->a", mu(i), mu(j))
                t_n = t_n + (torch.einsum("ab,ab->a", mu(i) ** 2, rho2(j))
                             + torch.einsum("ab,ab->a", mu(j) ** 2, rho2(i))
                             + torch.einsum("ab,ab->a", rho2(i), rho2(j)))
        a = torch.abs(self.alpha)
        partial = (0.5 * a.log() - a / 2 * ((y - y_bar) ** 2 + t_n)).sum()
        return lik, kls, partial, aux


def closed_loss(model: ClosedPort, kls, partial, aux, n_train: int, batch: int,
                train_counts: torch.Tensor) -> torch.Tensor:
    """vfm-tomasrch.py:569-588."""
    G = model.G
    present, counts = aux["present"], aux["counts"]
    w = torch.cat([(model.group_sizes[g] / (counts[g] / train_counts[present[g]]).sum()
                    ).reshape(1).repeat(len(present[g])) for g in range(G)])
    rescaled = ((kls[1] + kls[2].sum(dim=1)) * w * torch.cat(counts)
                / train_counts[torch.cat(present)]).sum()
    return -n_train * partial / batch + kls[0] + rescaled


def closed_port_step(model: ClosedPort, optimizer, x, y, n_train: int,
                     train_counts: torch.Tensor, update: bool = True) -> dict:
    lik, kls, partial, aux = model(x, y)
    loss = closed_loss(model, kls, partial, aux, n_train, len(x), train_counts)
    out = {"loss": loss.detach().clone(), "pred": lik.mean.detach().clone(),
           "partial_loss": partial.detach().clone(), **aux}
    if update:
        optimizer.zero_grad()
        loss.backward()
        out["grads"] = {k: (p.grad.detach().clone() if p.grad is not None else None)
                        for k, p in model.named_parameters()}
        optimizer.step()
    return out


def adam_bias_corrections(step: int, lr: float, beta1: float = 0.9, beta2: float = 0.999):
    """Scalars of torch's ``_single_tensor_adam`` for 1-based ``step``."""
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    return lr / bc1, math.sqrt(bc2)
