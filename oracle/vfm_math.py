"""Independent fp64 numpy restatement of the VFM step with hand-derived
gradients -- the arithmetic specification the CUDA kernels implement.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

Nothing here uses autograd: the forward follows the cited reference lines,
the backward is the closed-form derivative, and Adam is
torch's ``_single_tensor_adam`` recurrence.  It is cross-checked against
autograd on the reference classes (``tests/test_oracle_vs_reference.py``)
and against the torch port, and is the only oracle for the F>2 pairwise
sampled model, which the reference does not implement.

Reference lines: sampled ``vfm-torch.py:189-324, 359``; closed form
``vfm-tomasrch.py:262-453, 569-588``.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

LOG_2PI = float(np.log(2.0 * np.pi))


# ----------------------------------------------------------------------------- plan
def build_plan(x: np.ndarray) -> Dict[str, np.ndarray]:
    """Sorted-unique plan of a ``[B,F]`` id batch.

    ``uniq/inverse/counts`` equal ``torch.unique(x, return_inverse=True,
    return_counts=True)`` (vfm-torch.py:190).  ``order`` lists the B*F
    occurrences (flat index n*F+f) grouped by unique row, ascending inside a
    row; ``seg_off[u]:seg_off[u+1]`` is row u's segment.  ``field_counts[u,f]``
    counts occurrences of row u in column f (vfm-torch.py:191-192,
    vfm-tomasrch.py:538-542)."""
    B, F = x.shape
    flat = x.reshape(-1)
    uniq, inverse, counts = np.unique(flat, return_inverse=True, return_counts=True)
    order = np.argsort(inverse, kind="stable")
    seg_off = np.concatenate(([0], np.cumsum(counts)))
    field_counts = np.zeros((len(uniq), F), dtype=np.int64)
    np.add.at(field_counts, (inverse, np.tile(np.arange(F), B)), 1)
    return {"uniq": uniq.astype(np.int64), "inverse": inverse.reshape(B, F).astype(np.int64),
            "counts": counts.astype(np.int64), "order": order.astype(np.int64),
            "seg_off": seg_off.astype(np.int64), "field_counts": field_counts}


# ----------------------------------------------------------------------------- links
def _link(raw, kind):
    if kind == "abs":
        return np.abs(raw)
    return np.logaddexp(0.0, raw)            # softplus


def _dlink(raw, kind):
    if kind == "abs":
        return np.sign(raw)                    # sign(0) = 0, as torch.abs
    return 1.0 / (1.0 + np.exp(-raw))


def _kl(m, s, pm=0.0, ps=1.0):
    """KL(N(m,s) || N(pm,ps)) as torch ``_kl_normal_normal``."""
    vr = (s / ps) ** 2
    return 0.5 * (vr + ((m - pm) / ps) ** 2 - 1.0 - np.log(vr))


def _group_of(uniq: np.ndarray, field_sizes: Sequence[int]) -> np.ndarray:
    bounds = np.cumsum(field_sizes)
    return np.searchsorted(bounds, uniq, side="right")


def kl_weights(plan, train_counts: np.ndarray, field_sizes: Sequence[int], weighting: str):
    """c_u of SURVEY section 8-Maths.  ``"torch"``: vfm-torch.py:298-317 with the
    ``uniq <= N`` selection; ``"group"``: vfm-tomasrch.py:574-587."""
    uniq, cnt, fc = plan["uniq"], plan["counts"], plan["field_counts"]
    tc = train_counts[uniq].astype(np.float64)
    z = (fc / tc[:, None]).sum(axis=0)                        # per column
    if weighting == "torch":
        N, M = field_sizes[0], field_sizes[1]
        factor = np.where(uniq <= N, N / z[0], M / z[1])
    else:
        g = _group_of(uniq, field_sizes)
        factor = (np.asarray(field_sizes, dtype=np.float64) / z)[g]
    return cnt / tc * factor, z


# ----------------------------------------------------------------------------- sampled
def sampled_step(params: Dict[str, np.ndarray], x: np.ndarray, y: np.ndarray,
                 noise: Sequence[np.ndarray], train_counts: np.ndarray, n_train: int,
                 field_sizes: Sequence[int], output: str = "reg", link: str = "abs",
                 interaction: str = "prod", kl_weighting: str = "torch", kl_scale: float = 1.0,
                 kl_weight: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
    """Forward + analytic backward of the sampled ELBO.

    ``kl_scale`` multiplies every KL contribution (0 = data term only, as each rank computes it
    in data-parallel mode A); ``kl_weight`` overrides the per-unique-row KL weights c_u (e.g. the
    global-batch weights restricted to this batch's rows).

    ``params``: ``alpha[1], global_bias_mean[1], global_bias_scale[1],
    bias[R,2], entity[R,2d]``.  ``noise`` = (eps0 ``[S,1]``, eps_bias ``[S,U]``,
    eps_entity ``[S,U,d]``), indexed by unique rank (one draw per unique row,
    vfm-torch.py:238-241).  Returns loss terms, predictions and gradients
    (dense tables, zero on untouched rows)."""
    f8 = np.float64
    alpha, mu0, rho0 = (params[k].astype(f8)[0] for k in
                        ("alpha", "global_bias_mean", "global_bias_scale"))
    bias, ent = params["bias"].astype(f8), params["entity"].astype(f8)
    d = ent.shape[1] // 2
    B, F = x.shape
    plan = build_plan(x)
    uniq, inv = plan["uniq"], plan["inverse"]
    U = len(uniq)
    e0, eb, ee = (np.asarray(t, dtype=f8) for t in noise)
    S = e0.shape[0]
    a, b = bias[uniq, 0], bias[uniq, 1]
    mu, rho = ent[uniq, :d], ent[uniq, d:]
    sig0, tau, sig = _link(rho0, link), _link(b, link), _link(rho, link)
    w0 = mu0 + e0[:, 0] * sig0                                  # [S]
    w = a[None] + eb * tau[None]                                # [S,U]
    v = mu[None] + ee * sig[None]                               # [S,U,d]
    vg = v[:, inv]                                              # [S,B,F,d]
    if interaction == "prod":
        inter = vg.prod(axis=2).sum(axis=2)
    else:
        ssum = vg.sum(axis=2)
        inter = (0.5 * (ssum ** 2 - (vg ** 2).sum(axis=2))).sum(axis=2)
    h = w[:, inv].sum(axis=2) + inter                           # [S,B]
    pred = w0[:, None] + h.mean(axis=0)[None]                   # [S,B]
    yy = y.astype(f8)[None]
    if output == "reg":
        ap = _link(alpha, link)
        nll = 0.5 * ap * (yy - pred) ** 2 - 0.5 * np.log(ap) + 0.5 * LOG_2PI
        dnll = ap * (pred - yy)
        mean_out = pred
    else:
        nll = np.logaddexp(0.0, pred) - yy * pred
        dnll = 1.0 / (1.0 + np.exp(-pred)) - yy
        mean_out = 1.0 / (1.0 + np.exp(-pred))
    c, z = kl_weights(plan, train_counts, field_sizes, kl_weighting)
    if kl_weight is not None:
        c = np.asarray(kl_weight, dtype=f8)
    c = kl_scale * c
    kl_rows = _kl(a, tau) + _kl(mu, sig).sum(axis=1)
    kl0 = kl_scale * _kl(mu0, sig0)
    kl = kl0 + (c * kl_rows).sum()
    loss = n_train * nll.mean() + kl

    # ---- backward
    r = n_train / (S * B) * dnll                                # [S,B] dloss/dpred
    rbar = r.mean(axis=0)                                       # dloss/dh[s',n] for every s'
    rsum = r.sum(axis=1)                                        # [S]
    g_mu0 = rsum.sum() + kl_scale * mu0
    g_rho0 = _dlink(rho0, link) * ((e0[:, 0] * rsum).sum() + kl_scale * (sig0 - 1.0 / sig0))
    if output == "reg":
        g_alpha = _dlink(alpha, link) * n_train / (S * B) * (
            0.5 * (yy - pred) ** 2 - 0.5 / ap).sum()
    else:
        g_alpha = None                                          # no gradient
    gw = np.zeros((S, U))
    gv = np.zeros((S, U, d))
    for f in range(F):
        if interaction == "prod":
            others = [g for g in range(F) if g != f]
            partner = vg[:, :, others].prod(axis=2) if others else np.ones_like(vg[:, :, 0])
        else:
            partner = vg.sum(axis=2) - vg[:, :, f]
        for s in range(S):
            np.add.at(gw[s], inv[:, f], rbar)
            np.add.at(gv[s], inv[:, f], rbar[:, None] * partner[s])
    g_a = gw.sum(axis=0) + c * a
    g_b = _dlink(b, link) * ((gw * eb).sum(axis=0) + c * (tau - 1.0 / tau))
    g_mu = gv.sum(axis=0) + c[:, None] * mu
    g_rho = _dlink(rho, link) * ((gv * ee).sum(axis=0) + c[:, None] * (sig - 1.0 / sig))
    g_bias = np.zeros_like(bias)
    g_ent = np.zeros_like(ent)
    g_bias[uniq, 0], g_bias[uniq, 1] = g_a, g_b
    g_ent[uniq, :d], g_ent[uniq, d:] = g_mu, g_rho
    grads = {"global_bias_mean": np.array([g_mu0]), "global_bias_scale": np.array([g_rho0]),
             "bias": g_bias, "entity": g_ent}
    if g_alpha is not None:
        grads["alpha"] = np.array([g_alpha])
    return {"loss": loss, "kl": kl, "nll_mean": nll.mean(), "nll_sum": nll.sum(), "pred": pred,
            "mean": mean_out, "resid": r, "kl_weight": c, "z": z, "plan": plan, "grads": grads,
            "sq_err": ((yy - pred) ** 2).sum()}


# ----------------------------------------------------------------------------- closed form
def closed_step(params: Dict[str, np.ndarray], x: np.ndarray, y: np.ndarray,
                train_counts: np.ndarray, n_train: int, field_sizes: Sequence[int]
                ) -> Dict[str, np.ndarray]:
    """Closed-form expected-Gaussian ELBO (vfm-tomasrch.py:323-453, 569-588).

    ``params``: ``alpha[1], mean_global_bias[1], scale_global_bias[1],
    mean_global_bias_prior[1], scale_global_bias_prior[1], bias[R,2],
    entity[R,2d], prior_bias_mean[G], prior_bias_scale[G],
    prior_entity_mean[G,d], prior_entity_scale[G,d]``."""
    f8 = np.float64
    P = {k: np.asarray(v, dtype=f8) for k, v in params.items()}
    alpha, mu0, rho0 = P["alpha"][0], P["mean_global_bias"][0], P["scale_global_bias"][0]
    m0, s0 = P["mean_global_bias_prior"][0], P["scale_global_bias_prior"][0]
    bias, ent = P["bias"], P["entity"]
    d = ent.shape[1] // 2
    B, F = x.shape
    plan = build_plan(x)
    uniq, inv = plan["uniq"], plan["inverse"]
    g_of = _group_of(uniq, field_sizes)
    a, b = bias[uniq, 0], bias[uniq, 1]
    mu, rho = ent[uniq, :d], ent[uniq, d:]
    tau, sig = np.abs(b), np.abs(rho)
    pbm, pbs = P["prior_bias_mean"][g_of], np.abs(P["prior_bias_scale"][g_of])
    pem, pes = P["prior_entity_mean"][g_of], np.abs(P["prior_entity_scale"][g_of])
    mug, r2g = mu[inv], (rho ** 2)[inv]                         # [B,F,d]
    sm, sm2, sr2 = mug.sum(axis=1), (mug ** 2).sum(axis=1), r2g.sum(axis=1)
    e2 = lambda s1, s2: 0.5 * (s1 ** 2 - s2)
    y_bar = mu0 + a[inv].sum(axis=1) + e2(sm, sm2).sum(axis=1)
    # sum_{i<j} (mu_i^2 rho_j^2 + mu_j^2 rho_i^2 + rho_i^2 rho_j^2) = e2(mu^2+rho^2) - e2(mu^2)
    z2 = mug ** 2 + r2g
    t_n = rho0 ** 2 + (b ** 2)[inv].sum(axis=1) + (
        e2(z2.sum(axis=1), (z2 ** 2).sum(axis=1)) - e2(sm2, (mug ** 4).sum(axis=1))).sum(axis=1)
    ap = abs(alpha)
    delta = y.astype(f8) - y_bar
    partial = (0.5 * np.log(ap) - 0.5 * ap * (delta ** 2 + t_n)).sum()
    pred = mu0 + a[inv].sum(axis=1) + mug.prod(axis=1).sum(axis=1)   # product over groups (:342-348)
    c, z = kl_weights(plan, train_counts, field_sizes, "group")
    kl0 = _kl(mu0, abs(rho0), m0, abs(s0))
    kl_b, kl_e = _kl(a, tau, pbm, pbs), _kl(mu, sig, pem, pes)
    loss = -n_train * partial / B + kl0 + (c * (kl_b + kl_e.sum(axis=1))).sum()

    # ---- backward
    kappa = n_train / B * ap
    U = len(uniq)
    ga, gb = np.zeros(U), np.zeros(U)
    gmu, grho = np.zeros((U, d)), np.zeros((U, d))
    for f in range(F):
        u = inv[:, f]
        psm = sm - mug[:, f]                                    # partner sums over j != f
        psr2 = sr2 - r2g[:, f]
        psm2 = sm2 - mug[:, f] ** 2
        np.add.at(ga, u, -kappa * delta)
        np.add.at(gb, u, kappa * b[u])
        np.add.at(gmu, u, kappa * (-delta[:, None] * psm + mug[:, f] * psr2))
        np.add.at(grho, u, kappa * rho[u] * (psm2 + psr2))
    ga += c * (a - pbm) / pbs ** 2
    gb += np.sign(b) * c * (tau / pbs ** 2 - 1.0 / tau)
    gmu += c[:, None] * (mu - pem) / pes ** 2
    grho += np.sign(rho) * c[:, None] * (sig / pes ** 2 - 1.0 / sig)
    G = len(field_sizes)
    g_pbm, g_pbs = np.zeros(G), np.zeros(G)
    g_pem, g_pes = np.zeros((G, d)), np.zeros((G, d))
    np.add.at(g_pbm, g_of, c * (pbm - a) / pbs ** 2)
    np.add.at(g_pbs, g_of, c * (1.0 / pbs - (tau ** 2 + (a - pbm) ** 2) / pbs ** 3))
    np.add.at(g_pem, g_of, c[:, None] * (pem - mu) / pes ** 2)
    np.add.at(g_pes, g_of, c[:, None] * (1.0 / pes - (sig ** 2 + (mu - pem) ** 2) / pes ** 3))
    g_pbs *= np.sign(P["prior_bias_scale"])
    g_pes *= np.sign(P["prior_entity_scale"])
    as0 = abs(s0)
    g_bias = np.zeros_like(bias)
    g_ent = np.zeros_like(ent)
    g_bias[uniq, 0], g_bias[uniq, 1] = ga, gb
    g_ent[uniq, :d], g_ent[uniq, d:] = gmu, grho
    grads = {
        "alpha": np.array([-np.sign(alpha) * n_train / B * (0.5 / ap - 0.5 * (delta ** 2 + t_n)).sum()]),
        "mean_global_bias": np.array([-kappa * delta.sum() + (mu0 - m0) / as0 ** 2]),
        "scale_global_bias": np.array([kappa * B * rho0
                                       + np.sign(rho0) * (abs(rho0) / as0 ** 2 - 1.0 / abs(rho0))]),
        "mean_global_bias_prior": np.array([(m0 - mu0) / as0 ** 2]),
        "scale_global_bias_prior": np.array([np.sign(s0) * (1.0 / as0 - (rho0 ** 2 + (mu0 - m0) ** 2) / as0 ** 3)]),
        "bias": g_bias, "entity": g_ent,
        "prior_bias_mean": g_pbm, "prior_bias_scale": g_pbs,
        "prior_entity_mean": g_pem, "prior_entity_scale": g_pes,
    }
    return {"loss": loss, "partial_loss": partial, "pred": pred, "y_bar": y_bar, "t_n": t_n,
            "kl_weight": c, "z": z, "plan": plan, "grads": grads}


# ----------------------------------------------------------------------------- Adam
def adam_update(p, g, m, v, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8,
                rows: Optional[np.ndarray] = None):
    """torch ``_single_tensor_adam`` (defaults; vfm-torch.py:339,
    vfm-tomasrch.py:518) for the 1-based ``step``.  ``rows=None`` is the
    reference's dense update; ``rows=uniq`` is the touched-rows ("lazy")
    update of north_star, identical to dense when every row is touched or
    when m = v = 0 on the untouched rows."""
    p, m, v = p.copy(), m.copy(), v.copy()
    sel = slice(None) if rows is None else rows
    gs = g[sel]
    m[sel] = m[sel] + (1.0 - beta1) * (gs - m[sel])
    v[sel] = beta2 * v[sel] + (1.0 - beta2) * gs * gs
    step_size = lr / (1.0 - beta1 ** step)
    bc2_sqrt = np.sqrt(1.0 - beta2 ** step)
    p[sel] = p[sel] - step_size * m[sel] / (np.sqrt(v[sel]) / bc2_sqrt + eps)
    return p, m, v
