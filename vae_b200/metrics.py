"""Evaluation metrics of the reference's epoch end (vfm-torch.py:377-384, 402-422), computed on the
device the predictions live on (no host round trip): RMSE for the regression variant, ROC AUC and
average precision for the Bernoulli variant.  The reference calls sklearn
(``mean_squared_error ** 0.5``, ``roc_auc_score``, ``average_precision_score``); these functions
return the same numbers, ties included (``tests/test_metrics.py`` checks them against sklearn).

Plain tensor code (sort + prefix sums): evaluation runs once per displayed epoch and is not on the
measured path; the kernels of the training step are in ``csrc/``."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch


def rmse(truth: torch.Tensor, pred: torch.Tensor, clip: Optional[Tuple[float, float]] = None) -> torch.Tensor:
    """``mean_squared_error(truth, pred) ** 0.5`` (vfm-torch.py:381); ``clip=(1, 5)`` applies the
    reference's ``np.clip(pred, 1, 5)`` (vfm-torch.py:379, 405) first."""
    pred = pred.reshape(-1).to(torch.float64)
    if clip is not None:
        pred = pred.clamp(clip[0], clip[1])
    return ((truth.reshape(-1).to(torch.float64) - pred) ** 2).mean().sqrt()


def _groups(truth: torch.Tensor, score: torch.Tensor):
    """Distinct score values in descending order with their positive / total counts."""
    score = score.reshape(-1).to(torch.float64)
    pos = (truth.reshape(-1) > 0).to(torch.float64)
    order = torch.argsort(score, descending=True, stable=True)
    s, p = score[order], pos[order]
    _, counts = torch.unique_consecutive(s, return_counts=True)
    ends = torch.cumsum(counts, 0) - 1                     # last element of every tie group
    tp = torch.cumsum(p, 0)[ends]                          # positives with score >= the group's value
    n = (ends + 1).to(torch.float64)                       # samples   with score >= the group's value
    return tp, n, pos.sum(), pos.numel()


def roc_auc(truth: torch.Tensor, score: torch.Tensor) -> torch.Tensor:
    """``roc_auc_score(truth, score)`` (vfm-torch.py:383, 420): trapezoidal area under the ROC curve
    whose thresholds are the distinct score values (ties form one point)."""
    tp, n, n_pos, n_all = _groups(truth, score)
    n_neg = n_all - n_pos
    if float(n_pos) == 0.0 or float(n_neg) == 0.0:
        raise ValueError("roc_auc: only one class present in truth")
    fp = n - tp
    zero = tp.new_zeros(1)
    tpr, fpr = torch.cat((zero, tp / n_pos)), torch.cat((zero, fp / n_neg))
    return ((fpr[1:] - fpr[:-1]) * (tpr[1:] + tpr[:-1]) * 0.5).sum()


def average_precision(truth: torch.Tensor, score: torch.Tensor) -> torch.Tensor:
    """``average_precision_score(truth, score)`` (vfm-torch.py:384, 421):
    ``sum_k (R_k - R_{k-1}) P_k`` over the distinct score thresholds, descending."""
    tp, n, n_pos, _ = _groups(truth, score)
    if float(n_pos) == 0.0:
        raise ValueError("average_precision: no positive sample in truth")
    precision, recall = tp / n, tp / n_pos
    prev = torch.cat((recall.new_zeros(1), recall[:-1]))
    return ((recall - prev) * precision).sum()


@torch.no_grad()
def evaluate(model, x_test: torch.Tensor, y_test: torch.Tensor, all_preds: Optional[list] = None,
             clip: Sequence[float] = (1.0, 5.0)) -> dict:
    """The test block of the reference's display epoch (vfm-torch.py:402-422) on a drop-in sampled
    ``CF``: a SAMPLED forward on the whole test set (N4: the reference samples at test time too),
    plus -- for regression, once ``model.save_weights()`` has been called -- the deterministic
    predictions from the last and the epoch-averaged posterior means.  ``all_preds`` is the caller's
    running list of per-epoch predictions (``all_preds`` in the script) for ``rmse_all``."""
    likelihood, last_logits, mean_logits, _ = model(x_test)
    y_pred = likelihood.mean.squeeze().detach()
    y = y_test.to(y_pred.device)
    if model.output == "reg":
        out = {"rmse": rmse(y, y_pred, clip)}
        if all_preds is not None:
            all_preds.append(y_pred.clamp(clip[0], clip[1]))
            out["rmse_all"] = rmse(y, torch.stack(all_preds).mean(dim=0))
        if last_logits is not None:
            out["rmse_of_last"] = rmse(y, last_logits)           # (the script clips only y_pred_of_mean)
            out["rmse_of_mean"] = rmse(y, mean_logits, clip)
        return out
    return {"auc": roc_auc(y, y_pred), "map": average_precision(y, y_pred)}
