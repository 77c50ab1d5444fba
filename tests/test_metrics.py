"""vae_b200.metrics against sklearn (what the reference calls, vfm-torch.py:381-384, 410-422)."""
import numpy as np
import pytest
import torch
from sklearn.metrics import average_precision_score, mean_squared_error, roc_auc_score

from vae_b200 import metrics


@pytest.mark.parametrize("n,ties", [(50, False), (1000, False), (1000, True), (7, True)])
def test_auc_and_map_equal_sklearn(n, ties):
    rng = np.random.default_rng(n + ties)
    y = (rng.random(n) < 0.4).astype(np.float32)
    y[:2] = [0, 1]
    s = rng.random(n).astype(np.float32)
    if ties:
        s = np.round(s * 5) / 5                       # heavy ties, like sigmoid outputs saturating
    got_auc = metrics.roc_auc(torch.from_numpy(y), torch.from_numpy(s)).item()
    got_map = metrics.average_precision(torch.from_numpy(y), torch.from_numpy(s)).item()
    np.testing.assert_allclose(got_auc, roc_auc_score(y, s), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(got_map, average_precision_score(y, s), rtol=1e-12, atol=1e-12)


def test_rmse_equals_sklearn_with_the_reference_clip():
    rng = np.random.default_rng(0)
    y = np.clip(np.round(3.5 + rng.standard_normal(500)), 1, 5)
    p = 3.5 + 2 * rng.standard_normal(500)
    want = mean_squared_error(y, np.clip(p, 1, 5)) ** 0.5          # vfm-torch.py:379-381
    got = metrics.rmse(torch.from_numpy(y), torch.from_numpy(p), clip=(1, 5)).item()
    np.testing.assert_allclose(got, want, rtol=1e-12)
    np.testing.assert_allclose(metrics.rmse(torch.from_numpy(y), torch.from_numpy(p)).item(),
                               mean_squared_error(y, p) ** 0.5, rtol=1e-12)


def test_single_class_raises_like_sklearn():
    with pytest.raises(ValueError):
        metrics.roc_auc(torch.ones(5), torch.rand(5))
