"""Data-parallel mode A on the GPU: two ranks emulated on one device (their packed buffers are
added instead of all-reduced) must reproduce the single-process reference step on the global
batch, including the reference's dense Adam over every row."""
import numpy as np
import pytest
import torch

import golden_util as gu
import test_gpu_sampled as S

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _local_noise(noise, uniq_global, uniq_local):
    idx = torch.from_numpy(np.searchsorted(uniq_global, uniq_local)).to(DEV)
    return [noise[0], noise[1][:, idx].contiguous(), noise[2][:, idx].contiguous()]


@pytest.mark.parametrize("name,t", [("sampled_reg_d64", 0), ("sampled_reg_d64", 1), ("sampled_reg_d5", 2),
                                    ("sampled_fraction", 1)])
def test_two_emulated_ranks_equal_reference_dense_adam(name, t):
    from vae_b200.dist import DataParallelSampled, local_slice
    meta, g = gu.load(name)
    m = S._model(meta, g, t)
    x, y = gu.batch_of(meta, g, t)
    if len(x) % 2:
        x, y = x[:-1], y[:-1]
    noise = S._noise(g, t)
    even = len(x) == len(gu.batch_of(meta, g, t)[0])
    uniq_g = g[f"step{t}.uniq"] if even else np.unique(x)
    dp = DataParallelSampled(m, world=2, dense_adam=True)
    total = None
    for rank in range(2):
        sl = local_slice(len(x), rank, 2)
        xl, yl = x[sl], y[sl]
        nl = _local_noise(noise, g[f"step{t}.uniq"], np.unique(xl))
        flat = dp.local_backward(torch.from_numpy(xl).to(DEV), torch.from_numpy(yl).to(DEV), nl).clone()
        total = flat if total is None else total + flat
    out = dp.apply(total)
    assert int(m.adam_step.item()) == t + 1
    if not even:
        return
    np.testing.assert_allclose(out["loss"].item(), g[f"step{t}.loss"][0], rtol=1e-5)
    after = gu.state(g, f"step{t}.after")
    lr = meta["lr"]
    for key in ("entity_params.weight", "bias_params.weight"):        # dense Adam: EVERY row
        got, want = m.state_dict()[key].cpu().numpy(), after[key]
        bad = np.abs(got - want) > 1e-5 * np.abs(want) + 2e-5 * lr + 1e-6
        assert bad.mean() <= 1e-4, (name, t, key, float(bad.mean()))
    for key in ("global_bias_mean", "global_bias_scale") + (("alpha",) if meta["output"] == "reg" else ()):
        np.testing.assert_allclose(m.state_dict()[key].cpu().numpy(), after[key], rtol=1e-5, atol=2e-5 * lr + 1e-6)


def test_touched_mode_equals_single_process_fused_step():
    from vae_b200.dist import DataParallelSampled, local_slice
    meta, g = gu.load("sampled_reg_d64")
    x, y = gu.batch_of(meta, g, 1)
    noise = S._noise(g, 1)
    ref = S._model(meta, g, 1)
    ref.fused_step(torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV), noise=noise)
    m = S._model(meta, g, 1)
    dp = DataParallelSampled(m, world=2, dense_adam=False)
    total = None
    for rank in range(2):
        sl = local_slice(len(x), rank, 2)
        nl = _local_noise(noise, g["step1.uniq"], np.unique(x[sl]))
        flat = dp.local_backward(torch.from_numpy(x[sl]).to(DEV), torch.from_numpy(y[sl]).to(DEV), nl).clone()
        total = flat if total is None else total + flat
    out = dp.apply(total)
    assert torch.allclose(m.entity_params.weight, ref.entity_params.weight, rtol=1e-5, atol=2e-6)
    assert torch.allclose(m.bias_params.weight, ref.bias_params.weight, rtol=1e-5, atol=2e-6)
    assert torch.allclose(m._scalars, ref._scalars, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(out["loss"].item(), ref._buf.stats[0].item(), rtol=1e-5)
