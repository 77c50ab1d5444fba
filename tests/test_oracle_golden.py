"""Pin the oracle (torch port + fp64 maths) to the golden vectors produced by
the UNMODIFIED reference classes (``oracle/gen_golden.py``).  CPU only."""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import vfm_math, vfm_port


@pytest.fixture(autouse=True)
def _one_thread():
    """ATen's CPU scatter-add order depends on the thread count; the goldens are
    the single-thread result (oracle/gen_golden.py), so replay them the same way."""
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


@pytest.mark.parametrize("name", gu.SAMPLED)
def test_sampled_port_and_math_match_reference_golden(name):
    meta, g = gu.load(name)
    N, M, d, S = meta["N"], meta["M"], meta["d"], meta["S"]
    tc = torch.from_numpy(g["train_counts"])
    port = vfm_port.SampledPort(N, M, d, tc, output=meta["output"], n_var_samples=S,
                                link=meta["link"])
    for t in range(meta["steps"]):
        opt = gu.restore(port, g, t, meta["lr"])      # replay each step from its exact pre-state
        x, y = gu.batch_of(meta, g, t)
        noise = [torch.from_numpy(g[f"step{t}.noise{i}"]) for i in range(3)]
        sd_before = {k: v.detach().numpy().copy() for k, v in port.state_dict().items()}
        out = vfm_port.sampled_port_step(port, opt, torch.from_numpy(x), torch.from_numpy(y),
                                         meta["n_train"], noise)
        # integer plan: bit-exact
        assert np.array_equal(out["uniq"].numpy(), g[f"step{t}.uniq"])
        assert np.array_equal(out["inverse"].numpy(), g[f"step{t}.inverse"])
        assert np.array_equal(out["counts"].numpy(), g[f"step{t}.counts"])
        np.testing.assert_allclose(out["loss"].numpy(), g[f"step{t}.loss"], rtol=1e-6)
        np.testing.assert_allclose(out["pred"].numpy(), g[f"step{t}.pred"], rtol=1e-5, atol=1e-6)
        # independent fp64 maths from the same pre-step state
        m = vfm_math.sampled_step(gu.sampled_math_params(sd_before), x, y,
                                  [n.numpy() for n in noise], g["train_counts"], meta["n_train"],
                                  [N, M], output=meta["output"], link=meta["link"])
        assert np.array_equal(m["plan"]["uniq"], g[f"step{t}.uniq"])
        assert np.array_equal(m["plan"]["inverse"], g[f"step{t}.inverse"])
        np.testing.assert_allclose(m["loss"], g[f"step{t}.loss"][0], rtol=2e-6)
        np.testing.assert_allclose(m["mean"].squeeze(), g[f"step{t}.pred"], rtol=1e-5, atol=2e-6)
        if t == 0:
            assert gu.rel_err(m["grads"]["entity"], g["step0.grad.entity_params.weight"]) < 2e-5
            assert gu.rel_err(m["grads"]["bias"], g["step0.grad.bias_params.weight"]) < 2e-5
            for k in ("global_bias_mean", "global_bias_scale"):
                assert gu.rel_err(m["grads"][k], g[f"step0.grad.{k}"]) < 2e-5
            if meta["output"] == "reg":
                assert gu.rel_err(m["grads"]["alpha"], g["step0.grad.alpha"]) < 2e-5
            else:
                assert "step0.grad.alpha" not in g and "alpha" not in m["grads"]
        for k, v in gu.state(g, f"step{t}.after").items():
            np.testing.assert_allclose(port.state_dict()[k].numpy(), v, rtol=1e-5, atol=2e-6 * max(1.0, meta["lr"]),
                                       err_msg=f"{name} step {t} {k}")


@pytest.mark.parametrize("name", gu.CLOSED)
def test_closed_port_and_math_match_reference_golden(name):
    meta, g = gu.load(name)
    fs, d = meta["group_sizes"], meta["d"]
    G = len(fs)
    tc = torch.from_numpy(g["train_counts"]).float()
    port = vfm_port.ClosedPort(d, fs, alpha_0=meta["alpha_0"])
    for t in range(meta["steps"]):
        opt = gu.restore(port, g, t, meta["lr"])
        x, y = gu.batch_of(meta, g, t)
        sd_before = {k: v.detach().numpy().copy() for k, v in port.state_dict().items()}
        out = vfm_port.closed_port_step(port, opt, torch.from_numpy(x), torch.from_numpy(y),
                                        meta["n_train"], tc)
        for gi in range(G):
            assert np.array_equal(out["present"][gi].numpy(), g[f"step{t}.present{gi}"])
            assert np.array_equal(out["inverse"][gi].numpy(), g[f"step{t}.inverse{gi}"])
            assert np.array_equal(out["counts"][gi].numpy(), g[f"step{t}.counts{gi}"])
        np.testing.assert_allclose(out["loss"].numpy(), g[f"step{t}.loss"], rtol=2e-6)
        np.testing.assert_allclose(out["pred"].numpy(), g[f"step{t}.pred"], rtol=1e-5, atol=1e-6)
        m = vfm_math.closed_step(gu.closed_math_params(sd_before, G), x, y,
                                 g["train_counts"].astype(np.float64), meta["n_train"], fs)
        np.testing.assert_allclose(m["loss"], g[f"step{t}.loss"], rtol=5e-6)
        np.testing.assert_allclose(m["pred"], g[f"step{t}.pred"], rtol=1e-5, atol=2e-6)
        # the concatenated per-group unique lists equal the global sorted unique (disjoint id ranges)
        assert np.array_equal(m["plan"]["uniq"],
                              np.concatenate([g[f"step{t}.present{gi}"] for gi in range(G)]))
        if t == 0:
            assert gu.rel_err(m["grads"]["entity"], g["step0.grad.entity_params"]) < 2e-5
            assert gu.rel_err(m["grads"]["bias"], g["step0.grad.bias_params"]) < 2e-5
            for k in ("alpha", "mean_global_bias", "scale_global_bias",
                      "mean_global_bias_prior", "scale_global_bias_prior"):
                assert gu.rel_err(m["grads"][k], g[f"step0.grad.{k}"]) < 2e-5, k
            for gi in range(G):
                assert gu.rel_err(m["grads"]["prior_entity_mean"][gi],
                                  g[f"step0.grad.mean_group_entity_prior.{gi}"]) < 2e-5
                assert gu.rel_err(m["grads"]["prior_entity_scale"][gi],
                                  g[f"step0.grad.scale_group_entity_prior.{gi}"]) < 2e-5
                assert gu.rel_err(m["grads"]["prior_bias_mean"][gi],
                                  g[f"step0.grad.mean_group_bias_prior.{gi}"]) < 2e-5
                assert gu.rel_err(m["grads"]["prior_bias_scale"][gi],
                                  g[f"step0.grad.scale_group_bias_prior.{gi}"]) < 2e-5
        for k, v in gu.state(g, f"step{t}.after").items():
            np.testing.assert_allclose(port.state_dict()[k].numpy(), v, rtol=2e-5, atol=2e-6,
                                       err_msg=f"{name} step {t} {k}")


def test_adam_recurrence_matches_torch():
    rng = np.random.default_rng(0)
    p0 = rng.standard_normal((7, 5))
    p = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    opt = torch.optim.Adam([p], lr=0.05)
    pm, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    for step in range(1, 5):
        g = rng.standard_normal((7, 5))
        p.grad = torch.from_numpy(g.copy())
        opt.step()
        pm, m, v = vfm_math.adam_update(pm, g, m, v, step, 0.05)
        np.testing.assert_allclose(pm, p.detach().numpy(), rtol=1e-12)


def test_saved_weight_logits_golden_is_the_mean_of_the_snapshots():
    """vfm-torch.py:179-185, 248-262: last_logits / mean_logits of the golden are the FM prediction from
    the last / the averaged posterior-mean snapshots stored beside them (restated here in numpy)."""
    meta, g = gu.load("sampled_saved_logits")
    d, x = meta["d"], g["x_eval"].astype(np.int64)
    snaps = [gu.state(g, f"step{t}.after") for t in range(meta["steps"])]
    gb = [s["global_bias_mean"] for s in snaps]
    mb = [s["bias_params.weight"][:, 0] for s in snaps]
    me = [s["entity_params.weight"][:, :d] for s in snaps]
    fm = lambda g0, b, e: g0 + b[x].sum(axis=1) + e[x].prod(axis=1).sum(axis=1)
    np.testing.assert_allclose(fm(gb[-1], mb[-1], me[-1]), g["last_logits"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(fm(np.mean(gb, 0), np.mean(mb, 0), np.mean(me, 0)), g["mean_logits"], rtol=1e-5, atol=1e-6)
