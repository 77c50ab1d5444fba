"""Parity of the CUDA closed-form path (through the C ABI) with the oracle: golden vectors of the
unmodified vfm-tomasrch.py ``CF`` + training loop, the torch port and the fp64 maths."""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import vfm_math, vfm_port

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(meta, g, t=0):
    from vae_b200.vfm_tomasrch import CF
    fs = meta["group_sizes"]
    m = CF(embedding_size=meta["d"], n_groups=len(fs), group_sizes=fs, alpha_0=meta["alpha_0"],
           train_counts=torch.from_numpy(g["train_counts"]), n_train=meta["n_train"],
           max_batch=meta["batch"], lr=meta["lr"])
    _restore(m, g, t)
    return m


def _restore(m, g, t):
    sd = gu.state(g, "init" if t == 0 else f"step{t - 1}.after")
    own = m.state_dict()
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)).reshape(own[k].shape) for k, v in sd.items()},
                      strict=True)
    for buf in (m.bias_m, m.bias_v, m.entity_m, m.entity_v, m._scalars_m, m._scalars_v):
        buf.zero_()
    m.adam_step.fill_(t)
    if t > 0:
        a = lambda k: torch.from_numpy(g[f"step{t - 1}.adam.{k}"]).to(DEV)
        m.bias_m.copy_(a("bias_params.m")), m.bias_v.copy_(a("bias_params.v"))
        m.entity_m.copy_(a("entity_params.m")), m.entity_v.copy_(a("entity_params.v"))
        base = m._scalars.data_ptr()
        for name, p in m.named_parameters():
            if name in ("bias_params", "entity_params"):
                continue
            off = (p.data_ptr() - base) // 4
            m._scalars_m[off:off + p.numel()] = a(f"{name}.m").reshape(-1)
            m._scalars_v[off:off + p.numel()] = a(f"{name}.v").reshape(-1)


def test_seeded_initialisation_equals_reference():
    meta, g = gu.load("closed_3groups")
    from vae_b200.vfm_tomasrch import CF
    fs = meta["group_sizes"]
    torch.manual_seed(42)
    m = CF(embedding_size=meta["d"], n_groups=len(fs), group_sizes=fs, alpha_0=meta["alpha_0"],
           train_counts=torch.from_numpy(g["train_counts"]), n_train=meta["n_train"], max_batch=meta["batch"])
    torch.manual_seed(42)
    port = vfm_port.ClosedPort(meta["d"], fs, alpha_0=meta["alpha_0"])
    assert [k for k, _ in m.named_parameters()] == [k for k, _ in port.named_parameters()]
    for (k, a), (_, b) in zip(m.named_parameters(), port.named_parameters()):
        assert torch.equal(a.detach().cpu(), b.detach()), k


@pytest.mark.parametrize("name", gu.CLOSED)
def test_forward_loss_and_prediction_match_reference_golden(name):
    meta, g = gu.load(name)
    for t in range(meta["steps"]):
        m = _model(meta, g, t)
        x, y = gu.batch_of(meta, g, t)
        out = m.fused_step(torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV), update=False)
        np.testing.assert_allclose(out["loss"].item(), g[f"step{t}.loss"], rtol=1e-5)
        np.testing.assert_allclose(out["nll_mean"].item(), g[f"step{t}.partial_loss"], rtol=1e-5)
        np.testing.assert_allclose(out["kl"].item(), g[f"step{t}.kl"], rtol=1e-5)
        np.testing.assert_allclose(out["pred"].cpu().numpy(), g[f"step{t}.pred"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("name", gu.CLOSED)
def test_gradients_match_reference_and_fp64_maths(name):
    meta, g = gu.load(name)
    fs = meta["group_sizes"]
    G = len(fs)
    m = _model(meta, g, 0)
    x, y = gu.batch_of(meta, g, 0)
    gr = m.gradients(torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV))
    exact = vfm_math.closed_step(gu.closed_math_params(gu.state(g, "init"), G), x, y,
                                 g["train_counts"].astype(np.float64), meta["n_train"], fs)
    np.testing.assert_allclose(gr["loss"].item(), exact["loss"], rtol=1e-5)
    eg = exact["grads"]
    want = {"entity_params": eg["entity"], "bias_params": eg["bias"]}
    for k in ("alpha", "mean_global_bias", "scale_global_bias", "mean_global_bias_prior", "scale_global_bias_prior"):
        want[k] = eg[k]
    for q in range(G):
        want[f"mean_group_bias_prior.{q}"] = eg["prior_bias_mean"][q:q + 1]
        want[f"scale_group_bias_prior.{q}"] = eg["prior_bias_scale"][q:q + 1]
        want[f"mean_group_entity_prior.{q}"] = eg["prior_entity_mean"][q]
        want[f"scale_group_entity_prior.{q}"] = eg["prior_entity_scale"][q]
    for k, w in want.items():
        got = gr[k].cpu().numpy()
        assert gu.rel_err(got, w) < 5e-6, (k, "vs fp64 maths", gu.rel_err(got, w))
        assert gu.rel_err(got, g[f"step0.grad.{k}"]) < 3e-5, (k, "vs reference fp32")
    touched = np.zeros(sum(fs), dtype=bool)
    touched[exact["plan"]["uniq"]] = True
    assert not gr["entity_params"].cpu().numpy()[~touched].any()


@pytest.mark.parametrize("name", gu.CLOSED)
def test_fused_step_updates_match_reference(name):
    meta, g = gu.load(name)
    lr, fs = meta["lr"], meta["group_sizes"]
    for t in range(meta["steps"]):
        m = _model(meta, g, t)
        x, y = gu.batch_of(meta, g, t)
        xd, yd = torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV)
        before = {k: v.detach().clone() for k, v in m.state_dict().items()}
        mv = (m.entity_m.clone(), m.entity_v.clone(), m.bias_m.clone(), m.bias_v.clone())
        gr = m.gradients(xd, yd)
        out = m.fused_step(xd, yd)
        assert int(m.adam_step.item()) == t + 1
        np.testing.assert_allclose(out["loss"].item(), g[f"step{t}.loss"], rtol=1e-5)
        uniq = np.unique(x)
        after = gu.state(g, f"step{t}.after")
        # (a) fused Adam == torch's recurrence applied to the kernel's own gradient
        for key, mm, vv in (("entity_params", mv[0], mv[1]), ("bias_params", mv[2], mv[3])):
            p1, _, _ = vfm_math.adam_update(before[key].cpu().numpy().astype(np.float64),
                                            gr[key].cpu().numpy().astype(np.float64),
                                            mm.cpu().numpy().astype(np.float64),
                                            vv.cpu().numpy().astype(np.float64), t + 1, lr, rows=uniq)
            np.testing.assert_allclose(m.state_dict()[key].cpu().numpy(), p1, rtol=2e-6, atol=2e-6 * max(1.0, lr))
        # (b) touched rows and every scalar / prior parameter against the reference's dense Adam.
        # Elements with |g| ~ Adam's eps are ill-conditioned in fp32 (the reference's own
        # scatter-add order moves them): allow a 1e-3 fraction of outliers on the tables.
        for key in ("entity_params", "bias_params"):
            got, want = m.state_dict()[key].cpu().numpy()[uniq], after[key][uniq]
            bad = np.abs(got - want) > 1e-5 * np.abs(want) + 2e-5 * lr + 1e-6
            assert bad.mean() <= 1e-3, (name, t, key, float(bad.mean()))
        for key in after:
            if key in ("entity_params", "bias_params"):
                continue
            np.testing.assert_allclose(m.state_dict()[key].cpu().numpy(), after[key], rtol=2e-5,
                                       atol=5e-5 * lr + 1e-6, err_msg=f"{name} step {t} {key}")
        mask = np.ones(sum(fs), dtype=bool)
        mask[uniq] = False
        assert torch.equal(m.state_dict()["entity_params"][mask], before["entity_params"][mask])


def test_closed_backward_is_bitwise_deterministic():
    meta, g = gu.load("closed_ml100k")
    x, y = gu.batch_of(meta, g, 0)
    outs = []
    for _ in range(3):
        m = _model(meta, g, 0)
        m.fused_step(torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV))
        outs.append((m.entity_params.detach().clone(), m.bias_params.detach().clone(), m._scalars.clone()))
    for o in outs[1:]:
        assert all(torch.equal(a, b) for a, b in zip(o, outs[0]))


def test_reference_forward_signature_values():
    """model(inverse_group, group_present, closed_form_loss=True, target=y) as in the script's loop."""
    meta, g = gu.load("closed_3groups")
    fs, G = meta["group_sizes"], len(meta["group_sizes"])
    m = _model(meta, g, 0)
    x, y = gu.batch_of(meta, g, 0)
    xt = torch.from_numpy(x)
    present, inverse = [], []
    for q in range(G):
        p, i = torch.unique(xt[:, q], return_inverse=True)
        present.append(p), inverse.append(i)
    likelihood, kls, partial = m(inverse, present, closed_form_loss=True, target=torch.from_numpy(y))
    port = vfm_port.ClosedPort(meta["d"], fs, alpha_0=meta["alpha_0"])
    gu.load_state(port, gu.state(g, "init"))
    lik, pkls, ppart, _ = port(xt, torch.from_numpy(y))
    np.testing.assert_allclose(likelihood.mean.detach().cpu().numpy(), lik.mean.detach().numpy(), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(partial.item(), ppart.item(), rtol=1e-5)
    np.testing.assert_allclose(kls[0].detach().cpu().numpy(), pkls[0].detach().numpy(), rtol=1e-5)
    np.testing.assert_allclose(kls[1].detach().cpu().numpy(), pkls[1].detach().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(kls[2].detach().cpu().numpy(), pkls[2].detach().numpy(), rtol=1e-5, atol=1e-6)
    assert kls[1].requires_grad and kls[2].requires_grad and partial.requires_grad   # training outputs carry gradients
    with torch.no_grad():                                                            # evaluation path: values only
        lik_ng, kls_ng, part_ng = m(inverse, present, closed_form_loss=True, target=torch.from_numpy(y))
    assert torch.equal(lik_ng.mean, likelihood.mean.detach()) and torch.equal(kls_ng[2], kls[2].detach())
    likelihood2, kls2 = m(inverse, present)
    assert torch.equal(likelihood2.mean, likelihood.mean.detach())
    np.testing.assert_allclose(m.predict(xt).cpu().numpy(), likelihood.mean.detach().cpu().numpy(), rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("name", gu.CLOSED)
def test_per_group_plan_is_bit_exact(name):
    """vfm-tomasrch.py:536-545: per column g, torch.unique(indices[:, g], return_inverse, return_counts).
    The CUDA plan holds the same integers: rows are grouped by id range, so group g's ``present`` is the
    slice [class_off[g], class_off[g+1]) of the sorted unique list, ``inverse`` the unique rank minus
    that offset, ``counts`` the segment lengths."""
    meta, g = gu.load(name)
    fs = meta["group_sizes"]
    G = len(fs)
    m = _model(meta, g, 0)
    for t in range(meta["steps"]):
        x, y = gu.batch_of(meta, g, t)
        m.fused_step(torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV), update=False)
        plan = m._plan
        plan.check_ids()
        uniq, inverse, counts = (a.cpu().numpy() for a in plan.as_unique())
        off = plan.class_off[: G + 1].cpu().numpy()
        assert off[0] == 0 and off[G] == len(uniq)
        for q in range(G):
            lo, hi = off[q], off[q + 1]
            assert np.array_equal(uniq[lo:hi], g[f"step{t}.present{q}"]), (t, q)
            assert np.array_equal(inverse[:, q] - lo, g[f"step{t}.inverse{q}"]), (t, q)
            assert np.array_equal(counts[lo:hi], g[f"step{t}.counts{q}"]), (t, q)


@pytest.mark.parametrize("name", gu.CLOSED)
def test_dropin_autograd_loop_equals_reference(name):
    """The reference's own training loop (vfm-tomasrch.py:535-594) on the drop-in module: per-group
    torch.unique, model(inverse_group, group_present, closed_form_loss=True, target=y), the script's loss
    expression, loss.backward(), torch.optim.Adam (dense) -- against the goldens of the unmodified script."""
    meta, g = gu.load(name)
    fs, lr = meta["group_sizes"], meta["lr"]
    n_groups = len(fs)
    m = _model(meta, g, 0)
    entity_count = torch.from_numpy(g["train_counts"]).float().to(DEV)
    group_sizes = fs
    n_train = meta["n_train"]
    optimizer = torch.optim.Adam(m.parameters(), lr=lr)
    for t in range(meta["steps"]):
        if t > 0:                                          # replay from the exact reference state
            _restore(m, g, t)
            for k, p in m.named_parameters():
                optimizer.state[p] = {"step": torch.tensor(float(t)),
                                      "exp_avg": torch.from_numpy(g[f"step{t - 1}.adam.{k}.m"]).to(DEV).reshape(p.shape),
                                      "exp_avg_sq": torch.from_numpy(g[f"step{t - 1}.adam.{k}.v"]).to(DEV).reshape(p.shape)}
        x, y = gu.batch_of(meta, g, t)
        indices, target = torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV)
        group_present, inverse_group, batch_group_count = [], [], []
        for i_group in range(n_groups):
            present, inverse, batch_count = torch.unique(indices[:, i_group], return_inverse=True, return_counts=True)
            group_present.append(present), inverse_group.append(inverse), batch_group_count.append(batch_count)
        outputs, kls, partial_loss = m(inverse_group, group_present, closed_form_loss=True, target=target)
        loss = (- n_train * partial_loss / len(indices)
                + kls[0]
                + ((kls[1] + kls[2].sum(axis=1))
                   * torch.cat([(group_sizes[i_group]
                                 / (batch_group_count[i_group] / entity_count[group_present[i_group]]).sum()
                                 ).repeat(len(group_present[i_group])) for i_group in range(n_groups)])
                   * torch.concat(batch_group_count)
                   / entity_count[torch.concat(group_present)]).sum())
        np.testing.assert_allclose(loss.item(), g[f"step{t}.loss"], rtol=1e-5)
        np.testing.assert_allclose(outputs.mean.detach().cpu().numpy(), g[f"step{t}.pred"], rtol=1e-5, atol=2e-6)
        optimizer.zero_grad()
        loss.backward()
        if t == 0:
            for k, p in m.named_parameters():
                assert p.grad is not None, k
                assert gu.rel_err(p.grad.cpu().numpy().reshape(-1), g[f"step0.grad.{k}"].reshape(-1)) < 3e-5, k
        optimizer.step()
        after = gu.state(g, f"step{t}.after")
        for k in after:                                   # dense Adam: every row, every prior parameter
            got, want = m.state_dict()[k].cpu().numpy().reshape(-1), after[k].reshape(-1)
            bad = np.abs(got - want) > 1e-5 * np.abs(want) + 5e-5 * lr + 1e-6
            assert bad.mean() <= 1e-3, (name, t, k, float(bad.mean()))


def test_dropin_autograd_rejects_non_uniform_kl_weights():
    meta, g = gu.load("closed_3groups")
    m = _model(meta, g, 0)
    x, y = gu.batch_of(meta, g, 0)
    xt = torch.from_numpy(x).to(DEV)
    present, inverse = zip(*[torch.unique(xt[:, q], return_inverse=True) for q in range(len(meta["group_sizes"]))])
    _, kls, partial = m(list(inverse), list(present), closed_form_loss=True, target=torch.from_numpy(y).to(DEV))
    wk = torch.linspace(0.5, 1.5, meta["d"], device=DEV)
    (-partial + (kls[2] * wk).sum() + kls[1].sum()).backward()
    assert torch.isnan(m.entity_params.grad).all()


def test_closed_form_wide_rows():
    """d = 128 (3d = 384 floats per gathered row: beyond the 256-float layout limit of round 1)."""
    from vae_b200.vfm_tomasrch import CF
    fs, d, B = [50, 30], 128, 600
    rng = np.random.default_rng(3)
    x = np.stack([rng.integers(0, fs[0], B), fs[0] + rng.integers(0, fs[1], B)], 1).astype(np.int64)
    y = np.clip(np.round(3.5 + rng.standard_normal(B)), 1, 5).astype(np.float32)
    tc = np.bincount(x.reshape(-1), minlength=sum(fs))
    tc[tc == 0] = 1
    torch.manual_seed(1)
    m = CF(embedding_size=d, n_groups=2, group_sizes=fs, alpha_0=2.0, train_counts=torch.from_numpy(tc), n_train=B,
           max_batch=B, lr=0.1)
    with torch.no_grad():
        m.entity_params[:, :d].normal_(0, 0.1)
    sd = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    gr = m.gradients(torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV))
    exact = vfm_math.closed_step(gu.closed_math_params(sd, 2), x, y, tc.astype(np.float64), B, fs)
    np.testing.assert_allclose(gr["loss"].item(), exact["loss"], rtol=1e-5)
    assert gu.rel_err(gr["entity_params"].cpu().numpy(), exact["grads"]["entity"]) < 5e-6
    assert gu.rel_err(gr["bias_params"].cpu().numpy(), exact["grads"]["bias"]) < 5e-6
