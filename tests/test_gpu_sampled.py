"""Parity of the CUDA sampled-ELBO path (through the C ABI) with the oracle:
golden vectors of the unmodified reference, the torch port and the fp64 maths.

Tolerances (north_star): integer plan bit-exact; fp32 predictions / ELBO /
updated parameters within 1e-5 relative with the same injected noise.  The
reference's own fp32 gradient carries ~1e-5 of summation-order noise
(sequential scatter-add over hot rows), so gradients are compared in max-norm."""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import vfm_math, vfm_port

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(meta, g, t=0, **kw):
    from vae_b200.vfm_torch import CF
    m = CF(meta["d"], output=meta["output"], n_users=meta["N"], n_items=meta["M"],
           train_counts=torch.from_numpy(g["train_counts"]), n_var_samples=meta["S"],
           link=meta["link"], n_train=meta["n_train"], max_batch=meta["batch"], lr=meta["lr"], **kw)
    _restore(m, g, t)
    return m


def _restore(m, g, t):
    sd = gu.state(g, "init" if t == 0 else f"step{t - 1}.after")
    own = m.state_dict()
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)).reshape(own[k].shape) for k, v in sd.items()
                       if k in own}, strict=False)
    for buf in (m.bias_m, m.bias_v, m.entity_m, m.entity_v, m._scalars_m, m._scalars_v):
        buf.zero_()
    m.adam_step.fill_(t)
    if t > 0:
        a = lambda k: torch.from_numpy(g[f"step{t - 1}.adam.{k}"]).to(DEV)
        m.bias_m.copy_(a("bias_params.weight.m")), m.bias_v.copy_(a("bias_params.weight.v"))
        m.entity_m.copy_(a("entity_params.weight.m")), m.entity_v.copy_(a("entity_params.weight.v"))
        from vae_b200 import _lib as L
        for name, idx in (("alpha", L.S_ALPHA), ("global_bias_mean", L.S_GB_MEAN),
                          ("global_bias_scale", L.S_GB_SCALE)):
            if f"step{t - 1}.adam.{name}.m" in g:
                m._scalars_m[idx] = float(g[f"step{t - 1}.adam.{name}.m"][0])
                m._scalars_v[idx] = float(g[f"step{t - 1}.adam.{name}.v"][0])


def _noise(g, t):
    return [torch.from_numpy(g[f"step{t}.noise{i}"]).to(DEV) for i in range(3)]


S1 = [n for n in gu.SAMPLED if n != "sampled_class_s2"]


@pytest.mark.parametrize("name", S1)
def test_plan_is_bit_exact(name):
    meta, g = gu.load(name)
    m = _model(meta, g)
    for t in range(meta["steps"]):
        x, _ = gu.batch_of(meta, g, t)
        plan = m.plan(torch.from_numpy(x).to(DEV))
        plan.check_ids()
        uniq, inverse, counts = plan.as_unique()
        assert np.array_equal(uniq.cpu().numpy(), g[f"step{t}.uniq"])
        assert np.array_equal(inverse.cpu().numpy(), g[f"step{t}.inverse"])
        assert np.array_equal(counts.cpu().numpy(), g[f"step{t}.counts"])
        ref = vfm_math.build_plan(x)
        U = len(ref["uniq"])
        assert np.array_equal(plan.seg_off[:U + 1].cpu().numpy(), ref["seg_off"])
        assert np.array_equal(plan.occ[:x.size].cpu().numpy(), ref["order"])     # touched-row segments


@pytest.mark.parametrize("B,F,R,hot", [(1, 2, 10, 0), (7, 2, 5, 0), (1000, 2, 50, 1), (4096, 3, 100000, 0),
                                       (5000, 8, 3000, 1), (65536, 2, 165237, 1),
                                       (1, 1, 1, 0), (300, 2, 2, 1),                 # one bin / two rows
                                       (1025, 2, 1024, 0), (2048, 2, 1025, 0),       # digit-width edges
                                       (65536, 2, 5_000_000, 1),                     # 23 bits: 3 passes
                                       (50000, 2, 100_000_000, 1),                   # 27 bits: 3 passes
                                       (65536, 8, 1_000_000, 1),                     # 512 tiles
                                       (100000, 8, 1_000_000, 1)])                   # tile of 2048 keys
def test_plan_random_shapes_vs_torch_unique(B, F, R, hot):
    from vae_b200 import _lib as L
    from vae_b200.engine import BatchPlan, make_config
    gen = torch.Generator().manual_seed(B + F)
    x = torch.randint(0, R, (B, F), generator=gen)
    if hot:
        x[torch.rand(B, generator=gen) < 0.3, F - 1] = R - 1          # a heavy row, and the max id
    tc = torch.ones(R, device=DEV)
    plan = BatchPlan(B, F, R, DEV)
    cfg = make_config(B, F, 8, R, 1, "reg", "abs", [R // 2 + 1], [R // 2, R - R // 2], 100, 7)
    plan.build(cfg, x.to(DEV), tc)
    plan.check_ids()
    u, i, c = torch.unique(x, return_inverse=True, return_counts=True)
    pu, pi, pc = plan.as_unique()
    assert torch.equal(pu.cpu(), u) and torch.equal(pi.cpu(), i) and torch.equal(pc.cpu(), c)
    U = len(u)
    order = torch.argsort(i.reshape(-1), stable=True)
    assert torch.equal(plan.occ[: B * F].cpu().long(), order)
    # flat records: pos_of inverts occ; partner = rank of the sample's other field (F=2) or n
    occ = plan.occ[: B * F].cpu().long()
    assert torch.equal(plan.pos_of[: B * F].cpu().long()[occ], torch.arange(B * F))
    want_partner = i.reshape(-1)[occ ^ 1] if F == 2 else occ // F
    assert torch.equal(plan.partner[: B * F].cpu().long(), want_partner)
    urec = plan.urec[: 4 * U].cpu().long().reshape(U, 4)
    seg = torch.cat((torch.zeros(1, dtype=torch.long), torch.cumsum(c, 0)))
    assert torch.equal(urec[:, 0], u) and torch.equal(urec[:, 1], c) and torch.equal(urec[:, 2], seg[:-1])
    assert torch.equal(urec[:, 3], c)
    assert torch.equal(plan.pos_rank[: B * F].cpu().long(), i.reshape(-1)[occ])
    # per-column normaliser Z_f = sum_n 1/cnt_train = B for unit counts
    np.testing.assert_allclose(plan.z[:F].cpu().numpy(), np.full(F, float(B)), rtol=1e-6)


def test_plan_flags_out_of_range_ids():
    from vae_b200.engine import BatchPlan, make_config
    x = torch.tensor([[0, 3], [1, 9]])
    plan = BatchPlan(2, 2, 5, DEV)
    cfg = make_config(2, 2, 4, 5, 1, "reg", "abs", [3], [2, 3], 2, 7)
    plan.build(cfg, x.to(DEV), torch.ones(5, device=DEV))
    with pytest.raises(IndexError):
        plan.check_ids()


@pytest.mark.parametrize("name", gu.SAMPLED)       # incl. S = 2 variational samples
def test_forward_matches_reference_golden(name):
    meta, g = gu.load(name)
    for t in range(meta["steps"]):
        m = _model(meta, g, t)
        x, y = gu.batch_of(meta, g, t)
        out = m.fused_step(torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV), noise=_noise(g, t),
                           update=False)
        np.testing.assert_allclose(out["pred"].cpu().numpy(), g[f"step{t}.pred"], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(out["loss"].item(), g[f"step{t}.loss"][0], rtol=1e-5)
        np.testing.assert_allclose(out["kl"].item(), g[f"step{t}.kl"][0], rtol=1e-5)
        np.testing.assert_allclose(out["nll_mean"].item(), g[f"step{t}.nll_mean"], rtol=1e-5)


@pytest.mark.parametrize("name", gu.SAMPLED)       # incl. S = 2 variational samples
def test_gradients_match_reference_and_fp64_maths(name):
    meta, g = gu.load(name)
    m = _model(meta, g, 0)
    x, y = gu.batch_of(meta, g, 0)
    noise = _noise(g, 0)
    gr = m.gradients(torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV), noise=noise)
    exact = vfm_math.sampled_step(gu.sampled_math_params(gu.state(g, "init")), x, y,
                                  [n.cpu().numpy() for n in noise], g["train_counts"], meta["n_train"],
                                  [meta["N"], meta["M"]], output=meta["output"], link=meta["link"])
    pairs = [("entity_params.weight", "entity"), ("bias_params.weight", "bias"),
             ("global_bias_mean",) * 2, ("global_bias_scale",) * 2]
    if meta["output"] == "reg":
        pairs.append(("alpha",) * 2)
    for k, kk in pairs:
        got = gr[k].cpu().numpy()
        assert gu.rel_err(got, exact["grads"][kk]) < 3e-6, (k, "vs fp64 maths")
        assert gu.rel_err(got, g[f"step0.grad.{k}"]) < 2e-5, (k, "vs reference fp32")
    # untouched rows: exactly zero gradient, touched-row set bit-exact
    touched = np.zeros(meta["N"] + meta["M"], dtype=bool)
    touched[g["step0.uniq"]] = True
    ge = gr["entity_params.weight"].cpu().numpy()
    assert not ge[~touched].any()
    assert (np.abs(ge[touched]).sum(axis=1) > 0).all()


@pytest.mark.parametrize("name", gu.SAMPLED)       # incl. S = 2 variational samples
def test_fused_step_updates_match_reference(name):
    """One fused step from the exact pre-step state of every golden step."""
    meta, g = gu.load(name)
    lr = meta["lr"]
    for t in range(meta["steps"]):
        m = _model(meta, g, t)
        x, y = gu.batch_of(meta, g, t)
        xd, yd, noise = torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV), _noise(g, t)
        before = {k: v.detach().clone() for k, v in m.state_dict().items()}
        mv = (m.entity_m.clone(), m.entity_v.clone(), m.bias_m.clone(), m.bias_v.clone())
        gr = m.gradients(xd, yd, noise=noise)
        out = m.fused_step(xd, yd, noise=noise)
        assert int(m.adam_step.item()) == t + 1
        uniq = g[f"step{t}.uniq"]
        after = gu.state(g, f"step{t}.after")
        # (a) the fused Adam arithmetic == torch's recurrence applied to the kernel's own gradient
        for key, mm, vv in (("entity_params.weight", mv[0], mv[1]), ("bias_params.weight", mv[2], mv[3])):
            p1, m1, v1 = vfm_math.adam_update(before[key].cpu().numpy().astype(np.float64),
                                              gr[key].cpu().numpy().astype(np.float64),
                                              mm.cpu().numpy().astype(np.float64),
                                              vv.cpu().numpy().astype(np.float64), t + 1, lr, rows=uniq)
            got = m.state_dict()[key].cpu().numpy()
            np.testing.assert_allclose(got, p1, rtol=2e-6, atol=2e-6 * max(1.0, lr))
        # (b) touched rows against the reference's dense-Adam result (identical on touched rows);
        # elements whose update is ill-conditioned in fp32 (|g| ~ Adam eps, or m ~ 0) are the
        # reference's own summation-order noise: allow a 1e-4 fraction of outliers
        for key in ("entity_params.weight", "bias_params.weight"):
            got, want = m.state_dict()[key].cpu().numpy()[uniq], after[key][uniq]
            bad = np.abs(got - want) > 1e-5 * np.abs(want) + 2e-5 * lr + 1e-6
            assert bad.mean() <= 1e-4, (name, t, key, float(bad.mean()))
        for key in ("global_bias_mean", "global_bias_scale") + (("alpha",) if meta["output"] == "reg" else ()):
            np.testing.assert_allclose(m.state_dict()[key].cpu().numpy(), after[key], rtol=1e-5,
                                       atol=2e-5 * lr + 1e-6)
        if meta["output"] != "reg":                       # Bernoulli: alpha must not move
            assert torch.equal(m.state_dict()["alpha"], before["alpha"])
        # untouched rows are untouched by the lazy update
        mask = np.ones(meta["N"] + meta["M"], dtype=bool)
        mask[uniq] = False
        assert torch.equal(m.state_dict()["entity_params.weight"][mask], before["entity_params.weight"][mask])
        np.testing.assert_allclose(out["loss"].item(), g[f"step{t}.loss"][0], rtol=1e-5)
        np.testing.assert_allclose(out["kl"].item(), g[f"step{t}.kl"][0], rtol=1e-5)


def test_backward_is_bitwise_deterministic():
    meta, g = gu.load("sampled_fraction")               # 20 item rows with ~430 occurrences each
    x, y = gu.batch_of(meta, g, 0)
    outs = []
    for _ in range(3):
        m = _model(meta, g, 0)
        m.fused_step(torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV), noise=_noise(g, 0))
        outs.append((m.entity_params.weight.clone(), m.bias_params.weight.clone(), m.entity_v.clone()))
    for o in outs[1:]:
        assert all(torch.equal(a, b) for a, b in zip(o, outs[0]))


@pytest.mark.parametrize("name", ["sampled_reg_d5", "sampled_class_s2"])
def test_dropin_autograd_path_equals_reference_loop(name):
    """The reference's own loop (vfm-torch.py:353-370) run against the drop-in module:
    model(x) -> likelihood/kl -> loss.backward() -> torch.optim.Adam(dense).step()
    (also with S = 2 variational samples: likelihood batch shape [S, B])."""
    meta, g = gu.load(name)
    m = _model(meta, g, 0)
    opt = torch.optim.Adam(m.parameters(), lr=meta["lr"])
    for t in range(meta["steps"]):
        if t > 0:                                         # replay from the exact reference state
            _restore(m, g, t)
            for k, p in m.named_parameters():
                if f"step{t - 1}.adam.{k}.m" in g:
                    opt.state[p] = {"step": torch.tensor(float(t)),
                                    "exp_avg": torch.from_numpy(g[f"step{t - 1}.adam.{k}.m"]).to(DEV).reshape(p.shape),
                                    "exp_avg_sq": torch.from_numpy(g[f"step{t - 1}.adam.{k}.v"]).to(DEV).reshape(p.shape)}
        x, y = gu.batch_of(meta, g, t)
        likelihood, last_logits, mean_logits, kl_term = m(torch.from_numpy(x).to(DEV), noise=_noise(g, t))
        assert last_logits is None and mean_logits is None
        loss = -likelihood.log_prob(torch.from_numpy(y).to(DEV).float()).mean() * meta["n_train"] + kl_term
        np.testing.assert_allclose(loss.item(), g[f"step{t}.loss"][0], rtol=1e-5)
        np.testing.assert_allclose(likelihood.mean.squeeze().detach().cpu().numpy(), g[f"step{t}.pred"],
                                   rtol=1e-5, atol=2e-6)
        opt.zero_grad()
        loss.backward()
        if t == 0:
            scal = ("alpha",) if meta["output"] == "reg" else ()    # Bernoulli: alpha has no gradient (N10)
            for k in ("entity_params.weight", "bias_params.weight", "global_bias_mean", "global_bias_scale") + scal:
                got = dict(m.named_parameters())[k].grad.cpu().numpy()
                assert gu.rel_err(got, g[f"step0.grad.{k}"]) < 2e-5, k
            assert m.prec_user_bias_prior.grad is None
            if not scal:
                assert m.alpha.grad is None
        opt.step()
        after = gu.state(g, f"step{t}.after")
        for k in ("entity_params.weight", "bias_params.weight", "alpha", "global_bias_mean", "global_bias_scale"):
            if k == "alpha" and meta["output"] != "reg":
                continue
            got, want = m.state_dict()[k].cpu().numpy(), after[k]       # dense Adam: every row
            bad = np.abs(got - want) > 1e-5 * np.abs(want) + 2e-5 * meta["lr"] + 1e-6
            assert bad.mean() <= 1e-4, (t, k, float(bad.mean()))


@pytest.mark.parametrize("name", ["sampled_reg_d64", "sampled_class_s2"])
def test_philox_mode_matches_oracle_with_exported_noise(name):
    meta, g = gu.load(name)
    m = _model(meta, g, 0, seed=1234)
    x, y = gu.batch_of(meta, g, 0)
    xd, yd = torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV)
    uniq = torch.from_numpy(g["step0.uniq"])
    noise = m.philox_noise(uniq)
    e = noise[2].reshape(-1).cpu().numpy()
    tol = 0.02 if e.size > 50_000 else 0.1                              # the S = 2 golden has 1 600 draws
    assert abs(e.mean()) < tol and abs(e.std() - 1.0) < tol and np.abs(e).max() < 6.5
    assert torch.equal(m.philox_noise(uniq)[2], noise[2])                 # counter-based: reproducible
    assert not torch.equal(m.philox_noise(uniq, step=1)[2], noise[2])     # new step, new draws
    out = m.fused_step(xd, yd)                                            # Philox inside the kernels
    pred_philox, loss_philox = out["pred"].clone(), out["loss"].item()
    w_after = m.entity_params.weight.clone()
    m2 = _model(meta, g, 0, seed=1234)
    out2 = m2.fused_step(xd, yd, noise=noise)                             # same draws, injected
    assert torch.allclose(pred_philox, out2["pred"], rtol=1e-6, atol=1e-6)
    assert abs(loss_philox - out2["loss"].item()) <= 1e-6 * abs(loss_philox)
    assert torch.allclose(w_after, m2.entity_params.weight, rtol=1e-6, atol=1e-6)
    exact = vfm_math.sampled_step(gu.sampled_math_params(gu.state(g, "init")), x, y,
                                  [n.cpu().numpy() for n in noise], g["train_counts"], meta["n_train"],
                                  [meta["N"], meta["M"]], output=meta["output"], link=meta["link"])
    np.testing.assert_allclose(pred_philox.cpu().numpy(), exact["mean"].squeeze(), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(loss_philox, exact["loss"], rtol=1e-5)


def test_dropin_forward_draws_fresh_noise_every_call():
    """The advertised drop-in loop (model(x) -> loss.backward() -> torch.optim.Adam.step(), default
    noise) never advances the Adam step counter of the fused path; the Philox stream is therefore
    indexed by its own device counter that every sampled forward advances (vfm-torch.py:238-241 draws
    fresh noise in every forward).  Each call must use philox_noise(step = call index), consecutive
    calls must differ, and the backward must re-create the draws of ITS forward."""
    meta, g = gu.load("sampled_reg_d64")
    m = _model(meta, g, 0, seed=99)
    x, y = gu.batch_of(meta, g, 0)
    xd, yd = torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV)
    uniq = torch.from_numpy(g["step0.uniq"])
    opt = torch.optim.Adam(m.parameters(), lr=meta["lr"])
    preds = []
    for call in range(3):
        assert m.noise_step.tolist()[0] == call
        noise = m.philox_noise(uniq)                                    # what the next forward will draw
        twin = _model(meta, g, 0, seed=99)
        twin.load_state_dict(m.state_dict())
        want = twin.gradients(xd, yd, noise=noise)                      # same draws, injected
        likelihood, _, _, kl_term = m(xd)
        assert m.noise_step.tolist() == [call + 1, call]
        loss = -likelihood.log_prob(yd).mean() * meta["n_train"] + kl_term
        pred = likelihood.mean.detach().clone()
        assert torch.allclose(pred, want["pred"], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(loss.item(), want["loss"].item(), rtol=1e-6)
        opt.zero_grad()
        loss.backward()
        for k in ("entity_params.weight", "bias_params.weight", "global_bias_scale"):
            got = dict(m.named_parameters())[k].grad
            assert gu.rel_err(got.cpu().numpy(), want[k].cpu().numpy()) < 2e-6, (call, k)
        opt.step()
        preds.append(pred)
    assert not torch.allclose(preds[0], preds[1], atol=1e-3)
    # evaluation forwards consume noise too (vfm-torch.py:402-403 samples at test time)
    a = m.fused_step(xd, yd, update=False)["pred"].clone()
    b = m.fused_step(xd, yd, update=False)["pred"].clone()
    assert not torch.allclose(a, b, atol=1e-3) and m.noise_step.tolist() == [5, 4]


def test_out_of_range_id_poisons_the_loss():
    """The reference raises IndexError from nn.Embedding; here the plan flags the id (meta[2]), maps it
    to row 0 and the forward turns the loss into NaN -- a signal that needs no host synchronisation."""
    from vae_b200.vfm_torch import CF
    torch.manual_seed(0)
    m = CF(8, output="reg", n_users=5, n_items=4, train_counts=torch.ones(9), n_train=6, max_batch=6, device=DEV)
    x = torch.tensor([[0, 5], [1, 6], [2, 7], [3, 8], [4, 5], [1, 8]])
    y = torch.ones(6)
    ok = m.fused_step(x.to(DEV), y.to(DEV))
    assert np.isfinite(ok["loss"].item())
    x[2, 1] = 9                                                         # one past the last row
    out = m.fused_step(x.to(DEV), y.to(DEV))
    assert np.isnan(out["loss"].item())
    with pytest.raises(IndexError):
        m._plan.check_ids()
    out = m.fused_step(x.to(DEV), y.to(DEV), noise=None, update=False)
    assert np.isnan(out["loss"].item())


@pytest.mark.parametrize("F,d", [(3, 8), (8, 64), (4, 20)])
def test_multi_field_pairwise_matches_fp64_maths(F, d):
    """Config-4 shape: F>2 fields, pairwise interaction, per-group KL weights.  The reference has
    no sampled implementation for F>2; the oracle is the fp64 maths (+ torch port)."""
    from vae_b200.vfm_torch import CF
    fs = [37, 23, 11, 9, 7, 5, 4, 3][:F]
    R, B = sum(fs), 700
    rng = np.random.default_rng(F)
    offs = np.concatenate(([0], np.cumsum(fs)[:-1]))
    x = np.stack([offs[f] + rng.integers(0, fs[f], B) for f in range(F)], 1).astype(np.int64)
    y = (rng.random(B) < 0.5).astype(np.float32)
    tc = np.bincount(x.reshape(-1), minlength=R)
    tc[tc == 0] = 1
    torch.manual_seed(3)
    m = CF(d, output="class", n_users=fs[0], n_items=fs[1], train_counts=torch.from_numpy(tc),
           field_sizes=fs, kl_weighting="group", n_train=B, max_batch=B, lr=0.05)
    with torch.no_grad():
        m.entity_params.weight.mul_(0.3)                 # keep 8-field sums in a sane range
    sd = {k: v.cpu().numpy() for k, v in m.state_dict().items()}
    U = len(np.unique(x))
    gen = torch.Generator().manual_seed(5)
    noise = [torch.randn(1, 1, generator=gen), torch.randn(1, U, generator=gen), torch.randn(1, U, d, generator=gen)]
    exact = vfm_math.sampled_step(gu.sampled_math_params(sd), x, y, [n.numpy() for n in noise], tc, B, fs,
                                  output="class", interaction="pairwise", kl_weighting="group")
    port = vfm_port.SampledPort(fs[0], fs[1], d, torch.from_numpy(tc), output="class", field_sizes=fs,
                                interaction="pairwise", kl_weighting="group")
    gu.load_state(port, sd)
    po = vfm_port.sampled_port_step(port, torch.optim.Adam(port.parameters(), lr=0.05), torch.from_numpy(x),
                                    torch.from_numpy(y), B, noise)
    np.testing.assert_allclose(po["loss"].item(), exact["loss"], rtol=1e-5)
    xd, yd = torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV)
    gr = m.gradients(xd, yd, noise=[n.to(DEV) for n in noise])
    np.testing.assert_allclose(gr["loss"].item(), exact["loss"], rtol=1e-5)
    np.testing.assert_allclose(gr["pred"].cpu().numpy(), exact["mean"].squeeze(), rtol=1e-5, atol=2e-6)
    assert gu.rel_err(gr["entity_params.weight"].cpu().numpy(), exact["grads"]["entity"]) < 1e-5
    assert gu.rel_err(gr["bias_params.weight"].cpu().numpy(), exact["grads"]["bias"]) < 1e-5
    m.fused_step(xd, yd, noise=[n.to(DEV) for n in noise])
    uniq = exact["plan"]["uniq"]
    got, want = m.entity_params.weight.detach().cpu().numpy()[uniq], port.entity_params.weight.detach().numpy()[uniq]
    bad = np.abs(got - want) > 1e-5 * np.abs(want) + 2e-5 * 0.05 + 1e-6
    assert bad.mean() <= 1e-4


def test_mean_prediction_and_saved_weights():
    meta, g = gu.load("sampled_reg_d64")
    m = _model(meta, g, 0)
    x, _ = gu.batch_of(meta, g, 0)
    xd = torch.from_numpy(x).to(DEV)
    d = meta["d"]
    W, Bw = m.entity_params.weight.detach().cpu().numpy(), m.bias_params.weight.detach().cpu().numpy()
    want = (m.global_bias_mean.item() + Bw[x, 0].sum(1) + W[x][:, :, :d].prod(1).sum(1))
    np.testing.assert_allclose(m.predict_mean(xd).cpu().numpy(), want, rtol=1e-5, atol=1e-5)
    m.save_weights()
    _, last_logits, mean_logits, _ = m(xd)
    np.testing.assert_allclose(last_logits.cpu().numpy(), want, rtol=1e-5, atol=1e-5)   # vfm-torch.py:248-259
    np.testing.assert_allclose(mean_logits.cpu().numpy(), want, rtol=1e-5, atol=1e-5)


def test_last_and_mean_logits_match_reference_golden():
    """vfm-torch.py:179-185, 248-262 on the unmodified reference: save_weights() after each of three
    training steps, then last_logits (last snapshot) and mean_logits (mean of the three snapshots)."""
    meta, g = gu.load("sampled_saved_logits")
    m = _model(meta, g, 0)
    for t in range(meta["steps"]):
        sd = gu.state(g, f"step{t}.after")
        own = m.state_dict()
        m.load_state_dict({k: torch.from_numpy(np.asarray(v)).reshape(own[k].shape) for k, v in sd.items() if k in own},
                          strict=False)
        m.save_weights()
    xe = torch.from_numpy(g["x_eval"].astype(np.int64)).to(DEV)
    with torch.no_grad():
        _, last_logits, mean_logits, _ = m(xe)
    np.testing.assert_allclose(last_logits.cpu().numpy(), g["last_logits"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(mean_logits.cpu().numpy(), g["mean_logits"], rtol=1e-5, atol=2e-6)
    assert not np.allclose(g["last_logits"], g["mean_logits"], atol=1e-3)      # the running mean matters


def test_mean_prediction_of_multi_field_model_is_pairwise():
    """F > 2: the sampled step optimises the pairwise FM, so predict_mean / last_logits evaluate the
    pairwise interaction of the posterior means (interaction="prod" gives the scripts' formula)."""
    from vae_b200.vfm_torch import CF
    fs, d, B = [11, 7, 5], 8, 64
    rng = np.random.default_rng(0)
    offs = np.concatenate(([0], np.cumsum(fs)[:-1]))
    x = np.stack([offs[f] + rng.integers(0, fs[f], B) for f in range(3)], 1).astype(np.int64)
    for inter in ("pairwise", "prod"):
        torch.manual_seed(1)
        m = CF(d, output="reg", n_users=fs[0], n_items=fs[1], train_counts=torch.ones(sum(fs)), field_sizes=fs,
               kl_weighting="group", interaction=inter if inter == "prod" else None, n_train=B, max_batch=B)
        W = m.entity_params.weight.detach().cpu().numpy()[:, :d].astype(np.float64)
        bw = m.bias_params.weight.detach().cpu().numpy()[:, 0].astype(np.float64)
        v = W[x]                                            # [B, F, d]
        fm = (0.5 * (v.sum(1) ** 2 - (v ** 2).sum(1))).sum(1) if inter == "pairwise" else v.prod(1).sum(1)
        want = m.global_bias_mean.item() + bw[x].sum(1) + fm
        np.testing.assert_allclose(m.predict_mean(torch.from_numpy(x).to(DEV)).cpu().numpy(), want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", ["sampled_reg_d64", "sampled_fraction", "sampled_reg_softplus"])
def test_pipelined_and_register_staged_row_updates_agree_bitwise(name):
    """k_adam_rows_pipe (cp.async staging, default) and k_adam_rows do the same arithmetic per element:
    identical parameters and moments; the KL sum is accumulated in a different (fixed) order."""
    from vae_b200 import _lib as L
    meta, g = gu.load(name)
    x, y = gu.batch_of(meta, g, 0)
    xd, yd = torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV)
    res = []
    for pipe in (1, 0):
        L.check(L.lib().vfmb_set_tuning(b"adam_pipe", pipe))
        try:
            m = _model(meta, g, 0, seed=3)
            losses = [m.fused_step(xd, yd)["loss"].item() for _ in range(3)]            # Philox path (flavour 2)
            losses.append(m.fused_step(xd, yd, noise=_noise(g, 0))["loss"].item())       # injected (flavour 1)
            res.append((losses, m.entity_params.weight.clone(), m.entity_m.clone(), m.entity_v.clone(),
                        m.bias_params.weight.clone(), m._scalars.clone()))
        finally:
            L.lib().vfmb_set_tuning(b"adam_pipe", 1)
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=1e-6)
    for a, b in zip(res[0][1:], res[1][1:]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("d", [32, 64, 20])
def test_gather_launch_knobs_do_not_change_a_bit(d):
    """Tile hand-out (static stride / counter), the lane mapping of k_stage, the L2 hints and the fence
    flavour of the cut-row finisher are scheduling only: parameters, moments and losses are identical.
    The lane mapping of the gather also sets its group-tile size, i.e. where a long row's sum is cut:
    same result up to the association of that sum (one setting per process).
    The batch has Zipf head rows of thousands of occurrences (cut by group and block tiles, global finisher)
    and thousands of short rows."""
    from vae_b200 import _lib as L
    from vae_b200.vfm_torch import CF
    fs, B = [3000, 400], 16384
    rng = np.random.default_rng(11)
    p1 = 1.0 / np.arange(1, fs[1] + 1); p1 /= p1.sum()
    x = np.stack([rng.integers(0, fs[0], B), fs[0] + rng.choice(fs[1], B, p=p1)], 1).astype(np.int64)
    y = rng.normal(size=B).astype(np.float32)
    xd, yd = torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV)
    counts = torch.from_numpy(np.bincount(x.reshape(-1), minlength=sum(fs)).astype(np.float32))
    knobs = ("gather_dyn", "gather_fence", "gather_wide", "stage_wide", "l2_keep", "stage_chunk", "score_chunk")
    defaults = (1, 1, 1, 0, 1, 16, 32)
    res = []
    for setting in (defaults, (0, 0, 1, 1, 0, 32, 16), (1, 0, 1, 0, 1, 32, 32), (1, 1, 0, 0, 1, 16, 32)):
        for k, v in zip(knobs, setting):
            L.check(L.lib().vfmb_set_tuning(k.encode(), v))
        try:
            torch.manual_seed(2)
            m = CF(d, output="reg", n_users=fs[0], n_items=fs[1], train_counts=counts, n_train=B, max_batch=B, lr=0.01)
            losses = [m.fused_step(xd, yd)["loss"].item() for _ in range(4)]
            res.append((losses, m.entity_params.weight.clone(), m.entity_m.clone(), m.bias_params.weight.clone()))
        finally:
            for k, v in zip(knobs, defaults):
                L.lib().vfmb_set_tuning(k.encode(), v)
    for r in res[1:3]:
        assert r[0] == res[0][0]
        for a, b in zip(res[0][1:], r[1:]):
            assert torch.equal(a, b)
    np.testing.assert_allclose(res[3][0], res[0][0], rtol=1e-6)         # gather_wide = 0: another association
    for a, b in zip(res[0][1:], res[3][1:]):
        assert (a - b).norm().item() <= 1e-4 * a.norm().item()            # (four steps amplify the last bits)


@pytest.mark.parametrize("output,F", [("class", 2), ("reg", 2), ("class", 3)])
def test_predict_proba_many_samples_matches_oracle(output, F):
    """vfm.py:1047-1057: mean over S = 64 variational samples of the likelihood mean and the population
    variance of the logits, from ONE launch; the oracle gets the same Philox draws (exported) in fp64."""
    from vae_b200.vfm_torch import CF
    fs, d, B, S = [40, 25, 9][:F], 16, 300, 64
    rng = np.random.default_rng(5)
    offs = np.concatenate(([0], np.cumsum(fs)[:-1]))
    x = np.stack([offs[f] + rng.integers(0, fs[f], B) for f in range(F)], 1).astype(np.int64)
    torch.manual_seed(2)
    m = CF(d, output=output, n_users=fs[0], n_items=fs[1], train_counts=torch.ones(sum(fs)), field_sizes=fs,
           kl_weighting="torch" if F == 2 else "group", n_train=B, max_batch=B, seed=21)
    with torch.no_grad():
        m.entity_params.weight.mul_(0.4)
        m.global_bias_mean.fill_(0.3), m.global_bias_scale.fill_(0.5)
    uniq = np.unique(x)
    rank = np.searchsorted(uniq, x)                                   # [B, F] unique rank of every occurrence
    from vae_b200 import _lib as L
    from vae_b200.engine import make_config
    import ctypes as C
    U = len(uniq)
    cfgS = make_config(1, F, d, sum(fs), S, output, "abs", m._class_bounds, m._class_sizes, B, 21)
    e0 = torch.empty(S, device=DEV); eb = torch.empty(S * U, device=DEV); ee = torch.empty(S * U * d, device=DEV)
    step = int(m.noise_step[0].item())
    L.check(L.lib().vfmb_philox_normals(C.byref(cfgS), torch.from_numpy(uniq).int().to(DEV).data_ptr(), U, step,
                                        e0.data_ptr(), eb.data_ptr(), ee.data_ptr(), torch.cuda.current_stream().cuda_stream))
    e0, eb, ee = e0.cpu().numpy().astype(np.float64), eb.cpu().numpy().reshape(S, U).astype(np.float64), \
        ee.cpu().numpy().reshape(S, U, d).astype(np.float64)
    W = m.entity_params.weight.detach().cpu().numpy().astype(np.float64)
    Bw = m.bias_params.weight.detach().cpu().numpy().astype(np.float64)
    v = W[uniq][None, :, :d] + ee * np.abs(W[uniq][None, :, d:])     # [S, U, d]
    w = Bw[uniq][None, :, 0] + eb * np.abs(Bw[uniq][None, :, 1])      # [S, U]
    vg = v[:, rank]                                                   # [S, B, F, d]
    fm = vg.prod(2).sum(2) if F == 2 else (0.5 * (vg.sum(2) ** 2 - (vg ** 2).sum(2))).sum(2)
    logits = (0.3 + e0 * 0.5)[:, None] + w[:, rank].sum(2) + fm       # [S, B]
    lik_mean = 1 / (1 + np.exp(-logits)) if output == "class" else logits
    pm, lv, lm = m.predict_proba(torch.from_numpy(x).to(DEV), n_samples=S, return_logit_mean=True)
    assert m.noise_step.tolist()[0] == step + 1
    np.testing.assert_allclose(lm.cpu().numpy(), logits.mean(0), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(pm.cpu().numpy(), lik_mean.mean(0), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(lv.cpu().numpy(), logits.var(0), rtol=1e-4, atol=1e-5)
    # per-occurrence draws (vfm.py:440-445): same per-row marginals -- statistical comparison of the batch
    # averages (the shared global-bias draw alone gives the mean logit a standard error of 0.5 / sqrt(S))
    pm2, lv2, lm2 = m.predict_proba(torch.from_numpy(x).to(DEV), n_samples=4 * S, per_occurrence=True, return_logit_mean=True)
    assert abs(lv2.mean().item() - logits.var(0).mean()) < 0.25 * logits.var(0).mean()
    if F == 2:                                                        # E[logit] = the mean prediction when F = 2
        assert abs(lm2.mean().item() - m.predict_mean(torch.from_numpy(x).to(DEV)).mean().item()) < 0.2


def _guarded(t, pad=1024):
    """A same-shape view into a larger allocation whose margins hold a canary pattern."""
    flat = torch.empty(t.numel() + 2 * pad, dtype=t.dtype, device=t.device)
    canary = torch.full((pad,), 0x5A if t.dtype == torch.uint8 else 12345, dtype=t.dtype, device=t.device)
    flat[:pad], flat[-pad:] = canary, canary
    view = flat[pad:pad + t.numel()].view(t.shape)
    view.copy_(t)
    return view, flat, canary


@pytest.mark.parametrize("name", ["sampled_fraction", "sampled_reg_d64", "sampled_reg_d5"])
def test_kernels_stay_inside_their_buffers(name):
    """compute-sanitizer is closed on this GPU pool, so out-of-bounds WRITES are checked the plain way:
    every plan array, scratch buffer and output of a step is re-homed inside a larger allocation with
    canary margins (4 KB on both sides), full training steps run (plan, Philox and injected paths, F = 2
    fused and unfused backward), and the margins must be untouched."""
    from vae_b200 import _lib as L
    from vae_b200.engine import BatchPlan
    meta, g = gu.load(name)
    m = _model(meta, g, 0, seed=13)
    x, y = gu.batch_of(meta, g, 0)
    xd, yd = torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV)
    m.fused_step(xd, yd)                                    # allocates the pipeline and the step buffers
    guards = []
    buf = m._buf
    for nm in ("vs", "ws", "es", "ebs", "cq", "grow", "gws", "pred", "mean", "resid", "rsorted", "partials",
               "counters", "stats", "grad_scalars"):
        t = getattr(buf, nm)
        if t is not None:
            v, flat, can = _guarded(t)
            setattr(buf, nm, v)
            guards.append((nm, flat, can))
    m._fast_io = None
    for plan in m._pipe.ring:
        for nm in ("uniq", "inverse", "seg_off", "occ", "pos_of", "pos_rank", "partner", "urec", "class_off", "z",
                   "meta", "hot", "workspace"):
            v, flat, can = _guarded(getattr(plan, nm))
            setattr(plan, nm, v)
            guards.append((f"plan.{nm}", flat, can))
        plan.struct = L.Plan(L.ptr(plan.uniq), L.ptr(plan.inverse), L.ptr(plan.seg_off), L.ptr(plan.occ),
                             L.ptr(plan.pos_of), L.ptr(plan.pos_rank), L.ptr(plan.partner), L.ptr(plan.urec),
                             L.ptr(plan.class_off), L.ptr(plan.z), L.ptr(plan.meta), L.ptr(plan.hot))
    for reserve in (0, 1):                                  # fused k_gather_score / k_score + k_gather
        L.check(L.lib().vfmb_set_grid_reserve(reserve))
        try:
            for _ in range(2):
                out = m.fused_step(xd, yd)
            m.fused_step(xd, yd, noise=_noise(g, 0))
            m.gradients(xd, yd, noise=_noise(g, 0))
        finally:
            L.lib().vfmb_set_grid_reserve(0)
    torch.cuda.synchronize()
    assert np.isfinite(out["loss"].item())
    pad = 1024
    for nm, flat, can in guards:
        assert torch.equal(flat[:pad], can) and torch.equal(flat[-pad:], can), f"write outside {nm}"


def test_no_cpu_fallback():
    from vae_b200.vfm_torch import CF
    with pytest.raises(RuntimeError):
        CF(4, n_users=3, n_items=3, train_counts=torch.ones(6), device="cpu")


def test_prefetched_and_static_plans_give_identical_steps():
    """The plan depends only on the ids: building it ahead on a side stream, or once per
    recurring batch, must give bit-identical training steps."""
    meta, g = gu.load("sampled_reg_d64")
    xs = [torch.from_numpy(gu.batch_of(meta, g, t)[0]).to(DEV) for t in range(3)]
    ys = [torch.from_numpy(gu.batch_of(meta, g, t)[1]).to(DEV) for t in range(3)]
    results = []
    for mode in ("inline", "prefetch", "static"):
        m = _model(meta, g, 0, seed=11)
        plans = [m.static_plan(x) for x in xs] if mode == "static" else None
        losses = []
        if mode == "prefetch":
            m.prefetch_plan(xs[0])
        for t in range(6):
            i = t % 3
            if mode == "prefetch" and t + 1 < 6:
                m.prefetch_plan(xs[(t + 1) % 3])
            out = m.fused_step(xs[i], ys[i], plan=plans[i] if plans else None)
            losses.append(out["loss"].item())
        torch.cuda.synchronize()
        results.append((losses, m.entity_params.weight.detach().clone(), m.bias_params.weight.detach().clone()))
    for losses, ent, bias in results[1:]:
        assert losses == results[0][0]
        assert torch.equal(ent, results[0][1]) and torch.equal(bias, results[0][2])


def test_graphed_loop_is_bitwise_identical_to_eager_steps():
    """CUDA-graph replay of (step on batch i || plan of batch i+1) == eager fused_step sequence."""
    meta, g = gu.load("sampled_reg_d64")
    xs = [torch.from_numpy(gu.batch_of(meta, g, t)[0]).to(DEV) for t in range(3)]
    ys = [torch.from_numpy(gu.batch_of(meta, g, t)[1]).to(DEV) for t in range(3)]
    order = [0, 1, 2, 0, 1, 2, 0]
    ref = _model(meta, g, 0, seed=5)
    ref_losses = [ref.fused_step(xs[i], ys[i])["loss"].item() for i in order]
    m = _model(meta, g, 0, seed=5)
    loop = m.graphed_loop(meta["batch"])
    loop.start(xs[order[0]], ys[order[0]])
    losses = []
    for k, i in enumerate(order):
        nxt = order[k + 1] if k + 1 < len(order) else None
        res = loop.step(xs[nxt], ys[nxt]) if nxt is not None else loop.step()
        losses.append(res["loss"].item())
    assert losses == ref_losses
    assert torch.equal(m.entity_params.weight, ref.entity_params.weight)
    assert torch.equal(m.bias_params.weight, ref.bias_params.weight)
    assert torch.equal(m._scalars, ref._scalars) and int(m.adam_step) == len(order)


def test_graphed_loop_depth3_from_pinned_host_matches_eager():
    meta, g = gu.load("sampled_reg_d64")
    xh = [torch.from_numpy(gu.batch_of(meta, g, t)[0]).pin_memory() for t in range(3)]
    yh = [torch.from_numpy(gu.batch_of(meta, g, t)[1]).pin_memory() for t in range(3)]
    order = [0, 1, 2, 2, 1, 0, 1, 2]
    ref = _model(meta, g, 0, seed=9)
    ref_losses = [ref.fused_step(xh[i].to(DEV), yh[i].to(DEV))["loss"].item() for i in order]
    m = _model(meta, g, 0, seed=9)
    loop = m.graphed_loop(meta["batch"], depth=3)
    loop.start(xh[order[0]], yh[order[0]])
    loop.stage(xh[order[1]], yh[order[1]])
    losses = []
    for k in range(len(order)):
        if k + 2 < len(order):
            loop.stage(xh[order[k + 2]], yh[order[k + 2]])
        losses.append(loop.step()["loss"].item())
    assert losses == ref_losses
    assert torch.equal(m.entity_params.weight, ref.entity_params.weight)


def test_training_loop_learns_and_metrics_path():
    """The reference's loop shape (vfm-torch.py:347-384, 402-422) on the bundled fraction data (golden
    batch): fused steps with in-kernel Philox noise drive the ELBO down and the train AUC up; the
    evaluation helper returns the script's test metrics."""
    from vae_b200 import metrics
    meta, g = gu.load("sampled_fraction")
    m = _model(meta, g, 0, seed=11)
    m.configure_adam(0.05)
    x, y = gu.batch_of(meta, g, 0)
    xd, yd = torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV)
    first = m.fused_step(xd, yd)
    loss0, auc0 = first["loss"].item(), metrics.roc_auc(yd, first["pred"]).item()
    for _ in range(150):
        out = m.fused_step(xd, yd)
    loss1, auc1 = out["loss"].item(), metrics.roc_auc(yd, out["pred"]).item()
    assert np.isfinite(loss1) and loss1 < 0.7 * loss0, (loss0, loss1)
    assert auc1 > max(0.75, auc0 + 0.1), (auc0, auc1)
    ev = metrics.evaluate(m, xd[:2000], yd[:2000])
    assert set(ev) == {"auc", "map"} and 0.7 < ev["auc"].item() <= 1.0 and 0.5 < ev["map"].item() <= 1.0


@pytest.mark.parametrize("output", ["reg", "class"])
def test_epoch_level_metrics_equal_reference_loop(output):
    """Two epochs of the script's loop (vfm-torch.py:347-384: model(x) -> loss -> backward -> dense Adam, the
    epoch's train RMSE / AUC / MAP from the per-batch predictions, save_weights()) and its display block
    (:402-422: sampled forward on the test set, RMSE / RMSE-all / RMSE of the last and of the epoch-mean
    posterior, or AUC / MAP), run on the drop-in module with the metrics of vae_b200.metrics on the device,
    against the same loop on the reference restatement with sklearn's metrics -- same noise injected."""
    from sklearn.metrics import average_precision_score, mean_squared_error, roc_auc_score

    from oracle import vfm_port
    from vae_b200 import metrics
    from vae_b200.vfm_torch import CF
    N, M, d, B, n_train, n_test, lr = 80, 50, 8, 256, 1024, 300, 0.02
    rng = np.random.default_rng(21)
    x = np.stack([rng.integers(0, N, n_train + n_test), N + rng.integers(0, M, n_train + n_test)], 1).astype(np.int64)
    y = np.clip(np.round(3.5 + rng.standard_normal(len(x))), 1, 5).astype(np.float32)
    if output == "class":
        y = (y >= 4).astype(np.float32)
    xtr, ytr, xte, yte = x[:n_train], y[:n_train], x[n_train:], y[n_train:]
    tc = np.bincount(xtr.reshape(-1), minlength=N + M)
    tc[tc == 0] = 1
    torch.manual_seed(5)
    port = vfm_port.SampledPort(N, M, d, torch.from_numpy(tc), output=output, faithful_cost=False)
    m = CF(d, output=output, n_users=N, n_items=M, train_counts=torch.from_numpy(tc), n_train=n_train,
           max_batch=max(B, n_test), lr=lr, device=DEV)
    own = m.state_dict()
    m.load_state_dict({k: v.to(DEV).reshape(own[k].shape) for k, v in port.state_dict().items() if k in own}, strict=False)
    opt_m, opt_p = torch.optim.Adam(m.parameters(), lr=lr), torch.optim.Adam(port.parameters(), lr=lr)
    gen = torch.Generator().manual_seed(9)

    def draw(xb):
        U = len(np.unique(xb))
        return [torch.randn(1, 1, generator=gen), torch.randn(1, U, generator=gen), torch.randn(1, U, d, generator=gen)]

    saved = []                                               # the port's save_weights() (vfm-torch.py:179-185)
    all_m, all_p = [], []
    xte_d, yte_d = torch.from_numpy(xte).to(DEV), torch.from_numpy(yte).to(DEV)
    for epoch in range(2):
        pred_m, pred_p = [], []
        for lo in range(0, n_train, B):
            xb, yb = xtr[lo:lo + B], ytr[lo:lo + B]
            noise = draw(xb)
            lik, _, _, kl = m(torch.from_numpy(xb).to(DEV), noise=[n.to(DEV) for n in noise])
            loss = -lik.log_prob(torch.from_numpy(yb).to(DEV)).mean() * n_train + kl
            pred_m.append(lik.mean.squeeze().detach())
            opt_m.zero_grad()
            loss.backward()
            opt_m.step()
            po = vfm_port.sampled_port_step(port, opt_p, torch.from_numpy(xb), torch.from_numpy(yb), n_train, noise)
            pred_p.extend(po["pred"].numpy().tolist())
            np.testing.assert_allclose(loss.item(), po["loss"].item(), rtol=2e-5)
        pm, truth_d = torch.cat(pred_m), torch.from_numpy(ytr).to(DEV)
        if output == "reg":
            m.save_weights()
            saved.append((port.global_bias_mean.detach().numpy().copy(), port.bias_params.weight[:, 0].detach().numpy().copy(),
                          port.entity_params.weight[:, :d].detach().numpy().copy()))
            want = mean_squared_error(ytr, np.clip(pred_p, 1, 5)) ** 0.5
            np.testing.assert_allclose(metrics.rmse(truth_d, pm, clip=(1, 5)).item(), want, rtol=1e-5)
        else:
            np.testing.assert_allclose(metrics.roc_auc(truth_d, pm).item(), roc_auc_score(ytr, pred_p), rtol=1e-5)
            np.testing.assert_allclose(metrics.average_precision(truth_d, pm).item(), average_precision_score(ytr, pred_p), rtol=1e-5)
        # display block: a sampled forward on the test set
        noise = draw(xte)
        with torch.no_grad():
            lik_p, _, _ = port(torch.from_numpy(xte), noise)
            yp = lik_p.mean.squeeze().numpy()

            class _Injected:                                 # evaluate() calls model(x_test): inject the same draws
                output = m.output

                def __call__(self, xq):
                    return m(xq, noise=[n.to(DEV) for n in noise])
            ev = metrics.evaluate(_Injected(), xte_d, yte_d, all_preds=all_m)
        if output == "reg":
            all_p.append(np.clip(yp, 1, 5).tolist())
            gb, bm, em = saved[-1]
            last = gb + bm[xte].sum(axis=1) + em[xte].prod(axis=1).sum(axis=1)
            mgb, mbm, mem = (np.mean([s[i] for s in saved], axis=0) for i in range(3))
            mean_l = mgb + mbm[xte].sum(axis=1) + mem[xte].prod(axis=1).sum(axis=1)
            want = {"rmse": mean_squared_error(yte, np.clip(yp, 1, 5)) ** 0.5,
                    "rmse_all": mean_squared_error(yte, np.array(all_p).mean(axis=0)) ** 0.5,
                    "rmse_of_last": mean_squared_error(yte, last) ** 0.5,
                    "rmse_of_mean": mean_squared_error(yte, np.clip(mean_l, 1, 5)) ** 0.5}
        else:
            want = {"auc": roc_auc_score(yte, yp), "map": average_precision_score(yte, yp)}
        assert set(ev) == set(want), (sorted(ev), sorted(want))
        for k, v in want.items():
            np.testing.assert_allclose(ev[k].item(), v, rtol=2e-5, err_msg=f"epoch {epoch} {k}")


@pytest.mark.parametrize("F,d,B,hot", [
    (2, 4, 1, 0), (2, 5, 33, 1), (2, 12, 511, 1), (2, 16, 512, 0), (2, 20, 513, 1), (2, 32, 1025, 1),
    (2, 64, 2049, 1), (2, 100, 777, 1), (2, 128, 1536, 1), (2, 256, 600, 1), (3, 8, 257, 1), (5, 36, 1000, 1),
    (8, 128, 300, 0)])
def test_gradients_on_ragged_shapes_match_fp64_maths(F, d, B, hot):
    """Every lane layout (d = 4 ... 256: 4 / 8 / 16 / 32 lanes per row, one or two vectors per lane, the scalar
    layout for d % 4 != 0), batch sizes around the gather's group- and block-tile sizes (B * F = 1 ... 4 098
    positions: empty groups, a last short tile, several block tiles), with and without a row that takes a
    third of all occurrences (cut by group tiles, block tiles, or both): gradients, loss and predictions of
    one step against the fp64 maths, and the fused update against torch's Adam recurrence on those gradients."""
    from vae_b200.vfm_torch import CF
    fs = [max(3, 40 - 4 * f) for f in range(F)]
    R = sum(fs)
    rng = np.random.default_rng(1000 * F + d + B)
    offs = np.concatenate(([0], np.cumsum(fs)[:-1]))
    x = np.stack([offs[f] + rng.integers(0, fs[f], B) for f in range(F)], 1).astype(np.int64)
    if hot:
        x[rng.random(B) < 0.33, F - 1] = offs[F - 1]          # one row of the last field in a third of the samples
    out = "reg" if d % 8 else "class"
    y = rng.normal(3.0, 1.0, B).astype(np.float32) if out == "reg" else (rng.random(B) < 0.5).astype(np.float32)
    tc = np.bincount(x.reshape(-1), minlength=R)
    tc[tc == 0] = 1
    kl = "torch" if F == 2 else "group"
    inter = "prod" if F == 2 else "pairwise"
    torch.manual_seed(17)
    lr = 0.01
    m = CF(d, output=out, n_users=fs[0], n_items=fs[1], train_counts=torch.from_numpy(tc), field_sizes=fs,
           kl_weighting=kl, n_train=4 * B, max_batch=B, lr=lr)
    with torch.no_grad():
        m.entity_params.weight.mul_(0.5 if F == 2 else 0.25)
    sd = {k: v.cpu().numpy() for k, v in m.state_dict().items()}
    uniq = np.unique(x)
    U = len(uniq)
    gen = torch.Generator().manual_seed(3)
    noise = [torch.randn(1, 1, generator=gen), torch.randn(1, U, generator=gen), torch.randn(1, U, d, generator=gen)]
    ex = vfm_math.sampled_step(gu.sampled_math_params(sd), x, y, [n.numpy() for n in noise], tc, 4 * B, fs,
                               output=out, interaction=inter, kl_weighting=kl)
    xd, yd = torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV)
    nd = [n.to(DEV) for n in noise]
    gr = m.gradients(xd, yd, noise=nd)
    np.testing.assert_allclose(gr["loss"].item(), ex["loss"], rtol=1e-5)
    want = ex["mean"].squeeze().reshape(-1)
    np.testing.assert_allclose(gr["pred"].cpu().numpy().reshape(-1), want, rtol=1e-5,
                               atol=2e-6 + 1e-6 * float(np.sqrt(np.mean(want ** 2))))
    assert gu.rel_err(gr["entity_params.weight"].cpu().numpy(), ex["grads"]["entity"]) < 1e-5
    assert gu.rel_err(gr["bias_params.weight"].cpu().numpy(), ex["grads"]["bias"]) < 1e-5
    m.fused_step(xd, yd, noise=nd)
    g64 = ex["grads"]["entity"][uniq]
    step1 = sd["entity_params.weight"][uniq] - lr * g64 / (np.abs(g64) + 1e-8)      # Adam, t = 1, zero moments
    got = m.entity_params.weight.detach().cpu().numpy()[uniq]
    err = np.abs(got - step1)
    bad = err > 1e-5 * np.abs(step1) + 1e-4 * lr
    rowmax = np.abs(g64).max(axis=1, keepdims=True)
    assert bad.mean() <= 1e-3 and (not bad.any() or (np.abs(g64) / np.maximum(rowmax, 1e-300))[bad].max() < 1e-3), \
        (float(bad.mean()), float(err.max() / lr))
