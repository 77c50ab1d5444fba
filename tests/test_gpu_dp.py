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


def _run_sharded_emulated(ranks, xs, ys):
    """Lock-step emulation of P ranks on one device: all-to-all = transpose of the send buffers."""
    P = len(ranks)
    req = [r.phase_request(x, y) for r, x, y in zip(ranks, xs, ys)]
    z = sum(q[1] for q in req)
    a2a = lambda bufs: [torch.stack([bufs[src][dst] for src in range(P)]).contiguous() for dst in range(P)]
    recv = a2a([q[0] for q in req])
    replies = [r.phase_owner_stage(rv, z) for r, rv in zip(ranks, recv)]
    rows = a2a(replies)
    loc = [r.phase_local(rw) for r, rw in zip(ranks, rows)]
    tail = sum(t[1].clone() for t in loc)
    grads = a2a([t[0] for t in loc])
    return [r.phase_owner_update(g_, tail.clone()) for r, g_ in zip(ranks, grads)]


def _run_sharded_fused(ranks, xs, ys):
    """Lock-step emulation of the fused peer-memory step: the emulated ranks share a LocalPeerGroup, every
    phase runs on all ranks before the next one starts (what the cross-rank barriers guarantee)."""
    for r, x, y in zip(ranks, xs, ys):
        r._use_slot(0)
        r.phase_request(x, y)                 # local plan, requests into the owners' regions, routing tables
    for r in ranks:
        r._owner_prepare(None, None)          # owner's plan of the received ids, global batch counts
    for r in ranks:
        r._b_stage()
    for r in ranks:
        r._b_local()
    return [r._b_update() for r in ranks]


@pytest.mark.parametrize("mode", ["collective", "fused"])
@pytest.mark.parametrize("P", [2, 3, 4, 8])
@pytest.mark.parametrize("name", ["sampled_reg_d64", "sampled_fraction"])
def test_row_sharded_mode_b_equals_single_process_step(name, P, mode):
    """Mode B (row-sharded tables, all-to-all of sampled rows and gradients), P ranks emulated on one
    GPU, against the single-process fused step on the global batch with the same per-entity noise."""
    from vae_b200.dist import ShardedSampled
    meta, g = gu.load(name)
    N, M, d = meta["N"], meta["M"], meta["d"]
    R = N + M
    x, y = gu.batch_of(meta, g, 0)
    n = (len(x) // P) * P
    x, y = x[:n], y[:n]
    gen = torch.Generator().manual_seed(3)
    e0 = torch.randn(1, generator=gen).to(DEV)
    eb_t, ee_t = torch.randn(R, generator=gen).to(DEV), torch.randn(R, d, generator=gen).to(DEV)
    init = gu.state(g, "init")
    # single process on the global batch
    ref = S._model(meta, g, 0)
    ref._cfg_cache.clear()
    uniq = torch.from_numpy(np.unique(x)).to(DEV)
    for step in range(2):
        out_ref = ref.fused_step(torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV),
                                 noise=(e0.reshape(1, 1), eb_t[uniq][None], ee_t[uniq][None]))
    loss_ref = out_ref["loss"].item()
    # P ranks
    ini = {"bias": torch.from_numpy(init["bias_params.weight"]), "entity": torch.from_numpy(init["entity_params.weight"]),
           "alpha": init["alpha"][0], "global_bias_mean": init["global_bias_mean"][0],
           "global_bias_scale": init["global_bias_scale"][0]}
    from vae_b200.dist import LocalPeerGroup
    ex = LocalPeerGroup(P) if mode == "fused" else object()
    ranks = [ShardedSampled(d, [N, M], torch.from_numpy(g["train_counts"]), meta["n_train"], n // P, P, p,
                            output=meta["output"], link=meta["link"], lr=meta["lr"], init=ini,
                            noise_tables=(e0, eb_t, ee_t), exchange=ex) for p in range(P)]
    xs = [torch.from_numpy(x[p * (n // P):(p + 1) * (n // P)]).to(DEV) for p in range(P)]
    ys = [torch.from_numpy(y[p * (n // P):(p + 1) * (n // P)]).to(DEV) for p in range(P)]
    for step in range(2):
        outs = (_run_sharded_fused if mode == "fused" else _run_sharded_emulated)(ranks, xs, ys)
    for r in ranks:
        r.check_overflow()
        assert int(r.adam_step.item()) == 2
    ent = torch.zeros(R, 2 * d, device=DEV)
    bias = torch.zeros(R, 2, device=DEV)
    for r in ranks:
        gid, b, e = r.gather_tables()
        ent[gid], bias[gid] = e, b
    lr = meta["lr"]
    for got, want in ((ent, ref.entity_params.weight.detach()), (bias, ref.bias_params.weight.detach())):
        bad = (got - want).abs() > 1e-5 * want.abs() + 2e-5 * lr + 1e-6
        assert bad.float().mean().item() <= 1e-4, float(bad.float().mean())
    for o in outs:
        np.testing.assert_allclose(o["loss"].item(), loss_ref, rtol=1e-5)
    for r in ranks:
        assert torch.allclose(r.scalars[:3], ref._scalars[:3], rtol=1e-5, atol=2e-6)
    # predictions of each rank's slice
    pred = torch.cat([o["pred"] for o in outs])
    assert torch.allclose(pred, out_ref["pred"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("mode", ["collective", "fused"])
@pytest.mark.parametrize("F,d,P", [(3, 8, 2), (8, 64, 4)])
def test_row_sharded_mode_b_multi_field_equals_single_process_step(F, d, P, mode):
    """Mode B with F > 2 fields (config 4: pairwise interaction, per-group KL weights): P ranks
    emulated on one GPU against the single-process fused step on the global batch."""
    from vae_b200.dist import ShardedSampled
    from vae_b200.vfm_torch import CF
    fs = [37, 23, 11, 9, 7, 5, 4, 3][:F]
    R, B = sum(fs), 256 * P
    rng = np.random.default_rng(F * 10 + P)
    offs = np.concatenate(([0], np.cumsum(fs)[:-1]))
    x = np.stack([offs[f] + rng.integers(0, fs[f], B) for f in range(F)], 1).astype(np.int64)
    y = (rng.random(B) < 0.5).astype(np.float32)
    tc = np.bincount(x.reshape(-1), minlength=R)
    tc[tc == 0] = 1
    torch.manual_seed(3)
    ref = CF(d, output="class", n_users=fs[0], n_items=fs[1], train_counts=torch.from_numpy(tc),
             field_sizes=fs, kl_weighting="group", n_train=B, max_batch=B, lr=0.05)
    with torch.no_grad():
        ref.entity_params.weight.mul_(0.3)
    ini = {"bias": ref.bias_params.weight.detach().cpu().clone(), "entity": ref.entity_params.weight.detach().cpu().clone(),
           "alpha": float(ref.alpha.item()), "global_bias_mean": float(ref.global_bias_mean.item()),
           "global_bias_scale": float(ref.global_bias_scale.item())}
    gen = torch.Generator().manual_seed(5)
    e0 = torch.randn(1, generator=gen).to(DEV)
    eb_t, ee_t = torch.randn(R, generator=gen).to(DEV), torch.randn(R, d, generator=gen).to(DEV)
    uniq = torch.from_numpy(np.unique(x)).to(DEV)
    xd, yd = torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV)
    for _ in range(2):
        out_ref = ref.fused_step(xd, yd, noise=(e0.reshape(1, 1), eb_t[uniq][None], ee_t[uniq][None]))
    from vae_b200.dist import LocalPeerGroup
    ex = LocalPeerGroup(P) if mode == "fused" else object()
    ranks = [ShardedSampled(d, fs, torch.from_numpy(tc), B, B // P, P, p, output="class", kl_weighting="group",
                            lr=0.05, init=ini, noise_tables=(e0, eb_t, ee_t), exchange=ex) for p in range(P)]
    xs = [xd[p * (B // P):(p + 1) * (B // P)] for p in range(P)]
    ys = [yd[p * (B // P):(p + 1) * (B // P)] for p in range(P)]
    for _ in range(2):
        outs = (_run_sharded_fused if mode == "fused" else _run_sharded_emulated)(ranks, xs, ys)
    ent, bias = torch.zeros(R, 2 * d, device=DEV), torch.zeros(R, 2, device=DEV)
    for r in ranks:
        r.check_overflow()
        gid, b, e = r.gather_tables()
        ent[gid], bias[gid] = e, b
    for got, want in ((ent, ref.entity_params.weight.detach()), (bias, ref.bias_params.weight.detach())):
        bad = (got - want).abs() > 1e-5 * want.abs() + 2e-5 * 0.05 + 1e-6
        assert bad.float().mean().item() <= 1e-4, float(bad.float().mean())
    for o in outs:
        np.testing.assert_allclose(o["loss"].item(), out_ref["loss"].item(), rtol=1e-5)
    pred = torch.cat([o["pred"] for o in outs])
    assert torch.allclose(pred, out_ref["pred"], rtol=1e-5, atol=2e-6)
