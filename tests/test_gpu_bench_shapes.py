"""Parity of the CUDA step with the oracle ON THE SHAPES THE BENCHMARK RUNS (BASELINE.json configs
3, 4, 5 at full batch size), through the code paths the benchmark takes:

* config 3  ml20m      F = 2, d = 64,  B = 65 536, R = 165 237 -- Zipf head rows with thousands of
                       occurrences per batch: they cross a dozen 512-position block tiles of ``k_gather``,
                       so the shared-memory combine AND the global finisher (``finish_cut_row``) run;
* hot3                 F = 2, d = 64, B = 65 536, an item field of THREE rows (~21 800 occurrences each =
                       43 block tiles): the second level of the finisher's tree (rows spanning more than
                       32 block tiles) provably runs;
* config 5  d = 128    F = 2, B = 65 536, Zipf 1.05, table scaled to 300 000 rows (LPR = 32 lanes/row);
* config 4  sideinfo   F = 8, d = 64, B = 65 536, R = 10^6, Bernoulli, pairwise interaction.

Oracles: the fp64 maths (``oracle/vfm_math.py``) everywhere, plus the op-for-op torch port of the
reference (``oracle/vfm_port.py``) for the F = 2 cases.  Two consecutive steps are checked:

  step 1  injected noise (north_star's protocol): ``gradients`` and ``fused_step(noise=...)``;
  step 2  the benchmarked path -- in-kernel Philox noise, lean stage, KL / scalars / step counter
          folded into ``k_adam_rows``, both with and without a block slot reserved for the plan
          (grid sizes differ) -- replayed in the oracle from
          the GPU's exact state (p, m, v, t = 1) and the exported Philox draws.

Tolerances (north_star: fp32 within 1e-5 relative): ELBO / KL rtol 1e-5; predictions rtol 1e-5 plus an
absolute floor of 1e-6 x rms(pred) (a prediction is a d-term dot product of O(1) factors that may cancel
to ~0: fp32 rounding of the terms, ~6e-8 x sum |terms|, is all that is left of it there -- the fp32
reference restatement is off by the same amount against fp64); gradients in max-norm 3e-6 against fp64 maths; updated parameters elementwise |d| <= 1e-5 |p| + 1e-4 lr, where
every element outside that bound must be PROVABLY ill-conditioned: Adam maps g to ~ lr * g/(|g|+eps),
so an element whose fp64 gradient is below 1e-3 of its row's largest (its fp32 value carries >= 1000 x
the usual relative rounding error, the sums being dominated by the row's large terms) turns that
rounding into a change of 1e-3 lr ... lr (the reference's own 1-thread vs N-thread runs differ on exactly those elements; the test
measures that too).  The fraction and the largest error of those elements are recorded in
``gpurun_out/parity_bench_shapes.jsonl``.
"""
import json
import os

import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import vfm_math, vfm_port
from vae_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = 65536


def _case(name):
    if name == "ml20m":
        w = synth.make_workload("ml20m", n_rows=4 * B)
        fs, d, out, inter, kl = w.field_sizes, w.d, "reg", "prod", "torch"
        x, y = w.x, w.y
    elif name == "hot3":
        fs, d, out, inter, kl = [40_000, 3], 64, "reg", "prod", "torch"
        rng = np.random.default_rng(5)
        x = np.stack([synth.make_ids([fs[0]], [0.8], 4 * B)[:, 0], fs[0] + rng.integers(0, 3, 4 * B)], 1).astype(np.int64)
        y = synth._ratings(np.random.default_rng(2), 4 * B)
    elif name == "d128":
        fs, d, out, inter, kl = [210_000, 90_000], 128, "reg", "prod", "torch"
        x = synth.make_ids(fs, [1.05, 1.05], 4 * B)
        y = synth._ratings(np.random.default_rng(1), 4 * B)
    else:
        w = synth.make_workload("sideinfo", n_rows=4 * B)
        fs, d, out, inter, kl = w.field_sizes, w.d, "class", "pairwise", "group"
        x, y = w.x, w.y
    tc = np.bincount(x.reshape(-1), minlength=sum(fs)).astype(np.int64)
    tc[tc == 0] = 1
    return dict(name=name, fs=fs, d=d, output=out, interaction=inter, kl=kl, x=x, y=y, tc=tc, n_train=len(x),
                lr=1.0 / (1 + len(x) // B))


def _model(c, seed=7):
    from vae_b200.vfm_torch import CF
    torch.manual_seed(42)
    m = CF(c["d"], output=c["output"], n_users=c["fs"][0], n_items=c["fs"][1],
           train_counts=torch.from_numpy(c["tc"]), field_sizes=c["fs"], kl_weighting=c["kl"],
           n_train=c["n_train"], max_batch=B, lr=c["lr"], seed=seed, device=DEV)
    if len(c["fs"]) > 2:
        with torch.no_grad():
            m.entity_params.weight.mul_(0.3)             # keep the 8-field sums in a sane range
    return m


def _params(m):
    sd = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    return gu.sampled_math_params(sd), sd


def _adam_rows(p, g, mm, vv, t, lr):
    """torch _single_tensor_adam (fp64) on the given (touched) rows, cf. vfm_math.adam_update."""
    p, g, mm, vv = (np.asarray(a, dtype=np.float64) for a in (p, g, mm, vv))
    mm = mm + 0.1 * (g - mm)
    vv = 0.999 * vv + (1.0 - 0.999) * g * g
    return p - (lr / (1.0 - 0.9 ** t)) * mm / (np.sqrt(vv) / np.sqrt(1.0 - 0.999 ** t) + 1e-8)


def _param_check(got, want, g64, lr, label, rec):
    """Elementwise bound; everything outside it must be ill-conditioned (tiny gradient in its row)."""
    err = np.abs(got - want)
    bad = err > 1e-5 * np.abs(want) + 1e-4 * lr
    rec[label] = {"elements": int(bad.size), "outside_bound": int(bad.sum()), "fraction": float(bad.mean()),
                  "max_err_over_lr_outside": float(err[bad].max() / lr) if bad.any() else 0.0,
                  "max_err_over_lr_inside": float(err[~bad].max() / lr)}
    if bad.any():
        rowmax = np.abs(g64).max(axis=1, keepdims=True)
        rel_g = (np.abs(g64) / np.maximum(rowmax, 1e-300))[bad]
        rec[label]["max_rel_grad_outside"] = float(rel_g.max())
        assert rel_g.max() < 1e-3, (label, "a well-conditioned element is outside the bound", rec[label])
        assert err[bad].max() <= 2.5 * lr, (label, rec[label])          # at worst a sign flip of lr * g/|g|
    assert bad.mean() <= 2e-5, (label, rec[label])


def _pred_close(got, want, msg=""):
    want = np.asarray(want, dtype=np.float64)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6 * float(np.sqrt(np.mean(want ** 2))), err_msg=msg)


def _record(rec):
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_bench_shapes.jsonl"), "a") as fh:
        fh.write(json.dumps(rec) + "\n")


@pytest.mark.parametrize("name", ["ml20m", "d128", "sideinfo", "hot3"])
def test_two_steps_at_benchmark_shape_match_oracle(name):
    from vae_b200 import _lib as L
    c = _case(name)
    fs, d, lr = c["fs"], c["d"], c["lr"]
    F = len(fs)
    rec = {"case": name, "B": B, "F": F, "d": d, "R": int(sum(fs)), "lr": lr}
    m = _model(c)
    x0, y0 = c["x"][:B], c["y"][:B]
    x1, y1 = c["x"][B:2 * B], c["y"][B:2 * B]
    xd0, yd0 = torch.from_numpy(x0).to(DEV), torch.from_numpy(y0).to(DEV)
    xd1, yd1 = torch.from_numpy(x1).to(DEV), torch.from_numpy(y1).to(DEV)

    # ------------------------------------------------------------ step 1: injected noise
    P0, sd0 = _params(m)
    uniq0 = np.unique(x0)
    U0 = len(uniq0)
    gen = torch.Generator().manual_seed(11)
    noise = [torch.randn(1, 1, generator=gen), torch.randn(1, U0, generator=gen), torch.randn(1, U0, d, generator=gen)]
    ex = vfm_math.sampled_step(P0, x0, y0, [n.numpy() for n in noise], c["tc"], c["n_train"], fs, output=c["output"],
                               interaction=c["interaction"], kl_weighting=c["kl"])
    nd = [n.to(DEV) for n in noise]
    gr = m.gradients(xd0, yd0, noise=nd)
    plan = m._plan
    n_hot, n_cut = int(plan.meta[3].item()), int(plan.meta[5].item())
    rec.update(U0=U0, hot_rows=n_hot, cut_rows=n_cut, max_occurrences=int(ex["plan"]["counts"].max()))
    if name != "sideinfo":
        assert n_hot > 0, "rows spanning > 32 nominal tiles (1 024 occurrences) must be present"
    if name == "hot3":
        assert rec["max_occurrences"] > 32 * 512, "a row must span more than 32 block tiles (second finisher level)"
    assert np.array_equal(plan.as_unique()[0].cpu().numpy(), ex["plan"]["uniq"])
    np.testing.assert_allclose(gr["loss"].item(), ex["loss"], rtol=1e-5)
    _pred_close(gr["pred"].cpu().numpy(), ex["mean"].squeeze())
    ge, gb = gr["entity_params.weight"].cpu().numpy(), gr["bias_params.weight"].cpu().numpy()
    rec["grad_entity_vs_fp64"] = gu.rel_err(ge, ex["grads"]["entity"])
    rec["grad_bias_vs_fp64"] = gu.rel_err(gb, ex["grads"]["bias"])
    assert rec["grad_entity_vs_fp64"] < 3e-6 and rec["grad_bias_vs_fp64"] < 3e-6, rec
    # the hot rows alone (they carry the largest sums): row-wise relative error
    hot = np.argsort(ex["plan"]["counts"])[-8:]
    rows_hot = uniq0[hot]
    hot_err = np.abs(ge[rows_hot] - ex["grads"]["entity"][rows_hot]).max(axis=1) / np.abs(ex["grads"]["entity"][rows_hot]).max(axis=1)
    rec["grad_hot_rows_rowwise"] = float(hot_err.max())
    assert hot_err.max() < 1e-5, rec
    for k in ("global_bias_mean", "global_bias_scale") + (("alpha",) if c["output"] == "reg" else ()):
        np.testing.assert_allclose(gr[k].item(), ex["grads"][k][0], rtol=2e-5, err_msg=k)

    if F == 2:      # the reference restatement (fp32 torch, same ATen ops), 1 thread vs all threads
        ports = {}
        for nt in (1, max(1, os.cpu_count() or 1)):
            torch.set_num_threads(nt)
            port = vfm_port.SampledPort(fs[0], fs[1], d, torch.from_numpy(c["tc"]), output=c["output"], faithful_cost=False)
            gu.load_state(port, sd0)
            opt = torch.optim.SGD(port.parameters(), lr=0.0)          # gradients only; Adam is checked in fp64 below
            po = vfm_port.sampled_port_step(port, opt, torch.from_numpy(x0), torch.from_numpy(y0), c["n_train"], noise)
            ports[nt] = po
        p1, pn = ports[1], ports[max(ports)]
        rec["reference_self_diff_grad"] = gu.rel_err(p1["grads"]["entity_params.weight"].numpy(),
                                                     pn["grads"]["entity_params.weight"].numpy())
        rec["grad_entity_vs_port"] = gu.rel_err(ge, pn["grads"]["entity_params.weight"].numpy())
        assert rec["grad_entity_vs_port"] < max(3e-5, 2 * rec["reference_self_diff_grad"]), rec
        np.testing.assert_allclose(gr["loss"].item(), pn["loss"].item(), rtol=1e-5)
        _pred_close(gr["pred"].cpu().numpy(), pn["pred"].numpy(), "vs port")

    before = {k: v.detach().clone() for k, v in m.state_dict().items()}
    out = m.fused_step(xd0, yd0, noise=nd)
    np.testing.assert_allclose(out["loss"].item(), ex["loss"], rtol=1e-5)
    np.testing.assert_allclose(out["kl"].item(), ex["kl"], rtol=1e-5)
    for key, gk in (("entity_params.weight", "entity"), ("bias_params.weight", "bias")):
        p_rows, g_rows = sd0[key][uniq0], ex["grads"][gk][uniq0]
        want = _adam_rows(p_rows, g_rows, np.zeros_like(g_rows), np.zeros_like(g_rows), 1, lr)
        _param_check(m.state_dict()[key].cpu().numpy()[uniq0], want, ex["grads"][gk][uniq0], lr, f"step1.{key}", rec)
    mask = torch.ones(sum(fs), dtype=torch.bool, device=DEV)
    mask[torch.from_numpy(uniq0).to(DEV)] = False
    assert torch.equal(m.entity_params.weight[mask], before["entity_params.weight"][mask])   # untouched rows
    assert int(m.adam_step.item()) == 1

    # ------------------------------------------------------------ step 2: the benchmarked path (Philox)
    for reserve in (1, 0):          # 1: grids as in the graphed bench (one block slot per SM left to the plan);  0: full
        m2 = _model(c)
        m2.load_state_dict(m.state_dict())
        for dst, src in ((m2.entity_m, m.entity_m), (m2.entity_v, m.entity_v), (m2.bias_m, m.bias_m),
                         (m2.bias_v, m.bias_v), (m2._scalars_m, m._scalars_m), (m2._scalars_v, m._scalars_v),
                         (m2.adam_step, m.adam_step), (m2.noise_step, m.noise_step)):
            dst.copy_(src)
        P1, sd1 = _params(m2)
        mom = {k: getattr(m2, k).cpu().numpy() for k in ("entity_m", "entity_v", "bias_m", "bias_v")}
        uniq1 = np.unique(x1)
        ph = m2.philox_noise(torch.from_numpy(uniq1))
        ex2 = vfm_math.sampled_step(P1, x1, y1, [n.cpu().numpy() for n in ph], c["tc"], c["n_train"], fs,
                                    output=c["output"], interaction=c["interaction"], kl_weighting=c["kl"])
        L.check(L.lib().vfmb_set_grid_reserve(reserve))
        try:
            out2 = m2.fused_step(xd1, yd1)                     # fast path: Philox inside the kernels
            torch.cuda.synchronize()
        finally:
            L.lib().vfmb_set_grid_reserve(0)
        tag = f"step2.reserve{reserve}"
        np.testing.assert_allclose(out2["loss"].item(), ex2["loss"], rtol=1e-5, err_msg=tag)
        np.testing.assert_allclose(out2["kl"].item(), ex2["kl"], rtol=1e-5, err_msg=tag)
        _pred_close(out2["pred"].cpu().numpy(), ex2["mean"].squeeze(), tag)
        for key, gk, mk, vk in (("entity_params.weight", "entity", "entity_m", "entity_v"),
                                ("bias_params.weight", "bias", "bias_m", "bias_v")):
            want = _adam_rows(sd1[key][uniq1], ex2["grads"][gk][uniq1], mom[mk][uniq1], mom[vk][uniq1], 2, lr)
            _param_check(m2.state_dict()[key].cpu().numpy()[uniq1], want, ex2["grads"][gk][uniq1], lr, f"{tag}.{key}", rec)
        assert int(m2.adam_step.item()) == 2 and m2.noise_step.tolist() == [m.noise_step[0].item() + 1, m.noise_step[0].item()]
        # scalar parameters (alpha, global bias) after the in-kernel update
        t = 2
        for k, idx in (("global_bias_mean", L.S_GB_MEAN), ("global_bias_scale", L.S_GB_SCALE)) + \
                ((("alpha", L.S_ALPHA),) if c["output"] == "reg" else ()):
            g = ex2["grads"][k]
            want = _adam_rows(sd1[k].reshape(1, 1), np.asarray(g).reshape(1, 1),
                              m._scalars_m[idx].cpu().numpy().reshape(1, 1), m._scalars_v[idx].cpu().numpy().reshape(1, 1),
                              t, lr)
            np.testing.assert_allclose(m2.state_dict()[k].cpu().numpy().reshape(-1), want.reshape(-1), rtol=1e-5,
                                       atol=1e-4 * lr, err_msg=f"{tag}.{k}")
    _record(rec)


def test_graphed_loop_at_benchmark_shape_is_bitwise_identical_to_eager_steps():
    """The default bench configuration (CUDA-graph replay of step i || plan i+1, one block slot per SM
    reserved) against the eager fused steps, at config 3's full size."""
    c = _case("ml20m")
    xs = [torch.from_numpy(c["x"][i * B:(i + 1) * B]).to(DEV) for i in range(4)]
    ys = [torch.from_numpy(c["y"][i * B:(i + 1) * B]).to(DEV) for i in range(4)]
    order = [0, 1, 2, 3, 0, 1]
    ref = _model(c, seed=5)
    ref_losses = [ref.fused_step(xs[i], ys[i])["loss"].item() for i in order]
    m = _model(c, seed=5)
    loop = m.graphed_loop(B)
    loop.start(xs[order[0]], ys[order[0]])
    losses = []
    for k, i in enumerate(order):
        nxt = order[k + 1] if k + 1 < len(order) else None
        res = loop.step(xs[nxt], ys[nxt]) if nxt is not None else loop.step()
        losses.append(res["loss"].item())
    assert np.isfinite(losses).all()
    # unfused score/gather (graph) and fused k_gather_score (eager) score a sample with the same
    # arithmetic; the summation order of the segmented reduction is the plan's in both
    assert losses == ref_losses
    assert torch.equal(m.entity_params.weight, ref.entity_params.weight)
    assert torch.equal(m.bias_params.weight, ref.bias_params.weight)
    assert torch.equal(m._scalars, ref._scalars) and int(m.adam_step) == len(order)
