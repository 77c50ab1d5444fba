"""Generate ``tests/golden/*.npz`` from the UNMODIFIED reference classes.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Run in the build
container (needs ``/root/reference``):

    python -m oracle.gen_golden            # rewrites tests/golden/*.npz

Each file holds the inputs (ids, targets, train counts, initial parameters,
the injected N(0,1) draws) and what ``class CF`` of the reference + the loss /
backward / dense-Adam lines of its training loop produced: per-step loss, KL
and predictions, the step-0 gradients, and the parameters after every step.
The plan arrays (``torch.unique`` outputs) of every step are stored too --
they are the bit-exact targets for the CUDA plan kernels.

fp32 throughout (the reference's dtype).  Seeds: data 20221217, parameters
42 (vfm-torch.py:15), noise 7.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_slice                       # noqa: E402
from vae_b200 import synth                          # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy().copy()


def _state(model):
    return {f"{k}": _np(v) for k, v in model.state_dict().items()}


def _adam_state(model, opt, skip=("prec_",)):
    """exp_avg / exp_avg_sq after the step, so that every golden step can be
    replayed from its exact pre-step (p, m, v, t) state (SURVEY N5 protocol)."""
    out = {}
    for k, p in model.named_parameters():
        st = opt.state.get(p, None)
        if not st or k.startswith(skip):
            continue
        out[f"{k}.m"], out[f"{k}.v"] = _np(st["exp_avg"]), _np(st["exp_avg_sq"])
    return out


def _save(name, meta, arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, meta=np.array(json.dumps(meta)), **arrays)
    print(f"{path}: {os.path.getsize(path) / 1024:.0f} KiB")


def _batch_lo(t, batch, rows):
    """Batch t of the never-shuffled loader (vfm-torch.py:121-122); epochs wrap."""
    n_batches = -(-rows // batch)
    return (t % n_batches) * batch


def fraction_data():
    """Config 1 inputs: data/fraction/data.csv -> X=[user, 536+item], y=outcome,
    seeded 80/20 split."""
    import pandas as pd
    df = pd.read_csv(os.path.join(ref_slice.REFERENCE_DIR, "data", "fraction", "data.csv"))
    N, M = int(df["user"].nunique()), int(df["item"].nunique())
    x = np.stack([df["user"].to_numpy(), N + df["item"].to_numpy()], axis=1).astype(np.int64)
    y = df["outcome"].to_numpy().astype(np.float32)
    perm = np.random.default_rng(synth.DATA_SEED).permutation(len(x))
    n_train = int(0.8 * len(x))
    tr, te = perm[:n_train], perm[n_train:]
    return N, M, x[tr], y[tr], x[te], y[te]


def gen_sampled(name, N, M, d, x, y, n_train, batch, steps, output, lr, S=1, link="abs",
                train_counts=None):
    torch.manual_seed(synth.PARAM_SEED)
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    tc = torch.from_numpy(train_counts) if train_counts is not None else \
        torch.bincount(xt[:n_train].flatten(), minlength=N + M)
    CF = ref_slice.sampled_cf_class(N, M, d, tc, n_var_samples=S, link=link)
    model = CF(d, output=output)
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    arrays = {"x": x[: batch * steps].astype(np.int32), "y": y[: batch * steps],
              "train_counts": _np(tc).astype(np.int64)}
    arrays.update({f"init.{k}": v for k, v in _state(model).items()})
    gen = torch.Generator().manual_seed(synth.NOISE_SEED)
    for t in range(steps):
        lo = _batch_lo(t, batch, len(xt))
        xb, yb = xt[lo:lo + batch], yt[lo:lo + batch]
        uniq, inverse, counts = torch.unique(xb, return_inverse=True, return_counts=True)
        U = len(uniq)
        noise = [torch.randn(S, 1, generator=gen), torch.randn(S, U, generator=gen),
                 torch.randn(S, U, d, generator=gen)]
        out = ref_slice.sampled_step(model, opt, xb, yb, n_train, noise)
        for i, e in enumerate(noise):
            arrays[f"step{t}.noise{i}"] = _np(e)
        arrays[f"step{t}.uniq"] = _np(uniq)
        arrays[f"step{t}.inverse"] = _np(inverse).astype(np.int32)
        arrays[f"step{t}.counts"] = _np(counts)
        for k in ("loss", "kl", "pred", "nll_mean"):
            arrays[f"step{t}.{k}"] = _np(out[k])
        if t == 0:
            for k, g in out["grads"].items():
                if g is not None:
                    arrays[f"step0.grad.{k}"] = _np(g)
        arrays.update({f"step{t}.after.{k}": v for k, v in _state(model).items()
                       if not k.startswith("prec_")})
        arrays.update({f"step{t}.adam.{k}": v for k, v in _adam_state(model, opt).items()})
    meta = dict(variant="sampled", N=N, M=M, d=d, S=S, link=link, output=output, lr=lr,
                n_train=int(n_train), batch=int(batch), steps=steps,
                source="vfm-torch.py:129-324,351-370 (AST-sliced, unmodified)",
                torch=torch.__version__)
    _save(name, meta, arrays)


def gen_saved_logits(name, N, M, d, x, y, n_train, batch, steps, lr):
    """Evaluation outputs of the reference (vfm-torch.py:179-185, 248-262): ``save_weights()`` after
    each of ``steps`` training steps, then ``last_logits`` (last snapshot) and ``mean_logits`` (mean
    over the snapshots) of a forward on an evaluation batch."""
    torch.manual_seed(synth.PARAM_SEED)
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    tc = torch.bincount(xt[:n_train].flatten(), minlength=N + M)
    CF = ref_slice.sampled_cf_class(N, M, d, tc)
    model = CF(d, output="reg")
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    arrays = {"x": x.astype(np.int32), "y": y, "train_counts": _np(tc).astype(np.int64)}
    arrays.update({f"init.{k}": v for k, v in _state(model).items()})
    gen = torch.Generator().manual_seed(synth.NOISE_SEED)
    for t in range(steps):
        lo = _batch_lo(t, batch, len(xt))
        xb, yb = xt[lo:lo + batch], yt[lo:lo + batch]
        U = len(torch.unique(xb))
        noise = [torch.randn(1, 1, generator=gen), torch.randn(1, U, generator=gen), torch.randn(1, U, d, generator=gen)]
        ref_slice.sampled_step(model, opt, xb, yb, n_train, noise)
        model.save_weights()
        arrays.update({f"step{t}.after.{k}": v for k, v in _state(model).items() if not k.startswith("prec_")})
    x_eval = xt[-batch:]
    with torch.no_grad():
        _, last_logits, mean_logits, _ = model(x_eval)
    arrays["x_eval"] = _np(x_eval).astype(np.int32)
    arrays["last_logits"] = np.asarray(last_logits, dtype=np.float32)
    arrays["mean_logits"] = np.asarray(mean_logits, dtype=np.float32)
    meta = dict(variant="sampled", N=N, M=M, d=d, S=1, link="abs", output="reg", lr=lr, n_train=int(n_train),
                batch=int(batch), steps=steps, source="vfm-torch.py:179-185,248-262 (AST-sliced, unmodified)",
                torch=torch.__version__)
    _save(name, meta, arrays)


def gen_closed(name, group_sizes, d, x, y, n_train, batch, steps, lr, alpha_0, perturb=0.0):
    torch.manual_seed(synth.PARAM_SEED)
    G = len(group_sizes)
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    tc = torch.bincount(xt[:n_train].flatten(), minlength=sum(group_sizes)).float()   # :182
    CF = ref_slice.closed_cf_class(group_sizes[0], group_sizes[1])
    model = CF(embedding_size=d, n_groups=G, group_sizes=list(group_sizes), alpha_0=alpha_0)
    if perturb:
        gen = torch.Generator().manual_seed(synth.NOISE_SEED)
        with torch.no_grad():
            model.entity_params[:, :d] += perturb * torch.randn(sum(group_sizes), d, generator=gen)
            for g in range(G):
                model.mean_group_entity_prior[g] += 0.1 * torch.randn(d, generator=gen)
                model.scale_group_entity_prior[g] *= 1 + 0.1 * torch.randn(d, generator=gen)
                model.mean_group_bias_prior[g] += 0.05 * (g + 1)
                model.scale_group_bias_prior[g] *= 0.9
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    arrays = {"x": x[: batch * steps].astype(np.int32), "y": y[: batch * steps],
              "train_counts": _np(tc).astype(np.int64)}
    arrays.update({f"init.{k}": v for k, v in _state(model).items()})
    for t in range(steps):
        lo = _batch_lo(t, batch, len(xt))
        xb, yb = xt[lo:lo + batch], yt[lo:lo + batch]
        present, inverse, count = ref_slice.closed_plan(xb)
        out = ref_slice.closed_step(model, opt, xb, yb, n_train, tc, list(group_sizes))
        for g in range(G):
            arrays[f"step{t}.present{g}"] = _np(present[g])
            arrays[f"step{t}.inverse{g}"] = _np(inverse[g]).astype(np.int32)
            arrays[f"step{t}.counts{g}"] = _np(count[g])
        for k in ("loss", "kl", "pred", "partial_loss"):
            arrays[f"step{t}.{k}"] = _np(out[k])
        if t == 0:
            for k, g in out["grads"].items():
                if g is not None:
                    arrays[f"step0.grad.{k}"] = _np(g)
        arrays.update({f"step{t}.after.{k}": v for k, v in _state(model).items()})
        arrays.update({f"step{t}.adam.{k}": v for k, v in _adam_state(model, opt).items()})
    meta = dict(variant="closed", group_sizes=list(map(int, group_sizes)), d=d, lr=lr,
                alpha_0=alpha_0, n_train=int(n_train), batch=int(batch), steps=steps,
                source="vfm-tomasrch.py:186-453,535-594 (AST-sliced, unmodified)",
                torch=torch.__version__)
    _save(name, meta, arrays)


def _small_ids(field_sizes, exps, rows, seed):
    return synth.make_ids(field_sizes, exps, rows, seed)


def main():
    assert ref_slice.available(), "needs /root/reference"
    # one thread: ATen's CPU scatter-add order (and so the fp32 gradient bits) depends on the
    # thread count; the goldens are the single-thread result (reproducible).
    torch.set_num_threads(1)
    # config 1: fraction, Bernoulli, d=20, full batch (every row touched every step)
    N, M, xtr, ytr, _, _ = fraction_data()
    n_train = len(xtr)
    gen_sampled("sampled_fraction", N, M, 20, xtr, ytr, n_train, n_train, 3, "class",
                lr=1.0 / (1 + n_train // 100000))
    # sampled Gaussian, partial coverage, d=64 (the config-3 row width), three dense-Adam steps
    rng = np.random.default_rng(synth.DATA_SEED)
    fs = [180, 76]
    x = _small_ids(fs, [0.5, 1.0], 1536, synth.DATA_SEED)
    y = np.clip(np.round(3.5 + rng.standard_normal(len(x))), 1, 5).astype(np.float32)
    gen_sampled("sampled_reg_d64", fs[0], fs[1], 64, x, y, len(x), 512, 3, "reg", lr=0.05)
    # odd width (d=5, the scripts' default EMBEDDING_SIZE), softplus link, S=2
    fs = [60, 40]
    x = _small_ids(fs, [0.5, 1.0], 1536, synth.DATA_SEED + 1)
    y = np.clip(np.round(3.5 + rng.standard_normal(len(x))), 1, 5).astype(np.float32)
    gen_sampled("sampled_reg_d5", fs[0], fs[1], 5, x, y, len(x), 512, 3, "reg", lr=0.05)
    gen_sampled("sampled_reg_softplus", fs[0], fs[1], 8, x, y, len(x), 512, 2, "reg", lr=0.05,
                link="softplus")
    gen_sampled("sampled_class_s2", fs[0], fs[1], 8, x, (y > 3).astype(np.float32), len(x), 512,
                2, "class", lr=0.05, S=2)
    gen_saved_logits("sampled_saved_logits", fs[0], fs[1], 8, x, y, len(x), 512, 3, lr=0.05)
    # config 2: ML-100K-shaped closed form, d=20, B=8000, lr 0.1, alpha_0 = 0.5*ceil(80000/8000)
    w = synth.make_workload("ml100k")
    gen_closed("closed_ml100k", w.field_sizes, 20, w.x, w.y, w.n_train, 8000, 2, lr=0.1,
               alpha_0=5.0)
    # closed form, three groups (the fr_en layout [3, M, N], vfm-tomasrch.py:160), perturbed init
    fs = [3, 40, 60]
    x = _small_ids(fs, [0.5, 1.0, 0.5], 1536, synth.DATA_SEED + 2)
    y = np.clip(np.round(3.5 + rng.standard_normal(len(x))), 1, 5).astype(np.float32)
    gen_closed("closed_3groups", fs, 8, x, y, len(x), 512, 3, lr=0.1, alpha_0=1.5, perturb=0.3)
    gen_closed("closed_2groups_d64", [160, 96], 64,
               _small_ids([160, 96], [0.5, 1.0], 1536, synth.DATA_SEED + 3),
               np.clip(np.round(3.5 + rng.standard_normal(1536)), 1, 5).astype(np.float32),
               1536, 512, 3, lr=0.1, alpha_0=1.5, perturb=0.3)


if __name__ == "__main__":
    main()
