"""Helpers shared by the parity tests: load ``tests/golden/*.npz`` (outputs of
the unmodified reference, see ``oracle/gen_golden.py``) and rebuild oracle
models / parameter dictionaries from them."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SAMPLED = ["sampled_fraction", "sampled_reg_d64", "sampled_reg_d5", "sampled_reg_softplus",
           "sampled_class_s2"]
CLOSED = ["closed_ml100k", "closed_3groups", "closed_2groups_d64"]


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return meta, {k: z[k] for k in z.files if k != "meta"}


def batch_of(meta, g, t):
    rows, batch = len(g["x"]), meta["batch"]
    lo = (t % (-(-rows // batch))) * batch
    return g["x"][lo:lo + batch].astype(np.int64), g["y"][lo:lo + batch]


def state(g, prefix):
    """Reference ``state_dict`` stored under ``prefix`` ('init' or 'step{t}.after')."""
    p = prefix + "."
    return {k[len(p):]: v for k, v in g.items() if k.startswith(p)}


def sampled_math_params(sd):
    return {"alpha": sd["alpha"], "global_bias_mean": sd["global_bias_mean"],
            "global_bias_scale": sd["global_bias_scale"],
            "bias": sd["bias_params.weight"], "entity": sd["entity_params.weight"]}


def closed_math_params(sd, G):
    return {"alpha": sd["alpha"], "mean_global_bias": sd["mean_global_bias"],
            "scale_global_bias": sd["scale_global_bias"],
            "mean_global_bias_prior": sd["mean_global_bias_prior"],
            "scale_global_bias_prior": sd["scale_global_bias_prior"],
            "bias": sd["bias_params"], "entity": sd["entity_params"],
            "prior_bias_mean": np.concatenate([sd[f"mean_group_bias_prior.{g}"] for g in range(G)]),
            "prior_bias_scale": np.concatenate([sd[f"scale_group_bias_prior.{g}"] for g in range(G)]),
            "prior_entity_mean": np.stack([sd[f"mean_group_entity_prior.{g}"] for g in range(G)]),
            "prior_entity_scale": np.stack([sd[f"scale_group_entity_prior.{g}"] for g in range(G)])}


def load_state(module, sd):
    own = module.state_dict()
    module.load_state_dict({k: torch.from_numpy(np.asarray(v)).reshape(own[k].shape)
                            for k, v in sd.items() if k in own}, strict=False)


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def restore(module, g, t, lr):
    """Module parameters and a dense ``torch.optim.Adam`` positioned exactly
    before golden step ``t`` (parameters, exp_avg, exp_avg_sq, step count)."""
    load_state(module, state(g, "init" if t == 0 else f"step{t - 1}.after"))
    opt = torch.optim.Adam(module.parameters(), lr=lr)
    if t > 0:
        for k, p in module.named_parameters():
            if f"step{t - 1}.adam.{k}.m" not in g:
                continue
            opt.state[p] = {"step": torch.tensor(float(t)),
                            "exp_avg": torch.from_numpy(g[f"step{t - 1}.adam.{k}.m"].copy()),
                            "exp_avg_sq": torch.from_numpy(g[f"step{t - 1}.adam.{k}.v"].copy())}
    return opt
