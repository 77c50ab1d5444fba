"""Run the *unmodified* reference model classes as the parity pin.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

The reference scripts cannot be imported (they load absent datasets and train
at import time), so ``class CF`` is cut out of each script's AST at run time
and executed with the module-level names it expects injected as globals.
No reference source is copied into this repository: the text is read from
``/root/reference`` (override with ``$VFM_REFERENCE_DIR``), which exists only
in the build container.  Everything here therefore degrades to
``available() == False`` on the GPU box.

The few lines of the scripts' training loops that live outside the class
(loss assembly, ``zero_grad/backward/step``) are driven here in the same
order as the scripts do:

* sampled ELBO      ``vfm-torch.py:353-370``
* closed form       ``vfm-tomasrch.py:536-594``
"""
from __future__ import annotations

import ast
import contextlib
import os
from typing import Iterable, Sequence

import numpy as np
import torch

REFERENCE_DIR = os.environ.get("VFM_REFERENCE_DIR", "/root/reference")
_TORCH_SCRIPT = "vfm-torch.py"
_CLOSED_SCRIPT = "vfm-tomasrch.py"


def available() -> bool:
    return all(os.path.isfile(os.path.join(REFERENCE_DIR, s))
               for s in (_TORCH_SCRIPT, _CLOSED_SCRIPT))


def _slice_class(script: str, name: str = "CF") -> str:
    path = os.path.join(REFERENCE_DIR, script)
    with open(path) as fh:
        src = fh.read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name == name:
            return ast.get_source_segment(src, node)
    raise RuntimeError(f"class {name} not found in {path}")


def _exec_class(script: str, env: dict):
    code = compile(_slice_class(script), os.path.join(REFERENCE_DIR, script), "exec")
    exec(code, env)
    return env["CF"]


_LINKS = {"abs": torch.abs, "softplus": torch.nn.functional.softplus}


def sampled_cf_class(N: int, M: int, embedding_size: int, nb_occ: torch.Tensor,
                     n_var_samples: int = 1, link: str = "abs"):
    """``class CF`` of vfm-torch.py:129-324 bound to the globals the script
    keeps at module level (vfm-torch.py:18-19, 87-89, 125-126)."""
    env = {
        "torch": torch, "nn": torch.nn, "distributions": torch.distributions, "np": np,
        "N": N, "M": M, "EMBEDDING_SIZE": embedding_size,
        "N_VARIATIONAL_SAMPLES": n_var_samples, "LINK": _LINKS[link], "nb_occ": nb_occ,
    }
    return _exec_class(_TORCH_SCRIPT, env)


def closed_cf_class(N: int, M: int):
    """``class CF`` of vfm-tomasrch.py:186-453 (N, M only feed the default
    ``group_sizes``, vfm-tomasrch.py:199)."""
    env = {"torch": torch, "nn": torch.nn, "distributions": torch.distributions,
           "np": np, "N": N, "M": M}
    return _exec_class(_CLOSED_SCRIPT, env)


@contextlib.contextmanager
def injected_noise(draws: Iterable[torch.Tensor]):
    """Replace the N(0,1) source behind ``Normal.rsample`` by a queue.

    ``rsample`` is ``loc + eps * scale`` with ``eps = _standard_normal(shape)``
    (torch/distributions/normal.py), so feeding eps in draw order -- global
    ``[S,1]``, bias ``[S,U]``, entity ``[S,U,d]`` (vfm-torch.py:238-241) --
    makes the reference deterministic and lets the CUDA path consume the same
    noise."""
    import torch.distributions.normal as _normal
    queue = list(draws)
    original = _normal._standard_normal

    def _pop(shape, dtype, device):
        eps = queue.pop(0)
        assert tuple(eps.shape) == tuple(shape), (tuple(eps.shape), tuple(shape))
        return eps.to(dtype=dtype, device=device)

    _normal._standard_normal = _pop
    try:
        yield
    finally:
        _normal._standard_normal = original
    assert not queue, "unused noise draws"


def sampled_step(model, optimizer, x: torch.Tensor, y: torch.Tensor, n_train: int,
                 noise: Sequence[torch.Tensor] | None = None, update: bool = True) -> dict:
    """One batch of the loop at vfm-torch.py:351-370."""
    ctx = injected_noise(noise) if noise is not None else contextlib.nullcontext()
    with ctx:
        outputs, last_logits, mean_logits, kl_term = model(x)
    nll_mean = -outputs.log_prob(y.float()).mean()
    loss = nll_mean * n_train + kl_term
    pred = outputs.mean.squeeze().detach().clone()
    out = {"loss": loss.detach().clone(), "kl": kl_term.detach().clone(),
           "nll_mean": nll_mean.detach().clone(), "pred": pred}
    if update:
        optimizer.zero_grad()
        loss.backward()
        out["grads"] = {k: (p.grad.detach().clone() if p.grad is not None else None)
                        for k, p in model.named_parameters()}
        optimizer.step()
    return out


def closed_plan(x: torch.Tensor):
    """Per-column unique of vfm-tomasrch.py:536-545."""
    present, inverse, count = [], [], []
    for g in range(x.shape[1]):
        p, i, c = torch.unique(x[:, g], return_inverse=True, return_counts=True)
        present.append(p), inverse.append(i), count.append(c)
    return present, inverse, count


def closed_step(model, optimizer, x: torch.Tensor, y: torch.Tensor, n_train: int,
                entity_count: torch.Tensor, group_sizes: Sequence[int],
                bounds=(1, 5), update: bool = True) -> dict:
    """One batch of the loop at vfm-tomasrch.py:535-594 (Adam branch)."""
    n_groups = len(group_sizes)
    present, inverse, count = closed_plan(x)
    outputs, kls, partial_loss = model(inverse, present, closed_form_loss=True, target=y)
    pred = outputs.mean.detach().clone()
    weights = torch.cat([
        torch.Tensor(group_sizes[g] / (count[g] / entity_count[present[g]]).sum()
                     ).repeat(len(present[g]))
        for g in range(n_groups)])
    kl_rescaled = ((kls[1] + kls[2].sum(axis=1)) * weights
                   * torch.concat(count) / entity_count[torch.concat(present)]).sum()
    loss = -n_train * partial_loss / len(x) + kls[0] + kl_rescaled
    out = {"loss": loss.detach().clone(), "pred": pred, "pred_clipped": pred.clip(*bounds),
           "partial_loss": partial_loss.detach().clone(),
           "kl": (kls[0] + kl_rescaled).detach().clone()}
    if update:
        optimizer.zero_grad()
        loss.backward()
        out["grads"] = {k: (p.grad.detach().clone() if p.grad is not None else None)
                        for k, p in model.named_parameters()}
        optimizer.step()
    return out
