"""Data preparation of the reference (``prepare.py:10-64``; the preamble of ``vfm-torch.py:87-122``),
restated for this package: ``data/<name>/data.csv`` (+ optional fold files) -> id arrays in the
model's row numbering, train counts, and contiguous never-shuffled batches.

Host-side I/O (pandas / numpy), off the measured path -- the step kernels take device tensors.

Differences from the reference, all documented in SURVEY.md section 8a:
* N9: ``train_counts`` has ``minlength = rows`` (``np.bincount(X_train.flatten())`` without it fails
  when the largest id never occurs in training) and rows never seen in training get count 1;
* N13: datasets without fold files (the bundled ``fraction``) get a seeded 80/20 split, as the
  TF script's fallback does unseeded (``vfm.py:211-212``).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Iterator, Optional, Tuple

import numpy as np


@dataclass
class Interactions:
    n_users: int
    n_items: int
    x_train: np.ndarray        # int64 [R_train, 2] = [user, n_users + item]   (prepare.py:47)
    y_train: np.ndarray        # float32 [R_train]
    x_test: np.ndarray
    y_test: np.ndarray
    folds: dict                # {"trainval": row indices, "test": row indices} into data.csv

    @property
    def rows(self) -> int:
        return self.n_users + self.n_items

    def train_counts(self) -> np.ndarray:
        """``nb_occ`` of vfm-torch.py:89 with ``minlength = rows``; unseen rows count 1 (N9)."""
        tc = np.bincount(self.x_train.reshape(-1), minlength=self.rows).astype(np.int64)
        tc[tc == 0] = 1
        return tc


def reindex(users: np.ndarray, items: np.ndarray) -> Tuple[np.ndarray, np.ndarray, int, int]:
    """``prepare_data`` (prepare.py:45-47): users and items to 0..N-1 / 0..M-1 in sorted order."""
    u_vals, u = np.unique(users, return_inverse=True)
    i_vals, i = np.unique(items, return_inverse=True)
    return u.astype(np.int64), i.astype(np.int64), len(u_vals), len(i_vals)


def load_data(data_dir: str, output_type: str = "reg", test_fraction: float = 0.2,
              seed: int = 20221217) -> Interactions:
    """``load_data`` (prepare.py:10-37) on ``<data_dir>/data.csv`` with columns ``user, item`` and
    ``rating`` (``output_type='reg'``) or ``outcome`` (``'class'``).  ``trainval.csv`` / ``test.csv``
    (one ``index`` column of row numbers) are used when present, else a seeded split (N13).
    Column 1 of the returned ids is ``shifted_item = n_users + item``."""
    import pandas as pd
    df = pd.read_csv(os.path.join(data_dir, "data.csv"))
    outcome = "rating" if output_type == "reg" else "outcome"
    if outcome not in df.columns:
        raise KeyError(f"{data_dir}/data.csv has no '{outcome}' column (output_type={output_type!r})")
    u, i, n_users, n_items = reindex(df["user"].to_numpy(), df["item"].to_numpy())
    x = np.stack([u, n_users + i], axis=1)
    y = df[outcome].to_numpy().astype(np.float32)
    tv, te = os.path.join(data_dir, "trainval.csv"), os.path.join(data_dir, "test.csv")
    if os.path.isfile(tv) and os.path.isfile(te):
        folds = {"trainval": pd.read_csv(tv)["index"].to_numpy(), "test": pd.read_csv(te)["index"].to_numpy()}
    else:
        perm = np.random.default_rng(seed).permutation(len(df))
        n_test = int(round(test_fraction * len(df)))
        folds = {"trainval": np.sort(perm[n_test:]), "test": np.sort(perm[:n_test])}
    a, b = folds["trainval"], folds["test"]
    return Interactions(n_users, n_items, x[a], y[a], x[b], y[b], folds)


def batches(x: np.ndarray, y: np.ndarray, batch_size: int, device=None) -> Iterator[Tuple["torch.Tensor", "torch.Tensor"]]:
    """``DataLoader(TensorDataset(X, y), batch_size)`` of vfm-torch.py:119-122: contiguous slices in
    file order, never shuffled (N8) -- which is what lets ``CF.static_plan`` cache a batch's plan.
    With ``device`` the whole split is uploaded once and the batches are device views."""
    import torch
    xt, yt = torch.from_numpy(np.ascontiguousarray(x)), torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32))
    if device is not None:
        xt, yt = xt.to(device), yt.to(device)
    for lo in range(0, len(xt), batch_size):
        yield xt[lo:lo + batch_size], yt[lo:lo + batch_size]


def write_libfm(path: str, x: np.ndarray, y: np.ndarray) -> None:
    """The libFM export of ``prepare_data`` (prepare.py:57-61): ``<outcome> <user>:1 <shifted_item>:1``."""
    with open(path, "w") as fh:
        for (user, item), outcome in zip(x, y):
            fh.write("{:d} {:d}:1 {:d}:1\n".format(int(outcome), int(user), int(item)))
