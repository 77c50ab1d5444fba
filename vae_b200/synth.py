"""Synthetic workloads of the BASELINE.json shapes.

The reference trains on CSV datasets that are not available offline
(``prepare.py:10-37``); the benchmark and the parity tests feed id batches of
the same shapes instead.  Ids are drawn from a bounded Zipf law by inverse
transform and scattered over the field with a fixed odd multiplier so that
hot rows do not cluster (matters for row-sharded tables, ``owner = row mod P``).

Layout contract (same as the reference, ``prepare.py:47`` / ``vfm-torch.py:88``):
``x`` is int64 ``[R, F]`` with column f holding *global* row ids
``field_offset[f] + local id``; ``y`` is float32 ``[R]``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

DATA_SEED = 20221217
PARAM_SEED = 42          # the scripts' torch.manual_seed(42), vfm-torch.py:15
NOISE_SEED = 7


def zipf_ids(n: int, size: int, s: float, rng: np.random.Generator, scatter: bool = True) -> np.ndarray:
    """``size`` ids in [0, n) with P(rank r) ~ (r+1)^-s (continuous inverse CDF)."""
    u = rng.random(size)
    if abs(s - 1.0) < 1e-12:
        xx = np.power(float(n + 1), u)
    else:
        xx = np.power((np.power(float(n + 1), 1.0 - s) - 1.0) * u + 1.0, 1.0 / (1.0 - s))
    rank = np.minimum(np.floor(xx).astype(np.int64) - 1, n - 1)
    rank = np.maximum(rank, 0)
    if not scatter:
        return rank
    a = _odd_multiplier(n)
    return (rank * a) % n


def _odd_multiplier(n: int) -> int:
    a = 2654435761 % n
    a |= 1
    while np.gcd(a, n) != 1:
        a += 2
    return int(a)


@dataclass
class Workload:
    name: str
    field_sizes: List[int]
    d: int
    batch: int
    x: np.ndarray                 # int64 [R, F] global row ids
    y: np.ndarray                 # float32 [R]
    n_train: int
    output: str                   # "reg" | "class"
    variant: str                  # "sampled" | "closed"
    interaction: str = "prod"
    field_offsets: List[int] = field(default_factory=list)

    @property
    def rows(self) -> int:
        return int(sum(self.field_sizes))

    @property
    def n_fields(self) -> int:
        return len(self.field_sizes)

    def train_counts(self) -> np.ndarray:
        """``bincount(X_train.flatten())`` with ``minlength=rows`` (the reference
        omits minlength, vfm-torch.py:89 -- SURVEY N9)."""
        return np.bincount(self.x[: self.n_train].reshape(-1), minlength=self.rows).astype(np.int64)

    def batches(self):
        for lo in range(0, self.n_train, self.batch):       # contiguous, never shuffled (N8)
            hi = min(lo + self.batch, self.n_train)
            yield self.x[lo:hi], self.y[lo:hi]


def _ratings(rng, size):
    return np.clip(np.round(3.5 + rng.standard_normal(size)), 1, 5).astype(np.float32)


def make_ids(field_sizes: Sequence[int], exponents: Sequence[float], n_rows: int,
             seed: int = DATA_SEED) -> np.ndarray:
    rng = np.random.default_rng(seed)
    offs = np.concatenate(([0], np.cumsum(field_sizes)[:-1]))
    cols = [offs[f] + zipf_ids(int(n), n_rows, float(s), rng)
            for f, (n, s) in enumerate(zip(field_sizes, exponents))]
    return np.stack(cols, axis=1).astype(np.int64)


def _fraction_workload() -> Workload:
    """BASELINE config 1: the reference's bundled fixture ``data/fraction/data.csv`` (536 students x 20
    items, binary outcome; SURVEY N13: X = [user, 536 + item], seeded 80/20 split, full batch).  The
    training split travels as the inputs of the committed golden ``tests/golden/sampled_fraction.npz``
    (written by ``oracle/gen_golden.py`` from the reference's file), so nothing outside the repo is read."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                        "sampled_fraction.npz")
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    x, y = z["x"].astype(np.int64), z["y"].astype(np.float32)
    return Workload("fraction", [meta["N"], meta["M"]], meta["d"], len(x), x, y, len(x), "class", "sampled",
                    "prod", [0, meta["N"]])


def make_workload(name: str, n_rows: int | None = None, seed: int = DATA_SEED) -> Workload:
    """Named BASELINE.json configurations (sizes in SURVEY.md section 8a/8d)."""
    rng_y = np.random.default_rng(seed + 1)
    if name == "fraction":          # config 1 (real fixture, full batch)
        return _fraction_workload()
    if name == "ml100k":            # config 2
        fs, ex, d, B, R, frac = [943, 1682], [0.5, 1.0], 20, 8000, 100_000, 0.8
        variant, output, inter = "closed", "reg", "prod"
    elif name == "ml20m":           # config 3
        fs, ex, d, B, R, frac = [138_493, 26_744], [0.5, 1.0], 64, 65_536, 20_000_263, 1.0
        variant, output, inter = "sampled", "reg", "prod"
    elif name == "sideinfo":        # config 4
        fs = [500_000, 300_000, 100_000, 50_000, 30_000, 15_000, 4_000, 1_000]
        ex, d, B, R, frac = [0.5, 1, 1, 1, 1, 1, 1, 0.5], 64, 65_536, 4_000_000, 1.0
        variant, output, inter = "sampled", "class", "pairwise"
    elif name == "big100m":         # config 5
        fs, ex, d, B, R, frac = [70_000_000, 30_000_000], [1.05, 1.05], 128, 65_536, 4_000_000, 1.0
        variant, output, inter = "sampled", "reg", "prod"
    else:
        raise ValueError(name)
    if n_rows is not None:
        R = n_rows
    x = make_ids(fs, ex, R, seed)
    y = (_ratings(rng_y, R) if output == "reg"
         else (rng_y.random(R) < 0.5).astype(np.float32))
    offs = np.concatenate(([0], np.cumsum(fs)[:-1])).tolist()
    return Workload(name, fs, d, B, x, y, int(R * frac), output, variant, inter, offs)
