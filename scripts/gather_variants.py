"""Backward-gather launch knobs side by side (run under an ncu launch list restricted to k_gather):
for every (workload, gather_dyn, gather_wide) four fused steps, in this order."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                   # noqa: E402
from vae_b200 import synth                                      # noqa: E402
from vae_b200 import _lib as L                                  # noqa: E402

VARIANTS = [(0, 0), (0, 1), (1, 0), (1, 1)]      # (gather_dyn, gather_wide)
L.check(L.lib().vfmb_set_grid_reserve(1))
for name, rows in (("sideinfo", 1_000_000), ("ml20m", 2_000_000)):
    w = synth.make_workload(name, n_rows=rows)
    model = bench.make_model(w, torch.device("cuda", 0), w.train_counts(), 1.0 / (1 + w.n_train // w.batch))
    B = w.batch
    x = torch.from_numpy(w.x).cuda()
    y = torch.from_numpy(w.y).cuda()
    nb = w.n_train // B
    i = 0
    for dyn, wide in VARIANTS:
        for k, v in (("gather_dyn", dyn), ("gather_wide", wide)):
            L.check(L.lib().vfmb_set_tuning(k.encode(), v))
        for _ in range(4):
            j = i % nb
            i += 1
            out = model.fused_step(x[j * B:(j + 1) * B], y[j * B:(j + 1) * B])
        torch.cuda.synchronize()
        print(name, dyn, wide, "loss", out["loss"].item(), flush=True)
