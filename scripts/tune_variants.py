"""Launch knobs side by side (run under an ncu launch list): for every workload and every knob setting
four fused steps, in this order.  Usage: tune_variants.py key=v[,key=v...] [key=v...] ..."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                   # noqa: E402
from vae_b200 import synth                                      # noqa: E402
from vae_b200 import _lib as L                                  # noqa: E402

VARIANTS = [dict(kv.split("=") for kv in a.split(",")) for a in sys.argv[1:]]
L.check(L.lib().vfmb_set_grid_reserve(1))
for name, rows in (("sideinfo", 1_000_000), ("ml20m", 2_000_000)):
    w = synth.make_workload(name, n_rows=rows)
    B = w.batch
    x = torch.from_numpy(w.x).cuda()
    y = torch.from_numpy(w.y).cuda()
    nb = w.n_train // B
    for var in VARIANTS:
        for k, v in var.items():
            L.check(L.lib().vfmb_set_tuning(k.encode(), int(v)))
        model = bench.make_model(w, torch.device("cuda", 0), w.train_counts(), 1.0 / (1 + w.n_train // w.batch))
        for i in range(4):
            j = i % nb
            out = model.fused_step(x[j * B:(j + 1) * B], y[j * B:(j + 1) * B])
        torch.cuda.synchronize()
        print(name, var, "loss", out["loss"].item(), flush=True)
