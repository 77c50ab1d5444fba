"""A handful of fused training steps on a BASELINE workload -- the command profiled under ncu."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                   # noqa: E402
from vae_b200 import synth                                      # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="ml20m")
ap.add_argument("--rows", type=int, default=2_000_000)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--warmup", type=int, default=4)
ap.add_argument("--reserve", type=int, default=0,
                help="1: the kernels of the default (graphed, plan-overlapped) bench: unfused score/gather, "
                     "one block slot per SM left free")
ap.add_argument("--tune", action="append", default=[], metavar="KEY=VALUE")
a = ap.parse_args()
from vae_b200 import _lib as L                                  # noqa: E402
L.check(L.lib().vfmb_set_grid_reserve(a.reserve))
for kv in a.tune:
    k, v = kv.split("=")
    L.check(L.lib().vfmb_set_tuning(k.encode(), int(v)))
w = synth.make_workload(a.workload, n_rows=a.rows)
model = bench.make_model(w, torch.device("cuda", 0), w.train_counts(), 1.0 / (1 + w.n_train // w.batch))
B = w.batch
x = torch.from_numpy(w.x).cuda()
y = torch.from_numpy(w.y).cuda()
nb = w.n_train // B
for i in range(a.warmup + a.steps):
    j = i % nb
    out = model.fused_step(x[j * B:(j + 1) * B], y[j * B:(j + 1) * B])
torch.cuda.synchronize()
print("loss", out["loss"].item(), "U", int(model._plan.meta[0]), "W", int(model._plan.meta[1]))
