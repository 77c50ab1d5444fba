"""Where the graphed step's time outside its four kernels goes: the same loop with and without the plan
of the next batch on the side branch (the latter replays stale plans: timing only)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                   # noqa: E402
from vae_b200 import synth                                      # noqa: E402
from vae_b200 import _lib as L                                  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ml20m"
steps = 1000
w = synth.make_workload(name, n_rows=2_000_000 if name == "ml20m" else 1_000_000)
B = w.batch
x = torch.from_numpy(w.x).cuda()
y = torch.from_numpy(w.y).cuda()
nb = w.n_train // B
# (tried and dropped: forking the plan after the gradient gather, next to the row update only -- the plan's
#  nine-kernel chain then outlasts the update: 126.9 vs 111.2 us per ml20m step)
for label, kw in (("graph + plan", {}), ("graph, no plan", {"plan_in_graph": False})):
    model = bench.make_model(w, torch.device("cuda", 0), w.train_counts(), 1.0 / (1 + w.n_train // w.batch))
    loop = model.graphed_loop(B, depth=2, **kw)
    loop.start(x[:B], y[:B])
    if not kw.get("plan_in_graph", True):                      # leave a valid plan in every slot
        for s in range(loop.depth):
            loop.xs[s].copy_(x[s * B:(s + 1) * B])
            loop.plans[s].build(loop.cfg, loop.xs[s], model.train_counts)
    for i in range(1, 30):
        j = i % nb
        loop.step(x[j * B:(j + 1) * B], y[j * B:(j + 1) * B])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        j = (30 + i) % nb
        loop.step(x[j * B:(j + 1) * B], y[j * B:(j + 1) * B])
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {label:28s} {e0.elapsed_time(e1) / steps * 1e3:7.1f} us/step", flush=True)
