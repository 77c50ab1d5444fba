"""Host-side cost of one step (launch-bound check): wall time of the Python loop alone vs GPU time."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vae_b200 import synth
w = synth.make_workload("ml20m", n_rows=4_000_000)
model = bench.make_model(w, torch.device("cuda", 0), w.train_counts(), 0.01)
B = w.batch
x = torch.from_numpy(w.x).cuda(); y = torch.from_numpy(w.y).cuda()
nb = w.n_train // B
def batch(i):
    j = i % nb
    return x[j*B:(j+1)*B], y[j*B:(j+1)*B]
for mode in ("inline", "prefetch", "cached"):
    plans = {j: model.static_plan(batch(j)[0]) for j in range(nb)} if mode == "cached" else None
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 1000
        if mode == "prefetch": model.prefetch_plan(batch(0)[0])
        for i in range(K):
            xb, yb = batch(i)
            if mode == "prefetch": model.prefetch_plan(batch(i+1)[0])
            if mode == "cached": model.fused_step(xb, yb, plan=plans[i % nb])
            else: model.fused_step(xb, yb)
        t1 = time.perf_counter()
        e1.record(); torch.cuda.synchronize()
        t2 = time.perf_counter()
    print(f"{mode:9s} cpu-issue {1e6*(t1-t0)/K:7.1f} us/step   gpu {1e3*e0.elapsed_time(e1)/K:7.1f} us/step   wall {1e6*(t2-t0)/K:7.1f}")
