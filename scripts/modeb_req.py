import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_b200 import synth
from vae_b200.dist import ShardedSampled, bucket_by_owner
dev = torch.device("cuda", 0)
w = synth.make_workload("ml20m", n_rows=2_000_000)
B, d = w.batch, w.d
tc = w.train_counts(); tc[tc == 0] = 1
r = ShardedSampled(d, w.field_sizes, torch.from_numpy(tc), w.n_train, B, 2, 0, output="reg", lr=1e-3, device=dev, exchange=object(), slack=0.75)
x = torch.from_numpy(w.x).to(dev)
def T(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{name:20s} host {1e3*(t1-t0):7.3f} ms  total {1e3*(t2-t0):7.3f} ms"); return out
for it in range(3):
    xb = x[it*B:(it+1)*B].contiguous()
    T("plan.build", lambda: r.plan_l.build(r.cfg_l, xb, r.train_counts))
    pl = r.plan_l
    T("bucket", lambda: bucket_by_owner(pl.uniq, pl.urec.view(-1, 4)[:, 1], pl.meta[0], 2, r.CAP))
    ids = pl.uniq; n_valid = pl.meta[0]
    valid = T(" arange<", lambda: torch.arange(ids.numel(), device=dev) < n_valid)
    owner = T(" where", lambda: torch.where(valid, ids % 2, torch.zeros_like(ids)).long())
    onehot = T(" onehot", lambda: (owner[None, :] == torch.arange(2, device=dev)[:, None]) & valid[None, :])
    ordinal = T(" cumsum", lambda: (onehot.cumsum(1, dtype=torch.int32) - 1).gather(0, owner[None, :]).squeeze(0).long())
    dest = T(" dest", lambda: torch.where(valid & (ordinal < r.CAP), owner * r.CAP + ordinal, torch.full_like(owner, r.M)))
    send = torch.full((r.M + 1, 2), -1, dtype=torch.int32, device=dev)
    T(" index_put", lambda: send.__setitem__(dest, torch.stack((ids.int(), pl.urec.view(-1, 4)[:, 1].int()), dim=1)))
y = torch.from_numpy(w.y).to(dev)
for it in range(3):
    xb = x[it*B:(it+1)*B]; yb = y[it*B:(it+1)*B]
    T("phase_request", lambda: r.phase_request(xb, yb))
    T(" y.to", lambda: yb.to(dev, torch.float32).contiguous())
    T(" x.to", lambda: xb.to(dev).contiguous())
    T(" z.clone", lambda: r.plan_l.z.clone())
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); r.phase_request(xb, yb); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
