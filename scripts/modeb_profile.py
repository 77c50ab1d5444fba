"""Phase timing of mode B (row-sharded step) with P ranks emulated on one GPU."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_b200 import synth
from vae_b200.dist import ShardedSampled

P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
NIT = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda", 0)
w = synth.make_workload("ml20m", n_rows=2_000_000)
B, d = w.batch, w.d
tc = w.train_counts(); tc[tc == 0] = 1
ranks = [ShardedSampled(d, w.field_sizes, torch.from_numpy(tc), w.n_train, B, P, p, output="reg", lr=1e-3,
                        device=dev, exchange=object(), slack=0.75) for p in range(P)]
x = torch.from_numpy(w.x).to(dev); y = torch.from_numpy(w.y).to(dev)
a2a = lambda bufs: [torch.stack([bufs[src][dst] for src in range(P)]).contiguous() for dst in range(P)]

def timed(name, fn, acc):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = fn(); e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    acc.setdefault(name, []).append((e0.elapsed_time(e1), (t1 - t0) * 1e3))
    return out

acc = {}
nb = x.shape[0] // B
import cProfile, pstats
for it in range(NIT):
    xs = [x[((it * P + p) % nb) * B:((it * P + p) % nb + 1) * B] for p in range(P)]
    ys = [y[((it * P + p) % nb) * B:((it * P + p) % nb + 1) * B] for p in range(P)]
    if it == 5 and NIT == 6:
        torch.cuda.synchronize()
        pr = cProfile.Profile(); pr.enable(); ranks[0].phase_request(xs[0], ys[0]); pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(8)
    req = [timed("request", lambda r=r, a=a, b=b: r.phase_request(a, b), acc) for r, a, b in zip(ranks, xs, ys)]
    z = sum(q[1] for q in req)
    recv = a2a([q[0] for q in req])
    replies = [timed("owner_stage", lambda r=r, rv=rv: r.phase_owner_stage(rv, z), acc) for r, rv in zip(ranks, recv)]
    rows = a2a(replies)
    loc = [timed("local", lambda r=r, rw=rw: r.phase_local(rw), acc) for r, rw in zip(ranks, rows)]
    tail = sum(t[1].clone() for t in loc)
    grads = a2a([t[0] for t in loc])
    outs = [timed("owner_update", lambda r=r, g=g: r.phase_owner_update(g, tail.clone()), acc) for r, g in zip(ranks, grads)]
    if NIT > 6 and it % 5 == 0:
        print("it", it, "loss", outs[0]["loss"].item(), "kl", outs[0]["kl"].item(), "scalars", ranks[0].scalars.tolist(),
              "finite params", all(bool(torch.isfinite(r.entity).all()) for r in ranks), flush=True)
for k, v in acc.items():
    v = np.array(v[P * 2:])
    print(f"{k:14s} gpu {v[:, 0].mean():8.3f} ms   host-enqueue {v[:, 1].mean():8.3f} ms (per rank)")
print("loss", outs[0]["loss"].item())
