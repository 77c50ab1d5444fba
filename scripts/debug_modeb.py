import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import golden_util as gu
import test_gpu_sampled as S
from vae_b200.dist import ShardedSampled
DEV = "cuda"
meta, g = gu.load("sampled_reg_d64")
N, M, d = meta["N"], meta["M"], meta["d"]; R = N + M; P = 2
x, y = gu.batch_of(meta, g, 0)
n = (len(x) // P) * P
gen = torch.Generator().manual_seed(3)
e0 = torch.randn(1, generator=gen).to(DEV)
eb_t, ee_t = torch.randn(R, generator=gen).to(DEV), torch.randn(R, d, generator=gen).to(DEV)
init = gu.state(g, "init")
ref = S._model(meta, g, 0)
uniq = torch.from_numpy(np.unique(x)).to(DEV)
out_ref = ref.fused_step(torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV),
                         noise=(e0.reshape(1, 1), eb_t[uniq][None], ee_t[uniq][None]), update=False)
U = len(uniq)
ref_vs = ref._buf.vs[: U * d].view(U, d).clone(); ref_ws = ref._buf.ws[:U].clone()
ref_pred = out_ref["pred"].clone(); ref_loss = out_ref["loss"].item()
ref_cq = ref._buf.cq[:U].clone()
gr = ref.gradients(torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV), noise=(e0.reshape(1, 1), eb_t[uniq][None], ee_t[uniq][None]))
ref_grow = ref._buf.grow[: U * d].view(U, d).clone(); ref_gws = ref._buf.gws[:U].clone()
ini = {"bias": torch.from_numpy(init["bias_params.weight"]), "entity": torch.from_numpy(init["entity_params.weight"]),
       "alpha": init["alpha"][0], "global_bias_mean": init["global_bias_mean"][0], "global_bias_scale": init["global_bias_scale"][0]}
ranks = [ShardedSampled(d, [N, M], torch.from_numpy(g["train_counts"]), meta["n_train"], n // P, P, p, output=meta["output"],
                        link=meta["link"], lr=meta["lr"], init=ini, noise_tables=(e0, eb_t, ee_t), exchange=object()) for p in range(P)]
xs = [torch.from_numpy(x[p * (n // P):(p + 1) * (n // P)]).to(DEV) for p in range(P)]
ys = [torch.from_numpy(y[p * (n // P):(p + 1) * (n // P)]).to(DEV) for p in range(P)]
req = [r.phase_request(a, b) for r, a, b in zip(ranks, xs, ys)]
z = sum(q[1] for q in req)
print("z", z[:2].tolist(), "ref z", ref._plan.z[:2].tolist())
a2a = lambda bufs: [torch.stack([bufs[src][dst] for src in range(P)]).contiguous() for dst in range(P)]
recv = a2a([q[0] for q in req])
replies = [r.phase_owner_stage(rv, z) for r, rv in zip(ranks, recv)]
for p, r in enumerate(ranks):
    Uo = int(r.plan_o.meta[0]); uo = r.plan_o.uniq[:Uo].long()
    gid = uo * P + p
    ok = gid < R
    pos = torch.searchsorted(uniq, gid[ok])
    hit = uniq[pos.clamp(max=U - 1)] == gid[ok]
    vs_o = r.buf_o.vs[: Uo * d].view(Uo, d)[ok][hit]
    print("rank", p, "Uo", Uo, "owned real rows", int(hit.sum()), "vs diff", (vs_o - ref_vs[pos[hit]]).abs().max().item(),
          "cq diff", (r.buf_o.cq[:Uo][ok][hit] - ref_cq[pos[hit]]).abs().max().item(), "kl rows", r.buf_o.stats[5].item())
print("ref kl rows", ref._buf.stats[5].item())
rows = a2a(replies)
loc = [r.phase_local(rw) for r, rw in zip(ranks, rows)]
pred = torch.cat([r.buf_l.mean[: r.B] for r in ranks])
print("pred diff", (pred - ref_pred).abs().max().item())
for p, r in enumerate(ranks):
    Ul = int(r.plan_l.meta[0]); ul = r.plan_l.uniq[:Ul].long()
    print("rank", p, "local U", Ul)
tail = sum(t[1].clone() for t in loc)
print("tail", tail[8:12].tolist(), "ref nll_sum", ref._buf.stats[1].item() * n, "resid", ref._buf.stats[3].item())
grads = a2a([t[0] for t in loc])
outs = [r.phase_owner_update(g_, tail.clone()) for r, g_ in zip(ranks, grads)]
for p, r in enumerate(ranks):
    Uo = int(r.plan_o.meta[0]); uo = r.plan_o.uniq[:Uo].long(); gid = uo * P + p
    ok = gid < R
    pos = torch.searchsorted(uniq, gid[ok]); hit = uniq[pos.clamp(max=U - 1)] == gid[ok]
    go = r.buf_o.grow[: Uo * d].view(Uo, d)[ok][hit]
    print("rank", p, "grow diff", (go - ref_grow[pos[hit]]).abs().max().item(), "scale", ref_grow.abs().max().item(),
          "gws diff", (r.buf_o.gws[:Uo][ok][hit] - ref_gws[pos[hit]]).abs().max().item())
print("loss", [o["loss"].item() for o in outs], "ref", ref_loss)
