import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import golden_util as gu
import test_gpu_closed as T
meta, g = gu.load("closed_ml100k")
x, y = gu.batch_of(meta, g, 0)
xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
outs = []
for r in range(4):
    m = T._model(meta, g, 0)
    gr = m.gradients(xd, yd)
    buf = m._buf
    snap = {k: v.clone() for k, v in gr.items()}
    snap["grow"] = buf.grow.clone(); snap["gws"] = buf.gws.clone(); snap["cq"] = buf.cq.clone()
    snap["vs"] = buf.vs.clone(); snap["stats"] = buf.stats.clone(); snap["rsorted"] = buf.rsorted.clone()
    U = int(m._plan.meta[0])
    snap["grow"] = snap["grow"][: U * 3 * meta["d"]]; snap["gws"] = snap["gws"][:U]; snap["cq"] = snap["cq"][:U]
    snap["vs"] = snap["vs"][: U * 2 * meta["d"]]
    outs.append(snap)
for k in outs[0]:
    for r in range(1, 4):
        a, b = outs[0][k], outs[r][k]
        if not torch.equal(a, b):
            d = (a - b).abs()
            print("DIFF", k, "run", r, "max", d.max().item(), "n", int((d > 0).sum()), "where", torch.nonzero(d.reshape(-1) > 0)[:5].reshape(-1).tolist())
print("done")
