"""Staged GPU smoke test with progress prints (debug aid for gpurun sessions)."""
import faulthandler
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
faulthandler.dump_traceback_later(150, exit=True)
t0 = time.time()


def say(*a):
    print(f"[{time.time() - t0:6.1f}s]", *a, flush=True)


say("import torch")
import torch
say("torch", torch.__version__, torch.cuda.get_device_name(0))
from vae_b200 import _lib as L
say("stale?", L._stale())
L.lib()
say("lib loaded")
from vae_b200.engine import BatchPlan, make_config
B, F, R = 1000, 2, 50
x = torch.randint(0, R, (B, F))
plan = BatchPlan(B, F, R, "cuda")
cfg = make_config(B, F, 8, R, 1, "reg", "abs", [26], [25, 25], 100, 7)
say("plan.build")
plan.build(cfg, x.cuda(), torch.ones(R, device="cuda"))
torch.cuda.synchronize()
say("plan built U=", plan.num_unique(), "W=", int(plan.meta[1]))
u, i, c = torch.unique(x, return_inverse=True, return_counts=True)
pu, pi, pc = plan.as_unique()
say("plan ok:", torch.equal(pu.cpu(), u), torch.equal(pi.cpu(), i), torch.equal(pc.cpu(), c))
import golden_util as gu
meta, g = gu.load("sampled_reg_d64")
from vae_b200.vfm_torch import CF
m = CF(meta["d"], output=meta["output"], n_users=meta["N"], n_items=meta["M"],
       train_counts=torch.from_numpy(g["train_counts"]), n_train=meta["n_train"], max_batch=meta["batch"],
       lr=meta["lr"])
own = m.state_dict()
m.load_state_dict({k: torch.from_numpy(v).reshape(own[k].shape) for k, v in gu.state(g, "init").items() if k in own}, strict=False)
xx, yy = gu.batch_of(meta, g, 0)
noise = [torch.from_numpy(g[f"step0.noise{i}"]).cuda() for i in range(3)]
say("forward")
out = m.fused_step(torch.from_numpy(xx).cuda(), torch.from_numpy(yy).cuda(), noise=noise, update=False)
torch.cuda.synchronize()
say("loss", out["loss"].item(), "golden", g["step0.loss"], "pred err",
    float((out["pred"].cpu() - torch.from_numpy(g["step0.pred"])).abs().max()))
say("gradients")
gr = m.gradients(torch.from_numpy(xx).cuda(), torch.from_numpy(yy).cuda(), noise=noise)
torch.cuda.synchronize()
say("grad err", gu.rel_err(gr["entity_params.weight"].cpu().numpy(), g["step0.grad.entity_params.weight"]),
    gu.rel_err(gr["bias_params.weight"].cpu().numpy(), g["step0.grad.bias_params.weight"]))
say("fused step")
out = m.fused_step(torch.from_numpy(xx).cuda(), torch.from_numpy(yy).cuda(), noise=noise)
torch.cuda.synchronize()
after = gu.state(g, "step0.after")
say("param err", float((m.entity_params.weight.cpu() - torch.from_numpy(after["entity_params.weight"])).abs().max()),
    "step", int(m.adam_step))
say("done")
