"""Time k_adam_rows (library events around the launch) at controlled row densities.

Every batch touches each `stride`-th row of both fields exactly once, so the kernel's access
pattern is a monotone sweep over the tables at density 1/stride."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_b200 import _lib as L
from vae_b200.vfm_torch import CF

ap = argparse.ArgumentParser()
ap.add_argument("--d", type=int, default=64)
ap.add_argument("--B", type=int, default=32768)
a = ap.parse_args()
dev = torch.device("cuda", 0)
B, d = a.B, a.d
for stride in (1, 2, 3, 4, 8):
    N = M = B * stride
    tc = torch.ones(N + M)
    torch.manual_seed(0)
    model = CF(d, output="reg", n_users=N, n_items=M, train_counts=tc, n_train=B * 100, max_batch=B, lr=1e-3, device=dev)
    rng = np.random.default_rng(0)
    times = []
    for it in range(12):
        off = it % stride
        u = np.arange(B) * stride + off
        i = N + rng.permutation(B) * stride + off
        x = torch.from_numpy(np.stack([u, i], 1).astype(np.int64)).to(dev)
        y = torch.randn(B, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); e1.record()
        L.lib().vfmb_profile_events(e0.cuda_event, e1.cuda_event)
        model.fused_step(x, y)
        torch.cuda.synchronize()
        if it >= 4:
            times.append(e0.elapsed_time(e1))
    L.lib().vfmb_profile_events(None, None)
    t = float(np.median(times)) * 1e-3
    U = 2 * B
    by = U * (2 * d + 2) * 4 * 6
    print(f"stride {stride}: U={U} adam {t*1e6:.1f} us  {by/t/1e9:.0f} GB/s  (variant {os.environ.get('VFMB_VARIANT','default')})", flush=True)
    del model

# streaming reference: dense Adam over the same number of bytes (p, m, v read+write, g read)
import ctypes as C
n = 2 * a.B * 2 * d
p_, m_, v_, g_ = (torch.randn(n, device=dev) for _ in range(4))
v_.abs_()
adam = L.Adam(1e-3, 0.9, 0.999, 1e-8)
stepc = torch.zeros(1, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
ts = []
for it in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(L.lib().vfmb_adam_dense(p_.data_ptr(), m_.data_ptr(), v_.data_ptr(), g_.data_ptr(), n, C.byref(adam), stepc.data_ptr(), s))
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = float(np.median(ts[3:])) * 1e-3
print(f"dense adam {n} floats: {t*1e6:.1f} us  {n*4*7/t/1e9:.0f} GB/s (7 streams)")
