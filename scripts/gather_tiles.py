"""Per-block-tile timing of the backward gather (measurement build: VFMB_VARIANT=tt VFMB_NVCC_EXTRA=-DVFMB_TILE_TIMING)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                   # noqa: E402
from vae_b200 import synth                                      # noqa: E402
from vae_b200 import _lib as L                                  # noqa: E402

L.check(L.lib().vfmb_set_grid_reserve(1))
for name, rows in (("sideinfo", 1_000_000), ("ml20m", 2_000_000)):
    w = synth.make_workload(name, n_rows=rows)
    model = bench.make_model(w, torch.device("cuda", 0), w.train_counts(), 1.0 / (1 + w.n_train // w.batch))
    B = w.batch
    x = torch.from_numpy(w.x).cuda()
    y = torch.from_numpy(w.y).cuda()
    nt = (B * len(w.field_sizes) + 511) // 512
    for dyn, wide in ((0, 0), (0, 1), (1, 1)):
        for k, v in (("gather_dyn", dyn), ("gather_wide", wide)):
            L.check(L.lib().vfmb_set_tuning(k.encode(), v))
        for i in range(5):
            model.fused_step(x[i * B:(i + 1) * B], y[i * B:(i + 1) * B])
        torch.cuda.synchronize()
        buf = np.zeros((nt, 4), np.int32)
        rc = L.lib().vfmb_debug_tile_times(buf.ctypes.data_as(C.c_void_p), nt)
        main, fin, sm, t0 = buf[:, 0].astype(np.int64), buf[:, 1].astype(np.int64), buf[:, 2], buf[:, 3].astype(np.uint32).astype(np.int64)
        t0 = t0 - t0.min()
        end = t0 + (main + fin) / 1.9
        print(f"{name} dyn={dyn} wide={wide} rc={rc} tiles={nt}")
        print(f"  main cycles: mean {main.mean():.0f} p50 {np.median(main):.0f} p90 {np.percentile(main, 90):.0f} p99 {np.percentile(main, 99):.0f} max {main.max()}")
        print(f"  fin  cycles: mean {fin.mean():.0f} p50 {np.median(fin):.0f} p90 {np.percentile(fin, 90):.0f} p99 {np.percentile(fin, 99):.0f} max {fin.max()}")
        print(f"  span: last start {t0.max() / 1e3:.1f} us, last end {end.max() / 1e3:.1f} us")
        q = np.linspace(0, nt, 9).astype(int)
        for a, b in zip(q[:-1], q[1:]):
            print(f"    tiles {a:6d}-{b:6d}: main {main[a:b].mean():8.0f} fin {fin[a:b].mean():8.0f} start {t0[a:b].mean() / 1e3:6.1f} us  end max {end[a:b].max() / 1e3:6.1f} us")
        persm = np.zeros(160)
        np.maximum.at(persm, sm, end)
        busy = persm[persm > 0]
        print(f"  per-SM last end: min {busy.min() / 1e3:.1f} mean {busy.mean() / 1e3:.1f} max {busy.max() / 1e3:.1f} us")
