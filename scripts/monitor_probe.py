"""Monitoring pass on the full test set (10,000 images, k=50): fused kernel vs fp32 val_step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import iwae_b200 as ib
P, X, _ = ib.data.synthetic_binarised(10000)
m = ib.IWAE2([200, 100], [100, 50], out_bias=ib.data.get_bias(P.mean(0)))
x = torch.from_numpy(X).cuda()
for name, fn in (("monitor (fused tcgen05)", lambda: m.monitor(x, 50)), ("val_step (fp32 kernels)", lambda: m.val_step(x, 50, 1.0))):
    r = fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): r = fn()
    float(r["iwae_elbo"]); torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter()-t0)/3*1e3:.2f} ms  iwae_elbo {float(r['iwae_elbo']):.3f} vae_elbo {float(r['vae_elbo']):.3f}")
