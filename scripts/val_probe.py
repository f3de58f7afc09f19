"""val_step(Xtest [10000,784], k=50) - the monitoring call of main.py:152 - on the fused engine vs the fp32 kernels."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import iwae_b200 as ib
P, X, _ = ib.data.synthetic_binarised(10000)
bias = ib.data.get_bias(P.mean(0))
x = torch.from_numpy(X).cuda()
for name, m in (("iwae2", ib.IWAE2([200, 100], [100, 50], out_bias=bias)), ("iwae1", ib.IWAE(200, 100, out_bias=bias))):
    for prec in ("auto", "fp32"):
        for _ in range(3):
            r = m.val_step(x, 50, 1.0, precision=prec)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            r = m.val_step(x, 50, 1.0, precision=prec)
        e1.record(); torch.cuda.synchronize()
        print(f"{name} val_step(10000 images, k=50) precision={prec}: {e0.elapsed_time(e1)/10:.3f} ms  iwae_elbo {float(r['iwae_elbo']):.4f} keys {len(r)}")
