"""Short driver for ncu: a few fused-evaluation launches on a reduced image count."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import iwae_b200 as ib

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=592)
ap.add_argument("--k", type=int, default=5000)
ap.add_argument("--layers", type=int, default=2)
ap.add_argument("--precision", default="tc_f16_fast")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
P, X, _ = ib.data.synthetic_binarised(a.images)
bias = ib.data.get_bias(P.mean(0))
m = ib.IWAE2([200, 100], [100, 50], out_bias=bias) if a.layers == 2 else ib.IWAE(200, 100, out_bias=bias)
x = torch.from_numpy(X).cuda()
for i in range(a.reps):
    per, tot = m.llh(x, a.k, precision=a.precision, seed=1)
    torch.cuda.synchronize()
    print(i, float(tot) / a.images, m.last_eval_kernel_ms(), "ms")
