"""A few val_step(Xtest [10000,784], k=50) calls on the fused engine, for the ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import iwae_b200 as ib
P, X, _ = ib.data.synthetic_binarised(10000)
x = torch.from_numpy(X).cuda()
m = ib.IWAE2([200, 100], [100, 50], out_bias=ib.data.get_bias(P.mean(0)))
for _ in range(4):
    r = m.val_step(x, 50, 1.0, precision="auto")
torch.cuda.synchronize()
print(float(r["iwae_elbo"]))
