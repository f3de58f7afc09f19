"""Short train_step loop for the ncu launch list (k=50, batch 20, 2 stochastic layers)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import iwae_b200 as ib
P, X, _ = ib.data.synthetic_binarised(400)
m = ib.IWAE2([200, 100], [100, 50], out_bias=ib.data.get_bias(P.mean(0)))
opt = ib.Adam(1e-3, epsilon=1e-4)
xb = [torch.from_numpy(X[i * 20:(i + 1) * 20]).cuda() for i in range(20)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for i in range(5): m.train_step(xb[i], 50, 1.0, opt, objective="iwae_elbo")
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(n): res = m.train_step(xb[i % 20], 50, 1.0, opt, objective="iwae_elbo")
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"{n} steps, {1e3*dt/n:.3f} ms/step wall, iwae_elbo {float(res['iwae_elbo']):.3f}")
