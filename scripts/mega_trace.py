"""Per-level timeline of the train-step megakernel (IWAE_MEGA_TRACE): eager steps, the last one's trace is kept."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["IWAE_MEGA"] = "1"
os.environ["IWAE_NO_GRAPH"] = "1"
os.environ.setdefault("IWAE_MEGA_TRACE", "gpurun_out/mega_trace.txt")
import numpy as np, torch
import iwae_b200 as ib
P, X, _ = ib.data.synthetic_binarised(400)
m = ib.IWAE2([200, 100], [100, 50], out_bias=ib.data.get_bias(P.mean(0)))
opt = ib.Adam(1e-3, epsilon=1e-4)
xb = [torch.from_numpy(X[i * 20:(i + 1) * 20]).cuda() for i in range(20)]
for i in range(8): res = m.train_step(xb[i], 50, 1.0, opt, objective="iwae_elbo")
torch.cuda.synchronize()
print(open(os.environ["IWAE_MEGA_TRACE"]).read())
