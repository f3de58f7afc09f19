"""Where does the host time of a train step go?  Wall time of the whole Python train_step against the time inside the
C-ABI call (74 launches) alone."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import iwae_b200 as ib
from iwae_b200 import _abi, model as M
P, X, _ = ib.data.synthetic_binarised(400)
m = ib.IWAE2([200, 100], [100, 50], out_bias=ib.data.get_bias(P.mean(0)))
opt = ib.Adam(1e-3, epsilon=1e-4)
xb = [torch.from_numpy(X[i * 20:(i + 1) * 20]).cuda() for i in range(20)]
orig = M.lib.iwae_train_step
acc = [0.0, 0]
class Wrap:
    def __call__(self, *a):
        t0 = time.perf_counter(); r = orig(*a); acc[0] += time.perf_counter() - t0; acc[1] += 1; return r
class LibProxy:
    def __init__(self, lib): self._lib = lib; self.iwae_train_step = Wrap()
    def __getattr__(self, k): return getattr(self._lib, k)
M.lib = LibProxy(M.lib)
n = 1000
for i in range(10): m.train_step(xb[i], 50, 1.0, opt, objective="iwae_elbo")
torch.cuda.synchronize(); acc[0] = 0.0; acc[1] = 0; t0 = time.perf_counter()
for i in range(n): m.train_step(xb[i % 20], 50, 1.0, opt, objective="iwae_elbo")
host = time.perf_counter() - t0
torch.cuda.synchronize(); wall = time.perf_counter() - t0
print(f"per step: wall {1e3*wall/n:.3f} ms, host loop {1e3*host/n:.3f} ms, inside iwae_train_step {1e3*acc[0]/acc[1]:.3f} ms")
