"""Is the k=50 train step launch-bound?  Times (a) the normal eager train_step loop, (b) the same step captured
once into a CUDA graph and replayed (identical kernels, no per-launch host cost) and (c) the GPU-busy time of the
eager loop measured with CUDA events around each step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import iwae_b200 as ib
P, X, _ = ib.data.synthetic_binarised(400)
m = ib.IWAE2([200, 100], [100, 50], out_bias=ib.data.get_bias(P.mean(0)))
opt = ib.Adam(1e-3, epsilon=1e-4)
xb = [torch.from_numpy(X[i * 20:(i + 1) * 20]).cuda() for i in range(20)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
for i in range(10): m.train_step(xb[i], 50, 1.0, opt, objective="iwae_elbo")
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(n): m.train_step(xb[i % 20], 50, 1.0, opt, objective="iwae_elbo")
torch.cuda.synchronize(); eager = (time.perf_counter() - t0) / n
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    m.train_step(xb[0], 50, 1.0, opt, objective="iwae_elbo")
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        m.train_step(xb[0], 50, 1.0, opt, objective="iwae_elbo")
    torch.cuda.synchronize()
    for _ in range(10): g.replay()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): g.replay()
    torch.cuda.synchronize(); graph = (time.perf_counter() - t0) / n
print(f"eager {1e3*eager:.3f} ms/step   graph replay {1e3*graph:.3f} ms/step   ({n} steps)")
