"""Host-side cost of one model.llh call (tiny workload: the GPU work is negligible, the wall time per call is the Python +
ctypes + launch path that sits inside every timed bench step)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import iwae_b200 as ib
P, X, _ = ib.data.synthetic_binarised(64)
m = ib.IWAE2([200, 100], [100, 50], out_bias=ib.data.get_bias(P.mean(0)))
x = torch.from_numpy(X[:2]).cuda()
for _ in range(20): m.llh(x, 128, precision="tc_f16_fast", seed=1)
torch.cuda.synchronize()
n = 500
t0 = time.perf_counter()
for _ in range(n): per, tot = m.llh(x, 128, precision="tc_f16_fast", seed=1)
torch.cuda.synchronize()
print(f"llh call, async: {1e6 * (time.perf_counter() - t0) / n:.1f} us per call")
t0 = time.perf_counter()
for _ in range(n): per, tot = m.llh(x, 128, precision="tc_f16_fast", seed=1); v = float(tot.item())
print(f"llh call + .item(): {1e6 * (time.perf_counter() - t0) / n:.1f} us per call")
pr = cProfile.Profile(); pr.enable()
for _ in range(n): per, tot = m.llh(x, 128, precision="tc_f16_fast", seed=1)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
