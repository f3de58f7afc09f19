"""Multi-GPU check (torch.distributed.run, one rank per GPU): the sharded k-sample evaluation gives the
single-GPU per-example values bit for bit, and the data-parallel train step (image shards + one all-reduce
of the flat gradient + replicated Adam) follows the single-GPU train step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import iwae_b200 as ib
from iwae_b200.dist import model_llh_sharded, dp_train_step, init_comm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
P, X, _ = ib.data.synthetic_binarised(64)
bias = ib.data.get_bias(P.mean(0))
m = ib.IWAE2([200, 100], [100, 50], seed=1, out_bias=bias)
mean, per = model_llh_sharded(m, X[:37], 300, seed=11, gather=True)
ref_per, ref_tot = m.llh(X[:37], 300, seed=11)
ok_eval = torch.equal(per, ref_per) and abs(float(mean) - float(ref_tot) / 37) < 1e-9 * abs(float(mean))
# data-parallel training vs single-GPU training with the same noise (same seed, global image index keying)
m1 = ib.IWAE2([200, 100], [100, 50], seed=1, out_bias=bias)
if os.environ.get('IWAE_DP_TORCH') is None:
    init_comm(m)          # library-owned NCCL communicator: the all-reduce runs inside the step's CUDA graph
o, o1 = ib.Adam(1e-3, epsilon=1e-4), ib.Adam(1e-3, epsilon=1e-4)
for t in range(3):
    xb = X[t * 20:(t + 1) * 20]
    dp_train_step(m, xb, 50, 1.0, o, "iwae_elbo", seed=100 + t)
    m1.train_step(xb, 50, 1.0, o1, objective="iwae_elbo", seed=100 + t)
diff = float((m.params - m1.params).abs().max())
flag = torch.tensor([1.0 if (ok_eval and diff < 2e-5) else 0.0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
# throughput of the data-parallel step (latency-bound: one ~2 MB all-reduce per step)
import time
for t in range(20): dp_train_step(m, X[:20], 50, 1.0, o, "iwae_elbo", seed=500 + t)
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
for t in range(300): dp_train_step(m, X[:20], 50, 1.0, o, "iwae_elbo", seed=600 + t)
torch.cuda.synchronize(); dist.barrier(); dt = (time.perf_counter() - t0) / 300
if rank == 0:
    print(f"world={world} data-parallel train step (k=50, global batch 20): {1e3*dt:.3f} ms/step = {1/dt:.0f} steps/s")
    print(f"world={world} eval bit-exact={ok_eval} dp-vs-single max |dparam|={diff:.2e} ->", "OK" if flag.item() == 1 else "FAIL")
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
