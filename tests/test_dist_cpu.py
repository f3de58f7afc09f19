"""CPU, world_size 2 over gloo: the host-side logic of the sharded evaluation and of data-parallel
training, with the oracle standing in for the per-rank kernels."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from iwae_b200.dist import shard_range
    assert [shard_range(20, r, 8) for r in range(8)] == [(0, 3), (3, 6), (6, 9), (9, 12), (12, 14), (14, 16), (16, 18), (18, 20)]
    for n in (0, 1, 7, 10000):
        for w in (1, 2, 4, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1
    with pytest.raises(ValueError):
        shard_range(5, 2, 2)


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import iwae_oracle as oracle
    from conftest import make_case
    from iwae_b200.dist import llh_eval_sharded, shard_range
    N, k = 5, 12
    case = make_case(oracle, "iwae1", 32, 4, N, k, seed=3)
    p = oracle.to_torch(case["params"])
    x = torch.tensor(case["x"], dtype=torch.float64)

    def local_eval(a, b):
        eps = torch.tensor(oracle.philox_normal(9, 1, a, b - a, k, 4), dtype=torch.float64)
        r = oracle.iwae1_call(p, x[a:b], k, 1.0, eps)
        per = r["iwae_elbo_per_example"]
        return per, per.sum()

    mean, per_all = llh_eval_sharded(local_eval, N, gather=True)
    # data-parallel gradient: sum over ranks of shard gradients scaled by 1/B_global == full-batch gradient
    a, b = shard_range(N, rank, world)
    pg = oracle.to_torch(case["params"], requires_grad=True)
    eps = torch.tensor(case["eps1"], dtype=torch.float64)
    r = oracle.iwae1_call(pg, x[a:b], k, 1.0, eps[:, a:b])
    loss = -(r["iwae_elbo_per_example"].sum() / N)
    g = torch.cat([t.reshape(-1) for t in torch.autograd.grad(loss, list(pg.values()))])
    dist.all_reduce(g)
    if rank == 0:
        np.savez(tmp, mean=mean.numpy(), per=per_all.numpy(), grad=g.numpy())
    dist.destroy_process_group()


def test_two_rank_eval_and_dp_gradient(tmp_path, oracle):
    from conftest import make_case
    out = str(tmp_path / "r0.npz")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    z = np.load(out)
    N, k = 5, 12
    case = make_case(oracle, "iwae1", 32, 4, N, k, seed=3)
    p = oracle.to_torch(case["params"], requires_grad=True)
    x = torch.tensor(case["x"], dtype=torch.float64)
    eps = torch.tensor(oracle.philox_normal(9, 1, 0, N, k, 4), dtype=torch.float64)
    ref = oracle.iwae1_call(p, x, k, 1.0, eps)["iwae_elbo_per_example"].detach().numpy()
    assert np.allclose(z["per"], ref, rtol=1e-12)            # sharding does not change per-example values
    assert np.isclose(float(z["mean"]), ref.mean(), rtol=1e-12)
    g, _ = oracle.grads_objective(oracle.iwae1_call, p, "iwae_elbo", x, k, 1.0, torch.tensor(case["eps1"], dtype=torch.float64))
    gref = torch.cat([t.reshape(-1) for t in g.values()]).numpy()
    assert np.allclose(z["grad"], gref, rtol=1e-9, atol=1e-12)
