"""GPU: the tools either side of the hot path - device re-binarisation + shuffle in front
of train_step, `sample` / `conditional_sample` (iwae2.py:184-213, task04.py:190-209) and the true-posterior
grid of tasks/plot_task01.py:46-70 - against the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import make_case
from test_gpu_parity import build, rel, RTOL

pytestmark = pytest.mark.gpu


def test_device_bernoullisample_bit_exact(oracle):
    from iwae_b200._abi import lib, check
    rs = np.random.RandomState(0)
    for n, seed, epoch in ((1, 1, 0), (7, 123, 3), (784 * 33 + 2, 2 ** 40 + 17, 1092)):
        p = (rs.rand(n) ** 4).astype(np.float32)
        p[::5] = 0.0
        p[1::7] = 1.0
        tp = torch.tensor(p, device="cuda")
        out = torch.empty_like(tp)
        check(lib.iwae_bernoulli_sample(C.c_void_p(tp.data_ptr()), n, seed, epoch, C.c_void_p(out.data_ptr()),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        got = out.cpu().numpy()
        assert np.array_equal(got, oracle.philox_bernoulli(p, seed, epoch))
        assert not got[p == 0.0].any() and got[p == 1.0].all()     # p = 0 never fires, p = 1 always (u < 1)


def test_device_dataset_epochs(oracle):
    import iwae_b200 as ib
    P, _, y = oracle.synthetic_data(203, seed=11)
    ds = ib.data.DeviceDataset(P, labels=y, batch_size=20, seed=7)
    assert len(ds) == 10
    ds.start_epoch(0)
    X0, perm0 = ds.X.clone(), ds.perm.clone()
    assert np.array_equal(X0.cpu().numpy(), oracle.philox_bernoulli(P.astype(np.float32), 7, 0))
    assert sorted(perm0.tolist()) == list(range(203))
    ds.start_epoch(1)
    assert not torch.equal(ds.X, X0) and not torch.equal(ds.perm, perm0)
    assert abs(float(ds.X.mean()) - float(P.mean())) < 0.01
    ds.start_epoch(0)                                          # an epoch is a pure function of (seed, epoch)
    assert torch.equal(ds.X, X0) and torch.equal(ds.perm, perm0)
    xb, yb = ds.batch(3)
    idx = perm0[60:80]
    assert xb.is_cuda and xb.shape == (20, 784) and torch.equal(xb, X0[idx]) and torch.equal(yb.cpu(), torch.tensor(y, dtype=torch.float32)[idx.cpu()])
    # a device batch goes straight into train_step
    m = ib.IWAE(200, 50)
    opt = ib.Adam(1e-3, epsilon=1e-4)
    r1 = m.train_step(xb, 5, 1.0, opt, objective="iwae_elbo", seed=3)
    m2 = ib.IWAE(200, 50)
    r2 = m2.train_step(xb.cpu().numpy(), 5, 1.0, ib.Adam(1e-3, epsilon=1e-4), objective="iwae_elbo", seed=3)
    assert float(r1["iwae_elbo"]) == float(r2["iwae_elbo"])
    assert torch.equal(m.params, m2.params)


def test_iwae2_sample_and_conditional_sample(oracle):
    case = make_case(oracle, "iwae2", [200, 100], [100, 50], 2, 1, seed=43, scale=0.7)
    m = build("iwae2", dict(h=[200, 100], d=[100, 50]), case)
    p = oracle.to_torch(case["params"])
    z2 = np.random.RandomState(1).randn(6, 50).astype(np.float32)
    mu, sd = oracle.basic_block(p, "decoder.decode_z2_to_z1", torch.tensor(z2, dtype=torch.float64))
    for n, call in ((1, lambda: m.sample(z2, seed=9)), (10, lambda: m.conditional_sample(z2, seed=9))):
        xs, probs = call()
        eps = torch.tensor(oracle.philox_normal(9, 1, 0, 6, n, 100), dtype=torch.float64)       # [n, 6, 100]
        z1 = mu + sd * eps
        ref = torch.sigmoid(oracle.mlp3(p, "decoder.decode_z1_to_x", z1))
        if n == 1:
            ref = ref[0]
        assert probs.shape == tuple(ref.shape) and xs.shape == probs.shape
        assert np.max(np.abs(probs.numpy() - ref.numpy())) < 1e-5
        assert np.array_equal(xs.numpy(), oracle.philox_bernoulli(probs.numpy(), 9, 0))


@pytest.mark.parametrize("learned", [True, False])
def test_ciwae_sample(oracle, learned):
    kind = "ciwae" if learned else "ciwae_std"
    case = make_case(oracle, kind, 200, 100, 2, 1, seed=44, scale=0.7, n_cond=10)
    m = build(kind, dict(h=200, d=100), case)
    p = oracle.to_torch(case["params"])
    z = np.random.RandomState(2).randn(5, 100).astype(np.float32)
    label = 7
    xs, probs = m.sample(z, np.array([label]), seed=4)
    onehot = torch.zeros(5, 10, dtype=torch.float64)
    onehot[:, label] = 1.0
    zt = torch.tensor(z, dtype=torch.float64)
    if learned:
        mu, sd = oracle.basic_block(p, "conditional_prior_network", onehot)
        zt = mu + sd * zt                                       # task04.py:197
    ref = torch.sigmoid(oracle.mlp3(p, "decoder.decode_z_to_x", torch.cat([zt, onehot], -1)))
    assert probs.shape == (5, 784)
    assert np.max(np.abs(probs.numpy() - ref.numpy())) < 1e-5
    assert set(np.unique(xs.numpy())) <= {0.0, 1.0}


def test_posterior_grid(oracle):
    """tasks/plot_task01.py:46-70 on a 2-D latent: normalised true log-posterior and log q(z|x) on the grid."""
    case = make_case(oracle, "iwae1", 200, 2, 1, 1, seed=45)
    m = build("iwae1", dict(h=200, d=2), case)
    p = oracle.to_torch(case["params"])
    x = torch.tensor(case["x"], dtype=torch.float64)
    q_mu, q_sd = oracle.basic_block(p, "encoder.encode_x_to_z", x)
    q_mu, q_sd = q_mu[0].numpy(), q_sd[0].numpy()
    rx = [max(-3, q_mu[0] - 2 * q_sd[0]), min(3, q_mu[0] + 2 * q_sd[0])]
    ry = [max(-3, q_mu[1] - 2 * q_sd[1]), min(3, q_mu[1] + 2 * q_sd[1])]
    n1, n2 = 40, 30
    lp, lq, grid = m.posterior_grid(case["x"][0], rx, ry, n1, n2)
    gx, gy = np.meshgrid(np.linspace(rx[0], rx[1], n1), np.linspace(ry[0], ry[1], n2))
    g = np.asarray([gx.reshape(-1), gy.reshape(-1)]).T
    assert np.max(np.abs(grid.numpy() - g)) < 1e-5
    gt = torch.tensor(g)
    logits = oracle.mlp3(p, "decoder.decode_z_to_x", gt)
    un = oracle.bernoulli_log_prob(logits, x).sum(-1) + (-0.5 * gt ** 2 - 0.5 * np.log(2 * np.pi)).sum(-1)
    ref_lp = (un - torch.logsumexp(un, 0)).reshape(n2, n1).numpy()
    ref_lq = torch.distributions.Normal(torch.tensor(q_mu), torch.tensor(q_sd)).log_prob(gt).sum(-1).reshape(n2, n1).numpy()
    assert lp.shape == (n2, n1) and abs(float(np.exp(lp.numpy().astype(np.float64)).sum()) - 1.0) < 1e-4
    assert np.max(np.abs(lp.numpy() - ref_lp)) < 1e-3
    assert np.max(np.abs(lq.numpy() - ref_lq)) < 1e-3 * max(1.0, np.max(np.abs(ref_lq)))


_TRAIN_SCRIPT = r"""
import hashlib, sys, numpy as np, torch
sys.path.insert(0, {root!r})
import iwae_b200 as ib
P, X, y = ib.data.synthetic_binarised(160, seed=5)
kind = sys.argv[1]
if kind == "iwae2":
    m = ib.IWAE2([200, 100], [100, 50], out_bias=ib.data.get_bias(P.mean(0)))
elif kind == "dreg":
    m = ib.IWAEDReG(200, 50, out_bias=ib.data.get_bias(P.mean(0)))
else:
    m = ib.CIWAE(200, 50, out_bias=ib.data.get_bias(P.mean(0)))
opt = ib.Adam(1e-3, epsilon=1e-4)
vals = []
for i in range(8):
    B = 13 if i == 5 else 20                       # one odd batch in the middle: falls back to the eager path
    xb = torch.from_numpy(X[i * 20:i * 20 + B]).cuda()
    if i == 6: opt.learning_rate.assign(3e-4)      # the learning rate is a per-step scalar, not part of the graph
    if kind == "ciwae":
        res = m.train_step(xb, y[i * 20:i * 20 + B], 7, 1.0, opt, objective="iwae_elbo", seed=100 + i)
    elif kind == "dreg":
        res = m.train_step(xb, 7, 1.0, opt, seed=100 + i)
    else:
        res = m.train_step(xb, 7, 1.0, opt, objective="iwae_elbo", seed=100 + i)
    vals.append(float(res["iwae_elbo"]))
    vals.append(float(np.asarray(res[{{"iwae2": "lpxz1", "dreg": "lpxz", "ciwae": "lpxzy"}}[kind]].numpy()).sum()))
torch.cuda.synchronize()
print("HASH", hashlib.sha256(m.params.cpu().numpy().tobytes()).hexdigest(), " ".join("%.9g" % v for v in vals))
"""


@pytest.mark.parametrize("kind", ["iwae2", "dreg", "ciwae"])
def test_train_step_graph_replay_is_bit_identical_to_eager(kind, tmp_path):
    """iwae_train_step replays a captured CUDA graph from the third call with one signature on; parameters and
    returned values must be bit-identical to the eager launches (IWAE_NO_GRAPH=1), across a learning-rate change
    and an odd batch size."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "train.py"
    script.write_text(_TRAIN_SCRIPT.format(root=root))
    outs = []
    for no_graph in (False, True):
        env = dict(os.environ)
        env.pop("IWAE_NO_GRAPH", None)
        if no_graph:
            env["IWAE_NO_GRAPH"] = "1"
        r = subprocess.run([sys.executable, str(script), kind], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append([ln for ln in r.stdout.splitlines() if ln.startswith("HASH")][-1])
    assert outs[0] == outs[1]
