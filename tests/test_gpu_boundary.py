"""GPU: the drop-in boundary beyond the result dict - the tfd-like distribution shims the reference's
scripts use (`pxz.log_prob(x)`, `qzx.log_prob(z)`, `.sample()`, tasks/plot_task01.py:40-72), `save_weights` /
`load_weights` (main.py:165, plot_task01.py:194), the lazily materialised `z` / `logits` of a train step, and the
packed-weight cache of the fused evaluation (must follow every parameter change)."""
import numpy as np
import pytest
import torch

from conftest import make_case
from test_gpu_parity import build, rel, RTOL

pytestmark = pytest.mark.gpu


def test_distribution_shims_match_torch_distributions(oracle):
    case = make_case(oracle, "iwae1", 200, 2, 3, 7, seed=51)
    m = build("iwae1", dict(h=200, d=2), case)
    p = oracle.to_torch(case["params"])
    x = torch.tensor(case["x"], dtype=torch.float64)
    # encoder: (z, qzx) with loc / scale / log_prob / sample (plot_task01.py:40-42,72)
    z, qzx = m.encoder(case["x"], 7)
    mu, sd = oracle.basic_block(p, "encoder.encode_x_to_z", x)
    assert rel(qzx.loc.numpy(), mu.numpy()) < RTOL and rel(qzx.scale.numpy(), sd.numpy()) < RTOL
    grid = np.random.RandomState(0).randn(11, 3, 2).astype(np.float32)
    ref = torch.distributions.Normal(mu, sd).log_prob(torch.tensor(grid, dtype=torch.float64))
    got = qzx.log_prob(grid)
    assert got.shape == (11, 3, 2) and np.max(np.abs(got.numpy() - ref.numpy())) < 1e-4 * max(1.0, float(ref.abs().max()))
    s = qzx.sample(5, seed=17)
    eps = oracle.philox_normal(17, 1, 0, 3, 5, 2)
    assert s.shape == (5, 3, 2) and np.max(np.abs(s.numpy() - (eps * sd.numpy() + mu.numpy()))) < 1e-5
    assert qzx.sample(seed=17).shape == (3, 2)
    assert not np.array_equal(qzx.sample(5).numpy(), qzx.sample(5).numpy())       # fresh noise per call
    # decoder: (logits, pxz) with elementwise log_prob (plot_task01.py:57-63), sample, probs
    logits, pxz = m.decoder(z)
    ref_logits = oracle.mlp3(p, "decoder.decode_z_to_x", torch.tensor(z.numpy(), dtype=torch.float64))
    lp = pxz.log_prob(case["x"])
    ref_lp = oracle.bernoulli_log_prob(ref_logits, x)
    assert lp.shape == (7, 3, 784) and np.max(np.abs(lp.numpy() - ref_lp.numpy())) < 1e-4
    assert rel(lp.numpy().sum(-1), pxz.log_prob_sum(case["x"]).numpy()) < 1e-5          # the fused reduction agrees
    assert np.max(np.abs(pxz.probs.numpy() - torch.sigmoid(ref_logits).numpy())) < 1e-5
    xs = pxz.sample(seed=3)
    assert np.array_equal(xs.numpy(), oracle.philox_bernoulli(pxz.probs.numpy(), 3, 0))


def test_train_step_large_entries_are_pre_update_or_refused(oracle):
    case = make_case(oracle, "iwae2", [200, 100], [100, 50], 4, 6, seed=52)
    import iwae_b200 as ib
    m = build("iwae2", dict(h=[200, 100], d=[100, 50]), case)
    p64 = oracle.to_torch(case["params"])
    ref = oracle.iwae2_call(p64, torch.tensor(case["x"], dtype=torch.float64), 6, 1.0,
                            torch.tensor(case["eps1"], dtype=torch.float64), torch.tensor(case["eps2"], dtype=torch.float64))
    opt = ib.Adam(1e-3, epsilon=1e-4)
    res = m.train_step(case["x"], 6, 1.0, opt, objective="iwae_elbo", eps=case["eps1"], eps2=case["eps2"],
                       materialize=("z1", "z2", "logits"))
    # the PRE-update forward, like the reference's tape (src/iwae2.py:169-178)
    assert rel(res["z1"].numpy(), ref["z1"].numpy()) < RTOL and rel(res["z2"].numpy(), ref["z2"].numpy()) < RTOL
    assert rel(res["logits"].numpy(), ref["logits"].numpy()) < RTOL
    res = m.train_step(case["x"], 6, 1.0, opt, objective="iwae_elbo", eps=case["eps1"], eps2=case["eps2"])
    assert float(res["iwae_elbo"]) == float(res["iwae_elbo"])           # scalars are there
    with pytest.raises(RuntimeError):
        res["logits"]                                                   # not materialised, parameters already updated
    # a plain call stays lazy and valid ...
    r = m(case["x"], 6, eps=case["eps1"], eps2=case["eps2"])
    assert r["logits"].shape == (6, 4, 784)
    # ... until the weights change under it
    r = m(case["x"], 6, eps=case["eps1"], eps2=case["eps2"])
    m.trainable_weights[0].assign(m.trainable_weights[0].numpy() * 0.5)
    with pytest.raises(RuntimeError):
        r["z1"]
    assert opt.iterations == 2


@pytest.mark.parametrize("kind", ["iwae1", "iwae2", "ciwae"])
def test_save_load_round_trip(oracle, kind, tmp_path):
    import iwae_b200 as ib
    kw = dict(iwae1=dict(h=200, d=50), iwae2=dict(h=[200, 100], d=[100, 50]), ciwae=dict(h=200, d=100))[kind]
    case = make_case(oracle, kind, kw["h"], kw["d"], 5, 9, seed=53, n_cond=10 if kind == "ciwae" else 0)
    m = build(kind, kw, case)
    path = str(tmp_path / "final_weights")
    m.save_weights(path)                                               # main.py:165
    m2 = {"iwae1": lambda: ib.IWAE(200, 50, seed=99), "iwae2": lambda: ib.IWAE2([200, 100], [100, 50], seed=99),
          "ciwae": lambda: ib.CIWAE(200, 100, seed=99)}[kind]()
    assert not torch.equal(m.params, m2.params)
    m2.load_weights(path)                                              # plot_task01.py:194
    assert torch.equal(m.params, m2.params)
    assert [v.name for v in m.trainable_weights] == [v.name for v in m2.trainable_weights]
    y = case["y"] if kind == "ciwae" else None
    a, _ = m.llh(case["x"], 200, y=y, seed=5)
    b, _ = m2.llh(case["x"], 200, y=y, seed=5)
    assert torch.equal(a, b)


def test_packed_weight_cache_follows_every_parameter_change(oracle):
    """iwae_set_params_version: the fused evaluation may reuse its packed fp16 weights only while the parameters are
    unchanged - after train_step, Variable.assign, set_weights and load into the flat buffer it must re-pack."""
    import iwae_b200 as ib
    case = make_case(oracle, "iwae2", [200, 100], [100, 50], 6, 4, seed=54)
    m = build("iwae2", dict(h=[200, 100], d=[100, 50]), case)

    def fresh_value():
        f = ib.IWAE2([200, 100], [100, 50], seed=1)
        f.params.copy_(m.params)
        return f.llh(case["x"], 256, seed=8)[0]

    a = m.llh(case["x"], 256, seed=8)[0]
    assert torch.equal(a, m.llh(case["x"], 256, seed=8)[0])            # cached image, same result
    assert torch.equal(a, fresh_value())
    opt = ib.Adam(1e-2, epsilon=1e-4)
    for change in ("train", "assign", "set_weights", "inplace"):
        if change == "train":
            m.train_step(case["x"], 4, 1.0, opt, objective="iwae_elbo", seed=2)
        elif change == "assign":
            v = m.trainable_weights[-2]
            v.assign(v.numpy() * 1.1)
        elif change == "set_weights":
            m.set_weights([w * 0.97 for w in m.get_weights().values()])
        else:
            m.params.mul_(1.01)
        b = m.llh(case["x"], 256, seed=8)[0]
        assert not torch.equal(a, b), change
        assert torch.equal(b, fresh_value()), change
        a = b
