"""CPU: pin the oracle.  The reference has no tests/golden vectors and its TF2 stack is not installable
(SURVEY.md §8c -> "parity unpinned"), so the oracle is pinned against independent implementations of
the same third-party semantics (torch.distributions, torch.autograd), against the one piece of the
reference path that DOES run here (numpy RandomState.choice), and against the committed fixtures."""
import os

import numpy as np
import pytest
import torch

from conftest import make_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_log_probs_match_torch_distributions(oracle):
    g = torch.Generator().manual_seed(0)
    z = torch.randn(7, 3, 5, generator=g, dtype=torch.float64)
    mu = torch.randn(3, 5, generator=g, dtype=torch.float64)
    sd = torch.rand(3, 5, generator=g, dtype=torch.float64) + 0.1
    ref = torch.distributions.Normal(mu, sd).log_prob(z)
    assert torch.allclose(oracle.normal_log_prob(z, mu, sd), ref, atol=1e-12)
    assert torch.allclose(oracle.std_normal_log_prob(z), torch.distributions.Normal(0.0, 1.0).log_prob(z), atol=1e-12)
    l = torch.randn(7, 3, 11, generator=g, dtype=torch.float64) * 4
    x = (torch.rand(3, 11, generator=g) < 0.3).double()
    ref = torch.distributions.Bernoulli(logits=l).log_prob(x.expand_as(l))
    assert torch.allclose(oracle.bernoulli_log_prob(l, x), ref, atol=1e-12)
    kl = torch.distributions.kl_divergence(torch.distributions.Normal(mu, sd), torch.distributions.Normal(0.0, 1.0))
    assert torch.allclose(oracle.kl_normal_std(mu, sd), kl.to(torch.float64), atol=1e-12)


def test_logmeanexp_and_softmax_identities(oracle):
    g = torch.Generator().manual_seed(1)
    lw = torch.randn(50, 4, generator=g, dtype=torch.float64) * 30 - 500
    ref = torch.logsumexp(lw, 0) - np.log(50)
    assert torch.allclose(oracle.logmeanexp(lw, 0), ref, atol=1e-10)


@pytest.mark.parametrize("objective", ["iwae_elbo", "iwae_eq14", "vae_elbo"])
def test_upstream_gradient_identity(oracle, objective):
    """d objective / d log_w: softmax/B for iwae_elbo and eq14, 1/(kB) for vae_elbo."""
    g = torch.Generator().manual_seed(2)
    lw = (torch.randn(6, 3, generator=g, dtype=torch.float64) * 3).requires_grad_(True)
    res = oracle._objectives(lw)
    (grad,) = torch.autograd.grad(res[objective], lw)
    want = torch.softmax(lw, 0) / 3 if objective != "vae_elbo" else torch.full_like(lw, 1.0 / 18)
    assert torch.allclose(grad, want.detach(), atol=1e-14)


def test_param_counts_match_survey(oracle):
    n = lambda s: sum(int(np.prod(v)) for v in s.values())
    assert n(oracle.param_shapes("iwae1", 200, 100)) == 455384
    assert n(oracle.param_shapes("iwae1", 200, 50)) == 425284
    assert n(oracle.param_shapes("iwae2", [200, 100], [100, 50])) == 521084
    assert n(oracle.param_shapes("ciwae", 200, 100, n_cond=10)) == 541984
    assert len(oracle.param_shapes("iwae2", [200, 100], [100, 50])) == 30


def test_sir_matches_numpy_choice_bit_exact(oracle):
    """tasks/plot_task01.py:88 runs as-is here: np.random.choice(replace=False, p)."""
    for seed in range(12):
        rs = np.random.RandomState(seed)
        L = [10000, 500, 37][seed % 3]
        size = min(200, L // 2)
        lw = (rs.randn(L) * (1 + seed % 4)).astype(np.float32)
        al = oracle.softmax_f32(lw)
        rs_a, rs_b = np.random.RandomState(100 + seed), np.random.RandomState(100 + seed)
        want = rs_a.choice(np.arange(L), size=size, replace=False, p=al)
        got, used = oracle.sir_choice(al, size, rs_b.random_sample(20 * size))
        assert np.array_equal(want, got)
        # stream position after the call is identical too
        rs_c = np.random.RandomState(100 + seed)
        rs_c.random_sample(used)
        assert rs_a.random_sample() == rs_c.random_sample()


def test_philox_known_answers(oracle):
    """Random123 known-answer vectors for philox4x32-10."""
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        got = oracle.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert tuple(int(v) for v in got) == want


def test_philox_normal_moments(oracle):
    e = oracle.philox_normal(123, 1, 0, 8, 2000, 50).astype(np.float64)
    assert e.shape == (2000, 8, 50)
    assert abs(e.mean()) < 5e-3 and abs(e.std() - 1) < 5e-3
    # keyed on the global image index: shard [4:8] equals rows 4..7 of the whole
    e2 = oracle.philox_normal(123, 1, 4, 4, 2000, 50)
    assert np.array_equal(e[:, 4:8].astype(np.float32), e2)


def test_keras_adam_rule(oracle):
    rs = np.random.RandomState(0)
    th, g = rs.randn(10), rs.randn(10)
    th1, m, v = oracle.keras_adam_step(th, g, np.zeros(10), np.zeros(10), 1, 1e-3)
    lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    assert np.allclose(th1, th - lr_t * (0.1 * g) / (np.sqrt(0.001 * g * g) + 1e-4))
    sched, epochs = oracle.lr_schedule()
    assert epochs == 3280 and sorted(sched)[:4] == [0, 1, 4, 13] and abs(sched[0] - 1e-3) < 1e-12


def test_fp32_restatement_tracks_fp64(oracle):
    c = make_case(oracle, "iwae2", [200, 100], [100, 50], 2, 64)
    out = {}
    for dt in (torch.float64, torch.float32):
        p = oracle.to_torch(c["params"], dt)
        r = oracle.iwae2_call(p, torch.tensor(c["x"], dtype=dt), 64, 1.0, torch.tensor(c["eps1"], dtype=dt),
                              torch.tensor(c["eps2"], dtype=dt))
        out[dt] = r["iwae_elbo_per_example"].double().numpy()
    rel = np.abs(out[torch.float32] - out[torch.float64]) / np.abs(out[torch.float64])
    assert rel.max() < 2e-6


def test_dreg_gradient_structure(oracle):
    """Encoder grads of inference_loss flow only through z (task02.py:61-76): check against a manual
    path-wise computation."""
    c = make_case(oracle, "iwae1", 20, 4, 3, 6)
    p = oracle.to_torch(c["params"], torch.float64, requires_grad=True)
    x = torch.tensor(c["x"], dtype=torch.float64)
    eps = torch.tensor(c["eps1"], dtype=torch.float64)
    g, res = oracle.grads_dreg(p, x, 6, 1.0, eps)
    assert all(torch.isfinite(v).all() for v in g.values())
    # decoder grads equal those of -iwae_elbo
    g2, _ = oracle.grads_objective(oracle.iwae1_call, p, "iwae_elbo", x, 6, 1.0, eps)
    for k in g:
        if k.startswith("decoder."):
            assert torch.allclose(g[k], g2[k], atol=1e-12)


def test_golden_fixtures_reproduce(oracle):
    """tests/golden/*.npz were written by tests/golden/make_golden.py from this oracle (fp64)."""
    import glob
    files = sorted(glob.glob(os.path.join(GOLD, "case_*.npz")))
    assert files, "golden fixtures missing: run python tests/golden/make_golden.py"
    from golden.make_golden import run_case, CASES
    for f in files:
        with np.load(f) as z:
            name = str(z["name"])
            want = {k: z[k] for k in z.files if k.startswith("out_")}
        got = run_case(oracle, CASES[name])[1]
        for k, v in want.items():
            assert np.allclose(got[k[4:]], v, rtol=1e-10, atol=1e-10), (name, k)


def test_full_model_against_independent_torch_distributions_build(oracle):
    """The whole two-layer call, rebuilt independently of the oracle's formulas with torch.distributions objects
    (Normal / Bernoulli log_prob, as the reference builds it with tfd), must agree with the restatement,
    including the gradient of -iwae_elbo obtained by autograd through the distribution objects."""
    D = torch.distributions
    c = make_case(oracle, "iwae2", [200, 100], [100, 50], 3, 7, seed=5, scale=0.5)
    k = 7
    p = oracle.to_torch(c["params"], torch.float64, requires_grad=True)
    x = torch.tensor(c["x"], dtype=torch.float64)
    e1 = torch.tensor(c["eps1"], dtype=torch.float64)
    e2 = torch.tensor(c["eps2"], dtype=torch.float64)

    def block(prefix, inp):
        h = torch.tanh(inp @ p[prefix + ".l1.kernel"] + p[prefix + ".l1.bias"])
        h = torch.tanh(h @ p[prefix + ".l2.kernel"] + p[prefix + ".l2.bias"])
        return D.Normal(h @ p[prefix + ".lmu.kernel"] + p[prefix + ".lmu.bias"],
                        torch.exp(h @ p[prefix + ".lstd.kernel"] + p[prefix + ".lstd.bias"]) + 1e-6)

    qz1 = block("encoder.encode_x_to_z1", x)
    z1 = qz1.loc + qz1.scale * e1
    qz2 = block("encoder.encode_z1_to_z2", z1)
    z2 = qz2.loc + qz2.scale * e2
    pz1 = block("decoder.decode_z2_to_z1", z2)
    h = z1
    for i, act in enumerate((torch.tanh, torch.tanh, lambda t: t)):
        h = act(h @ p[f"decoder.decode_z1_to_x.dense{i}.kernel"] + p[f"decoder.decode_z1_to_x.dense{i}.bias"])
    log_w = (D.Bernoulli(logits=h).log_prob(x.expand(k, -1, -1)).sum(-1) + pz1.log_prob(z1).sum(-1)
             + D.Normal(0.0, 1.0).log_prob(z2).sum(-1) - qz1.log_prob(z1).sum(-1) - qz2.log_prob(z2).sum(-1))
    iwae = (torch.logsumexp(log_w, 0) - np.log(k)).mean()
    g_ind = torch.autograd.grad(-iwae, list(p.values()))
    g_or, res = oracle.grads_objective(oracle.iwae2_call, p, "iwae_elbo", x, k, 1.0, e1, e2)
    assert torch.allclose(res["log_w"], log_w, rtol=1e-10, atol=1e-9)
    assert abs(float(res["iwae_elbo"]) - float(iwae)) < 1e-10 * abs(float(iwae))
    for (name, go), gi in zip(g_or.items(), g_ind):
        assert torch.allclose(go, gi, rtol=1e-8, atol=1e-12), name


def test_philox_bernoulli_restatement(oracle):
    """The device re-binarisation rule (utils.bernoullisample on the GPU): p = 0 / 1 are never / always drawn, the mean
    follows p, an epoch is a pure function of (seed, epoch), and the four draws of one Philox call are the call's four
    words in order (checked against the Philox restatement that the Random123 known-answer tests pin)."""
    rs = np.random.RandomState(3)
    p = rs.rand(50, 784).astype(np.float32) ** 4
    p[0, :10] = 0.0
    p[0, 10:20] = 1.0
    a = oracle.philox_bernoulli(p, 123, 0)
    assert a.shape == p.shape and set(np.unique(a)) <= {0.0, 1.0}
    assert not a[0, :10].any() and a[0, 10:20].all()
    assert abs(a.mean() - p.mean()) < 0.01
    assert np.array_equal(a, oracle.philox_bernoulli(p, 123, 0))
    assert not np.array_equal(a, oracle.philox_bernoulli(p, 123, 1)) and not np.array_equal(a, oracle.philox_bernoulli(p, 124, 0))
    words = oracle.philox4x32_10(np.array([[1, 0, 7, 0x42]], dtype=np.uint32), np.array([[9, 0]], dtype=np.uint32))[0]
    u = (words >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    probe = np.full(8, 0.5, dtype=np.float32)                   # elements 4..7 use Philox call q = 1
    assert np.array_equal(oracle.philox_bernoulli(probe, 9, 7)[4:], (u < 0.5).astype(np.float32))


def test_two_layer_call_against_numpy_scipy(oracle):
    """A third, torch-free statement of iwae2.IWAE.call (iwae2.py:109-167): plain NumPy matmuls with SciPy's
    norm.logpdf / expit / logsumexp (another library lineage than torch.distributions) must reproduce every per-row
    term, log_w, the importance weights and the three ELBOs of the oracle."""
    from scipy.special import expit, logsumexp
    from scipy.stats import norm
    c = make_case(oracle, "iwae2", [200, 100], [100, 50], 4, 6, seed=9, scale=0.7)
    k = 6
    P = {n: np.asarray(v, dtype=np.float64) for n, v in c["params"].items()}
    x = c["x"].astype(np.float64)
    e1, e2 = c["eps1"].astype(np.float64), c["eps2"].astype(np.float64)

    def block(prefix, inp):
        h = np.tanh(inp @ P[prefix + ".l1.kernel"] + P[prefix + ".l1.bias"])
        h = np.tanh(h @ P[prefix + ".l2.kernel"] + P[prefix + ".l2.bias"])
        return h @ P[prefix + ".lmu.kernel"] + P[prefix + ".lmu.bias"], np.exp(h @ P[prefix + ".lstd.kernel"] + P[prefix + ".lstd.bias"]) + 1e-6

    mu1, sd1 = block("encoder.encode_x_to_z1", x)
    z1 = mu1 + sd1 * e1
    mu2, sd2 = block("encoder.encode_z1_to_z2", z1)
    z2 = mu2 + sd2 * e2
    pmu, psd = block("decoder.decode_z2_to_z1", z2)
    h = z1
    for i, act in enumerate((np.tanh, np.tanh, lambda t: t)):
        h = act(h @ P[f"decoder.decode_z1_to_x.dense{i}.kernel"] + P[f"decoder.decode_z1_to_x.dense{i}.bias"])
    pr = expit(h)
    terms = {
        "lpxz1": (x * np.log(pr) + (1.0 - x) * np.log1p(-pr)).sum(-1),
        "lpz1z2": norm.logpdf(z1, pmu, psd).sum(-1),
        "lpz2": norm.logpdf(z2).sum(-1),
        "lqz1x": norm.logpdf(z1, mu1, sd1).sum(-1),
        "lqz2z1": norm.logpdf(z2, mu2, sd2).sum(-1),
    }
    log_w = terms["lpxz1"] + terms["lpz1z2"] + terms["lpz2"] - terms["lqz1x"] - terms["lqz2z1"]
    p64 = oracle.to_torch(c["params"], torch.float64)
    res = oracle.iwae2_call(p64, torch.tensor(x), k, 1.0, torch.tensor(e1), torch.tensor(e2))
    for name, v in terms.items():
        assert np.allclose(res[name].numpy(), v, rtol=1e-9, atol=1e-8), name
    assert np.allclose(res["log_w"].numpy(), log_w, rtol=1e-10, atol=1e-8)
    assert np.allclose(res["al"].numpy(), np.exp(log_w - logsumexp(log_w, axis=0)), rtol=1e-9, atol=1e-12)
    assert abs(float(res["iwae_elbo"]) - float((logsumexp(log_w, axis=0) - np.log(k)).mean())) < 1e-9
    assert abs(float(res["vae_elbo"]) - float(log_w.mean())) < 1e-9
