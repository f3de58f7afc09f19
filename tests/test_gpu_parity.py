"""GPU parity: the CUDA path (through the C-ABI, via the reference-shaped Python classes) against the
fp64 oracle on identical weights, inputs and injected noise, and against the committed fixtures.

Tolerances (BASELINE.json north_star): per-example log-weights / ELBO / LLH <= 1e-4 relative in fp32
mode; gradients <= 1e-3 relative; SIR indices bit-exact; the tensor-core (fp16-operand) modes carry
their own stated tolerance (TC_RTOL below)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import make_case

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-4          # fp32 mode, per-row terms and per-example ELBOs
GRAD_RTOL = 1e-3     # gradients, per tensor, relative to the tensor's norm
TC_RTOL = 1e-4       # fp16-operand tensor-core eval, per-example LLH (relative)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))


def build(kind, c, case):
    import iwae_b200 as ib
    if kind == "iwae1":
        m = ib.IWAE(c["h"], c["d"])
    elif kind == "dreg":
        m = ib.IWAEDReG(c["h"], c["d"])
    elif kind == "iwae2":
        m = ib.IWAE2(c["h"], c["d"])
    elif kind == "ciwae":
        m = ib.CIWAE(c["h"], c["d"])
    else:
        m = ib.CIWAE(c["h"], c["d"], learned_prior=False)
    m.set_weights(case["params"])
    return m


def run_model(kind, m, case, k, beta=1.0):
    if kind == "iwae2":
        return m(case["x"], k, beta, eps=case["eps1"], eps2=case["eps2"])
    if kind.startswith("ciwae"):
        return m(case["x"], case["y"], k, beta, eps=case["eps1"])
    return m(case["x"], k, beta, eps=case["eps1"])


def flat_grads(kind, m, case, k, objective, beta=1.0):
    kw = dict(eps=case["eps1"])
    if kind == "iwae2":
        kw["eps2"] = case["eps2"]
    if kind.startswith("ciwae"):
        kw["y"] = case["y"]
    return m.grads(case["x"], k, beta, objective, **kw).cpu().numpy()


def oracle_call(oracle, kind, case, k, beta, dtype=torch.float64, requires_grad=False):
    p = oracle.to_torch(case["params"], dtype, requires_grad=requires_grad)
    x = torch.tensor(case["x"], dtype=dtype)
    e1 = torch.tensor(case["eps1"], dtype=dtype)
    if kind == "iwae1":
        return p, oracle.iwae1_call, (x, k, beta, e1)
    if kind == "dreg":
        return p, oracle.dreg_call, (x, k, beta, e1)
    if kind == "iwae2":
        return p, oracle.iwae2_call, (x, k, beta, e1, torch.tensor(case["eps2"], dtype=dtype))
    learned = kind == "ciwae"
    fn = lambda p_, x_, y_, k_, b_, e_: oracle.ciwae_call(p_, x_, y_, k_, b_, e_, learned_prior=learned)
    return p, fn, (x, case["y"], k, beta, e1)


# ---------------------------------------------------------------------------------------------
# golden fixtures (BASELINE.json configs 0,1,3,4 and the task01/task03 latent sizes)
# ---------------------------------------------------------------------------------------------
def _golden_names():
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from golden.make_golden import CASES
    return list(CASES)


@pytest.mark.parametrize("name", _golden_names())
def test_forward_and_grads_match_golden(oracle, name):
    from golden.make_golden import CASES, GRAD_STRIDE
    c = CASES[name]
    kind = c["kind"]
    okind = "iwae1" if kind == "dreg" else kind
    case = make_case(oracle, okind, c["h"], c["d"], c["B"], c["k"], seed=c["seed"], scale=c["scale"],
                     n_cond=10 if kind.startswith("ciwae") else 0)
    with np.load(os.path.join(GOLD, f"case_{name}.npz")) as z:
        want = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    m = build(kind, c, case)
    res = run_model(kind, m, case, c["k"])
    assert set(k for k in want if not k.startswith("grad") and k != "kl") <= set(res.keys())
    for key, w in want.items():
        if key.startswith("grad") or key == "kl":
            continue
        got = res[key].numpy()
        assert got.shape == w.shape, key
        tol = 2e-3 if key in ("al", "snis_z", "snis_z1", "snis_z2") else RTOL   # softmax amplifies log_w error
        assert rel(got, w) < tol, (name, key, rel(got, w))
    obj = "dreg" if kind == "dreg" else "iwae_elbo"
    g = flat_grads(kind, m, case, c["k"], obj)
    assert np.all(np.isfinite(g))
    sample = g[::GRAD_STRIDE]
    scale = np.max(np.abs(want["grad_sample"]))
    assert np.max(np.abs(sample - want["grad_sample"])) < GRAD_RTOL * scale, name
    norms = np.array([float(np.linalg.norm(g[v.offset:v.offset + int(np.prod(v.shape))])) for v in m.trainable_weights])
    assert np.allclose(norms, want["grad_norms"], rtol=GRAD_RTOL, atol=1e-7), name


# ---------------------------------------------------------------------------------------------
# shape / objective coverage against the live oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,h,d,B,k", [
    ("iwae1", 200, 100, 1, 1), ("iwae1", 200, 50, 20, 5), ("iwae1", 200, 2, 3, 130), ("iwae1", 64, 4, 7, 33),
    ("iwae2", [200, 100], [100, 50], 1, 1), ("iwae2", [200, 100], [4, 2], 5, 17), ("iwae2", [200, 100], [2, 2], 2, 64),
    ("ciwae", 200, 100, 6, 9), ("ciwae_std", 200, 100, 3, 4),
])
def test_forward_shapes(oracle, kind, h, d, B, k):
    case = make_case(oracle, kind, h, d, B, k, seed=11, n_cond=10 if kind.startswith("ciwae") else 0)
    m = build(kind, dict(h=h, d=d), case)
    res = run_model(kind, m, case, k, beta=0.7)
    p, fn, args = oracle_call(oracle, kind, case, k, 0.7)
    ref = fn(p, *args)
    for key in res.keys():
        if key not in ref:
            continue
        got, w = res[key].numpy(), ref[key].detach().numpy()
        assert got.shape == w.shape, key
        tol = 2e-3 if key in ("al", "snis_z", "snis_z1", "snis_z2") else (1e-3 if key in ("z", "z1", "z2", "logits") else RTOL)
        assert rel(got, w) < tol, (key, rel(got, w))


@pytest.mark.parametrize("kind,objective", [
    ("iwae1", "vae_elbo"), ("iwae1", "iwae_elbo"), ("iwae1", "iwae_eq14"), ("iwae1", "vae_elbo_kl"),
    ("iwae2", "vae_elbo"), ("iwae2", "iwae_elbo"), ("iwae2", "iwae_eq14"),
    ("ciwae", "vae_elbo"), ("ciwae", "iwae_elbo"), ("ciwae", "vae_elbo_kl"), ("ciwae_std", "iwae_elbo"),
    ("dreg", "dreg"),
])
def test_gradients_all_objectives(oracle, kind, objective):
    h, d = ([200, 100], [100, 50]) if kind == "iwae2" else (200, 100)
    B, k, beta = 5, 12, 0.8
    okind = "iwae1" if kind == "dreg" else kind
    case = make_case(oracle, okind, h, d, B, k, seed=21, scale=0.5, n_cond=10 if kind.startswith("ciwae") else 0)
    m = build(kind, dict(h=h, d=d), case)
    g = flat_grads(kind, m, case, k, objective, beta)
    p, fn, args = oracle_call(oracle, kind, case, k, beta, requires_grad=True)
    if kind == "dreg":
        gref, _ = oracle.grads_dreg(p, *args)
    else:
        gref, _ = oracle.grads_objective(fn, p, objective, *args)
    for v in m.trainable_weights:
        n = int(np.prod(v.shape))
        a, b = g[v.offset:v.offset + n], gref[v.name].reshape(-1).numpy()
        denom = max(np.linalg.norm(b), 1e-8)
        assert np.linalg.norm(a - b) / denom < GRAD_RTOL or np.linalg.norm(a - b) < 1e-7, (v.name, np.linalg.norm(a - b) / denom)


def test_unknown_objective_is_keyerror(oracle):
    import iwae_b200 as ib
    case = make_case(oracle, "iwae2", [200, 100], [100, 50], 2, 3)
    m = build("iwae2", dict(h=[200, 100], d=[100, 50]), case)
    opt = ib.Adam(1e-3, epsilon=1e-4)
    with pytest.raises(KeyError):       # iwae2 has no 'vae_elbo_kl' (iwae2.py:154-167,173)
        m.train_step(case["x"], 3, 1.0, opt, objective="vae_elbo_kl")
    with pytest.raises(KeyError):
        m.train_step(case["x"], 3, 1.0, opt, objective="nope")
    with pytest.raises(ValueError):
        m(case["x"][:, :100], 3)
    with pytest.raises(ValueError):
        m(case["x"], 3, eps=np.zeros((2, 2, 100), np.float32))


def test_adam_matches_keras_rule(oracle):
    import iwae_b200 as ib
    import ctypes as C
    from iwae_b200._abi import lib, check
    rs = np.random.RandomState(0)
    n = 100003       # odd length: exercises the scalar tail
    th, m_, v_ = rs.randn(n).astype(np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
    d = lambda a: torch.tensor(a, device="cuda")
    tth, tm, tv = d(th), d(m_), d(v_)
    ref = (th.astype(np.float64), m_.astype(np.float64), v_.astype(np.float64))
    for t in range(1, 6):
        g = (rs.randn(n) * 10.0 ** rs.randint(-4, 1)).astype(np.float32)
        tg = d(g)
        check(lib.iwae_adam_step(C.c_void_p(tth.data_ptr()), C.c_void_p(tg.data_ptr()), C.c_void_p(tm.data_ptr()),
                                 C.c_void_p(tv.data_ptr()), n, 1e-3, 0.9, 0.999, 1e-4, t,
                                 torch.cuda.current_stream().cuda_stream))
        ref = oracle.keras_adam_step(ref[0], g.astype(np.float64), ref[1], ref[2], t, 1e-3)
    assert np.max(np.abs(tth.cpu().numpy() - ref[0])) < 2e-6
    assert rel(tm.cpu().numpy(), ref[1]) < 1e-5 and rel(tv.cpu().numpy(), ref[2]) < 1e-5


@pytest.mark.parametrize("kind", ["iwae1", "iwae2", "dreg", "ciwae"])
def test_train_steps_track_oracle(oracle, kind):
    """A few train_step calls with injected noise follow the oracle's autograd + Keras-Adam trajectory."""
    import iwae_b200 as ib
    h, d = ([200, 100], [100, 50]) if kind == "iwae2" else (200, 100)
    B, k, steps = 6, 8, 4
    okind = "iwae1" if kind == "dreg" else kind
    case = make_case(oracle, okind, h, d, B, k, seed=31, n_cond=10 if kind.startswith("ciwae") else 0)
    m = build(kind, dict(h=h, d=d), case)
    opt = ib.Adam(1e-3, epsilon=1e-4)
    p = oracle.to_torch(case["params"], torch.float64, requires_grad=True)
    ms = {n: torch.zeros_like(v) for n, v in p.items()}
    vs = {n: torch.zeros_like(v) for n, v in p.items()}
    rs = np.random.RandomState(5)
    for t in range(1, steps + 1):
        case["eps1"] = rs.randn(k, B, 100).astype(np.float32)
        if kind == "iwae2":
            case["eps2"] = rs.randn(k, B, 50).astype(np.float32)
        if kind == "iwae2":
            res = m.train_step(case["x"], k, 1.0, opt, objective="iwae_elbo", eps=case["eps1"], eps2=case["eps2"])
        elif kind == "dreg":
            res = m.train_step(case["x"], k, 1.0, opt, eps=case["eps1"])
        elif kind == "ciwae":
            res = m.train_step(case["x"], case["y"], k, 1.0, opt, objective="iwae_elbo", eps=case["eps1"])
        else:
            res = m.train_step(case["x"], k, 1.0, opt, objective="iwae_elbo", eps=case["eps1"])
        _, fn, args = oracle_call(oracle, kind, case, k, 1.0)
        if kind == "dreg":
            g, ref = oracle.grads_dreg(p, *args)
        else:
            g, ref = oracle.grads_objective(fn, p, "iwae_elbo", *args)
        assert rel(float(res["iwae_elbo"]), float(ref["iwae_elbo"])) < RTOL, t
        newp = {}
        for n in p:
            th, ms[n], vs[n] = oracle.keras_adam_step(p[n].detach(), g[n], ms[n], vs[n], t, 1e-3)
            newp[n] = th.clone().requires_grad_(True)
        p = newp
    got = m.get_weights()
    for n in p:
        # Adam's sign-like update amplifies tiny gradient differences where |g| ~ eps; bound by a few lr
        assert np.max(np.abs(got[n] - p[n].detach().numpy())) < 4e-4, n
    assert opt.iterations == steps


def test_sir_bit_exact_vs_numpy(oracle):
    """tasks/plot_task01.py:85-88: scipy softmax -> np.random.choice(L, 200, replace=False, p=al)."""
    from iwae_b200 import posterior
    for seed in range(6):
        rs = np.random.RandomState(seed)
        L = [10000, 777, 10000][seed % 3]
        size = min(200, L // 3)
        lw = (rs.randn(L) * (1 + seed)).astype(np.float32) - 400
        al_dev = posterior.softmax_f32(lw)
        al_ref = oracle.softmax_f32(lw)
        assert rel(al_dev.cpu().numpy(), al_ref) < 1e-5
        # resampling itself: identical weights in -> identical indices out, bit for bit
        want = np.random.RandomState(50 + seed).choice(np.arange(L), size=size, replace=False, p=al_ref)
        rs2 = np.random.RandomState(50 + seed)
        idx = posterior.sir_from_randomstate(al_ref, size, rs2)
        assert np.array_equal(idx.cpu().numpy(), want)
        rs3 = np.random.RandomState(50 + seed)
        rs3.choice(np.arange(L), size=size, replace=False, p=al_ref)
        assert rs2.random_sample() == rs3.random_sample()      # RandomState left in the same position


def test_philox_noise_matches_restatement(oracle):
    import ctypes as C
    from iwae_b200._abi import lib, check
    for (B, k, d, off, stream) in [(3, 40, 100, 0, 1), (2, 7, 50, 9998, 2), (1, 5, 2, 17, 1)]:
        out = torch.empty(k, B, d, device="cuda")
        check(lib.iwae_philox_normal(123, stream, off, B, k, d, C.c_void_p(out.data_ptr()),
                                     torch.cuda.current_stream().cuda_stream))
        ref = oracle.philox_normal(123, stream, off, B, k, d)
        assert np.max(np.abs(out.cpu().numpy() - ref)) < 2e-5


def test_decoder_tools(oracle):
    case = make_case(oracle, "iwae1", 200, 2, 2, 9, seed=41)
    m = build("iwae1", dict(h=200, d=2), case)
    z = np.random.RandomState(0).randn(9, 2, 2).astype(np.float32)
    logits, pxz = m.decoder(z)
    p = oracle.to_torch(case["params"])
    ref = oracle.mlp3(p, "decoder.decode_z_to_x", torch.tensor(z, dtype=torch.float64))
    assert rel(logits.numpy(), ref.numpy()) < RTOL
    lp = m._decode(z, None, case["x"], want_logits=False)[1]
    ref_lp = oracle.bernoulli_log_prob(ref, torch.tensor(case["x"], dtype=torch.float64)).sum(-1)
    assert rel(lp.numpy(), ref_lp.numpy()) < RTOL
    zz, qzx = m.encoder(case["x"], 4)
    mu, sd = oracle.basic_block(p, "encoder.encode_x_to_z", torch.tensor(case["x"], dtype=torch.float64))
    assert rel(qzx.loc.numpy(), mu.numpy()) < RTOL and rel(qzx.scale.numpy(), sd.numpy()) < RTOL
    assert zz.shape == (4, 2, 2)
    xs, probs = m.sample(z[:, 0])               # (x_sample, probs), iwae1.py:178
    assert probs.shape == (9, 784) and set(np.unique(xs.numpy())) <= {0.0, 1.0}


def test_val_step_full_test_set_shape(oracle):
    """val_step on a test-set-sized batch (main.py:152: all test images at k=n_samples).  B=2,000 here keeps
    the fp64 oracle to seconds; the kernels are the same at B=10,000."""
    B, k = 2000, 20
    P, X, _ = oracle.synthetic_data(B, seed=7)
    rs = np.random.RandomState(3)
    shapes = oracle.param_shapes("iwae2", [200, 100], [100, 50])
    params = oracle.glorot_params(shapes, rs, out_bias=oracle.get_bias(P.mean(0)), scale=0.5)
    import iwae_b200 as ib
    m = ib.IWAE2([200, 100], [100, 50])
    m.set_weights(params)
    eps1 = rs.randn(k, B, 100).astype(np.float32)
    eps2 = rs.randn(k, B, 50).astype(np.float32)
    res = m(X, k, 1.0, eps=eps1, eps2=eps2)          # injected noise, fp32 kernels: comparable with the oracle value by value
    p = oracle.to_torch(params)
    ref = oracle.iwae2_call(p, torch.tensor(X, dtype=torch.float64), k, 1.0, torch.tensor(eps1, dtype=torch.float64),
                            torch.tensor(eps2, dtype=torch.float64))
    for key in ("vae_elbo", "iwae_elbo", "iwae_eq14"):
        assert rel(float(res[key]), float(ref[key])) < RTOL, key
    assert rel(res["iwae_elbo_per_example"].numpy(), ref["iwae_elbo_per_example"].numpy()) < RTOL
    assert res["lpxz1"].shape == (k, B) and res["snis_z1"].shape == (B, 100)
    # the monitoring call itself (Philox noise) runs and returns finite scalars
    mon = m.val_step(X, k, 1.0)
    assert np.isfinite(float(mon["iwae_elbo"])) and float(mon["iwae_elbo"]) > float(mon["vae_elbo"])


def test_sharded_eval_single_rank(oracle):
    """iwae_b200.dist.model_llh_sharded without a process group == one rank: mean of the per-example LLH."""
    from iwae_b200.dist import model_llh_sharded
    case = make_case(oracle, "iwae2", [200, 100], [100, 50], 6, 1, seed=71)
    m = build("iwae2", dict(h=[200, 100], d=[100, 50]), case)
    mean, per = model_llh_sharded(m, case["x"], 200, precision="tc_f16_fast", seed=5)
    assert per.shape == (6,) and abs(float(mean) - float(per.double().mean())) < 1e-9 * abs(float(mean))
