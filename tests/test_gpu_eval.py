"""GPU: the k-sample test log-likelihood estimator (main.py:170-184) - fp32 parity mode and the fused
tcgen05 kernel - against the oracle fed with the SAME Philox noise (oracle.philox_normal restates the
device generator), plus size-independent properties at BASELINE.json's k=5000."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import make_case
from test_gpu_parity import build, rel, RTOL, TC_RTOL

pytestmark = pytest.mark.gpu
SEED = 20251018


def oracle_llh(oracle, kind, case, k, img_offset=0, dtype=torch.float64):
    p = oracle.to_torch(case["params"], dtype)
    x = torch.tensor(case["x"], dtype=dtype)
    B = x.shape[0]
    if kind == "iwae2":
        d1, d2 = case["d"]
        e1 = torch.tensor(oracle.philox_normal(SEED, 1, img_offset, B, k, d1), dtype=dtype)
        e2 = torch.tensor(oracle.philox_normal(SEED, 2, img_offset, B, k, d2), dtype=dtype)
        r = oracle.iwae2_call(p, x, k, 1.0, e1, e2)
    elif kind == "iwae1":
        e1 = torch.tensor(oracle.philox_normal(SEED, 1, img_offset, B, k, case["d"]), dtype=dtype)
        r = oracle.iwae1_call(p, x, k, 1.0, e1)
    else:
        e1 = torch.tensor(oracle.philox_normal(SEED, 1, img_offset, B, k, case["d"]), dtype=dtype)
        r = oracle.ciwae_call(p, x, case["y"], k, 1.0, e1, learned_prior=(kind == "ciwae"))
    return r["iwae_elbo_per_example"].numpy()


def _case(oracle, kind, B, scale=1.0, seed=51):
    h, d = ([200, 100], [100, 50]) if kind == "iwae2" else (200, 100)
    case = make_case(oracle, kind, h, d, B, 1, seed=seed, scale=scale, n_cond=10 if kind.startswith("ciwae") else 0)
    case["d"] = d
    case["h"] = h
    return case


@pytest.mark.parametrize("N,K", [(16, 16), (112, 112), (96, 208), (256, 64), (48, 100), (112, 7)])
def test_umma_descriptor_selftest(N, K):
    """One tcgen05.mma chain through the exact operand staging of the fused kernel."""
    from iwae_b200._abi import diag_lib, check_diag as check
    lib = diag_lib()               # libiwae_b200_diag.so: the unit-test hook is not part of the product library
    rs = np.random.RandomState(N * 1000 + K)
    A = rs.randn(128, K).astype(np.float32)
    Bm = rs.randn(N, K).astype(np.float32)
    tA, tB = torch.tensor(A, device="cuda"), torch.tensor(Bm, device="cuda")
    D = torch.zeros(128, N, device="cuda")
    check(lib.iwae_umma_selftest(C.c_void_p(tA.data_ptr()), C.c_void_p(tB.data_ptr()), N, K, 0,
                                 C.c_void_p(D.data_ptr()), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = A.astype(np.float16).astype(np.float64) @ Bm.astype(np.float16).astype(np.float64).T
    assert np.max(np.abs(D.cpu().numpy() - ref)) < 2e-3 * np.sqrt(K)


@pytest.mark.parametrize("kind", ["iwae1", "iwae2", "ciwae", "ciwae_std"])
def test_llh_fp32_matches_oracle(oracle, kind):
    N, k = 3, 300
    case = _case(oracle, kind, N)
    m = build(kind, case, case)
    y = case["y"] if kind.startswith("ciwae") else None
    per, total = m.llh(case["x"], k, y=y, precision="fp32", seed=SEED, img_offset=7)
    ref = oracle_llh(oracle, kind, case, k, img_offset=7)
    assert rel(per.cpu().numpy(), ref) < RTOL
    assert abs(float(total) - ref.sum()) < RTOL * abs(ref.sum())


@pytest.mark.parametrize("precision", ["tc_f16", "tc_f16_fast"])
@pytest.mark.parametrize("kind,scale", [("iwae2", 1.0), ("iwae2", 0.5), ("iwae1", 1.0), ("iwae1", 0.5), ("ciwae", 1.0),
                                        ("ciwae_std", 1.0)])
def test_llh_tensor_core_matches_oracle(oracle, kind, scale, precision):
    """Fused tcgen05 kernel (fp16 operands, fp32 accumulate).  Stated tolerance: TC_RTOL relative on
    the per-example LLH."""
    N, k = 5, 300       # 3 tiles of 100 rows per image: exercises ragged tiles and the cross-tile combine
    case = _case(oracle, kind, N, scale=scale)
    m = build(kind, case, case)
    y = case["y"] if kind.startswith("ciwae") else None
    per, total = m.llh(case["x"], k, y=y, precision=precision, seed=SEED, img_offset=3)
    torch.cuda.synchronize()
    ref = oracle_llh(oracle, kind, case, k, img_offset=3)
    err = rel(per.cpu().numpy(), ref)
    print(f"\n[tc parity] {kind} scale={scale} {precision}: max rel err {err:.3e}")
    assert err < TC_RTOL
    assert abs(float(total) - float(per.double().sum())) < 1e-6 * abs(float(total))


@pytest.mark.parametrize("kind,d", [("iwae1", 2), ("iwae1", 50), ("iwae2", [4, 2]), ("iwae2", [2, 2])])
@pytest.mark.parametrize("k", [1, 5, 129])
def test_llh_tensor_core_small_latents(oracle, kind, d, k):
    h = [200, 100] if kind == "iwae2" else 200
    case = make_case(oracle, kind, h, d, 3, 1, seed=61)
    case["d"], case["h"] = d, h
    m = build(kind, case, case)
    per, _ = m.llh(case["x"], k, precision="tc_f16", seed=SEED)
    ref = oracle_llh(oracle, kind, case, k)
    assert rel(per.cpu().numpy(), ref) < TC_RTOL


def test_llh_edge_cases(oracle):
    case = _case(oracle, "iwae2", 4)
    m = build("iwae2", case, case)
    for prec in ("fp32", "tc_f16_fast"):
        per, total = m.llh(case["x"][:0], 50, precision=prec, seed=SEED)
        assert per.numel() == 0 and float(total) == 0.0
    # host-buffer entry point == device entry point
    per_d, tot_d = m.llh(case["x"], 200, precision="tc_f16_fast", seed=SEED)
    per_h, tot_h = m.llh_host(case["x"], 200, precision="tc_f16_fast", seed=SEED)
    assert np.array_equal(per_d.cpu().numpy(), per_h) and float(tot_d) == tot_h


def test_llh_sharding_invariance(oracle):
    """Noise is keyed on the global image index: evaluating shards with img_offset reproduces the
    single-pass per-example values bit for bit (what makes the 1/2/4/8-GPU runs identical)."""
    case = _case(oracle, "iwae2", 12)
    m = build("iwae2", case, case)
    for prec in ("fp32", "tc_f16_fast"):
        whole, tot = m.llh(case["x"], 150, precision=prec, seed=SEED)
        parts = [m.llh(case["x"][a:b], 150, precision=prec, seed=SEED, img_offset=a)[0] for a, b in ((0, 5), (5, 6), (6, 12))]
        assert torch.equal(whole, torch.cat(parts)), prec


def test_llh_full_size_properties(oracle):
    """BASELINE.json k=5000 on a bounded set of images: the tensor-core fast path agrees with the fp32
    FFMA path per example (same Philox noise), LLH rises with k (Jensen), and the fp64 sum is the sum
    of the per-example values."""
    case = _case(oracle, "iwae2", 24, scale=0.5)
    m = build("iwae2", case, case)
    per_tc, tot = m.llh(case["x"], 5000, precision="tc_f16_fast", seed=SEED)
    per_32, _ = m.llh(case["x"], 5000, precision="fp32", seed=SEED)
    err = rel(per_tc.cpu().numpy(), per_32.cpu().numpy())
    print(f"\n[k=5000] tc_f16_fast vs fp32: max rel err {err:.3e}")
    assert err < TC_RTOL
    assert abs(float(tot) - float(per_tc.double().sum())) < 1e-6 * abs(float(tot))
    per_small, _ = m.llh(case["x"], 50, precision="tc_f16_fast", seed=SEED)
    assert float(per_tc.mean()) > float(per_small.mean())
    # one image against the fp64 oracle at the full k
    ref = oracle_llh(oracle, "iwae2", {**case, "x": case["x"][:2]}, 5000)
    assert rel(per_tc[:2].cpu().numpy(), ref) < TC_RTOL


@pytest.mark.parametrize("precision", ["fp32", "tc_f16_fast"])
@pytest.mark.parametrize("kind", ["iwae1", "iwae2"])
def test_monitor_pass_matches_oracle(oracle, kind, precision):
    """Fused monitoring pass (val_step's iwae_elbo / vae_elbo on a test set, main.py:152) at the training k."""
    N, k = 7, 50
    case = _case(oracle, kind, N, scale=0.5)
    m = build(kind, case, case)
    res = m.monitor(case["x"], k, precision=precision, seed=SEED)
    p = oracle.to_torch(case["params"])
    x = torch.tensor(case["x"], dtype=torch.float64)
    if kind == "iwae2":
        e1 = torch.tensor(oracle.philox_normal(SEED, 1, 0, N, k, 100), dtype=torch.float64)
        e2 = torch.tensor(oracle.philox_normal(SEED, 2, 0, N, k, 50), dtype=torch.float64)
        ref = oracle.iwae2_call(p, x, k, 1.0, e1, e2)
    else:
        e1 = torch.tensor(oracle.philox_normal(SEED, 1, 0, N, k, 100), dtype=torch.float64)
        ref = oracle.iwae1_call(p, x, k, 1.0, e1)
    tol = RTOL if precision == "fp32" else TC_RTOL
    assert rel(res["iwae_elbo_per_example"].numpy(), ref["iwae_elbo_per_example"].numpy()) < tol
    assert rel(res["vae_elbo_per_example"].numpy(), ref["vae_elbo_per_example"].numpy()) < tol
    assert rel(float(res["iwae_elbo"]), float(ref["iwae_elbo"])) < tol
    assert rel(float(res["vae_elbo"]), float(ref["vae_elbo"])) < tol


@pytest.mark.parametrize("seed", range(8))
def test_random_architectures_both_paths(oracle, seed):
    """Random layer sizes / k / N inside the tensor-core path's limits (h <= 207, d1 + n_cond + 1 <= 112, d2 <= 56):
    fp32 and fused tensor-core LLH against the oracle on the same Philox noise."""
    rs = np.random.RandomState(900 + seed)
    two = bool(seed % 2)
    h1 = int(rs.choice([16, 37, 64, 128, 200, 207]))
    d1 = int(rs.choice([1, 2, 3, 8, 17, 50, 64, 100]))
    N, k = int(rs.randint(1, 6)), int(rs.choice([1, 2, 31, 128, 129, 257]))
    if two:
        h2, d2 = int(rs.choice([8, 33, 100, 150])), int(rs.choice([1, 2, 7, 24, 50, 56]))
        kind, h, d = "iwae2", [h1, h2], [d1, d2]
    else:
        kind, h, d = "iwae1", h1, d1
    case = make_case(oracle, kind, h, d, N, 1, seed=70 + seed, scale=0.7)
    case["d"], case["h"] = d, h
    m = build(kind, case, case)
    ref = oracle_llh(oracle, kind, case, k, img_offset=5)
    for prec, tol in (("fp32", RTOL), ("tc_f16", TC_RTOL), ("tc_f16_fast", TC_RTOL)):
        per, _ = m.llh(case["x"], k, precision=prec, seed=SEED, img_offset=5)
        assert rel(per.cpu().numpy(), ref) < tol, (kind, h, d, N, k, prec, rel(per.cpu().numpy(), ref))


def test_auto_precision_falls_back_to_fp32_kernels_outside_the_fused_kernels_limits(oracle):
    """precision="auto" (the default) uses the fused tensor-core kernel when the architecture fits it and the fp32
    CUDA kernels otherwise (h = 300 here) - never an error, never the CPU."""
    case = make_case(oracle, "iwae1", 300, 20, 3, 1, seed=77, scale=0.7)
    case["d"], case["h"] = 20, 300
    m = build("iwae1", case, case)
    assert not m.tc_supported()
    per, _ = m.llh(case["x"], 64, seed=SEED)
    per32, _ = m.llh(case["x"], 64, precision="fp32", seed=SEED)
    assert torch.equal(per, per32)
    assert rel(per.cpu().numpy(), oracle_llh(oracle, "iwae1", case, 64)) < RTOL
    with pytest.raises(Exception):
        m.llh(case["x"], 64, precision="tc_f16_fast", seed=SEED)
    small = _case(oracle, "iwae2", 2)
    assert build("iwae2", small, small).tc_supported()


@pytest.mark.parametrize("precision", ["tc_f16", "tc_f16_fast"])
@pytest.mark.parametrize("kind", ["iwae1", "iwae2", "ciwae", "ciwae_std"])
def test_val_step_on_the_fused_engine_matches_oracle(oracle, kind, precision):
    """model.val_step (src/iwae1.py:164-166, main.py:152) through the fused tensor-core kernel (iwae_forward_tc): every
    entry of the result dict except logits against the fp64 oracle fed with the same Philox noise, at a k that leaves
    tiles ragged (k = 50: 50 of 128 rows) and one that spans tiles (k = 300); beta != 1 for the one-layer models."""
    for k, beta in ((50, 1.0), (300, 0.7)):
        N = 5
        case = _case(oracle, kind, N, scale=0.7)
        m = build(kind, case, case)
        p = oracle.to_torch(case["params"])
        x = torch.tensor(case["x"], dtype=torch.float64)
        e1 = torch.tensor(oracle.philox_normal(SEED, 1, 0, N, k, 100), dtype=torch.float64)
        if kind == "iwae2":
            e2 = torch.tensor(oracle.philox_normal(SEED, 2, 0, N, k, 50), dtype=torch.float64)
            ref = oracle.iwae2_call(p, x, k, beta, e1, e2)
            res = m.call(case["x"], k, beta, seed=SEED, precision=precision)
            keys = ["lpxz1", "lpz1z2", "lpz2", "lqz1x", "lqz2z1", "al", "snis_z1", "snis_z2", "vae_elbo", "iwae_elbo", "iwae_eq14"]
        elif kind == "iwae1":
            ref = oracle.iwae1_call(p, x, k, beta, e1)
            res = m.call(case["x"], k, beta, seed=SEED, precision=precision)
            keys = ["lpxz", "lpz", "lqzx", "al", "snis_z", "vae_elbo", "vae_elbo_kl", "iwae_elbo", "iwae_eq14"]
        else:
            learned = kind == "ciwae"
            ref = oracle.ciwae_call(p, x, case["y"], k, beta, e1, learned_prior=learned)
            res = m.call(case["x"], case["y"], k, beta, seed=SEED, precision=precision)
            keys = ["lpxzy", "lpzy" if learned else "lpz", "lqzxy", "al", "snis_z", "vae_elbo", "vae_elbo_kl", "iwae_elbo", "iwae_eq14"]
        for key in keys + ["iwae_elbo_per_example", "vae_elbo_per_example", "log_w"]:
            got, want = res[key].numpy(), ref[key].detach().numpy()
            # stated tolerances: log_w and every ELBO <= 1e-4 relative; a SINGLE log-density term of a row <= 3e-4;
            # softmax weights `al` and SNIS means: 1e-2 ABSOLUTE (rel() floors the denominator at 1) - al = softmax(log_w)
            # turns an absolute log_w error of 0.02 nats (5e-5 of |log_w| = 450, inside the tolerance) into 2 % of a weight
            tol = 1e-2 if key in ("al", "snis_z", "snis_z1", "snis_z2") else (3e-4 if key.startswith(("lp", "lq")) else TC_RTOL)
            assert got.shape == want.shape, (key, got.shape, want.shape)
            assert rel(got, want) < tol, (kind, k, key, rel(got, want))
        # the lazily recomputed entries use the same seed on the fp32 kernels
        zk = "z1" if kind == "iwae2" else "z"
        assert rel(res[zk].numpy(), ref[zk].detach().numpy()) < RTOL
        assert rel(res["logits"].numpy(), ref["logits"].detach().numpy()) < RTOL


def test_val_step_full_test_set_shape(oracle):
    """The driver's monitoring call: val_step(Xtest [10000,784], 50, 1.0) (main.py:152).  Fused engine vs the fp32
    kernels on the same Philox noise at the full size (sampled), and the batch scalars."""
    import iwae_b200 as ib
    P, X, _ = ib.data.synthetic_binarised(10000, seed=123)
    m = ib.IWAE2([200, 100], [100, 50], seed=3, out_bias=ib.data.get_bias(P.mean(0)))
    a = m.val_step(X, 50, 1.0)
    assert a["al"].shape == (50, 10000) and a["snis_z1"].shape == (10000, 100) and a["lpxz1"].shape == (50, 10000)
    b = m.call(X[:512], 50, 1.0, seed=a_seed(a), precision="fp32")
    for key in ("lpxz1", "lpz1z2", "lpz2", "lqz1x", "lqz2z1", "log_w", "iwae_elbo_per_example", "vae_elbo_per_example"):
        ga, gb = a[key].numpy()[..., :512], b[key].numpy()
        err = rel(ga, gb)
        mean_err = abs(float(ga.mean()) - float(gb.mean())) / abs(float(gb.mean()))
        print(f"[val_step 10000 x 50] {key}: worst entry rel err vs fp32 kernels {err:.2e}, mean over the set {mean_err:.2e}")
        # Stated tolerances of the fused engine (fp16 GEMM operands): the per-example ELBOs and the test-set MEAN of every
        # term (the scalars the driver logs, src/iwae2.py:315-323) <= 1e-4 relative (measured 1e-6); an image's k=50 ELBO <= 3e-4
        # (measured 1.3e-4 worst of 512); a single row's term / log_w <= 2e-3 in
        # the worst of 25,600 rows (rows with |z2| >> 1, whose importance weight is e^-1000: z2 = mu2 + e^a eps with
        # mu2, a from fp16 GEMMs).  precision="fp32" gives the exact parity kernels.
        assert mean_err < TC_RTOL, key
        assert err < (3e-4 if key.endswith("per_example") else 2e-3), key       # k=50: one or two rows dominate an image's ELBO
    assert abs(float(a["iwae_elbo"]) - float(a["iwae_elbo_per_example"].numpy().mean())) < 1e-3


def a_seed(res):
    """the noise seed a result was drawn with (kept for the lazily recomputed entries)"""
    for cell in (res.__dict__.get("_seed"),):
        if cell is not None:
            return cell
    raise AssertionError("result carries no seed")


@pytest.mark.parametrize("kind,scale,B", [("iwae2", 1.0, 300), ("iwae1", 3.0, 129), ("ciwae", 1.0, 77), ("iwae2", 0.5, 5)])
def test_tensor_core_encoder_matches_fp32_kernels(oracle, kind, scale, B):
    """q(z1|x) on tcgen05 (csrc/encoder_tc.cu: the pre-pass of the fused engine; fp16 hi + scaled-lo operand pairs, three MMAs
    per product into two fp32 accumulators) against the fp32 kernels AND the fp64 oracle: mu and std of every image, incl. a partly
    filled last tile of 128 images and the label-conditioned input [x | onehot(y)] (src/iwae1.py:24-60)."""
    case = _case(oracle, kind, B, scale=scale)
    m = build(kind, case, case)
    y = case.get("y") if kind.startswith("ciwae") else None
    want = ("q_mu", "q_std", "log_w")
    f32 = m._forward(case["x"], y, 3, 1.0, want, seed=SEED)
    tcr = m._forward_tc(case["x"], y, 3, 1.0, want, seed=SEED)
    p = oracle.to_torch(case["params"], torch.float64)
    x64 = torch.tensor(case["x"], dtype=torch.float64)
    if kind.startswith("ciwae"):
        yoh = torch.nn.functional.one_hot(torch.tensor(np.asarray(y), dtype=torch.int64), 10).to(torch.float64)
        x64 = torch.cat([x64, yoh], dim=1)
    prefix = "encoder.encode_x_to_z1" if kind == "iwae2" else "encoder.encode_x_to_z"
    mu64, sd64 = oracle.basic_block(p, prefix, x64)
    for key, ref in (("q_mu", mu64), ("q_std", sd64)):
        a, b, r = tcr[key].cpu().numpy().astype(np.float64), f32[key].cpu().numpy().astype(np.float64), ref.numpy()
        # errors relative to the tensor's scale (mu has entries near zero; their absolute error is what moves z1 = mu + std*eps)
        den = float(np.max(np.abs(r)))
        e_tc, e_32 = float(np.max(np.abs(a - r))) / den, float(np.max(np.abs(b - r))) / den
        print(f"\n[tc encoder] {kind} scale={scale} B={B} {key}: tensor-core {e_tc:.2e}, fp32 kernels {e_32:.2e} of max |{key}| vs fp64 oracle")
        # measured: tensor-core 2.3e-7 .. 9.2e-7, fp32 kernels 2.1e-7 .. 6.6e-7 (both 5.7e-6 for std at Glorot x3: std = exp(.)
        # turns the absolute error of a large pre-activation into a relative one): fp16 hi + scaled-lo pairs carry 22 bits
        assert e_32 < 1e-5 and e_tc < 1e-5 and e_tc < 2.0 * max(e_32, 5e-7), (key, e_tc, e_32)
