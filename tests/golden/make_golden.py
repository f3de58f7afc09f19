"""Writes the committed parity fixtures tests/golden/case_*.npz.

The reference (TF2/TFP/Keras) cannot be imported in this image, so the fixtures come from the fp64
oracle restatement (oracle/iwae_oracle.py), which is itself pinned by tests/test_oracle.py.  Each
fixture holds ONLY what is needed to re-create the inputs (a case name -> seeds and shapes, below) and
the expected outputs; weights, x and the injected noise are regenerated from the seeds.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)

GRAD_STRIDE = 211

CASES = {
    # BASELINE.json configs[0]: 1 stochastic layer k=5 batch 20 (d=100 per main.py:85, d=50 per the paper)
    "c1_iwae1_d100_k5_b20": dict(kind="iwae1", h=200, d=100, B=20, k=5, seed=1, scale=1.0),
    "c1_iwae1_d50_k5_b20": dict(kind="iwae1", h=200, d=50, B=20, k=5, seed=2, scale=1.0),
    # configs[1]: 2 stochastic layers k=50 batch 20
    "c2_iwae2_k50_b20": dict(kind="iwae2", h=[200, 100], d=[100, 50], B=20, k=50, seed=3, scale=1.0),
    # spread importance weights (Glorot x0.5: max normalised weight ~0.4, like a trained model; at plain
    # Glorot init one sample dominates each image)
    "iwae2_k50_b4_spread": dict(kind="iwae2", h=[200, 100], d=[100, 50], B=4, k=50, seed=4, scale=0.5),
    # task01 / task03 latent sizes
    "iwae1_d2_k64_b3": dict(kind="iwae1", h=200, d=2, B=3, k=64, seed=5, scale=1.0),
    "iwae2_d4_2_k16_b3": dict(kind="iwae2", h=[200, 100], d=[4, 2], B=3, k=16, seed=6, scale=1.0),
    # configs[3]: DReG k=50
    "c4_dreg_k50_b20": dict(kind="dreg", h=200, d=100, B=20, k=50, seed=7, scale=1.0),
    # configs[4]: conditional IWAE
    "c5_ciwae_k50_b20": dict(kind="ciwae", h=200, d=100, B=20, k=50, seed=8, scale=1.0),
    "ciwae_std_k8_b5": dict(kind="ciwae_std", h=200, d=100, B=5, k=8, seed=9, scale=1.0),
}


def run_case(oracle, c):
    from conftest import make_case
    kind = c["kind"]
    okind = "iwae1" if kind == "dreg" else kind
    case = make_case(oracle, okind, c["h"], c["d"], c["B"], c["k"], seed=c["seed"], scale=c["scale"],
                     n_cond=10 if kind.startswith("ciwae") else 0)
    dt = torch.float64
    p = oracle.to_torch(case["params"], dt, requires_grad=True)
    x = torch.tensor(case["x"], dtype=dt)
    e1 = torch.tensor(case["eps1"], dtype=dt)
    out = {}
    if kind == "iwae1":
        g, r = oracle.grads_objective(oracle.iwae1_call, p, "iwae_elbo", x, c["k"], 1.0, e1)
        keys = ["lpxz", "lpz", "lqzx", "al", "snis_z", "iwae_elbo_per_example", "vae_elbo_per_example", "kl"]
        scal = ["vae_elbo", "vae_elbo_kl", "iwae_elbo", "iwae_eq14"]
    elif kind == "iwae2":
        e2 = torch.tensor(case["eps2"], dtype=dt)
        g, r = oracle.grads_objective(oracle.iwae2_call, p, "iwae_elbo", x, c["k"], 1.0, e1, e2)
        keys = ["lpxz1", "lpz1z2", "lpz2", "lqz1x", "lqz2z1", "al", "snis_z1", "snis_z2", "iwae_elbo_per_example",
                "vae_elbo_per_example"]
        scal = ["vae_elbo", "iwae_elbo", "iwae_eq14"]
    elif kind == "dreg":
        g, r = oracle.grads_dreg(p, x, c["k"], 1.0, e1)
        keys = ["lpxz", "lpz", "lqzx", "snis_z", "iwae_elbo_per_example"]
        scal = ["iwae_elbo", "inference_loss"]
    else:
        learned = kind == "ciwae"
        fn = lambda p_, x_, y_, k_, b_, e_: oracle.ciwae_call(p_, x_, y_, k_, b_, e_, learned_prior=learned)
        g, r = oracle.grads_objective(fn, p, "iwae_elbo", x, case["y"], c["k"], 1.0, e1)
        keys = ["lpxzy", "lpzy" if learned else "lpz", "lqzxy", "al", "snis_z", "iwae_elbo_per_example", "kl"]
        scal = ["vae_elbo", "vae_elbo_kl", "iwae_elbo", "iwae_eq14"]
    for k_ in keys + scal:
        out[k_] = r[k_].detach().numpy()
    flat = torch.cat([v.reshape(-1) for v in g.values()]).detach().numpy()
    # gradient digest (the full ~0.5M-float gradient is compared against the live oracle in the GPU
    # tests): per-tensor L2 norms and every 211th entry of the flat gradient
    out["grad_norms"] = np.array([float(v.norm()) for v in g.values()])
    out["grad_sample"] = flat[::GRAD_STRIDE].copy()
    return case, out


def main():
    from oracle import iwae_oracle as oracle
    for name, c in CASES.items():
        _, out = run_case(oracle, c)
        save = {"out_" + k: v for k, v in out.items()}
        np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), name=name, **save)
        print(name, float(out["iwae_elbo"]))


if __name__ == "__main__":
    main()
