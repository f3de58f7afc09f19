"""GPU: the two opt-in executions of the train step - IWAE_MEGA=1 (the whole step as one persistent cooperative kernel,
kernels_f32.cu "train-step megakernel") and IWAE_CHAIN=1 (a BasicBlock's Dense layers in one launch, dense_chain_kernel).
Both are switched by an environment variable the library reads once, so each runs in a child process that trains a few
steps and prints its parameters' checksum and last ELBO; the default multi-kernel step is the reference.
The megakernel runs the SAME *_body code for every virtual block: results must be bit-identical.  The chain kernel adds
the products in a different order: 1e-5 relative on the parameters after 6 steps."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import json, sys, numpy as np, torch
sys.path.insert(0, %r)
import iwae_b200 as ib
kind = sys.argv[1]
P, X, Y = ib.data.synthetic_binarised(120, seed=5)
bias = ib.data.get_bias(P.mean(0))
if kind == "iwae2": m = ib.IWAE2([200, 100], [100, 50], out_bias=bias, seed=11)
elif kind == "iwae1": m = ib.IWAE(200, 50, out_bias=bias, seed=11)
else: m = ib.CIWAE(200, 50, out_bias=bias, seed=11)
opt = ib.Adam(1e-3, epsilon=1e-4)
if kind == "dreg": m = ib.IWAEDReG(200, 50, out_bias=bias, seed=11)
elbos = []
for i in range(6):       # the first step is eager, the second captures the CUDA graph, the rest replay it
    xb = torch.from_numpy(X[i * 20:(i + 1) * 20]).cuda()
    yb = torch.from_numpy(Y[i * 20:(i + 1) * 20].astype(np.float32)).cuda()
    if kind == "ciwae": res = m.train_step(xb, yb, 10, 1.0, opt, objective="iwae_elbo", seed=100 + i)
    elif kind == "dreg": res = m.train_step(xb, 10, 1.0, opt, seed=100 + i)
    else: res = m.train_step(xb, 10, 1.0, opt, objective="iwae_elbo", seed=100 + i)
    elbos.append(float(res["iwae_elbo"]))
torch.cuda.synchronize()
p = m.params.detach().cpu().numpy()
print("RESULT " + json.dumps({"elbos": elbos, "sum": float(p.astype(np.float64).sum()), "params": p[::97].tolist()}))
""" % ROOT


def run_child(kind, **env):
    e = dict(os.environ)
    for k in ("IWAE_MEGA", "IWAE_CHAIN", "IWAE_RECOMPUTE"):
        e.pop(k, None)
    e.update(env)
    out = subprocess.run([sys.executable, "-c", CHILD, kind], env=e, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


@pytest.mark.parametrize("kind", ["iwae2", "iwae1", "ciwae", "dreg"])
def test_megakernel_step_is_bit_identical(kind):
    ref = run_child(kind)
    got = run_child(kind, IWAE_MEGA="1")
    assert got["elbos"] == ref["elbos"]
    assert got["sum"] == ref["sum"] and got["params"] == ref["params"]


@pytest.mark.parametrize("kind", ["iwae2", "iwae1", "ciwae", "dreg"])
def test_recompute_backward_is_bit_identical(kind):
    """IWAE_RECOMPUTE=1: the forward of the train step does not write the [k*B, 784] logits and the backward runs the
    output layer's GEMM again with the gradient in its epilogue (abi.cu, backward_internal; north_star "recomputes decoder
    activations instead of storing them").  Same tiles, same order: parameters after six steps equal those from stored
    logits (IWAE_RECOMPUTE=0) bit for bit - also inside the megakernel."""
    ref = run_child(kind, IWAE_RECOMPUTE="0")
    got = run_child(kind, IWAE_RECOMPUTE="1")
    assert got["elbos"] == ref["elbos"]
    assert got["sum"] == ref["sum"] and got["params"] == ref["params"]
    if kind == "iwae2":
        mega = run_child(kind, IWAE_RECOMPUTE="1", IWAE_MEGA="1")
        assert mega["sum"] == ref["sum"] and mega["params"] == ref["params"]


def test_chain_kernel_step_matches():
    ref = run_child("iwae2")
    got = run_child("iwae2", IWAE_CHAIN="1")
    a, b = np.array(ref["params"]), np.array(got["params"])
    assert np.max(np.abs(a - b)) <= 1e-5 * max(1.0, float(np.max(np.abs(a))))
    assert np.allclose(ref["elbos"], got["elbos"], rtol=1e-5)
