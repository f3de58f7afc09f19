import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import iwae_oracle
    return iwae_oracle


def make_case(oracle, kind, n_hidden, n_latent, B, k, seed=0, scale=1.0, n_cond=0):
    """Seeded weights + inputs + injected noise shared by the oracle and the CUDA path."""
    rs = np.random.RandomState(1000 + seed)
    P, X, y = oracle.synthetic_data(max(B, 64), seed=123)
    shapes = oracle.param_shapes(kind, n_hidden, n_latent, n_cond=n_cond)
    params = oracle.glorot_params(shapes, rs, out_bias=oracle.get_bias(P.mean(0)), scale=scale)
    d1 = n_latent[0] if isinstance(n_latent, (list, tuple)) else n_latent
    eps1 = rs.randn(k, B, d1).astype(np.float32)
    eps2 = None
    if isinstance(n_latent, (list, tuple)):
        eps2 = rs.randn(k, B, n_latent[1]).astype(np.float32)
    return dict(params=params, x=X[:B].copy(), y=y[:B].astype(np.float32), eps1=eps1, eps2=eps2)
