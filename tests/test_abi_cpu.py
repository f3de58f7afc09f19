"""CPU: the C-ABI library loads, exports every symbol include/iwae_b200.h declares, and its
device-free host logic (parameter layout) agrees with the oracle.  No compute calls without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "iwae_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(iwae_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from iwae_b200 import _abi
    names = _header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(_abi.lib, n), f"libiwae_b200.so does not export {n}"
    assert set(names) <= set(_abi.EXPORTS) | {"iwae_layout_for_config"}
    assert _abi.lib.iwae_version() >= 100


def test_diagnostics_library_exports_every_header_symbol():
    """libiwae_b200_diag.so (tensor-core unit test hook and the micro-benchmarks behind profiles/r02_microbench_*.txt) is a
    separate library: every function of include/iwae_b200_diag.h must be exported, none of them by the product library."""
    from iwae_b200 import _abi
    src = open(os.path.join(ROOT, "include", "iwae_b200_diag.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(iwae_[a-z0-9_]+)\s*\(", src)))
    assert {"iwae_umma_selftest", "iwae_umma_bench", "iwae_ldtm_bench", "iwae_diag_last_error"} <= set(names)
    d = _abi.diag_lib()
    for n in names:
        assert hasattr(d, n), f"libiwae_b200_diag.so does not export {n}"
        assert not hasattr(_abi.lib, n), f"{n} belongs to the diagnostics library, not to libiwae_b200.so"


@pytest.mark.parametrize("kind,cfg,okind,h,d", [
    (1, (784, 0, 200, 0, 100, 0), "iwae1", 200, 100),
    (1, (784, 0, 200, 0, 50, 0), "iwae1", 200, 50),
    (2, (784, 0, 200, 100, 100, 50), "iwae2", [200, 100], [100, 50]),
    (2, (784, 0, 200, 100, 4, 2), "iwae2", [200, 100], [4, 2]),
    (3, (784, 10, 200, 0, 100, 0), "ciwae", 200, 100),
    (4, (784, 10, 200, 0, 100, 0), "ciwae_std", 200, 100),
])
def test_param_layout_is_keras_order(oracle, kind, cfg, okind, h, d):
    from iwae_b200 import _abi
    n, layout = _abi.layout_for_config(_abi.Config(kind, *cfg))
    shapes = oracle.param_shapes(okind, h, d, n_cond=cfg[1])
    assert len(layout) == len(shapes)
    off = 0
    for (o, rows, cols), shp in zip(layout, shapes.values()):
        assert o == off
        assert (rows, cols) == (shp[0], shp[1] if len(shp) == 2 else 0)
        off += int(np.prod(shp))
    assert n == off


def test_bad_config_and_missing_gpu_fail_loudly():
    import torch
    from iwae_b200 import _abi
    with pytest.raises(_abi.IwaeError):
        _abi.layout_for_config(_abi.Config(9, 784, 0, 200, 0, 100, 0))
    with pytest.raises(_abi.IwaeError):
        _abi.layout_for_config(_abi.Config(2, 784, 0, 200, 0, 100, 0))     # IWAE2 without h2/d2
    if not torch.cuda.is_available():
        ctx = C.c_void_p()
        cfg = _abi.Config(1, 784, 0, 200, 0, 100, 0)
        assert _abi.lib.iwae_create(C.byref(cfg), C.byref(ctx)) != 0       # no CPU fallback
        assert b"no CUDA device" in _abi.lib.iwae_last_error()
        import iwae_b200
        with pytest.raises(RuntimeError):
            iwae_b200.IWAE(200, 100)


def test_product_code_never_imports_the_oracle():
    for sub in ("iwae_b200", "src"):
        for dp, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dp, f)).read()
                    assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dp, f)


def test_host_helpers():
    from iwae_b200 import data
    from iwae_b200.model import Result, _splitmix64
    sched, epochs = data.lr_schedule()
    assert epochs == 3280 and len(sched) == 8
    P, X, y = data.synthetic_binarised(16)
    assert X.dtype == np.float32 and set(np.unique(X)) <= {0.0, 1.0} and P.shape == (16, 784)
    b = data.get_bias(np.array([0.0, 0.5, 1.0]))
    assert np.allclose(b, [-np.log(1 / 0.001 - 1), 0.0, -np.log(1 / 0.999 - 1)])
    met = data.MyMetric()
    met.update_state(np.array([[1.0]])); met.update_state(np.array([[3.0]]))
    assert met.result() == 2.0
    r = Result(a=1, b=lambda: 2)
    assert r["b"] == 2 and dict.__getitem__(r, "b") == 2 and set(r.keys()) == {"a", "b"}
    assert _splitmix64(0) == 0xE220A8397B1DCDAF
