"""ctypes binding of libiwae_b200.so (include/iwae_b200.h).

This is the whole FFI surface: plain pointers and sizes, no torch types.  PyTorch tensors are used by
the callers only as device-buffer handles (`tensor.data_ptr()`).  There is no CPU fallback: if the
shared library is missing or cannot be loaded, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# IWAE_B200_LIB: an alternative build of the SAME library (A/B measurements of kernel variants, scripts/build_variants.sh)
LIB_PATH = os.environ.get("IWAE_B200_LIB") or os.path.join(_HERE, "lib", "libiwae_b200.so")

KIND_IWAE1, KIND_IWAE2, KIND_CIWAE, KIND_CIWAE_STD = 1, 2, 3, 4
OBJ = {"vae_elbo": 0, "iwae_elbo": 1, "iwae_eq14": 2, "vae_elbo_kl": 3, "dreg": 4}
PREC_FP32, PREC_TC_F16, PREC_TC_F16_FAST = 0, 1, 2
BLOCK_ENCODER, BLOCK_ENCODER2, BLOCK_DECODER21, BLOCK_PRIOR = 0, 1, 2, 3
PRECISIONS = {"fp32": PREC_FP32, "tc_f16": PREC_TC_F16, "tc_f16_fast": PREC_TC_F16_FAST}


class IwaeError(RuntimeError):
    """An entry point of libiwae_b200.so returned non-zero; the message is iwae_last_error()."""


class Config(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_x", C.c_int32), ("n_cond", C.c_int32), ("h1", C.c_int32),
                ("h2", C.c_int32), ("d1", C.c_int32), ("d2", C.c_int32)]


OUTPUT_FIELDS = ("lpxz", "lpz", "lq1", "lq2", "lp12", "log_w", "al", "iwae", "vae", "eq14", "kl", "dreg",
                 "snis_z1", "snis_z2", "z1", "z2", "logits", "q_mu", "q_std", "scalars")


class Outputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in OUTPUT_FIELDS]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C iwae_b200/csrc`).  iwae_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float
    P = C.POINTER
    sigs = {
        "iwae_last_error": (C.c_char_p, []),
        "iwae_version": (C.c_int, []),
        "iwae_create": (C.c_int, [P(Config), P(vp)]),
        "iwae_destroy": (C.c_int, [vp]),
        "iwae_param_count": (C.c_int, [vp, P(i64), P(i32)]),
        "iwae_param_layout": (C.c_int, [vp, P(i64), P(i32), P(i32)]),
        "iwae_layout_for_config": (C.c_int, [P(Config), P(i64), P(i32), P(i64), P(i32), P(i32)]),
        "iwae_forward": (C.c_int, [vp, vp, vp, vp, i32, i32, f32, vp, vp, u64, i64, P(Outputs), vp]),
        "iwae_forward_tc": (C.c_int, [vp, vp, vp, vp, i32, i32, f32, u64, i64, i32, P(Outputs), vp]),
        "iwae_llh_eval": (C.c_int, [vp, vp, vp, vp, i64, i32, u64, i64, i32, vp, vp, vp]),
        "iwae_llh_eval_eps": (C.c_int, [vp, vp, vp, vp, i64, i32, vp, vp, i32, vp, vp, vp]),
        "iwae_monitor_eval": (C.c_int, [vp, vp, vp, vp, i64, i32, u64, i64, i32, vp, vp, vp, vp]),
        "iwae_llh_eval_host": (C.c_int, [vp, vp, vp, vp, i64, i32, u64, i64, i32, vp, vp, vp]),
        "iwae_grads": (C.c_int, [vp, vp, vp, vp, i32, i32, f32, i32, i32, vp, vp, u64, i64, vp, P(Outputs), vp]),
        "iwae_adam_step": (C.c_int, [vp, vp, vp, vp, i64, f32, f32, f32, f32, i64, vp]),
        "iwae_train_step": (C.c_int, [vp, vp, vp, vp, vp, vp, i32, i32, f32, i32, vp, vp, u64, i64, f32, f32, f32,
                                      f32, i64, P(Outputs), vp]),
        "iwae_philox_normal": (C.c_int, [u64, i32, i64, i32, i32, i32, vp, vp]),
        "iwae_decode": (C.c_int, [vp, vp, vp, vp, vp, i32, i32, vp, vp, vp]),
        "iwae_block_forward": (C.c_int, [vp, vp, i32, vp, i32, vp, vp, vp]),
        "iwae_sir": (C.c_int, [vp, i32, vp, i32, i32, vp, vp, vp]),
        "iwae_softmax_f32": (C.c_int, [vp, i32, vp, vp]),
        "iwae_bernoulli_sample": (C.c_int, [vp, i64, u64, u64, vp, vp]),
        "iwae_launch_count": (i64, []),
        "iwae_llh_kernel_ms": (C.c_int, [vp, P(C.c_float)]),
        "iwae_set_params_version": (C.c_int, [vp, C.c_uint64]),
        "iwae_bernoulli_log_prob": (C.c_int, [vp, vp, i64, i32, i32, vp, vp]),
        "iwae_normal_log_prob": (C.c_int, [vp, vp, vp, i64, i32, i32, vp, vp]),
        "iwae_normal_sample": (C.c_int, [vp, vp, vp, i32, i32, i32, u64, i32, vp, vp]),
        "iwae_bernoulli_sample_logits": (C.c_int, [vp, i64, u64, u64, vp, vp, vp]),
        "iwae_posterior_grid": (C.c_int, [vp, vp, vp, C.c_double, C.c_double, C.c_double, C.c_double, i32, i32, vp, vp, vp, vp]),
        "iwae_comm_unique_id": (C.c_int, [vp]),
        "iwae_comm_init": (C.c_int, [vp, vp, i32, i32]),
        "iwae_comm_destroy": (C.c_int, [vp]),
        "iwae_comm_info": (C.c_int, [vp, P(i32), P(i32), P(i32)]),
        "iwae_allreduce_sum_f32": (C.c_int, [vp, vp, i64, vp]),
        "iwae_allreduce_sum_f64": (C.c_int, [vp, vp, i64, vp]),
        "iwae_dp_train_step": (C.c_int, [vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, i32, u64, i64, f32, f32, f32,
                                         f32, i64, P(Outputs), vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)      # AttributeError here = the library does not export the header's symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()

DIAG_LIB_PATH = os.path.join(_HERE, "lib", "libiwae_b200_diag.so")
_diag = None


def diag_lib():
    """libiwae_b200_diag.so (include/iwae_b200_diag.h): the tensor-core unit-test hook and the tcgen05.mma
    micro-benchmark.  Tests and scripts/ only - the product library does not carry them."""
    global _diag
    if _diag is None:
        d = C.CDLL(DIAG_LIB_PATH)
        vp, i32 = C.c_void_p, C.c_int32
        d.iwae_diag_last_error.restype = C.c_char_p
        d.iwae_umma_selftest.restype = C.c_int
        d.iwae_umma_selftest.argtypes = [vp, vp, i32, i32, i32, vp, vp]
        d.iwae_umma_bench.restype = C.c_int
        d.iwae_umma_bench.argtypes = [i32, i32, i32, vp, vp]
        d.iwae_ldtm_bench.restype = C.c_int
        d.iwae_ldtm_bench.argtypes = [i32, i32, i32, i32, vp, vp, vp]
        _diag = d
    return _diag


def check_diag(rc: int, what: str = ""):
    if rc != 0:
        msg = diag_lib().iwae_diag_last_error()
        raise IwaeError(f"{what} failed (rc={rc}): {msg.decode('utf-8', 'replace') if msg else ''}")
EXPORTS = ("iwae_last_error", "iwae_version", "iwae_create", "iwae_destroy", "iwae_param_count",
           "iwae_param_layout", "iwae_layout_for_config", "iwae_forward", "iwae_llh_eval", "iwae_llh_eval_host",
           "iwae_grads", "iwae_adam_step", "iwae_train_step", "iwae_philox_normal", "iwae_decode", "iwae_sir",
           "iwae_softmax_f32", "iwae_bernoulli_sample", "iwae_launch_count", "iwae_llh_kernel_ms", "iwae_set_params_version", "iwae_comm_unique_id", "iwae_comm_init", "iwae_comm_destroy", "iwae_comm_info",
           "iwae_allreduce_sum_f32", "iwae_allreduce_sum_f64", "iwae_dp_train_step", "iwae_forward_tc", "iwae_llh_eval_eps", "iwae_monitor_eval", "iwae_bernoulli_log_prob", "iwae_normal_log_prob",
           "iwae_normal_sample", "iwae_bernoulli_sample_logits", "iwae_posterior_grid",
           "iwae_block_forward")


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib.iwae_last_error()
        msg = msg.decode("utf-8", "replace") if msg else ""
        if msg.startswith("KeyError"):
            # the reference signals an unknown objective with KeyError (src/iwae1.py:157)
            raise KeyError(msg.split(":", 1)[1].strip())
        raise IwaeError(f"{what} failed (rc={rc}): {msg}")


def layout_for_config(cfg: Config):
    """(n_params, [(offset, rows, cols)]) of model.trainable_weights; works without a GPU."""
    n_params = C.c_int64(0)
    n_t = C.c_int32(64)
    off = (C.c_int64 * 64)()
    rows = (C.c_int32 * 64)()
    cols = (C.c_int32 * 64)()
    check(lib.iwae_layout_for_config(C.byref(cfg), C.byref(n_params), C.byref(n_t), off, rows, cols),
          "iwae_layout_for_config")
    return n_params.value, [(off[i], rows[i], cols[i]) for i in range(n_t.value)]
