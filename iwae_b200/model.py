"""Host-side mirror of the reference's model/objective API.

Same names, argument meaning and error behaviour as `src/iwae1.py`, `src/iwae2.py`,
`tasks/task02.py` (IWAEDReG) and `tasks/task04.py` / `tasks/task05.py` (CIWAE) of nbip/IWAE, so the
reference's drivers can run on it unchanged; every number is produced by libiwae_b200.so
(hand-written sm_100a kernels) through the ctypes C-ABI in `_abi.py`.  PyTorch tensors appear only as
device-buffer handles.  There is no CPU path: constructing a model without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _abi
from ._abi import Config, Outputs, check, lib

_F32 = torch.float32


def _stream_ptr() -> int:
    # (the raw handle of torch's current stream on the current device; torch.cuda.current_stream() builds a Stream
    #  object and re-resolves the device twice: ~14 us per call, on the path of every timed step)
    try:
        return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())
    except AttributeError:          # private binding gone in some torch build
        return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


class DeviceValue:
    """What the result dict holds: a device buffer that answers the few things the reference's drivers
    ask of a tf.Tensor - `.numpy()`, indexing (`res['iwae_elbo'][None, None]`, main.py:177), `.shape`,
    float() and `format(v, '.2f')` (main.py:161-162)."""

    __slots__ = ("tensor",)

    def __init__(self, tensor: torch.Tensor):
        self.tensor = tensor

    def numpy(self) -> np.ndarray:
        return self.tensor.detach().cpu().numpy()

    @property
    def shape(self):
        return tuple(self.tensor.shape)

    @property
    def dtype(self):
        return self.tensor.dtype

    def __getitem__(self, idx):
        return DeviceValue(self.tensor[idx])

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def __float__(self):
        return float(self.tensor.item())

    def __format__(self, spec):
        return format(float(self), spec) if self.tensor.numel() == 1 else format(self.numpy(), spec)

    def __len__(self):
        return self.tensor.shape[0]

    def __repr__(self):
        return f"DeviceValue(shape={self.shape}, device={self.tensor.device})"


class Result(dict):
    """The dict `IWAE.call` returns.  The large entries (`logits` [k,B,784], `z`) are thunks resolved on
    first access, so a training or monitoring step does not write GBs nobody reads (the reference
    materialises them on every call, src/iwae1.py:141-151)."""

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if callable(v):
            v = v()
            dict.__setitem__(self, key, v)
        return v

    def get(self, key, default=None):
        return self[key] if key in self else default

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]


class Variable:
    """One entry of `model.trainable_weights`: a view into the model's flat parameter buffer with the
    Keras layout (`kernel [in,out]`, `bias [out]`)."""

    def __init__(self, name: str, view: torch.Tensor, offset: int):
        self.name = name
        self.value = view
        self.offset = offset

    @property
    def shape(self):
        return tuple(self.value.shape)

    def numpy(self) -> np.ndarray:
        return self.value.detach().cpu().numpy()

    def assign(self, v):
        self.value.copy_(torch.as_tensor(np.asarray(v), dtype=_F32).reshape(self.value.shape))
        return self

    def __repr__(self):
        return f"<Variable {self.name} shape={self.shape}>"


def _as_dev(a, device) -> torch.Tensor:
    if isinstance(a, DeviceValue):
        a = a.tensor
    if not isinstance(a, torch.Tensor):
        a = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32)))
    return a.to(device=device, dtype=_F32).contiguous()


class Normal:
    """Stand-in for `tfd.Normal(loc, scale)` as the reference uses it: `.loc`, `.scale` (plot_task01.py:40-42,
    task02.py:61), `.log_prob(z)` (plot_task01.py:72, iwae1.py:109) and `.sample(n)` (iwae1.py:59), all backed by the
    library's kernels (iwae_normal_log_prob / iwae_normal_sample; TFP's op order)."""

    def __init__(self, loc: torch.Tensor, scale: torch.Tensor, seed_source=None):
        self.loc = DeviceValue(loc)
        self.scale = DeviceValue(scale)
        self._seed_source = seed_source

    def log_prob(self, z) -> DeviceValue:
        """Elementwise log N(z; loc, scale); z [..., *loc.shape] broadcast over leading (sample) dims."""
        loc, scale = self.loc.tensor.contiguous(), self.scale.tensor.contiguous()
        z = _as_dev(z, loc.device)
        d = loc.shape[-1]
        if z.shape[-1] != d or z.numel() % loc.numel():
            raise ValueError(f"z {tuple(z.shape)} does not broadcast against loc {tuple(loc.shape)}")
        out = torch.empty_like(z)
        check(lib.iwae_normal_log_prob(_ptr(z), _ptr(loc), _ptr(scale), z.numel() // d, loc.numel() // d, d, _ptr(out),
                                       _stream_ptr()), "iwae_normal_log_prob")
        return DeviceValue(out)

    def sample(self, n: Optional[int] = None, seed: Optional[int] = None) -> DeviceValue:
        """eps * scale + loc with eps ~ N(0,1) from the device Philox generator: [n, *loc.shape], or loc.shape if n is
        None (`q(z2|z1).sample()`, iwae2.py:65)."""
        loc, scale = self.loc.tensor.contiguous(), self.scale.tensor.contiguous()
        d = loc.shape[-1]
        k = 1 if n is None else int(n)
        if seed is None:
            seed = self._seed_source() if self._seed_source is not None else int(np.random.randint(0, 2 ** 31 - 1))
        out = torch.empty((k,) + tuple(loc.shape), dtype=_F32, device=loc.device)
        check(lib.iwae_normal_sample(_ptr(loc), _ptr(scale), None, loc.numel() // d, k, d, int(seed), 1, _ptr(out),
                                     _stream_ptr()), "iwae_normal_sample")
        return DeviceValue(out[0] if n is None else out)


class Bernoulli:
    """Stand-in for `tfd.Bernoulli(logits)`: `.logits`, `.log_prob(x)` (plot_task01.py:63; elementwise),
    `.sample()` and `.probs` (iwae1.py:172-176), backed by iwae_bernoulli_log_prob / iwae_bernoulli_sample_logits."""

    def __init__(self, owner: "IWAE", z: torch.Tensor, y, logits: torch.Tensor):
        self._owner, self._z, self._y = owner, z, y
        self.logits = DeviceValue(logits)

    def log_prob(self, x) -> DeviceValue:
        """Elementwise log p(x_j | z) [..., n_x]; x [B, n_x] (or [n_x]) broadcast over the leading dims of logits."""
        lg = self.logits.tensor.contiguous()
        n = lg.shape[-1]
        x = _as_dev(x, lg.device).reshape(-1, n)
        R = lg.numel() // n
        if R % x.shape[0]:
            raise ValueError("logits rows must be a multiple of the rows of x")
        out = torch.empty_like(lg)
        check(lib.iwae_bernoulli_log_prob(_ptr(lg), _ptr(x), R, x.shape[0], n, _ptr(out), _stream_ptr()),
              "iwae_bernoulli_log_prob")
        return DeviceValue(out)

    def log_prob_sum(self, x) -> DeviceValue:
        """sum_j log p(x_j | z): the reduction the reference applies right after log_prob
        (iwae1.py:111), computed in the decoder kernel's fused epilogue."""
        return self._owner._decode(self._z, self._y, x, want_logits=False)[1]

    @property
    def probs(self) -> DeviceValue:
        return self._owner._bernoulli(self.logits.tensor, 0)[1]

    def sample(self, seed: Optional[int] = None) -> DeviceValue:
        return self._owner._bernoulli(self.logits.tensor, self._owner._next_seed(seed))[0]


def _block_names(prefix):
    out = []
    for lyr in ("l1", "l2", "lmu", "lstd"):
        out += [f"{prefix}.{lyr}.kernel", f"{prefix}.{lyr}.bias"]
    return out


def _mlp3_names(prefix):
    out = []
    for i in range(3):
        out += [f"{prefix}.dense{i}.kernel", f"{prefix}.dense{i}.bias"]
    return out


class _SubModel:
    """`model.encoder` / `model.decoder` / `model.conditional_prior_network`: owns a contiguous range of
    the flat parameter buffer (task02.py:95-99 reads `.trainable_weights` of each)."""

    def __init__(self, owner: "IWAE", prefix: str):
        self._owner, self._prefix = owner, prefix

    @property
    def trainable_weights(self) -> List[Variable]:
        return [v for v in self._owner.trainable_weights if v.name.startswith(self._prefix + ".")]


class _Encoder(_SubModel):
    def __call__(self, x, n_samples, y=None):
        """Encoder.call (iwae1.py:56-61): returns (z [k,B,d], qzx)."""
        o = self._owner
        res = o._forward(x, y, int(n_samples), 1.0, want=("z1", "q_mu", "q_std"))
        return DeviceValue(res["z1"]), Normal(res["q_mu"], res["q_std"], seed_source=o._next_seed)


class _Decoder(_SubModel):
    def __call__(self, z, y=None):
        """Decoder.call (iwae1.py:79-85): returns (logits, pxz)."""
        logits, _ = self._owner._decode(z, y, None, want_logits=True)
        zt = z.tensor if isinstance(z, DeviceValue) else z
        return logits, Bernoulli(self._owner, zt, y, logits.tensor)

    def decode_z_to_x(self, z, y=None):
        return self._owner._decode(z, y, None, want_logits=True)[0]


class IWAE:
    """iwae1.IWAE (src/iwae1.py:88-166): one stochastic layer.

    `IWAE(n_hidden, n_latent)`; `model(x, n_samples, beta=1.0)` -> dict with the reference's keys;
    `train_step(x, n_samples, beta, optimizer, objective=...)`; `val_step(x, n_samples, beta)`.
    Extras over the reference: `eps=` (inject the N(0,1) noise, [k,B,d]), `seed=`, and the per-example
    vectors `iwae_elbo_per_example` / `vae_elbo_per_example` / `log_w` in the result.
    """

    KIND = _abi.KIND_IWAE1
    # result-dict names of (lpxz, prior term, q term): src/iwae1.py:141-151
    TERM_KEYS = ("lpxz", "lpz", "lqzx")

    def __init__(self, n_hidden, n_latent, n_x: int = 784, n_cond: int = 0, seed: int = 123, device=None,
                 out_bias: Optional[np.ndarray] = None, **kwargs):
        if not torch.cuda.is_available():
            raise RuntimeError("iwae_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._cfg = self._make_config(n_hidden, n_latent, n_x, n_cond)
        self._ctx = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.iwae_create(C.byref(self._cfg), C.byref(self._ctx)), "iwae_create")
        n_params, layout = _abi.layout_for_config(self._cfg)
        self.n_params = n_params
        self.params = torch.zeros(n_params, dtype=_F32, device=self.device)
        self._names = self._param_names()
        assert len(self._names) == len(layout)
        self.trainable_weights: List[Variable] = []
        for name, (off, rows, cols) in zip(self._names, layout):
            n = rows * cols if cols else rows
            view = self.params[off:off + n].view(rows, cols) if cols else self.params[off:off + n]
            self.trainable_weights.append(Variable(name, view, off))
        self._seed = int(seed)
        self._calls = 0
        self._native_writes = 0        # bumped whenever the library writes self.params through a raw pointer
        self._build_submodels()
        self.init_weights(np.random.RandomState(seed), out_bias)

    # ---- construction ---------------------------------------------------------------------------
    def _make_config(self, n_hidden, n_latent, n_x, n_cond) -> Config:
        return Config(self.KIND, int(n_x), 0, int(n_hidden), 0, int(n_latent), 0)

    def _param_names(self) -> List[str]:
        return _block_names("encoder.encode_x_to_z") + _mlp3_names("decoder.decode_z_to_x")

    def _build_submodels(self):
        self.encoder = _Encoder(self, "encoder")
        self.decoder = _Decoder(self, "decoder")

    def close(self):
        """Destroy the library context now (scratch arena, packed weights, CUDA graph, NCCL communicator).  Models hold
        reference cycles, so without this the context is freed whenever the cyclic GC runs - and destroying an NCCL
        communicator is a blocking call that every rank must reach at the same point of its program: multi-rank
        callers close their models explicitly (bench.py, tests)."""
        ctx = getattr(self, "_ctx", None)
        if ctx is not None and ctx.value:
            try:
                lib.iwae_destroy(ctx)
            except Exception:
                pass
            self._ctx = C.c_void_p()

    def __del__(self):
        self.close()

    # ---- weights --------------------------------------------------------------------------------
    def init_weights(self, rs: np.random.RandomState, out_bias: Optional[np.ndarray] = None, scale: float = 1.0):
        """Keras defaults: kernels glorot_uniform, biases zero; the 784-way output bias is utils.get_bias()
        (iwae1.py:74-75).  TF's seeded initialiser stream is not reproducible outside TF, so values
        come from `rs`; only the order and shapes are the reference's.  out_bias=None: the get_bias() of the synthetic
        train set (data.default_out_bias), like a reference model built with its defaults."""
        flat = np.zeros(self.n_params, dtype=np.float32)
        for v in self.trainable_weights:
            shp = v.shape
            if len(shp) == 2:
                lim = math.sqrt(6.0 / (shp[0] + shp[1])) * scale
                flat[v.offset:v.offset + shp[0] * shp[1]] = rs.uniform(-lim, lim, size=shp).astype(np.float32).ravel()
        if out_bias is None:
            from . import data as _data
            out_bias = _data.default_out_bias(self._cfg.n_x)     # the reference ALWAYS starts from utils.get_bias()
        if out_bias is not None:
            last = [v for v in self.trainable_weights if v.name.endswith("dense2.bias")][-1]
            flat[last.offset:last.offset + last.shape[0]] = np.asarray(out_bias, dtype=np.float64).astype(np.float32)
        self.params.copy_(torch.from_numpy(flat))

    def get_weights(self) -> "OrderedDict[str, np.ndarray]":
        return OrderedDict((v.name, v.numpy()) for v in self.trainable_weights)

    def set_weights(self, weights):
        """dict name -> array, or a list in `trainable_weights` order."""
        if isinstance(weights, dict):
            for v in self.trainable_weights:
                v.assign(weights[v.name])
        else:
            weights = list(weights)
            assert len(weights) == len(self.trainable_weights)
            for v, w in zip(self.trainable_weights, weights):
                v.assign(w)

    def save_weights(self, path: str):
        """model.save_weights (main.py:165).  Own format (.npz keyed by the Keras-style names); the
        TF checkpoint format cannot be written without TF."""
        import os
        os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
        np.savez(path if path.endswith(".npz") else path + ".npz", **self.get_weights())

    def load_weights(self, path: str):
        with np.load(path if path.endswith(".npz") else path + ".npz") as f:
            self.set_weights({k: f[k] for k in f.files})

    # ---- plumbing -------------------------------------------------------------------------------
    def mark_params_modified(self):
        """Tell the model that `self.params` was written through its raw pointer (the library's Adam kernels do that;
        torch-side writes are tracked by the tensor's own version counter)."""
        self._native_writes += 1

    def _params_version(self) -> int:
        return (((self.params._version + 1) & 0xFFFFFFFF) << 32) | (self._native_writes & 0xFFFFFFFF)

    def _declare_params_version(self):
        """iwae_set_params_version: lets the fused evaluation reuse its packed fp16 weight image while the parameters
        are unchanged.  Version = torch's in-place version counter of the flat buffer (every torch write to it or to a
        Variable view bumps it) combined with the count of native writes."""
        check(lib.iwae_set_params_version(self._ctx, C.c_uint64(self._params_version())), "iwae_set_params_version")

    def _dev(self, a, dtype=_F32) -> Optional[torch.Tensor]:
        if a is None:
            return None
        if isinstance(a, DeviceValue):
            a = a.tensor
        if not isinstance(a, torch.Tensor):
            a = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32 if dtype == _F32 else None)))
        return a.to(device=self.device, dtype=dtype, non_blocking=True).contiguous()

    def _next_seed(self, seed=None) -> int:
        if seed is not None:
            return int(seed) & 0xFFFFFFFFFFFFFFFF
        self._calls += 1
        return _splitmix64(self._seed * 0x100000001B3 + self._calls)

    @property
    def _d1(self):
        return self._cfg.d1

    @property
    def _d2(self):
        return self._cfg.d2

    def _alloc(self, *shape):
        return torch.empty(*shape, dtype=_F32, device=self.device)

    def _outputs(self, B, k, want: Sequence[str]):
        """Allocate the requested output buffers; returns (Outputs struct, dict name -> tensor)."""
        cfg = self._cfg
        shapes = {
            "lpxz": (k, B), "lpz": (k, B), "lq1": (k, B), "lq2": (k, B), "lp12": (k, B), "log_w": (k, B),
            "al": (k, B), "iwae": (B,), "vae": (B,), "eq14": (B,), "kl": (B,), "dreg": (B,),
            "snis_z1": (B, cfg.d1), "snis_z2": (B, max(cfg.d2, 1)), "z1": (k, B, cfg.d1),
            "z2": (k, B, max(cfg.d2, 1)), "logits": (k, B, cfg.n_x), "q_mu": (B, cfg.d1), "q_std": (B, cfg.d1),
            "scalars": (8,),
        }
        bufs: Dict[str, torch.Tensor] = {}
        o = Outputs()
        for name in want:
            t = self._alloc(*shapes[name])
            bufs[name] = t
            setattr(o, name, t.data_ptr())
        return o, bufs

    def _check_x(self, x, y):
        x = self._dev(x)
        if x.dim() == 1:
            x = x[None, :]
        if x.dim() != 2 or x.shape[1] != self._cfg.n_x:
            raise ValueError(f"x must be [B, {self._cfg.n_x}], got {tuple(x.shape)}")
        if self._cfg.n_cond:
            if y is None:
                raise TypeError("this model is label-conditioned: call(x, y, n_samples, beta)")
            y = self._dev(np.asarray(y, dtype=np.float32) if not isinstance(y, (torch.Tensor, DeviceValue)) else y)
            y = y.reshape(-1)
            if y.shape[0] != x.shape[0]:
                raise ValueError("x and y disagree on the batch size")
        else:
            y = None
        return x, y

    def _check_eps(self, eps, k, B, d, what):
        if eps is None:
            return None
        eps = self._dev(eps)
        if tuple(eps.shape) != (k, B, d):
            raise ValueError(f"{what} must be [{k}, {B}, {d}], got {tuple(eps.shape)}")
        return eps

    def _forward(self, x, y, k, beta, want, eps1=None, eps2=None, seed=None, img_offset=0):
        x, y = self._check_x(x, y)
        B = x.shape[0]
        if k <= 0:
            raise ValueError("n_samples must be positive")
        eps1 = self._check_eps(eps1, k, B, self._d1, "eps")
        eps2 = self._check_eps(eps2, k, B, self._d2, "eps2") if self._d2 else None
        o, bufs = self._outputs(B, k, want)
        seed = self._next_seed(seed)
        with torch.cuda.device(self.device):
            check(lib.iwae_forward(self._ctx, _ptr(self.params), _ptr(x), _ptr(y), B, k, float(beta), _ptr(eps1),
                                   _ptr(eps2), seed, int(img_offset), C.byref(o), _stream_ptr()), "iwae_forward")
        bufs["_seed"] = seed
        bufs["_x"], bufs["_y"], bufs["_eps1"], bufs["_eps2"] = x, y, eps1, eps2
        return bufs

    def _forward_tc(self, x, y, k, beta, want, seed=None, img_offset=0, precision="tc_f16_fast"):
        """The same outputs through the fused tensor-core kernel (iwae_forward_tc): per-row terms, log_w, al, ELBOs, SNIS
        means, scalars - no [k,B,784] tensor is ever written (a full-test-set val_step at k=50 takes ~1.5 ms instead of
        17 ms).  `logits` (and the noise-injection / DReG extras) are not available here."""
        x, y = self._check_x(x, y)
        B = x.shape[0]
        if k <= 0:
            raise ValueError("n_samples must be positive")
        o, bufs = self._outputs(B, k, want)
        seed = self._next_seed(seed)
        with torch.cuda.device(self.device):
            self._declare_params_version()
            check(lib.iwae_forward_tc(self._ctx, _ptr(self.params), _ptr(x), _ptr(y), B, k, float(beta), seed,
                                      int(img_offset), _abi.PRECISIONS[precision], C.byref(o), _stream_ptr()),
                  "iwae_forward_tc")
        bufs["_seed"] = seed
        bufs["_x"], bufs["_y"], bufs["_eps1"], bufs["_eps2"] = x, y, None, None
        return bufs

    def _engine(self, precision, want, eps=None, eps2=None):
        """Which kernels run a forward call: "fp32" (parity mode, takes injected noise, can write logits) or one of the
        fused tensor-core modes.  "auto" picks the fused kernel when the architecture and the request allow it."""
        tc_ok = self.tc_supported() and eps is None and eps2 is None and not ({"logits", "dreg"} & set(want))
        if precision in (None, "auto"):
            return "tc_f16_fast" if tc_ok else "fp32"
        if precision != "fp32" and not tc_ok:
            raise ValueError("the fused tensor-core forward takes Philox noise only and does not produce logits / dreg")
        return precision

    def _decode(self, z, y, x, want_logits: bool):
        """decoder(z) and/or sum_j Bernoulli(logits).log_prob(x) for arbitrary z [..., d]
        (plot_task01.py:57-63; model.sample, iwae1.py:168-178)."""
        z = self._dev(z)
        lead = tuple(z.shape[:-1])
        if z.shape[-1] != self._d1:
            raise ValueError(f"z must have last dim {self._d1}")
        z2 = z.reshape(-1, self._d1)
        R = z2.shape[0]
        xd = self._dev(x) if x is not None else None
        if xd is not None and xd.dim() == 1:
            xd = xd[None, :]
        if self._cfg.n_cond:
            if y is None:
                raise TypeError("this model is label-conditioned: pass y")
            yd = self._dev(np.asarray(y, dtype=np.float32) if not isinstance(y, (torch.Tensor, DeviceValue)) else y).reshape(-1)
            B = yd.shape[0]
        else:
            yd = None
            B = xd.shape[0] if xd is not None else 1
        if R % B:
            raise ValueError("number of z rows must be a multiple of the batch size of x / y")
        logits = self._alloc(R, self._cfg.n_x) if want_logits else None
        lp = self._alloc(R) if xd is not None else None
        with torch.cuda.device(self.device):
            check(lib.iwae_decode(self._ctx, _ptr(self.params), _ptr(z2), _ptr(xd), _ptr(yd), R, B, _ptr(logits),
                                  _ptr(lp), _stream_ptr()), "iwae_decode")
        return (DeviceValue(logits.view(*lead, self._cfg.n_x)) if logits is not None else None,
                DeviceValue(lp.view(*lead)) if lp is not None else None)

    # ---- the reference API ----------------------------------------------------------------------
    _WANT = ("lpxz", "lpz", "lq1", "log_w", "al", "iwae", "vae", "eq14", "snis_z1", "scalars")

    def _lazy(self, bufs, k, beta, names):
        """Entries of the result dict that are large and rarely read (`z`, `logits`: [k,B,784] is 1.6 GB for a
        full-test-set val_step).  If the step wrote them (train_step(materialize=...)) they are returned as they are.
        Otherwise a thunk re-runs the forward with the same noise on first access - valid only while the parameters
        are the ones the call used: after a train_step (whose Adam update has already changed them) or any other
        parameter change the thunk raises instead of silently returning values of a different model
        (the reference's tape returns the PRE-update forward, src/iwae1.py:153-162)."""
        version = self._params_version()
        after_train = bool(bufs.get("_after_train"))

        def run():
            if after_train or self._params_version() != version:
                raise RuntimeError(
                    "this result's 'z'/'logits' were not materialised and the parameters have changed since the call; "
                    "pass materialize=('z', 'logits') to train_step (or read them before updating the weights)")
            return self._forward(bufs["_x"], bufs["_y"], k, beta, names, bufs["_eps1"], bufs["_eps2"], bufs["_seed"],
                                 bufs.get("_img_offset", 0))
        cache = {}

        def get(name):
            if name in bufs:
                return DeviceValue(bufs[name])

            def thunk():
                if not cache:
                    cache.update(run())
                return DeviceValue(cache[name])
            return thunk
        return get

    def _result(self, bufs, k, beta) -> Result:
        s = bufs["scalars"]
        kx, kp, kq = self.TERM_KEYS
        lazy = self._lazy(bufs, k, beta, ("z1", "logits"))
        res = Result()
        res._seed = bufs.get("_seed")
        res["vae_elbo"] = DeviceValue(s[0])
        res["vae_elbo_kl"] = DeviceValue(s[1])
        res["iwae_elbo"] = DeviceValue(s[2])
        res["iwae_eq14"] = DeviceValue(s[3])
        res["z"] = lazy("z1")
        res["snis_z"] = DeviceValue(bufs["snis_z1"])
        res["al"] = DeviceValue(bufs["al"])
        res["logits"] = lazy("logits")
        res[kx] = DeviceValue(bufs["lpxz"])
        res[kp] = DeviceValue(bufs["lpz"])
        res[kq] = DeviceValue(bufs["lq1"])
        self._extras(res, bufs)
        return res

    @staticmethod
    def _extras(res, bufs):
        res["iwae_elbo_per_example"] = DeviceValue(bufs["iwae"])
        res["vae_elbo_per_example"] = DeviceValue(bufs["vae"])
        res["log_w"] = DeviceValue(bufs["log_w"])

    def call(self, x, n_samples, beta=1.0, eps=None, seed=None, img_offset=0, precision="fp32"):
        """IWAE.call (src/iwae1.py:98-151).  precision: "fp32" (default: the parity kernels), "auto" / "tc_f16" /
        "tc_f16_fast" (the fused tensor-core kernel as the engine, see val_step)."""
        k = int(n_samples)
        engine = self._engine(precision, self._WANT, eps)
        if engine == "fp32":
            bufs = self._forward(x, None, k, beta, self._WANT, eps, None, seed, img_offset)
        else:
            bufs = self._forward_tc(x, None, k, beta, self._WANT, seed, img_offset, engine)
        bufs["_img_offset"] = img_offset
        return self._result(bufs, k, beta)

    __call__ = call

    def val_step(self, x, n_samples, beta, precision="auto"):
        """IWAE.val_step (src/iwae1.py:164-166; the driver calls it on the WHOLE test set every 200 steps, main.py:152).
        Runs on the fused tensor-core kernel when the architecture fits it: every dict entry of `call` with no
        [k,B,784] tensor written; `logits` / `z` stay lazy (first access recomputes them with the fp32 kernels and the
        same noise seed)."""
        return self.call(x, n_samples, beta, precision=precision)

    def _objective_id(self, objective: str, res_keys) -> int:
        if objective not in res_keys or objective not in _abi.OBJ:
            raise KeyError(objective)       # the reference's error for an unknown objective (iwae1.py:157)
        return _abi.OBJ[objective]

    _OBJECTIVES = ("vae_elbo", "iwae_elbo", "iwae_eq14", "vae_elbo_kl")

    _MATERIALIZE = {"z": ("z1",), "z1": ("z1",), "z2": ("z2",), "logits": ("logits",)}

    def _train(self, x, y, k, beta, optimizer, objective_id, eps1=None, eps2=None, seed=None, want=None,
               materialize=()):
        x, y = self._check_x(x, y)
        B = x.shape[0]
        eps1 = self._check_eps(eps1, k, B, self._d1, "eps")
        eps2 = self._check_eps(eps2, k, B, self._d2, "eps2") if self._d2 else None
        want = tuple(want or self._WANT)
        for key in materialize:
            for f in self._MATERIALIZE[key]:        # KeyError for anything but z / z1 / z2 / logits
                if f not in want and (f != "z2" or self._d2):
                    want += (f,)
        o, bufs = self._outputs(B, k, want)
        seed = self._next_seed(seed)
        m, v = optimizer._state_for(self)
        with torch.cuda.device(self.device):
            check(lib.iwae_train_step(self._ctx, _ptr(self.params), _ptr(m), _ptr(v), _ptr(x), _ptr(y), B, k,
                                      float(beta), objective_id, _ptr(eps1), _ptr(eps2), seed, 0,
                                      float(optimizer.learning_rate.value), optimizer.beta_1, optimizer.beta_2,
                                      optimizer.epsilon, optimizer.iterations + 1, C.byref(o), _stream_ptr()),
                  "iwae_train_step")
        optimizer.iterations += 1          # only once the step has been issued (a failed call leaves the count alone)
        self.mark_params_modified()
        bufs["_seed"] = seed
        bufs["_x"], bufs["_y"], bufs["_eps1"], bufs["_eps2"] = x, y, eps1, eps2
        bufs["_after_train"] = True
        return bufs

    def train_step(self, x, n_samples, beta, optimizer, objective="vae_elbo", eps=None, seed=None, materialize=()):
        """IWAE.train_step (src/iwae1.py:153-162): forward, reverse pass of -res[objective], Keras-Adam
        update.  The returned dict holds the PRE-update forward values, like the reference's tape; the large entries
        `z` and `logits` are written only if named in `materialize` (reading them otherwise raises: by then the
        parameters are the updated ones, so they cannot be recomputed)."""
        oid = self._objective_id(objective, self._OBJECTIVES)
        k = int(n_samples)
        bufs = self._train(x, None, k, beta, optimizer, oid, eps, None, seed, materialize=materialize)
        return self._result(bufs, k, beta)

    def grads(self, x, n_samples, beta=1.0, objective="iwae_elbo", eps=None, eps2=None, seed=None, y=None,
              B_norm: int = 0, img_offset: int = 0) -> torch.Tensor:
        """Flat gradient of -res[objective] in `trainable_weights` order (tape.gradient, iwae1.py:159)."""
        oid = _abi.OBJ[objective] if objective == "dreg" else self._objective_id(objective, self._OBJECTIVES)
        x, y = self._check_x(x, y)
        B, k = x.shape[0], int(n_samples)
        eps = self._check_eps(eps, k, B, self._d1, "eps")
        eps2 = self._check_eps(eps2, k, B, self._d2, "eps2") if self._d2 else None
        g = torch.empty_like(self.params)
        with torch.cuda.device(self.device):
            check(lib.iwae_grads(self._ctx, _ptr(self.params), _ptr(x), _ptr(y), B, k, float(beta), oid, int(B_norm),
                                 _ptr(eps), _ptr(eps2), self._next_seed(seed), int(img_offset), _ptr(g), None,
                                 _stream_ptr()), "iwae_grads")
        return g

    # ---- k-sample test log-likelihood (main.py:170-184) ------------------------------------------
    def tc_supported(self) -> bool:
        """Whether the fused tcgen05 evaluation kernel covers this architecture (the limits `build_program` in
        csrc/eval_tc.cu enforces: one 208-column activation buffer, one 112-column z1 buffer, <= 7 logit chunks)."""
        c = self._cfg
        ok = c.h1 + 1 <= 208 and c.d1 <= 100 and c.n_x <= 784 and c.d1 + (c.n_cond if self._cfg.n_cond else 0) + 1 <= 112
        if c.kind == _abi.KIND_IWAE2:
            ok = ok and c.h2 + 1 <= 208 and c.d2 <= 56
        return bool(ok)

    def _precision(self, precision):
        """None / "auto": the fused tensor-core kernel when the architecture fits it, else the fp32 CUDA kernels.
        The tensor-core precisions assume BINARISED images (x in {0,1} exactly, as main.py feeds them): pass
        precision="fp32" for real-valued x."""
        if precision in (None, "auto"):
            return "tc_f16_fast" if self.tc_supported() else "fp32"
        return precision

    def llh(self, x, n_samples=5000, y=None, precision="auto", seed=None, img_offset=0, eps=None, eps2=None):
        """Per-example log-mean-exp of `n_samples` importance weights for every row of x (device
        resident), and their fp64 sum: returns (per_example [N] tensor, sum as 0-dim fp64 tensor).
        `eps` [k,N,d1] (and `eps2` [k,N,d2]) inject the N(0,1) noise instead of the device Philox generator."""
        precision = self._precision(precision)
        x, y = self._check_x(x, y)
        N = x.shape[0]
        per = self._alloc(max(N, 1))[:N]
        total = torch.empty((), dtype=torch.float64, device=self.device)     # written by every path below (memset when N == 0)
        if eps is not None:
            k = int(n_samples)
            eps = self._check_eps(eps, k, N, self._d1, "eps")
            eps2 = self._check_eps(eps2, k, N, self._d2, "eps2") if self._d2 else None
            if precision == "fp32":      # the fp32 kernels take injected noise through IWAE.call
                r = self._forward(x, y, k, 1.0, ("iwae",), eps, eps2)
                return r["iwae"], r["iwae"].double().sum()
            with torch.cuda.device(self.device):
                self._declare_params_version()
                check(lib.iwae_llh_eval_eps(self._ctx, _ptr(self.params), _ptr(x), _ptr(y), N, k, _ptr(eps), _ptr(eps2),
                                            _abi.PRECISIONS[precision], _ptr(per), _ptr(total), _stream_ptr()),
                      "iwae_llh_eval_eps")
            return per, total
        with torch.cuda.device(self.device):
            self._declare_params_version()
            check(lib.iwae_llh_eval(self._ctx, _ptr(self.params), _ptr(x), _ptr(y), N, int(n_samples),
                                    self._next_seed(seed), int(img_offset), _abi.PRECISIONS[precision], _ptr(per),
                                    _ptr(total), _stream_ptr()), "iwae_llh_eval")
        return per, total

    def monitor(self, x, n_samples, y=None, precision="auto", seed=None):
        """The monitoring numbers of `val_step` on a whole test set (main.py:152) at beta = 1 through the fused
        evaluation kernel: {'iwae_elbo', 'vae_elbo'} as batch means plus the per-example vectors, without
        materialising any [k,B,.] tensor."""
        precision = self._precision(precision)
        x, y = self._check_x(x, y)
        N = x.shape[0]
        per_i, per_v = self._alloc(max(N, 1))[:N], self._alloc(max(N, 1))[:N]
        sums = torch.zeros(2, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            self._declare_params_version()
            check(lib.iwae_monitor_eval(self._ctx, _ptr(self.params), _ptr(x), _ptr(y), N, int(n_samples),
                                        self._next_seed(seed), 0, _abi.PRECISIONS[precision], _ptr(per_i), _ptr(per_v),
                                        _ptr(sums), _stream_ptr()), "iwae_monitor_eval")
        res = Result()
        res["iwae_elbo"] = DeviceValue((sums[0] / max(N, 1)).float())
        res["vae_elbo"] = DeviceValue((sums[1] / max(N, 1)).float())
        res["iwae_elbo_per_example"] = DeviceValue(per_i)
        res["vae_elbo_per_example"] = DeviceValue(per_v)
        return res

    def llh_host(self, x_host: np.ndarray, n_samples=5000, y_host=None, precision="auto", seed=None,
                 img_offset=0):
        """Same through the reference-facing HOST-buffer entry point (iwae_llh_eval_host): x (and y) are
        host arrays, copied to the device, evaluated, results copied back; synchronous."""
        precision = self._precision(precision)
        x_host = np.ascontiguousarray(x_host, dtype=np.float32)
        N = x_host.shape[0]
        per = np.empty(N, dtype=np.float32)
        total = C.c_double(0.0)
        yh = None if y_host is None else np.ascontiguousarray(y_host, dtype=np.float32)
        with torch.cuda.device(self.device):
            self._declare_params_version()
            check(lib.iwae_llh_eval_host(self._ctx, _ptr(self.params), x_host.ctypes.data_as(C.c_void_p),
                                         None if yh is None else yh.ctypes.data_as(C.c_void_p), N, int(n_samples),
                                         self._next_seed(seed), int(img_offset), _abi.PRECISIONS[precision],
                                         per.ctypes.data_as(C.c_void_p), C.byref(total), _stream_ptr()),
                  "iwae_llh_eval_host")
        return per, total.value

    def last_eval_kernel_ms(self) -> float:
        """Device time of the fused evaluation kernel of the most recent tensor-core `llh` call."""
        ms = C.c_float(0.0)
        check(lib.iwae_llh_kernel_ms(self._ctx, C.byref(ms)), "iwae_llh_kernel_ms")
        return float(ms.value)

    # ---- sampling (iwae1.py:168-178) ------------------------------------------------------------
    def _bernoulli(self, logits: torch.Tensor, seed):
        """(x_sample, probs) of tfd.Bernoulli(logits) - the order every `sample` of the reference returns."""
        logits = logits.contiguous()
        probs, out = torch.empty_like(logits), torch.empty_like(logits)
        with torch.cuda.device(self.device):
            check(lib.iwae_bernoulli_sample_logits(_ptr(logits), logits.numel(), int(seed), 0, _ptr(out), _ptr(probs),
                                                   _stream_ptr()), "iwae_bernoulli_sample_logits")
        return DeviceValue(out), DeviceValue(probs)

    def _block(self, block_id: int, inp: torch.Tensor, n_out: int):
        """(loc, scale) of one BasicBlock of the model on rows `inp` (iwae1.py:36-44)."""
        rows = inp.reshape(-1, inp.shape[-1]).contiguous()
        mu, sd = self._alloc(rows.shape[0], n_out), self._alloc(rows.shape[0], n_out)
        with torch.cuda.device(self.device):
            check(lib.iwae_block_forward(self._ctx, _ptr(self.params), block_id, _ptr(rows), rows.shape[0], _ptr(mu),
                                         _ptr(sd), _stream_ptr()), "iwae_block_forward")
        lead = tuple(inp.shape[:-1])
        return mu.view(*lead, n_out), sd.view(*lead, n_out)

    def _normal(self, seed, stream_id, n, rows, d):
        """eps [n, rows, d] from the device Philox generator (the noise source of every sampling call)."""
        eps = self._alloc(n, rows, d)
        with torch.cuda.device(self.device):
            check(lib.iwae_philox_normal(int(seed), stream_id, 0, rows, n, d, _ptr(eps), _stream_ptr()), "iwae_philox_normal")
        return eps

    def sample(self, z, y=None, seed=None):
        """model.sample(z) -> (x_sample, probs) (iwae1.py:168-178): decode prior samples to Bernoulli
        probabilities and one binary draw each (a fresh draw per call unless `seed` is given)."""
        logits, _ = self._decode(z, y, None, want_logits=True)
        return self._bernoulli(logits.tensor, self._next_seed(seed))

    def posterior_grid(self, x, range_x, range_y, n1=200, n2=200):
        """The true-posterior grid of tasks/plot_task01.py:46-70 for a 2-D latent: log p(z|x) on an n2 x n1 grid
        (normalised over the grid with logsumexp) and the variational log q(z|x) on the same grid.  Returns
        (log_posterior [n2,n1], variational_log_posterior [n2,n1], grid [n1*n2, 2])."""
        if self._d1 != 2 or self._cfg.n_cond:
            raise ValueError("posterior_grid needs an unconditional model with a 2-D latent (tasks/task01.py:84)")
        x1 = self._dev(x).reshape(-1)
        if x1.numel() != self._cfg.n_x:
            raise ValueError("posterior_grid takes ONE image")
        log_post, lq = self._alloc(n2, n1), self._alloc(n2, n1)
        grid = self._alloc(n1 * n2, 2)
        with torch.cuda.device(self.device):        # grid, decoder + Bernoulli log-lik, prior, normalisation, q: all in the library
            check(lib.iwae_posterior_grid(self._ctx, _ptr(self.params), _ptr(x1), float(range_x[0]), float(range_x[1]),
                                          float(range_y[0]), float(range_y[1]), int(n1), int(n2), _ptr(log_post), _ptr(lq),
                                          _ptr(grid), _stream_ptr()), "iwae_posterior_grid")
        return DeviceValue(log_post), DeviceValue(lq), DeviceValue(grid)

    @staticmethod
    def write_to_tensorboard(res, step, writer=None):
        """iwae1.py:226-232: the five scalars the reference logs; returns them (and writes them if a
        torch.utils.tensorboard SummaryWriter is given)."""
        vals = {"Evaluation/vae_elbo": float(res["vae_elbo"]), "Evaluation/iwae_elbo": float(res["iwae_elbo"])}
        for key in res.keys():
            if key.startswith("lp") or key.startswith("lq"):
                vals[f"Evaluation/{key}"] = float(res[key].numpy().mean())
        if writer is not None:
            for name, v in vals.items():
                writer.add_scalar(name, v, step)
        return vals


class IWAE2(IWAE):
    """iwae2.IWAE (src/iwae2.py:99-182): two stochastic layers; `beta` is accepted and ignored and there
    is no `vae_elbo_kl` key (iwae2.py:128,154-167)."""

    KIND = _abi.KIND_IWAE2
    _OBJECTIVES = ("vae_elbo", "iwae_elbo", "iwae_eq14")
    _WANT = ("lpxz", "lpz", "lq1", "lq2", "lp12", "log_w", "al", "iwae", "vae", "eq14", "snis_z1", "snis_z2",
             "scalars")

    def _make_config(self, n_hidden, n_latent, n_x, n_cond):
        return Config(self.KIND, int(n_x), 0, int(n_hidden[0]), int(n_hidden[1]), int(n_latent[0]), int(n_latent[1]))

    def _param_names(self):
        return (_block_names("encoder.encode_x_to_z1") + _block_names("encoder.encode_z1_to_z2") +
                _block_names("decoder.decode_z2_to_z1") + _mlp3_names("decoder.decode_z1_to_x"))

    def _result(self, bufs, k, beta) -> Result:
        s = bufs["scalars"]
        lazy = self._lazy(bufs, k, beta, ("z1", "z2", "logits"))
        res = Result()
        res._seed = bufs.get("_seed")
        res["vae_elbo"] = DeviceValue(s[0])
        res["iwae_elbo"] = DeviceValue(s[2])
        res["iwae_eq14"] = DeviceValue(s[3])
        res["z1"] = lazy("z1")
        res["z2"] = lazy("z2")
        res["snis_z1"] = DeviceValue(bufs["snis_z1"])
        res["snis_z2"] = DeviceValue(bufs["snis_z2"])
        res["al"] = DeviceValue(bufs["al"])
        res["logits"] = lazy("logits")
        res["lpxz1"] = DeviceValue(bufs["lpxz"])
        res["lpz1z2"] = DeviceValue(bufs["lp12"])
        res["lpz2"] = DeviceValue(bufs["lpz"])
        res["lqz1x"] = DeviceValue(bufs["lq1"])
        res["lqz2z1"] = DeviceValue(bufs["lq2"])
        self._extras(res, bufs)
        return res

    def call(self, x, n_samples, beta=1.0, eps=None, eps2=None, seed=None, img_offset=0, precision="fp32"):
        """iwae2.IWAE.call (src/iwae2.py:109-167); eps [k,B,d1], eps2 [k,B,d2] inject the noise."""
        k = int(n_samples)
        engine = self._engine(precision, self._WANT, eps, eps2)
        if engine == "fp32":
            bufs = self._forward(x, None, k, beta, self._WANT, eps, eps2, seed, img_offset)
        else:
            bufs = self._forward_tc(x, None, k, beta, self._WANT, seed, img_offset, engine)
        bufs["_img_offset"] = img_offset
        return self._result(bufs, k, beta)

    __call__ = call

    def train_step(self, x, n_samples, beta, optimizer, objective="vae_elbo", eps=None, eps2=None, seed=None,
                   materialize=()):
        oid = self._objective_id(objective, self._OBJECTIVES)
        k = int(n_samples)
        bufs = self._train(x, None, k, beta, optimizer, oid, eps, eps2, seed, materialize=materialize)
        return self._result(bufs, k, beta)


    def _z1_given_z2(self, z2, n, seed):
        z2 = self._dev(z2)
        if z2.shape[-1] != self._d2:
            raise ValueError(f"z2 must have last dim {self._d2}")
        mu, sd = self._block(_abi.BLOCK_DECODER21, z2, self._d1)
        return Normal(mu, sd).sample(n, seed=seed).tensor                   # pz1z2.sample(n) (A.2)

    def sample(self, z2, seed=None):
        """iwae2.IWAE.sample (iwae2.py:184-198): z1 ~ p(z1|z2), x ~ Bernoulli(decode(z1)) -> (x_sample, probs)."""
        seed = self._next_seed(seed)
        z1 = self._z1_given_z2(z2, 1, seed)[0]
        logits, _ = self._decode(z1, None, None, want_logits=True)
        return self._bernoulli(logits.tensor, seed)

    def conditional_sample(self, z2, seed=None, n=10):
        """iwae2.IWAE.conditional_sample (iwae2.py:200-213): ten z1 draws per fixed z2 -> ([10, ..., 784] x2)."""
        seed = self._next_seed(seed)
        z1 = self._z1_given_z2(z2, n, seed)
        logits, _ = self._decode(z1, None, None, want_logits=True)
        return self._bernoulli(logits.tensor, seed)

class IWAEDReG(IWAE):
    """tasks/task02.py:27-101: the doubly-reparameterised gradient estimator.  `call` returns the
    reference's keys (no `al`, no vae variants, plus `inference_loss`); `train_step` takes no objective:
    encoder gradients come from `inference_loss`, decoder gradients from `-iwae_elbo`."""

    _WANT = ("lpxz", "lpz", "lq1", "log_w", "al", "iwae", "vae", "eq14", "snis_z1", "dreg", "scalars")

    def _result(self, bufs, k, beta) -> Result:
        s = bufs["scalars"]
        lazy = self._lazy(bufs, k, beta, ("z1", "logits"))
        res = Result()
        res["iwae_elbo"] = DeviceValue(s[2])
        res["inference_loss"] = DeviceValue(s[4])
        res["z"] = lazy("z1")
        res["snis_z"] = DeviceValue(bufs["snis_z1"])
        res["logits"] = lazy("logits")
        res["lpxz"] = DeviceValue(bufs["lpxz"])
        res["lpz"] = DeviceValue(bufs["lpz"])
        res["lqzx"] = DeviceValue(bufs["lq1"])
        res["iwae_elbo_per_example"] = DeviceValue(bufs["iwae"])
        res["log_w"] = DeviceValue(bufs["log_w"])
        return res

    def train_step(self, x, n_samples, beta, optimizer, eps=None, seed=None, materialize=()):
        k = int(n_samples)
        bufs = self._train(x, None, k, beta, optimizer, _abi.OBJ["dreg"], eps, None, seed, materialize=materialize)
        return self._result(bufs, k, beta)

    @staticmethod
    def write_to_tensorboard(res, step, writer=None):
        """tasks/task02.py:103-108: iwae_elbo and the three term means (this dict has no vae_elbo)."""
        vals = {"Evaluation/iwae_elbo": float(res["iwae_elbo"])}
        for key in ("lpxz", "lqzx", "lpz"):
            vals[f"Evaluation/{key}"] = float(res[key].numpy().mean())
        if writer is not None:
            for name, v in vals.items():
                writer.add_scalar(name, v, step)
        return vals


class CIWAE(IWAE):
    """tasks/task04.py:101-188: label-conditioned IWAE with a learned prior p(z|y); `call(x, y, n_samples,
    beta)`.  `learned_prior=False` gives tasks/task05.py (N(0,1) prior)."""

    KIND = _abi.KIND_CIWAE
    TERM_KEYS = ("lpxzy", "lpzy", "lqzxy")

    def __init__(self, n_hidden, n_latent, learned_prior: bool = True, **kwargs):
        self.KIND = _abi.KIND_CIWAE if learned_prior else _abi.KIND_CIWAE_STD
        self._learned_prior = learned_prior
        if not learned_prior:
            self.TERM_KEYS = ("lpxzy", "lpz", "lqzxy")      # tasks/task05.py keeps the key "lpz" (N(0,1) prior)
        super().__init__(n_hidden, n_latent, **kwargs)

    def _make_config(self, n_hidden, n_latent, n_x, n_cond):
        return Config(self.KIND, int(n_x), int(n_cond or 10), int(n_hidden), 0, int(n_latent), 0)

    def _param_names(self):
        names = _block_names("encoder.encode_x_to_z") + _mlp3_names("decoder.decode_z_to_x")
        if self._learned_prior:
            names += _block_names("conditional_prior_network")
        return names

    def _build_submodels(self):
        super()._build_submodels()
        if self._learned_prior:
            self.conditional_prior_network = _SubModel(self, "conditional_prior_network")

    def call(self, x, y, n_samples, beta=1.0, eps=None, seed=None, img_offset=0, precision="fp32"):
        k = int(n_samples)
        engine = self._engine(precision, self._WANT, eps)
        if engine == "fp32":
            bufs = self._forward(x, y, k, beta, self._WANT, eps, None, seed, img_offset)
        else:
            bufs = self._forward_tc(x, y, k, beta, self._WANT, seed, img_offset, engine)
        bufs["_img_offset"] = img_offset
        return self._result(bufs, k, beta)

    __call__ = call

    def val_step(self, x, y, n_samples, beta, precision="auto"):
        return self.call(x, y, n_samples, beta, precision=precision)

    def train_step(self, x, y, n_samples, beta, optimizer, objective="vae_elbo", eps=None, seed=None, materialize=()):
        oid = self._objective_id(objective, self._OBJECTIVES)
        k = int(n_samples)
        bufs = self._train(x, y, k, beta, optimizer, oid, eps, None, seed, materialize=materialize)
        return self._result(bufs, k, beta)


    def sample(self, z, y, seed=None):
        """CIWAE.sample(z, y) (task04.py:190-209 / task05.py:185-199): every row of z decoded under the single
        label y; with the learned prior z is first mapped through p(z|y): z_new = loc + scale * z."""
        z = self._dev(z)
        rows = z.reshape(-1, self._d1)
        label = float(np.asarray(y).reshape(-1)[0])
        if self._learned_prior:
            onehot = torch.zeros(1, self._cfg.n_cond, device=self.device, dtype=_F32)
            onehot[0, int(label)] = 1.0
            mu, sd = self._block(_abi.BLOCK_PRIOR, onehot, self._d1)
            moved = self._alloc(rows.shape[0], 1, self._d1)                 # z_new = loc + scale * z (task04.py:196)
            with torch.cuda.device(self.device):
                check(lib.iwae_normal_sample(_ptr(mu), _ptr(sd), _ptr(rows.contiguous()), 1, rows.shape[0], self._d1, 0, 1,
                                             _ptr(moved), _stream_ptr()), "iwae_normal_sample")
            rows = moved.view(rows.shape[0], self._d1)
        yv = torch.full((rows.shape[0],), label, device=self.device, dtype=_F32)
        logits, _ = self._decode(rows, yv, None, want_logits=True)
        return self._bernoulli(logits.tensor.view(*z.shape[:-1], self._cfg.n_x), self._next_seed(seed))

class BasicBlock:
    """iwae1.BasicBlock(n_hidden, n_latent) as tasks/task04.py:108 constructs it directly.  In this
    implementation blocks live inside a model's flat parameter buffer, so a free-standing block is only a
    shape record; CIWAE creates its prior network itself."""

    def __init__(self, n_hidden, n_latent, **kwargs):
        self.n_hidden, self.n_latent = int(n_hidden), int(n_latent)
