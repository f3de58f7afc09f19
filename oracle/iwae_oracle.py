"""CPU oracle for the importance-weighted ELBO hot path of nbip/IWAE.

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped package (`iwae_b200/`, `src/`) may import
this module; only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` do, and there only as the checker / the timed CPU baseline.

PARITY UNPINNED.  The reference executes all arithmetic inside TensorFlow 2 / TensorFlow-Probability /
Keras (un-vendored, version-unpinned; none importable in this image, no network) and ships no tests,
golden vectors or fixtures.  This file is therefore a *restatement* of the
reference's algorithm, following the reference call sites line by line (cited below) with the
third-party op semantics of SURVEY.md Appendix A.  What IS pinned:
  * the SIR resampler is checked bit-exactly against the real `numpy.random.RandomState.choice`
    (NumPy is installed, so this piece of the reference path (`tasks/plot_task01.py:88`) runs as-is);
  * log-probabilities are cross-checked against `torch.distributions.{Normal,Bernoulli}` and
    gradients against `torch.autograd` (tests/test_oracle.py).

Everything takes the noise `eps` as an explicit input ("injected-eps mode"), so the oracle and the
CUDA path see identical inputs.  `dtype=torch.float64` is the truth; `dtype=torch.float32` follows the
reference's fp32 op order and estimates the reference's own rounding noise.

Tensor layout is the reference's: sample-major `[k, B, .]`, Dense kernels `[in, out]`, `y = x @ W + b`.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

LOG2PI_HALF = 0.5 * math.log(2.0 * math.pi)

# --------------------------------------------------------------------------------------------
# Parameter containers (canonical Keras order)
# --------------------------------------------------------------------------------------------


def _block_names(prefix: str) -> List[str]:
    # BasicBlock attribute order l1, l2, lmu, lstd (iwae1.py:31-34, iwae2.py:32-35)
    out = []
    for lyr in ("l1", "l2", "lmu", "lstd"):
        out += [f"{prefix}.{lyr}.kernel", f"{prefix}.{lyr}.bias"]
    return out


def _mlp3_names(prefix: str) -> List[str]:
    # keras.Sequential of three Dense (iwae1.py:70-77, iwae2.py:80-87)
    out = []
    for i in range(3):
        out += [f"{prefix}.dense{i}.kernel", f"{prefix}.dense{i}.bias"]
    return out


def param_shapes(kind: str, n_hidden, n_latent, n_x: int = 784, n_cond: int = 0) -> "OrderedDict[str, Tuple[int, ...]]":
    """Shapes of `model.trainable_weights` in Keras creation order.

    kind: 'iwae1' (iwae1.py:88-96), 'iwae2' (iwae2.py:99-107), 'ciwae' (task04.py:101-108; learned
    conditional prior), 'ciwae_std' (task05.py: CIWAE with N(0,1) prior).
    """
    shapes: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def block(prefix, n_in, h, d):
        for name, shp in zip(_block_names(prefix),
                             [(n_in, h), (h,), (h, h), (h,), (h, d), (d,), (h, d), (d,)]):
            shapes[name] = shp

    def mlp3(prefix, n_in, h, n_out):
        for name, shp in zip(_mlp3_names(prefix),
                             [(n_in, h), (h,), (h, h), (h,), (h, n_out), (n_out,)]):
            shapes[name] = shp

    if kind == "iwae1":
        h, d = int(n_hidden), int(n_latent)
        block("encoder.encode_x_to_z", n_x, h, d)
        mlp3("decoder.decode_z_to_x", d, h, n_x)
    elif kind in ("ciwae", "ciwae_std"):
        h, d = int(n_hidden), int(n_latent)
        nc = n_cond or 10
        block("encoder.encode_x_to_z", n_x + nc, h, d)
        mlp3("decoder.decode_z_to_x", d + nc, h, n_x)
        if kind == "ciwae":
            block("conditional_prior_network", nc, h, d)
    elif kind == "iwae2":
        h1, h2 = int(n_hidden[0]), int(n_hidden[1])
        d1, d2 = int(n_latent[0]), int(n_latent[1])
        block("encoder.encode_x_to_z1", n_x, h1, d1)
        block("encoder.encode_z1_to_z2", d1, h2, d2)
        block("decoder.decode_z2_to_z1", d2, h2, d1)
        mlp3("decoder.decode_z1_to_x", d1, h1, n_x)
    else:
        raise ValueError(kind)
    return shapes


def get_bias(train_mean: np.ndarray) -> np.ndarray:
    """utils.get_bias (utils.py:11-23): logit of the clipped mean train pixel (float64)."""
    return -np.log(1.0 / np.clip(train_mean, 0.001, 0.999) - 1.0)


def glorot_params(shapes: "OrderedDict[str, Tuple[int, ...]]", rs: np.random.RandomState,
                  out_bias: Optional[np.ndarray] = None, scale: float = 1.0) -> "OrderedDict[str, np.ndarray]":
    """Keras default Dense init: kernel glorot_uniform, bias zeros; the 784-way output bias is
    `get_bias()` (iwae1.py:74-75).  Values come from our own seeded generator (TF's stream is not
    reproducible outside TF) - only order and shapes are inherited."""
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    names = list(shapes)
    for name in names:
        shp = shapes[name]
        if len(shp) == 2:
            lim = math.sqrt(6.0 / (shp[0] + shp[1])) * scale
            out[name] = rs.uniform(-lim, lim, size=shp).astype(np.float32)
        else:
            out[name] = np.zeros(shp, dtype=np.float32)
    if out_bias is not None:
        last_bias = [n for n in names if n.endswith("dense2.bias")][-1]
        out[last_bias] = np.asarray(out_bias, dtype=np.float64).astype(np.float32)
    return out


def synthetic_data(n: int, seed: int = 123, n_x: int = 784):
    """Synthetic binarised 28x28 stand-in for MNIST: intensities rand**4, binarised
    exactly like utils.bernoullisample (utils.py:26-27)."""
    rs = np.random.RandomState(seed)
    P = rs.rand(n, n_x) ** 4
    X = rs.binomial(1, P, size=P.shape).astype("float32")
    y = rs.randint(0, 10, n)
    return P, X, y


def to_torch(params: Dict[str, np.ndarray], dtype=torch.float64, requires_grad: bool = False):
    out = OrderedDict()
    for k, v in params.items():
        t = torch.tensor(np.asarray(v), dtype=dtype)
        if requires_grad:
            t.requires_grad_(True)
        out[k] = t
    return out


# --------------------------------------------------------------------------------------------
# Third-party op semantics
# --------------------------------------------------------------------------------------------


def normal_log_prob(x, loc, scale):
    """tfd.Normal.log_prob (A.2): TFP divides both terms by scale before differencing."""
    return -0.5 * (x / scale - loc / scale) ** 2 - (LOG2PI_HALF + torch.log(scale))


def std_normal_log_prob(z):
    return -0.5 * z * z - LOG2PI_HALF


def bernoulli_log_prob(logits, x):
    """tfd.Bernoulli(logits).log_prob(x) (A.3) = -(1-x) softplus(l) - x softplus(-l)."""
    sp = torch.nn.functional.softplus
    return -(1.0 - x) * sp(logits) - x * sp(-logits)


def kl_normal_std(loc, scale):
    """tfd.kl_divergence(N(mu,sigma), N(0,1)) (A.2, iwae1.py:116)."""
    ls = torch.log(scale)
    return 0.5 * loc * loc + 0.5 * torch.expm1(2.0 * ls) - ls


def logmeanexp(log_w, axis=0):
    """utils.logmeanexp (utils.py:6-8): max without keepdims, broadcast [k,B]-[B]."""
    m = torch.amax(log_w, dim=axis)
    return torch.log(torch.mean(torch.exp(log_w - m), dim=axis)) + m


def dense(p, name, x, act=None):
    y = x @ p[name + ".kernel"] + p[name + ".bias"]
    if act == "tanh":
        return torch.tanh(y)
    if act == "exp":
        return torch.exp(y)
    return y


def basic_block(p, prefix, inp):
    """BasicBlock.call (iwae1.py:36-44): returns (loc, scale) of tfd.Normal(q_mu, q_std + 1e-6)."""
    h1 = dense(p, prefix + ".l1", inp, "tanh")
    h2 = dense(p, prefix + ".l2", h1, "tanh")
    mu = dense(p, prefix + ".lmu", h2)
    std = dense(p, prefix + ".lstd", h2, "exp")
    return mu, std + 1e-6


def mlp3(p, prefix, z):
    """Decoder Sequential (iwae1.py:70-77): tanh, tanh, linear 784 logits."""
    h1 = dense(p, prefix + ".dense0", z, "tanh")
    h2 = dense(p, prefix + ".dense1", h1, "tanh")
    return dense(p, prefix + ".dense2", h2)


def _objectives(log_w, extra=None):
    """The ELBO variants shared by iwae1/iwae2/CIWAE (iwae1.py:119-139)."""
    res = {}
    per_ex_iwae = logmeanexp(log_w, 0)
    res["vae_elbo"] = torch.mean(torch.mean(log_w, dim=0), dim=-1)
    res["iwae_elbo"] = torch.mean(per_ex_iwae, dim=-1)
    m = torch.amax(log_w, dim=0, keepdim=True)
    w = torch.exp(log_w - m)
    wn = (w / torch.sum(w, dim=0, keepdim=True)).detach()
    res["iwae_eq14"] = torch.mean(torch.sum(wn * log_w, dim=0))
    res["al"] = torch.softmax(log_w, dim=0)
    # additions over the reference dict: per-example vectors before reduce_mean(axis=-1)
    res["iwae_elbo_per_example"] = per_ex_iwae
    res["vae_elbo_per_example"] = torch.mean(log_w, dim=0)
    res["log_w"] = log_w
    return res


def iwae1_call(p, x, n_samples: int, beta: float = 1.0, eps=None, prefix_enc="encoder.encode_x_to_z",
               prefix_dec="decoder.decode_z_to_x"):
    """iwae1.IWAE.call (iwae1.py:98-151).  x [B,784]; eps [k,B,d] ~ N(0,1) injected."""
    mu, std = basic_block(p, prefix_enc, x)                 # iwae1.py:57
    z = eps * std + mu                                       # qzx.sample(k) (iwae1.py:59, A.2)
    assert z.shape[0] == n_samples
    logits = mlp3(p, prefix_dec, z)                          # iwae1.py:81
    lpz = torch.sum(std_normal_log_prob(z), dim=-1)          # iwae1.py:107
    lqzx = torch.sum(normal_log_prob(z, mu, std), dim=-1)    # iwae1.py:109
    lpxz = torch.sum(bernoulli_log_prob(logits, x), dim=-1)  # iwae1.py:111
    log_w = lpxz + beta * (lpz - lqzx)                       # iwae1.py:113
    kl = torch.sum(kl_normal_std(mu, std), dim=-1)           # iwae1.py:116
    res = _objectives(log_w)
    res["vae_elbo_kl"] = torch.mean(lpxz) - beta * torch.mean(kl)   # iwae1.py:121
    res["snis_z"] = torch.sum(res["al"][:, :, None] * z, dim=0)     # iwae1.py:139
    res.update(z=z, logits=logits, lpxz=lpxz, lpz=lpz, lqzx=lqzx, q_mu=mu, q_std=std, kl=kl)
    return res


def iwae2_call(p, x, n_samples: int, beta: float = 1.0, eps1=None, eps2=None):
    """iwae2.IWAE.call (iwae2.py:109-167).  eps1 [k,B,d1], eps2 [k,B,d2]; beta is ignored (iwae2.py:128)."""
    mu1, std1 = basic_block(p, "encoder.encode_x_to_z1", x)          # iwae2.py:59
    z1 = eps1 * std1 + mu1                                            # iwae2.py:61
    assert z1.shape[0] == n_samples
    mu2, std2 = basic_block(p, "encoder.encode_z1_to_z2", z1)        # iwae2.py:63
    z2 = eps2 * std2 + mu2                                            # iwae2.py:65 (.sample())
    pmu, pstd = basic_block(p, "decoder.decode_z2_to_z1", z2)        # iwae2.py:90
    logits = mlp3(p, "decoder.decode_z1_to_x", z1)                   # iwae2.py:92
    lpz2 = torch.sum(std_normal_log_prob(z2), dim=-1)                 # iwae2.py:116
    lqz2z1 = torch.sum(normal_log_prob(z2, mu2, std2), dim=-1)        # iwae2.py:118
    lpz1z2 = torch.sum(normal_log_prob(z1, pmu, pstd), dim=-1)        # iwae2.py:120
    lqz1x = torch.sum(normal_log_prob(z1, mu1, std1), dim=-1)         # iwae2.py:122
    lpxz1 = torch.sum(bernoulli_log_prob(logits, x), dim=-1)          # iwae2.py:124
    log_w = lpxz1 + lpz1z2 + lpz2 - lqz1x - lqz2z1                    # iwae2.py:126
    res = _objectives(log_w)
    res["snis_z1"] = torch.sum(res["al"][:, :, None] * z1, dim=0)     # iwae2.py:150
    res["snis_z2"] = torch.sum(res["al"][:, :, None] * z2, dim=0)     # iwae2.py:152
    res.update(z1=z1, z2=z2, logits=logits, lpxz1=lpxz1, lpz1z2=lpz1z2, lpz2=lpz2, lqz1x=lqz1x,
               lqz2z1=lqz2z1, q_mu=mu1, q_std=std1)
    return res


def dreg_call(p, x, n_samples: int, beta: float = 1.0, eps=None):
    """IWAEDReG.call (task02.py:34-85)."""
    mu, std = basic_block(p, "encoder.encode_x_to_z", x)
    z = eps * std + mu
    logits = mlp3(p, "decoder.decode_z_to_x", z)
    lpz = torch.sum(std_normal_log_prob(z), dim=-1)
    lqzx = torch.sum(normal_log_prob(z, mu, std), dim=-1)
    lpxz = torch.sum(bernoulli_log_prob(logits, x), dim=-1)
    log_w = lpxz + beta * (lpz - lqzx)                                # task02.py:49
    per_ex = logmeanexp(log_w, 0)
    iwae_elbo = torch.mean(per_ex, dim=-1)                            # task02.py:53
    al = torch.softmax(log_w, dim=0)                                  # task02.py:56
    snis_z = torch.sum(al[:, :, None] * z, dim=0)
    # qzx.scale already holds exp(a)+1e-6; a second +1e-6 is added (task02.py:61-63)
    lqzx_stopped = torch.sum(normal_log_prob(z, mu.detach(), std.detach() + 1e-6), dim=-1)
    sq = al.detach() ** 2                                             # task02.py:67-68
    stopped_log_w = lpz + lpxz - lqzx_stopped                         # task02.py:70 (no beta)
    dreg = torch.sum(sq * stopped_log_w, dim=0)                       # task02.py:73
    inference_loss = -torch.mean(dreg, dim=-1)                        # task02.py:76
    return {"iwae_elbo": iwae_elbo, "inference_loss": inference_loss, "z": z, "snis_z": snis_z,
            "logits": logits, "lpxz": lpxz, "lpz": lpz, "lqzx": lqzx, "al": al, "log_w": log_w,
            "iwae_elbo_per_example": per_ex}


def one_hot(y, depth=10, dtype=torch.float64):
    y = torch.as_tensor(np.asarray(y).astype(np.uint8).astype(np.int64))     # tf.cast(y, uint8) (task04.py:112)
    return torch.nn.functional.one_hot(y, depth).to(dtype)


def ciwae_call(p, x, y, n_samples: int, beta: float = 1.0, eps=None, learned_prior: bool = True):
    """CIWAE.call (task04.py:110-173); learned_prior=False gives task05's N(0,1) prior."""
    yo = one_hot(y, 10, x.dtype)
    xy = torch.cat([x, yo], dim=-1)                                   # task04.py:115
    mu, std = basic_block(p, "encoder.encode_x_to_z", xy)
    z = eps * std + mu
    zy = torch.cat([z, yo[None].expand(n_samples, -1, -1)], dim=-1)   # task04.py:119
    logits = mlp3(p, "decoder.decode_z_to_x", zy)
    if learned_prior:
        pmu, pstd = basic_block(p, "conditional_prior_network", yo)   # task04.py:123
        lpzy = torch.sum(normal_log_prob(z, pmu, pstd), dim=-1)       # task04.py:129
    else:
        lpzy = torch.sum(std_normal_log_prob(z), dim=-1)
    lqzxy = torch.sum(normal_log_prob(z, mu, std), dim=-1)
    lpxzy = torch.sum(bernoulli_log_prob(logits, x), dim=-1)
    log_w = lpxzy + beta * (lpzy - lqzxy)                             # task04.py:135
    kl = torch.sum(kl_normal_std(mu, std), dim=-1)                    # task04.py:138 (global pz = N(0,1))
    res = _objectives(log_w)
    res["vae_elbo_kl"] = torch.mean(lpxzy) - beta * torch.mean(kl)
    res["snis_z"] = torch.sum(res["al"][:, :, None] * z, dim=0)
    res.update(z=z, logits=logits, lpxzy=lpxzy, lqzxy=lqzxy, q_mu=mu, q_std=std, kl=kl)
    res["lpzy" if learned_prior else "lpz"] = lpzy                    # task04.py:169 'lpzy' / task05.py:167 'lpz'
    return res


# --------------------------------------------------------------------------------------------
# Gradients (reference: tape.gradient of -res[objective], iwae1.py:155-159) and Keras Adam
# --------------------------------------------------------------------------------------------


def grads_objective(call_fn, p, objective: str, *args, **kw):
    """d(-res[objective]) / d(params) in canonical order.  `p` tensors need requires_grad."""
    res = call_fn(p, *args, **kw)
    loss = -res[objective]       # unknown objective -> KeyError, the reference's convention
    g = torch.autograd.grad(loss, list(p.values()), allow_unused=True)
    out = OrderedDict()
    for (name, t), gi in zip(p.items(), g):
        out[name] = torch.zeros_like(t) if gi is None else gi
    return out, res


def grads_dreg(p, x, n_samples, beta, eps):
    """IWAEDReG.train_step (task02.py:87-101): encoder grads from inference_loss, decoder grads
    from -iwae_elbo."""
    res = dreg_call(p, x, n_samples, beta, eps)
    enc = [k for k in p if k.startswith("encoder.")]
    dec = [k for k in p if k.startswith("decoder.")]
    ge = torch.autograd.grad(res["inference_loss"], [p[k] for k in enc], retain_graph=True)
    gd = torch.autograd.grad(-res["iwae_elbo"], [p[k] for k in dec])
    out = OrderedDict()
    for k, g in zip(enc, ge):
        out[k] = g
    for k, g in zip(dec, gd):
        out[k] = g
    return out, res


def keras_adam_step(theta, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-4):
    """Keras Adam (A.6): epsilon added to the UNcorrected sqrt(v); t is the step AFTER increment.
    Works on numpy arrays or torch tensors; returns (theta, m, v)."""
    m = beta1 * m + (1.0 - beta1) * g
    v = beta2 * v + (1.0 - beta2) * g * g
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    sq = torch.sqrt(v) if isinstance(v, torch.Tensor) else np.sqrt(v)
    theta = theta - lr_t * m / (sq + eps)
    return theta, m, v


def lr_schedule():
    """main.py:44-51: {epoch: lr}; 3280 epochs in total."""
    epochs, d = 0, {}
    for i in range(8):
        d[epochs] = 0.001 * 10 ** (-i / 7)
        epochs += 3 ** i
    return d, epochs


# --------------------------------------------------------------------------------------------
# SIR (plot_task01.py:75-89): NumPy legacy RandomState.choice(replace=False, p) restated (A.7)
# --------------------------------------------------------------------------------------------


def softmax_f32(log_w: np.ndarray) -> np.ndarray:
    """scipy.special.softmax on a float32 vector (plot_task01.py:85): exp(x - logsumexp(x))."""
    from scipy.special import softmax
    return softmax(np.asarray(log_w, dtype=np.float32), axis=0)


def sir_choice(al: np.ndarray, size: int, uniforms: np.ndarray):
    """Restatement of `RandomState.choice(arange(L), size, replace=False, p=al)` consuming float64
    uniforms from `uniforms` in stream order.  Returns (idx int64 [size], n_uniforms_consumed)."""
    p = np.array(al, dtype=np.float64, copy=True)
    found = np.zeros(size, dtype=np.int64)
    n_uniq, pos = 0, 0
    while n_uniq < size:
        n = size - n_uniq
        x = np.asarray(uniforms[pos:pos + n], dtype=np.float64)
        if x.shape[0] < n:
            raise ValueError("not enough uniforms supplied")
        pos += n
        if n_uniq > 0:
            p[found[:n_uniq]] = 0
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        new = cdf.searchsorted(x, side="right")
        _, first = np.unique(new, return_index=True)
        first.sort()
        new = new.take(first)
        found[n_uniq:n_uniq + new.size] = new
        n_uniq += new.size
    return found, pos


# --------------------------------------------------------------------------------------------
# Philox4x32-10 + Box-Muller (the device RNG of the CUDA path, restated for the tests)
# --------------------------------------------------------------------------------------------

_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = np.uint32(0x9E3779B9)
_PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """Philox4x32-10 (Salmon et al. 2011).  ctr [...,4] uint32, key [...,2] uint32 -> [...,4] uint32."""
    c = [ctr[..., i].astype(np.uint32) for i in range(4)]
    k0 = key[..., 0].astype(np.uint32).copy()
    k1 = key[..., 1].astype(np.uint32).copy()
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PHILOX_M0 * c[0].astype(np.uint64)
            p1 = _PHILOX_M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = (k0 + _PHILOX_W0).astype(np.uint32)
            k1 = (k1 + _PHILOX_W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def philox_normal(seed: int, stream: int, img_offset: int, B: int, k: int, d: int) -> np.ndarray:
    """eps [k,B,d] float32 as generated by the CUDA path (iwae_b200/csrc/rng.cuh):
    counter = (sample s, dim-octet g, global image index, stream id), key = (seed lo, seed hi);
    each Philox call gives 4 uint32; word h -> one Box-Muller pair for dims 8g+2h, 8g+2h+1 from two 16-bit
    uniforms: u1 = (hi16 + 0.5) / 65536, u2 = (lo16 + 0.5) / 65536;
    n0 = sqrt(-2 ln u1) cos(2 pi u2), n1 = sqrt(-2 ln u1) sin(2 pi u2)."""
    ng = (d + 7) // 8
    s = np.arange(k, dtype=np.uint32)[:, None, None]
    b = (np.arange(B, dtype=np.uint64) + np.uint64(img_offset)).astype(np.uint32)[None, :, None]
    g = np.arange(ng, dtype=np.uint32)[None, None, :]
    ctr = np.stack(np.broadcast_arrays(s, g, b, np.uint32(stream)), axis=-1)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    key = np.broadcast_to(key, ctr.shape[:-1] + (2,))
    r = philox4x32_10(ctr, key)
    out = np.empty((k, B, ng * 8), dtype=np.float64)
    for h in range(4):
        w = r[..., h]
        u1 = ((w >> np.uint32(16)).astype(np.float64) + 0.5) * 2.0 ** -16
        u2 = ((w & np.uint32(0xFFFF)).astype(np.float64) + 0.5) * 2.0 ** -16
        rad = np.sqrt(-2.0 * np.log(u1))
        out[..., 2 * h::8] = rad * np.cos(2.0 * np.pi * u2)
        out[..., 2 * h + 1::8] = rad * np.sin(2.0 * np.pi * u2)
    return out[..., :d].astype(np.float32)


def philox_bernoulli(p: np.ndarray, seed: int, epoch: int = 0) -> np.ndarray:
    """utils.bernoullisample (src/utils.py:26-27) as the CUDA path draws it (kernels_f32.cu,
    bernoulli_sample_kernel): flat element i uses word i%4 of Philox4x32-10 with counter
    (i//4 lo, i//4 hi, epoch, 0x42), key = (seed lo, seed hi); u = (r >> 8) * 2^-24 in fp32; x = 1 if u < p."""
    flat = np.asarray(p, dtype=np.float32).reshape(-1)
    n = flat.size
    nq = (n + 3) // 4
    q = np.arange(nq, dtype=np.uint64)
    ctr = np.stack([(q & np.uint64(0xFFFFFFFF)).astype(np.uint32), (q >> np.uint64(32)).astype(np.uint32),
                    np.full(nq, epoch & 0xFFFFFFFF, dtype=np.uint32), np.full(nq, 0x42, dtype=np.uint32)], axis=-1)
    key = np.broadcast_to(np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32), (nq, 2))
    r = philox4x32_10(ctr, key).reshape(-1)[:n]
    u = ((r >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)
    return (u < flat).astype(np.float32).reshape(np.shape(p))
