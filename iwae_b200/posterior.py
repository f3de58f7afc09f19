"""task01's posterior tools (tasks/plot_task01.py:75-89): self-normalised importance weights and
sampling-importance-resampling, computed on the device through the C-ABI.

`sir_choice` reproduces `np.random.RandomState.choice(arange(L), size, replace=False, p=al)` bit-exactly
given the float64 uniforms that RandomState would draw (`rs.random_sample`), SURVEY.md A.7."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._abi import check, lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def softmax_f32(log_w, device=None) -> torch.Tensor:
    """scipy.special.softmax on a float32 vector (plot_task01.py:85)."""
    t = getattr(log_w, "tensor", log_w)
    if not isinstance(t, torch.Tensor):
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(t, dtype=np.float32)))
    t = t.to(device=device or (t.device if t.is_cuda else "cuda"), dtype=torch.float32).contiguous().reshape(-1)
    out = torch.empty_like(t)
    with torch.cuda.device(t.device):
        check(lib.iwae_softmax_f32(C.c_void_p(t.data_ptr()), t.numel(), C.c_void_p(out.data_ptr()), _stream()),
              "iwae_softmax_f32")
    return out


def sir_choice(al, size: int, uniforms, device=None):
    """Indices [size] int64 (device) and the number of uniforms consumed.

    al: normalised weights [L] (fp32); uniforms: float64 stream in RandomState.random_sample order -
    supply at least a few times `size` (each rejection round draws `size - found` more)."""
    a = getattr(al, "tensor", al)
    if not isinstance(a, torch.Tensor):
        a = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32)))
    dev = device or (a.device if a.is_cuda else "cuda")
    a = a.to(device=dev, dtype=torch.float32).contiguous().reshape(-1)
    u = uniforms if isinstance(uniforms, torch.Tensor) else torch.from_numpy(
        np.ascontiguousarray(np.asarray(uniforms, dtype=np.float64)))
    u = u.to(device=dev, dtype=torch.float64).contiguous()
    idx = torch.empty(size, dtype=torch.int64, device=dev)
    used = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(a.device):
        check(lib.iwae_sir(C.c_void_p(a.data_ptr()), a.numel(), C.c_void_p(u.data_ptr()), u.numel(), int(size),
                           C.c_void_p(idx.data_ptr()), C.c_void_p(used.data_ptr()), _stream()), "iwae_sir")
    n_used = int(used.item())
    if n_used < 0:
        raise ValueError("sir_choice: not enough uniforms supplied")
    return idx, n_used


def sir_from_randomstate(al, size: int, rs: np.random.RandomState, device=None):
    """Drop-in for `rs.choice(np.arange(L), size=size, replace=False, p=al)`: draws the uniforms from
    `rs` exactly as NumPy would (round by round) so `rs` ends in the same state."""
    a = getattr(al, "tensor", al)
    a_host = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float32)
    # Uniform consumption depends on the collisions of each round, so replay round by round: ask for
    # the first round's draws, run, and extend while the kernel reports it needs more.
    state = rs.get_state()
    budget = size
    while True:
        rs.set_state(state)
        u = rs.random_sample(budget)
        try:
            idx, used = sir_choice(a_host, size, u, device)
        except ValueError:
            budget *= 2
            continue
        rs.set_state(state)
        if used:
            # consume exactly what choice() consumes: successive random_sample(n_i) calls draw the same
            # stream as one random_sample(sum n_i)
            rs.random_sample(used)
        return idx
