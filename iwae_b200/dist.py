"""Multi-GPU: one process per GPU, `torch.distributed` (NCCL over NVLink on the GPU box, gloo in the
CPU tests) for the plumbing.

* k-sample LLH evaluation shards the IMAGES contiguously over ranks - independent units, weights
  replicated, Philox noise keyed on the global image index so per-example results do not depend on
  the world size - and ends with ONE all-reduce of the local fp64 sum.
* Data-parallel training shards the batch (each rank keeps all k samples of its images, so the
  log-sum-exp over k stays local), scales local gradients by 1/B_global inside the backward kernels and
  all-reduces the flat gradient once per step (~2 MB: latency-bound), then every rank applies the same
  Adam update.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of `n` units for `rank`: the first n % world ranks get one extra
    (20 images over 8 ranks -> 3,3,3,3,2,2,2,2)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def llh_eval_sharded(local_eval: Callable, n_images: int, device=None, gather: bool = False):
    """Run `local_eval(start, stop) -> (per_example [stop-start] tensor, local_sum 0-dim fp64 tensor)`
    on this rank's image shard and all-reduce the sum.  Returns (mean LLH as a 0-dim fp64 tensor,
    per-example vector of all images if gather else of this shard)."""
    rank, world = _world()
    start, stop = shard_range(n_images, rank, world)
    per, local_sum = local_eval(start, stop)
    total = local_sum.to(torch.float64).reshape(1).clone()
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
        if gather:
            # shards differ by at most one image: pad to the longest so every rank sends the same size
            sizes = [b - a for a, b in (shard_range(n_images, r, world) for r in range(world))]
            padded = torch.zeros(max(sizes), dtype=per.dtype, device=per.device)
            padded[:per.numel()] = per
            parts = [torch.empty_like(padded) for _ in range(world)]
            dist.all_gather(parts, padded)
            per = torch.cat([p[:n] for p, n in zip(parts, sizes)])
    return (total / max(n_images, 1)).reshape(()), per


def model_llh_sharded(model, x_all, n_samples: int = 5000, y_all=None, precision: str = "auto",
                      seed: int = 123, gather: bool = False):
    """The k=5000 test-set log-likelihood of main.py:170-184 over all ranks.  `x_all` [N,784] (host or
    device; every rank passes the same array and evaluates its own slice)."""
    n = x_all.shape[0]

    def local_eval(a, b):
        ys = None if y_all is None else y_all[a:b]
        if b == a:
            return (torch.empty(0, dtype=torch.float32, device=model.device),
                    torch.zeros((), dtype=torch.float64, device=model.device))
        return model.llh(x_all[a:b], n_samples, y=ys, precision=precision, seed=seed, img_offset=a)

    return llh_eval_sharded(local_eval, n, gather=gather)


def init_comm(model) -> Tuple[int, int]:
    """Give `model`'s context a library-owned NCCL communicator spanning the torch.distributed world
    (iwae_comm_unique_id on rank 0 -> broadcast of the 128-byte id -> iwae_comm_init on every rank).
    `dp_train_step` then runs the whole data-parallel step - gradient all-reduce included - inside the library
    (one CUDA graph per step).  Returns (rank, world)."""
    import ctypes as C
    from ._abi import check, lib
    rank, world = _world()
    buf = (C.c_uint8 * 128)()
    if rank == 0:
        check(lib.iwae_comm_unique_id(buf), "iwae_comm_unique_id")
    box = [bytes(buf)]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    ident = (C.c_uint8 * 128).from_buffer_copy(box[0])
    with torch.cuda.device(model.device):
        check(lib.iwae_comm_init(model._ctx, ident, rank, world), "iwae_comm_init")
    model._comm = (rank, world)
    return rank, world


def dp_train_step(model, x, n_samples: int, beta: float, optimizer, objective: str = "iwae_elbo", y=None,
                  seed: Optional[int] = None):
    """Data-parallel train step: local gradients of this rank's image shard (already divided by the
    GLOBAL batch size), one SUM all-reduce of the flat gradient, replicated Adam update.
    Every rank passes the same global batch x [B,784].

    With a library communicator (`init_comm(model)`) the step is ONE call into the library: forward, backward,
    ncclAllReduce and Keras-Adam on one stream, replayed from a CUDA graph (iwae_dp_train_step); returns None.
    Without one (the gloo CPU tests of the host logic, or no NCCL) it falls back to iwae_grads ->
    torch.distributed.all_reduce -> iwae_adam_step and returns the reduced gradient."""
    rank, world = _world()
    B = x.shape[0]
    a, b = shard_range(B, rank, world)
    if seed is None:
        seed = model._next_seed()
    if getattr(model, "_comm", None) == (rank, world):
        import ctypes as C
        from ._abi import OBJ, check, lib
        xs, ys = model._check_x(x[a:b], None if y is None else y[a:b]) if b > a else (None, None)
        m, v = optimizer._state_for(model)
        stream = torch.cuda.current_stream().cuda_stream
        ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        with torch.cuda.device(model.device):
            check(lib.iwae_dp_train_step(model._ctx, ptr(model.params), ptr(m), ptr(v), ptr(xs), ptr(ys), b - a, B,
                                         int(n_samples), float(beta), OBJ[objective], int(seed), a,
                                         float(optimizer.learning_rate.value), optimizer.beta_1, optimizer.beta_2,
                                         optimizer.epsilon, optimizer.iterations + 1, None, stream),
                  "iwae_dp_train_step")
        optimizer.iterations += 1
        model.mark_params_modified()
        return None
    if b > a:
        ys = None if y is None else y[a:b]
        g = model.grads(x[a:b], n_samples, beta, objective, seed=seed, y=ys, B_norm=B, img_offset=a)
    else:
        g = torch.zeros_like(model.params)
    if world > 1:
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
    optimizer.apply_gradients([(g, _FlatVar(model))])
    model.mark_params_modified()
    return g


class _FlatVar:
    """The whole flat parameter buffer as one optimizer variable."""

    def __init__(self, model):
        self.value = model.params
        self.offset = 0
