#!/usr/bin/env python
"""bench.py - BASELINE.json's headline metric on B200: k=5000 test-LLH image-samples/s (and k=50 train steps/s).

A "step" is one pass of the k=5000 importance-weighted log-likelihood estimator (main.py:170-184 of
nbip/IWAE) over 10,000 synthetic binarised 28x28 images with the 2-stochastic-layer IWAE
(784-200-200-100 / 100-100-100-50), sharded over the ranks (fixed total work: strong scaling), ending
with one all-reduce of the fp64 LLH sum and its read-back on the host.

  python bench.py [--gpus N --steps K --warmup W]            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                       the reference's CPU path (oracle port: the
                                                             TF2 stack is not installable here); imports
                                                             nothing of iwae_b200 and loads no .so of this repo

Prints ONE JSON line (see README / DESIGN.md §Measurement for the keys).  Besides the headline it carries
`train` (M2 of BASELINE.md: k=50 train steps/s for BASELINE configs[1], at EVERY N - data-parallel with the gradient
all-reduce inside the step for N > 1), `train_configs` (configs[0], [3], [4]), `cpu_baseline` and `cpu_baseline_train`.
"""
import argparse
import csv
import glob
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "k=5000 test-LLH image-samples/s"
UNIT = "image-samples/s"
FLOP_PER_ROW_2L = 563600.0      # 281,800 MAC per (image, sample) row, 2-layer [100,50]
FLOP_PER_ROW_1L = 433600.0      # 1-layer d=100

# training configurations of BASELINE.json (batch 20 everywhere, main.py:21)
TRAIN_CONFIGS = {
    "C2": dict(config="configs[1]: IWAE 2 stochastic layers (100/50) k=50", kind="iwae2", h=[200, 100], d=[100, 50], k=50,
               objective="iwae_elbo"),
    "C1": dict(config="configs[0]: IWAE 1 stochastic layer k=5 (784-200-200-50)", kind="iwae1", h=200, d=50, k=5,
               objective="iwae_elbo"),
    "C4": dict(config="configs[3]: DReG estimator (task02) 1 stochastic layer k=50", kind="dreg", h=200, d=100, k=50,
               objective="dreg"),
    "C5": dict(config="configs[4]: conditional IWAE (task04) label-conditioned prior k=50", kind="ciwae", h=200, d=100, k=50,
               objective="iwae_elbo"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=10000)
    ap.add_argument("--k", type=int, default=5000)
    ap.add_argument("--layers", type=int, default=2, choices=[1, 2])
    ap.add_argument("--precision", default="tc_f16_fast", choices=["fp32", "tc_f16", "tc_f16_fast"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of each cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--train-steps", type=int, default=1000)
    ap.add_argument("--ref-images", type=int, default=32, help="images per step of --impl reference")
    return ap.parse_args()


def config_dict(args, n_local_note="contiguous image shards", precision=None):
    arch = "784-200-200-100/100-100-100-50" if args.layers == 2 else "784-200-200-100"
    return {
        "workload": f"k={args.k} test-set LLH, {args.images} synthetic binarised 28x28 images, "
                    f"IWAE {args.layers} stochastic layer(s) {arch} (BASELINE.json configs[2])",
        "images": args.images, "k": args.k, "stochastic_layers": args.layers, "precision": precision or args.precision,
        "sharding": n_local_note, "l2": "256 MiB buffer rewritten between timed steps (L2 flush)",
        "noise": "Philox4x32-10 + Box-Muller (16-bit uniforms, 8 normals per call) in registers, keyed on the global image index",
    }


# ---------------------------------------------------------------------------------------------
# CPU legs: the reference's path as restated by the oracle (fp32, PyTorch-CPU, reference call shapes).
# Nothing below imports iwae_b200: the reference arm must not load this repo's native library.
# ---------------------------------------------------------------------------------------------
def cpu_threads(torch):
    # torch.distributed.run exports OMP_NUM_THREADS=1; the CPU arm uses every host core it can
    try:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        torch.set_num_threads(os.cpu_count() or 1)
    return torch.get_num_threads()


def cpu_llh_pass(params32, X, k, layers, torch, oracle, gen):
    """main.py:175-177: one eager call per image with B=1, k=5000, every dict entry computed."""
    total = 0.0
    for i in range(X.shape[0]):
        x = torch.from_numpy(X[i:i + 1])
        if layers == 2:
            e1 = torch.randn(k, 1, 100, generator=gen)
            e2 = torch.randn(k, 1, 50, generator=gen)
            r = oracle.iwae2_call(params32, x, k, 1.0, e1, e2)
        else:
            e1 = torch.randn(k, 1, 100, generator=gen)
            r = oracle.iwae1_call(params32, x, k, 1.0, e1)
        total += float(r["iwae_elbo"])
    return total


def cpu_setup(args):
    import torch
    cpu_threads(torch)
    from oracle import iwae_oracle as oracle
    P, X, _ = oracle.synthetic_data(max(args.ref_images, 1000), seed=123)
    rs = np.random.RandomState(123)
    if args.layers == 2:
        shapes = oracle.param_shapes("iwae2", [200, 100], [100, 50])
    else:
        shapes = oracle.param_shapes("iwae1", 200, 100)
    params = oracle.glorot_params(shapes, rs, out_bias=oracle.get_bias(P.mean(0)))
    p32 = oracle.to_torch(params, torch.float32)
    gen = torch.Generator().manual_seed(123)
    return torch, oracle, p32, X, gen


def cpu_baseline(args):
    torch, oracle, p32, X, gen = cpu_setup(args)
    with torch.no_grad():
        cpu_llh_pass(p32, X[:2], args.k, args.layers, torch, oracle, gen)        # warm-up
        t0 = time.perf_counter()
        cpu_llh_pass(p32, X[:2], args.k, args.layers, torch, oracle, gen)
        per_img = (time.perf_counter() - t0) / 2
        n = int(min(max(args.cpu_seconds / max(per_img, 1e-4), 4), 1000))
        t0 = time.perf_counter()
        cpu_llh_pass(p32, X[:n], args.k, args.layers, torch, oracle, gen)
        dt = time.perf_counter() - t0
    return {"value": n * args.k / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} images x k={args.k}, one eager B=1 call per image computing every dict entry "
                      f"(main.py:175-177), PyTorch-CPU fp32 oracle port standing in for TF2-CPU (not installable); "
                      f"{dt:.1f} s, host has {os.cpu_count()} logical cores"}


def cpu_baseline_train(args):
    """M2 of BASELINE.md §2 on the host cores: model.train_step (src/iwae1.py:153-162) restated as oracle forward +
    torch.autograd + Keras-rule Adam (eps = 1e-4, main.py:93), batch 20, >= 200 steps after 20 warm-up steps, for the
    shapes of BASELINE configs[0] (C1) and configs[1] (C2)."""
    import torch
    threads = cpu_threads(torch)
    from oracle import iwae_oracle as oracle
    P, X, _ = oracle.synthetic_data(1000, seed=123)
    out = {}
    for name in ("C1", "C2"):
        c = TRAIN_CONFIGS[name]
        shapes = oracle.param_shapes(c["kind"], c["h"], c["d"])
        params = oracle.glorot_params(shapes, np.random.RandomState(123), out_bias=oracle.get_bias(P.mean(0)))
        p = oracle.to_torch(params, torch.float32, requires_grad=True)
        m = {k_: torch.zeros_like(v) for k_, v in p.items()}
        v2 = {k_: torch.zeros_like(v) for k_, v in p.items()}
        gen = torch.Generator().manual_seed(1)
        k = c["k"]

        def step(t):
            x = torch.from_numpy(X[(t % 50) * 20:(t % 50) * 20 + 20])
            if c["kind"] == "iwae2":
                e1, e2 = torch.randn(k, 20, 100, generator=gen), torch.randn(k, 20, 50, generator=gen)
                g, _ = oracle.grads_objective(oracle.iwae2_call, p, "iwae_elbo", x, k, 1.0, e1, e2)
            else:
                e1 = torch.randn(k, 20, c["d"], generator=gen)
                g, _ = oracle.grads_objective(oracle.iwae1_call, p, "iwae_elbo", x, k, 1.0, e1)
            with torch.no_grad():
                for n_ in p:
                    th, m[n_], v2[n_] = oracle.keras_adam_step(p[n_], g[n_], m[n_], v2[n_], t + 1, 1e-3)
                    p[n_].copy_(th)

        for t in range(20):
            step(t)
        t0 = time.perf_counter()
        n_steps = 0
        while n_steps < 200 or (time.perf_counter() - t0 < min(args.cpu_seconds, 5.0) and n_steps < 5000):
            step(20 + n_steps)
            n_steps += 1
        dt = time.perf_counter() - t0
        out[name] = {"value": n_steps / dt, "unit": "steps/s", "steps": n_steps, "config": c["config"]}
    return {"metric": "k=50 / k=5 train steps/s on the host cores", "cores": threads, "kind": "port",
            "sample": "oracle forward + torch.autograd + Keras-rule Adam, batch 20, >= 200 steps after 20 warm-up "
                      "(main.py:135-143); PyTorch-CPU fp32 port standing in for TF2-CPU",
            **out}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch, oracle, p32, X, gen = cpu_setup(args)
    n = args.ref_images
    times = []
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            cpu_llh_pass(p32, X[:n], args.k, args.layers, torch, oracle, gen)
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
    dt = float(np.sum(times))
    value = n * args.k * args.steps / dt
    cb = {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
          "sample": f"each step = {n} images x k={args.k} (bounded sample of the 10,000-image workload), one eager "
                    f"B=1 call per image (main.py:175-177), PyTorch-CPU fp32 oracle port standing in for TF2-CPU"}
    assert "iwae_b200" not in sys.modules, "the reference arm must not import the product package"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, "CPU, rank 0 only", precision="f32 (PyTorch-CPU)"),
        "cpu_baseline": cb, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def ncu_traffic(layers):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the fused kernel on the default
    single-GPU workload, parsed from the newest committed `ncu` capture profiles/r*_ncu_traffic_full_launch*.csv
    (scripts/gpu_traffic.sh writes it; the file name carries the round).  None when no capture of this workload exists."""
    pat = "profiles/r*_ncu_traffic_full_launch.csv" if layers == 2 else "profiles/r*_ncu_traffic_full_launch_1layer.csv"
    files = sorted(glob.glob(os.path.join(ROOT, pat)))
    if not files:
        return None, None
    path = files[-1]
    rd = wr = None
    try:
        for row in csv.reader(open(path)):
            if len(row) < 3 or "iwae_eval_tc_kernel" not in ",".join(row):
                continue
            name = next((c for c in row if c.startswith("dram__bytes_")), None)
            if name is None:
                continue
            unit, val = row[-2], float(row[-1].replace(",", ""))
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            if name == "dram__bytes_read.sum":
                rd = val * mult
            elif name == "dram__bytes_write.sum":
                wr = val * mult
    except Exception:
        return None, None
    if rd is None or wr is None:
        return None, None
    return rd + wr, os.path.relpath(path, ROOT)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import iwae_b200 as ib
    from iwae_b200 import _abi
    from iwae_b200.dist import shard_range

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    P, X, Y = ib.data.synthetic_binarised(args.images, seed=123)
    bias = ib.data.get_bias(P.mean(0))
    if args.layers == 2:
        model = ib.IWAE2([200, 100], [100, 50], seed=123, out_bias=bias)
    else:
        model = ib.IWAE(200, 100, seed=123, out_bias=bias)
    a, b = shard_range(args.images, rank, world)
    x_dev = torch.from_numpy(X[a:b]).to(dev)
    x_pin = torch.from_numpy(X[a:b]).pin_memory()
    x_host = x_pin.numpy()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    total = torch.zeros(1, dtype=torch.float64, device=dev)
    seed = 123

    def step_device():
        per, s = model.llh(x_dev, args.k, precision=args.precision, seed=seed, img_offset=a)
        total.copy_(s.reshape(1))
        if world > 1:
            dist.all_reduce(total)
        return float(total.item()) / args.images            # scalar LLH on the host

    def step_host():
        per, s = model.llh_host(x_host, args.k, precision=args.precision, seed=seed, img_offset=a)
        if world > 1:
            total.copy_(torch.tensor([s], dtype=torch.float64))
            dist.all_reduce(total)
            s = float(total.item())
        return s / args.images

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, collect=None):
        for _ in range(warmup):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        out = None
        for e0, e1 in ev:
            flush.zero_()
            e0.record()
            out = fn()
            e1.record()
            if collect is not None:
                collect()
        barrier()
        ms = torch.tensor([sum(e0.elapsed_time(e1) for e0, e1 in ev)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    kernel_ms = []
    collect = (lambda: kernel_ms.append(model.last_eval_kernel_ms())) if args.precision != "fp32" else None
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _abi.lib.iwae_launch_count()
    ms_total, llh = timed(step_device, args.steps, args.warmup, collect)
    launches = (_abi.lib.iwae_launch_count() - l0) * args.steps // (args.steps + args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, llh_e2e = timed(step_host, args.steps, max(args.warmup, 1))

    rows_total = float(args.images) * args.k
    value = rows_total * args.steps / (ms_total * 1e-3)
    e2e = rows_total * args.steps / (ms_e2e * 1e-3)

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f16" if args.precision != "fp32" else "f32",
        "dtype_detail": ("fp16 tcgen05 operands, fp32 TMEM accumulation, fp32 log-density / log-sum-exp arithmetic"
                         if args.precision != "fp32" else "fp32 FFMA kernels"),
        "data": "synthetic", "config": config_dict(args), "impl": "ours",
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(args.images) * 784 * 4,
                "d2h_bytes_per_step": int(args.images) * 4 + 8 * world,
                "api": "IWAE2.llh_host -> iwae_llh_eval_host (pinned host x in, per-example LLH + fp64 sum out)"},
        "gpu_launches": int(launches), "clocks": clocks, "test_llh": llh, "test_llh_e2e": llh_e2e,
    }
    if kernel_ms:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained")
        which = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel runs for the whole multi-ms step)"
        if peak is None:
            peak, which = 1400.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"
        kms = float(np.mean(kernel_ms))
        flop_row = FLOP_PER_ROW_2L if args.layers == 2 else FLOP_PER_ROW_1L
        ach = (b - a) * args.k * flop_row / (kms * 1e-3) / 1e12
        default_workload = world == 1 and args.images == 10000 and args.k == 5000
        traffic, traffic_src = ncu_traffic(args.layers) if default_workload else (None, None)
        out["roofline"] = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                           "traffic": traffic, "traffic_unit": "bytes/launch (DRAM, ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                           "traffic_source": traffic_src,
                           "kernel": "tc::iwae_eval_tc_kernel", "kernel_ms": kms,
                           "kernel_share_of_step": kms / (ms_total / args.steps),
                           "flop_per_row": flop_row, "rows_per_launch": (b - a) * args.k, "peak_source": which,
                           "note": "fp16 operands; peak is the measured dense bf16 cuBLAS figure (same tensor rate)"}

    if not args.no_train:
        tr = bench_train(ib, torch, dist, rank, world, dev, X, Y, bias, "C2", args.train_steps)
        extra = {n: bench_train(ib, torch, dist, rank, world, dev, X, Y, bias, n, args.train_steps) for n in ("C1", "C4", "C5")}
        if rank == 0:
            out["train"] = tr
            out["train_configs"] = extra
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args)
        out["cpu_baseline_train"] = cpu_baseline_train(args)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def bench_train(ib, torch, dist, rank, world, dev, X, Y, bias, name, steps=1000, warmup=50):
    """Second half of BASELINE.json's metric: train steps/s, global batch 20 already on the device, forward + backward
    + Keras-Adam per step.  One GPU: model.train_step (the step replayed from a CUDA graph).  N GPUs: data-parallel over
    image shards (3,3,3,3,2,2,2,2 at N=8) with ONE ncclAllReduce of the flat gradient inside the step's graph
    (iwae_dp_train_step); timed with CUDA events between barriers, max over ranks."""
    from iwae_b200 import _abi
    from iwae_b200.dist import dp_train_step, init_comm
    c = TRAIN_CONFIGS[name]
    if c["kind"] == "iwae2":
        model = ib.IWAE2(c["h"], c["d"], seed=123, out_bias=bias)
    elif c["kind"] == "dreg":
        model = ib.IWAEDReG(c["h"], c["d"], seed=123, out_bias=bias)
    elif c["kind"] == "ciwae":
        model = ib.CIWAE(c["h"], c["d"], seed=123, out_bias=bias)
    else:
        model = ib.IWAE(c["h"], c["d"], seed=123, out_bias=bias)
    opt = ib.Adam(1e-3, epsilon=1e-4)
    k = c["k"]
    xb = [torch.from_numpy(X[i * 20:(i + 1) * 20]).to(dev) for i in range(16)]
    yb = [torch.from_numpy(Y[i * 20:(i + 1) * 20].astype(np.float32)).to(dev) for i in range(16)]
    if world > 1:
        init_comm(model)

    def step(i):
        j = i % 16
        if world > 1:
            dp_train_step(model, xb[j], k, 1.0, opt, c["objective"], y=yb[j] if c["kind"] == "ciwae" else None, seed=1000 + i)
            return None
        if c["kind"] == "dreg":
            return model.train_step(xb[j], k, 1.0, opt, seed=1000 + i)
        if c["kind"] == "ciwae":
            return model.train_step(xb[j], yb[j], k, 1.0, opt, objective=c["objective"], seed=1000 + i)
        return model.train_step(xb[j], k, 1.0, opt, objective=c["objective"], seed=1000 + i)

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = _abi.lib.iwae_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = None
    for i in range(steps):
        res = step(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    launches = (_abi.lib.iwae_launch_count() - l0) / steps
    allreduce_us = None
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # the collective by itself: the same ~2 MB sum all-reduce, back to back on one stream
        import ctypes as C
        g = torch.zeros(model.n_params, dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(20):
            _abi.check(_abi.lib.iwae_allreduce_sum_f32(model._ctx, C.c_void_p(g.data_ptr()), model.n_params, st))
        torch.cuda.synchronize(); dist.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(200):
            _abi.check(_abi.lib.iwae_allreduce_sum_f32(model._ctx, C.c_void_p(g.data_ptr()), model.n_params, st))
        a1.record()
        torch.cuda.synchronize()
        t = torch.tensor([a0.elapsed_time(a1) / 200 * 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        allreduce_us = float(t.item())
    ms = float(ms.item())
    rec = {"metric": f"k={k} train steps/s", "value": steps / (ms * 1e-3), "unit": "steps/s", "ms_per_step": ms / steps,
           "config": c["config"], "global_batch": 20, "k": k, "objective": c["objective"], "n_gpus": world,
           "parallelism": "single GPU, step replayed from a CUDA graph" if world == 1 else
                          f"data-parallel over image shards, dp{world}; ncclAllReduce of the {model.n_params}-float gradient inside the step's CUDA graph",
           "launches_per_step": launches, "n_params": int(model.n_params),
           "allreduce_us": allreduce_us,
           "note": "latency-bound (north_star): ~1.7 GFLOP per C2 step; fp32 FFMA kernels"}
    if res is not None:
        rec["last_iwae_elbo"] = float(res["iwae_elbo"])
    # tear the context (and its NCCL communicator) down at the same program point on every rank, not whenever the
    # garbage collector gets to it (models hold reference cycles)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    model.close()
    return rec


def main():
    args = parse()
    if os.environ.get("IWAE_BENCH_DEBUG"):        # where is a rank stuck: dump every thread's stack after N seconds and exit
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["IWAE_BENCH_DEBUG"]), exit=True)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
