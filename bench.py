#!/usr/bin/env python
"""bench.py - BASELINE.json's headline metric on B200: k=5000 test-LLH image-samples/s.

A "step" is one pass of the k=5000 importance-weighted log-likelihood estimator (main.py:170-184 of
nbip/IWAE) over 10,000 synthetic binarised 28x28 images with the 2-stochastic-layer IWAE
(784-200-200-100 / 100-100-100-50), sharded over the ranks (fixed total work: strong scaling), ending
with one all-reduce of the fp64 LLH sum and its read-back on the host.

  python bench.py [--gpus N --steps K --warmup W]            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                       the reference's CPU path (oracle port: the
                                                             TF2 stack is not installable here)

Prints ONE JSON line (see README / DESIGN.md §Measurement for the keys).
"""
import argparse
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
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--ref-images", type=int, default=32, help="images per step of --impl reference")
    return ap.parse_args()


def synthetic(n_images):
    from iwae_b200 import data
    P, X, _ = data.synthetic_binarised(n_images, seed=123)
    return P, X


def config_dict(args, n_local_note="contiguous image shards"):
    arch = "784-200-200-100/100-100-100-50" if args.layers == 2 else "784-200-200-100"
    return {
        "workload": f"k={args.k} test-set LLH, {args.images} synthetic binarised 28x28 images, "
                    f"IWAE {args.layers} stochastic layer(s) {arch} (BASELINE.json configs[2])",
        "images": args.images, "k": args.k, "stochastic_layers": args.layers, "precision": args.precision,
        "sharding": n_local_note, "l2": "256 MiB buffer rewritten between timed steps (L2 flush)",
        "noise": "Philox4x32-10 + Box-Muller in registers, keyed on the global image index",
    }


# ---------------------------------------------------------------------------------------------
# CPU leg: the reference's path as restated by the oracle (fp32, PyTorch-CPU, reference call shape:
# one eager call per image with B=1, k=5000, every dict entry computed - main.py:175-177)
# ---------------------------------------------------------------------------------------------
def cpu_llh_pass(params32, X, k, layers, torch, oracle, gen):
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
    # torch.distributed.run exports OMP_NUM_THREADS=1; the CPU arm uses every host core it can
    try:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        torch.set_num_threads(os.cpu_count() or 1)
    from oracle import iwae_oracle as oracle
    P, X = synthetic(max(args.ref_images, 1000))
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
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args, "CPU, rank 0 only"),
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

    P, X = synthetic(args.images)
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
        # DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of this kernel on the default
        # single-GPU workload, from the ncu capture committed as profiles/r01_v5_ncu_traffic_full_launch.csv;
        # other workloads were not captured
        traffic = 40567040 + 34816 if (world == 1 and args.images == 10000 and args.k == 5000 and args.layers == 2) else None
        out["roofline"] = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                           "traffic": traffic, "traffic_unit": "bytes/launch (DRAM; L2->SM weight streaming is 221 GB/launch)",
                           "kernel": "tc::iwae_eval_tc_kernel", "kernel_ms": kms,
                           "kernel_share_of_step": kms / (ms_total / args.steps),
                           "flop_per_row": flop_row, "rows_per_launch": (b - a) * args.k, "peak_source": which,
                           "note": "fp16 operands; peak is the measured dense bf16 cuBLAS figure (same tensor rate)"}

    if rank == 0 and world == 1 and not args.no_train:
        out["train"] = bench_train(ib, torch, X, bias)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def bench_train(ib, torch, X, bias, steps=1000, warmup=50):
    """Second half of BASELINE.json's metric: k=50 train steps/s, batch 20, 2 stochastic layers
    (configs[1]); batch already on the device, forward + backward + Keras-Adam per step."""
    from iwae_b200 import _abi
    model = ib.IWAE2([200, 100], [100, 50], seed=123, out_bias=bias)
    opt = ib.Adam(1e-3, epsilon=1e-4)
    xb = [torch.from_numpy(X[i * 20:(i + 1) * 20]).cuda() for i in range(16)]
    for i in range(warmup):
        model.train_step(xb[i % 16], 50, 1.0, opt, objective="iwae_elbo")
    torch.cuda.synchronize()
    l0 = _abi.lib.iwae_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        res = model.train_step(xb[i % 16], 50, 1.0, opt, objective="iwae_elbo")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"metric": "k=50 train steps/s", "value": steps / (ms * 1e-3), "unit": "steps/s", "batch": 20, "k": 50,
            "stochastic_layers": 2, "objective": "iwae_elbo", "launches_per_step": (_abi.lib.iwae_launch_count() - l0) / steps,
            "note": "latency-bound: ~1.7 GFLOP per step; fp32 FFMA kernels, the step replayed from a CUDA graph inside iwae_train_step", "last_iwae_elbo": float(res["iwae_elbo"])}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
