"""Absolute latency of short tcgen05.mma bursts (issue of first MMA -> commit observed)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iwae_b200._abi import diag_lib, check_diag as check
lib = diag_lib()
out = torch.zeros(148, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for N in (112, 256):
    for reps in (1, 2, 4, 7, 14, 26, 52, 128):
        for _ in range(2):
            check(lib.iwae_umma_bench(N, reps, 148, C.c_void_p(out.data_ptr()), st)); torch.cuda.synchronize()
        t = float(out.double().mean())
        print(f"N={N:3d} reps={reps:3d}: {t:8.0f} cycles total, {t/reps:7.1f} per MMA")
