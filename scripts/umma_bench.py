"""tcgen05.mma cost versus N (M=128, K=16, SS mode, no-swizzle K-major operands): cycles per MMA."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iwae_b200._abi import diag_lib, check_diag as check
lib = diag_lib()
out = torch.zeros(148, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for nctas in (1, 148):
    for N in (16, 64, 96, 112, 128, 160, 208, 224, 256):
        res = []
        for reps in (64, 256):
            check(lib.iwae_umma_bench(N, reps, nctas, C.c_void_p(out.data_ptr()), st)); torch.cuda.synchronize()
            check(lib.iwae_umma_bench(N, reps, nctas, C.c_void_p(out.data_ptr()), st)); torch.cuda.synchronize()
            res.append(float(out[:nctas].double().mean()))
        per = (res[1] - res[0]) / (256 - 64)
        print(f"ctas={nctas:3d} N={N:3d}: {per:7.1f} cycles/MMA  (ideal N/2 = {N/2:.0f};  {128*N*16/per:7.0f} MAC/clk/SM)")
