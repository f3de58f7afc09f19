"""tcgen05.mma issue-pattern variants (IWAE_UMMA_BENCH_MODE bits: 1 commit per 2 MMAs, 2 tcgen05 fence per
2 MMAs, 4 elect + __syncwarp per 2 MMAs) - which part of the fused kernel's slab loop costs the time."""
import ctypes as C, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from iwae_b200._abi import diag_lib, check_diag as check
    lib = diag_lib()
    out = torch.zeros(148, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for reps in (4, 8, 14, 26, 52):
        for _ in range(2):
            check(lib.iwae_umma_bench(112, reps, 148, C.c_void_p(out.data_ptr()), st)); torch.cuda.synchronize()
        t = float(out.double().mean())
        print(f"mode={os.environ.get('IWAE_UMMA_BENCH_MODE','0')} N=112 reps={reps:3d}: {t:8.0f} cycles total, {t/reps:7.1f} per MMA")
else:
    for mode in (0, 4, 16, 21, 32, 37):
        subprocess.run([sys.executable, __file__, "run"], env={**os.environ, "IWAE_UMMA_BENCH_MODE": str(mode)})
