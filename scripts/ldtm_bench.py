"""TMEM-read (tcgen05.ld) latency / throughput for the epilogue's load patterns: cycles per 16-column group and warp."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iwae_b200._abi import diag_lib, check_diag as check
lib = diag_lib()
st = torch.cuda.current_stream().cuda_stream
sink = torch.zeros(4, device="cuda")
names = {0: "x16 serial", 1: "x16 double-buffered", 2: "x32 serial", 3: "x32 double-buffered", 4: "x16 dbuf + 16 tanh", 5: "16 tanh only",
         6: "x32 dbuf + 32 tanh", 7: "x16 dbuf + Bernoulli math", 8: "x16 dbuf + tanh+pack+STS", 9: "x16 dbuf + pack+STS"}
reps = 4096
for nctas in (1, 296):
    for nwarps in (4, 8):
        for mode in (int(m) for m in os.environ.get('MODES', '0,1,2,3,4,5,6,7,8,9').split(',')):
            out = torch.zeros(nctas * nwarps, dtype=torch.int64, device="cuda")
            for _ in range(2):
                check(lib.iwae_ldtm_bench(mode, nwarps, reps, nctas, C.c_void_p(out.data_ptr()), C.c_void_p(sink.data_ptr()), st))
                torch.cuda.synchronize()
            t = out.double() / reps
            print(f"ctas={nctas:3d} warps/CTA={nwarps:2d} {names[mode]:22s}: {float(t.mean()):7.1f} cycles per 16-col group (max {float(t.max()):7.1f})")
