/* iwae_b200_diag.h - diagnostics library libiwae_b200_diag.so (not part of the product library libiwae_b200.so):
 * the tensor-core unit test hook and the tcgen05.mma micro-benchmark.  extern "C", device pointers, explicit stream;
 * 0 = ok, else iwae_diag_last_error(). */
#ifndef IWAE_B200_DIAG_H
#define IWAE_B200_DIAG_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* iwae_diag_last_error(void);

/* Tensor-core unit test hook: D[128,N] = A[128,K] * B[N,K]^T with fp16 operands staged exactly like
 * the fused evaluation kernel stages them (no-swizzle K-major core matrices, one tcgen05.mma chain).
 * A, B fp32 on the device (rounded to fp16 inside).  variant 0 = the descriptor convention the fused
 * kernel uses (LBO = K-direction core-matrix stride, SBO = M/N-direction); 1 = the two swapped
 * (diagnostic only).  Used by tests/test_gpu_eval.py::test_umma_descriptor_selftest. */
int iwae_umma_selftest(const float* A, const float* B, int32_t N, int32_t K, int32_t variant, float* D,
                       void* stream);

/* Tensor-core micro-benchmark (diagnostic): `reps` back-to-back tcgen05.mma of shape M=128, N, K=16 in the
 * fused kernel's operand layout on `nctas` CTAs; out_cycles [nctas] (device) receives the SM cycles from
 * first issue to completion.  Used to size the fused kernel's MMA width (scripts/umma_bench.py). */
int iwae_umma_bench(int32_t N, int32_t reps, int32_t nctas, int64_t* out_cycles, void* stream);

/* TMEM-read micro-benchmark (diagnostic): cycles of `reps` 16-column accumulator groups per warp for the epilogue's
 * load patterns (mode 0/1: tcgen05.ld .x16 serial / double-buffered, 2/3: .x32, 4/6: with the tanh MUFU work, 5: MUFU
 * only), `nwarps` warps per CTA (nwarps/4 per TMEM lane quarter).  out_cycles [nctas*nwarps].  scripts/ldtm_bench.py. */
int iwae_ldtm_bench(int32_t mode, int32_t nwarps, int32_t reps, int32_t nctas, int64_t* out_cycles, float* sink, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IWAE_B200_DIAG_H */
