// Diagnostics library (libiwae_b200_diag.so, NOT part of the product library): the tensor-core unit test and the
// tcgen05.mma micro-benchmark that sized the fused evaluation kernel.  Declared in include/iwae_b200_diag.h.
#include "tc_ptx.cuh"
#include <stdarg.h>
#include <stdlib.h>

static thread_local char g_diag_err[512] = "";
static void iwae_set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_diag_err, sizeof(g_diag_err), fmt, ap); va_end(ap);
}
#define CUDA_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { iwae_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); return 1; } } while (0)
#define LAUNCH_CHECK() CUDA_TRY(cudaGetLastError())
#define IWAE_REQUIRE(cond, ...) do { if (!(cond)) { iwae_set_error(__VA_ARGS__); return 2; } } while (0)

namespace tc {

// ------------------------------------------------------------------------------------------------
// Self-test GEMM: D[128,N] = A[128,K] B[N,K]^T through exactly the staging the fused kernel uses.
// variant 0: LBO = K-direction stride, SBO = M/N-direction stride (CUTLASS canonical); 1: swapped.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, int N,
                                                            int K, float* __restrict__ D, int variant) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int Kp = (K + 15) & ~15;
  uint8_t* sA = smem;                          // [Kp/8][128][8] halves
  uint8_t* sB = smem + (size_t)Kp * 128 * 2;   // [Kp/8][N][8] halves
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * Kp; i += 128) {
    const int r = i / Kp, k = i % Kp;
    const float v = (k < K) ? A[(size_t)r * K + k] : 0.f;
    *reinterpret_cast<__half*>(sA + (size_t)(k >> 3) * 128 * 16 + r * 16 + (k & 7) * 2) = __float2half_rn(v);
  }
  for (int i = tid; i < N * Kp; i += 128) {
    const int n = i / Kp, k = i % Kp;
    const float v = (k < K) ? B[(size_t)n * K + k] : 0.f;
    *reinterpret_cast<__half*>(sB + (size_t)(k >> 3) * N * 16 + n * 16 + (k & 7) * 2) = __float2half_rn(v);
  }
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncwarp();
  if (warp == 0) { tmem_alloc(&tmem_base_s, ncols); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = make_idesc((uint32_t)N);
    const uint32_t lboA = 128 * 16, lboB = (uint32_t)N * 16, sbo = 128;
    for (int ks = 0; ks < Kp / 16; ++ks) {
      const uint32_t a_addr = smem_u32(sA) + ks * 2 * lboA;
      const uint32_t b_addr = smem_u32(sB) + ks * 2 * lboB;
      const uint64_t ad = variant ? make_smem_desc(a_addr, sbo, lboA) : make_smem_desc(a_addr, lboA, sbo);
      const uint64_t bd = variant ? make_smem_desc(b_addr, sbo, lboB) : make_smem_desc(b_addr, lboB, sbo);
      umma_f16(tmem_base, ad, bd, idesc, ks > 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0, 900);
  tc_fence_after();
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c0 + j < N) D[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

// Micro-benchmark: `reps` back-to-back tcgen05.mma (M=128, K=16, fp16, smem A and B in the fused kernel's
// no-swizzle K-major layout) of width N into one accumulator; returns cycles from first issue to the
// commit's mbarrier completion.  Sizes the chunk width of the fused kernel (SS-mode A-operand fetch).
__global__ void __launch_bounds__(128) umma_bench_kernel(int N, int reps, int mode, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar, bar2, bar3;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + N) * 16 * 2 * 4 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init(&bar3, 1); fence_barrier_init(); }
  __syncwarp();
  if (warp == 0) { tmem_alloc(&tmem_base_s, 256); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  long long t0 = 0;
  if (warp == 1) {
    const uint32_t idesc = make_idesc((uint32_t)N);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem) + 128 * 16 * 2 * 4;
    t0 = clock64();
    if (mode == 0) {
      if (elect_one_sync()) {
        for (int r = 0; r < reps; ++r) {
          const uint32_t ks = r & 3;
          const uint64_t ad = make_smem_desc(a_addr + ks * 2 * 128 * 16, 128 * 16, 128);
          const uint64_t bd = make_smem_desc(b_addr + ks * 2 * N * 16, (uint32_t)N * 16, 128);
          umma_f16(tmem_base, ad, bd, idesc, r > 0);
        }
        umma_commit(&bar);
      }
      __syncwarp();
    } else if (mode & 32) {
      // as the fused kernel's slab loop (elect per slab of two MMAs) but unrolled by four, so that consecutive
      // slabs use different uniform registers for their descriptors
#pragma unroll 4
      for (int r = 0; r < reps; r += 2) {
        if (mode & 2) tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t ks = r & 3;
          const uint64_t ad = make_smem_desc(a_addr + ks * 2 * 128 * 16, 128 * 16, 128);
          const uint64_t bd = make_smem_desc(b_addr + ks * 2 * N * 16, (uint32_t)N * 16, 128);
          umma_f16(tmem_base, ad, bd, idesc, r > 0);
          if (r + 1 < reps) umma_f16(tmem_base, ad + (uint64_t)(2u * 128u), bd + (uint64_t)(2u * (uint32_t)N), idesc, true);
          if (mode & 1) umma_commit(&bar2);
        }
        if (mode & 4) __syncwarp();
      }
      if (elect_one_sync()) umma_commit(&bar);
      __syncwarp();
    } else if (mode & 16) {
      // warp-uniform descriptor arithmetic OUTSIDE the elected region (uniform datapath), only the
      // tcgen05 instructions inside it
      for (int r = 0; r < reps; r += 2) {
        if (mode & 2) tc_fence_after();
        const uint32_t ks = r & 3;
        const uint64_t ad = make_smem_desc(a_addr + ks * 2 * 128 * 16, 128 * 16, 128);
        const uint64_t bd = make_smem_desc(b_addr + ks * 2 * N * 16, (uint32_t)N * 16, 128);
        const uint64_t ad1 = ad + (uint64_t)(2u * 128u), bd1 = bd + (uint64_t)(2u * (uint32_t)N);
        const bool two = r + 1 < reps;
        if (elect_one_sync()) {
          umma_f16(tmem_base, ad, bd, idesc, r > 0);
          if (two) umma_f16(tmem_base, ad1, bd1, idesc, true);
          if (mode & 1) umma_commit(&bar2);
        }
        if (mode & 4) __syncwarp();
      }
      if (elect_one_sync()) umma_commit(&bar);
      __syncwarp();
    } else if (mode & 8) {
      // everything inside ONE elected thread: per slab (bit 0) a commit to a second mbarrier, (bit 1) a
      // try_wait on an already-completed mbarrier
      if (elect_one_sync()) {
        mbar_arrive(&bar3);
        for (int r = 0; r < reps; r += 2) {
          if (mode & 2) { while (!mbar_try_wait(&bar3, 0)) {} tc_fence_after(); }
          for (int rr = r; rr < min(r + 2, reps); ++rr) {
            const uint32_t ks = rr & 3;
            const uint64_t ad = make_smem_desc(a_addr + ks * 2 * 128 * 16, 128 * 16, 128);
            const uint64_t bd = make_smem_desc(b_addr + ks * 2 * N * 16, (uint32_t)N * 16, 128);
            umma_f16(tmem_base, ad, bd, idesc, rr > 0);
          }
          if (mode & 1) umma_commit(&bar2);
        }
        umma_commit(&bar);
      }
      __syncwarp();
    } else {
      // the issue pattern of the fused kernel: per slab of two MMAs a warp-uniform iteration with
      // (mode bit 0) a commit to a second mbarrier, (bit 1) a tcgen05 fence, (bit 2) an elect + __syncwarp
      for (int r = 0; r < reps; r += 2) {
        if (mode & 2) tc_fence_after();
        if (elect_one_sync()) {
          for (int rr = r; rr < min(r + 2, reps); ++rr) {
            const uint32_t ks = rr & 3;
            const uint64_t ad = make_smem_desc(a_addr + ks * 2 * 128 * 16, 128 * 16, 128);
            const uint64_t bd = make_smem_desc(b_addr + ks * 2 * N * 16, (uint32_t)N * 16, 128);
            umma_f16(tmem_base, ad, bd, idesc, rr > 0);
          }
          if (mode & 1) umma_commit(&bar2);
        }
        if (mode & 4) __syncwarp();
      }
      if (elect_one_sync()) umma_commit(&bar);
      __syncwarp();
    }
    mbar_wait(&bar, 0, 901);
    const long long t1 = clock64();
    if (tid == 32) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace tc

int umma_bench(int N, int reps, int nctas, long long* out_cycles_dev, cudaStream_t st) {
  IWAE_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && reps > 0 && nctas > 0, "umma_bench: bad argument");
  const size_t smem = (size_t)(128 + N) * 16 * 2 * 4;
  CUDA_TRY(cudaFuncSetAttribute(tc::umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const char* md = getenv("IWAE_UMMA_BENCH_MODE");      // diagnostics: issue-pattern variants, see the kernel
  tc::umma_bench_kernel<<<nctas, 128, smem, st>>>(N, reps, md ? atoi(md) : 0, out_cycles_dev);
  LAUNCH_CHECK();
  return 0;
}

int umma_selftest(const float* A, const float* B, int N, int K, int variant, float* D, cudaStream_t st) {
  IWAE_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0, "umma_selftest: N must be a multiple of 16 in [16,256]");
  IWAE_REQUIRE(K >= 1 && K <= 256, "umma_selftest: K must be in [1,256]");
  const int Kp = (K + 15) & ~15;
  const size_t smem = (size_t)Kp * (128 + N) * 2;
  CUDA_TRY(cudaFuncSetAttribute(tc::umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc::umma_selftest_kernel<<<1, 128, smem, st>>>(A, B, N, K, D, variant);
  LAUNCH_CHECK();
  return 0;
}

extern "C" const char* iwae_diag_last_error(void) { return g_diag_err; }

extern "C" int iwae_umma_selftest(const float* A, const float* B, int32_t N, int32_t K, int32_t variant, float* D,
                                  void* stream) {
  IWAE_REQUIRE(A && B && D, "iwae_umma_selftest: null argument");
  return umma_selftest(A, B, N, K, variant, D, (cudaStream_t)stream);
}

extern "C" int iwae_umma_bench(int32_t N, int32_t reps, int32_t nctas, int64_t* out_cycles, void* stream) {
  IWAE_REQUIRE(out_cycles, "iwae_umma_bench: null argument");
  return umma_bench(N, reps, nctas, reinterpret_cast<long long*>(out_cycles), (cudaStream_t)stream);
}
