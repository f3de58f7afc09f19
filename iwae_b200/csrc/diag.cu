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


// Micro-benchmark of the epilogue's TMEM reads: every warp of the CTA (wpq warps per TMEM lane quarter) runs `reps`
// iterations of one pattern and reports its cycles per 16-column group.
//   mode 0: tcgen05.ld .x16, wait, consume (serial latency)          mode 1: .x16 double-buffered like the fused kernel
//   mode 2: .x32 serial                                              mode 3: .x32 double-buffered
//   mode 4: mode 1 + 16 MUFU.TANH per group (the tanh epilogue without its stores)
//   mode 5: 16 MUFU.TANH per group, no TMEM read                     mode 6: mode 3 + 32 MUFU.TANH per .x32 group
__device__ __forceinline__ void tmem_ld32_d(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
      "%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ float tanh_fast(float v) { float r; asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
template <int NR>
__device__ __forceinline__ float consume(const uint32_t (&r)[NR], bool mufu) {
  float a = 0.f;
#pragma unroll
  for (int i = 0; i < NR; ++i) a += mufu ? tanh_fast(__uint_as_float(r[i])) : __uint_as_float(r[i]);
  return a;
}
__device__ __forceinline__ float ex2_fast(float v) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__global__ void __launch_bounds__(512) ldtm_bench_kernel(int mode, int reps, long long* out, float* sink) {
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float sgn_s[144];
  __shared__ __align__(16) uint8_t act_s[14 * 128 * 16];
  if (threadIdx.x < 144) sgn_s[threadIdx.x] = (threadIdx.x & 1) ? 1.f : -1.f;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (warp == 0) { tmem_alloc(&tmem_base_s, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (mode == 0) {
    for (int r = 0; r < reps; ++r) { uint32_t a[16]; tmem_ld16(taddr + (r & 7) * 16, a); tmem_ld_wait(); acc += consume<16>(a, false); }
  } else if (mode == 1 || mode == 4) {
    uint32_t a[16], b[16];
    tmem_ld16(taddr, a); tmem_ld_wait();
    for (int r = 0; r < reps; r += 2) {
      tmem_ld16(taddr + 16, b); acc += consume<16>(a, mode == 4); tmem_ld_wait();
      tmem_ld16(taddr + 32, a); acc += consume<16>(b, mode == 4); tmem_ld_wait();
    }
  } else if (mode == 2) {
    for (int r = 0; r < reps; r += 2) { uint32_t a[32]; tmem_ld32_d(taddr + (r & 6) * 16, a); tmem_ld_wait(); acc += consume<32>(a, false); }
  } else if (mode == 3 || mode == 6) {
    uint32_t a[32], b[32];
    tmem_ld32_d(taddr, a); tmem_ld_wait();
    for (int r = 0; r < reps; r += 4) {
      tmem_ld32_d(taddr + 32, b); acc += consume<32>(a, mode == 6); tmem_ld_wait();
      tmem_ld32_d(taddr + 64, a); acc += consume<32>(b, mode == 6); tmem_ld_wait();
    }
  } else if (mode == 7) {
    // the Bernoulli epilogue's arithmetic per 16-column group (eval_tc.cu, EPI_BERN fast path)
    uint32_t a[16], b[16];
    float p0 = 1.f, p1 = 1.f, p2 = 1.f, p3 = 1.f, s0 = 0.f, s1 = 0.f, a0 = 0.f, a1 = 0.f;
    auto bern = [&](const uint32_t (&r)[16], int g) {
      const float4* sg4 = reinterpret_cast<const float4*>(&sgn_s[g & 127]);
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        const float4 sg = sg4[i4];
        const float l0 = __uint_as_float(r[4 * i4]), l1 = __uint_as_float(r[4 * i4 + 1]);
        const float l2 = __uint_as_float(r[4 * i4 + 2]), l3 = __uint_as_float(r[4 * i4 + 3]);
        s0 = fmaf(sg.x, l0, s0); s1 = fmaf(sg.y, l1, s1); s0 = fmaf(sg.z, l2, s0); s1 = fmaf(sg.w, l3, s1);
        a0 += fabsf(l0); a1 += fabsf(l1); a0 += fabsf(l2); a1 += fabsf(l3);
        p0 = fmaf(p0, ex2_fast(-fabsf(l0)), p0); p1 = fmaf(p1, ex2_fast(-fabsf(l1)), p1);
        p2 = fmaf(p2, ex2_fast(-fabsf(l2)), p2); p3 = fmaf(p3, ex2_fast(-fabsf(l3)), p3);
      }
    };
    tmem_ld16(taddr, a); tmem_ld_wait();
    for (int r = 0; r < reps; r += 2) {
      tmem_ld16(taddr + 16, b); bern(a, (r * 16) & 96); tmem_ld_wait();
      tmem_ld16(taddr + 32, a); bern(b, (r * 16 + 16) & 96); tmem_ld_wait();
    }
    acc = p0 * p1 + p2 * p3 + s0 + s1 + a0 + a1;
  } else if (mode == 8 || mode == 9) {
    // the tanh epilogue per 16-column group: (mode 8: 16 MUFU.TANH,) 8 cvt.rn.satfinite.f16x2.f32, two 16-byte stores
    uint32_t a[16], b[16];
    const int row = (warp & 3) * 32 + (tid & 31);
    auto th = [&](const uint32_t (&r)[16], int g) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = (mode == 8) ? tanh_fast(__uint_as_float(r[i])) : __uint_as_float(r[i]);
      uint4 x = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
      uint4 y = make_uint4(pack_h2(v[8], v[9]), pack_h2(v[10], v[11]), pack_h2(v[12], v[13]), pack_h2(v[14], v[15]));
      *reinterpret_cast<uint4*>(act_s + (size_t)(g >> 3) * (128 * 16) + row * 16) = x;
      *reinterpret_cast<uint4*>(act_s + (size_t)((g >> 3) + 1) * (128 * 16) + row * 16) = y;
    };
    tmem_ld16(taddr, a); tmem_ld_wait();
    for (int r = 0; r < reps; r += 2) {
      tmem_ld16(taddr + 16, b); th(a, (r * 16) & 96); tmem_ld_wait();
      tmem_ld16(taddr + 32, a); th(b, (r * 16 + 16) & 96); tmem_ld_wait();
    }
    __syncwarp();
    acc = __uint_as_float(a[0]) + (float)act_s[(tid * 16) % (14 * 128 * 16)];
  } else {
    float v = (float)tid;
    for (int r = 0; r < reps; ++r) {
      uint32_t a[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = __float_as_uint(v + (float)i);
      const float t = consume<16>(a, true);
      acc += t; v = t * 1e-3f;
    }
  }
  const long long t1 = clock64();
  if ((tid & 31) == 0) out[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 256);
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

/* TMEM-read micro-benchmark (diagnostic): `nwarps` warps (a multiple of 4) per CTA, `nctas` CTAs; out_cycles [nctas*nwarps] */
extern "C" int iwae_ldtm_bench(int32_t mode, int32_t nwarps, int32_t reps, int32_t nctas, int64_t* out_cycles, float* sink, void* stream) {
  IWAE_REQUIRE(out_cycles && sink && nwarps >= 4 && nwarps <= 16 && nwarps % 4 == 0 && reps % 4 == 0, "iwae_ldtm_bench: bad argument");
  tc::ldtm_bench_kernel<<<nctas, nwarps * 32, 0, (cudaStream_t)stream>>>(mode, reps, reinterpret_cast<long long*>(out_cycles), sink);
  LAUNCH_CHECK();
  return 0;
}
