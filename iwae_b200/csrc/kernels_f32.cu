// fp32 (FFMA) kernels: the parity-mode forward, the backward pass, Adam, SIR and the data step.
// Row index convention everywhere: r = s*B + b (sample-major, the reference's [k,B,.] layout).
#include "common.cuh"
#include "rng.cuh"
#include <math.h>
#include <algorithm>
#include <utility>
#include <vector>
#include <string.h>
#include <stdio.h>

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------
// The train step is a chain of ~35 dependent kernels of a few microseconds each: the gap between two of them (drain,
// launch, ramp) is as long as the kernels.  Every kernel of this file starts with griddepcontrol.wait and is launched
// with the programmatic-stream-serialization attribute, so its CTAs are scheduled while its predecessor in the stream is
// still finishing and block at that wait until the predecessor has completed and flushed (a no-op without the
// attribute).  Under stream capture the dependency becomes a programmatic edge of the CUDA graph.  IWAE_NO_PDL=1: off.
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// Every train-step kernel is a `*_body(vb_, vg_, args...)` device function plus a thin __global__ wrapper: the body sees
// its block index / grid size through these two shadowing locals, so the same code runs as a kernel of its own
// (vb_ = blockIdx) or as one virtual block of the train-step megakernel (end of this file).
#define VBLOCK_SHADOW const uint3 blockIdx = vb_; const dim3 gridDim = vg_; (void)blockIdx; (void)gridDim
static bool pdl_enabled() { static const bool on = getenv("IWAE_NO_PDL") == nullptr; return on; }
// ---- argument packs of recorded launches (train-step megakernel, end of this file) ----
constexpr int MEGA_ARG_BYTES = 512;
constexpr size_t MEGA_SMEM_BYTES = 36 * 1024;       // dynamic shared-memory pool: the largest GEMM ring (64x64 tiles: 34,816 bytes)
template <class... A> struct ArgPack;
template <> struct ArgPack<> {};
template <class H, class... T> struct ArgPack<H, T...> { H head; ArgPack<T...> tail; };
template <class... A> struct PackMaker;
template <> struct PackMaker<> { static ArgPack<> make() { return {}; } };
template <class H, class... T> struct PackMaker<H, T...> {
  static ArgPack<H, T...> make(H h, T... t) { return ArgPack<H, T...>{h, PackMaker<T...>::make(t...)}; }
};
template <size_t I, class H, class... T>
__host__ __device__ __forceinline__ const auto& pack_get(const ArgPack<H, T...>& p) {
  if constexpr (I == 0) return p.head; else return pack_get<I - 1>(p.tail);
}
bool mega_record(const void* kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const void* pack, size_t pack_bytes);

template <class... KArgs, class... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  if (mega_recording()) {        // train-step megakernel: append to the program instead of launching
    const ArgPack<KArgs...> pack = PackMaker<KArgs...>::make(static_cast<KArgs>(args)...);
    mega_record((const void*)kern, grid, block, smem, st, &pack, sizeof(pack));
    return cudaSuccess;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// =============================================================================================
// Tiled SGEMM with fused epilogues.  C(m,n) = epi( sum_k A(m,k) B(k,n) ).
// Generic strides cover x@W (NN), dY@W^T (NT) and X^T@dY (TN); the tile loaders pick the thread
// mapping that walks the unit-stride dimension so global loads stay coalesced in all three.
// =============================================================================================
template <int BM, int BN, int TM, int TN>
__device__ __forceinline__ void gemm_f32_body(const uint3 vb_, const dim3 vg_, GemmArgs a) {
  VBLOCK_SHADOW;
  constexpr int BK = 16;
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int TX = BN / TN;   // threads along n
  constexpr int EA = BM * BK / NT, EB = BN * BK / NT;   // elements of the A / B tile each thread stages
  static_assert(EA * NT == BM * BK && EB * NT == BN * BK, "tile must divide evenly among the threads");
  // NSTG-stage cp.async ring (4-byte copies with zero fill, so any stride / alignment / ragged edge works): with
  // rows = k*B = 1,000 these GEMMs are a chain of 7-50 K-tiles of a few hundred cycles of FFMAs each, and a
  // register-staged double buffer (prefetch distance 1) left every tile exposed to the full global-load latency -
  // ~11 us per launch, 43 launches per train step.  Three tiles in flight hide it.
  constexpr int NSTG = 4;
  // the ring lives in DYNAMIC shared memory (gemm_smem_bytes): inside the train-step megakernel every operation shares one pool
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  typedef float AsT[BK][BM + 4];
  typedef float BsT[BK][BN + 4];
  AsT* As = reinterpret_cast<AsT*>(dyn_smem);
  BsT* Bs = reinterpret_cast<BsT*>(dyn_smem + sizeof(AsT) * NSTG);

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;   // M tiles on x (no 65535 limit)
  const bool a_kcontig = (a.sa_k == 1);
  const bool b_ncontig = (a.sb_n == 1);
  // split-K: blockIdx.z owns K range [k_begin, k_end) and writes its partial tile to C + z * split_stride
  const int k_chunk = (a.split_k > 1) ? ((a.K + a.split_k - 1) / a.split_k + BK - 1) / BK * BK : a.K;
  const int k_begin = blockIdx.z * k_chunk;
  const int k_end = min(a.K, k_begin + k_chunk);
  float* __restrict__ Cout = a.C ? a.C + (size_t)blockIdx.z * a.split_stride : nullptr;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  auto cp4 = [](float* dst_smem, const float* src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    const int nbytes = valid ? 4 : 0;                       // src-size 0: the 4 bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(nbytes) : "memory");
  };
  // per-thread element slots of the A / B tiles: base pointer at k = k_begin, validity of the row / column and the
  // k offset inside a tile are fixed for the whole K loop, so a tile issue is one add + compare + cp.async per element
  const float* pa[EA]; float* sa[EA]; int ka[EA]; bool oka[EA]; bool onea[EA];
  const float* pb[EB]; float* sb[EB]; int kb[EB]; bool okb[EB];
#pragma unroll
  for (int i = 0; i < EA; ++i) {
    const int e = tid + i * NT;
    int mm, kk;
    if (a_kcontig) { mm = e / BK; kk = e % BK; } else { kk = e / BM; mm = e % BM; }
    const int gm = m0 + mm;
    oka[i] = gm < a.M; onea[i] = a.ones_row && gm == a.M; ka[i] = kk;
    pa[i] = a.A + (oka[i] ? (int64_t)gm * a.sa_m + (int64_t)(k_begin + kk) * a.sa_k : 0);
    sa[i] = &As[0][kk][mm];
  }
#pragma unroll
  for (int i = 0; i < EB; ++i) {
    const int e = tid + i * NT;
    int nn, kk;
    if (b_ncontig) { kk = e / BN; nn = e % BN; } else { nn = e / BK; kk = e % BK; }
    const int gn = n0 + nn;
    okb[i] = gn < a.N; kb[i] = kk;
    pb[i] = a.B + (okb[i] ? (int64_t)(k_begin + kk) * a.sb_k + (int64_t)gn * a.sb_n : 0);
    sb[i] = &Bs[0][kk][nn];
  }
  constexpr int A_STAGE = BK * (BM + 4), B_STAGE = BK * (BN + 4);      // floats per ring stage
  auto issue_tile = [&](int t, int buf) {                               // tile t covers k_begin + t*BK ...
    const int krem = k_end - (k_begin + t * BK);                       // valid k offsets of this tile: kk < krem
#pragma unroll
    for (int i = 0; i < EA; ++i) {
      const bool ok = oka[i] && ka[i] < krem;
      if (onea[i]) sa[i][buf * A_STAGE] = (ka[i] < krem) ? 1.f : 0.f;  // virtual row of ones (bias gradient)
      else cp4(sa[i] + buf * A_STAGE, ok ? pa[i] + (int64_t)t * BK * a.sa_k : a.A, ok);
    }
#pragma unroll
    for (int i = 0; i < EB; ++i) {
      const bool ok = okb[i] && kb[i] < krem;
      cp4(sb[i] + buf * B_STAGE, ok ? pb[i] + (int64_t)t * BK * a.sb_k : a.B, ok);
    }
  };

  if (k_begin < k_end) {
    const int ntiles = (k_end - k_begin + BK - 1) / BK;
    // prologue: tiles 0 .. NSTG-2 (one commit group per slot, empty groups keep the count uniform)
#pragma unroll
    for (int t = 0; t < NSTG - 1; ++t) {
      if (t < ntiles) issue_tile(t, t);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int t = 0; t < ntiles; ++t) {
      asm volatile("cp.async.wait_group %0;" ::"n"(NSTG - 2) : "memory");      // tile t has landed (this thread's part)
      __syncthreads();                                                          // ... everyone's part; tile t-1 is consumed
      if (t + NSTG - 1 < ntiles) issue_tile(t + NSTG - 1, (t + NSTG - 1) % NSTG);
      asm volatile("cp.async.commit_group;" ::: "memory");
      const int buf = t % NSTG;
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float av[TM], bv[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) av[i] = As[buf][kk][ty * TM + i];
#pragma unroll
        for (int j = 0; j < TN; ++j) bv[j] = Bs[buf][kk][tx * TN + j];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
  }

  // ---- epilogue
  const int m_rows = a.M + (a.ones_row ? 1 : 0);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + ty * TM + i;
    float bern = 0.f;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tx * TN + j;
      if (gm < m_rows && gn < a.N) {
        float v = acc[i][j];
        if (a.bias) v += __ldg(a.bias + gn);
        if (a.act == ACT_TANH) v = tanhf(v);
        else if (a.act == ACT_EXP_EPS) v = expf(v) + 1e-6f;
        if (a.accumulate) v += Cout[(int64_t)gm * a.ldc + gn];
        if (a.mul_h) {
          const float h = __ldg(a.mul_h + (int64_t)gm * a.ld_h + gn);
          v *= (1.f - h * h);
        }
        if (a.dl_x) {
          // recomputed logit -> its gradient (the arithmetic of dlogits_body on the stored logit, bit for bit)
          const float sg = 1.f / (1.f + expf(-v));
          v = __ldg(a.dl_g + gm) * (__ldg(a.dl_x + (int64_t)(gm % a.dl_B) * a.N + gn) - sg);
        }
        if (Cout) Cout[(int64_t)gm * a.ldc + gn] = v;
        if (a.bern_x) {
          // tfd.Bernoulli(logits).log_prob(x) = -(1-x) softplus(l) - x softplus(-l)   (A.3)
          const float xv = __ldg(a.bern_x + (int64_t)(gm % a.bern_B) * a.N + gn);
          const float sp = fmaxf(v, 0.f) + log1pf(expf(-fabsf(v)));
          bern += -(1.f - xv) * sp - xv * (sp - v);
        }
      }
    }
    if (a.bern_x) {
      // reduce over the TX threads that share this row (TX is a power of two <= 32, lane-aligned)
#pragma unroll
      for (int o = TX / 2; o > 0; o >>= 1) bern += __shfl_xor_sync(0xffffffffu, bern, o);
      if (tx == 0 && gm < a.M) a.bern_partial[(int64_t)gm * gridDim.y + blockIdx.y] = bern;
    }
  }
}

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) gemm_f32_kernel(GemmArgs a) {
  pdl_wait();
  gemm_f32_body<BM, BN, TM, TN>(blockIdx, gridDim, a);
}
template <int BM, int BN> constexpr size_t gemm_smem_bytes() { return sizeof(float) * 4 * 16 * ((BM + 4) + (BN + 4)); }

int gemm_f32_ntiles_n(int N, int small_tile) { return small_tile ? (N + 31) / 32 : (N + 63) / 64; }

int gemm_f32(const GemmArgs& a, cudaStream_t st) {
  if (a.M <= 0 || a.N <= 0) return 0;
  const int z = a.split_k > 1 ? a.split_k : 1;
  IWAE_REQUIRE(z == 1 || (!a.bias && a.act == ACT_NONE && !a.accumulate && !a.mul_h && !a.bern_x),
               "gemm_f32: split-K supports plain products only");
  IWAE_REQUIRE(!a.ones_row || (!a.bias && a.act == ACT_NONE && !a.accumulate && !a.mul_h && !a.bern_x),
               "gemm_f32: ones_row supports plain products only");
  IWAE_REQUIRE(!a.dl_x || (z == 1 && !a.ones_row && !a.bern_x && !a.mul_h && !a.accumulate && a.dl_g && a.dl_B > 0 && a.C),
               "gemm_f32: the logit-gradient epilogue needs a plain product with g, B and an output");
  const int m_rows = a.M + (a.ones_row ? 1 : 0);
  if (a.small_tile && m_rows >= 256 && a.K >= 512 && z == 1) {
    // tall, narrow outputs (rows = k*B = 1,000 by <= 200 units): 4x2 register tiles halve the shared-memory loads
    // per FFMA of the 2x2 kernel, which is LDS-bound on the long-K products (dX of the 784-logit layer: 48 -> ~20 us)
    dim3 grid((m_rows + 63) / 64, (a.N + 31) / 32, z);
    CUDA_TRY(launch_pdl(gemm_f32_kernel<64, 32, 4, 2>, dim3(grid), dim3(256), gemm_smem_bytes<64, 32>(), st, a));
  } else if (a.small_tile) {
    dim3 grid((m_rows + 31) / 32, (a.N + 31) / 32, z);
    CUDA_TRY(launch_pdl(gemm_f32_kernel<32, 32, 2, 2>, dim3(grid), dim3(256), gemm_smem_bytes<32, 32>(), st, a));
  } else {
    dim3 grid((m_rows + 63) / 64, (a.N + 63) / 64, z);
    CUDA_TRY(launch_pdl(gemm_f32_kernel<64, 64, 4, 4>, dim3(grid), dim3(256), gemm_smem_bytes<64, 64>(), st, a));
  }
  LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// Small elementwise / reduction kernels
// =============================================================================================
__global__ void colsum_kernel(const float* __restrict__ A, int64_t lda, int M, int N, float* __restrict__ out,
                              int64_t split_stride) {
  pdl_wait();
  // one strip of 32 columns per blockIdx.x, one row range per blockIdx.y (partial sum -> out + y*split_stride);
  // blockDim = (32, 8): 8 row-lanes then an smem reduce in a fixed order
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float s = 0.f;
  if (n < N)
    for (int m = r0 + threadIdx.y; m < r1; m += 8) s += A[(int64_t)m * lda + n];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    out[(int64_t)blockIdx.y * split_stride + n] = t;
  }
}
int colsum_f32(const float* A, int64_t lda, int M, int N, float* out, cudaStream_t st, int splits, int64_t split_stride) {
  if (splits < 1 || M < 64 * splits) splits = 1;
  CUDA_TRY(launch_pdl(colsum_kernel, dim3(dim3((N + 31) / 32, splits)), dim3(dim3(32, 8)), 0, st, A, lda, M, N, out, split_stride));
  LAUNCH_CHECK();
  return 0;
}

__device__ __forceinline__ int label_of(const float* y, int b) {
  // tf.cast(y, tf.uint8) then one_hot(depth) (task04.py:112): out-of-range labels give a zero row
  return (int)(unsigned char)(int)y[b];
}

__device__ __forceinline__ void concat_onehot_body(const uint3 vb_, const dim3 vg_, const float* __restrict__ x, int64_t ldx, int n, const float* __restrict__ y,
                                     int B, int rows, int nc, float* __restrict__ out) {
  VBLOCK_SHADOW;
  const int64_t total = (int64_t)rows * (n + nc);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (n + nc);
    const int c = (int)(i % (n + nc));
    float v;
    if (c < n) v = x[r * ldx + c];
    else v = (label_of(y, (int)(r % B)) == c - n) ? 1.f : 0.f;
    out[i] = v;
  }
}
__global__ void __launch_bounds__(256) concat_onehot_kernel(const float* __restrict__ x, int64_t ldx, int n, const float* __restrict__ y, int B, int rows, int nc, float* __restrict__ out) {
  pdl_wait();
  concat_onehot_body(blockIdx, gridDim, x, ldx, n, y, B, rows, nc, out);
}
int concat_onehot(const float* x, int64_t ldx, int n, const float* y, int B, int rows, int nc, float* out,
                  cudaStream_t st) {
  const int64_t total = (int64_t)rows * (n + nc);
  const int blocks = (int)((total + 255) / 256 > 65535 ? 65535 : (total + 255) / 256);
  CUDA_TRY(launch_pdl(concat_onehot_kernel, dim3(blocks), dim3(256), 0, st, x, ldx, n, y, B, rows, nc, out));
  LAUNCH_CHECK();
  return 0;
}
int onehot_f32(const float* y, int B, int nc, float* out, cudaStream_t st) {
  return concat_onehot(nullptr, 0, 0, y, B, B, nc, out, st);
}

// ---- reparameterised sampling: z = eps*std + mu (tfd.Normal.sample, A.2) ----
__device__ __forceinline__ void sample_latent_body(const uint3 vb_, const dim3 vg_, const float* __restrict__ mu, const float* __restrict__ sd, int bcast,
                                     const float* __restrict__ eps_in, float* __restrict__ eps_out,
                                     float* __restrict__ z, int B, int k, int d, uint64_t seed, int stream_id,
                                     int64_t img_offset, const StepDyn* __restrict__ dyn) {
  VBLOCK_SHADOW;
  if (dyn) { seed = dyn->seed; img_offset = dyn->img_offset; }      // graph replay: per-step values live on the device
  const int nq = (d + 7) / 8;                // one Philox call per dim-octet (rng.cuh)
  const int64_t total = (int64_t)k * B * nq;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % nq);
    const int64_t r = i / nq;
    const int b = (int)(r % B);
    const int s = (int)(r / B);
    float e[8];
    if (eps_in) {
#pragma unroll
      for (int j = 0; j < 8; ++j) e[j] = (8 * q + j < d) ? eps_in[r * d + 8 * q + j] : 0.f;
    } else {
      philox_normal8(seed, (uint32_t)stream_id, (uint32_t)(img_offset + b), (uint32_t)s, (uint32_t)q, e);
    }
    const int64_t prow = bcast ? b : r;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = 8 * q + j;
      if (c < d) {
        const float m = mu ? mu[prow * d + c] : 0.f;
        const float sdev = sd ? sd[prow * d + c] : 1.f;
        if (z) z[r * d + c] = e[j] * sdev + m;
        if (eps_out) eps_out[r * d + c] = e[j];
      }
    }
  }
}
__global__ void __launch_bounds__(256) sample_latent_kernel(const float* __restrict__ mu, const float* __restrict__ sd, int bcast, const float* __restrict__ eps_in, float* __restrict__ eps_out, float* __restrict__ z, int B, int k, int d, uint64_t seed, int stream_id, int64_t img_offset, const StepDyn* __restrict__ dyn) {
  pdl_wait();
  sample_latent_body(blockIdx, gridDim, mu, sd, bcast, eps_in, eps_out, z, B, k, d, seed, stream_id, img_offset, dyn);
}
int sample_latent(const float* mu, const float* sd, int bcast, const float* eps_in, float* eps_out, float* z,
                  int B, int k, int d, uint64_t seed, int stream_id, int64_t img_offset, cudaStream_t st) {
  const int64_t total = (int64_t)k * B * ((d + 7) / 8);
  if (total == 0) return 0;
  const int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  CUDA_TRY(launch_pdl(sample_latent_kernel, dim3(blocks), dim3(256), 0, st, mu, sd, bcast, eps_in, eps_out, z, B, k, d, seed, stream_id, img_offset,
                                               eps_in ? nullptr : g_step_dyn));
  LAUNCH_CHECK();
  return 0;
}
int philox_fill(uint64_t seed, int stream_id, int64_t img_offset, int B, int k, int d, float* out, cudaStream_t st) {
  return sample_latent(nullptr, nullptr, 1, nullptr, out, nullptr, B, k, d, seed, stream_id, img_offset, st);
}

// ---- Gaussian log-density row sums: one warp per row ----
#define LOG2PI_HALF 0.91893853320467274178f
__device__ __forceinline__ void normal_logp_rows_body(const uint3 vb_, const dim3 vg_, const float* __restrict__ z, const float* __restrict__ mu,
                                        const float* __restrict__ sd, int bcast, float std_extra, int B, int64_t R, int d,
                                        float* __restrict__ out) {
  VBLOCK_SHADOW;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < R; r += nwarps) {
    const int64_t prow = bcast ? (r % B) : r;
    float s = 0.f;
    for (int j = lane; j < d; j += 32) {
      const float zz = z[r * d + j];
      if (mu) {
        const float sc = sd[prow * d + j] + std_extra;
        const float t = zz / sc - mu[prow * d + j] / sc;            // TFP: both terms divided first
        s += -0.5f * t * t - (LOG2PI_HALF + logf(sc));
      } else {
        s += -0.5f * zz * zz - LOG2PI_HALF;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[r] = s;
  }
}
__global__ void __launch_bounds__(256) normal_logp_rows_kernel(const float* __restrict__ z, const float* __restrict__ mu, const float* __restrict__ sd, int bcast, float std_extra, int B, int64_t R, int d, float* __restrict__ out) {
  pdl_wait();
  normal_logp_rows_body(blockIdx, gridDim, z, mu, sd, bcast, std_extra, B, R, d, out);
}
int normal_logp_rows(const float* z, const float* mu, const float* sd, int bcast, float std_extra, int B, int k,
                     int d, float* out, cudaStream_t st) {
  const int64_t R = (int64_t)B * k;
  if (R == 0) return 0;
  int64_t blocks = (R + 7) / 8;
  if (blocks > 148 * 32) blocks = 148 * 32;
  CUDA_TRY(launch_pdl(normal_logp_rows_kernel, dim3((int)blocks), dim3(256), 0, st, z, mu, sd, bcast, std_extra, B, R, d, out));
  LAUNCH_CHECK();
  return 0;
}

__device__ __forceinline__ void normal_logp_rows_multi_body(const uint3 vb_, const dim3 vg_, NlpJobs jobs, int B, int64_t R) {
  VBLOCK_SHADOW;
  const NlpJob& jb = jobs.job[blockIdx.y];
  const float* __restrict__ z = jb.z; const float* __restrict__ mu = jb.mu; const float* __restrict__ sd = jb.sd;
  const int d = jb.d, bcast = jb.bcast; const float std_extra = jb.std_extra;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < R; r += nwarps) {      // the same per-row arithmetic as normal_logp_rows_kernel
    const int64_t prow = bcast ? (r % B) : r;
    float s = 0.f;
    for (int j = lane; j < d; j += 32) {
      const float zz = z[r * d + j];
      if (mu) {
        const float sc = sd[prow * d + j] + std_extra;
        const float t = zz / sc - mu[prow * d + j] / sc;
        s += -0.5f * t * t - (LOG2PI_HALF + logf(sc));
      } else {
        s += -0.5f * zz * zz - LOG2PI_HALF;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) jb.out[r] = s;
  }
}
__global__ void __launch_bounds__(256) normal_logp_rows_multi_kernel(NlpJobs jobs, int B, int64_t R) {
  pdl_wait();
  normal_logp_rows_multi_body(blockIdx, gridDim, jobs, B, R);
}
int normal_logp_rows_multi(const NlpJobs& jobs, int B, int k, cudaStream_t st) {
  const int64_t R = (int64_t)B * k;
  if (R == 0 || jobs.count == 0) return 0;
  int64_t blocks = (R + 7) / 8;
  if (blocks > 148 * 32) blocks = 148 * 32;
  CUDA_TRY(launch_pdl(normal_logp_rows_multi_kernel, dim3(dim3((unsigned)blocks, (unsigned)jobs.count)), dim3(256), 0, st, jobs, B, R));
  LAUNCH_CHECK();
  return 0;
}

__device__ __forceinline__ void kl_rows_body(const uint3 vb_, const dim3 vg_, const float* __restrict__ mu, const float* __restrict__ sd, int B, int d,
                               float* __restrict__ out) {
  VBLOCK_SHADOW;
  // tfd.kl_divergence(N(mu,sd), N(0,1)) = 0.5 mu^2 + 0.5 expm1(2 log sd) - log sd   (A.2)
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  float s = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float m = mu[b * d + j], ls = logf(sd[b * d + j]);
    s += 0.5f * m * m + 0.5f * expm1f(2.f * ls) - ls;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[b] = s;
}
__global__ void __launch_bounds__(256) kl_rows_kernel(const float* __restrict__ mu, const float* __restrict__ sd, int B, int d, float* __restrict__ out) {
  pdl_wait();
  kl_rows_body(blockIdx, gridDim, mu, sd, B, d, out);
}
int kl_rows(const float* mu, const float* sd, int B, int d, float* out, cudaStream_t st) {
  CUDA_TRY(launch_pdl(kl_rows_kernel, dim3((B + 7) / 8), dim3(256), 0, st, mu, sd, B, d, out));
  LAUNCH_CHECK();
  return 0;
}

__device__ __forceinline__ void bern_reduce_partials_body(const uint3 vb_, const dim3 vg_, const float* __restrict__ p, int64_t R, int nt, float* __restrict__ out) {
  VBLOCK_SHADOW;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < R; r += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < nt; ++t) s += p[r * nt + t];
    out[r] = s;
  }
}
__global__ void __launch_bounds__(256) bern_reduce_partials_kernel(const float* __restrict__ p, int64_t R, int nt, float* __restrict__ out) {
  pdl_wait();
  bern_reduce_partials_body(blockIdx, gridDim, p, R, nt, out);
}
int bern_reduce_partials(const float* partial, int R, int nt, float* lpxz, cudaStream_t st) {
  if (R == 0) return 0;
  int blocks = (R + 255) / 256; if (blocks > 148 * 8) blocks = 148 * 8;
  CUDA_TRY(launch_pdl(bern_reduce_partials_kernel, dim3(blocks), dim3(256), 0, st, partial, R, nt, lpxz));
  LAUNCH_CHECK();
  return 0;
}

__global__ void bern_logp_rows_kernel(const float* __restrict__ logits, const float* __restrict__ x, int64_t R,
                                      int B, int n, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < R; r += nwarps) {
    const float* l = logits + r * n;
    const float* xr = x + (r % B) * (int64_t)n;
    float s = 0.f;
    for (int j = lane; j < n; j += 32) {
      const float v = l[j], xv = xr[j];
      const float sp = fmaxf(v, 0.f) + log1pf(expf(-fabsf(v)));
      s += -(1.f - xv) * sp - xv * (sp - v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[r] = s;
  }
}
int bern_logp_rows(const float* logits, const float* x, int R, int B, int n, float* lpxz, cudaStream_t st) {
  if (R == 0) return 0;
  int64_t blocks = ((int64_t)R + 7) / 8; if (blocks > 148 * 32) blocks = 148 * 32;
  bern_logp_rows_kernel<<<(int)blocks, 256, 0, st>>>(logits, x, R, B, n, lpxz);
  LAUNCH_CHECK();
  return 0;
}

// ---- log-weights, log-mean-exp over k, softmax, SNIS: one CTA per image ----
__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < nw; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

__device__ __forceinline__ void logw_reduce_body(const uint3 vb_, const dim3 vg_, LogwArgs a) {
  VBLOCK_SHADOW;
  __shared__ float red[8];
  extern __shared__ float s_al[];      // normalised weights of this image when k fits (a.al_in_smem)
  const int b = blockIdx.x, B = a.B, k = a.k;
  // pass 1: log_w and its max
  float mx = -INFINITY, sm = 0.f;
  for (int s = threadIdx.x; s < k; s += blockDim.x) {
    const int64_t r = (int64_t)s * B + b;
    float lw;
    if (a.two_layer) lw = a.lpxz[r] + a.lp12[r] + a.lpz[r] - a.lq1[r] - a.lq2[r];     // iwae2.py:126
    else lw = a.lpxz[r] + a.beta * (a.lpz[r] - a.lq1[r]);                              // iwae1.py:113
    if (a.log_w) a.log_w[r] = lw;
    if (a.g_out && a.g_vae) a.g_out[r] = -1.f / ((float)k * (float)a.B_norm);
    mx = fmaxf(mx, lw);
    sm += lw;
  }
  mx = block_reduce(mx, red, true);
  sm = block_reduce(sm, red, false);
  // pass 2: sum exp(lw - max)
  float se = 0.f;
  for (int s = threadIdx.x; s < k; s += blockDim.x) {
    const int64_t r = (int64_t)s * B + b;
    float lw = a.two_layer ? a.lpxz[r] + a.lp12[r] + a.lpz[r] - a.lq1[r] - a.lq2[r]
                           : a.lpxz[r] + a.beta * (a.lpz[r] - a.lq1[r]);
    se += expf(lw - mx);
  }
  se = block_reduce(se, red, false);
  if (threadIdx.x == 0) {
    if (a.iwae) a.iwae[b] = logf(se / (float)k) + mx;       // utils.logmeanexp (utils.py:6-8)
    if (a.vae) a.vae[b] = sm / (float)k;
  }
  const bool need_al = a.al || a.eq14 || a.dreg || a.snis1 || a.snis2 || (a.g_out && !a.g_vae);
  if (!need_al) return;
  // pass 3: normalised weights, eq14, DReG
  float e14 = 0.f, dr = 0.f;
  for (int s = threadIdx.x; s < k; s += blockDim.x) {
    const int64_t r = (int64_t)s * B + b;
    float lw = a.two_layer ? a.lpxz[r] + a.lp12[r] + a.lpz[r] - a.lq1[r] - a.lq2[r]
                           : a.lpxz[r] + a.beta * (a.lpz[r] - a.lq1[r]);
    const float w = expf(lw - mx) / se;
    if (a.al) a.al[r] = w;
    if (a.g_out && !a.g_vae) a.g_out[r] = -w / (float)a.B_norm;
    if (a.al_in_smem) s_al[s] = w;
    e14 += w * lw;
    if (a.dreg) dr += w * w * (a.lpz[r] + a.lpxz[r] - a.lq_stopped[r]);                // task02.py:70-73
  }
  if (a.eq14) { e14 = block_reduce(e14, red, false); if (threadIdx.x == 0) a.eq14[b] = e14; }
  if (a.dreg) { dr = block_reduce(dr, red, false); if (threadIdx.x == 0) a.dreg[b] = dr; }
  // pass 4: snis_z = sum_k al * z  (thread per latent dim, coalesced over the dim)
  __syncthreads();
  for (int which = 0; which < 2; ++which) {
    float* out = which ? a.snis2 : a.snis1;
    const float* z = which ? a.z2 : a.z1;
    const int d = which ? a.d2 : a.d1;
    if (!out) continue;
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
      float acc = 0.f;
#pragma unroll 10
      for (int s = 0; s < k; ++s) {          // (unrolled: the loads of z are independent, the sum keeps its order)
        const int64_t r = (int64_t)s * B + b;
        float w;
        if (a.al_in_smem) w = s_al[s];
        else {
          float lw = a.two_layer ? a.lpxz[r] + a.lp12[r] + a.lpz[r] - a.lq1[r] - a.lq2[r]
                                 : a.lpxz[r] + a.beta * (a.lpz[r] - a.lq1[r]);
          w = expf(lw - mx) / se;
        }
        acc += w * z[r * d + j];
      }
      out[(int64_t)b * d + j] = acc;
    }
  }
}
__global__ void __launch_bounds__(256) logw_reduce_kernel(LogwArgs a) {
  pdl_wait();
  logw_reduce_body(blockIdx, gridDim, a);
}
int logw_reduce(const LogwArgs& a0, cudaStream_t st) {
  if (a0.B == 0) return 0;
  LogwArgs a = a0;
  a.al_in_smem = (a.k <= 10240 && (a.snis1 || a.snis2)) ? 1 : 0;
  CUDA_TRY(launch_pdl(logw_reduce_kernel, dim3(a.B), dim3(256), a.al_in_smem ? sizeof(float) * a.k : 0, st, a));
  LAUNCH_CHECK();
  return 0;
}

// Stage 1 of the batch means for large calls: CTA c sums slice c of each of the eight arrays (fixed order -> deterministic)
__device__ __forceinline__ void scalars_partial_body(const uint3 vb_, const dim3 vg_, ScalarArgs a) {
  VBLOCK_SHADOW;
  __shared__ float red[8][8];
  const float* src[8] = {a.vae, a.iwae, a.eq14, a.kl, a.dreg, a.lpxz, a.lpz, a.lq1};
  const int64_t len[8] = {a.B, a.B, a.B, a.B, a.B, a.R, a.R, a.R};
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s[j] = 0.f;
    if (!src[j]) continue;
    const int64_t per = (len[j] + gridDim.x - 1) / gridDim.x;
    const int64_t lo = min((int64_t)blockIdx.x * per, len[j]), hi = min(lo + per, len[j]);
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) s[j] += src[j][i];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int j = 0; j < 8; ++j) red[j][threadIdx.x >> 5] = s[j];
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    a.partial[(int64_t)blockIdx.x * 8 + threadIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) scalars_partial_kernel(ScalarArgs a) {
  pdl_wait();
  scalars_partial_body(blockIdx, gridDim, a);
}

// Batch-mean scalars of the result dict (iwae1.py:119-134): one CTA, fixed order -> deterministic.
__device__ __forceinline__ void finalize_scalars_body(const uint3 vb_, const dim3 vg_, ScalarArgs a) {
  VBLOCK_SHADOW;
  // eight means in ONE pass: all loads are issued before the first reduction (eight serial block reductions, each
  // waiting for its own global loads, were a 12 us kernel on the train step's critical path)
  __shared__ float red[8][8];
  const float* src[8] = {a.vae, a.iwae, a.eq14, a.kl, a.dreg, a.lpxz, a.lpz, a.lq1};
  const int64_t len[8] = {a.B, a.B, a.B, a.B, a.B, a.R, a.R, a.R};
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s[j] = 0.f;
    if (!src[j]) continue;
    if (a.partial) { for (int i = threadIdx.x; i < SCALAR_PARTS; i += blockDim.x) s[j] += a.partial[i * 8 + j]; }
    else for (int64_t i = threadIdx.x; i < len[j]; i += blockDim.x) s[j] += src[j][i];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int j = 0; j < 8; ++j) red[j][threadIdx.x >> 5] = s[j];
  __syncthreads();
  float mean[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[j][w];
    mean[j] = (src[j] && len[j] > 0) ? t / (float)len[j] : 0.f;
  }
  const float m_vae = mean[0], m_iwae = mean[1], m_eq14 = mean[2], m_kl = mean[3], m_dreg = mean[4];
  const float m_lpxz = mean[5], m_lpz = mean[6], m_lq1 = mean[7];
  if (threadIdx.x == 0) {
    a.out[0] = m_vae;
    a.out[1] = m_lpxz - a.beta * m_kl;      // vae_elbo_kl (iwae1.py:121)
    a.out[2] = m_iwae;
    a.out[3] = m_eq14;
    a.out[4] = -m_dreg;                     // inference_loss (task02.py:76)
    a.out[5] = m_lpxz; a.out[6] = m_lpz; a.out[7] = m_lq1;
  }
}
__global__ void __launch_bounds__(256) finalize_scalars_kernel(ScalarArgs a) {
  pdl_wait();
  finalize_scalars_body(blockIdx, gridDim, a);
}
int finalize_scalars(const ScalarArgs& a, cudaStream_t st) {
  if (a.partial) { CUDA_TRY(launch_pdl(scalars_partial_kernel, dim3(SCALAR_PARTS), dim3(256), 0, st, a)); LAUNCH_CHECK(); }
  CUDA_TRY(launch_pdl(finalize_scalars_kernel, dim3(1), dim3(256), 0, st, a));
  LAUNCH_CHECK();
  return 0;
}

__global__ void sum_to_double_kernel(const float* __restrict__ v, int64_t n, double* acc, int accumulate) {
  // single CTA, fixed summation order -> deterministic
  __shared__ double red[256];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += (double)v[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *acc = (accumulate ? *acc : 0.0) + red[0];
}
int sum_to_double(const float* v, int64_t n, double* acc, int accumulate, cudaStream_t st) {
  sum_to_double_kernel<<<1, 256, 0, st>>>(v, n, acc, accumulate);
  LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// Backward
// =============================================================================================
// g[r] = d(-objective)/d(log_w[r])  (A.4): iwae_elbo / iwae_eq14 / DReG-decoder -> -al/B ; vae -> -1/(kB)
__device__ __forceinline__ void row_coef_body(const uint3 vb_, const dim3 vg_, const float* __restrict__ al, int64_t R, int k, int objective, int B_norm,
                                float* __restrict__ g) {
  VBLOCK_SHADOW;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < R; r += (int64_t)gridDim.x * blockDim.x) {
    float v;
    if (objective == IWAE_OBJ_VAE_ELBO || objective == IWAE_OBJ_VAE_ELBO_KL) v = -1.f / ((float)k * (float)B_norm);
    else v = -al[r] / (float)B_norm;
    g[r] = v;
  }
}
__global__ void __launch_bounds__(256) row_coef_kernel(const float* __restrict__ al, int64_t R, int k, int objective, int B_norm, float* __restrict__ g) {
  pdl_wait();
  row_coef_body(blockIdx, gridDim, al, R, k, objective, B_norm, g);
}
int row_coef(const float* al, int B, int k, int objective, int B_norm, float* g, cudaStream_t st) {
  const int64_t R = (int64_t)B * k;
  if (R == 0) return 0;
  int blocks = (int)((R + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
  CUDA_TRY(launch_pdl(row_coef_kernel, dim3(blocks), dim3(256), 0, st, al, R, k, objective, B_norm, g));
  LAUNCH_CHECK();
  return 0;
}

// dl[r,j] = g[r] * (x[b,j] - sigmoid(l[r,j]))
__device__ __forceinline__ void dlogits_body(const uint3 vb_, const dim3 vg_, const float* __restrict__ logits, const float* __restrict__ x,
                               const float* __restrict__ g, int64_t R, int B, int n, float* __restrict__ dl) {
  VBLOCK_SHADOW;
  const int64_t total = R * n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / n;
    const int j = (int)(i % n);
    const float l = logits[i];
    const float sg = 1.f / (1.f + expf(-l));
    dl[i] = g[r] * (x[(r % B) * (int64_t)n + j] - sg);
  }
}
__global__ void __launch_bounds__(256) dlogits_kernel(const float* __restrict__ logits, const float* __restrict__ x, const float* __restrict__ g, int64_t R, int B, int n, float* __restrict__ dl) {
  pdl_wait();
  dlogits_body(blockIdx, gridDim, logits, x, g, R, B, n, dl);
}
int dlogits_f32(const float* logits, const float* x, const float* g, int R, int B, int n, float* dl, cudaStream_t st) {
  const int64_t total = (int64_t)R * n;
  if (total == 0) return 0;
  int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  CUDA_TRY(launch_pdl(dlogits_kernel, dim3(blocks), dim3(256), 0, st, logits, x, g, R, B, n, dl));
  LAUNCH_CHECK();
  return 0;
}

// First stochastic layer.  One CTA per (image, chunk of 32 latent dims); thread (ks, j) sums dim j over the samples
// s = ks, ks+8, ... and the 8 partial sums are added in a fixed order through shared memory.  (One CTA per image with
// d threads x 256/d sample slices was a 23 us serial chain of load latencies on the step's critical path.)
// Uses the total-derivative identities of A.4: d lq/d mu = 0, d lq/d std = -1/std.
__device__ __forceinline__ void latent1_bwd_body(const uint3 vb_, const dim3 vg_, Latent1BwdArgs a) {
  VBLOCK_SHADOW;
  __shared__ float red[4][256];
  const int b = blockIdx.x;
  constexpr int KS = 8;                             // sample slices handled in parallel (one warp each)
  const int jl = threadIdx.x & 31, ks = threadIdx.x >> 5;
  const int j = blockIdx.y * 32 + jl;
  const bool active = j < a.d;
  const bool dreg = (a.objective == IWAE_OBJ_DREG);
  const bool klobj = (a.objective == IWAE_OBJ_VAE_ELBO_KL);
  float acc_mu = 0.f, acc_sd = 0.f, acc_pmu = 0.f, acc_psd = 0.f;
  const int idx = b * a.d + j;
  float mu = 0.f, sd = 1.f;
  if (active) {
    mu = a.mu[idx]; sd = a.std[idx];
    // (unrolled: the ~7 global loads of different samples are independent; one CTA per image leaves most SMs idle,
    //  so this kernel's time is its serial load-latency chain)
#pragma unroll 7
    for (int s = ks; s < a.k; s += KS) {
      const int64_t r = (int64_t)s * a.B + b;
      const float g = a.g[r];
      const float z = a.z[r * a.d + j], e = a.eps[r * a.d + j];
      float dz = a.dz_dec[r * a.ld_dz + j];
      if (a.dz_q2) dz = __fadd_rn(dz, a.dz_q2[r * a.d + j]);
      // (explicit roundings / fmaf below: the body is compiled into two kernels - its own and the train-step megakernel -
      //  and the compiler's choice of which a*b+c to contract must not differ between them)
      if (dreg) {
        // task02.py:60-76: encoder signal = al * (decoder-path dz) - (al^2/B) * d(lpz - lq_stopped)/dz
        const float al = a.al[r];
        const float sd2 = __fadd_rn(sd, 1e-6f);
        const float inner = __fadd_rn(-z, __fdiv_rn(__fsub_rn(z, mu), __fmul_rn(sd2, sd2)));
        dz = __fsub_rn(__fmul_rn(al, dz), __fmul_rn(__fdiv_rn(__fmul_rn(al, al), (float)a.B_norm), inner));
        acc_mu = __fadd_rn(acc_mu, dz);
        acc_sd = fmaf(dz, e, acc_sd);
        continue;
      }
      if (!klobj) {
        if (a.prior_rows == 0) {
          dz = fmaf(__fmul_rn(g, a.beta), -z, dz);
        } else {
          const int64_t pr = (a.prior_rows == 1) ? (int64_t)b * a.d + j : r * a.d + j;
          const float pm = a.pmu[pr], ps = a.pstd[pr];
          const float gbt = (a.prior_rows == 1) ? __fmul_rn(g, a.beta) : g;    // iwae2 ignores beta (iwae2.py:128)
          const float zc = __fsub_rn(z, pm);
          const float t = __fdiv_rn(zc, __fmul_rn(ps, ps));
          dz = fmaf(gbt, -t, dz);
          const float dpm = __fmul_rn(gbt, t);
          const float dps = __fmul_rn(gbt, __fsub_rn(__fdiv_rn(__fmul_rn(zc, t), ps), __fdiv_rn(1.f, ps)));
          if (a.prior_rows == 1) { acc_pmu = __fadd_rn(acc_pmu, dpm); acc_psd = __fadd_rn(acc_psd, dps); }
          else if (a.dpmu) { a.dpmu[pr] = dpm; a.dpa[pr] = __fmul_rn(dps, __fsub_rn(ps, 1e-6f)); }
        }
        const float gb = (a.prior_rows == 2) ? g : __fmul_rn(g, a.beta);
        acc_sd = __fadd_rn(acc_sd, __fdiv_rn(gb, sd));                 // -beta * d lq/d std (total) = +beta/std
      }
      acc_mu = __fadd_rn(acc_mu, dz);
      acc_sd = fmaf(dz, e, acc_sd);
    }
  }
  red[0][threadIdx.x] = acc_mu; red[1][threadIdx.x] = acc_sd; red[2][threadIdx.x] = acc_pmu; red[3][threadIdx.x] = acc_psd;
  __syncthreads();
  if (active && ks == 0) {
    for (int q = 1; q < KS; ++q) {
      acc_mu = __fadd_rn(acc_mu, red[0][q * 32 + jl]); acc_sd = __fadd_rn(acc_sd, red[1][q * 32 + jl]);
      acc_pmu = __fadd_rn(acc_pmu, red[2][q * 32 + jl]); acc_psd = __fadd_rn(acc_psd, red[3][q * 32 + jl]);
    }
    if (klobj) {
      // loss = -mean lpxz + beta * mean_b kl ; d kl/d mu = mu, d kl/d std = std - 1/std
      const float bn = __fdiv_rn(a.beta, (float)a.B_norm);
      acc_mu = fmaf(bn, mu, acc_mu);
      acc_sd = fmaf(bn, __fsub_rn(sd, __fdiv_rn(1.f, sd)), acc_sd);
    }
    a.dmu[idx] = acc_mu;
    a.da[idx] = __fmul_rn(acc_sd, __fsub_rn(sd, 1e-6f));      // std = exp(a) + 1e-6
    if (a.prior_rows == 1 && a.dpmu) {
      a.dpmu[idx] = acc_pmu;
      a.dpa[idx] = __fmul_rn(acc_psd, __fsub_rn(a.pstd[idx], 1e-6f));
    }
  }
}
__global__ void __launch_bounds__(256) latent1_bwd_kernel(Latent1BwdArgs a) {
  pdl_wait();
  latent1_bwd_body(blockIdx, gridDim, a);
}
int latent1_bwd(const Latent1BwdArgs& a, cudaStream_t st) {
  if (a.B * a.d == 0) return 0;
  CUDA_TRY(launch_pdl(latent1_bwd_kernel, dim3(a.B, (a.d + 31) / 32), dim3(256), 0, st, a));
  LAUNCH_CHECK();
  return 0;
}

// Gradients reaching the parameters (pmu, pstd pre-activation) of a per-row Gaussian log p(z; pmu, pstd)
// weighted by g[r] (IWAE2's lpz1z2, iwae2.py:120).
__device__ __forceinline__ void gauss_rows_bwd_body(const uint3 vb_, const dim3 vg_, const float* __restrict__ g, const float* __restrict__ z,
                                      const float* __restrict__ pmu, const float* __restrict__ pstd, int64_t R, int d,
                                      float* __restrict__ dpmu, float* __restrict__ dpa) {
  VBLOCK_SHADOW;
  const int64_t total = R * d;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float gg = g[i / d];
    const float pm = pmu[i], ps = pstd[i];
    const float t = (z[i] - pm) / (ps * ps);
    dpmu[i] = gg * t;
    dpa[i] = gg * ((z[i] - pm) * t / ps - 1.f / ps) * (ps - 1e-6f);
  }
}
__global__ void __launch_bounds__(256) gauss_rows_bwd_kernel(const float* __restrict__ g, const float* __restrict__ z, const float* __restrict__ pmu, const float* __restrict__ pstd, int64_t R, int d, float* __restrict__ dpmu, float* __restrict__ dpa) {
  pdl_wait();
  gauss_rows_bwd_body(blockIdx, gridDim, g, z, pmu, pstd, R, d, dpmu, dpa);
}
int gauss_rows_bwd(const float* g, const float* z, const float* pmu, const float* pstd, int64_t R, int d,
                   float* dpmu, float* dpa, cudaStream_t st) {
  const int64_t total = R * d;
  if (total == 0) return 0;
  int blocks = (int)((total + 255) / 256 > 148 * 8 ? 148 * 8 : (total + 255) / 256);
  CUDA_TRY(launch_pdl(gauss_rows_bwd_kernel, dim3(blocks), dim3(256), 0, st, g, z, pmu, pstd, R, d, dpmu, dpa));
  LAUNCH_CHECK();
  return 0;
}

// Second stochastic layer of IWAE2 (all per row): z2 = mu2 + std2*eps2, lpz2 = N(0,1), -lq2 total.
__device__ __forceinline__ void latent2_bwd_body(const uint3 vb_, const dim3 vg_, Latent2BwdArgs a) {
  VBLOCK_SHADOW;
  const int64_t total = (int64_t)a.R * a.d;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / a.d;
    const float g = a.g[r];
    const float sd = a.std[i];
    const float dz = a.dz_p[i] + g * (-a.z[i]);
    a.dmu[i] = dz;
    a.da[i] = (dz * a.eps[i] + g / sd) * (sd - 1e-6f);
  }
}
__global__ void __launch_bounds__(256) latent2_bwd_kernel(Latent2BwdArgs a) {
  pdl_wait();
  latent2_bwd_body(blockIdx, gridDim, a);
}
int latent2_bwd(const Latent2BwdArgs& a, cudaStream_t st) {
  const int64_t total = (int64_t)a.R * a.d;
  if (total == 0) return 0;
  int blocks = (int)((total + 255) / 256 > 148 * 8 ? 148 * 8 : (total + 255) / 256);
  CUDA_TRY(launch_pdl(latent2_bwd_kernel, dim3(blocks), dim3(256), 0, st, a));
  LAUNCH_CHECK();
  return 0;
}

// ---- Keras Adam (A.6), 128-bit vectorised with a scalar tail ----
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, int vec, float lr_t, float b1, float b2, float eps) {
  const int64_t n4 = vec ? (n >> 2) : 0;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (int64_t i = tid; i < n4; i += stride) {
    float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
#define ADAM1(c)                                             \
    mm.c = b1 * mm.c + (1.f - b1) * gg.c;                    \
    vv.c = b2 * vv.c + (1.f - b2) * gg.c * gg.c;             \
    pp.c = pp.c - lr_t * mm.c / (sqrtf(vv.c) + eps);
    ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  for (int64_t i = (n4 << 2) + tid; i < n; i += stride) {
    const float gg = g[i];
    const float mm = b1 * m[i] + (1.f - b1) * gg;
    const float vv = b2 * v[i] + (1.f - b2) * gg * gg;
    m[i] = mm; v[i] = vv;
    p[i] = p[i] - lr_t * mm / (sqrtf(vv) + eps);
  }
}
int adam_f32(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
             int64_t t, cudaStream_t st) {
  if (n == 0) return 0;
  // lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t), computed in double on the host like Keras does in fp32 ops
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)b2, (double)t)) / (1.0 - pow((double)b1, (double)t));
  const bool aligned = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
  int blocks = (int)((n / 4 + 255) / 256); if (blocks > 148 * 4) blocks = 148 * 4; if (blocks < 1) blocks = 1;
  adam_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, n, aligned ? 1 : 0, (float)lr_t, b1, b2, eps);
  LAUNCH_CHECK();
  return 0;
}

// ---- split-K gradient partials: fixed-order reduction, optionally fused with the Keras-Adam update ----
__device__ __forceinline__ void grad_reduce_body(const uint3 vb_, const dim3 vg_, const float* __restrict__ ws, int splits, int64_t n, float* __restrict__ g,
                                   float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, float lr_t,
                                   float b1, float b2, float eps, const StepDyn* __restrict__ dyn) {
  VBLOCK_SHADOW;
  if (dyn) { lr_t = dyn->lr_t; b1 = dyn->b1; b2 = dyn->b2; eps = dyn->eps; }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float gg = 0.f;
    for (int s = 0; s < splits; ++s) gg += ws[(int64_t)s * n + i];
    if (g) g[i] = gg;
    if (p) {
      const float mm = b1 * m[i] + (1.f - b1) * gg;
      const float vv = b2 * v[i] + (1.f - b2) * gg * gg;
      m[i] = mm; v[i] = vv;
      p[i] = p[i] - lr_t * mm / (sqrtf(vv) + eps);
    }
  }
}
__global__ void __launch_bounds__(256) grad_reduce_kernel(const float* __restrict__ ws, int splits, int64_t n, float* __restrict__ g, float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, float lr_t, float b1, float b2, float eps, const StepDyn* __restrict__ dyn) {
  pdl_wait();
  grad_reduce_body(blockIdx, gridDim, ws, splits, n, g, p, m, v, lr_t, b1, b2, eps, dyn);
}
int grad_reduce(const float* ws, int splits, int64_t n, float* grads, cudaStream_t st) {
  if (n == 0) return 0;
  int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  CUDA_TRY(launch_pdl(grad_reduce_kernel, dim3(blocks), dim3(256), 0, st, ws, splits, n, grads, nullptr, nullptr, nullptr, 0.f, 0.f, 0.f, 0.f, nullptr));
  LAUNCH_CHECK();
  return 0;
}
int adam_from_partials(float* p, const float* ws, int splits, float* m, float* v, int64_t n, float lr, float b1, float b2,
                       float eps, int64_t t, cudaStream_t st) {
  if (n == 0) return 0;
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)b2, (double)t)) / (1.0 - pow((double)b1, (double)t));
  int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  CUDA_TRY(launch_pdl(grad_reduce_kernel, dim3(blocks), dim3(256), 0, st, ws, splits, n, nullptr, p, m, v, (float)lr_t, b1, b2, eps, g_step_dyn));
  LAUNCH_CHECK();
  return 0;
}

// ---- utils.bernoullisample on the device ----
__global__ void bernoulli_sample_kernel(const float* __restrict__ p, int64_t n, uint64_t seed, uint64_t epoch,
                                        float* __restrict__ out) {
  const int64_t nq = (n + 3) / 4;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)epoch, 0x42u),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t i = 4 * q + j;
      if (i < n) {
        const float u = (float)(rr[j] >> 8) * 5.9604644775390625e-8f;
        out[i] = (u < p[i]) ? 1.f : 0.f;
      }
    }
  }
}
int bernoulli_sample_f32(const float* p, int64_t n, uint64_t seed, uint64_t epoch, float* out, cudaStream_t st) {
  if (n == 0) return 0;
  int64_t blocks = ((n + 3) / 4 + 255) / 256; if (blocks > 148 * 16) blocks = 148 * 16;
  bernoulli_sample_kernel<<<(int)blocks, 256, 0, st>>>(p, n, seed, epoch, out);
  LAUNCH_CHECK();
  return 0;
}

// ---- scipy.special.softmax on an fp32 vector: exp(x - logsumexp(x)) ----
__global__ void __launch_bounds__(256) softmax_vec_kernel(const float* __restrict__ x, int L, float* __restrict__ out) {
  __shared__ float red[8];
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < L; i += blockDim.x) mx = fmaxf(mx, x[i]);
  mx = block_reduce(mx, red, true);
  float se = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) se += expf(x[i] - mx);
  se = block_reduce(se, red, false);
  const float lse = logf(se) + mx;
  for (int i = threadIdx.x; i < L; i += blockDim.x) out[i] = expf(x[i] - lse);
}
int softmax_vec_f32(const float* x, int L, float* out, cudaStream_t st) {
  softmax_vec_kernel<<<1, 256, 0, st>>>(x, L, out);
  LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// SIR: numpy legacy RandomState.choice(L, size, replace=False, p) (A.7), bit-exact.
// The cdf is a SEQUENTIAL fp64 inclusive scan (np.cumsum order) done by one thread; searchsorted and
// the first-occurrence de-duplication run on the whole CTA.  L = 10,000: ~50 us per round.
// =============================================================================================
__global__ void __launch_bounds__(256) sir_kernel(const float* __restrict__ al, int L, const double* __restrict__ u,
                                                  int n_u, int size, int64_t* __restrict__ out_idx,
                                                  int32_t* __restrict__ out_consumed, double* __restrict__ p,
                                                  double* __restrict__ cdf, int32_t* __restrict__ cand) {
  __shared__ int s_nuniq, s_pos, s_fail;
  __shared__ int s_keep[1024];
  for (int i = threadIdx.x; i < L; i += blockDim.x) p[i] = (double)al[i];
  if (threadIdx.x == 0) { s_nuniq = 0; s_pos = 0; s_fail = 0; }
  __syncthreads();
  while (true) {
    const int n_uniq = s_nuniq, pos = s_pos;
    if (n_uniq >= size || s_fail) break;
    const int n = size - n_uniq;
    if (pos + n > n_u) { if (threadIdx.x == 0) s_fail = 1; __syncthreads(); break; }
    for (int i = threadIdx.x; i < n_uniq; i += blockDim.x) p[out_idx[i]] = 0.0;       // p[found] = 0
    __syncthreads();
    if (threadIdx.x == 0) {                                                           // cdf = cumsum(p)
      double c = 0.0;
      for (int i = 0; i < L; ++i) { c += p[i]; cdf[i] = c; }
    }
    __syncthreads();
    const double tot = cdf[L - 1];
    __syncthreads();
    for (int i = threadIdx.x; i < L; i += blockDim.x) cdf[i] = cdf[i] / tot;          // cdf /= cdf[-1]
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {                               // searchsorted(side='right')
      const double x = u[pos + i];
      int lo = 0, hi = L;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (cdf[mid] <= x) lo = mid + 1; else hi = mid; }
      cand[i] = lo;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {                               // first occurrences, draw order
      int keep = 1;
      const int c = cand[i];
      for (int j = 0; j < i; ++j) if (cand[j] == c) { keep = 0; break; }
      s_keep[i] = keep;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int w = n_uniq;
      for (int i = 0; i < n; ++i) if (s_keep[i]) out_idx[w++] = (int64_t)cand[i];
      s_nuniq = w;
      s_pos = pos + n;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *out_consumed = s_fail ? -1 : s_pos;
}

int sir_launch(const float* al, int L, const double* u, int n_u, int size, int64_t* out_idx, int32_t* out_consumed,
               cudaStream_t st) {
  IWAE_REQUIRE(size <= 1024 && size <= L, "iwae_sir: size must be <= min(L, 1024)");
  // scratch: p[L], cdf[L] doubles + cand[size] ints, allocated stream-ordered
  double* scratch = nullptr;
  const size_t bytes = sizeof(double) * 2 * (size_t)L + sizeof(int32_t) * 1024;
  CUDA_TRY(cudaMallocAsync((void**)&scratch, bytes, st));
  sir_kernel<<<1, 256, 0, st>>>(al, L, u, n_u, size, out_idx, out_consumed, scratch, scratch + L,
                               reinterpret_cast<int32_t*>(scratch + 2 * (size_t)L));
  LAUNCH_CHECK();
  CUDA_TRY(cudaFreeAsync(scratch, st));
  return 0;
}


// ---- split-K forward epilogue ----
__device__ __forceinline__ void splitk_bias_act_body(const uint3 vb_, const dim3 vg_, const float* __restrict__ partial, int splits, int M, int N,
                                       const float* __restrict__ bias, int act, float* __restrict__ out) {
  VBLOCK_SHADOW;
  const int total = M * N;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    float v = 0.f;
    for (int s = 0; s < splits; ++s) v += partial[(size_t)s * total + i];
    v += bias[i % N];
    if (act == ACT_TANH) v = tanhf(v);
    else if (act == ACT_EXP_EPS) v = expf(v) + 1e-6f;
    out[i] = v;
  }
}
__global__ void __launch_bounds__(256) splitk_bias_act_kernel(const float* __restrict__ partial, int splits, int M, int N, const float* __restrict__ bias, int act, float* __restrict__ out) {
  pdl_wait();
  splitk_bias_act_body(blockIdx, gridDim, partial, splits, M, N, bias, act, out);
}
int splitk_bias_act(const float* partial, int splits, int M, int N, const float* bias, int act, float* out, cudaStream_t st) {
  if (M * N == 0) return 0;
  CUDA_TRY(launch_pdl(splitk_bias_act_kernel, dim3((M * N + 255) / 256), dim3(256), 0, st, partial, splits, M, N, bias, act, out));
  LAUNCH_CHECK();
  return 0;
}

__device__ __forceinline__ void splitk_mul_h_body(const uint3 vb_, const dim3 vg_, const float* __restrict__ partial, int splits, int M, int N, const float* __restrict__ H,
                                    int64_t ld_h, int accumulate, float* __restrict__ dX, int64_t ld_dx) {
  VBLOCK_SHADOW;
  const int total = M * N;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int m = i / N, n = i % N;
    float v = 0.f;
    for (int s = 0; s < splits; ++s) v += partial[(size_t)s * total + i];
    if (accumulate) v += dX[(int64_t)m * ld_dx + n];
    if (H) { const float h = H[(int64_t)m * ld_h + n]; v *= (1.f - h * h); }
    dX[(int64_t)m * ld_dx + n] = v;
  }
}
__global__ void __launch_bounds__(256) splitk_mul_h_kernel(const float* __restrict__ partial, int splits, int M, int N, const float* __restrict__ H, int64_t ld_h, int accumulate, float* __restrict__ dX, int64_t ld_dx) {
  pdl_wait();
  splitk_mul_h_body(blockIdx, gridDim, partial, splits, M, N, H, ld_h, accumulate, dX, ld_dx);
}
int splitk_mul_h(const float* partial, int splits, int M, int N, const float* H, int64_t ld_h, int accumulate, float* dX,
                 int64_t ld_dx, cudaStream_t st) {
  if (M * N == 0) return 0;
  const int blocks = std::min((M * N + 255) / 256, 148 * 8);
  CUDA_TRY(launch_pdl(splitk_mul_h_kernel, dim3(blocks), dim3(256), 0, st, partial, splits, M, N, H, ld_h, accumulate, dX, ld_dx));
  LAUNCH_CHECK();
  return 0;
}

// ---- train-step graph replay helpers ----
thread_local const StepDyn* g_step_dyn = nullptr;   // per host thread: one context per thread may capture a train-step graph

__global__ void set_step_dyn_kernel(StepDyn* d, StepDyn v) { *d = v; }
int set_step_dyn(StepDyn* dev, const StepDyn& v, cudaStream_t st) {
  set_step_dyn_kernel<<<1, 1, 0, st>>>(dev, v);
  LAUNCH_CHECK();
  return 0;
}

__global__ void copy_table_kernel(CopyTable t) {
  const float* __restrict__ src = t.src[blockIdx.y];
  float* __restrict__ dst = t.dst[blockIdx.y];
  const long long n = t.n[blockIdx.y];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
}
int copy_table(const CopyTable& t, cudaStream_t st) {
  if (t.count == 0) return 0;
  long long nmax = 0;
  for (int i = 0; i < t.count; ++i) nmax = std::max(nmax, t.n[i]);
  const int bx = (int)std::min<long long>((nmax + 255) / 256, 296);
  copy_table_kernel<<<dim3(bx, t.count), 256, 0, st>>>(t);
  LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// Fused Dense chain (train step): a tile of TM rows goes through up to THREE Dense layers in one launch.
// At the training shapes (k*B = 1,000 rows, 50..200 units) a Dense layer is ~1 us of FFMAs behind ~6 us of launch,
// drain and pipeline-fill latency, and a BasicBlock (iwae1.py:36-44) is four of them in a row.  Here one CTA keeps its
// rows' activations in shared memory, streams each layer's weights from L2 (coalesced: thread n reads W[k][n]) and
// writes every layer's output to global memory as well (the backward pass needs them).  The backward pass of a block
// is the SAME kernel on transposed weights (dX = dY W^T is a forward product with W^T; transpose_jobs builds them once
// per step, off the critical path) with the activation replaced by the multiplication with tanh' = 1 - H^2.
// =============================================================================================
constexpr int CHAIN_KT = 32;          // weight rows per shared-memory chunk
constexpr int CHAIN_TM = 8;           // rows per CTA

template <int V>
__device__ __forceinline__ void chain_cp_async(float* dst, const float* src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  if (V == 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
  else if (V == 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}
template <int V>
__device__ __forceinline__ void chain_copy_flat(float* dst, const float* src, int cnt, int tid) {
  for (int i = tid * V; i < cnt; i += 256 * V) chain_cp_async<V>(dst + i, src + i);
}
// Rows [c*KT, c*KT + KT) of layer L's weights -> wb, one cp.async group.  A kernel's rows are contiguous (ldw == its
// width), so the chunk is a flat copy; a second head goes to wb + KT*ldw.  Rows beyond K are zero up to a multiple of 4.
__device__ __forceinline__ void chain_load_chunk(const ChainLayer& L, int c, float* wb, int tid) {
  const int k0 = c * CHAIN_KT;
  const int rows = min(CHAIN_KT, L.K - k0);
  const int cnt = rows * L.ldw, cnt4 = ((rows + 3) & ~3) * L.ldw;
  const int parts = L.Wb ? 2 : 1;
  for (int p = 0; p < parts; ++p) {
    const float* src = (p ? L.Wb : L.W) + (int64_t)k0 * L.ldw;
    float* dst = wb + p * CHAIN_KT * L.ldw;
    if (L.vec == 4) chain_copy_flat<4>(dst, src, cnt, tid);
    else if (L.vec == 2) chain_copy_flat<2>(dst, src, cnt, tid);
    else chain_copy_flat<1>(dst, src, cnt, tid);
    for (int i = cnt + tid; i < cnt4; i += 256) dst[i] = 0.f;
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// acc[r] += sum_kk x[r][kk] * w[kk] over the valid rows of one chunk (steps of 4; rows beyond K are zero in both operands)
template <int RG>
__device__ __forceinline__ void chain_fma_chunk(const float* __restrict__ wb, int ws, const float* __restrict__ xin, int kpad,
                                                int ksteps, float (&acc)[8]) {
#pragma unroll 4
  for (int s = 0; s < ksteps; ++s) {
    const float w0 = wb[0], w1 = wb[ws], w2 = wb[2 * ws], w3 = wb[3 * ws];
#pragma unroll
    for (int r = 0; r < RG; ++r) {
      const float4 x = *reinterpret_cast<const float4*>(&xin[r * kpad]);                         // warp-wide broadcast
      acc[r] = fmaf(x.x, w0, acc[r]); acc[r] = fmaf(x.y, w1, acc[r]);
      acc[r] = fmaf(x.z, w2, acc[r]); acc[r] = fmaf(x.w, w3, acc[r]);
    }
    wb += 4 * ws; xin += 4;
  }
}

// wait until at most `n` of this thread's cp.async groups are pending (the operand must be an immediate)
__device__ __forceinline__ void chain_wait_pending(int n) {
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    case 7: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
    case 8: asm volatile("cp.async.wait_group 8;" ::: "memory"); break;
    case 9: asm volatile("cp.async.wait_group 9;" ::: "memory"); break;
    case 10: asm volatile("cp.async.wait_group 10;" ::: "memory"); break;
    case 11: asm volatile("cp.async.wait_group 11;" ::: "memory"); break;
    case 12: asm volatile("cp.async.wait_group 12;" ::: "memory"); break;
    case 13: asm volatile("cp.async.wait_group 13;" ::: "memory"); break;
    case 14: asm volatile("cp.async.wait_group 14;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 15;" ::: "memory"); break;
  }
}

__global__ void __launch_bounds__(256) dense_chain_kernel(const __grid_constant__ ChainArgs a) {
  extern __shared__ __align__(16) float chain_sm[];
  constexpr int TM = CHAIN_TM;
  const int kpad = a.kpad;
  float* in = chain_sm;
  float* nxt = chain_sm + (size_t)TM * kpad;
  float* wbuf = chain_sm + 2 * (size_t)TM * kpad;          // [a.njobs][a.wsz]: every weight chunk of the chain
  const int tid = threadIdx.x;
  const int r_base = blockIdx.x * TM;
  // The weights were written long before the kernel this launch programmatically depends on (Adam of the previous
  // step / the transposes at the start of this one): ALL of them are requested BEFORE the dependency wait, one cp.async
  // group per chunk, so the round trips to L2 overlap the tail of the previous kernel and each other.
  {
    int job = 0;
    for (int l = 0; l < a.nlayers; ++l) {
      const int nchunks = (a.L[l].K + CHAIN_KT - 1) / CHAIN_KT;
      for (int c = 0; c < nchunks; ++c, ++job) chain_load_chunk(a.L[l], c, wbuf + (size_t)job * a.wsz, tid);
    }
  }
  pdl_wait();
  // stage-0 input: [X | X2] rows of this tile, zero beyond the data (rows >= M, columns >= kx + kx2)
  {
    const int kin = a.kx + a.kx2;
#pragma unroll 4
    for (int i = tid; i < TM * kpad; i += 256) {
      const int r = i / kpad, c = i - r * kpad;
      const int gr = r_base + r;
      const bool ok = gr < a.M && c < kin;
      const float* src = (c < a.kx) ? a.X + (int64_t)gr * a.ldx + c : a.X2 + (int64_t)gr * a.ldx2 + (c - a.kx);
      in[i] = ok ? __ldg(src) : 0.f;
    }
  }
  int job = 0;                                               // flat (layer, chunk) counter: selects the weight slot
#pragma unroll 1
  for (int l = 0; l < a.nlayers; ++l) {
    const ChainLayer& L = a.L[l];
    const int NC = min(256, (L.N + 31) & ~31);          // column slots (N <= 256 is checked on the host)
    const int G = 256 / NC;                              // row groups that fit the CTA (1, 2, 4 or 8)
    const int RG = TM / G;                               // rows per thread
    const int n = tid % NC, g = tid / NC;
    const bool active = g < G && n < L.N;
    const bool second = active && n >= L.nsplit;
    const int r0 = g * RG;
    const int nchunks = (L.K + CHAIN_KT - 1) / CHAIN_KT;
    const int wcol = second ? CHAIN_KT * L.ldw + (n - L.nsplit) : n;
    // epilogue operands fetched now: their latency hides behind the products
    float hmul[8];
    float bias = 0.f;
    const int act = second ? L.actb : L.act;
    if (active) {
      const float* bp = second ? L.biasb : L.bias;
      if (bp) bias = __ldg(bp + (second ? n - L.nsplit : n));
      if (act == CH_MULH) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int gr = r_base + r0 + r;
          hmul[r] = (r < RG && gr < a.M) ? __ldg(L.H + (int64_t)gr * L.ldh + n) : 0.f;
        }
      }
    }
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c, ++job) {
      chain_wait_pending(a.njobs - 1 - job);
      __syncthreads();      // chunk `job` has landed for every thread; the previous layer's outputs are in `in`
      if (active) {
        const float* wb = wbuf + (size_t)job * a.wsz + wcol;
        const float* xin = in + (size_t)r0 * kpad + c * CHAIN_KT;
        const int ksteps = (min(CHAIN_KT, L.K - c * CHAIN_KT) + 3) >> 2;
        if (RG == 8) chain_fma_chunk<8>(wb, L.ldw, xin, kpad, ksteps, acc);
        else if (RG == 4) chain_fma_chunk<4>(wb, L.ldw, xin, kpad, ksteps, acc);
        else if (RG == 2) chain_fma_chunk<2>(wb, L.ldw, xin, kpad, ksteps, acc);
        else chain_fma_chunk<1>(wb, L.ldw, xin, kpad, ksteps, acc);
      }
    }
    // the next layer reads the columns of `nxt` up to its K rounded to 4 (<= kpad): clear what this layer does not write
    for (int i = tid; i < TM * 4; i += 256) { const int c = L.N + (i & 3); if (c < kpad) nxt[(i >> 2) * kpad + c] = 0.f; }
    if (active) {
      float* og = second ? L.outb : L.out;
      const int64_t ldo = second ? L.ldob : L.ldo;
      const int on = second ? n - L.nsplit : n;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (r < RG) {
          const int gr = r_base + r0 + r;
          float v = acc[r] + bias;
          if (act == CH_TANH) v = tanhf(v);
          else if (act == CH_EXP_EPS) v = expf(v) + 1e-6f;
          else if (act == CH_MULH) v *= (1.f - hmul[r] * hmul[r]);
          nxt[(size_t)(r0 + r) * kpad + n] = v;
          if (og && gr < a.M) og[(int64_t)gr * ldo + on] = v;
        }
      }
    }
    float* t = in; in = nxt; nxt = t;
    // no barrier here: the first chunk of the next layer starts with one
  }
}

// fills the derived fields of `a`; returns the dynamic shared memory the launch needs (0: the chain does not qualify)
static size_t chain_plan(ChainArgs& a) {
  int kmax = (a.kx + a.kx2 + 3) & ~3, wsz = 0, njobs = 0;
  for (int l = 0; l < a.nlayers; ++l) {
    ChainLayer& L = a.L[l];
    if (!(L.N <= 256 && L.N > 0 && L.K <= kmax)) return 0;
    if (!(L.ldw == (L.Wb ? L.nsplit : L.N) && (!L.Wb || L.N == 2 * L.nsplit))) return 0;     // contiguous weight rows
    // widest cp.async every chunk of this layer's weights is aligned for
    auto al = [](const void* p, int v) { return p == nullptr || ((uintptr_t)p % (4 * v)) == 0; };
    int v = 4;
    while (v > 1 && !(L.ldw % v == 0 && al(L.W, v) && al(L.Wb, v))) v >>= 1;
    L.vec = v;
    kmax = std::max(kmax, (L.N + 3) & ~3);
    wsz = std::max(wsz, CHAIN_KT * L.ldw * (L.Wb ? 2 : 1));
    njobs += (L.K + CHAIN_KT - 1) / CHAIN_KT;
  }
  a.kpad = kmax;
  a.wsz = (wsz + 3) & ~3;
  a.njobs = njobs;
  a.TM = CHAIN_TM;
  if (njobs > 16) return 0;
  return (2 * (size_t)CHAIN_TM * a.kpad + (size_t)njobs * a.wsz) * sizeof(float);
}
constexpr size_t CHAIN_SMEM_MAX = 200 * 1024;
bool dense_chain_fits(const ChainArgs& a_in) {
  ChainArgs a = a_in;
  const size_t smem = chain_plan(a);
  return smem > 0 && smem <= CHAIN_SMEM_MAX;
}
int dense_chain(const ChainArgs& a_in, cudaStream_t st) {
  ChainArgs a = a_in;
  const size_t smem = chain_plan(a);
  IWAE_REQUIRE(smem > 0 && smem <= CHAIN_SMEM_MAX, "dense_chain: the chain does not fit shared memory (%zu bytes)", smem);
  if (a.M <= 0) return 0;
  // (per device and cheap: set on every launch rather than remembered in a process-wide flag)
  CUDA_TRY(cudaFuncSetAttribute(dense_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHAIN_SMEM_MAX));
  CUDA_TRY(launch_pdl(dense_chain_kernel, dim3((a.M + CHAIN_TM - 1) / CHAIN_TM), dim3(256), smem, st, a));
  LAUNCH_CHECK();
  return 0;
}

// dst[c][r] = src[r][c] for a table of matrices (the transposed weights of the fused backward chains), one launch
__global__ void __launch_bounds__(256) transpose_jobs_kernel(TransposeJobs t) {
  pdl_wait();
  __shared__ float tile[32][33];
  const TransposeJob j = t.job[blockIdx.y];
  const int tiles_c = (j.cols + 31) / 32, tiles_r = (j.rows + 31) / 32;
  for (int tt = blockIdx.x; tt < tiles_c * tiles_r; tt += gridDim.x) {
    const int tr = tt / tiles_c, tc = tt % tiles_c;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
      const int r = tr * 32 + i, c = tc * 32 + tx;
      tile[i][tx] = (r < j.rows && c < j.cols) ? j.src[(int64_t)r * j.ld_src + c] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
      const int c = tc * 32 + i, r = tr * 32 + tx;
      if (r < j.rows && c < j.cols) j.dst[(int64_t)c * j.ld_dst + r] = tile[tx][i];
    }
  }
}
int transpose_jobs(const TransposeJobs& t, cudaStream_t st) {
  if (t.count == 0) return 0;
  CUDA_TRY(launch_pdl(transpose_jobs_kernel, dim3(24, t.count), dim3(256), 0, st, t));
  LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// Train-step megakernel.
// A k=50 step is ~60 kernels of 1-6 us on a dependency chain ~31 kernels long; even replayed from a CUDA graph with
// programmatic dependent launch the boundary between two dependent kernels (drain, launch, ramp-up) costs more than the
// kernels do.  In record mode (mega_begin .. mega_flush) the launch wrappers of this file do not launch: they append
// (kernel, grid, arguments) to a program, and each operation gets a LEVEL from the stream it was issued on and the
// event edges between the streams (lane_dep) - exactly the dependency structure the multi-stream step already spells
// out.  mega_flush uploads the program and launches ONE persistent cooperative kernel: for every level its CTAs share
// out the virtual blocks of all operations of that level (each runs the unchanged *_body of its kernel), then meet at a
// grid-wide barrier (one atomic + an acquire spin per CTA, ~1 us) instead of a kernel boundary.
// Per-step scalars come from StepDyn as before, so the captured CUDA graph of a step is [inputs, step scalars, this
// kernel, outputs] and the program is uploaded once, at capture.
// Memory ordering: every buffer of a step is written at one level and read at later ones (the arena never reuses memory
// within a step); writers fence before arriving at the barrier and readers first touch the data after their acquire.
// =============================================================================================
enum MegaType {
  MT_GEMM_64_32 = 0, MT_GEMM_32_32, MT_GEMM_64_64, MT_CONCAT_ONEHOT, MT_SAMPLE_LATENT, MT_NLP_ROWS, MT_NLP_MULTI, MT_KL_ROWS,
  MT_BERN_REDUCE, MT_LOGW, MT_SCAL_PART, MT_FINALIZE, MT_ROW_COEF, MT_DLOGITS, MT_LATENT1, MT_GAUSS_BWD, MT_LATENT2,
  MT_GRAD_REDUCE, MT_SPLITK_BIAS, MT_SPLITK_MULH, MT_ZERO, MT_COUNT
};
struct MegaOp {
  int type; unsigned gx, gy, gz;
  unsigned vb_start, nblocks; unsigned pad[2];
  alignas(16) unsigned char args[MEGA_ARG_BYTES];
};
struct MegaLevel { int op_begin, op_end; unsigned total, pad; };

// zero n floats (the split-K gradient workspace): a level-0 operation of the megakernel, a memset outside it
__device__ __forceinline__ void zero_f32_body(const uint3 vb_, const dim3 vg_, float* p, int64_t n) {
  VBLOCK_SHADOW;
  const int64_t n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
    p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0.f;
}
__global__ void __launch_bounds__(256) zero_f32_kernel(float* p, int64_t n) { pdl_wait(); zero_f32_body(blockIdx, gridDim, p, n); }

template <class... KArgs, size_t... I>
__device__ __forceinline__ void mega_apply(void (*body)(uint3, dim3, KArgs...), const void* raw, uint3 vb, dim3 vg,
                                           std::index_sequence<I...>) {
  const ArgPack<KArgs...>& p = *static_cast<const ArgPack<KArgs...>*>(raw);
  body(vb, vg, pack_get<I>(p)...);
}
template <class... KArgs>
__device__ __forceinline__ void mega_call(void (*body)(uint3, dim3, KArgs...), const void* raw, uint3 vb, dim3 vg) {
  mega_apply(body, raw, vb, vg, std::index_sequence_for<KArgs...>{});
}

__device__ __forceinline__ void mega_grid_barrier(unsigned long long* bar) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();                                                   // this CTA's writes are visible device-wide ...
    const unsigned long long t = atomicAdd(bar, 1ULL);                 // ... before its arrival is
    const unsigned long long target = (t / gridDim.x + 1ULL) * gridDim.x;
    unsigned long long seen;
    do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(bar) : "memory"); } while (seen < target);
  }
  __syncthreads();
}

__device__ __forceinline__ unsigned long long mega_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// trace (IWAE_MEGA_TRACE, diagnostics): CTA 0 stamps [2L] = its work of level L done, [2L+1] = barrier of level L passed
__global__ void __launch_bounds__(256, 2) mega_kernel(const MegaLevel* __restrict__ levels, const MegaOp* __restrict__ ops,
                                                      int nlevels, unsigned long long* bar, unsigned long long* trace) {
  __shared__ __align__(16) unsigned char s_args[MEGA_ARG_BYTES];
  int last_op = -1;
  if (trace && blockIdx.x == 0 && threadIdx.x == 0) trace[2 * nlevels] = mega_now();
  for (int L = 0; L < nlevels; ++L) {
    const int ob = levels[L].op_begin, oe = levels[L].op_end;
    const unsigned total = levels[L].total;
    for (unsigned v = blockIdx.x; v < total; v += gridDim.x) {
      int o = ob;
      while (o + 1 < oe && v >= ops[o + 1].vb_start) ++o;
      __syncthreads();                     // the previous virtual block is done with s_args and the shared-memory pool
      if (o != last_op) {
        if (threadIdx.x < MEGA_ARG_BYTES / 16)
          reinterpret_cast<uint4*>(s_args)[threadIdx.x] = reinterpret_cast<const uint4*>(ops[o].args)[threadIdx.x];
        last_op = o;
        __syncthreads();
      }
      const unsigned gx = ops[o].gx, gy = ops[o].gy, gz = ops[o].gz;
      const unsigned local = v - ops[o].vb_start;
      uint3 vb;
      vb.x = local % gx; vb.y = (local / gx) % gy; vb.z = local / (gx * gy);
      const dim3 vg(gx, gy, gz);
      const void* raw = s_args;
      switch (ops[o].type) {
        case MT_GEMM_64_32: mega_call(gemm_f32_body<64, 32, 4, 2>, raw, vb, vg); break;
        case MT_GEMM_32_32: mega_call(gemm_f32_body<32, 32, 2, 2>, raw, vb, vg); break;
        case MT_GEMM_64_64: mega_call(gemm_f32_body<64, 64, 4, 4>, raw, vb, vg); break;
        case MT_CONCAT_ONEHOT: mega_call(concat_onehot_body, raw, vb, vg); break;
        case MT_SAMPLE_LATENT: mega_call(sample_latent_body, raw, vb, vg); break;
        case MT_NLP_ROWS: mega_call(normal_logp_rows_body, raw, vb, vg); break;
        case MT_NLP_MULTI: mega_call(normal_logp_rows_multi_body, raw, vb, vg); break;
        case MT_KL_ROWS: mega_call(kl_rows_body, raw, vb, vg); break;
        case MT_BERN_REDUCE: mega_call(bern_reduce_partials_body, raw, vb, vg); break;
        case MT_LOGW: mega_call(logw_reduce_body, raw, vb, vg); break;
        case MT_SCAL_PART: mega_call(scalars_partial_body, raw, vb, vg); break;
        case MT_FINALIZE: mega_call(finalize_scalars_body, raw, vb, vg); break;
        case MT_ROW_COEF: mega_call(row_coef_body, raw, vb, vg); break;
        case MT_DLOGITS: mega_call(dlogits_body, raw, vb, vg); break;
        case MT_LATENT1: mega_call(latent1_bwd_body, raw, vb, vg); break;
        case MT_GAUSS_BWD: mega_call(gauss_rows_bwd_body, raw, vb, vg); break;
        case MT_LATENT2: mega_call(latent2_bwd_body, raw, vb, vg); break;
        case MT_GRAD_REDUCE: mega_call(grad_reduce_body, raw, vb, vg); break;
        case MT_SPLITK_BIAS: mega_call(splitk_bias_act_body, raw, vb, vg); break;
        case MT_SPLITK_MULH: mega_call(splitk_mul_h_body, raw, vb, vg); break;
        case MT_ZERO: mega_call(zero_f32_body, raw, vb, vg); break;
        default: break;
      }
    }
    if (trace && blockIdx.x == 0 && threadIdx.x == 0) trace[2 * L] = mega_now();
    if (L + 1 < nlevels) mega_grid_barrier(bar);
    if (trace && blockIdx.x == 0 && threadIdx.x == 0) trace[2 * L + 1] = mega_now();
  }
}

// ---- host side: recorder, program builder, launcher ----
struct MegaOpHost { int type; int level; dim3 grid; size_t smem; int seq; unsigned char args[MEGA_ARG_BYTES]; };
struct MegaBuf {
  unsigned char* dev = nullptr; size_t cap = 0;
  unsigned char* host = nullptr; size_t host_cap = 0;       // pinned staging
  cudaEvent_t used = nullptr;                                 // the last upload from `host` / launch reading `dev`
};
struct MegaState {
  bool recording = false, failed = false;
  std::vector<MegaOpHost> ops;
  std::vector<std::pair<cudaStream_t, int>> lvl;              // next free level of each stream seen while recording
  MegaBuf buf[2];                                             // [0] eager launches, [1] the captured train-step graph
  unsigned long long* bar = nullptr;
  unsigned long long* trace = nullptr;
  cudaStream_t upload = nullptr;
  int grid = 0;
  char why[160] = {0};
};
static thread_local MegaState* g_mega = nullptr;              // the recorder of the step this host thread is issuing

bool mega_recording() { return g_mega != nullptr; }
static int& mega_level_of(MegaState* m, cudaStream_t s) {
  for (auto& e : m->lvl) if (e.first == s) return e.second;
  m->lvl.emplace_back(s, 0);
  return m->lvl.back().second;
}
void mega_dep(cudaStream_t from, cudaStream_t to) {
  if (!g_mega || from == to) return;
  const int lf = mega_level_of(g_mega, from);
  int& lt = mega_level_of(g_mega, to);
  lt = std::max(lt, lf);
}
static int mega_type_of(const void* kern) {
  static const std::pair<const void*, int> table[] = {
    {(const void*)gemm_f32_kernel<64, 32, 4, 2>, MT_GEMM_64_32}, {(const void*)gemm_f32_kernel<32, 32, 2, 2>, MT_GEMM_32_32},
    {(const void*)gemm_f32_kernel<64, 64, 4, 4>, MT_GEMM_64_64}, {(const void*)concat_onehot_kernel, MT_CONCAT_ONEHOT},
    {(const void*)sample_latent_kernel, MT_SAMPLE_LATENT}, {(const void*)normal_logp_rows_kernel, MT_NLP_ROWS},
    {(const void*)normal_logp_rows_multi_kernel, MT_NLP_MULTI}, {(const void*)kl_rows_kernel, MT_KL_ROWS},
    {(const void*)bern_reduce_partials_kernel, MT_BERN_REDUCE}, {(const void*)logw_reduce_kernel, MT_LOGW},
    {(const void*)scalars_partial_kernel, MT_SCAL_PART}, {(const void*)finalize_scalars_kernel, MT_FINALIZE},
    {(const void*)row_coef_kernel, MT_ROW_COEF}, {(const void*)dlogits_kernel, MT_DLOGITS},
    {(const void*)latent1_bwd_kernel, MT_LATENT1}, {(const void*)gauss_rows_bwd_kernel, MT_GAUSS_BWD},
    {(const void*)latent2_bwd_kernel, MT_LATENT2}, {(const void*)grad_reduce_kernel, MT_GRAD_REDUCE},
    {(const void*)splitk_bias_act_kernel, MT_SPLITK_BIAS}, {(const void*)splitk_mul_h_kernel, MT_SPLITK_MULH},
    {(const void*)zero_f32_kernel, MT_ZERO}};
  for (const auto& e : table) if (e.first == kern) return e.second;
  return -1;
}
bool mega_record(const void* kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const void* pack, size_t pack_bytes) {
  MegaState* m = g_mega;
  const int type = mega_type_of(kern);
  if (type < 0 || block.x != 256 || block.y != 1 || block.z != 1 || pack_bytes > MEGA_ARG_BYTES || smem > MEGA_SMEM_BYTES) {
    if (!m->failed) snprintf(m->why, sizeof(m->why), "operation not supported by the megakernel (type %d, block %ux%u, %zu bytes of shared memory)",
                             type, block.x, block.y, smem);
    m->failed = true;
    return true;                 // keep recording; mega_flush reports
  }
  MegaOpHost op;
  op.type = type; op.grid = grid; op.smem = smem; op.seq = (int)m->ops.size();
  int& lv = mega_level_of(m, st);
  op.level = lv; lv += 1;
  memset(op.args, 0, sizeof(op.args));
  memcpy(op.args, pack, pack_bytes);
  m->ops.push_back(op);
  return true;
}

int zero_f32(float* p, int64_t n, cudaStream_t st) {
  if (n <= 0) return 0;
  if (!mega_recording()) { CUDA_TRY(cudaMemsetAsync(p, 0, sizeof(float) * (size_t)n, st)); return 0; }
  IWAE_REQUIRE(((uintptr_t)p & 15) == 0, "zero_f32: unaligned buffer");
  const int blocks = (int)std::min<int64_t>((n / 4 + 2047) / 2048 + 1, 592);
  CUDA_TRY(launch_pdl(zero_f32_kernel, dim3(blocks), dim3(256), 0, st, p, n));
  return 0;
}

static int mega_state(iwae_ctx* ctx, MegaState** out) {
  if (!ctx->mega) {
    MegaState* m = new MegaState();
    ctx->mega = m;
    CUDA_TRY(cudaMalloc((void**)&m->bar, sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(m->bar, 0, sizeof(unsigned long long)));
    CUDA_TRY(cudaStreamCreateWithFlags(&m->upload, cudaStreamNonBlocking));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mega_kernel, 256, MEGA_SMEM_BYTES));
    IWAE_REQUIRE(per_sm >= 1, "megakernel does not fit an SM");
    m->grid = ctx->sm_count * std::min(per_sm, 2);
  }
  *out = ctx->mega;
  return 0;
}
void mega_release(iwae_ctx* ctx) {
  MegaState* m = ctx->mega;
  if (!m) return;
  for (MegaBuf& b : m->buf) {
    if (b.dev) cudaFree(b.dev);
    if (b.host) cudaFreeHost(b.host);
    if (b.used) cudaEventDestroy(b.used);
  }
  if (m->bar) cudaFree(m->bar);
  if (m->trace) cudaFree(m->trace);
  if (m->upload) cudaStreamDestroy(m->upload);
  delete m;
  ctx->mega = nullptr;
}

int mega_begin(iwae_ctx* ctx, cudaStream_t st) {
  MegaState* m;
  IWAE_TRY(mega_state(ctx, &m));
  IWAE_REQUIRE(!g_mega, "mega_begin: a recording is already open on this thread");
  m->recording = true; m->failed = false; m->ops.clear(); m->lvl.clear(); m->why[0] = 0;
  mega_level_of(m, st) = 0;
  g_mega = m;
  return 0;
}
void mega_abort(iwae_ctx* ctx) { if (ctx->mega && g_mega == ctx->mega) { g_mega = nullptr; ctx->mega->recording = false; } }

// close the recording, build the level program, upload it and launch the megakernel on `st`
int mega_flush(iwae_ctx* ctx, cudaStream_t st) {
  MegaState* m = ctx->mega;
  IWAE_REQUIRE(m && g_mega == m, "mega_flush: no open recording");
  g_mega = nullptr; m->recording = false;
  IWAE_REQUIRE(!m->failed, "train-step megakernel: %s", m->why);
  if (m->ops.empty()) return 0;
  // order: by level; inside a level the matrix products first (the longest virtual blocks start first), then issue order
  std::vector<MegaOpHost*> ord;
  for (auto& o : m->ops) ord.push_back(&o);
  std::stable_sort(ord.begin(), ord.end(), [](const MegaOpHost* a, const MegaOpHost* b) {
    if (a->level != b->level) return a->level < b->level;
    const bool ga = a->type <= MT_GEMM_64_64, gb = b->type <= MT_GEMM_64_64;
    if (ga != gb) return ga;
    return a->seq < b->seq;
  });
  const int nlevels = ord.back()->level + 1;
  const size_t lv_bytes = ((sizeof(MegaLevel) * (size_t)nlevels + 255) / 256) * 256;
  const size_t bytes = lv_bytes + sizeof(MegaOp) * ord.size();
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  CUDA_TRY(cudaStreamIsCapturing(st, &cs));
  const bool capturing = cs != cudaStreamCaptureStatusNone;
  MegaBuf& b = m->buf[capturing ? 1 : 0];
  if (b.cap < bytes || b.host_cap < bytes) {
    IWAE_REQUIRE(!capturing, "train-step megakernel: program buffer too small during capture (%zu bytes)", bytes);
    if (b.used) CUDA_TRY(cudaEventSynchronize(b.used));
    if (b.dev) CUDA_TRY(cudaFree(b.dev));
    if (b.host) CUDA_TRY(cudaFreeHost(b.host));
    const size_t cap = bytes * 2;
    CUDA_TRY(cudaMalloc((void**)&b.dev, cap)); b.cap = cap;
    CUDA_TRY(cudaMallocHost((void**)&b.host, cap)); b.host_cap = cap;
    // the captured graph's buffer is sized with the eager one (an eager step always precedes a capture)
    MegaBuf& g = m->buf[1];
    if (&g != &b && g.cap < cap) {
      if (g.dev) CUDA_TRY(cudaFree(g.dev));
      if (g.host) CUDA_TRY(cudaFreeHost(g.host));
      CUDA_TRY(cudaMalloc((void**)&g.dev, cap)); g.cap = cap;
      CUDA_TRY(cudaMallocHost((void**)&g.host, cap)); g.host_cap = cap;
    }
  }
  if (!b.used) {
    IWAE_REQUIRE(!capturing, "train-step megakernel: first use during capture");
    CUDA_TRY(cudaEventCreateWithFlags(&b.used, cudaEventDisableTiming));
    MegaBuf& g = m->buf[1];
    if (!g.used) CUDA_TRY(cudaEventCreateWithFlags(&g.used, cudaEventDisableTiming));
  } else if (!capturing) {
    CUDA_TRY(cudaEventSynchronize(b.used));        // the previous upload has left the staging buffer
  }
  MegaLevel* lv = reinterpret_cast<MegaLevel*>(b.host);
  MegaOp* ops = reinterpret_cast<MegaOp*>(b.host + lv_bytes);
  memset(lv, 0, lv_bytes);
  size_t smem = 0;
  int cur = -1; unsigned vb = 0;
  for (size_t i = 0; i < ord.size(); ++i) {
    const MegaOpHost& h = *ord[i];
    if (h.level != cur) {
      // levels without operations cannot occur: a level is only created by an operation
      cur = h.level; lv[cur].op_begin = (int)i; vb = 0;
    }
    MegaOp& o = ops[i];
    o.type = h.type; o.gx = h.grid.x; o.gy = h.grid.y; o.gz = h.grid.z;
    o.nblocks = h.grid.x * h.grid.y * h.grid.z; o.vb_start = vb; o.pad[0] = o.pad[1] = 0;
    memcpy(o.args, h.args, MEGA_ARG_BYTES);
    vb += o.nblocks;
    lv[cur].op_end = (int)i + 1; lv[cur].total = vb;
    smem = std::max(smem, h.smem);
  }
  for (int L = 0; L < nlevels; ++L) IWAE_REQUIRE(lv[L].op_end > lv[L].op_begin, "train-step megakernel: empty level %d", L);
  if (capturing) {
    // not a node of the graph: the program is constant for the life of the captured step
    CUDA_TRY(cudaMemcpyAsync(b.dev, b.host, bytes, cudaMemcpyHostToDevice, m->upload));
    CUDA_TRY(cudaStreamSynchronize(m->upload));
  } else {
    CUDA_TRY(cudaMemcpyAsync(b.dev, b.host, bytes, cudaMemcpyHostToDevice, st));
  }
  const MegaLevel* d_lv = reinterpret_cast<const MegaLevel*>(b.dev);
  const MegaOp* d_ops = reinterpret_cast<const MegaOp*>(b.dev + lv_bytes);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(m->grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = MEGA_SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  static const char* trace_path = getenv("IWAE_MEGA_TRACE");
  unsigned long long* trace = nullptr;
  if (trace_path && !capturing) {
    if (!m->trace) CUDA_TRY(cudaMalloc((void**)&m->trace, sizeof(unsigned long long) * 2 * 257));
    IWAE_REQUIRE(nlevels <= 256, "trace: too many levels");
    trace = m->trace;
  }
  CUDA_TRY(cudaLaunchKernelEx(&cfg, mega_kernel, d_lv, d_ops, nlevels, m->bar, trace));
  LAUNCH_CHECK();
  if (!capturing) CUDA_TRY(cudaEventRecord(b.used, st));
  if (trace) {
    CUDA_TRY(cudaStreamSynchronize(st));
    std::vector<unsigned long long> t(2 * nlevels + 1);
    CUDA_TRY(cudaMemcpy(t.data(), trace, sizeof(unsigned long long) * t.size(), cudaMemcpyDeviceToHost));
    static const char* names[MT_COUNT] = {"gemm64x32", "gemm32x32", "gemm64x64", "concat_onehot", "sample_latent", "nlp_rows", "nlp_multi",
      "kl_rows", "bern_reduce", "logw", "scal_part", "finalize", "row_coef", "dlogits", "latent1", "gauss_bwd", "latent2", "grad_reduce",
      "splitk_bias", "splitk_mulh", "zero"};
    FILE* fp = fopen(trace_path, "w");
    if (fp) {
      unsigned long long prev = t[2 * nlevels];
      fprintf(fp, "# level  work_us(CTA 0)  barrier_us  ops (type:virtual blocks)   grid=%d CTAs\n", m->grid);
      for (int L = 0; L < nlevels; ++L) {
        fprintf(fp, "%3d  %7.2f  %7.2f ", L, (t[2 * L] - prev) * 1e-3, (t[2 * L + 1] - t[2 * L]) * 1e-3);
        for (int i = lv[L].op_begin; i < lv[L].op_end; ++i) fprintf(fp, " %s:%u", names[ops[i].type], ops[i].nblocks);
        fprintf(fp, "\n");
        prev = t[2 * L + 1];
      }
      fprintf(fp, "# total %.2f us\n", (t[2 * nlevels - 1] - t[2 * nlevels]) * 1e-3);
      fclose(fp);
    }
  }
  return 0;
}
