// Encoder BasicBlock q(z1|x) on the tensor cores: x -> tanh -> tanh -> (mu | exp(.)+1e-6)   (src/iwae1.py:24-44, 47-60)
//
// The per-image pre-pass of the fused evaluation (eval_tc.cu) and of val_step / IWAE.call on the fused engine: one CTA owns
// a tile of 128 IMAGES and runs the three Dense layers of the block back to back with tcgen05.mma (cta_group::1, M = 128,
// N = the padded layer width <= 208 in ONE instruction per K step), accumulators in TMEM, bias + tanh / exp fused into the
// epilogue, activations handed from layer to layer through shared memory - nothing but mu and std reaches HBM.
//
// Precision: the outputs feed every sample's z1 = mu + std*eps, so they must be fp32-accurate.  Every operand is an fp16
// hi + lo PAIR, hi = rn(v) and lo = rn((v - hi) * 2048): the factor 2^11 keeps the lo halves of small weights out of the
// fp16 subnormals (22 bits in all).  A product is three MMAs into TWO fp32 accumulators in TMEM, D1 += hi*hi and
// D2 += lo*hi + hi*lo (the lo*lo term is 2^-22 relative), and the epilogue adds them as D1 + D2 / 2048.  The first layer's
// input is the binarised image (or image | one-hot label): exactly representable, so it needs two MMAs per K step.
// Measured against the fp64 oracle (tests/test_gpu_eval.py::test_tensor_core_encoder_matches_fp32_kernels): mu and std as
// close as the fp32 FFMA kernels' (both 0.2 - 0.7e-6 of the tensor's scale).
//
// Data movement: the first layer streams K (784 pixels) in slabs of 32: all 128 threads convert their image row's 32
// pixels to fp16 into a double-buffered A slab while one thread fetches the matching hi/lo weight slab with cp.async.bulk
// (weights are pre-packed per slab in the exact shared-memory image, cached against the parameter version like the fused
// kernel's).  Layers 2 and 3 read the resident activations and stream only weights.  Shared memory: 2 x 8 KB A slabs +
// 2 x 26 KB weight slabs + 2 x 52 KB activations (hi, lo) = 172 KB, one CTA per SM; 79 tiles for 10,000 images.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace tc {

constexpr int ENC_M = 128;          // images per tile
constexpr int ENC_SK = 32;          // K rows per slab
constexpr int ENC_NMAX = 208;       // widest layer (padded to 16)
constexpr int ENC_ABYTES = ENC_M * ENC_SK * 2;                 // 8192: one A slab (fp16)
constexpr int ENC_WBYTES = 2 * ENC_SK * ENC_NMAX * 2;          // 26624: one weight slab, hi block then lo block
constexpr int ENC_HBYTES = ENC_M * ENC_NMAX * 2;               // 53248: activations, one of (hi, lo)
constexpr int ENC_TMEM_COLS = 512;                             // D1 at columns 0.., D2 at columns 256.. (one CTA per SM)
constexpr float ENC_LO_SCALE = 2048.f;                         // lo halves are stored times 2^11

struct EncLayer {
  int k_real, k_pad, n_real, n_pad;     // k_pad, n_pad: multiples of 16
  int nslabs;
  int64_t wofs;                         // byte offset of the layer's first slab in the packed buffer
  const float* bias;                    // [n_real] (layer 3: see bias2)
};
struct EncArgs {
  EncLayer L[3];
  const float* bias2;                   // log-std head bias (layer 3 columns d..2d-1)
  const uint8_t* wpack;
  const float* x; int64_t ldx;          // [N, k_real of layer 1]
  int64_t N;
  int d;                                // latent dims
  float* mu; float* sd;                 // [N, d]
};

struct EncSmem {
  uint8_t a[2][ENC_ABYTES];
  uint8_t w[2][ENC_WBYTES];
  uint8_t h_hi[ENC_HBYTES];
  uint8_t h_lo[ENC_HBYTES];
  uint64_t full[2], mma_done[2], acc_full;
  uint32_t tmem_base;
};

// bytes of slab s of a layer: hi and lo block of `rows` K rows each
__host__ __device__ inline int enc_slab_rows(const EncLayer& L, int s) { return (L.k_pad - s * ENC_SK) < ENC_SK ? (L.k_pad - s * ENC_SK) : ENC_SK; }

__global__ void __launch_bounds__(128, 1) encoder_tc_kernel(const __grid_constant__ EncArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  EncSmem& S = *reinterpret_cast<EncSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&S.full[i], 1); mbar_init(&S.mma_done[i], 1); }
    mbar_init(&S.acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(&S.tmem_base, ENC_TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = S.tmem_base;
  const int64_t row_g = (int64_t)blockIdx.x * ENC_M + tid;       // this thread's image
  const bool row_ok = row_g < a.N;
  const float* xrow = a.x + (row_ok ? row_g : (a.N - 1)) * a.ldx;
  const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);

  uint32_t it = 0;                   // running slab counter over all three layers: buffer it & 1, use number it >> 1
  for (int l = 0; l < 3; ++l) {
    const EncLayer& L = a.L[l];
    const uint32_t idesc = make_idesc((uint32_t)L.n_pad);
    const uint8_t* wsrc = a.wpack + L.wofs;
    for (int s = 0; s < L.nslabs; ++s, ++it) {
      const uint32_t buf = it & 1u, use = it >> 1;
      const int rows = enc_slab_rows(L, s);
      const uint32_t wbytes = 2u * (uint32_t)rows * (uint32_t)L.n_pad * 2u;
      // the buffer pair is free once the MMAs of its previous use have retired
      if (use > 0) mbar_wait(&S.mma_done[buf], (use - 1u) & 1u, 700);
      if (tid == 0) {
        mbar_expect_tx(&S.full[buf], wbytes);
        bulk_g2s(S.w[buf], wsrc, wbytes, &S.full[buf]);
      }
      wsrc += wbytes;
      if (l == 0) {
        // this image's pixels s*32 .. s*32+rows-1 -> fp16 (exact: x is 0/1) -> A slab, K-major core matrices
        const int k0 = s * ENC_SK;
#pragma unroll
        for (int kg = 0; kg < ENC_SK / 8; ++kg) {
          if (kg * 8 < rows) {
            float v[8];
            const int kb = k0 + kg * 8;
            if ((a.ldx & 3) == 0 && kb + 8 <= L.k_real) {
              const float4 p0 = __ldg(reinterpret_cast<const float4*>(xrow + kb)), p1 = __ldg(reinterpret_cast<const float4*>(xrow + kb + 4));
              v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = (kb + i < L.k_real) ? __ldg(xrow + kb + i) : 0.f;
            }
            const uint4 pk = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
            *reinterpret_cast<uint4*>(S.a[buf] + (size_t)kg * (ENC_M * 16) + tid * 16) = pk;
          }
        }
        fence_proxy_async();
      }
      __syncthreads();               // A slab complete (layer 1); every thread is past its wait on this buffer
      if (tid == 0) {
        mbar_wait(&S.full[buf], use & 1u, 701);
        tc_fence_after();
        const uint32_t wb = smem_u32(S.w[buf]);
        const uint32_t lo_ofs = (uint32_t)rows * (uint32_t)L.n_pad * 2u;          // the lo block follows the hi block
        const uint32_t lbo_w = (uint32_t)L.n_pad * 16u;
        for (int ks = 0; ks < rows / 16; ++ks) {
          const uint64_t bhi = make_smem_desc(wb + ks * 2 * lbo_w, lbo_w, 128);
          const uint64_t blo = make_smem_desc(wb + lo_ofs + ks * 2 * lbo_w, lbo_w, 128);
          const uint32_t first = (s == 0 && ks == 0) ? 0u : 1u;
          const uint32_t d1 = tmem_base, d2 = tmem_base + 256u;        // hi*hi | the two cross terms (times 2^11)
          if (l == 0) {
            const uint64_t ad = make_smem_desc(smem_u32(S.a[buf]) + ks * 2 * (ENC_M * 16), ENC_M * 16, 128);
            umma_f16(d1, ad, bhi, idesc, first);
            umma_f16(d2, ad, blo, idesc, first);
          } else {
            const uint32_t kofs = (uint32_t)(s * ENC_SK / 8 + ks * 2) * (ENC_M * 16);
            const uint64_t ahi = make_smem_desc(smem_u32(S.h_hi) + kofs, ENC_M * 16, 128);
            const uint64_t alo = make_smem_desc(smem_u32(S.h_lo) + kofs, ENC_M * 16, 128);
            umma_f16(d1, ahi, bhi, idesc, first);
            umma_f16(d2, alo, bhi, idesc, first);
            umma_f16(d2, ahi, blo, idesc, 1u);
          }
        }
        umma_commit(&S.mma_done[buf]);
        if (s == L.nslabs - 1) umma_commit(&S.acc_full);
      }
    }
    // ---- epilogue of layer l: all of its MMAs have retired, so the activations may be overwritten in place
    mbar_wait(&S.acc_full, (uint32_t)l & 1u, 702);
    tc_fence_after();
    for (int g = 0; g < L.n_pad; g += 16) {
      uint32_t r[16], r2[16];
      tmem_ld16(taddr + g, r);
      tmem_ld16(taddr + 256u + g, r2);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(fmaf(__uint_as_float(r2[i]), 1.f / ENC_LO_SCALE, __uint_as_float(r[i])));
      if (l < 2) {
        float hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = g + i;
          const float v = (c < L.n_real) ? tanhf(__uint_as_float(r[i]) + __ldg(L.bias + c)) : 0.f;
          const float h = __half2float(__float2half_rn(v));
          hi[i] = h; lo[i] = (v - h) * ENC_LO_SCALE;
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const size_t o = (size_t)((g >> 3) + q) * (ENC_M * 16) + tid * 16;
          *reinterpret_cast<uint4*>(S.h_hi + o) = make_uint4(pack_h2(hi[8 * q], hi[8 * q + 1]), pack_h2(hi[8 * q + 2], hi[8 * q + 3]),
                                                             pack_h2(hi[8 * q + 4], hi[8 * q + 5]), pack_h2(hi[8 * q + 6], hi[8 * q + 7]));
          *reinterpret_cast<uint4*>(S.h_lo + o) = make_uint4(pack_h2(lo[8 * q], lo[8 * q + 1]), pack_h2(lo[8 * q + 2], lo[8 * q + 3]),
                                                             pack_h2(lo[8 * q + 4], lo[8 * q + 5]), pack_h2(lo[8 * q + 6], lo[8 * q + 7]));
        }
      } else if (row_ok) {
        // heads: columns [0,d) = mu, [d,2d) = log-std pre-activation -> std = exp(.) + 1e-6 (iwae1.py:41-43)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = g + i;
          const float v = __uint_as_float(r[i]);
          if (c < a.d) a.mu[row_g * a.d + c] = v + __ldg(L.bias + c);
          else if (c < 2 * a.d) a.sd[row_g * a.d + (c - a.d)] = expf(v + __ldg(a.bias2 + (c - a.d))) + 1e-6f;
        }
      }
    }
    // the next layer's MMAs read these activations through the async proxy and overwrite the accumulator
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem_base, ENC_TMEM_COLS);
}

// Weights of one layer -> hi / lo fp16 slabs in the shared-memory image: slab s = [hi block | lo block], each
// [rows/8 K groups][n_pad columns][8 K] halves.  W2 (optional) continues the columns after n1 (the two heads of layer 3).
__global__ void encoder_pack_kernel(const float* __restrict__ W, int ldw, int n1, const float* __restrict__ W2, int ldw2, EncLayer L,
                                    __half* __restrict__ out) {
  const int total = L.k_pad * L.n_pad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i / L.n_pad, n = i % L.n_pad;
    const int s = k / ENC_SK, kin = k % ENC_SK;
    const int rows = enc_slab_rows(L, s);
    float v = 0.f;
    if (k < L.k_real && n < L.n_real) v = (n < n1) ? W[(int64_t)k * ldw + n] : W2[(int64_t)k * ldw2 + (n - n1)];
    const float hi = __half2float(__float2half_rn(v));
    const int64_t slab_base = (int64_t)s * 2 * ENC_SK * L.n_pad;                  // halves; full slabs precede
    const int64_t pos = ((int64_t)(kin >> 3) * L.n_pad + n) * 8 + (kin & 7);
    out[slab_base + pos] = __float2half_rn(hi);
    out[slab_base + (int64_t)rows * L.n_pad + pos] = __float2half_rn((v - hi) * ENC_LO_SCALE);
  }
}

}  // namespace tc

struct EncPack {
  uint8_t* wpack = nullptr; size_t bytes = 0;
  const float* packed_params = nullptr; uint64_t packed_version = 0; int64_t packed_w = -1;
  tc::EncLayer L[3];
  bool attr_set = false;
};

void tc_encoder_release(iwae_ctx* ctx) {
  if (!ctx->enc) return;
  if (ctx->enc->wpack) cudaFree(ctx->enc->wpack);
  delete ctx->enc;
  ctx->enc = nullptr;
}

static int r16(int v) { return (v + 15) & ~15; }

// mu, std [M, d] of the BasicBlock `blk` applied to X [M, ldx].  Returns 0 when done, -1 when the block does not fit the
// kernel (wider than 208 units / 104 latent dims) or the path is switched off (IWAE_TC_ENCODER=0): the caller then runs
// the fp32 kernels (block_fwd_f32).
int tc_encoder_fwd(iwae_ctx* ctx, const float* params, const BlockP& blk, const float* X, int64_t ldx, int64_t M, float* mu, float* sd,
                   cudaStream_t st) {
  static const bool off = getenv("IWAE_TC_ENCODER") && atoi(getenv("IWAE_TC_ENCODER")) == 0;
  const int h = blk.l1.n_out, d = blk.lmu.n_out;
  if (off || M <= 0 || h > tc::ENC_NMAX || blk.l2.n_out != h || blk.l2.n_in != h || 2 * d > tc::ENC_NMAX || blk.lstd.n_out != d) return -1;
  if (!ctx->enc) ctx->enc = new EncPack();
  EncPack& ep = *ctx->enc;
  const bool cached = ctx->params_version != 0 && ep.wpack && ep.packed_params == params && ep.packed_version == ctx->params_version &&
                      ep.packed_w == blk.l1.w;
  if (!cached) {
    const DenseP* dl[3] = {&blk.l1, &blk.l2, &blk.lmu};
    int64_t ofs = 0;
    for (int l = 0; l < 3; ++l) {
      tc::EncLayer& L = ep.L[l];
      L.k_real = dl[l]->n_in; L.k_pad = r16(L.k_real);
      L.n_real = (l == 2) ? 2 * d : dl[l]->n_out; L.n_pad = r16(L.n_real);
      L.nslabs = (L.k_pad + tc::ENC_SK - 1) / tc::ENC_SK;
      L.wofs = ofs; L.bias = params + dl[l]->b;
      ofs += (int64_t)2 * L.k_pad * L.n_pad * 2;
    }
    if (ep.bytes < (size_t)ofs) {
      if (ep.wpack) CUDA_TRY(cudaFree(ep.wpack));
      ep.wpack = nullptr; ep.bytes = 0;
      CUDA_TRY(cudaMalloc((void**)&ep.wpack, (size_t)ofs));
      ep.bytes = (size_t)ofs;
    }
    for (int l = 0; l < 3; ++l) {
      const float* W = params + dl[l]->w;
      tc::encoder_pack_kernel<<<64, 256, 0, st>>>(W, dl[l]->n_out, dl[l]->n_out, l == 2 ? params + blk.lstd.w : nullptr, d, ep.L[l],
                                                  reinterpret_cast<__half*>(ep.wpack + ep.L[l].wofs));
      LAUNCH_CHECK();
    }
    ep.packed_params = params; ep.packed_version = ctx->params_version; ep.packed_w = blk.l1.w;
  }
  tc::EncArgs a;
  for (int l = 0; l < 3; ++l) { a.L[l] = ep.L[l]; }
  a.L[0].bias = params + blk.l1.b; a.L[1].bias = params + blk.l2.b; a.L[2].bias = params + blk.lmu.b;
  a.bias2 = params + blk.lstd.b;
  a.wpack = ep.wpack; a.x = X; a.ldx = ldx; a.N = M; a.d = d; a.mu = mu; a.sd = sd;
  const size_t smem = sizeof(tc::EncSmem);
  if (!ep.attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(tc::encoder_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ep.attr_set = true;
  }
  tc::encoder_tc_kernel<<<(int)((M + tc::ENC_M - 1) / tc::ENC_M), 128, smem, st>>>(a);
  LAUNCH_CHECK();
  return 0;
}
