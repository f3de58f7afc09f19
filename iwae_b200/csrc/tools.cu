// Distribution shims and posterior tools around the hot path (SURVEY.md §8b "sub-modules", §8f rank 3): what the
// reference's scripts ask of the tfd objects a model returns -
//   pxz.log_prob(x)              tasks/plot_task01.py:63      -> iwae_bernoulli_log_prob
//   qzx.log_prob(z), .sample(n)  tasks/plot_task01.py:72, src/iwae1.py:59 -> iwae_normal_log_prob, iwae_normal_sample
//   Bernoulli(logits).sample()   src/iwae1.py:172-176         -> iwae_bernoulli_sample_logits
//   the 200x200 true-posterior grid, tasks/plot_task01.py:46-73 -> iwae_posterior_grid
// - as kernels, so that the Python mirror holds no eager tensor arithmetic.  Log-densities follow TFP's op order
// (SURVEY.md Appendix A.2, A.3).
#include "common.cuh"
#include "rng.cuh"
#include <math.h>

#define LOG2PI_HALF_T 0.91893853320467274178f

namespace {

__device__ __forceinline__ float softplus_f(float v) { return fmaxf(v, 0.f) + log1pf(expf(-fabsf(v))); }

// tfd.Bernoulli(logits).log_prob(x) = -(1-x) softplus(l) - x softplus(-l), elementwise; x row = r % B (A.3)
__global__ void bernoulli_logp_elem_kernel(const float* __restrict__ logits, const float* __restrict__ x, int64_t R, int B, int n,
                                           float* __restrict__ out) {
  const int64_t total = R * n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / n;
    const int j = (int)(i - r * n);
    const float l = logits[i], xv = x[(r % B) * n + j];
    out[i] = -(1.f - xv) * softplus_f(l) - xv * softplus_f(-l);
  }
}

// tfd.Normal(loc, scale).log_prob(z), elementwise; parameter row = r % Rm; loc == null: N(0,1)  (A.2)
__global__ void normal_logp_elem_kernel(const float* __restrict__ z, const float* __restrict__ loc, const float* __restrict__ scale,
                                        int64_t R, int Rm, int d, float* __restrict__ out) {
  const int64_t total = R * d;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / d;
    const int j = (int)(i - r * d);
    const float zv = z[i];
    if (!loc) { out[i] = -0.5f * zv * zv - LOG2PI_HALF_T; continue; }
    const float m = loc[(r % Rm) * d + j], s = scale[(r % Rm) * d + j];
    const float t = zv / s - m / s;
    out[i] = -0.5f * t * t - (LOG2PI_HALF_T + logf(s));
  }
}

// probs = sigmoid(logits), x = (u < probs) with the Philox uniforms of iwae_bernoulli_sample (element i: word i%4 of
// counter (i/4 lo, i/4 hi, epoch, 0x42))
__global__ void bernoulli_sample_logits_kernel(const float* __restrict__ logits, int64_t n, uint64_t seed, uint64_t epoch,
                                               float* __restrict__ x_out, float* __restrict__ p_out) {
  const int64_t nq = (n + 3) / 4;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)epoch, 0x42u),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t i = 4 * q + j;
      if (i < n) {
        const float p = 1.f / (1.f + expf(-logits[i]));
        const float u = (float)(rr[j] >> 8) * 5.9604644775390625e-8f;
        if (p_out) p_out[i] = p;
        x_out[i] = (u < p) ? 1.f : 0.f;
      }
    }
  }
}

// grid of plot_task01.py:21-30 (get_grid): np.meshgrid(linspace(x0,x1,n1), linspace(y0,y1,n2)), row i = (x[i % n1], y[i / n1])
__global__ void make_grid_kernel(double x0, double x1, double y0, double y1, int n1, int n2, float* __restrict__ grid) {
  const int total = n1 * n2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ix = i % n1, iy = i / n1;
    const double sx = n1 > 1 ? (x1 - x0) / (double)(n1 - 1) : 0.0, sy = n2 > 1 ? (y1 - y0) / (double)(n2 - 1) : 0.0;
    grid[2 * i] = (float)(x0 + ix * sx);
    grid[2 * i + 1] = (float)(y0 + iy * sy);
  }
}

// unnorm[i] = lpxz[i] + sum_j N(0,1).log_prob(z_ij)   (plot_task01.py:60-66)   lq[i] = sum_j N(mu,sd).log_prob(z_ij) (:72)
__global__ void posterior_terms_kernel(const float* __restrict__ grid, const float* lpxz, const float* __restrict__ mu,
                                       const float* __restrict__ sd, int n, int d, float* unnorm, float* __restrict__ lq) {   // unnorm may alias lpxz
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float lp = 0.f, q = 0.f;
    for (int j = 0; j < d; ++j) {
      const float zv = grid[(int64_t)i * d + j];
      lp += -0.5f * zv * zv - LOG2PI_HALF_T;
      const float t = zv / sd[j] - mu[j] / sd[j];
      q += -0.5f * t * t - (LOG2PI_HALF_T + logf(sd[j]));
    }
    unnorm[i] = lpxz[i] + lp;
    lq[i] = q;
  }
}

__device__ __forceinline__ float block_reduce_t(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

// x - logsumexp(x) in place (scipy.special.logsumexp normalisation, plot_task01.py:68-70); one CTA
__global__ void __launch_bounds__(256) log_softmax_vec_kernel(float* __restrict__ x, int L) {
  __shared__ float red[8];
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < L; i += blockDim.x) mx = fmaxf(mx, x[i]);
  mx = block_reduce_t(mx, red, true);
  float se = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) se += expf(x[i] - mx);
  se = block_reduce_t(se, red, false);
  const float lse = logf(se) + mx;
  for (int i = threadIdx.x; i < L; i += blockDim.x) x[i] -= lse;
}

int nblocks(int64_t total) { return (int)std::min<int64_t>((total + 255) / 256, 148 * 16); }

}  // namespace

extern "C" int iwae_bernoulli_log_prob(const float* logits, const float* x, int64_t R, int32_t B, int32_t n, float* out,
                                       void* stream) {
  IWAE_REQUIRE(logits && x && out, "iwae_bernoulli_log_prob: null argument");
  IWAE_REQUIRE(R >= 0 && B > 0 && n > 0, "iwae_bernoulli_log_prob: bad shape");
  if (R == 0) return 0;
  bernoulli_logp_elem_kernel<<<nblocks(R * n), 256, 0, (cudaStream_t)stream>>>(logits, x, R, B, n, out);
  LAUNCH_CHECK();
  return 0;
}

extern "C" int iwae_normal_log_prob(const float* z, const float* loc, const float* scale, int64_t R, int32_t Rm, int32_t d,
                                    float* out, void* stream) {
  IWAE_REQUIRE(z && out && ((loc == nullptr) == (scale == nullptr)), "iwae_normal_log_prob: bad argument");
  IWAE_REQUIRE(R >= 0 && Rm > 0 && d > 0, "iwae_normal_log_prob: bad shape");
  if (R == 0) return 0;
  normal_logp_elem_kernel<<<nblocks(R * d), 256, 0, (cudaStream_t)stream>>>(z, loc, scale, R, Rm, d, out);
  LAUNCH_CHECK();
  return 0;
}

extern "C" int iwae_normal_sample(const float* loc, const float* scale, const float* eps_in, int32_t Rm, int32_t n, int32_t d,
                                  uint64_t seed, int32_t stream_id, float* out, void* stream) {
  IWAE_REQUIRE(loc && scale && out && Rm > 0 && n > 0 && d > 0, "iwae_normal_sample: bad argument");
  // eps [n, Rm, d]: injected, or from the Philox generator (counter: sample, dim-octet, row, stream id);
  // z = eps * scale + loc (A.2)
  return sample_latent(loc, scale, 1, eps_in, nullptr, out, Rm, n, d, seed, stream_id, 0, (cudaStream_t)stream);
}

extern "C" int iwae_bernoulli_sample_logits(const float* logits, int64_t n, uint64_t seed, uint64_t epoch, float* x_out,
                                            float* probs_out, void* stream) {
  IWAE_REQUIRE(logits && x_out && n >= 0, "iwae_bernoulli_sample_logits: bad argument");
  if (n == 0) return 0;
  bernoulli_sample_logits_kernel<<<nblocks((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(logits, n, seed, epoch, x_out, probs_out);
  LAUNCH_CHECK();
  return 0;
}

extern "C" int iwae_posterior_grid(iwae_ctx* ctx, const float* params, const float* x, double x0, double x1, double y0,
                                   double y1, int32_t n1, int32_t n2, float* out_log_post, float* out_log_q, float* out_grid,
                                   void* stream) {
  IWAE_REQUIRE(ctx && params && x && out_log_post && out_log_q && out_grid, "iwae_posterior_grid: null argument");
  IWAE_REQUIRE(n1 > 0 && n2 > 0, "iwae_posterior_grid: bad grid");
  const ModelP& m = ctx->model;
  IWAE_REQUIRE(m.cfg.d1 == 2 && !m.cond() && !m.two_layer(), "iwae_posterior_grid: needs an unconditional one-layer model with a 2-D latent (tasks/task01.py:84)");
  cudaStream_t st = (cudaStream_t)stream;
  const int n = n1 * n2;
  make_grid_kernel<<<nblocks(n), 256, 0, st>>>(x0, x1, y0, y1, n1, n2, out_grid);
  LAUNCH_CHECK();
  // log p(x | z) on the grid: the decoder with the fused Bernoulli log-likelihood (same entry point as model.decoder)
  IWAE_TRY(iwae_decode(ctx, params, out_grid, x, nullptr, n, 1, nullptr, out_log_post, st));
  // q(z | x): the encoder block on the one image (iwae_decode has reset the arena; take scratch after it)
  float *h1, *h2, *mu, *sd;
  Arena& ar = ctx->arena;
  IWAE_TRY(ar.get(&h1, (size_t)m.enc1.l1.n_out)); IWAE_TRY(ar.get(&h2, (size_t)m.enc1.l2.n_out));
  IWAE_TRY(ar.get(&mu, 2)); IWAE_TRY(ar.get(&sd, 2));
  IWAE_TRY(block_fwd_f32(params, m.enc1, x, m.cfg.n_x, 1, h1, h2, mu, sd, st));
  posterior_terms_kernel<<<nblocks(n), 256, 0, st>>>(out_grid, out_log_post, mu, sd, n, 2, out_log_post, out_log_q);
  LAUNCH_CHECK();
  log_softmax_vec_kernel<<<1, 256, 0, st>>>(out_log_post, n);
  LAUNCH_CHECK();
  return 0;
}
