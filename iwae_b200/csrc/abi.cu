// C-ABI entry points of libiwae_b200.so and the host-side drivers that sequence the kernels.
// See include/iwae_b200.h for the contract and the reference interfaces each entry point replaces.
#include "common.cuh"
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void iwae_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* iwae_last_error(void) { return g_err; }
extern "C" int iwae_version(void) { return 100; }
long long g_iwae_launches = 0;
extern "C" int64_t iwae_launch_count(void) { return (int64_t)g_iwae_launches; }

// ---------------------------------------------------------------------------------------------
// arena
// ---------------------------------------------------------------------------------------------
int Arena::alloc(void** out, size_t bytes) {
  bytes = (bytes + 255) & ~(size_t)255;
  if (bytes == 0) bytes = 256;
  for (auto& b : blks) {
    if (b.cap - b.off >= bytes) {
      *out = b.p + b.off;
      b.off += bytes;
      return 0;
    }
  }
  const size_t cap = std::max(bytes, (size_t)64 << 20);
  char* p = nullptr;
  CUDA_TRY(cudaMalloc((void**)&p, cap));
  blks.push_back({p, cap, bytes});
  *out = p;
  return 0;
}
void Arena::release() {
  for (auto& b : blks) cudaFree(b.p);
  blks.clear();
}

// ---------------------------------------------------------------------------------------------
// model layout (Keras trainable_weights order)
// ---------------------------------------------------------------------------------------------
static void add_dense(ModelP& m, DenseP& d, int n_in, int n_out) {
  d.n_in = n_in; d.n_out = n_out;
  d.w = m.n_params; m.t_off.push_back(d.w); m.t_rows.push_back(n_in); m.t_cols.push_back(n_out);
  m.n_params += (int64_t)n_in * n_out;
  d.b = m.n_params; m.t_off.push_back(d.b); m.t_rows.push_back(n_out); m.t_cols.push_back(0);
  m.n_params += n_out;
}
static void add_block(ModelP& m, BlockP& b, int n_in, int h, int d) {
  add_dense(m, b.l1, n_in, h); add_dense(m, b.l2, h, h); add_dense(m, b.lmu, h, d); add_dense(m, b.lstd, h, d);
}
static void add_mlp3(ModelP& m, Mlp3P& p, int n_in, int h, int n_out) {
  add_dense(m, p.d0, n_in, h); add_dense(m, p.d1, h, h); add_dense(m, p.d2, h, n_out);
}

static int build_model(const iwae_config& c, ModelP& m) {
  m = ModelP();
  m.cfg = c;
  IWAE_REQUIRE(c.n_x > 0 && c.h1 > 0 && c.d1 > 0, "iwae_create: n_x, h1, d1 must be positive");
  switch (c.kind) {
    case IWAE_KIND_IWAE1:
      IWAE_REQUIRE(c.n_cond == 0, "iwae_create: IWAE1 takes n_cond == 0");
      add_block(m, m.enc1, c.n_x, c.h1, c.d1);
      add_mlp3(m, m.dec, c.d1, c.h1, c.n_x);
      break;
    case IWAE_KIND_IWAE2:
      IWAE_REQUIRE(c.n_cond == 0 && c.h2 > 0 && c.d2 > 0, "iwae_create: IWAE2 needs h2, d2 > 0 and n_cond == 0");
      add_block(m, m.enc1, c.n_x, c.h1, c.d1);
      add_block(m, m.enc2, c.d1, c.h2, c.d2);
      add_block(m, m.dec21, c.d2, c.h2, c.d1);
      add_mlp3(m, m.dec, c.d1, c.h1, c.n_x);
      break;
    case IWAE_KIND_CIWAE:
    case IWAE_KIND_CIWAE_STD:
      IWAE_REQUIRE(c.n_cond > 0, "iwae_create: CIWAE needs n_cond > 0");
      add_block(m, m.enc1, c.n_x + c.n_cond, c.h1, c.d1);
      add_mlp3(m, m.dec, c.d1 + c.n_cond, c.h1, c.n_x);
      if (c.kind == IWAE_KIND_CIWAE) add_block(m, m.prior, c.n_cond, c.h1, c.d1);
      break;
    default:
      IWAE_REQUIRE(false, "iwae_create: unknown kind %d", c.kind);
  }
  return 0;
}

extern "C" int iwae_create(const iwae_config* cfg, iwae_ctx** out) {
  IWAE_REQUIRE(cfg && out, "iwae_create: null argument");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    iwae_set_error("iwae_create: no CUDA device (%s); libiwae_b200 has no CPU fallback",
                   e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return 3;
  }
  iwae_ctx* ctx = new iwae_ctx();
  int r = build_model(*cfg, ctx->model);
  if (r) { delete ctx; return r; }
  CUDA_TRY(cudaGetDevice(&ctx->device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, ctx->device));
  ctx->sm_count = prop.multiProcessorCount;
  *out = ctx;
  return 0;
}

static void train_graph_release(iwae_ctx* ctx);
extern "C" int iwae_destroy(iwae_ctx* ctx) {
  if (!ctx) return 0;
  cudaDeviceSynchronize();
  tc_release(ctx);
  tc_encoder_release(ctx);
  ctx->arena.release();
  if (ctx->per_buf) cudaFree(ctx->per_buf);
  if (ctx->stage_x) cudaFree(ctx->stage_x);
  if (ctx->stage_aux) cudaFree(ctx->stage_aux);
  train_graph_release(ctx);
  comm_release(ctx);
  for (cudaStream_t& l : ctx->lane) if (l) { cudaStreamDestroy(l); l = nullptr; }
  for (cudaEvent_t& e : ctx->lane_ev) if (e) { cudaEventDestroy(e); e = nullptr; }
  mega_release(ctx);
  if (ctx->wt_ev) { cudaEventDestroy(ctx->wt_ev); ctx->wt_ev = nullptr; }
  if (ctx->wt_buf) { cudaFree(ctx->wt_buf); ctx->wt_buf = nullptr; }
  if (ctx->ev_k0) cudaEventDestroy(ctx->ev_k0);
  if (ctx->ev_k1) cudaEventDestroy(ctx->ev_k1);
  delete ctx;
  return 0;
}

// iwae_param_count / iwae_param_layout work without a device: a NULL-device context is never built,
// so they are also offered through a config-only helper used by the CPU tests.
extern "C" int iwae_param_count(const iwae_ctx* ctx, int64_t* n_params, int32_t* n_tensors) {
  IWAE_REQUIRE(ctx, "iwae_param_count: null ctx");
  if (n_params) *n_params = ctx->model.n_params;
  if (n_tensors) *n_tensors = (int32_t)ctx->model.t_off.size();
  return 0;
}
extern "C" int iwae_param_layout(const iwae_ctx* ctx, int64_t* offsets, int32_t* rows, int32_t* cols) {
  IWAE_REQUIRE(ctx, "iwae_param_layout: null ctx");
  for (size_t i = 0; i < ctx->model.t_off.size(); ++i) {
    if (offsets) offsets[i] = ctx->model.t_off[i];
    if (rows) rows[i] = ctx->model.t_rows[i];
    if (cols) cols[i] = ctx->model.t_cols[i];
  }
  return 0;
}
// Device-free variant (host logic only): layout for a config.  n_tensors in/out = capacity / count.
extern "C" int iwae_layout_for_config(const iwae_config* cfg, int64_t* n_params, int32_t* n_tensors,
                                      int64_t* offsets, int32_t* rows, int32_t* cols) {
  IWAE_REQUIRE(cfg, "iwae_layout_for_config: null cfg");
  ModelP m;
  IWAE_TRY(build_model(*cfg, m));
  const int cap = n_tensors ? *n_tensors : 0;
  if (n_params) *n_params = m.n_params;
  if (n_tensors) *n_tensors = (int32_t)m.t_off.size();
  for (size_t i = 0; i < m.t_off.size() && (int)i < cap; ++i) {
    if (offsets) offsets[i] = m.t_off[i];
    if (rows) rows[i] = m.t_rows[i];
    if (cols) cols[i] = m.t_cols[i];
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// forward (fp32 path)
// ---------------------------------------------------------------------------------------------
struct Fwd {
  int B = 0, k = 0; int64_t R = 0;
  const float* x = nullptr; const float* y = nullptr;
  const float* xin = nullptr; int64_t ld_xin = 0;      // encoder input [B, enc_in]
  float *h1 = nullptr, *h2 = nullptr, *mu1 = nullptr, *std1 = nullptr;
  float *eps1 = nullptr, *z1 = nullptr;
  const float* zin = nullptr; int64_t ld_zin = 0;      // decoder input [R, dec_in]
  float *a1 = nullptr, *a2 = nullptr, *logits = nullptr, *bpart = nullptr;
  float *g1 = nullptr, *g2 = nullptr, *mu2 = nullptr, *std2 = nullptr, *eps2 = nullptr, *z2 = nullptr;
  float *p1 = nullptr, *p2 = nullptr, *pmu = nullptr, *pstd = nullptr;   // IWAE2: [R,.]; CIWAE prior: [B,.]
  float *yoh = nullptr;
  float *lpxz = nullptr, *lpz = nullptr, *lq1 = nullptr, *lq2 = nullptr, *lp12 = nullptr, *lqst = nullptr;
  float *log_w = nullptr, *al = nullptr, *iwae = nullptr, *vae = nullptr, *eq14 = nullptr, *kl = nullptr,
        *dreg = nullptr, *snis1 = nullptr, *snis2 = nullptr;
  float* g = nullptr;       // row coefficient of the backward pass, written by logw_reduce in gradient calls
};

struct FwdOpts {
  bool keep_for_bwd = false;   // materialise logits, eps, al
  bool want_dreg = false;
  bool lanes = false;          // train step: run independent branches on the context's side streams
  int objective = -1;          // >= 0 (gradient calls): logw_reduce also writes the backward's row coefficient g
  int B_norm = 0;
};

// `to` waits for everything issued so far on `from` (an event from the context's pool; re-recording an event after the
// wait on it has been issued is fine, eagerly and under stream capture)
static int lane_dep(iwae_ctx* ctx, cudaStream_t from, cudaStream_t to) {
  if (from == to) return 0;
  if (mega_recording()) { mega_dep(from, to); return 0; }     // train-step megakernel: the edge only orders the levels
  cudaEvent_t& ev = ctx->lane_ev[ctx->lane_ev_next];
  ctx->lane_ev_next = (ctx->lane_ev_next + 1) % 16;
  if (!ev) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(ev, from));
  CUDA_TRY(cudaStreamWaitEvent(to, ev, 0));
  return 0;
}
static int lanes_init(iwae_ctx* ctx) {
  for (int i = 0; i < 3; ++i)
    if (!ctx->lane[i]) CUDA_TRY(cudaStreamCreateWithFlags(&ctx->lane[i], cudaStreamNonBlocking));
  return 0;
}

// Training forward of the encoder's first layer: 20 rows x 784 inputs -> 200 units is 7 CTAs walking 49 K-tiles each
// (22-32 us).  With `scratch` the product is split over K into partial sums that one small kernel reduces with the
// bias and activation.  Only the train step passes a scratch buffer: the evaluation paths must round identically for
// any number of rows (per-example results are bit-identical across shards).
constexpr int FWD_SPLITS = 7;
static int dense_fwd(const float* P, const DenseP& L, const float* X, int64_t ldx, int64_t M, int act, float* Y,
                     cudaStream_t st, float* scratch = nullptr) {
  if (scratch && M <= 64 && L.n_in >= 400) {
    GemmArgs g;
    g.A = X; g.sa_m = ldx; g.sa_k = 1;
    g.B = P + L.w; g.sb_k = L.n_out; g.sb_n = 1;
    g.C = scratch; g.ldc = L.n_out;
    g.M = (int)M; g.N = L.n_out; g.K = L.n_in;
    g.small_tile = 1; g.split_k = FWD_SPLITS; g.split_stride = (int64_t)M * L.n_out;
    IWAE_TRY(gemm_f32(g, st));
    return splitk_bias_act(scratch, FWD_SPLITS, (int)M, L.n_out, P + L.b, act, Y, st);
  }
  GemmArgs g;
  g.A = X; g.sa_m = ldx; g.sa_k = 1;
  g.B = P + L.w; g.sb_k = L.n_out; g.sb_n = 1;
  g.C = Y; g.ldc = L.n_out;
  g.M = (int)M; g.N = L.n_out; g.K = L.n_in;
  g.bias = P + L.b; g.act = act;
  g.small_tile = (((M + 63) / 64) * ((L.n_out + 63) / 64) < 100) ? 1 : 0;
  return gemm_f32(g, st);
}

// Train step only (M <= CHAIN_MAX_ROWS rows, widths <= 256): the block's Dense layers in ONE launch (dense_chain_kernel).
// A first layer with a long reduction (784 pixels on 20 rows) stays on the split-K product, the chain starts after it.
// The evaluation paths keep the tiled GEMMs: their per-row rounding must not depend on which kernel a shape selects.
constexpr int64_t CHAIN_MAX_ROWS = 4096;
// experimental until it beats the per-layer kernels: IWAE_CHAIN=1 turns it on
static bool chain_enabled() { static const bool on = getenv("IWAE_CHAIN") != nullptr; return on; }
// Backward that RECOMPUTES the decoder's output activations (north_star): the forward of a train step does not write
// the [k*B, 784] logits (its GEMM reduces the Bernoulli log-likelihood in the epilogue); the backward runs that GEMM again
// with an epilogue that turns each logit straight into its gradient g*(x - sigmoid(l)).  Same tiles, same summation order:
// gradients and trajectories are bit-identical to the ones from stored logits (tests/test_gpu_parity.py).
// Measured at the training shape (k*B = 1,000 rows): 0.299 ms per step against 0.285 ms with stored logits - the step is
// latency-bound and the second GEMM sits on the decoder's backward chain, while the 3 MB it saves are nothing on a
// 180 GB part.  So the mode is chosen by size: IWAE_RECOMPUTE=1 forces it, =0 forbids it, unset = recompute only when
// the logits of the call would take >= 1 GiB (k*B >= 342,393 rows).
static bool recompute_enabled(int64_t logit_elems) {
  static const int mode = getenv("IWAE_RECOMPUTE") ? (atoi(getenv("IWAE_RECOMPUTE")) != 0 ? 1 : 0) : -1;
  return mode >= 0 ? mode == 1 : logit_elems * (int64_t)sizeof(float) >= ((int64_t)1 << 30);
}
static void chain_dense(ChainLayer& L, const float* P, const DenseP& D, int act, float* out) {
  L.W = P + D.w; L.ldw = D.n_out; L.bias = P + D.b; L.K = D.n_in; L.N = D.n_out; L.nsplit = D.n_out;
  L.act = act; L.out = out; L.ldo = D.n_out;
}
static void chain_heads(ChainLayer& L, const float* P, const BlockP& blk, float* mu, float* sd) {
  const int d = blk.lmu.n_out;
  L.W = P + blk.lmu.w; L.Wb = P + blk.lstd.w; L.ldw = d; L.bias = P + blk.lmu.b; L.biasb = P + blk.lstd.b;
  L.K = blk.lmu.n_in; L.N = 2 * d; L.nsplit = d;
  L.act = CH_NONE; L.actb = CH_EXP_EPS;                         // exp(a) + 1e-6 (iwae1.py:34,42)
  L.out = mu; L.ldo = d; L.outb = sd; L.ldob = d;
}
static ChainArgs block_fwd_chain_args(const float* P, const BlockP& blk, const float* X, int64_t ldx, int64_t M, float* h1,
                                      float* h2, float* mu, float* sd) {
  ChainArgs a;
  a.M = (int)M; a.X = X; a.ldx = ldx; a.kx = blk.l1.n_in; a.nlayers = 3;
  chain_dense(a.L[0], P, blk.l1, CH_TANH, h1);
  chain_dense(a.L[1], P, blk.l2, CH_TANH, h2);
  chain_heads(a.L[2], P, blk, mu, sd);
  return a;
}

int block_fwd_f32(const float* P, const BlockP& blk, const float* X, int64_t ldx, int64_t M, float* h1, float* h2,
                     float* mu, float* sd, cudaStream_t st, float* l1_scratch, iwae_ctx* lane_ctx, cudaStream_t lane) {
  if (lane_ctx && lane && chain_enabled() && M <= CHAIN_MAX_ROWS) {
    const ChainArgs a = block_fwd_chain_args(P, blk, X, ldx, M, h1, h2, mu, sd);
    if (dense_chain_fits(a)) return dense_chain(a, st);
  }
  IWAE_TRY(dense_fwd(P, blk.l1, X, ldx, M, ACT_TANH, h1, st, l1_scratch));
  IWAE_TRY(dense_fwd(P, blk.l2, h1, blk.l1.n_out, M, ACT_TANH, h2, st));
  // the two heads read the same h2 and are independent: in a train step the std head runs on a side stream
  cudaStream_t s2 = (lane_ctx && lane) ? lane : st;
  if (s2 != st) IWAE_TRY(lane_dep(lane_ctx, st, s2));
  IWAE_TRY(dense_fwd(P, blk.lmu, h2, blk.l2.n_out, M, ACT_NONE, mu, st));
  IWAE_TRY(dense_fwd(P, blk.lstd, h2, blk.l2.n_out, M, ACT_EXP_EPS, sd, s2));   // exp(a) + 1e-6 (iwae1.py:34,42)
  if (s2 != st) IWAE_TRY(lane_dep(lane_ctx, s2, st));
  return 0;
}

template <class T>
static int pick(Arena& ar, T* user, size_t n, T** out) {
  if (user) { *out = user; return 0; }
  return ar.get(out, n);
}

static int forward_internal(iwae_ctx* ctx, const float* P, const float* x, const float* y, int B, int k, float beta,
                            const float* eps1_in, const float* eps2_in, uint64_t seed, int64_t img_offset,
                            const iwae_outputs* o, const FwdOpts& opt, Fwd& f, cudaStream_t st) {
  const ModelP& m = ctx->model;
  const iwae_config& c = m.cfg;
  Arena& ar = ctx->arena;
  static const iwae_outputs none = {};
  if (!o) o = &none;
  IWAE_REQUIRE(B > 0 && k > 0, "forward: B and k must be positive (B=%d k=%d)", B, k);
  IWAE_REQUIRE(P && x, "forward: null params or x");
  IWAE_REQUIRE(!m.cond() || y, "forward: CIWAE needs labels y");
  const int64_t R = (int64_t)B * k;
  IWAE_REQUIRE(R < ((int64_t)1 << 31) / 8, "forward: k*B too large for one call (%lld rows); chunk the batch", (long long)R);
  f.B = B; f.k = k; f.R = R; f.x = x; f.y = y;
  const int d1 = c.d1, d2 = c.d2, h1 = c.h1, h2 = c.h2, nx = c.n_x, nc = m.cond() ? c.n_cond : 0;

  // ---- encoder on x (per image): q(z1|x)                                   iwae1.py:56-57
  if (nc) {
    float* xin; IWAE_TRY(ar.get(&xin, (size_t)B * (nx + nc)));
    IWAE_TRY(concat_onehot(x, nx, nx, y, B, B, nc, xin, st));                 // task04.py:112-115
    f.xin = xin; f.ld_xin = nx + nc;
  } else { f.xin = x; f.ld_xin = nx; }
  IWAE_TRY(ar.get(&f.h1, (size_t)B * h1)); IWAE_TRY(ar.get(&f.h2, (size_t)B * h1));
  IWAE_TRY(pick(ar, o->q_mu, (size_t)B * d1, &f.mu1)); IWAE_TRY(pick(ar, o->q_std, (size_t)B * d1, &f.std1));
  float* l1_scratch = nullptr;       // split-K partial sums of the encoder's first layer (train step only, see dense_fwd)
  if (opt.keep_for_bwd && B <= 64) IWAE_TRY(ar.get(&l1_scratch, (size_t)FWD_SPLITS * B * h1));
  cudaStream_t hl = nullptr;           // side stream for the std heads of the BasicBlocks (train step)
  if (opt.lanes) { IWAE_TRY(lanes_init(ctx)); hl = ctx->lane[2]; }
  IWAE_TRY(block_fwd_f32(P, m.enc1, f.xin, f.ld_xin, B, f.h1, f.h2, f.mu1, f.std1, st, l1_scratch, hl ? ctx : nullptr, hl));

  // ---- z1 = mu + std*eps, k samples                                        iwae1.py:59
  IWAE_TRY(pick(ar, o->z1, (size_t)R * d1, &f.z1));
  if (opt.keep_for_bwd) IWAE_TRY(ar.get(&f.eps1, (size_t)R * d1));
  IWAE_TRY(sample_latent(f.mu1, f.std1, 1, eps1_in, f.eps1, f.z1, B, k, d1, seed, 1, img_offset, st));

  // The decoder p(x|z1) depends on z1 only: in a train step of the two-layer model it runs on lane 0 beside the
  // q(z2|z1) -> z2 -> p(z1|z2) chain (9 dependent small kernels), which hides its three GEMMs incl. the 784-logit one.
  cudaStream_t sd = st;
  if (opt.lanes && m.two_layer()) { IWAE_TRY(lanes_init(ctx)); sd = ctx->lane[0]; IWAE_TRY(lane_dep(ctx, st, sd)); }

  // ---- IWAE2: q(z2|z1), z2, p(z1|z2)                                       iwae2.py:63-65,90
  if (m.two_layer()) {
    IWAE_TRY(ar.get(&f.g1, (size_t)R * h2)); IWAE_TRY(ar.get(&f.g2, (size_t)R * h2));
    IWAE_TRY(ar.get(&f.mu2, (size_t)R * d2)); IWAE_TRY(ar.get(&f.std2, (size_t)R * d2));
    IWAE_TRY(block_fwd_f32(P, m.enc2, f.z1, d1, R, f.g1, f.g2, f.mu2, f.std2, st, nullptr, hl ? ctx : nullptr, hl));
    IWAE_TRY(pick(ar, o->z2, (size_t)R * d2, &f.z2));
    if (opt.keep_for_bwd) IWAE_TRY(ar.get(&f.eps2, (size_t)R * d2));
    IWAE_TRY(sample_latent(f.mu2, f.std2, 0, eps2_in, f.eps2, f.z2, B, k, d2, seed, 2, img_offset, st));
    IWAE_TRY(ar.get(&f.p1, (size_t)R * h2)); IWAE_TRY(ar.get(&f.p2, (size_t)R * h2));
    IWAE_TRY(ar.get(&f.pmu, (size_t)R * d1)); IWAE_TRY(ar.get(&f.pstd, (size_t)R * d1));
    IWAE_TRY(block_fwd_f32(P, m.dec21, f.z2, d2, R, f.p1, f.p2, f.pmu, f.pstd, st, nullptr, hl ? ctx : nullptr, hl));
  }
  // ---- CIWAE learned prior p(z|y)                                          task04.py:123
  if (m.learned_prior()) {
    IWAE_TRY(ar.get(&f.yoh, (size_t)B * nc));
    IWAE_TRY(onehot_f32(y, B, nc, f.yoh, st));
    IWAE_TRY(ar.get(&f.p1, (size_t)B * h1)); IWAE_TRY(ar.get(&f.p2, (size_t)B * h1));
    IWAE_TRY(ar.get(&f.pmu, (size_t)B * d1)); IWAE_TRY(ar.get(&f.pstd, (size_t)B * d1));
    IWAE_TRY(block_fwd_f32(P, m.prior, f.yoh, nc, B, f.p1, f.p2, f.pmu, f.pstd, st));
  }

  // ---- decoder p(x|z1)                                                     iwae1.py:79-85
  if (nc) {
    float* zin; IWAE_TRY(ar.get(&zin, (size_t)R * (d1 + nc)));
    IWAE_TRY(concat_onehot(f.z1, d1, d1, y, B, (int)R, nc, zin, sd));         // task04.py:119
    f.zin = zin; f.ld_zin = d1 + nc;
  } else { f.zin = f.z1; f.ld_zin = d1; }
  IWAE_TRY(ar.get(&f.a1, (size_t)R * h1)); IWAE_TRY(ar.get(&f.a2, (size_t)R * h1));
  bool dec_chained = false;
  if (opt.lanes && chain_enabled() && R <= CHAIN_MAX_ROWS) {
    ChainArgs a;
    a.M = (int)R; a.X = f.zin; a.ldx = f.ld_zin; a.kx = m.dec.d0.n_in; a.nlayers = 2;
    chain_dense(a.L[0], P, m.dec.d0, CH_TANH, f.a1);
    chain_dense(a.L[1], P, m.dec.d1, CH_TANH, f.a2);
    if (dense_chain_fits(a)) { IWAE_TRY(dense_chain(a, sd)); dec_chained = true; }
  }
  if (!dec_chained) {
    IWAE_TRY(dense_fwd(P, m.dec.d0, f.zin, f.ld_zin, R, ACT_TANH, f.a1, sd));
    IWAE_TRY(dense_fwd(P, m.dec.d1, f.a1, h1, R, ACT_TANH, f.a2, sd));
  }
  {
    const bool mat = (opt.keep_for_bwd && !recompute_enabled((int64_t)R * nx)) || o->logits;
    if (mat) IWAE_TRY(pick(ar, o->logits, (size_t)R * nx, &f.logits));
    const int nt = gemm_f32_ntiles_n(nx, 0);
    IWAE_TRY(ar.get(&f.bpart, (size_t)R * nt));
    GemmArgs g;
    g.A = f.a2; g.sa_m = h1; g.sa_k = 1;
    g.B = P + m.dec.d2.w; g.sb_k = nx; g.sb_n = 1;
    g.C = f.logits; g.ldc = nx;
    g.M = (int)R; g.N = nx; g.K = h1; g.bias = P + m.dec.d2.b;
    g.bern_x = x; g.bern_B = B; g.bern_partial = f.bpart;      // fused Bernoulli log-lik (iwae1.py:111)
    IWAE_TRY(gemm_f32(g, sd));
    IWAE_TRY(pick(ar, o->lpxz, (size_t)R, &f.lpxz));
    IWAE_TRY(bern_reduce_partials(f.bpart, (int)R, nt, f.lpxz, sd));
  }
  IWAE_TRY(lane_dep(ctx, sd, st));

  // ---- Gaussian terms                                                      iwae1.py:107-109, iwae2.py:116-122
  IWAE_TRY(pick(ar, o->lpz, (size_t)R, &f.lpz));
  IWAE_TRY(pick(ar, o->lq1, (size_t)R, &f.lq1));
  // all Gaussian log-density row sums of the call in ONE launch (they are independent; up to four kernels before)
  NlpJobs jobs;
  auto add_job = [&](const float* z, const float* mu, const float* sd, int bcast, float std_extra, int d, float* out) {
    NlpJob& j = jobs.job[jobs.count++];
    j.z = z; j.mu = mu; j.sd = sd; j.bcast = bcast; j.std_extra = std_extra; j.d = d; j.out = out;
  };
  add_job(f.z1, f.mu1, f.std1, 1, 0.f, d1, f.lq1);
  if (m.two_layer()) {
    IWAE_TRY(pick(ar, o->lq2, (size_t)R, &f.lq2));
    IWAE_TRY(pick(ar, o->lp12, (size_t)R, &f.lp12));
    add_job(f.z2, nullptr, nullptr, 0, 0.f, d2, f.lpz);        // lpz2
    add_job(f.z2, f.mu2, f.std2, 0, 0.f, d2, f.lq2);           // lqz2z1
    add_job(f.z1, f.pmu, f.pstd, 0, 0.f, d1, f.lp12);          // lpz1z2
  } else if (m.learned_prior()) {
    add_job(f.z1, f.pmu, f.pstd, 1, 0.f, d1, f.lpz);           // lpzy
  } else {
    add_job(f.z1, nullptr, nullptr, 0, 0.f, d1, f.lpz);        // lpz
  }
  if (opt.want_dreg || o->dreg) {
    IWAE_TRY(ar.get(&f.lqst, (size_t)R));
    add_job(f.z1, f.mu1, f.std1, 1, 1e-6f, d1, f.lqst);        // task02.py:61-65
  }
  IWAE_TRY(normal_logp_rows_multi(jobs, B, k, st));
  if ((o->kl || o->scalars) && !m.two_layer()) {
    IWAE_TRY(pick(ar, o->kl, (size_t)B, &f.kl));
    IWAE_TRY(kl_rows(f.mu1, f.std1, B, d1, f.kl, st));                                      // iwae1.py:116
  }

  // ---- log-weights and reductions over k                                   iwae1.py:113-139
  LogwArgs la;
  la.B = B; la.k = k; la.beta = beta; la.two_layer = m.two_layer() ? 1 : 0;
  la.lpxz = f.lpxz; la.lpz = f.lpz; la.lq1 = f.lq1; la.lq2 = f.lq2; la.lp12 = f.lp12; la.lq_stopped = f.lqst;
  la.z1 = f.z1; la.z2 = f.z2; la.d1 = d1; la.d2 = d2;
  f.log_w = o->log_w; f.al = o->al; f.iwae = o->iwae; f.vae = o->vae; f.eq14 = o->eq14;
  f.snis1 = o->snis_z1; f.snis2 = m.two_layer() ? o->snis_z2 : nullptr;
  if (opt.keep_for_bwd && !f.al) IWAE_TRY(ar.get(&f.al, (size_t)R));
  if (o->scalars) {
    if (!f.iwae) IWAE_TRY(ar.get(&f.iwae, (size_t)B));
    if (!f.vae) IWAE_TRY(ar.get(&f.vae, (size_t)B));
    if (!f.eq14) IWAE_TRY(ar.get(&f.eq14, (size_t)B));
  }
  if (f.lqst) IWAE_TRY(pick(ar, o->dreg, (size_t)B, &f.dreg));
  la.log_w = f.log_w; la.al = f.al; la.iwae = f.iwae; la.vae = f.vae; la.eq14 = f.eq14; la.dreg = f.dreg;
  la.snis1 = f.snis1; la.snis2 = f.snis2;
  if (opt.objective >= 0) {
    // DReG's decoder part is the IWAE-ELBO gradient (task02.py:96)
    IWAE_TRY(ar.get(&f.g, (size_t)R));
    la.g_out = f.g; la.B_norm = opt.B_norm;
    la.g_vae = (opt.objective == IWAE_OBJ_VAE_ELBO || opt.objective == IWAE_OBJ_VAE_ELBO_KL) ? 1 : 0;
  }
  IWAE_TRY(logw_reduce(la, st));
  if (o->scalars) {
    ScalarArgs sa;
    sa.B = B; sa.R = R; sa.beta = beta;
    sa.iwae = f.iwae; sa.vae = f.vae; sa.eq14 = f.eq14; sa.kl = f.kl; sa.dreg = f.dreg;
    sa.lpxz = f.lpxz; sa.lpz = f.lpz; sa.lq1 = f.lq1; sa.out = o->scalars;
    if (R >= SCALAR_PARTS_MIN_ROWS) IWAE_TRY(ar.get(&sa.partial, (size_t)SCALAR_PARTS * 8));
    // train step: the batch scalars are outputs only - nothing of the backward pass reads them - so they leave the
    // critical path for lane 1 (joined at the end of backward_internal together with the weight gradients)
    cudaStream_t ss = st;
    if (opt.lanes) { IWAE_TRY(lanes_init(ctx)); ss = ctx->lane[1]; IWAE_TRY(lane_dep(ctx, st, ss)); }
    IWAE_TRY(finalize_scalars(sa, ss));
  }
  return 0;
}

// IWAE.call / val_step with the fused tensor-core kernel as the engine: the kernel (TERMS mode) writes the per-row
// log-density terms and the latent samples, the same reduction kernels as the fp32 path turn them into log_w, al, the
// ELBO variants, SNIS means and the batch scalars.  logits [k,B,784] never exist (not available from this entry point:
// ask iwae_forward for them); DReG's extra outputs likewise.
extern "C" int iwae_forward_tc(iwae_ctx* ctx, const float* params, const float* x, const float* y, int32_t B, int32_t k,
                               float beta, uint64_t seed, int64_t img_offset, int32_t precision, const iwae_outputs* o,
                               void* stream) {
  IWAE_REQUIRE(ctx && params && x && o, "iwae_forward_tc: null argument");
  IWAE_REQUIRE(B > 0 && k > 0, "iwae_forward_tc: B and k must be positive (B=%d k=%d)", B, k);
  IWAE_REQUIRE(precision == IWAE_PREC_TC_F16 || precision == IWAE_PREC_TC_F16_FAST, "iwae_forward_tc: tensor-core precisions only");
  IWAE_REQUIRE(!o->logits && !o->dreg, "iwae_forward_tc: logits / dreg are not produced by the fused kernel (use iwae_forward)");
  cudaStream_t st = (cudaStream_t)stream;
  const ModelP& m = ctx->model;
  const iwae_config& c = m.cfg;
  Arena& ar = ctx->arena;
  ar.reset();
  const int64_t R = (int64_t)B * k;
  const int d1 = c.d1, d2 = c.d2;
  // buffers that must outlive tc_llh_eval's own arena use: take them first (the arena only grows within a call)
  Fwd f;
  f.B = B; f.k = k; f.R = R;
  IWAE_TRY(pick(ar, o->q_mu, (size_t)B * d1, &f.mu1)); IWAE_TRY(pick(ar, o->q_std, (size_t)B * d1, &f.std1));
  IWAE_TRY(pick(ar, o->lpxz, (size_t)R, &f.lpxz)); IWAE_TRY(pick(ar, o->lpz, (size_t)R, &f.lpz));
  IWAE_TRY(pick(ar, o->lq1, (size_t)R, &f.lq1));
  const bool want_snis1 = o->snis_z1 != nullptr, want_snis2 = m.two_layer() && o->snis_z2 != nullptr;
  if (o->z1 || want_snis1) IWAE_TRY(pick(ar, o->z1, (size_t)R * d1, &f.z1));
  if (m.two_layer()) {
    IWAE_TRY(pick(ar, o->lq2, (size_t)R, &f.lq2)); IWAE_TRY(pick(ar, o->lp12, (size_t)R, &f.lp12));
    if (o->z2 || want_snis2) IWAE_TRY(pick(ar, o->z2, (size_t)R * d2, &f.z2));
  }
  float* per; IWAE_TRY(ar.get(&per, (size_t)B));
  TcTerms t = {f.lpxz, f.lpz, f.lq1, f.lq2, f.lp12, f.z1, f.z2, f.mu1, f.std1};
  IWAE_TRY(tc_llh_eval(ctx, params, x, y, B, k, seed, img_offset, precision == IWAE_PREC_TC_F16_FAST, per, nullptr, st, nullptr,
                       nullptr, &t));
  if ((o->kl || o->scalars) && !m.two_layer()) {
    IWAE_TRY(pick(ar, o->kl, (size_t)B, &f.kl));
    IWAE_TRY(kl_rows(f.mu1, f.std1, B, d1, f.kl, st));                                      // iwae1.py:116
  }
  LogwArgs la;
  la.B = B; la.k = k; la.beta = beta; la.two_layer = m.two_layer() ? 1 : 0;
  la.lpxz = f.lpxz; la.lpz = f.lpz; la.lq1 = f.lq1; la.lq2 = f.lq2; la.lp12 = f.lp12;
  la.z1 = f.z1; la.z2 = f.z2; la.d1 = d1; la.d2 = d2;
  f.iwae = o->iwae; f.vae = o->vae; f.eq14 = o->eq14;
  if (o->scalars) {
    if (!f.iwae) IWAE_TRY(ar.get(&f.iwae, (size_t)B));
    if (!f.vae) IWAE_TRY(ar.get(&f.vae, (size_t)B));
    if (!f.eq14) IWAE_TRY(ar.get(&f.eq14, (size_t)B));
  }
  la.log_w = o->log_w; la.al = o->al; la.iwae = f.iwae; la.vae = f.vae; la.eq14 = f.eq14;
  la.snis1 = want_snis1 ? o->snis_z1 : nullptr; la.snis2 = want_snis2 ? o->snis_z2 : nullptr;
  IWAE_TRY(logw_reduce(la, st));
  if (o->scalars) {
    ScalarArgs sa;
    sa.B = B; sa.R = R; sa.beta = beta;
    sa.iwae = f.iwae; sa.vae = f.vae; sa.eq14 = f.eq14; sa.kl = f.kl;
    sa.lpxz = f.lpxz; sa.lpz = f.lpz; sa.lq1 = f.lq1; sa.out = o->scalars;
    if (R >= SCALAR_PARTS_MIN_ROWS) IWAE_TRY(ar.get(&sa.partial, (size_t)SCALAR_PARTS * 8));
    IWAE_TRY(finalize_scalars(sa, st));
  }
  return 0;
}

extern "C" int iwae_forward(iwae_ctx* ctx, const float* params, const float* x, const float* y, int32_t B, int32_t k,
                            float beta, const float* eps1, const float* eps2, uint64_t seed, int64_t img_offset,
                            const iwae_outputs* outs, void* stream) {
  IWAE_REQUIRE(ctx, "iwae_forward: null ctx");
  ctx->arena.reset();
  Fwd f; FwdOpts opt;
  return forward_internal(ctx, params, x, y, B, k, beta, eps1, eps2, seed, img_offset, outs, opt, f,
                          (cudaStream_t)stream);
}

static int grow_dev(float** p, size_t* cap, size_t bytes) {
  if (*cap >= bytes) return 0;
  if (*p) CUDA_TRY(cudaFree(*p));
  *p = nullptr; *cap = 0;
  CUDA_TRY(cudaMalloc((void**)p, bytes));
  *cap = bytes;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// k-sample log-likelihood evaluation (main.py:170-184)
// ---------------------------------------------------------------------------------------------
static int llh_eval_impl(iwae_ctx* ctx, const float* params, const float* x, const float* y, int64_t N, int32_t k,
                         uint64_t seed, int64_t img_offset, int32_t precision, float* out_per_example, float* out_vae,
                         double* out_sum, void* stream);

extern "C" int iwae_llh_eval(iwae_ctx* ctx, const float* params, const float* x, const float* y, int64_t N, int32_t k,
                             uint64_t seed, int64_t img_offset, int32_t precision, float* out_per_example,
                             double* out_sum, void* stream) {
  return llh_eval_impl(ctx, params, x, y, N, k, seed, img_offset, precision, out_per_example, nullptr, out_sum, stream);
}

extern "C" int iwae_monitor_eval(iwae_ctx* ctx, const float* params, const float* x, const float* y, int64_t N, int32_t k,
                                 uint64_t seed, int64_t img_offset, int32_t precision, float* out_iwae, float* out_vae,
                                 double* out_sums, void* stream) {
  IWAE_REQUIRE(out_iwae && out_vae, "iwae_monitor_eval: null output");
  IWAE_TRY(llh_eval_impl(ctx, params, x, y, N, k, seed, img_offset, precision, out_iwae, out_vae, out_sums, stream));
  if (out_sums && N > 0) IWAE_TRY(sum_to_double(out_vae, N, out_sums + 1, 0, (cudaStream_t)stream));
  else if (out_sums) CUDA_TRY(cudaMemsetAsync(out_sums + 1, 0, sizeof(double), (cudaStream_t)stream));
  return 0;
}

static int llh_eval_impl(iwae_ctx* ctx, const float* params, const float* x, const float* y, int64_t N, int32_t k,
                         uint64_t seed, int64_t img_offset, int32_t precision, float* out_per_example, float* out_vae,
                         double* out_sum, void* stream) {
  IWAE_REQUIRE(ctx && params && (x || N == 0), "iwae_llh_eval: null argument");
  IWAE_REQUIRE(N >= 0 && k > 0, "iwae_llh_eval: bad N=%lld k=%d", (long long)N, k);
  cudaStream_t st = (cudaStream_t)stream;
  const ModelP& m = ctx->model;
  ctx->arena.reset();
  float* per = out_per_example;
  if (!per) {
    IWAE_TRY(grow_dev(&ctx->per_buf, &ctx->per_cap, (size_t)std::max<int64_t>(N, 1) * sizeof(float)));
    per = ctx->per_buf;
  }
  if (N == 0) {
    if (out_sum) CUDA_TRY(cudaMemsetAsync(out_sum, 0, sizeof(double), st));
    return 0;
  }
  if (precision == IWAE_PREC_TC_F16 || precision == IWAE_PREC_TC_F16_FAST) {
    IWAE_TRY(tc_llh_eval(ctx, params, x, y, N, k, seed, img_offset, precision == IWAE_PREC_TC_F16_FAST, per, out_vae, st));
  } else {
    IWAE_REQUIRE(precision == IWAE_PREC_FP32, "iwae_llh_eval: unknown precision %d", precision);
    // chunk the images so that one chunk's activations stay around 256k rows
    int64_t Bc = std::max<int64_t>(1, (int64_t)262144 / k);
    Bc = std::min(Bc, N);
    for (int64_t i0 = 0; i0 < N; i0 += Bc) {
      const int Bi = (int)std::min<int64_t>(Bc, N - i0);
      ctx->arena.reset();
      iwae_outputs o = {};
      o.iwae = per + i0;
      if (out_vae) o.vae = out_vae + i0;
      Fwd f; FwdOpts opt;
      IWAE_TRY(forward_internal(ctx, params, x + i0 * m.cfg.n_x, y ? y + i0 : nullptr, Bi, k, 1.0f, nullptr, nullptr,
                                seed, img_offset + i0, &o, opt, f, st));
    }
  }
  if (out_sum) IWAE_TRY(sum_to_double(per, N, out_sum, 0, st));
  return 0;
}

extern "C" int iwae_llh_eval_eps(iwae_ctx* ctx, const float* params, const float* x, const float* y, int64_t N, int32_t k,
                                 const float* eps1, const float* eps2, int32_t precision, float* out_per_example,
                                 double* out_sum, void* stream) {
  IWAE_REQUIRE(ctx && params && x && eps1 && out_per_example, "iwae_llh_eval_eps: null argument");
  IWAE_REQUIRE(N > 0 && k > 0, "iwae_llh_eval_eps: bad N=%lld k=%d", (long long)N, k);
  IWAE_REQUIRE(precision == IWAE_PREC_TC_F16 || precision == IWAE_PREC_TC_F16_FAST,
               "iwae_llh_eval_eps: tensor-core precisions only (iwae_forward takes injected noise on the fp32 path)");
  cudaStream_t st = (cudaStream_t)stream;
  ctx->arena.reset();
  IWAE_TRY(tc_llh_eval(ctx, params, x, y, N, k, 0, 0, precision == IWAE_PREC_TC_F16_FAST, out_per_example, nullptr, st, eps1, eps2));
  if (out_sum) IWAE_TRY(sum_to_double(out_per_example, N, out_sum, 0, st));
  return 0;
}

extern "C" int iwae_set_params_version(iwae_ctx* ctx, uint64_t version) {
  IWAE_REQUIRE(ctx, "iwae_set_params_version: null context");
  ctx->params_version = version;
  return 0;
}

extern "C" int iwae_llh_kernel_ms(iwae_ctx* ctx, float* ms) {
  IWAE_REQUIRE(ctx && ms, "iwae_llh_kernel_ms: null argument");
  IWAE_REQUIRE(ctx->ev_valid, "iwae_llh_kernel_ms: no tensor-core evaluation has run on this context");
  CUDA_TRY(cudaEventSynchronize(ctx->ev_k1));
  CUDA_TRY(cudaEventElapsedTime(ms, ctx->ev_k0, ctx->ev_k1));
  return 0;
}

extern "C" int iwae_llh_eval_host(iwae_ctx* ctx, const float* params, const float* x_host, const float* y_host,
                                  int64_t N, int32_t k, uint64_t seed, int64_t img_offset, int32_t precision,
                                  float* out_per_example_host, double* out_sum_host, void* stream) {
  IWAE_REQUIRE(ctx && params && (x_host || N == 0), "iwae_llh_eval_host: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int nx = ctx->model.cfg.n_x;
  // persistent device staging (not the arena: iwae_llh_eval resets it)
  const size_t xb = (size_t)N * nx * sizeof(float);
  IWAE_TRY(grow_dev(&ctx->stage_x, &ctx->stage_x_cap, std::max(xb, (size_t)256)));
  IWAE_TRY(grow_dev(&ctx->stage_aux, &ctx->stage_aux_cap, (size_t)N * 2 * sizeof(float) + 64));
  float* d_x = ctx->stage_x;
  float* d_aux = ctx->stage_aux;
  double* d_sum = reinterpret_cast<double*>(d_aux);
  float* d_per = d_aux + 16;
  float* d_y = d_per + N;
  CUDA_TRY(cudaMemcpyAsync(d_x, x_host, xb, cudaMemcpyHostToDevice, st));
  if (y_host) CUDA_TRY(cudaMemcpyAsync(d_y, y_host, (size_t)N * sizeof(float), cudaMemcpyHostToDevice, st));
  IWAE_TRY(iwae_llh_eval(ctx, params, d_x, y_host ? d_y : nullptr, N, k, seed, img_offset, precision, d_per, d_sum, st));
  if (out_per_example_host)
    CUDA_TRY(cudaMemcpyAsync(out_per_example_host, d_per, (size_t)N * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (out_sum_host) CUDA_TRY(cudaMemcpyAsync(out_sum_host, d_sum, sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// backward (fp32 path)
// ---------------------------------------------------------------------------------------------
// Parameter gradients are produced as GRAD_SPLITS partial sums over the rows (split-K): a weight gradient
// dW = X^T dY has few output tiles (100x100 .. 200x784) but a long reduction (k*B = 1,000 rows), so the
// row range is split over gridDim.z to fill the SMs.  `grads` points to slice 0 of a zeroed
// [GRAD_SPLITS][n_params] workspace; grad_reduce / adam_from_partials sum the slices in a fixed order.
constexpr int GRAD_SPLITS = 8;
static thread_local int g_grad_stride = 0;      // n_params of the model whose backward is running on this host thread (slice stride)

// dW = X^T dY, db = colsum(dY)
// `st` is the stream dY was produced on; with a weight-gradient lane `wl` the product runs there (after st's work so
// far) so that the dX chain on st does not wait for it
static int dense_bwd_w(iwae_ctx* ctx, const DenseP& L, const float* X, int64_t ldx, int64_t M, const float* dY, float* grads,
                       cudaStream_t st, cudaStream_t wl = nullptr) {
  if (wl && wl != st) { IWAE_TRY(lane_dep(ctx, st, wl)); st = wl; }
  GemmArgs g;
  g.A = X; g.sa_m = 1; g.sa_k = ldx;                 // A(i_in, row)
  g.B = dY; g.sb_k = L.n_out; g.sb_n = 1;            // B(row, o)
  g.C = grads + L.w; g.ldc = L.n_out;
  g.M = L.n_in; g.N = L.n_out; g.K = (int)M;
  g.small_tile = 1;
  // db = colsum(dY) rides along as a virtual row of ones in X^T: the bias follows its kernel in the flat
  // parameter layout (add_dense), so row n_in of the [n_in + 1, n_out] product lands on the bias gradient
  const bool fused_bias = (L.b == L.w + (int64_t)L.n_in * L.n_out);
  g.ones_row = fused_bias ? 1 : 0;
  const int tiles = ((L.n_in + g.ones_row + 31) / 32) * ((L.n_out + 31) / 32);
  // enough slices to fill the SMs AND to keep each CTA's serial K loop short (<= ~128 rows): the slices are summed
  // by grad_reduce anyway, and a 200 x 784 gradient over 1,000 rows in two slices was a 42 us chain of 32 K-tiles
  int split = std::max<int>((296 + tiles - 1) / tiles, (int)((M + 127) / 128));
  split = std::min<int>(GRAD_SPLITS, split);
  split = (int)std::max<int64_t>(1, std::min<int64_t>(split, M / 64));     // at least 64 rows per slice
  g.split_k = split; g.split_stride = g_grad_stride;
  IWAE_TRY(gemm_f32(g, st));
  if (fused_bias) return 0;
  return colsum_f32(dY, L.n_out, (int)M, L.n_out, grads + L.b, st, GRAD_SPLITS, g_grad_stride);
}
// dX = dY W^T, optionally accumulated into dX and/or multiplied by tanh'(H) = 1 - H^2
constexpr int BWDX_SPLITS = 4;
static int dense_bwd_x(const float* P, const DenseP& L, const float* dY, int64_t M, float* dX, int64_t ld_dx,
                       int accumulate, const float* H, int64_t ld_h, cudaStream_t st, float* scratch = nullptr) {
  if (scratch && L.n_out >= 400) {
    // long reduction (the 784 logits of the output layer): K split into partial products, one epilogue kernel
    GemmArgs g;
    g.A = dY; g.sa_m = L.n_out; g.sa_k = 1;
    g.B = P + L.w; g.sb_k = 1; g.sb_n = L.n_out;
    g.C = scratch; g.ldc = L.n_in;
    g.M = (int)M; g.N = L.n_in; g.K = L.n_out;
    g.small_tile = 1; g.split_k = BWDX_SPLITS; g.split_stride = M * L.n_in;
    IWAE_TRY(gemm_f32(g, st));
    return splitk_mul_h(scratch, BWDX_SPLITS, (int)M, L.n_in, H, ld_h, accumulate, dX, ld_dx, st);
  }
  GemmArgs g;
  g.A = dY; g.sa_m = L.n_out; g.sa_k = 1;
  g.B = P + L.w; g.sb_k = 1; g.sb_n = L.n_out;       // B(o, i) = W[i, o]
  g.C = dX; g.ldc = ld_dx;
  g.M = (int)M; g.N = L.n_in; g.K = L.n_out;
  g.accumulate = accumulate; g.mul_h = H; g.ld_h = ld_h;
  g.small_tile = (((M + 63) / 64) * ((L.n_in + 63) / 64) < 100) ? 1 : 0;
  return gemm_f32(g, st);
}

// Transposed weights for the fused backward chains: the copy of a Dense kernel [n_in, n_out] lives at the kernel's own
// offset in ctx->wt_buf as [n_out, n_in]; a block's two heads are stored back to back ([2d, h]) so that
// d(h2) = [dmu | da] [Wmu^T ; Wstd^T] is one product (the bias slot between the two kernels gives the room).
static void add_transpose(TransposeJobs& t, const float* P, float* WT, const DenseP& L, int64_t dst_off) {
  TransposeJob& j = t.job[t.count++];
  j.src = P + L.w; j.dst = WT + dst_off; j.rows = L.n_in; j.cols = L.n_out; j.ld_src = L.n_out; j.ld_dst = L.n_in;
}
static void add_block_transposes(TransposeJobs& t, const float* P, float* WT, const BlockP& blk, bool need_l1) {
  add_transpose(t, P, WT, blk.lmu, blk.lmu.w);
  add_transpose(t, P, WT, blk.lstd, blk.lmu.w + (int64_t)blk.lmu.n_in * blk.lmu.n_out);
  add_transpose(t, P, WT, blk.l2, blk.l2.w);
  if (need_l1) add_transpose(t, P, WT, blk.l1, blk.l1.w);
}
static bool bwd_chain_ok(const ModelP& m, int64_t R) {
  if (!chain_enabled() || R > CHAIN_MAX_ROWS) return false;
  auto blk_ok = [](const BlockP& b) {
    return b.l1.n_out <= 256 && 2 * b.lmu.n_out <= 256 && b.lstd.w >= b.lmu.w + (int64_t)b.lmu.n_in * b.lmu.n_out;
  };
  if (!blk_ok(m.enc1) || m.cfg.h1 > 256 || m.dec_in() > 256) return false;
  if (m.two_layer() && (!blk_ok(m.enc2) || !blk_ok(m.dec21) || m.cfg.d1 > 256 || m.cfg.d2 > 256)) return false;
  if (m.learned_prior() && !blk_ok(m.prior)) return false;
  return true;
}
// one launch on `ts` that rebuilds every transposed copy the step's backward chains read; `done` is recorded after it
static int launch_transposes(iwae_ctx* ctx, const float* P, cudaStream_t ts) {
  const ModelP& m = ctx->model;
  if (!ctx->wt_buf) {
    CUDA_TRY(cudaMalloc((void**)&ctx->wt_buf, sizeof(float) * (size_t)m.n_params));
    ctx->wt_cap = (size_t)m.n_params;
  }
  TransposeJobs t;
  add_block_transposes(t, P, ctx->wt_buf, m.enc1, false);
  if (m.two_layer()) {
    add_block_transposes(t, P, ctx->wt_buf, m.enc2, true);
    add_block_transposes(t, P, ctx->wt_buf, m.dec21, true);
  }
  if (m.learned_prior()) add_block_transposes(t, P, ctx->wt_buf, m.prior, false);
  add_transpose(t, P, ctx->wt_buf, m.dec.d0, m.dec.d0.w);
  add_transpose(t, P, ctx->wt_buf, m.dec.d1, m.dec.d1.w);
  return transpose_jobs(t, ts);
}
static void chain_bwd_dense(ChainLayer& L, const float* WT, const DenseP& D, const float* H, float* out) {
  L.W = WT + D.w; L.ldw = D.n_in; L.K = D.n_out; L.N = D.n_in; L.nsplit = D.n_in;
  L.act = H ? CH_MULH : CH_NONE; L.H = H; L.ldh = D.n_in; L.out = out; L.ldo = D.n_in;
}
// BasicBlock backward with the dX products of the whole block in one launch
static int block_bwd_chain(iwae_ctx* ctx, const BlockP& blk, const float* X, int64_t ldx, int64_t M,
                           const float* h1, const float* h2, const float* dmu, const float* da, float* grads, float* dX,
                           cudaStream_t st, cudaStream_t wl) {
  const int h = blk.l1.n_out, d = blk.lmu.n_out;
  ChainArgs a;
  a.M = (int)M; a.X = dmu; a.ldx = d; a.kx = d; a.X2 = da; a.ldx2 = d; a.kx2 = d;
  ChainLayer& L0 = a.L[0];                                  // dh2 = ([dmu | da] [Wmu^T ; Wstd^T]) * (1 - h2^2)
  L0.W = ctx->wt_buf + blk.lmu.w; L0.ldw = h; L0.K = 2 * d; L0.N = h; L0.nsplit = h;
  L0.act = CH_MULH; L0.H = h2; L0.ldh = h; L0.out = nullptr; L0.ldo = h;
  chain_bwd_dense(a.L[1], ctx->wt_buf, blk.l2, h1, nullptr);   // dh1 = (dh2 W2^T) * (1 - h1^2)
  a.nlayers = 2;
  if (dX) { chain_bwd_dense(a.L[2], ctx->wt_buf, blk.l1, nullptr, dX); a.nlayers = 3; }
  if (!dense_chain_fits(a)) return 1;                       // the caller falls back to the per-layer products
  float *dh2, *dh1;
  IWAE_TRY(ctx->arena.get(&dh2, (size_t)M * h)); IWAE_TRY(ctx->arena.get(&dh1, (size_t)M * h));
  a.L[0].out = dh2; a.L[1].out = dh1;
  IWAE_TRY(dense_bwd_w(ctx, blk.lmu, h2, h, M, dmu, grads, st, wl));
  IWAE_TRY(dense_bwd_w(ctx, blk.lstd, h2, h, M, da, grads, st, wl));
  IWAE_TRY(dense_chain(a, st));
  IWAE_TRY(dense_bwd_w(ctx, blk.l2, h1, h, M, dh2, grads, st, wl));
  IWAE_TRY(dense_bwd_w(ctx, blk.l1, X, ldx, M, dh1, grads, st, wl));
  return 0;
}

// BasicBlock backward from (dmu, da) where a is the std pre-activation.  dX may be null.
static int block_bwd(iwae_ctx* ctx, const float* P, const BlockP& blk, const float* X, int64_t ldx, int64_t M,
                     const float* h1, const float* h2, const float* dmu, const float* da, float* grads, float* dX,
                     cudaStream_t st, cudaStream_t wl = nullptr, bool chain = false) {
  if (chain) {
    const int rc = block_bwd_chain(ctx, blk, X, ldx, M, h1, h2, dmu, da, grads, dX, st, wl);
    if (rc != 1) return rc;                                  // 1: this block's chain does not fit shared memory
  }
  const int h = blk.l1.n_out;
  float *dh2, *dh1;
  IWAE_TRY(ctx->arena.get(&dh2, (size_t)M * h)); IWAE_TRY(ctx->arena.get(&dh1, (size_t)M * h));
  IWAE_TRY(dense_bwd_w(ctx, blk.lmu, h2, h, M, dmu, grads, st, wl));
  IWAE_TRY(dense_bwd_w(ctx, blk.lstd, h2, h, M, da, grads, st, wl));
  IWAE_TRY(dense_bwd_x(P, blk.lmu, dmu, M, dh2, h, 0, nullptr, 0, st));
  IWAE_TRY(dense_bwd_x(P, blk.lstd, da, M, dh2, h, 1, h2, h, st));            // (.. + ..) * (1 - h2^2)
  IWAE_TRY(dense_bwd_w(ctx, blk.l2, h1, h, M, dh2, grads, st, wl));
  IWAE_TRY(dense_bwd_x(P, blk.l2, dh2, M, dh1, h, 0, h1, h, st));
  IWAE_TRY(dense_bwd_w(ctx, blk.l1, X, ldx, M, dh1, grads, st, wl));
  if (dX) IWAE_TRY(dense_bwd_x(P, blk.l1, dh1, M, dX, blk.l1.n_in, 0, nullptr, 0, st));
  return 0;
}

static int backward_internal(iwae_ctx* ctx, const float* P, const Fwd& f, float beta, int objective, int B_norm,
                             float* grads, cudaStream_t st, bool lanes = false, bool chain = false) {
  const ModelP& m = ctx->model;
  const iwae_config& c = m.cfg;
  Arena& ar = ctx->arena;
  const int B = f.B, k = f.k; const int64_t R = f.R;
  const int d1 = c.d1, d2 = c.d2, h1 = c.h1, nx = c.n_x;
  // DReG's decoder part is the IWAE-ELBO gradient (task02.py:96)
  const int dec_obj = (objective == IWAE_OBJ_DREG) ? IWAE_OBJ_IWAE_ELBO : objective;

  // Lanes (train step): the decoder's backward chain (lane 0) and the p(z1|z2) -> z2 -> q(z2|z1) chain (caller's stream)
  // are independent until they meet in the first latent layer; every weight gradient dW = X^T dY leaves its chain for
  // lane 1 / lane 2, so a chain is only its dX products.  All weight gradients write disjoint slices of `grads`.
  cudaStream_t s_dec = st, w_dec = nullptr, w_lat = nullptr;
  if (lanes) {
    IWAE_TRY(lanes_init(ctx));
    w_dec = ctx->lane[1]; w_lat = ctx->lane[2];
    if (m.two_layer()) s_dec = ctx->lane[0];
  }
  float *g = f.g, *dl, *da2, *da1, *dzin;
  if (!g) {                                        // (a forward that did not write the coefficient)
    IWAE_TRY(ar.get(&g, (size_t)R));
    IWAE_TRY(row_coef(f.al, B, k, dec_obj, B_norm, g, st));
  }
  IWAE_TRY(lane_dep(ctx, st, s_dec));
  IWAE_TRY(ar.get(&dl, (size_t)R * nx));
  if (f.logits) {
    IWAE_TRY(dlogits_f32(f.logits, f.x, g, (int)R, B, nx, dl, s_dec));
  } else {
    // recompute: logits = a2 W + b once more (the forward's GEMM, same tiling), gradient in the epilogue
    GemmArgs ga;
    ga.A = f.a2; ga.sa_m = h1; ga.sa_k = 1;
    ga.B = P + m.dec.d2.w; ga.sb_k = nx; ga.sb_n = 1;
    ga.C = dl; ga.ldc = nx;
    ga.M = (int)R; ga.N = nx; ga.K = h1; ga.bias = P + m.dec.d2.b;
    ga.dl_x = f.x; ga.dl_g = g; ga.dl_B = B;
    IWAE_TRY(gemm_f32(ga, s_dec));
  }
  // decoder MLP
  IWAE_TRY(ar.get(&da2, (size_t)R * h1)); IWAE_TRY(ar.get(&da1, (size_t)R * h1));
  IWAE_TRY(ar.get(&dzin, (size_t)R * m.dec_in()));
  IWAE_TRY(dense_bwd_w(ctx, m.dec.d2, f.a2, h1, R, dl, grads, s_dec, w_dec));
  float* dx_scratch = nullptr;       // split-K partial sums of the output layer's dX (training shapes only)
  if (R <= 4096) IWAE_TRY(ar.get(&dx_scratch, (size_t)BWDX_SPLITS * R * h1));
  IWAE_TRY(dense_bwd_x(P, m.dec.d2, dl, R, da2, h1, 0, f.a2, h1, s_dec, dx_scratch));
  IWAE_TRY(dense_bwd_w(ctx, m.dec.d1, f.a1, h1, R, da2, grads, s_dec, w_dec));
  bool dec_chain = false;
  if (chain) {
    ChainArgs a;                                             // da1 = (da2 W1^T) * (1 - a1^2) ; dzin = da1 W0^T
    a.M = (int)R; a.X = da2; a.ldx = h1; a.kx = h1; a.nlayers = 2;
    chain_bwd_dense(a.L[0], ctx->wt_buf, m.dec.d1, f.a1, da1);
    chain_bwd_dense(a.L[1], ctx->wt_buf, m.dec.d0, nullptr, dzin);
    if (dense_chain_fits(a)) { IWAE_TRY(dense_chain(a, s_dec)); dec_chain = true; }
  }
  if (!dec_chain) IWAE_TRY(dense_bwd_x(P, m.dec.d1, da2, R, da1, h1, 0, f.a1, h1, s_dec));
  IWAE_TRY(dense_bwd_w(ctx, m.dec.d0, f.zin, f.ld_zin, R, da1, grads, s_dec, w_dec));
  if (!dec_chain) IWAE_TRY(dense_bwd_x(P, m.dec.d0, da1, R, dzin, m.dec_in(), 0, nullptr, 0, s_dec));

  float* dz1_q = nullptr;
  if (m.two_layer()) {
    // lpz1z2 -> parameters of p(z1|z2)                                        iwae2.py:120
    float *dpmu, *dpa, *dz2_p, *dmu2, *da_2;
    IWAE_TRY(ar.get(&dpmu, (size_t)R * d1)); IWAE_TRY(ar.get(&dpa, (size_t)R * d1));
    IWAE_TRY(gauss_rows_bwd(g, f.z1, f.pmu, f.pstd, R, d1, dpmu, dpa, st));
    IWAE_TRY(ar.get(&dz2_p, (size_t)R * d2));
    IWAE_TRY(block_bwd(ctx, P, m.dec21, f.z2, d2, R, f.p1, f.p2, dpmu, dpa, grads, dz2_p, st, w_lat, chain));
    // z2 = mu2 + std2 eps2 ; lpz2 ; -lqz2z1                                   iwae2.py:65,116-118
    IWAE_TRY(ar.get(&dmu2, (size_t)R * d2)); IWAE_TRY(ar.get(&da_2, (size_t)R * d2));
    Latent2BwdArgs l2;
    l2.R = (int)R; l2.d = d2; l2.g = g; l2.dz_p = dz2_p; l2.z = f.z2; l2.eps = f.eps2; l2.std = f.std2;
    l2.dmu = dmu2; l2.da = da_2;
    IWAE_TRY(latent2_bwd(l2, st));
    IWAE_TRY(ar.get(&dz1_q, (size_t)R * d1));
    IWAE_TRY(block_bwd(ctx, P, m.enc2, f.z1, d1, R, f.g1, f.g2, dmu2, da_2, grads, dz1_q, st, w_lat, chain));
  }
  IWAE_TRY(lane_dep(ctx, s_dec, st));            // the first latent layer needs dz from both chains

  // first stochastic layer
  float *dmu1, *da_1, *dpmu_b = nullptr, *dpa_b = nullptr;
  IWAE_TRY(ar.get(&dmu1, (size_t)B * d1)); IWAE_TRY(ar.get(&da_1, (size_t)B * d1));
  Latent1BwdArgs l1;
  l1.B = B; l1.k = k; l1.d = d1; l1.beta = beta; l1.objective = objective; l1.B_norm = B_norm;
  l1.g = g; l1.al = f.al; l1.dz_dec = dzin; l1.ld_dz = m.dec_in(); l1.dz_q2 = dz1_q;
  l1.z = f.z1; l1.eps = f.eps1; l1.mu = f.mu1; l1.std = f.std1;
  if (m.two_layer()) { l1.prior_rows = 2; l1.pmu = f.pmu; l1.pstd = f.pstd; }
  else if (m.learned_prior()) {
    l1.prior_rows = 1; l1.pmu = f.pmu; l1.pstd = f.pstd;
    IWAE_TRY(ar.get(&dpmu_b, (size_t)B * d1)); IWAE_TRY(ar.get(&dpa_b, (size_t)B * d1));
    l1.dpmu = dpmu_b; l1.dpa = dpa_b;
  }
  l1.dmu = dmu1; l1.da = da_1;
  IWAE_TRY(latent1_bwd(l1, st));
  IWAE_TRY(block_bwd(ctx, P, m.enc1, f.xin, f.ld_xin, B, f.h1, f.h2, dmu1, da_1, grads, nullptr, st, w_lat, chain));
  if (m.learned_prior())
    IWAE_TRY(block_bwd(ctx, P, m.prior, f.yoh, c.n_cond, B, f.p1, f.p2, dpmu_b, dpa_b, grads, nullptr, st, w_dec, chain));
  if (w_dec) IWAE_TRY(lane_dep(ctx, w_dec, st));     // every weight gradient is in `grads` before the reduction / Adam
  if (w_lat) IWAE_TRY(lane_dep(ctx, w_lat, st));
  return 0;
}

struct AdamArgs { float *params, *m, *v; float lr, b1, b2, eps; int64_t t; };

static int grads_internal(iwae_ctx* ctx, const float* params, const float* x, const float* y, int B, int k, float beta,
                          int objective, int B_norm, const float* eps1, const float* eps2, uint64_t seed,
                          int64_t img_offset, float* grads, const iwae_outputs* outs, cudaStream_t st,
                          const AdamArgs* adam = nullptr) {
  const ModelP& m = ctx->model;
  IWAE_REQUIRE(grads || adam, "iwae_grads: null grads");
  IWAE_REQUIRE(objective >= IWAE_OBJ_VAE_ELBO && objective <= IWAE_OBJ_DREG, "unknown objective %d", objective);
  // iwae2.IWAE.call has no 'vae_elbo_kl' key -> KeyError in the reference (iwae2.py:154-167,173)
  IWAE_REQUIRE(!(m.two_layer() && objective == IWAE_OBJ_VAE_ELBO_KL), "KeyError: 'vae_elbo_kl' (two stochastic layers)");
  IWAE_REQUIRE(!(objective == IWAE_OBJ_DREG && m.cfg.kind != IWAE_KIND_IWAE1), "DReG is defined for IWAE1 only (task02.py)");
  if (B_norm <= 0) B_norm = B;
  Fwd f; FwdOpts opt;
  opt.keep_for_bwd = true;
  opt.want_dreg = (objective == IWAE_OBJ_DREG);
  static const bool no_lanes = getenv("IWAE_NO_LANES") != nullptr;
  opt.lanes = !no_lanes;
  opt.objective = objective; opt.B_norm = B_norm;
  // Training shapes: forward + backward + update as ONE persistent kernel (kernels_f32.cu, "train-step megakernel"): the
  // launches below are recorded with their stream / event structure and run level by level behind grid barriers.
  static const bool mega_on = getenv("IWAE_MEGA") != nullptr;
  const bool mega = mega_on && opt.lanes && !chain_enabled() && (int64_t)B * k <= MEGA_MAX_ROWS && k <= MEGA_MAX_K;
  struct MegaGuard { iwae_ctx* c; bool armed; ~MegaGuard() { if (armed) mega_abort(c); } } guard{ctx, false};
  if (mega) { IWAE_TRY(lanes_init(ctx)); IWAE_TRY(mega_begin(ctx, st)); guard.armed = true; }
  // transposed weights for the backward chains: built on lane 1 while the forward pass runs
  const bool chain = opt.lanes && bwd_chain_ok(m, (int64_t)B * k);
  cudaEvent_t wt_done = nullptr;
  if (chain) {
    IWAE_TRY(lanes_init(ctx));
    IWAE_TRY(lane_dep(ctx, st, ctx->lane[1]));
    IWAE_TRY(launch_transposes(ctx, params, ctx->lane[1]));
    if (!ctx->wt_ev) CUDA_TRY(cudaEventCreateWithFlags(&ctx->wt_ev, cudaEventDisableTiming));
    wt_done = ctx->wt_ev;
    CUDA_TRY(cudaEventRecord(wt_done, ctx->lane[1]));
  }
  IWAE_TRY(forward_internal(ctx, params, x, y, B, k, beta, eps1, eps2, seed, img_offset, outs, opt, f, st));
  if (wt_done) CUDA_TRY(cudaStreamWaitEvent(st, wt_done, 0));
  // backward into the zeroed [GRAD_SPLITS][n_params] workspace, then one fixed-order reduction
  float* ws;
  const int64_t n = m.n_params;
  IWAE_TRY(ctx->arena.get(&ws, (size_t)GRAD_SPLITS * n));
  // recorded: a level-0 operation beside the forward pass (levels are global, no edge needed); else a memset in stream order
  IWAE_TRY(zero_f32(ws, (int64_t)GRAD_SPLITS * n, mega ? MEGA_SIDE_STREAM : st));
  g_grad_stride = (int)n;
  IWAE_TRY(backward_internal(ctx, params, f, beta, objective, B_norm, ws, st, opt.lanes, chain));
  if (adam) IWAE_TRY(adam_from_partials(adam->params, ws, GRAD_SPLITS, adam->m, adam->v, n, adam->lr, adam->b1, adam->b2,
                                        adam->eps, adam->t, st));
  else IWAE_TRY(grad_reduce(ws, GRAD_SPLITS, n, grads, st));
  if (mega) { guard.armed = false; IWAE_TRY(mega_flush(ctx, st)); }
  return 0;
}

extern "C" int iwae_grads(iwae_ctx* ctx, const float* params, const float* x, const float* y, int32_t B, int32_t k,
                          float beta, int32_t objective, int32_t B_norm, const float* eps1, const float* eps2,
                          uint64_t seed, int64_t img_offset, float* grads, const iwae_outputs* outs, void* stream) {
  IWAE_REQUIRE(ctx, "iwae_grads: null ctx");
  ctx->arena.reset();
  return grads_internal(ctx, params, x, y, B, k, beta, objective, B_norm, eps1, eps2, seed, img_offset, grads, outs,
                        (cudaStream_t)stream);
}

extern "C" int iwae_adam_step(float* params, const float* grads, float* m, float* v, int64_t n, float lr, float beta1,
                              float beta2, float eps, int64_t t, void* stream) {
  IWAE_REQUIRE(params && grads && m && v, "iwae_adam_step: null argument");
  IWAE_REQUIRE(t >= 1, "iwae_adam_step: t is the step count after increment (>= 1)");
  return adam_f32(params, grads, m, v, n, lr, beta1, beta2, eps, t, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// train step: eager, or replayed from a CUDA graph
// ---------------------------------------------------------------------------------------------
// The k=50 step is ~59 dependent small kernels; launched in stream order they leave ~2 us between kernels that a
// graph does not (scripts/train_graph_probe.py: 0.62 -> 0.51 ms per step).  The graph is captured inside the
// library on the SECOND consecutive call with the same signature (the first, eager one has grown the arena, so the
// capture allocates nothing), on a library-owned stream (the caller's may be the legacy default stream, which cannot
// be captured), against library-owned copies of x / y and of the requested outputs: the caller's pointers change
// from call to call.  Per-step scalars (Philox seed, Adam's lr_t) live in a device struct the kernels read.
// Injected eps (the parity path) and IWAE_NO_GRAPH=1 stay eager.
static long long out_field_len(int f, const iwae_config& c, int B, int k) {
  const long long R = (long long)B * k;
  switch (f) {
    case 0: case 1: case 2: case 3: case 4: case 5: case 6: return R;             // lpxz .. al
    case 7: case 8: case 9: case 10: case 11: return B;                           // iwae, vae, eq14, kl, dreg
    case 12: return (long long)B * c.d1;                                          // snis_z1
    case 13: return (long long)B * std::max(c.d2, 1);                             // snis_z2
    case 14: return R * c.d1;                                                     // z1
    case 15: return R * std::max(c.d2, 1);                                        // z2
    case 16: return R * c.n_x;                                                    // logits
    case 17: case 18: return (long long)B * c.d1;                                 // q_mu, q_std
    default: return 8;                                                            // scalars
  }
}
constexpr int N_OUT_FIELDS = (int)(sizeof(iwae_outputs) / sizeof(float*));
static_assert(N_OUT_FIELDS == 20, "iwae_outputs changed: update out_field_len");

static void train_graph_release(iwae_ctx* ctx) {
  TrainGraph& g = ctx->tg;
  if (g.exec) cudaGraphExecDestroy(g.exec);
  if (g.x_buf) cudaFree(g.x_buf);
  if (g.y_buf) cudaFree(g.y_buf);
  if (g.out_mem) cudaFree(g.out_mem);
  g = TrainGraph();
  if (ctx->dyn_dev) { cudaFree(ctx->dyn_dev); ctx->dyn_dev = nullptr; }
  if (ctx->ev_gin) { cudaEventDestroy(ctx->ev_gin); ctx->ev_gin = nullptr; }
  if (ctx->ev_gout) { cudaEventDestroy(ctx->ev_gout); ctx->ev_gout = nullptr; }
  if (ctx->gstream) { cudaStreamDestroy(ctx->gstream); ctx->gstream = nullptr; }
}

// B_norm: the batch size the objective's mean_b divides by (B, or the GLOBAL batch of a data-parallel step).
// dp: data-parallel step - local gradient (already divided by the global batch) -> ONE NCCL sum all-reduce on the
// step's own stream (so it is captured inside the step's CUDA graph) -> the same fused reduce + Keras-Adam kernel.
static int train_step_eager(iwae_ctx* ctx, float* params, float* adam_m, float* adam_v, const float* x, const float* y,
                            int B, int k, float beta, int objective, const float* eps1, const float* eps2, uint64_t seed,
                            int64_t img_offset, float lr, float beta1, float beta2, float adam_eps, int64_t t,
                            const iwae_outputs* outs, cudaStream_t st, int B_norm = 0, bool dp = false) {
  ctx->arena.reset();
  if (!dp) {
    AdamArgs ad = {params, adam_m, adam_v, lr, beta1, beta2, adam_eps, t};
    return grads_internal(ctx, params, x, y, B, k, beta, objective, B_norm > 0 ? B_norm : B, eps1, eps2, seed, img_offset,
                          nullptr, outs, st, &ad);
  }
  float* g = comm_grad_buffer(ctx);
  IWAE_REQUIRE(g, "iwae_dp_train_step: no communicator (iwae_comm_init)");
  const int64_t n = ctx->model.n_params;
  if (B > 0)
    IWAE_TRY(grads_internal(ctx, params, x, y, B, k, beta, objective, B_norm, eps1, eps2, seed, img_offset, g, outs, st));
  else
    CUDA_TRY(cudaMemsetAsync(g, 0, sizeof(float) * (size_t)n, st));      // a rank without images still joins the all-reduce
  IWAE_TRY(comm_allreduce_f32(ctx, g, n, st));
  return adam_from_partials(params, g, 1, adam_m, adam_v, n, lr, beta1, beta2, adam_eps, t, st);
}

// capture the step for the signature in g.cand; returns 0 and g.valid on success
static int train_graph_capture(iwae_ctx* ctx, float* params, float* adam_m, float* adam_v, bool has_y, int B, int k,
                               float beta, int objective, const iwae_outputs* outs, int B_norm, bool dp) {
  TrainGraph& g = ctx->tg;
  const iwae_config& c = ctx->model.cfg;
  if (!ctx->gstream) CUDA_TRY(cudaStreamCreateWithFlags(&ctx->gstream, cudaStreamNonBlocking));
  if (!ctx->ev_gin) { CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_gin, cudaEventDisableTiming)); CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_gout, cudaEventDisableTiming)); }
  if (!ctx->dyn_dev) CUDA_TRY(cudaMalloc((void**)&ctx->dyn_dev, sizeof(StepDyn)));
  IWAE_TRY(grow_dev(&g.x_buf, &g.x_cap, (size_t)B * c.n_x * sizeof(float)));
  IWAE_TRY(grow_dev(&g.y_buf, &g.y_cap, (size_t)B * sizeof(float) + 16));
  // shadow outputs: one block, each requested field 256-byte aligned
  size_t need = 0;
  float* const* user = reinterpret_cast<float* const*>(outs);
  for (int f = 0; f < N_OUT_FIELDS; ++f)
    if (outs && user[f]) need += ((size_t)out_field_len(f, c, B, k) * sizeof(float) + 255) & ~(size_t)255;
  IWAE_TRY(grow_dev(&g.out_mem, &g.out_cap, std::max(need, (size_t)256)));
  g.shadow = iwae_outputs{};
  float** sh = reinterpret_cast<float**>(&g.shadow);
  size_t off = 0;
  for (int f = 0; f < N_OUT_FIELDS; ++f)
    if (outs && user[f]) {
      sh[f] = reinterpret_cast<float*>(reinterpret_cast<char*>(g.out_mem) + off);
      off += ((size_t)out_field_len(f, c, B, k) * sizeof(float) + 255) & ~(size_t)255;
    }
  if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
  g.valid = false;

  cudaGraph_t graph = nullptr;
  const long long launches0 = g_iwae_launches;
  CUDA_TRY(cudaStreamBeginCapture(ctx->gstream, cudaStreamCaptureModeThreadLocal));
  g_step_dyn = ctx->dyn_dev;
  const int rc = train_step_eager(ctx, params, adam_m, adam_v, g.x_buf, has_y ? g.y_buf : nullptr, B, k, beta, objective,
                                  nullptr, nullptr, 0, 0, 0.f, 0.f, 0.f, 0.f, 1, &g.shadow, ctx->gstream, B_norm, dp);
  g_step_dyn = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(ctx->gstream, &graph);
  g.launches = g_iwae_launches - launches0;
  g_iwae_launches = launches0;
  if (rc != 0 || ce != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    g.disabled = true;
    return 0;       // not an error: the caller sees !g.valid and runs this (and every later) step eagerly
  }
  const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) { cudaGetLastError(); g.exec = nullptr; g.disabled = true; return 0; }
  memcpy(g.key, g.cand, sizeof(g.key));
  g.valid = true;
  return 0;
}

static int train_step_impl(iwae_ctx* ctx, float* params, float* adam_m, float* adam_v, const float* x,
                           const float* y, int32_t B, int32_t k, float beta, int32_t objective, const float* eps1,
                           const float* eps2, uint64_t seed, int64_t img_offset, float lr, float beta1, float beta2,
                           float adam_eps, int64_t t, const iwae_outputs* outs, void* stream, int B_norm, bool dp) {
  IWAE_REQUIRE(ctx && params && adam_m && adam_v, "iwae_train_step: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  IWAE_REQUIRE(t >= 1, "iwae_train_step: t is the step count after increment (>= 1)");
  TrainGraph& g = ctx->tg;
  static const bool no_graph = getenv("IWAE_NO_GRAPH") != nullptr;
  const bool graphable = !no_graph && !g.disabled && !eps1 && !eps2 && x && B > 0 && k > 0;
  if (!graphable)
    return train_step_eager(ctx, params, adam_m, adam_v, x, y, B, k, beta, objective, eps1, eps2, seed, img_offset, lr,
                            beta1, beta2, adam_eps, t, outs, st, B_norm, dp);

  // signature: everything the captured kernels hold by value or by address
  unsigned long long key[12] = {};
  key[0] = (unsigned long long)(uintptr_t)params; key[1] = (unsigned long long)(uintptr_t)adam_m;
  key[2] = (unsigned long long)(uintptr_t)adam_v; key[3] = (unsigned long long)(unsigned)B;
  key[4] = (unsigned long long)(unsigned)k; key[5] = (unsigned long long)(unsigned)objective;
  { unsigned int bb; memcpy(&bb, &beta, 4); key[6] = bb; }
  key[7] = y ? 1 : 0;
  float* const* user = reinterpret_cast<float* const*>(outs);
  unsigned long long mask = 0;
  for (int f = 0; f < N_OUT_FIELDS; ++f) if (outs && user[f]) mask |= 1ull << f;
  key[8] = mask;
  key[9] = (unsigned long long)(unsigned)B_norm; key[10] = dp ? 1 : 0;

  if (!(g.valid && memcmp(key, g.key, sizeof(key)) == 0)) {
    if (g.have_cand && memcmp(key, g.cand, sizeof(key)) == 0) {
      // second consecutive call with this signature: capture (the eager call before it sized the arena)
      IWAE_TRY(train_graph_capture(ctx, params, adam_m, adam_v, y != nullptr, B, k, beta, objective, outs, B_norm, dp));
    } else {
      memcpy(g.cand, key, sizeof(key)); g.have_cand = true;
    }
    if (!(g.valid && memcmp(key, g.key, sizeof(key)) == 0))
      return train_step_eager(ctx, params, adam_m, adam_v, x, y, B, k, beta, objective, eps1, eps2, seed, img_offset, lr,
                              beta1, beta2, adam_eps, t, outs, st, B_norm, dp);
  }

  // replay: order the library stream after the caller's, refresh inputs and scalars, launch, hand the outputs back
  const iwae_config& c = ctx->model.cfg;
  cudaStream_t gs = ctx->gstream;
  CUDA_TRY(cudaEventRecord(ctx->ev_gin, st));
  CUDA_TRY(cudaStreamWaitEvent(gs, ctx->ev_gin, 0));
  CUDA_TRY(cudaMemcpyAsync(g.x_buf, x, (size_t)B * c.n_x * sizeof(float), cudaMemcpyDeviceToDevice, gs));
  if (y) CUDA_TRY(cudaMemcpyAsync(g.y_buf, y, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, gs));
  StepDyn d;
  d.seed = seed; d.img_offset = img_offset;
  d.lr_t = (float)((double)lr * sqrt(1.0 - pow((double)beta2, (double)t)) / (1.0 - pow((double)beta1, (double)t)));
  d.b1 = beta1; d.b2 = beta2; d.eps = adam_eps;
  IWAE_TRY(set_step_dyn(ctx->dyn_dev, d, gs));
  CUDA_TRY(cudaGraphLaunch(g.exec, gs));
  g_iwae_launches += g.launches;
  CopyTable tb; tb.count = 0;
  float* const* sh = reinterpret_cast<float* const*>(&g.shadow);
  for (int f = 0; f < N_OUT_FIELDS; ++f)
    if (outs && user[f]) { tb.src[tb.count] = sh[f]; tb.dst[tb.count] = user[f]; tb.n[tb.count] = out_field_len(f, c, B, k); ++tb.count; }
  IWAE_TRY(copy_table(tb, gs));
  CUDA_TRY(cudaEventRecord(ctx->ev_gout, gs));
  CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_gout, 0));
  return 0;
}

extern "C" int iwae_train_step(iwae_ctx* ctx, float* params, float* adam_m, float* adam_v, const float* x,
                               const float* y, int32_t B, int32_t k, float beta, int32_t objective, const float* eps1,
                               const float* eps2, uint64_t seed, int64_t img_offset, float lr, float beta1, float beta2,
                               float adam_eps, int64_t t, const iwae_outputs* outs, void* stream) {
  return train_step_impl(ctx, params, adam_m, adam_v, x, y, B, k, beta, objective, eps1, eps2, seed, img_offset, lr, beta1,
                         beta2, adam_eps, t, outs, stream, 0, false);
}

extern "C" int iwae_dp_train_step(iwae_ctx* ctx, float* params, float* adam_m, float* adam_v, const float* x_local,
                                  const float* y_local, int32_t B_local, int32_t B_global, int32_t k, float beta,
                                  int32_t objective, uint64_t seed, int64_t img_offset, float lr, float beta1, float beta2,
                                  float adam_eps, int64_t t, const iwae_outputs* outs, void* stream) {
  IWAE_REQUIRE(ctx && comm_grad_buffer(ctx), "iwae_dp_train_step: no communicator (iwae_comm_init)");
  IWAE_REQUIRE(B_local >= 0 && B_global >= B_local && B_global > 0, "iwae_dp_train_step: bad batch sizes %d / %d", B_local,
               B_global);
  return train_step_impl(ctx, params, adam_m, adam_v, x_local, y_local, B_local, k, beta, objective, nullptr, nullptr, seed,
                         img_offset, lr, beta1, beta2, adam_eps, t, outs, stream, B_global, true);
}

extern "C" int iwae_philox_normal(uint64_t seed, int32_t stream_id, int64_t img_offset, int32_t B, int32_t k, int32_t d,
                                  float* out, void* stream) {
  IWAE_REQUIRE(out && B > 0 && k > 0 && d > 0, "iwae_philox_normal: bad argument");
  return philox_fill(seed, stream_id, img_offset, B, k, d, out, (cudaStream_t)stream);
}

extern "C" int iwae_decode(iwae_ctx* ctx, const float* params, const float* z, const float* x, const float* y,
                           int32_t R, int32_t B, float* logits, float* lpxz, void* stream) {
  IWAE_REQUIRE(ctx && params && z, "iwae_decode: null argument");
  IWAE_REQUIRE(R > 0 && B > 0, "iwae_decode: R and B must be positive");
  IWAE_REQUIRE(!lpxz || x, "iwae_decode: lpxz needs x");
  cudaStream_t st = (cudaStream_t)stream;
  const ModelP& m = ctx->model;
  const iwae_config& c = m.cfg;
  Arena& ar = ctx->arena;
  ar.reset();
  const int nc = m.cond() ? c.n_cond : 0;
  IWAE_REQUIRE(!nc || y, "iwae_decode: CIWAE needs labels y");
  const float* zin = z; int64_t ld = c.d1;
  if (nc) {
    float* t; IWAE_TRY(ar.get(&t, (size_t)R * (c.d1 + nc)));
    IWAE_TRY(concat_onehot(z, c.d1, c.d1, y, B, R, nc, t, st));
    zin = t; ld = c.d1 + nc;
  }
  float *a1, *a2, *bpart = nullptr;
  IWAE_TRY(ar.get(&a1, (size_t)R * c.h1)); IWAE_TRY(ar.get(&a2, (size_t)R * c.h1));
  IWAE_TRY(dense_fwd(params, m.dec.d0, zin, ld, R, ACT_TANH, a1, st));
  IWAE_TRY(dense_fwd(params, m.dec.d1, a1, c.h1, R, ACT_TANH, a2, st));
  GemmArgs g;
  g.A = a2; g.sa_m = c.h1; g.sa_k = 1;
  g.B = params + m.dec.d2.w; g.sb_k = c.n_x; g.sb_n = 1;
  g.C = logits; g.ldc = c.n_x; g.M = R; g.N = c.n_x; g.K = c.h1; g.bias = params + m.dec.d2.b;
  const int nt = gemm_f32_ntiles_n(c.n_x, 0);
  if (lpxz) {
    IWAE_TRY(ar.get(&bpart, (size_t)R * nt));
    g.bern_x = x; g.bern_B = B; g.bern_partial = bpart;
  }
  IWAE_REQUIRE(logits || lpxz, "iwae_decode: nothing to compute");
  IWAE_TRY(gemm_f32(g, st));
  if (lpxz) IWAE_TRY(bern_reduce_partials(bpart, R, nt, lpxz, st));
  return 0;
}

// BasicBlock.call (iwae1.py:36-44) of one of the model's four blocks on caller-provided rows: what
// IWAE.sample / conditional_sample (iwae2.py:184-213) and CIWAE.sample (task04.py:190-209) evaluate outside `call`.
extern "C" int iwae_block_forward(iwae_ctx* ctx, const float* params, int32_t block, const float* in, int32_t R,
                                  float* mu, float* sd, void* stream) {
  IWAE_REQUIRE(ctx && params && in && mu && sd, "iwae_block_forward: null argument");
  IWAE_REQUIRE(R > 0, "iwae_block_forward: R must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  const ModelP& m = ctx->model;
  const BlockP* blk = nullptr;
  switch (block) {
    case IWAE_BLOCK_ENCODER: blk = &m.enc1; break;
    case IWAE_BLOCK_ENCODER2: IWAE_REQUIRE(m.two_layer(), "iwae_block_forward: model has no q(z2|z1) block"); blk = &m.enc2; break;
    case IWAE_BLOCK_DECODER21: IWAE_REQUIRE(m.two_layer(), "iwae_block_forward: model has no p(z1|z2) block"); blk = &m.dec21; break;
    case IWAE_BLOCK_PRIOR: IWAE_REQUIRE(m.learned_prior(), "iwae_block_forward: model has no learned prior block"); blk = &m.prior; break;
    default: IWAE_REQUIRE(false, "iwae_block_forward: unknown block %d", block);
  }
  Arena& ar = ctx->arena;
  ar.reset();
  float *h1, *h2;
  IWAE_TRY(ar.get(&h1, (size_t)R * blk->l1.n_out)); IWAE_TRY(ar.get(&h2, (size_t)R * blk->l2.n_out));
  return block_fwd_f32(params, *blk, in, blk->l1.n_in, R, h1, h2, mu, sd, st);
}

extern "C" int iwae_sir(const float* al, int32_t L, const double* uniforms, int32_t n_uniforms, int32_t size,
                        int64_t* out_idx, int32_t* out_consumed, void* stream) {
  IWAE_REQUIRE(al && uniforms && out_idx && out_consumed, "iwae_sir: null argument");
  IWAE_REQUIRE(L > 0 && size > 0, "iwae_sir: L and size must be positive");
  return sir_launch(al, L, uniforms, n_uniforms, size, out_idx, out_consumed, (cudaStream_t)stream);
}

extern "C" int iwae_softmax_f32(const float* log_w, int32_t L, float* out, void* stream) {
  IWAE_REQUIRE(log_w && out && L > 0, "iwae_softmax_f32: bad argument");
  return softmax_vec_f32(log_w, L, out, (cudaStream_t)stream);
}

extern "C" int iwae_bernoulli_sample(const float* p, int64_t n, uint64_t seed, uint64_t epoch, float* out,
                                     void* stream) {
  IWAE_REQUIRE(p && out && n >= 0, "iwae_bernoulli_sample: bad argument");
  return bernoulli_sample_f32(p, n, seed, epoch, out, (cudaStream_t)stream);
}
