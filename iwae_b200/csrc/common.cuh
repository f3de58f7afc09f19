// Internal declarations shared by the translation units of libiwae_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <vector>
#include "../../include/iwae_b200.h"

void iwae_set_error(const char* fmt, ...);

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      iwae_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__));    \
      return 1;                                                                                 \
    }                                                                                           \
  } while (0)

extern long long g_iwae_launches;
// after every kernel launch: count it and surface launch-configuration errors
bool mega_recording();   // the launch wrappers of kernels_f32.cu are recording a train-step program instead of launching
#define LAUNCH_CHECK()                                                                          \
  do {                                                                                          \
    if (!mega_recording()) ++g_iwae_launches;                                                   \
    CUDA_TRY(cudaGetLastError());                                                               \
  } while (0)

#define IWAE_REQUIRE(cond, ...)                                                                 \
  do {                                                                                          \
    if (!(cond)) {                                                                              \
      iwae_set_error(__VA_ARGS__);                                                              \
      return 2;                                                                                 \
    }                                                                                           \
  } while (0)

#define IWAE_TRY(expr)                                                                          \
  do {                                                                                          \
    int r__ = (expr);                                                                           \
    if (r__ != 0) return r__;                                                                   \
  } while (0)

// ---- model description: offsets into the flat parameter buffer (Keras order) ----
struct DenseP {
  int64_t w = 0, b = 0;   // offsets (floats)
  int n_in = 0, n_out = 0;
};
struct BlockP { DenseP l1, l2, lmu, lstd; };   // BasicBlock (iwae1.py:24-44)
struct Mlp3P { DenseP d0, d1, d2; };           // decoder Sequential (iwae1.py:70-77)

struct ModelP {
  iwae_config cfg{};
  BlockP enc1;    // encoder.encode_x_to_z / encode_x_to_z1
  BlockP enc2;    // encoder.encode_z1_to_z2        (IWAE2)
  BlockP dec21;   // decoder.decode_z2_to_z1        (IWAE2)
  Mlp3P dec;      // decoder.decode_z_to_x / decode_z1_to_x
  BlockP prior;   // conditional_prior_network      (CIWAE)
  int64_t n_params = 0;
  std::vector<int64_t> t_off;
  std::vector<int32_t> t_rows, t_cols;
  bool two_layer() const { return cfg.kind == IWAE_KIND_IWAE2; }
  bool cond() const { return cfg.kind == IWAE_KIND_CIWAE || cfg.kind == IWAE_KIND_CIWAE_STD; }
  bool learned_prior() const { return cfg.kind == IWAE_KIND_CIWAE; }
  int enc_in() const { return cfg.n_x + (cond() ? cfg.n_cond : 0); }
  int dec_in() const { return cfg.d1 + (cond() ? cfg.n_cond : 0); }
};

// ---- grow-only scratch arena (device memory owned by the context) ----
struct Arena {
  struct Blk { char* p; size_t cap; size_t off; };
  std::vector<Blk> blks;
  int alloc(void** out, size_t bytes);
  void reset() { for (auto& b : blks) b.off = 0; }
  void release();
  template <class T> int get(T** out, size_t n) { return alloc((void**)out, n * sizeof(T)); }
};

struct TcPack;   // packed fp16 weights + program for the fused tcgen05 evaluation (eval_tc.cu)
struct DpComm;   // library-owned NCCL communicator + reduced-gradient buffer (comm.cu)

// Per-step scalars of a train step, device-resident so that a captured CUDA graph of the step can be replayed with
// new values (kernels read them through g_step_dyn when it is non-null, else from their by-value arguments).
struct StepDyn {
  unsigned long long seed;
  long long img_offset;
  float lr_t, b1, b2, eps;
};
extern thread_local const StepDyn* g_step_dyn;

// One cached CUDA graph of iwae_train_step (see abi.cu): signature it was captured for, library-owned input and
// output buffers (the caller's pointers change from call to call), and the captured launch count.
struct TrainGraph {
  bool valid = false;
  unsigned long long key[12] = {};          // signature of the captured step
  unsigned long long cand[12] = {};         // signature seen on the previous (eager) call
  bool have_cand = false;
  bool disabled = false;                    // capture failed once: stay eager
  cudaGraphExec_t exec = nullptr;
  float* x_buf = nullptr; float* y_buf = nullptr; size_t x_cap = 0, y_cap = 0;
  float* out_mem = nullptr; size_t out_cap = 0;
  iwae_outputs shadow{};                    // the graph writes here; a gather kernel copies to the caller's pointers
  long long launches = 0;
};

struct iwae_ctx {
  ModelP model;
  Arena arena;
  int device = 0;
  int sm_count = 148;
  TcPack* tc = nullptr;
  struct EncPack* enc = nullptr;    // packed hi/lo weights of the tensor-core encoder pre-pass (encoder_tc.cu)
  DpComm* dp = nullptr;
  uint64_t params_version = 0;     // iwae_set_params_version: content version of the caller's parameter buffer (0 = unknown)
  // persistent device staging for iwae_llh_eval / iwae_llh_eval_host (outlives arena resets)
  float* per_buf = nullptr; size_t per_cap = 0;       // per-example results
  float* stage_x = nullptr; size_t stage_x_cap = 0;   // x (and y) copied from the host
  float* stage_aux = nullptr; size_t stage_aux_cap = 0;
  cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr; bool ev_valid = false;   // around the fused eval kernel
  // train-step graph replay
  TrainGraph tg;
  StepDyn* dyn_dev = nullptr;
  cudaStream_t gstream = nullptr;
  cudaEvent_t ev_gin = nullptr, ev_gout = nullptr;
  // train step: side streams ("lanes") for the branches of the step that do not depend on each other - the decoder
  // p(x|z1) beside the q(z2|z1) / p(z1|z2) blocks, every weight gradient dW = X^T dY beside the dX chain - joined with
  // events (inside a captured step they become parallel branches of the CUDA graph)
  cudaStream_t lane[3] = {nullptr, nullptr, nullptr};
  struct MegaState* mega = nullptr;                // train-step megakernel: recorder, program buffers, barrier counter
  cudaEvent_t wt_ev = nullptr;
  float* wt_buf = nullptr; size_t wt_cap = 0;      // transposed weights of the fused backward chains (rebuilt every step)
  cudaEvent_t lane_ev[16] = {};
  int lane_ev_next = 0;
};

// ---- fp32 kernels (kernels_f32.cu) ----
enum { ACT_NONE = 0, ACT_TANH = 1, ACT_EXP_EPS = 2 };

struct GemmArgs {
  const float* A = nullptr; int64_t sa_m = 0, sa_k = 0;   // A(m,k) = A[m*sa_m + k*sa_k]
  const float* B = nullptr; int64_t sb_k = 0, sb_n = 0;   // B(k,n) = B[k*sb_k + n*sb_n]
  float* C = nullptr; int64_t ldc = 0;                    // C(m,n) row-major (may be null with bern)
  int M = 0, N = 0, K = 0;
  const float* bias = nullptr;                            // [N]
  int act = ACT_NONE;
  int accumulate = 0;                                     // v += C_old before mul_h
  const float* mul_h = nullptr; int64_t ld_h = 0;         // v *= (1 - H(m,n)^2)   (tanh backward)
  const float* bern_x = nullptr; int bern_B = 0;          // fused Bernoulli log-lik: x[(m % B), n]
  float* bern_partial = nullptr;                          // [M, n_tiles_n] partial row sums
  const float* dl_x = nullptr; const float* dl_g = nullptr; int dl_B = 0;
                                                          // backward recompute of the output layer: the stored value is
                                                          // d(objective)/d logit = g[m] * (x[(m % B), n] - sigmoid(v))
  int small_tile = 0;                                     // 32x32 tiles (more CTAs for skinny outputs)
  int split_k = 1;                                        // >1: gridDim.z partial products, C + z*split_stride (plain products only)
  int64_t split_stride = 0;
  int ones_row = 0;                                       // A gets a virtual row m = M of ones: C row M = column sums of B
                                                          // (dW = X^T dY then yields db = colsum(dY) in the same launch)
};
int gemm_f32(const GemmArgs& a, cudaStream_t st);
int gemm_f32_ntiles_n(int N, int small_tile);

int colsum_f32(const float* A, int64_t lda, int M, int N, float* out, cudaStream_t st, int splits = 1,
               int64_t split_stride = 0);   // splits > 1: per-row-range partial sums at out + s*split_stride
int concat_onehot(const float* x, int64_t ldx, int n, const float* y, int B, int rows, int nc,
                  float* out, cudaStream_t st);   // out[r] = [x[r, :n] | onehot(y[r % B])]
int onehot_f32(const float* y, int B, int nc, float* out, cudaStream_t st);

// z = mu + std*eps.  bcast: mu/std are [B,d] shared by all k samples, else [R,d].
// eps_in null -> Philox (seed, stream_id, img_offset); eps_out (optional) receives the noise used.
int sample_latent(const float* mu, const float* std, int bcast, const float* eps_in, float* eps_out,
                  float* z, int B, int k, int d, uint64_t seed, int stream_id, int64_t img_offset,
                  cudaStream_t st);
int philox_fill(uint64_t seed, int stream_id, int64_t img_offset, int B, int k, int d, float* out,
                cudaStream_t st);

// out[r] = sum_j log N(z[r,j]; mu, std) in TFP's op order; mu==null -> standard normal.
// std_extra is added to std (DReG's second +1e-6, task02.py:61-63).
int normal_logp_rows(const float* z, const float* mu, const float* std, int bcast, float std_extra,
                     int B, int k, int d, float* out, cudaStream_t st);
// several independent normal_logp_rows jobs in one launch (blockIdx.y = job)
struct NlpJob { const float* z; const float* mu; const float* sd; int bcast; float std_extra; int d; float* out; };
struct NlpJobs { NlpJob job[6]; int count = 0; };
int normal_logp_rows_multi(const NlpJobs& jobs, int B, int k, cudaStream_t st);
int kl_rows(const float* mu, const float* std, int B, int d, float* out, cudaStream_t st);
int bern_reduce_partials(const float* partial, int R, int nt, float* lpxz, cudaStream_t st);
int bern_logp_rows(const float* logits, const float* x, int R, int B, int n, float* lpxz, cudaStream_t st);

struct LogwArgs {
  int B = 0, k = 0; float beta = 1.f; int two_layer = 0;
  const float *lpxz = nullptr, *lpz = nullptr, *lq1 = nullptr, *lq2 = nullptr, *lp12 = nullptr;
  const float* lq_stopped = nullptr;        // DReG
  const float *z1 = nullptr, *z2 = nullptr; int d1 = 0, d2 = 0;
  float *log_w = nullptr, *al = nullptr, *iwae = nullptr, *vae = nullptr, *eq14 = nullptr, *dreg = nullptr;
  float *snis1 = nullptr, *snis2 = nullptr;
  int al_in_smem = 0;                       // set by logw_reduce
  // train step: the backward's per-row coefficient g[r] = d(-objective)/d(log_w[r]) (A.4) written here instead of by a
  // kernel of its own: -al/B_norm for the importance-weighted objectives, -1/(k B_norm) for the VAE ones (g_vae)
  float* g_out = nullptr; int g_vae = 0; int B_norm = 0;
};
int logw_reduce(const LogwArgs& a, cudaStream_t st);
struct ScalarArgs {
  int B = 0; int64_t R = 0; float beta = 1.f;
  const float *iwae = nullptr, *vae = nullptr, *eq14 = nullptr, *kl = nullptr, *dreg = nullptr;
  const float *lpxz = nullptr, *lpz = nullptr, *lq1 = nullptr;
  float* out = nullptr;   // [8]
  // large calls (val_step on the whole test set: k*B = 500,000 rows): `partial` [SCALAR_PARTS][8] receives per-CTA sums of
  // contiguous slices first; the single finalising CTA then adds SCALAR_PARTS values per mean in a fixed order
  float* partial = nullptr;
};
constexpr int SCALAR_PARTS = 296;
constexpr int64_t SCALAR_PARTS_MIN_ROWS = 32768;     // below this one CTA reads everything itself
int finalize_scalars(const ScalarArgs& a, cudaStream_t st);
int sum_to_double(const float* v, int64_t n, double* acc, int accumulate, cudaStream_t st);

// backward helpers
int row_coef(const float* al, int B, int k, int objective, int B_norm, float* g, cudaStream_t st);
int dlogits_f32(const float* logits, const float* x, const float* g, int R, int B, int n, float* dl,
                cudaStream_t st);
struct Latent1BwdArgs {   // first stochastic layer: mu/std broadcast over k
  int B = 0, k = 0, d = 0; float beta = 1.f; int objective = 0; int B_norm = 1;
  const float *g = nullptr, *al = nullptr;
  const float* dz_dec = nullptr; int64_t ld_dz = 0;    // [R, ld] gradient reaching z through the decoder
  const float* dz_q2 = nullptr;                        // [R, d] through q(z2|z1) (IWAE2) or null
  const float *z = nullptr, *eps = nullptr, *mu = nullptr, *std = nullptr;
  const float *pmu = nullptr, *pstd = nullptr; int prior_rows = 0;  // 0: N(0,1); 1: [B,d] (CIWAE); 2: [R,d] (IWAE2 p(z1|z2))
  float *dmu = nullptr, *da = nullptr;                 // [B,d] grads wrt mu and the std pre-activation
  float *dpmu = nullptr, *dpa = nullptr;               // prior block grads: [B,d] (CIWAE) or [R,d] (IWAE2)
};
int latent1_bwd(const Latent1BwdArgs& a, cudaStream_t st);
struct Latent2BwdArgs {   // second stochastic layer of IWAE2: everything per row
  int R = 0, d = 0; const float* g = nullptr;
  const float *dz_p = nullptr, *z = nullptr, *eps = nullptr, *std = nullptr;
  float *dmu = nullptr, *da = nullptr;
};
int latent2_bwd(const Latent2BwdArgs& a, cudaStream_t st);
int gauss_rows_bwd(const float* g, const float* z, const float* pmu, const float* pstd, int64_t R, int d,
                   float* dpmu, float* dpa, cudaStream_t st);

int grad_reduce(const float* ws, int splits, int64_t n, float* grads, cudaStream_t st);
int adam_from_partials(float* p, const float* ws, int splits, float* m, float* v, int64_t n, float lr, float b1, float b2,
                       float eps, int64_t t, cudaStream_t st);
int adam_f32(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2,
             float eps, int64_t t, cudaStream_t st);
int bernoulli_sample_f32(const float* p, int64_t n, uint64_t seed, uint64_t epoch, float* out, cudaStream_t st);
// graph replay helpers: write the per-step scalars / copy up to 24 (src, dst, n) float ranges in one launch
struct CopyTable { const float* src[24]; float* dst[24]; long long n[24]; int count; };
int set_step_dyn(StepDyn* dev, const StepDyn& v, cudaStream_t st);
int copy_table(const CopyTable& t, cudaStream_t st);
int sir_launch(const float* al, int L, const double* u, int n_u, int size, int64_t* out_idx,
               int32_t* out_consumed, cudaStream_t st);
int softmax_vec_f32(const float* x, int L, float* out, cudaStream_t st);

// ---- train-step megakernel (kernels_f32.cu) ----
struct MegaState;
int mega_begin(iwae_ctx* ctx, cudaStream_t st);            // the launch wrappers record from here on ...
int mega_flush(iwae_ctx* ctx, cudaStream_t st);            // ... until the program is launched as one kernel
void mega_abort(iwae_ctx* ctx);
void mega_dep(cudaStream_t from, cudaStream_t to);         // `to` waits for what has been recorded on `from`
void mega_release(iwae_ctx* ctx);
int zero_f32(float* p, int64_t n, cudaStream_t st);
constexpr int64_t MEGA_MAX_ROWS = 8192;                    // k*B of a recorded step (training shapes)
constexpr int MEGA_MAX_K = 4096;                           // logw_reduce keeps k weights in the shared-memory pool
#define MEGA_SIDE_STREAM ((cudaStream_t)(uintptr_t)0x10)   // key of a level-0 side branch while recording (never launched on)

// ---- fused Dense chain of the train step (kernels_f32.cu) ----
enum { CH_NONE = 0, CH_TANH = 1, CH_EXP_EPS = 2, CH_MULH = 3 };
struct ChainLayer {
  const float* W = nullptr; int ldw = 0;      // W(k, n) = W[k*ldw + n] for n < nsplit
  const float* Wb = nullptr;                  // columns n >= nsplit read Wb[k*ldw + (n - nsplit)] (second head), else unused
  const float* bias = nullptr; const float* biasb = nullptr;
  int K = 0, N = 0, nsplit = 0;               // nsplit == N: one matrix
  int act = CH_NONE, actb = CH_NONE;
  const float* H = nullptr; int64_t ldh = 0;  // CH_MULH: v *= 1 - H[r*ldh + n]^2
  float* out = nullptr; int64_t ldo = 0;      // global output for n < nsplit (may be null)
  float* outb = nullptr; int64_t ldob = 0;    // ... for n >= nsplit
  int vec = 1;                                // cp.async width in floats (set by dense_chain)
};
struct ChainArgs {
  const float* X = nullptr; int64_t ldx = 0; int kx = 0;        // input rows: [X (kx columns) | X2 (kx2 columns)]
  const float* X2 = nullptr; int64_t ldx2 = 0; int kx2 = 0;
  int M = 0, nlayers = 0;
  int TM = 0, kpad = 0, wsz = 0, njobs = 0;                      // set by dense_chain
  ChainLayer L[3];
};
int dense_chain(const ChainArgs& a, cudaStream_t st);
bool dense_chain_fits(const ChainArgs& a);       // every weight chunk of the chain resident in shared memory?
struct TransposeJob { const float* src; float* dst; int rows, cols; int64_t ld_src, ld_dst; };
struct TransposeJobs { TransposeJob job[16]; int count = 0; };
int transpose_jobs(const TransposeJobs& t, cudaStream_t st);

// ---- fused tcgen05 evaluation (eval_tc.cu) ----
int tc_pack_weights(iwae_ctx* ctx, const float* params, bool fast, bool side_ok, cudaStream_t st);
// TERMS mode of the fused kernel: per-row log-density terms [k,N], latent samples [k,N,d] and the per-image q(z1|x)
// parameters [N,d1] (q_mu / q_std must be given; every other pointer may be null)
struct TcTerms { float *lpxz, *prior, *lq1, *lq2, *lp12, *z1, *z2, *q_mu, *q_std; };
int tc_llh_eval(iwae_ctx* ctx, const float* params, const float* x, const float* y, int64_t N, int k,
                uint64_t seed, int64_t img_offset, int fast, float* out_per_example, float* out_vae, cudaStream_t st,
                const float* eps1 = nullptr, const float* eps2 = nullptr, const TcTerms* terms = nullptr);
// encoder_tc.cu: the BasicBlock q(z1|x) on tcgen05 (fp16 hi+lo operands, fp32-accurate); 0 = done, -1 = shape not supported /
// switched off (run block_fwd_f32 instead), > 0 = error
int tc_encoder_fwd(iwae_ctx* ctx, const float* params, const BlockP& blk, const float* X, int64_t ldx, int64_t M, float* mu, float* sd,
                   cudaStream_t st);
void tc_encoder_release(iwae_ctx* ctx);
int block_fwd_f32(const float* P, const BlockP& blk, const float* X, int64_t ldx, int64_t M, float* h1, float* h2,
                  float* mu, float* sd, cudaStream_t st, float* l1_scratch = nullptr, iwae_ctx* lane_ctx = nullptr,
                  cudaStream_t lane = nullptr);
// out[m, n] = act(sum_s partial[s][m, n] + bias[n])   (split-K forward, see dense_fwd in abi.cu)
int splitk_bias_act(const float* partial, int splits, int M, int N, const float* bias, int act, float* out, cudaStream_t st);
// dX[m, n] = (sum_s partial[s][m, n] (+ dX[m, n])) * (1 - H[m, n]^2 if H)   (split-K backward through tanh)
int splitk_mul_h(const float* partial, int splits, int M, int N, const float* H, int64_t ld_h, int accumulate, float* dX,
                 int64_t ld_dx, cudaStream_t st);
void tc_release(iwae_ctx* ctx);
// ---- communicator (comm.cu) ----
int comm_allreduce_f32(iwae_ctx* ctx, float* buf, int64_t n, cudaStream_t st);
float* comm_grad_buffer(iwae_ctx* ctx);
int comm_world(const iwae_ctx* ctx);
void comm_release(iwae_ctx* ctx);
