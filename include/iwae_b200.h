/* iwae_b200.h - C-ABI of libiwae_b200.so: the B200-native importance-weighted ELBO hot path.
 *
 * The reference (nbip/IWAE) has no FFI: its boundary is the Python class API that main.py and
 * the tasks/ scripts call.  This header is what the Python shim (iwae_b200/model.py, reached
 * from src/iwae1.py, src/iwae2.py) binds through ctypes; every entry point names the reference
 * interface it replaces.
 *
 * Conventions
 *   - extern "C"; every function returns int: 0 = ok, non-zero = error; iwae_last_error() gives a
 *     thread-local message.  No exception crosses the ABI.
 *   - All data pointers are DEVICE pointers owned by the caller unless the name ends in _host.
 *     The library owns only the opaque iwae_ctx (scratch arena, packed fp16 weights, staging buffers).
 *   - Every launch takes an explicit CUDA stream (cudaStream_t passed as void*).  Device entry points
 *     never synchronise; *_host entry points synchronise the stream before returning.
 *   - Tensors use the reference layout: sample-major [k, B, .] row-major fp32, Dense kernels
 *     [in, out]; parameters live in ONE flat fp32 buffer in Keras `trainable_weights` order
 *     (SURVEY.md A.9), so gradient / Adam / all-reduce act on a single span.
 *   - One context per device and per host thread (the reference is single-threaded).
 */
#ifndef IWAE_B200_H
#define IWAE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct iwae_ctx iwae_ctx;

/* Model family.  IWAE1: src/iwae1.py:88-96.  IWAE2: src/iwae2.py:99-107.
 * CIWAE: tasks/task04.py:101-108 (label-conditioned encoder/decoder + learned prior p(z|y)).
 * CIWAE_STD: tasks/task05.py (same, N(0,1) prior). */
enum { IWAE_KIND_IWAE1 = 1, IWAE_KIND_IWAE2 = 2, IWAE_KIND_CIWAE = 3, IWAE_KIND_CIWAE_STD = 4 };

/* Objective = the key train_step negates (src/iwae1.py:154-157); DREG = tasks/task02.py:87-101. */
enum {
  IWAE_OBJ_VAE_ELBO = 0, IWAE_OBJ_IWAE_ELBO = 1, IWAE_OBJ_IWAE_EQ14 = 2, IWAE_OBJ_VAE_ELBO_KL = 3,
  IWAE_OBJ_DREG = 4
};

/* Arithmetic of the k-sample evaluation (iwae_llh_eval).
 * FP32        : FFMA kernels, fp32 everywhere (parity mode, <=1e-4 rel on every per-row term).
 * TC_F16      : fused tcgen05 kernel, fp16 operands / fp32 TMEM accumulation, libm-accurate
 *               tanh/exp/log in the epilogues.
 * TC_F16_FAST : same kernel with MUFU approximations (tanh.approx, ex2.approx, lg2.approx); the
 *               throughput mode the bench reports. */
enum { IWAE_PREC_FP32 = 0, IWAE_PREC_TC_F16 = 1, IWAE_PREC_TC_F16_FAST = 2 };

/* The BasicBlocks of a model (src/iwae1.py:24-44), for iwae_block_forward. */
enum {
  IWAE_BLOCK_ENCODER = 0,   /* q(z1|x[,y])  encoder.encode_x_to_z   in: n_x + n_cond */
  IWAE_BLOCK_ENCODER2 = 1,  /* q(z2|z1)     iwae2.py:53             in: d1 */
  IWAE_BLOCK_DECODER21 = 2, /* p(z1|z2)     iwae2.py:78             in: d2 */
  IWAE_BLOCK_PRIOR = 3      /* p(z|y)       task04.py:108           in: n_cond (one-hot rows) */
};

typedef struct {
  int32_t kind;      /* IWAE_KIND_* */
  int32_t n_x;       /* 784 */
  int32_t n_cond;    /* 0, or 10 for CIWAE (one-hot label width, tasks/task04.py:112) */
  int32_t h1, h2;    /* n_hidden (h2 only for IWAE2: main.py:90) */
  int32_t d1, d2;    /* n_latent (d2 only for IWAE2: main.py:89) */
} iwae_config;

/* Optional outputs of iwae_forward / iwae_grads; NULL = not wanted.  Names follow the dict keys of
 * IWAE.call (src/iwae1.py:141-151, src/iwae2.py:154-167, tasks/task04.py:163-173).
 * R = k*B rows, sample-major. */
typedef struct {
  float* lpxz;      /* [k,B]  lpxz / lpxz1 / lpxzy                                   */
  float* lpz;       /* [k,B]  lpz / lpz2 / lpzy (prior term of the top latent)      */
  float* lq1;       /* [k,B]  lqzx / lqz1x / lqzxy                                   */
  float* lq2;       /* [k,B]  lqz2z1           (IWAE2 only)                          */
  float* lp12;      /* [k,B]  lpz1z2           (IWAE2 only)                          */
  float* log_w;     /* [k,B]                                                         */
  float* al;        /* [k,B]  softmax_k(log_w)                                       */
  float* iwae;      /* [B]    logmeanexp_k(log_w)   (per-example; mean_b = iwae_elbo)*/
  float* vae;       /* [B]    mean_k(log_w)         (mean_b = vae_elbo)              */
  float* eq14;      /* [B]    sum_k al*log_w        (mean_b = iwae_eq14)             */
  float* kl;        /* [B]    KL(q(z|x) || N(0,1))  (IWAE1/CIWAE: vae_elbo_kl)       */
  float* dreg;      /* [B]    sum_k al^2 (lpz+lpxz-lq_stopped) (tasks/task02.py:73)  */
  float* snis_z1;   /* [B,d1] sum_k al*z                                             */
  float* snis_z2;   /* [B,d2] (IWAE2 only)                                           */
  float* z1;        /* [k,B,d1]                                                      */
  float* z2;        /* [k,B,d2] (IWAE2 only)                                         */
  float* logits;    /* [k,B,n_x]  (materialised only when asked for)                 */
  float* q_mu;      /* [B,d1]  qzx.loc                                               */
  float* q_std;     /* [B,d1]  qzx.scale (= exp(a)+1e-6)                             */
  float* scalars;   /* [8] batch-mean scalars of the result dict (src/iwae1.py:119-134):
                       [0] vae_elbo  [1] vae_elbo_kl (IWAE1/CIWAE)  [2] iwae_elbo  [3] iwae_eq14
                       [4] inference_loss = -mean_b dreg (tasks/task02.py:76, when dreg is computed)
                       [5] mean lpxz  [6] mean lpz  [7] mean lq1 (the TensorBoard scalars)      */
} iwae_outputs;

const char* iwae_last_error(void);
int iwae_version(void);

/* Replaces the constructors iwae1.IWAE(n_hidden, n_latent) / iwae2.IWAE(...) / CIWAE(...)
 * (src/iwae1.py:89-96, src/iwae2.py:100-107, tasks/task04.py:102-108) on the current device. */
int iwae_create(const iwae_config* cfg, iwae_ctx** out);
int iwae_destroy(iwae_ctx* ctx);

/* model.trainable_weights (Keras order): total fp32 count, tensor count, and per
 * tensor (offset, rows, cols) with cols==0 for a bias vector of length rows. */
int iwae_param_count(const iwae_ctx* ctx, int64_t* n_params, int32_t* n_tensors);
int iwae_param_layout(const iwae_ctx* ctx, int64_t* offsets, int32_t* rows, int32_t* cols);

/* Replaces IWAE.call / val_step (src/iwae1.py:98-151,164-166; src/iwae2.py:109-167; CIWAE.call
 * tasks/task04.py:110-173; IWAEDReG.call tasks/task02.py:34-85).
 *   params [n_params]; x [B,n_x] fp32 {0,1}; y [B] fp32 labels or NULL; eps1 [k,B,d1] / eps2
 *   [k,B,d2] injected N(0,1) noise, or NULL to draw Philox4x32-10 noise keyed on
 *   (seed, global image index img_offset+b, sample, dim) - identical under any sharding. */
int iwae_forward(iwae_ctx* ctx, const float* params, const float* x, const float* y,
                 int32_t B, int32_t k, float beta, const float* eps1, const float* eps2,
                 uint64_t seed, int64_t img_offset, const iwae_outputs* outs, void* stream);

/* The same call with the fused tensor-core kernel as the engine (model.val_step on the whole test set at k = n_samples,
 * main.py:152; generate_and_save_posteriors' call(Xtest, 10), src/iwae1.py:207): every output of iwae_forward except
 * `logits` and `dreg` - per-row terms, log_w, al, per-example ELBOs, SNIS means, z1 / z2, q_mu / q_std, scalars - with
 * no [k,B,784] tensor ever materialised.  Philox noise only.  precision: IWAE_PREC_TC_F16 or IWAE_PREC_TC_F16_FAST.
 * The tensor-core precisions (here and in iwae_llh_eval / iwae_monitor_eval / iwae_llh_eval_eps) REQUIRE binarised images,
 * x in {0,1} exactly, as the reference feeds them (utils.bernoullisample, main.py:117-120): the Bernoulli term is evaluated
 * as -softplus((1-2x) l) and the encoder's first GEMM takes x as an exact fp16 operand.  Real-valued x: IWAE_PREC_FP32. */
int iwae_forward_tc(iwae_ctx* ctx, const float* params, const float* x, const float* y, int32_t B, int32_t k,
                    float beta, uint64_t seed, int64_t img_offset, int32_t precision, const iwae_outputs* outs,
                    void* stream);

/* Replaces the k=5000 test-set loop of main.py:170-184 (+ utils.MyMetric, src/utils.py:30-45):
 * per-example logmeanexp_k(log_w) for N images, and their fp64 sum (mean = sum / N is the LLH).
 *   x [N,n_x]; y [N] or NULL; out_per_example [N] or NULL; out_sum: one double on the device.
 *   precision = IWAE_PREC_*.  Philox noise only (use iwae_forward for injected eps). */
int iwae_llh_eval(iwae_ctx* ctx, const float* params, const float* x, const float* y,
                  int64_t N, int32_t k, uint64_t seed, int64_t img_offset, int32_t precision,
                  float* out_per_example, double* out_sum, void* stream);

/* Same estimator through the fused tensor-core kernel with INJECTED noise instead of Philox (north_star: "an injected-eps
 * mode takes an externally supplied eps tensor"): eps1 [k,N,d1], eps2 [k,N,d2] (IWAE2 only, else NULL), sample-major like
 * qzx.sample(k) (src/iwae1.py:59).  The parity tests feed the oracle the very same tensors, so the comparison does not
 * rest on the restatement of the device RNG.  precision: IWAE_PREC_TC_F16 or IWAE_PREC_TC_F16_FAST. */
int iwae_llh_eval_eps(iwae_ctx* ctx, const float* params, const float* x, const float* y, int64_t N, int32_t k,
                      const float* eps1, const float* eps2, int32_t precision, float* out_per_example, double* out_sum,
                      void* stream);

/* Monitoring pass (model.val_step on the whole test set, main.py:152) through the same
 * fused kernel: per-example logmeanexp_k(log_w) AND mean_k(log_w) for N images at beta = 1, i.e. the
 * iwae_elbo / vae_elbo the training loop prints, without materialising [k,B,784].
 *   out_iwae, out_vae [N]; out_sums: two doubles on the device (sum of each) or NULL. */
int iwae_monitor_eval(iwae_ctx* ctx, const float* params, const float* x, const float* y, int64_t N,
                      int32_t k, uint64_t seed, int64_t img_offset, int32_t precision, float* out_iwae,
                      float* out_vae, double* out_sums, void* stream);

/* Declares the CONTENT version of the parameter buffer the caller passes to iwae_llh_eval / iwae_monitor_eval (the
 * Python shim bumps it in set_weights / load_weights / train_step / Variable.assign, i.e. whenever
 * model.trainable_weights change, src/iwae1.py:160).  While (pointer, version) stay the same the fused evaluation
 * reuses its packed fp16 weight image instead of re-packing it on every call.  version 0 (the default) = unknown:
 * always re-pack. */
int iwae_set_params_version(iwae_ctx* ctx, uint64_t version);

/* Device time (ms) of the fused tensor-core evaluation kernel in the most recent iwae_llh_eval /
 * iwae_llh_eval_host call with a TC precision on this context (CUDA events recorded around that one
 * launch on the caller's stream); synchronises on the second event.  bench.py's roofline uses it. */
int iwae_llh_kernel_ms(iwae_ctx* ctx, float* ms);

/* Same, reference-facing with HOST buffers: copies x (and y) host->device, runs, copies the
 * per-example vector and the sum back and synchronises.  This is the call the e2e number times. */
int iwae_llh_eval_host(iwae_ctx* ctx, const float* params, const float* x_host, const float* y_host,
                       int64_t N, int32_t k, uint64_t seed, int64_t img_offset, int32_t precision,
                       float* out_per_example_host, double* out_sum_host, void* stream);

/* Gradient of -res[objective] w.r.t. the flat parameter buffer (tape.gradient in
 * src/iwae1.py:155-159; DREG: tasks/task02.py:90-96).  grads [n_params] is overwritten.
 * grad_scale multiplies every gradient (1/world for data-parallel image shards use B_global via
 * B_norm: the batch size the objective's mean_b divides by; pass B when not sharded). */
int iwae_grads(iwae_ctx* ctx, const float* params, const float* x, const float* y,
               int32_t B, int32_t k, float beta, int32_t objective, int32_t B_norm,
               const float* eps1, const float* eps2, uint64_t seed, int64_t img_offset,
               float* grads, const iwae_outputs* outs, void* stream);

/* keras.optimizers.Adam(lr, epsilon=1e-4).apply_gradients (main.py:93, src/iwae1.py:160):
 * t is the step count AFTER increment; epsilon is added to the uncorrected sqrt(v) (A.6). */
int iwae_adam_step(float* params, const float* grads, float* m, float* v, int64_t n,
                   float lr, float beta1, float beta2, float eps, int64_t t, void* stream);

/* Replaces model.train_step (src/iwae1.py:153-162; tasks/task02.py:87-101; tasks/task04.py:175-184):
 * forward + backward + Adam on one device in one call.  For data-parallel training call iwae_grads,
 * all-reduce, then iwae_adam_step.
 * With Philox noise (eps1 == eps2 == NULL) the step is captured into a CUDA graph on the second consecutive call
 * with one signature (params / m / v pointers, B, k, beta, objective, set of requested outputs) and replayed from
 * then on - same kernels, bit-identical results; x, y and the outputs may live anywhere (they are staged through
 * library-owned buffers on the device).  IWAE_NO_GRAPH=1 in the environment keeps every call eager. */
int iwae_train_step(iwae_ctx* ctx, float* params, float* adam_m, float* adam_v,
                    const float* x, const float* y, int32_t B, int32_t k, float beta,
                    int32_t objective, const float* eps1, const float* eps2, uint64_t seed,
                    int64_t img_offset, float lr, float beta1, float beta2, float adam_eps,
                    int64_t t, const iwae_outputs* outs, void* stream);

/* ---- multi-GPU (the reference is single-device: main.py:32) -------------------------------------
 * One process per GPU.  The library owns an NCCL communicator per context (NCCL is bound at run time with
 * dlopen("libnccl.so.2"), i.e. the copy PyTorch has already loaded when there is one).
 *   iwae_comm_unique_id : rank 0 creates the 128-byte ncclUniqueId, the host side broadcasts it (torch.distributed).
 *   iwae_comm_init      : every rank joins (collective).   iwae_comm_destroy: leave.
 *   iwae_allreduce_sum_*: in-place sum all-reduce of a device buffer on `stream` (the fp64 LLH sum of the sharded
 *                         k-sample evaluation, main.py:181). */
int iwae_comm_unique_id(uint8_t* out128);
int iwae_comm_init(iwae_ctx* ctx, const uint8_t* id128, int32_t rank, int32_t world);
int iwae_comm_destroy(iwae_ctx* ctx);
int iwae_comm_info(const iwae_ctx* ctx, int32_t* rank, int32_t* world, int32_t* nccl_version);
int iwae_allreduce_sum_f32(iwae_ctx* ctx, float* buf, int64_t n, void* stream);
int iwae_allreduce_sum_f64(iwae_ctx* ctx, double* buf, int64_t n, void* stream);

/* Data-parallel model.train_step (src/iwae1.py:153-162 / tasks/task02.py:87-101 over image shards): this rank's
 * B_local images (global indices img_offset .. img_offset + B_local of a batch of B_global; B_local may be 0), local
 * gradient divided by B_global, ONE ncclAllReduce(sum) of the flat gradient, the same Keras-Adam update on every
 * rank.  Like iwae_train_step the whole step - including the all-reduce - is captured into a CUDA graph on the second
 * consecutive call with one signature and replayed afterwards.  `outs` describe the LOCAL shard. */
int iwae_dp_train_step(iwae_ctx* ctx, float* params, float* adam_m, float* adam_v, const float* x_local,
                       const float* y_local, int32_t B_local, int32_t B_global, int32_t k, float beta,
                       int32_t objective, uint64_t seed, int64_t img_offset, float lr, float beta1, float beta2,
                       float adam_eps, int64_t t, const iwae_outputs* outs, void* stream);

/* The device noise source by itself: eps [k,B,d] for stream_id 1 (first latent) or 2 (second). */
int iwae_philox_normal(uint64_t seed, int32_t stream_id, int64_t img_offset, int32_t B, int32_t k,
                       int32_t d, float* out, void* stream);

/* Decoder-only tools (model.decoder(z) + pxz.log_prob(x), tasks/plot_task01.py:57-63;
 * model.sample, src/iwae1.py:168-178): z [R, d1(+n_cond folded by y)] -> logits [R,n_x] and/or
 * lpxz [R] = sum_j Bernoulli(logits).log_prob(x[r % B]).  Either output may be NULL. */
int iwae_decode(iwae_ctx* ctx, const float* params, const float* z, const float* x, const float* y,
                int32_t R, int32_t B, float* logits, float* lpxz, void* stream);

/* One BasicBlock on caller-provided rows: in [R, n_in(block)] -> mu, sd [R, n_out] with sd = exp(.) + 1e-6.
 * Replaces decoder.decode_z2_to_z1(z2) in IWAE.sample / conditional_sample (src/iwae2.py:184-213) and
 * conditional_prior_network(y_onehot) in CIWAE.sample (tasks/task04.py:190-196). */
int iwae_block_forward(iwae_ctx* ctx, const float* params, int32_t block, const float* in, int32_t R,
                       float* mu, float* sd, void* stream);

/* Distribution shims: what the reference's scripts ask of the tfd objects model.encoder / model.decoder return.
 *   iwae_bernoulli_log_prob : pxz.log_prob(x) elementwise (tasks/plot_task01.py:63): logits [R,n], x [B,n] (row r % B)
 *                             -> out [R,n] = -(1-x) softplus(l) - x softplus(-l)        
 *   iwae_normal_log_prob    : qzx.log_prob(z) elementwise (tasks/plot_task01.py:72): z [R,d], loc/scale [Rm,d]
 *                             (row r % Rm; both NULL = N(0,1)) -> out [R,d], TFP op order (A.2)
 *   iwae_normal_sample      : qzx.sample(n) (src/iwae1.py:59): out [n,Rm,d] = eps*scale + loc; eps_in [n,Rm,d], or NULL to
 *                             draw it from the device Philox generator (seed, stream_id)
 *   iwae_bernoulli_sample_logits : Bernoulli(logits).sample() and .probs (src/iwae1.py:172-176): probs = sigmoid(logits),
 *                             x = (u < probs) with the uniforms of iwae_bernoulli_sample; probs_out may be NULL */
int iwae_bernoulli_log_prob(const float* logits, const float* x, int64_t R, int32_t B, int32_t n, float* out, void* stream);
int iwae_normal_log_prob(const float* z, const float* loc, const float* scale, int64_t R, int32_t Rm, int32_t d, float* out,
                         void* stream);
int iwae_normal_sample(const float* loc, const float* scale, const float* eps_in, int32_t Rm, int32_t n, int32_t d,
                       uint64_t seed, int32_t stream_id, float* out, void* stream);
int iwae_bernoulli_sample_logits(const float* logits, int64_t n, uint64_t seed, uint64_t epoch, float* x_out,
                                 float* probs_out, void* stream);

/* The true-posterior grid of tasks/plot_task01.py:46-73 for a 2-D latent, entirely on the device: grid =
 * meshgrid(linspace(x0,x1,n1), linspace(y0,y1,n2)) [n1*n2, 2]; out_log_post [n2*n1] = lpxz + lpz - logsumexp over the
 * grid; out_log_q [n2*n1] = sum_j log q(z_j | x).  x: one image [n_x]. */
int iwae_posterior_grid(iwae_ctx* ctx, const float* params, const float* x, double x0, double x1, double y0, double y1,
                        int32_t n1, int32_t n2, float* out_log_post, float* out_log_q, float* out_grid, void* stream);

/* SIR: numpy.random.RandomState.choice(arange(L), size, replace=False, p=al)
 * (tasks/plot_task01.py:88) bit-exactly, given the float64 uniforms the RandomState would draw.
 *   al [L] fp32 normalised weights; uniforms [n_uniforms] fp64 (device); out_idx [size] int64;
 *   out_consumed: number of uniforms used (device int32).  Single-CTA kernel, fp64 sequential cdf. */
int iwae_sir(const float* al, int32_t L, const double* uniforms, int32_t n_uniforms, int32_t size,
             int64_t* out_idx, int32_t* out_consumed, void* stream);

/* scipy.special.softmax on a float32 vector (tasks/plot_task01.py:85): out[i] = exp(x[i]-lse(x)). */
int iwae_softmax_f32(const float* log_w, int32_t L, float* out, void* stream);

/* utils.bernoullisample on the device (src/utils.py:26-27): x = (u < p) as fp32 {0,1} with
 * Philox uniforms keyed on (seed, epoch, element).  Statistically equivalent, not stream-identical,
 * to NumPy's legacy binomial. */
int iwae_bernoulli_sample(const float* p, int64_t n, uint64_t seed, uint64_t epoch, float* out,
                          void* stream);

/* Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t iwae_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* IWAE_B200_H */
