// Fused tcgen05 evaluation kernel: the k-sample importance-weighted log-likelihood of main.py:170-184
// in ONE pass per 128-row tile, with every activation and all 784 logits staying on chip.
//
// One CTA owns one tile (<=128 samples of ONE image) at a time and walks a static layer "program":
//   z1 = mu1(x) + std1(x)*eps   (Philox in registers)              -> fp16 A operand in smem
//   [IWAE2]  q(z2|z1): tanh,tanh,(mu2|log-std2) -> z2 ; p(z1|z2): tanh,tanh,(mu|log-std) -> log p(z1|z2)
//   p(x|z1): tanh, tanh, 784 logits -> Bernoulli log-lik reduced in the epilogue (logits never leave TMEM)
// The kernel runs as CTA PAIRS (clusters of two CTAs on the two SMs of a TPC): every tcgen05.mma is a cta_group::2
// instruction with M = 256 - rows 0..127 are the tile of the even (leader) CTA, rows 128..255 the tile of its peer -
// and each CTA stores only its half (N/2 columns) of every weight slab.
// Roles (320 threads): warp 0 = weight producer (cp.async.bulk half-slabs through an mbarrier ring; weights are fp16,
// pre-packed in the exact smem image, ~0.31 MB per tile and CTA from L2), warp 1 = MMA issuer in the leader (tcgen05.mma,
// kind::f16, fp32 accumulators in TMEM, two 128-column accumulator buffers; completion multicast to both CTAs' mbarriers)
// and slab-arrival relay in the peer (the issuer runs with all 32 lanes converged and warp-uniform control flow, so its
// descriptors live in uniform registers; producer and relay run inside one elected thread each), warps 2-9 = epilogue
// (tcgen05.ld -> the bias is already folded into the GEMM through a constant-1 column -> tanh / exp / softplus -> fp16
// back to smem for the next layer; warps e and e+4 share a TMEM lane quarter and split the accumulator columns;
// log-density terms accumulate per row in registers; they arrive on the LEADER's mbarriers).  Two CTAs (of two
// different pairs) are resident per SM (<=113 KB smem, 256 TMEM columns each) so one CTA's epilogue overlaps the other
// CTA's MMAs.  Every pair owns a contiguous range of tile pairs (consecutive tiles share the image, whose
// constants are loaded once); the second decoder layer parks its first output chunk in the by then dead z1 buffer so
// that its epilogue overlaps the MMAs of its second chunk.  -DIWAE_PAIR=0 builds the round-1 schedule (independent
// CTAs, cta_group::1).
// Operand layout: no-swizzle K-major core matrices, element (row,k) at (k/8)*rows*16 + row*16 + (k%8)*2 bytes,
// i.e. LBO = rows*16 (K direction), SBO = 128 (M/N direction).  DESIGN.md section 4.1 has the roofline and the
// measured timeline.
#include "common.cuh"
#include "rng.cuh"
#include "tc_ptx.cuh"
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <algorithm>
#include <stdlib.h>
#include <stddef.h>
#include <vector>

namespace tc {

// ------------------------------------------------------------------------------------------------
// Fused evaluation kernel
// ------------------------------------------------------------------------------------------------
#ifndef IWAE_PRODUCER_SLEEP
#define IWAE_PRODUCER_SLEEP 32
#endif
// IWAE_PAIR = 1 (default): the kernel runs as CTA PAIRS (clusters of two CTAs on the two SMs of a TPC) and every
// tcgen05.mma is a cta_group::2 instruction with M = 256: rows 0..127 are the tile of the even ("leader") CTA, rows
// 128..255 the tile of its peer; each CTA keeps only HALF of every weight slab (N/2 columns of B) in its shared
// memory, so the L2 -> shared-memory weight stream, the shared-memory operand reads of B and the number of MMA
// instructions per row are all halved.  Only the leader issues MMAs; completion is multicast to both CTAs' mbarriers.
// IWAE_PAIR = 0 builds the round-1 schedule (independent CTAs, cta_group::1) for A/B measurements.
#ifndef IWAE_PAIR
#define IWAE_PAIR 1
#endif
#if IWAE_PAIR
#define IWAE_CG "cta_group::2"
#else
#define IWAE_CG "cta_group::1"
#endif
constexpr int PAIR = IWAE_PAIR ? 2 : 1;               // CTAs that share one MMA (and split every weight slab)
constexpr int TILE_M = 128;
constexpr int MAX_LAYERS = 10;
constexpr int MAX_CHUNKS = 7;
constexpr int CHUNK_N_MAX = 112;
constexpr int CTAS_PER_SM = 2;                        // two resident CTAs per SM: one CTA's epilogue overlaps the other's MMAs
constexpr int NSTAGES = 4;                            // weight-slab ring depth
constexpr int SLAB_K = 32 * PAIR;                     // K rows per slab: the ring holds NSTAGES * SLAB_K rows of B ahead of the MMAs
constexpr int SLAB_BYTES = (CHUNK_N_MAX / PAIR) * SLAB_K * 2;   // 7168: this CTA's share of a slab
constexpr int KZ_MAX = 112;                            // K capacity of the z1 buffer
constexpr int KP_MAX = 208;                            // K capacity of the activation buffer
constexpr int BUFZ_BYTES = TILE_M * KZ_MAX * 2;        // 28672
constexpr int BUFP_BYTES = TILE_M * KP_MAX * 2;        // 53248
constexpr int NX_MAX = 784;
constexpr int D_MAX = 100;
constexpr int NUM_EPI_WARPS = 8;                      // epilogue warps per CTA
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;  // producer warp, MMA warp, epilogue warps
constexpr int TMEM_COLS = 256;                        // 2 accumulator buffers x 128 columns

// ---- cluster / cta_group-dependent PTX wrappers of the fused kernel (all tcgen05 instructions of one kernel must use
// ---- the same cta_group; the self-test and micro-benchmark kernels above stay cta_group::1) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` (a shared::cta address) inside the CTA of rank `rank`
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// Arrive on an mbarrier of any CTA of the cluster.  Default semantics (.release.cta), like CUTLASS's
// ClusterBarrier::arrive(cta_id): an explicit .release.cluster compiles to MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in front
// of every arrive (~1,000 cycles each; the first pair-mode build was 20 % SLOWER than independent CTAs because of it).
// What the arrive publishes is either "my tcgen05.ld of this accumulator have completed" (ordered by
// tcgen05.fence::before_thread_sync) or shared-memory stores that a fence.proxy.async has already made visible to the
// tensor core of the SM that owns that shared memory - neither needs a cluster-scope memory fence.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a barrier that CTAs of the whole pair arrive on
__device__ __forceinline__ void mbar_wait_pair(uint64_t* bar, uint32_t parity, int tag, uint32_t sleep_ns = 32) {
  mbar_wait(bar, parity, tag, sleep_ns);
}
// IWAE_MMA_CONVERGED = 1: the MMA role runs with all 32 lanes converged (only the tcgen05 instructions are predicated on the
// elected lane) and its barrier waits keep the control flow warp-UNIFORM (the loop condition is a vote), so that the
// compiler can keep the descriptor arithmetic in uniform registers - the operand form UTCHMMA reads - instead of moving
// seven vector registers to uniform ones (R2UR) in front of every MMA.
#ifndef IWAE_MMA_SLEEP
#define IWAE_MMA_SLEEP 32
#endif
#ifndef IWAE_MMA_CONVERGED
#define IWAE_MMA_CONVERGED 1
#endif
__device__ __forceinline__ void mma_wait(uint64_t* bar, uint32_t parity, int tag) {
#if IWAE_MMA_CONVERGED
  uint32_t polls = 0;
  while (!__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) {
#if IWAE_MMA_SLEEP
    __nanosleep(IWAE_MMA_SLEEP);
#endif
    if (++polls > (1u << 26)) { if ((threadIdx.x & 31) == 0) printf("iwae_b200: mbarrier timeout tag=%d block=%d\n", tag, blockIdx.x); __trap(); }
  }
#else
  mbar_wait(bar, parity, tag, 32);
#endif
}
__device__ __forceinline__ void tmem_alloc_cg(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc." IWAE_CG ".sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_cg() {
  asm volatile("tcgen05.relinquish_alloc_permit." IWAE_CG ".sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc." IWAE_CG ".sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 operands, fp32 accumulate; M = 128 per CTA of the group
__device__ __forceinline__ void umma_f16_cg(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma." IWAE_CG ".kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
               "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
               : "memory");
}
// mbarrier arrive when all MMAs issued so far have retired; in pair mode the arrive is multicast to the barrier at
// the same shared-memory offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_cg(uint64_t* bar) {
#if IWAE_PAIR
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
#else
  umma_commit(bar);
#endif
}
// Predicated forms for the warp-convergent MMA role: all 32 lanes run the (warp-uniform) control flow and descriptor
// arithmetic - which the compiler then keeps in UNIFORM registers, the operand form tcgen05.mma wants - and only the
// elected lane (issue != 0) executes the tcgen05 instruction.
__device__ __forceinline__ void umma_f16_cg_pred(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                                 uint32_t issue) {
#ifdef IWAE_DIAG_NOMMA
  return;                                      // timing diagnostic only (wrong results): the schedule without tensor work
#endif
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t@q tcgen05.mma." IWAE_CG
      ".kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(issue)
      : "memory");
}
__device__ __forceinline__ void umma_commit_cg_pred(uint64_t* bar, uint32_t issue) {
#if IWAE_PAIR
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;\n\t}" ::"r"(smem_u32(bar)),
      "h"((uint16_t)3), "r"(issue)
      : "memory");
#else
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(
          smem_u32(bar)),
      "r"(issue)
      : "memory");
#endif
}
// instruction descriptor of the group MMA: M = 128 * PAIR
__device__ __forceinline__ uint32_t make_idesc_cg(uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | (((128u * PAIR) >> 4) << 24); }

enum { EPI_TANH = 0, EPI_GSAMPLE = 1, EPI_GLOGP = 2, EPI_BERN = 3 };

struct Layer {
  int kind, src, dst;          // src/dst: 0 = z1 buffer, 1 = activation buffer
  int k_pad;                   // multiple of 16 (includes the constant-1 bias column)
  int n_real;                  // TANH: units, GSAMPLE/GLOGP: latent dims, BERN: pixels
  int next_kpad;               // K of the consumer of this layer's output (0 = none)
  int nchunks;
  int chunk_n[MAX_CHUNKS];     // MMA N of the chunk (multiple of 16, <= 112)
  int chunk_cnt[MAX_CHUNKS];   // real units / dims / pixels in the chunk
  int chunk_off[MAX_CHUNKS];   // first unit / dim / pixel of the chunk
  int wait_act, signal_act, inplace;
  int split_dst;               // 1: chunk 0 of this (two-chunk, otherwise in-place) layer fills the z1 buffer completely (KZ_MAX columns)
                               //    and the later chunks go to the START of the activation buffer, which follows the z1 buffer in
                               //    shared memory: the output is then ONE contiguous K-major operand that begins at the z1 buffer
  int src_split;               // 1: this layer reads that contiguous operand (src == 0, k_pad > KZ_MAX)
  // Early hand-off between two dependent layers: the epilogue warps of a TANH layer signal `act_early` as soon as the first
  // SLAB_K output columns of every row are in shared memory (early_sig), and the MMA issuer of the next layer starts
  // with the K steps of its first weight slab then, waiting for `act_ready` (all columns) only before the second slab
  // (wait_split).  Hides ~half of the issue + tensor latency of the dependent chain of small layers.
  int early_sig, wait_split;
  int side;                    // host-side only: an independent layer placed between a producer and its consumer (see build_program,
                               // "side layer"); it signals nothing and is skipped when producers are linked to consumers
  int early_two;               // wait_split consumer of a TWO-chunk producer: its first chunk accumulates into the TMEM buffer of the
                               // producer's chunk 0, so the issuer waits for that buffer's acc_empty explicitly (after a one-chunk
                               // producer the consumer's first chunk gets the other buffer, drained a layer earlier)
};
struct Program {
  int nlayers, two_layer, cond, learned_prior;
  int d1, d2, nc, nx, kz_pad;  // kz_pad: K of the layers that read z1
  int z1_lo;                   // 1: room for the fp16 "lo" halves of z1 (hi + lo = 22 bits) next to the operands, see stage 0
  int stage0_early;            // 1: stage 0 signals act_early once the first SLAB_K columns of z1 are written (Layer::early_sig)
  Layer L[MAX_LAYERS];
};

struct EvalArgs {
  const uint8_t* wpack;        // packed fp16 weight slabs, program order
  const float* x;              // [N, nx]
  const float* y;              // [N] or null
  const float* mu1;            // [N, d1]
  const float* std1;           // [N, d1]
  const float* pmu;            // [N, d1] learned prior (CIWAE) or null
  const float* pstd;
  float2* partial;             // [N * tiles_per_img] (max, sum exp)
  float* partial_lw;           // [N * tiles_per_img] sum of log_w over the tile's valid rows (monitoring), or null
  int64_t n_tiles;
  int tiles_per_img, rows_per_tile, k;
  uint64_t seed; int64_t img_offset;
  // injected-eps mode (parity vehicle: the oracle is fed the SAME tensors, no dependence on the RNG restatement):
  const float* eps1;           // [k, n_images, d1] or null (Philox)
  const float* eps2;           // [k, n_images, d2] or null
  int64_t n_images;
  // TERMS mode (the val_step / IWAE.call engine): per-row log-density terms [k, n_images] sample-major like the reference
  // (src/iwae1.py:141-151, src/iwae2.py:154-167) and the latent samples [k, n_images, d]; any pointer may be null
  float* term_lpxz; float* term_prior; float* term_lq1; float* term_lq2; float* term_lp12;
  float* z1_out; float* z2_out;
  long long* trace;            // debug timeline (IWAE_TC_TRACE=<file>): [cta<4][role<3][512] (tag, clock64) pairs, else null
};

#define TRACE_SLOTS 1024
// tag = tile_local*10000 + layer*100 + chunk*10 + phase
template <bool TRACE>
__device__ __forceinline__ void trace_ev(const EvalArgs& a, int role, int& n, int tag) {
  if (TRACE && a.trace && blockIdx.x < 4 && n < TRACE_SLOTS) {
    long long* p = a.trace + (((size_t)blockIdx.x * 3 + role) * TRACE_SLOTS + n) * 2;
    p[0] = tag; p[1] = clock64();
    ++n;
  }
}


struct __align__(16) SubTile {       // the CTA's 128-row tile
  uint8_t bufZ[BUFZ_BYTES];
  uint8_t bufP[BUFP_BYTES];
  float sgn[NX_MAX];           // 1 - 2x per pixel of the tile's image (0 for padded columns); 16-byte aligned
  float mu[D_MAX], sd[D_MAX], pmu[D_MAX];
  float psd[D_MAX];            // learned prior: 1/std in FAST mode, std in precise mode
  float red[16];               // [0,4) max, [4,8) sum-exp, [8,16) per-warp constant partials
  float red_lw[4];             // per-quarter sums of log_w (monitoring)
  float red_p[8];              // per-warp partial of sum_j log std of a learned prior (TERMS mode keeps it apart from red[8..16))
  // (the partial log-weights of the second column half of each row are parked in bufP, which is dead
  //  once the last layer's MMAs have retired: a 4th weight slab fits in the shared memory this frees)
  int label; int pad_[3];
};
// Both CTAs of a pair use the SAME layout: cta_group::2 descriptors and multicast commits address the two CTAs'
// shared memories by one common offset.
struct __align__(16) SmemLayout {
  SubTile sub;
  uint8_t ring[NSTAGES][SLAB_BYTES];
  // full / empty: weight ring.  acc_full[b]: accumulator buffer b complete (commit, multicast).  acc_empty[b],
  // act_ready: arrived on by the epilogue warps of BOTH CTAs, in the leader's shared memory (the MMA issuer waits there)
  uint64_t full[NSTAGES], empty[NSTAGES], acc_full[2], acc_empty[2], act_ready, act_early;
  uint32_t tmem_base;
};

// two resident CTAs per SM: the dynamic shared memory of one CTA is at most ~113 KB
static_assert(CTAS_PER_SM * (sizeof(SmemLayout) + 1024) <= 228 * 1024, "fused eval kernel: shared memory exceeds the SM");

#define LOG2PI_HALF_F 0.91893853320467274178f

// named barrier of the CTA's 8 epilogue warps
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <bool FAST>
__device__ __forceinline__ float act_tanh(float v) {
#ifdef IWAE_DIAG_NOTANH
  return v * 0.25f;                            // timing diagnostic only (wrong results): the layer without its MUFU.TANH
#endif
  if (FAST) { float r; asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
  return tanhf(v);
}
template <bool FAST>
__device__ __forceinline__ float act_exp(float v) { return FAST ? __expf(v) : expf(v); }
template <bool FAST>
__device__ __forceinline__ float act_log(float v) { return FAST ? __logf(v) : logf(v); }

// store 16 consecutive k values (two 16-byte vectors) of one row into an A-operand buffer
__device__ __forceinline__ void store16(uint8_t* buf, int row, int k0, const float (&v)[16]) {
  uint4 a = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
  uint4 b = make_uint4(pack_h2(v[8], v[9]), pack_h2(v[10], v[11]), pack_h2(v[12], v[13]), pack_h2(v[14], v[15]));
  *reinterpret_cast<uint4*>(buf + (size_t)(k0 >> 3) * (TILE_M * 16) + row * 16) = a;
  *reinterpret_cast<uint4*>(buf + (size_t)((k0 >> 3) + 1) * (TILE_M * 16) + row * 16) = b;
}

// store 8 consecutive k values (one 16-byte core-matrix row) of one row
__device__ __forceinline__ void store8(uint8_t* buf, int row, int k0, const float (&v)[8]) {
  uint4 a = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
  *reinterpret_cast<uint4*>(buf + (size_t)(k0 >> 3) * (TILE_M * 16) + row * 16) = a;
}

// ---- math helpers ----
__device__ __forceinline__ float ex2_approx(float v) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float lg2_approx(float v) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float rcp_approx(float v) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float sqrt_approx(float v) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float sin_approx(float v) { float r; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float cos_approx(float v) { float r; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
// 2^x for x <= 0 on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f, f in [-0.5, 0.5], degree-4 minimax
// polynomial for 2^f (max relative error 2.7e-6 incl. fp32 evaluation), exponent inserted with one integer add.  The
// Bernoulli epilogue evaluates a compile-time share of its exponentials this way: the XU pipe (16 MUFU/clk/SM) is the
// most contended unit of the kernel while the FMA pipe is under 20 % busy (the FlashAttention-4 trick).
__device__ __forceinline__ float ex2_poly_neg(float x) {
  x = fmaxf(x, -125.f);
  const float r = x + 12582912.f;            // 1.5 * 2^23: the low mantissa bits of r hold round(x)
  const float f = x - (r - 12582912.f);
  float p = 0.009570102207362652f;
  p = fmaf(p, f, 0.05591785907745361f);
  p = fmaf(p, f, 0.240247443318367f);
  p = fmaf(p, f, 0.6931217908859253f);
  p = fmaf(p, f, 0.9999992847442627f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}
#ifndef IWAE_EX2_POLY
#define IWAE_EX2_POLY 0                       // of every 4 groups of 4 pixels, how many use ex2_poly_neg instead of MUFU.EX2
#endif
#define LN2_F 0.69314718055994530942f
#define LOG2E_F 1.44269504088896340736f

// Eight N(0,1) for dims 8g..8g+7: Philox4x32-10 + Box-Muller on 16-bit uniforms (rng.cuh defines the mapping).
// FAST builds the uniforms with two ALU ops each (no integer->float conversion, which would share the XU pipe
// with the transcendentals) and uses the MUFU approximations (lg2, sqrt, sin, cos on (-pi,pi)): same uniforms,
// normals equal to ~1e-6.
template <bool FAST>
__device__ __forceinline__ void normal8(uint64_t seed, uint32_t stream_id, uint32_t img, uint32_t s, uint32_t g, float (&o)[8]) {
  if (!FAST) { philox_normal8(seed, stream_id, img, s, g, o); return; }
#ifdef IWAE_DIAG_NONOISE
  {                                            // timing diagnostic only (wrong results): no generator at all
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.125f * (float)((int)((s + g + i) & 15u) - 8);
    return;
  }
#endif
#ifdef IWAE_DIAG_NOPHILOX
  // timing diagnostic only (wrong results): Box-Muller kept, the Philox rounds replaced by one multiply per word
  const uint4 r = make_uint4(s * 2654435761u + g, g * 40503u + img, (s ^ img) * 2246822519u, (s + stream_id) * 3266489917u);
#else
  const uint4 r = philox4x32_10(make_uint4(s, g, img, stream_id), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
#endif
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const float f1 = __uint_as_float(((w[h] >> 9) & 0x007FFF80u) | 0x3F800000u);     // 1 + hi16/65536
    const float f2 = __uint_as_float(((w[h] << 7) & 0x007FFF80u) | 0x3F800000u);     // 1 + lo16/65536
    const float u1 = f1 - 0.99999237060546875f;                                      // (hi16 + 0.5)/65536, exact
    const float rad = sqrt_approx(-2.0f * LN2_F * lg2_approx(u1));
    // theta = 2 pi u2 - pi in (-pi, pi), u2 = (lo16 + 0.5)/65536;  cos(2 pi u2) = -cos(theta), sin likewise
    const float th = fmaf(f2, 6.283185307179586f, -9.4247300721120986f);             // -(3 pi - 2 pi 2^-17)
    o[2 * h] = -rad * cos_approx(th);
    o[2 * h + 1] = -rad * sin_approx(th);
  }
}

// TMEM -> registers, 16 columns at a time, software-pipelined: the load of this warp's next group is
// in flight while the current one is processed (tcgen05.wait::ld waits for ALL outstanding loads, so
// the order is wait(g) -> issue(next) -> process(g)).
template <class F>
__device__ __forceinline__ void for_each_16cols(uint32_t taddr, int g_first, int ncols, F&& f) {
  // this warp's groups: g_first, g_first+32, ... (the partner warp of the lane quarter takes the others)
  if (g_first >= ncols) return;
  uint32_t ra[16], rb[16];
  tmem_ld16(taddr + g_first, ra);
  tmem_ld_wait();
  for (int g = g_first; g < ncols; g += 64) {
    const bool has_b = g + 32 < ncols;
    if (has_b) tmem_ld16(taddr + g + 32, rb);
    f(ra, g);
    if (has_b) {
      tmem_ld_wait();
      const bool has_a = g + 64 < ncols;
      if (has_a) tmem_ld16(taddr + g + 64, ra);
      f(rb, g + 32);
      if (has_a) tmem_ld_wait();
    }
  }
}

// TRACE: the timeline instrumentation (IWAE_TC_TRACE) is compiled into a separate instantiation
// The five log-density terms of a row (sums over columns).  The evaluation needs only their signed sum (one register);
// TERMS mode keeps them apart (five registers) and writes them per row.
enum { T_LPXZ = 0, T_PRIOR = 1, T_LQ1 = 2, T_LQ2 = 3, T_LP12 = 4 };
template <bool TERMS>
struct RowAcc {
  float v[TERMS ? 5 : 1];
  __device__ __forceinline__ RowAcc() {
#pragma unroll
    for (int i = 0; i < (TERMS ? 5 : 1); ++i) v[i] = 0.f;
  }
  // term T of log_w = lpxz + prior + lp12 - lq1 - lq2
  template <int T>
  __device__ __forceinline__ void add(float x) {
    if (TERMS) v[T] += x;
    else v[0] += (T == T_LQ1 || T == T_LQ2) ? -x : x;
  }
  __device__ __forceinline__ float log_w() const {
    if (TERMS) return v[T_LPXZ] + v[T_PRIOR] + v[TERMS ? T_LP12 : 0] - v[TERMS ? T_LQ1 : 0] - v[TERMS ? T_LQ2 : 0];
    return v[0];
  }
};

// INJECT: the N(0,1) noise is read from args.eps1 / args.eps2 instead of being drawn in registers
// TERMS: per-row terms and latent samples are written out (RowAcc above)
template <bool FAST, bool TRACE, bool INJECT, bool TERMS>
__global__ void
#if IWAE_PAIR
__cluster_dims__(2, 1, 1)
#endif
__launch_bounds__(NUM_THREADS, CTAS_PER_SM) iwae_eval_tc_kernel(const __grid_constant__ EvalArgs args, const __grid_constant__ Program P) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemLayout& S = *reinterpret_cast<SmemLayout*>(smem_raw);
  // (the warp index through a shuffle: the compiler then knows it is warp-uniform, so the role branches below are uniform
  //  branches and the code inside them may use the uniform datapath - what CUTLASS calls canonical_warp_idx_sync)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const uint32_t rank = IWAE_PAIR ? cluster_ctarank() : 0u;          // 0 = leader (issues the MMAs), 1 = peer
  // Work unit of a CTA group: PAIR consecutive tiles (one per CTA).  Each group owns a CONTIGUOUS range of units:
  // consecutive tiles then belong to the same image (k/128 tiles per image), so the per-image constants (x, mu, std:
  // ~3k cycles of global-load latency and two CTA barriers per tile when reloaded) are fetched once per image.
  // (tile indices fit 32 bits: n_tiles < 2^31 is checked on the host)
  const uint32_t n_units = (uint32_t)((args.n_tiles + PAIR - 1) / PAIR);
  const uint32_t n_groups = gridDim.x / PAIR, group = blockIdx.x / PAIR;
  const uint32_t per_group = (n_units + n_groups - 1u) / n_groups;
  const uint32_t mt_begin = min(group * per_group, n_units);
  const uint32_t mt_end = min(mt_begin + per_group, n_units);

  if (tid == 0) {
    for (int i = 0; i < NSTAGES; ++i) {
      // the leader's MMA thread must see BOTH halves of a slab: its own producer's expect_tx + the peer's relay
      mbar_init(&S.full[i], (IWAE_PAIR && rank == 0) ? 2 : 1);
      mbar_init(&S.empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&S.acc_full[i], 1); mbar_init(&S.acc_empty[i], NUM_EPI_WARPS * PAIR); }
    mbar_init(&S.act_ready, NUM_EPI_WARPS * PAIR);
    mbar_init(&S.act_early, NUM_EPI_WARPS * PAIR);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc_cg(&S.tmem_base, TMEM_COLS); tmem_relinquish_cg(); }
  tc_fence_before();
  __syncthreads();
#if IWAE_PAIR
  cluster_sync_all();           // the peer's mbarriers must be initialised before anything arrives on them remotely
#endif
  tc_fence_after();
  const uint32_t tmem_base = S.tmem_base;

#ifdef IWAE_PHASE_SHIFT_NS
  // Experiment: the second wave of CTA groups (co-resident with the first on the same SMs) starts half a tile late, so
  // that the two resident CTAs of an SM run their MMA-heavy and their XU-heavy layers out of phase.
  if (group >= n_groups / 2) __nanosleep(IWAE_PHASE_SHIFT_NS);
#endif

  if (warp == 0) {
    // ===================== weight producer: one elected thread =====================
    // One pass over the packed weights per tile; this CTA copies ITS share (N / PAIR columns) of every slab.  The
    // ring keeps NSTAGES slabs in flight ahead of the MMA issuer.  Like the MMA role below, the whole role runs
    // inside a single elected region (no per-slab elect / reconverge).
    if (elect_one_sync()) {
      uint32_t stage = 0, phase = 0;
      for (uint32_t mt = mt_begin; mt < mt_end; ++mt) {
        const uint8_t* src = args.wpack;
        for (int l = 0; l < P.nlayers; ++l) {
          const int nchunks = P.L[l].nchunks, k_pad = P.L[l].k_pad;
          for (int c = 0; c < nchunks; ++c) {
            const uint32_t cn = (uint32_t)P.L[l].chunk_n[c];
            for (int k0 = 0; k0 < k_pad; k0 += SLAB_K) {
              const uint32_t share = (cn / PAIR) * (uint32_t)min(SLAB_K, k_pad - k0) * 2u;   // bytes of one CTA's share
              mbar_wait(&S.empty[stage], phase ^ 1, 100, IWAE_PRODUCER_SLEEP);
              mbar_expect_tx(&S.full[stage], share);
              bulk_g2s(S.ring[stage], src + rank * share, share, &S.full[stage]);
              src += share * PAIR;
              if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 && rank != 0) {
    // ===================== peer CTA: relay "my half of slab s has landed" to the leader =====================
    // (a bulk copy signals the mbarrier of the CTA it writes to; the leader's full[s] needs the peer's arrival too)
    if (elect_one_sync()) {
      uint32_t stage = 0, phase = 0;
      const uint32_t full0 = map_to_rank(smem_u32(&S.full[0]), 0);
      for (uint32_t mt = mt_begin; mt < mt_end; ++mt)
        for (int l = 0; l < P.nlayers; ++l) {
          const int nslabs = (P.L[l].k_pad + SLAB_K - 1) / SLAB_K * P.L[l].nchunks;
          for (int i = 0; i < nslabs; ++i) {
            mbar_wait(&S.full[stage], phase, 150, IWAE_PRODUCER_SLEEP);
            mbar_arrive_cluster(full0 + stage * 8u);
            if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
          }
        }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    // The whole warp runs the role CONVERGED: control flow, barrier waits and descriptor arithmetic are warp-uniform,
    // so the compiler keeps them in uniform registers (the operand form tcgen05.mma reads); only the tcgen05
    // instructions themselves are predicated on the elected lane.  This needs the role branch itself to be a UNIFORM
    // branch (the warp index comes from a shuffle, see the top of the kernel): inside a branch on `tid >> 5` the
    // compiler must assume divergence and moves every operand with R2UR / VOTEU in front of each MMA.
    // (-DIWAE_MMA_CONVERGED=0: the role inside one elected thread, 7 R2UR per MMA, 2 % slower.)
#if IWAE_MMA_CONVERGED
    {
    const uint32_t issue = elect_one_sync();
#else
    if (elect_one_sync()) {
    const uint32_t issue = 1u;
#endif
    uint32_t stage = 0, phase = 0, cc = 0;
    uint32_t act_phase = 0, early_phase = 0;
    uint32_t empty_bits = 0;     // phase bit of acc_empty[buf] at bit buf (a register, not a dynamically indexed array)
    const uint32_t ring0 = smem_u32(S.ring[0]);
    const uint32_t z0 = smem_u32(S.sub.bufZ), p0 = smem_u32(S.sub.bufP);
    int tn = 0, tl = 0;

    for (uint32_t mt = mt_begin; mt < mt_end; ++mt, ++tl) {
      for (int l = 0; l < P.nlayers; ++l) {
        const int nchunks = P.L[l].nchunks, k_pad = P.L[l].k_pad, wait_act = P.L[l].wait_act, wait_split = P.L[l].wait_split;
        const uint32_t a_base0 = P.L[l].src ? p0 : z0;
        trace_ev<TRACE>(args, 1, tn, tl * 10000 + l * 100 + 0);
        for (int c = 0; c < nchunks; ++c) {
          const uint32_t buf = cc & 1;
          const uint32_t n = (uint32_t)P.L[l].chunk_n[c];
          const uint32_t idesc = make_idesc_cg(n);
          if (wait_act && c == 0) {
            if (wait_split) { mma_wait(&S.act_early, early_phase, 203); early_phase ^= 1; }     // first slab's columns only
            else { mma_wait(&S.act_ready, act_phase, 200); act_phase ^= 1; }
            trace_ev<TRACE>(args, 1, tn, tl * 10000 + l * 100 + 1);
          }
          // Every epilogue warp signals act_ready only AFTER it has drained (and signalled acc_empty for) all chunks of
          // the layer before: once act_ready has completed, both accumulator buffers are known to be free, so the first
          // two chunks of such a layer skip the acc_empty wait (a try_wait on a completed barrier still costs ~100
          // cycles of the issuing warp's dependent chain); the phase bookkeeping advances either way.
          if (!(wait_act && c < 2) || (c == 0 && P.L[l].early_two)) mma_wait(&S.acc_empty[buf], ((empty_bits >> buf) & 1u) ^ 1u, 201);
          empty_bits ^= 1u << buf;
          tc_fence_after();
          trace_ev<TRACE>(args, 1, tn, tl * 10000 + l * 100 + c * 10 + 2);
          const uint32_t d_tmem = tmem_base + buf * 128u;
          const uint32_t n_share = n / PAIR;               // columns of this CTA's share of B; K groups of 8 rows are n_share*16 bytes apart
          // descriptors: only the 14-bit start-address field (16-byte units) changes between MMAs
          uint64_t ad = make_smem_desc(a_base0, TILE_M * 16, 128);
          for (int k0 = 0; k0 < k_pad; k0 += SLAB_K) {
            if (wait_split && c == 0 && k0 == SLAB_K) { mma_wait(&S.act_ready, act_phase, 204); act_phase ^= 1; }
            mma_wait(&S.full[stage], phase, 202);
            tc_fence_after();
            uint64_t bd = make_smem_desc(ring0 + stage * SLAB_BYTES, n_share * 16u, 128);
            const int ksteps = min(SLAB_K, k_pad - k0) >> 4;
#pragma unroll
            for (int j = 0; j < SLAB_K / 16; ++j) {
              if (j < ksteps) {
                umma_f16_cg_pred(d_tmem, ad, bd, idesc, (k0 | j) != 0, issue);
                ad += 2u * TILE_M;            // 16 K columns = two 8-column groups of TILE_M*16 bytes
                bd += 2u * n_share;
              }
            }
            umma_commit_cg_pred(&S.empty[stage], issue);     // slab consumed (in both CTAs) when these MMAs retire
            if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit_cg_pred(&S.acc_full[buf], issue);
          trace_ev<TRACE>(args, 1, tn, tl * 10000 + l * 100 + c * 10 + 3);
          ++cc;
        }
      }
    }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps =====================
    // Eight warps.  Warps e and e+4 share one TMEM lane quarter (one row per lane) and split the COLUMNS of every
    // accumulator chunk between them (alternating 16-column groups).  Every log-density term is a sum over columns,
    // so each thread keeps one partial log-weight; the two halves of a row are added through shared memory at the
    // end of the tile.
    const int e_warp = warp - 2;             // 0..7
    const int half = (e_warp >> 2) & 1;      // which half of the column groups this warp takes
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int et = e_warp * 32 + lane;       // 0..255 among the epilogue threads
    SubTile& T = S.sub;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t full_bits = 0;      // phase bit of acc_full[buf] at bit buf
    uint32_t cc = 0;
    // the barriers the MMA issuer waits on live in the leader CTA
#if IWAE_PAIR
    const uint32_t act_ready_l = map_to_rank(smem_u32(&S.act_ready), 0), acc_empty_l = map_to_rank(smem_u32(&S.acc_empty[0]), 0);
    const uint32_t act_early_l = map_to_rank(smem_u32(&S.act_early), 0);
#else
    const uint32_t act_ready_l = smem_u32(&S.act_ready), acc_empty_l = smem_u32(&S.acc_empty[0]);
    const uint32_t act_early_l = smem_u32(&S.act_early);
#endif
    const int d1 = P.d1, nx = P.nx;
    const int nx_pad = (nx + 15) & ~15;
    const int d1_full = d1 & ~7;             // dims covered by complete 8-wide groups
    int tn = ((e_warp == 0 || e_warp == 7) && lane == 0) ? 0 : TRACE_SLOTS, tl = 0;     // two threads record: one per column half
    const int trole = (e_warp == 0) ? 2 : 0;
    uint32_t prev_img = 0xFFFFFFFFu;

    for (uint32_t mt = mt_begin; mt < mt_end; ++mt, ++tl) {
      trace_ev<TRACE>(args, trole, tn, tl * 10000 + 9000);
      // (tile indices fit 32 bits: n_tiles < 2^31 is checked on the host; 32-bit division is much cheaper)
      uint32_t tile = mt * PAIR + rank;
      const bool tile_ok = tile < (uint32_t)args.n_tiles;       // the last unit of a pair may be half empty
      if (!tile_ok) tile = (uint32_t)args.n_tiles - 1u;
      const uint32_t img32 = tile / (uint32_t)args.tiles_per_img;
      const int64_t img = img32;
      const int tin = (int)(tile - img32 * (uint32_t)args.tiles_per_img);
      const int s = tin * args.rows_per_tile + row;                 // sample index of this row
      const bool valid = tile_ok && (row < args.rows_per_tile) && (s < args.k);
      // the 8 normals of dims 8g..8g+7 of this row in noise stream 1 (first latent, d = d1) or 2 (second latent)
      auto noise8 = [&](uint32_t stream_id, uint32_t g, int d, float (&e)[8]) {
        if (INJECT) {
          const float* src = (stream_id == 1u ? args.eps1 : args.eps2) +
                             ((int64_t)min(s, args.k - 1) * args.n_images + img) * d + 8 * (int)g;
#pragma unroll
          for (int i = 0; i < 8; ++i) e[i] = (8 * (int)g + i < d) ? __ldg(src + i) : 0.f;
        } else {
          normal8<FAST>(args.seed, stream_id, (uint32_t)(args.img_offset + img), (uint32_t)s, g, e);
        }
      };

      // ---- per-image constants into smem (only when the image changes)
      if (img32 != prev_img) {
      prev_img = img32;
      epi_bar_sync();
      for (int j = et; j < nx_pad; j += 256)
        T.sgn[j] = (j < nx) ? 1.f - 2.f * __ldg(args.x + img * nx + j) : 0.f;
      float csum = 0.f, csum_p = 0.f;
      for (int j = et; j < d1; j += 256) {
        const float sd = __ldg(args.std1 + img * d1 + j);
        const float ls = logf(sd);
        T.mu[j] = __ldg(args.mu1 + img * d1 + j); T.sd[j] = sd;
        csum += ls;
        if (P.learned_prior) {
          const float ps = __ldg(args.pstd + img * d1 + j);
          const float lps = logf(ps);
          T.pmu[j] = __ldg(args.pmu + img * d1 + j); T.psd[j] = FAST ? 1.f / ps : ps;
          if (TERMS) csum_p += lps; else csum -= lps;
        }
      }
      if (FAST) {
        // constants of the FAST accumulation, once per tile: padded pixel columns each add
        // softplus(0) = ln 2 to the Bernoulli sum; sum_j log sd_j of q(z1|x) minus that of a learned
        // prior; the log(2 pi)/2 terms of lq1 and of a first-layer prior cancel, in the two-layer
        // model lq1's constant has no first-layer N(0,1) partner.
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
        if (lane == 0) T.red[8 + (e_warp & 7)] = csum;
        if (TERMS) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) csum_p += __shfl_xor_sync(0xffffffffu, csum_p, o);
          if (lane == 0) T.red_p[e_warp & 7] = csum_p;
        }
      }
      if (et == 0 && P.cond) T.label = (int)(unsigned char)(int)__ldg(args.y + img);
      epi_bar_sync();
      }
      trace_ev<TRACE>(args, trole, tn, tl * 10000 + 9050);

      // This thread's partial log-weight: log_w = lpxz + lp(prior terms) - lq(terms), all sums over
      // columns.  Precise mode adds log-probabilities term by term in TFP's op order; FAST mode
      // accumulates sums of squares / logs and applies the constants once (algebraically identical:
      // (z-mu)/sd == eps for a reparameterised sample).
      RowAcc<TERMS> acc;
      // latent samples of this row, for the TERMS outputs (fp32, before the fp16 rounding of the operand)
      const int64_t out_row = (int64_t)min(s, args.k - 1) * args.n_images + img;
      auto put_z = [&](float* base, int d, int k0, const float (&v)[8]) {
        if (TERMS && base && valid) {
          float* dstp = base + out_row * d + k0;
          if (k0 + 8 <= d && (d & 3) == 0) {            // rows are 16-byte aligned: one full 32-byte sector per thread
            *reinterpret_cast<float4*>(dstp) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(dstp + 4) = make_float4(v[4], v[5], v[6], v[7]);
          } else if (k0 + 8 <= d && (d & 1) == 0) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) *reinterpret_cast<float2*>(dstp + i) = make_float2(v[i], v[i + 1]);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) if (k0 + i < d) dstp[i] = v[i];
          }
        }
      };

      // Accurate modes (TERMS, or the libm-accurate tc_f16) keep z1 as an fp16 hi + lo PAIR for log p(z1|z2): that term
      // amplifies the 2^-11 rounding of the operand copy by |z1 - mu| / sigma^2 (measured: 1.4e-3 relative on the single
      // term at Glorot init).  The lo halves of dims 8g..8g+7 live in columns that are dead while q(z2|z1) / p(z1|z2)
      // run: K groups 14..25 of the activation buffer (those layers use at most 112 of its 208 columns) for g < 12, and
      // the zero-weight padding columns 104..111 of the z1 buffer for g = 12.
      constexpr bool Z1LO = TERMS || !FAST;
      auto put_lo = [&](int k0, const float (&v)[8]) {
        if (Z1LO && P.z1_lo) {
          float lo[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) lo[i] = v[i] - __half2float(__float2half_rn(v[i]));
          const int g = k0 >> 3;
          if (g < 12) store8(T.bufP, row, (14 + g) * 8, lo); else store8(T.bufZ, row, 104, lo);
        }
      };

      // ---- stage 0: z1 = mu + std*eps (A.2), lq1, lpz; fp16 z1 | onehot(y) | 1 | 0.. into bufZ
      {
        const int one_col = d1 + P.nc;
        float e1sum = 0.f, z1sum = 0.f;
        // complete group of 8 latent dims (FAST): no per-element predicates, vector constant loads
        auto full_group = [&](int k0, const float (&e)[8], float (&v)[8]) {
          const float4 m0 = *reinterpret_cast<const float4*>(&T.mu[k0]), m1 = *reinterpret_cast<const float4*>(&T.mu[k0 + 4]);
          const float4 s0 = *reinterpret_cast<const float4*>(&T.sd[k0]), s1 = *reinterpret_cast<const float4*>(&T.sd[k0 + 4]);
          const float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
          const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[i] = fmaf(e[i], ss[i], mm[i]);
            e1sum = fmaf(e[i], e[i], e1sum);
          }
          if (P.learned_prior) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float tp = (v[i] - T.pmu[k0 + i]) * T.psd[k0 + i]; z1sum = fmaf(tp, tp, z1sum); }
          } else if (!P.two_layer) {
#pragma unroll
            for (int i = 0; i < 8; ++i) z1sum = fmaf(v[i], v[i], z1sum);
          }
          store8(T.bufZ, row, k0, v);
          put_lo(k0, v);
          put_z(args.z1_out, d1, k0, v);
        };
        // early hand-off to the first layer: after this warp's last 8-dim group below SLAB_K
        auto maybe_early = [&](int kgrp) {
          if (P.stage0_early && kgrp < SLAB_K && kgrp + 16 >= SLAB_K) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(act_early_l);
          }
        };
        int k0 = 8 * half;
        if (FAST) {
          // two groups per iteration: two independent Philox / Box-Muller chains in flight per thread (the
          // generator is a long dependent chain - 10 rounds - so one chain alone issues at ~0.25 IPC)
          for (; k0 + 24 <= d1_full; k0 += 32) {
            float ea[8], eb[8], va[8], vb[8];
            noise8(1u, (uint32_t)(k0 >> 3), d1, ea);
            noise8(1u, (uint32_t)(k0 >> 3) + 2u, d1, eb);
            full_group(k0, ea, va);
            full_group(k0 + 16, eb, vb);
            maybe_early(k0 + 16);
          }
        }
        for (; k0 < P.kz_pad; k0 += 16) {
          float v[8];
          if (FAST && k0 < d1_full) {
            float ea[8];
            noise8(1u, (uint32_t)(k0 >> 3), d1, ea);
            full_group(k0, ea, v);
            maybe_early(k0);
            continue;
          } else {
            float e[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (k0 < d1) noise8(1u, (uint32_t)(k0 >> 3), d1, e);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int j = k0 + i;
              float val = 0.f;
              if (j < d1) {
                const float sd = T.sd[j], m = T.mu[j];
                const float z = fmaf(e[i], sd, m);
                if (FAST) {
                  e1sum = fmaf(e[i], e[i], e1sum);
                  if (P.learned_prior) { const float tp = (z - T.pmu[j]) * T.psd[j]; z1sum = fmaf(tp, tp, z1sum); }
                  else if (!P.two_layer) z1sum = fmaf(z, z, z1sum);
                } else {
                  const float t = z / sd - m / sd;                         // TFP op order
                  acc.template add<T_LQ1>(-0.5f * t * t - (LOG2PI_HALF_F + logf(sd)));       // lq(z1|x)
                  if (!P.two_layer) {
                    if (P.learned_prior) {
                      const float ps = T.psd[j];
                      const float tp = z / ps - T.pmu[j] / ps;
                      acc.template add<T_PRIOR>(-0.5f * tp * tp - (LOG2PI_HALF_F + logf(ps)));
                    } else {
                      acc.template add<T_PRIOR>(-0.5f * z * z - LOG2PI_HALF_F);
                    }
                  }
                }
                val = z;
              } else if (j < one_col) {
                val = (T.label == j - d1) ? 1.f : 0.f;
              } else if (j == one_col || j == one_col + 1) {      // two constant-1 columns: bias hi + lo
                val = 1.f;
              }
              v[i] = val;
            }
          }
          store8(T.bufZ, row, k0, v);
          if (k0 < d1) put_lo(k0, v);
          put_z(args.z1_out, d1, k0, v);
          maybe_early(k0);
        }
        if (FAST) { acc.template add<T_LQ1>(-0.5f * e1sum); acc.template add<T_PRIOR>(-0.5f * z1sum); }     // constants are added once per row at the end
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(act_ready_l);
      trace_ev<TRACE>(args, trole, tn, tl * 10000 + 9100);

      // ---- layers
      float bern_s = 0.f, bern_a = 0.f, bern_lg = 0.f;     // FAST Bernoulli accumulators (log2 units)
      for (int l = 0; l < P.nlayers; ++l) {
        const Layer& Ly = P.L[l];
        // layer constants once per layer (the asm statements' memory clobbers otherwise force re-reads per chunk)
        const int kind = Ly.kind, nchunks = Ly.nchunks;
        const bool wait_all = Ly.inplace && nchunks > 1, split_dst = Ly.split_dst != 0, signal_act = Ly.signal_act != 0;
        const bool early_sig = Ly.early_sig != 0;
        uint8_t* dst = Ly.dst ? T.bufP : T.bufZ;
        if (wait_all) {
          // all chunks' MMAs must have retired before the input buffer is overwritten
          for (int c = 0; c < nchunks; ++c) {
            const uint32_t b = (cc + c) & 1;
            mbar_wait(&S.acc_full[b], (full_bits >> b) & 1u, 300);
            full_bits ^= 1u << b;
          }
        }
        for (int c = 0; c < nchunks; ++c) {
          const uint32_t buf = cc & 1;
          if (!wait_all) {
            mbar_wait(&S.acc_full[buf], (full_bits >> buf) & 1u, 301);
            full_bits ^= 1u << buf;
          }
          tc_fence_after();
          trace_ev<TRACE>(args, trole, tn, tl * 10000 + l * 100 + c * 10 + 5);
          const uint32_t tbase = tmem_base + lane_addr + buf * 128u;
          const int cnt = Ly.chunk_cnt[c], off = Ly.chunk_off[c], cn = Ly.chunk_n[c];
          const int g_first = 16 * ((half + c) & 1);     // this warp's first 16-column group of the chunk

          if (kind == EPI_TANH) {
            // padded units have zero weights (tanh(0) = 0) and the constant-1 unit has weight 0 and
            // bias 30 (tanh(30) == 1.0f): no per-element special cases
            // split destination: unit j of the layer lands at K column j of the contiguous operand that starts at the z1
            // buffer (the activation buffer follows it in shared memory; bufZ holds exactly KZ_MAX columns)
            uint8_t* dstc = split_dst ? T.bufZ : dst;
            for_each_16cols(tbase, g_first, cn, [&](const uint32_t (&r)[16], int g) {
              float v[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = act_tanh<FAST>(__uint_as_float(r[i]));
              store16(dstc, row, off + g, v);
              // this warp's last group below SLAB_K of chunk 0: the first weight slab's K columns are complete for its rows
              if (early_sig && c == 0 && g < SLAB_K && g + 32 >= SLAB_K) {
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(act_early_l);
              }
              trace_ev<TRACE>(args, trole, tn, tl * 10000 + l * 100 + c * 10 + 8);
            });
          } else if (kind == EPI_GSAMPLE) {
            // Gaussian-head column layout: 16 columns per 8 latent dims, [mu x8 | log-std pre-activation x8]
            // (FAST: the log-std head is pre-scaled by log2(e) at pack time).
            // z2 = mu2 + std2*eps2 (iwae2.py:65); one 16-byte A-operand row segment per group.
                float e2sum = 0.f, z2sum = 0.f, lgsum = 0.f;
            for (int g8 = 8 * half; g8 < Ly.next_kpad; g8 += 16) {
              float v[8];
              if (FAST && g8 + 8 <= cnt) {
                // complete group of 8 latent dims: branch-free
                uint32_t r[16];
                tmem_ld16(tbase + 2 * g8, r);
                float e[8];
                noise8(2u, (uint32_t)((off + g8) >> 3), P.d2, e);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float sd = ex2_approx(__uint_as_float(r[8 + i])) + 1e-6f;
                  v[i] = fmaf(e[i], sd, __uint_as_float(r[i]));
                  e2sum = fmaf(e[i], e[i], e2sum);
                  z2sum = fmaf(v[i], v[i], z2sum);
                  lgsum += lg2_approx(sd);
                }
              } else if (g8 < cnt) {
                uint32_t r[16];
                tmem_ld16(tbase + 2 * g8, r);
                tmem_ld_wait();
                float e[8];
                noise8(2u, (uint32_t)((off + g8) >> 3), P.d2, e);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const int j = g8 + i;
                  float val = 0.f;
                  if (j < cnt) {
                    const float m = __uint_as_float(r[i]);
                    const float a = __uint_as_float(r[8 + i]);
                    if (FAST) {
                      const float sd = ex2_approx(a) + 1e-6f;
                      const float z = fmaf(e[i], sd, m);
                      e2sum = fmaf(e[i], e[i], e2sum);
                      z2sum = fmaf(z, z, z2sum);
                      lgsum += lg2_approx(sd);
                      val = z;
                    } else {
                      const float sd = expf(a) + 1e-6f;
                      const float z = e[i] * sd + m;
                      const float t = z / sd - m / sd;
                      acc.template add<T_LQ2>(-0.5f * t * t - (LOG2PI_HALF_F + logf(sd)));      // lqz2z1 (iwae2.py:118)
                      acc.template add<T_PRIOR>(-0.5f * z * z - LOG2PI_HALF_F);                  // lpz2   (iwae2.py:116)
                      val = z;
                    }
                  } else if (j == cnt || j == cnt + 1) {
                    val = 1.f;
                  }
                  v[i] = val;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = (g8 + i == cnt || g8 + i == cnt + 1) ? 1.f : 0.f;
              }
              store8(dst, row, g8, v);
              put_z(args.z2_out, cnt, g8, v);
            }
            // -lq2 + lpz2: the two log(2 pi)/2 constants cancel (TERMS mode adds them back per term at the end)
            if (FAST) { acc.template add<T_LQ2>(-0.5f * e2sum - LN2_F * lgsum); acc.template add<T_PRIOR>(-0.5f * z2sum); }
          } else if (kind == EPI_GLOGP) {
            // same column layout; log p(z1|z2) over dims off..off+cnt (iwae2.py:120), z1 re-read from bufZ
            float t2sum = 0.f, lssum = 0.f;
            int n_mine = 0;
            for (int g8 = 8 * ((half + c) & 1); g8 < cnt; g8 += 16) {
              uint32_t r[16];
              tmem_ld16(tbase + 2 * g8, r);
              tmem_ld_wait();
              const uint4 za = *reinterpret_cast<const uint4*>(T.bufZ + (size_t)((off + g8) >> 3) * (TILE_M * 16) + row * 16);
              const uint32_t zw[4] = {za.x, za.y, za.z, za.w};
              float zf[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const __half2 h2 = *reinterpret_cast<const __half2*>(&zw[i >> 1]);
                zf[i] = (i & 1) ? __high2float(h2) : __low2float(h2);
              }
              if (Z1LO && P.z1_lo) {
                const int g = (off + g8) >> 3;
                const uint8_t* lp = (g < 12) ? T.bufP + (size_t)(14 + g) * (TILE_M * 16) : T.bufZ + (size_t)13 * (TILE_M * 16);
                const uint4 zl = *reinterpret_cast<const uint4*>(lp + row * 16);
                const uint32_t lw4[4] = {zl.x, zl.y, zl.z, zl.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const __half2 h2 = *reinterpret_cast<const __half2*>(&lw4[i >> 1]);
                  zf[i] += (i & 1) ? __high2float(h2) : __low2float(h2);
                }
              }
              const int nv = min(8, cnt - g8);       // valid dims of this group (8 except possibly the last)
              n_mine += nv;
              if (FAST) {
                // sd = e^a + 1e-6 with a' = a*log2(e) in TMEM.  One MUFU per dim: inv = 2^-a' = e^-a, then
                // 1/sd = inv*(1 - 1e-6*inv) and log sd = a + 1e-6*inv to first order in 1e-6*inv (error
                // <= (1e-6*inv)^2).  Branch-free over the group; a warp vote re-does the group exactly in
                // the rare case that some sd is < ~1e-3.
                float inv[8];
                float invmax = 0.f;                       // padded dims have a' = 0 -> inv = 1: never "small"
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  inv[i] = ex2_approx(-__uint_as_float(r[8 + i]));
                  invmax = fmaxf(invmax, inv[i]);
                }
                if (!__any_sync(0xffffffffu, !(invmax < 1024.f))) {
                  if (nv == 8) {                          // complete group (warp-uniform): no per-dim masks
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                      const float z = zf[i];
                      const float c = 1e-6f * inv[i];
                      const float t = (z - __uint_as_float(r[i])) * fmaf(-c, inv[i], inv[i]);
                      t2sum = fmaf(t, t, t2sum);
                      lssum += fmaf(__uint_as_float(r[8 + i]), LN2_F, c);
                    }
                  } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                      const float z = zf[i];
                      const float isd = fmaf(-1e-6f * inv[i], inv[i], inv[i]);
                      const float ls = fmaf(__uint_as_float(r[8 + i]), LN2_F, 1e-6f * inv[i]);
                      const float t = (i < nv) ? (z - __uint_as_float(r[i])) * isd : 0.f;
                      t2sum = fmaf(t, t, t2sum);
                      lssum += (i < nv) ? ls : 0.f;
                    }
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    if (i < nv) {
                      const float z = zf[i];
                      const float sd = ex2_approx(__uint_as_float(r[8 + i])) + 1e-6f;
                      const float t = (z - __uint_as_float(r[i])) * rcp_approx(sd);
                      t2sum = fmaf(t, t, t2sum);
                      lssum += LN2_F * lg2_approx(sd);
                    }
                  }
                }
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  if (i < nv) {
                    const __half2 h2 = *reinterpret_cast<const __half2*>(&zw[i >> 1]);
                    const float z = (i & 1) ? __high2float(h2) : __low2float(h2);
                    const float m = __uint_as_float(r[i]);
                    const float sd = expf(__uint_as_float(r[8 + i])) + 1e-6f;
                    const float t = z / sd - m / sd;
                    acc.template add<T_LP12>(-0.5f * t * t - (LOG2PI_HALF_F + logf(sd)));         // lpz1z2 (iwae2.py:120)
                  }
                }
              }
            }
            if (FAST) acc.template add<T_LP12>(-0.5f * t2sum - lssum - (float)n_mine * LOG2PI_HALF_F);
          } else {
            // EPI_BERN: log p(x|z) = -sum_j softplus(s_j), s_j = (1-2x_j) l_j   (A.3 with x in {0,1})
            //   softplus(s) = max(s,0) + log(1 + e^-|s|),  max(s,0) = (s + |l|)/2,  |s| = |l|
            // FAST: the accumulators hold l' = l*log2(e) (folded into W3/b3 at pack time), so the
            // exponential is a bare ex2 of -|l'| and the sign only enters through one FFMA; the logs
            // are batched: one lg2 per <= 32 factors (1 + 2^-|l'|) in (1, 2].
            if (FAST) {
              float p0 = 1.f, p1 = 1.f, p2 = 1.f, p3 = 1.f, s0 = 0.f, s1 = 0.f, a0 = 0.f, a1 = 0.f;
              for_each_16cols(tbase, g_first, cn, [&](const uint32_t (&r)[16], int g) {
                const float4* sg4 = reinterpret_cast<const float4*>(&T.sgn[off + g]);
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
#ifndef IWAE_DIAG_NOSIGN
                  const float4 sg = sg4[i4];
#endif
                  const float l0 = __uint_as_float(r[4 * i4]), l1 = __uint_as_float(r[4 * i4 + 1]);
                  const float l2 = __uint_as_float(r[4 * i4 + 2]), l3 = __uint_as_float(r[4 * i4 + 3]);
#ifndef IWAE_DIAG_NOSIGN
                  s0 = fmaf(sg.x, l0, s0); s1 = fmaf(sg.y, l1, s1); s0 = fmaf(sg.z, l2, s0); s1 = fmaf(sg.w, l3, s1);
#endif
                  a0 += fabsf(l0); a1 += fabsf(l1); a0 += fabsf(l2); a1 += fabsf(l3);
#ifdef IWAE_DIAG_NOEX2
                  if (true) {                  // timing diagnostic only (wrong results): the pixel without its MUFU.EX2
                    p0 = fmaf(p0, -fabsf(l0), p0); p1 = fmaf(p1, -fabsf(l1), p1); p2 = fmaf(p2, -fabsf(l2), p2); p3 = fmaf(p3, -fabsf(l3), p3);
                  } else
#endif
                  if (i4 < 4 - IWAE_EX2_POLY) {
                    p0 = fmaf(p0, ex2_approx(-fabsf(l0)), p0); p1 = fmaf(p1, ex2_approx(-fabsf(l1)), p1);
                    p2 = fmaf(p2, ex2_approx(-fabsf(l2)), p2); p3 = fmaf(p3, ex2_approx(-fabsf(l3)), p3);
                  } else {
                    p0 = fmaf(p0, ex2_poly_neg(-fabsf(l0)), p0); p1 = fmaf(p1, ex2_poly_neg(-fabsf(l1)), p1);
                    p2 = fmaf(p2, ex2_poly_neg(-fabsf(l2)), p2); p3 = fmaf(p3, ex2_poly_neg(-fabsf(l3)), p3);
                  }
                }
                trace_ev<TRACE>(args, trole, tn, tl * 10000 + l * 100 + c * 10 + 8);
              });
              bern_s += s0 + s1; bern_a += a0 + a1;
              bern_lg += lg2_approx(p0 * p1) + lg2_approx(p2 * p3);
            } else {
              float bsum = 0.f;
              for_each_16cols(tbase, g_first, cn, [&](const uint32_t (&r)[16], int g) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float lgt = __uint_as_float(r[i]);
                  const float sg = T.sgn[off + g + i];
                  // padded columns (sgn 0, logit 0) must contribute nothing
                  bsum += (sg != 0.f) ? fmaxf(lgt * sg, 0.f) + log1pf(expf(-fabsf(lgt))) : 0.f;
                }
              });
              acc.template add<T_LPXZ>(-bsum);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc_empty_l + buf * 8u);
          trace_ev<TRACE>(args, trole, tn, tl * 10000 + l * 100 + c * 10 + 6);
          ++cc;
        }
        if (signal_act) {
          // (issuing this proxy fence BEFORE the acc_empty arrive of the last chunk, so that its MEMBAR does not wait for the
          //  remote arrive, measured 0.5 % slower: the accumulator is handed back later)
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(act_ready_l);
          trace_ev<TRACE>(args, trole, tn, tl * 10000 + l * 100 + 7);
        }
      }
      if (FAST) acc.template add<T_LPXZ>(-LN2_F * (0.5f * (bern_s + bern_a) + bern_lg));

      // ---- combine the two column halves of each row, then the tile's (max, sum exp) partial
      float* lwpart = reinterpret_cast<float*>(T.bufP);
      constexpr int NACC = TERMS ? 5 : 1;
      if (half == 1) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) lwpart[i * TILE_M + row] = acc.v[i];
      }
      epi_bar_sync();
      float lw = 0.f;
      if (half == 0) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc.v[i] += lwpart[i * TILE_M + row];
        if (FAST) {
          // constants of the FAST accumulation: padded pixel columns each added softplus(0) = ln 2 to the Bernoulli sum;
          // lq1 lacks -sum_j log sd_j - d1 log(2 pi)/2, a first-layer prior lacks -d1 log(2 pi)/2 (and a learned one
          // -sum_j log psd_j), lq2 and lpz2 lack -d2 log(2 pi)/2 each.  In the signed sum most of them cancel.
          float cq = 0.f, cp = 0.f;
#pragma unroll
          for (int w = 0; w < 8; ++w) { cq += T.red[8 + w]; if (TERMS) cp += T.red_p[w]; }
          if (TERMS) {
            acc.v[T_LPXZ] += (float)(nx_pad - nx) * LN2_F;
            acc.v[TERMS ? T_LQ1 : 0] += -cq - (float)d1 * LOG2PI_HALF_F;
            if (P.two_layer) {
              acc.v[T_PRIOR] += -(float)P.d2 * LOG2PI_HALF_F;
              acc.v[TERMS ? T_LQ2 : 0] += -(float)P.d2 * LOG2PI_HALF_F;
            } else {
              acc.v[T_PRIOR] += -cp - (float)d1 * LOG2PI_HALF_F;
            }
          } else {
            acc.v[0] += (float)(nx_pad - nx) * LN2_F + (P.two_layer ? (float)d1 * LOG2PI_HALF_F : 0.f) + cq;
          }
        }
        lw = acc.log_w();
        if (TERMS && valid) {
          if (args.term_lpxz) args.term_lpxz[out_row] = acc.v[T_LPXZ];
          if (args.term_prior) args.term_prior[out_row] = acc.v[T_PRIOR];
          if (args.term_lq1) args.term_lq1[out_row] = acc.v[TERMS ? T_LQ1 : 0];
          if (P.two_layer) {
            if (args.term_lq2) args.term_lq2[out_row] = acc.v[TERMS ? T_LQ2 : 0];
            if (args.term_lp12) args.term_lp12[out_row] = acc.v[TERMS ? T_LP12 : 0];
          }
        }
        if (!valid) lw = -INFINITY;
      }
      float mx = (half == 0) ? lw : -INFINITY;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (lane == 0 && half == 0) T.red[q] = mx;
      epi_bar_sync();
      mx = fmaxf(fmaxf(T.red[0], T.red[1]), fmaxf(T.red[2], T.red[3]));
      float se = (half == 0 && valid) ? expf(lw - mx) : 0.f;
      float sl = (half == 0 && valid) ? lw : 0.f;          // for vae_elbo = mean_k log_w (iwae1.py:119)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        se += __shfl_xor_sync(0xffffffffu, se, o);
        sl += __shfl_xor_sync(0xffffffffu, sl, o);
      }
      if (lane == 0 && half == 0) { T.red[4 + q] = se; T.red_lw[q] = sl; }
      epi_bar_sync();
      if (et == 0 && tile_ok) {
        args.partial[tile] = make_float2(mx, T.red[4] + T.red[5] + T.red[6] + T.red[7]);
        if (args.partial_lw) args.partial_lw[tile] = T.red_lw[0] + T.red_lw[1] + T.red_lw[2] + T.red_lw[3];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
#if IWAE_PAIR
  cluster_sync_all();           // neither CTA may exit (or free TMEM) while its peer can still arrive on its barriers
#endif
  if (warp == 1) tmem_dealloc_cg(tmem_base, TMEM_COLS);
}

// per image: logmeanexp over its tiles' partials (utils.logmeanexp, utils.py:6-8)
__global__ void combine_partials_kernel(const float2* __restrict__ partial, const float* __restrict__ partial_lw, int64_t N,
                                        int tiles_per_img, int k, float* __restrict__ out, float* __restrict__ out_vae) {
  const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= N) return;
  float mx = -INFINITY;
  for (int t = 0; t < tiles_per_img; ++t) mx = fmaxf(mx, partial[n * tiles_per_img + t].x);
  float se = 0.f;
  for (int t = 0; t < tiles_per_img; ++t) {
    const float2 p = partial[n * tiles_per_img + t];
    se += p.y * expf(p.x - mx);
  }
  out[n] = logf(se / (float)k) + mx;
  if (out_vae) {
    float sl = 0.f;
    for (int t = 0; t < tiles_per_img; ++t) sl += partial_lw[n * tiles_per_img + t];
    out_vae[n] = sl / (float)k;
  }
}

// ------------------------------------------------------------------------------------------------
// Weight packing: fp32 [in,out] Keras kernels -> fp16 slabs in the exact smem image, bias folded in as
// K-row `k_real`.  One launch per layer; one thread per packed half.
// ------------------------------------------------------------------------------------------------
struct PackLayer {
  const float *W, *b;       // primary kernel [k_real, ldw] and bias
  const float *W2, *b2;     // second head (log-std) for Gaussian layers, else null
  int k_real, k_pad, ldw;
  int nchunks, gauss;
  int one_unit;             // TANH layers: output units one_unit, one_unit+1 must evaluate to 1 (the next layer's bias inputs), else -1
  float scale, scale2;      // multiplies (W,b) / (W2,b2): log2(e) where the FAST epilogue works in base 2
  int chunk_n[MAX_CHUNKS], chunk_cnt[MAX_CHUNKS], chunk_off[MAX_CHUNKS];
  int64_t chunk_base[MAX_CHUNKS];   // half offset of the chunk's first slab inside the packed buffer
};

__global__ void pack_layer_kernel(PackLayer pl, __half* __restrict__ out) {
  for (int c = 0; c < pl.nchunks; ++c) {
    const int cn = pl.chunk_n[c], cnt = pl.chunk_cnt[c], off = pl.chunk_off[c];
    const int total = cn * pl.k_pad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      // i enumerates (k, n); locate the slab and the position inside it
      const int k = i / cn, n = i % cn;
      const int slab = k / SLAB_K, kin = k % SLAB_K;
      const int rows = min(SLAB_K, pl.k_pad - slab * SLAB_K);                     // K rows of this slab
      const int cs = cn / PAIR, rk = n / cs, nn = n % cs;                         // CTA `rk` of the pair holds columns [rk*cs, rk*cs + cs)
      const int64_t slab_base = pl.chunk_base[c] + (int64_t)slab * cn * SLAB_K;   // full slabs precede
      const int64_t pos = slab_base + (int64_t)rk * rows * cs + ((int64_t)(kin >> 3) * cs + nn) * 8 + (kin & 7);
      float v = 0.f;
      const float* W = pl.W; const float* b = pl.b; int col = -1;
      if (!pl.gauss) { if (n < cnt) col = off + n; }
      else {
        // 16 columns per 8 latent dims: [mu x8 | log-std pre-activation x8]
        const int dl = 8 * (n >> 4) + (n & 7);
        if (dl < cnt) { col = off + dl; if (n & 8) { W = pl.W2; b = pl.b2; } }
      }
      if (col >= 0) {
        const float sc = (W == pl.W2 && pl.gauss) ? pl.scale2 : pl.scale;
        if (k < pl.k_real) v = sc * W[(int64_t)k * pl.ldw + col];
        else if (k == pl.k_real || k == pl.k_real + 1) {
          // the bias rides on TWO constant-1 activations as an fp16 hi + lo pair (K is padded to a multiple
          // of 16, the second column is free): ~22 bits instead of 11 - the 784 output biases are logits of
          // pixel means (|b| up to 7), whose fp16 rounding alone was a 4e-5 relative error on the LLH
          const float bf = sc * b[col];
          const float hi = __half2float(__float2half_rn(bf));
          v = (k == pl.k_real) ? hi : bf - hi;
        }
      } else if (!pl.gauss && pl.one_unit >= 0 && (off + n == pl.one_unit || off + n == pl.one_unit + 1) && k == pl.k_real) {
        v = 30.f;           // tanh(30) == 1.0f: the constant-1 activation that carries the next layer's bias
      }
      out[pos] = __float2half_rn(v);
    }
  }
}

}  // namespace tc

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct TcPack {
  tc::Program prog;
  tc::PackLayer pl[tc::MAX_LAYERS];
  uint8_t* wpack = nullptr;
  size_t wpack_bytes = 0;
  float2* partial = nullptr;
  float* partial_lw = nullptr;
  size_t partial_cap = 0;
  bool attr_set = false;
  // cache of the packed image: valid for (parameter buffer, its content version, arithmetic mode)
  const float* packed_params = nullptr; uint64_t packed_version = 0; int packed_fast = -1;
};

static int round16(int v) { return (v + 15) & ~15; }

static void split_chunks(int n_pad, tc::Layer& L) {
  // N (multiple of 16) into chunks of <= 112 columns
  L.nchunks = 0;
  int off = 0;
  while (off < n_pad) {
    const int cn = std::min(tc::CHUNK_N_MAX, n_pad - off);
    L.chunk_n[L.nchunks] = cn; L.chunk_off[L.nchunks] = off; L.chunk_cnt[L.nchunks] = cn;
    ++L.nchunks; off += cn;
  }
}

static int build_program(const ModelP& m, TcPack& tp, const float* P, bool fast, bool side_ok) {
  using namespace tc;
  const iwae_config& c = m.cfg;
  Program& pr = tp.prog;
  memset(&pr, 0, sizeof(pr));
  memset(tp.pl, 0, sizeof(tp.pl));
  pr.two_layer = m.two_layer(); pr.cond = m.cond(); pr.learned_prior = m.learned_prior();
  pr.d1 = c.d1; pr.d2 = c.d2; pr.nc = m.cond() ? c.n_cond : 0; pr.nx = c.n_x;
  const int kz = round16(m.dec_in() + 1);
  pr.kz_pad = kz;
  IWAE_REQUIRE(kz <= KZ_MAX, "tensor-core path: d1 + n_cond + 1 must be <= %d", KZ_MAX);
  IWAE_REQUIRE(c.h1 + 1 <= KP_MAX && c.d1 <= D_MAX && c.n_x <= NX_MAX, "tensor-core path: needs h1 <= %d, d1 <= %d, n_x <= %d",
               KP_MAX - 1, D_MAX, NX_MAX);
  int nl = 0;
  bool side = false;
  auto tanh_layer = [&](const DenseP& D, int src, int dst, int k_pad) {
    Layer& L = pr.L[nl]; PackLayer& K = tp.pl[nl]; ++nl;
    L.kind = EPI_TANH; L.src = src; L.dst = dst; L.k_pad = k_pad; L.n_real = D.n_out;
    L.next_kpad = round16(D.n_out + 1);
    split_chunks(L.next_kpad, L);
    for (int i = 0; i < L.nchunks; ++i) L.chunk_cnt[i] = std::max(0, std::min(L.chunk_n[i], D.n_out - L.chunk_off[i]));
    K.W = P + D.w; K.b = P + D.b; K.k_real = D.n_in; K.k_pad = k_pad; K.ldw = D.n_out; K.gauss = 0;
    K.one_unit = D.n_out; K.scale = 1.f; K.scale2 = 1.f;
  };
  // one chunk (`which`) of a tanh layer as a layer of its own: the same columns of W, written to the same columns of dst
  auto tanh_part = [&](const DenseP& D, int src, int dst, int k_pad, int which, int side) {
    tanh_layer(D, src, dst, k_pad);
    Layer& L = pr.L[nl - 1];
    L.chunk_n[0] = L.chunk_n[which]; L.chunk_off[0] = L.chunk_off[which]; L.chunk_cnt[0] = L.chunk_cnt[which];
    L.nchunks = 1; L.side = side;
  };
  auto gauss_layer = [&](const BlockP& B, int kind, int src, int dst, int k_pad, int next_kpad, int per_chunk) {
    Layer& L = pr.L[nl]; PackLayer& K = tp.pl[nl]; ++nl;
    const int d = B.lmu.n_out;
    L.kind = kind; L.src = src; L.dst = dst; L.k_pad = k_pad; L.n_real = d; L.next_kpad = next_kpad;
    L.nchunks = 0;
    for (int off = 0; off < d; off += per_chunk) {
      const int cnt = std::min(per_chunk, d - off);
      L.chunk_cnt[L.nchunks] = cnt; L.chunk_off[L.nchunks] = off; L.chunk_n[L.nchunks] = 16 * ((cnt + 7) / 8);
      ++L.nchunks;
    }
    K.W = P + B.lmu.w; K.b = P + B.lmu.b; K.W2 = P + B.lstd.w; K.b2 = P + B.lstd.b;
    K.k_real = B.lmu.n_in; K.k_pad = k_pad; K.ldw = d; K.gauss = 1;
    K.one_unit = -1; K.scale = 1.f; K.scale2 = fast ? 1.44269504088896340736f : 1.f;
  };
  if (m.two_layer()) {
    IWAE_REQUIRE(c.h2 + 1 <= KP_MAX && c.d2 <= 56, "tensor-core path: needs h2 <= %d and d2 <= 56", KP_MAX - 1);
    const int kh2 = round16(c.h2 + 1), kd2 = round16(c.d2 + 1);
    // hi + lo copy of z1 for log p(z1|z2): the lo halves need activation-buffer columns 112..207 (free while the
    // 112-column-or-narrower layers of the two Gaussian blocks run) and z1-buffer columns 104..111 (padding)
    pr.z1_lo = (kh2 <= 112 && kd2 <= 112 && kz == 112 && c.d1 + 2 <= 104) ? 1 : 0;
    // Side layer: the first decoder layer p(x|z1) reads z1 only, and the columns 112.. of the activation buffer that its
    // second chunk writes are free while the two Gaussian blocks run (they stay inside the first 112; the fast evaluation
    // keeps no lo halves of z1 there).  That chunk therefore runs as a layer of its own right after the first layer of
    // q(z2|z1): its MMAs go out together with that layer's, and its epilogue fills the time the epilogue warps would
    // otherwise spend waiting for the MMAs of the second layer.  The decoder itself then starts with one chunk less.
    // Measured (IWAE_TC_SIDE=1, bit-identical results): 11.67 ms against 11.65 ms without on 2,368 images - with two resident
    // CTAs per SM the other CTA already fills that gap, so the order stays the plain one by default.
    static const bool want_side = getenv("IWAE_TC_SIDE") && atoi(getenv("IWAE_TC_SIDE")) != 0;
    side = side_ok && want_side && kh2 <= KZ_MAX && kd2 <= KZ_MAX && round16(c.h1 + 1) > CHUNK_N_MAX && round16(c.h1 + 1) <= 2 * CHUNK_N_MAX;
    tanh_layer(m.enc2.l1, 0, 1, kz);                                  // q(z2|z1) l1
    if (side) tanh_part(m.dec.d0, 0, 1, kz, 1, 1);                    // p(x|z1) dense0, units 112.. (side layer)
    tanh_layer(m.enc2.l2, 1, 1, kh2);                                 //          l2
    gauss_layer(m.enc2, EPI_GSAMPLE, 1, 1, kh2, kd2, 56);             //          (mu2|std2) -> z2
    IWAE_REQUIRE(pr.L[nl - 1].nchunks == 1, "tensor-core path: d2 too large");
    tanh_layer(m.dec21.l1, 1, 1, kd2);                                // p(z1|z2) l1
    tanh_layer(m.dec21.l2, 1, 1, kh2);                                //          l2
    gauss_layer(m.dec21, EPI_GLOGP, 1, 1, kh2, 0, 56);                //          (mu|std) -> log p(z1|z2)
  }
  const int kh1 = round16(c.h1 + 1);
  if (side) tanh_part(m.dec.d0, 0, 1, kz, 0, 0);                      // p(x|z1) dense0, units 0..111 (the rest ran as the side layer)
  else tanh_layer(m.dec.d0, 0, 1, kz);                                // p(x|z1) dense0
  tanh_layer(m.dec.d1, 1, 1, kh1);                                    //         dense1
  {                                                                   //         dense2 + Bernoulli
    Layer& L = pr.L[nl]; PackLayer& K = tp.pl[nl]; ++nl;
    L.kind = EPI_BERN; L.src = 1; L.dst = 1; L.k_pad = kh1; L.n_real = c.n_x; L.next_kpad = 0;
    split_chunks(round16(c.n_x), L);
    for (int i = 0; i < L.nchunks; ++i) L.chunk_cnt[i] = std::max(0, std::min(L.chunk_n[i], c.n_x - L.chunk_off[i]));
    IWAE_REQUIRE(L.nchunks <= MAX_CHUNKS, "tensor-core path: n_x too large");
    K.W = P + m.dec.d2.w; K.b = P + m.dec.d2.b; K.k_real = c.h1; K.k_pad = kh1; K.ldw = c.n_x; K.gauss = 0;
    K.one_unit = -1; K.scale = fast ? 1.44269504088896340736f : 1.f; K.scale2 = 1.f;
  }
  pr.nlayers = nl;
  // synchronisation flags: a layer waits for act_ready iff its input was produced by the previous
  // layer's epilogue (or by the sampling stage for the first layer); layers reading z1 later do not.
  for (int l = 0; l < nl; ++l) {
    Layer& L = pr.L[l];
    L.wait_act = (l == 0) || (L.src == 1);
    L.inplace = (L.kind != EPI_BERN && L.kind != EPI_GLOGP && L.src == L.dst) ? 1 : 0;
  }
  // a consumer's producer is the layer before it, side layers skipped
  for (int l = 1; l < nl; ++l)
    if (pr.L[l].wait_act) {
      int p = l - 1;
      while (p > 0 && pr.L[p].side) --p;
      pr.L[p].signal_act = 1;
    }
  // The second decoder layer (h1 -> h1 units, two chunks, in place) would have to wait for ALL its MMAs before the
  // first epilogue may overwrite the input.  The z1 buffer is dead once the first decoder layer's MMAs have retired,
  // so chunk 0 (exactly KZ_MAX columns) goes there instead and its epilogue overlaps the MMAs of chunk 1; chunk 1 is
  // written after ALL MMAs of the layer have retired, to the START of the activation buffer.  The activation buffer
  // follows the z1 buffer in shared memory and a K group of 8 columns is TILE_M*16 bytes in both, so the logit layer
  // reads ONE contiguous operand that begins at the z1 buffer (no per-MMA address select in the issue loop).
  {
    Layer& D1 = pr.L[nl - 2];
    Layer& BL = pr.L[nl - 1];
    static_assert(offsetof(SubTile, bufP) == offsetof(SubTile, bufZ) + BUFZ_BYTES, "bufP must follow bufZ");
    if (D1.kind == EPI_TANH && D1.inplace && D1.nchunks == 2 && D1.chunk_n[0] == KZ_MAX && pr.L[nl - 3].src == 0 &&
        BL.kind == EPI_BERN) {
      D1.split_dst = 1; D1.inplace = 0;
      BL.src_split = 1; BL.src = 0;
    }
  }
  // early hand-off (see Layer::early_sig): TANH producer whose chunk 0 covers the consumer's whole first slab, written
  // to where the consumer reads it (not the split destination), consumer with more than one slab
  static const bool no_early = getenv("IWAE_TC_NO_EARLY") != nullptr;
  for (int l = 0; l + 1 < nl && !no_early; ++l) {
    Layer& A = pr.L[l]; Layer& B = pr.L[l + 1];
    // (single-chunk producers only: the consumer's first chunk then accumulates into the OTHER TMEM buffer, which every
    //  epilogue warp has drained before it reached this layer; a two-chunk producer would still be reading it)
    // A two-chunk producer qualifies too (IWAE_TC_EARLY2=0 turns it off): the first slab's columns all come from its chunk 0,
    // whose TMEM buffer is the one the consumer's first chunk accumulates into - the issuer then waits for that buffer's
    // acc_empty (Layer::early_two).  The producer's chunk 0 must land where the consumer's operand begins: the same buffer,
    // or the z1 buffer for the split pair (second decoder layer -> logit layer).
    static const bool early2 = !(getenv("IWAE_TC_EARLY2") && atoi(getenv("IWAE_TC_EARLY2")) == 0);
    const bool same_base = (A.split_dst && B.src_split) || (!A.split_dst && !B.src_split && B.src == A.dst);
    if (A.kind == EPI_TANH && (A.nchunks == 1 || (A.nchunks == 2 && early2 && !A.inplace)) && A.signal_act && A.chunk_n[0] >= SLAB_K &&
        A.chunk_off[0] == 0 && B.wait_act && same_base && B.k_pad > SLAB_K && (A.nchunks == 2 || !A.split_dst)) {
      A.early_sig = 1; B.wait_split = 1; B.early_two = (A.nchunks == 2) ? 1 : 0;
    }
  }
  // ... and the sampling stage hands the first slab of z1 to the first layer the same way (needs every dim below SLAB_K to
  // be a latent dim written by complete 8-dim groups)
  if (!no_early && pr.L[0].wait_act && pr.L[0].src == 0 && pr.L[0].k_pad > SLAB_K && (c.d1 & ~7) >= SLAB_K) {
    pr.stage0_early = 1; pr.L[0].wait_split = 1;
  }
  // packed offsets
  int64_t off = 0;
  for (int l = 0; l < nl; ++l) {
    Layer& L = pr.L[l]; PackLayer& K = tp.pl[l];
    K.nchunks = L.nchunks;
    for (int i = 0; i < L.nchunks; ++i) {
      K.chunk_n[i] = L.chunk_n[i]; K.chunk_cnt[i] = L.chunk_cnt[i]; K.chunk_off[i] = L.chunk_off[i];
      K.chunk_base[i] = off;
      off += (int64_t)L.chunk_n[i] * L.k_pad;
    }
  }
  tp.wpack_bytes = (size_t)off * 2;
  return 0;
}

int tc_pack_weights(iwae_ctx* ctx, const float* params, bool fast, bool side_ok, cudaStream_t st) {
  if (!ctx->tc) ctx->tc = new TcPack();
  TcPack& tp = *ctx->tc;
  // The packed fp16 image depends only on the parameter values: when the caller has declared a content version of its
  // parameter buffer (iwae_set_params_version) and nothing changed since the last pack, the 9 pack launches are skipped
  // (they were ~15 % of the fixed per-pass cost that limits the 8-GPU efficiency).
  if (ctx->params_version != 0 && tp.wpack && tp.packed_params == params && tp.packed_version == ctx->params_version &&
      tp.packed_fast == (int)fast + 2 * (int)side_ok)
    return 0;
  IWAE_TRY(build_program(ctx->model, tp, params, fast, side_ok));
  if (!tp.wpack) CUDA_TRY(cudaMalloc((void**)&tp.wpack, tp.wpack_bytes));
  for (int l = 0; l < tp.prog.nlayers; ++l) {
    tc::pack_layer_kernel<<<64, 256, 0, st>>>(tp.pl[l], reinterpret_cast<__half*>(tp.wpack));
    LAUNCH_CHECK();
  }
  tp.packed_params = params; tp.packed_version = ctx->params_version; tp.packed_fast = (int)fast + 2 * (int)side_ok;
  return 0;
}

void tc_release(iwae_ctx* ctx) {
  if (!ctx->tc) return;
  if (ctx->tc->wpack) cudaFree(ctx->tc->wpack);
  if (ctx->tc->partial) cudaFree(ctx->tc->partial);
  if (ctx->tc->partial_lw) cudaFree(ctx->tc->partial_lw);
  delete ctx->tc;
  ctx->tc = nullptr;
}

int tc_llh_eval(iwae_ctx* ctx, const float* params, const float* x, const float* y, int64_t N, int k, uint64_t seed,
                int64_t img_offset, int fast, float* out_per_example, float* out_vae, cudaStream_t st, const float* eps1,
                const float* eps2, const TcTerms* terms) {
  const ModelP& m = ctx->model;
  const iwae_config& c = m.cfg;
  IWAE_REQUIRE(!m.cond() || y, "tensor-core path: CIWAE needs labels");
  // (the layer order with the side layer needs activation-buffer columns that the TERMS mode uses for the lo halves of z1)
  IWAE_TRY(tc_pack_weights(ctx, params, fast != 0, fast != 0 && !terms, st));
  TcPack& tp = *ctx->tc;
  Arena& ar = ctx->arena;

  // encoder pre-pass q(z1|x) (per image, k-independent, ~1% of the FLOPs)
  float *h1, *h2, *mu1, *std1, *pmu = nullptr, *pstd = nullptr;
  const float* xin = x; int64_t ldx = c.n_x;
  const int nc = m.cond() ? c.n_cond : 0;
  if (nc) {
    float* t; IWAE_TRY(ar.get(&t, (size_t)N * (c.n_x + nc)));
    IWAE_TRY(concat_onehot(x, c.n_x, c.n_x, y, (int)N, (int)N, nc, t, st));
    xin = t; ldx = c.n_x + nc;
  }
  IWAE_TRY(ar.get(&h1, (size_t)N * c.h1)); IWAE_TRY(ar.get(&h2, (size_t)N * c.h1));
  if (terms) { mu1 = terms->q_mu; std1 = terms->q_std; IWAE_REQUIRE(mu1 && std1, "tensor-core path: terms mode needs q_mu / q_std buffers"); }
  else { IWAE_TRY(ar.get(&mu1, (size_t)N * c.d1)); IWAE_TRY(ar.get(&std1, (size_t)N * c.d1)); }
  {
    // tensor cores (encoder_tc.cu) when the block fits, else the fp32 kernels
    const int rc = tc_encoder_fwd(ctx, params, m.enc1, xin, ldx, N, mu1, std1, st);
    if (rc > 0) return rc;
    if (rc < 0) IWAE_TRY(block_fwd_f32(params, m.enc1, xin, ldx, N, h1, h2, mu1, std1, st));
  }
  if (m.learned_prior()) {
    float* yoh;
    IWAE_TRY(ar.get(&yoh, (size_t)N * nc));
    IWAE_TRY(onehot_f32(y, (int)N, nc, yoh, st));
    IWAE_TRY(ar.get(&pmu, (size_t)N * c.d1)); IWAE_TRY(ar.get(&pstd, (size_t)N * c.d1));
    IWAE_TRY(block_fwd_f32(params, m.prior, yoh, nc, N, h1, h2, pmu, pstd, st));
  }

  tc::EvalArgs a;
  a.wpack = tp.wpack; a.x = x; a.y = y; a.mu1 = mu1; a.std1 = std1; a.pmu = pmu; a.pstd = pstd;
  a.tiles_per_img = (k + tc::TILE_M - 1) / tc::TILE_M;
  a.rows_per_tile = (k + a.tiles_per_img - 1) / a.tiles_per_img;
  a.n_tiles = N * a.tiles_per_img;
  IWAE_REQUIRE(a.n_tiles < ((int64_t)1 << 31), "tensor-core path: too many tiles in one call (N * ceil(k/128) must be < 2^31)");
  a.k = k; a.seed = seed; a.img_offset = img_offset;
  a.eps1 = eps1; a.eps2 = eps2; a.n_images = N;
  a.term_lpxz = a.term_prior = a.term_lq1 = a.term_lq2 = a.term_lp12 = a.z1_out = a.z2_out = nullptr;
  if (terms) {
    IWAE_REQUIRE(!eps1, "tensor-core path: terms mode takes Philox noise only");
    a.term_lpxz = terms->lpxz; a.term_prior = terms->prior; a.term_lq1 = terms->lq1; a.term_lq2 = terms->lq2;
    a.term_lp12 = terms->lp12; a.z1_out = terms->z1; a.z2_out = terms->z2;
  }
  IWAE_REQUIRE(!eps1 || !m.two_layer() || eps2, "tensor-core path: injected noise needs eps2 for two stochastic layers");
  const size_t need = (size_t)a.n_tiles * sizeof(float2);
  if (tp.partial_cap < need) {
    if (tp.partial) CUDA_TRY(cudaFree(tp.partial));
    if (tp.partial_lw) CUDA_TRY(cudaFree(tp.partial_lw));
    tp.partial = nullptr; tp.partial_lw = nullptr; tp.partial_cap = 0;
    CUDA_TRY(cudaMalloc((void**)&tp.partial, need));
    CUDA_TRY(cudaMalloc((void**)&tp.partial_lw, need / 2));
    tp.partial_cap = need;
  }
  a.partial = tp.partial;
  a.partial_lw = out_vae ? tp.partial_lw : nullptr;
  a.trace = nullptr;
  const char* trace_path = getenv("IWAE_TC_TRACE");
  if (trace_path && *trace_path) {
    CUDA_TRY(cudaMalloc((void**)&a.trace, sizeof(long long) * 4 * 3 * TRACE_SLOTS * 2));
    CUDA_TRY(cudaMemsetAsync(a.trace, 0, sizeof(long long) * 4 * 3 * TRACE_SLOTS * 2, st));
  }

  const size_t smem = sizeof(tc::SmemLayout);     // (the dynamic window is 128-byte aligned: no static shared memory in this kernel)
  using KernelT = void (*)(const tc::EvalArgs, const tc::Program);
  KernelT kern;
  if (terms) kern = fast ? tc::iwae_eval_tc_kernel<true, false, false, true> : tc::iwae_eval_tc_kernel<false, false, false, true>;
  else if (eps1) kern = fast ? tc::iwae_eval_tc_kernel<true, false, true, false> : tc::iwae_eval_tc_kernel<false, false, true, false>;
  else if (a.trace) kern = fast ? tc::iwae_eval_tc_kernel<true, true, false, false> : tc::iwae_eval_tc_kernel<false, true, false, false>;
  else kern = fast ? tc::iwae_eval_tc_kernel<true, false, false, false> : tc::iwae_eval_tc_kernel<false, false, false, false>;
  if (!tp.attr_set) {
    KernelT all[8] = {tc::iwae_eval_tc_kernel<true, false, false, false>, tc::iwae_eval_tc_kernel<false, false, false, false>,
                      tc::iwae_eval_tc_kernel<true, true, false, false>,  tc::iwae_eval_tc_kernel<false, true, false, false>,
                      tc::iwae_eval_tc_kernel<true, false, true, false>,  tc::iwae_eval_tc_kernel<false, false, true, false>,
                      tc::iwae_eval_tc_kernel<true, false, false, true>,  tc::iwae_eval_tc_kernel<false, false, false, true>};
    for (KernelT f : all) CUDA_TRY(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tp.attr_set = true;
  }
  // one CTA group (a pair of CTAs on the two SMs of a TPC) per unit of PAIR tiles, at most CTAS_PER_SM CTAs per SM
  int grid = tc::PAIR * (int)std::min<int64_t>((a.n_tiles + tc::PAIR - 1) / tc::PAIR, (int64_t)ctx->sm_count * tc::CTAS_PER_SM / tc::PAIR);
  if (const char* g = getenv("IWAE_TC_GRID")) { const int v = atoi(g) / tc::PAIR * tc::PAIR; if (v > 0 && v < grid) grid = v; }   // diagnostics only
  if (!ctx->ev_k0) { CUDA_TRY(cudaEventCreate(&ctx->ev_k0)); CUDA_TRY(cudaEventCreate(&ctx->ev_k1)); }
  CUDA_TRY(cudaEventRecord(ctx->ev_k0, st));
  kern<<<grid, tc::NUM_THREADS, smem, st>>>(a, tp.prog);
  LAUNCH_CHECK();
  CUDA_TRY(cudaEventRecord(ctx->ev_k1, st));
  ctx->ev_valid = true;
  tc::combine_partials_kernel<<<(int)((N + 255) / 256), 256, 0, st>>>(tp.partial, a.partial_lw, N, a.tiles_per_img, k, out_per_example,
                                                                         out_vae);
  LAUNCH_CHECK();
  if (a.trace) {     // debug only: synchronous dump of the timeline
    std::vector<long long> h(4 * 3 * TRACE_SLOTS * 2);
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaMemcpy(h.data(), a.trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaFree(a.trace));
    if (FILE* f = fopen(trace_path, "w")) {
      for (int cta = 0; cta < 4; ++cta)
        for (int role = 0; role < 3; ++role)
          for (int i = 0; i < TRACE_SLOTS; ++i) {
            const long long* p = &h[(((size_t)cta * 3 + role) * TRACE_SLOTS + i) * 2];
            if (p[1]) fprintf(f, "%d %d %lld %lld\n", cta, role, p[0], p[1]);
          }
      fclose(f);
    }
  }
  return 0;
}

