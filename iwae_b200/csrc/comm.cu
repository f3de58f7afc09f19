// Library-owned NCCL communicator for the two exchange steps the path has:
//   * data-parallel training: ONE sum all-reduce of the flat gradient (~2 MB) per step, issued on the step's own
//     stream so that it is captured INSIDE the train step's CUDA graph (iwae_dp_train_step, abi.cu);
//   * sharded k-sample evaluation: one all-reduce of the fp64 LLH sum.
// NCCL is bound at run time with dlopen("libnccl.so.2"): a process that has already loaded PyTorch's bundled NCCL gets
// that very library (same SONAME), so the library and torch.distributed never mix two NCCL versions; a process
// without NCCL can still use every single-GPU entry point.  Only the types come from <nccl.h>.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle) return 0;
  const char* names[] = {getenv("IWAE_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n || !*n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  IWAE_REQUIRE(h, "NCCL not found (dlopen libnccl.so.2: %s); set IWAE_NCCL_LIB", dlerror());
  NcclApi a;
  a.handle = h;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
  a.AllReduce = (decltype(a.AllReduce))dlsym(h, "ncclAllReduce");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
  a.GetVersion = (decltype(a.GetVersion))dlsym(h, "ncclGetVersion");
  IWAE_REQUIRE(a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GetErrorString,
               "libnccl.so.2 lacks a required symbol");
  g_nccl = a;
  return 0;
}
}  // namespace

#define NCCL_TRY(expr)                                                                                  \
  do {                                                                                                  \
    ncclResult_t r__ = (expr);                                                                          \
    if (r__ != ncclSuccess) {                                                                           \
      iwae_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(r__));         \
      return 3;                                                                                         \
    }                                                                                                   \
  } while (0)

struct DpComm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  float* grads = nullptr;      // [n_params] reduced gradient of the data-parallel step (persistent: the step's graph holds it)
};

int comm_allreduce_f32(iwae_ctx* ctx, float* buf, int64_t n, cudaStream_t st) {
  IWAE_REQUIRE(ctx && ctx->dp && ctx->dp->comm, "no communicator: call iwae_comm_init first");
  NCCL_TRY(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat, ncclSum, ctx->dp->comm, st));
  return 0;
}
float* comm_grad_buffer(iwae_ctx* ctx) { return ctx->dp ? ctx->dp->grads : nullptr; }
int comm_world(const iwae_ctx* ctx) { return ctx->dp ? ctx->dp->world : 1; }

void comm_release(iwae_ctx* ctx) {
  if (!ctx->dp) return;
  if (ctx->dp->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->dp->comm);
  if (ctx->dp->grads) cudaFree(ctx->dp->grads);
  delete ctx->dp;
  ctx->dp = nullptr;
}

extern "C" int iwae_comm_unique_id(uint8_t* out128) {
  IWAE_REQUIRE(out128, "iwae_comm_unique_id: null argument");
  IWAE_TRY(load_nccl());
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  NCCL_TRY(g_nccl.GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  return 0;
}

extern "C" int iwae_comm_init(iwae_ctx* ctx, const uint8_t* id128, int32_t rank, int32_t world) {
  IWAE_REQUIRE(ctx && id128, "iwae_comm_init: null argument");
  IWAE_REQUIRE(world >= 1 && rank >= 0 && rank < world, "iwae_comm_init: bad rank %d / world %d", rank, world);
  IWAE_TRY(load_nccl());
  comm_release(ctx);
  ctx->dp = new DpComm();
  ctx->dp->rank = rank; ctx->dp->world = world;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NCCL_TRY(g_nccl.CommInitRank(&ctx->dp->comm, world, id, rank));
  CUDA_TRY(cudaMalloc((void**)&ctx->dp->grads, sizeof(float) * (size_t)std::max<int64_t>(ctx->model.n_params, 1)));
  return 0;
}

extern "C" int iwae_comm_destroy(iwae_ctx* ctx) {
  IWAE_REQUIRE(ctx, "iwae_comm_destroy: null context");
  comm_release(ctx);
  return 0;
}

extern "C" int iwae_comm_info(const iwae_ctx* ctx, int32_t* rank, int32_t* world, int32_t* nccl_version) {
  IWAE_REQUIRE(ctx, "iwae_comm_info: null context");
  if (rank) *rank = ctx->dp ? ctx->dp->rank : 0;
  if (world) *world = ctx->dp ? ctx->dp->world : 1;
  if (nccl_version) { int v = 0; if (g_nccl.GetVersion) g_nccl.GetVersion(&v); *nccl_version = v; }
  return 0;
}

extern "C" int iwae_allreduce_sum_f32(iwae_ctx* ctx, float* buf, int64_t n, void* stream) {
  IWAE_REQUIRE(buf && n >= 0, "iwae_allreduce_sum_f32: bad argument");
  return comm_allreduce_f32(ctx, buf, n, (cudaStream_t)stream);
}

extern "C" int iwae_allreduce_sum_f64(iwae_ctx* ctx, double* buf, int64_t n, void* stream) {
  IWAE_REQUIRE(buf && n >= 0, "iwae_allreduce_sum_f64: bad argument");
  IWAE_REQUIRE(ctx && ctx->dp && ctx->dp->comm, "no communicator: call iwae_comm_init first");
  NCCL_TRY(g_nccl.AllReduce(buf, buf, (size_t)n, ncclDouble, ncclSum, ctx->dp->comm, (cudaStream_t)stream));
  return 0;
}
