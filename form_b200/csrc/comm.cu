// C-ABI (include/formgpu.h): point-sharded mode of ONE dense sequence over several GPUs of a
// node (BASELINE.json configs[4]) - the collectives.
//
// The reference has no distributed code; this is the B200-native extension the north star names:
// every rank keeps a replica of the sequence's keypoint store, but the reparative MAP - the large
// and expensive object of this configuration (1.5 M points, a million voxels, rebuilt every scan) -
// is sharded by scans: rank r transforms, hashes and cell-orders only the scans whose window slot
// is congruent to r.  Every rank searches its sub-map for all keypoints of the scan; one in-place
// ncclAllGather per keypoint type collects the candidates and a combine kernel applies the full
// rule-R5 key across ranks, so matches, segments, pair counts and the novel-keypoint commit are
// replicated and bit-identical to one GPU.  The pair moments are accumulated over the rank's
// share of every pair's correspondences, and a linearisation is the cached evaluation
// of those partial moments followed by ONE ncclAllReduce(sum, f64, 91 * P) of the blocks on the
// context's stream - the only bytes of stage 3 that cross NVLink.  NCCL is resolved at run time
// (dlopen of libnccl.so.2 - the copy the process already holds, e.g. PyTorch's, if any), so the
// library has no link-time dependency on it and single-GPU users never load it.
#include "api_common.hpp"

#include <dlfcn.h>

using namespace formgpu;

namespace {

// the slice of nccl.h this file uses (NCCL 2.x ABI)
typedef struct ncclComm *ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t; // ncclSuccess = 0
enum { kNcclSum = 0 };
enum { kNcclUint64 = 5, kNcclFloat64 = 8 }; // ncclDataType_t

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi &nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD); // a copy the process already loaded
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    api.error = std::string("NCCL not found: ") + dlerror();
    return api;
  }
  api.handle = h;
  auto sym = [&](const char *name) {
    void *p = dlsym(h, name);
    if (!p) api.error = std::string("NCCL symbol missing: ") + name;
    return p;
  };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  return api;
}

int nccl_fail(formgpu_ctx *ctx, const char *what, ncclResult_t r) {
  const NcclApi &api = nccl();
  return fail(ctx, FORMGPU_ERR_CUDA,
              std::string(what) + ": " + (api.GetErrorString ? api.GetErrorString(r) : "NCCL error"));
}

} // namespace

namespace formgpu {

int comm_allgather_matches(formgpu_ctx *ctx, int n_planar, int n_point) {
  NcclApi &api = nccl();
  const size_t chunk = (size_t)n_planar + (size_t)n_point; // a candidate per query and rank
  if (chunk == 0) return FORMGPU_OK;
  MatchRec *buf = ctx->d_gather;
  // in place: this rank's candidates already sit at their position
  const ncclResult_t r = api.AllGather(buf + chunk * ctx->comm_rank, buf, 2 * chunk, kNcclUint64,
                                       static_cast<ncclComm_t>(ctx->comm), ctx->stream);
  if (r != 0) return nccl_fail(ctx, "ncclAllGather(matches)", r);
  return FORMGPU_OK;
}

int comm_allreduce_f64(formgpu_ctx *ctx, double *dev, size_t count) {
  if (count == 0) return FORMGPU_OK;
  NcclApi &api = nccl();
  const ncclResult_t r = api.AllReduce(dev, dev, count, kNcclFloat64, kNcclSum,
                                       static_cast<ncclComm_t>(ctx->comm), ctx->stream);
  if (r != 0) return nccl_fail(ctx, "ncclAllReduce(blocks)", r);
  return FORMGPU_OK;
}

int comm_ensure_reduce(formgpu_ctx *ctx, size_t doubles) {
  if (doubles <= ctx->red_cap) return FORMGPU_OK;
  FORMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->d_red) cudaFree(ctx->d_red);
  if (ctx->h_red) cudaFreeHost(ctx->h_red);
  ctx->d_red = nullptr;
  ctx->h_red = nullptr;
  size_t cap = 91 * 64;
  while (cap < doubles) cap *= 2;
  FORMGPU_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&ctx->d_red), cap * sizeof(double)));
  FORMGPU_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_red), cap * sizeof(double), cudaHostAllocDefault));
  ctx->red_cap = cap;
  return FORMGPU_OK;
}

void comm_release(formgpu_ctx *ctx) {
  if (ctx->comm) {
    NcclApi &api = nccl();
    if (api.CommDestroy) api.CommDestroy(static_cast<ncclComm_t>(ctx->comm));
    ctx->comm = nullptr;
  }
  if (ctx->d_red) cudaFree(ctx->d_red);
  if (ctx->h_red) cudaFreeHost(ctx->h_red);
  if (ctx->d_gather) cudaFree(ctx->d_gather);
  ctx->d_gather = nullptr;
  ctx->d_red = nullptr;
  ctx->h_red = nullptr;
  ctx->red_cap = 0;
  ctx->comm_rank = 0;
  ctx->comm_world = 1;
  ctx->map_built = false; // the map of a sharded context holds a share of the scans only
}

} // namespace formgpu

extern "C" {

int formgpu_comm_unique_id(void *id128) {
  if (!id128) return FORMGPU_ERR_INVALID_ARG;
  NcclApi &api = nccl();
  if (!api.handle || !api.GetUniqueId) return FORMGPU_ERR_UNSUPPORTED;
  ncclUniqueId id;
  if (api.GetUniqueId(&id) != 0) return FORMGPU_ERR_CUDA;
  std::memcpy(id128, id.internal, sizeof(id.internal));
  return FORMGPU_OK;
}

int formgpu_comm_init(formgpu_ctx *ctx, const void *id128, int rank, int world) {
  if (!ctx) return FORMGPU_ERR_INVALID_ARG;
  if (!id128 || world < 1 || rank < 0 || rank >= world || world > 64)
    return fail(ctx, FORMGPU_ERR_INVALID_ARG, "formgpu_comm_init: need an id and 0 <= rank < world <= 64");
  if (ctx->comm) return fail(ctx, FORMGPU_ERR_STATE, "formgpu_comm_init: the context already has a communicator");
  if (ctx->shard_world > 1)
    return fail(ctx, FORMGPU_ERR_STATE, "formgpu_comm_init: formgpu_set_shard is active on this context");
  NcclApi &api = nccl();
  if (!api.handle || !api.error.empty())
    return fail(ctx, FORMGPU_ERR_UNSUPPORTED, api.error.empty() ? "NCCL not available" : api.error);
  FORMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId id;
  std::memcpy(id.internal, id128, sizeof(id.internal));
  ncclComm_t comm = nullptr;
  const ncclResult_t r = api.CommInitRank(&comm, world, id, rank);
  if (r != 0) return nccl_fail(ctx, "ncclCommInitRank", r);
  ctx->comm = comm;
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  ctx->map_built = false; // rebuilt sharded from now on
  if (world > 1)
    FORMGPU_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&ctx->d_gather),
                                 (size_t)world * (ctx->kp_cap + ctx->kq_cap) * sizeof(MatchRec)));
  return FORMGPU_OK;
}

int formgpu_comm_destroy(formgpu_ctx *ctx) {
  if (!ctx) return FORMGPU_ERR_INVALID_ARG;
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  comm_release(ctx);
  return FORMGPU_OK;
}

} // extern "C"
