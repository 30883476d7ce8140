// Helpers shared by the C-ABI translation units.
#pragma once

#include "ctx.hpp"
#include "kernels.hpp"

#include <cstdio>
#include <cstring>

namespace formgpu {

inline int fail(const formgpu_ctx *ctx, int code, const std::string &msg) {
  if (ctx) ctx->err = msg;
  return code;
}

#define FORMGPU_CUDA(ctx, expr)                                                         \
  do {                                                                                  \
    cudaError_t e_ = (expr);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return ::formgpu::fail((ctx), FORMGPU_ERR_CUDA,                                   \
                             std::string(#expr) + ": " + cudaGetErrorString(e_));      \
  } while (0)

/// Collects the event timings of an API call on scope exit (profiling mode only).
struct ProfScope {
  formgpu_ctx *ctx;
  explicit ProfScope(formgpu_ctx *c) : ctx(c) {}
  ~ProfScope() {
    if (ctx->prof.timing) ctx->prof.collect();
  }
};

template <typename T> inline cudaError_t dev_alloc(T **p, size_t count) {
  return cudaMalloc(reinterpret_cast<void **>(p), count * sizeof(T));
}

inline int find_slot(const formgpu_ctx *ctx, uint64_t scan) {
  auto it = ctx->slot_of.find(scan);
  return it == ctx->slot_of.end() ? -1 : it->second;
}

/// Spin until a kernel has published `seq` in the mapped flag.  Falls back to the
/// stream state every few thousand polls so a faulted kernel cannot hang the caller.
int wait_flag(formgpu_ctx *ctx, int which, unsigned long long seq);

/// Stage 3 building blocks (api_stage3.cu), also used by the fused association call.
void relative_pose(const formgpu_pose &Ti, const formgpu_pose &Tj, double rel[12]);
/// Launch one cluster per task (no wait).  Assigns and returns the sequence number.
int lin_launch(formgpu_ctx *ctx, const std::vector<LinTask> &tasks, bool error_only,
               unsigned long long *seq_out);
/// Spin until every word of the listed pairs carries the call's tag, decoding them into
/// dst (per_pair = 91 doubles for blocks, 1 for errors; dst[k] belongs to out_indices[k]).
int lin_wait(formgpu_ctx *ctx, const int *out_indices, size_t n, unsigned long long seq,
             size_t per_pair, double *dst);

/// Ensure the pinned upload / result staging buffers are large enough.
int ensure_upload(formgpu_ctx *ctx, size_t bytes);
int ensure_out(formgpu_ctx *ctx, size_t pairs);

} // namespace formgpu
