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

/// Brackets one stage with CUDA events when profiling is on and counts launches.
struct StageScope {
  formgpu_ctx *ctx;
  int stage;
  StageScope(formgpu_ctx *c, int s) : ctx(c), stage(s) {
    if (ctx->profiling) cudaEventRecord(ctx->ev_a, ctx->stream);
  }
  void launches(int n) {
    ctx->launches += (uint64_t)n;
    ctx->prof[stage].launches += (uint64_t)n;
  }
  ~StageScope() {
    ctx->prof[stage].calls += 1;
    if (ctx->profiling) {
      cudaEventRecord(ctx->ev_b, ctx->stream);
      cudaEventSynchronize(ctx->ev_b);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b);
      ctx->prof[stage].ms += (double)ms;
    }
  }
};

template <typename T> inline cudaError_t dev_alloc(T **p, size_t count) {
  return cudaMalloc(reinterpret_cast<void **>(p), count * sizeof(T));
}

inline int find_slot(const formgpu_ctx *ctx, uint64_t scan) {
  auto it = ctx->slot_of.find(scan);
  return it == ctx->slot_of.end() ? -1 : it->second;
}

/// Ensure the pinned upload / result staging buffers are large enough.
int ensure_upload(formgpu_ctx *ctx, size_t bytes);
int ensure_out(formgpu_ctx *ctx, size_t pairs);

} // namespace formgpu
