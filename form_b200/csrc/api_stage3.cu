// C-ABI (include/formgpu.h): stage 3 - linearisation and error evaluation.
#include "api_common.hpp"

#include <algorithm>

using namespace formgpu;

namespace {

// relative pose of scan j seen from scan i: R_i^T R_j and R_i^T (t_j - t_i)
void relative_pose(const formgpu_pose &Ti, const formgpu_pose &Tj, double rel[12]) {
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b)
      rel[3 * a + b] = Ti.R[a] * Tj.R[b] + Ti.R[3 + a] * Tj.R[3 + b] + Ti.R[6 + a] * Tj.R[6 + b];
  const double d[3] = {Tj.t[0] - Ti.t[0], Tj.t[1] - Ti.t[1], Tj.t[2] - Ti.t[2]};
  for (int a = 0; a < 3; ++a) rel[9 + a] = Ti.R[a] * d[0] + Ti.R[3 + a] * d[1] + Ti.R[6 + a] * d[2];
}

// Shared body of formgpu_linearize / formgpu_error.  Builds one task per pair that has
// correspondences, launches one cluster per task, then polls the per-pair flags the
// kernel raises in mapped host memory.
int run(formgpu_ctx *ctx, const formgpu_pair *pairs, size_t n_pairs, const formgpu_scan_pose *poses,
        size_t n_poses, bool error_only, double *out) {
  const int W = ctx->W;
  const size_t per_pair = error_only ? 1 : 91;
  std::vector<int> pose_idx(W, -1);
  for (size_t p = 0; p < n_poses; ++p) {
    const int s = find_slot(ctx, poses[p].scan);
    if (s >= 0) pose_idx[s] = (int)p;
  }
  int rc = ensure_out(ctx, n_pairs);
  if (rc) return rc;

  static thread_local std::vector<LinTask> tasks;
  tasks.clear();
  for (size_t p = 0; p < n_pairs; ++p) {
    const int si = find_slot(ctx, pairs[p].i), sj = find_slot(ctx, pairs[p].j);
    const PairEntry *e = (si >= 0 && sj >= 0) ? &ctx->h_pair_table[(size_t)sj * W + si] : nullptr;
    if (!e || (e->n_planar == 0 && e->n_point == 0)) {
      // no correspondences (or unknown scans): zero block, never touched by a CTA
      std::memset(ctx->h_out + p * per_pair, 0, per_pair * sizeof(double));
      continue;
    }
    if (pose_idx[si] < 0 || pose_idx[sj] < 0)
      return fail(ctx, FORMGPU_ERR_INVALID_ARG,
                  "linearize: no pose given for a scan of pair (" + std::to_string(pairs[p].i) +
                      ", " + std::to_string(pairs[p].j) + ")");
    LinTask t{};
    relative_pose(poses[pose_idx[si]].pose, poses[pose_idx[sj]].pose, t.rel);
    t.off_planar = e->off_planar;
    t.n_planar = e->n_planar;
    t.off_point = e->off_point;
    t.n_point = e->n_point;
    t.slot_j = sj;
    t.out_index = (int)p;
    tasks.push_back(t);
  }

  LinArgs a{};
  a.kp_cap = ctx->kp_cap;
  a.kq_cap = ctx->kq_cap;
  a.seg_planar = ctx->d_seg_planar;
  a.seg_point = ctx->d_seg_point;
  a.n_tasks = (int)tasks.size();
  // CTAs per pair: as many as keep the launch within about one wave (2 CTAs per SM)
  a.cluster = kLinCluster;
  while (a.cluster > 1 && (size_t)a.cluster * tasks.size() > 2 * 148) a.cluster >>= 1;
  a.inv_sigma2 = 1.0 / (ctx->P.sigma * ctx->P.sigma);
  a.out = ctx->h_out;
  a.flags = ctx->h_pair_flags;
  a.seq = ++ctx->seq;
  if (!tasks.empty()) {
    if ((int)tasks.size() <= kLinInlineTasks) {
      // the whole request rides in the kernel parameters: no upload, no dependent loads
      static thread_local LinInline inl;
      std::memcpy(inl.tasks, tasks.data(), tasks.size() * sizeof(LinTask));
      FORMGPU_CUDA(ctx, linearize_launch(a, &inl, error_only, ctx->stream, ctx->prof));
    } else {
      const size_t bytes = tasks.size() * sizeof(LinTask);
      FORMGPU_CUDA(ctx, cudaEventSynchronize(ctx->ev_upload));
      rc = ensure_upload(ctx, bytes);
      if (rc) return rc;
      std::memcpy(ctx->h_upload, tasks.data(), bytes);
      FORMGPU_CUDA(ctx, cudaMemcpyAsync(ctx->d_request, ctx->h_upload, bytes, cudaMemcpyHostToDevice,
                                        ctx->stream));
      FORMGPU_CUDA(ctx, cudaEventRecord(ctx->ev_upload, ctx->stream));
      a.tasks = static_cast<const LinTask *>(ctx->d_request);
      FORMGPU_CUDA(ctx, linearize_launch(a, nullptr, error_only, ctx->stream, ctx->prof));
    }
    // wait for every pair's flag (they complete in no particular order)
    unsigned spins = 0;
    for (const LinTask &t : tasks) {
      volatile unsigned long long *f = ctx->h_pair_flags + t.out_index;
      while (*f != a.seq) {
        if ((++spins & 0xfff) == 0) {
          const cudaError_t e = cudaStreamQuery(ctx->stream);
          if (e == cudaSuccess) {
            if (*f == a.seq) break;
            return fail(ctx, FORMGPU_ERR_STATE, "linearize kernel finished without publishing a pair");
          }
          if (e != cudaErrorNotReady)
            return fail(ctx, FORMGPU_ERR_CUDA, std::string("linearize kernel failed: ") + cudaGetErrorString(e));
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
      }
    }
  }
  std::memcpy(out, ctx->h_out, n_pairs * per_pair * sizeof(double));
  return FORMGPU_OK;
}

} // namespace

extern "C" {

int formgpu_linearize(formgpu_ctx *ctx, const formgpu_pair *pairs, size_t n_pairs,
                      const formgpu_scan_pose *poses, size_t n_poses, double *out91) {
  if (!ctx) return FORMGPU_ERR_INVALID_ARG;
  if (n_pairs == 0) return FORMGPU_OK;
  if (!pairs || !out91 || (n_poses && !poses))
    return fail(ctx, FORMGPU_ERR_INVALID_ARG, "formgpu_linearize: null argument");
  FORMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  ProfScope scope(ctx);
  return run(ctx, pairs, n_pairs, poses, n_poses, false, out91);
}

int formgpu_error(formgpu_ctx *ctx, const formgpu_pair *pairs, size_t n_pairs,
                  const formgpu_scan_pose *poses, size_t n_poses, double *out) {
  if (!ctx) return FORMGPU_ERR_INVALID_ARG;
  if (n_pairs == 0) return FORMGPU_OK;
  if (!pairs || !out || (n_poses && !poses))
    return fail(ctx, FORMGPU_ERR_INVALID_ARG, "formgpu_error: null argument");
  FORMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  ProfScope scope(ctx);
  return run(ctx, pairs, n_pairs, poses, n_poses, true, out);
}

} // extern "C"
