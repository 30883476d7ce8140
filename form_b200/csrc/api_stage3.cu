// C-ABI (include/formgpu.h): stage 3 - linearisation and error evaluation.
#include "api_common.hpp"

#include <algorithm>

using namespace formgpu;

namespace {

// doubles per pair in the result buffer: 91 for blocks, 1 for errors
inline size_t values_per_chunk_out(int values_per_chunk) { return values_per_chunk == 28 ? 91 : 1; }

size_t next_pow2(size_t v) {
  size_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Builds the request (poses by slot, pair descriptors, chunk table) in pinned
// memory, uploads it with one copy and fills LinArgs.  The chunk size only
// depends on the request's total size, so the partition - and with it the
// floating-point reduction order - is a pure function of the inputs.
int prepare(formgpu_ctx *ctx, const formgpu_pair *pairs, size_t n_pairs,
            const formgpu_scan_pose *poses, size_t n_poses, int values_per_chunk, LinArgs &a) {
  const int W = ctx->W;
  std::vector<int> pose_idx(W, -1);
  for (size_t p = 0; p < n_poses; ++p) {
    const int s = find_slot(ctx, poses[p].scan);
    if (s >= 0) pose_idx[s] = (int)p;
  }
  std::vector<LinPair> lp(n_pairs);
  size_t total = 0;
  for (size_t p = 0; p < n_pairs; ++p) {
    LinPair &d = lp[p];
    d = LinPair{};
    const int si = find_slot(ctx, pairs[p].i), sj = find_slot(ctx, pairs[p].j);
    if (si < 0 || sj < 0) continue; // unknown scans: no correspondences, zero block
    const PairEntry &e = ctx->h_pair_table[(size_t)sj * W + si];
    if (e.n_planar == 0 && e.n_point == 0) continue;
    if (pose_idx[si] < 0 || pose_idx[sj] < 0)
      return fail(ctx, FORMGPU_ERR_INVALID_ARG,
                  "linearize: no pose given for a scan of pair (" + std::to_string(pairs[p].i) +
                      ", " + std::to_string(pairs[p].j) + ")");
    d.slot_i = si;
    d.slot_j = sj;
    d.off_planar = e.off_planar;
    d.n_planar = e.n_planar;
    d.off_point = e.off_point;
    d.n_point = e.n_point;
    total += (size_t)e.n_planar + e.n_point;
  }
  // chunk length: about 8 CTAs per SM on a full-window request, never below 1024
  size_t chunk = (total + 148 * 8 - 1) / (148 * 8);
  chunk = std::max<size_t>(1024, (chunk + 127) / 128 * 128);
  std::vector<LinChunk> chunks;
  chunks.reserve(total / chunk + 2 * n_pairs + 1);
  for (size_t p = 0; p < n_pairs; ++p) {
    LinPair &d = lp[p];
    d.chunk_begin_planar = (int)chunks.size();
    for (uint32_t s = 0; s < d.n_planar; s += (uint32_t)chunk)
      chunks.push_back({(int)p, 0, d.off_planar + s, std::min<uint32_t>((uint32_t)chunk, d.n_planar - s)});
    d.n_chunks_planar = (int)chunks.size() - d.chunk_begin_planar;
    d.chunk_begin_point = (int)chunks.size();
    for (uint32_t s = 0; s < d.n_point; s += (uint32_t)chunk)
      chunks.push_back({(int)p, 1, d.off_point + s, std::min<uint32_t>((uint32_t)chunk, d.n_point - s)});
    d.n_chunks_point = (int)chunks.size() - d.chunk_begin_point;
  }

  const size_t pose_bytes = (size_t)W * 12 * sizeof(double);
  const size_t pair_bytes = (n_pairs * sizeof(LinPair) + 15) / 16 * 16;
  const size_t chunk_bytes = chunks.size() * sizeof(LinChunk);
  const size_t bytes = pose_bytes + pair_bytes + chunk_bytes;
  FORMGPU_CUDA(ctx, cudaEventSynchronize(ctx->ev_upload));
  int rc = ensure_upload(ctx, bytes);
  if (rc) return rc;
  rc = ensure_out(ctx, n_pairs);
  if (rc) return rc;
  if (chunks.size() * (size_t)values_per_chunk > ctx->partial_cap) {
    FORMGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->d_partials) cudaFree(ctx->d_partials);
    ctx->d_partials = nullptr;
    ctx->partial_cap = 0;
    const size_t cap = next_pow2(chunks.size() * 28);
    FORMGPU_CUDA(ctx, dev_alloc(&ctx->d_partials, cap));
    ctx->partial_cap = cap;
  }
  unsigned char *h = static_cast<unsigned char *>(ctx->h_upload);
  double *hp = reinterpret_cast<double *>(h);
  for (int s = 0; s < W; ++s) {
    if (pose_idx[s] >= 0) std::memcpy(hp + 12 * s, &poses[pose_idx[s]].pose, 12 * sizeof(double));
    else std::memset(hp + 12 * s, 0, 12 * sizeof(double));
  }
  if (n_pairs) std::memcpy(h + pose_bytes, lp.data(), n_pairs * sizeof(LinPair));
  if (!chunks.empty()) std::memcpy(h + pose_bytes + pair_bytes, chunks.data(), chunk_bytes);
  FORMGPU_CUDA(ctx, cudaMemcpyAsync(ctx->d_request, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
  FORMGPU_CUDA(ctx, cudaEventRecord(ctx->ev_upload, ctx->stream));
  unsigned char *d = static_cast<unsigned char *>(ctx->d_request);
  a.W = W;
  a.kp_cap = ctx->kp_cap;
  a.kq_cap = ctx->kq_cap;
  a.seg_planar = ctx->d_seg_planar;
  a.seg_point = ctx->d_seg_point;
  a.poses = reinterpret_cast<const double *>(d);
  a.pairs = reinterpret_cast<const LinPair *>(d + pose_bytes);
  a.chunks = reinterpret_cast<const LinChunk *>(d + pose_bytes + pair_bytes);
  a.n_pairs = (int)n_pairs;
  a.n_chunks = (int)chunks.size();
  a.inv_sigma2 = 1.0 / (ctx->P.sigma * ctx->P.sigma);
  a.partials = ctx->d_partials;
  a.out = ctx->h_out; // mapped pinned: the kernel writes the results straight to the host
  a.pair_counter = ctx->d_counters;
  a.done_counter = ctx->d_counters + ctx->counter_cap;
  int work = 0;
  for (const LinPair &d : lp) work += (d.n_chunks_planar + d.n_chunks_point) > 0;
  a.n_work_pairs = work;
  a.flag = ctx->h_flags + 0;
  a.seq = ++ctx->seq;
  // pairs without correspondences are never touched by a CTA: zero them here
  for (size_t p = 0; p < n_pairs; ++p)
    if (lp[p].n_chunks_planar + lp[p].n_chunks_point == 0)
      std::memset(ctx->h_out + p * (size_t)values_per_chunk_out(values_per_chunk), 0,
                  values_per_chunk_out(values_per_chunk) * sizeof(double));
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
  LinArgs a{};
  const int rc = prepare(ctx, pairs, n_pairs, poses, n_poses, 28, a);
  if (rc) return rc;
  linearize_launch(a, ctx->stream, ctx->prof);
  FORMGPU_CUDA(ctx, cudaGetLastError());
  if (a.n_chunks > 0) {
    const int w = wait_flag(ctx, 0, a.seq);
    if (w) return w;
  }
  std::memcpy(out91, ctx->h_out, n_pairs * 91 * sizeof(double));
  return FORMGPU_OK;
}

int formgpu_error(formgpu_ctx *ctx, const formgpu_pair *pairs, size_t n_pairs,
                  const formgpu_scan_pose *poses, size_t n_poses, double *out) {
  if (!ctx) return FORMGPU_ERR_INVALID_ARG;
  if (n_pairs == 0) return FORMGPU_OK;
  if (!pairs || !out || (n_poses && !poses))
    return fail(ctx, FORMGPU_ERR_INVALID_ARG, "formgpu_error: null argument");
  FORMGPU_CUDA(ctx, cudaSetDevice(ctx->device));
  ProfScope scope(ctx);
  LinArgs a{};
  const int rc = prepare(ctx, pairs, n_pairs, poses, n_poses, 1, a);
  if (rc) return rc;
  error_launch(a, ctx->stream, ctx->prof);
  FORMGPU_CUDA(ctx, cudaGetLastError());
  if (a.n_chunks > 0) {
    const int w = wait_flag(ctx, 0, a.seq);
    if (w) return w;
  }
  std::memcpy(out, ctx->h_out, n_pairs * sizeof(double));
  return FORMGPU_OK;
}

} // extern "C"
