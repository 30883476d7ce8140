// Kernel argument blocks and host-side launchers shared between the .cu files.
#pragma once

#include "ctx.hpp"

namespace formgpu {

// ---- stage 1 (extract.cu, compiled with -fmad=false) ----
struct ExtractArgs {
  int rows, cols, cols_pad, words;
  int np, num_sectors, pps;
  int planar_per_sector, point_per_sector, min_points;
  int pr_cap, qr_cap;
  size_t kp_cap, kq_cap;
  double min_norm2, max_norm2, planar_threshold, radius;
  const float4 *scan;      // [B][rows*cols]
  uint32_t *valid_bits;    // [B][rows][words]
  float4 *row_box;         // [B][rows][words][2]  (lo, hi) of the valid points of a 32-column chunk
  uint16_t *planar_cols;   // [B][rows][pr_cap]
  int *planar_cnt;         // [B][rows]
  uint16_t *point_cols;    // [B][rows][qr_cap]
  int *point_cnt;          // [B][rows]
  float4 *normals;         // [B][rows][pr_cap]  (nx, ny, nz, keep)
  int *closest;            // [B][rows][pr_cap][2] or nullptr
  int *keep_cnt;           // [B][rows]
  PlanarRec *cur_planar;   // [B][kp_cap]
  PointRec *cur_point;     // [B][kq_cap]
  int *cur_counts;         // [B][2]
  uint8_t *dbg_valid, *dbg_pvalid; // [B][rows*cols] or nullptr
  float *dbg_curv;                 // [B][rows*cols] or nullptr
  // zero-copy results (mapped pinned host memory); host_planar/host_point may be null
  // when the caller only needs the counts (device-resident mode)
  PlanarRec *host_planar;
  PointRec *host_point;
  // ... or the caller's own page-locked buffers, filled with the API's f64 structs
  formgpu_planar_feat *host_planar_f64;
  formgpu_point_feat *host_point_f64;
  unsigned long long scan_idx;
  int *host_counts;                // [2]
  unsigned *done_counter;          // zero on entry, self-cleaning
  volatile unsigned long long *flag;
  unsigned long long seq;
  unsigned done_target;            // pack CTAs that share done_counter (rows x scans of the launch)
  unsigned pad_;
};

constexpr int kExtractMaxSectors = 64; // sectors per row the select kernel validates in order
/// Entries of a sector's speculative pick list in the select kernel (planar or point picks).
/// Picks of one sector are pairwise >= np columns apart (suppression +-(np-1)); the last sector is
/// the longest (it takes the remainder columns).
__host__ __device__ inline int extract_spec_cap(const ExtractArgs &a) {
  const int longest = a.cols - (a.num_sectors - 1) * a.pps;
  return (longest + a.np - 1) / a.np + 1;
}
cudaError_t extract_configure(const ExtractArgs &shape);
size_t extract_select_smem(int cols, int cols_pad, int words, int sectors, int spec_cap);
inline size_t extract_select_smem(const ExtractArgs &a) {
  return extract_select_smem(a.cols, a.cols_pad, a.words, a.num_sectors, extract_spec_cap(a));
}
size_t extract_normals_smem(int cols, int words, int pr_cap);
/// Launches the three stage-1 kernels.
void extract_launch(const ExtractArgs &a, int n_scans, cudaStream_t stream, Profiler &prof);
/// Same three kernels for n_items scans of DIFFERENT contexts (one ExtractArgs each, in
/// device memory); `shape` supplies the common geometry.
/// Launches with at least `many_rows_min` rows (rows x items) use the many-row variants: 128-thread
/// select CTAs and the thread-per-pick normals kernel, which trade per-row latency for rows in flight.
constexpr int kManyRowsMin = 1; // measured: +5 % batched throughput over switching at 2 x 148 rows
void extract_batch_launch(const ExtractArgs &shape, const ExtractArgs *items_dev, int n_items,
                          int many_rows_min, cudaStream_t stream, Profiler &prof);

// ---- stage 2 (map_assoc.cu, compiled with -fmad=false) ----
struct MapArgs {
  int type;            // 0 planar, 1 point
  int W;
  size_t kcap;         // keypoints per slot in the store
  const void *store;   // PlanarRec* or PointRec*
  const int *slot_off; // [W+1] exclusive prefix of store counts (device)
  const double *slot_pose;   // [W][12]
  const uint64_t *slot_scan; // [W]
  int n_total;
  double voxel_width;
  double inv_voxel_width; // 1 / voxel_width (host, IEEE): fast path of voxel_coord
  int cells;              // != 0: pass 4 orders the buckets of >= kCellMin points by cell code
  int pad_cells;
  HashSlot *hash;
  uint32_t hash_mask;
  WorldPoint *world_tmp; // store order; with cells: the FINAL voxel- and cell-sorted points
  uint32_t *world_slot;  // hash slot per store-order point; with cells: the final world_src
  uint32_t *world_src;   // voxel-sorted -> (window slot << 24 | k)
  WorldPoint *world;     // voxel-sorted; with cells: the per-voxel cell tables (pass 4)
  uint32_t *cursor;      // [0] allocation cursor, [1] occupied voxels (zeroed before the build)
  uint32_t *voxel_list;  // hash slots of the occupied voxels, in creation order (insert pass)
};
/// Cell-ordered buckets (map_assoc.cu, pass 4): a voxel holding kCellMin..kCellMax points gets
/// kCellFlag in its slot's count; its bucket is ordered by a kCellSub^3 cell code and carries a
/// table of 64 u16 end offsets.
constexpr uint32_t kCellFlag = 0x80000000u;
constexpr uint32_t kCellMin = 16, kCellMax = 65535;
constexpr int kCellSub = 4;
void map_build_launch(const MapArgs &planar, const MapArgs &point, cudaStream_t stream, Profiler &prof);
struct MapClearRegion {
  void *base;   // 16-byte aligned
  size_t bytes; // multiple of 16
  // batched rebuilds: the request of the context (slot offsets, poses, scan ids) rides in the
  // submission's argument upload and the clear kernel - the first launch of the rebuild - moves it
  // to its place, instead of one upload + event pair per sequence (copy_bytes a multiple of 8)
  size_t copy_src_off; // offset of the staged request inside the submission's argument ring
  void *copy_dst;
  size_t copy_bytes;
};
/// Rebuild of n_items maps in one pass of four launches: items_dev = [item][type] MapArgs,
/// regions_dev[item] = the cursor + hash tables to clear first.
/// `ring_dev`: base of the argument ring the regions' copy_src_off refer to.
void map_build_batch_launch(const MapArgs *items_dev, const MapClearRegion *regions_dev, const unsigned char *ring_dev,
                            int n_items, int max_points, uint32_t max_hash, size_t max_clear_bytes, bool cells,
                            cudaStream_t stream, Profiler &prof);

struct AssocArgs {
  int type;
  int n_query;
  int n_map;
  // queries this launch searches: [q_begin, q_end) (the whole scan)
  int q_begin, q_end;
  // point-sharded mode: the map holds only this rank's scans; the winner's shift rank travels in
  // bits 8..12 of MatchRec::slot so that assoc_combine can apply rule R5 across the ranks
  int pack_rank, pad_rank;
  const void *queries; // PlanarRec* / PointRec* of the current scan
  double pose[12];     // pose of the current scan
  double voxel_width;
  double inv_voxel_width; // 1 / voxel_width (host, IEEE): fast path of voxel_coord
  // cell-ordered buckets: the offset table of a flagged voxel whose bucket starts at `start` lives
  // in cell_tab[start .. start+3] (64 x u16 end offsets); nullptr = buckets are not cell-ordered
  const WorldPoint *cell_tab;
  const HashSlot *hash;
  uint32_t hash_mask;
  const WorldPoint *world;
  const uint32_t *world_src;
  MatchRec *match;
  // fused histogram: matches per (256-query block, matched slot), entry W = novel keypoints
  int W;
  double max_dist2, min_dist2;
  uint32_t *hist_cnt;      // [blocks256][W+1] atomic counters, zero on entry; nullptr: the
                           // histogram is built by assoc_hist_launch once all matches are there
};
/// Point-sharded mode: per query the rule-R5 minimum over the `world` candidate records the ranks
/// found in their sub-maps (gathered[r * n_query + q], shift rank packed into the slot), written to
/// a.match with the histogram the association kernels otherwise build themselves.
struct CombineArgs {
  AssocArgs a;                // n_query, match (output), hist_cnt, W, max / min dist
  const MatchRec *gathered;   // rank r's candidate of query q: gathered[r * stride + q]
  const uint64_t *slot_scan;  // [W] scan id of every window slot (tie-break, rule R4)
  int world;
  int stride;                 // records per rank (planar + point candidates of the scan)
};
void assoc_combine_launch(const CombineArgs &planar, const CombineArgs &point, cudaStream_t stream, Profiler &prof);
/// One sequence: `cell_search` = the one-thread-per-query search over the cell-ordered buckets (what
/// batched submits always use); otherwise one WARP per query scans whole buckets - the lower
/// latency for a single scan's ~30 k queries, which cannot fill the GPU either way.
void assoc_launch(const AssocArgs &planar, const AssocArgs &point, bool cell_search, cudaStream_t stream,
                  Profiler &prof);
/// `lanes` = lanes that share one query in the batched kernel (2, 4 or 8; kAssocLanes by default,
/// FORMGPU_ASSOC_LANES overrides it per batch for tuning).
constexpr int kAssocLanes = 4;
void assoc_batch_launch(const AssocArgs *items_dev, int n_items, int max_query, int lanes, cudaStream_t stream,
                        Profiler &prof);

struct SegmentArgs {
  int type;
  int W;
  int n_query;
  double max_dist2;
  double min_dist2;
  size_t kcap;
  const void *queries;  // current scan keypoints
  const void *store;    // keypoint store
  const MatchRec *match;
  const uint32_t *hist_cnt; // [blocks256][W+1] counters filled by the NN kernel
  uint32_t *hist_next;      // the other counter buffer, cleared after this launch
  size_t hist_bytes;
  uint32_t *host_pair_off;  // [W+1] mapped pinned: start of every pair in the segment
  uint32_t *host_pair_cnt;  // [W+1] mapped pinned: counts (entry W = novel keypoints)
  uint32_t *dev_pair_off;   // device copies for a linearisation queued right behind
  uint32_t *dev_pair_cnt;
  unsigned *done_counter;   // zero on entry, self-cleaning
  volatile unsigned long long *flag; // mapped pinned: set to `seq` once both types are published
  unsigned long long seq;
  float *seg;           // segment base of the current slot: [9 or 6][kcap]
};
void segment_build_launch(const SegmentArgs &planar, const SegmentArgs &point, cudaStream_t stream, Profiler &prof);
void segment_build_batch_launch(const SegmentArgs *items_dev, int n_items, int max_query,
                                cudaStream_t stream, Profiler &prof);

struct CommitArgs {
  int type;
  int W;
  int n_query;
  double min_dist2;
  const void *queries;
  const MatchRec *match;
  const uint32_t *hist_cnt; // counters of the association the matches come from
  void *store_dst;   // store slot base of the scan being appended to
  uint32_t dst_count; // keypoints already stored there
};
void commit_launch(const CommitArgs &planar, const CommitArgs &point, cudaStream_t stream, Profiler &prof);
void commit_batch_launch(const CommitArgs *items_dev, int n_items, int max_query, cudaStream_t stream,
                         Profiler &prof);

struct WorldExportArgs {
  int type;
  int W;
  size_t kcap;
  const void *store;
  const int *slot_off;  // [W+1] prefix in scan-id order
  const int *order;     // [W] slot at each position of that order
  const double *slot_pose;
  const uint64_t *slot_scan;
  int n_total;
  void *out; // formgpu_planar_feat* / formgpu_point_feat* (device)
};
void world_export_launch(const WorldExportArgs &a, cudaStream_t stream, Profiler &prof);

// ---- stage 3 (linearize.cu, FMA allowed: tolerance class) ----
constexpr int kLinThreads = 256;   // threads per CTA
constexpr int kLinCluster = 8;     // largest cluster (CTAs per scan pair)
constexpr int kLinInlineTasks = 48; // pairs that travel in the kernel parameters (6 KB)

struct LinTask { // one scan pair with at least one correspondence (144 B)
  double rel[12];      // R_i^T R_j row-major, then R_i^T (t_j - t_i): precomputed by the host
  uint32_t off_planar, n_planar, off_point, n_point; // ranges inside the segment of slot_j
  int slot_j;
  int out_index;       // position of the pair in the caller's list
  uint32_t dyn_slot_i_plus1; // != 0: the ranges are read from the device pair row of slot_i
                             // (association and linearisation queued back to back)
  uint32_t ctx_index;        // batched launches: which context's LinArgs the task uses
  int slot_i;                // window slot of scan i: the pair's entry in the moment cache
  int pad_;
  const double *entry;       // the pair's cache entry (ctx->d_moments + (slot_j * W + slot_i) * kMomentStride):
                             // resolved by the host so that the evaluation kernel's first load is the entry itself
};
static_assert(sizeof(LinTask) == 144, "LinTask size");

struct LinInline {
  LinTask tasks[kLinInlineTasks];
};

struct LinArgs {
  size_t kp_cap, kq_cap;
  const float *seg_planar; // [W][9][kp_cap]
  const float *seg_point;  // [W][6][kq_cap]
  const LinTask *tasks;    // device copy of the request when it does not fit the parameters
  const uint32_t *pair_row; // device [type][off|cnt][W+1] written by the last association
  int W;
  int n_tasks;
  int cluster;             // CTAs per pair for this launch: 1, 2, 4 or 8
  int debug_flags;         // timing experiments only: 1 = skip system fence, 2 = skip expansion,
                           // 4 = record %globaltimer checkpoints of pair 0 into debug_ts
  unsigned long long *debug_ts; // [16] mapped pinned
  double inv_sigma2;
  // mapped pinned host memory (zero-copy), sequence-tagged words: [n_pairs][182] for
  // blocks, [n_pairs][2] for errors; word = 32 bits of payload | (seq & 0xffffffff) << 32
  volatile unsigned long long *out;
  unsigned long long seq;
  // point-sharded mode: this context reduces only the shard_rank-th of
  // shard_world contiguous shares of every pair's correspondences; the caller sums the
  // per-pair blocks over the ranks (NCCL all-reduce of 91 * P doubles)
  int shard_rank, shard_world;
  // != nullptr: blocks / errors are written as plain doubles into DEVICE memory
  // ([n_pairs][91] or [n_pairs]) instead of the tagged host words, so that a collective
  // queued on the same stream can consume them without a host round trip
  double *out_plain;
  // pair-moment cache of the context (moments.cu): [W(j)][W(i)][kMomentStride] doubles
  const double *moments;
};
static_assert(sizeof(LinArgs) <= 128 && sizeof(LinArgs) % 8 == 0, "LinArgs: lin_warp_kernel copies it with lanes 16..31");
/// Per-device attributes of the stage-3 kernels (formgpu_create, with the device current).
cudaError_t linearize_configure();
cudaError_t linearize_launch(const LinArgs &a, const LinInline *inline_req, bool error_only,
                             cudaStream_t stream, Profiler &prof);
/// Batched launches: correspondences per warp slice the host aims for, and the slice table.
constexpr uint32_t kLinWarpSlice = 384;     // shortest slice
constexpr uint32_t kLinWarpSliceMax = 1536; // longest slice
constexpr uint32_t kLinWarpTarget = 3552;   // warps aimed for: 2 x 148 SMs x 12 resident warps
struct LinCta { // one warp-sized slice of a pair in a batched linearisation launch (32 B)
  uint32_t task;  // index into the task array
  uint32_t first; // index of the task's first slice (= base of its partial sums)
  uint16_t rank;  // this slice's share of the pair: rank of n_cta
  uint16_t n_cta; // slices of the pair
  uint32_t ctx_index;      // = tasks[task].ctx_index, so task and context load together
  const uint32_t *pair_row; // device [type][off|cnt][W+1] of the task's context (dynamic ranges)
  uint32_t dyn_slot_i_plus1; // = tasks[task].dyn_slot_i_plus1
  uint32_t row_stride;     // W + 1
};
static_assert(sizeof(LinCta) == 32, "LinCta size");
/// Tasks of several contexts in one launch, one warp per slice (eight per CTA):
/// ctx_args_dev[task.ctx_index] is the task's context; `partials` holds 56 doubles per
/// slice, `tickets` one zeroed counter per task.
cudaError_t linearize_warp_launch(const LinArgs *ctx_args_dev, const LinTask *tasks_dev,
                                  const LinCta *entries_dev, int n_entries, double *partials,
                                  unsigned *tickets, bool error_only, cudaStream_t stream, Profiler &prof);

// ---- stage 3, cached form (moments.cu, FMA allowed: tolerance class) ----
// Every association leaves, per pair (i, current scan), the POSE-INDEPENDENT second moments of
// its correspondences at a reference relative pose rel0 (header of moments.cu); a later
// linearisation / error evaluation at any poses is a 13x13 congruence of those moments - no
// correspondence is streamed again.
constexpr int kMomentPlanar = 91;    // packed upper triangle of sum phi phi^T, phi in R^13
constexpr int kMomentPoint = 28;     // packed upper triangle of sum zeta zeta^T, zeta in R^7
constexpr int kMomentStride = 132;   // doubles per entry: M_p[91] | M_q[28] | rel0[12] | counts (2 x u32)
constexpr int kMomentAcc = 73;       // distinct planar sums a thread accumulates
constexpr int kMomentPartial = 104;  // doubles per unit partial: 73 planar | 28 point | pad
// correspondences per warp-sized unit of work (MomentArgs::unit): batched submits amortise the
// per-unit reduction over 512 correspondences; a single sequence's association (~15 k accepted
// matches) would occupy ~40 warps of the GPU that way, so it is cut into units of 128
constexpr uint32_t kMomentUnit = 512, kMomentUnitSingle = 128;

struct MomentArgs { // one association of one context
  size_t kp_cap, kq_cap;
  const float *seg_planar;   // segment of the current slot: [9][kp_cap]
  const float *seg_point;    // [6][kq_cap]
  const uint32_t *pair_row;  // device [type][off|cnt][W+1] written by the scatter kernel just before
  const double *slot_pose;   // [W][12] poses of the last map rebuild (device)
  double pose_k[12];         // pose the current scan was associated at
  double *moments;           // cache row of the current slot: [W][kMomentStride]
  double *partials;          // [max_units][kMomentPartial]
  unsigned *tickets;         // [W], zero on entry, self-cleaning
  int W;
  int n_pairs;               // map slots listed in `slots`
  int shard_rank, shard_world;
  uint32_t unit, pad_unit;   // correspondences per unit (kMomentUnit or kMomentUnitSingle)
  unsigned char slots[kMaxWindow];
};
static_assert(sizeof(MomentArgs) % 8 == 0, "MomentArgs is copied in 8-byte words");
/// Upper bound of the units (= warps) one association with n_corr accepted matches over n_pairs
/// pairs can need.
inline int moment_max_units(size_t n_corr, int n_pairs, uint32_t unit) { return (int)(n_corr / unit) + n_pairs; }
void moments_launch(const MomentArgs &a, int max_units, cudaStream_t stream, Profiler &prof);
void moments_batch_launch(const MomentArgs *items_dev, int n_items, int max_units, cudaStream_t stream,
                          Profiler &prof);
/// Blocks / errors of the listed tasks from the moment cache: one warp per task.  Tasks travel in
/// the kernel parameters (inline_req, <= kLinInlineTasks) or in a.tasks (device memory).
cudaError_t eval_launch(const LinArgs &a, const LinInline *inline_req, bool error_only, cudaStream_t stream,
                        Profiler &prof);
cudaError_t eval_batch_launch(const LinArgs *ctx_args_dev, const LinTask *tasks_dev, int n_tasks,
                              bool error_only, cudaStream_t stream, Profiler &prof);

} // namespace formgpu
