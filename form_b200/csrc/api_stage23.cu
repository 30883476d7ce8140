// TEMPORARY stubs (replaced as stages 2/3 land).
#include "api_common.hpp"
using namespace formgpu;
extern "C" {
int formgpu_map_rebuild(formgpu_ctx *ctx, const formgpu_scan_pose *, size_t) { return fail(ctx, FORMGPU_ERR_UNSUPPORTED, "not implemented"); }
int formgpu_associate(formgpu_ctx *ctx, const formgpu_pose *, formgpu_pair_count *, size_t, size_t *) { return fail(ctx, FORMGPU_ERR_UNSUPPORTED, "not implemented"); }
int formgpu_get_matches(formgpu_ctx *ctx, int, formgpu_match *, size_t, size_t *) { return fail(ctx, FORMGPU_ERR_UNSUPPORTED, "not implemented"); }
int formgpu_commit_scan(formgpu_ctx *ctx, size_t *, size_t *) { return fail(ctx, FORMGPU_ERR_UNSUPPORTED, "not implemented"); }
int formgpu_remove_scans(formgpu_ctx *ctx, const uint64_t *, size_t) { return fail(ctx, FORMGPU_ERR_UNSUPPORTED, "not implemented"); }
int formgpu_get_keypoints(formgpu_ctx *ctx, int, uint64_t, void *, size_t, size_t *) { return fail(ctx, FORMGPU_ERR_UNSUPPORTED, "not implemented"); }
int formgpu_world_keypoints(formgpu_ctx *ctx, const formgpu_scan_pose *, size_t, formgpu_planar_feat *, size_t, size_t *, formgpu_point_feat *, size_t, size_t *) { return fail(ctx, FORMGPU_ERR_UNSUPPORTED, "not implemented"); }
int formgpu_linearize(formgpu_ctx *ctx, const formgpu_pair *, size_t, const formgpu_scan_pose *, size_t, double *) { return fail(ctx, FORMGPU_ERR_UNSUPPORTED, "not implemented"); }
int formgpu_error(formgpu_ctx *ctx, const formgpu_pair *, size_t, const formgpu_scan_pose *, size_t, double *) { return fail(ctx, FORMGPU_ERR_UNSUPPORTED, "not implemented"); }
}
