// Pieces shared by the farthest-point-sampling kernels (fps.cu, fps_bucket.cu).
#pragma once
#include "pcc_common.cuh"

namespace pcc {

// Optional fused epilogue: the sampled point itself (the index_points gather that follows every FPS call), optionally
// snapped to the octree grid: floor(c / cube) * cube + cube / 2 (octree_np.getDecodeFromPc, octree_np.py:114-133).
__device__ __forceinline__ void store_centre(float *o, float x, float y, float z, float cube) {
    if (cube > 0.0f) {
        const float h = __fmul_rn(cube, 0.5f);
        x = __fadd_rn(__fmul_rn(floorf(__fdiv_rn(x, cube)), cube), h);
        y = __fadd_rn(__fmul_rn(floorf(__fdiv_rn(y, cube)), cube), h);
        z = __fadd_rn(__fmul_rn(floorf(__fdiv_rn(z, cube)), cube), h);
    }
    o[0] = x;
    o[1] = y;
    o[2] = z;
}

// Scene-scale form (fps_bucket.cu): spatial buckets with exact skipping, one CTA per cloud.
bool fps_bucket_takes(int N);                       // shape gate of the bucketed form
int64_t fps_bucket_workspace_bytes(int B, int N);
int fps_bucket_run(const float *xyz, int B, int N, int npoint, const int64_t *start_idx, float init_dist, int64_t *out_idx,
                   float *out_xyz, float quant_cube, void *workspace, cudaStream_t st);

}  // namespace pcc
