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

// Co-resident multi-CTA form (fps.cu), N > 8192.  out_stride >= npoint is the row length of out_idx / out_xyz; md_out (nullable)
// [B, N] receives the running distances at exit (after the centres out[0 .. npoint - 2]).
int64_t fps_grid_workspace_bytes();
int fps_grid_run(const float *xyz, int B, int N, int npoint, const int64_t *start_idx, float init_dist, int64_t *out_idx,
                 float *out_xyz, float quant_cube, int out_stride, float *md_out, void *workspace, cudaStream_t st);

// Scene-scale form (fps_bucket.cu): spatial buckets with exact skipping, one CTA per cloud, after a head of iterations of the
// co-resident form.
bool fps_bucket_takes(int B, int N, int npoint);          // shape gate of the bucketed form
int64_t fps_bucket_workspace_bytes(int B, int N);
int fps_bucket_run(const float *xyz, int B, int N, int npoint, const int64_t *start_idx, float init_dist, int64_t *out_idx,
                   float *out_xyz, float quant_cube, void *workspace, cudaStream_t st);

}  // namespace pcc
