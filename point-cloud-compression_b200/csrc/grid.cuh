// Uniform grid over one cloud (built by grid_build_kernel, chamfer_grid.cu): shared by the exact grid-pruned nearest-neighbour
// search of the Chamfer distance (chamfer_grid.cu) and the scene-scale kNN (knn.cu).
#pragma once
#include <cuda_runtime.h>

namespace pcc {

struct GridInfo {
    float mnx, mny, mnz, h, inv_h;
    int G;
    float margin;   // slack taken off every face / gap distance before it is trusted (rounding of cell assignment / faces)
    int pad;
};

__device__ __forceinline__ int cell_coord(float p, float mn, float inv_h, int G) {
    const int c = static_cast<int>((p - mn) * inv_h);
    return c < 0 ? 0 : (c >= G ? G - 1 : c);
}

// scratch: B * ((G^3 + 1) + 8) * 4 bytes for the multi-CTA build of big clouds (NULL: one CTA per cloud)
int grid_build_single(const float *pts, int B, int P, int G, float4 *sorted, unsigned *starts, GridInfo *info, unsigned *rowmask,
                      void *scratch, cudaStream_t st);

}  // namespace pcc
