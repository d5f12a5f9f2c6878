// Dispatch hook of the warp-specialised fixed-shape chains (chain_ws.cu) used by pcc_mlp_chain (mlp_chain.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pcc_b200.h"

namespace pcc {

// Runs the call on a specialised kernel when it matches one of the AE's three chains; *handled tells whether it did.
int ws_dispatch(const PccMlpInput *inputs, int n_inputs, int64_t rows, const PccMlpLayer *layers, int n_layers, int group,
                void *out, int out_dtype, cudaStream_t stream, bool *handled);

// CUtensorMap (passed as void*) over a [rows, cols] bf16 row-major tensor with row pitch `ld` elements; box = 64 columns x
// box_rows rows, CU_TENSOR_MAP_SWIZZLE_128B.  The encoder is resolved from the driver at run time (no libcuda link).
int make_tmap_bf16_2d(void *map, const void *ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

}  // namespace pcc
