// Pooling pieces of the training path (the shared-MLP layers themselves run on gemm_ws.cu / wgrad_ws.cu):
//   pcc_groupmax_fwd_bf16   pooled[r / g, c] = max over the g rows of a group of x[r, c], arg = the first row that attains it --
//       the torch.max(..., dim)[0] that ends pn_kit.SetAbstraction / PointNet (/root/reference/pn_kit.py:139-143, 207) applied to
//       the bf16 activation the last layer wrote (training keeps it for the backward pass).
//   pcc_groupmax_bwd_bf16   dy[r, c] = dout[r / g, c] if r is the group's arg-max row (and, after a ReLU, the maximum is > 0),
//       else 0 -- the backward of that max, emitted as the bf16 operand of the weight / data gradient GEMMs.
// Both are HBM bound: one pass over the [M, C] activation.
#include <cuda_bf16.h>

#include "pcc_common.cuh"

namespace pcc {
namespace {

// thread = (group row, 8-column chunk); consecutive threads take consecutive chunks of the same rows (16-byte loads, coalesced)
__global__ void __launch_bounds__(256)
groupmax_fwd_kernel(const __nv_bfloat16 *__restrict__ x, long long ldx, long long n_groups, int chunks, int g,
                    float *__restrict__ pooled, short *__restrict__ arg) {
    const long long total = n_groups * chunks;
    for (long long t = blockIdx.x * 256ll + threadIdx.x; t < total; t += static_cast<long long>(gridDim.x) * 256) {
        const long long gr = t / chunks;
        const int ch = static_cast<int>(t - gr * chunks);
        const __nv_bfloat16 *src = x + gr * g * ldx + ch * 8;
        float best[8];
        short at[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            best[e] = -INFINITY;
            at[e] = 0;
        }
        for (int j = 0; j < g; ++j) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + j * ldx));
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float f = __uint_as_float((e & 1) ? (w[e >> 1] & 0xffff0000u) : (w[e >> 1] << 16));
                if (f > best[e]) {   // strict: the first row wins ties, like torch.max's backward
                    best[e] = f;
                    at[e] = static_cast<short>(j);
                }
            }
        }
        const long long o = gr * (chunks * 8ll) + ch * 8;
        *reinterpret_cast<float4 *>(pooled + o) = make_float4(best[0], best[1], best[2], best[3]);
        *reinterpret_cast<float4 *>(pooled + o + 4) = make_float4(best[4], best[5], best[6], best[7]);
        *reinterpret_cast<uint4 *>(arg + o) = make_uint4((at[0] & 0xffff) | (at[1] << 16), (at[2] & 0xffff) | (at[3] << 16),
                                                          (at[4] & 0xffff) | (at[5] << 16), (at[6] & 0xffff) | (at[7] << 16));
    }
}

// thread = (group, 8-column chunk of the ld_dy-wide output): the group's arg / dout / pooled values are read once and its g rows
// are written from registers (consecutive threads -> consecutive 16-byte chunks of the same row: coalesced); columns >= C are
// written as zeros (operand padding)
__global__ void __launch_bounds__(256)
groupmax_bwd_kernel(const float *__restrict__ dout, const float *__restrict__ pooled, const short *__restrict__ arg, long long M, int C,
                    int g, int out_chunks, __nv_bfloat16 *__restrict__ dy, long long ld_dy) {
    const long long n_groups = M / g, total = n_groups * out_chunks;
    for (long long t = blockIdx.x * 256ll + threadIdx.x; t < total; t += static_cast<long long>(gridDim.x) * 256) {
        const long long gr = t / out_chunks;
        const int ch = static_cast<int>(t - gr * out_chunks);
        __nv_bfloat16 *dst = dy + gr * g * ld_dy + ch * 8;
        if (ch * 8 >= C) {
            for (int j = 0; j < g; ++j) *reinterpret_cast<uint4 *>(dst + j * ld_dy) = make_uint4(0u, 0u, 0u, 0u);
            continue;
        }
        const long long s = gr * C + ch * 8;
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(arg + s));
        const float4 d0 = __ldg(reinterpret_cast<const float4 *>(dout + s)), d1 = __ldg(reinterpret_cast<const float4 *>(dout + s + 4));
        float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        if (pooled) {   // a ReLU preceded the max: no gradient where the maximum is not positive
            const float4 p0 = __ldg(reinterpret_cast<const float4 *>(pooled + s)), p1 = __ldg(reinterpret_cast<const float4 *>(pooled + s + 4));
            const float p[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) d[e] = p[e] > 0.0f ? d[e] : 0.0f;
        }
        const unsigned aw[4] = {a.x, a.y, a.z, a.w};
        int at[8];
        unsigned hv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            at[e] = static_cast<int>((e & 1) ? (aw[e >> 1] >> 16) : (aw[e >> 1] & 0xffffu));
            hv[e] = static_cast<unsigned>(__bfloat16_as_ushort(__float2bfloat16_rn(d[e])));
        }
        for (int j = 0; j < g; ++j) {
            unsigned w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) w[e] = (at[2 * e] == j ? hv[2 * e] : 0u) | ((at[2 * e + 1] == j ? hv[2 * e + 1] : 0u) << 16);
            *reinterpret_cast<uint4 *>(dst + j * ld_dy) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

int grid_of(long long total) {
    long long blocks = (total + 255) / 256;
    const long long cap = static_cast<long long>(num_sms()) * 16;
    return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace
}  // namespace pcc

PCC_API int pcc_groupmax_fwd_bf16(const void *x, int64_t M, int C, int64_t ldx, int group, float *pooled, int16_t *arg, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(x && pooled && arg, "pcc_groupmax_fwd_bf16: null pointer");
    PCC_REQUIRE(M >= 0 && C >= 8 && C % 8 == 0 && ldx >= C && ldx % 8 == 0 && group >= 1 && group <= 32767 && M % group == 0,
                "pcc_groupmax_fwd_bf16: bad shape M=%lld C=%d ldx=%lld group=%d", static_cast<long long>(M), C, static_cast<long long>(ldx), group);
    PCC_REQUIRE(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(pooled) % 16 == 0 && reinterpret_cast<uintptr_t>(arg) % 16 == 0,
                "pcc_groupmax_fwd_bf16: pointers must be 16-byte aligned");
    if (M == 0) return 0;
    const long long n_groups = M / group;
    groupmax_fwd_kernel<<<grid_of(n_groups * (C / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16 *>(x), ldx, n_groups, C / 8, group, pooled, arg);
    return check_launch("groupmax_fwd_kernel");
}

PCC_API int pcc_groupmax_bwd_bf16(const float *dout, const float *pooled, const int16_t *arg, int64_t M, int C, int group, void *dy,
                                  int64_t ld_dy, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(dout && arg && dy, "pcc_groupmax_bwd_bf16: null pointer");
    PCC_REQUIRE(M >= 0 && C >= 8 && C % 8 == 0 && ld_dy >= C && ld_dy % 8 == 0 && group >= 1 && M % group == 0,
                "pcc_groupmax_bwd_bf16: bad shape M=%lld C=%d ld_dy=%lld group=%d", static_cast<long long>(M), C, static_cast<long long>(ld_dy), group);
    PCC_REQUIRE(reinterpret_cast<uintptr_t>(dout) % 16 == 0 && reinterpret_cast<uintptr_t>(arg) % 16 == 0 && reinterpret_cast<uintptr_t>(dy) % 16 == 0 &&
                    (!pooled || reinterpret_cast<uintptr_t>(pooled) % 16 == 0),
                "pcc_groupmax_bwd_bf16: pointers must be 16-byte aligned");
    if (M == 0) return 0;
    const int out_chunks = static_cast<int>(ld_dy / 8);
    groupmax_bwd_kernel<<<grid_of(M / group * out_chunks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        dout, pooled, arg, M, C, group, out_chunks, static_cast<__nv_bfloat16 *>(dy), ld_dy);
    return check_launch("groupmax_bwd_kernel");
}
