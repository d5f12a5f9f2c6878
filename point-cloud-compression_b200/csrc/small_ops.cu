// Small fp32 pieces around the tensor-core layers: per-cloud vectors and the "folded" first layer of a stage whose input is
// cat([per-point coordinates, one per-cloud vector tiled over the points]).
//
//   pcc_linear_small_f32   out[M, N] = act(x[M, K] . w[N, K]^T + b)      for SKINNY problems (M = clouds, not points):
//       PPPF_AE.enc_proj / dec_proj (/root/reference/PPPF_AE.py:122-123,139,145: Linear 1024 -> d and d -> 1024 on ONE row per
//       cloud), the latent part of FoldingNet's first Conv1d of each stage (PPPF_AE.py:100-109: the 1024-wide latent is the same
//       for every grid point, so W[:, n_local:] . latent is a per-cloud vector), the pooled-feature part of
//       AE.ConditionalProbabilityModel's first Conv2d (AE.py:115-116).  These are weight-bandwidth bound (every weight is used
//       M <= a few hundred times), so they run on the CUDA cores in fp32 -- which also keeps the value that is about to be
//       ROUNDED to a symbol (enc_proj) in the reference's precision.  Every output is one fixed-order sum (32 strided FMA chains
//       + a shuffle tree), so a row's result does not depend on how many rows are in the call (batch invariance: the entropy coder needs the decoder to
//       reproduce the encoder's PMFs bit for bit whatever the batch size).
//
//   pcc_fold_first_bf16    out[r, c] = relu(per_cloud[r / n_pts, c] + sum_j local[r, j] * w[c, j])   (bf16 rows)
//       the per-point part of those first layers (2 grid coordinates, 3 coarse coordinates, 3 centre coordinates) in fp32,
//       emitted as the bf16 A operand of the next layer's tensor-core GEMM (pcc_linear_bf16 / pcc_mlp_chain).  HBM bound
//       (2 * C bytes written per row).
#include <cuda_bf16.h>

#include "pcc_common.cuh"

namespace pcc {
namespace {

constexpr int LS_TM = 8, LS_WARPS = 4, LS_THREADS = LS_WARPS * 32;

// One WARP per (output column n, block of LS_TM rows): lane l owns k = l, l + 32, ... (the weight row is read once, coalesced,
// and reused for the 8 rows), partial sums are combined by a fixed xor-shuffle tree.  Every output is the same fixed-order sum
// whatever M is (batch invariance: the entropy coder needs the decoder to reproduce the encoder's PMFs bit for bit), and the
// M x N outputs spread over M / 8 x N warps -- the serial-chain form of round 2's first version kept 64 CTAs busy for 110 us on
// FoldingNet's 1024-wide latent.
__global__ void __launch_bounds__(LS_THREADS)
linear_small_kernel(const float *__restrict__ x, long long ldx, const float *__restrict__ w, long long ldw,
                    const float *__restrict__ bias, int M, int K, int N, int relu, float *__restrict__ out, long long ldo) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * LS_WARPS + (threadIdx.x >> 5);
    const int m0 = blockIdx.y * LS_TM;
    if (n >= N) return;
    const float *wr = w + static_cast<long long>(n) * ldw;
    float acc[LS_TM];
#pragma unroll
    for (int i = 0; i < LS_TM; ++i) acc[i] = 0.0f;
    for (int k = lane; k < K; k += 32) {
        const float wv = __ldg(wr + k);
#pragma unroll
        for (int i = 0; i < LS_TM; ++i) {
            const int r = m0 + i < M ? m0 + i : M - 1;   // rows past the end replay the last row and are not stored
            acc[i] = fmaf(__ldg(x + static_cast<long long>(r) * ldx + k), wv, acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < LS_TM; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(FULL_MASK, acc[i], o);
    }
    if (lane == 0) {
        const float b = bias ? __ldg(bias + n) : 0.0f;
#pragma unroll
        for (int i = 0; i < LS_TM; ++i) {
            if (m0 + i < M) {
                const float v = acc[i] + b;
                out[static_cast<long long>(m0 + i) * ldo + n] = relu ? fmaxf(v, 0.0f) : v;
            }
        }
    }
}

// thread = (row, 8-channel chunk): consecutive threads write consecutive 16-byte chunks (coalesced); the layer's weights sit in
// shared memory as (w_j0..w_j3) float4 per channel; rows and chunks are split with 32-bit arithmetic inside a block's row window
__global__ void __launch_bounds__(256)
fold_first_kernel(const float *__restrict__ local, int n_local, long long ld_local, const float *__restrict__ w, long long ldw,
                  const float *__restrict__ per_cloud, long long M, int n_pts, int C, int relu, __nv_bfloat16 *__restrict__ out,
                  long long ldo) {
    extern __shared__ float4 ws4[];   // [roundup(C, 8)]
    const int chunks = static_cast<int>(ldo >> 3);  // 8 channels (16 bytes) per thread
    const int cpad = (C + 7) & ~7;
    for (int c = threadIdx.x; c < cpad; c += 256) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < C) {
            v.x = __ldg(w + static_cast<long long>(c) * ldw);
            if (n_local > 1) v.y = __ldg(w + static_cast<long long>(c) * ldw + 1);
            if (n_local > 2) v.z = __ldg(w + static_cast<long long>(c) * ldw + 2);
            if (n_local > 3) v.w = __ldg(w + static_cast<long long>(c) * ldw + 3);
        }
        ws4[c] = v;
    }
    __syncthreads();
    const int rows_per_pass = 256 / chunks > 0 ? 256 / chunks : 1;       // ldo <= 2048: every thread owns one (row, chunk)
    const int my_row = threadIdx.x / chunks, my_chunk = threadIdx.x - my_row * chunks;
    if (my_row >= rows_per_pass) return;
    const int c0 = my_chunk * 8;
    for (long long r = static_cast<long long>(blockIdx.x) * rows_per_pass + my_row; r < M; r += static_cast<long long>(gridDim.x) * rows_per_pass) {
        uint32_t pk[4] = {0u, 0u, 0u, 0u};
        if (c0 < C) {
            const float *lp = local + r * ld_local;
            const float l0 = __ldg(lp), l1 = n_local > 1 ? __ldg(lp + 1) : 0.f, l2 = n_local > 2 ? __ldg(lp + 2) : 0.f,
                        l3 = n_local > 3 ? __ldg(lp + 3) : 0.f;
            const float *pc = per_cloud + (n_pts >= M ? 0 : r / n_pts) * C;
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                float v[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int c = c0 + 2 * h + q;
                    float a = 0.0f;
                    if (c < C) {
                        const float4 wv = ws4[c];
                        a = fmaf(l3, wv.w, fmaf(l2, wv.z, fmaf(l1, wv.y, fmaf(l0, wv.x, __ldg(pc + c)))));
                        if (relu) a = fmaxf(a, 0.0f);
                    }
                    v[q] = a;
                }
                const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[0], v[1]);
                pk[h] = *reinterpret_cast<const uint32_t *>(&b2);
            }
        }
        *reinterpret_cast<uint4 *>(out + r * ldo + c0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

}  // namespace
}  // namespace pcc

PCC_API int pcc_linear_small_f32(const float *x, int M, int K, int64_t ldx, const float *w, int64_t ldw, const float *bias, int N,
                                 int relu, float *out, int64_t ld_out, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(x && w && out, "pcc_linear_small_f32: null pointer");
    PCC_REQUIRE(M >= 1 && M <= 65535 * LS_TM && K >= 1 && N >= 1, "pcc_linear_small_f32: M=%d K=%d N=%d out of range", M, K, N);
    PCC_REQUIRE(ldx >= K && ldw >= K && ld_out >= N, "pcc_linear_small_f32: row pitch smaller than the row");
    const dim3 grid((N + LS_WARPS - 1) / LS_WARPS, (M + LS_TM - 1) / LS_TM);
    linear_small_kernel<<<grid, LS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, ldx, w, ldw, bias, M, K, N, relu, out, ld_out);
    return check_launch("linear_small_kernel");
}

PCC_API int pcc_fold_first_bf16(const float *local, int n_local, int64_t ld_local, const float *w, int64_t ldw, const float *per_cloud,
                                int64_t M, int n_pts, int C, int relu, void *out, int64_t ld_out, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(local && w && per_cloud && out, "pcc_fold_first_bf16: null pointer");
    PCC_REQUIRE(n_local >= 1 && n_local <= 4 && ld_local >= n_local && ldw >= n_local, "pcc_fold_first_bf16: n_local=%d outside [1,4]", n_local);
    PCC_REQUIRE(M >= 0 && n_pts >= 1 && M % n_pts == 0 && C >= 1, "pcc_fold_first_bf16: M=%lld must be a multiple of n_pts=%d",
                static_cast<long long>(M), n_pts);
    PCC_REQUIRE(ld_out >= C && ld_out % 8 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0,
                "pcc_fold_first_bf16: out must be 16-byte aligned with a row pitch that is a multiple of 8 elements");
    PCC_REQUIRE(ld_out <= 2048 && C <= 2048, "pcc_fold_first_bf16: at most 2048 output channels");
    if (M == 0) return 0;
    const int chunks = static_cast<int>(ld_out >> 3);
    const int rows_per_pass = 256 / chunks > 0 ? 256 / chunks : 1;
    long long blocks = (M + rows_per_pass - 1) / rows_per_pass;
    const long long cap = static_cast<long long>(num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    const size_t smem = static_cast<size_t>((C + 7) & ~7) * sizeof(float4);
    fold_first_kernel<<<static_cast<unsigned>(blocks), 256, smem, static_cast<cudaStream_t>(stream)>>>(
        local, n_local, ld_local, w, ldw, per_cloud, M, n_pts, C, relu, static_cast<__nv_bfloat16 *>(out), ld_out);
    return check_launch("fold_first_kernel");
}
