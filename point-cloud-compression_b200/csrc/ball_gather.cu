// Ball query and row gather / scatter-add for sm_100a.
//
// ball_query_kernel replaces pytorch3d.ops.ball_query (/root/reference/pointnet_sa_module.py:16-19,71):
//   the first K candidate indices IN INDEX ORDER with d2 < radius^2 (strict, fp32 radius*radius), padded with
//   -1 (idx) / 0 (d2).  One warp per query: 32 candidates per step from a shared-memory tile, ordered
//   compaction with ballot + popc, early exit (per warp, and per CTA once every warp has K hits).
//
// gather kernels replace pn_kit.index_points (/root/reference/pn_kit.py:332-360), pytorch3d knn_gather
//   (/root/reference/pointnet_sa_module.py:28) and torch.gather (pointnet_sa_module.py:68): pure data movement,
//   HBM/L2 bound; 16-byte vector path when C % 4 == 0.
#include <cuda_bf16.h>

#include "pcc_common.cuh"

namespace pcc {

constexpr int BQ_TILE = 1024;
constexpr int BQ_WARPS = 8;

__global__ void __launch_bounds__(BQ_WARPS * 32)
ball_query_kernel(const float *__restrict__ q, const float *__restrict__ p, int P1, int P2, int K, float r2,
                  int64_t *__restrict__ out_idx, float *__restrict__ out_d2) {
    __shared__ float tx[BQ_TILE], ty[BQ_TILE], tz[BQ_TILE];
    const int b = blockIdx.y;
    const unsigned lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int qi = blockIdx.x * BQ_WARPS + warp;
    const bool active = qi < P1;
    const float *pc = p + static_cast<size_t>(b) * P2 * 3;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        const float *qp = q + (static_cast<size_t>(b) * P1 + qi) * 3;
        qx = qp[0];
        qy = qp[1];
        qz = qp[2];
    }
    const size_t obase = (static_cast<size_t>(b) * P1 + (active ? qi : 0)) * K;
    int count = active ? 0 : K;  // inactive warps count as finished

    for (int t0 = 0; t0 < P2; t0 += BQ_TILE) {
        const int tn = min(BQ_TILE, P2 - t0);
        if (__syncthreads_and(count >= K)) break;  // also orders the previous tile's reads before the refill
        for (int e = threadIdx.x; e < tn * 3; e += BQ_WARPS * 32) {
            const float v = pc[static_cast<size_t>(t0) * 3 + e];
            const int pt = e / 3, c = e - pt * 3;
            (c == 0 ? tx : (c == 1 ? ty : tz))[pt] = v;
        }
        __syncthreads();
        for (int c0 = 0; c0 < tn && count < K; c0 += 32) {
            const int j = c0 + lane;
            float d = 0.f;
            bool hit = false;
            if (j < tn) {
                d = dist2_rn(qx, qy, qz, tx[j], ty[j], tz[j]);
                hit = d < r2;
            }
            const unsigned m = __ballot_sync(FULL_MASK, hit);
            if (m == 0u) continue;
            const int pos = count + __popc(m & ((1u << lane) - 1u));
            if (hit && pos < K) {
                out_idx[obase + pos] = static_cast<int64_t>(t0 + j);
                if (out_d2) out_d2[obase + pos] = d;
            }
            count += __popc(m);
        }
    }
    if (!active) return;
    for (int k = min(count, K) + lane; k < K; k += 32) {
        out_idx[obase + k] = -1;
        if (out_d2) out_d2[obase + k] = 0.0f;
    }
}

// out[r, :] = feat[(r / M) * N + idx[r], :], r over B*M rows.
template <typename VT>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const VT *__restrict__ feat, const int64_t *__restrict__ idx, long long rows, int N, int CV,
                   long long M, VT *__restrict__ out) {
    const long long total = rows * CV;
    for (long long e = blockIdx.x * 256ll + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256ll) {
        const long long r = e / CV;
        const int c = static_cast<int>(e - r * CV);
        const long long b = r / M;
        const long long j = idx[r];
        out[e] = feat[(b * N + j) * CV + c];
    }
}

__global__ void __launch_bounds__(256)
gather_bwd_kernel(const float *__restrict__ grad_out, const int64_t *__restrict__ idx, long long rows, int N, int C,
                  long long M, float *__restrict__ grad_feat) {
    const long long total = rows * C;
    for (long long e = blockIdx.x * 256ll + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256ll) {
        const long long r = e / C;
        const int c = static_cast<int>(e - r * C);
        const long long b = r / M;
        const long long j = idx[r];
        atomicAdd(grad_feat + (b * N + j) * C + c, grad_out[e]);
    }
}

static int grid_for(long long total) {
    long long blocks = (total + 255) / 256;
    const long long cap = static_cast<long long>(num_sms()) * 16;
    return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

__device__ __forceinline__ unsigned pack2_bf16(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const unsigned *>(&h);
}

// Grouping of a PointNet++ set-abstraction level in one pass (/root/reference/pointnet_sa_module.py:73-85): row r of the output
// is [features[b, idx[r], 0..C) | xyz[b, idx[r], 0..3) | zeros up to kpad] in bf16 -- the gathers, the torch.cat and the
// operand conversion of the GEMM that follows.  Negative indices (ball-query padding) read point 0 (pointnet_sa_module.py:27).
// One warp per group of 4 output rows: the 4 indices are read first (independent loads), then lanes walk each row's 16-byte
// chunks with the rows unrolled, so several rows' feature loads are in flight at once.
__global__ void __launch_bounds__(256)
gather_concat_bf16_kernel(const float *__restrict__ feat, int C, const float *__restrict__ xyz, const int64_t *__restrict__ idx,
                          long long rows, int N, long long M, int kpad, uint4 *__restrict__ out,
                          const float *__restrict__ centre, int nsample) {
    constexpr int R = 4;
    const int chunks = kpad >> 3, lane = threadIdx.x & 31;
    const bool vec = feat && (C & 7) == 0;
    // rows < 2^31 (checked by the launcher): the two divisions per row are 32-bit (the 64-bit form was ~100 instructions each,
    // more than the rest of the row's work)
    const unsigned Mu = static_cast<unsigned>(M), nsu = static_cast<unsigned>(nsample);
    for (long long r0 = (blockIdx.x * 8ll + (threadIdx.x >> 5)) * R; r0 < rows; r0 += static_cast<long long>(gridDim.x) * 8 * R) {
        long long src[R];
        const float *cq[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const unsigned r = static_cast<unsigned>(r0 + i < rows ? r0 + i : rows - 1);
            long long j = __ldg(idx + r);
            j = j < 0 ? 0 : j;
            src[i] = static_cast<long long>(r / Mu) * N + j;
            // recentred grouping (pppe_pcd_ae.py:600): row r belongs to query r / nsample, whose xyz is subtracted in fp32
            cq[i] = centre ? centre + static_cast<size_t>(r / nsu) * 3 : nullptr;
        }
        for (int ch = lane; ch < chunks; ch += 32) {
            const int c0 = ch * 8;
            uint4 q[R];
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const float *f = feat ? feat + src[i] * C : nullptr;
                const float *p = xyz ? xyz + src[i] * 3 : nullptr;
                float v[8];
                if (vec && c0 + 8 <= C) {
                    const float4 a = __ldg(reinterpret_cast<const float4 *>(f + c0));
                    const float4 d = __ldg(reinterpret_cast<const float4 *>(f + c0 + 4));
                    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = d.x, v[5] = d.y, v[6] = d.z, v[7] = d.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int c = c0 + e;
                        v[e] = (f && c < C) ? __ldg(f + c)
                               : (p && c >= C && c < C + 3) ? (cq[i] ? __fsub_rn(__ldg(p + (c - C)), __ldg(cq[i] + (c - C))) : __ldg(p + (c - C)))
                                                            : 0.0f;
                    }
                }
                q[i].x = pack2_bf16(v[0], v[1]);
                q[i].y = pack2_bf16(v[2], v[3]);
                q[i].z = pack2_bf16(v[4], v[5]);
                q[i].w = pack2_bf16(v[6], v[7]);
            }
#pragma unroll
            for (int i = 0; i < R; ++i)
                if (r0 + i < rows) out[(r0 + i) * chunks + ch] = q[i];
        }
    }
}

}  // namespace pcc

PCC_API int pcc_ball_query_f32(const float *q, const float *p, int B, int P1, int P2, int K, float radius,
                               int64_t *out_idx, float *out_d2, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(q && p && out_idx, "pcc_ball_query_f32: null pointer");
    PCC_REQUIRE(B >= 0 && P1 >= 0 && P2 >= 1 && K >= 1, "pcc_ball_query_f32: bad shape B=%d P1=%d P2=%d K=%d", B, P1,
                P2, K);
    PCC_REQUIRE(B <= 65535, "pcc_ball_query_f32: B=%d exceeds 65535", B);
    if (B == 0 || P1 == 0) return 0;
    const float r2 = radius * radius;  // fp32 product, as ball_query_cpu.cpp does
    dim3 grid((P1 + BQ_WARPS - 1) / BQ_WARPS, B);
    ball_query_kernel<<<grid, BQ_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(q, p, P1, P2, K, r2, out_idx,
                                                                                    out_d2);
    return check_launch("ball_query_kernel");
}

PCC_API int pcc_gather_f32(const float *feat, const int64_t *idx, int B, int N, int C, int64_t M, float *out,
                           void *stream) {
    using namespace pcc;
    PCC_REQUIRE(feat && idx && out, "pcc_gather_f32: null pointer");
    PCC_REQUIRE(B >= 0 && N >= 1 && C >= 1 && M >= 0, "pcc_gather_f32: bad shape");
    const long long rows = static_cast<long long>(B) * M;
    if (rows == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec = (C % 4 == 0) && (reinterpret_cast<uintptr_t>(feat) % 16 == 0) &&
                     (reinterpret_cast<uintptr_t>(out) % 16 == 0);
    if (vec) {
        gather_rows_kernel<float4><<<grid_for(rows * (C / 4)), 256, 0, st>>>(
            reinterpret_cast<const float4 *>(feat), idx, rows, N, C / 4, M, reinterpret_cast<float4 *>(out));
    } else {
        gather_rows_kernel<float><<<grid_for(rows * C), 256, 0, st>>>(feat, idx, rows, N, C, M, out);
    }
    return check_launch("gather_rows_kernel");
}

PCC_API int pcc_gather_bwd_f32(const float *grad_out, const int64_t *idx, int B, int N, int C, int64_t M,
                               float *grad_feat, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(grad_out && idx && grad_feat, "pcc_gather_bwd_f32: null pointer");
    PCC_REQUIRE(B >= 0 && N >= 1 && C >= 1 && M >= 0, "pcc_gather_bwd_f32: bad shape");
    const long long rows = static_cast<long long>(B) * M;
    if (rows == 0) return 0;
    gather_bwd_kernel<<<grid_for(rows * C), 256, 0, static_cast<cudaStream_t>(stream)>>>(grad_out, idx, rows, N, C, M,
                                                                                         grad_feat);
    return check_launch("gather_bwd_kernel");
}

PCC_API int pcc_gather_concat_bf16(const float *feat, int C, const float *xyz, const int64_t *idx, int B, int N, int64_t M, int kpad,
                                   void *out, const float *centre, int nsample, void *stream) {
    using namespace pcc;
    PCC_REQUIRE((feat || xyz) && idx && out, "pcc_gather_concat_bf16: null pointer");
    PCC_REQUIRE(B >= 0 && N >= 1 && M >= 0 && C >= 0 && (feat || C == 0), "pcc_gather_concat_bf16: bad shape");
    PCC_REQUIRE(kpad % 8 == 0 && kpad >= C + (xyz ? 3 : 0), "pcc_gather_concat_bf16: kpad=%d must be a multiple of 8 and >= %d", kpad,
                C + (xyz ? 3 : 0));
    PCC_REQUIRE(reinterpret_cast<uintptr_t>(out) % 16 == 0 && (!feat || reinterpret_cast<uintptr_t>(feat) % 16 == 0),
                "pcc_gather_concat_bf16: feat / out must be 16-byte aligned");
    PCC_REQUIRE(!centre || (xyz && nsample >= 1 && M % nsample == 0), "pcc_gather_concat_bf16: centre needs xyz and nsample | M");
    const long long rows = static_cast<long long>(B) * M;
    if (rows == 0) return 0;
    PCC_REQUIRE(rows < (1ll << 31), "pcc_gather_concat_bf16: %lld rows exceed 2^31", rows);
    gather_concat_bf16_kernel<<<grid_for(rows * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        feat, C, xyz, idx, rows, N, M, kpad, static_cast<uint4 *>(out), centre, centre ? nsample : 1);
    return check_launch("gather_concat_bf16_kernel");
}
