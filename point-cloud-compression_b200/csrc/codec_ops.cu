// Small fused element-wise kernels of the batched codec driver (HBM-bound glue between the big kernels).
//
//   normalize_kernel  pn_kit.normalize   (/root/reference/pn_kit.py:47-60), one CTA per cloud: bbox reduction then
//                     out = (p - center) * (1 - margin) / longest + 0.5, same fp32 op order as the reference.
//   assemble_kernel   decompress.py:104-116: patches / scale + patch centre, then pn_kit.denormalize (pn_kit.py:62-66).
#include <cuda_bf16.h>

#include "pcc_common.cuh"

namespace pcc {

__global__ void __launch_bounds__(1024)
normalize_kernel(const float *__restrict__ xyz, int N, float scale_mul, float *__restrict__ out, float *__restrict__ center,
                 float *__restrict__ longest, float *__restrict__ bbox) {
    __shared__ float red[6][32];
    __shared__ float s_c[3], s_l;
    const int b = blockIdx.x, tid = threadIdx.x;
    const float *pc = xyz + static_cast<size_t>(b) * N * 3;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = tid; i < N; i += 1024) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = pc[static_cast<size_t>(i) * 3 + c];
            mn[c] = fminf(mn[c], v);
            mx[c] = fmaxf(mx[c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(FULL_MASK, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(FULL_MASK, mx[c], o));
        }
        if ((tid & 31) == 0) {
            red[c][tid >> 5] = mn[c];
            red[3 + c][tid >> 5] = mx[c];
        }
    }
    __syncthreads();
    if (tid < 32) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float a = red[c][tid], z = red[3 + c][tid];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a = fminf(a, __shfl_xor_sync(FULL_MASK, a, o));
                z = fmaxf(z, __shfl_xor_sync(FULL_MASK, z, o));
            }
            mn[c] = a;
            mx[c] = z;
        }
        if (tid == 0) {
            float l = -INFINITY;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                s_c[c] = __fdiv_rn(__fadd_rn(mx[c], mn[c]), 2.0f);  // (max + min) / 2      pn_kit.py:53
                l = fmaxf(l, __fsub_rn(mx[c], mn[c]));             // max extent           pn_kit.py:54
                if (center) center[b * 3 + c] = s_c[c];
                if (bbox) {
                    bbox[b * 6 + c] = mn[c];
                    bbox[b * 6 + 3 + c] = mx[c];
                }
            }
            s_l = l;
            if (longest) longest[b] = l;
        }
    }
    __syncthreads();
    const float l = s_l;
    float *o = out + static_cast<size_t>(b) * N * 3;
    for (int e = tid; e < N * 3; e += 1024) {
        const int c = e % 3;
        float v = __fsub_rn(pc[e], s_c[c]);                      // pc - center            pn_kit.py:56
        v = __fdiv_rn(__fmul_rn(v, scale_mul), l);              // * (1 - margin) / longest   :57
        o[e] = __fadd_rn(v, 0.5f);                              // + 0.5                  :58
    }
}

// rec[b, s*k + j, :] = ((patch[b,s,j,:] * inv_scale + centres[b,s,:]) - 0.5) * longest[b] / (1 - margin) + center[b,:]
__global__ void __launch_bounds__(256)
assemble_kernel(const float *__restrict__ patches, const float *__restrict__ centres, const float *__restrict__ center,
                const float *__restrict__ longest, int S, int k, float inv_scale, float one_minus_margin, long long total,
                float *__restrict__ out) {
    for (long long e = blockIdx.x * 256ll + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256ll) {
        const int c = static_cast<int>(e % 3);
        const long long pt = e / 3;
        const long long bs = pt / k;        // patch index b*S + s
        const long long b = bs / S;
        float v = __fadd_rn(__fmul_rn(patches[e], inv_scale), centres[bs * 3 + c]);
        if (center) {
            v = __fsub_rn(v, 0.5f);
            v = __fdiv_rn(__fmul_rn(v, longest[b]), one_minus_margin);
            v = __fadd_rn(v, center[b * 3 + c]);
        }
        out[e] = v;
    }
}

// eval.py's per-cloud metrics from the Chamfer by-products, one CTA per cloud (all arithmetic in double, fixed order):
//   chamfer = per_cloud / (max - min)^2                  eval.py:199-205 (the scale + shift normalisation divides d2)
//   d1 mse  = mean_i dx[i]                               eval.py:84 (recon -> original)
//   d1 psnr = 10 log10(|bbox diagonal|^2 / mse)          eval.py:88-92
__global__ void __launch_bounds__(256)
eval_metrics_kernel(const float *__restrict__ dx, const float *__restrict__ per_cloud, const float *__restrict__ bbox, int P1,
                    double *__restrict__ out) {
    __shared__ double red[8];
    const int b = blockIdx.x, tid = threadIdx.x;
    double s = 0.0;
    for (int i = tid; i < P1; i += 256) s += static_cast<double>(dx[static_cast<size_t>(b) * P1 + i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL_MASK, s, o);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        const double mse = t / static_cast<double>(P1);
        const float *bb = bbox + static_cast<size_t>(b) * 6;
        const double lo = fmin(fmin(static_cast<double>(bb[0]), static_cast<double>(bb[1])), static_cast<double>(bb[2]));
        const double hi = fmax(fmax(static_cast<double>(bb[3]), static_cast<double>(bb[4])), static_cast<double>(bb[5]));
        const double scale = hi - lo;
        double diag2 = 0.0;
        for (int a = 0; a < 3; ++a) {
            const double e = static_cast<double>(bb[3 + a]) - static_cast<double>(bb[a]);
            diag2 += e * e;
        }
        out[static_cast<size_t>(b) * 3 + 0] = static_cast<double>(per_cloud[b]) / (scale * scale);
        out[static_cast<size_t>(b) * 3 + 1] = 10.0 * log10(diag2 / mse);
        out[static_cast<size_t>(b) * 3 + 2] = mse;
    }
}

// AE.py:42-45 / compress.py:125-127 in one pass: latent = sigmoid(raw) * spread - spread / 2 (same fp32 operation order as the
// reference), latent_q = round(latent) (half to even, torch.round), and the rounded latent again as bf16 rows zero padded to
// kpad columns -- the A operand of inv_pool's first GEMM (AE.py:48).
__global__ void __launch_bounds__(256)
quantise_latent_kernel(const float *__restrict__ raw, long long rows, int d, int kpad, float spread, float *__restrict__ latent,
                       float *__restrict__ latent_q, __nv_bfloat16 *__restrict__ q_bf16) {
    const long long total = rows * kpad;
    const float half = spread / 2;
    for (long long e = blockIdx.x * 256ll + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256ll) {
        const long long r = e / kpad;
        const int c = static_cast<int>(e - r * kpad);
        float q = 0.0f;
        if (c < d) {
            const float x = raw[r * d + c];
            const float sg = 1.0f / (1.0f + expf(-x));                    // torch.sigmoid (fp32)
            const float l = __fsub_rn(__fmul_rn(sg, spread), half);
            q = rintf(l);
            latent[r * d + c] = l;
            latent_q[r * d + c] = q;
        }
        if (q_bf16) q_bf16[e] = __float2bfloat16_rn(q);
    }
}

}  // namespace pcc

PCC_API int pcc_normalize_f32(const float *xyz, int B, int N, float margin, float *out, float *center, float *longest,
                              float *bbox, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(xyz && out && B >= 0 && N >= 1, "pcc_normalize_f32: bad argument");
    if (B == 0) return 0;
    const float one_minus = static_cast<float>(1.0 - static_cast<double>(margin));  // python computes 1 - margin in double
    normalize_kernel<<<B, 1024, 0, static_cast<cudaStream_t>(stream)>>>(xyz, N, one_minus, out, center, longest, bbox);
    return check_launch("normalize_kernel");
}

PCC_API int pcc_assemble_f32(const float *patches, const float *centres, const float *center, const float *longest, int B,
                             int S, int k, float patch_scale, float margin, float *out, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(patches && centres && out && B >= 0 && S >= 1 && k >= 1, "pcc_assemble_f32: bad argument");
    PCC_REQUIRE((center == nullptr) == (longest == nullptr), "pcc_assemble_f32: center and longest go together");
    const long long total = static_cast<long long>(B) * S * k * 3;
    if (total == 0) return 0;
    long long blocks = (total + 255) / 256;
    const long long cap = static_cast<long long>(num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    const float one_minus = static_cast<float>(1.0 - static_cast<double>(margin));
    assemble_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        patches, centres, center, longest, S, k, 1.0f / patch_scale, one_minus, total, out);
    return check_launch("assemble_kernel");
}

PCC_API int pcc_eval_metrics_f32(const float *dx, const float *per_cloud, const float *bbox, int B, int P1, double *out,
                                 void *stream) {
    using namespace pcc;
    PCC_REQUIRE(dx && per_cloud && bbox && out, "pcc_eval_metrics_f32: null pointer");
    PCC_REQUIRE(B >= 1 && P1 >= 1, "pcc_eval_metrics_f32: bad shape B=%d P1=%d", B, P1);
    eval_metrics_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(dx, per_cloud, bbox, P1, out);
    return check_launch("eval_metrics_kernel");
}

PCC_API int pcc_quantise_latent_f32(const float *raw, int64_t rows, int d, int kpad, float spread, float *out_latent, float *out_q,
                                    void *out_q_bf16, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(raw && out_latent && out_q, "pcc_quantise_latent_f32: null pointer");
    PCC_REQUIRE(rows >= 0 && d >= 1 && kpad >= d, "pcc_quantise_latent_f32: bad shape rows=%lld d=%d kpad=%d", static_cast<long long>(rows), d, kpad);
    if (rows == 0) return 0;
    const long long total = rows * kpad;
    long long blocks = (total + 255) / 256;
    const long long cap = static_cast<long long>(num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    quantise_latent_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        raw, rows, d, kpad, spread, out_latent, out_q, static_cast<__nv_bfloat16 *>(out_q_bf16));
    return check_launch("quantise_latent_kernel");
}
