// pn_kit.PointNet of the IPDAE encoder in ONE kernel (/root/reference/pn_kit.py:124-144 as AE.py:17,39 calls it):
//     y[patch, d] = max over the patch's 256 positions of  W3 . relu(W2 . relu(W1 . relu(W0 . [feat | xyz] + b0) + b1) + b2) + b3
// feat [M, 128] bf16 (the SetAbstraction kernel's output), xyz [M, 3] fp32; W0 [128, 131], W1 [256, 128], W2 [512, 256], W3 [d, 512].
//
// Round 1 ran the 131 -> 128 -> 256 front (ws::pnf_chain_kernel, HBM bound) and the 256 -> 512 -> d tail (pnt::pn_tail_kernel) as
// two launches with a [M, 256] bf16 activation (268 MB per 32-cloud step) written to HBM and read straight back.  Here a CTA
// keeps a tile of 128 positions on chip through all four layers:
//   * every weight matrix is STREAMED from L2 through one 4-stage TMA ring of [128 ch x 64 k] slabs (W0: 2 slabs, W1: 4, W2: 16
//     per tile); only the [128 x 16] xyz + bias block of W0, the biases and the packed W3 are resident;
//   * the feature tile (2 slabs, TMA) and X1 = relu(layer 0) (2 slabs) alias the two X3 buffers of the tail, which are idle until
//     the first 512-wide chunk's epilogue; X2 = relu(layer 1) (4 slabs) is where the tail kernel's TMA-loaded input used to be;
//   * accumulators ping-pong between two 128-column TMEM regions through the seven contractions of a tile (A, B0, B1, T0..T3),
//     the 512 -> d layer accumulates in a third region across the four chunks, exactly as in pn_tail.cu.
// Ordering between the aliased regions needs no extra barriers: MMAs complete in issue order, so a tcgen05.commit that signals
// "accumulator full" also says every earlier MMA has finished reading its operands.  The one exception is the feature tile of
// the NEXT tile, loaded by TMA: the producer waits for `f_empty`, committed after the last MMA that reads X3 buffer 0.
//   warps 0-7 epilogue (warp w: TMEM lanes 32 (w % 4).., columns 64 (w / 4)..), warp 8 MMA issue, warp 9 TMA producer
// HBM traffic: 134 MB of features in, d floats per patch out.
#include <cuda.h>
#include <cuda_bf16.h>

#include "chain_ws.h"
#include "pcc_common.cuh"
#include "tc_ptx.cuh"

namespace pcc {
namespace pnf2 {

constexpr int P = 128, SLAB = P * 128, K16 = 4096;
constexpr int C_MID = 512, C_OUT_MAX = 16, KP3 = 528;
constexpr int NST = 4;
constexpr int OFF_X2 = 0;                               // 4 slabs
constexpr int OFF_W = OFF_X2 + 4 * SLAB;                // 65536: ring
constexpr int OFF_X3 = OFF_W + NST * SLAB;              // 131072: 2 buffers x 2 slabs
constexpr int OFF_F = OFF_X3;                           //   feature tile   (aliases X3 buffer 0)
constexpr int OFF_X1 = OFF_X3 + 2 * SLAB;               //   relu(layer 0)  (aliases X3 buffer 1)
constexpr int OFF_W3 = OFF_X3 + 4 * SLAB;               // 196608: packed [16 x 528]
constexpr int OFF_ONES = OFF_W3 + 17 * 1024;            // 214016
constexpr int OFF_XIN = OFF_ONES + K16;                 // 218112: (x y z 1 0...) per position, canonical K16 block
constexpr int OFF_W0X = OFF_XIN + K16;                  // 222208: W0[:, 128..143] (xyz weights + bias), canonical, SBO 256
constexpr int OFF_B1 = OFF_W0X + K16;                   // 226304: 256 floats
constexpr int OFF_B2 = OFF_B1 + 1024;                   // 227328: 512 floats
constexpr int OFF_RED = OFF_B2 + 2048;                  // 229376
constexpr int OFF_BAR = OFF_RED + 512;                  // 229888
constexpr int SMEM = OFF_BAR + 384 + 1024;              // 231296
constexpr int THREADS = 320;
constexpr int TMEM_COLS = 512;
static_assert(16 * KP3 * 2 <= 17 * 1024 && SMEM <= 227 * 1024, "layout");

struct Params {
    const float *xyz;
    long long ld_xyz;
    const void *w0p;        // packed [128 x 144] (pcc_mlp_pack_weights_f32 of the rotated [feat | xyz] layer: bias in column 131)
    const float *b1;        // [256]
    const float *b2;        // [512]
    const void *w3p;        // packed [16(128) x 528]
    float *out;             // [n_patches, c_out]
    int n_patches;
    int c_out;
    int relu3;
};

__device__ __forceinline__ uint32_t bf16_bits(float v) {
    return static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(v)));
}

// 32 accumulator columns (channels c0 .. c0+31 of position `row`) -> (+ bias) ReLU -> bf16 -> swizzled slabs starting at base
__device__ __forceinline__ void store_chunk32(uint32_t base, int row, int c0, const uint32_t (&v)[32], const float *bias) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + 8 * i;
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float lo = __uint_as_float(v[8 * i + 2 * e]), hi = __uint_as_float(v[8 * i + 2 * e + 1]);
            if (bias) {
                lo += bias[c + 2 * e];
                hi += bias[c + 2 * e + 1];
            }
            pk[e] = pack_relu_bf16x2(lo, hi);
        }
        st_shared_v4(base + (c >> 6) * SLAB + row * 128 + ((((c >> 3) & 7) ^ (row & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
    }
}

__global__ void __launch_bounds__(THREADS, 1)
pn_fused_kernel(const __grid_constant__ Params prm, const __grid_constant__ CUtensorMap tm_f, const __grid_constant__ CUtensorMap tm_w0,
                const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t sb = smem_u32(smem);
    const int tid = threadIdx.x, warp = __shfl_sync(FULL_MASK, tid >> 5, 0), lane = tid & 31;
    const uint32_t bar = sb + OFF_BAR;
    const uint32_t w_full = bar, w_empty = bar + 32;                                                          // [4] each
    const uint32_t acc_full = bar + 64, acc_empty = bar + 80, x3_full = bar + 96, x3_empty = bar + 112;        // [2] each
    const uint32_t acc3_full = bar + 128, acc3_empty = bar + 136, f_full = bar + 144, f_empty = bar + 152;
    const uint32_t xin_ready = bar + 160, x1_full = bar + 168, x2r = bar + 176;                                // x2r [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 200);

    // ---- prologue: resident pieces ----
    {
        const int4 *src = static_cast<const int4 *>(prm.w3p);
        int4 *dst = reinterpret_cast<int4 *>(smem + OFF_W3);
        for (int i = tid; i < 16 * KP3 * 2 / 16; i += THREADS) dst[i] = __ldg(src + i);
        for (int i = tid; i < K16 / 16; i += THREADS)
            reinterpret_cast<uint4 *>(smem + OFF_ONES)[i] = make_uint4(i < 128 ? 0x3f80u : 0u, 0u, 0u, 0u);
        // K step 8 (k = 128..143) of the packed W0: 256 bytes per 8-channel group, groups 144 * 16 bytes apart -> compact, SBO 256
        const unsigned char *w0b = static_cast<const unsigned char *>(prm.w0p);
        for (int i = tid; i < 16 * 16; i += THREADS)
            reinterpret_cast<int4 *>(smem + OFF_W0X)[i] = __ldg(reinterpret_cast<const int4 *>(w0b + (i >> 4) * (144 * 16) + 2048) + (i & 15));
        for (int i = tid; i < 256; i += THREADS) reinterpret_cast<float *>(smem + OFF_B1)[i] = __ldg(prm.b1 + i);
        for (int i = tid; i < C_MID; i += THREADS) reinterpret_cast<float *>(smem + OFF_B2)[i] = __ldg(prm.b2 + i);
    }
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) {
            mbar_init(w_full + 8 * i, 1);
            mbar_init(w_empty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(acc_full + 8 * i, 1);
            mbar_init(acc_empty + 8 * i, 8);
            mbar_init(x3_full + 8 * i, 8);
            mbar_init(x3_empty + 8 * i, 1);
            mbar_init(x2r + 8 * i, 8);
        }
        mbar_init(acc3_full, 1);
        mbar_init(acc3_empty, 8);
        mbar_init(f_full, 1);
        mbar_init(f_empty, 1);
        mbar_init(xin_ready, 8);
        mbar_init(x1_full, 8);
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 9) {
        // ---- TMA producer: the feature tile, then the 22 weight slabs of a tile through the ring ----
        if (lane == 0) {
            uint32_t it = 0, fph = 1;   // a fresh barrier passes a parity-1 wait: the feature region starts out free
            auto ring = [&](const CUtensorMap *tm, int c0, int c1) {
                const uint32_t st = it % NST, ph = ((it / NST) & 1u) ^ 1u;
                mbar_wait(w_empty + 8 * st, ph);
                mbar_arrive_expect_tx(w_full + 8 * st, SLAB);
                tma_load_2d(sb + OFF_W + st * SLAB, tm, c0, c1, w_full + 8 * st);
                ++it;
            };
            for (int patch = blockIdx.x; patch < prm.n_patches; patch += gridDim.x) {
                for (int t2 = 0; t2 < 2; ++t2) {
                    const int tile = 2 * patch + t2;
                    mbar_wait(f_empty, fph);
                    fph ^= 1u;
                    mbar_arrive_expect_tx(f_full, 2 * SLAB);
                    tma_load_2d(sb + OFF_F, &tm_f, 0, tile * P, f_full);
                    tma_load_2d(sb + OFF_F + SLAB, &tm_f, 64, tile * P, f_full);
                    for (int kb = 0; kb < 2; ++kb) ring(&tm_w0, kb * 64, 0);
                    for (int c = 0; c < 2; ++c)
                        for (int kb = 0; kb < 2; ++kb) ring(&tm_w1, kb * 64, c * 128);
                    for (int c = 0; c < 4; ++c)
                        for (int kb = 0; kb < 4; ++kb) ring(&tm_w2, kb * 64, c * 128);
                }
            }
        }
    } else if (warp == 8) {
        // ---- MMA issuer (warp-uniform control flow, one elected lane issues) ----
        const uint32_t tb = __shfl_sync(FULL_MASK, tmem_base, 0);
        const uint32_t id128 = umma_idesc(128, 128), id16 = umma_idesc(128, 16);
        const uint64_t d_f = umma_desc_sw128(sb + OFF_F), d_x1 = umma_desc_sw128(sb + OFF_X1), d_x2 = umma_desc_sw128(sb + OFF_X2);
        const uint64_t d_w = umma_desc_sw128(sb + OFF_W), d_x3 = umma_desc_sw128(sb + OFF_X3);
        const uint64_t d_w3 = umma_desc(sb + OFF_W3, 128, KP3 * 16), d_ones = umma_desc(sb + OFF_ONES, 2048, 128);
        const uint64_t d_xin = umma_desc(sb + OFF_XIN, 2048, 128), d_w0x = umma_desc(sb + OFF_W0X, 128, 256);
        uint32_t it = 0, ph_acc_empty[2] = {1, 1}, ph_x3_full[2] = {0, 0}, ph_acc3_empty = 1;
        uint32_t ph_f = 0, ph_xin = 0, ph_x1 = 0, ph_x2r = 0;
        // one contraction of K = 64 * n_kb: acc[buf] = A slabs (a_desc + kb slabs) . ring slabs^T (+ the xyz / bias K block of
        // layer 0); the "accumulator full" commit is issued by the same elected lane, in the same block, as the last MMAs
        auto contract = [&](int buf, uint64_t a_desc, int n_kb, bool with_xin) {
            for (int kb = 0; kb < n_kb; ++kb, ++it) {
                const uint32_t st = it % NST, ph = (it / NST) & 1u;
                mbar_wait(w_full + 8 * st, ph);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_bf16(tb + buf * 128, a_desc + kb * (SLAB >> 4) + ks * 2, d_w + st * (SLAB >> 4) + ks * 2, id128, (kb | ks) > 0);
                    umma_commit(w_empty + 8 * st);
                    if (kb == n_kb - 1) {
                        if (with_xin) umma_bf16(tb + buf * 128, d_xin, d_w0x, id128, 1u);
                        umma_commit(acc_full + 8 * buf);
                    }
                }
                __syncwarp();
            }
        };
        auto acquire = [&](int buf) {
            mbar_wait(acc_empty + 8 * buf, ph_acc_empty[buf]);
            ph_acc_empty[buf] ^= 1u;
            tc_fence_after();
        };
        for (int patch = blockIdx.x; patch < prm.n_patches; patch += gridDim.x) {
            for (int t2 = 0; t2 < 2; ++t2) {
                // ---- A: acc[0] = [feat | xyz 1] . W0^T ----
                acquire(0);
                mbar_wait(f_full, ph_f);
                ph_f ^= 1u;
                mbar_wait(xin_ready, ph_xin);
                ph_xin ^= 1u;
                tc_fence_after();
                contract(0, d_f, 2, true);
                // ---- B0 / B1: acc[1], acc[0] = X1 . W1[c * 128 ..]^T (bias added by the epilogue) ----
                acquire(1);
                mbar_wait(x1_full, ph_x1);
                ph_x1 ^= 1u;
                tc_fence_after();
                contract(1, d_x1, 2, false);
                acquire(0);
                contract(0, d_x1, 2, false);
                // ---- T0..T3: 512-wide layer in chunks of 128 channels, each consumed as a K = 128 slice of the 512 -> d layer ----
#pragma unroll
                for (int c = 0; c <= 4; ++c) {
                    if (c < 4) {
                        const int buf = (c + 1) & 1;
                        acquire(buf);
                        if (c == 0) {
                            mbar_wait(x2r, ph_x2r);
                            mbar_wait(x2r + 8, ph_x2r);
                            ph_x2r ^= 1u;
                            tc_fence_after();
                        }
                        contract(buf, d_x2, 4, false);
                    }
                    if (c > 0) {
                        const int pc = c - 1, pb = pc & 1;
                        mbar_wait(x3_full + 8 * pb, ph_x3_full[pb]);
                        ph_x3_full[pb] ^= 1u;
                        if (pc == 0) {
                            mbar_wait(acc3_empty, ph_acc3_empty);
                            ph_acc3_empty ^= 1u;
                        }
                        tc_fence_after();
                        if (elect_one()) {
#pragma unroll
                            for (int ks = 0; ks < 8; ++ks)
                                umma_bf16(tb + 256, d_x3 + (pb * 2 + (ks >> 2)) * (SLAB >> 4) + (ks & 3) * 2, d_w3 + 16 * (pc * 8 + ks), id16,
                                          (pc | ks) > 0);
                            if (pc == 3) umma_bf16(tb + 256, d_ones, d_w3 + 16 * 32, id16, 1u);   // bias column (k = 512)
                            umma_commit(x3_empty + 8 * pb);
                            if (pc == 2) umma_commit(f_empty);       // X3 buffer 0 has had its last reader: the next feature tile may land
                            if (pc == 3) umma_commit(acc3_full);
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else {
        // ---- epilogue warps ----
        const int q = warp & 3, h = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const float *b1s = reinterpret_cast<const float *>(smem + OFF_B1);
        const float *b2s = reinterpret_cast<const float *>(smem + OFF_B2);
        float *red = reinterpret_cast<float *>(smem + OFF_RED);
        uint32_t ph_acc_full[2] = {0, 0}, ph_x3_empty[2] = {1, 1}, ph_acc3_full = 0;
        float px = 0.f, py = 0.f, pz = 0.f;
        auto load_xyz = [&](long long tile) {
            const float *src = prm.xyz + (tile * P + tid) * prm.ld_xyz;
            px = __ldg(src);
            py = __ldg(src + 1);
            pz = __ldg(src + 2);
        };
        if (tid < P && blockIdx.x < prm.n_patches) load_xyz(2ll * blockIdx.x);
        // acc[buf] columns h*64 .. h*64+63 of my row -> (+ bias) ReLU -> bf16 slabs at `base`, then the two signals
        auto drain = [&](int buf, uint32_t base, const float *bias) {
            mbar_wait(acc_full + 8 * buf, ph_acc_full[buf]);
            ph_acc_full[buf] ^= 1u;
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t v[32];
                tmem_ld32(lane_base + buf * 128 + h * 64 + j * 32, v);
                store_chunk32(base, row, h * 64 + j * 32, v, bias);
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
        };
        for (int patch = blockIdx.x; patch < prm.n_patches; patch += gridDim.x) {
            for (int t2 = 0; t2 < 2; ++t2) {
                const long long tile = 2ll * patch + t2;
                if (tid < P) {   // K block of layer 0's last step: (x y z 1 0 ...), the 1 multiplies the bias column (k = 131)
                    const uint32_t w0 = bf16_bits(px) | (bf16_bits(py) << 16), w1 = bf16_bits(pz) | (0x3f80u << 16);
                    st_shared_v4(sb + OFF_XIN + tid * 16, w0, w1, 0u, 0u);
                    st_shared_v4(sb + OFF_XIN + 2048 + tid * 16, 0u, 0u, 0u, 0u);
                    const long long nt = t2 == 0 ? tile + 1 : 2ll * (patch + gridDim.x);
                    if (nt < 2ll * prm.n_patches) load_xyz(nt);
                }
                fence_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive1(xin_ready);
                // ---- A -> X1 ----
                drain(0, sb + OFF_X1, nullptr);
                if (lane == 0) {
                    mbar_arrive1(x1_full);
                    mbar_arrive1(acc_empty);
                }
                // ---- B0, B1 -> X2 slabs 0-1, 2-3 ----
                drain(1, sb + OFF_X2, b1s);
                if (lane == 0) {
                    mbar_arrive1(x2r);
                    mbar_arrive1(acc_empty + 8);
                }
                drain(0, sb + OFF_X2 + 2 * SLAB, b1s + 128);
                if (lane == 0) {
                    mbar_arrive1(x2r + 8);
                    mbar_arrive1(acc_empty);
                }
                // ---- T0..T3 -> X3 buffers ----
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int buf = (c + 1) & 1, pb = c & 1;
                    mbar_wait(x3_empty + 8 * pb, ph_x3_empty[pb]);
                    ph_x3_empty[pb] ^= 1u;
                    drain(buf, sb + OFF_X3 + pb * 2 * SLAB, b2s + c * 128);
                    if (lane == 0) {
                        mbar_arrive1(x3_full + 8 * pb);
                        mbar_arrive1(acc_empty + 8 * buf);
                    }
                }
                // ---- layer 3 result of this tile: 16 columns, max over the 128 positions ----
                mbar_wait(acc3_full, ph_acc3_full);
                ph_acc3_full ^= 1u;
                tc_fence_after();
                if (h == 0) {
                    uint32_t v[16];
                    tmem_ld16(lane_base + 256, v);
                    float mine = 0.0f;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float m;
                        asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(__uint_as_float(v[i])));
                        if (lane == i) mine = m;
                    }
                    if (lane < 16) red[q * 16 + lane] = mine;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive1(acc3_empty);
                named_bar_sync(1, 256);
                if (tid < 16) {
                    float m = fmaxf(fmaxf(red[tid], red[16 + tid]), fmaxf(red[32 + tid], red[48 + tid]));
                    if (t2 == 0) {
                        red[64 + tid] = m;
                    } else {
                        m = fmaxf(m, red[64 + tid]);
                        if (prm.relu3) m = fmaxf(m, 0.0f);
                        if (tid < prm.c_out) prm.out[static_cast<long long>(patch) * prm.c_out + tid] = m;
                    }
                }
                named_bar_sync(1, 256);   // red[] is rewritten by the next tile
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

}  // namespace pnf2
}  // namespace pcc

/*
 * feat [rows, 128] bf16 (row pitch ld_feat), xyz [rows, 3] fp32 (row pitch ld_xyz), rows % 256 == 0.
 * w0f [128, 128] bf16 = the feature columns of the rotated first layer [feat | xyz]; w0_packed = pcc_mlp_pack_weights_f32 of the
 * whole rotated layer (cin = 131: its K step 8 holds the xyz columns and the bias); w1 [256, 128] bf16, b1 [256];
 * w2 [512, 256] bf16, b2 [512]; w3_packed: pcc_mlp_pack_weights_f32(cin = 512, cout <= 16).  out [rows / 256, cout] fp32.
 */
PCC_API int pcc_pointnet_fused_bf16(const void *feat, int64_t rows, int64_t ld_feat, const float *xyz, int64_t ld_xyz, const void *w0f_bf16,
                                    const void *w0_packed, const void *w1_bf16, const float *b1, const void *w2_bf16, const float *b2,
                                    const void *w3_packed, int cout, int relu3, float *out, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(feat && xyz && w0f_bf16 && w0_packed && w1_bf16 && b1 && w2_bf16 && b2 && w3_packed && out, "pcc_pointnet_fused_bf16: null pointer");
    PCC_REQUIRE(rows >= 0 && rows % 256 == 0 && rows / 256 < (1ll << 30), "pcc_pointnet_fused_bf16: rows=%lld must be a multiple of 256",
                static_cast<long long>(rows));
    PCC_REQUIRE(cout >= 1 && cout <= pnf2::C_OUT_MAX, "pcc_pointnet_fused_bf16: cout=%d outside [1,16]", cout);
    PCC_REQUIRE(ld_feat >= 128 && ld_feat % 8 == 0 && ld_xyz >= 3 && reinterpret_cast<uintptr_t>(feat) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(w0f_bf16) % 16 == 0 && reinterpret_cast<uintptr_t>(w1_bf16) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(w2_bf16) % 16 == 0 && reinterpret_cast<uintptr_t>(w0_packed) % 16 == 0,
                "pcc_pointnet_fused_bf16: operands must be 16-byte aligned, the feature pitch a multiple of 8 elements");
    if (rows == 0) return 0;
    CUtensorMap tm_f, tm_w0, tm_w1, tm_w2;
    if (int r = make_tmap_bf16_2d(&tm_f, feat, static_cast<uint64_t>(rows), 128, static_cast<uint64_t>(ld_feat), 128)) return r;
    if (int r = make_tmap_bf16_2d(&tm_w0, w0f_bf16, 128, 128, 128, 128)) return r;
    if (int r = make_tmap_bf16_2d(&tm_w1, w1_bf16, 256, 128, 128, 128)) return r;
    if (int r = make_tmap_bf16_2d(&tm_w2, w2_bf16, 512, 256, 256, 128)) return r;
    pnf2::Params p{};
    p.xyz = xyz;
    p.ld_xyz = ld_xyz;
    p.w0p = w0_packed;
    p.b1 = b1;
    p.b2 = b2;
    p.w3p = w3_packed;
    p.out = out;
    p.n_patches = static_cast<int>(rows / 256);
    p.c_out = cout;
    p.relu3 = relu3;
    const cudaError_t e = cudaFuncSetAttribute(pnf2::pn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pnf2::SMEM);
    if (e != cudaSuccess) {
        set_error("pcc_pointnet_fused_bf16: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    const int grid = p.n_patches < num_sms() ? p.n_patches : num_sms();
    pnf2::pn_fused_kernel<<<grid, pnf2::THREADS, pnf2::SMEM, static_cast<cudaStream_t>(stream)>>>(p, tm_f, tm_w0, tm_w1, tm_w2);
    return check_launch("pn_fused_kernel");
}
