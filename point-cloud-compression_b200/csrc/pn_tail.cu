// PointNet tail of the IPDAE encoder, fused:  y[patch, 16] = max over the patch's 256 positions of
//     W3 . relu(W2 . x + b2) + b3            x [M, 256] bf16 (output of the PNF chain), W2 [512, 256], W3 [16, 512]
// (pn_kit.PointNet layers 2-3 + the max of /root/reference/pn_kit.py:136-143; AE.py:17,39).
//
// The 512-wide activation never exists in HBM: the CTA keeps a tile of 128 positions (4 swizzled K slabs, TMA) in shared
// memory and streams W2 from L2 through a 4-stage TMA ring in chunks of 128 output channels.  Each chunk's accumulator
// (TMEM, double buffered) is turned into bf16 by the epilogue warps (+ bias, ReLU) and immediately consumed as one K = 128
// slice of the 512 -> 16 layer, which accumulates in a third TMEM region across the four chunks.  The final 16 columns
// are reduced over the tile's 128 positions with redux.sync.max.f32 and over the patch's two tiles in shared memory.
//   warps 0-7  epilogue (warp w: TMEM lanes 32*(w%4).., columns 64*(w/4)..), warp 8 MMA issue, warp 9 TMA producer
// Bound: L2 -> SM traffic (256 KB of W2 + 64 KB of x per 48 MFLOP tile, ~80 B/clk/SM at the tensor rate).
#include <cuda.h>
#include <cuda_bf16.h>

#include "chain_ws.h"
#include "pcc_common.cuh"
#include "tc_ptx.cuh"

namespace pcc {
namespace pnt {

constexpr int P = 128, SLAB = P * 128, K16 = 4096;
constexpr int C_IN = 256, C_MID = 512, C_OUT_MAX = 16, KP3 = 528;
constexpr int NST = 4;                                  // W2 ring stages (one [128 ch x 64 k] slab each)
constexpr int OFF_X2 = 0;                               // 4 slabs
constexpr int OFF_W = OFF_X2 + 4 * SLAB;                // 65536
constexpr int OFF_X3 = OFF_W + NST * SLAB;              // 131072: 2 buffers x 2 slabs
constexpr int OFF_W3 = OFF_X3 + 4 * SLAB;               // 196608: packed [16 x 528]
constexpr int OFF_ONES = OFF_W3 + 17 * 1024;            // 214016
constexpr int OFF_B2 = OFF_ONES + K16;                  // 218112: 512 floats
constexpr int OFF_RED = OFF_B2 + 2048;                  // 220160: [4 quadrants][16] + running [16] floats
constexpr int OFF_BAR = OFF_RED + 512;                  // 220672
constexpr int SMEM = OFF_BAR + 256 + 1024;
constexpr int THREADS = 320;
constexpr int TMEM_COLS = 512;                          // acc[2] x 128 + acc3 (16) -> next power of two
static_assert(16 * KP3 * 2 <= 17 * 1024 && SMEM <= 227 * 1024, "layout");

struct Params {
    const float *b2;        // [512]
    const void *w3p;        // packed [16(128) x 528] (pcc_mlp_pack_weights_f32: bias in column 512)
    float *out;             // [n_patches, c_out]
    int n_patches;          // patch = 2 tiles of 128 positions
    int c_out;              // <= 16
    int relu3;
};

__global__ void __launch_bounds__(THREADS, 1)
pn_tail_kernel(const __grid_constant__ Params prm, const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w2) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t sb = smem_u32(smem);
    const int tid = threadIdx.x, warp = __shfl_sync(FULL_MASK, tid >> 5, 0), lane = tid & 31;
    // barriers (8 bytes each)
    const uint32_t bar = sb + OFF_BAR;
    const uint32_t x2_full = bar, x2_empty = bar + 32, w_full = bar + 64, w_empty = bar + 96;            // [4] each
    const uint32_t acc_full = bar + 128, acc_empty = bar + 144, x3_full = bar + 160, x3_empty = bar + 176;  // [2] each
    const uint32_t acc3_full = bar + 192, acc3_empty = bar + 200;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 208);

    // ---- prologue ----
    {
        const int4 *src = static_cast<const int4 *>(prm.w3p);
        int4 *dst = reinterpret_cast<int4 *>(smem + OFF_W3);
        for (int i = tid; i < 16 * KP3 * 2 / 16; i += THREADS) dst[i] = __ldg(src + i);
        for (int i = tid; i < K16 / 16; i += THREADS)
            reinterpret_cast<uint4 *>(smem + OFF_ONES)[i] = make_uint4(i < 128 ? 0x3f80u : 0u, 0u, 0u, 0u);
        for (int i = tid; i < C_MID; i += THREADS) reinterpret_cast<float *>(smem + OFF_B2)[i] = __ldg(prm.b2 + i);
    }
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) {
            mbar_init(x2_full + 8 * i, 1);
            mbar_init(x2_empty + 8 * i, 1);
            mbar_init(w_full + 8 * i, 1);
            mbar_init(w_empty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(acc_full + 8 * i, 1);
            mbar_init(acc_empty + 8 * i, 8);
            mbar_init(x3_full + 8 * i, 8);
            mbar_init(x3_empty + 8 * i, 1);
        }
        mbar_init(acc3_full, 1);
        mbar_init(acc3_empty, 8);
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
    const int n_tiles = 2 * prm.n_patches;

    if (warp == 9) {
        // ---- TMA producer ----
        if (lane == 0) {
            uint32_t it = 0;              // W ring position
            uint32_t xph = 1;             // x2_empty parity: a fresh barrier passes a parity-1 wait
            for (int patch = blockIdx.x; patch < prm.n_patches; patch += gridDim.x) {
                for (int t2 = 0; t2 < 2; ++t2) {
                    const int tile = 2 * patch + t2;
                    for (int c = 0; c < 4; ++c) {
                        for (int kb = 0; kb < 4; ++kb, ++it) {
                            if (c == 0) {   // x slab kb of this tile (freed by the previous tile's last chunk)
                                mbar_wait(x2_empty + 8 * kb, xph);
                                mbar_arrive_expect_tx(x2_full + 8 * kb, SLAB);
                                tma_load_2d(sb + OFF_X2 + kb * SLAB, &tm_x, kb * 64, tile * P, x2_full + 8 * kb);
                            }
                            const uint32_t st = it % NST, ph = ((it / NST) & 1u) ^ 1u;
                            mbar_wait(w_empty + 8 * st, ph);
                            mbar_arrive_expect_tx(w_full + 8 * st, SLAB);
                            tma_load_2d(sb + OFF_W + st * SLAB, &tm_w2, kb * 64, c * 128, w_full + 8 * st);
                        }
                    }
                    xph ^= 1u;
                }
            }
        }
    } else if (warp == 8) {
        // ---- MMA issuer (warp-uniform control flow, one elected lane issues) ----
        const uint32_t tb = __shfl_sync(FULL_MASK, tmem_base, 0);
        const uint32_t id128 = umma_idesc(128, 128), id16 = umma_idesc(128, 16);
        const uint64_t d_x2 = umma_desc_sw128(sb + OFF_X2), d_w = umma_desc_sw128(sb + OFF_W), d_x3 = umma_desc_sw128(sb + OFF_X3);
        const uint64_t d_w3 = umma_desc(sb + OFF_W3, 128, KP3 * 16), d_ones = umma_desc(sb + OFF_ONES, 2048, 128);
        uint32_t it = 0, xph = 0, ph_acc_empty[2] = {1, 1}, ph_x3_full[2] = {0, 0}, ph_acc3_empty = 1;
        for (int patch = blockIdx.x; patch < prm.n_patches; patch += gridDim.x) {
            for (int t2 = 0; t2 < 2; ++t2) {
#pragma unroll
                for (int c = 0; c <= 4; ++c) {
                    if (c < 4) {
                        // ---- chunk c of layer 2: acc[c & 1] = x . W2[c*128 .. +128]^T ----
                        const int buf = c & 1;
                        mbar_wait(acc_empty + 8 * buf, ph_acc_empty[buf]);
                        ph_acc_empty[buf] ^= 1u;
                        tc_fence_after();
#pragma unroll
                        for (int kb = 0; kb < 4; ++kb, ++it) {
                            if (c == 0) mbar_wait(x2_full + 8 * kb, xph);
                            const uint32_t st = it % NST, ph = (it / NST) & 1u;
                            mbar_wait(w_full + 8 * st, ph);
                            tc_fence_after();
                            if (elect_one()) {
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks)
                                    umma_bf16(tb + buf * 128, d_x2 + kb * (SLAB >> 4) + ks * 2, d_w + st * (SLAB >> 4) + ks * 2, id128,
                                              (kb | ks) > 0);
                                umma_commit(w_empty + 8 * st);
                                if (c == 3) umma_commit(x2_empty + 8 * kb);   // the x slab can take the next tile
                                if (kb == 3) umma_commit(acc_full + 8 * buf);
                            }
                            __syncwarp();
                        }
                    }
                    if (c > 0) {
                        // ---- K slice c-1 of layer 3: acc3 += relu(chunk c-1) . W3[:, (c-1)*128 .. +128]^T ----
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
                            if (pc == 3) umma_commit(acc3_full);
                        }
                        __syncwarp();
                    }
                }
                xph ^= 1u;
            }
        }
    } else {
        // ---- epilogue warps ----
        const int q = warp & 3, h = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const float *b2s = reinterpret_cast<const float *>(smem + OFF_B2);
        float *red = reinterpret_cast<float *>(smem + OFF_RED);   // [4][16] per-quadrant maxima, then [16] running maxima
        uint32_t ph_acc_full[2] = {0, 0}, ph_x3_empty[2] = {1, 1}, ph_acc3_full = 0;
        for (int patch = blockIdx.x; patch < prm.n_patches; patch += gridDim.x) {
            for (int t2 = 0; t2 < 2; ++t2) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int buf = c & 1;
                    mbar_wait(acc_full + 8 * buf, ph_acc_full[buf]);
                    ph_acc_full[buf] ^= 1u;
                    mbar_wait(x3_empty + 8 * buf, ph_x3_empty[buf]);
                    ph_x3_empty[buf] ^= 1u;
                    tc_fence_after();
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        uint32_t v[32];
                        tmem_ld32(lane_base + buf * 128 + h * 64 + j * 32, v);
                        const float4 *bb = reinterpret_cast<const float4 *>(b2s + c * 128 + h * 64 + j * 32);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 b4 = bb[i];
                            v[4 * i] = __float_as_uint(__uint_as_float(v[4 * i]) + b4.x);
                            v[4 * i + 1] = __float_as_uint(__uint_as_float(v[4 * i + 1]) + b4.y);
                            v[4 * i + 2] = __float_as_uint(__uint_as_float(v[4 * i + 2]) + b4.z);
                            v[4 * i + 3] = __float_as_uint(__uint_as_float(v[4 * i + 3]) + b4.w);
                        }
                        // channel (h*64 + j*32 + i) of this chunk -> slab h of X3[buf], 16-byte chunk (j*4 + i/8)
                        const uint32_t xb = sb + OFF_X3 + (buf * 2 + h) * SLAB + row * 128;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            uint32_t pk[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                pk[e] = pack_relu_bf16x2(__uint_as_float(v[8 * i + 2 * e]), __uint_as_float(v[8 * i + 2 * e + 1]));
                            st_shared_v4(xb + (((j * 4 + i) ^ (row & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
                        }
                    }
                    fence_async_smem();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive1(x3_full + 8 * buf);
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

}  // namespace pnt
}  // namespace pcc

/*
 * x [rows, 256] bf16 (row pitch ldx elements, 16-byte aligned), rows % 256 == 0;  w2 [512, 256] bf16 row-major, b2 [512] fp32;
 * w3_packed: pcc_mlp_pack_weights_f32(cin = 512, cout <= 16);  out [rows / 256, cout] fp32.
 */
PCC_API int pcc_pn_tail_bf16(const void *x, int64_t rows, int64_t ldx, const void *w2_bf16, const float *b2, const void *w3_packed,
                             int cout, int relu3, float *out, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(x && w2_bf16 && b2 && w3_packed && out, "pcc_pn_tail_bf16: null pointer");
    PCC_REQUIRE(rows >= 0 && rows % 256 == 0 && rows / 256 < (1ll << 30), "pcc_pn_tail_bf16: rows=%lld must be a multiple of 256",
                static_cast<long long>(rows));
    PCC_REQUIRE(cout >= 1 && cout <= pnt::C_OUT_MAX, "pcc_pn_tail_bf16: cout=%d outside [1,16]", cout);
    PCC_REQUIRE(ldx >= 256 && ldx % 8 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(w2_bf16) % 16 == 0,
                "pcc_pn_tail_bf16: x / w2 must be 16-byte aligned with a row pitch that is a multiple of 8 elements");
    if (rows == 0) return 0;
    CUtensorMap tm_x, tm_w;
    if (int r = make_tmap_bf16_2d(&tm_x, x, static_cast<uint64_t>(rows), 256, static_cast<uint64_t>(ldx), 128)) return r;
    if (int r = make_tmap_bf16_2d(&tm_w, w2_bf16, 512, 256, 256, 128)) return r;
    pnt::Params p{};
    p.b2 = b2;
    p.w3p = w3_packed;
    p.out = out;
    p.n_patches = static_cast<int>(rows / 256);
    p.c_out = cout;
    p.relu3 = relu3;
    const cudaError_t e = cudaFuncSetAttribute(pnt::pn_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pnt::SMEM);
    if (e != cudaSuccess) {
        set_error("pcc_pn_tail_bf16: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    const int grid = p.n_patches < num_sms() ? p.n_patches : num_sms();
    pnt::pn_tail_kernel<<<grid, pnt::THREADS, pnt::SMEM, static_cast<cudaStream_t>(stream)>>>(p, tm_x, tm_w);
    return check_launch("pn_tail_kernel");
}
