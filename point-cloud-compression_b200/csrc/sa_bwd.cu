// Backward pass of the SetAbstraction shared MLP (3 -> 32 -> 64 -> 128, ReLU, max over the 16 neighbours) in ONE kernel that
// recomputes the forward activations per tile instead of reading them back: the counterpart of autograd's backward through
// pn_kit.SetAbstraction's Conv2d(1x1) stack + torch.max (/root/reference/pn_kit.py:196-207 under train.py:193-221).
//
// The unfused training path writes and re-reads every activation of the 8.4 M-row stack (X1, X2, the pre-pool Y2, the
// 94 %-zero dY2, dX2, dX1: ~20 GB of HBM traffic per 32-cloud step, ~6 ms).  Here a tile of 128 positions (8 points x 16
// neighbours) lives in shared memory / TMEM from the gathered coordinates to the weight-gradient accumulators; HBM sees the
// patches (12 B / point), the neighbour bytes (16 B / point), the pooled gradient (512 B / point) and, once per CTA, the weights
// and the weight gradients.
//
// Per tile, five tensor phases alternate with five CUDA-core phases of the slot's four epilogue warps (thread = position, or
// = channel in the T-form phase):
//   P0  gather + recentre (pn_kit.py:190-191), layer 0 in fp32 (fma.rn.f32x2) -> X1+ = [X1 (32) | 1 | x y z | 0 ..] bf16
//   L1  Y1[pos, 64]   = X1+ . W1+^T          (W1+ = [W1 | b1 | 0 ..]: the bias rides on the ones channel, as in the forward kernel)
//   P1  X2 = relu(Y1) -> bf16
//   L2  Y2^T[ch, pos] = W2 . X2^T            (T-form: a lane holds one channel, its 16 neighbours are 16 registers)
//   P2  per (point, channel): max + arg-max over the 16 neighbours, pooled = max + b2, g' = g * [pooled > 0]  (db2 += g')
//       dY2^T[ch, pos] = g' at the arg-max position, 0 elsewhere -> bf16
//   M3  dX2[pos, 64]  = dY2 . W2             and   dW2[ch, 64] += dY2^T . X2
//   P3  dX2 *= [X2 > 0] -> bf16 (E)
//   M4  dX1[pos, 32]  = E . W1
//   P4  dX1 *= [X1 > 0] -> bf16 (F = [dX1 | 0])
//   M5  [dW1 | db1 | . ; . | db0 | dW0] += [E ; F]^T . X1+      (rows 0..63 = layer 1, rows 64..95 = layer 0: one accumulator)
// Every operand is one of the 128-byte-swizzled [128 x 64] bf16 slabs, read K-major or MN-major as the contraction needs
// (the MN-major form is the one wgrad_ws.cu uses), so nothing is ever transposed in memory.
// Two slots per CTA (two independent tiles in flight, each with its own accumulators: 2 x (128 + 128) TMEM columns).
#include <cuda.h>
#include <cuda_bf16.h>

#include "pcc_common.cuh"
#include "tc_ptx.cuh"

namespace pcc {
namespace sab {

constexpr int P = 128;                    // positions per tile
constexpr int SLAB = P * 128;             // [128 x 64] bf16, 16 KB
constexpr int NS = 2;                     // slots per CTA
constexpr int OFF_W2 = 0;                 // [128 ch x 64] bf16
constexpr int OFF_W1 = OFF_W2 + SLAB;     // [64 x 64] bf16: cols 0..31 = W1, col 32 = b1, rest 0 (8 KB)
constexpr int OFF_SLOT = OFF_W1 + SLAB / 2;
constexpr int SL_X1 = 0, SL_X2 = SLAB, SL_DY = 2 * SLAB;   // dY2^T: two slabs (positions 0..63, 64..127); E, F alias them
constexpr int SLOT_BYTES = 4 * SLAB;
constexpr int OFF_W0 = OFF_SLOT + NS * SLOT_BYTES;   // layer-0 weights as channel pairs (512 B)
constexpr int OFF_BAR = OFF_W0 + 512;
constexpr int SMEM = OFF_BAR + 64 + 1024;
constexpr int THREADS = 160 * NS;         // per slot: 4 epilogue warps + 1 issuing warp
constexpr int TMEM_COLS = 512;            // per slot: 128 (tile accumulator) + 64 (dW2) + 64 (dW1 / dW0)

struct Params {
    const float *xyz;            // patches [BS * Pp, 3]
    const unsigned char *idx8;   // [points, 16]
    const float *g;              // [points, 128] gradient of the pooled output: fp32 rows of g_ld elements, or ...
    const __nv_bfloat16 *g16;    // ... bf16 rows (non-NULL: autograd hands the PointNet stack's input gradient over as a bf16 slice)
    long long g_ld;
    const float *w0, *b0, *w1, *b1, *w2, *b2;
    float *dw0, *db0, *dw1, *db1, *dw2, *db2;   // accumulated with atomics (zeroed by the caller)
    int n_tiles, pts_per_patch, pts_shift;
};

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = static_cast<uint64_t>((saddr & 0x3ffffu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc_major(int M, int N, uint32_t a_mn, uint32_t b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn << 15) | (b_mn << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
           (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t bf16_bits(float v) { return static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(v))); }
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
// address of 16-byte chunk `c` of row `r` in a swizzled slab
__device__ __forceinline__ uint32_t chunk_addr(uint32_t slab, int r, int c) { return slab + r * 128 + ((c ^ (r & 7)) << 4); }

template <int SLOT>
__device__ __forceinline__ void slot_sync() {   // the slot's four epilogue warps + its issuing warp
    if (SLOT == 0) asm volatile("bar.sync 1, 160;" ::: "memory");
    else asm volatile("bar.sync 2, 160;" ::: "memory");
}

__global__ void __launch_bounds__(THREADS, 1) sa_bwd_kernel(const __grid_constant__ Params prm) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t sb = smem_u32(smem);
    const int tid = threadIdx.x, warp = __shfl_sync(FULL_MASK, tid >> 5, 0), lane = tid & 31;
    const uint32_t bar_mma = sb + OFF_BAR;   // [NS]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 32);

    // ---- weights -> swizzled bf16 slabs; zero the activation slabs once (their padding columns are never written again) ----
    for (int e = tid; e < 128 * 8; e += THREADS) {          // W2: row = channel, chunk = 8 input channels
        const int r = e >> 3, c = e & 7;
        const float *s = prm.w2 + r * 64 + c * 8;
        uint32_t pk[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) pk[h] = bf16_bits(__ldg(s + 2 * h)) | (bf16_bits(__ldg(s + 2 * h + 1)) << 16);
        st_shared_v4(chunk_addr(sb + OFF_W2, r, c), pk[0], pk[1], pk[2], pk[3]);
    }
    for (int e = tid; e < 64 * 8; e += THREADS) {           // W1+: cols 0..31 = W1, col 32 = b1, rest 0
        const int r = e >> 3, c = e & 7;
        uint32_t pk[4] = {0u, 0u, 0u, 0u};
        if (c < 4) {
            const float *s = prm.w1 + r * 32 + c * 8;
#pragma unroll
            for (int h = 0; h < 4; ++h) pk[h] = bf16_bits(__ldg(s + 2 * h)) | (bf16_bits(__ldg(s + 2 * h + 1)) << 16);
        } else if (c == 4) {
            pk[0] = bf16_bits(__ldg(prm.b1 + r));
        }
        st_shared_v4(chunk_addr(sb + OFF_W1, r, c), pk[0], pk[1], pk[2], pk[3]);
    }
    for (int e = tid; e < NS * SLOT_BYTES / 16; e += THREADS) st_shared_v4(sb + OFF_SLOT + e * 16, 0u, 0u, 0u, 0u);
    // layer-0 weights as channel PAIRS for the packed fp32 FMA: pair j = channels (2j, 2j + 1) -> {wx wx' | wy wy' | wz wz' | b b'}
    for (int e = tid; e < 32 * 4; e += THREADS) {
        const int ch = 2 * (e >> 3) + (e & 1), comp = (e >> 1) & 3;
        reinterpret_cast<float *>(smem + OFF_W0)[e] = comp < 3 ? __ldg(prm.w0 + ch * 3 + comp) : __ldg(prm.b0 + ch);
    }
    if (tid == 0)
        for (int s = 0; s < NS; ++s) mbar_init(bar_mma + 8 * s, 1);
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_tiles = prm.n_tiles;
    const int tstride = NS * gridDim.x;
    const int s = warp / 5, wq = warp - 5 * s;          // slot, warp within the slot (0..3 epilogue, 4 issuer)
    const uint32_t slot = sb + OFF_SLOT + s * SLOT_BYTES;
    const uint32_t acc_col = s * 256;                   // accumulator columns of the slot: [0,128) tile, [128,192) dW2, [192,256) dW1/dW0
    const long long tile0 = static_cast<long long>(NS) * blockIdx.x + s;

    if (wq == 4) {
        // ---- issuing warp: five tensor phases per tile, each after the epilogue group's hand-over ----
        const uint32_t acc = __shfl_sync(FULL_MASK, tmem_base, 0) + acc_col;
        const uint32_t id_kk = umma_idesc(128, 64), id_kk128 = umma_idesc(128, 128);
        const uint32_t id_mm = idesc_major(128, 64, 1u, 1u), id_km = idesc_major(128, 64, 0u, 1u);
        const uint64_t d_x1 = umma_desc_sw128(slot + SL_X1), d_x2 = umma_desc_sw128(slot + SL_X2);
        const uint64_t d_w1 = umma_desc_sw128(sb + OFF_W1), d_w2 = umma_desc_sw128(sb + OFF_W2);
        const uint64_t d_dy_k0 = umma_desc_sw128(slot + SL_DY), d_dy_k1 = umma_desc_sw128(slot + SL_DY + SLAB);   // dY2^T, K = positions
        const uint64_t m_dy = desc_mn_sw128(slot + SL_DY, SLAB);        // dY2^T read MN-major (M = positions), also [E ; F]
        const uint64_t m_w2 = desc_mn_sw128(sb + OFF_W2, SLAB), m_w1 = desc_mn_sw128(sb + OFF_W1, SLAB);
        const uint64_t m_x1 = desc_mn_sw128(slot + SL_X1, SLAB), m_x2 = desc_mn_sw128(slot + SL_X2, SLAB);
        bool first = true;
        for (long long tile = tile0; tile < n_tiles; tile += tstride) {
            // L1: Y1[pos, 64] = X1+ (K = 48: X1, ones, xyz) . W1+^T
            if (s == 0) slot_sync<0>(); else slot_sync<1>();
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 3; ++ks) umma_bf16(acc, d_x1 + 2 * ks, d_w1 + 2 * ks, id_kk, ks > 0);
                umma_commit(bar_mma + 8 * s);
            }
            __syncwarp();
            // L2 (T-form): Y2^T[ch, pos] = W2 . X2^T
            if (s == 0) slot_sync<0>(); else slot_sync<1>();
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_bf16(acc, d_w2 + 2 * ks, d_x2 + 2 * ks, id_kk128, ks > 0);
                umma_commit(bar_mma + 8 * s);
            }
            __syncwarp();
            // M3: dX2[pos, 64] = dY2 . W2 (K = 128 channels)  and  dW2[ch, 64] += dY2^T . X2 (K = 128 positions)
            if (s == 0) slot_sync<0>(); else slot_sync<1>();
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) umma_bf16(acc, m_dy + 128 * ks, m_w2 + 128 * ks, id_mm, ks > 0);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma_bf16(acc + 128, (ks < 4 ? d_dy_k0 : d_dy_k1) + 2 * (ks & 3), m_x2 + 128 * ks, id_km, (first && ks == 0) ? 0u : 1u);
                umma_commit(bar_mma + 8 * s);
            }
            __syncwarp();
            // M4: dX1[pos, 64 (32 used)] = E . W1+ (K = 64 channels)
            if (s == 0) slot_sync<0>(); else slot_sync<1>();
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_bf16(acc, d_dy_k0 + 2 * ks, m_w1 + 128 * ks, id_km, ks > 0);
                umma_commit(bar_mma + 8 * s);
            }
            __syncwarp();
            // M5: [dW1 | db1 ; db0 | dW0] += [E ; F]^T . X1+ (K = 128 positions)
            if (s == 0) slot_sync<0>(); else slot_sync<1>();
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) umma_bf16(acc + 192, m_dy + 128 * ks, m_x1 + 128 * ks, id_mm, (first && ks == 0) ? 0u : 1u);
                umma_commit(bar_mma + 8 * s);
            }
            __syncwarp();
            first = false;
        }
    } else {
        // ---- epilogue group of slot s; a warp can only read the TMEM lanes of its own quarter, 32 (warp % 4) .. (slot 1's warps
        // 5..8 are quarters 1, 2, 3, 0: every quarter is covered once per slot) ----
        const int q = warp & 3;
        const int row = q * 32 + lane;                        // position (P0, P1, P3, P4) or channel (P2)
        const uint32_t acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc_col;
        const uint32_t w0s = sb + OFF_W0;
        const float bias2 = __ldg(prm.b2 + row);
        float db2 = 0.0f;
        uint32_t ph = 0;
        auto hand_over = [&]() {        // my slab writes are done -> issuing warp; then wait for the tensor phase
            fence_async_smem();
            tc_fence_before();
            if (s == 0) slot_sync<0>(); else slot_sync<1>();
            mbar_wait(bar_mma + 8 * s, ph);
            ph ^= 1u;
            tc_fence_after();
        };
        // gather of a tile: neighbour byte, then centre + neighbour coordinates (loads in flight across the previous tile)
        float gc[3], gn[3];
        auto issue_gather = [&](long long tl) {
            const unsigned nb = __ldg(prm.idx8 + tl * P + row);
            const unsigned g0 = static_cast<unsigned>(tl) * 8u;
            const unsigned pbase = prm.pts_shift >= 0 ? (g0 >> prm.pts_shift) << prm.pts_shift
                                                      : g0 / static_cast<unsigned>(prm.pts_per_patch) * static_cast<unsigned>(prm.pts_per_patch);
            const float *c = prm.xyz + static_cast<size_t>(g0 + (row >> 4)) * 3;
            const float *n = prm.xyz + static_cast<size_t>(pbase + nb) * 3;
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                gc[e] = __ldg(c + e);
                gn[e] = __ldg(n + e);
            }
        };
        if (tile0 < n_tiles) issue_gather(tile0);
        for (long long tile = tile0; tile < n_tiles; tile += tstride) {
            float gq[8];                                       // pooled gradient of my channel, the tile's 8 points (used in P2)
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const long long o = (static_cast<long long>(tile) * 8 + p) * prm.g_ld + row;
                gq[p] = prm.g16 ? __bfloat162float(prm.g16[o]) : __ldg(prm.g + o);
            }
            // ---- P0: recentre, layer 0 (fp32), X1+ row ----
            {
                const float lx = __fsub_rn(gn[0], gc[0]), ly = __fsub_rn(gn[1], gc[1]), lz = __fsub_rn(gn[2], gc[2]);
                if (tile + tstride < n_tiles) issue_gather(tile + tstride);
                const uint64_t qx = dup_f32x2(lx), qy = dup_f32x2(ly), qz = dup_f32x2(lz);
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    uint32_t pk[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const ulonglong2 wxy = ld_shared_v2u64(w0s + 32 * (4 * c8 + j)), wzb = ld_shared_v2u64(w0s + 32 * (4 * c8 + j) + 16);
                        pk[j] = pack_relu_bf16x2_pair(fma_f32x2(wzb.x, qz, fma_f32x2(wxy.y, qy, fma_f32x2(wxy.x, qx, wzb.y))));
                    }
                    st_shared_v4(chunk_addr(slot + SL_X1, row, c8), pk[0], pk[1], pk[2], pk[3]);
                }
                st_shared_v4(chunk_addr(slot + SL_X1, row, 4), 0x3f80u | (bf16_bits(lx) << 16), bf16_bits(ly) | (bf16_bits(lz) << 16), 0u, 0u);
            }
            hand_over();   // -> L1
            // ---- P1: X2 = relu(Y1) -> bf16 row ----
            {
                uint32_t v0[32], v1[32];
                tmem_ld32_issue(acc, v0);
                tmem_ld32_issue(acc + 32, v1);
                tmem_ld32_wait(v0);
                tmem_ld32_wait(v1);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t(&v)[32] = c < 4 ? v0 : v1;
                    const int b = (c & 3) * 8;
                    st_shared_v4(chunk_addr(slot + SL_X2, row, c),
                                 pack_relu_bf16x2(__uint_as_float(v[b]), __uint_as_float(v[b + 1])),
                                 pack_relu_bf16x2(__uint_as_float(v[b + 2]), __uint_as_float(v[b + 3])),
                                 pack_relu_bf16x2(__uint_as_float(v[b + 4]), __uint_as_float(v[b + 5])),
                                 pack_relu_bf16x2(__uint_as_float(v[b + 6]), __uint_as_float(v[b + 7])));
                }
            }
            hand_over();   // -> L2
            // ---- P2 (thread = channel `row`): arg-max over each point's 16 neighbours, masked gradient -> dY2^T row ----
#pragma unroll
            for (int h = 0; h < 4; ++h) {                      // 32 accumulator columns = 2 points
                uint32_t v[32];
                tmem_ld32(acc + 32 * h, v);
#pragma unroll
                for (int pp = 0; pp < 2; ++pp) {
                    const int p = 2 * h + pp;
                    float m = __uint_as_float(v[16 * pp]);
                    int j = 0;
#pragma unroll
                    for (int i = 1; i < 16; ++i) {
                        const float x = __uint_as_float(v[16 * pp + i]);
                        if (x > m) {                           // strict: ties keep the first neighbour
                            m = x;
                            j = i;
                        }
                    }
                    const float gp = (m + bias2 > 0.0f) ? gq[p] : 0.0f;
                    db2 += gp;
                    const uint32_t val = bf16_bits(gp) << (16 * (j & 1));
                    const int wsel = (j & 7) >> 1, csel = j >> 3;
                    uint32_t w[2][4];
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int k = 0; k < 4; ++k) w[c][k] = (c == csel && k == wsel) ? val : 0u;
                    // positions 16 p .. 16 p + 15 of row `row`: slab p / 4, chunks 2 (p % 4), 2 (p % 4) + 1
                    const uint32_t dslab = slot + SL_DY + (p >> 2) * SLAB;
                    st_shared_v4(chunk_addr(dslab, row, 2 * (p & 3)), w[0][0], w[0][1], w[0][2], w[0][3]);
                    st_shared_v4(chunk_addr(dslab, row, 2 * (p & 3) + 1), w[1][0], w[1][1], w[1][2], w[1][3]);
                }
            }
            hand_over();   // -> M3
            // ---- P3 (thread = position): dX2 masked by X2 > 0 -> E row (aliases dY2^T's first slab: the tensor phase is over) ----
            {
                uint32_t v0[32], v1[32];
                tmem_ld32_issue(acc, v0);
                tmem_ld32_issue(acc + 32, v1);
                tmem_ld32_wait(v0);
                tmem_ld32_wait(v1);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t(&v)[32] = c < 4 ? v0 : v1;
                    const int b = (c & 3) * 8;
                    const uint4 x = ld_shared_v4(chunk_addr(slot + SL_X2, row, c));   // X2 >= +0 after the ReLU: > 0 <=> bits != 0
                    const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
                    uint32_t pk[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float lo = (xs[k] & 0xffffu) ? __uint_as_float(v[b + 2 * k]) : 0.0f;
                        const float hi = (xs[k] >> 16) ? __uint_as_float(v[b + 2 * k + 1]) : 0.0f;
                        pk[k] = pack_bf16x2(lo, hi);
                    }
                    st_shared_v4(chunk_addr(slot + SL_DY, row, c), pk[0], pk[1], pk[2], pk[3]);
                }
            }
            hand_over();   // -> M4
            // ---- P4 (thread = position): dX1 masked by X1 > 0 -> F row = [dX1 (32) | 0] (dY2^T's second slab) ----
            {
                uint32_t v[32];
                tmem_ld32(acc, v);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint32_t pk[4] = {0u, 0u, 0u, 0u};
                    if (c < 4) {
                        const uint4 x = ld_shared_v4(chunk_addr(slot + SL_X1, row, c));
                        const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float lo = (xs[k] & 0xffffu) ? __uint_as_float(v[8 * c + 2 * k]) : 0.0f;
                            const float hi = (xs[k] >> 16) ? __uint_as_float(v[8 * c + 2 * k + 1]) : 0.0f;
                            pk[k] = pack_bf16x2(lo, hi);
                        }
                    }
                    st_shared_v4(chunk_addr(slot + SL_DY + SLAB, row, c), pk[0], pk[1], pk[2], pk[3]);
                }
            }
            hand_over();   // -> M5 (reads X1+, E, F: the next tile's P0 / P2 may overwrite them once it is complete)
        }
        // ---- the CTA's weight gradients: TMEM -> global (atomics over the CTAs) ----
        if (tile0 < n_tiles) {
            atomicAdd(prm.db2 + row, db2);
            uint32_t v0[32], v1[32];
            tmem_ld32_issue(acc + 128, v0);           // dW2[row = channel, 0..63]
            tmem_ld32_issue(acc + 160, v1);
            tmem_ld32_wait(v0);
            tmem_ld32_wait(v1);
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                atomicAdd(prm.dw2 + row * 64 + k, __uint_as_float(v0[k]));
                atomicAdd(prm.dw2 + row * 64 + 32 + k, __uint_as_float(v1[k]));
            }
            tmem_ld32_issue(acc + 192, v0);           // rows 0..63: [dW1 | db1 | .], rows 64..95: [. | db0 | dW0 x y z]
            tmem_ld32_issue(acc + 224, v1);
            tmem_ld32_wait(v0);
            tmem_ld32_wait(v1);
            if (row < 64) {
#pragma unroll
                for (int k = 0; k < 32; ++k) atomicAdd(prm.dw1 + row * 32 + k, __uint_as_float(v0[k]));
                atomicAdd(prm.db1 + row, __uint_as_float(v1[0]));
            } else if (row < 96) {
                const int c = row - 64;
                atomicAdd(prm.db0 + c, __uint_as_float(v1[0]));
                atomicAdd(prm.dw0 + c * 3 + 0, __uint_as_float(v1[1]));
                atomicAdd(prm.dw0 + c * 3 + 1, __uint_as_float(v1[2]));
                atomicAdd(prm.dw0 + c * 3 + 2, __uint_as_float(v1[3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

}  // namespace sab
}  // namespace pcc

// Weight gradients of the SetAbstraction stack whose forward pass is pcc_sa_chain_indexed: see include/pcc_b200.h
PCC_API int pcc_sa_chain_indexed_bwd(const float *patches, const unsigned char *idx8, int64_t points, int pts_per_patch, const float *w0,
                                     const float *b0, const float *w1, const float *b1, const float *w2, const float *b2,
                                     const void *grad_out, int grad_dtype, int64_t grad_ld, float *dw0, float *db0, float *dw1, float *db1,
                                     float *dw2, float *db2, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(patches && idx8 && w0 && b0 && w1 && b1 && w2 && b2 && grad_out && dw0 && db0 && dw1 && db1 && dw2 && db2,
                "pcc_sa_chain_indexed_bwd: null pointer");
    PCC_REQUIRE(points >= 0 && points % 8 == 0 && pts_per_patch >= 8 && pts_per_patch <= 256 && pts_per_patch % 8 == 0 &&
                    points % pts_per_patch == 0 && points / 8 < (1ll << 28),
                "pcc_sa_chain_indexed_bwd: bad shape points=%lld pts_per_patch=%d", static_cast<long long>(points), pts_per_patch);
    PCC_REQUIRE((grad_dtype == 0 || grad_dtype == 1) && grad_ld >= 128, "pcc_sa_chain_indexed_bwd: grad_dtype=%d grad_ld=%lld", grad_dtype,
                static_cast<long long>(grad_ld));
    if (points == 0) return 0;
    sab::Params p{};
    p.xyz = patches;
    p.idx8 = idx8;
    p.g = grad_dtype == 0 ? static_cast<const float *>(grad_out) : nullptr;
    p.g16 = grad_dtype == 1 ? static_cast<const __nv_bfloat16 *>(grad_out) : nullptr;
    p.g_ld = grad_ld;
    p.w0 = w0; p.b0 = b0; p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2;
    p.dw0 = dw0; p.db0 = db0; p.dw1 = dw1; p.db1 = db1; p.dw2 = dw2; p.db2 = db2;
    p.n_tiles = static_cast<int>(points / 8);
    p.pts_per_patch = pts_per_patch;
    p.pts_shift = -1;
    for (int sh = 3; sh <= 8; ++sh)
        if ((1 << sh) == pts_per_patch) p.pts_shift = sh;
    static bool attr_done_dev[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (!attr_done_dev[dev]) {
        const cudaError_t e = cudaFuncSetAttribute(sab::sa_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sab::SMEM);
        if (e != cudaSuccess) {
            set_error("pcc_sa_chain_indexed_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
            return static_cast<int>(e);
        }
        attr_done_dev[dev] = true;
    }
    const int sms = num_sms();
    const int pairs = (p.n_tiles + sab::NS - 1) / sab::NS;
    sab::sa_bwd_kernel<<<pairs < sms ? pairs : sms, sab::THREADS, sab::SMEM, static_cast<cudaStream_t>(stream)>>>(p);
    return check_launch("sa_bwd_kernel");
}
