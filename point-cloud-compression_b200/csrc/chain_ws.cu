// Warp-specialised, compile-time-shaped fused chains for the three shared-MLP stacks of the IPDAE patch auto-encoder
// (sm_100a: tcgen05.mma + TMEM + TMA).  They are what pcc_mlp_chain() runs when the call matches one of
//   SA   pn_kit.SetAbstraction body   3 -> 32 -> 64 -> 128 (+ReLU), max over 16 neighbours   /root/reference/pn_kit.py:196-207
//   PNF  pn_kit.PointNet, first half   [feat 128 | xyz 3] -> 128 -> 256 (+ReLU)                /root/reference/pn_kit.py:124-144
//   DEC  pn_kit.MLP (AE.inv_mlp)       [lin 128 | latent 16] -> 128 -> 64 -> 32 -> 3           /root/reference/AE.py:50-53
// (anything else keeps using the runtime-shaped kernel in mlp_chain.cu).
//
// Why a second kernel family: the runtime-shaped kernel is CUDA-core *issue* bound (ncu: 1300 thread-instructions per
// position for SA, tensor pipe 19 % active) - TMEM read-back itself sustains > 600 B/clk/SM (profiles/r01_ubench_tmem.txt).
// Here every shape is a constant, so the epilogues reduce to tcgen05.ld -> F2FP.RELU pack -> 16-byte st.shared, and the
// roles are split across warps so the tensor pipe and the epilogue overlap inside one CTA:
//   warps 0-7  epilogue: warp w owns TMEM lanes 32*(w%4).. and column half w/4 of every accumulator; they also convert
//              the small fp32 inputs (xyz / latent) into a 16-wide K block of the first MMA
//   warp 8     one thread issues every tcgen05.mma and commits to mbarriers
//   warp 9     one thread issues the TMA loads of the bf16 feature tiles (PNF / DEC)
// Layouts: activations are [128 positions][64 ch] bf16 slabs with the 128-byte swizzle - exactly what TMA writes and
// what a K-major UMMA descriptor (layout type 2) reads, whether the slab is the A operand (N-form, D[pos, ch]) or the B
// operand (T-form pooled layer, D^T[ch, pos]).  Weights stay in the packed no-swizzle image of pcc_mlp_pack_weights_f32
// (bias in column cin).  The bias column multiplies a constant [128 x 16] "ones" block (column 0 = 1) that is shared by
// all layers, so activation slabs hold real channels only and never need padding writes.
// SA keeps two tiles in flight per CTA (ping-pong slots, 2 CTAs / SM); its 3-channel first layer runs on the tensor
// core at ~fp32 accuracy by splitting coordinates and weights into bf16 hi + lo parts inside one K = 16 step.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "chain_ws.h"
#include "pcc_common.cuh"
#include "tc_ptx.cuh"

namespace pcc {
namespace ws {

constexpr int P = 128;              // positions per tile
constexpr int SLAB = P * 128;       // bytes of one [128 x 64] bf16 slab (SWIZZLE_128B)
constexpr int K16 = 4096;           // bytes of one [128 x 16] bf16 block, canonical no-swizzle (LBO 2048, SBO 128)
constexpr int EPI_THREADS = 256;
constexpr uint32_t ONE_BF16 = 0x3f80u;

__device__ __forceinline__ uint32_t bf16_bits(float v) {
    return static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(v)));
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// canonical K16 block descriptors (activation side: rows = positions)
__device__ __forceinline__ uint64_t desc_k16(uint32_t saddr) { return umma_desc(saddr, 2048, 128); }
// packed weights (rows = channels, kp columns): K step ks starts ks * 256 bytes in; LBO 128, SBO kp * 16
__device__ __forceinline__ uint64_t desc_w(uint32_t saddr, int kp, int ks) { return umma_desc(saddr + ks * 256, 128, kp * 16); }

// row r of a [128 x 16] canonical block: two 16-byte chunks (k 0..7 at r*16, k 8..15 at 2048 + r*16)
__device__ __forceinline__ void store_k16_row(uint32_t blk, int r, const uint32_t (&w)[8]) {
    st_shared_v4(blk + r * 16, w[0], w[1], w[2], w[3]);
    st_shared_v4(blk + 2048 + r * 16, w[4], w[5], w[6], w[7]);
}

// NC accumulator columns (channels c0 .. c0+NC-1 of row r) -> ReLU -> bf16 -> the swizzled slabs starting at xbase
template <int NC, bool RELU>
__device__ __forceinline__ void store_row_chunks(uint32_t xbase, int r, int c0, const uint32_t (&v)[NC]) {
#pragma unroll
    for (int i = 0; i < NC / 8; ++i) {
        const int c = c0 + 8 * i;
        const uint32_t addr = xbase + (c >> 6) * SLAB + r * 128 + ((((c >> 3) & 7) ^ (r & 7)) << 4);
        uint32_t pk[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const float lo = __uint_as_float(v[8 * i + 2 * h]), hi = __uint_as_float(v[8 * i + 2 * h + 1]);
            pk[h] = RELU ? pack_relu_bf16x2(lo, hi) : pack_bf16x2(lo, hi);
        }
        st_shared_v4(addr, pk[0], pk[1], pk[2], pk[3]);
    }
}

__device__ __forceinline__ void copy_to_smem(unsigned char *dst, const void *src, int bytes, int tid, int nthreads) {
    const int4 *s = static_cast<const int4 *>(src);
    int4 *d = reinterpret_cast<int4 *>(dst);
    for (int i = tid; i < bytes / 16; i += nthreads) d[i] = __ldg(s + i);
}

__device__ __forceinline__ void fill_ones_block(unsigned char *blk, int tid, int nthreads) {
    for (int i = tid; i < K16 / 16; i += nthreads) {
        // chunk i < 128 is (row i, k 0..7): k = 0 holds 1.0; the k 8..15 half is zero
        reinterpret_cast<uint4 *>(blk)[i] = make_uint4(i < 128 ? ONE_BF16 : 0u, 0u, 0u, 0u);
    }
}

__device__ __forceinline__ uint32_t tmem_alloc_and_sync(uint32_t *slot, int cols, int warp, int alloc_warp) {
    if (warp == alloc_warp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return *slot;
}
__device__ __forceinline__ void tmem_free(uint32_t base, int cols, int warp, int alloc_warp) {
    tc_fence_before();
    __syncthreads();
    if (warp == alloc_warp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}

// ======================================================================================================================
// SA: xyz[M,3] fp32 -> 32 -> 64 -> 128, max over each run of 16 positions -> out[M/16, 128]
// ======================================================================================================================
// A tile's critical path is a chain of MMA -> epilogue -> MMA hand-offs (0.5-1 k clocks each, measured with the ticks
// below), and TMEM (128 accumulator columns per tile for the pooled T-form layer) caps the tiles in flight at 4 per SM.
// So (i) the 3 -> 32 layer does not take a tensor round trip: it runs in fp32 on the CUDA cores (the reference's own
// arithmetic) with the weights of a thread's 8 channels held in registers, and (ii) the four tiles in flight on an SM
// are four independent chains: 2 CTAs x 2 slots, each slot served by its own group of four epilogue warps, and one MMA
// thread per CTA that polls both slots and issues whichever layer is ready.
struct SaParams {
    const float *xyz;
    long long ld;
    const float *w0, *b0;      // fp32 [32,3], [32]
    const float *b2;           // fp32 [128] or NULL (sa_chain2_kernel): layer 2's bias, added after the max in the epilogue; NULL:
                               // the bf16 bias column of the packed weights goes through one more MMA, as in layers 1
    const void *w1p, *w2p;     // packed [64 x 48], [128 x 80]
    void *out;
    int out_bf16;
    int n_tiles;
    long long *dbg;            // bring-up aid (NULL in production): clock64() ticks of CTA 0's MMA thread and epilogue thread 0
    // indexed form (sa_chain2_kernel<1>): position (point g, neighbour n) reads pts[patch(g) * P + idx8[16 g + n]] - pts[g]
    const unsigned char *idx8; // [points, 16] neighbour indices inside the point's own patch
    int pts_per_patch;         // P <= 256
    int pts_shift;             // log2(P) when P is a power of two, else -1
};

namespace sa {
constexpr int KP1 = 48, KP2 = 80;
constexpr int OFF_W1 = 0;
constexpr int OFF_W2 = OFF_W1 + 64 * KP1 * 2;        // 6144
constexpr int OFF_ONES = OFF_W2 + 128 * KP2 * 2;     // 26624 = 26 * 1024
constexpr int OFF_SLOT = OFF_ONES + K16;             // 30720 = 30 * 1024
constexpr int SL_X1 = 0, SL_X2 = SLAB;
constexpr int SLOT_BYTES = 2 * SLAB;                 // 32768
constexpr int OFF_BAR = OFF_SLOT + 2 * SLOT_BYTES;   // 96256
constexpr int OFF_W0 = OFF_BAR + 64;                 // float4 (w0, w1, w2, bias) x 32 channels
constexpr int SMEM = OFF_W0 + 512 + 1024;            // + alignment slack
constexpr int THREADS = 320;                         // 2 groups of 4 epilogue warps + one MMA warp per slot
constexpr int TMEM_COLS = 256;                       // 128 per slot
// indexed form only: recentred neighbour coordinates of the next tiles, SoA x[128] y[128] z[128], 2 slots x 2 buffers, and
// one "filled" barrier per buffer
constexpr int OFF_XYZ = OFF_W0 + 512;
constexpr int XYZ_BUF = 3 * P * 4;                   // 1536
constexpr int OFF_BAR2 = OFF_XYZ + 4 * XYZ_BUF;      // 4 mbarriers
constexpr int SMEM_IDX = OFF_BAR2 + 32 + 1024;
static_assert(OFF_SLOT % 1024 == 0 && SLOT_BYTES % 1024 == 0, "slab alignment");
// sa_chain2_kernel's layout for NS accumulator slots per CTA (NS = 2 is the layout above)
template <int NS>
struct Lay {
    static constexpr int OFF_BAR = OFF_SLOT + NS * SLOT_BYTES;   // 3 x NS mbarriers + the TMEM address word
    static constexpr int OFF_W0 = OFF_BAR + 32 * NS;
    static constexpr int SMEM = OFF_W0 + 512 + 1024;
    static constexpr int THREADS = 160 * NS;                     // per slot: 4 epilogue warps + one MMA warp
    static constexpr int TMEM_COLS = 128 * NS;
    static constexpr int OFF_XYZ = OFF_W0 + 512;
    static constexpr int OFF_BAR2 = OFF_XYZ + 2 * NS * XYZ_BUF;  // 2 x NS mbarriers
    static constexpr int SMEM_IDX = OFF_BAR2 + 16 * NS + 1024;
};
static_assert(Lay<2>::OFF_BAR == OFF_BAR && Lay<2>::OFF_W0 == OFF_W0 && Lay<2>::SMEM == SMEM && Lay<2>::SMEM_IDX == SMEM_IDX, "Lay<2>");
}  // namespace sa

template <bool TIMED>
__global__ void __launch_bounds__(sa::THREADS, 2) sa_chain_kernel(const __grid_constant__ SaParams prm) {
    using namespace sa;
    int tick = 0;
#define SA_TICK(base)                                                                                              \
    do {                                                                                                           \
        if constexpr (TIMED) {                                                                                     \
            if (blockIdx.x == 0 && warp != 9 && tick < 500) prm.dbg[(base) + tick++] = clock64();                   \
        }                                                                                                          \
    } while (0)
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t sb = smem_u32(smem);
    const int tid = threadIdx.x, warp = __shfl_sync(FULL_MASK, tid >> 5, 0), lane = tid & 31;
    const uint32_t bar_acc = sb + OFF_BAR, bar_act = sb + OFF_BAR + 16;  // [2] each
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 48);

    // ---- prologue ----
    copy_to_smem(smem + OFF_W1, prm.w1p, 64 * KP1 * 2, tid, THREADS);
    copy_to_smem(smem + OFF_W2, prm.w2p, 128 * KP2 * 2, tid, THREADS);
    fill_ones_block(smem + OFF_ONES, tid, THREADS);
    for (int e = tid; e < 32 * 4; e += THREADS)
        reinterpret_cast<float *>(smem + OFF_W0)[e] = (e & 3) < 3 ? prm.w0[(e >> 2) * 3 + (e & 3)] : prm.b0[e >> 2];
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_acc + 8 * s, 1);
            mbar_init(bar_act + 8 * s, 4);
        }
    }
    const uint32_t tmem_base = tmem_alloc_and_sync(tmem_slot, TMEM_COLS, warp, 8);

    // slot s of this CTA walks tiles 2 * blockIdx.x + s, + 2 * gridDim.x, ...
    const int n_tiles = prm.n_tiles;
    const int tstride = 2 * gridDim.x;

    if (warp >= 8) {
        // ---- MMA issuers: warp 8 + s serves slot s; warp-uniform control flow, one elected lane issues ----
        const int s = warp - 8;
        const uint32_t acc = __shfl_sync(FULL_MASK, tmem_base, 0) + s * 128;
        const uint32_t id64 = umma_idesc(128, 64), id128 = umma_idesc(128, 128);
        const uint64_t d_ones = desc_k16(sb + OFF_ONES);
        const uint64_t d_w1 = desc_w(sb + OFF_W1, KP1, 0), d_w2 = desc_w(sb + OFF_W2, KP2, 0);   // + 16 per K step (256 B)
        const uint64_t d_x1 = umma_desc_sw128(sb + OFF_SLOT + s * SLOT_BYTES + SL_X1);            // + 2 per K step (32 B)
        const uint64_t d_x2 = umma_desc_sw128(sb + OFF_SLOT + s * SLOT_BYTES + SL_X2);
        uint32_t ph = 0;
        for (int tile = 2 * blockIdx.x + s; tile < n_tiles; tile += tstride) {
            mbar_wait(bar_act + 8 * s, ph);
            ph ^= 1u;
            tc_fence_after();
            if (lane == 0) SA_TICK(512);
            if (elect_one()) {
                umma_bf16(acc, d_x1, d_w1, id64, 0u);
                umma_bf16(acc, d_x1 + 2, d_w1 + 16, id64, 1u);
                umma_bf16(acc, d_ones, d_w1 + 32, id64, 1u);
                umma_commit(bar_acc + 8 * s);
            }
            __syncwarp();
            if (lane == 0) SA_TICK(512);
            mbar_wait(bar_act + 8 * s, ph);
            ph ^= 1u;
            tc_fence_after();
            if (lane == 0) SA_TICK(512);
            if (elect_one()) {  // T-form: D^T[128 channels, 128 positions]
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_bf16(acc, d_w2 + 16 * ks, d_x2 + 2 * ks, id128, ks > 0);
                umma_bf16(acc, d_w2 + 64, d_ones, id128, 1u);
                umma_commit(bar_acc + 8 * s);
            }
            __syncwarp();
            if (lane == 0) SA_TICK(512);
        }
    } else {
        // ---- epilogue group g = warp / 4 serves slot g; warp q = warp % 4 owns TMEM lanes 32q.. ----
        const int s = warp >> 2, q = warp & 3;
        const int t = tid & 127;                         // thread index inside the group
        const int row = q * 32 + lane;                   // TMEM lane = position (N-form) or channel (T-form)
        const uint32_t acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + s * 128;
        const uint32_t slot = sb + OFF_SLOT + s * SLOT_BYTES;
        uint32_t ph_acc = 0;
        // layer 0: this thread computes channels 8*c8 .. 8*c8+7 of positions (t >> 2) + 32 i, i = 0..3
        const int c8 = t & 3, p0 = t >> 2;
        const float4 *w0s = reinterpret_cast<const float4 *>(smem + OFF_W0) + c8 * 8;   // re-read per tile: 8 LDS.128, no conflicts
        float px[4], py[4], pz[4];
        auto load_xyz = [&](long long tile) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float *src = prm.xyz + (tile * P + p0 + 32 * i) * prm.ld;
                px[i] = __ldg(src);
                py[i] = __ldg(src + 1);
                pz[i] = __ldg(src + 2);
            }
        };
        if (2 * blockIdx.x + s < n_tiles) load_xyz(2 * blockIdx.x + s);
        float *out_f = prm.out_bf16 ? nullptr : static_cast<float *>(prm.out);
        __nv_bfloat16 *out_h = prm.out_bf16 ? static_cast<__nv_bfloat16 *>(prm.out) : nullptr;
        const bool timer = TIMED && t == 0 && s == 0;
        for (long long tile = 2 * blockIdx.x + s; tile < n_tiles; tile += tstride) {
            if (timer) SA_TICK(0);
            // ---- layer 0 (3 -> 32, fp32, ReLU) -> X1 ----
            float4 w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = w0s[i];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int p = p0 + 32 * i;
                uint32_t pk[4];
#pragma unroll
                for (int hh = 0; hh < 4; ++hh) {
                    const float4 wa = w[2 * hh], wb = w[2 * hh + 1];
                    const float a = fmaf(wa.z, pz[i], fmaf(wa.y, py[i], fmaf(wa.x, px[i], wa.w)));
                    const float b = fmaf(wb.z, pz[i], fmaf(wb.y, py[i], fmaf(wb.x, px[i], wb.w)));
                    pk[hh] = pack_relu_bf16x2(a, b);
                }
                st_shared_v4(slot + SL_X1 + p * 128 + ((c8 ^ (p & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
            }
            fence_async_smem();
            tc_fence_before();   // this warp's reads of the previous tile's accumulator are complete as well
            __syncwarp();
            if (lane == 0) mbar_arrive1(bar_act + 8 * s);
            if (timer) SA_TICK(0);
            if (tile + tstride < n_tiles) load_xyz(tile + tstride);   // in flight for the rest of the tile
            // ---- layer 1 epilogue: 64 channels of my position -> X2 (one 128-byte row of the slab) ----
            mbar_wait(bar_acc + 8 * s, ph_acc);
            ph_acc ^= 1u;
            tc_fence_after();
            if (timer) SA_TICK(0);
            {
                uint32_t v0[32], v1[32];
                tmem_ld32_issue(acc, v0);
                tmem_ld32_issue(acc + 32, v1);
                tmem_ld32_wait(v0);
                store_row_chunks<32, true>(slot + SL_X2, row, 0, v0);
                tmem_ld32_wait(v1);
                store_row_chunks<32, true>(slot + SL_X2, row, 32, v1);
            }
            if (timer) SA_TICK(0);
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive1(bar_act + 8 * s);
            if (timer) SA_TICK(0);
            // ---- layer 2 epilogue (T-form): lane = channel `row`, columns = positions; max over each run of 16 ----
            mbar_wait(bar_acc + 8 * s, ph_acc);
            ph_acc ^= 1u;
            tc_fence_after();
            if (timer) SA_TICK(0);
            const long long o = tile * 8 * 128 + row;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t v0[32], v1[32];
                tmem_ld32_issue(acc + j * 64, v0);
                tmem_ld32_issue(acc + j * 64 + 32, v1);
                tmem_ld32_wait(v0);
                tmem_ld32_wait(v1);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint32_t(&v)[32] = g < 2 ? v0 : v1;
                    const int b = (g & 1) * 16;
                    float mm = fmax3(__uint_as_float(v[b]), __uint_as_float(v[b + 1]), __uint_as_float(v[b + 2]));
#pragma unroll
                    for (int i = 3; i < 15; i += 2) mm = fmax3(mm, __uint_as_float(v[b + i]), __uint_as_float(v[b + i + 1]));
                    mm = fmax3(mm, __uint_as_float(v[b + 15]), 0.0f);  // the ReLU commutes with the max
                    if (out_h) out_h[o + (j * 4 + g) * 128] = __float2bfloat16_rn(mm); else out_f[o + (j * 4 + g) * 128] = mm;
                }
            }
            if (timer) SA_TICK(0);
        }
    }
    tmem_free(tmem_base, TMEM_COLS, warp, 8);
#undef SA_TICK
}

// ======================================================================================================================
// SA, second form: same slots (2 CTAs / SM x 2 tiles in flight), but the fp32 3 -> 32 layer of the NEXT tile is computed by
// the slot's MMA warp in the shadow of the current tile's epilogues.  In the first form the four epilogue warps of a slot run
// layer 0, epilogue 1 and epilogue 2 back to back, one warp per scheduler, and their per-warp instruction latency IS the tile
// time (adding 80 instructions to epilogue 1 cost 12 %; a fully decoupled one-CTA pipeline with dedicated producer /
// epilogue-1 / epilogue-2 warp groups was measured slower, 382 us vs 330 us: it has a single MMA issuer per SM).
// Moving layer 0 (46 % of the group's instructions) onto the otherwise idle MMA warp shortens the slot's critical path to
// MMA1 -> epilogue 1 -> MMA2 -> epilogue 2.
//   MMA warp of slot s, per tile:  wait "accumulator free" -> MMA1 (X1 -> acc) -> wait MMA1 done (X1 consumed) -> layer 0 of
//       the next tile -> X1, polling "X2 ready" every 4 channel pairs to issue MMA2 (X2 -> acc) as soon as it can
//   epilogue group of slot s, per tile:  wait MMA1 -> epilogue 1 (acc -> X2) -> signal -> wait MMA2 -> epilogue 2 (max over 16
//       neighbours -> out) -> signal "accumulator free"
// ======================================================================================================================
// IDX = 1: the kernel takes the patches themselves ([points, 3], prm.xyz) and the in-patch kNN table as bytes (prm.idx8) and
// forms the recentred neighbour coordinates on the fly -- the [points, 16, 3] fp32 tensor (100 MB per 32-cloud step, written by
// the kNN kernel and read straight back here) never exists.  Same fp32 subtraction, so the result is bit-identical.
// PCC_SA_TICKS (a build flag, never set in the shipped library): clock64() of CTA 0 / slot 0 -- epilogue thread 0 at dbg[0..],
// the MMA warp's lane 0 at dbg[2048..]; read by tools/time_sa_ticks.py
#ifdef PCC_SA_TICKS
#define SA_TICK(base) do { if (ticker && tick < 2000) prm.dbg[(base) + tick++] = clock64(); } while (0)
#else
#define SA_TICK(base) do { } while (0)
#endif
template <int IDX, int NS>
__global__ void __launch_bounds__(sa::Lay<NS>::THREADS, NS == 2 ? 2 : 1) sa_chain2_kernel(const __grid_constant__ SaParams prm) {
    using namespace sa;
    // NS accumulator slots per CTA: 2 (two CTAs per SM) or 4 (one CTA per SM, whose four MMA warps land on four different
    // SM sub-partitions -- with 2 + 2 the MMA warps of both CTAs share sub-partitions 0 and 1)
    using L = Lay<NS>;
    constexpr int OFF_BAR = L::OFF_BAR, OFF_W0 = L::OFF_W0, OFF_XYZ = L::OFF_XYZ, OFF_BAR2 = L::OFF_BAR2, THREADS = L::THREADS;
    constexpr int TMEM_COLS = L::TMEM_COLS, EW = 4 * NS;   // EW: epilogue warps
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t sb = smem_u32(smem);
    const int tid = threadIdx.x, warp = __shfl_sync(FULL_MASK, tid >> 5, 0), lane = tid & 31;
    const uint32_t bar_acc = sb + OFF_BAR, bar_act = sb + OFF_BAR + 8 * NS, bar_x1 = sb + OFF_BAR + 16 * NS;  // [NS] each
    const uint32_t bar_xyz = sb + OFF_BAR2;                                                           // [NS slots][2 buffers] (IDX)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 24 * NS);

    copy_to_smem(smem + OFF_W1, prm.w1p, 64 * KP1 * 2, tid, THREADS);
    copy_to_smem(smem + OFF_W2, prm.w2p, 128 * KP2 * 2, tid, THREADS);
    fill_ones_block(smem + OFF_ONES, tid, THREADS);
    // layer-0 weights as channel PAIRS for the packed fp32 FMA: pair j = channels (2j, 2j + 1) -> {wx wx' | wy wy' | wz wz' | b b'}
    for (int e = tid; e < 32 * 4; e += THREADS) {
        const int ch = 2 * (e >> 3) + (e & 1), comp = (e >> 1) & 3;
        reinterpret_cast<float *>(smem + OFF_W0)[e] = comp < 3 ? prm.w0[ch * 3 + comp] : prm.b0[ch];
    }
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(bar_acc + 8 * s, 1);
            mbar_init(bar_act + 8 * s, 4);
            mbar_init(bar_x1 + 8 * s, 1);
        }
        if constexpr (IDX) {
            for (int i = 0; i < 2 * NS; ++i) mbar_init(bar_xyz + 8 * i, 4);
        }
    }
    const uint32_t tmem_base = tmem_alloc_and_sync(tmem_slot, TMEM_COLS, warp, EW);
    const int n_tiles = prm.n_tiles;
    const int tstride = NS * gridDim.x;
#ifdef PCC_SA_TICKS
    const bool ticker = prm.dbg && blockIdx.x == 0 && (tid == 0 || tid == 32 * EW);   // slot 0: epilogue thread 0, MMA lane 0
    int tick = 0, tick2 = 0;
#endif

    if (warp >= EW) {
        // ---- MMA warp of slot s: issues layer 1 and computes layer 0 of the next tile (layer 2 is issued by the slot's epilogue
        // group the moment X2 is complete: polling for that from inside layer 0 cost this warp ~150 clocks per poll) ----
        const int s = warp - EW;
        const uint32_t acc = __shfl_sync(FULL_MASK, tmem_base, 0) + s * 128;
        const uint32_t id64 = umma_idesc(128, 64);
        const uint64_t d_ones = desc_k16(sb + OFF_ONES);
        const uint64_t d_w1 = desc_w(sb + OFF_W1, KP1, 0);
        const uint32_t x1 = sb + OFF_SLOT + s * SLOT_BYTES + SL_X1;
        const uint64_t d_x1 = umma_desc_sw128(x1);
        const uint32_t w0s = sb + OFF_W0;   // (explicit ld.shared below: through the re-aligned generic pointer these were LD.E)
        float px[4], py[4], pz[4];   // lane owns positions lane + 32 i of the tile
        auto load_xyz = [&](long long tile) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float *src = prm.xyz + (tile * P + lane + 32 * i) * prm.ld;
                px[i] = __ldg(src);
                py[i] = __ldg(src + 1);
                pz[i] = __ldg(src + 2);
            }
        };
        // IDX: the slot's epilogue group gathers and recentres the next tiles' neighbours into small shared buffers (k-th tile
        // of the slot -> buffer k & 1, its barrier completes once per two tiles); this warp only reads 12 floats per lane
        uint32_t kx = 0;
        auto take_xyz = [&]() {
            mbar_wait(bar_xyz + 8 * (2 * s + (kx & 1u)), (kx >> 1) & 1u);
            // (plain generic loads on purpose: with ld.shared here and st.shared in finish_gather the kernel measured 247 us against
            // 224 us -- this warp has slack, and its layer 0 then starts later, out of the way of the slot's layer-1 epilogue)
            const float *xg = reinterpret_cast<const float *>(smem + OFF_XYZ + (2 * s + (kx & 1u)) * XYZ_BUF);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                px[i] = xg[lane + 32 * i];
                py[i] = xg[P + lane + 32 * i];
                pz[i] = xg[2 * P + lane + 32 * i];
            }
            ++kx;
        };
        uint32_t ph_act = 0, ph_x1 = 0;
        // layer 0 (3 -> 32, fp32) of the lane's four positions, two channels per instruction (fma.rn.f32x2: each half is the same
        // IEEE fma as the scalar form, so the result is bit-identical), one 16-byte chunk (8 channels) of X1 per position at a time
        auto layer0_tile = [&]() {
            uint64_t qx[4], qy[4], qz[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                qx[i] = dup_f32x2(px[i]);
                qy[i] = dup_f32x2(py[i]);
                qz[i] = dup_f32x2(pz[i]);
            }
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                uint32_t pk[4][4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const ulonglong2 wxy = ld_shared_v2u64(w0s + 32 * (4 * c8 + j)), wzb = ld_shared_v2u64(w0s + 32 * (4 * c8 + j) + 16);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        pk[i][j] = pack_relu_bf16x2_pair(fma_f32x2(wzb.x, qz[i], fma_f32x2(wxy.y, qy[i], fma_f32x2(wxy.x, qx[i], wzb.y))));
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int p = lane + 32 * i;
                    st_shared_v4(x1 + p * 128 + ((c8 ^ (p & 7)) << 4), pk[i][0], pk[i][1], pk[i][2], pk[i][3]);
                }
            }
        };
        long long tile = static_cast<long long>(NS) * blockIdx.x + s;
        if (tile < n_tiles) {
            if constexpr (IDX) take_xyz(); else load_xyz(tile);
            layer0_tile();
            fence_async_smem();
            if constexpr (!IDX) { if (tile + tstride < n_tiles) load_xyz(tile + tstride); }
        }
        for (; tile < n_tiles; tile += tstride) {
            SA_TICK(2048);   // 0 top
            const bool more = tile + tstride < n_tiles;
            mbar_wait(bar_act + 8 * s, ph_act);   // accumulator columns 0..63 are free (epilogue 2 of the previous tile has read them)
            ph_act ^= 1u;
            SA_TICK(2048);   // 1 accumulator free
            tc_fence_after();
            if (elect_one()) {
                umma_bf16(acc, d_x1, d_w1, id64, 0u);
                umma_bf16(acc, d_x1 + 2, d_w1 + 16, id64, 1u);
                umma_bf16(acc, d_ones, d_w1 + 32, id64, 1u);
                umma_commit(bar_x1 + 8 * s);
                umma_commit(bar_acc + 8 * s);
            }
            __syncwarp();
            SA_TICK(2048);   // 2 layer 1 issued
            if constexpr (IDX) { if (more) take_xyz(); }   // (layer 0 of this tile is done with px/py/pz: the loads land under the wait)
            SA_TICK(2048);   // 3 xyz taken
            // layer 1 has consumed X1: the next tile's layer 0 may overwrite it.  (Its own barrier, one phase per tile: waiting
            // for the even phases of bar_acc instead deadlocks if this warp is ever delayed past layer 2's completion, which
            // flips bar_acc's parity back -- seen as a rare launch failure under the 4-stream replay)
            mbar_wait(bar_x1 + 8 * s, ph_x1);
            ph_x1 ^= 1u;
            SA_TICK(2048);   // 4 X1 consumed
            if (more) {
                tc_fence_after();
                layer0_tile();
                fence_async_smem();
                if constexpr (!IDX) { if (tile + 2 * tstride < n_tiles) load_xyz(tile + 2 * tstride); }
            }
            SA_TICK(2048);   // 5 layer 0 of the next tile done
        }
    } else {
        // ---- epilogue group g = warp / 4 serves slot g; warp q = warp % 4 owns TMEM lanes 32q.. ----
        const int s = warp >> 2, q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + s * 128;
        const uint32_t slot = sb + OFF_SLOT + s * SLOT_BYTES;
        uint32_t ph_acc = 0;
        // layer 2 is issued from here: warp (s & 3) of the group (a different SM sub-partition per slot), one elected lane
        const uint32_t id128 = umma_idesc(128, 128);
        const uint64_t d_ones = desc_k16(sb + OFF_ONES), d_w2 = desc_w(sb + OFF_W2, KP2, 0), d_x2 = umma_desc_sw128(slot + SL_X2);
        const uint32_t acc_slot = tmem_base + s * 128;
        const float bias2 = prm.b2 ? __ldg(prm.b2 + row) : 0.0f;
        float *out_f = prm.out_bf16 ? nullptr : static_cast<float *>(prm.out);
        __nv_bfloat16 *out_h = prm.out_bf16 ? static_cast<__nv_bfloat16 *>(prm.out) : nullptr;
        if (lane == 0) mbar_arrive1(bar_act + 8 * s);   // the accumulator starts out free
        // IDX: thread t of the group owns position t of a tile = neighbour (t & 15) of point 8 * tile + (t >> 4); it gathers
        // that neighbour from the point's own patch and recentres it (pn_kit.py:190-191), two tiles ahead of the epilogues
        const int t = row;
        unsigned nbyte = 0u;                 // neighbour byte of the tile whose gather is issued next
        float gc[3], gn[3];                  // centre / neighbour coordinates in flight
        uint32_t kg = 0;                     // tiles of this slot gathered so far
        auto load_byte = [&](long long tl) { nbyte = __ldg(prm.idx8 + tl * P + t); };
        auto issue_gather = [&](long long tl) {
            const unsigned g0 = static_cast<unsigned>(tl) * 8u;      // a tile is 8 points x 16 neighbours; P % 8 == 0: one patch
            const unsigned pbase = prm.pts_shift >= 0 ? (g0 >> prm.pts_shift) << prm.pts_shift
                                                      : g0 / static_cast<unsigned>(prm.pts_per_patch) * static_cast<unsigned>(prm.pts_per_patch);
            const float *c = prm.xyz + static_cast<size_t>(g0 + (t >> 4)) * 3;
            const float *n = prm.xyz + static_cast<size_t>(pbase + nbyte) * 3;
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                gc[e] = __ldg(c + e);
                gn[e] = __ldg(n + e);
            }
        };
        auto finish_gather = [&]() {
            float *xg = reinterpret_cast<float *>(smem + OFF_XYZ + (2 * s + (kg & 1u)) * XYZ_BUF);
#pragma unroll
            for (int e = 0; e < 3; ++e) xg[e * P + t] = __fsub_rn(gn[e], gc[e]);
            __syncwarp();
            if (lane == 0) mbar_arrive1(bar_xyz + 8 * (2 * s + (kg & 1u)));
            ++kg;
        };
        if constexpr (IDX) {
            const long long t0 = static_cast<long long>(NS) * blockIdx.x + s;
            if (t0 < n_tiles) {
                load_byte(t0);
                issue_gather(t0);
                if (t0 + tstride < n_tiles) load_byte(t0 + tstride);
                finish_gather();
                if (t0 + tstride < n_tiles) {
                    issue_gather(t0 + tstride);
                    if (t0 + 2 * tstride < n_tiles) load_byte(t0 + 2 * tstride);
                    finish_gather();
                }
            }
        }
        for (long long tile = static_cast<long long>(NS) * blockIdx.x + s; tile < n_tiles; tile += tstride) {
            SA_TICK(0);   // 0 top
            if constexpr (IDX) {            // the tile after next: loads in flight under the wait for layer 1 and its epilogue
                if (tile + 2 * tstride < n_tiles) issue_gather(tile + 2 * tstride);
                if (tile + 3 * tstride < n_tiles) load_byte(tile + 3 * tstride);
            }
            // ---- layer 1 epilogue: 64 channels of my position -> X2 (one 128-byte row of the slab) ----
            mbar_wait(bar_acc + 8 * s, ph_acc);
            ph_acc ^= 1u;
            tc_fence_after();
            SA_TICK(0);   // 1 layer 1 accumulator ready
            {
                uint32_t v0[32], v1[32];
                tmem_ld32_issue(acc, v0);
                tmem_ld32_issue(acc + 32, v1);
                tmem_ld32_wait(v0);
                store_row_chunks<32, true>(slot + SL_X2, row, 0, v0);
                tmem_ld32_wait(v1);
                store_row_chunks<32, true>(slot + SL_X2, row, 32, v1);
            }
            SA_TICK(0);   // 2 X2 stored
            fence_async_smem();
            tc_fence_before();
            // the four warps of the group: X2 is complete (and layer 1's accumulator read); literal ids, so that ptxas
            // reserves NS + 1 hardware barriers and not all 16
            if (s == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
            else if (s == 1) asm volatile("bar.sync 2, 128;" ::: "memory");
            else if (s == 2) asm volatile("bar.sync 3, 128;" ::: "memory");
            else asm volatile("bar.sync 4, 128;" ::: "memory");
            if (q == (s & 3)) {
#ifdef PCC_SA_TICKS
                if (ticker && tick2 < 500) prm.dbg[4096 + tick2++] = clock64();
#endif
                tc_fence_after();
                if (elect_one()) {  // T-form: D^T[128 channels, 128 positions]
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_bf16(acc_slot, d_w2 + 16 * ks, d_x2 + 2 * ks, id128, ks > 0);
                    if (!prm.b2) umma_bf16(acc_slot, d_w2 + 64, d_ones, id128, 1u);
                    umma_commit(bar_acc + 8 * s);
                }
                __syncwarp();
            }
            SA_TICK(0);   // 3 layer 2 issued
            if constexpr (IDX) {            // (late: the loads issued at the top of the iteration have landed by now)
                if (tile + 2 * tstride < n_tiles) finish_gather();
            }
            SA_TICK(0);   // 4 gather finished
            // ---- layer 2 epilogue (T-form): lane = channel `row`, columns = positions; max over each run of 16 ----
            mbar_wait(bar_acc + 8 * s, ph_acc);
            ph_acc ^= 1u;
            tc_fence_after();
            SA_TICK(0);   // 5 layer 2 accumulator ready
            const long long o = tile * 8 * 128 + row;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t v0[32], v1[32];
                tmem_ld32_issue(acc + j * 64, v0);
                tmem_ld32_issue(acc + j * 64 + 32, v1);
                tmem_ld32_wait(v0);
                tmem_ld32_wait(v1);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint32_t(&v)[32] = g < 2 ? v0 : v1;
                    const int b = (g & 1) * 16;
                    float mm = fmax3(__uint_as_float(v[b]), __uint_as_float(v[b + 1]), __uint_as_float(v[b + 2]));
#pragma unroll
                    for (int i = 3; i < 15; i += 2) mm = fmax3(mm, __uint_as_float(v[b + i]), __uint_as_float(v[b + i + 1]));
                    // the bias (one per channel = per lane) and the ReLU commute with the max over the 16 neighbours
                    mm = fmaxf(fmaxf(mm, __uint_as_float(v[b + 15])) + bias2, 0.0f);
                    if (out_h) out_h[o + (j * 4 + g) * 128] = __float2bfloat16_rn(mm); else out_f[o + (j * 4 + g) * 128] = mm;
                }
            }
            SA_TICK(0);   // 6 pooled rows stored
            tc_fence_before();   // (the reads of columns 64..127 are ordered before the group's next bar.sync -> layer 2 of the next tile)
            __syncwarp();
            if (lane == 0) mbar_arrive1(bar_act + 8 * s);   // accumulator free
        }
    }
    tmem_free(tmem_base, TMEM_COLS, warp, EW);
}

// ======================================================================================================================
// PNF: [feat[M,128] bf16 | xyz[M,3] fp32] -> 128 -> 256 (ReLU both) -> out[M,256] bf16
// ======================================================================================================================
struct PnfParams {
    const float *xyz;
    long long ld;
    const void *w0p, *w1p;     // packed [128 x 144], [256 x 144]
    int n_tiles;
};

namespace pnf {
constexpr int KP = 144;
constexpr int OFF_W0 = 0;
constexpr int OFF_W1 = OFF_W0 + 128 * KP * 2;        // 36864
constexpr int OFF_ONES = OFF_W1 + 256 * KP * 2;      // 110592
constexpr int OFF_XIN = OFF_ONES + K16;              // 114688
constexpr int OFF_F = OFF_XIN + K16;                 // 118784 = 116 * 1024
constexpr int OFF_X1 = OFF_F + 2 * SLAB;             // 151552   (X1 slabs 0,1; with the two slabs behind them: output staging 0..3)
constexpr int OFF_BAR = OFF_X1 + 4 * SLAB;           // 217088
constexpr int SMEM = OFF_BAR + 64 + 1024;
constexpr int THREADS = 320;
constexpr int TMEM_COLS = 256;
static_assert(OFF_F % 1024 == 0 && OFF_X1 % 1024 == 0, "slab alignment");
}  // namespace pnf

__global__ void __launch_bounds__(pnf::THREADS, 1)
pnf_chain_kernel(const __grid_constant__ PnfParams prm, const __grid_constant__ CUtensorMap tm_feat,
                 const __grid_constant__ CUtensorMap tm_out) {
    using namespace pnf;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t sb = smem_u32(smem);
    const int tid = threadIdx.x, warp = __shfl_sync(FULL_MASK, tid >> 5, 0), lane = tid & 31;
    const uint32_t bar_in = sb + OFF_BAR, bar_empty = sb + OFF_BAR + 8, bar_acc = sb + OFF_BAR + 16, bar_act = sb + OFF_BAR + 24;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 48);

    copy_to_smem(smem + OFF_W0, prm.w0p, 128 * KP * 2, tid, THREADS);
    copy_to_smem(smem + OFF_W1, prm.w1p, 256 * KP * 2, tid, THREADS);
    fill_ones_block(smem + OFF_ONES, tid, THREADS);
    if (tid == 0) {
        mbar_init(bar_in, 9);      // 8 epilogue warps (xyz block) + the producer's expect_tx arrival
        mbar_init(bar_empty, 1);
        mbar_init(bar_acc, 1);
        mbar_init(bar_act, 8);
    }
    const uint32_t tmem_base = tmem_alloc_and_sync(tmem_slot, TMEM_COLS, warp, 8);
    const int n_tiles = prm.n_tiles;

    if (warp == 9) {
        if (lane == 0) {  // ---- TMA producer: the two feature slabs of each tile ----
            uint32_t ph = 1;  // a fresh barrier passes a parity-1 wait: the buffer starts empty
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(bar_empty, ph);
                ph ^= 1u;
                mbar_arrive_expect_tx(bar_in, 2 * SLAB);
                tma_load_2d(sb + OFF_F, &tm_feat, 0, tile * P, bar_in);
                tma_load_2d(sb + OFF_F + SLAB, &tm_feat, 64, tile * P, bar_in);
            }
        }
    } else if (warp == 8) {
        {   // ---- MMA issuer: warp-uniform control flow, one elected lane issues ----
            const uint32_t tb = __shfl_sync(FULL_MASK, tmem_base, 0);
            uint32_t ph_in = 0, ph_act = 0;
            const uint32_t id128 = umma_idesc(128, 128), id256 = umma_idesc(128, 256);
            const uint64_t d_f = umma_desc_sw128(sb + OFF_F), d_x1 = umma_desc_sw128(sb + OFF_X1);   // slab j: + j * (SLAB >> 4)
            const uint64_t d_w0 = desc_w(sb + OFF_W0, KP, 0), d_w1 = desc_w(sb + OFF_W1, KP, 0);
            const uint64_t d_xin = desc_k16(sb + OFF_XIN), d_ones = desc_k16(sb + OFF_ONES);
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(bar_in, ph_in);
                ph_in ^= 1u;
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        umma_bf16(tb, d_f + (ks >> 2) * (SLAB >> 4) + (ks & 3) * 2, d_w0 + 16 * ks, id128, ks > 0);
                    umma_bf16(tb, d_xin, d_w0 + 16 * 8, id128, 1u);
                    umma_commit(bar_acc);
                    umma_commit(bar_empty);   // the feature slabs may be refilled
                }
                __syncwarp();
                mbar_wait(bar_act, ph_act);
                ph_act ^= 1u;
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        umma_bf16(tb, d_x1 + (ks >> 2) * (SLAB >> 4) + (ks & 3) * 2, d_w1 + 16 * ks, id256, ks > 0);
                    umma_bf16(tb, d_ones, d_w1 + 16 * 8, id256, 1u);
                    umma_commit(bar_acc);
                }
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3, h = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        uint32_t ph_acc = 0;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (tid < P && blockIdx.x < n_tiles) {
            const float *src = prm.xyz + (static_cast<long long>(blockIdx.x) * P + tid) * prm.ld;
            px = __ldg(src);
            py = __ldg(src + 1);
            pz = __ldg(src + 2);
        }
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            if (tid < P) {  // K block of the 9th step: (x y z 1 0 ...), the 1 multiplies the bias column (k = 131)
                const uint32_t w[8] = {bf16_bits(px) | (bf16_bits(py) << 16), bf16_bits(pz) | (ONE_BF16 << 16), 0u, 0u, 0u, 0u, 0u, 0u};
                store_k16_row(sb + OFF_XIN, tid, w);
                const long long nt = static_cast<long long>(tile) + gridDim.x;
                if (nt < n_tiles) {
                    const float *src = prm.xyz + (nt * P + tid) * prm.ld;
                    px = __ldg(src);
                    py = __ldg(src + 1);
                    pz = __ldg(src + 2);
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive1(bar_in);
            // the previous tile's TMA store must have finished reading the staging slabs (which alias X1)
            if (tid == 0) tma_store_wait_read0();
            named_bar_sync(1, EPI_THREADS);
            // ---- layer 0 epilogue: 64 channels per warp -> X1 slab h ----
            mbar_wait(bar_acc, ph_acc);
            ph_acc ^= 1u;
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t v[32];
                tmem_ld32(lane_base + h * 64 + j * 32, v);
                store_row_chunks<32, true>(sb + OFF_X1, row, h * 64 + j * 32, v);
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive1(bar_act);
            // ---- layer 1 epilogue: 128 channels per warp -> staging slabs 2h, 2h+1 -> TMA store ----
            mbar_wait(bar_acc, ph_acc);
            ph_acc ^= 1u;
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t v[32];
                tmem_ld32(lane_base + h * 128 + j * 32, v);
                store_row_chunks<32, true>(sb + OFF_X1, row, h * 128 + j * 32, v);
            }
            fence_async_smem();
            tc_fence_before();
            named_bar_sync(1, EPI_THREADS);
            if (tid == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) tma_store_2d(&tm_out, j * 64, tile * P, sb + OFF_X1 + j * SLAB);
                tma_store_commit();
            }
        }
        if (tid == 0) tma_store_wait_all0();
    }
    tmem_free(tmem_base, TMEM_COLS, warp, 8);
}

// ======================================================================================================================
// DEC: [lin[M,128] bf16 | latent[M/128,16] fp32] -> 128 -> 64 -> 32 (ReLU) -> 3 -> out[M,3] fp32
// ======================================================================================================================
struct DecParams {
    const float *lat;          // [n_tiles, ld_lat] fp32: one latent per tile (row_div == 128)
    long long ld_lat;
    const void *w0p, *w1p, *w2p, *w3p;   // packed [128 x 160], [64 x 144], [32 x 80], [16 x 48]
    float *out;                // [M, 3]
    int n_tiles;
};

namespace dec {
constexpr int KP0 = 160, KP1 = 144, KP2 = 80, KP3 = 48;
constexpr int OFF_W0 = 0;
constexpr int OFF_W1 = OFF_W0 + 128 * KP0 * 2;       // 40960
constexpr int OFF_W2 = OFF_W1 + 64 * KP1 * 2;        // 59392
constexpr int OFF_W3 = OFF_W2 + 32 * KP2 * 2;        // 64512
constexpr int OFF_ONES = 66560;                      // 65 * 1024
constexpr int OFF_XLAT = OFF_ONES + K16;
// (a second input buffer -- the TMA load of tile t + 1 under the whole chain of tile t instead of under its layers 1..3 -- was
// measured: 45 us either way; the tile's four MMA -> epilogue hand-offs, ~1.5 k clocks each with one tile in flight per SM, are
// what the kernel takes: issue 9 %, tensor 18 %, DRAM 18 %)
constexpr int OFF_LIN = OFF_XLAT + K16;              // 74752 = 73 * 1024
constexpr int OFF_X1 = OFF_LIN + 2 * SLAB;
constexpr int OFF_X2 = OFF_X1 + 2 * SLAB;
constexpr int OFF_X3 = OFF_X2 + SLAB;
constexpr int OFF_BAR = OFF_X3 + SLAB;               // 173056
constexpr int SMEM = OFF_BAR + 64 + 1024;
constexpr int THREADS = 320;
constexpr int TMEM_COLS = 128;
static_assert(OFF_W3 + 16 * KP3 * 2 <= OFF_ONES && OFF_LIN % 1024 == 0, "layout");
}  // namespace dec

__global__ void __launch_bounds__(dec::THREADS, 1)
dec_chain_kernel(const __grid_constant__ DecParams prm, const __grid_constant__ CUtensorMap tm_lin) {
    using namespace dec;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t sb = smem_u32(smem);
    const int tid = threadIdx.x, warp = __shfl_sync(FULL_MASK, tid >> 5, 0), lane = tid & 31;
    const uint32_t bar_in = sb + OFF_BAR, bar_empty = sb + OFF_BAR + 8, bar_acc = sb + OFF_BAR + 16, bar_act = sb + OFF_BAR + 24;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 48);

    copy_to_smem(smem + OFF_W0, prm.w0p, 128 * KP0 * 2, tid, THREADS);
    copy_to_smem(smem + OFF_W1, prm.w1p, 64 * KP1 * 2, tid, THREADS);
    copy_to_smem(smem + OFF_W2, prm.w2p, 32 * KP2 * 2, tid, THREADS);
    copy_to_smem(smem + OFF_W3, prm.w3p, 16 * KP3 * 2, tid, THREADS);
    fill_ones_block(smem + OFF_ONES, tid, THREADS);
    if (tid == 0) {
        mbar_init(bar_in, 9);
        mbar_init(bar_empty, 1);
        mbar_init(bar_acc, 1);
        mbar_init(bar_act, 8);
    }
    const uint32_t tmem_base = tmem_alloc_and_sync(tmem_slot, TMEM_COLS, warp, 8);
    const int n_tiles = prm.n_tiles;

    if (warp == 9) {
        if (lane == 0) {
            uint32_t ph = 1;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(bar_empty, ph);
                ph ^= 1u;
                mbar_arrive_expect_tx(bar_in, 2 * SLAB);
                tma_load_2d(sb + OFF_LIN, &tm_lin, 0, tile * P, bar_in);
                tma_load_2d(sb + OFF_LIN + SLAB, &tm_lin, 64, tile * P, bar_in);
            }
        }
    } else if (warp == 8) {
        {   // ---- MMA issuer: warp-uniform control flow, one elected lane issues ----
            const uint32_t tb = __shfl_sync(FULL_MASK, tmem_base, 0);
            uint32_t ph_in = 0, ph_act = 0;
            const uint32_t id128 = umma_idesc(128, 128), id64 = umma_idesc(128, 64), id32 = umma_idesc(128, 32), id16 = umma_idesc(128, 16);
            const uint64_t d_lin = umma_desc_sw128(sb + OFF_LIN), d_x1 = umma_desc_sw128(sb + OFF_X1), d_x2 = umma_desc_sw128(sb + OFF_X2),
                           d_x3 = umma_desc_sw128(sb + OFF_X3);
            const uint64_t d_w0 = desc_w(sb + OFF_W0, KP0, 0), d_w1 = desc_w(sb + OFF_W1, KP1, 0), d_w2 = desc_w(sb + OFF_W2, KP2, 0),
                           d_w3 = desc_w(sb + OFF_W3, KP3, 0);
            const uint64_t d_xlat = desc_k16(sb + OFF_XLAT), d_ones = desc_k16(sb + OFF_ONES);
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(bar_in, ph_in);
                ph_in ^= 1u;
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        umma_bf16(tb, d_lin + (ks >> 2) * (SLAB >> 4) + (ks & 3) * 2, d_w0 + 16 * ks, id128, ks > 0);
                    umma_bf16(tb, d_xlat, d_w0 + 16 * 8, id128, 1u);
                    umma_bf16(tb, d_ones, d_w0 + 16 * 9, id128, 1u);
                    umma_commit(bar_acc);
                    umma_commit(bar_empty);
                }
                __syncwarp();
                // layer 1: 128 -> 64
                mbar_wait(bar_act, ph_act);
                ph_act ^= 1u;
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        umma_bf16(tb, d_x1 + (ks >> 2) * (SLAB >> 4) + (ks & 3) * 2, d_w1 + 16 * ks, id64, ks > 0);
                    umma_bf16(tb, d_ones, d_w1 + 16 * 8, id64, 1u);
                    umma_commit(bar_acc);
                }
                __syncwarp();
                // layer 2: 64 -> 32
                mbar_wait(bar_act, ph_act);
                ph_act ^= 1u;
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_bf16(tb, d_x2 + ks * 2, d_w2 + 16 * ks, id32, ks > 0);
                    umma_bf16(tb, d_ones, d_w2 + 16 * 4, id32, 1u);
                    umma_commit(bar_acc);
                }
                __syncwarp();
                // layer 3: 32 -> 3 (N = 16)
                mbar_wait(bar_act, ph_act);
                ph_act ^= 1u;
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) umma_bf16(tb, d_x3 + ks * 2, d_w3 + 16 * ks, id16, ks > 0);
                    umma_bf16(tb, d_ones, d_w3 + 16 * 2, id16, 1u);
                    umma_commit(bar_acc);
                }
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3, h = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        uint32_t ph_acc = 0;
        auto arrive_act = [&]() {
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive1(bar_act);
        };
        auto wait_acc = [&]() {
            mbar_wait(bar_acc, ph_acc);
            ph_acc ^= 1u;
            tc_fence_after();
        };
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            if (tid < P) {  // every row of the tile carries the same latent (AE.py:51 tiles it over the k points)
                const float4 *lp = reinterpret_cast<const float4 *>(prm.lat + static_cast<long long>(tile) * prm.ld_lat);
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 f = __ldg(lp + i);
                    w[2 * i] = bf16_bits(f.x) | (bf16_bits(f.y) << 16);
                    w[2 * i + 1] = bf16_bits(f.z) | (bf16_bits(f.w) << 16);
                }
                store_k16_row(sb + OFF_XLAT, tid, w);
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive1(bar_in);
            wait_acc();   // layer 0: 64 channels per warp -> X1 slab h
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t v[32];
                tmem_ld32(lane_base + h * 64 + j * 32, v);
                store_row_chunks<32, true>(sb + OFF_X1, row, h * 64 + j * 32, v);
            }
            arrive_act();
            wait_acc();   // layer 1: 32 channels per warp
            {
                uint32_t v[32];
                tmem_ld32(lane_base + h * 32, v);
                store_row_chunks<32, true>(sb + OFF_X2, row, h * 32, v);
            }
            arrive_act();
            wait_acc();   // layer 2: 16 channels per warp
            {
                uint32_t v[16];
                tmem_ld16(lane_base + h * 16, v);
                store_row_chunks<16, true>(sb + OFF_X3, row, h * 16, v);
            }
            arrive_act();
            wait_acc();   // layer 3: x, y, z of this position (no ReLU)
            if (h == 0) {
                uint32_t v[4];
                tmem_ld4(lane_base, v);
                float *o = prm.out + (static_cast<long long>(tile) * P + row) * 3;
                o[0] = __uint_as_float(v[0]);
                o[1] = __uint_as_float(v[1]);
                o[2] = __uint_as_float(v[2]);
            }
            tc_fence_before();
        }
    }
    tmem_free(tmem_base, TMEM_COLS, warp, 8);
}

// ---- DEC with two tiles in flight per SM -------------------------------------------------------------------------------------
// The one-tile kernel above is a chain of four MMA -> epilogue hand-offs per tile with nothing else to run on the SM (issue 9 %,
// tensor 18 %, DRAM 18 %).  Here a CTA carries TWO independent chains ("slots"): per slot 8 epilogue warps, one MMA warp, one TMA
// producer warp, 128 TMEM columns and its own barriers, so one slot's epilogue runs under the other's MMAs.  To fit twice in
// shared memory a slot keeps only [latent block | lin | X1]: X2 is written over X1's first slab and X3 over its second -- each
// only after the MMAs that read the bytes underneath have completed (the epilogue that writes waits for exactly that commit).
namespace dec2 {
using dec::KP0;
using dec::KP1;
using dec::KP2;
using dec::KP3;
using dec::OFF_ONES;
using dec::OFF_W0;
using dec::OFF_W1;
using dec::OFF_W2;
using dec::OFF_W3;
constexpr int SLOTS = 2;
constexpr int S_XLAT = 0, S_LIN = K16, S_X1 = S_LIN + 2 * SLAB;     // X2 = S_X1, X3 = S_X1 + SLAB
constexpr int SLOT_BYTES = S_X1 + 2 * SLAB;                          // 69632
constexpr int OFF_SLOT = OFF_ONES + K16;                             // 70656 = 69 * 1024
constexpr int OFF_BAR = OFF_SLOT + SLOTS * SLOT_BYTES;               // 209920
constexpr int SMEM = OFF_BAR + 128 + 1024;
constexpr int EPI_WARPS = 8;
constexpr int MMA_WARP0 = SLOTS * EPI_WARPS, TMA_WARP0 = MMA_WARP0 + SLOTS;
constexpr int THREADS = (TMA_WARP0 + SLOTS) * 32;                    // 640
constexpr int TMEM_COLS = 128 * SLOTS;
static_assert(OFF_SLOT % 1024 == 0 && SLOT_BYTES % 1024 == 0 && S_LIN % 1024 == 0 && S_X1 % 1024 == 0, "swizzled slabs are 1 KB aligned");
}  // namespace dec2

__global__ void __launch_bounds__(dec2::THREADS, 1)
dec_chain2_kernel(const __grid_constant__ DecParams prm, const __grid_constant__ CUtensorMap tm_lin) {
    using namespace dec2;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t sb = smem_u32(smem);
    const int tid = threadIdx.x, warp = __shfl_sync(FULL_MASK, tid >> 5, 0), lane = tid & 31;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 96);

    copy_to_smem(smem + OFF_W0, prm.w0p, 128 * KP0 * 2, tid, THREADS);
    copy_to_smem(smem + OFF_W1, prm.w1p, 64 * KP1 * 2, tid, THREADS);
    copy_to_smem(smem + OFF_W2, prm.w2p, 32 * KP2 * 2, tid, THREADS);
    copy_to_smem(smem + OFF_W3, prm.w3p, 16 * KP3 * 2, tid, THREADS);
    fill_ones_block(smem + OFF_ONES, tid, THREADS);
    if (tid == 0) {
#pragma unroll
        for (int sl = 0; sl < SLOTS; ++sl) {
            mbar_init(sb + OFF_BAR + 32 * sl, 9);         // in: 8 epilogue warps (latent block) + the producer's expect_tx
            mbar_init(sb + OFF_BAR + 32 * sl + 8, 1);     // empty: lin consumed
            mbar_init(sb + OFF_BAR + 32 * sl + 16, 1);    // acc: a layer's MMAs completed
            mbar_init(sb + OFF_BAR + 32 * sl + 24, 8);    // act: the next layer's input is in shared memory
        }
    }
    const uint32_t tmem_base = tmem_alloc_and_sync(tmem_slot, TMEM_COLS, warp, MMA_WARP0);
    const int n_tiles = prm.n_tiles;
    // role and slot of this warp; slot sl takes tiles blockIdx.x + gridDim.x * (2 j + sl)
    const int sl = warp < MMA_WARP0 ? warp / EPI_WARPS : (warp < TMA_WARP0 ? warp - MMA_WARP0 : warp - TMA_WARP0);
    const uint32_t ss = sb + OFF_SLOT + sl * SLOT_BYTES;
    const uint32_t bar_in = sb + OFF_BAR + 32 * sl, bar_empty = bar_in + 8, bar_acc = bar_in + 16, bar_act = bar_in + 24;
    const int tile0 = blockIdx.x + gridDim.x * sl, tstep = gridDim.x * SLOTS;

    if (warp >= TMA_WARP0) {
        if (lane == 0) {
            uint32_t ph = 1;
            for (int tile = tile0; tile < n_tiles; tile += tstep) {
                mbar_wait(bar_empty, ph);
                ph ^= 1u;
                mbar_arrive_expect_tx(bar_in, 2 * SLAB);
                tma_load_2d(ss + S_LIN, &tm_lin, 0, tile * P, bar_in);
                tma_load_2d(ss + S_LIN + SLAB, &tm_lin, 64, tile * P, bar_in);
            }
        }
    } else if (warp >= MMA_WARP0) {
        // ---- MMA issuer of slot sl: warp-uniform control flow, one elected lane issues ----
        const uint32_t tb = __shfl_sync(FULL_MASK, tmem_base, 0) + 128u * sl;
        uint32_t ph_in = 0, ph_act = 0;
        const uint32_t id128 = umma_idesc(128, 128), id64 = umma_idesc(128, 64), id32 = umma_idesc(128, 32), id16 = umma_idesc(128, 16);
        const uint64_t d_lin = umma_desc_sw128(ss + S_LIN), d_x1 = umma_desc_sw128(ss + S_X1), d_x2 = d_x1,
                       d_x3 = umma_desc_sw128(ss + S_X1 + SLAB);
        const uint64_t d_w0 = desc_w(sb + OFF_W0, KP0, 0), d_w1 = desc_w(sb + OFF_W1, KP1, 0), d_w2 = desc_w(sb + OFF_W2, KP2, 0),
                       d_w3 = desc_w(sb + OFF_W3, KP3, 0);
        const uint64_t d_xlat = desc_k16(ss + S_XLAT), d_ones = desc_k16(sb + OFF_ONES);
        for (int tile = tile0; tile < n_tiles; tile += tstep) {
            mbar_wait(bar_in, ph_in);
            ph_in ^= 1u;
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma_bf16(tb, d_lin + (ks >> 2) * (SLAB >> 4) + (ks & 3) * 2, d_w0 + 16 * ks, id128, ks > 0);
                umma_bf16(tb, d_xlat, d_w0 + 16 * 8, id128, 1u);
                umma_bf16(tb, d_ones, d_w0 + 16 * 9, id128, 1u);
                umma_commit(bar_acc);
                umma_commit(bar_empty);
            }
            __syncwarp();
            mbar_wait(bar_act, ph_act);   // layer 1: 128 -> 64
            ph_act ^= 1u;
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma_bf16(tb, d_x1 + (ks >> 2) * (SLAB >> 4) + (ks & 3) * 2, d_w1 + 16 * ks, id64, ks > 0);
                umma_bf16(tb, d_ones, d_w1 + 16 * 8, id64, 1u);
                umma_commit(bar_acc);
            }
            __syncwarp();
            mbar_wait(bar_act, ph_act);   // layer 2: 64 -> 32
            ph_act ^= 1u;
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_bf16(tb, d_x2 + ks * 2, d_w2 + 16 * ks, id32, ks > 0);
                umma_bf16(tb, d_ones, d_w2 + 16 * 4, id32, 1u);
                umma_commit(bar_acc);
            }
            __syncwarp();
            mbar_wait(bar_act, ph_act);   // layer 3: 32 -> 3 (N = 16)
            ph_act ^= 1u;
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) umma_bf16(tb, d_x3 + ks * 2, d_w3 + 16 * ks, id16, ks > 0);
                umma_bf16(tb, d_ones, d_w3 + 16 * 2, id16, 1u);
                umma_commit(bar_acc);
            }
            __syncwarp();
        }
    } else {
        const int q = warp & 3, h = (warp >> 2) & 1;
        const int row = q * 32 + lane;
        const int et = tid - sl * (EPI_WARPS * 32);   // thread of the slot's epilogue group
        const uint32_t lane_base = tmem_base + 128u * sl + (static_cast<uint32_t>(q * 32) << 16);
        uint32_t ph_acc = 0;
        auto arrive_act = [&]() {
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive1(bar_act);
        };
        auto wait_acc = [&]() {
            mbar_wait(bar_acc, ph_acc);
            ph_acc ^= 1u;
            tc_fence_after();
        };
        for (int tile = tile0; tile < n_tiles; tile += tstep) {
            if (et < P) {  // every row of the tile carries the same latent (AE.py:51 tiles it over the k points)
                const float4 *lp = reinterpret_cast<const float4 *>(prm.lat + static_cast<long long>(tile) * prm.ld_lat);
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 f = __ldg(lp + i);
                    w[2 * i] = bf16_bits(f.x) | (bf16_bits(f.y) << 16);
                    w[2 * i + 1] = bf16_bits(f.z) | (bf16_bits(f.w) << 16);
                }
                store_k16_row(ss + S_XLAT, et, w);
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive1(bar_in);
            wait_acc();   // layer 0: 64 channels per warp -> X1 slab h
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t v[32];
                tmem_ld32(lane_base + h * 64 + j * 32, v);
                store_row_chunks<32, true>(ss + S_X1, row, h * 64 + j * 32, v);
            }
            arrive_act();
            wait_acc();   // layer 1: 32 channels per warp -> X2, over X1's first slab (layer 1 has read it)
            {
                uint32_t v[32];
                tmem_ld32(lane_base + h * 32, v);
                store_row_chunks<32, true>(ss + S_X1, row, h * 32, v);
            }
            arrive_act();
            wait_acc();   // layer 2: 16 channels per warp -> X3, over X1's second slab
            {
                uint32_t v[16];
                tmem_ld16(lane_base + h * 16, v);
                store_row_chunks<16, true>(ss + S_X1 + SLAB, row, h * 16, v);
            }
            arrive_act();
            wait_acc();   // layer 3: x, y, z of this position (no ReLU)
            if (h == 0) {
                uint32_t v[4];
                tmem_ld4(lane_base, v);
                float *o = prm.out + (static_cast<long long>(tile) * P + row) * 3;
                o[0] = __uint_as_float(v[0]);
                o[1] = __uint_as_float(v[1]);
                o[2] = __uint_as_float(v[2]);
            }
            tc_fence_before();
        }
    }
    tmem_free(tmem_base, TMEM_COLS, warp, MMA_WARP0);
}

// ---- host side ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

}  // namespace ws

// [rows, cols] bf16 row-major (row pitch ld elements), box = 64 columns x box_rows rows, 128-byte swizzle
int make_tmap_bf16_2d(void *map, const void *ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    ws::EncodeTiledFn fn = ws::encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return PCC_ERR_UNSUPPORTED;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t es[2] = {1, 1};
    const CUresult r = fn(static_cast<CUtensorMap *>(map), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides,
                          box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for a [%llu x %llu] bf16 tensor, pitch %llu", static_cast<int>(r),
                  static_cast<unsigned long long>(rows), static_cast<unsigned long long>(cols), static_cast<unsigned long long>(ld));
        return PCC_ERR_UNSUPPORTED;
    }
    return 0;
}

static bool dims_are(const PccMlpLayer *layers, int n_layers, const int *dims, const int *relu, int n) {
    if (n_layers != n) return false;
    for (int l = 0; l < n; ++l)
        if (layers[l].cin != dims[l] || layers[l].cout != dims[l + 1] || (layers[l].relu != 0) != (relu[l] != 0) || !layers[l].packed_w) return false;
    return true;
}

static long long *g_ws_dbg = nullptr;

template <typename K>
static int set_smem(K kernel, int bytes) {
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        set_error("chain_ws: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    return 0;
}

int ws_dispatch(const PccMlpInput *in, int n_inputs, int64_t rows, const PccMlpLayer *layers, int n_layers, int group, void *out,
                int out_dtype, cudaStream_t st, bool *handled) {
    using namespace ws;
    *handled = false;
    static const bool disabled = getenv("PCC_NO_WS") != nullptr;
    if (disabled || rows <= 0 || rows % P != 0 || rows / P > 0x7fffffff / 2) return 0;
    const int n_tiles = static_cast<int>(rows / P);
    const int sms = num_sms();
    static const int sa_dims[] = {3, 32, 64, 128}, sa_relu[] = {1, 1, 1};
    static const int pnf_dims[] = {131, 128, 256}, pnf_relu[] = {1, 1};
    static const int dec_dims[] = {144, 128, 64, 32, 3}, dec_relu[] = {1, 1, 1, 0};
    auto aligned16 = [](const void *p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };

    if (n_inputs == 1 && group == 16 && dims_are(layers, n_layers, sa_dims, sa_relu, 3) && in[0].dtype == 0 && in[0].channels == 3 &&
        in[0].row_div == 1 && layers[0].w_f32 && layers[0].b_f32) {
        SaParams p{};
        p.xyz = static_cast<const float *>(in[0].ptr);
        p.ld = in[0].ld;
        p.w0 = layers[0].w_f32;
        p.b0 = layers[0].b_f32;
        p.b2 = layers[2].b_f32;
        p.w1p = layers[1].packed_w;
        p.w2p = layers[2].packed_w;
        p.out = out;
        p.out_bf16 = out_dtype;
        p.n_tiles = n_tiles;
        p.dbg = g_ws_dbg;
        static const bool form1 = getenv("PCC_SA_KERNEL") && !strcmp(getenv("PCC_SA_KERNEL"), "1");
        const int pairs = (n_tiles + 1) / 2;
        const int grid = pairs < 2 * sms ? pairs : 2 * sms;  // 2 CTAs per SM, 2 tiles in flight each
        if (g_ws_dbg) {
            if (int r = set_smem(sa_chain_kernel<true>, sa::SMEM)) return r;
            sa_chain_kernel<true><<<grid, sa::THREADS, sa::SMEM, st>>>(p);
        } else if (!form1) {
            if (int r = set_smem(sa_chain2_kernel<0, 2>, sa::SMEM)) return r;
            sa_chain2_kernel<0, 2><<<grid, sa::THREADS, sa::SMEM, st>>>(p);
            *handled = true;
            return check_launch("sa_chain2_kernel");
        } else {
            if (int r = set_smem(sa_chain_kernel<false>, sa::SMEM)) return r;
            sa_chain_kernel<false><<<grid, sa::THREADS, sa::SMEM, st>>>(p);
        }
        *handled = true;
        return check_launch("sa_chain_kernel");
    }
    if (n_inputs == 2 && group <= 1 && out_dtype == 1 && dims_are(layers, n_layers, pnf_dims, pnf_relu, 2) && in[0].dtype == 1 &&
        in[0].channels == 128 && in[0].row_div == 1 && in[0].ld % 8 == 0 && aligned16(in[0].ptr) && in[1].dtype == 0 &&
        in[1].channels == 3 && in[1].row_div == 1 && aligned16(out)) {
        CUtensorMap tm_feat, tm_out;
        if (int r = make_tmap_bf16_2d(&tm_feat, in[0].ptr, rows, 128, in[0].ld, P)) return r;
        if (int r = make_tmap_bf16_2d(&tm_out, out, rows, 256, 256, P)) return r;
        PnfParams p{};
        p.xyz = static_cast<const float *>(in[1].ptr);
        p.ld = in[1].ld;
        p.w0p = layers[0].packed_w;
        p.w1p = layers[1].packed_w;
        p.n_tiles = n_tiles;
        if (int r = set_smem(pnf_chain_kernel, pnf::SMEM)) return r;
        const int grid = n_tiles < sms ? n_tiles : sms;
        pnf_chain_kernel<<<grid, pnf::THREADS, pnf::SMEM, st>>>(p, tm_feat, tm_out);
        *handled = true;
        return check_launch("pnf_chain_kernel");
    }
    if (n_inputs == 2 && group <= 1 && out_dtype == 0 && dims_are(layers, n_layers, dec_dims, dec_relu, 4) && in[0].dtype == 1 &&
        in[0].channels == 128 && in[0].row_div == 1 && in[0].ld % 8 == 0 && aligned16(in[0].ptr) && in[1].dtype == 0 &&
        in[1].channels == 16 && in[1].row_div == P && in[1].ld % 4 == 0 && aligned16(in[1].ptr)) {
        CUtensorMap tm_lin;
        if (int r = make_tmap_bf16_2d(&tm_lin, in[0].ptr, rows, 128, in[0].ld, P)) return r;
        DecParams p{};
        p.lat = static_cast<const float *>(in[1].ptr);
        p.ld_lat = in[1].ld;
        p.w0p = layers[0].packed_w;
        p.w1p = layers[1].packed_w;
        p.w2p = layers[2].packed_w;
        p.w3p = layers[3].packed_w;
        p.out = static_cast<float *>(out);
        p.n_tiles = n_tiles;
        static const bool one_slot = getenv("PCC_DEC_SLOTS") != nullptr && getenv("PCC_DEC_SLOTS")[0] == '1';   // A/B
        if (n_tiles > sms && !one_slot) {   // more than one tile per SM: two tiles in flight per CTA
            if (int r = set_smem(dec_chain2_kernel, dec2::SMEM)) return r;
            const int grid = (n_tiles + 1) / 2 < sms ? (n_tiles + 1) / 2 : sms;
            dec_chain2_kernel<<<grid, dec2::THREADS, dec2::SMEM, st>>>(p, tm_lin);
            *handled = true;
            return check_launch("dec_chain2_kernel");
        }
        if (int r = set_smem(dec_chain_kernel, dec::SMEM)) return r;
        const int grid = n_tiles < sms ? n_tiles : sms;
        dec_chain_kernel<<<grid, dec::THREADS, dec::SMEM, st>>>(p, tm_lin);
        *handled = true;
        return check_launch("dec_chain_kernel");
    }
    return 0;
}

}  // namespace pcc

PCC_API int pcc_sa_chain_indexed(const float *patches, const uint8_t *idx8, int64_t points, int pts_per_patch, const PccMlpLayer *layers,
                                 int n_layers, void *out, int out_dtype, void *stream) {
    using namespace pcc;
    using namespace pcc::ws;
    PCC_REQUIRE(patches && idx8 && layers && out, "pcc_sa_chain_indexed: null pointer");
    PCC_REQUIRE(points >= 0 && points < (1ll << 28) && (out_dtype == 0 || out_dtype == 1), "pcc_sa_chain_indexed: bad points / out_dtype");
    static const int sa_dims[] = {3, 32, 64, 128}, sa_relu[] = {1, 1, 1};
    if (!dims_are(layers, n_layers, sa_dims, sa_relu, 3) || !layers[0].w_f32 || !layers[0].b_f32 || pts_per_patch < 8 ||
        pts_per_patch > 256 || pts_per_patch % 8 != 0 || points % pts_per_patch != 0) {
        set_error("pcc_sa_chain_indexed: only the SetAbstraction shape 3-32-64-128 (ReLU), K = 16, patches of 8..256 points (multiple of 8)");
        return PCC_ERR_UNSUPPORTED;
    }
    if (points == 0) return 0;
    SaParams p{};
    p.xyz = patches;
    p.ld = 3;
    p.w0 = layers[0].w_f32;
    p.b0 = layers[0].b_f32;
    p.b2 = layers[2].b_f32;
    p.w1p = layers[1].packed_w;
    p.w2p = layers[2].packed_w;
    p.out = out;
    p.out_bf16 = out_dtype;
    p.n_tiles = static_cast<int>(points / 8);
    p.dbg = g_ws_dbg;
    p.idx8 = idx8;
    p.pts_per_patch = pts_per_patch;
    p.pts_shift = -1;
    for (int sh = 3; sh <= 8; ++sh)
        if ((1 << sh) == pts_per_patch) p.pts_shift = sh;
    const int sms = num_sms();
    // (a third form -- the epilogue group also computes the fp32 3 -> 32 layer of the next tile in its wait for layer 2, the MMA
    // warp only issues -- was measured at 362 us against 312 us: the epilogue groups, not the MMA warps, pace the kernel)
    // (eight epilogue warps per slot -- 18 warps per CTA at 56 registers, each slot's epilogues split over two groups, one
    // 32-column tcgen05.ld in flight per warp -- was measured at 398 us against 310 us: more warps lose here, as in round 1)
    static const int slots = [] {
        const char *e = getenv("PCC_SA_SLOTS");
        return e && atoi(e) == 2 ? 2 : 4;
    }();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (slots == 4) {
        const int quads = (p.n_tiles + 3) / 4;
        if (int r = set_smem(sa_chain2_kernel<1, 4>, sa::Lay<4>::SMEM_IDX)) return r;
        sa_chain2_kernel<1, 4><<<quads < sms ? quads : sms, sa::Lay<4>::THREADS, sa::Lay<4>::SMEM_IDX, st>>>(p);
    } else {
        const int pairs = (p.n_tiles + 1) / 2;
        if (int r = set_smem(sa_chain2_kernel<1, 2>, sa::SMEM_IDX)) return r;
        sa_chain2_kernel<1, 2><<<pairs < 2 * sms ? pairs : 2 * sms, sa::THREADS, sa::SMEM_IDX, st>>>(p);
    }
    return check_launch("sa_chain2_kernel<indexed>");
}

/* bring-up aid, not part of the public header: device buffer of 1024 int64 for clock64() ticks of the SA chain's CTA 0 */
PCC_API void pcc_debug_ws_timing(long long *buf) { pcc::g_ws_dbg = buf; }
