// Fused shared-MLP chain on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// Replaces the 1x1-convolution stacks + max-pool of the reference's point-set blocks, whose intermediates the
// reference materialises in HBM ([B*S,128,16,256] fp32 per SetAbstraction call):
//   pn_kit.SetAbstraction  conv0/1/2 + ReLU + max over K   /root/reference/pn_kit.py:196-207
//   pn_kit.PointNet        4 convs + max over points        /root/reference/pn_kit.py:124-144
//   pn_kit.MLP             4 convs                          /root/reference/pn_kit.py:289-305
//   PointnetSAModule.mlp   Conv2d+BN(folded)+ReLU + max     /root/reference/pointnet_sa_module.py:87-91
//
// A CTA (128 threads) walks tiles of P = 128 positions (rows).  Weights are packed once (bf16, canonical no-swizzle
// K-major core matrices, the bias folded in as an extra K column that multiplies a constant-one input channel) and stay
// resident in shared memory; activations live in ONE shared buffer, K-major ([8-channel chunk][position][8] bf16), that
// each epilogue overwrites in place (the MMA that read it has completed); accumulators live in TMEM.
//
// Two MMA orientations, chosen per layer:
//   N-form (every layer but a pooled last one):   D[128 positions, Cout] = X[128, K] . W[Cout, K]^T
//       activations are the A operand (M = 128), weights the B operand (N = Cout <= 256 per instruction).  A TMEM lane is
//       a position, so all 128 epilogue threads are busy whatever Cout is (32-channel layers included) and each thread
//       writes its position's next-layer channels as 16-byte chunks (8 channels), 512 contiguous bytes per warp.
//   T-form (last layer when a max over G consecutive positions follows):   D^T[Cout, 128] = W . X^T
//       weights are the A operand (M = 128 channels per instruction), activations the B operand.  A TMEM lane is a channel
//       and the columns are positions, so the max over a group is a plain in-register reduction over consecutive columns
//       and the pooled row is written coalesced across the warp's 32 channels.
// The epilogue is ReLU + bf16 pack only (bias rides in the MMA).  tcgen05.mma is issued by thread 0; completion is
// tracked with tcgen05.commit on an mbarrier; MMA/epilogue overlap comes from the co-resident CTAs of an SM.
#include <cuda_bf16.h>

#include "chain_ws.h"
#include "pcc_common.cuh"
#include "tc_ptx.cuh"

namespace pcc {

constexpr int MLP_P = 128;  // positions per tile
constexpr int MLP_THREADS = 128;
constexpr int MLP_COMPUTE_THREADS = 128;
constexpr int MLP_MAX_LAYERS = PCC_MLP_MAX_LAYERS;
constexpr int MLP_CHUNK = (MLP_P / 8) * 128;  // bytes between 8-channel chunks of the activation buffer (LBO)

struct MlpChainParams {
    int n_layers;
    int cin[MLP_MAX_LAYERS], cout[MLP_MAX_LAYERS], kp[MLP_MAX_LAYERS], relu[MLP_MAX_LAYERS];
    int tform[MLP_MAX_LAYERS];    // 1: channels on lanes (pooled last layer)
    int ncol[MLP_MAX_LAYERS];     // TMEM columns the layer's accumulator uses
    int w_rows[MLP_MAX_LAYERS];   // weight rows kept in shared memory (Cout rounded to 16, or to 128 for T-form)
    int w_off[MLP_MAX_LAYERS];    // shared-memory byte offset of the packed weights of layer l
    const void *w[MLP_MAX_LAYERS];
    int first_fp32;               // 1: layer 0 runs in fp32 on the CUDA cores inside the producer warp (cin <= 8)
    int w0f_off;                  // its [cout][cin + 1] fp32 weights (bias last) in shared memory
    const float *w0f, *b0f;
    int x0_off;                   // input tile buffer (filled by the producer warp)
    int x_off;                    // the activation buffer
    int x_bytes;
    int tmem_cols;
    int ctrl_off;                 // mbarrier + TMEM base address slot
};

// ---- weight packing -------------------------------------------------------------------------------------------
// Packed layer = [rows128 x kp] bf16, kp = roundup(cin + 1, 16), in the K-major no-swizzle core-matrix layout:
//   offset(ch, k) = (ch / 8) * (kp * 16) + (k / 8) * 128 + (ch % 8) * 16 + (k % 8) * 2        bytes
// (SBO = kp*16 between 8-channel groups, LBO = 128 between 8-wide k chunks).  Column k = cin holds the bias (it
// multiplies the constant-one channel the kernel appends to every activation tile); everything else is zero padded.
__global__ void __launch_bounds__(256)
mlp_pack_kernel(const float *__restrict__ w, const float *__restrict__ bias, int cin, int cout, int kp, int rows128,
                __nv_bfloat16 *__restrict__ packed) {
    const long long total = static_cast<long long>(rows128) * kp;
    for (long long e = blockIdx.x * 256ll + threadIdx.x; e < total; e += static_cast<long long>(gridDim.x) * 256ll) {
        const int grp = static_cast<int>(e / (kp * 8));  // 8-channel group
        const int r2 = static_cast<int>(e - static_cast<long long>(grp) * kp * 8);
        const int kc = r2 / 64;                          // 8-wide k chunk
        const int r3 = r2 - kc * 64;
        const int chl = r3 / 8, kl = r3 - chl * 8;
        const int ch = grp * 8 + chl, k = kc * 8 + kl;
        float v = 0.0f;
        if (ch < cout) {
            if (k < cin) v = w[static_cast<size_t>(ch) * cin + k];
            else if (k == cin) v = bias[ch];
        }
        packed[e] = __float2bfloat16_rn(v);
    }
}

// ---- the chain kernel -------------------------------------------------------------------------------------------
struct MlpSeg {
    const void *ptr;   // [rows / row_div, ld] row-major
    long long ld;
    int dtype;         // 0 = fp32, 1 = bf16
    int ch;            // channels taken from this segment (they are concatenated in segment order)
    int row_div;       // source row = position row / row_div (broadcast of a per-group vector, e.g. the latent)
    int vec;           // 1: bf16 rows can be moved as 16-byte chunks (8 channels)
};

struct MlpIo {
    MlpSeg seg[PCC_MLP_MAX_INPUTS];
    int n_seg;
    int out_bf16;
    long long *timing;  // bring-up aid (NULL in production): CTA 0 / thread 0 stores clock64() at phase boundaries
};

#define MLP_TICK(slot)                                                                                 \
    do {                                                                                               \
        if constexpr (TIMED) {                                                                         \
            if (io.timing && blockIdx.x == 0 && tid == 0 && tick < 128) io.timing[tick++] = clock64(); \
        }                                                                                              \
    } while (0)

__device__ __forceinline__ void store_out(float *__restrict__ out_f, __nv_bfloat16 *__restrict__ out_h, long long o, float v) {
    if (out_h) out_h[o] = __float2bfloat16_rn(v); else out_f[o] = v;
}

// v = 32 consecutive positions of one channel; writes the 32 / G pooled values.  `o` = output offset of the first group.
template <int G>
__device__ __forceinline__ void pool_store_small(const uint32_t (&v)[32], int relu, float *__restrict__ out_f,
                                                 __nv_bfloat16 *__restrict__ out_h, long long o, int groups_left, int CL) {
#pragma unroll
    for (int g = 0; g < 32 / G; ++g) {
        float m = __uint_as_float(v[g * G]);
#pragma unroll
        for (int i = 1; i < G; ++i) m = fmaxf(m, __uint_as_float(v[g * G + i]));
        if (relu) m = fmaxf(m, 0.0f);  // ReLU commutes with max
        if (g < groups_left) store_out(out_f, out_h, o + g * CL, m);
    }
}

// for (j = 0; j < nj; ++j) body(v_j, j)   with the TMEM load of chunk j+1 in flight while chunk j is processed
template <typename Body>
__device__ __forceinline__ void for_each_chunk(uint32_t base, int nj, Body body) {
    uint32_t v0[32], v1[32];
    if (nj <= 0) return;
    tmem_ld32_issue(base, v0);
    for (int j = 0; j < nj; j += 2) {
        tmem_ld32_wait(v0);
        if (j + 1 < nj) tmem_ld32_issue(base + (j + 1) * 32, v1);
        body(v0, j);
        if (j + 1 < nj) {
            tmem_ld32_wait(v1);
            if (j + 2 < nj) tmem_ld32_issue(base + (j + 2) * 32, v0);
            body(v1, j + 1);
        }
    }
}

__device__ __forceinline__ void compute_sync() { __syncthreads(); }
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16 : 0) : "memory");
}

// ---- input staging ---------------------------------------------------------------------------------------------------
// Element (p, c) of the first MMA operand lives at (p/8)*128 + (c/8)*MLP_CHUNK + (p%8)*16 + (c%8)*2 of the x0 buffer.
// The next tile's input is fetched while the current tile is computed:
//   * bf16 segments whose rows can move as 16-byte chunks go global -> shared with cp.async (no registers), issued as
//     soon as the current tile's first MMA has finished reading x0;
//   * narrow fp32 / unaligned segments (<= MLP_PF elements per thread) are prefetched into registers at the top of the
//     current tile and converted / stored at the top of the next one;
//   * a first layer with <= 8 input channels runs here in fp32 on the CUDA cores (pn_kit.SetAbstraction's 3 -> 32):
//     reading a 32-column accumulator back from TMEM costs more than 96 FMAs per position, and the coordinates stay fp32.
constexpr int MLP_PF = 8;

// Source row of position row r for a broadcast segment.  A 64-bit division costs ~100 instructions, so the common
// row_div == 1 case must be a real (uniform) branch, and the rest uses 32-bit math (the host checks rows < 2^31).
__device__ __forceinline__ long long src_row(long long r, int row_div) {
    if (row_div == 1) return r;
    return static_cast<long long>(static_cast<unsigned>(r) / static_cast<unsigned>(row_div));
}

struct Prefetch {
    uint32_t v[PCC_MLP_MAX_INPUTS][MLP_PF];  // raw bits (fp32, or bf16 in the low half): converted when staged, so the
};                                           // loads have no dependent instruction and stay in flight
__device__ __forceinline__ float pf_value(uint32_t raw, int dtype) {
    return dtype == 0 ? __uint_as_float(raw) : __uint_as_float(raw << 16);
}

__device__ __forceinline__ void prefetch_regs(const MlpIo &io, const MlpChainParams &prm, long long rows, long long row0,
                                              int tid, Prefetch &pf) {
    if (prm.first_fp32) {  // thread = position: its cin input values
        const MlpSeg sg = io.seg[0];
        const long long r = row0 + tid;
        const long long o = src_row(r, sg.row_div) * sg.ld;
#pragma unroll
        for (int k = 0; k < MLP_PF; ++k) {
            pf.v[0][k] = 0u;
            if (k < sg.ch && r < rows) pf.v[0][k] = __ldg(static_cast<const unsigned *>(sg.ptr) + o + k);
        }
        return;
    }
#pragma unroll
    for (int s = 0; s < PCC_MLP_MAX_INPUTS; ++s) {
        if (s >= io.n_seg) break;
        const MlpSeg sg = io.seg[s];
        if (sg.vec || sg.ch > MLP_PF) continue;
        int p = tid / sg.ch, c = tid - p * sg.ch;
        const int dp = MLP_COMPUTE_THREADS / sg.ch, dc = MLP_COMPUTE_THREADS - dp * sg.ch;
#pragma unroll
        for (int u = 0; u < MLP_PF; ++u) {  // element e = u*128 + tid of the [P, ch] block
            pf.v[s][u] = 0u;
            if (u < sg.ch) {
                const long long r = row0 + p;
                if (r < rows) {
                    const long long o = src_row(r, sg.row_div) * sg.ld + c;
                    if (sg.dtype == 0) pf.v[s][u] = __ldg(static_cast<const unsigned *>(sg.ptr) + o);
                    else pf.v[s][u] = __ldg(static_cast<const unsigned short *>(sg.ptr) + o);
                }
                p += dp;
                c += dc;
                if (c >= sg.ch) {
                    c -= sg.ch;
                    ++p;
                }
            }
        }
    }
}

__device__ __forceinline__ void issue_cp_async(const MlpIo &io, long long rows, long long row0, unsigned char *x0, int tid) {
    int coff = 0;
    for (int s = 0; s < io.n_seg; ++s) {
        const MlpSeg sg = io.seg[s];
        if (sg.vec) {
            // unit e = (row_group * n8 + chunk) * 8 + row_in_group: 8 consecutive lanes copy the same 16-byte chunk of 8
            // consecutive rows, so a warp instruction touches 8 rows x 64 contiguous bytes in global memory (8 L1
            // wavefronts instead of 32) and 4 x 128 contiguous bytes in shared memory (conflict free)
            const int n8 = sg.ch >> 3;
            const int per_rg = n8 * 8;
            int rg = tid / per_rg, rem = tid - rg * per_rg;
            const int drg = MLP_COMPUTE_THREADS / per_rg, drem = MLP_COMPUTE_THREADS - drg * per_rg;
            const uint32_t dst0 = smem_u32(x0) + (coff >> 3) * MLP_CHUNK;
            for (int e = tid; e < MLP_P * n8; e += MLP_COMPUTE_THREADS) {
                const int c8 = rem >> 3, p = rg * 8 + (rem & 7);
                const long long r = row0 + p;
                const bool ok = r < rows;
                const long long sr = ok ? src_row(r, sg.row_div) : 0;
                cp_async16(dst0 + rg * 128 + (rem & 7) * 16 + c8 * MLP_CHUNK,
                           static_cast<const __nv_bfloat16 *>(sg.ptr) + sr * sg.ld + c8 * 8, ok);
                rg += drg;
                rem += drem;
                if (rem >= per_rg) {
                    rem -= per_rg;
                    ++rg;
                }
            }
        }
        coff += sg.ch;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// Registers (and, for wide scalar segments, a synchronous load) -> x0; also the constant-one channel and the K padding.
__device__ __forceinline__ void stage_tile(const MlpIo &io, const MlpChainParams &prm, long long rows, long long row0,
                                           unsigned char *x0, const float *w0s, int tid, const Prefetch &pf) {
    const uint32_t ONE_BF16 = 0x3f80u;
    unsigned char *xr = x0 + (tid >> 3) * 128 + (tid & 7) * 16;
    if (prm.first_fp32) {
        // weights in shared memory: per output channel two float4 (w0, w1, w2, bias | w3, w4, w5, w6), zero padded to
        // a multiple of 8 channels, so the loop is branch free
        const int cin = prm.cin[0], cout = prm.cout[0], relu = prm.relu[0], kp1 = prm.kp[1];
        const float4 *w4 = reinterpret_cast<const float4 *>(w0s);
        const bool wide = cin > 3;
        for (int c8 = 0; c8 * 8 < cout; ++c8) {
            float a[8];
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                const float4 wa = w4[(c8 * 8 + ch) * 2];
                float acc = fmaf(wa.x, __uint_as_float(pf.v[0][0]), wa.w);
                acc = fmaf(wa.y, __uint_as_float(pf.v[0][1]), acc);
                acc = fmaf(wa.z, __uint_as_float(pf.v[0][2]), acc);
                if (wide) {
                    const float4 wb = w4[(c8 * 8 + ch) * 2 + 1];
                    acc = fmaf(wb.x, __uint_as_float(pf.v[0][3]), acc);
                    acc = fmaf(wb.y, __uint_as_float(pf.v[0][4]), acc);
                    acc = fmaf(wb.z, __uint_as_float(pf.v[0][5]), acc);
                    acc = fmaf(wb.w, __uint_as_float(pf.v[0][6]), acc);
                }
                a[ch] = relu ? fmaxf(acc, 0.0f) : acc;
            }
            *reinterpret_cast<uint4 *>(xr + c8 * MLP_CHUNK) =
                make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7]));
        }
        const int covered = (cout + 7) / 8 * 8;
        if (cout < covered)
            *reinterpret_cast<unsigned short *>(xr + (cout >> 3) * MLP_CHUNK + (cout & 7) * 2) = static_cast<unsigned short>(ONE_BF16);
        for (int cc = covered; cc < kp1; cc += 8)
            *reinterpret_cast<uint4 *>(xr + (cc >> 3) * MLP_CHUNK) = make_uint4(cc == cout ? ONE_BF16 : 0u, 0u, 0u, 0u);
        return;
    }
    int coff = 0;
#pragma unroll
    for (int s = 0; s < PCC_MLP_MAX_INPUTS; ++s) {
        if (s >= io.n_seg) break;
        const MlpSeg sg = io.seg[s];
        if (!sg.vec) {
            if (sg.ch <= MLP_PF) {
                int p = tid / sg.ch, c = tid - p * sg.ch;
                const int dp = MLP_COMPUTE_THREADS / sg.ch, dc = MLP_COMPUTE_THREADS - dp * sg.ch;
#pragma unroll
                for (int u = 0; u < MLP_PF; ++u) {
                    if (u < sg.ch) {
                        const int cc = coff + c;
                        *reinterpret_cast<__nv_bfloat16 *>(x0 + (p >> 3) * 128 + (cc >> 3) * MLP_CHUNK + (p & 7) * 16 + (cc & 7) * 2) =
                            __float2bfloat16_rn(pf_value(pf.v[s][u], sg.dtype));
                        p += dp;
                        c += dc;
                        if (c >= sg.ch) {
                            c -= sg.ch;
                            ++p;
                        }
                    }
                }
            } else {  // wide scalar segment: synchronous, coalesced along the channels of a row
                const int total = MLP_P * sg.ch;
                for (int e = tid; e < total; e += MLP_COMPUTE_THREADS) {
                    const int p = e / sg.ch, c = e - p * sg.ch;
                    const long long r = row0 + p;
                    float v = 0.0f;
                    if (r < rows) {
                        const long long o = src_row(r, sg.row_div) * sg.ld + c;
                        v = sg.dtype == 0 ? __ldg(static_cast<const float *>(sg.ptr) + o)
                                          : __bfloat162float(static_cast<const __nv_bfloat16 *>(sg.ptr)[o]);
                    }
                    const int cc = coff + c;
                    *reinterpret_cast<__nv_bfloat16 *>(x0 + (p >> 3) * 128 + (cc >> 3) * MLP_CHUNK + (p & 7) * 16 + (cc & 7) * 2) =
                        __float2bfloat16_rn(v);
                }
            }
        }
        coff += sg.ch;
    }
    const int c0 = prm.cin[0], kp0 = prm.kp[0];
    int cc = c0;
    for (; (cc & 7) != 0 && cc < kp0; ++cc)
        *reinterpret_cast<unsigned short *>(xr + (cc >> 3) * MLP_CHUNK + (cc & 7) * 2) = static_cast<unsigned short>(cc == c0 ? ONE_BF16 : 0u);
    for (; cc < kp0; cc += 8) *reinterpret_cast<uint4 *>(xr + (cc >> 3) * MLP_CHUNK) = make_uint4(cc == c0 ? ONE_BF16 : 0u, 0u, 0u, 0u);
}

// Work unit = max(1, group / P) consecutive tiles of P positions.
// out: [rows, CL] (group <= 1) or [rows / group, CL] (max over each run of `group` consecutive rows), fp32 or bf16.
// POOL: 0 = no pooling, 2..32 = max over that many consecutive rows (in-register), 64 = any larger group.
template <int POOL, int MINB, bool TIMED>
__global__ void __launch_bounds__(MLP_THREADS, MINB)
mlp_chain_kernel(const __grid_constant__ MlpChainParams prm, const __grid_constant__ MlpIo io, long long rows, int group,
                 void *__restrict__ out, long long n_units, int tiles_per_unit) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t mbar = smem_base + prm.ctrl_off;        // MMA completion
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + prm.ctrl_off + 8);
    float *out_f = io.out_bf16 ? nullptr : static_cast<float *>(out);
    __nv_bfloat16 *out_h = io.out_bf16 ? static_cast<__nv_bfloat16 *>(out) : nullptr;
    unsigned char *xbuf = smem + prm.x_off;
    unsigned char *x0buf = smem + prm.x0_off;
    const float *w0s = reinterpret_cast<const float *>(smem + prm.w0f_off);
    const uint32_t ONE_BF16 = 0x3f80u;

    // ---- prologue: weights -> shared memory, barrier init, TMEM allocation ----
    for (int l = 0; l < prm.n_layers; ++l) {
        const int4 *src = static_cast<const int4 *>(prm.w[l]);
        int4 *dst = reinterpret_cast<int4 *>(smem + prm.w_off[l]);
        const int n16 = prm.w_rows[l] * prm.kp[l] * 2 / 16;
        for (int i = tid; i < n16; i += MLP_THREADS) dst[i] = src[i];
    }
    if (prm.first_fp32) {  // [cout_pad8][8] floats: (w0, w1, w2, bias, w3, w4, w5, w6), zeros beyond cin / cout
        float *w0w = reinterpret_cast<float *>(smem + prm.w0f_off);
        const int cin0 = prm.cin[0], n0 = prm.cout[0], n0p = (n0 + 7) / 8 * 8;
        for (int i = tid; i < n0p * 8; i += MLP_THREADS) {
            const int c = i >> 3, slot = i & 7;
            const int k = slot < 3 ? slot : slot - 1;  // slot 3 is the bias
            float v = 0.0f;
            if (c < n0) {
                if (slot == 3) v = prm.b0f[c];
                else if (k < cin0) v = prm.w0f[c * cin0 + k];
            }
            w0w[i] = v;
        }
    }
    if (tid == 0) mbar_init(mbar, 1);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(static_cast<uint32_t>(prm.tmem_cols))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int L = prm.n_layers;
    const int CL = prm.cout[L - 1];
    const int lstart = prm.first_fp32;
    const long long n_tiles_total = n_units * tiles_per_unit;

    unsigned char *xrow = xbuf + (tid >> 3) * 128 + (tid & 7) * 16;  // this thread's position inside every chunk
    uint32_t phase = 0;
    int tick = 0;
    Prefetch pf;
    {   // first tile of this CTA: fetch synchronously
        const long long t0 = static_cast<long long>(blockIdx.x) * tiles_per_unit;
        if (t0 < n_tiles_total) {
            prefetch_regs(io, prm, rows, t0 * MLP_P, tid, pf);
            issue_cp_async(io, rows, t0 * MLP_P, x0buf, tid);
        }
    }
    for (long long unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        float run_max = -INFINITY;  // running max across the tiles of one unit (pooled last layer, group > P, Cout <= 128)

        for (int sub = 0; sub < tiles_per_unit; ++sub) {
            const long long row0 = (unit * tiles_per_unit + sub) * MLP_P;
            // tile that follows this one in this CTA's sequence (prefetch target)
            const long long next_tile = sub + 1 < tiles_per_unit ? unit * tiles_per_unit + sub + 1
                                                                 : (unit + gridDim.x) * tiles_per_unit;
            const bool has_next = next_tile < n_tiles_total;
            MLP_TICK(0);
            stage_tile(io, prm, rows, row0, x0buf, w0s, tid, pf);          // registers -> x0 (data fetched during the previous tile)
            MLP_TICK(0);
            asm volatile("cp.async.wait_group 0;" ::: "memory");             // ... and the cp.async part has landed
            MLP_TICK(0);
            if (has_next) prefetch_regs(io, prm, rows, next_tile * MLP_P, tid, pf);  // in flight for the whole tile
            MLP_TICK(0);
            fence_async_smem();
            compute_sync();

            for (int l = lstart; l < L; ++l) {
                const int kp = prm.kp[l];
                const int tform = prm.tform[l];
                const int ncol = prm.ncol[l];
                // ---- MMA: one elected thread ----
                if (tid == 0) {
                    tc_fence_after();
                    MLP_TICK(1);
                    const uint32_t w_base = smem_base + prm.w_off[l];
                    const uint32_t x_base = smem_base + (l == lstart ? prm.x0_off : prm.x_off);
                    if (tform) {   // D^T[128 channels, P] per M tile: A = weights, B = activations
                        const uint32_t idesc = umma_idesc(128, MLP_P);
                        for (int t = 0; t < ncol / MLP_P; ++t)
                            for (int ks = 0; ks < kp / 16; ++ks)
                                umma_bf16(tmem_base + t * MLP_P, umma_desc(w_base + t * 128 * kp * 2 + ks * 256, 128, kp * 16),
                                          umma_desc(x_base + ks * 2 * MLP_CHUNK, MLP_CHUNK, 128), idesc, ks > 0 ? 1u : 0u);
                    } else {       // D[P positions, Cout]: A = activations, B = weights (N <= 256 per instruction)
                        for (int n0 = 0; n0 < ncol; n0 += 256) {
                            const int n = ncol - n0 < 256 ? ncol - n0 : 256;
                            const uint32_t idesc = umma_idesc(128, n);
                            for (int ks = 0; ks < kp / 16; ++ks)
                                umma_bf16(tmem_base + n0, umma_desc(x_base + ks * 2 * MLP_CHUNK, MLP_CHUNK, 128),
                                          umma_desc(w_base + n0 * kp * 2 + ks * 256, 128, kp * 16), idesc, ks > 0 ? 1u : 0u);
                        }
                    }
                    MLP_TICK(2);
                    umma_commit(mbar);
                }
                MLP_TICK(3);
                mbar_wait(mbar, phase);
                phase ^= 1u;
                tc_fence_after();
                MLP_TICK(4);
                // x0 has been consumed: start moving the next tile's 16-byte-chunk segments into it
                if (l == lstart && has_next) issue_cp_async(io, rows, next_tile * MLP_P, x0buf, tid);
                MLP_TICK(4);

                const bool last = (l == L - 1);
                const int cout = prm.cout[l];
                const int relu = prm.relu[l];
                const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
                if (!tform) {
                    // ---- N-form epilogue: thread = position (TMEM lane), columns = channels ----
                    if (!last) {
                        const int kpn = prm.kp[l + 1];
                        for_each_chunk(lane_addr, (ncol + 31) / 32, [&](const uint32_t (&v)[32], int j) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (j * 32 + q * 8 < ncol) {  // columns past ncol are stale and never stored
                                    uint32_t pk[4];
#pragma unroll
                                    for (int h = 0; h < 4; ++h) {  // ReLU after the (monotone) bf16 rounding: one packed max
                                        pk[h] = pack_bf16x2(__uint_as_float(v[q * 8 + 2 * h]), __uint_as_float(v[q * 8 + 2 * h + 1]));
                                        if (relu) pk[h] = relu_bf16x2(pk[h]);
                                    }
                                    *reinterpret_cast<uint4 *>(xrow + (j * 4 + q) * MLP_CHUNK) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                                }
                            }
                        });
                        // constant-one channel at index cout, zero chunks up to the next layer's padded K
                        if (cout < ncol)
                            *reinterpret_cast<unsigned short *>(xrow + (cout >> 3) * MLP_CHUNK + (cout & 7) * 2) =
                                static_cast<unsigned short>(ONE_BF16);
                        for (int cc = ncol; cc < kpn; cc += 8)
                            *reinterpret_cast<uint4 *>(xrow + (cc >> 3) * MLP_CHUNK) = make_uint4(cc == cout ? ONE_BF16 : 0u, 0u, 0u, 0u);
                    } else if constexpr (POOL == 0) {
                        const long long r = row0 + tid;
                        const bool row_ok = r < rows;
                        const long long obase = r * CL;
                        const bool stage_ok = prm.x_bytes >= 4 * 32 * 144;  // the activation buffer is free: staging space
                        for_each_chunk(lane_addr, (cout + 31) / 32, [&](const uint32_t (&v)[32], int j) {
                            const long long o = obase + j * 32;
                            if (out_h && (CL & 7) == 0 && stage_ok) {
                                // bf16 rows: stage the warp's 32 rows x 64 B (pitch 80 B), then every store instruction
                                // writes 8 rows x 64 contiguous bytes (8 L1 wavefronts instead of 32)
                                unsigned char *st = xbuf + warp * (32 * 144);
                                __syncwarp();
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    uint32_t pk[4];
#pragma unroll
                                    for (int h = 0; h < 4; ++h) {
                                        float a = __uint_as_float(v[q * 8 + 2 * h]), b = __uint_as_float(v[q * 8 + 2 * h + 1]);
                                        if (relu) {
                                            a = fmaxf(a, 0.0f);
                                            b = fmaxf(b, 0.0f);
                                        }
                                        pk[h] = pack_bf16x2(a, b);
                                    }
                                    *reinterpret_cast<uint4 *>(st + lane * 80 + q * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                                }
                                __syncwarp();
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const int rl = i * 8 + (lane >> 2), piece = lane & 3;
                                    const long long rr = row0 + warp * 32 + rl;
                                    if (rr < rows && j * 32 + piece * 8 < cout)
                                        *reinterpret_cast<uint4 *>(out_h + rr * CL + j * 32 + piece * 8) =
                                            *reinterpret_cast<const uint4 *>(st + rl * 80 + piece * 16);
                                }
                            } else if (out_f && (CL & 3) == 0 && stage_ok) {
                                // fp32 rows: 32 rows x 128 B (pitch 144 B); a store instruction writes 4 rows x 128 B
                                unsigned char *st = xbuf + warp * (32 * 144);
                                __syncwarp();
#pragma unroll
                                for (int q = 0; q < 8; ++q) {
                                    float4 f = make_float4(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]),
                                                           __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
                                    if (relu) f = make_float4(fmaxf(f.x, 0.f), fmaxf(f.y, 0.f), fmaxf(f.z, 0.f), fmaxf(f.w, 0.f));
                                    *reinterpret_cast<float4 *>(st + lane * 144 + q * 16) = f;
                                }
                                __syncwarp();
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const int rl = i * 4 + (lane >> 3), piece = lane & 7;
                                    const long long rr = row0 + warp * 32 + rl;
                                    if (rr < rows && j * 32 + piece * 4 < cout)
                                        *reinterpret_cast<float4 *>(out_f + rr * CL + j * 32 + piece * 4) =
                                            *reinterpret_cast<const float4 *>(st + rl * 144 + piece * 16);
                                }
                            } else if (!row_ok) {
                                return;
                            } else if (out_h && (CL & 7) == 0) {  // no staging space (single-layer chain): direct 16-byte stores
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    if (j * 32 + q * 8 < cout) {
                                        uint32_t pk[4];
#pragma unroll
                                        for (int h = 0; h < 4; ++h) {
                                            float a = __uint_as_float(v[q * 8 + 2 * h]), b = __uint_as_float(v[q * 8 + 2 * h + 1]);
                                            if (relu) {
                                                a = fmaxf(a, 0.0f);
                                                b = fmaxf(b, 0.0f);
                                            }
                                            pk[h] = pack_bf16x2(a, b);
                                        }
                                        *reinterpret_cast<uint4 *>(out_h + o + q * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                                    }
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 32; ++i) {
                                    if (j * 32 + i < cout) {
                                        float a = __uint_as_float(v[i]);
                                        if (relu) a = fmaxf(a, 0.0f);
                                        store_out(out_f, out_h, o + i, a);
                                    }
                                }
                            }
                        });
                    }
                } else if constexpr (POOL != 0) {
                    // ---- T-form epilogue (pooled last layer): thread = channel (TMEM lane), columns = positions ----
                    for (int t = 0; t < ncol / MLP_P; ++t) {
                        if (t * 128 + warp * 32 >= cout) break;  // warp-uniform: no real channel in this quadrant
                        const int c = t * 128 + warp * 32 + lane;
                        const bool real = c < cout;
                        float gmax = (group > 32 && t == 0) ? run_max : -INFINITY;
                        for_each_chunk(lane_addr + t * MLP_P, MLP_P / 32, [&](const uint32_t (&v)[32], int j) {
                            if (!real) return;
                            if constexpr (POOL >= 2 && POOL <= 32) {
                                // rows % POOL == 0 (checked by the host): whole groups only
                                const long long g0 = (row0 + j * 32) / POOL;
                                const long long gl = rows / POOL - g0;
                                pool_store_small<POOL>(v, relu, out_f, out_h, g0 * CL + c, gl > 32 ? 32 : static_cast<int>(gl), CL);
                            } else {
                                float m = __uint_as_float(v[0]);
#pragma unroll
                                for (int i = 1; i < 32; ++i) m = fmaxf(m, __uint_as_float(v[i]));
                                gmax = fmaxf(gmax, m);
                                const long long pos = static_cast<long long>(sub) * MLP_P + (j + 1) * 32;  // within the unit
                                const long long gsz = group < MLP_P ? group : static_cast<long long>(tiles_per_unit) * MLP_P;
                                if (pos % gsz == 0) {
                                    const long long r = unit * tiles_per_unit * MLP_P + pos - gsz;
                                    if (r < rows) store_out(out_f, out_h, (r / group) * CL + c, relu ? fmaxf(gmax, 0.0f) : gmax);
                                    gmax = -INFINITY;
                                }
                            }
                        });
                        if (group > 32 && t == 0) run_max = gmax;
                    }
                }
                MLP_TICK(5);
                tc_fence_before();
                fence_async_smem();
                compute_sync();
                MLP_TICK(6);
            }
        }
    }

    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(static_cast<uint32_t>(prm.tmem_cols))
                     : "memory");
    }
}

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

}  // namespace pcc

static long long *g_mlp_timing = nullptr;
/* bring-up aid, not part of the public header: device buffer of 256 int64 that CTA 0 fills with clock64() ticks */
PCC_API void pcc_debug_mlp_timing(long long *buf) { g_mlp_timing = buf; }

PCC_API int64_t pcc_mlp_packed_bytes(int cin, int cout) {
    if (cin < 1 || cout < 1) return 0;
    return static_cast<int64_t>(pcc::round_up(cout, 128)) * pcc::round_up(cin + 1, 16) * 2;
}

PCC_API int pcc_mlp_pack_weights_f32(const float *w, const float *bias, int cin, int cout, void *packed, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(w && bias && packed && cin >= 1 && cout >= 1, "pcc_mlp_pack_weights_f32: bad argument");
    const int kp = round_up(cin + 1, 16), rows128 = round_up(cout, 128);
    const long long total = static_cast<long long>(rows128) * kp;
    const int blocks = static_cast<int>((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    mlp_pack_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, bias, cin, cout, kp, rows128,
                                                                          static_cast<__nv_bfloat16 *>(packed));
    return check_launch("mlp_pack_kernel");
}

PCC_API int pcc_mlp_chain(const PccMlpInput *inputs, int n_inputs, int64_t rows, const PccMlpLayer *layers, int n_layers,
                          int group, void *out, int out_dtype, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(inputs && layers && out, "pcc_mlp_chain: null pointer");
    PCC_REQUIRE(n_inputs >= 1 && n_inputs <= PCC_MLP_MAX_INPUTS, "pcc_mlp_chain: n_inputs=%d outside [1,%d]", n_inputs,
                PCC_MLP_MAX_INPUTS);
    PCC_REQUIRE(n_layers >= 1 && n_layers <= MLP_MAX_LAYERS, "pcc_mlp_chain: n_layers=%d outside [1,%d]", n_layers,
                MLP_MAX_LAYERS);
    PCC_REQUIRE(rows >= 0 && rows < (1ll << 31) && group >= 0 && (out_dtype == 0 || out_dtype == 1),
                "pcc_mlp_chain: bad rows / group / out_dtype");
    if (rows == 0) return 0;
    MlpIo io{};
    io.n_seg = n_inputs;
    io.out_bf16 = out_dtype;
    io.timing = g_mlp_timing;
    int ctot = 0;
    for (int s = 0; s < n_inputs; ++s) {
        const PccMlpInput &in = inputs[s];
        PCC_REQUIRE(in.ptr && in.channels >= 1 && in.ld >= in.channels && in.row_div >= 1 && (in.dtype == 0 || in.dtype == 1),
                    "pcc_mlp_chain: bad input segment %d", s);
        io.seg[s].ptr = in.ptr;
        io.seg[s].ld = in.ld;
        io.seg[s].dtype = in.dtype;
        io.seg[s].ch = in.channels;
        io.seg[s].row_div = in.row_div;
        io.seg[s].vec = (in.dtype == 1 && in.channels % 8 == 0 && ctot % 8 == 0 && in.ld % 8 == 0 &&
                         reinterpret_cast<uintptr_t>(in.ptr) % 16 == 0) ? 1 : 0;
        ctot += in.channels;
    }
    PCC_REQUIRE(ctot == layers[0].cin, "pcc_mlp_chain: input segments carry %d channels, layer 0 expects %d", ctot,
                layers[0].cin);
    const bool pooled = group > 1;
    if (pooled) {
        PCC_REQUIRE(rows % group == 0, "pcc_mlp_chain: rows=%lld is not a multiple of group=%d",
                    static_cast<long long>(rows), group);
        const bool ok = (group <= MLP_P) ? (MLP_P % group == 0 && (group <= 32 ? 32 % group == 0 : group % 32 == 0))
                                         : (group % MLP_P == 0);
        if (!ok) {
            set_error("pcc_mlp_chain: group=%d must divide %d (and 32) or be a multiple of %d", group, MLP_P, MLP_P);
            return PCC_ERR_UNSUPPORTED;
        }
        if (group > MLP_P && layers[n_layers - 1].cout > 128) {
            set_error("pcc_mlp_chain: a max over more than %d rows needs the last layer to have <= 128 channels", MLP_P);
            return PCC_ERR_UNSUPPORTED;
        }
    }
    if (!g_mlp_timing) {  // the AE's three chains have compile-time-shaped, warp-specialised kernels (chain_ws.cu)
        bool handled = false;
        const int r = ws_dispatch(inputs, n_inputs, rows, layers, n_layers, group, out, out_dtype,
                                  static_cast<cudaStream_t>(stream), &handled);
        if (handled || r != 0) return r;
    }
    MlpChainParams prm{};
    prm.n_layers = n_layers;
    prm.first_fp32 = (n_layers >= 2 && n_inputs == 1 && inputs[0].dtype == 0 && layers[0].cin <= 7 && layers[0].cout <= 128 &&
                      layers[0].w_f32 && layers[0].b_f32) ? 1 : 0;
    prm.w0f = layers[0].w_f32;
    prm.b0f = layers[0].b_f32;
    int off = 0, max_cols = 0, max_kp = 0;
    for (int l = 0; l < n_layers; ++l) {
        PCC_REQUIRE(layers[l].packed_w && layers[l].cin >= 1 && layers[l].cout >= 1, "pcc_mlp_chain: bad layer %d", l);
        if (l > 0) PCC_REQUIRE(layers[l].cin == layers[l - 1].cout, "pcc_mlp_chain: layer %d cin != previous cout", l);
        prm.cin[l] = layers[l].cin;
        prm.cout[l] = layers[l].cout;
        prm.kp[l] = round_up(layers[l].cin + 1, 16);
        prm.relu[l] = layers[l].relu;
        prm.tform[l] = (pooled && l == n_layers - 1) ? 1 : 0;
        prm.w_rows[l] = prm.tform[l] ? round_up(layers[l].cout, 128) : round_up(layers[l].cout, 16);
        prm.ncol[l] = prm.w_rows[l];
        prm.w[l] = layers[l].packed_w;
        prm.w_off[l] = off;
        if (l == 0 && prm.first_fp32) {
            prm.w_rows[l] = 0;  // not staged: layer 0 runs on the CUDA cores
            prm.ncol[l] = 0;
        }
        off += prm.w_rows[l] * prm.kp[l] * 2;
        if (prm.ncol[l] > max_cols) max_cols = prm.ncol[l];
        if (prm.kp[l] > max_kp) max_kp = prm.kp[l];
    }
    if (max_cols > 512) {
        set_error("pcc_mlp_chain: a layer wider than 512 channels (TMEM columns %d) is not supported by this kernel", max_cols);
        return PCC_ERR_UNSUPPORTED;
    }
    off = round_up(off, 128);
    prm.w0f_off = off;
    if (prm.first_fp32) off += round_up(round_up(layers[0].cout, 8) * 8 * 4, 128);
    const int lfirst = prm.first_fp32;  // first layer that runs as an MMA; its operand is the producer's buffer
    // Without cp.async segments nothing lands in the input buffer while a tile is being computed (the next tile waits in
    // registers), so the activation buffer may alias it: one buffer of the larger size, and room for a 4th CTA per SM.
    bool any_vec = false;
    for (int s = 0; s < n_inputs; ++s) any_vec = any_vec || io.seg[s].vec;
    int max_kp_rest = 0;
    for (int l = lfirst + 1; l < n_layers; ++l) max_kp_rest = prm.kp[l] > max_kp_rest ? prm.kp[l] : max_kp_rest;
    prm.x0_off = off;
    if (any_vec) {
        off += MLP_P * prm.kp[lfirst] * 2;
        prm.x_off = off;
        prm.x_bytes = MLP_P * max_kp_rest * 2;
        off += prm.x_bytes;
    } else {
        const int kp_max = prm.kp[lfirst] > max_kp_rest ? prm.kp[lfirst] : max_kp_rest;
        prm.x_off = off;
        prm.x_bytes = max_kp_rest > 0 ? MLP_P * kp_max * 2 : 0;
        off += MLP_P * kp_max * 2;
    }
    prm.ctrl_off = off;
    off += 16;
    int cols = 32;
    while (cols < max_cols) cols <<= 1;
    prm.tmem_cols = cols;
    const size_t smem_bytes = static_cast<size_t>(off);
    if (smem_bytes > 227 * 1024) {
        set_error("pcc_mlp_chain: chain needs %zu bytes of shared memory (resident weights); max is %d", smem_bytes,
                  227 * 1024);
        return PCC_ERR_UNSUPPORTED;
    }
    const int tiles_per_unit = group > MLP_P ? group / MLP_P : 1;
    const long long n_tiles = (rows + MLP_P - 1) / MLP_P;
    const long long n_units = (n_tiles + tiles_per_unit - 1) / tiles_per_unit;
    // co-resident CTAs per SM: limited by shared memory and by TMEM columns (512 per SM)
    int per_sm = static_cast<int>((227 * 1024) / (smem_bytes + 1024));
    if (per_sm > 512 / cols) per_sm = 512 / cols;
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    long long grid = static_cast<long long>(num_sms()) * per_sm;
    if (grid > n_units) grid = n_units;
    const int pool = group <= 1 ? 0 : (group <= 32 ? group : 64);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaSuccess;
#define PCC_MLP_LAUNCH(POOL, MINB)                                                                                      \
    do {                                                                                                                \
        e = cudaFuncSetAttribute(mlp_chain_kernel<POOL, MINB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                 static_cast<int>(smem_bytes));                                                         \
        if (e == cudaSuccess)                                                                                           \
            mlp_chain_kernel<POOL, MINB, false><<<static_cast<unsigned>(grid), MLP_THREADS, smem_bytes, st>>>(          \
                prm, io, rows, group, out, n_units, tiles_per_unit);                                                    \
    } while (0)
#define PCC_MLP_LAUNCH_POOL(MINB)                                                                                       \
    switch (pool) {                                                                                                     \
        case 0: PCC_MLP_LAUNCH(0, MINB); break;                                                                         \
        case 2: PCC_MLP_LAUNCH(2, MINB); break;                                                                         \
        case 4: PCC_MLP_LAUNCH(4, MINB); break;                                                                         \
        case 8: PCC_MLP_LAUNCH(8, MINB); break;                                                                         \
        case 16: PCC_MLP_LAUNCH(16, MINB); break;                                                                       \
        case 32: PCC_MLP_LAUNCH(32, MINB); break;                                                                       \
        default: PCC_MLP_LAUNCH(64, MINB); break;                                                                       \
    }
    if (io.timing) {  // bring-up instantiations with clock64() ticks (tools/time_chain.py): pool 16 / none only
        if (pool == 16) {
            e = cudaFuncSetAttribute(mlp_chain_kernel<16, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes));
            if (e == cudaSuccess)
                mlp_chain_kernel<16, 4, true><<<static_cast<unsigned>(grid), MLP_THREADS, smem_bytes, st>>>(prm, io, rows, group, out, n_units, tiles_per_unit);
        } else if (pool == 0) {
            e = cudaFuncSetAttribute(mlp_chain_kernel<0, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes));
            if (e == cudaSuccess)
                mlp_chain_kernel<0, 2, true><<<static_cast<unsigned>(grid), MLP_THREADS, smem_bytes, st>>>(prm, io, rows, group, out, n_units, tiles_per_unit);
        } else {
            set_error("pcc_mlp_chain: timing instantiation exists for group 16 or none only");
            return PCC_ERR_UNSUPPORTED;
        }
    } else if (per_sm >= 3) { PCC_MLP_LAUNCH_POOL(4) } else { PCC_MLP_LAUNCH_POOL(2) }
    if (e != cudaSuccess) {
        set_error("pcc_mlp_chain: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    return check_launch("mlp_chain_kernel");
}

PCC_API int pcc_mlp_chain_f32(const float *x, int64_t rows, int ldx, const PccMlpLayer *layers, int n_layers, int group,
                              float *out, void *stream) {
    PCC_REQUIRE(x && layers && n_layers >= 1, "pcc_mlp_chain_f32: null pointer");
    PccMlpInput in{x, 0, layers[0].cin, ldx, 1};
    return pcc_mlp_chain(&in, 1, rows, layers, n_layers, group, out, 0, stream);
}
