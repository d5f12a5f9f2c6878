// Farthest point sampling for sm_100a.
//
// Replaces pn_kit.farthest_point_sample_batch (/root/reference/pn_kit.py:309-330) and PyTorch3D
// sample_farthest_points (/root/reference/pointnet_sa_module.py:10-13).  Bit-exact indices:
//   * running min distance per point, d2 computed with un-fused _rn arithmetic (pcc_common.cuh);
//   * arg-max with ties to the LOWEST index (torch.max(dim) / std::max_element return the first maximum).
//
// fps_block_kernel: one CTA per cloud, the whole cloud and its running distances live in REGISTERS
//   (THREADS x PPT points, N <= 8192).  One __syncthreads per iteration: every warp reduces with two
//   redux.sync (max over the float bits -- valid because d2 >= +0 -- then min over the indices that hold
//   the maximum), the owning lane publishes (bits, idx, x, y, z) in a double-buffered shared slot, and after
//   the barrier every warp reduces the <= 32 slots again, so no second barrier / broadcast is needed.
//   The kernel is latency bound (npoint dependent iterations); HBM traffic is 12*N bytes per cloud, once.
//
// fps_grid_kernel: clouds with N > 8192 (the 1M-point scene).  The cloud is spread
//   over C = ceil(N/8192) co-resident CTAs (cooperative launch), still register resident; per iteration each
//   CTA publishes its best candidate as tagged 64-bit words in its own slot and polls the other CTAs' slots
//   (no atomics, no fences: see the comment above the kernel).  For big clouds and long samplings the bucketed form of
//   fps_bucket.cu takes over after a head of iterations of this kernel (same output, one CTA per cloud, exact spatial
//   skipping); PCC_FPS_PATH=grid keeps this one throughout.
#include <stdlib.h>

#include "fps.cuh"

namespace pcc {

struct FpsSlots {
    unsigned bits[2][32];
    unsigned idx[2][32];
    float x[2][32], y[2][32], z[2][32];
};

// Block-wide arg-max of (bits, idx) with coordinates riding along.  Every thread returns the winner.
template <int NWARPS>
__device__ __forceinline__ void block_argmax(FpsSlots &s, int buf, unsigned bits, unsigned idx, float x, float y,
                                             float z, unsigned &w_bits, unsigned &w_idx, float &wx, float &wy,
                                             float &wz) {
    const unsigned lane = lane_id();
    const unsigned warp = threadIdx.x >> 5;
    const unsigned m = __reduce_max_sync(FULL_MASK, bits);
    const unsigned cand = (bits == m) ? idx : 0xffffffffu;
    const unsigned mi = __reduce_min_sync(FULL_MASK, cand);
    if (bits == m && idx == mi) {  // exactly one lane (indices are unique)
        s.bits[buf][warp] = m;
        s.idx[buf][warp] = mi;
        s.x[buf][warp] = x;
        s.y[buf][warp] = y;
        s.z[buf][warp] = z;
    }
    __syncthreads();
    unsigned sb = 0u, si = 0xffffffffu;
    if (lane < NWARPS) {
        sb = s.bits[buf][lane];
        si = s.idx[buf][lane];
    }
    const unsigned m2 = __reduce_max_sync(FULL_MASK, sb);
    const unsigned cand2 = (sb == m2) ? si : 0xffffffffu;
    const unsigned mi2 = __reduce_min_sync(FULL_MASK, cand2);
    const unsigned owner = __ffs(__ballot_sync(FULL_MASK, sb == m2 && si == mi2)) - 1;
    w_bits = m2;
    w_idx = mi2;
    wx = s.x[buf][owner];
    wy = s.y[buf][owner];
    wz = s.z[buf][owner];
}

// Update the PPT register-resident points of this thread against centre c and return the thread's best.
// Even PPT: two points per packed-fp32 instruction for the three differences and the three squares (each half is the IEEE
// operation of dist2_rn: p + (-c) == p - c exactly); the two sums per point stay scalar -- ptxas contracts mul.rn.f32x2 +
// add.rn.f32x2 into FFMA2 whatever --fmad says, which would change dist2_rn's rounding.  10 issue slots per pair instead of 16:
// the single-CTA kernel is issue bound (~150 instructions per thread and iteration on a full SM).
template <int PPT>
__device__ __forceinline__ void fps_local_update(const float (&px)[PPT], const float (&py)[PPT],
                                                 const float (&pz)[PPT], float (&md)[PPT], float cx, float cy,
                                                 float cz, unsigned &bits, int &slot) {
    float best = -1.0f;
    slot = 0;
    if constexpr (PPT % 2 == 0) {
        const unsigned long long ncx = dup_neg_f32x2(cx), ncy = dup_neg_f32x2(cy), ncz = dup_neg_f32x2(cz);
#pragma unroll
        for (int p = 0; p < PPT; p += 2) {
            const unsigned long long dx = add_f32x2(pack_f32x2(px[p], px[p + 1]), ncx), dy = add_f32x2(pack_f32x2(py[p], py[p + 1]), ncy),
                                     dz = add_f32x2(pack_f32x2(pz[p], pz[p + 1]), ncz);
            const float2 sx = unpack_f32x2(mul_f32x2(dx, dx)), sy = unpack_f32x2(mul_f32x2(dy, dy)), sz = unpack_f32x2(mul_f32x2(dz, dz));
            const float d0 = __fadd_rn(__fadd_rn(sx.x, sy.x), sz.x), d1 = __fadd_rn(__fadd_rn(sx.y, sy.y), sz.y);
            md[p] = fminf(md[p], d0);
            md[p + 1] = fminf(md[p + 1], d1);
            if (md[p] > best) {       // strict: the lowest p (lowest global index) keeps a tie
                best = md[p];
                slot = p;
            }
            if (md[p + 1] > best) {
                best = md[p + 1];
                slot = p + 1;
            }
        }
    } else {
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const float d = dist2_rn(px[p], py[p], pz[p], cx, cy, cz);
            md[p] = fminf(md[p], d);  // == `if (d < md) md = d` (pn_kit.py:327-328) for non-NaN inputs
            if (md[p] > best) {       // strict: the lowest p (lowest global index) keeps a tie
                best = md[p];
                slot = p;
            }
        }
    }
    bits = __float_as_uint(best);
}

template <int PPT>
__device__ __forceinline__ void select_slot(const float (&px)[PPT], const float (&py)[PPT], const float (&pz)[PPT],
                                            int slot, float &x, float &y, float &z) {
    x = px[0];
    y = py[0];
    z = pz[0];
#pragma unroll
    for (int p = 1; p < PPT; ++p)
        if (slot == p) {
            x = px[p];
            y = py[p];
            z = pz[p];
        }
}

template <int THREADS, int PPT>
__global__ void __launch_bounds__(THREADS, 1)
fps_block_kernel(const float *__restrict__ xyz, int N, int npoint, const int64_t *__restrict__ start_idx,
                 float init_dist, int64_t *__restrict__ out_idx, float *__restrict__ out_xyz, float quant_cube) {
    constexpr int NWARPS = THREADS / 32;
    __shared__ FpsSlots slots;
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const float *pc = xyz + static_cast<size_t>(b) * N * 3;
    int64_t *out = out_idx + static_cast<size_t>(b) * npoint;

    float px[PPT], py[PPT], pz[PPT], md[PPT];
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        const int g = p * THREADS + tid;
        if (g < N) {
            px[p] = pc[g * 3 + 0];
            py[p] = pc[g * 3 + 1];
            pz[p] = pc[g * 3 + 2];
            md[p] = init_dist;
        } else {  // padding: distance pinned at 0, index above every real one -> never preferred
            px[p] = py[p] = pz[p] = 0.0f;
            md[p] = 0.0f;
        }
    }
    const int k_n = npoint < N ? npoint : N;
    for (int k = k_n + tid; k < npoint; k += THREADS) {  // PyTorch3D padding when npoint > N
        out[k] = -1;
        if (out_xyz) {
            float *o = out_xyz + (static_cast<size_t>(b) * npoint + k) * 3;
            o[0] = o[1] = o[2] = 0.0f;
        }
    }

    int far = 0;
    if (start_idx) {  // an out-of-range start (the reference would raise an IndexError on the host) must not read outside the cloud
        const long long s0 = start_idx[b];
        far = s0 < 0 ? 0 : (s0 >= N ? N - 1 : static_cast<int>(s0));
    }
    float cx = pc[far * 3 + 0], cy = pc[far * 3 + 1], cz = pc[far * 3 + 2];
    for (int i = 0; i < k_n; ++i) {
        if (tid == 0) {
            out[i] = far;
            if (out_xyz) store_centre(out_xyz + (static_cast<size_t>(b) * npoint + i) * 3, cx, cy, cz, quant_cube);
        }
        if (i == k_n - 1) break;
        unsigned bits;
        int slot;
        fps_local_update<PPT>(px, py, pz, md, cx, cy, cz, bits, slot);
        float x, y, z;
        select_slot<PPT>(px, py, pz, slot, x, y, z);
        unsigned w_bits, w_idx;
        block_argmax<NWARPS>(slots, i & 1, bits, static_cast<unsigned>(slot * THREADS + tid), x, y, z, w_bits,
                             w_idx, cx, cy, cz);
        far = static_cast<int>(w_idx);
    }
}

// ---- multi-CTA variant ----------------------------------------------------------------------------------
// Grid-wide arg-max without atomics or fences.  Every CTA publishes its candidate per iteration into its own 32-byte slot of a
// double-buffered table as self-describing 64-bit words,
//     w0 = d2 bits << 32 | tag << 21 | (2^21 - 1 - index)      tag = (iteration + 1) mod 2^11
//     w1 = x bits  << 32 | tag << 21 | z bits [31:11]          (WIDE slots only)
//     w2 = y bits  << 32 | tag << 21 | z bits [10:0] << 10     (WIDE slots only)
// and warp 0 of every CTA polls all the slots of its cloud until their tags are this iteration's, taking the maximum w0
// (largest d2, then lowest index: all fresh words carry the same tag).  Each word is written and read as one relaxed 8-byte
// access, so a value and its tag can never be seen apart, and nothing else needs ordering (the coordinates are read-only
// input).  A slot of parity (i & 1) is rewritten at iteration i + 2, which a CTA reaches only after every CTA has published
// iteration i + 1, i.e. has finished reading iteration i.
// WIDE slots carry the winner's coordinates, saving the dependent load of xyz[winner] (one L2 round trip) at three times the
// polling traffic; that traffic (C^2 slot reads per iteration on a few L2 lines) is what costs at large C, so WIDE is used up
// to 128 CTAs per cloud.  Measured per iteration on B200 (C = 2 / 13 / 123 CTAs per cloud; C = 147 is 4.0-4.9 us for every
// variant, with more spread between boxes than between variants):
//     atomicMax + arrival counter + two fences + load of the winner (previous version)      -    / -         / 4.60 us
//     one word per slot + load of the winner                                                  1.93 / 2.06-2.72 / 3.76-3.89 us
//     WIDE slots                                                                              1.84 / 1.88-1.96 / 3.87-4.00 us
// and, not kept: one polling thread per slot instead of one warp (4.16 at C = 123: more pollers, more contention), eight
// replicas of the table (no better), a reducer CTA that republishes the winner to an outbox (3.8 us already at C = 2) and
// two-level groups of 8..32 CTAs (5.4-5.7 us at C = 123): a second, dependent store -> poll hop costs 1.3 us or more, far more
// than the polling traffic it saves.  Round 2 repeated that with thread-block clusters of 8 (cooperative + cluster launch, 16
// clusters for 1M points): (a) candidates pushed into every peer's shared memory (st.shared::cluster, local polling), the
// cluster winner through a 16-slot global table: 5.76 us / iteration; (b) the flat global table polled by one CTA per cluster
// only, which pushes the winner to its peers over DSMEM: 5.15 us -- against 4.40 us for this kernel on the same box.  Both were
// bit-exact against the oracle and both lose: the dependent DSMEM hop costs more than the polling traffic it removes.
// Also measured and not kept: polling all of a lane's slots together (five loads in flight per round, first words only, the
// winner's coordinate words fetched by one lane afterwards) instead of one slot after the other -- 6.0 us per iteration against
// 4.7 at 123 CTAs, 3.3 against 1.9 at 13: every round re-reads every slot, and that traffic is what the exchange is bound by.
// Instrumented with clock64 (1.97 GHz): CTA-local update + block arg-max 2020 clocks; publish -> every slot fresh 3300 clocks at
// C = 13 (one slot per lane: the bare store -> visible -> polled latency across the two dies) and 7750 at C = 123; barrier +
// unpacking the coordinates 215 (the instrumented kernel is ~30 % slower than the plain one: read these as proportions).
struct FpsGridWs {             // workspace unit per SM: 2 buffers x 32-byte slot per CTA, zero-filled by the host
    unsigned long long word[8];
};

constexpr int GRID_THREADS = 1024;
constexpr int GRID_PPT = 8;
constexpr int GRID_PTS_PER_CTA = GRID_THREADS * GRID_PPT;
constexpr unsigned GRID_IDX_MASK = (1u << 21) - 1u;   // 148 CTAs x 8192 points < 2^21
constexpr int GRID_WIDE_MAX_CTAS = 128;

__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <bool WIDE, bool HEAD>   // HEAD: the first centres of a longer row, running distances handed over at exit (fps_bucket.cu)
__global__ void __launch_bounds__(GRID_THREADS, 1)
fps_grid_kernel(const float *__restrict__ xyz, int N, int npoint, const int64_t *__restrict__ start_idx,
                float init_dist, int64_t *__restrict__ out_idx, unsigned long long *ws, int ctas_per_cloud, int cloud0,
                float *__restrict__ out_xyz, float quant_cube, int out_stride_arg, float *__restrict__ md_out) {
    constexpr int NWARPS = GRID_THREADS / 32;
    const int out_stride = HEAD ? out_stride_arg : npoint;
    __shared__ FpsSlots slots;
    __shared__ unsigned long long s_w0, s_w1, s_w2;
    const int cloud_local = blockIdx.x / ctas_per_cloud;
    const int part = blockIdx.x % ctas_per_cloud;
    const int b = cloud0 + cloud_local;
    const int tid = threadIdx.x;
    const float *pc = xyz + static_cast<size_t>(b) * N * 3;
    int64_t *out = out_idx + static_cast<size_t>(b) * out_stride;   // out_stride >= npoint: a prefix of a longer row (hand-over)
    unsigned long long *table = ws + static_cast<size_t>(cloud_local) * 8 * ctas_per_cloud;   // [2][ctas_per_cloud][4]
    const int base = part * GRID_PTS_PER_CTA;

    float px[GRID_PPT], py[GRID_PPT], pz[GRID_PPT], md[GRID_PPT];
#pragma unroll
    for (int p = 0; p < GRID_PPT; ++p) {
        const int g = base + p * GRID_THREADS + tid;
        if (g < N) {
            px[p] = pc[static_cast<size_t>(g) * 3 + 0];
            py[p] = pc[static_cast<size_t>(g) * 3 + 1];
            pz[p] = pc[static_cast<size_t>(g) * 3 + 2];
            md[p] = init_dist;
        } else {
            px[p] = py[p] = pz[p] = 0.0f;
            md[p] = 0.0f;
        }
    }
    const int k_n = npoint < N ? npoint : N;
    if (part == 0)
        for (int k = k_n + tid; k < npoint; k += GRID_THREADS) {
            out[k] = -1;
            if (out_xyz) {
                float *o = out_xyz + (static_cast<size_t>(b) * out_stride + k) * 3;
                o[0] = o[1] = o[2] = 0.0f;
            }
        }

    int far = 0;
    if (start_idx) {  // an out-of-range start (the reference would raise an IndexError on the host) must not read outside the cloud
        const long long s0 = start_idx[b];
        far = s0 < 0 ? 0 : (s0 >= N ? N - 1 : static_cast<int>(s0));
    }
    float cx = pc[static_cast<size_t>(far) * 3 + 0], cy = pc[static_cast<size_t>(far) * 3 + 1],
          cz = pc[static_cast<size_t>(far) * 3 + 2];
    for (int i = 0; i < k_n; ++i) {
        if (part == 0 && tid == 0) {
            out[i] = far;
            if (out_xyz) store_centre(out_xyz + (static_cast<size_t>(b) * out_stride + i) * 3, cx, cy, cz, quant_cube);
        }
        if (i == k_n - 1) break;
        unsigned bits;
        int slot;
        fps_local_update<GRID_PPT>(px, py, pz, md, cx, cy, cz, bits, slot);
        float x, y, z;
        select_slot<GRID_PPT>(px, py, pz, slot, x, y, z);
        unsigned w_bits, w_idx;
        float bx, by, bz;
        unsigned gidx = static_cast<unsigned>(base + slot * GRID_THREADS + tid);
        if (gidx >= static_cast<unsigned>(N)) gidx = 0x7fffffffu;  // padding lanes: above every real index
        block_argmax<NWARPS>(slots, i & 1, bits, gidx, x, y, z, w_bits, w_idx, bx, by, bz);
        // every CTA holds at least one real point and padding lanes lose ties, so w_idx < N < 2^21
        const unsigned long long tag = static_cast<unsigned long long>((i + 1) & 2047) << 21;
        constexpr unsigned long long TAG_MASK = 2047ull << 21;
        unsigned long long *row = table + static_cast<size_t>(i & 1) * ctas_per_cloud * 4;
        const unsigned long long my0 = (static_cast<unsigned long long>(w_bits) << 32) | tag | static_cast<unsigned long long>(GRID_IDX_MASK - w_idx);
        if constexpr (!WIDE) {
            if (tid < 32) {
                if (tid == 0) st_relaxed_u64(row + part * 4, my0);
                unsigned long long best = 0ull;
                for (int p = tid; p < ctas_per_cloud; p += 32) {
                    unsigned long long v;
                    do {
                        v = ld_relaxed_u64(row + p * 4);
                    } while ((v & TAG_MASK) != tag);
                    best = v > best ? v : best;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                    best = other > best ? other : best;
                }
                if (tid == 0) s_w0 = best;
            }
            __syncthreads();
            far = static_cast<int>(GRID_IDX_MASK - (static_cast<unsigned>(s_w0) & GRID_IDX_MASK));
            cx = __ldg(pc + static_cast<size_t>(far) * 3 + 0);
            cy = __ldg(pc + static_cast<size_t>(far) * 3 + 1);
            cz = __ldg(pc + static_cast<size_t>(far) * 3 + 2);
        } else {
            if (tid < 32) {
                if (tid == 0) {
                    const unsigned zb = __float_as_uint(bz);
                    st_relaxed_u64(row + part * 4 + 0, my0);
                    st_relaxed_u64(row + part * 4 + 1, (static_cast<unsigned long long>(__float_as_uint(bx)) << 32) | tag | (zb >> 11));
                    st_relaxed_u64(row + part * 4 + 2, (static_cast<unsigned long long>(__float_as_uint(by)) << 32) | tag |
                                                           static_cast<unsigned long long>((zb & 2047u) << 10));
                }
                unsigned long long best = 0ull, b1 = 0ull, b2 = 0ull;
                for (int p = tid; p < ctas_per_cloud; p += 32) {
                    unsigned long long v0, v1, v2;
                    do {
                        v0 = ld_relaxed_u64(row + p * 4 + 0);
                        v1 = ld_relaxed_u64(row + p * 4 + 1);
                        v2 = ld_relaxed_u64(row + p * 4 + 2);
                    } while ((v0 & TAG_MASK) != tag || (v1 & TAG_MASK) != tag || (v2 & TAG_MASK) != tag);
                    if (v0 > best || p == tid) {
                        best = v0;
                        b1 = v1;
                        b2 = v2;
                    }
                }
                unsigned long long top = best;   // lanes past the last slot hold 0, below every fresh word unless tag == 0 ...
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, top, o);
                    top = other > top ? other : top;
                }
                // ... where an empty lane's 0 can at most tie with a real word (d2 = 0, index 2^21 - 1): first REAL owner
                const unsigned owner = __ffs(__ballot_sync(0xffffffffu, tid < ctas_per_cloud && best == top)) - 1;
                if (tid == owner) {
                    s_w0 = best;
                    s_w1 = b1;
                    s_w2 = b2;
                }
            }
            __syncthreads();
            const unsigned long long t1 = s_w1, t2 = s_w2;
            far = static_cast<int>(GRID_IDX_MASK - (static_cast<unsigned>(s_w0) & GRID_IDX_MASK));
            cx = __uint_as_float(static_cast<unsigned>(t1 >> 32));
            cy = __uint_as_float(static_cast<unsigned>(t2 >> 32));
            cz = __uint_as_float(((static_cast<unsigned>(t1) & GRID_IDX_MASK) << 11) | ((static_cast<unsigned>(t2) >> 10) & 2047u));
        }
        // s_w0..2 are rewritten only after the next block_argmax's __syncthreads, which every thread reaches
        // after reading it here.
    }
    if (HEAD && md_out) {   // hand-over to the bucketed form: running distances after the centres out[0 .. k_n - 2]
#pragma unroll
        for (int p = 0; p < GRID_PPT; ++p) {
            const int g = base + p * GRID_THREADS + tid;
            if (g < N) md_out[static_cast<size_t>(b) * N + g] = md[p];
        }
    }
}

template <int THREADS, int PPT>
static int launch_block(const float *xyz, int B, int N, int npoint, const int64_t *start_idx, float init_dist,
                        int64_t *out_idx, float *out_xyz, float quant_cube, cudaStream_t st) {
    fps_block_kernel<THREADS, PPT><<<B, THREADS, 0, st>>>(xyz, N, npoint, start_idx, init_dist, out_idx, out_xyz, quant_cube);
    return check_launch("fps_block_kernel");
}

// multi-CTA path: C co-resident CTAs per cloud, as many clouds per cooperative launch as fit.  `out_stride` >= npoint is the row
// length of out_idx / out_xyz (the bucketed form lets this kernel write the head of a longer row); `md_out` (nullable) [B, N]
// receives the running distances at exit.  `workspace`: pcc_fps_grid_workspace_bytes().
int64_t fps_grid_workspace_bytes() { return static_cast<int64_t>(sizeof(FpsGridWs)) * num_sms(); }

int fps_grid_run(const float *xyz, int B, int N, int npoint, const int64_t *start_idx, float init_dist, int64_t *out_idx,
                 float *out_xyz, float quant_cube, int out_stride, float *md_out, void *workspace, cudaStream_t st) {
    const int sms = num_sms();
    const int C = (N + GRID_PTS_PER_CTA - 1) / GRID_PTS_PER_CTA;
    if (C > sms || static_cast<long long>(C) * GRID_PTS_PER_CTA > GRID_IDX_MASK) {
        set_error("pcc_fps_f32: N=%d exceeds the co-resident capacity %d", N, sms * GRID_PTS_PER_CTA);
        return PCC_ERR_UNSUPPORTED;
    }
    unsigned long long *ws = static_cast<unsigned long long *>(workspace);   // 2 x 32 bytes per CTA of a launch
    const int clouds_per_launch = sms / C;
    for (int c0 = 0; c0 < B; c0 += clouds_per_launch) {
        int nc = B - c0 < clouds_per_launch ? B - c0 : clouds_per_launch;
        cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(unsigned long long) * 8 * nc * C, st);
        if (e != cudaSuccess) {
            set_error("pcc_fps_f32: memset failed: %s", cudaGetErrorString(e));
            return static_cast<int>(e);
        }
        int ctas = C, cloud0 = c0;
        void *args[] = {&xyz, &N, &npoint, &start_idx, &init_dist, &out_idx, &ws, &ctas, &cloud0, &out_xyz, &quant_cube, &out_stride, &md_out};
        const bool head = md_out != nullptr || out_stride != npoint;
        void *fn = C <= GRID_WIDE_MAX_CTAS ? (head ? reinterpret_cast<void *>(fps_grid_kernel<true, true>) : reinterpret_cast<void *>(fps_grid_kernel<true, false>))
                                           : (head ? reinterpret_cast<void *>(fps_grid_kernel<false, true>) : reinterpret_cast<void *>(fps_grid_kernel<false, false>));
        e = cudaLaunchCooperativeKernel(fn, dim3(nc * C), dim3(GRID_THREADS),
                                        args, 0, st);
        if (e != cudaSuccess) {
            set_error("pcc_fps_f32: cooperative launch failed: %s", cudaGetErrorString(e));
            return static_cast<int>(e);
        }
    }
    return check_launch("fps_grid_kernel");
}

}  // namespace pcc

PCC_API int64_t pcc_fps_workspace_bytes(int B, int N, int npoint) {
    if (N <= pcc::GRID_PTS_PER_CTA || B <= 0) return 0;
    if (pcc::fps_bucket_takes(B, N, npoint)) return pcc::fps_bucket_workspace_bytes(B, N);
    return pcc::fps_grid_workspace_bytes();
}

PCC_API int pcc_fps_f32(const float *xyz, int B, int N, int npoint, const int64_t *start_idx, float init_dist,
                        int64_t *out_idx, float *out_xyz, float quant_cube, void *workspace, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(xyz && out_idx, "pcc_fps_f32: null pointer");
    PCC_REQUIRE(B >= 0 && N >= 1 && npoint >= 0, "pcc_fps_f32: bad shape B=%d N=%d npoint=%d", B, N, npoint);
    if (B == 0 || npoint == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (N <= 128) return launch_block<128, 1>(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_xyz, quant_cube, st);
    if (N <= 256) return launch_block<128, 2>(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_xyz, quant_cube, st);
    if (N <= 512) return launch_block<128, 4>(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_xyz, quant_cube, st);
    if (N <= 1024) return launch_block<256, 4>(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_xyz, quant_cube, st);
    // eight points per thread on half the warps: the per-iteration barrier and the second-level reduction get cheaper, the packed
    // update keeps the issue slots (measured: 196 -> 180 us at 64 x 2048 -> 512, 592 -> 535 us at 64 x 4096 -> 1024)
    if (N <= 2048) return launch_block<256, 8>(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_xyz, quant_cube, st);
    if (N <= 4096) return launch_block<512, 8>(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_xyz, quant_cube, st);
    // (sixteen points per thread on 16 warps for 4097 .. 8192 points: 75.9 vs 74.2 us at 32 x 8192 -> 64 -- not kept)
    if (N <= 8192) return launch_block<1024, 8>(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_xyz, quant_cube, st);

    PCC_REQUIRE(workspace, "pcc_fps_f32: N=%d needs a workspace of pcc_fps_workspace_bytes()", N);
    // scene scale: Morton buckets with exact skipping, one CTA per cloud (fps_bucket.cu), after a head of co-resident iterations
    if (fps_bucket_takes(B, N, npoint))
        return fps_bucket_run(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_xyz, quant_cube, workspace, st);
    return fps_grid_run(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_xyz, quant_cube, npoint, nullptr, workspace, st);
}
