// Brute-force K nearest neighbours for sm_100a, bit-exact to PyTorch3D's CPU knn_points
// (call sites /root/reference/train.py:185, compress.py:71, pn_kit.py:190, pppe_pcd_ae.py:599, eval.py:132).
//
// Result definition: the K smallest (d2, idx) pairs in lexicographic order, ascending -- exactly what the
// upstream max-heap of (dist, idx) tuples produces.  Because the result is fully determined by the set of
// un-fused d2 values, any selection algorithm gives identical output; two are used:
//
// knn_warp_kernel  (any 1 <= K <= 1024): one WARP per query, candidates staged in shared-memory tiles shared
//   by the CTA's warps.  Each lane evaluates one candidate per step and packs (d2 bits, idx) into a u64 key
//   whose unsigned order is the lexicographic order.  Keys below the current K-th best are appended (ballot +
//   popc compaction) to a per-warp shared buffer of 2*Kp keys; when it fills, a warp-level bitonic sort keeps
//   the best K and tightens the threshold.  After warm-up almost every step is rejected: the loop takes 128
//   candidates per step (one LDS.128 per candidate from an (x, y, z, -) tile), compares the smallest of a
//   lane's four d2 bit patterns with the threshold's and leaves on one vote -- 12 instructions per 32 pairs
//   (ncu, 7812 x 1M, K = 256: the earlier one-ballot-per-32 loop with an SoA tile executed 47).
//
// knn_thread_kernel (K <= 32, small candidate sets, many queries -- the in-patch 256x256 K=16 search of
//   pn_kit.SetAbstraction): one THREAD per query, sorted top-K in registers, candidates read as float4
//   shared-memory broadcasts, results staged through shared memory so global writes are coalesced.
//
// Both kernels are CUDA-core / shared-memory bound: inputs are 12 B per point, read once per CTA.
#include <stdlib.h>

#include "grid.cuh"
#include "pcc_common.cuh"

namespace pcc {

constexpr int KNN_TILE = 1024;

// ---- warp-per-query kernel -------------------------------------------------------------------------------
__device__ __forceinline__ void warp_bitonic_sort(unsigned long long *keys, int cap) {
    const unsigned lane = lane_id();
    for (int k = 2; k <= cap; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (cap >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const bool up = (i & k) == 0;
                const unsigned long long a = keys[i], b = keys[l];
                if ((a > b) == up) {
                    keys[i] = b;
                    keys[l] = a;
                }
            }
            __syncwarp();
        }
    }
}

// Keep the K smallest of keys[0..count): pad to cap, sort, return min(count, K).
__device__ __forceinline__ int warp_select(unsigned long long *keys, int count, int cap, int K) {
    for (int t = count + lane_id(); t < cap; t += 32) keys[t] = KEY_MAX;
    __syncwarp();
    warp_bitonic_sort(keys, cap);
    return count < K ? count : K;
}

// Cheaper than sorting when the buffer fills (cap <= 512): find T = the K-th smallest d2 BIT PATTERN of keys[0..count) by
// bisection over the lane's register copy of the high words, then compact the keys with d2 <= T in place (order is irrelevant
// until the final sort).  Ties at T are all kept, so the new threshold (T, 0xffffffff) stays conservative; returns -1 when ties
// leave no room (the caller then falls back to the exact sort).  ~1 k instructions against ~8 k for the 512-key bitonic sort,
// which matters twice: the warps of a CTA meet at a barrier every tile, so one warp's selection stalls the other seven.
constexpr int PRUNE_NPL = 16;
__device__ __forceinline__ int warp_prune(unsigned long long *keys, int count, int cap, int K, unsigned &T_out) {
    const unsigned lane = lane_id();
    unsigned hi[PRUNE_NPL];
    unsigned vmin = 0xffffffffu, vmax = 0u;
#pragma unroll
    for (int i = 0; i < PRUNE_NPL; ++i) {
        const int t = i * 32 + static_cast<int>(lane);
        hi[i] = 0xffffffffu;
        if (t < count) {
            hi[i] = static_cast<unsigned>(keys[t] >> 32);
            vmin = min(vmin, hi[i]);
            vmax = max(vmax, hi[i]);
        }
    }
    unsigned lo = __reduce_min_sync(FULL_MASK, vmin), up = __reduce_max_sync(FULL_MASK, vmax);
    while (lo < up) {   // smallest T with #(hi <= T) >= K; count >= K valid entries, so T <= up
        const unsigned mid = lo + ((up - lo) >> 1);
        int c = 0;
#pragma unroll
        for (int i = 0; i < PRUNE_NPL; ++i) c += hi[i] <= mid ? 1 : 0;
        c = __reduce_add_sync(FULL_MASK, c);
        if (c >= K) up = mid;
        else lo = mid + 1u;
    }
    const unsigned T = lo;
    int kept = 0;
#pragma unroll
    for (int i = 0; i < PRUNE_NPL; ++i) kept += hi[i] <= T ? 1 : 0;
    kept = __reduce_add_sync(FULL_MASK, kept);
    if (kept > cap - 64) return -1;
    int out = 0;
    for (int c0 = 0; c0 < count; c0 += 32) {
        const int t = c0 + static_cast<int>(lane);
        unsigned long long key = KEY_MAX;
        if (t < count) key = keys[t];
        const bool keep = t < count && static_cast<unsigned>(key >> 32) <= T;
        const unsigned m = __ballot_sync(FULL_MASK, keep);
        __syncwarp();
        if (keep) keys[out + __popc(m & ((1u << lane) - 1u))] = key;   // out + rank <= t: never ahead of the reads
        out += __popc(m);
        __syncwarp();
    }
    T_out = T;
    return out;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
knn_warp_kernel(const float *__restrict__ q, const float *__restrict__ p, int P1, int P2, int K, int cap,
                float *__restrict__ out_d2, int64_t *__restrict__ out_idx, float *__restrict__ out_nn,
                int centre_sub, float nn_scale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *tile = reinterpret_cast<float4 *>(smem_raw);                                  // (x, y, z, -) per candidate
    unsigned long long *all_keys = reinterpret_cast<unsigned long long *>(tile + KNN_TILE);

    const int b = blockIdx.y;
    const unsigned lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int qi = blockIdx.x * WARPS + warp;
    const bool active = qi < P1;
    const float *pc = p + static_cast<size_t>(b) * P2 * 3;
    unsigned long long *keys = all_keys + static_cast<size_t>(warp) * cap;

    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        const float *qp = q + (static_cast<size_t>(b) * P1 + qi) * 3;
        qx = qp[0];
        qy = qp[1];
        qz = qp[2];
    }
    int count = 0;
    unsigned long long thresh = KEY_MAX;
    unsigned thresh_hi = 0xffffffffu;   // d2 bit pattern of the threshold: d2 >= +0, so unsigned order = value order

    // exact step for 32 candidates: append the keys below the threshold, reselect when the buffer is full
    auto consider = [&](unsigned d2_bits, unsigned idx, bool valid) {
        const unsigned long long key = valid ? ((static_cast<unsigned long long>(d2_bits) << 32) | idx) : KEY_MAX;
        const bool pass = key < thresh;
        const unsigned m = __ballot_sync(FULL_MASK, pass);
        if (m == 0u) return;
        if (pass) keys[count + __popc(m & ((1u << lane) - 1u))] = key;
        count += __popc(m);
        if (count + 32 > cap) {
            __syncwarp();
            int kept = -1;
            unsigned T = 0u;
            if (cap <= 32 * PRUNE_NPL) kept = warp_prune(keys, count, cap, K, T);   // count > cap - 32 >= K here
            if (kept >= 0) {
                count = kept;
                thresh = (static_cast<unsigned long long>(T) << 32) | 0xffffffffull;
            } else {
                count = warp_select(keys, count, cap, K);
                thresh = count >= K ? keys[K - 1] : KEY_MAX;
            }
            thresh_hi = static_cast<unsigned>(thresh >> 32);
        }
    };

    for (int t0 = 0; t0 < P2; t0 += KNN_TILE) {
        const int tn = min(KNN_TILE, P2 - t0);
        __syncthreads();  // previous tile fully consumed
        for (int pt = threadIdx.x; pt < tn; pt += WARPS * 32) {
            const float *src = pc + static_cast<size_t>(t0 + pt) * 3;
            tile[pt] = make_float4(src[0], src[1], src[2], 0.0f);
        }
        __syncthreads();
        if (!active) continue;
        // 128 candidates per step, one LDS.128 each; a step whose four d2 are all above the threshold's d2 costs one vote
        const int full = tn & ~127;
        for (int c0 = 0; c0 < full; c0 += 128) {
            const float4 a0 = tile[c0 + lane], a1 = tile[c0 + 32 + lane], a2 = tile[c0 + 64 + lane], a3 = tile[c0 + 96 + lane];
            const unsigned d0 = __float_as_uint(dist2_rn(qx, qy, qz, a0.x, a0.y, a0.z));
            const unsigned d1 = __float_as_uint(dist2_rn(qx, qy, qz, a1.x, a1.y, a1.z));
            const unsigned d2 = __float_as_uint(dist2_rn(qx, qy, qz, a2.x, a2.y, a2.z));
            const unsigned d3 = __float_as_uint(dist2_rn(qx, qy, qz, a3.x, a3.y, a3.z));
            const unsigned dmin = min(min(d0, d1), min(d2, d3));
            if (!__any_sync(FULL_MASK, dmin <= thresh_hi)) continue;
            const unsigned base = static_cast<unsigned>(t0 + c0) + lane;
            consider(d0, base, true);
            consider(d1, base + 32u, true);
            consider(d2, base + 64u, true);
            consider(d3, base + 96u, true);
        }
        for (int c0 = full; c0 < tn; c0 += 32) {
            const int j = c0 + lane;
            const bool valid = j < tn;
            const float4 a = tile[valid ? j : 0];
            consider(__float_as_uint(dist2_rn(qx, qy, qz, a.x, a.y, a.z)), static_cast<unsigned>(t0 + j), valid);
        }
    }
    if (!active) return;
    __syncwarp();
    count = warp_select(keys, count, cap, K);

    const size_t obase = (static_cast<size_t>(b) * P1 + qi) * K;
    for (int k = lane; k < K; k += 32) {
        float d = 0.0f;
        unsigned idx = 0u;
        if (k < count) {
            const unsigned long long key = keys[k];
            d = key_d2(key);
            idx = key_idx(key);
        }
        if (out_d2) out_d2[obase + k] = d;
        if (out_idx) out_idx[obase + k] = static_cast<int64_t>(idx);
        if (out_nn) {
            float nx = pc[static_cast<size_t>(idx) * 3 + 0], ny = pc[static_cast<size_t>(idx) * 3 + 1],
                  nz = pc[static_cast<size_t>(idx) * 3 + 2];
            if (centre_sub) {
                nx = __fsub_rn(nx, qx);
                ny = __fsub_rn(ny, qy);
                nz = __fsub_rn(nz, qz);
            }
            if (nn_scale != 1.0f) {
                nx = __fmul_rn(nx, nn_scale);
                ny = __fmul_rn(ny, nn_scale);
                nz = __fmul_rn(nz, nn_scale);
            }
            float *o = out_nn + (obase + k) * 3;
            o[0] = nx;
            o[1] = ny;
            o[2] = nz;
        }
    }
}

// ---- warp-per-query kernel over a uniform grid (scene scale: P2 in the 10^5 .. 10^7 range) ------------------------------------
// Same result as knn_warp_kernel, bit for bit (same un-fused d2, same (d2, idx) keys, same selection code); what changes is
// which candidates are evaluated.  The candidate cloud is sorted by cell of a G^3 grid (grid_build_kernel, chamfer_grid.cu);
// the warp scans the query's own cell, then shell after shell of cells around it (each row of a shell is one contiguous range
// of the sorted array, or its two end cells), and stops when the K-th smallest d2 found so far is smaller than the squared
// distance from the query to the nearest face of the scanned block that is still inside the grid (minus GridInfo::margin, with
// the relative slack the Chamfer search uses): every point outside the block is then farther than all K kept -- ties included.
// 7812 queries x 1M points, K = 256: ~1-3 k candidates per query instead of 10^6.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
knn_grid_kernel(const float *__restrict__ q, const float *__restrict__ p, const float4 *__restrict__ sorted,
                const unsigned *__restrict__ starts, const GridInfo *__restrict__ info, int P1, int P2, int K, int cap,
                float *__restrict__ out_d2, int64_t *__restrict__ out_idx, float *__restrict__ out_nn, int centre_sub, float nn_scale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *all_keys = reinterpret_cast<unsigned long long *>(smem_raw);
    const int b = blockIdx.y;
    const unsigned lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int qi = blockIdx.x * WARPS + warp;
    if (qi >= P1) return;                       // warps are independent: no block-level barrier below
    const float *pc = p + static_cast<size_t>(b) * P2 * 3;
    const float4 *cand = sorted + static_cast<size_t>(b) * P2;
    const GridInfo g = info[b];
    const int G = g.G;
    const unsigned *st = starts + static_cast<size_t>(b) * (G * G * G + 1);
    unsigned long long *keys = all_keys + static_cast<size_t>(warp) * cap;
    const float *qp = q + (static_cast<size_t>(b) * P1 + qi) * 3;
    const float qx = qp[0], qy = qp[1], qz = qp[2];
    int count = 0;
    unsigned long long thresh = KEY_MAX;
    unsigned thresh_hi = 0xffffffffu;

    auto reselect = [&]() {   // keep the K smallest of the buffer (ties at the K-th d2 kept when they fit), tighten the threshold
        int kept = -1;
        unsigned T = 0u;
        __syncwarp();
        if (cap <= 32 * PRUNE_NPL && count >= K) kept = warp_prune(keys, count, cap, K, T);
        if (kept >= 0) {
            count = kept;
            thresh = (static_cast<unsigned long long>(T) << 32) | 0xffffffffull;
        } else {
            count = warp_select(keys, count, cap, K);
            thresh = count >= K ? keys[K - 1] : KEY_MAX;
        }
        thresh_hi = static_cast<unsigned>(thresh >> 32);
    };
    auto consider = [&](unsigned d2_bits, unsigned idx, bool valid) {
        const unsigned long long key = valid ? ((static_cast<unsigned long long>(d2_bits) << 32) | idx) : KEY_MAX;
        const bool pass = key < thresh;
        const unsigned m = __ballot_sync(FULL_MASK, pass);
        if (m == 0u) return;
        if (pass) keys[count + __popc(m & ((1u << lane) - 1u))] = key;
        count += __popc(m);
        if (count + 32 > cap) reselect();
    };
    auto scan = [&](unsigned a, unsigned e) {   // candidates [a, e) of the sorted array: one coalesced 16-byte load per lane
        for (unsigned j0 = a; j0 < e; j0 += 32) {
            const unsigned j = j0 + lane;
            const bool valid = j < e;
            const float4 c = __ldg(cand + (valid ? j : a));
            const unsigned d = __float_as_uint(dist2_rn(qx, qy, qz, c.x, c.y, c.z));
            if (!__any_sync(FULL_MASK, valid && d <= thresh_hi)) continue;
            consider(d, __float_as_uint(c.w), valid);
        }
    };

    const int cx = cell_coord(qx, g.mnx, g.inv_h, G), cy = cell_coord(qy, g.mny, g.inv_h, G), cz = cell_coord(qz, g.mnz, g.inv_h, G);
    for (int R = 0; R < G; ++R) {
        const int z0 = max(cz - R, 0), z1 = min(cz + R, G - 1), y0 = max(cy - R, 0), y1 = min(cy + R, G - 1);
        const int xa = max(cx - R, 0), xb = min(cx + R, G - 1);
        for (int z = z0; z <= z1; ++z)
            for (int yy = y0; yy <= y1; ++yy) {
                const unsigned row = static_cast<unsigned>((z * G + yy) * G);
                if (z == cz - R || z == cz + R || yy == cy - R || yy == cy + R) {
                    scan(__ldg(st + row + xa), __ldg(st + row + xb + 1));                       // a face row of the shell
                } else {                                                                         // interior row: its two end cells
                    if (cx - R >= 0) scan(__ldg(st + row + cx - R), __ldg(st + row + cx - R + 1));
                    if (cx + R <= G - 1) scan(__ldg(st + row + cx + R), __ldg(st + row + cx + R + 1));
                }
            }
        // every unscanned point lies beyond a face of the block that is inside the grid
        float bound = INFINITY;
        if (cx - R > 0) bound = fminf(bound, qx - (g.mnx + static_cast<float>(cx - R) * g.h));
        if (cx + R + 1 < G) bound = fminf(bound, (g.mnx + static_cast<float>(cx + R + 1) * g.h) - qx);
        if (cy - R > 0) bound = fminf(bound, qy - (g.mny + static_cast<float>(cy - R) * g.h));
        if (cy + R + 1 < G) bound = fminf(bound, (g.mny + static_cast<float>(cy + R + 1) * g.h) - qy);
        if (cz - R > 0) bound = fminf(bound, qz - (g.mnz + static_cast<float>(cz - R) * g.h));
        if (cz + R + 1 < G) bound = fminf(bound, (g.mnz + static_cast<float>(cz + R + 1) * g.h) - qz);
        if (bound == INFINITY) break;            // the block covers the grid
        bound -= g.margin;
        if (count >= K && bound > 0.0f) {
            reselect();                          // thresh_hi = the K-th smallest d2 so far (its bit pattern)
            if (count >= K && __uint_as_float(thresh_hi) < bound * bound * 0.99998f) break;
        }
    }
    __syncwarp();
    count = warp_select(keys, count, cap, K);

    const size_t obase = (static_cast<size_t>(b) * P1 + qi) * K;
    for (int k = lane; k < K; k += 32) {
        float d = 0.0f;
        unsigned idx = 0u;
        if (k < count) {
            const unsigned long long key = keys[k];
            d = key_d2(key);
            idx = key_idx(key);
        }
        if (out_d2) out_d2[obase + k] = d;
        if (out_idx) out_idx[obase + k] = static_cast<int64_t>(idx);
        if (out_nn) {
            float nx = pc[static_cast<size_t>(idx) * 3 + 0], ny = pc[static_cast<size_t>(idx) * 3 + 1],
                  nz = pc[static_cast<size_t>(idx) * 3 + 2];
            if (centre_sub) {
                nx = __fsub_rn(nx, qx);
                ny = __fsub_rn(ny, qy);
                nz = __fsub_rn(nz, qz);
            }
            if (nn_scale != 1.0f) {
                nx = __fmul_rn(nx, nn_scale);
                ny = __fmul_rn(ny, nn_scale);
                nz = __fmul_rn(nz, nn_scale);
            }
            float *o = out_nn + (obase + k) * 3;
            o[0] = nx;
            o[1] = ny;
            o[2] = nz;
        }
    }
}

// ---- CTA-per-query kernel (P2 <= 8192) ---------------------------------------------------------------------------------
// One CTA (512 threads) per query; every thread keeps the d2 bit patterns of its CPT candidates in registers (candidate
// s*512 + tid).  The K-th smallest (d2, idx) key is found WITHOUT sorting: a bisection on the 32-bit pattern of d2 (each
// step = CPT compares per thread + one redux.sync.add per warp + one __syncthreads, double-buffered partials), stopping
// early when a threshold selects exactly K; exact distance ties at the K-th value are resolved by a second bisection on
// the index.  The K selected keys are compacted into shared memory, ranked by counting (K broadcast compares per key, no
// barriers) and written out in order.  Cost per query ~ 6-8 k cycles regardless of K, against ~0.3 ms of dependent warp
// bitonic sorts in knn_warp_kernel -- this is what makes the B=1 compress latency and the K=256 patching cheap.
constexpr int KB_THREADS = 512;
constexpr int KB_WARPS = KB_THREADS / 32;

__device__ __forceinline__ int block_sum(int v, int (*part)[KB_WARPS], int &buf) {
    const int w = __reduce_add_sync(FULL_MASK, v);
    if (lane_id() == 0) part[buf][threadIdx.x >> 5] = w;
    __syncthreads();
    const int r = lane_id() < KB_WARPS ? part[buf][lane_id()] : 0;
    buf ^= 1;
    return __reduce_add_sync(FULL_MASK, r);
}

// Block-wide bitonic sort of one value per thread (512 threads), ascending in thread order.  Exchanges at distance < 32 are
// warp shuffles; the 10 exchanges at distance >= 32 go through two alternating shared-memory buffers (one barrier each).
template <typename T>
__device__ __forceinline__ T block_sort_512(T v, T *buf0, T *buf1) {
    const unsigned tid = threadIdx.x;
    int flip = 0;
#pragma unroll
    for (unsigned k = 2; k <= KB_THREADS; k <<= 1) {
#pragma unroll
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            T other;
            if (j >= 32) {
                T *b = flip ? buf1 : buf0;
                flip ^= 1;
                b[tid] = v;
                __syncthreads();
                other = b[tid ^ j];
            } else {
                other = __shfl_xor_sync(FULL_MASK, v, j);
            }
            const bool lower = (tid & j) == 0, up = (tid & k) == 0;
            const T lo = v < other ? v : other, hi = v < other ? other : v;
            v = (lower == up) ? lo : hi;
        }
    }
    return v;
}

template <int CPT>
__global__ void __launch_bounds__(KB_THREADS)
knn_block_kernel(const float *__restrict__ q, const float *__restrict__ p, int P1, int P2, int K,
                 float *__restrict__ out_d2, int64_t *__restrict__ out_idx, float *__restrict__ out_nn, int centre_sub,
                 float nn_scale) {
    __shared__ unsigned long long sel[PCC_MAX_KNN_K];
    __shared__ unsigned long long sorted[PCC_MAX_KNN_K];
    __shared__ int part[2][KB_WARPS];
    __shared__ int sel_count;
    const int b = blockIdx.y, qi = blockIdx.x, tid = threadIdx.x;
    const float *pc = p + static_cast<size_t>(b) * P2 * 3;
    const float *qp = q + (static_cast<size_t>(b) * P1 + qi) * 3;
    const float qx = qp[0], qy = qp[1], qz = qp[2];
    if (tid == 0) sel_count = 0;

    unsigned d[CPT];
#pragma unroll
    for (int s = 0; s < CPT; ++s) {
        const int j = s * KB_THREADS + tid;
        d[s] = 0xffffffffu;  // beyond P2: above every real pattern (finite d2 and +inf are <= 0x7f800000)
        if (j < P2) d[s] = __float_as_uint(dist2_rn(qx, qy, qz, pc[static_cast<size_t>(j) * 3 + 0],
                                                    pc[static_cast<size_t>(j) * 3 + 1], pc[static_cast<size_t>(j) * 3 + 2]));
    }
    const int Keff = K < P2 ? K : P2;
    int buf = 0;
    // ---- fast path (K <= 512): selection by two small sorts instead of ~32 bisection passes over all candidates ----
    // The Keff-th smallest of the 512 per-thread minima is an upper bound tau of the Keff-th smallest distance (Keff
    // distinct candidates are <= it), and only ~1.4 Keff candidates pass it (each minimum is the best of CPT values).
    // Those are compacted and sorted as (d2, idx) keys; the first Keff are the answer -- the same set and order as any
    // other exact selection.  Falls through to the bisection when fewer than Keff threads hold a candidate or when ties
    // let more than 512 candidates through.
    bool fast = false;
    if (CPT >= 16 && Keff <= KB_THREADS) {   // with few candidates per thread the bisection passes are cheaper than two sorts (measured)
        unsigned m = d[0];
#pragma unroll
        for (int s = 1; s < CPT; ++s) m = d[s] < m ? d[s] : m;
        const unsigned ms = block_sort_512<unsigned>(m, reinterpret_cast<unsigned *>(sel), reinterpret_cast<unsigned *>(sorted));
        __shared__ unsigned s_tau;
        if (tid == Keff - 1) s_tau = ms;
        __syncthreads();
        const unsigned tau = s_tau;
        int c = 0;
#pragma unroll
        for (int s = 0; s < CPT; ++s) c += d[s] <= tau ? 1 : 0;
        const int n = block_sum(c, part, buf);   // (its barrier also orders the sort's last buffer reads before the writes below)
        if (tau != 0xffffffffu && n <= KB_THREADS) {
            fast = true;
#pragma unroll
            for (int s = 0; s < CPT; ++s) {
                const unsigned j = static_cast<unsigned>(s * KB_THREADS + tid);
                const bool take = d[s] <= tau;
                const unsigned mk = __ballot_sync(FULL_MASK, take);
                if (mk) {
                    int base = 0;
                    if (lane_id() == static_cast<unsigned>(__ffs(mk) - 1)) base = atomicAdd(&sel_count, __popc(mk));
                    base = __shfl_sync(FULL_MASK, base, __ffs(mk) - 1);
                    if (take) sel[base + __popc(mk & ((1u << lane_id()) - 1u))] = (static_cast<unsigned long long>(d[s]) << 32) | j;
                }
            }
            __syncthreads();
            unsigned long long key = tid < n ? sel[tid] : KEY_MAX;
            __syncthreads();   // sel doubles as an exchange buffer of the sort
            key = block_sort_512<unsigned long long>(key, sel, sorted);
            __syncthreads();   // the last exchange may still be reading `sorted`
            if (tid < Keff) sorted[tid] = key;
            __syncthreads();
        }
    }
    if (!fast) {
    // bisection on the distance pattern: smallest v with count(d <= v) >= Keff
    unsigned lo = 0u, hi = 0xfffffffeu;
    bool exact = false;
    while (lo < hi) {
        const unsigned mid = lo + ((hi - lo) >> 1);
        int c = 0;
#pragma unroll
        for (int s = 0; s < CPT; ++s) c += d[s] <= mid ? 1 : 0;
        c = block_sum(c, part, buf);
        if (c == Keff) {
            lo = mid;
            exact = true;
            break;
        }
        if (c > Keff) hi = mid; else lo = mid + 1u;
    }
    const unsigned v = lo;
    unsigned tie_max = 0xffffffffu;  // candidates with d == v are taken while idx <= tie_max
    if (!exact) {
        int c = 0;
#pragma unroll
        for (int s = 0; s < CPT; ++s) c += d[s] < v ? 1 : 0;
        const int need = Keff - block_sum(c, part, buf);  // >= 1 ties to take, lowest indices first
        unsigned tlo = 0u, thi = static_cast<unsigned>(P2 - 1);
        while (tlo < thi) {
            const unsigned mid = tlo + ((thi - tlo) >> 1);
            int t = 0;
#pragma unroll
            for (int s = 0; s < CPT; ++s) t += (d[s] == v && static_cast<unsigned>(s * KB_THREADS + tid) <= mid) ? 1 : 0;
            t = block_sum(t, part, buf);
            if (t >= need) thi = mid; else tlo = mid + 1u;
        }
        tie_max = tlo;
    }
    // compaction of exactly Keff keys (unordered), one shared atomic per warp and slot
#pragma unroll
    for (int s = 0; s < CPT; ++s) {
        const unsigned j = static_cast<unsigned>(s * KB_THREADS + tid);
        const bool take = d[s] < v || (d[s] == v && j <= tie_max);
        const unsigned m = __ballot_sync(FULL_MASK, take);
        if (m) {
            int base = 0;
            if (lane_id() == static_cast<unsigned>(__ffs(m) - 1)) base = atomicAdd(&sel_count, __popc(m));
            base = __shfl_sync(FULL_MASK, base, __ffs(m) - 1);
            if (take) sel[base + __popc(m & ((1u << lane_id()) - 1u))] = (static_cast<unsigned long long>(d[s]) << 32) | j;
        }
    }
    __syncthreads();
    // rank by counting: keys are unique, so ranks are a permutation of 0..Keff-1
    for (int e = tid; e < Keff; e += KB_THREADS) {
        const unsigned long long key = sel[e];
        int rank = 0;
        for (int o = 0; o < Keff; ++o) rank += sel[o] < key ? 1 : 0;
        sorted[rank] = key;
    }
    __syncthreads();
    }  // !fast
    const size_t obase = (static_cast<size_t>(b) * P1 + qi) * K;
    for (int k = tid; k < K; k += KB_THREADS) {
        float dd = 0.0f;
        unsigned idx = 0u;
        if (k < Keff) {
            dd = key_d2(sorted[k]);
            idx = key_idx(sorted[k]);
        }
        if (out_d2) out_d2[obase + k] = dd;
        if (out_idx) out_idx[obase + k] = static_cast<int64_t>(idx);
    }
    if (out_nn) {
        for (int e = tid; e < K * 3; e += KB_THREADS) {
            const int k = e / 3, c = e - k * 3;
            const unsigned idx = k < Keff ? key_idx(sorted[k]) : 0u;
            float val = pc[static_cast<size_t>(idx) * 3 + c];
            if (centre_sub) val = __fsub_rn(val, c == 0 ? qx : (c == 1 ? qy : qz));
            if (nn_scale != 1.0f) val = __fmul_rn(val, nn_scale);
            out_nn[obase * 3 + e] = val;
        }
    }
}

template <int CPT>
static int launch_block(const float *q, const float *p, int B, int P1, int P2, int K, float *out_d2, int64_t *out_idx,
                        float *out_nn, int centre_sub, float nn_scale, cudaStream_t st) {
    knn_block_kernel<CPT><<<dim3(P1, B), KB_THREADS, 0, st>>>(q, p, P1, P2, K, out_d2, out_idx, out_nn, centre_sub, nn_scale);
    return check_launch("knn_block_kernel");
}

// ---- thread-per-query kernel ---------------------------------------------------------------------------------
constexpr int KT_THREADS = 256;
constexpr int KT_PENDING = 32;

template <int KT>
__global__ void __launch_bounds__(KT_THREADS, KT <= 16 ? 4 : 2)
knn_thread_kernel(const float *__restrict__ q, const float *__restrict__ p, int P1, int P2, int K,
                  float *__restrict__ out_d2, int64_t *__restrict__ out_idx, float *__restrict__ out_nn,
                  int centre_sub, float nn_scale) {
    constexpr int STAGE_LD = KT + 1;
    constexpr int TILE_BYTES = KNN_TILE * 16;
    constexpr int STAGE_BYTES = KT_THREADS * STAGE_LD * 4;
    __shared__ __align__(16) unsigned char smem_raw[TILE_BYTES > STAGE_BYTES ? TILE_BYTES : STAGE_BYTES];
    __shared__ float sq[KT_THREADS * 3];
    constexpr int PB = KT <= 16 ? KT_PENDING : KT_PENDING / 2;  // pending queue depth: candidate indices only (the distance is re-evaluated when drained)
    __shared__ unsigned short pend_i[PB][KT_THREADS];
    float4 *tile = reinterpret_cast<float4 *>(smem_raw);
    unsigned *stage = reinterpret_cast<unsigned *>(smem_raw);

    const int b = blockIdx.y;
    const int q0 = blockIdx.x * KT_THREADS;
    const int qi = q0 + threadIdx.x;
    const bool active = qi < P1;
    const float *pc_ = p + static_cast<size_t>(b) * P2 * 3;

    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        const float *qp = q + (static_cast<size_t>(b) * P1 + qi) * 3;
        qx = qp[0];
        qy = qp[1];
        qz = qp[2];
    }
    sq[threadIdx.x * 3 + 0] = qx;
    sq[threadIdx.x * 3 + 1] = qy;
    sq[threadIdx.x * 3 + 2] = qz;


    // the candidate set is one shared-memory tile (the dispatch guarantees P2 <= KNN_TILE)
    const int tn = P2;
    __syncthreads();
    for (int pt = threadIdx.x; pt < tn; pt += KT_THREADS) {
        const float *s = pc_ + static_cast<size_t>(pt) * 3;
        tile[pt] = make_float4(s[0], s[1], s[2], 0.f);
    }
    __syncthreads();
    float worst = __int_as_float(0x7f800000);
    {
        if (tn >= 64) {
            // A-priori threshold: the KT-th smallest of G >= KT group minima is >= the KT-th smallest distance overall (KT
            // distinct candidates are <= it).  With 32 strided groups of 8 only ~8 % of the candidates pass it (rank ~21 of
            // 256 for K = 16), so one extra pass of 9 instructions per pair + a 32-value register sort replaces the whole
            // threshold warm-up and most of its ~85-instruction sorted insertions.  Groups are strided (j mod 32): the
            // callers' candidates are often ordered (kNN patches come sorted by distance from the patch centre), and
            // contiguous groups would then be radial shells with useless minima.
            constexpr int G = 32;
            float m[G];
#pragma unroll
            for (int g0 = 0; g0 < G; ++g0) m[g0] = __int_as_float(0x7f800000);
            const int full = tn / G * G;
            for (int j0 = 0; j0 < full; j0 += G) {
#pragma unroll
                for (int g0 = 0; g0 < G; ++g0) {
                    const float4 c = tile[j0 + g0];
                    m[g0] = fminf(m[g0], dist2_rn(qx, qy, qz, c.x, c.y, c.z));
                }
            }
            // bitonic sort of the 32 minima in registers (values only: one FMNMX pair per compare-exchange)
#pragma unroll
            for (int k = 2; k <= G; k <<= 1) {
#pragma unroll
                for (int jj = k >> 1; jj > 0; jj >>= 1) {
#pragma unroll
                    for (int i = 0; i < G; ++i) {
                        const int l = i ^ jj;
                        if (l > i) {
                            const float lo = fminf(m[i], m[l]), hi = fmaxf(m[i], m[l]);
                            const bool up = (i & k) == 0;
                            m[i] = up ? lo : hi;
                            m[l] = up ? hi : lo;
                        }
                    }
                }
            }
            const float tau = m[KT - 1 < G ? KT - 1 : G - 1];   // KT-th smallest (KT <= 32)
            worst = __uint_as_float(__float_as_uint(tau) + 1u);   // next float up: `d < worst` admits d == tau (d2 >= +0)
        }
    }
    float dl[KT];
    unsigned il[KT];
#pragma unroll
    for (int s = 0; s < KT; ++s) {
        dl[s] = __int_as_float(0x7f800000);  // +inf: unfilled
        il[s] = 0u;
    }
    int pc = 0;
    unsigned short *pq = &pend_i[0][threadIdx.x];
    // Sorted insertion costs ~5 instructions per list slot and, thread-per-query, a warp pays it whenever ANY lane
    // inserts -- i.e. for almost every candidate.  Candidates that beat the (possibly stale) K-th best are therefore only
    // parked in a per-thread shared-memory queue (2 predicated stores); the warp drains all queues, in candidate order,
    // when one of them is full.  The result is identical to immediate insertion: the list changes only through the same
    // insertions in the same order, and a stale threshold only lets extra candidates through to the exact re-test.
    auto insert = [&](float d, unsigned j) {
        if (d < dl[KT - 1]) {  // strict: an equal distance with a larger index never displaces
            dl[KT - 1] = d;
            il[KT - 1] = j;
#pragma unroll
            for (int s = KT - 1; s > 0; --s) {
                const bool sw = dl[s] < dl[s - 1];  // strict keeps earlier (lower) indices first on ties
                const float dlo = sw ? dl[s] : dl[s - 1], dhi = sw ? dl[s - 1] : dl[s];
                const unsigned ilo = sw ? il[s] : il[s - 1], ihi = sw ? il[s - 1] : il[s];
                dl[s - 1] = dlo;
                dl[s] = dhi;
                il[s - 1] = ilo;
                il[s] = ihi;
            }
        }
    };
    auto drain = [&]() {
        const int n = __reduce_max_sync(FULL_MASK, pc);
        for (int s = 0; s < n; ++s) {
            if (s < pc) {
                const unsigned j = pend_i[s][threadIdx.x];
                const float4 c = tile[j];
                insert(dist2_rn(qx, qy, qz, c.x, c.y, c.z), j);   // bit-identical to the value that passed the test
            }
        }
        pc = 0;
        pq = &pend_i[0][threadIdx.x];
        worst = fminf(worst, dl[KT - 1]);   // the list may not be full yet: keep the a-priori bound
    };

    // main pass: 8 candidates per queue-occupancy vote (a queue never gains more than 8 entries in between)
    constexpr int UN = 8;
    int j = 0;
    for (; j + UN <= tn; j += UN) {
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const float4 c = tile[j + u];
            const float d = dist2_rn(qx, qy, qz, c.x, c.y, c.z);
            if (d < worst) {
                *pq = static_cast<unsigned short>(j + u);
                pq += KT_THREADS;
                ++pc;
            }
        }
        if (__any_sync(FULL_MASK, pc > PB - UN)) drain();
    }
    for (; j < tn; ++j) {
        const float4 c = tile[j];
        const float d = dist2_rn(qx, qy, qz, c.x, c.y, c.z);
        if (d < worst) {
            *pq = static_cast<unsigned short>(j);
            pq += KT_THREADS;
            ++pc;
        }
        if (__any_sync(FULL_MASK, pc > PB - UN)) drain();
    }
    drain();
    __syncthreads();  // tile no longer needed: reuse as output staging

    const int nq = min(KT_THREADS, P1 - q0);
    const int valid = P2 < K ? P2 : K;
    const size_t obase = (static_cast<size_t>(b) * P1 + q0) * K;
    // distances
#pragma unroll
    for (int s = 0; s < KT; ++s) stage[threadIdx.x * STAGE_LD + s] = s < valid ? __float_as_uint(dl[s]) : 0u;
    __syncthreads();
    if (out_d2) {
        for (int e = threadIdx.x; e < nq * K; e += KT_THREADS) {
            const int ql = e / K, k = e - ql * K;
            out_d2[obase + e] = __uint_as_float(stage[ql * STAGE_LD + k]);
        }
    }
    __syncthreads();
    // indices (+ gathered neighbours)
#pragma unroll
    for (int s = 0; s < KT; ++s) stage[threadIdx.x * STAGE_LD + s] = s < valid ? il[s] : 0u;
    __syncthreads();
    if (out_idx) {
        for (int e = threadIdx.x; e < nq * K; e += KT_THREADS) {
            const int ql = e / K, k = e - ql * K;
            out_idx[obase + e] = static_cast<int64_t>(stage[ql * STAGE_LD + k]);
        }
    }
    if (out_nn && K == KT) {   // compile-time divisors
        for (int e = threadIdx.x; e < nq * KT * 3; e += KT_THREADS) {
            const int pk = e / 3, c = e - pk * 3;
            const int ql = pk / KT, k = pk - ql * KT;
            float v = pc_[static_cast<size_t>(stage[ql * STAGE_LD + k]) * 3 + c];
            if (centre_sub) v = __fsub_rn(v, sq[ql * 3 + c]);
            if (nn_scale != 1.0f) v = __fmul_rn(v, nn_scale);
            out_nn[obase * 3 + e] = v;
        }
    } else if (out_nn) {
        for (int e = threadIdx.x; e < nq * K * 3; e += KT_THREADS) {
            const int pk = e / 3, c = e - pk * 3;
            const int ql = pk / K, k = pk - ql * K;
            float v = pc_[static_cast<size_t>(stage[ql * STAGE_LD + k]) * 3 + c];
            if (centre_sub) v = __fsub_rn(v, sq[ql * 3 + c]);
            if (nn_scale != 1.0f) v = __fmul_rn(v, nn_scale);
            out_nn[obase * 3 + e] = v;
        }
    }
}


// ---- thread-per-query kernel, filter form (K <= 16) ------------------------------------------------------------------
// Same contract and the same exact result as knn_thread_kernel; what changes is how few instructions a (query, candidate)
// pair costs.  The exact d2 takes 8 dependent-rounding FP32 operations + a shared-memory load; here a pair is first seen
// through B = |c|^2 - 2 q.c (3 FFMA on a (x, y, z, |c|^2) tile: d2 = B + |q|^2 up to a PROVEN rounding bound E), twice:
//   pass 1  32 strided group minima of B, sorted in registers; the KT-th smallest, tau, satisfies
//           (K-th smallest exact d2) <= tau + |q|^2 + E                       (KT candidates have B <= tau)
//   pass 2  a candidate can be one of the K nearest only if B <= tau + 2 E: its index is parked in a per-thread queue
//           (~21 of 256 pass for K = 16)
//   final   the parked candidates' EXACT d2 (dist2_rn, the reference arithmetic) are packed with their queue slot into
//           32-bit keys ((d2 bits - base) << 5 | slot: slots are in index order, so the unsigned key order is the
//           (d2, idx) order) and sorted by a 32-key register network -- no data-dependent control flow, no insertion chain.
// Error bound (u = 2^-24, R = (|q| + |c|)^2 <= 2 (|q|^2 + max|c|^2)): the three FFMA and the two squared norms contribute
// <= 9.2 u R against the real-number d2, the reference's rounded d2 differs from it by <= 5.1 u R, so
// |B + |q|^2 - d2_rn| <= 2^-20 R <= 2^-19 (|q|^2 + max|c|^2); the kernel uses twice that.  A warp whose queues overflow
// (heavy ties, tiny candidate sets) or whose keys do not fit 27 bits of d2 range falls back to the exact insertion scan.
template <int KT>
__global__ void __launch_bounds__(KT_THREADS, 4)
knn_filter_kernel(const float *__restrict__ q, const float *__restrict__ p, int P1, int P2, int K,
                  float *__restrict__ out_d2, int64_t *__restrict__ out_idx, float *__restrict__ out_nn,
                  unsigned char *__restrict__ out_idx8, int centre_sub, float nn_scale) {
    static_assert(KT <= 16, "filter form: K <= 16");
    constexpr int STAGE_LD = KT + 1;
    constexpr int PB = 64;                                    // queue depth: 32 keys of the sorting network + overflow slots
    constexpr int FT = 256;                                   // candidates per search (byte queue entries)
    constexpr int TILE_BYTES = FT * 16;
    constexpr int STAGE_BYTES = KT_THREADS * STAGE_LD * 4;
    __shared__ __align__(16) unsigned char smem_raw[TILE_BYTES > STAGE_BYTES ? TILE_BYTES : STAGE_BYTES];
    __shared__ float sq[KT_THREADS * 3];
    __shared__ unsigned char pend_i[PB][KT_THREADS];
    __shared__ float s_cmax[KT_THREADS / 32];
    float4 *tile = reinterpret_cast<float4 *>(smem_raw);
    unsigned *stage = reinterpret_cast<unsigned *>(smem_raw);

    const int b = blockIdx.y;
    const int q0 = blockIdx.x * KT_THREADS;
    const int qi = q0 + threadIdx.x;
    const bool active = qi < P1;
    const float *pc_ = p + static_cast<size_t>(b) * P2 * 3;

    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        const float *qp = q + (static_cast<size_t>(b) * P1 + qi) * 3;
        qx = qp[0];
        qy = qp[1];
        qz = qp[2];
    }
    sq[threadIdx.x * 3 + 0] = qx;
    sq[threadIdx.x * 3 + 1] = qy;
    sq[threadIdx.x * 3 + 2] = qz;
    const int tn = P2;   // the dispatch guarantees P2 <= FT
    float cm = 0.f;
    for (int pt = threadIdx.x; pt < tn; pt += KT_THREADS) {
        const float *s = pc_ + static_cast<size_t>(pt) * 3;
        const float x = s[0], y = s[1], z = s[2];
        const float w = fmaf(z, z, fmaf(y, y, x * x));
        tile[pt] = make_float4(x, y, z, w);
        cm = fmaxf(cm, w);
    }
    cm = __uint_as_float(__reduce_max_sync(FULL_MASK, __float_as_uint(cm)));   // w >= +0: the bit patterns order like the values
    if ((threadIdx.x & 31) == 0) s_cmax[threadIdx.x >> 5] = cm;
    __syncthreads();
    float cmax2 = s_cmax[0];
#pragma unroll
    for (int i = 1; i < KT_THREADS / 32; ++i) cmax2 = fmaxf(cmax2, s_cmax[i]);
    const float qn = fmaf(qz, qz, fmaf(qy, qy, qx * qx));
    const float E = fmaf(qn + cmax2, 3.814697265625e-6f /* 2^-18 */, 1e-30f);
    const float mx = -2.0f * qx, my = -2.0f * qy, mz = -2.0f * qz;
    auto approx = [&](const float4 c) { return fmaf(mx, c.x, fmaf(my, c.y, fmaf(mz, c.z, c.w))); };

    float thr = __int_as_float(0x7f800000);   // +inf: everything passes (tiny candidate sets)
    if (tn >= 64) {
        constexpr int G = 32;
        float m[G];
#pragma unroll
        for (int g0 = 0; g0 < G; ++g0) m[g0] = __int_as_float(0x7f800000);
        const int full = tn / G * G;
        for (int j0 = 0; j0 < full; j0 += G) {
#pragma unroll
            for (int g0 = 0; g0 < G; ++g0) m[g0] = fminf(m[g0], approx(tile[j0 + g0]));
        }
#pragma unroll
        for (int k = 2; k <= G; k <<= 1) {
#pragma unroll
            for (int jj = k >> 1; jj > 0; jj >>= 1) {
#pragma unroll
                for (int i = 0; i < G; ++i) {
                    const int l = i ^ jj;
                    if (l > i) {
                        const float lo = fminf(m[i], m[l]), hi = fmaxf(m[i], m[l]);
                        const bool up = (i & k) == 0;
                        m[i] = up ? lo : hi;
                        m[l] = up ? hi : lo;
                    }
                }
            }
        }
        thr = m[KT - 1] + 2.0f * E;
    }

    // pass 2: one bit per candidate that passes (LDS + 3 FFMA + compare + a predicated OR with a constant per pair), 32
    // candidates per word; the set bits of each word are unpacked into the thread's byte queue (indices < 256).  A word that
    // would not fit the PB-deep queue marks the thread as overflowed (exact scan below).
    unsigned char *const pq0 = &pend_i[0][threadIdx.x];
    unsigned char *pq = pq0;
    bool over = false;
    for (int j0 = 0; j0 < tn; j0 += 32) {
        unsigned w = 0u;
        if (j0 + 32 <= tn) {
#pragma unroll
            for (int u = 0; u < 32; ++u)
                if (approx(tile[j0 + u]) <= thr) w |= 1u << u;
        } else {
            for (int u = 0; j0 + u < tn; ++u)
                if (approx(tile[j0 + u]) <= thr) w |= 1u << u;
        }
        if (pq + __popc(w) * KT_THREADS > pq0 + PB * KT_THREADS) {
            over = true;
            w = 0u;
        }
        while (w != 0u) {
            *pq = static_cast<unsigned char>(j0 + __ffs(w) - 1);
            w &= w - 1u;
            pq += KT_THREADS;
        }
    }
    const int pc = static_cast<int>(pq - pq0) / KT_THREADS;

    float dl[KT];
    unsigned il[KT];
    // final selection on the exact distances: the first 32 parked candidates through the sorting network
    constexpr int NS = 32;
    unsigned key[NS];
    unsigned dmax = 0u;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        key[s] = 0xffffffffu;
        if (s < pc) {
            const float4 c = tile[pend_i[s][threadIdx.x]];
            key[s] = __float_as_uint(dist2_rn(qx, qy, qz, c.x, c.y, c.z));
            dmax = max(dmax, key[s]);
        }
    }
    // keys: rel = 0 for d2 == +0, d2 bits - base + 1 otherwise; needs 1 <= rel < 2^27 for every non-zero candidate
    const unsigned base = dmax > 0x07fffffeu ? dmax - 0x07fffffdu : 1u;   // dmax - base + 1 <= 2^27 - 2
    bool fits = !over;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        if (s < pc) {
            const unsigned bits = key[s];
            fits = fits && (bits == 0u || bits >= base);
            key[s] = ((bits == 0u ? 0u : bits - base + 1u) << 5) | static_cast<unsigned>(s);
        }
    }
    auto insert = [&](float d, unsigned j) {   // sorted insertion, strict compares: equal distances keep index order
        if (d < dl[KT - 1]) {
            dl[KT - 1] = d;
            il[KT - 1] = j;
#pragma unroll
            for (int s = KT - 1; s > 0; --s) {
                const bool sw = dl[s] < dl[s - 1];
                const float dlo = sw ? dl[s] : dl[s - 1], dhi = sw ? dl[s - 1] : dl[s];
                const unsigned ilo = sw ? il[s] : il[s - 1], ihi = sw ? il[s - 1] : il[s];
                dl[s - 1] = dlo;
                dl[s] = dhi;
                il[s - 1] = ilo;
                il[s] = ihi;
            }
        }
    };
    if (fits) {
#pragma unroll
        for (int k = 2; k <= NS; k <<= 1) {
#pragma unroll
            for (int jj = k >> 1; jj > 0; jj >>= 1) {
#pragma unroll
                for (int i = 0; i < NS; ++i) {
                    const int l = i ^ jj;
                    if (l > i) {
                        const unsigned lo = min(key[i], key[l]), hi = max(key[i], key[l]);
                        const bool up = (i & k) == 0;
                        key[i] = up ? lo : hi;
                        key[l] = up ? hi : lo;
                    }
                }
            }
        }
#pragma unroll
        for (int s = 0; s < KT; ++s) {
            const unsigned rel = key[s] >> 5;
            const bool have = s < pc;
            dl[s] = have ? __uint_as_float(rel == 0u ? 0u : rel - 1u + base) : __int_as_float(0x7f800000);
            il[s] = have ? pend_i[key[s] & 31u][threadIdx.x] : 0u;
        }
        for (int s = NS; s < pc; ++s) {          // the few candidates beyond the network's 32 keys (index order: ties stay right)
            const unsigned j = pend_i[s][threadIdx.x];
            const float4 c = tile[j];
            insert(dist2_rn(qx, qy, qz, c.x, c.y, c.z), j);
        }
    } else {
        // exact insertion scan over every candidate that passes the filter (heavy ties, key range, tiny sets): rare
#pragma unroll
        for (int s = 0; s < KT; ++s) {
            dl[s] = __int_as_float(0x7f800000);
            il[s] = 0u;
        }
        for (int j = 0; j < tn; ++j) {
            const float4 c = tile[j];
            if (approx(c) <= thr) insert(dist2_rn(qx, qy, qz, c.x, c.y, c.z), static_cast<unsigned>(j));
        }
    }
    __syncthreads();  // tile no longer needed: reuse as output staging

    const int nq = min(KT_THREADS, P1 - q0);
    const int valid = P2 < K ? P2 : K;
    const size_t obase = (static_cast<size_t>(b) * P1 + q0) * K;
    if (out_d2) {
#pragma unroll
        for (int s = 0; s < KT; ++s) stage[threadIdx.x * STAGE_LD + s] = s < valid ? __float_as_uint(dl[s]) : 0u;
        __syncthreads();
        for (int e = threadIdx.x; e < nq * K; e += KT_THREADS) {
            const int ql = e / K, k = e - ql * K;
            out_d2[obase + e] = __uint_as_float(stage[ql * STAGE_LD + k]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int s = 0; s < KT; ++s) stage[threadIdx.x * STAGE_LD + s] = s < valid ? il[s] : 0u;
    __syncthreads();
    if (out_idx) {
        for (int e = threadIdx.x; e < nq * K; e += KT_THREADS) {
            const int ql = e / K, k = e - ql * K;
            out_idx[obase + e] = static_cast<int64_t>(stage[ql * STAGE_LD + k]);
        }
    }
    if (out_idx8 && active) {   // P2 <= 256, K == KT: neighbour indices as bytes, one 8 / 16-byte store per query
        unsigned pk[KT / 4];
#pragma unroll
        for (int i = 0; i < KT / 4; ++i)
            pk[i] = (il[4 * i] & 255u) | ((il[4 * i + 1] & 255u) << 8) | ((il[4 * i + 2] & 255u) << 16) | (il[4 * i + 3] << 24);
        unsigned *dst = reinterpret_cast<unsigned *>(out_idx8 + obase + static_cast<size_t>(threadIdx.x) * KT);
        if constexpr (KT == 16) *reinterpret_cast<uint4 *>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        else *reinterpret_cast<uint2 *>(dst) = make_uint2(pk[0], pk[1]);
    }
    if (out_nn) {
        for (int e = threadIdx.x; e < nq * K * 3; e += KT_THREADS) {
            const int pk = e / 3, c = e - pk * 3;
            const int ql = K == KT ? pk / KT : pk / K, k = pk - ql * K;
            float v = pc_[static_cast<size_t>(stage[ql * STAGE_LD + k]) * 3 + c];
            if (centre_sub) v = __fsub_rn(v, sq[ql * 3 + c]);
            if (nn_scale != 1.0f) v = __fmul_rn(v, nn_scale);
            out_nn[obase * 3 + e] = v;
        }
    }
}

template <int KT>
static int launch_filter(const float *q, const float *p, int B, int P1, int P2, int K, float *out_d2, int64_t *out_idx,
                         float *out_nn, unsigned char *out_idx8, int centre_sub, float nn_scale, cudaStream_t st) {
    dim3 grid((P1 + KT_THREADS - 1) / KT_THREADS, B);
    knn_filter_kernel<KT><<<grid, KT_THREADS, 0, st>>>(q, p, P1, P2, K, out_d2, out_idx, out_nn, out_idx8, centre_sub, nn_scale);
    return check_launch("knn_filter_kernel");
}

template <int KT>
static int launch_thread(const float *q, const float *p, int B, int P1, int P2, int K, float *out_d2,
                         int64_t *out_idx, float *out_nn, int centre_sub, float nn_scale, cudaStream_t st) {
    dim3 grid((P1 + KT_THREADS - 1) / KT_THREADS, B);
    knn_thread_kernel<KT><<<grid, KT_THREADS, 0, st>>>(q, p, P1, P2, K, out_d2, out_idx, out_nn, centre_sub, nn_scale);
    return check_launch("knn_thread_kernel");
}

}  // namespace pcc

PCC_API int pcc_knn_f32(const float *q, const float *p, int B, int P1, int P2, int K, float *out_d2, int64_t *out_idx,
                        float *out_nn, int centre_sub, float nn_scale, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(q && p && (out_d2 || out_idx || out_nn), "pcc_knn_f32: null pointer");
    PCC_REQUIRE(B >= 0 && P1 >= 0 && P2 >= 1, "pcc_knn_f32: bad shape B=%d P1=%d P2=%d", B, P1, P2);
    PCC_REQUIRE(K >= 1 && K <= PCC_MAX_KNN_K, "pcc_knn_f32: K=%d outside [1,%d]", K, PCC_MAX_KNN_K);
    PCC_REQUIRE(B <= 65535, "pcc_knn_f32: B=%d exceeds 65535", B);
    if (B == 0 || P1 == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    const long long nq = static_cast<long long>(B) * P1;
    if (K <= 32 && P2 <= KNN_TILE && nq >= 32768) {
        static const bool old_form = getenv("PCC_KNN_THREAD_OLD") != nullptr;   // A/B switch for the measurements in profiles/
        const bool filt = !old_form && P2 <= 256;   // filter form: candidate sets that fit byte indices (the in-patch searches)
        if (filt && K <= 8) return launch_filter<8>(q, p, B, P1, P2, K, out_d2, out_idx, out_nn, nullptr, centre_sub, nn_scale, st);
        if (filt && K <= 16) return launch_filter<16>(q, p, B, P1, P2, K, out_d2, out_idx, out_nn, nullptr, centre_sub, nn_scale, st);
        if (K <= 8) return launch_thread<8>(q, p, B, P1, P2, K, out_d2, out_idx, out_nn, centre_sub, nn_scale, st);
        if (K <= 16) return launch_thread<16>(q, p, B, P1, P2, K, out_d2, out_idx, out_nn, centre_sub, nn_scale, st);
        return launch_thread<32>(q, p, B, P1, P2, K, out_d2, out_idx, out_nn, centre_sub, nn_scale, st);
    }

    if (P2 <= 16 * KB_THREADS && P1 <= 65535) {
        if (P2 <= 1 * KB_THREADS) return launch_block<1>(q, p, B, P1, P2, K, out_d2, out_idx, out_nn, centre_sub, nn_scale, st);
        if (P2 <= 2 * KB_THREADS) return launch_block<2>(q, p, B, P1, P2, K, out_d2, out_idx, out_nn, centre_sub, nn_scale, st);
        if (P2 <= 4 * KB_THREADS) return launch_block<4>(q, p, B, P1, P2, K, out_d2, out_idx, out_nn, centre_sub, nn_scale, st);
        if (P2 <= 8 * KB_THREADS) return launch_block<8>(q, p, B, P1, P2, K, out_d2, out_idx, out_nn, centre_sub, nn_scale, st);
        return launch_block<16>(q, p, B, P1, P2, K, out_d2, out_idx, out_nn, centre_sub, nn_scale, st);
    }

    int kp = 32;
    while (kp < K) kp <<= 1;
    const int cap = 2 * kp;
    // 8 warps per CTA; fewer when the key buffers would not fit (K = 1024 -> 16 KB per warp).
    constexpr int W = 8;
    const size_t smem = KNN_TILE * sizeof(float4) + static_cast<size_t>(W) * cap * sizeof(unsigned long long);
    static bool attr_set_dev[64] = {false};   // per device: the attribute belongs to the device's copy of the kernel
    int attr_set_d = 0;
    if (cudaGetDevice(&attr_set_d) != cudaSuccess || attr_set_d < 0 || attr_set_d >= 64) attr_set_d = 0;
    if (!attr_set_dev[attr_set_d]) {
        cudaError_t e = cudaFuncSetAttribute(knn_warp_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             KNN_TILE * 16 + W * 2 * PCC_MAX_KNN_K * 8);
        if (e != cudaSuccess) {
            set_error("pcc_knn_f32: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
            return static_cast<int>(e);
        }
        attr_set_dev[attr_set_d] = true;
    }
    dim3 grid((P1 + W - 1) / W, B);
    knn_warp_kernel<W><<<grid, W * 32, smem, st>>>(q, p, P1, P2, K, cap, out_d2, out_idx, out_nn, centre_sub, nn_scale);
    return check_launch("knn_warp_kernel");
}

PCC_API int pcc_knn_patch_u8(const float *patches, int BS, int P, int K, uint8_t *out_idx, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(patches && out_idx, "pcc_knn_patch_u8: null pointer");
    PCC_REQUIRE(BS >= 0 && BS <= 65535 && P >= 1, "pcc_knn_patch_u8: bad shape BS=%d P=%d", BS, P);
    if ((K != 8 && K != 16) || P > 256 || P < K) {
        set_error("pcc_knn_patch_u8: K must be 8 or 16 and K <= P <= 256 (byte indices)");
        return PCC_ERR_UNSUPPORTED;
    }
    PCC_REQUIRE(reinterpret_cast<uintptr_t>(out_idx) % 16 == 0, "pcc_knn_patch_u8: out_idx must be 16-byte aligned");
    if (BS == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (K == 8) return launch_filter<8>(patches, patches, BS, P, P, K, nullptr, nullptr, nullptr, out_idx, 0, 1.0f, st);
    return launch_filter<16>(patches, patches, BS, P, P, K, nullptr, nullptr, nullptr, out_idx, 0, 1.0f, st);
}

/* Scene-scale form: the same result as pcc_knn_f32 through a uniform grid over p (workspace from pcc_knn_grid_workspace_bytes). */
PCC_API int64_t pcc_knn_grid_workspace_bytes(int B, int P2) {
    const int64_t G = 32;
    return static_cast<int64_t>(B) * P2 * 16 + ((static_cast<int64_t>(B) * (G * G * G + 1) * 4 + 15) / 16) * 16 +
           static_cast<int64_t>(B) * static_cast<int64_t>(sizeof(pcc::GridInfo)) + static_cast<int64_t>(B) * 32 * 4 + 64 +
           static_cast<int64_t>(B) * ((G * G * G + 1) + 8) * 4;
}

PCC_API int pcc_knn_grid_f32(const float *q, const float *p, int B, int P1, int P2, int K, float *out_d2, int64_t *out_idx,
                             float *out_nn, int centre_sub, float nn_scale, void *workspace, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(q && p && workspace && (out_d2 || out_idx || out_nn), "pcc_knn_grid_f32: null pointer");
    PCC_REQUIRE(B >= 0 && P1 >= 0 && P2 >= 1 && P2 <= (1 << 24), "pcc_knn_grid_f32: bad shape B=%d P1=%d P2=%d", B, P1, P2);
    PCC_REQUIRE(K >= 1 && K <= PCC_MAX_KNN_K, "pcc_knn_grid_f32: K=%d outside [1,%d]", K, PCC_MAX_KNN_K);
    PCC_REQUIRE(B <= 65535 && reinterpret_cast<uintptr_t>(workspace) % 16 == 0, "pcc_knn_grid_f32: B > 65535 or unaligned workspace");
    if (B == 0 || P1 == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // cells per axis: 32 for scenes; small clouds get ~2 points per cell so that a K = 256 ball spans a handful of shells
    static const int g_env = getenv("PCC_KNN_GRID_G") ? atoi(getenv("PCC_KNN_GRID_G")) : 0;   // tuning knob (profiles/)
    const int G = (g_env >= 4 && g_env <= 32) ? g_env : (P2 >= 200000 ? 32 : P2 >= 30000 ? 24 : 16);
    char *w = static_cast<char *>(workspace);
    float4 *sorted = reinterpret_cast<float4 *>(w);
    w += static_cast<int64_t>(B) * P2 * 16;
    unsigned *starts = reinterpret_cast<unsigned *>(w);
    w += ((static_cast<int64_t>(B) * (G * G * G + 1) * 4 + 15) / 16) * 16;
    GridInfo *info = reinterpret_cast<GridInfo *>(w);
    unsigned *rowmask = reinterpret_cast<unsigned *>(info + B);
    void *scratch = reinterpret_cast<void *>((reinterpret_cast<uintptr_t>(rowmask + static_cast<size_t>(B) * 32) + 15) & ~static_cast<uintptr_t>(15));
    if (int r = grid_build_single(p, B, P2, G, sorted, starts, info, rowmask, scratch, st)) return r;
    int kp = 32;
    while (kp < K) kp <<= 1;
    const int cap = 2 * kp;
    constexpr int W = 8;
    const size_t smem = static_cast<size_t>(W) * cap * sizeof(unsigned long long);
    static bool attr_set_dev[64] = {false};
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) d = 0;
    if (!attr_set_dev[d]) {
        const cudaError_t e = cudaFuncSetAttribute(knn_grid_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, W * 2 * PCC_MAX_KNN_K * 8);
        if (e != cudaSuccess) {
            set_error("pcc_knn_grid_f32: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
            return static_cast<int>(e);
        }
        attr_set_dev[d] = true;
    }
    dim3 grid((P1 + W - 1) / W, B);
    knn_grid_kernel<W><<<grid, W * 32, smem, st>>>(q, p, sorted, starts, info, P1, P2, K, cap, out_d2, out_idx, out_nn, centre_sub, nn_scale);
    return check_launch("knn_grid_kernel");
}
