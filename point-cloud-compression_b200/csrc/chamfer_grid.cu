// Exact nearest-neighbour search through a uniform grid, for the Chamfer distance / D1 PSNR of clouds with 1024..2^24 points.
//
// Same results as the brute-force kernels of chamfer.cu, bit for bit: every candidate that is evaluated goes through the
// same un-fused d2 (dist2_rn) and the same (d2, original index) key, and a cell (or a row of cells) is skipped only when
// its distance to the query provably exceeds the best d2 found so far (with explicit margins against the rounding of the
// cell assignment and of the bound itself), so the minimum over the visited candidates is the minimum over all of them
// and ties still go to the lowest original index.  What changes is the work: 30-1000 candidates per query instead of P2
// (8192): the algorithmic P1*P2 pair evaluations of SURVEY.md 8(d) are mostly pruned.  Replaces the two
// knn_points(K=1) passes of pytorch3d.loss.chamfer_distance (/root/reference/AE.py:67, eval.py:204) and the KD-tree loop
// of eval.py:68-81.  (A variant that culled dense 128-point Morton tiles against 256-query blocks was measured 2x slower
// than brute force: bounding boxes of Morton runs are too loose, and two directions cost two passes.)
//
//   grid_build_kernel   one CTA per (cloud, side): bounding box, G^3 cell histogram in shared memory, exclusive scan,
//                       counting-sort scatter of (x, y, z, original index) -> points sorted by cell (x fastest)
//   grid_nn_kernel      one thread per query, queries taken in *their own* sorted order so a warp's queries are
//                       neighbours in space and walk the same cells.  Step 1: the 3x3x3 block of cells around the query,
//                       which settles it when the best d2 is closer than the block's faces.  Step 2 (far neighbours),
//                       warp-cooperative: an upper bound from a strided sample of the candidates if nothing was found,
//                       then the warp walks the union of its unsettled queries' balls with uniform control flow
#include <stdlib.h>

#include "grid.cuh"
#include "pcc_common.cuh"

namespace pcc {

constexpr int GRID_BUILD_THREADS = 1024;

// grid (B, 2); dynamic smem: (G^3 + 1) u32
// PAIRED: the sorted points are stored two by two, 32 bytes per pair {x0 x1 y0 y1 z0 z1 i0 i1}, every cloud padded to an even
// count with a point at +inf -- the operand layout of the packed-fp32 distance loop of grid_nn_d2_kernel
template <bool PAIRED>
__global__ void __launch_bounds__(GRID_BUILD_THREADS)
grid_build_kernel(const float *__restrict__ x, const float *__restrict__ y, int P1, int P2, int G, float4 *__restrict__ sorted,
                  unsigned *__restrict__ starts, GridInfo *__restrict__ info, unsigned *__restrict__ rowmask) {
    extern __shared__ unsigned cnt[];
    __shared__ float red[6][32];
    __shared__ unsigned wsum[32];
    __shared__ GridInfo gi;
    const int b = blockIdx.x, side = blockIdx.y, B = gridDim.x;
    const int P = side ? P2 : P1;
    const float *pts = side ? y + static_cast<size_t>(b) * P2 * 3 : x + static_cast<size_t>(b) * P1 * 3;
    const size_t S1 = PAIRED ? (static_cast<size_t>(P1) + 1) & ~static_cast<size_t>(1) : P1;
    const size_t S2 = PAIRED ? (static_cast<size_t>(P2) + 1) & ~static_cast<size_t>(1) : P2;
    float4 *out = sorted + (side ? static_cast<size_t>(B) * S1 + static_cast<size_t>(b) * S2 : static_cast<size_t>(b) * S1);
    const int ncell = G * G * G;
    unsigned *st_out = starts + (static_cast<size_t>(side) * B + b) * (ncell + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- bounding box ----
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = tid; i < P; i += GRID_BUILD_THREADS) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = pts[static_cast<size_t>(i) * 3 + a];
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(FULL_MASK, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(FULL_MASK, mx[a], o));
        }
        if (lane == 0) {
            red[a][warp] = mn[a];
            red[3 + a][warp] = mx[a];
        }
    }
    for (int i = tid; i <= ncell; i += GRID_BUILD_THREADS) cnt[i] = 0u;
    __syncthreads();
    if (warp == 0) {
        float ext = 0.0f, lo[3], maxabs = 0.0f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float m0 = red[a][lane], m1 = red[3 + a][lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                m0 = fminf(m0, __shfl_xor_sync(FULL_MASK, m0, o));
                m1 = fmaxf(m1, __shfl_xor_sync(FULL_MASK, m1, o));
            }
            lo[a] = m0;
            ext = fmaxf(ext, m1 - m0);
            maxabs = fmaxf(maxabs, fmaxf(fabsf(m0), fabsf(m1)));
        }
        if (lane == 0) {
            // non-finite or zero extent: a cell size of 1 keeps every formula finite (all points land in clamped cells)
            const bool ok = ext > 0.0f && ext < 3.0e38f;
            gi.h = ok ? ext / static_cast<float>(G) : 1.0f;
            gi.inv_h = ok ? static_cast<float>(G) / ext : 1.0f;
            gi.mnx = lo[0];
            gi.mny = lo[1];
            gi.mnz = lo[2];
            gi.G = G;
            gi.margin = 1e-4f * gi.h + 1e-6f * maxabs;
            info[static_cast<size_t>(side) * B + b] = gi;
        }
    }
    __syncthreads();
    const GridInfo g = gi;

    // ---- histogram ----
    for (int i = tid; i < P; i += GRID_BUILD_THREADS) {
        const int cx = cell_coord(pts[static_cast<size_t>(i) * 3], g.mnx, g.inv_h, G);
        const int cy = cell_coord(pts[static_cast<size_t>(i) * 3 + 1], g.mny, g.inv_h, G);
        const int cz = cell_coord(pts[static_cast<size_t>(i) * 3 + 2], g.mnz, g.inv_h, G);
        atomicAdd(&cnt[(cz * G + cy) * G + cx], 1u);
    }
    __syncthreads();

    // ---- exclusive scan of cnt[0 .. ncell): warp w owns a contiguous segment, 32 entries per step ----
    const int seg = (ncell + 31) / 32;
    const int s0 = warp * seg, s1 = min(ncell, s0 + seg);
    unsigned carry = 0u;
    for (int base = s0; base < s1; base += 32) {
        const int i = base + lane;
        const unsigned v = i < s1 ? cnt[i] : 0u;
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(FULL_MASK, inc, o);
            if (lane >= o) inc += t;
        }
        if (i < s1) cnt[i] = carry + inc - v;   // exclusive, relative to the segment
        carry += __shfl_sync(FULL_MASK, inc, 31);
    }
    if (lane == 0) wsum[warp] = carry;
    __syncthreads();
    if (warp == 0) {
        const unsigned v = wsum[lane];
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(FULL_MASK, inc, o);
            if (lane >= o) inc += t;
        }
        wsum[lane] = inc - v;
    }
    __syncthreads();
    const unsigned off = wsum[warp];
    for (int i = s0 + lane; i < s1; i += 32) {
        const unsigned v = cnt[i] + off;
        cnt[i] = v;
        st_out[i] = v;
    }
    if (tid == 0) st_out[ncell] = static_cast<unsigned>(P);
    __syncthreads();
    // bit y of rowmask[z]: the row of cells (z, y, *) holds at least one point (G <= 32; warp z covers slab z).  The far walk
    // of grid_nn_kernel iterates over set bits instead of testing every row of its box.
    {
        const int z = warp, yy = lane;
        const bool occ = z < G && yy < G && cnt[(z * G + yy) * G] != (yy + 1 < G || z + 1 < G ? cnt[(z * G + yy) * G + G] : static_cast<unsigned>(P));
        const unsigned m = __ballot_sync(FULL_MASK, occ);
        if (lane == 0) rowmask[(static_cast<size_t>(side) * B + b) * 32 + z] = z < G ? m : 0u;
    }
    __syncthreads();

    // ---- scatter (cnt now holds the running cursor of every cell; the order inside a cell does not affect any result) ----
    for (int i = tid; i < P; i += GRID_BUILD_THREADS) {
        const float px = pts[static_cast<size_t>(i) * 3], py = pts[static_cast<size_t>(i) * 3 + 1], pz = pts[static_cast<size_t>(i) * 3 + 2];
        const int cx = cell_coord(px, g.mnx, g.inv_h, G), cy = cell_coord(py, g.mny, g.inv_h, G), cz = cell_coord(pz, g.mnz, g.inv_h, G);
        const unsigned pos = atomicAdd(&cnt[(cz * G + cy) * G + cx], 1u);
        if (PAIRED) {
            float *o = reinterpret_cast<float *>(out) + static_cast<size_t>(pos >> 1) * 8 + (pos & 1u);
            o[0] = px;
            o[2] = py;
            o[4] = pz;
            o[6] = __uint_as_float(static_cast<unsigned>(i));
        } else {
            out[pos] = make_float4(px, py, pz, __uint_as_float(static_cast<unsigned>(i)));
        }
    }
    if (PAIRED && (P & 1) && tid == 0) {   // the odd cloud's last pair: a point no query can be near
        float *o = reinterpret_cast<float *>(out) + static_cast<size_t>(P >> 1) * 8 + 1;
        o[0] = o[2] = o[4] = INFINITY;
        o[6] = __uint_as_float(0xffffffffu);
    }
}

__device__ __forceinline__ void scan_range(const float4 *__restrict__ cand, unsigned a, unsigned e, float qx, float qy, float qz,
                                           unsigned long long &best) {
    for (unsigned j = a; j < e; ++j) {
        const float4 c = __ldg(cand + j);
        const unsigned long long k = pack_key(dist2_rn(qx, qy, qz, c.x, c.y, c.z), __float_as_uint(c.w));
        best = k < best ? k : best;
    }
}

// distance from q to the slab [mn + c*h, mn + (c+1)*h] along one axis, made safe (never over-estimated) by the margin
__device__ __forceinline__ float slab_gap(float q, float mn, float h, int c, float margin) {
    const float lo = mn + static_cast<float>(c) * h;
    const float g = fmaxf(lo - q, q - (lo + h)) - margin;
    return g > 0.0f ? g : 0.0f;
}

// grid (ceil(max(P1,P2) / 128), B, n_dir): direction 0 = x queries against y's grid -> kx; direction 1 = y against x -> ky
__global__ void __launch_bounds__(128)
grid_nn_kernel(int P1, int P2, const float4 *__restrict__ sorted, const unsigned *__restrict__ starts,
               const GridInfo *__restrict__ info, const unsigned *__restrict__ rowmask, unsigned long long *__restrict__ kx,
               unsigned long long *__restrict__ ky) {
    const int b = blockIdx.y, dir = blockIdx.z, B = gridDim.y;
    const int Pq = dir ? P2 : P1, Pc = dir ? P1 : P2;
    if (blockIdx.x * 128 >= Pq) return;
    const int qi_raw = blockIdx.x * 128 + threadIdx.x;
    const int qi = qi_raw < Pq ? qi_raw : Pq - 1;   // lanes past the end replay the last query (warp-uniform step 2) and do not store
    const float4 *qpts = sorted + (dir ? static_cast<size_t>(B) * P1 + static_cast<size_t>(b) * P2 : static_cast<size_t>(b) * P1);
    const float4 *cand = sorted + (dir ? static_cast<size_t>(b) * P1 : static_cast<size_t>(B) * P1 + static_cast<size_t>(b) * P2);
    const int cside = dir ? 0 : 1;
    const GridInfo g = info[static_cast<size_t>(cside) * B + b];
    const int G = g.G;
    const unsigned *st = starts + (static_cast<size_t>(cside) * B + b) * (G * G * G + 1);
    unsigned long long *keys = dir ? ky + static_cast<size_t>(b) * P2 : kx + static_cast<size_t>(b) * P1;

    const float4 q = qpts[qi];
    const int cx = cell_coord(q.x, g.mnx, g.inv_h, G), cy = cell_coord(q.y, g.mny, g.inv_h, G), cz = cell_coord(q.z, g.mnz, g.inv_h, G);
    unsigned long long best = KEY_MAX;
    // ---- step 1: the 3x3x3 block around the query's cell ----
    {
        const int z0 = max(cz - 1, 0), z1 = min(cz + 1, G - 1), y0 = max(cy - 1, 0), y1 = min(cy + 1, G - 1);
        const int xa = max(cx - 1, 0), xb = min(cx + 1, G - 1);
        for (int z = z0; z <= z1; ++z)
            for (int yy = y0; yy <= y1; ++yy) {
                const unsigned row = static_cast<unsigned>((z * G + yy) * G);
                scan_range(cand, __ldg(st + row + xa), __ldg(st + row + xb + 1), q.x, q.y, q.z, best);
            }
    }
    // every unvisited point lies beyond one of the block's faces that are inside the grid.  Cell boundaries computed here
    // and the cell assignment of the build differ by rounding (~1e-6 cells + a few ulp of the coordinates): the bound is
    // shrunk by GridInfo::margin and compared with a 2e-5 relative slack on the square
    float bound = INFINITY;
    if (cx - 1 > 0) bound = fminf(bound, q.x - (g.mnx + static_cast<float>(cx - 1) * g.h));
    if (cx + 2 < G) bound = fminf(bound, (g.mnx + static_cast<float>(cx + 2) * g.h) - q.x);
    if (cy - 1 > 0) bound = fminf(bound, q.y - (g.mny + static_cast<float>(cy - 1) * g.h));
    if (cy + 2 < G) bound = fminf(bound, (g.mny + static_cast<float>(cy + 2) * g.h) - q.y);
    if (cz - 1 > 0) bound = fminf(bound, q.z - (g.mnz + static_cast<float>(cz - 1) * g.h));
    if (cz + 2 < G) bound = fminf(bound, (g.mnz + static_cast<float>(cz + 2) * g.h) - q.z);
    bound -= g.margin;
    const bool done = bound == INFINITY || (best != KEY_MAX && bound > 0.0f && key_d2(best) < bound * bound * 0.99998f);
    // ---- step 2 (warp-cooperative): some neighbour is farther than a cell.  Per-thread cell walks would make every load
    // a 32-way scattered access; instead the warp walks ONE region - the union of its unsettled queries' balls - with
    // uniform control flow, so every candidate is one broadcast load and 32 dense distance evaluations.
    const bool need = !done;
    if (__any_sync(FULL_MASK, need)) {
        if (__any_sync(FULL_MASK, need && best == KEY_MAX)) {
            // upper bound from a strided sample of the candidates (<= 128 points, the same for every lane)
            const unsigned stride = static_cast<unsigned>(Pc) / 128u + 1u;
            for (unsigned j = 0; j < static_cast<unsigned>(Pc); j += stride) {
                const float4 c = __ldg(cand + j);
                const unsigned long long k = pack_key(dist2_rn(q.x, q.y, q.z, c.x, c.y, c.z), __float_as_uint(c.w));
                best = k < best ? k : best;
            }
        }
        const float R = need ? sqrtf(key_d2(best)) * 1.0001f + g.margin : 0.0f;
        float lo[3] = {need ? q.x - R : INFINITY, need ? q.y - R : INFINITY, need ? q.z - R : INFINITY};
        float hi[3] = {need ? q.x + R : -INFINITY, need ? q.y + R : -INFINITY, need ? q.z + R : -INFINITY};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo[a] = fminf(lo[a], __shfl_xor_sync(FULL_MASK, lo[a], o));
                hi[a] = fmaxf(hi[a], __shfl_xor_sync(FULL_MASK, hi[a], o));
            }
        }
        const int xa = max(cell_coord(lo[0], g.mnx, g.inv_h, G) - 1, 0), xb = min(cell_coord(hi[0], g.mnx, g.inv_h, G) + 1, G - 1);
        const int ya = max(cell_coord(lo[1], g.mny, g.inv_h, G) - 1, 0), yb = min(cell_coord(hi[1], g.mny, g.inv_h, G) + 1, G - 1);
        const int za = max(cell_coord(lo[2], g.mnz, g.inv_h, G) - 1, 0), zb = min(cell_coord(hi[2], g.mnz, g.inv_h, G) + 1, G - 1);
        // lane z holds the occupancy bits of slab z's rows: empty rows (most of the box when the candidate cloud is clumpy --
        // e.g. the reconstruction of an untrained decoder) are never visited, no table loads, no pruning test
        const unsigned my_rows = __ldg(rowmask + (static_cast<size_t>(cside) * B + b) * 32 + (threadIdx.x & 31));
        const unsigned ymask = (yb >= 31 ? 0xffffffffu : (1u << (yb + 1)) - 1u) & ~((1u << ya) - 1u);
        for (int z = za; z <= zb; ++z) {
            unsigned rows = __shfl_sync(FULL_MASK, my_rows, z) & ymask;
            if (rows == 0u) continue;
            const float gz = slab_gap(q.z, g.mnz, g.h, z, g.margin);
            if (__all_sync(FULL_MASK, !need || gz * gz * 0.9999f > key_d2(best))) continue;
            for (; rows != 0u; rows &= rows - 1u) {
                const int yy = __ffs(rows) - 1;
                const float gy = slab_gap(q.y, g.mny, g.h, yy, g.margin);
                // a row is skipped only if it lies outside the ball of every unsettled query of the warp
                const float rem = key_d2(best) - (gz * gz + gy * gy) * 0.9999f;
                const bool hit = need && rem >= 0.0f;
                if (!__any_sync(FULL_MASK, hit)) continue;
                // x extent of the row that can still matter: the chord of each lane's ball at this (y, z) slab, in cells.
                // cell_coord is monotone in the coordinate and is the function the build used, so a point with
                // |p.x - q.x| <= rx lies in a cell of [cell(q.x - rx), cell(q.x + rx)]: no extra cell of slack is needed
                const float rx = hit ? sqrtf(rem) * 1.0001f + g.margin : 0.0f;
                const int lx = hit ? cell_coord(q.x - rx, g.mnx, g.inv_h, G) : G;
                const int hx = hit ? cell_coord(q.x + rx, g.mnx, g.inv_h, G) : -1;
                const int xa2 = max(xa, __reduce_min_sync(FULL_MASK, lx)), xb2 = min(xb, __reduce_max_sync(FULL_MASK, hx));
                if (xa2 > xb2) continue;   // (visiting the rows centre-out was measured: no gain, the balls are already tight)
                const unsigned row = static_cast<unsigned>((z * G + yy) * G);
                const unsigned a = __ldg(st + row + xa2), e = __ldg(st + row + xb2 + 1);
                for (unsigned j = a; j < e; ++j) {
                    const float4 c = __ldg(cand + j);
                    const unsigned long long k = pack_key(dist2_rn(q.x, q.y, q.z, c.x, c.y, c.z), __float_as_uint(c.w));
                    best = k < best ? k : best;
                }
            }
        }
    }
    if (qi_raw >= Pq) return;
    keys[__float_as_uint(q.w)] = best;
}

// ---- distance-only form (the caller wants no indices: eval.py's Chamfer / D1 PSNR, AE.get_loss without a backward pass) ----------
// Same search, same visited candidates (a superset: ranges are widened to whole pairs), same un-fused d2 per candidate -- so the
// same minima, bit for bit -- but the running best is a float and the distance of TWO candidates costs 8 packed fp32
// instructions (add / mul .f32x2: each half is the IEEE operation of dist2_rn; (c - q)^2 == (q - c)^2 exactly), 4 scalar adds and
// one 3-input min: 7.3 issue slots per candidate against 14.7 for the keyed loop, which was issue bound.
// (Measured beside it: the plain float4 layout with (x, y) of ONE candidate packed -- 8 issue slots per candidate, no paired
// scatter in the build: 260 us against 238 us on the bench's clumpy reconstructions, 107 us against ~125 us on close clouds where
// the 3x3x3 step and the build dominate (keyed kernel: 286 / 119 us).  Looking one row ahead with the table look-ups of the far
// walk, and double-buffering the candidate batches, changed nothing / lost 4 %.)
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float pair_min(float bd, ulonglong2 xy, unsigned long long zz, unsigned long long nqx, unsigned long long nqy,
                                          unsigned long long nqz) {
    const unsigned long long dx = add_f32x2(xy.x, nqx), dy = add_f32x2(xy.y, nqy), dz = add_f32x2(zz, nqz);
    // the three squares packed, the two sums per candidate scalar: ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into
    // FFMA2 whatever --fmad says, which would change the rounding of dist2_rn; it never fuses across the scalar add.rn
    const float2 sx = unpack_f32x2(mul_f32x2(dx, dx)), sy = unpack_f32x2(mul_f32x2(dy, dy)), sz = unpack_f32x2(mul_f32x2(dz, dz));
    return fmin3(bd, __fadd_rn(__fadd_rn(sx.x, sy.x), sz.x), __fadd_rn(__fadd_rn(sx.y, sy.y), sz.y));
}
// candidates [a, e) of a paired array, widened to whole pairs; pairs j, j + step, ... (step > 1: the strided sample).  Four pairs
// are loaded before the first is used: the loop is load-latency bound otherwise (55 % of its stall samples on the first add)
__device__ __forceinline__ void scan_pairs(const float4 *__restrict__ cand, unsigned a, unsigned e, unsigned long long nqx,
                                           unsigned long long nqy, unsigned long long nqz, float &bd, unsigned step = 1u) {
    if (a >= e) return;   // (an empty range with an odd start would otherwise widen to one pair)
    const unsigned long long *c = reinterpret_cast<const unsigned long long *>(cand);
    unsigned j = a >> 1;
    const unsigned je = (e + 1u) >> 1;
    for (; j + 3u * step < je; j += 4u * step) {
        ulonglong2 xy[4];
        unsigned long long zz[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned long long *p = c + 4 * static_cast<size_t>(j + u * step);
            xy[u] = __ldg(reinterpret_cast<const ulonglong2 *>(p));
            zz[u] = __ldg(p + 2);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) bd = pair_min(bd, xy[u], zz[u], nqx, nqy, nqz);
    }
#pragma unroll 1
    for (; j < je; j += step) {
        const unsigned long long *p = c + 4 * static_cast<size_t>(j);
        bd = pair_min(bd, __ldg(reinterpret_cast<const ulonglong2 *>(p)), __ldg(p + 2), nqx, nqy, nqz);
    }
}

// grid and arguments as grid_nn_kernel; `sorted` is the PAIRED array; kx / ky receive pack_key(d2, 0)
__global__ void __launch_bounds__(128)
grid_nn_d2_kernel(int P1, int P2, const float4 *__restrict__ sorted, const unsigned *__restrict__ starts,
                  const GridInfo *__restrict__ info, const unsigned *__restrict__ rowmask, unsigned long long *__restrict__ kx,
                  unsigned long long *__restrict__ ky) {
    const int b = blockIdx.y, dir = blockIdx.z, B = gridDim.y;
    const int Pq = dir ? P2 : P1, Pc = dir ? P1 : P2;
    if (blockIdx.x * 128 >= Pq) return;
    const int qi_raw = blockIdx.x * 128 + threadIdx.x;
    const int qi = qi_raw < Pq ? qi_raw : Pq - 1;   // lanes past the end replay the last query (warp-uniform step 2) and do not store
    const size_t S1 = (static_cast<size_t>(P1) + 1) & ~static_cast<size_t>(1), S2 = (static_cast<size_t>(P2) + 1) & ~static_cast<size_t>(1);
    const float4 *qpts = sorted + (dir ? static_cast<size_t>(B) * S1 + static_cast<size_t>(b) * S2 : static_cast<size_t>(b) * S1);
    const float4 *cand = sorted + (dir ? static_cast<size_t>(b) * S1 : static_cast<size_t>(B) * S1 + static_cast<size_t>(b) * S2);
    const int cside = dir ? 0 : 1;
    const GridInfo g = info[static_cast<size_t>(cside) * B + b];
    const int G = g.G;
    const unsigned *st = starts + (static_cast<size_t>(cside) * B + b) * (G * G * G + 1);
    unsigned long long *keys = dir ? ky + static_cast<size_t>(b) * P2 : kx + static_cast<size_t>(b) * P1;

    const float *qf = reinterpret_cast<const float *>(qpts) + static_cast<size_t>(qi >> 1) * 8 + (qi & 1);
    const float qx = qf[0], qy = qf[2], qz = qf[4];
    const unsigned qw = __float_as_uint(qf[6]);
    const unsigned long long nqx = dup_neg_f32x2(qx), nqy = dup_neg_f32x2(qy), nqz = dup_neg_f32x2(qz);
    const int cx = cell_coord(qx, g.mnx, g.inv_h, G), cy = cell_coord(qy, g.mny, g.inv_h, G), cz = cell_coord(qz, g.mnz, g.inv_h, G);
    float bd = INFINITY;
    // ---- step 1: the 3x3x3 block around the query's cell (the nine table look-ups first: they are independent loads) ----
    {
        const int xa = max(cx - 1, 0), xb = min(cx + 1, G - 1);
        unsigned ra[9], re[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int z = cz + k / 3 - 1, yy = cy + k % 3 - 1;
            const bool in = z >= 0 && z < G && yy >= 0 && yy < G;
            const unsigned row = static_cast<unsigned>((z * G + yy) * G);
            ra[k] = in ? __ldg(st + row + xa) : 0u;
            re[k] = in ? __ldg(st + row + xb + 1) : 0u;
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) scan_pairs(cand, ra[k], re[k], nqx, nqy, nqz, bd);
    }
    // (bounds and margins: see grid_nn_kernel)
    float bound = INFINITY;
    if (cx - 1 > 0) bound = fminf(bound, qx - (g.mnx + static_cast<float>(cx - 1) * g.h));
    if (cx + 2 < G) bound = fminf(bound, (g.mnx + static_cast<float>(cx + 2) * g.h) - qx);
    if (cy - 1 > 0) bound = fminf(bound, qy - (g.mny + static_cast<float>(cy - 1) * g.h));
    if (cy + 2 < G) bound = fminf(bound, (g.mny + static_cast<float>(cy + 2) * g.h) - qy);
    if (cz - 1 > 0) bound = fminf(bound, qz - (g.mnz + static_cast<float>(cz - 1) * g.h));
    if (cz + 2 < G) bound = fminf(bound, (g.mnz + static_cast<float>(cz + 2) * g.h) - qz);
    bound -= g.margin;
    const bool done = bound == INFINITY || (bd != INFINITY && bound > 0.0f && bd < bound * bound * 0.99998f);
    const bool need = !done;
    if (__any_sync(FULL_MASK, need)) {
        if (__any_sync(FULL_MASK, need && bd == INFINITY)) {
            // upper bound from a strided sample of the candidates (<= 128 points, the same for every lane)
            // (<= 64 pairs, every (Pc / 128 + 1)-th pair: a strided sample of the candidates, the same for every lane)
            scan_pairs(cand, 0u, static_cast<unsigned>(Pc), nqx, nqy, nqz, bd, static_cast<unsigned>(Pc) / 128u + 1u);
        }
        const float R = need ? sqrtf(bd) * 1.0001f + g.margin : 0.0f;
        float lo[3] = {need ? qx - R : INFINITY, need ? qy - R : INFINITY, need ? qz - R : INFINITY};
        float hi[3] = {need ? qx + R : -INFINITY, need ? qy + R : -INFINITY, need ? qz + R : -INFINITY};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo[a] = fminf(lo[a], __shfl_xor_sync(FULL_MASK, lo[a], o));
                hi[a] = fmaxf(hi[a], __shfl_xor_sync(FULL_MASK, hi[a], o));
            }
        }
        const int xa = max(cell_coord(lo[0], g.mnx, g.inv_h, G) - 1, 0), xb = min(cell_coord(hi[0], g.mnx, g.inv_h, G) + 1, G - 1);
        const int ya = max(cell_coord(lo[1], g.mny, g.inv_h, G) - 1, 0), yb = min(cell_coord(hi[1], g.mny, g.inv_h, G) + 1, G - 1);
        const int za = max(cell_coord(lo[2], g.mnz, g.inv_h, G) - 1, 0), zb = min(cell_coord(hi[2], g.mnz, g.inv_h, G) + 1, G - 1);
        const unsigned my_rows = __ldg(rowmask + (static_cast<size_t>(cside) * B + b) * 32 + (threadIdx.x & 31));
        const unsigned ymask = (yb >= 31 ? 0xffffffffu : (1u << (yb + 1)) - 1u) & ~((1u << ya) - 1u);
        for (int z = za; z <= zb; ++z) {
            unsigned rows = __shfl_sync(FULL_MASK, my_rows, z) & ymask;
            if (rows == 0u) continue;
            const float gz = slab_gap(qz, g.mnz, g.h, z, g.margin);
            if (__all_sync(FULL_MASK, !need || gz * gz * 0.9999f > bd)) continue;
            for (; rows != 0u; rows &= rows - 1u) {
                const int yy = __ffs(rows) - 1;
                const float gy = slab_gap(qy, g.mny, g.h, yy, g.margin);
                const float rem = bd - (gz * gz + gy * gy) * 0.9999f;
                const bool hit = need && rem >= 0.0f;
                if (!__any_sync(FULL_MASK, hit)) continue;
                const float rx = hit ? sqrtf(rem) * 1.0001f + g.margin : 0.0f;
                const int lx = hit ? cell_coord(qx - rx, g.mnx, g.inv_h, G) : G;
                const int hx = hit ? cell_coord(qx + rx, g.mnx, g.inv_h, G) : -1;
                const int xa2 = max(xa, __reduce_min_sync(FULL_MASK, lx)), xb2 = min(xb, __reduce_max_sync(FULL_MASK, hx));
                if (xa2 > xb2) continue;
                const unsigned row = static_cast<unsigned>((z * G + yy) * G);
                scan_pairs(cand, __ldg(st + row + xa2), __ldg(st + row + xb2 + 1), nqx, nqy, nqz, bd);
            }
        }
    }
    if (qi_raw >= Pq) return;
    keys[qw] = pack_key(bd, 0u);
}

int64_t chamfer_grid_extra_bytes(int B, int P1, int P2, int G) {
    const int64_t pts = static_cast<int64_t>(B) * (((static_cast<int64_t>(P1) + 1) & ~1ll) + ((static_cast<int64_t>(P2) + 1) & ~1ll)) * 16;
    const int64_t tab = 2ll * B * (static_cast<int64_t>(G) * G * G + 1) * 4;
    return pts + ((tab + 15) / 16) * 16 + 2ll * B * static_cast<int64_t>(sizeof(GridInfo)) + 2ll * B * 32 * 4;
}

// ---- multi-CTA build for one big cloud (scene scale): the same grid as grid_build_kernel makes, spread over the SMs ----------
__device__ __forceinline__ int ordered_int(float f) {   // monotone float -> int map (atomicMin / atomicMax on floats)
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// grid (blocks, B): bb[b][0..2] = min, bb[b][4..6] = min of the negated coordinates, as ordered ints (memset to 0x7f7f7f7f)
__global__ void __launch_bounds__(256)
big_bbox_kernel(const float *__restrict__ pts, int P, int *__restrict__ bb) {
    const int b = blockIdx.y;
    const float *p = pts + static_cast<size_t>(b) * P * 3;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * 256 + threadIdx.x; i < P; i += gridDim.x * 256) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = p[static_cast<size_t>(i) * 3 + a];
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(FULL_MASK, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(FULL_MASK, mx[a], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(bb + b * 8 + a, ordered_int(mn[a]));
            atomicMin(bb + b * 8 + 4 + a, ordered_int(-mx[a]));   // max as the min of the negated values: one memset pattern
        }
    }
}

__global__ void big_info_kernel(int *__restrict__ bb, int G, GridInfo *__restrict__ info, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float ext = 0.0f, lo[3], maxabs = 0.0f;
    for (int a = 0; a < 3; ++a) {
        const float m0 = ordered_float(bb[b * 8 + a]), m1 = -ordered_float(bb[b * 8 + 4 + a]);
        lo[a] = m0;
        ext = fmaxf(ext, m1 - m0);
        maxabs = fmaxf(maxabs, fmaxf(fabsf(m0), fabsf(m1)));
    }
    GridInfo gi;
    const bool ok = ext > 0.0f && ext < 3.0e38f;
    gi.h = ok ? ext / static_cast<float>(G) : 1.0f;
    gi.inv_h = ok ? static_cast<float>(G) / ext : 1.0f;
    gi.mnx = lo[0];
    gi.mny = lo[1];
    gi.mnz = lo[2];
    gi.G = G;
    gi.margin = 1e-4f * gi.h + 1e-6f * maxabs;
    gi.pad = 0;
    info[b] = gi;
}

// grid (blocks, B): SCATTER = false: cell histogram into cnt; true: counting-sort scatter with cnt as the running cursors
template <bool SCATTER>
__global__ void __launch_bounds__(256)
big_cells_kernel(const float *__restrict__ pts, int P, const GridInfo *__restrict__ info, unsigned *__restrict__ cnt, int ncell1,
                 float4 *__restrict__ sorted) {
    const int b = blockIdx.y;
    const GridInfo g = info[b];
    const int G = g.G;
    const float *p = pts + static_cast<size_t>(b) * P * 3;
    unsigned *c = cnt + static_cast<size_t>(b) * ncell1;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < P; i += gridDim.x * 256) {
        const float px = p[static_cast<size_t>(i) * 3], py = p[static_cast<size_t>(i) * 3 + 1], pz = p[static_cast<size_t>(i) * 3 + 2];
        const int cell = (cell_coord(pz, g.mnz, g.inv_h, G) * G + cell_coord(py, g.mny, g.inv_h, G)) * G + cell_coord(px, g.mnx, g.inv_h, G);
        const unsigned pos = atomicAdd(c + cell, 1u);
        if (SCATTER) sorted[static_cast<size_t>(b) * P + pos] = make_float4(px, py, pz, __uint_as_float(static_cast<unsigned>(i)));
    }
}

// one CTA (1024 threads) per cloud: exclusive scan of the histogram -> starts (and the scatter's cursors), row occupancy masks
__global__ void __launch_bounds__(GRID_BUILD_THREADS)
big_scan_kernel(unsigned *__restrict__ cnt, unsigned *__restrict__ starts, unsigned *__restrict__ rowmask, int G, int P) {
    __shared__ unsigned wsum[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ncell = G * G * G;
    unsigned *c = cnt + static_cast<size_t>(b) * (ncell + 1);
    unsigned *st_out = starts + static_cast<size_t>(b) * (ncell + 1);
    const int seg = (ncell + 31) / 32;
    const int s0 = warp * seg, s1 = min(ncell, s0 + seg);
    unsigned carry = 0u;
    for (int base = s0; base < s1; base += 32) {
        const int i = base + lane;
        const unsigned v = i < s1 ? c[i] : 0u;
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(FULL_MASK, inc, o);
            if (lane >= o) inc += t;
        }
        if (i < s1) c[i] = carry + inc - v;
        carry += __shfl_sync(FULL_MASK, inc, 31);
    }
    if (lane == 0) wsum[warp] = carry;
    __syncthreads();
    if (warp == 0) {
        const unsigned v = wsum[lane];
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(FULL_MASK, inc, o);
            if (lane >= o) inc += t;
        }
        wsum[lane] = inc - v;
    }
    __syncthreads();
    const unsigned off = wsum[warp];
    for (int i = s0 + lane; i < s1; i += 32) {
        const unsigned v = c[i] + off;
        c[i] = v;
        st_out[i] = v;
    }
    if (tid == 0) st_out[ncell] = static_cast<unsigned>(P);
    __syncthreads();
    {
        const int z = warp, yy = lane;
        const bool occ = z < G && yy < G &&
                         st_out[(z * G + yy) * G] != (yy + 1 < G || z + 1 < G ? st_out[(z * G + yy) * G + G] : static_cast<unsigned>(P));
        const unsigned m = __ballot_sync(FULL_MASK, occ);
        if (lane == 0) rowmask[static_cast<size_t>(b) * 32 + z] = z < G ? m : 0u;
    }
}

// One side only (the candidate cloud of the scene-scale kNN, knn.cu): sorted [B, P] float4, starts [B, G^3 + 1], info [B],
// rowmask [B, 32]
int grid_build_single(const float *pts, int B, int P, int G, float4 *sorted, unsigned *starts, GridInfo *info, unsigned *rowmask,
                      void *scratch, cudaStream_t st) {
    if (scratch && P >= 32768) {
        // big clouds: bounding box, histogram, scan and scatter as grid-wide kernels (a 1M-point cloud on one CTA takes ~1.1 ms)
        const int ncell1 = G * G * G + 1;
        unsigned *cnt = static_cast<unsigned *>(scratch);                       // [B][ncell1] histogram -> cursors
        int *bb = reinterpret_cast<int *>(cnt + static_cast<size_t>(B) * ncell1);   // [B][8]
        cudaError_t e = cudaMemsetAsync(cnt, 0, static_cast<size_t>(B) * ncell1 * 4, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(bb, 0x7f, static_cast<size_t>(B) * 8 * 4, st);        // mins: large positive ordered ints
        if (e != cudaSuccess) {
            set_error("grid build: cudaMemsetAsync failed: %s", cudaGetErrorString(e));
            return static_cast<int>(e);
        }
        const int blocks = min((P + 255) / 256, num_sms() * 8);
        big_bbox_kernel<<<dim3(blocks, B), 256, 0, st>>>(pts, P, bb);
        big_info_kernel<<<(B + 63) / 64, 64, 0, st>>>(bb, G, info, B);
        big_cells_kernel<false><<<dim3(blocks, B), 256, 0, st>>>(pts, P, info, cnt, ncell1, nullptr);
        big_scan_kernel<<<B, GRID_BUILD_THREADS, 0, st>>>(cnt, starts, rowmask, G, P);
        big_cells_kernel<true><<<dim3(blocks, B), 256, 0, st>>>(pts, P, info, cnt, ncell1, sorted);
        return check_launch("grid build (multi-CTA)");
    }
    static bool attr_done_dev[64] = {false};
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) d = 0;
    if (!attr_done_dev[d]) {
        const cudaError_t e = cudaFuncSetAttribute(grid_build_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 32 * 32 * 4 + 4);
        if (e != cudaSuccess) {
            set_error("grid build: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
            return static_cast<int>(e);
        }
        attr_done_dev[d] = true;
    }
    grid_build_kernel<false><<<dim3(B, 1), GRID_BUILD_THREADS, static_cast<size_t>(G) * G * G * 4 + 4, st>>>(pts, pts, P, P, G, sorted, starts, info,
                                                                                                  rowmask);
    return check_launch("grid_build_kernel");
}

int chamfer_grid_pick(int P1, int P2) {   // grid resolution, or 0: use the brute-force kernels
    const int lo = P1 < P2 ? P1 : P2, hi = P1 < P2 ? P2 : P1;
    if (lo < 1024 || hi > (1 << 24)) return 0;
    if (const char *e = getenv("PCC_CHAMFER_GRID")) {   // tuning knob: cells per axis, 4..32
        const int g = atoi(e);
        if (g >= 4 && g <= 32) return g;
    }
    return lo >= 4096 ? 32 : 16;
}

// extra = 16-byte aligned workspace region behind the two key arrays (chamfer_grid_extra_bytes); fills kx and (if ky) ky
int chamfer_grid_run(const float *x, const float *y, int B, int P1, int P2, int G, unsigned long long *kx,
                     unsigned long long *ky, bool want_idx, void *extra, cudaStream_t st) {
    float4 *sorted = static_cast<float4 *>(extra);
    const int64_t pts = static_cast<int64_t>(B) * (((static_cast<int64_t>(P1) + 1) & ~1ll) + ((static_cast<int64_t>(P2) + 1) & ~1ll)) * 16;
    static const bool keyed_only = getenv("PCC_CHAMFER_KEYED") != nullptr;   // A/B: the keyed kernel also when no index is wanted
    const bool d2_only = !want_idx && !keyed_only;
    const int64_t tab = 2ll * B * (static_cast<int64_t>(G) * G * G + 1) * 4;
    unsigned *starts = reinterpret_cast<unsigned *>(static_cast<char *>(extra) + pts);
    GridInfo *info = reinterpret_cast<GridInfo *>(static_cast<char *>(extra) + pts + ((tab + 15) / 16) * 16);
    unsigned *rowmask = reinterpret_cast<unsigned *>(info + 2ll * B);   // [2][B][32]
    const size_t smem = static_cast<size_t>(G) * G * G * 4 + 4;
    static bool attr_done_dev[64] = {false};   // per device: the attribute belongs to the device's copy of the kernel
    int attr_done_d = 0;
    if (cudaGetDevice(&attr_done_d) != cudaSuccess || attr_done_d < 0 || attr_done_d >= 64) attr_done_d = 0;
    if (!attr_done_dev[attr_done_d]) {
        cudaError_t e = cudaFuncSetAttribute(grid_build_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 32 * 32 * 4 + 4);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(grid_build_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 32 * 32 * 4 + 4);
        if (e != cudaSuccess) {
            set_error("chamfer grid: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
            return static_cast<int>(e);
        }
        attr_done_dev[attr_done_d] = true;
    }
    if (d2_only) grid_build_kernel<true><<<dim3(B, 2), GRID_BUILD_THREADS, smem, st>>>(x, y, P1, P2, G, sorted, starts, info, rowmask);
    else grid_build_kernel<false><<<dim3(B, 2), GRID_BUILD_THREADS, smem, st>>>(x, y, P1, P2, G, sorted, starts, info, rowmask);
    int rc = check_launch("grid_build_kernel");
    if (rc) return rc;
    const int pmax = P1 > P2 ? P1 : P2;
    const dim3 grid((pmax + 127) / 128, B, ky ? 2 : 1);
    if (d2_only) grid_nn_d2_kernel<<<grid, 128, 0, st>>>(P1, P2, sorted, starts, info, rowmask, kx, ky);
    else grid_nn_kernel<<<grid, 128, 0, st>>>(P1, P2, sorted, starts, info, rowmask, kx, ky);
    return check_launch("grid_nn_kernel");
}

}  // namespace pcc
