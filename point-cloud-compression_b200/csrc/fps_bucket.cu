// Scene-scale farthest point sampling (the 1M-point cloud of SURVEY.md 8a/a1 cfg5) with EXACT spatial skipping.
//
// Same contract and the same bits as fps_block_kernel / fps_grid_kernel (fps.cu): running min of the un-fused d2, arg-max
// with ties to the lowest index.  What changes is the work per iteration.  The centre picked at iteration k has the largest
// running distance D_k of the whole cloud, so only points closer than D_k to it can change -- a ball that shrinks like
// k^(-1/2) on a surface.  The cloud is therefore sorted once along a Hilbert curve (64^3 cells, counting sort) and cut into
// buckets of 256 consecutive points, each with its bounding box, the largest running distance it holds and that point's
// (index, coordinates).  A bucket whose box is farther from the new centre than its own largest running distance cannot
// change (margins below) and is skipped; the arg-max is taken over the bucket records.  One CTA per cloud owns the whole
// state: bucket records and boxes in shared memory (184 KB), points and running distances in L2; per iteration
//   A  every thread tests the union box of its 4 buckets (registers), then the 4 boxes, and queues the buckets that may change
//   B  a warp per queued bucket: 8 points per lane from L2, d2 / min / store-if-smaller, warp arg-max -> bucket record
//   C  threads whose buckets were queued refresh their best key (d2 bits, 2^20 - 1 - index, bucket); two redux.sync per warp
//      with a queued bucket -> that warp's key in shared memory
//   D  every warp reduces the 32 keys itself (two redux.sync, no broadcast barrier): the index is in the key, the coordinates
//      in the winner's bucket record
// i.e. three CTA barriers and one L2 round trip per iteration instead of a grid-wide exchange among 123 CTAs (fps.cu,
// 4.4 us per iteration at 1M points), and ~10 buckets instead of the whole cloud once a few hundred centres exist.
// The first centres move nearly every bucket -- a full pass over the cloud through ONE SM costs far more than the grid-wide
// exchange -- so the co-resident kernel runs the head of the sampling (head_iterations()) and hands over its running
// distances; this kernel continues from the centre that kernel picked last.
// Measured and not kept: the cloud's maximum by one shared-memory atomicMax per warp instead of phase D's reduction (a 64-bit
// shared atomicMax is a compare-and-swap loop, ATOMS.CAST.SPIN: 16.6 ms against 15.6); a second level inside phase B (eight
// 32-point sub-buckets per bucket with their own boxes and records in L2, only the reachable ones loaded: half the bytes) --
// 31.5 ms against 18.9 ms: phase B is one L2 round trip deep and the iteration waits for its slowest warp, so the extra
// dependent trip for the sub-records and the per-sub-bucket arg-max cost more than the bytes they save.
// clock64 ticks per iteration at 1M points (make ticks; 12.6 bucket visits on average): A 0.5 k, B 2.2 k, C + D 1.2 k clocks.
//
// Exactness of the skip.  For every point p of a bucket with box [lo, hi] and the exact gap vector g(c) to the box,
// |p - c|^2 >= |g|^2.  dist2_rn is 5 roundings deep, so dist2_rn(p, c) >= |p - c|^2 (1 - 5u), u = 2^-24; the box distance
// computed in fp32 satisfies lb <= |g|^2 (1 + 5u) (FMA contraction only removes roundings).  Hence dist2_rn(p, c) >
// lb (1 - 2^-20), and a bucket is skipped only if lb (1 - 2^-20) >= its largest running distance (>= every point's): no
// `d < md` can hold.  Squares that underflow break relative bounds, so nothing is skipped on lb < 1e-30 (absolute errors of
// underflowed terms are < 1e-44).  Skipping less is always correct; the queue order is arbitrary but every record is a pure
// function of the bucket's points, so the output does not depend on it.
#include <stdlib.h>

#include "fps.cuh"

namespace pcc {
namespace fpsb {

constexpr int BS = 256;                       // points per bucket
constexpr int PPL = BS / 32;                  // per lane
constexpr int THREADS = 1024;
constexpr int BPT = 4;                        // buckets per thread (adjacent on the curve: one union box)
constexpr int NB_MAX = THREADS * BPT;         // 4096 buckets = 1,048,576 points
constexpr int LG = 6, G = 1 << LG, NCELL = G * G * G;
constexpr int TAB_ROWS = 11;                  // lo xyz, hi xyz, d2, idx, best xyz -- [row][NB_MAX] per cloud
constexpr float SKIP_K = 0.99999904632568359375f;   // 1 - 2^-20
constexpr float SKIP_TINY = 1e-30f;

struct Info {
    float mnx, mny, mnz, inv_h;
};

struct Ws {
    float4 *sorted;   // [B][NB * BS]  x, y, z, original index (bits); tail padded with index 0xffffffff
    float *mind;      // [B][NB * BS]  running min distance, sorted order; pads 0
    float *tab;       // [B][TAB_ROWS][NB_MAX]
    unsigned *cnt;    // [B][NCELL]    histogram -> cursors
    int *bb;          // [B][8]
    Info *info;       // [B]
    float *md;        // [B][N]        running distances handed over by the co-resident head, original order
    void *grid_ws;    // the head kernel's exchange table
};

__host__ __device__ inline int64_t align16(int64_t v) { return (v + 15) & ~15ll; }

static int64_t carve(void *base, int B, int N, Ws *w) {
    const int64_t npad = static_cast<int64_t>((N + BS - 1) / BS) * BS;
    char *p = static_cast<char *>(base);
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        char *r = p ? p + off : nullptr;
        off += align16(bytes);
        return r;
    };
    w->sorted = reinterpret_cast<float4 *>(take(B * npad * 16));
    w->mind = reinterpret_cast<float *>(take(B * npad * 4));
    w->tab = reinterpret_cast<float *>(take(static_cast<int64_t>(B) * TAB_ROWS * NB_MAX * 4));
    w->cnt = reinterpret_cast<unsigned *>(take(static_cast<int64_t>(B) * NCELL * 4));
    w->bb = reinterpret_cast<int *>(take(static_cast<int64_t>(B) * 8 * 4));
    w->info = reinterpret_cast<Info *>(take(static_cast<int64_t>(B) * sizeof(Info)));
    w->md = reinterpret_cast<float *>(take(static_cast<int64_t>(B) * N * 4));
    w->grid_ws = take(fps_grid_workspace_bytes());
    return off;
}

// ---- build: bounding box, Morton cell histogram, scan, scatter ---------------------------------------------------------------
__device__ __forceinline__ int ordered_int(float f) {   // monotone float -> int map (atomicMin on floats)
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// grid (blocks, B): bb[b][0..2] = min, bb[b][4..6] = min of the negated coordinates (memset to 0x7f7f7f7f)
__global__ void __launch_bounds__(256)
bbox_kernel(const float *__restrict__ pts, int N, int *__restrict__ bb) {
    const int b = blockIdx.y;
    const float *p = pts + static_cast<size_t>(b) * N * 3;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * 256 + threadIdx.x; i < N; i += gridDim.x * 256) {
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
            atomicMin(bb + b * 8 + 4 + a, ordered_int(-mx[a]));
        }
    }
}

__global__ void info_kernel(const int *__restrict__ bb, Info *__restrict__ info, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float ext = 0.0f, lo[3];
    for (int a = 0; a < 3; ++a) {
        const float m0 = ordered_float(bb[b * 8 + a]), m1 = -ordered_float(bb[b * 8 + 4 + a]);
        lo[a] = m0;
        ext = fmaxf(ext, m1 - m0);
    }
    Info gi;
    gi.mnx = lo[0];
    gi.mny = lo[1];
    gi.mnz = lo[2];
    gi.inv_h = (ext > 0.0f && ext < 3.0e38f) ? static_cast<float>(G) / ext : 0.0f;   // the order only shapes the buckets
    info[b] = gi;
}

__device__ __forceinline__ unsigned spread3(unsigned v) {   // bit i of v (< 64) -> bit 3 i
    unsigned r = 0u;
#pragma unroll
    for (int i = 0; i < LG; ++i) r |= ((v >> i) & 1u) << (3 * i);
    return r;
}

// Cell (x, y, z) -> its position on the 64^3 Hilbert curve (Skilling's transpose form).  A run of a Hilbert curve is a connected
// blob, a run of a Z curve can straddle one of its jumps and get a box many cells wide: measured on the scene, 28 % fewer bucket
// visits over a whole sampling and 38 % fewer per late iteration (9 instead of 15).  Only the ORDER depends on this function.
__device__ __forceinline__ unsigned hilbert_cell(unsigned x, unsigned y, unsigned z) {
    unsigned X[3] = {x, y, z};
#pragma unroll
    for (unsigned Q = 1u << (LG - 1); Q > 1u; Q >>= 1) {
        const unsigned P = Q - 1u;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (X[i] & Q) {
                X[0] ^= P;
            } else {
                const unsigned t = (X[0] ^ X[i]) & P;
                X[0] ^= t;
                X[i] ^= t;
            }
        }
    }
    X[1] ^= X[0];
    X[2] ^= X[1];
    unsigned t = 0u;
#pragma unroll
    for (unsigned Q = 1u << (LG - 1); Q > 1u; Q >>= 1)
        if (X[2] & Q) t ^= Q - 1u;
    return (spread3(X[0] ^ t) << 2) | (spread3(X[1] ^ t) << 1) | spread3(X[2] ^ t);
}

__device__ __forceinline__ unsigned axis_cell(float p, float mn, float inv_h) {
    const float f = (p - mn) * inv_h;
    const int c = f > 0.0f ? static_cast<int>(fminf(f, static_cast<float>(G - 1))) : 0;   // NaN -> 0
    return static_cast<unsigned>(c);
}

// grid (blocks, B): SCATTER = false: histogram of the Morton cells; true: counting-sort scatter with cnt as the running cursors
template <bool SCATTER>
__global__ void __launch_bounds__(256)
cells_kernel(const float *__restrict__ pts, int N, int npad, const Info *__restrict__ info, unsigned *__restrict__ cnt,
             float4 *__restrict__ sorted, int hilbert) {
    const int b = blockIdx.y;
    const Info g = info[b];
    const float *p = pts + static_cast<size_t>(b) * N * 3;
    unsigned *c = cnt + static_cast<size_t>(b) * NCELL;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < N; i += gridDim.x * 256) {
        const float px = p[static_cast<size_t>(i) * 3], py = p[static_cast<size_t>(i) * 3 + 1], pz = p[static_cast<size_t>(i) * 3 + 2];
        const unsigned ax = axis_cell(px, g.mnx, g.inv_h), ay = axis_cell(py, g.mny, g.inv_h), az = axis_cell(pz, g.mnz, g.inv_h);
        const unsigned cell = hilbert ? hilbert_cell(ax, ay, az) : (spread3(ax) | (spread3(ay) << 1) | (spread3(az) << 2));
        const unsigned pos = atomicAdd(c + cell, 1u);
        if (SCATTER) sorted[static_cast<size_t>(b) * npad + pos] = make_float4(px, py, pz, __uint_as_float(static_cast<unsigned>(i)));
    }
}

// one CTA (1024 threads) per cloud: in-place exclusive scan of the histogram
__global__ void __launch_bounds__(THREADS)
scan_kernel(unsigned *__restrict__ cnt) {
    __shared__ unsigned wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned *c = cnt + static_cast<size_t>(blockIdx.x) * NCELL;
    constexpr int SEG = NCELL / 32;
    const int s0 = warp * SEG;
    unsigned carry = 0u;
    for (int base = s0; base < s0 + SEG; base += 32) {
        const unsigned v = c[base + lane];
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(FULL_MASK, inc, o);
            if (lane >= o) inc += t;
        }
        c[base + lane] = carry + inc - v;
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
    if (off)
        for (int i = s0 + lane; i < s0 + SEG; i += 32) c[i] += off;
}

// ---- one bucket: 8 points per lane ----------------------------------------------------------------------------------------
// (d2 bits, ~index) as one 64-bit key: the maximum is the largest running distance, then the LOWEST original index.
struct Best {
    unsigned hi, lo;
    float x, y, z;
};

__device__ __forceinline__ void best_take(Best &b, unsigned hi, unsigned lo, float x, float y, float z) {
    if (hi > b.hi || (hi == b.hi && lo > b.lo)) {
        b.hi = hi;
        b.lo = lo;
        b.x = x;
        b.y = y;
        b.z = z;
    }
}

// warp arg-max of the lanes' records; returns the owning lane (the same value in every lane)
__device__ __forceinline__ unsigned warp_best(unsigned hi, unsigned lo, unsigned &mh, unsigned &ml) {
    mh = __reduce_max_sync(FULL_MASK, hi);
    ml = __reduce_max_sync(FULL_MASK, hi == mh ? lo : 0u);
    return __ffs(__ballot_sync(FULL_MASK, hi == mh && lo == ml)) - 1;
}

// grid (NB_MAX / 8, B), 256 threads: a warp per bucket -- pads, running distances, boxes and the first records
__global__ void __launch_bounds__(256)
init_kernel(float4 *__restrict__ sorted_all, float *__restrict__ mind_all, float *__restrict__ tab_all, int N, int NB,
            float init_dist, const float *__restrict__ md_in) {
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int bucket = blockIdx.x * 8 + (threadIdx.x >> 5);
    const size_t npad = static_cast<size_t>(NB) * BS;
    float4 *sorted = sorted_all + b * npad;
    float *mind = mind_all + b * npad;
    float *tab = tab_all + static_cast<size_t>(b) * TAB_ROWS * NB_MAX;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    Best best = {0u, 0u, 0.0f, 0.0f, 0.0f};
    if (bucket < NB) {
#pragma unroll
        for (int p = 0; p < PPL; ++p) {
            const int j = bucket * BS + p * 32 + lane;
            if (j < N) {
                const float4 v = sorted[j];
                const float m0 = md_in ? md_in[static_cast<size_t>(b) * N + __float_as_uint(v.w)] : init_dist;
                mind[j] = m0;
                lo[0] = fminf(lo[0], v.x), lo[1] = fminf(lo[1], v.y), lo[2] = fminf(lo[2], v.z);
                hi[0] = fmaxf(hi[0], v.x), hi[1] = fmaxf(hi[1], v.y), hi[2] = fmaxf(hi[2], v.z);
                best_take(best, __float_as_uint(m0), ~__float_as_uint(v.w), v.x, v.y, v.z);
            } else {   // pad: distance 0 and the lowest possible key -> never preferred to a real point
                sorted[j] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0xffffffffu));
                mind[j] = 0.0f;
            }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo[a] = fminf(lo[a], __shfl_xor_sync(FULL_MASK, lo[a], o));
                hi[a] = fmaxf(hi[a], __shfl_xor_sync(FULL_MASK, hi[a], o));
            }
    }
    unsigned mh, ml;
    const unsigned owner = warp_best(best.hi, best.lo, mh, ml);
    if (lane == owner) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            tab[a * NB_MAX + bucket] = lo[a];
            tab[(3 + a) * NB_MAX + bucket] = hi[a];
        }
        tab[6 * NB_MAX + bucket] = __uint_as_float(mh);
        tab[7 * NB_MAX + bucket] = __uint_as_float(~ml);
        tab[8 * NB_MAX + bucket] = best.x;
        tab[9 * NB_MAX + bucket] = best.y;
        tab[10 * NB_MAX + bucket] = best.z;
    }
}

__device__ __forceinline__ float box_lb(float lx, float ly, float lz, float hx, float hy, float hz, float cx, float cy, float cz) {
    const float dx = fmaxf(fmaxf(lx - cx, cx - hx), 0.0f);
    const float dy = fmaxf(fmaxf(ly - cy, cy - hy), 0.0f);
    const float dz = fmaxf(fmaxf(lz - cz, cz - hz), 0.0f);
    return dx * dx + dy * dy + dz * dz;
}
// true: no point of a bucket with largest running distance d2 can change
__device__ __forceinline__ bool box_skips(float lb, float d2) { return d2 == 0.0f || (lb >= SKIP_TINY && lb * SKIP_K >= d2); }

struct Smem {
    float lo[3][NB_MAX], hi[3][NB_MAX];
    float d2[NB_MAX];
    unsigned idx[NB_MAX];
    float x[NB_MAX], y[NB_MAX], z[NB_MAX];
    unsigned short queue[NB_MAX];
    unsigned long long wkey[32];  // the warps' best keys (d2 bits, 2^20 - 1 - index, bucket)
    unsigned qn;
};

// Low word of an arg-max key: ties in d2 go to the LOWEST original index (N < 2^20), the bucket rides in the low 12 bits
// (NB_MAX = 4096) so that the winner's coordinates can be read from its bucket record; pads lose to every real point.
__device__ __forceinline__ unsigned key_low(unsigned idx, unsigned bucket) {
    return (idx == 0xffffffffu ? 0u : (0xfffffu - idx) << 12) | bucket;
}

__global__ void __launch_bounds__(THREADS, 1)
fps_bucket_kernel(const float *__restrict__ xyz, const float4 *__restrict__ sorted_all, float *__restrict__ mind_all,
                  const float *__restrict__ tab_all, int N, int NB, int npoint, const int64_t *__restrict__ start_idx,
                  int64_t *__restrict__ out_idx, float *__restrict__ out_xyz, float quant_cube, int i0) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &s = *reinterpret_cast<Smem *>(smem_raw);
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t npad = static_cast<size_t>(NB) * BS;
    const float *pc = xyz + static_cast<size_t>(b) * N * 3;
    const float4 *sorted = sorted_all + b * npad;
    float *mind = mind_all + b * npad;
    const float *tab = tab_all + static_cast<size_t>(b) * TAB_ROWS * NB_MAX;
    int64_t *out = out_idx + static_cast<size_t>(b) * npoint;

    for (int t = tid; t < NB_MAX; t += THREADS) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            s.lo[a][t] = tab[a * NB_MAX + t];
            s.hi[a][t] = tab[(3 + a) * NB_MAX + t];
        }
        s.d2[t] = tab[6 * NB_MAX + t];
        s.idx[t] = __float_as_uint(tab[7 * NB_MAX + t]);
        s.x[t] = tab[8 * NB_MAX + t];
        s.y[t] = tab[9 * NB_MAX + t];
        s.z[t] = tab[10 * NB_MAX + t];
    }
    if (tid == 0) s.qn = 0u;
    const int k_n = npoint < N ? npoint : N;
    for (int k = k_n + tid; k < npoint; k += THREADS) {  // PyTorch3D padding when npoint > N
        out[k] = -1;
        if (out_xyz) {
            float *o = out_xyz + (static_cast<size_t>(b) * npoint + k) * 3;
            o[0] = o[1] = o[2] = 0.0f;
        }
    }
    __syncthreads();

    // this thread's four buckets: union box, largest running distance, best record
    const int b0 = tid * BPT;
    float ulx = INFINITY, uly = INFINITY, ulz = INFINITY, uhx = -INFINITY, uhy = -INFINITY, uhz = -INFINITY, umax = 0.0f;
    unsigned t_hi = 0u, t_lo = 0u;
#pragma unroll
    for (int q = 0; q < BPT; ++q) {
        ulx = fminf(ulx, s.lo[0][b0 + q]), uly = fminf(uly, s.lo[1][b0 + q]), ulz = fminf(ulz, s.lo[2][b0 + q]);
        uhx = fmaxf(uhx, s.hi[0][b0 + q]), uhy = fmaxf(uhy, s.hi[1][b0 + q]), uhz = fmaxf(uhz, s.hi[2][b0 + q]);
        umax = fmaxf(umax, s.d2[b0 + q]);
        const unsigned h = __float_as_uint(s.d2[b0 + q]), l = key_low(s.idx[b0 + q], b0 + q);
        if (h > t_hi || (h == t_hi && l > t_lo)) t_hi = h, t_lo = l;
    }
    {   // the warps' first keys (phase C keeps them current)
        const unsigned w_hi = __reduce_max_sync(FULL_MASK, t_hi);
        const unsigned w_lo = __reduce_max_sync(FULL_MASK, t_hi == w_hi ? t_lo : 0u);
        if (lane == 0) s.wkey[warp] = (static_cast<unsigned long long>(w_hi) << 32) | w_lo;
    }

    int far = 0;
    if (i0 > 0) {     // hand-over: the head kernel picked out[i0] and has not applied it yet
        far = static_cast<int>(out[i0]);
    } else if (start_idx) {  // an out-of-range start (the reference would raise an IndexError on the host) must not read outside the cloud
        const long long s0 = start_idx[b];
        far = s0 < 0 ? 0 : (s0 >= N ? N - 1 : static_cast<int>(s0));
    }
    float cx = pc[static_cast<size_t>(far) * 3 + 0], cy = pc[static_cast<size_t>(far) * 3 + 1], cz = pc[static_cast<size_t>(far) * 3 + 2];
    __syncthreads();

#ifdef PCC_FPS_TICKS   // bring-up aid (make ticks): clocks per phase, summed over the iterations, seen by thread 0
    long long tk[5] = {0, 0, 0, 0, 0}, tq = 0;
#define FPS_TICK(n)                        \
    {                                      \
        const long long t_now = clock64(); \
        tk[n] += t_now - t_prev;           \
        t_prev = t_now;                    \
    }
    long long t_prev = clock64();
#else
#define FPS_TICK(n)
#endif
    for (int i = i0; i < k_n; ++i) {
        if (tid == 0) {
            out[i] = far;
            if (out_xyz) store_centre(out_xyz + (static_cast<size_t>(b) * npoint + i) * 3, cx, cy, cz, quant_cube);
        }
        if (i == k_n - 1) break;

        // A: which of my buckets may change?
        unsigned hit = 0u;
        if (!box_skips(box_lb(ulx, uly, ulz, uhx, uhy, uhz, cx, cy, cz), umax)) {
#pragma unroll
            for (int q = 0; q < BPT; ++q) {
                const int t = b0 + q;
                const float lb = box_lb(s.lo[0][t], s.lo[1][t], s.lo[2][t], s.hi[0][t], s.hi[1][t], s.hi[2][t], cx, cy, cz);
                if (!box_skips(lb, s.d2[t])) hit |= 1u << q;
            }
            if (hit) {
                unsigned pos = atomicAdd(&s.qn, static_cast<unsigned>(__popc(hit)));
#pragma unroll
                for (int q = 0; q < BPT; ++q)
                    if (hit & (1u << q)) s.queue[pos++] = static_cast<unsigned short>(b0 + q);
            }
        }
        FPS_TICK(0)
        const bool warp_hit = __any_sync(FULL_MASK, hit != 0u);
        __syncthreads();
        FPS_TICK(1)
        const int nq = static_cast<int>(s.qn);
#ifdef PCC_FPS_TICKS
        tq += nq;
#endif

        // B: a warp per queued bucket
        for (int w = warp; w < nq; w += 32) {
            const int bucket = s.queue[w];
            const float4 *sp = sorted + static_cast<size_t>(bucket) * BS + lane;
            float *mp = mind + static_cast<size_t>(bucket) * BS + lane;
            float4 v[PPL];
            float m[PPL];
#pragma unroll
            for (int p = 0; p < PPL; ++p) {
                v[p] = __ldg(sp + p * 32);
                m[p] = __ldcg(mp + p * 32);
            }
            Best best = {0u, 0u, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int p = 0; p < PPL; ++p) {
                const float d = dist2_rn(v[p].x, v[p].y, v[p].z, cx, cy, cz);
                if (d < m[p]) {   // pn_kit.py:327-328; pads hold 0 and never change
                    m[p] = d;
                    mp[p * 32] = d;
                }
                best_take(best, __float_as_uint(m[p]), ~__float_as_uint(v[p].w), v[p].x, v[p].y, v[p].z);
            }
            unsigned mh, ml;
            const unsigned owner = warp_best(best.hi, best.lo, mh, ml);
            if (lane == owner) {
                s.d2[bucket] = __uint_as_float(mh);
                s.idx[bucket] = ~ml;
                s.x[bucket] = best.x;
                s.y[bucket] = best.y;
                s.z[bucket] = best.z;
            }
        }
        __syncthreads();
        FPS_TICK(2)
        if (tid == 0) s.qn = 0u;   // next written after the third barrier

        // C: warps with queued buckets refresh their best key
        if (warp_hit) {
            if (hit) {
                const float4 d4 = *reinterpret_cast<const float4 *>(&s.d2[b0]);
                const uint4 i4 = *reinterpret_cast<const uint4 *>(&s.idx[b0]);
                const unsigned h[4] = {__float_as_uint(d4.x), __float_as_uint(d4.y), __float_as_uint(d4.z), __float_as_uint(d4.w)};
                const unsigned l[4] = {key_low(i4.x, b0), key_low(i4.y, b0 + 1), key_low(i4.z, b0 + 2), key_low(i4.w, b0 + 3)};
                umax = fmaxf(fmaxf(d4.x, d4.y), fmaxf(d4.z, d4.w));
                t_hi = 0u, t_lo = 0u;
#pragma unroll
                for (int q = 0; q < BPT; ++q)
                    if (h[q] > t_hi || (h[q] == t_hi && l[q] > t_lo)) t_hi = h[q], t_lo = l[q];
            }
            const unsigned w_hi = __reduce_max_sync(FULL_MASK, t_hi);
            const unsigned w_lo = __reduce_max_sync(FULL_MASK, t_hi == w_hi ? t_lo : 0u);
            if (lane == 0) s.wkey[warp] = (static_cast<unsigned long long>(w_hi) << 32) | w_lo;
        }
        __syncthreads();
        FPS_TICK(3)

        // D: every warp reduces the 32 keys itself (no broadcast barrier): two redux.sync; the winner's index is in the key,
        // its coordinates in its bucket's record
        {
            const unsigned long long k = s.wkey[lane];
            const unsigned kh = static_cast<unsigned>(k >> 32), mh = __reduce_max_sync(FULL_MASK, kh);
            const unsigned kl = __reduce_max_sync(FULL_MASK, kh == mh ? static_cast<unsigned>(k) : 0u);
            const unsigned wb = kl & 0xfffu;
            far = static_cast<int>(0xfffffu - (kl >> 12));
            cx = s.x[wb];
            cy = s.y[wb];
            cz = s.z[wb];
        }
        FPS_TICK(4)
    }
#ifdef PCC_FPS_TICKS
    if (tid == 0 || tid == 1023)
        printf("fps_bucket ticks (thread %d, %d iterations, %lld bucket visits): A %lld | barrier 1 %lld | B + barrier 2 %lld | C + barrier 3 %lld | D %lld clocks per iteration\n",
               tid, k_n - i0, tq, tk[0] / (k_n - i0), tk[1] / (k_n - i0), tk[2] / (k_n - i0), tk[3] / (k_n - i0), tk[4] / (k_n - i0));
#endif
}

}  // namespace fpsb

// Iterations left to the co-resident kernel: the first centres move nearly every bucket, which costs one SM far more than a
// grid-wide exchange; after ~N / 4096 centres a new one reaches a few dozen buckets and the bucketed iteration is the cheaper one
// (measured at 1M points: 15.96 ms without a head, 15.37 with 256 iterations, 15.61 with 390, 16.8 with 640).
static int head_iterations(int N, int npoint) {
    int k0 = N / 4096;
    k0 = k0 < 128 ? 128 : (k0 > 1024 ? 1024 : k0);
    if (const char *e = getenv("PCC_FPS_HEAD")) k0 = atoi(e);   // tuning knob; 0: no head
    const int k_n = npoint < N ? npoint : N;
    return (k0 >= 2 && k_n >= 4 * k0) ? k0 : 0;
}

bool fps_bucket_takes(int B, int N, int npoint) {
    const char *e = getenv("PCC_FPS_PATH");   // "grid": the co-resident multi-CTA kernel; "bucket": this form whenever it fits
    if (e && e[0] == 'g') return false;
    if (N >= fpsb::NB_MAX * fpsb::BS) return false;   // N < 2^20: the index field of the arg-max key
    if (e && e[0] == 'b') return true;
    if (N < 196608) return false;                     // measured: the co-resident kernel wins on smaller clouds
    if (head_iterations(N, npoint) > 0) return true;  // long samplings
    // short samplings: this form runs one CTA per cloud side by side; the co-resident kernel needs ceil(N / 8192) SMs per cloud
    // and takes the clouds it cannot fit one launch after the other (1M points: 2.4 ms per cloud for 512 centres against 2.7 ms
    // for any number of clouds here -- the two branches of the pppe encoder's first level)
    const int per_launch = num_sms() / ((N + 8191) / 8192);
    return B > 1 && per_launch < 2;
}

int64_t fps_bucket_workspace_bytes(int B, int N) {
    fpsb::Ws w;
    return fpsb::carve(nullptr, B, N, &w);
}

int fps_bucket_run(const float *xyz, int B, int N, int npoint, const int64_t *start_idx, float init_dist, int64_t *out_idx,
                   float *out_xyz, float quant_cube, void *workspace, cudaStream_t st) {
    using namespace fpsb;
    Ws w;
    carve(workspace, B, N, &w);
    const int NB = (N + BS - 1) / BS, npad = NB * BS;
    const int k0 = head_iterations(N, npoint);
    if (k0 > 0) {   // centres 0 .. k0 - 1 and the running distances after the first k0 - 1 of them
        const int rc = fps_grid_run(xyz, B, N, k0, start_idx, init_dist, out_idx, out_xyz, quant_cube, npoint, w.md, w.grid_ws, st);
        if (rc) return rc;
    }
    cudaError_t e = cudaMemsetAsync(w.cnt, 0, static_cast<size_t>(B) * NCELL * 4, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(w.bb, 0x7f, static_cast<size_t>(B) * 8 * 4, st);
    static bool attr_done_dev[64] = {false};   // per device: the attribute belongs to the device's copy of the kernel
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) d = 0;
    if (e == cudaSuccess && !attr_done_dev[d]) {
        e = cudaFuncSetAttribute(fps_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(Smem)));
        attr_done_dev[d] = e == cudaSuccess;
    }
    if (e != cudaSuccess) {
        set_error("pcc_fps_f32 (bucket form): setup failed: %s", cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    const char *ec = getenv("PCC_FPS_CURVE");   // A/B: "morton" orders the cells along a Z curve instead of a Hilbert curve
    const int hilbert = !(ec && ec[0] == 'm');
    const int blocks = min((N + 255) / 256, num_sms() * 8);
    bbox_kernel<<<dim3(blocks, B), 256, 0, st>>>(xyz, N, w.bb);
    info_kernel<<<(B + 63) / 64, 64, 0, st>>>(w.bb, w.info, B);
    cells_kernel<false><<<dim3(blocks, B), 256, 0, st>>>(xyz, N, npad, w.info, w.cnt, nullptr, hilbert);
    scan_kernel<<<B, THREADS, 0, st>>>(w.cnt);
    cells_kernel<true><<<dim3(blocks, B), 256, 0, st>>>(xyz, N, npad, w.info, w.cnt, w.sorted, hilbert);
    init_kernel<<<dim3(NB_MAX / 8, B), 256, 0, st>>>(w.sorted, w.mind, w.tab, N, NB, init_dist, k0 > 0 ? w.md : nullptr);
    fps_bucket_kernel<<<B, THREADS, sizeof(Smem), st>>>(xyz, w.sorted, w.mind, w.tab, N, NB, npoint, start_idx, out_idx, out_xyz,
                                                       quant_cube, k0 > 0 ? k0 - 1 : 0);
    return check_launch("fps_bucket_kernel");
}

}  // namespace pcc
