// Nearest neighbour (K=1) search, Chamfer distance forward / backward for sm_100a.
//
// Replaces pytorch3d.loss.chamfer_distance (/root/reference/AE.py:67, PPPF_AE.py:168, pppe_pcd_ae.py:820,
// eval.py:204) -- two knn_points(K=1) searches plus mean reductions -- and the 1-NN inner step of the D1 PSNR
// loop (/root/reference/eval.py:68-81).
//
// nn1_kernel: register-tiled brute force.  Each thread owns QPT queries; candidates are staged in shared
//   memory as float4 and read as LDS.128 broadcasts in sub-blocks of SUB; per sub-block and query the thread
//   takes min over the SUB un-fused d2 values (8 ops per pair + 1 FMNMX) and remembers the first sub-block
//   that improved its best (1 FSETP + 1 SEL per sub-block).  The arg-min index is resolved afterwards by
//   re-evaluating only that sub-block, so index tracking costs ~0.25 instructions per pair.  The per-point
//   minima are bit-exact to the CPU oracle and ties go to the lowest index.  The kernel is FP32-issue bound
//   (SURVEY.md 8d): 196 KB of input per 8192^2 cloud pair against 134 M pair evaluations.
//   When a launch would leave SMs idle the candidate range is split over several CTAs, which merge through
//   atomicMin on the packed (d2, idx) key.
// chamfer_finalize_kernel: unpacks keys, reduces the means in double (fixed order -> run-to-run deterministic).
#include <stdlib.h>

#include "pcc_common.cuh"

namespace pcc {

constexpr int NN_THREADS = 256;
constexpr int NN_QPT = 4;
constexpr int NN_SUB = 8;
constexpr int NN_TILE = 1024;
constexpr int NN_QPB = NN_THREADS * NN_QPT;

struct Nn1Dir {
    const float *q;  // [B,P1,3]
    const float *p;  // [B,P2,3]
    unsigned long long *keys;  // [B,P1]
    int P1, P2;
};

// grid: (max query blocks, splits, n_dir*B)
__global__ void __launch_bounds__(NN_THREADS)
nn1_kernel(Nn1Dir d0, Nn1Dir d1, int B, int splits, int use_atomic) {
    __shared__ float4 tile[NN_TILE];
    const int dir = blockIdx.z / B;
    const int b = blockIdx.z - dir * B;
    const Nn1Dir D = dir == 0 ? d0 : d1;
    const int P1 = D.P1, P2 = D.P2;
    const int q0 = blockIdx.x * NN_QPB;
    if (q0 >= P1) return;
    // candidate range of this split, rounded to whole sub-blocks so sub-block ids are split independent
    const int per = ((P2 + splits - 1) / splits + NN_SUB - 1) / NN_SUB * NN_SUB;
    const int c_begin = blockIdx.y * per;
    const int c_end = min(P2, c_begin + per);
    if (c_begin >= c_end) return;
    const float *qc = D.q + static_cast<size_t>(b) * P1 * 3;
    const float *pc = D.p + static_cast<size_t>(b) * P2 * 3;

    float qx[NN_QPT], qy[NN_QPT], qz[NN_QPT], best[NN_QPT];
    int blk[NN_QPT];
#pragma unroll
    for (int r = 0; r < NN_QPT; ++r) {
        const int qi = q0 + r * NN_THREADS + threadIdx.x;
        const int qs = qi < P1 ? qi : P1 - 1;
        qx[r] = qc[static_cast<size_t>(qs) * 3 + 0];
        qy[r] = qc[static_cast<size_t>(qs) * 3 + 1];
        qz[r] = qc[static_cast<size_t>(qs) * 3 + 2];
        best[r] = __int_as_float(0x7f800000);
        blk[r] = c_begin;
    }

    for (int t0 = c_begin; t0 < c_end; t0 += NN_TILE) {
        const int tn = min(NN_TILE, c_end - t0);
        const int tn_pad = (tn + NN_SUB - 1) / NN_SUB * NN_SUB;
        __syncthreads();
        for (int pt = threadIdx.x; pt < tn_pad; pt += NN_THREADS) {
            float4 v = make_float4(__int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0x7f800000), 0.f);
            if (pt < tn) {
                const float *s = pc + static_cast<size_t>(t0 + pt) * 3;
                v = make_float4(s[0], s[1], s[2], 0.f);
            }
            tile[pt] = v;  // padding = +inf coordinates -> d2 = +inf, never strictly below a best
        }
        __syncthreads();
        for (int j0 = 0; j0 < tn_pad; j0 += NN_SUB) {
            float4 c[NN_SUB];
#pragma unroll
            for (int s = 0; s < NN_SUB; ++s) c[s] = tile[j0 + s];
#pragma unroll
            for (int r = 0; r < NN_QPT; ++r) {
                float m = dist2_rn(qx[r], qy[r], qz[r], c[0].x, c[0].y, c[0].z);
#pragma unroll
                for (int s = 1; s < NN_SUB; ++s) m = fminf(m, dist2_rn(qx[r], qy[r], qz[r], c[s].x, c[s].y, c[s].z));
                if (m < best[r]) {  // strict: the first sub-block reaching the minimum is remembered
                    best[r] = m;
                    blk[r] = t0 + j0;
                }
            }
        }
    }

    // resolve the arg-min inside the remembered sub-block (first index whose d2 equals the minimum)
#pragma unroll
    for (int r = 0; r < NN_QPT; ++r) {
        const int qi = q0 + r * NN_THREADS + threadIdx.x;
        if (qi >= P1) continue;
        unsigned idx = static_cast<unsigned>(blk[r]);
        const int jend = min(blk[r] + NN_SUB, c_end);
        for (int j = jend - 1; j >= blk[r]; --j) {
            const float *s = pc + static_cast<size_t>(j) * 3;
            if (dist2_rn(qx[r], qy[r], qz[r], s[0], s[1], s[2]) == best[r]) idx = static_cast<unsigned>(j);
        }
        const unsigned long long key = pack_key(best[r], idx);
        unsigned long long *dst = D.keys + static_cast<size_t>(b) * P1 + qi;
        if (use_atomic)
            atomicMin(dst, key);
        else
            *dst = key;
    }
}

// ---- one-pass Chamfer: every pair is evaluated ONCE and feeds both directions -------------------------------------------
// CTA = 256 threads x 4 consecutive rows (x points) each = a block of 1024 rows, against a range of columns (y points).
//   row side   : running min + remembered sub-block per row in registers (as nn1_kernel);
//   column side: per sub-block the thread folds its 4 rows into 8 column minima, one redux.sync.min per column gives
//                the warp's minimum (float bits are monotone: d2 >= +0), lane s parks column s in the warp's own
//                shared array (no atomics: each (warp, column) is produced exactly once per tile); after the tile the
//                8 warp arrays are merged (lowest warp wins ties = lowest row index) into one global atomicMin on the
//                packed key (d2 bits << 32 | global warp id).  chamfer_resolve_cols_kernel then scans only the 128 rows
//                of the winning warp for the first row whose d2 equals the minimum.
// Cost: 8 (d2) + ~2.2 instructions per pair for BOTH directions, vs 2 x 9.25 for two nn1 passes.
constexpr int CH_RB = NN_THREADS * NN_QPT;  // rows per CTA (1024)
constexpr int CH_WARPS = NN_THREADS / 32;

__global__ void __launch_bounds__(NN_THREADS)
chamfer_onepass_kernel(const float *__restrict__ x, const float *__restrict__ y, int P1, int P2, int splits,
                       unsigned long long *__restrict__ kx, unsigned long long *__restrict__ ky) {
    __shared__ float4 tile[NN_TILE];
    __shared__ unsigned colmin[CH_WARPS][NN_TILE];
    const int b = blockIdx.z;
    const int rb = blockIdx.x;
    const int q0 = rb * CH_RB;
    const int per = ((P2 + splits - 1) / splits + NN_SUB - 1) / NN_SUB * NN_SUB;
    const int c_begin = blockIdx.y * per;
    const int c_end = min(P2, c_begin + per);
    if (c_begin >= c_end) return;
    const float *xc = x + static_cast<size_t>(b) * P1 * 3;
    const float *yc = y + static_cast<size_t>(b) * P2 * 3;
    const int warp = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    const float INF = __int_as_float(0x7f800000);

    float qx[NN_QPT], qy[NN_QPT], qz[NN_QPT], best[NN_QPT];
    int blk[NN_QPT];
#pragma unroll
    for (int r = 0; r < NN_QPT; ++r) {
        const int qi = q0 + threadIdx.x * NN_QPT + r;
        if (qi < P1) {
            qx[r] = xc[static_cast<size_t>(qi) * 3 + 0];
            qy[r] = xc[static_cast<size_t>(qi) * 3 + 1];
            qz[r] = xc[static_cast<size_t>(qi) * 3 + 2];
        } else {  // rows beyond P1: +inf coordinates -> d2 = +inf, never a column minimum
            qx[r] = qy[r] = qz[r] = INF;
        }
        best[r] = INF;
        blk[r] = c_begin;
    }

    for (int t0 = c_begin; t0 < c_end; t0 += NN_TILE) {
        const int tn = min(NN_TILE, c_end - t0);
        const int tn_pad = (tn + NN_SUB - 1) / NN_SUB * NN_SUB;
        __syncthreads();  // previous tile's merge finished
        for (int pt = threadIdx.x; pt < tn_pad; pt += NN_THREADS) {
            float4 v = make_float4(INF, INF, INF, 0.f);
            if (pt < tn) {
                const float *s = yc + static_cast<size_t>(t0 + pt) * 3;
                v = make_float4(s[0], s[1], s[2], 0.f);
            }
            tile[pt] = v;
        }
        __syncthreads();
        for (int j0 = 0; j0 < tn_pad; j0 += NN_SUB) {
            float4 c[NN_SUB];
#pragma unroll
            for (int s = 0; s < NN_SUB; ++s) c[s] = tile[j0 + s];
            float cm[NN_SUB];
#pragma unroll
            for (int r = 0; r < NN_QPT; ++r) {
                float m = INF;
#pragma unroll
                for (int s = 0; s < NN_SUB; ++s) {
                    const float d = dist2_rn(qx[r], qy[r], qz[r], c[s].x, c[s].y, c[s].z);
                    m = fminf(m, d);
                    cm[s] = r == 0 ? d : fminf(cm[s], d);
                }
                if (m < best[r]) {
                    best[r] = m;
                    blk[r] = t0 + j0;
                }
            }
            unsigned mine = 0u;
#pragma unroll
            for (int s = 0; s < NN_SUB; ++s) {
                // NaN-free inputs: +inf - +inf only happens for padded rows against padded columns, which are never read
                const unsigned w = __reduce_min_sync(FULL_MASK, __float_as_uint(cm[s]));
                if (lane == static_cast<unsigned>(s)) mine = w;
            }
            if (lane < NN_SUB) colmin[warp][j0 + lane] = mine;
        }
        __syncthreads();
        // merge the 8 warp arrays; lowest warp id wins ties (= lowest row index)
        for (int j = threadIdx.x; j < tn; j += NN_THREADS) {
            unsigned m = colmin[0][j];
            int w = 0;
#pragma unroll
            for (int k = 1; k < CH_WARPS; ++k) {
                const unsigned v = colmin[k][j];
                if (v < m) {
                    m = v;
                    w = k;
                }
            }
            const unsigned long long key = (static_cast<unsigned long long>(m) << 32) | static_cast<unsigned>(rb * CH_WARPS + w);
            atomicMin(ky + static_cast<size_t>(b) * P2 + t0 + j, key);
        }
    }

    // row side: resolve the arg-min inside the remembered sub-block
#pragma unroll
    for (int r = 0; r < NN_QPT; ++r) {
        const int qi = q0 + threadIdx.x * NN_QPT + r;
        if (qi >= P1) continue;
        unsigned idx = static_cast<unsigned>(blk[r]);
        const int jend = min(blk[r] + NN_SUB, c_end);
        for (int j = jend - 1; j >= blk[r]; --j) {
            const float *s = yc + static_cast<size_t>(j) * 3;
            if (dist2_rn(qx[r], qy[r], qz[r], s[0], s[1], s[2]) == best[r]) idx = static_cast<unsigned>(j);
        }
        atomicMin(kx + static_cast<size_t>(b) * P1 + qi, pack_key(best[r], idx));
    }
}

// ky holds (d2 bits, global warp id): replace the warp id by the first row of that warp's 128 rows whose d2 equals the min.
// One warp per column at a time: the 128 candidate rows are read coalesced, 4 per lane, and the first match is found
// with ballots.
__global__ void __launch_bounds__(256)
chamfer_resolve_cols_kernel(const float *__restrict__ x, const float *__restrict__ y, int P1, int P2, long long n_cols,
                            unsigned long long *__restrict__ ky) {
    const unsigned lane = lane_id();
    const long long warp_global = (blockIdx.x * 256ll + threadIdx.x) >> 5;
    const long long n_warps = (static_cast<long long>(gridDim.x) * 256ll) >> 5;
    for (long long col = warp_global; col < n_cols; col += n_warps) {
        const long long b = col / P2;
        const float *xc = x + static_cast<size_t>(b) * P1 * 3;
        const float *yp = y + static_cast<size_t>(col) * 3;
        const float yx = yp[0], yy = yp[1], yz = yp[2];
        const unsigned long long key = ky[col];
        const float d = key_d2(key);
        const int i0 = static_cast<int>(key_idx(key)) * (32 * NN_QPT);
        unsigned idx = static_cast<unsigned>(i0);
        for (int k = 0; k < NN_QPT; ++k) {
            const int i = i0 + k * 32 + static_cast<int>(lane);
            bool hit = false;
            if (i < P1) {
                const float *s = xc + static_cast<size_t>(i) * 3;
                hit = dist2_rn(s[0], s[1], s[2], yx, yy, yz) == d;
            }
            const unsigned m = __ballot_sync(FULL_MASK, hit);
            if (m) {
                idx = static_cast<unsigned>(i0 + k * 32 + __ffs(m) - 1);
                break;
            }
        }
        if (lane == 0) ky[col] = pack_key(d, idx);
    }
}

// One CTA per cloud: unpack keys -> (d2, idx), mean reductions in double.
__global__ void __launch_bounds__(1024)
chamfer_finalize_kernel(const unsigned long long *__restrict__ kx, const unsigned long long *__restrict__ ky, int P1,
                        int P2, float *__restrict__ out_dx, int64_t *__restrict__ out_ix, float *__restrict__ out_dy,
                        int64_t *__restrict__ out_iy, float *__restrict__ out_per_cloud) {
    __shared__ double red[2][32];
    const int b = blockIdx.x;
    double sx = 0.0, sy = 0.0;
    for (int i = threadIdx.x; i < P1; i += 1024) {
        const unsigned long long k = kx[static_cast<size_t>(b) * P1 + i];
        const float d = key_d2(k);
        out_dx[static_cast<size_t>(b) * P1 + i] = d;
        if (out_ix) out_ix[static_cast<size_t>(b) * P1 + i] = static_cast<int64_t>(key_idx(k));
        sx += static_cast<double>(d);
    }
    if (ky) {
        for (int j = threadIdx.x; j < P2; j += 1024) {
            const unsigned long long k = ky[static_cast<size_t>(b) * P2 + j];
            const float d = key_d2(k);
            out_dy[static_cast<size_t>(b) * P2 + j] = d;
            if (out_iy) out_iy[static_cast<size_t>(b) * P2 + j] = static_cast<int64_t>(key_idx(k));
            sy += static_cast<double>(d);
        }
    }
    if (!out_per_cloud) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(FULL_MASK, sx, o);
        sy += __shfl_xor_sync(FULL_MASK, sy, o);
    }
    if (lane_id() == 0) {
        red[0][threadIdx.x >> 5] = sx;
        red[1][threadIdx.x >> 5] = sy;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        sx = red[0][threadIdx.x];
        sy = red[1][threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sx += __shfl_xor_sync(FULL_MASK, sx, o);
            sy += __shfl_xor_sync(FULL_MASK, sy, o);
        }
        if (threadIdx.x == 0)
            out_per_cloud[b] = static_cast<float>(sx / static_cast<double>(P1) + sy / static_cast<double>(P2));
    }
}

__global__ void __launch_bounds__(32) batch_mean_kernel(const float *__restrict__ per_cloud, int B,
                                                         float *__restrict__ out_loss) {
    double s = 0.0;
    for (int b = threadIdx.x; b < B; b += 32) s += static_cast<double>(per_cloud[b]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL_MASK, s, o);
    if (threadIdx.x == 0) out_loss[0] = static_cast<float>(s / static_cast<double>(B));
}

// direction 0 (blockIdx.y == 0): rows i of x;  direction 1: rows j of y.
__global__ void __launch_bounds__(256)
chamfer_bwd_kernel(const float *__restrict__ x, const float *__restrict__ y, const int64_t *__restrict__ ix,
                   const int64_t *__restrict__ iy, int B, int P1, int P2, const float *__restrict__ grad_loss,
                   float *__restrict__ gx, float *__restrict__ gy) {
    const int dir = blockIdx.y;
    const float *a = dir == 0 ? x : y;
    const float *o = dir == 0 ? y : x;
    const int64_t *ia = dir == 0 ? ix : iy;
    float *ga = dir == 0 ? gx : gy;
    float *go = dir == 0 ? gy : gx;
    const int Pa = dir == 0 ? P1 : P2, Po = dir == 0 ? P2 : P1;
    const float w = grad_loss[0] / static_cast<float>(static_cast<long long>(B) * Pa);
    const long long total = static_cast<long long>(B) * Pa;
    for (long long r = blockIdx.x * 256ll + threadIdx.x; r < total; r += static_cast<long long>(gridDim.x) * 256ll) {
        const long long b = r / Pa;
        const long long j = ia[r];
        const float *ap = a + r * 3;
        const float *op = o + (b * Po + j) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float g = __fmul_rn(__fmul_rn(2.0f, w), __fsub_rn(ap[c], op[c]));
            atomicAdd(ga + r * 3 + c, g);
            atomicAdd(go + (b * Po + j) * 3 + c, -g);
        }
    }
}

static int pick_splits(int B, int ndir, int P1max, int P2min) {
    const long long ctas = static_cast<long long>((P1max + NN_QPB - 1) / NN_QPB) * B * ndir;
    const long long want = 2ll * num_sms();
    int s = 1;
    if (ctas < want) s = static_cast<int>((want + ctas - 1) / ctas);
    const int max_s = (P2min + NN_TILE - 1) / NN_TILE;  // at least one tile of candidates per split
    if (s > max_s) s = max_s;
    return s < 1 ? 1 : s;
}

static int launch_nn1(const Nn1Dir &d0, const Nn1Dir &d1, int ndir, int B, cudaStream_t st) {
    const int P1max = ndir == 2 ? (d0.P1 > d1.P1 ? d0.P1 : d1.P1) : d0.P1;
    const int P2min = ndir == 2 ? (d0.P2 < d1.P2 ? d0.P2 : d1.P2) : d0.P2;
    const int splits = pick_splits(B, ndir, P1max, P2min);
    if (splits > 1) {
        cudaError_t e = cudaMemsetAsync(d0.keys, 0xff, sizeof(unsigned long long) * B * d0.P1, st);
        if (e == cudaSuccess && ndir == 2) e = cudaMemsetAsync(d1.keys, 0xff, sizeof(unsigned long long) * B * d1.P1, st);
        if (e != cudaSuccess) {
            set_error("nn1: memset failed: %s", cudaGetErrorString(e));
            return static_cast<int>(e);
        }
    }
    dim3 grid((P1max + NN_QPB - 1) / NN_QPB, splits, ndir * B);
    nn1_kernel<<<grid, NN_THREADS, 0, st>>>(d0, d1, B, splits, splits > 1 ? 1 : 0);
    return check_launch("nn1_kernel");
}

}  // namespace pcc

PCC_API int64_t pcc_nn1_workspace_bytes(int B, int P1, int P2) {
    (void)P2;
    return static_cast<int64_t>(sizeof(unsigned long long)) * B * P1;
}

namespace pcc {
// chamfer_grid.cu: exact grid-pruned search for large clouds
int64_t chamfer_grid_extra_bytes(int B, int P1, int P2, int G);
int chamfer_grid_pick(int P1, int P2);
int chamfer_grid_run(const float *x, const float *y, int B, int P1, int P2, int G, unsigned long long *kx,
                     unsigned long long *ky, bool want_idx, void *extra, cudaStream_t st);
static int chamfer_path() {   // PCC_CHAMFER_PATH=brute forces the brute-force kernels (A/B measurements)
    static const int v = [] {
        const char *e = getenv("PCC_CHAMFER_PATH");
        return (e && e[0] == 'b') ? 1 : 0;
    }();
    return v;
}
}  // namespace pcc

PCC_API int64_t pcc_chamfer_workspace_bytes(int B, int P1, int P2) {
    const int64_t keys = static_cast<int64_t>(sizeof(unsigned long long)) * B * (static_cast<int64_t>(P1) + P2);
    const int G = pcc::chamfer_grid_pick(P1, P2);
    return keys + (G ? 16 + pcc::chamfer_grid_extra_bytes(B, P1, P2, G) : 0);   // + 16: the grid region is 16-byte aligned
}

PCC_API int pcc_nn1_f32(const float *q, const float *p, int B, int P1, int P2, float *out_d2, int64_t *out_idx,
                        void *workspace, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(q && p && out_d2 && workspace, "pcc_nn1_f32: null pointer");
    PCC_REQUIRE(B >= 0 && P1 >= 0 && P2 >= 1, "pcc_nn1_f32: bad shape B=%d P1=%d P2=%d", B, P1, P2);
    PCC_REQUIRE(B <= 32767, "pcc_nn1_f32: B=%d exceeds 32767", B);
    if (B == 0 || P1 == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Nn1Dir d0{q, p, static_cast<unsigned long long *>(workspace), P1, P2};
    int rc = launch_nn1(d0, d0, 1, B, st);
    if (rc) return rc;
    chamfer_finalize_kernel<<<B, 1024, 0, st>>>(d0.keys, nullptr, P1, P2, out_d2, out_idx, nullptr, nullptr, nullptr);
    return check_launch("chamfer_finalize_kernel");
}

PCC_API int pcc_chamfer_fwd_f32(const float *x, const float *y, int B, int P1, int P2, float *out_dx, int64_t *out_ix,
                                float *out_dy, int64_t *out_iy, float *out_per_cloud, float *out_loss,
                                void *workspace, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(x && y && out_dx && out_dy && out_per_cloud && workspace, "pcc_chamfer_fwd_f32: null pointer");
    PCC_REQUIRE(B >= 1 && P1 >= 1 && P2 >= 1, "pcc_chamfer_fwd_f32: bad shape B=%d P1=%d P2=%d", B, P1, P2);
    PCC_REQUIRE(B <= 32767, "pcc_chamfer_fwd_f32: B=%d exceeds 32767", B);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long *kx = static_cast<unsigned long long *>(workspace);
    unsigned long long *ky = kx + static_cast<size_t>(B) * P1;
    const int G = chamfer_path() ? 0 : chamfer_grid_pick(P1, P2);
    if (G) {
        const uintptr_t extra = (reinterpret_cast<uintptr_t>(ky + static_cast<size_t>(B) * P2) + 15) & ~static_cast<uintptr_t>(15);
        const int rc = chamfer_grid_run(x, y, B, P1, P2, G, kx, ky, out_ix || out_iy, reinterpret_cast<void *>(extra), st);
        if (rc) return rc;
    } else {
        // one-pass kernel: pick the column split that balances the grid over the SMs (~4 resident CTAs each)
        const int row_blocks = (P1 + CH_RB - 1) / CH_RB;
        const int max_s = (P2 + 255) / 256;  // at least 256 columns per split
        const long long slots = 4ll * num_sms();
        int best_s = 1;
        double best_eff = -1.0;
        for (int sgs = 1; sgs <= max_s; ++sgs) {
            const long long items = static_cast<long long>(row_blocks) * B * sgs;
            const long long waves = (items + slots - 1) / slots;
            const double eff = static_cast<double>(items) / static_cast<double>(waves * slots) - 0.002 * sgs;
            if (eff > best_eff) {
                best_eff = eff;
                best_s = sgs;
            }
        }
        cudaError_t e = cudaMemsetAsync(kx, 0xff, sizeof(unsigned long long) * B * (static_cast<size_t>(P1) + P2), st);
        if (e != cudaSuccess) {
            set_error("pcc_chamfer_fwd_f32: memset failed: %s", cudaGetErrorString(e));
            return static_cast<int>(e);
        }
        dim3 grid(row_blocks, best_s, B);
        chamfer_onepass_kernel<<<grid, NN_THREADS, 0, st>>>(x, y, P1, P2, best_s, kx, ky);
        int rc = check_launch("chamfer_onepass_kernel");
        if (rc) return rc;
        if (out_iy) {
            const long long n_cols = static_cast<long long>(B) * P2;
            long long blocks = (n_cols + 7) / 8;
            if (blocks > 16ll * num_sms()) blocks = 16ll * num_sms();
            chamfer_resolve_cols_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(x, y, P1, P2, n_cols, ky);
            rc = check_launch("chamfer_resolve_cols_kernel");
            if (rc) return rc;
        }
    }
    int rc = 0;
    chamfer_finalize_kernel<<<B, 1024, 0, st>>>(kx, ky, P1, P2, out_dx, out_ix, out_dy, out_iy, out_per_cloud);
    rc = check_launch("chamfer_finalize_kernel");
    if (rc) return rc;
    if (out_loss) {
        batch_mean_kernel<<<1, 32, 0, st>>>(out_per_cloud, B, out_loss);
        rc = check_launch("batch_mean_kernel");
    }
    return rc;
}

PCC_API int pcc_chamfer_bwd_f32(const float *x, const float *y, const int64_t *ix, const int64_t *iy, int B, int P1,
                                int P2, const float *grad_loss, float *gx, float *gy, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(x && y && ix && iy && grad_loss && gx && gy, "pcc_chamfer_bwd_f32: null pointer");
    PCC_REQUIRE(B >= 1 && P1 >= 1 && P2 >= 1, "pcc_chamfer_bwd_f32: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(gx, 0, sizeof(float) * 3 * B * P1, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(gy, 0, sizeof(float) * 3 * B * P2, st);
    if (e != cudaSuccess) {
        set_error("pcc_chamfer_bwd_f32: memset failed: %s", cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    const long long rows = static_cast<long long>(B) * (P1 > P2 ? P1 : P2);
    long long blocks = (rows + 255) / 256;
    const long long cap = static_cast<long long>(num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    chamfer_bwd_kernel<<<dim3(static_cast<unsigned>(blocks), 2), 256, 0, st>>>(x, y, ix, iy, B, P1, P2, grad_loss, gx, gy);
    return check_launch("chamfer_bwd_kernel");
}
