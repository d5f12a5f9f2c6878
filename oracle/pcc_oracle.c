/*
 * pcc_oracle.c -- CPU ORACLE for the hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path (point-cloud-compression_b200/) never does.
 *
 * It restates, in plain C, the CPU algorithms the reference runs for the hot path:
 *   - pn_kit.farthest_point_sample_batch         /root/reference/pn_kit.py:309-330   (in-tree torch code)
 *   - pn_kit.index_points                        /root/reference/pn_kit.py:332-360   (in-tree torch code)
 *   - pytorch3d.ops.knn_points / knn_gather      third-party, pytorch3d==0.7.5 (requirements_gpu.txt:33),
 *   - pytorch3d.ops.ball_query                   absent from /root/reference and from this image;
 *   - pytorch3d.ops.sample_farthest_points       restated from the published CPU algorithm
 *   - pytorch3d.loss.chamfer_distance            (csrc/knn/knn_cpu.cpp, csrc/ball_query/ball_query_cpu.cpp,
 *                                                 csrc/sample_farthest_points/sample_farthest_points.cpp,
 *                                                 loss/chamfer.py), call sites listed per function below.
 *
 * Pinning status:
 *   - FPS (a1) and index_points (a2) are PINNED: tests/golden/ holds outputs of the reference's own
 *     functions imported from /root/reference (tests/golden/make_golden.py), and this file reproduces them
 *     bit for bit.
 *   - The PyTorch3D ops (a3-a7) are "PARITY UNPINNED" against the real package (it cannot be installed
 *     here: no network).  They are cross-checked against an independent torch brute-force statement of
 *     the same published semantics, and against the reference's own modules run with these ops injected.
 *
 * Arithmetic rule shared by everything here: squared distance is
 *     d2 = fl(fl(fl(dx*dx) + fl(dy*dy)) + fl(dz*dz)),   no FMA contraction, left to right.
 * Build with -ffp-contract=off (see oracle/Makefile) so gcc keeps it that way.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

static inline float dist2(const float *a, const float *b) {
    float dx = a[0] - b[0];
    float dy = a[1] - b[1];
    float dz = a[2] - b[2];
    float d = dx * dx;
    d = d + dy * dy;
    d = d + dz * dz;
    return d;
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * Farthest point sampling.
 *   mode a1: pn_kit.farthest_point_sample_batch, pn_kit.py:309-330
 *            start_idx[b] = the reference's torch.randint draw (:321), init_dist = 1e10 (:320),
 *            centroids[:,i] written before the update (:324), dist<distance update (:327-328),
 *            first-max argmax (:329).
 *   mode a6: pytorch3d sample_farthest_points (pointnet_sa_module.py:10-13): start_idx == NULL (=> 0),
 *            init_dist = FLT_MAX, already selected points forced to 0, k_n = min(N, npoint), idx padded -1.
 *            Forcing selected points to 0 is what the plain update computes anyway (d2(p,p) == 0).
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_fps(const float *xyz, int64_t B, int64_t N, int64_t npoint, const int64_t *start_idx,
                     float init_dist, int pad_beyond_n, int64_t *out_idx, int nthreads) {
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(dynamic, 1)
    for (int64_t b = 0; b < B; ++b) {
        const float *pc = xyz + b * N * 3;
        int64_t *out = out_idx + b * npoint;
        float *dist = (float *)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
        for (int64_t i = 0; i < N; ++i) dist[i] = init_dist;
        int64_t far = start_idx ? start_idx[b] : 0;
        int64_t k_n = npoint;
        if (pad_beyond_n && N < npoint) k_n = N;
        for (int64_t i = 0; i < npoint; ++i) out[i] = -1;
        for (int64_t i = 0; i < k_n; ++i) {
            out[i] = far;
            const float *c = pc + far * 3;
            float best = -1.0f;
            int64_t besti = 0;
            for (int64_t p = 0; p < N; ++p) {
                float d = dist2(pc + p * 3, c);
                if (d < dist[p]) dist[p] = d;
                if (dist[p] > best) {
                    best = dist[p];
                    besti = p;
                }
            }
            far = besti;
        }
        free(dist);
    }
}

/* ------------------------------------------------------------------------------------------------
 * index_points / knn_gather: out[b, m, :] = feat[b, idx[b, m], :]
 *   pn_kit.py:332-360; pytorch3d knn_gather (pointnet_sa_module.py:28).  idx is flattened to [B, M].
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_gather(const float *feat, const int64_t *idx, int64_t B, int64_t N, int64_t C, int64_t M,
                        float *out) {
    for (int64_t b = 0; b < B; ++b)
        for (int64_t m = 0; m < M; ++m) {
            int64_t j = idx[b * M + m];
            memcpy(out + (b * M + m) * C, feat + (b * N + j) * C, sizeof(float) * (size_t)C);
        }
}

/* ------------------------------------------------------------------------------------------------
 * knn_points: brute force K nearest of each p1 point in p2, ascending in (d2, idx).
 *   Call sites: train.py:185, compress.py:71, pn_kit.py:190, pppe_pcd_ae.py:599, eval.py:132.
 *   Published algorithm (knn_cpu.cpp): max-heap of (dist, idx) tuples, candidates visited in index
 *   order, insert iff size<K or dist < top.dist (strict), pop when over K; drained back to front.
 *   Slots beyond min(K, P2) keep the initial fill (dist 0, idx 0).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    float d;
    int32_t i;
} heap_ent;

static inline int ent_less(heap_ent a, heap_ent b) { /* tuple ordering (d, i) */
    return a.d < b.d || (a.d == b.d && a.i < b.i);
}

static void heap_sift_up(heap_ent *h, int pos) {
    while (pos > 0) {
        int par = (pos - 1) >> 1;
        if (ent_less(h[par], h[pos])) {
            heap_ent t = h[par];
            h[par] = h[pos];
            h[pos] = t;
            pos = par;
        } else
            break;
    }
}

static void heap_sift_down(heap_ent *h, int n, int pos) {
    for (;;) {
        int l = 2 * pos + 1, r = l + 1, m = pos;
        if (l < n && ent_less(h[m], h[l])) m = l;
        if (r < n && ent_less(h[m], h[r])) m = r;
        if (m == pos) break;
        heap_ent t = h[m];
        h[m] = h[pos];
        h[pos] = t;
        pos = m;
    }
}

ORC_API void orc_knn(const float *p1, const float *p2, int64_t B, int64_t P1, int64_t P2, int64_t K,
                     float *out_d2, int64_t *out_idx, int nthreads) {
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        heap_ent *h = (heap_ent *)malloc(sizeof(heap_ent) * (size_t)(K + 1));
#pragma omp for schedule(dynamic, 16)
        for (int64_t q = 0; q < B * P1; ++q) {
            int64_t b = q / P1;
            const float *a = p1 + q * 3;
            const float *pc = p2 + b * P2 * 3;
            int n = 0;
            for (int64_t j = 0; j < P2; ++j) {
                float d = dist2(a, pc + j * 3);
                if (n < K) {
                    h[n].d = d;
                    h[n].i = (int32_t)j;
                    heap_sift_up(h, n);
                    ++n;
                } else if (d < h[0].d) {
                    h[0].d = d;
                    h[0].i = (int32_t)j;
                    heap_sift_down(h, n, 0);
                }
            }
            float *od = out_d2 + q * K;
            int64_t *oi = out_idx + q * K;
            for (int64_t k = 0; k < K; ++k) {
                od[k] = 0.0f;
                oi[k] = 0;
            }
            for (int k = n - 1; k >= 0; --k) { /* drain: largest first, written back to front */
                od[k] = h[0].d;
                oi[k] = h[0].i;
                h[0] = h[k];
                heap_sift_down(h, k, 0);
            }
        }
        free(h);
    }
}

/* ------------------------------------------------------------------------------------------------
 * ball_query: first K p2 indices in index order with d2 < radius^2 (strict), idx padded -1, d2 padded 0.
 *   Call site: pointnet_sa_module.py:16-19,71.   Published algorithm: ball_query_cpu.cpp.
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_ball_query(const float *p1, const float *p2, int64_t B, int64_t P1, int64_t P2, int64_t K,
                            float radius, int64_t *out_idx, float *out_d2, int nthreads) {
    const float r2 = radius * radius;
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(dynamic, 16)
    for (int64_t q = 0; q < B * P1; ++q) {
        int64_t b = q / P1;
        const float *a = p1 + q * 3;
        const float *pc = p2 + b * P2 * 3;
        int64_t *oi = out_idx + q * K;
        float *od = out_d2 ? out_d2 + q * K : NULL;
        for (int64_t k = 0; k < K; ++k) {
            oi[k] = -1;
            if (od) od[k] = 0.0f;
        }
        int64_t cnt = 0;
        for (int64_t j = 0; j < P2 && cnt < K; ++j) {
            float d = dist2(a, pc + j * 3);
            if (d < r2) {
                oi[cnt] = j;
                if (od) od[cnt] = d;
                ++cnt;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * nn1: nearest neighbour (K=1 knn_points), the inner step of chamfer_distance (loss/chamfer.py) and of the
 *   D1 PSNR loop (eval.py:68-81).  First minimum in index order (ties -> lowest index).
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_nn1(const float *p1, const float *p2, int64_t B, int64_t P1, int64_t P2, float *out_d2,
                     int64_t *out_idx, int nthreads) {
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(dynamic, 64)
    for (int64_t q = 0; q < B * P1; ++q) {
        int64_t b = q / P1;
        const float *a = p1 + q * 3;
        const float *pc = p2 + b * P2 * 3;
        float best = INFINITY;
        int64_t bi = 0;
        for (int64_t j = 0; j < P2; ++j) {
            float d = dist2(a, pc + j * 3);
            if (d < best) {
                best = d;
                bi = j;
            }
        }
        out_d2[q] = best;
        if (out_idx) out_idx[q] = bi;
    }
}

/* ------------------------------------------------------------------------------------------------
 * chamfer_distance(x, y) with point_reduction="mean", batch_reduction="mean" (loss/chamfer.py):
 *   mean_b [ mean_i min_j |x_i - y_j|^2 + mean_j min_i |y_j - x_i|^2 ].
 *   Call sites: AE.py:67, PPPF_AE.py:168, pppe_pcd_ae.py:820, eval.py:204.
 *   Sums are taken in double here (the reference's fp32 torch sums have unspecified order; the parity
 *   tolerance for the scalar is 1e-5 relative, the per-point minima are compared bit for bit).
 *   per_cloud (nullable) receives the per-cloud value mean_i + mean_j.
 * ---------------------------------------------------------------------------------------------- */
ORC_API double orc_chamfer(const float *x, const float *y, int64_t B, int64_t P1, int64_t P2, float *dx,
                           int64_t *ix, float *dy, int64_t *iy, double *per_cloud, int nthreads) {
    orc_nn1(x, y, B, P1, P2, dx, ix, nthreads);
    orc_nn1(y, x, B, P2, P1, dy, iy, nthreads);
    double total = 0.0;
    for (int64_t b = 0; b < B; ++b) {
        double sx = 0.0, sy = 0.0;
        for (int64_t i = 0; i < P1; ++i) sx += dx[b * P1 + i];
        for (int64_t j = 0; j < P2; ++j) sy += dy[b * P2 + j];
        double v = sx / (double)P1 + sy / (double)P2;
        if (per_cloud) per_cloud[b] = v;
        total += v;
    }
    return total / (double)B;
}

/* Chamfer backward (pytorch3d knn_points_backward through K=1 in both directions), norm=2:
 *   gx[b,i] += 2*w_x*(x_i - y_ix[i]);  gy[b,ix[i]] -= same;   and symmetrically for the y->x direction.
 *   w_x = grad / (B*P1), w_y = grad / (B*P2).  gx, gy must be zero-initialised by the caller. */
ORC_API void orc_chamfer_bwd(const float *x, const float *y, const int64_t *ix, const int64_t *iy, int64_t B,
                             int64_t P1, int64_t P2, float grad, float *gx, float *gy) {
    const float wx = grad / (float)(B * P1), wy = grad / (float)(B * P2);
    for (int64_t b = 0; b < B; ++b) {
        const float *xb = x + b * P1 * 3, *yb = y + b * P2 * 3;
        float *gxb = gx + b * P1 * 3, *gyb = gy + b * P2 * 3;
        for (int64_t i = 0; i < P1; ++i) {
            int64_t j = ix[b * P1 + i];
            for (int c = 0; c < 3; ++c) {
                float g = 2.0f * wx * (xb[i * 3 + c] - yb[j * 3 + c]);
                gxb[i * 3 + c] += g;
                gyb[j * 3 + c] -= g;
            }
        }
        for (int64_t j = 0; j < P2; ++j) {
            int64_t i = iy[b * P2 + j];
            for (int c = 0; c < 3; ++c) {
                float g = 2.0f * wy * (yb[j * 3 + c] - xb[i * 3 + c]);
                gyb[j * 3 + c] += g;
                gxb[i * 3 + c] -= g;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * D1 (point-to-point) PSNR inner step, eval.py:68-81,84,88-92: 1-NN of every recon point in the original,
 *   float64 like the reference's numpy, exact brute-force NN in place of the Open3D KD-tree (an exact NN
 *   search returns the same distances).  Returns 10*log10(diag^2 / mse), diag = |bbox(orig)|.
 * ---------------------------------------------------------------------------------------------- */
ORC_API double orc_d1_psnr(const float *orig, int64_t No, const float *recon, int64_t Nr, double *out_mse) {
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = 0; i < No; ++i)
        for (int c = 0; c < 3; ++c) {
            double v = orig[i * 3 + c];
            if (v < mn[c]) mn[c] = v;
            if (v > mx[c]) mx[c] = v;
        }
    double sum = 0.0;
    for (int64_t r = 0; r < Nr; ++r) {
        double best = INFINITY;
        for (int64_t i = 0; i < No; ++i) {
            double dx = (double)recon[r * 3 + 0] - (double)orig[i * 3 + 0];
            double dy = (double)recon[r * 3 + 1] - (double)orig[i * 3 + 1];
            double dz = (double)recon[r * 3 + 2] - (double)orig[i * 3 + 2];
            double d = dx * dx + dy * dy + dz * dz;
            if (d < best) best = d;
        }
        sum += best;
    }
    double mse = sum / (double)Nr;
    if (out_mse) *out_mse = mse;
    double diag2 = 0.0;
    for (int c = 0; c < 3; ++c) diag2 += (mx[c] - mn[c]) * (mx[c] - mn[c]);
    return mse > 0 ? 10.0 * log10(diag2 / mse) : INFINITY;
}

/* ================================================================================================
 * Octree centre coding.  Restates, literally, the reference's numpy coder:
 *   octree_np.getDecodeFromPc   /root/reference/octree_np.py:114-133   (float32 floor-divide snap + np.unique)
 *   octree_np.encode            /root/reference/octree_np.py:10-45     (stack DFS, children popped 7..0,
 *                                                                        bits appended per level, inclusive masks)
 *   octree_np.decode            /root/reference/octree_np.py:47-112    (as written: it consumes only the first
 *                                                                        8 bits, emits depth-1 octant centres and
 *                                                                        pads to 64 rows -- SURVEY.md appendix B-2)
 *   pn_kit.encode_sampled_np    /root/reference/pn_kit.py:380-401      (depth search 1..16)
 *   pn_kit.binary_array_to_byte_array / byte_array_to_binary_array   pn_kit.py:463-475
 * PINNED: tests/golden/ref_octree.npz holds the outputs of these reference functions themselves.
 * ============================================================================================== */

/* numpy's float32 floor_divide (npy_divmodf, numpy/core/src/npymath/npy_math_internal.h.src) */
static float np_floor_divide_f32(float a, float b) {
    float mod = fmodf(a, b);
    if (b == 0.0f) return a / b;
    float div = (a - mod) / b;
    if (mod != 0.0f) {
        if ((b < 0) != (mod < 0)) div -= 1.0f;
    }
    float floordiv;
    if (div != 0.0f) {
        floordiv = floorf(div);
        if (div - floordiv > 0.5f) floordiv += 1.0f;
    } else {
        floordiv = copysignf(0.0f, a / b);
    }
    return floordiv;
}

static int cmp_row3(const void *pa, const void *pb) {
    const float *a = (const float *)pa, *b = (const float *)pb;
    for (int c = 0; c < 3; ++c) {
        if (a[c] < b[c]) return -1;
        if (a[c] > b[c]) return 1;
    }
    return 0;
}

/* getDecodeFromPc for one [S,3] cloud: snapped points (row order kept) into `snapped` (nullable), the np.unique'd rows
 * into `uniq` (lexicographically ascending); returns the number of unique rows. */
ORC_API int64_t orc_octree_quantise(const float *pc, int64_t S, double resolution, int depth, float *snapped, float *uniq) {
    int capped = depth < 30 ? depth : 30;
    double divisor = pow(2.0, (double)capped);
    if (divisor < 1.0) divisor = 1.0;
    double cube = resolution / divisor;
    if (cube < 1e-6) cube = 1e-6;
    const float cr = (float)cube;          /* python float is a weak scalar against a float32 array (NEP 50) */
    const float half = (float)(cube / 2);
    for (int64_t i = 0; i < S * 3; ++i) {
        float v = np_floor_divide_f32(pc[i], cr) * cr + half;
        if (isnan(v)) v = 0.0f;            /* np.nan_to_num */
        else if (isinf(v)) v = v > 0 ? FLT_MAX : -FLT_MAX;
        uniq[i] = v;
        if (snapped) snapped[i] = v;
    }
    qsort(uniq, (size_t)S, 3 * sizeof(float), cmp_row3);
    int64_t n = 0;
    for (int64_t i = 0; i < S; ++i)
        if (i == 0 || cmp_row3(uniq + 3 * i, uniq + 3 * (n - 1)) != 0) {
            memmove(uniq + 3 * n, uniq + 3 * i, 3 * sizeof(float));
            ++n;
        }
    return n;
}

typedef struct { double x, y, z; int d; } OrcNode;

/* octree_np.encode: returns the number of bits written to out_bits (one uint8 per bit), or -1 if cap is too small. */
ORC_API int64_t orc_octree_encode(const float *pc, int64_t S, double resolution, int depth, uint8_t *out_bits, int64_t cap,
                                  int64_t *out_unique) {
    float *uniq = (float *)malloc((size_t)(S > 0 ? S : 1) * 3 * sizeof(float));
    const int64_t U = orc_octree_quantise(pc, S, resolution, depth, NULL, uniq);
    if (out_unique) *out_unique = U;
    /* per-level bit lists */
    int64_t *len = (int64_t *)calloc((size_t)depth + 1, sizeof(int64_t));
    int64_t *capl = (int64_t *)calloc((size_t)depth + 1, sizeof(int64_t));
    uint8_t **lev = (uint8_t **)calloc((size_t)depth + 1, sizeof(uint8_t *));
    int64_t scap = 64, sp = 0;
    OrcNode *stack = (OrcNode *)malloc((size_t)scap * sizeof(OrcNode));
    stack[sp++] = (OrcNode){0.0, 0.0, 0.0, 0};
    while (sp > 0) {
        const OrcNode nd = stack[--sp];
        const double reso = resolution / pow(2.0, (double)nd.d);
        int any = 0;
        for (int64_t i = 0; i < U && !any; ++i) {
            /* float32 array against python floats: the scalars are cast to float32 (weak scalars) */
            const float x = uniq[3 * i], y = uniq[3 * i + 1], z = uniq[3 * i + 2];
            any = (float)nd.x <= x && x <= (float)(nd.x + reso) && (float)nd.y <= y && y <= (float)(nd.y + reso) &&
                  (float)nd.z <= z && z <= (float)(nd.z + reso);
        }
        if (len[nd.d] == capl[nd.d]) {
            capl[nd.d] = capl[nd.d] ? 2 * capl[nd.d] : 64;
            lev[nd.d] = (uint8_t *)realloc(lev[nd.d], (size_t)capl[nd.d]);
        }
        lev[nd.d][len[nd.d]++] = (uint8_t)any;
        if (any && nd.d < depth) {
            const double h = reso / 2;
            if (sp + 8 > scap) {
                scap *= 2;
                stack = (OrcNode *)realloc(stack, (size_t)scap * sizeof(OrcNode));
            }
            for (int c = 0; c < 8; ++c)  /* pushed 0..7 (x is the slowest bit), popped 7..0 */
                stack[sp++] = (OrcNode){nd.x + ((c >> 2) & 1) * h, nd.y + ((c >> 1) & 1) * h, nd.z + (c & 1) * h, nd.d + 1};
        }
    }
    int64_t n = 0, ok = 1;
    for (int l = 0; l <= depth; ++l) {
        if (n + len[l] > cap) ok = 0;
        if (ok) memcpy(out_bits + n, lev[l], (size_t)len[l]);
        n += len[l];
        free(lev[l]);
    }
    free(lev); free(len); free(capl); free(stack); free(uniq);
    return ok ? n : -1;
}

/* pn_kit.encode_sampled_np for one cloud: depth search 1..16 (pn_kit.py:386-396); returns nbits, *out_depth = the depth
 * of the returned code (the reference's DEPTH counter overshoots to 17 when nothing converged; the code is depth 16). */
ORC_API int64_t orc_octree_encode_sampled(const float *pc, int64_t S, double scale, int64_t N, double min_bpp, uint8_t *out_bits,
                                          int64_t cap, int *out_depth) {
    int64_t nb = -1;
    int depth = 1;
    for (int attempt = 0; attempt < 16; ++attempt) {
        int64_t U = 0;
        nb = orc_octree_encode(pc, S, scale, depth, out_bits, cap, &U);
        if (nb < 0) return -1;
        const double bpp = (double)nb / (double)N;
        if (bpp > min_bpp && U == S) break;
        if (attempt < 15) depth += 1;
    }
    if (out_depth) *out_depth = depth;
    return nb;
}

/* octree_np.decode exactly as written (see the header note): out [64,3]. */
ORC_API void orc_octree_decode_ref(const uint8_t *bits, int64_t nbits, double resolution, float *out) {
    const int64_t g = nbits < 8 ? nbits : 8;
    int64_t n = 0;
    const double reso = resolution / 2.0;  /* curr_cube_reso at depth 1 */
    for (int64_t j = 0; j < g; ++j) {      /* j-th popped child is octant 7 - j */
        if (bits[j] != 1) continue;
        const int c = 7 - (int)j;
        out[3 * n + 0] = (float)(((c >> 2) & 1) * reso + reso / 2);
        out[3 * n + 1] = (float)(((c >> 1) & 1) * reso + reso / 2);
        out[3 * n + 2] = (float)((c & 1) * reso + reso / 2);
        ++n;
    }
    if (n == 0) {
        memset(out, 0, 64 * 3 * sizeof(float));
        return;
    }
    for (int64_t i = n; i < 64; ++i) memcpy(out + 3 * i, out + 3 * (n - 1), 3 * sizeof(float));
}

/* pn_kit.binary_array_to_byte_array: chunks of 8 bits, MSB first; a short last chunk is read as a short binary number. */
ORC_API int64_t orc_bits_to_bytes(const uint8_t *bits, int64_t nbits, uint8_t *out) {
    int64_t nb = 0;
    for (int64_t i = 0; i < nbits; i += 8) {
        unsigned v = 0;
        for (int64_t t = i; t < i + 8 && t < nbits; ++t) v = (v << 1) | (bits[t] & 1u);
        out[nb++] = (uint8_t)v;
    }
    return nb;
}

/* ================================================================================================
 * Entropy stage: pn_kit.pmf_to_cdf (/root/reference/pn_kit.py:452-461) and the arithmetic coder the
 * reference calls through torchac.encode_float_cdf / decode_float_cdf (compress.py:134-136, decompress.py:92-93).
 * torchac (torchac==0.9.3, requirements_gpu.txt:30) is a third-party package that is absent from /root/reference and from
 * this image: PARITY UNPINNED.  This restates its published algorithm (torchac/torchac.py::_convert_to_int_and_normalize
 * and torchac/backend/torchac_backend.cpp: 32-bit low / high range coder, 16-bit CDFs, pending-bit carry handling,
 * MSB-first bit packing); it is anchored by round trips and by the reference's call sites.
 * ============================================================================================== */

/* pmf [rows, L] float32 -> cdf uint16 [rows, L + 1]:  cdf_float = clamp(cat(0, cumsum(pmf)), max = 1)  (pn_kit.pmf_to_cdf),
 * then torchac's needs_normalization path: round(cdf_float * (2^16 - L)) as int16 bit pattern, + arange(L + 1). */
ORC_API void orc_pmf_to_cdf_u16(const float *pmf, int64_t rows, int L, uint16_t *cdf) {
    const int Lp = L + 1;
    const float new_max = (float)(65536 - (Lp - 1));
    for (int64_t r = 0; r < rows; ++r) {
        double run = 0.0;
        for (int k = 0; k < Lp; ++k) {
            float c = 0.0f;
            if (k > 0) {
                run = run + (double)pmf[r * L + k - 1];   /* torch CPU cumsum accumulates float inputs in double (acc_type) */
                c = (float)run;
                c = c > 1.0f ? 1.0f : c;
            }
            const float scaled = nearbyintf(c * new_max);   /* torch.round: half to even */
            cdf[r * Lp + k] = (uint16_t)((int32_t)scaled + k);
        }
    }
}

typedef struct { uint8_t *out; int64_t cap, n; uint8_t cache; int count; } OrcBitOut;
static void bo_append(OrcBitOut *o, int bit) {
    o->cache = (uint8_t)((o->cache << 1) | (bit & 1));
    if (++o->count == 8) {
        if (o->n < o->cap) o->out[o->n] = o->cache;
        o->n++;
        o->count = 0;
        o->cache = 0;
    }
}
static void bo_append_and_pending(OrcBitOut *o, int bit, uint64_t *pending) {
    bo_append(o, bit);
    while (*pending > 0) {
        bo_append(o, !bit);
        *pending -= 1;
    }
}

/* torchac encode_cdf: cdf uint16 [n_sym, Lp], sym int16 [n_sym] in [0, Lp - 2]; returns the number of bytes (written up to cap). */
ORC_API int64_t orc_range_encode(const uint16_t *cdf, const int16_t *sym, int64_t n_sym, int Lp, uint8_t *out, int64_t cap) {
    OrcBitOut bo = {out, cap, 0, 0, 0};
    uint32_t low = 0, high = 0xFFFFFFFFu;
    uint64_t pending = 0;
    const int max_symbol = Lp - 2;
    for (int64_t i = 0; i < n_sym; ++i) {
        const int s = sym[i];
        const uint64_t span = (uint64_t)high - (uint64_t)low + 1;
        const uint32_t c_low = cdf[i * Lp + s];
        const uint32_t c_high = s == max_symbol ? 0x10000u : cdf[i * Lp + s + 1];
        high = (low - 1) + (uint32_t)((span * (uint64_t)c_high) >> 16);
        low = low + (uint32_t)((span * (uint64_t)c_low) >> 16);
        for (;;) {
            if (high < 0x80000000u) {
                bo_append_and_pending(&bo, 0, &pending);
                low <<= 1; high <<= 1; high |= 1;
            } else if (low >= 0x80000000u) {
                bo_append_and_pending(&bo, 1, &pending);
                low <<= 1; high <<= 1; high |= 1;
            } else if (low >= 0x40000000u && high < 0xC0000000u) {
                pending++;
                low <<= 1; low &= 0x7FFFFFFFu; high <<= 1; high |= 0x80000001u;
            } else {
                break;
            }
        }
    }
    pending += 1;
    bo_append_and_pending(&bo, low < 0x40000000u ? 0 : 1, &pending);
    while (bo.count != 0) bo_append(&bo, 0);   /* flush the partial byte */
    return bo.n;
}

typedef struct { const uint8_t *in; int64_t n, pos; uint8_t cache; int cached; } OrcBitIn;
static void bi_get(OrcBitIn *b, uint32_t *value) {
    if (b->cached == 0) {
        if (b->pos == b->n) { *value <<= 1; return; }
        b->cache = b->in[b->pos++];
        b->cached = 8;
    }
    *value <<= 1;
    *value |= (uint32_t)((b->cache >> (b->cached - 1)) & 1);
    b->cached--;
}

/* torchac decode_cdf */
ORC_API void orc_range_decode(const uint16_t *cdf, int64_t n_sym, int Lp, const uint8_t *in, int64_t n_bytes, int16_t *sym) {
    OrcBitIn bi = {in, n_bytes, 0, 0, 0};
    uint32_t low = 0, high = 0xFFFFFFFFu, value = 0;
    const int max_symbol = Lp - 2;
    for (int i = 0; i < 32; ++i) bi_get(&bi, &value);
    for (int64_t i = 0; i < n_sym; ++i) {
        const uint64_t span = (uint64_t)high - (uint64_t)low + 1;
        const uint16_t count = (uint16_t)(((((uint64_t)value - (uint64_t)low + 1) << 16) - 1) / span);
        int left = 0, right = max_symbol + 1;
        while (left + 1 < right) {   /* largest s with cdf[s] <= count */
            const int m = (left + right) / 2;
            const uint16_t v = cdf[i * Lp + m];
            if (v < count) left = m;
            else if (v > count) right = m;
            else { left = m; break; }
        }
        const int s = left;
        sym[i] = (int16_t)s;
        const uint32_t c_low = cdf[i * Lp + s];
        const uint32_t c_high = s == max_symbol ? 0x10000u : cdf[i * Lp + s + 1];
        high = (low - 1) + (uint32_t)((span * (uint64_t)c_high) >> 16);
        low = low + (uint32_t)((span * (uint64_t)c_low) >> 16);
        for (;;) {
            if (low >= 0x80000000u || high < 0x80000000u) {
                low <<= 1; high <<= 1; high |= 1;
                bi_get(&bi, &value);
            } else if (low >= 0x40000000u && high < 0xC0000000u) {
                low <<= 1; low &= 0x7FFFFFFFu; high <<= 1; high |= 0x80000001u;
                value -= 0x40000000u;
                bi_get(&bi, &value);
            } else {
                break;
            }
        }
    }
}
