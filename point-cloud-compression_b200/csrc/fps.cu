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
//   CTA publishes its best key with one atomicMax and the C CTAs meet at a monotonic-counter barrier.
#include "pcc_common.cuh"

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
template <int PPT>
__device__ __forceinline__ void fps_local_update(const float (&px)[PPT], const float (&py)[PPT],
                                                 const float (&pz)[PPT], float (&md)[PPT], float cx, float cy,
                                                 float cz, unsigned &bits, int &slot) {
    float best = -1.0f;
    slot = 0;
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        const float d = dist2_rn(px[p], py[p], pz[p], cx, cy, cz);
        md[p] = fminf(md[p], d);  // == `if (d < md) md = d` (pn_kit.py:327-328) for non-NaN inputs
        if (md[p] > best) {       // strict: the lowest p (lowest global index) keeps a tie
            best = md[p];
            slot = p;
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

// Optional fused epilogue: the sampled point itself (the index_points gather that follows every FPS call), optionally
// snapped to the octree grid: floor(c / cube) * cube + cube / 2 (octree_np.getDecodeFromPc, octree_np.py:114-133).
__device__ __forceinline__ void store_centre(float *o, float x, float y, float z, float cube) {
    if (cube > 0.0f) {
        const float h = __fmul_rn(cube, 0.5f);
        x = __fadd_rn(__fmul_rn(floorf(__fdiv_rn(x, cube)), cube), h);
        y = __fadd_rn(__fmul_rn(floorf(__fdiv_rn(y, cube)), cube), h);
        z = __fadd_rn(__fmul_rn(floorf(__fdiv_rn(z, cube)), cube), h);
    }
    o[0] = x;
    o[1] = y;
    o[2] = z;
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

    int far = start_idx ? static_cast<int>(start_idx[b]) : 0;
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
struct FpsGridWs {             // one per cloud, zero-filled by the host before the launch
    unsigned long long key[3]; // rotating arg-max slots: (d2 bits << 32) | ~idx
    unsigned counter;          // monotonic barrier counter
    unsigned pad;
};

constexpr int GRID_THREADS = 1024;
constexpr int GRID_PPT = 8;
constexpr int GRID_PTS_PER_CTA = GRID_THREADS * GRID_PPT;

__global__ void __launch_bounds__(GRID_THREADS, 1)
fps_grid_kernel(const float *__restrict__ xyz, int N, int npoint, const int64_t *__restrict__ start_idx,
                float init_dist, int64_t *__restrict__ out_idx, FpsGridWs *ws, int ctas_per_cloud, int cloud0,
                float *__restrict__ out_xyz, float quant_cube) {
    constexpr int NWARPS = GRID_THREADS / 32;
    __shared__ FpsSlots slots;
    __shared__ unsigned long long s_win;
    const int cloud_local = blockIdx.x / ctas_per_cloud;
    const int part = blockIdx.x % ctas_per_cloud;
    const int b = cloud0 + cloud_local;
    const int tid = threadIdx.x;
    const float *pc = xyz + static_cast<size_t>(b) * N * 3;
    int64_t *out = out_idx + static_cast<size_t>(b) * npoint;
    FpsGridWs *w = ws + cloud_local;
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
                float *o = out_xyz + (static_cast<size_t>(b) * npoint + k) * 3;
                o[0] = o[1] = o[2] = 0.0f;
            }
        }

    int far = start_idx ? static_cast<int>(start_idx[b]) : 0;
    float cx = pc[static_cast<size_t>(far) * 3 + 0], cy = pc[static_cast<size_t>(far) * 3 + 1],
          cz = pc[static_cast<size_t>(far) * 3 + 2];
    for (int i = 0; i < k_n; ++i) {
        if (part == 0 && tid == 0) {
            out[i] = far;
            if (out_xyz) store_centre(out_xyz + (static_cast<size_t>(b) * npoint + i) * 3, cx, cy, cz, quant_cube);
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
        if (tid == 0) {
            const unsigned long long key = (static_cast<unsigned long long>(w_bits) << 32) | (0xffffffffu - w_idx);
            if (part == 0) w->key[(i + 1) % 3] = 0ull;  // slot of iteration i+1: last read before barrier i-1
            atomicMax(&w->key[i % 3], key);
            __threadfence();
            atomicAdd(&w->counter, 1u);
            const unsigned target = static_cast<unsigned>(ctas_per_cloud) * static_cast<unsigned>(i + 1);
            while (*reinterpret_cast<volatile unsigned *>(&w->counter) < target) {
            }
            __threadfence();
            s_win = *reinterpret_cast<volatile unsigned long long *>(&w->key[i % 3]);
        }
        __syncthreads();
        const unsigned long long win = s_win;
        far = static_cast<int>(0xffffffffu - static_cast<unsigned>(win & 0xffffffffu));
        cx = __ldcg(pc + static_cast<size_t>(far) * 3 + 0);
        cy = __ldcg(pc + static_cast<size_t>(far) * 3 + 1);
        cz = __ldcg(pc + static_cast<size_t>(far) * 3 + 2);
        // s_win is rewritten only after the next block_argmax's __syncthreads, which every thread reaches
        // after reading it here.
    }
}

template <int THREADS, int PPT>
static int launch_block(const float *xyz, int B, int N, int npoint, const int64_t *start_idx, float init_dist,
                        int64_t *out_idx, float *out_xyz, float quant_cube, cudaStream_t st) {
    fps_block_kernel<THREADS, PPT><<<B, THREADS, 0, st>>>(xyz, N, npoint, start_idx, init_dist, out_idx, out_xyz, quant_cube);
    return check_launch("fps_block_kernel");
}

}  // namespace pcc

PCC_API int64_t pcc_fps_workspace_bytes(int B, int N, int npoint) {
    (void)npoint;
    if (N <= pcc::GRID_PTS_PER_CTA || B <= 0) return 0;
    return static_cast<int64_t>(sizeof(pcc::FpsGridWs)) * pcc::num_sms();
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
    if (N <= 2048) return launch_block<512, 4>(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_xyz, quant_cube, st);
    if (N <= 4096) return launch_block<1024, 4>(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_xyz, quant_cube, st);
    if (N <= 8192) return launch_block<1024, 8>(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_xyz, quant_cube, st);

    // multi-CTA path: C co-resident CTAs per cloud, as many clouds per cooperative launch as fit.
    PCC_REQUIRE(workspace, "pcc_fps_f32: N=%d needs a workspace of pcc_fps_workspace_bytes()", N);
    const int sms = num_sms();
    const int C = (N + GRID_PTS_PER_CTA - 1) / GRID_PTS_PER_CTA;
    if (C > sms) {
        set_error("pcc_fps_f32: N=%d exceeds the co-resident capacity %d", N, sms * GRID_PTS_PER_CTA);
        return PCC_ERR_UNSUPPORTED;
    }
    const int clouds_per_launch = sms / C;
    FpsGridWs *ws = static_cast<FpsGridWs *>(workspace);
    for (int c0 = 0; c0 < B; c0 += clouds_per_launch) {
        int nc = B - c0 < clouds_per_launch ? B - c0 : clouds_per_launch;
        cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(FpsGridWs) * nc, st);
        if (e != cudaSuccess) {
            set_error("pcc_fps_f32: memset failed: %s", cudaGetErrorString(e));
            return static_cast<int>(e);
        }
        int ctas = C, cloud0 = c0;
        void *args[] = {&xyz, &N, &npoint, &start_idx, &init_dist, &out_idx, &ws, &ctas, &cloud0, &out_xyz, &quant_cube};
        e = cudaLaunchCooperativeKernel(reinterpret_cast<void *>(fps_grid_kernel), dim3(nc * C), dim3(GRID_THREADS),
                                        args, 0, st);
        if (e != cudaSuccess) {
            set_error("pcc_fps_f32: cooperative launch failed: %s", cudaGetErrorString(e));
            return static_cast<int>(e);
        }
    }
    return check_launch("fps_grid_kernel");
}
