// Weight-gradient contraction of a shared-MLP layer as a streamed tcgen05 GEMM over the ROWS of two activation tensors:
//     C[Na, Nb] = sum_r A[r, 0..Na) (x) B[r, 0..Nb)          colsum[Na] = sum_r A[r, 0..Na)
// A = the (masked) output gradient dY [M, Na], B = the layer's input activation X [M, Nb], both bf16 row-major as the forward
// kernels wrote them: dW = dY^T X and db = column sums of dY -- the backward of the Conv2d(1x1) / Linear layers the reference
// trains under autograd (/root/reference/train.py:193-221 -> pn_kit.py:124-211,289-305, AE.py:19-55).
//
// The contraction runs over rows, so neither operand is K-major in memory.  No transposed copies are made: a [64 rows x 64
// columns] TMA box (128-byte swizzle) IS the canonical MN-major SWIZZLE_128B operand of tcgen05.mma (64 contiguous M/N elements
// per 128-byte line, 8 lines per swizzle atom along K, atoms 1024 bytes apart along K = SBO, 64-column blocks one slab apart =
// LBO), so both descriptors just set the MN-major bits of the instruction descriptor.  The bias gradient rides along as a
// second, 16-column MMA against a constant block of ones.
//   warp 5: TMA producer (ring of NST stages: 2 A slabs + 2 B slabs of 64 rows);  warp 4: MMA issue (elect.sync), 4 K steps of
//   16 rows per stage, accumulators in TMEM (128 + 16 columns);  warps 0-3: epilogue -- plain fp32 stores when one CTA owns the
//   whole row range of a tile, red.global.add.f32 when the rows are split over CTAs (long M, small Na x Nb: the usual case).
// Bound: HBM on the two activation streams ((Na + Nb) * 2 B per row) -- 256 tensor clocks per 24..32 KB stage.
#include <cuda.h>
#include <cuda_bf16.h>

#include "chain_ws.h"
#include "pcc_common.cuh"
#include "tc_ptx.cuh"

namespace pcc {
namespace wg {

constexpr int ROWS = 64;                 // rows (K of the contraction) per stage
constexpr int SLAB = ROWS * 128;         // [64 rows x 64 bf16], 8 KB
constexpr int STAGE = 4 * SLAB;          // A: columns na0 .. na0+127 (2 slabs), B: nb0 .. nb0+127 (2 slabs)
constexpr int NST = 6;
constexpr int OFF_ONES = NST * STAGE;    // [16 x 16] bf16 ones, K-major no-swizzle (512 B)
constexpr int OFF_BAR = OFF_ONES + 1024;
constexpr int SMEM = OFF_BAR + 256 + 1024;
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 256;           // 128 (C tile) + 16 (column sums) -> next power of two

struct Params {
    float *c;            // [Na, ldc]
    float *colsum;       // [Na] or NULL
    long long M, ldc;
    int Na, Nb, n_ta, n_tb, n_chunks;
    long long rows_per_chunk;   // multiple of 64
    int atomic;          // 1: several chunks add into C (zeroed by the host side)
};

// MN-major SWIZZLE_128B operand: LBO = bytes between 64-element blocks along M/N, SBO = bytes between 8-row groups along K
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = static_cast<uint64_t>((saddr & 0x3ffffu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}
// kind::f16, D = f32, A = B = bf16, shape M x N, operand majors as given (1 = MN-major)
__device__ __forceinline__ uint32_t idesc_major(int M, int N, uint32_t a_mn, uint32_t b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn << 15) | (b_mn << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
           (static_cast<uint32_t>(M >> 4) << 24);
}

__global__ void __launch_bounds__(THREADS, 1)
wgrad_kernel(const __grid_constant__ Params prm, const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t sb = smem_u32(smem);
    const int tid = threadIdx.x, warp = __shfl_sync(FULL_MASK, tid >> 5, 0), lane = tid & 31;
    const uint32_t bar = sb + OFF_BAR;
    const uint32_t full = bar, empty = bar + 64, acc_full = bar + 128, acc_empty = bar + 136;   // [NST], [NST], 1, 1
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 144);

    for (int i = tid; i < 512 / 4; i += THREADS) reinterpret_cast<uint32_t *>(smem + OFF_ONES)[i] = 0x3f803f80u;
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) {
            mbar_init(full + 8 * i, 1);
            mbar_init(empty + 8 * i, 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 4);
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long n_items = static_cast<long long>(prm.n_ta) * prm.n_tb * prm.n_chunks;
    const bool want_sum = prm.colsum != nullptr;

    if (warp == 5) {
        if (lane == 0) {   // ---- TMA producer ----
            uint32_t it = 0;
            for (long long w = blockIdx.x; w < n_items; w += gridDim.x) {
                const int ta = static_cast<int>(w % prm.n_ta), tb = static_cast<int>((w / prm.n_ta) % prm.n_tb);
                const long long ch = w / (static_cast<long long>(prm.n_ta) * prm.n_tb);
                const long long r0 = ch * prm.rows_per_chunk, r1 = min(prm.M, r0 + prm.rows_per_chunk);
                for (long long r = r0; r < r1; r += ROWS, ++it) {
                    const uint32_t st = it % NST, ph = ((it / NST) & 1u) ^ 1u;
                    mbar_wait(empty + 8 * st, ph);
                    mbar_arrive_expect_tx(full + 8 * st, STAGE);
                    const uint32_t base = sb + st * STAGE;
                    tma_load_2d(base, &tm_a, ta * 128, static_cast<int>(r), full + 8 * st);
                    tma_load_2d(base + SLAB, &tm_a, ta * 128 + 64, static_cast<int>(r), full + 8 * st);
                    tma_load_2d(base + 2 * SLAB, &tm_b, tb * 128, static_cast<int>(r), full + 8 * st);
                    tma_load_2d(base + 3 * SLAB, &tm_b, tb * 128 + 64, static_cast<int>(r), full + 8 * st);
                }
            }
        }
    } else if (warp == 4) {
        // ---- MMA issuer (warp-uniform control flow, one elected lane issues) ----
        const uint32_t tb_ = __shfl_sync(FULL_MASK, tmem_base, 0);
        const uint32_t id_c = idesc_major(128, 128, 1u, 1u), id_s = idesc_major(128, 16, 1u, 0u);
        const uint64_t d_ones = umma_desc(sb + OFF_ONES, 128, 256);
        uint32_t it = 0, ph_acc_empty = 1;
        for (long long w = blockIdx.x; w < n_items; w += gridDim.x) {
            const long long ch = w / (static_cast<long long>(prm.n_ta) * prm.n_tb);
            const long long r0 = ch * prm.rows_per_chunk, r1 = min(prm.M, r0 + prm.rows_per_chunk);
            mbar_wait(acc_empty, ph_acc_empty);
            ph_acc_empty ^= 1u;
            tc_fence_after();
            bool first = true;
            for (long long r = r0; r < r1; r += ROWS, ++it) {
                const uint32_t st = it % NST, ph = (it / NST) & 1u;
                mbar_wait(full + 8 * st, ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t d_a = desc_mn_sw128(sb + st * STAGE, SLAB), d_b = desc_mn_sw128(sb + st * STAGE + 2 * SLAB, SLAB);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {   // 16 rows per K step = two 8-row groups = 2048 bytes
                        umma_bf16(tb_, d_a + ks * 128, d_b + ks * 128, id_c, (first && ks == 0) ? 0u : 1u);
                        if (want_sum) umma_bf16(tb_ + 128, d_a + ks * 128, d_ones, id_s, (first && ks == 0) ? 0u : 1u);
                    }
                    umma_commit(empty + 8 * st);
                    if (r + ROWS >= r1) umma_commit(acc_full);
                }
                __syncwarp();
                first = false;
            }
        }
    } else {
        // ---- epilogue: lane = row of the C tile ----
        const int row = warp * 32 + lane;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        uint32_t ph_acc_full = 0;
        // (the accumulator starts out free: the MMA warp's first wait is on parity 1 of the fresh barrier)
        for (long long w = blockIdx.x; w < n_items; w += gridDim.x) {
            const int ta = static_cast<int>(w % prm.n_ta), tb = static_cast<int>((w / prm.n_ta) % prm.n_tb);
            const long long ch = w / (static_cast<long long>(prm.n_ta) * prm.n_tb);
            const bool has_rows = ch * prm.rows_per_chunk < prm.M;
            if (has_rows) {
                mbar_wait(acc_full, ph_acc_full);
                ph_acc_full ^= 1u;
                tc_fence_after();
                const int na = ta * 128 + row;
                float *crow = prm.c + static_cast<long long>(na) * prm.ldc + tb * 128;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    uint32_t v[32];
                    tmem_ld32(lane_base + c4 * 32, v);
                    if (na < prm.Na) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int nb = tb * 128 + c4 * 32 + i;
                            if (nb < prm.Nb) {
                                if (prm.atomic) atomicAdd(crow + c4 * 32 + i, __uint_as_float(v[i]));
                                else crow[c4 * 32 + i] = __uint_as_float(v[i]);
                            }
                        }
                    }
                }
                if (want_sum && tb == 0) {
                    uint32_t s4[4];
                    tmem_ld4(lane_base + 128, s4);
                    if (na < prm.Na) {
                        if (prm.atomic) atomicAdd(prm.colsum + na, __uint_as_float(s4[0]));
                        else prm.colsum[na] = __uint_as_float(s4[0]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive1(acc_empty);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

}  // namespace wg
}  // namespace pcc

PCC_API int pcc_wgrad_bf16(const void *a, int64_t lda, int Na, const void *b, int64_t ldb, int Nb, int64_t M, float *c, int64_t ldc,
                           float *colsum, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(a && b && c, "pcc_wgrad_bf16: null pointer");
    PCC_REQUIRE(M >= 1 && M < (1ll << 31) - 256 && Na >= 1 && Nb >= 1 && ldc >= Nb, "pcc_wgrad_bf16: bad shape M=%lld Na=%d Nb=%d",
                static_cast<long long>(M), Na, Nb);
    PCC_REQUIRE(Na % 8 == 0 && Nb % 8 == 0 && lda >= Na && ldb >= Nb && lda % 8 == 0 && ldb % 8 == 0 &&
                    reinterpret_cast<uintptr_t>(a) % 16 == 0 && reinterpret_cast<uintptr_t>(b) % 16 == 0,
                "pcc_wgrad_bf16: operands must be 16-byte aligned bf16 with widths / row pitches that are multiples of 8 elements");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUtensorMap ta, tb;
    if (int r = make_tmap_bf16_2d(&ta, a, static_cast<uint64_t>(M), static_cast<uint64_t>(Na), static_cast<uint64_t>(lda), wg::ROWS)) return r;
    if (int r = make_tmap_bf16_2d(&tb, b, static_cast<uint64_t>(M), static_cast<uint64_t>(Nb), static_cast<uint64_t>(ldb), wg::ROWS)) return r;
    wg::Params p{};
    p.c = c;
    p.colsum = colsum;
    p.M = M;
    p.ldc = ldc;
    p.Na = Na;
    p.Nb = Nb;
    p.n_ta = (Na + 127) / 128;
    p.n_tb = (Nb + 127) / 128;
    const long long tiles = static_cast<long long>(p.n_ta) * p.n_tb;
    const long long stages = (M + wg::ROWS - 1) / wg::ROWS;
    const long long sms = num_sms();
    long long chunks = (2 * sms + tiles - 1) / tiles;           // ~2 work items per SM ...
    if (chunks > (stages + 7) / 8) chunks = (stages + 7) / 8;    // ... of at least 8 stages each
    if (chunks < 1) chunks = 1;
    p.rows_per_chunk = ((stages + chunks - 1) / chunks) * wg::ROWS;
    p.n_chunks = static_cast<int>((M + p.rows_per_chunk - 1) / p.rows_per_chunk);
    p.atomic = p.n_chunks > 1;
    if (p.atomic) {
        cudaError_t e = cudaMemset2DAsync(c, static_cast<size_t>(ldc) * 4, 0, static_cast<size_t>(Nb) * 4, static_cast<size_t>(Na), st);
        if (e == cudaSuccess && colsum) e = cudaMemsetAsync(colsum, 0, static_cast<size_t>(Na) * 4, st);
        if (e != cudaSuccess) {
            set_error("pcc_wgrad_bf16: cudaMemsetAsync failed: %s", cudaGetErrorString(e));
            return static_cast<int>(e);
        }
    }
    const cudaError_t e = cudaFuncSetAttribute(wg::wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::SMEM);
    if (e != cudaSuccess) {
        set_error("pcc_wgrad_bf16: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    const long long items = tiles * p.n_chunks;
    const int grid = static_cast<int>(items < sms ? items : sms);
    wg::wgrad_kernel<<<grid, wg::THREADS, wg::SMEM, st>>>(p, ta, tb);
    return check_launch("wgrad_kernel");
}
