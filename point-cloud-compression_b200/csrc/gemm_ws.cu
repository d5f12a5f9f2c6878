// One shared-MLP layer whose weights do not fit beside the activations in shared memory, as a streamed tcgen05 GEMM:
//     out[M, N] = act(A[M, K] . W[N, K]^T + bias)            (bf16 operands, fp32 accumulation in TMEM)
// optionally followed by the max over every run of `group` consecutive rows (the pooling that ends every PointNet++ set
// abstraction stack, /root/reference/pointnet_sa_module.py:87-91, and pn_kit.PointNet, pn_kit.py:136-143).
// Used for the wide Conv2d+BN+ReLU layers of PointNetPP (PPPF_AE.py:29-33: 128..1024 channels), and for AE.inv_pool's
// Linear layers (AE.py:19-26) -- the layers that were library GEMMs before.
//
//   persistent CTAs, one 128 x BN output tile at a time (BN = 256 or 128), tiles ordered n-fastest so the CTAs that run
//   together share the A tile in L2;
//   warp 9: TMA producer -- ring of NST stages, each one [128 x 64] slab of A and one [BN x 64] slab of W (128-byte swizzle);
//   warp 8: MMA issue (elect.sync, warp-uniform control flow), 4 x (M128 N=BN K16) per stage, accumulators double buffered
//           in TMEM (2 x BN columns), so the epilogue of tile i overlaps the main loop of tile i + 1;
//   warps 0-7: epilogue -- tcgen05.ld, + bias, ReLU, then either bf16 -> swizzled staging slabs -> TMA store, or
//           redux.sync.max over the warp's 32 rows (a run of `group` rows never straddles a warp: group % 32 == 0) and a
//           plain store (group == 32) / atomicMax on the bit pattern (group > 32: values are >= 0 after the ReLU).
// Bound: L2 -> SM operand traffic, (128 + BN) * 128 B per 2 * 128 * BN * 64 FLOP -> 85 FLOP/B at BN = 256, i.e. ~1050 TFLOP/s at
// the measured ~12.4 TB/s L2 cap; layers with small K * N are bound by the activations' HBM traffic instead.
#include <cuda.h>
#include <cuda_bf16.h>

#include "chain_ws.h"
#include "pcc_common.cuh"
#include "tc_ptx.cuh"

namespace pcc {
namespace gws {

constexpr int P = 128;
constexpr int SLAB = P * 128;  // [128 rows x 64 bf16]
constexpr int THREADS = 320;
constexpr int EPI_THREADS = 256;

template <int BN>
struct Lay {
    static constexpr int NST = BN == 256 ? 3 : 4;
    static constexpr int STAGE = SLAB + BN * 128;
    static constexpr int OFF_STG = NST * STAGE;
    static constexpr int STG = (BN / 64) * SLAB;
    static constexpr int OFF_BIAS = OFF_STG + STG;  // [2][BN] floats
    static constexpr int OFF_BAR = OFF_BIAS + 2 * BN * 4;
    static constexpr int SMEM = OFF_BAR + 256 + 1024;
    static_assert(SMEM <= 227 * 1024, "shared memory");
};

struct Params {
    const float *bias;  // [N]
    float *out_pool;    // [M / group, N] fp32 (pooled mode)
    long long M;
    int N, K, n_n;      // n_n = N / BN
    long long n_tiles;
    int relu, group;
    // training (bf16 output only): columns >= n_store are not written (narrow layers inside the 128-column granule); with a
    // mask tensor [M, >= n_store] bf16 the output is zeroed where mask <= 0 -- the ReLU backward of the layer below, fused
    const __nv_bfloat16 *mask;
    long long ld_mask;
    int n_store;
};

template <bool RELU>
__device__ __forceinline__ void stage_row_chunk32(uint32_t stg, int r, int c0, const uint32_t (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + 8 * i;
        const uint32_t addr = stg + (c >> 6) * SLAB + r * 128 + ((((c >> 3) & 7) ^ (r & 7)) << 4);
        uint32_t pk[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const float lo = __uint_as_float(v[8 * i + 2 * h]), hi = __uint_as_float(v[8 * i + 2 * h + 1]);
            pk[h] = RELU ? pack_relu_bf16x2(lo, hi) : pack_bf16x2(lo, hi);
        }
        st_shared_v4(addr, pk[0], pk[1], pk[2], pk[3]);
    }
}

template <int BN, bool POOL, bool MASK>
__global__ void __launch_bounds__(THREADS, 1)
linear_kernel(const __grid_constant__ Params prm, const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
              const __grid_constant__ CUtensorMap tm_o) {
    using L = Lay<BN>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t sb = smem_u32(smem);
    const int tid = threadIdx.x, warp = __shfl_sync(FULL_MASK, tid >> 5, 0), lane = tid & 31;
    const uint32_t bar = sb + L::OFF_BAR;
    const uint32_t full = bar, empty = bar + 32;                 // [NST] each
    const uint32_t acc_full = bar + 64, acc_empty = bar + 80;    // [2] each
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::OFF_BAR + 96);
    constexpr int TMEM_COLS = 2 * BN;

    if (tid == 0) {
        for (int i = 0; i < L::NST; ++i) {
            mbar_init(full + 8 * i, 1);
            mbar_init(empty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(acc_full + 8 * i, 1);
            mbar_init(acc_empty + 8 * i, 8);
        }
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nk = prm.K >> 6;
    const long long n_tiles = prm.n_tiles;

    if (warp == 9) {
        // ---- TMA producer ----
        if (lane == 0) {
            uint32_t it = 0;
            for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int m0 = static_cast<int>(t / prm.n_n) * P, n0 = static_cast<int>(t % prm.n_n) * BN;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const uint32_t st = it % L::NST, ph = ((it / L::NST) & 1u) ^ 1u;
                    mbar_wait(empty + 8 * st, ph);
                    mbar_arrive_expect_tx(full + 8 * st, L::STAGE);
                    tma_load_2d(sb + st * L::STAGE, &tm_a, kb * 64, m0, full + 8 * st);
                    tma_load_2d(sb + st * L::STAGE + SLAB, &tm_w, kb * 64, n0, full + 8 * st);
                }
            }
        }
    } else if (warp == 8) {
        // ---- MMA issuer ----
        const uint32_t tb = __shfl_sync(FULL_MASK, tmem_base, 0);
        const uint32_t idesc = umma_idesc(128, BN);
        const uint64_t d_a = umma_desc_sw128(sb), d_w = umma_desc_sw128(sb + SLAB);
        uint32_t it = 0, ph_acc_empty[2] = {1, 1};
        int j = 0;
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++j) {
            const int buf = j & 1;
            mbar_wait(acc_empty + 8 * buf, ph_acc_empty[buf]);
            ph_acc_empty[buf] ^= 1u;
            tc_fence_after();
            for (int kb = 0; kb < nk; ++kb, ++it) {
                const uint32_t st = it % L::NST, ph = (it / L::NST) & 1u;
                mbar_wait(full + 8 * st, ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t so = static_cast<uint64_t>((st * L::STAGE) >> 4);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_bf16(tb + buf * BN, d_a + so + ks * 2, d_w + so + ks * 2, idesc, (kb | ks) > 0);
                    umma_commit(empty + 8 * st);
                    if (kb == nk - 1) umma_commit(acc_full + 8 * buf);
                }
                __syncwarp();
            }
        }
    } else {
        // ---- epilogue warps ----
        const int q = warp & 3, h = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        float *bias_s = reinterpret_cast<float *>(smem + L::OFF_BIAS);
        const uint32_t stg = sb + L::OFF_STG;
        uint32_t ph_acc_full[2] = {0, 0};
        int j = 0;
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++j) {
            const int buf = j & 1;
            const long long m0 = (t / prm.n_n) * P;
            const int n0 = static_cast<int>(t % prm.n_n) * BN;
            if (tid < BN) bias_s[buf * BN + tid] = __ldg(prm.bias + n0 + tid);
            if (!POOL && tid == 0) tma_store_wait_read0();  // the previous tile's store has finished reading the staging slabs
            named_bar_sync(1, EPI_THREADS);
            mbar_wait(acc_full + 8 * buf, ph_acc_full[buf]);
            ph_acc_full[buf] ^= 1u;
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) {
                const int col = h * (BN / 2) + c * 32;
                uint32_t v[32];
                tmem_ld32(lane_base + buf * BN + col, v);
                const float4 *bb = reinterpret_cast<const float4 *>(bias_s + buf * BN + col);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 b4 = bb[i];
                    v[4 * i] = __float_as_uint(__uint_as_float(v[4 * i]) + b4.x);
                    v[4 * i + 1] = __float_as_uint(__uint_as_float(v[4 * i + 1]) + b4.y);
                    v[4 * i + 2] = __float_as_uint(__uint_as_float(v[4 * i + 2]) + b4.z);
                    v[4 * i + 3] = __float_as_uint(__uint_as_float(v[4 * i + 3]) + b4.w);
                }
                if (POOL && prm.group == 1) {
                    // fp32 rows, no pooling: lane = row, 32 consecutive columns (one 128-byte line per lane)
                    const long long r = m0 + row;
                    if (r < prm.M) {
                        float4 *o = reinterpret_cast<float4 *>(prm.out_pool + r * prm.N + n0 + col);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float4 f = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                                   __uint_as_float(v[4 * i + 3]));
                            if (prm.relu) f = make_float4(fmaxf(f.x, 0.0f), fmaxf(f.y, 0.0f), fmaxf(f.z, 0.0f), fmaxf(f.w, 0.0f));
                            o[i] = f;
                        }
                    }
                } else if (POOL) {
                    float mine = 0.0f;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float m;
                        asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(__uint_as_float(v[i])));
                        if (lane == i) mine = m;
                    }
                    if (prm.relu) mine = fmaxf(mine, 0.0f);  // the ReLU commutes with the max
                    const long long r0 = m0 + q * 32;
                    if (r0 < prm.M) {
                        float *o = prm.out_pool + (r0 / prm.group) * prm.N + n0 + col + lane;
                        if (prm.group == 32) *o = mine;
                        else atomicMax(reinterpret_cast<unsigned *>(o), __float_as_uint(mine));
                    }
                } else {
                    if constexpr (MASK) {   // zero the 8-column chunks' elements whose mask value (the activation below) is <= 0
                        const long long r = m0 + row;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int c = n0 + col + 8 * i;
                            uint4 mk = make_uint4(0u, 0u, 0u, 0u);
                            if (r < prm.M && c + 8 <= prm.n_store) mk = __ldg(reinterpret_cast<const uint4 *>(prm.mask + r * prm.ld_mask + c));
                            const uint32_t mw[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {   // bf16 > 0  <=>  sign clear and not zero
                                const uint32_t lo = mw[e] & 0xffffu, hi = mw[e] >> 16;
                                if (lo == 0u || lo >= 0x8000u) v[8 * i + 2 * e] = 0u;
                                if (hi == 0u || hi >= 0x8000u) v[8 * i + 2 * e + 1] = 0u;
                            }
                        }
                    }
                    if (prm.relu) stage_row_chunk32<true>(stg, row, col, v);
                    else stage_row_chunk32<false>(stg, row, col, v);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive1(acc_empty + 8 * buf);
            if (!POOL) {
                fence_async_smem();
                named_bar_sync(2, EPI_THREADS);
                if (tid == 0) {
#pragma unroll
                    for (int s = 0; s < BN / 64; ++s)
                        if (n0 + s * 64 < prm.n_store) tma_store_2d(&tm_o, n0 + s * 64, static_cast<int>(m0), stg + s * SLAB);
                    tma_store_commit();
                }
            }
        }
        if (!POOL && tid == 0) tma_store_wait_all0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

template <int BN, bool POOL, bool MASK = false>
static int launch(const Params &p, const CUtensorMap &ta, const CUtensorMap &tw, const CUtensorMap &to, cudaStream_t st) {
    using L = Lay<BN>;
    const cudaError_t e = cudaFuncSetAttribute(linear_kernel<BN, POOL, MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM);
    if (e != cudaSuccess) {
        set_error("pcc_linear_bf16: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    const long long sms = num_sms();
    const int grid = static_cast<int>(p.n_tiles < sms ? p.n_tiles : sms);
    linear_kernel<BN, POOL, MASK><<<grid, THREADS, L::SMEM, st>>>(p, ta, tw, to);
    return check_launch("linear_kernel");
}

}  // namespace gws
}  // namespace pcc

static int linear_run(const void *a, int64_t M, int K, int64_t lda, const void *w, int64_t ldw, const float *bias, int N, int relu,
                      int group, void *out, int64_t ld_out, const void *mask, int64_t ld_mask, int n_store, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(a && w && bias && out, "pcc_linear_bf16: null pointer");
    PCC_REQUIRE(M >= 1 && M < (1ll << 31) - 256, "pcc_linear_bf16: M=%lld out of range", static_cast<long long>(M));
    PCC_REQUIRE(K >= 64 && K % 64 == 0, "pcc_linear_bf16: K=%d must be a positive multiple of 64 (pad the operands with zero columns)", K);
    PCC_REQUIRE(N >= 128 && N % 128 == 0, "pcc_linear_bf16: N=%d must be a positive multiple of 128", N);
    PCC_REQUIRE(lda >= K && lda % 8 == 0 && ldw >= K && ldw % 8 == 0 && reinterpret_cast<uintptr_t>(a) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(w) % 16 == 0,
                "pcc_linear_bf16: operands must be 16-byte aligned with row pitches that are multiples of 8 elements");
    const bool pool = group >= 1;  // group == 1: dense fp32 rows through the pooled kernel's plain-store epilogue
    if (group == 1) {
        PCC_REQUIRE(ld_out == N && reinterpret_cast<uintptr_t>(out) % 16 == 0, "pcc_linear_bf16: the fp32 output is dense [M, N], 16-byte aligned");
    } else if (pool) {
        PCC_REQUIRE(group % 32 == 0 && (128 % group == 0 || group % 128 == 0) && M % group == 0,
                    "pcc_linear_bf16: group=%d must be a multiple of 32 that divides or is a multiple of 128, and divide M", group);
        PCC_REQUIRE(group == 32 || relu, "pcc_linear_bf16: pooling over more than 32 rows needs the ReLU (atomicMax on non-negative values)");
        PCC_REQUIRE(ld_out == N, "pcc_linear_bf16: the pooled output is dense [M / group, N] fp32");
    } else {
        PCC_REQUIRE(n_store >= 8 && n_store <= N && n_store % 8 == 0, "pcc_linear_bf16: n_store=%d must be a multiple of 8 in [8, N]", n_store);
        PCC_REQUIRE(ld_out >= n_store && ld_out % 8 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0,
                    "pcc_linear_bf16: out must be 16-byte aligned bf16 with a row pitch that is a multiple of 8 elements");
        PCC_REQUIRE(!mask || (ld_mask >= n_store && ld_mask % 8 == 0 && reinterpret_cast<uintptr_t>(mask) % 16 == 0),
                    "pcc_linear_bf16: mask must be 16-byte aligned bf16 [M, >= n_store] with a row pitch that is a multiple of 8");
    }
    PCC_REQUIRE(!pool || !mask, "pcc_linear_bf16: the mask applies to the bf16 output only");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int BN = (N % 256 == 0) ? 256 : 128;
    CUtensorMap ta, tw, to;
    if (int r = make_tmap_bf16_2d(&ta, a, static_cast<uint64_t>(M), static_cast<uint64_t>(K), static_cast<uint64_t>(lda), 128)) return r;
    if (int r = make_tmap_bf16_2d(&tw, w, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(ldw), BN)) return r;
    if (pool) {
        to = ta;  // unused
        if (group > 32) {
            const cudaError_t e = cudaMemsetAsync(out, 0, static_cast<size_t>(M / group) * N * sizeof(float), st);
            if (e != cudaSuccess) {
                set_error("pcc_linear_bf16: cudaMemsetAsync failed: %s", cudaGetErrorString(e));
                return static_cast<int>(e);
            }
        }
    } else if (int r = make_tmap_bf16_2d(&to, out, static_cast<uint64_t>(M), static_cast<uint64_t>(n_store), static_cast<uint64_t>(ld_out), 128)) {
        return r;   // the store map ends at n_store: TMA clips the columns past it
    }
    gws::Params p{};
    p.bias = bias;
    p.out_pool = pool ? static_cast<float *>(out) : nullptr;
    p.M = M;
    p.N = N;
    p.K = K;
    p.n_n = N / BN;
    p.n_tiles = ((M + 127) / 128) * p.n_n;
    p.relu = relu;
    p.group = pool ? group : 0;
    p.mask = static_cast<const __nv_bfloat16 *>(mask);
    p.ld_mask = ld_mask;
    p.n_store = pool ? N : n_store;
    if (mask) return BN == 256 ? gws::launch<256, false, true>(p, ta, tw, to, st) : gws::launch<128, false, true>(p, ta, tw, to, st);
    if (BN == 256) return pool ? gws::launch<256, true>(p, ta, tw, to, st) : gws::launch<256, false>(p, ta, tw, to, st);
    return pool ? gws::launch<128, true>(p, ta, tw, to, st) : gws::launch<128, false>(p, ta, tw, to, st);
}

PCC_API int pcc_linear_bf16(const void *a, int64_t M, int K, int64_t lda, const void *w, int64_t ldw, const float *bias, int N, int relu,
                            int group, void *out, int64_t ld_out, void *stream) {
    return linear_run(a, M, K, lda, w, ldw, bias, N, relu, group, out, ld_out, nullptr, 0, N, stream);
}

PCC_API int pcc_linear_train_bf16(const void *a, int64_t M, int K, int64_t lda, const void *w, int64_t ldw, const float *bias, int N,
                                  int relu, void *out, int64_t ld_out, int n_store, const void *mask, int64_t ld_mask, void *stream) {
    return linear_run(a, M, K, lda, w, ldw, bias, N, relu, 0, out, ld_out, mask, ld_mask, n_store, stream);
}
