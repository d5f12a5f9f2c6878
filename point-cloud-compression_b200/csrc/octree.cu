// Octree centre coding on the device: the host-side stage the reference runs between FPS and kNN patching
// (train.py:175-179, compress.py:98-101), which otherwise forces a D2H / H2D round trip and 100+ ms of numpy per cloud.
//
//   octree_encode_kernel   octree_np.encode + the depth search of pn_kit.encode_sampled_np
//                          (/root/reference/octree_np.py:10-45,114-133, pn_kit.py:380-401), bit-exact bit stream; optional
//                          by-products: the byte stream of pn_kit.binary_array_to_byte_array (pn_kit.py:463-467), the snapped
//                          centres of octree_np.getDecodeFromPc in input order, and the output of the reference's decoder.
//   octree_decode_kernel   mode 0: octree_np.decode exactly as written (octree_np.py:47-112: it reads the first 8 bits only,
//                          emits depth-1 octant centres and pads to 64 rows);  mode 1: the inverse of encode (leaf centres in
//                          the stream's own order), which the reference does not have.
//
// The reference walks the tree with a stack and tests every point against every visited cube.  Here one CTA owns one cloud:
// the 48-bit Morton codes of the depth-16 cells (x the slowest bit of each triple, as the reference's child order) are sorted
// once in DESCENDING order -- the order in which the reference's DFS pops children (7 first) -- and every coarser level is a
// prefix of those codes.  The level-l section of the stream is, for every distinct level-(l-1) prefix in sorted order, the
// occupancy of its children 7..0; node counts per level come from one histogram of "first level at which neighbours differ",
// so the depth search costs nothing.  Coordinates must lie in [0, 1) (octree_np.py:5-7); scale is 1 at every call site.
#include "pcc_common.cuh"

namespace pcc {

constexpr int OCT_MAXD = 16;
constexpr int OCT_THREADS = 256;
constexpr unsigned long long OCT_MASK48 = 0xffffffffffffull;

__device__ __forceinline__ unsigned long long spread3_16(unsigned v) {
    unsigned long long x = v & 0xffffu;
    x = (x | (x << 16)) & 0x0000ff0000ffull;
    x = (x | (x << 8)) & 0x00f00f00f00full;
    x = (x | (x << 4)) & 0x0c30c30c30c3ull;
    x = (x | (x << 2)) & 0x249249249249ull;
    return x;
}

// block-wide exclusive prefix of a 0/1 flag over the 256 threads (+ running carry); returns this thread's exclusive rank
__device__ __forceinline__ int block_flag_scan(bool flag, int *warp_tot, int &carry) {
    const unsigned m = __ballot_sync(FULL_MASK, flag);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int in_warp = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[w] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int i = 0; i < OCT_THREADS / 32; ++i) {
        const int t = warp_tot[i];
        before += (i < w) ? t : 0;
        total += t;
    }
    const int r = carry + before + in_warp;
    carry += total;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(OCT_THREADS)
octree_encode_kernel(const float *__restrict__ xyz, int S, int P, int n_points, double min_bpp, int fixed_depth,
                     uint8_t *__restrict__ bits, int max_bits, int *__restrict__ nbits_out, int *__restrict__ depth_out,
                     uint8_t *__restrict__ bytes, float *__restrict__ quant, float *__restrict__ rec_ref,
                     float *__restrict__ stream_xyz) {
    extern __shared__ unsigned long long oct_sm[];
    unsigned long long *keys = oct_sm;                                   // [P] inverted Morton codes, ascending
    unsigned short *rank_prev = reinterpret_cast<unsigned short *>(keys + P);  // [P] rank of the level-(l-1) node
    unsigned short *rank_cur = rank_prev + P;
    __shared__ int hist[OCT_MAXD + 2];
    __shared__ int n_level[OCT_MAXD + 1];
    __shared__ int warp_tot[OCT_THREADS / 32];
    __shared__ int s_depth, s_nbits, s_bad;

    const int b = blockIdx.x, tid = threadIdx.x;
    const float *pc = xyz + static_cast<size_t>(b) * S * 3;
    uint8_t *out = bits + static_cast<size_t>(b) * max_bits;
    if (tid < OCT_MAXD + 2) hist[tid] = 0;
    if (tid == 0) s_bad = 0;
    __syncthreads();

    // ---- depth-16 cell codes (floor(p * 2^16) is exact for float32 p in [0,1)) ----
    for (int i = tid; i < P; i += OCT_THREADS) {
        unsigned long long k = KEY_MAX;  // padding sorts last
        if (i < S) {
            const float x = pc[i * 3 + 0], y = pc[i * 3 + 1], z = pc[i * 3 + 2];
            if (!(x >= 0.0f && x < 1.0f && y >= 0.0f && y < 1.0f && z >= 0.0f && z < 1.0f)) s_bad = 1;
            const unsigned cx = static_cast<unsigned>(floorf(x * 65536.0f)), cy = static_cast<unsigned>(floorf(y * 65536.0f)),
                           cz = static_cast<unsigned>(floorf(z * 65536.0f));
            const unsigned long long m = (spread3_16(cx) << 2) | (spread3_16(cy) << 1) | spread3_16(cz);
            k = (~m) & OCT_MASK48;
        }
        keys[i] = k;
    }
    __syncthreads();
    if (s_bad) {  // outside the unit cube: unsupported by the reference coder as well (octree_np.py:5-7)
        if (tid == 0) {
            nbits_out[b] = 0;
            depth_out[b] = -1;
        }
        return;
    }
    // ---- bitonic sort, ascending in the inverted code = descending Morton order ----
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P; i += OCT_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], c = keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > c) == up) {
                        keys[i] = c;
                        keys[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    // ---- first level at which sorted neighbours differ -> distinct nodes per level ----
    for (int i = tid + 1; i < S; i += OCT_THREADS) {
        const unsigned long long d = keys[i] ^ keys[i - 1];
        const int lvl = d ? OCT_MAXD - (63 - __clzll(static_cast<long long>(d))) / 3 : OCT_MAXD + 1;
        atomicAdd(&hist[lvl], 1);
    }
    __syncthreads();
    if (tid == 0) {
        n_level[0] = 1;
        for (int l = 1; l <= OCT_MAXD; ++l) n_level[l] = n_level[l - 1] + hist[l];
        int depth = fixed_depth, cum = 0;
        if (fixed_depth <= 0) {  // pn_kit.encode_sampled_np: first depth with bpp > min_bpp and no two centres in one cell
            depth = OCT_MAXD;
            for (int D = 1; D <= OCT_MAXD; ++D) {
                cum += n_level[D - 1];
                const int nb = 1 + 8 * cum;
                if (static_cast<double>(nb) / static_cast<double>(n_points) > min_bpp && n_level[D] == S) {
                    depth = D;
                    break;
                }
            }
        }
        cum = 0;
        for (int l = 0; l < depth; ++l) cum += n_level[l];
        s_depth = depth;
        s_nbits = 1 + 8 * cum;
        nbits_out[b] = s_nbits;
        depth_out[b] = depth;
    }
    __syncthreads();
    const int depth = s_depth, nbits = s_nbits;
    for (int i = tid; i < max_bits; i += OCT_THREADS) out[i] = 0;
    for (int i = tid; i < P; i += OCT_THREADS) rank_prev[i] = 0;
    __syncthreads();
    if (tid == 0) out[0] = 1;  // the root cube holds every point
    // ---- one section per level: child occupancy (7..0) of every distinct parent, parents in sorted order ----
    int base = 1;
    for (int l = 1; l <= depth; ++l) {
        const int sh = 3 * (OCT_MAXD - l);
        int carry = 0;
        for (int i0 = 0; i0 < S; i0 += OCT_THREADS) {
            const int i = i0 + tid;
            bool fresh = false;
            unsigned long long pre = 0;
            if (i < S) {
                pre = keys[i] >> sh;
                fresh = (i == 0) || (pre != (keys[i - 1] >> sh));
            }
            const int r = block_flag_scan(fresh, warp_tot, carry);  // exclusive count of fresh nodes before i
            if (i < S) {
                rank_cur[i] = static_cast<unsigned short>(fresh ? r : r - 1);
                if (fresh) out[base + 8 * rank_prev[i] + static_cast<int>(pre & 7ull)] = 1;  // inverted code: 7 - child
            }
        }
        __syncthreads();
        base += 8 * n_level[l - 1];
        unsigned short *t = rank_prev;
        rank_prev = rank_cur;
        rank_cur = t;
    }
    __syncthreads();
    // ---- by-products ----
    if (bytes) {  // pn_kit.binary_array_to_byte_array: MSB first; a short last chunk is read as a short binary number
        const int nbytes = (nbits + 7) / 8, cap = (max_bits + 7) / 8;
        uint8_t *ob = bytes + static_cast<size_t>(b) * cap;
        for (int j = tid; j < cap; j += OCT_THREADS) {
            unsigned v = 0;
            if (j < nbytes)
                for (int t = 8 * j; t < 8 * j + 8 && t < nbits; ++t) v = (v << 1) | out[t];
            ob[j] = static_cast<uint8_t>(v);
        }
    }
    if (quant) {  // octree_np.getDecodeFromPc at the chosen depth, input order kept
        const float inv = static_cast<float>(1 << depth), cube = 1.0f / inv, half = 0.5f * cube;
        for (int i = tid; i < S * 3; i += OCT_THREADS)
            quant[static_cast<size_t>(b) * S * 3 + i] = __fadd_rn(__fmul_rn(floorf(pc[i] * inv), cube), half);
    }
    if (stream_xyz) {  // what the inverse of this coder returns: the distinct leaf centres in stream order, last one repeated
        const int sh = 3 * (OCT_MAXD - depth), n_leaf = n_level[depth];
        const float cube = 1.0f / static_cast<float>(1 << depth), half = 0.5f * cube;
        float *o = stream_xyz + static_cast<size_t>(b) * S * 3;
        for (int i = tid; i < S; i += OCT_THREADS) {
            const unsigned long long pre = keys[i] >> sh;
            const bool fresh = (i == 0) || (pre != (keys[i - 1] >> sh));
            const bool last = (i == S - 1);
            if (!fresh && !last) continue;
            const unsigned long long m = (~(pre << sh)) & OCT_MASK48;  // true Morton code, low levels forced to 1s
            unsigned cx = 0, cy = 0, cz = 0;
            for (int l = 0; l < depth; ++l) {
                const unsigned c = static_cast<unsigned>(m >> (3 * (OCT_MAXD - 1 - l))) & 7u;
                cx = (cx << 1) | ((c >> 2) & 1u);
                cy = (cy << 1) | ((c >> 1) & 1u);
                cz = (cz << 1) | (c & 1u);
            }
            const float x = __fadd_rn(__fmul_rn(static_cast<float>(cx), cube), half),
                        y = __fadd_rn(__fmul_rn(static_cast<float>(cy), cube), half),
                        z = __fadd_rn(__fmul_rn(static_cast<float>(cz), cube), half);
            if (fresh) {
                const int r = rank_prev[i];
                o[r * 3 + 0] = x, o[r * 3 + 1] = y, o[r * 3 + 2] = z;
            }
            if (last)
                for (int r = n_leaf; r < S; ++r) o[r * 3 + 0] = x, o[r * 3 + 1] = y, o[r * 3 + 2] = z;
        }
    }
    if (rec_ref && tid < 64) {  // octree_np.decode as written: the first 8 bits select depth-1 octant centres, padded to 64
        const int g = nbits < 8 ? nbits : 8;
        int n = 0, mine = -1, last = -1;
        for (int j = 0; j < g; ++j)
            if (out[j] == 1) {
                if (n == tid) mine = 7 - j;
                last = 7 - j;
                ++n;
            }
        const int c = tid < n ? mine : last;
        float *o = rec_ref + (static_cast<size_t>(b) * 64 + tid) * 3;
        o[0] = c < 0 ? 0.0f : (((c >> 2) & 1) ? 0.75f : 0.25f);
        o[1] = c < 0 ? 0.0f : (((c >> 1) & 1) ? 0.75f : 0.25f);
        o[2] = c < 0 ? 0.0f : ((c & 1) ? 0.75f : 0.25f);
    }
}

// mode 0: the reference's decoder as written; mode 1: the true inverse of octree_encode_kernel.
__global__ void __launch_bounds__(OCT_THREADS)
octree_decode_kernel(const uint8_t *__restrict__ bits, const int *__restrict__ nbits_in, int max_bits, int mode, int cap,
                     float *__restrict__ out, int *__restrict__ count_out, int *__restrict__ depth_out) {
    extern __shared__ unsigned long long oct_sm[];
    unsigned long long *cur = oct_sm, *nxt = oct_sm + cap;  // node paths (3 bits per level), level by level
    __shared__ int warp_tot[OCT_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const uint8_t *in = bits + static_cast<size_t>(b) * max_bits;
    const int nbits = nbits_in[b] < max_bits ? (nbits_in[b] > 0 ? nbits_in[b] : 0) : max_bits;   // never read past the row
    float *o = out + static_cast<size_t>(b) * cap * 3;
    if (mode == 0) {
        if (tid < cap) {
            const int g = nbits < 8 ? nbits : 8;
            int n = 0, mine = -1, last = -1;
            for (int j = 0; j < g; ++j)
                if (in[j] == 1) {
                    if (n == tid) mine = 7 - j;
                    last = 7 - j;
                    ++n;
                }
            const int c = tid < n ? mine : last;
            o[tid * 3 + 0] = c < 0 ? 0.0f : (((c >> 2) & 1) ? 0.75f : 0.25f);
            o[tid * 3 + 1] = c < 0 ? 0.0f : (((c >> 1) & 1) ? 0.75f : 0.25f);
            o[tid * 3 + 2] = c < 0 ? 0.0f : ((c & 1) ? 0.75f : 0.25f);
            if (tid == 0) {
                if (count_out) count_out[b] = n;
                if (depth_out) depth_out[b] = 1;
            }
        }
        return;
    }
    int n_nodes = (nbits >= 1 && in[0] == 1) ? 1 : 0, pos = 1, depth = 0;
    if (tid == 0) cur[0] = 0;
    __syncthreads();
    while (n_nodes > 0 && pos + 8 * n_nodes <= nbits && depth < OCT_MAXD) {
        int carry = 0;
        const int slots = 8 * n_nodes;
        for (int s0 = 0; s0 < slots; s0 += OCT_THREADS) {
            const int s = s0 + tid;
            const bool occ = s < slots && in[pos + s] == 1;
            const int r = block_flag_scan(occ, warp_tot, carry);
            if (occ && r < cap) nxt[r] = (cur[s >> 3] << 3) | static_cast<unsigned long long>(7 - (s & 7));
        }
        __syncthreads();
        pos += slots;
        n_nodes = carry < cap ? carry : cap;
        ++depth;
        unsigned long long *t = cur;
        cur = nxt;
        nxt = t;
    }
    const float cube = 1.0f / static_cast<float>(1 << depth), half = 0.5f * cube;
    for (int i = tid; i < cap; i += OCT_THREADS) {
        float x = 0.0f, y = 0.0f, z = 0.0f;
        if (n_nodes > 0) {  // rows past the last leaf repeat it (the reference's padding rule, octree_np.py:101-105)
            const unsigned long long path = cur[i < n_nodes ? i : n_nodes - 1];
            unsigned cx = 0, cy = 0, cz = 0;
            for (int l = 0; l < depth; ++l) {
                const unsigned c = static_cast<unsigned>(path >> (3 * (depth - 1 - l))) & 7u;
                cx = (cx << 1) | ((c >> 2) & 1u);
                cy = (cy << 1) | ((c >> 1) & 1u);
                cz = (cz << 1) | (c & 1u);
            }
            x = __fadd_rn(__fmul_rn(static_cast<float>(cx), cube), half);
            y = __fadd_rn(__fmul_rn(static_cast<float>(cy), cube), half);
            z = __fadd_rn(__fmul_rn(static_cast<float>(cz), cube), half);
        }
        o[i * 3 + 0] = x;
        o[i * 3 + 1] = y;
        o[i * 3 + 2] = z;
    }
    if (tid == 0) {
        if (count_out) count_out[b] = n_nodes;
        if (depth_out) depth_out[b] = depth;
    }
}

static int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace pcc

PCC_API int pcc_octree_max_bits(int S) { return 1 + 8 * pcc::OCT_MAXD * (S > 0 ? S : 1); }

PCC_API int pcc_octree_encode_f32(const float *centres, int B, int S, int n_points, double min_bpp, int fixed_depth,
                                  uint8_t *out_bits, int max_bits, int32_t *out_nbits, int32_t *out_depth, uint8_t *out_bytes,
                                  float *out_quant, float *out_rec_ref, float *out_stream_xyz, void *stream) {
    PCC_REQUIRE(centres && out_bits && out_nbits && out_depth, "pcc_octree_encode_f32: null pointer");
    PCC_REQUIRE(B >= 1 && S >= 1 && S <= 8192, "pcc_octree_encode_f32: need B >= 1 and 1 <= S <= 8192 (got B=%d S=%d)", B, S);
    PCC_REQUIRE(n_points >= 1, "pcc_octree_encode_f32: n_points must be >= 1");
    PCC_REQUIRE(fixed_depth >= 0 && fixed_depth <= pcc::OCT_MAXD, "pcc_octree_encode_f32: depth must be in [0, 16] (0 = search)");
    const int need = fixed_depth > 0 ? 1 + 8 * fixed_depth * S : pcc_octree_max_bits(S);
    PCC_REQUIRE(max_bits >= need, "pcc_octree_encode_f32: max_bits %d < %d", max_bits, need);
    const int P = pcc::next_pow2(S);
    const size_t smem = static_cast<size_t>(P) * (8 + 2 + 2);
    if (smem + 1024 > 48 * 1024)  // per device, cheap: only clouds with more than 4096 centres get here
        cudaFuncSetAttribute(pcc::octree_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 12);
    pcc::octree_encode_kernel<<<B, pcc::OCT_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
        centres, S, P, n_points, min_bpp, fixed_depth, out_bits, max_bits, out_nbits, out_depth, out_bytes, out_quant, out_rec_ref,
        out_stream_xyz);
    return pcc::check_launch("octree_encode_kernel");
}

PCC_API int pcc_octree_decode_f32(const uint8_t *bits, const int32_t *nbits, int B, int max_bits, int mode, int cap, float *out_xyz,
                                  int32_t *out_count, int32_t *out_depth, void *stream) {
    PCC_REQUIRE(bits && nbits && out_xyz, "pcc_octree_decode_f32: null pointer");
    PCC_REQUIRE(B >= 1 && max_bits >= 1, "pcc_octree_decode_f32: need B >= 1 and max_bits >= 1");
    PCC_REQUIRE(mode == 0 || mode == 1, "pcc_octree_decode_f32: mode must be 0 (reference decoder) or 1 (inverse of encode)");
    if (mode == 0) PCC_REQUIRE(cap == 64, "pcc_octree_decode_f32: the reference decoder always returns 64 rows (octree_np.py:100)");
    PCC_REQUIRE(cap >= 1 && cap <= 8192, "pcc_octree_decode_f32: cap must be in [1, 8192]");
    const size_t smem = mode == 1 ? static_cast<size_t>(cap) * 16 : 0;
    if (smem + 1024 > 48 * 1024)
        cudaFuncSetAttribute(pcc::octree_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16);
    pcc::octree_decode_kernel<<<B, pcc::OCT_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(bits, nbits, max_bits, mode, cap,
                                                                                              out_xyz, out_count, out_depth);
    return pcc::check_launch("octree_decode_kernel");
}
