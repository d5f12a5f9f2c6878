// Entropy stage of the codec on the device: what compress.py:131-140 / decompress.py:88-93 do on the host
// between the probability model and the .p.bin file.
//
//   pmf_to_cdf_kernel     pn_kit.pmf_to_cdf (/root/reference/pn_kit.py:452-461: cat(0, cumsum(pmf)), clamp(max = 1)) fused with
//                         torchac's _convert_to_int_and_normalize(needs_normalization = True): round(cdf * (2^16 - L)) as a
//                         16-bit pattern, + arange(L + 1).  The running sum is kept in double and rounded to float at every
//                         step, which is what torch's CPU cumsum does for float inputs (bit-exact to pn_kit.pmf_to_cdf run on
//                         CPU tensors: tests/golden/ref_entropy.npz).  compress.py:134 calls pmf_to_cdf on the DEVICE tensor,
//                         i.e. torch's CUDA cumsum, whose summation order is unspecified; together with torchac being
//                         unpinned this means .p.bin streams are self-consistent (encode <-> decode of this library, any
//                         batch size), not guaranteed interchangeable with streams the reference wrote on a GPU.
//   cdf_to_u16_kernel     the same conversion for a float CDF that the caller already holds (torchac.encode_float_cdf's input)
//   range_encode_kernel   torchac's arithmetic coder (32-bit low / high, 16-bit CDFs, pending-bit carry handling, MSB-first
//   range_decode_kernel   bits), one stream per cloud.  The recurrence is serial per stream: one thread per stream, streams in
//                         different warps so they do not serialise each other; a cloud's 1024 symbols take ~0.2 ms of latency
//                         on one lane, which the driver hides behind the next batch's kernels on a side stream.
// torchac (0.9.3) is not in /root/reference nor in this image: its published algorithm is restated (parity unpinned), and the
// tests anchor it with round trips and against the CPU oracle.
#include "pcc_common.cuh"

namespace pcc {

__global__ void __launch_bounds__(256)
pmf_to_cdf_kernel(const float *__restrict__ pmf, long long rows, int L, uint16_t *__restrict__ cdf) {
    const long long r = blockIdx.x * 256ll + threadIdx.x;
    if (r >= rows) return;
    const int Lp = L + 1;
    const float new_max = static_cast<float>(65536 - L);
    double run = 0.0;
    cdf[r * Lp] = 0;
    for (int k = 1; k < Lp; ++k) {
        run += static_cast<double>(pmf[r * L + k - 1]);
        float c = static_cast<float>(run);
        c = c > 1.0f ? 1.0f : c;
        cdf[r * Lp + k] = static_cast<uint16_t>(static_cast<int>(rintf(__fmul_rn(c, new_max))) + k);
    }
}

__global__ void __launch_bounds__(256)
cdf_to_u16_kernel(const float *__restrict__ cdf_f, long long total, int Lp, uint16_t *__restrict__ cdf) {
    const long long e = blockIdx.x * 256ll + threadIdx.x;
    if (e >= total) return;
    const int k = static_cast<int>(e % Lp);
    const float new_max = static_cast<float>(65536 - (Lp - 1));
    cdf[e] = static_cast<uint16_t>(static_cast<int>(rintf(__fmul_rn(cdf_f[e], new_max))) + k);
}

struct BitOut {
    uint8_t *out;
    int cap, n;
    unsigned cache;
    int count;
    __device__ __forceinline__ void append(unsigned bit) {
        cache = (cache << 1) | (bit & 1u);
        if (++count == 8) {
            if (n < cap) out[n] = static_cast<uint8_t>(cache);
            ++n;
            count = 0;
            cache = 0;
        }
    }
    __device__ __forceinline__ void append_and_pending(unsigned bit, unsigned long long &pending) {
        append(bit);
        while (pending > 0) {
            append(bit ^ 1u);
            --pending;
        }
    }
};

// one warp per stream, lane 0 runs the recurrence
__global__ void __launch_bounds__(32)
range_encode_kernel(const uint16_t *__restrict__ cdf, const int16_t *__restrict__ sym, int n_sym, int Lp, uint8_t *__restrict__ out,
                    int cap, int *__restrict__ nbytes) {
    if (threadIdx.x != 0) return;   // (staging the CDFs in shared memory first was measured: no gain, the recurrence is the latency)
    const int b = blockIdx.x;
    const uint16_t *c = cdf + static_cast<size_t>(b) * n_sym * Lp;
    const int16_t *s = sym + static_cast<size_t>(b) * n_sym;
    BitOut bo{out + static_cast<size_t>(b) * cap, cap, 0, 0u, 0};
    unsigned low = 0u, high = 0xFFFFFFFFu;
    unsigned long long pending = 0;
    const int max_symbol = Lp - 2;
    for (int i = 0; i < n_sym; ++i) {
        int v = s[i];
        v = v < 0 ? 0 : (v > max_symbol ? max_symbol : v);   // out-of-range symbols are the caller's error; never index past the row
        const unsigned long long span = static_cast<unsigned long long>(high) - static_cast<unsigned long long>(low) + 1ull;
        const unsigned c_low = c[static_cast<size_t>(i) * Lp + v];
        const unsigned c_high = v == max_symbol ? 0x10000u : c[static_cast<size_t>(i) * Lp + v + 1];
        high = (low - 1u) + static_cast<unsigned>((span * c_high) >> 16);
        low = low + static_cast<unsigned>((span * c_low) >> 16);
        for (;;) {
            if (high < 0x80000000u) {
                bo.append_and_pending(0u, pending);
                low <<= 1;
                high = (high << 1) | 1u;
            } else if (low >= 0x80000000u) {
                bo.append_and_pending(1u, pending);
                low <<= 1;
                high = (high << 1) | 1u;
            } else if (low >= 0x40000000u && high < 0xC0000000u) {
                ++pending;
                low = (low << 1) & 0x7FFFFFFFu;
                high = (high << 1) | 0x80000001u;
            } else {
                break;
            }
        }
    }
    pending += 1;
    bo.append_and_pending(low < 0x40000000u ? 0u : 1u, pending);
    while (bo.count != 0) bo.append(0u);
    nbytes[b] = bo.n;
}

struct BitIn {
    const uint8_t *in;
    int n, pos;
    unsigned cache;
    int cached;
    __device__ __forceinline__ void get(unsigned &value) {
        if (cached == 0) {
            if (pos == n) {
                value <<= 1;
                return;
            }
            cache = in[pos++];
            cached = 8;
        }
        value = (value << 1) | ((cache >> (cached - 1)) & 1u);
        --cached;
    }
};

__global__ void __launch_bounds__(32)
range_decode_kernel(const uint16_t *__restrict__ cdf, const uint8_t *__restrict__ in, const int *__restrict__ nbytes, int cap, int n_sym,
                    int Lp, int16_t *__restrict__ sym) {
    if (threadIdx.x != 0) return;
    const int b = blockIdx.x;
    const uint16_t *c = cdf + static_cast<size_t>(b) * n_sym * Lp;
    int16_t *s = sym + static_cast<size_t>(b) * n_sym;
    const int nb = nbytes[b] < cap ? nbytes[b] : cap;
    BitIn bi{in + static_cast<size_t>(b) * cap, nb, 0, 0u, 0};
    unsigned low = 0u, high = 0xFFFFFFFFu, value = 0u;
    const int max_symbol = Lp - 2;
    for (int i = 0; i < 32; ++i) bi.get(value);
    for (int i = 0; i < n_sym; ++i) {
        const unsigned long long span = static_cast<unsigned long long>(high) - static_cast<unsigned long long>(low) + 1ull;
        const unsigned count =
            static_cast<unsigned>(((((static_cast<unsigned long long>(value) - static_cast<unsigned long long>(low) + 1ull) << 16) - 1ull) / span)) & 0xffffu;
        int left = 0, right = max_symbol + 1;
        while (left + 1 < right) {
            const int m = (left + right) / 2;
            const unsigned v = c[static_cast<size_t>(i) * Lp + m];
            if (v < count) left = m;
            else if (v > count) right = m;
            else {
                left = m;
                break;
            }
        }
        s[i] = static_cast<int16_t>(left);
        const unsigned c_low = c[static_cast<size_t>(i) * Lp + left];
        const unsigned c_high = left == max_symbol ? 0x10000u : c[static_cast<size_t>(i) * Lp + left + 1];
        high = (low - 1u) + static_cast<unsigned>((span * c_high) >> 16);
        low = low + static_cast<unsigned>((span * c_low) >> 16);
        for (;;) {
            if (low >= 0x80000000u || high < 0x80000000u) {
                low <<= 1;
                high = (high << 1) | 1u;
                bi.get(value);
            } else if (low >= 0x40000000u && high < 0xC0000000u) {
                low = (low << 1) & 0x7FFFFFFFu;
                high = (high << 1) | 0x80000001u;
                value -= 0x40000000u;
                bi.get(value);
            } else {
                break;
            }
        }
    }
}

}  // namespace pcc

PCC_API int pcc_pmf_to_cdf_u16(const float *pmf, int64_t rows, int L, uint16_t *out_cdf, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(pmf && out_cdf, "pcc_pmf_to_cdf_u16: null pointer");
    PCC_REQUIRE(rows >= 0 && L >= 1 && L <= 4096, "pcc_pmf_to_cdf_u16: bad shape rows=%lld L=%d", static_cast<long long>(rows), L);
    if (rows == 0) return 0;
    pmf_to_cdf_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(pmf, rows, L, out_cdf);
    return check_launch("pmf_to_cdf_kernel");
}

PCC_API int pcc_cdf_to_u16(const float *cdf_float, int64_t rows, int Lp, uint16_t *out_cdf, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(cdf_float && out_cdf, "pcc_cdf_to_u16: null pointer");
    PCC_REQUIRE(rows >= 0 && Lp >= 2 && Lp <= 4097, "pcc_cdf_to_u16: bad shape");
    if (rows == 0) return 0;
    const long long total = rows * Lp;
    cdf_to_u16_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(cdf_float, total, Lp, out_cdf);
    return check_launch("cdf_to_u16_kernel");
}

PCC_API int pcc_range_encode_u16(const uint16_t *cdf, const int16_t *sym, int B, int n_sym, int Lp, uint8_t *out_bytes, int cap,
                                 int32_t *out_nbytes, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(cdf && sym && out_bytes && out_nbytes, "pcc_range_encode_u16: null pointer");
    PCC_REQUIRE(B >= 1 && n_sym >= 0 && Lp >= 2, "pcc_range_encode_u16: bad shape");
    PCC_REQUIRE(cap >= 2 * n_sym + 8, "pcc_range_encode_u16: cap=%d must be at least 2 * n_sym + 8 = %d bytes", cap, 2 * n_sym + 8);
    range_encode_kernel<<<B, 32, 0, static_cast<cudaStream_t>(stream)>>>(cdf, sym, n_sym, Lp, out_bytes, cap, out_nbytes);
    return check_launch("range_encode_kernel");
}

PCC_API int pcc_range_decode_u16(const uint16_t *cdf, const uint8_t *bytes, const int32_t *nbytes, int B, int n_sym, int Lp, int cap,
                                 int16_t *out_sym, void *stream) {
    using namespace pcc;
    PCC_REQUIRE(cdf && bytes && nbytes && out_sym, "pcc_range_decode_u16: null pointer");
    PCC_REQUIRE(B >= 1 && n_sym >= 0 && Lp >= 2 && cap >= 1, "pcc_range_decode_u16: bad shape");
    range_decode_kernel<<<B, 32, 0, static_cast<cudaStream_t>(stream)>>>(cdf, bytes, nbytes, cap, n_sym, Lp, out_sym);
    return check_launch("range_decode_kernel");
}
