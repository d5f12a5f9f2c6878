// Shared device helpers and host-side error plumbing for libpcc_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pcc_b200.h"

#define PCC_API extern "C" __attribute__((visibility("default")))

namespace pcc {

// ---- host: error reporting ------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int check_launch(const char *what);  // cudaGetLastError -> 0 or positive cudaError_t (+ message)

#define PCC_REQUIRE(cond, ...)                \
    do {                                      \
        if (!(cond)) {                        \
            pcc::set_error(__VA_ARGS__);      \
            return PCC_ERR_INVALID_ARGUMENT;  \
        }                                     \
    } while (0)

int num_sms();  // SM count of the current device (cached)

// ---- device: exact (non-contracted) squared distance --------------------------------------------------------
// d2 = fl(fl(fl(dx*dx) + fl(dy*dy)) + fl(dz*dz)); the _rn intrinsics are never fused into FMA by nvcc.
// Matches torch CPU sum((a-b)**2,-1) (pn_kit.py:326) and PyTorch3D's `dist += diff*diff` CPU loops.
__device__ __forceinline__ float dist2_rn(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = __fsub_rn(ax, bx);
    const float dy = __fsub_rn(ay, by);
    const float dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// (d2, idx) packed so that unsigned 64-bit order == lexicographic (d2, idx) order; valid because d2 >= +0.
__device__ __forceinline__ unsigned long long pack_key(float d2, unsigned idx) {
    return (static_cast<unsigned long long>(__float_as_uint(d2)) << 32) | idx;
}
__device__ __forceinline__ float key_d2(unsigned long long k) { return __uint_as_float(static_cast<unsigned>(k >> 32)); }
__device__ __forceinline__ unsigned key_idx(unsigned long long k) { return static_cast<unsigned>(k & 0xffffffffu); }

constexpr unsigned long long KEY_MAX = 0xffffffffffffffffull;
constexpr unsigned FULL_MASK = 0xffffffffu;

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// packed fp32 (sm_100): two independent IEEE operations per instruction, operands are {lo, hi} register pairs.  They issue at
// half rate -- no extra FP32 throughput, half the issue slots.  NOTE: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into
// FFMA2 whatever --fmad says; code that needs un-fused sums keeps them scalar (chamfer_grid.cu).
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t dup_f32x2(float x) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ unsigned long long add_f32x2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long mul_f32x2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long dup_neg_f32x2(float x) {
    unsigned long long r;
    const float n = -x;
    asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(n));
    return r;
}
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack_f32x2(uint64_t v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

}  // namespace pcc
