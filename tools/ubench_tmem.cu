// TMEM read-back micro-benchmark for B200 (sm_100a): bytes/clk/SM that tcgen05.ld sustains, as a function of the number
// of warps reading and of co-resident CTAs.  Sets the epilogue roof of the fused MLP chain kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_tmem tools/ubench_tmem.cu && ./ubench_tmem
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define REPS 256

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void ld_wait(uint32_t (&a)[32], uint32_t (&b)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[7]), "+r"(a[15]), "+r"(a[31]), "+r"(b[0]), "+r"(b[7]), "+r"(b[15]), "+r"(b[31])
                 :
                 : "memory");
}

// COLS = TMEM columns of this CTA; WORK = extra ALU instructions per value (0 = xor of 4 probes per load only)
template <int COLS, int WORK>
__global__ void __launch_bounds__(256) k(unsigned *out, long long *clk) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(&slot)),
                     "r"(COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    // warps 4..7 (if present) share the lane quadrant of warps 0..3 and start at a different column
    const int nw = blockDim.x >> 5;
    const int col0 = (warp >> 2) * (COLS / 2);
    uint32_t a[32], b[32];
    float acc = 0.f;
    unsigned x = 0;
    const long long t0 = clock64();
    for (int r = 0; r < REPS; ++r) {
#pragma unroll 1
        for (int c = 0; c < COLS; c += 64) {
            const int cc = (col0 + c) % COLS;
            ld32(base + cc, a);
            ld32(base + cc + 32, b);
            ld_wait(a, b);
            if (WORK) {
#pragma unroll
                for (int i = 0; i < 32; ++i) acc = fmaxf(acc, fmaxf(__uint_as_float(a[i]), __uint_as_float(b[i])));
            } else {
                x ^= a[0] ^ a[7] ^ a[15] ^ a[31] ^ b[0] ^ b[7] ^ b[15] ^ b[31];
            }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + __float_as_uint(acc);
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(COLS) : "memory");
    (void)nw;
}

template <int COLS, int WORK>
void run(const char *name, int threads, int ctas_per_sm, unsigned *out, long long *clk) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<COLS, WORK><<<grid, threads>>>(out, clk);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<COLS, WORK><<<grid, threads>>>(out, clk);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    long long h[8];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    // bytes read per CTA: every warp reads REPS * COLS columns x 32 lanes x 4 B
    const double bytes_cta = (double)(threads / 32) * REPS * COLS * 32 * 4;
    printf("%-44s thr=%3d ctas/sm=%d  %7.3f ms  cta clk=%8lld  %6.1f B/clk/CTA  %7.1f B/clk/SM (event time)\n", name, threads,
           ctas_per_sm, best, h[0], bytes_cta / (double)h[0], bytes_cta * ctas_per_sm / (best * 1e-3 * 1.965e9));
}

int main() {
    unsigned *out;
    long long *clk;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaMalloc(&clk, 148 * 8 * 8);
    run<512, 0>("ld 32x32b.x32, 512 cols, no work", 128, 1, out, clk);
    run<512, 0>("ld 32x32b.x32, 512 cols, no work", 256, 1, out, clk);
    run<128, 0>("ld 32x32b.x32, 128 cols, no work", 128, 1, out, clk);
    run<128, 0>("ld 32x32b.x32, 128 cols, no work", 128, 2, out, clk);
    run<128, 0>("ld 32x32b.x32, 128 cols, no work", 128, 4, out, clk);
    run<128, 0>("ld 32x32b.x32, 128 cols, no work", 256, 4, out, clk);
    run<512, 1>("ld + 1 FMNMX / value, 512 cols", 128, 1, out, clk);
    run<512, 1>("ld + 1 FMNMX / value, 512 cols", 256, 1, out, clk);
    run<128, 1>("ld + 1 FMNMX / value, 128 cols", 128, 4, out, clk);
    run<128, 1>("ld + 1 FMNMX / value, 128 cols", 256, 4, out, clk);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
