// FP32 CUDA-core pipe micro-benchmark for B200 (sm_100a): how many lane-ops/s do FFMA / FADD / FMUL / FMNMX and the
// packed f32x2 forms sustain?  Gives the measured "FP32 issue" roof the kNN / Chamfer kernels are compared against.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench tools/ubench.cu && ./ubench
#include <cuda_runtime.h>
#include <stdio.h>

#define ITERS 4096
#define CHAINS 8

template <int OP>
__global__ void __launch_bounds__(256) k(float *out, float a, float b) {
    float x[CHAINS];
    unsigned long long px[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        x[i] = a + threadIdx.x + i;
        px[i] = (unsigned long long)__float_as_uint(x[i]) << 32 | __float_as_uint(b + i);
    }
    unsigned long long pa = (unsigned long long)__float_as_uint(a) << 32 | __float_as_uint(a);
    unsigned long long pb = (unsigned long long)__float_as_uint(b) << 32 | __float_as_uint(b);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (OP == 0) x[i] = __fmaf_rn(x[i], a, b);
            if (OP == 1) x[i] = __fadd_rn(x[i], b);
            if (OP == 2) x[i] = __fmul_rn(x[i], a);
            if (OP == 3) x[i] = fminf(x[i], b + it);
            if (OP == 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(px[i]) : "l"(pa), "l"(pb));
            if (OP == 5) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(px[i]) : "l"(pb));
            if (OP == 6) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(px[i]) : "l"(pa));
            if (OP == 7) {  // the un-fused d2 mix: 3 sub, 3 mul, 2 add, 1 min  (9 instr)
                float dx = __fsub_rn(x[i], a), dy = __fsub_rn(x[i], b), dz = __fsub_rn(x[i], a + b);
                float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                x[i] = fminf(x[i], d);
            }
            if (OP == 8) {  // FMA form: 3 sub, 1 mul, 2 fma, 1 min (7 instr)
                float dx = __fsub_rn(x[i], a), dy = __fsub_rn(x[i], b), dz = __fsub_rn(x[i], a + b);
                float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                x[i] = fminf(x[i], d);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += x[i] + __uint_as_float((unsigned)(px[i] >> 32)) + __uint_as_float((unsigned)px[i]);
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int OP>
void run(const char *name, double lane_ops_per_iter, float *out) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    dim3 grid(sms * 8), block(256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<OP><<<grid, block>>>(out, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<OP><<<grid, block>>>(out, 1.0001f, 0.5f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double ops = (double)grid.x * 256 * ITERS * CHAINS * lane_ops_per_iter;
    printf("%-28s %8.3f ms  %8.2f T lane-instr/s  (%.1f instr/clk/SM at 1.965 GHz)\n", name, best, ops / best / 1e9,
           ops / best / 1e9 * 1e12 / 32 / sms / 1.965e9);
}

int main() {
    float *out;
    cudaMalloc(&out, 148 * 8 * 256 * 4 * 2);
    run<0>("FFMA", 1, out);
    run<1>("FADD", 1, out);
    run<2>("FMUL", 1, out);
    run<3>("FMNMX", 1, out);
    run<4>("FFMA2 (f32x2, counted as 1)", 1, out);
    run<5>("FADD2 (f32x2, counted as 1)", 1, out);
    run<6>("FMUL2 (f32x2, counted as 1)", 1, out);
    run<7>("d2 un-fused + min (9 instr)", 9, out);
    run<8>("d2 fma + min (7 instr)", 7, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
