// Error plumbing, version and device queries of libpcc_b200.so.
#include <stdarg.h>

#include "pcc_common.cuh"

namespace pcc {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static unsigned long long g_launches = 0;  // kernels enqueued through check_launch (bench.py's gpu_launches)

int check_launch(const char *what) {
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) {
        __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED);
        return 0;
    }
    set_error("%s: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace pcc

PCC_API int pcc_version(void) { return 100; /* 0.1.0 */ }

PCC_API const char *pcc_last_error_string(void) { return pcc::g_err; }

PCC_API int64_t pcc_launch_count(void) { return static_cast<int64_t>(__atomic_load_n(&pcc::g_launches, __ATOMIC_RELAXED)); }
