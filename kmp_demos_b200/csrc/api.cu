// Library-level state of libkmpx: version, thread-local error string, launch counter.
#include "common.cuh"

namespace kmpx {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
int g_host_dbg = 0;   // profiling experiments (kmpx_debug_set)
double g_predict_kskip = 3.3881317890172014e-21;      // 2^-68, kmpx_predict_skip_threshold
int g_predict_tile_queries = 0;                       // 0 = automatic, 32, 64: kmpx_predict_tile_queries
}  // namespace kmpx

extern "C" int kmpx_version(void) { return KMPX_VERSION; }
extern "C" const char* kmpx_last_error(void) { return kmpx::g_err; }
extern "C" long long kmpx_launch_count(void) { return kmpx::g_launches.load(); }

// Accuracy probe of the two branch-free exponentials every hot kernel evaluates (common.cuh): y[i] = exp(x[i]) by
// exp_nonpos (mode 0) or by the table-driven exp_tab (mode 1).  Test infrastructure only.
namespace kmpx {
__global__ void k_debug_exp(const double* x, double* y, long long n, int mode) {
    __shared__ double tab[kExpTabN];
    fill_exp2_table(tab, threadIdx.x, blockDim.x);
    __syncthreads();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = mode ? exp_tab(x[i], tab) : exp_nonpos(x[i]);
}
}  // namespace kmpx
extern "C" int kmpx_debug_exp(const double* x, double* y, long long n, int mode, void* stream) {
    if (!x || !y) return kmpx::fail_arg(1, "x or y is NULL");
    if (n <= 0) return kmpx::fail_arg(3, "n must be > 0");
    kmpx::k_debug_exp<<<(int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, (cudaStream_t)stream>>>(x, y, n, mode);
    return kmpx::check_launch("k_debug_exp");
}
