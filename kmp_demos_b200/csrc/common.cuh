// Shared device/host helpers for libkmpx (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <atomic>
#include "../../include/kmpx.h"

namespace kmpx {

// ---------------------------------------------------------------- host side: errors, launch counting
extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;
extern int g_host_dbg;

inline int fail_arg(int k, const char* what) {
    snprintf(g_err, sizeof(g_err), "kmpx: bad argument %d: %s", k, what);
    return -k;
}
inline int check_launch(const char* name) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "kmpx: launch of %s failed: %s", name, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}
inline int check_cuda(cudaError_t e, const char* what) {
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "kmpx: %s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}
// Opt-in to more than 48 KB of dynamic shared memory; `cache` holds, per device, the largest size requested so far for
// this kernel (the attribute is per device and only ever raised; the limit is 227 KB minus the kernel's static shared
// memory).  One cache object per kernel: entry points that launch the same kernel share it.
constexpr int kMaxDevices = 64;
struct SmemCache {
    size_t v[kMaxDevices];
    SmemCache() { for (int i = 0; i < kMaxDevices; ++i) v[i] = 48 * 1024; }
};
inline int current_device() { int d = 0; (void)cudaGetDevice(&d); return (d >= 0 && d < kMaxDevices) ? d : 0; }
template <typename Kern>
inline int ensure_dyn_smem(Kern kernel, size_t bytes, SmemCache& cache, const char* name) {
    size_t& cur = cache.v[current_device()];
    if (bytes <= cur) return 0;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        snprintf(g_err, sizeof(g_err), "kmpx: %s needs %zu bytes of shared memory: %s", name, bytes, cudaGetErrorString(e));
        return (int)e;
    }
    cur = bytes;
    return 0;
}
#define KMPX_TRY(expr) do { int _rc = (expr); if (_rc != 0) return _rc; } while (0)

constexpr int kNumSMsB200 = 148;   // upper bound used for workspace sizing (the only target is the B200)
// SM count of the current device (queried once per device); grids of the persistent kernels are sized from it
inline int num_sms() {
    static int cache[kMaxDevices] = {};
    const int dev = current_device();
    if (cache[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) { (void)cudaGetLastError(); v = kNumSMsB200; }
        cache[dev] = v;
    }
    return cache[dev];
}
__host__ __device__ inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
__host__ __device__ inline long long round_up(long long a, long long b) { return (a + b - 1) / b * b; }

// ---------------------------------------------------------------- device side
constexpr double kFdDt = 0.001;   // finite-difference step of the time-driven kernel, KMP.py:113

// FP64 tensor-core MMA: D(8x8) += A(8x4,row) * B(4x8,col).  Lowers to DMMA.8x8x4 on sm_100a.
//   a: A[row = lane/4][k = lane%4]     b: B[k = lane%4][col = lane/4]
//   c[0..1]: C[row = lane/4][col = (lane%4)*2 + {0,1}]
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// 16-byte async global->shared copy; src_bytes in {0,8,16}: the remainder is zero-filled.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src, int src_bytes) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" :: "r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

// RBF coefficient k = exp(-sigma_f * ||s1-s2||^2), norm first then squared (KMP.py:109-110).
template <typename P1, typename P2>
__device__ __forceinline__ double rbf(P1 s1, P2 s2, int I, double sigma_f, double d1, double d2) {
    double r;
    if (I == 1) {
        r = fabs((s1[0] + d1) - (s2[0] + d2));
    } else {
        double acc = 0.0;
        for (int t = 0; t < I; ++t) { double d = (s1[t] + d1) - (s2[t] + d2); acc += d * d; }
        r = sqrt(acc);
    }
    return exp(-sigma_f * (r * r));
}

// exp(x) for x <= 0 without a single branch, so that several evaluations interleave in one instruction stream (the
// library exp carries a rare-path branch that serialises them; a dependent FP64 operation costs 30-60 cycles here).
// Cody-Waite reduction x = k ln2 + r, |r| <= ln2/2, degree-13 Taylor polynomial (truncation 4e-18), scaling through
// the exponent field; results below the normal range (x < -708) flush to zero.  Error <= ~1 ulp.
__device__ __forceinline__ double exp_nonpos(double x) {
    const double shift = 6755399441055744.0;                   // 1.5 * 2^52: the integer part lands in the low word
    const double t = fma(x, 1.4426950408889634, shift);
    const int k = __double2loint(t);
    const double kd = t - shift;
    double r = fma(kd, -6.93147180369123816490e-01, x);        // ln2 split: the high part has 33 significant bits
    r = fma(kd, -1.90821492927058770002e-10, r);
    double q = 1.6059043836821613e-10;                         // 1/13!
    q = fma(q, r, 2.08767569878681e-09);                       // 1/12!
    q = fma(q, r, 2.505210838544172e-08);
    q = fma(q, r, 2.755731922398589e-07);
    q = fma(q, r, 2.7557319223985893e-06);
    q = fma(q, r, 2.48015873015873e-05);
    q = fma(q, r, 1.984126984126984e-04);
    q = fma(q, r, 1.388888888888889e-03);
    q = fma(q, r, 8.333333333333333e-03);
    q = fma(q, r, 4.1666666666666664e-02);
    q = fma(q, r, 1.6666666666666666e-01);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = fma(q, r, 1.0);
    const double res = __hiloint2double(__double2hiint(q) + (k << 20), __double2loint(q));
    return x < -708.0 ? 0.0 : (x != x ? x : res);            // NaN propagates (select, still branch-free)
}
// Table-driven variant for kernels that keep a 32-entry table of 2^(j/32) in shared memory (fill_exp2_table): with
// x = (32 k + j) ln2/32 + r, |r| <= ln2/64, a degree-6 polynomial reaches 3e-18 truncation error, so an exponential costs
// 11 dependent FP64 operations + one shared-memory load instead of 18.  exp(x) = 2^k * T[j] * p(r); error ~1 ulp (the
// table entries are correctly rounded).  Same domain rules as exp_nonpos: x < -708 -> 0, NaN propagates, no overflow check.
constexpr int kExpTabN = 32;
__device__ __forceinline__ void fill_exp2_table(double* tab, int tid, int nthreads) {
    for (int j = tid; j < kExpTabN; j += nthreads) tab[j] = exp2((double)j / kExpTabN);   // library exp2: correctly rounded here
}
__device__ __forceinline__ double exp_tab(double x, const double* tab) {
    const double shift = 6755399441055744.0;
    const double t = fma(x, 46.16624130844682903551, shift);                 // 32 / ln 2
    const int n = __double2loint(t);
    const double kd = t - shift;
    double r = fma(kd, -2.16608493865351192653e-02, x);                      // ln2/32 split like exp_nonpos (hi has 33 bits)
    r = fma(kd, -5.96317165397058656257e-12, r);
    double q = 1.388888888888889e-03;                                        // 1/6!
    q = fma(q, r, 8.333333333333333e-03);
    q = fma(q, r, 4.1666666666666664e-02);
    q = fma(q, r, 1.6666666666666666e-01);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    const double tj = tab[n & (kExpTabN - 1)];
    q = fma(q * r, tj, tj);                                                  // T (1 + r p(r))
    const double res = __hiloint2double(__double2hiint(q) + ((n >> 5) << 20), __double2loint(q));
    return x < -708.0 ? 0.0 : (x != x ? x : res);
}
// rbf() with the branch-free exponential (plain kernel, hot loops that evaluate many coefficients per thread)
template <typename P1, typename P2>
__device__ __forceinline__ double rbf_fast(P1 s1, P2 s2, int I, double sigma_f) {
    double r2;
    if (I == 1) {
        const double r = fabs(s1[0] - s2[0]);
        r2 = r * r;
    } else {
        double acc = 0.0;
        for (int t = 0; t < I; ++t) { const double d = s1[t] - s2[t]; acc += d * d; }
        const double r = sqrt(acc);                            // norm first, then squared, as the reference does
        r2 = r * r;
    }
    return exp_nonpos(-sigma_f * r2);
}

// rbf_fast with the table-driven exponential (tab: fill_exp2_table in shared memory)
template <typename P1, typename P2>
__device__ __forceinline__ double rbf_tab(P1 s1, P2 s2, int I, double sigma_f, const double* tab) {
    double r2;
    if (I == 1) { const double d = fabs(s1[0] - s2[0]); r2 = d * d; }
    else {
        double acc = 0.0;
        for (int t = 0; t < I; ++t) { const double d = s1[t] - s2[t]; acc += d * d; }
        const double r = sqrt(acc);                            // norm first, then squared, as the reference does
        r2 = r * r;
    }
    return exp_tab(-sigma_f * r2, tab);
}

// The coefficient block of one kernel evaluation: plain -> c[0]; time-driven -> c = {ktt, ktdt, kdtt, kdtdt}
// by forward differences with dt = 1e-3 added to every input coordinate (KMP.py:113-123).
template <typename P1, typename P2>
__device__ __forceinline__ void kernel_coeffs(P1 s1, P2 s2, int I, double sigma_f, bool time_driven, double (&c)[4]) {
    double ktt = rbf(s1, s2, I, sigma_f, 0.0, 0.0);
    c[0] = ktt;
    if (!time_driven) { c[1] = c[2] = c[3] = 0.0; return; }
    double a = rbf(s1, s2, I, sigma_f, 0.0, kFdDt);
    double b = rbf(s1, s2, I, sigma_f, kFdDt, 0.0);
    double d = rbf(s1, s2, I, sigma_f, kFdDt, kFdDt);
    c[1] = (a - ktt) / kFdDt;
    c[2] = (b - ktt) / kFdDt;
    c[3] = (d - a - b + ktt) / (kFdDt * kFdDt);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace kmpx
