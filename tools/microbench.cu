// Micro-benchmarks that size the FP64 roofline on the box (not part of the product path).
//   DMMA m8n8k4 issue rate, DFMA issue rate, fp64 exp() rate, HBM write bandwidth.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double* out, int iters) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_exp(double* out, int iters) {
    double x = -1e-3 * threadIdx.x, s = 0;
    for (int it = 0; it < iters; ++it) { s += exp(x); x -= 1e-4; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_write(double2* p, size_t n2) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += stride) p[i] = make_double2(1.0, 2.0);
}

template <typename F>
float timeit(F f, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 32 * 1024));
    int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int bps = 1; bps <= 2; ++bps) {
            int threads = warps * 32;
            if (threads > 1024) continue;
            float ms = timeit([&] { k_dmma<16><<<sms * bps, threads>>>(out, iters); });
            double fl = 2.0 * 256 * 16 * (double)iters * warps * sms * bps;
            printf("{\"bench\": \"dmma884\", \"warps_per_cta\": %d, \"cta_per_sm\": %d, \"tflops\": %.3f}\n", warps, bps, fl / ms * 1e-9);
            ms = timeit([&] { k_dfma<16><<<sms * bps, threads>>>(out, iters); });
            fl = 2.0 * 32 * 16 * (double)iters * warps * sms * bps;
            printf("{\"bench\": \"dfma\", \"warps_per_cta\": %d, \"cta_per_sm\": %d, \"tflops\": %.3f}\n", warps, bps, fl / ms * 1e-9);
        }
    }
    {
        float ms = timeit([&] { k_dmma<4><<<sms, 256>>>(out, iters); });
        double fl = 2.0 * 256 * 4 * (double)iters * 8 * sms;
        printf("{\"bench\": \"dmma884_ilp4_8warps\", \"tflops\": %.3f}\n", fl / ms * 1e-9);
        ms = timeit([&] { k_dmma<1><<<sms, 32>>>(out, iters); });
        printf("{\"bench\": \"dmma884_latency_ns\", \"value\": %.3f}\n", ms * 1e6 / iters);
        ms = timeit([&] { k_dfma<1><<<sms, 32>>>(out, iters); });
        printf("{\"bench\": \"dfma_latency_ns\", \"value\": %.3f}\n", ms * 1e6 / iters);
    }
    {
        int it2 = 2000;
        float ms = timeit([&] { k_exp<<<sms * 8, 256>>>(out, it2); });
        printf("{\"bench\": \"exp_f64\", \"gexp_per_s\": %.3f}\n", (double)it2 * sms * 8 * 256 / ms * 1e-6);
    }
    {
        size_t bytes = (size_t)8 << 30; double2* buf; CK(cudaMalloc(&buf, bytes));
        float ms = timeit([&] { k_write<<<sms * 16, 512>>>(buf, bytes / 16); });
        printf("{\"bench\": \"hbm_write\", \"gbs\": %.1f}\n", bytes / ms * 1e-6);
        CK(cudaFree(buf));
    }
    return 0;
}
