// DMMA throughput with shared-memory fragment loads interleaved (mimics the real kernels).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench2 tools/microbench2.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
// warp tile MI x NI m8n8 tiles; per k4 step: MI + NI LDS.64, MI*NI DMMA
template <int MI, int NI, bool LOADS>
__global__ void k(double* out, int iters) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1e-3 * i;
    __syncthreads();
    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const double* A = sm + warp * 64 + g * 20 + t;
    const double* B = sm + 2048 + g * 20 + t;
    double a[MI], b[NI];
#pragma unroll
    for (int i = 0; i < MI; ++i) a[i] = A[i * 160];
#pragma unroll
    for (int j = 0; j < NI; ++j) b[j] = B[j * 160];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            if (LOADS) {
#pragma unroll
                for (int i = 0; i < MI; ++i) a[i] = A[i * 160 + kk * 4 + (it & 1) * 16];
#pragma unroll
                for (int j = 0; j < NI; ++j) b[j] = B[j * 160 + kk * 4 + (it & 1) * 16];
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma884(acc[i][j], a[i], b[j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) s += acc[i][j][0] + acc[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MI, int NI, bool LOADS>
void run(const char* name, double* out, int warps) {
    int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MI, NI, LOADS><<<148, warps * 32, 40000>>>(out, iters); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MI, NI, LOADS><<<148, warps * 32, 40000>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 256 * MI * NI * 4.0 * iters * warps * 148;
    printf("%-28s warps/SM=%2d  %.2f TFLOP/s\n", name, warps, fl / ms * 1e-9);
}
int main() {
    double* out; cudaMalloc(&out, 148 * 1024 * 8);
    for (int w : {4, 8, 16}) {
        run<2, 8, false>("2x8 no loads", out, w);
        run<2, 8, true>("2x8 10 LDS per 16 DMMA", out, w);
        run<4, 4, true>("4x4  8 LDS per 16 DMMA", out, w);
        run<8, 4, true>("8x4 12 LDS per 32 DMMA", out, w);
        run<2, 4, true>("2x4  6 LDS per  8 DMMA", out, w);
    }
    return 0;
}
