// General (non-symmetric) path: in-place Gauss-Jordan inverse with partial (row) pivoting, one CTA per
// problem, matrix resident in L2.  Stands in for numpy.linalg.inv (LU) at KMP.py:157 when the stored
// covariance blocks are not symmetric (demos.py:251,285 -> SURVEY.md H2).  Only reference-sized systems
// take this path (n <= KMPX_SMALL_MAX), so it is written for clarity, not for the roofline:
// n steps, each a rank-1 update of the whole matrix (2 n^3 flops, 16 n^3 bytes through L2).
#include "common.cuh"

namespace kmpx {

constexpr int GJ_THREADS = 1024;

struct GjArgs { double* A; int lda; long long strideA; const int* n_sys; int n_max; int batch; int* info; int* piv; };

__global__ void __launch_bounds__(GJ_THREADS) k_inverse_gj(GjArgs p) {
    extern __shared__ __align__(16) double smem[];
    double* rowk = smem;                    // pivot row, scaled
    double* colk = smem + p.n_max;          // column k before elimination
    __shared__ double red_v[GJ_THREADS / 32];
    __shared__ int red_i[GJ_THREADS / 32];
    __shared__ int sh_piv;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int b = blockIdx.x; b < p.batch; b += gridDim.x) {
        const int n = p.n_sys ? p.n_sys[b] : p.n_max;
        double* A = p.A + (size_t)b * p.strideA;
        int* piv = p.piv + (size_t)b * p.n_max;
        int bad = 0;
        for (int k = 0; k < n; ++k) {
            // pivot search: first row of maximal |A[r][k]|, r >= k
            double bv = -1.0; int bi = n;
            for (int r = k + tid; r < n; r += GJ_THREADS) {
                const double v = fabs(A[(size_t)r * p.lda + k]);
                if (v > bv || (v == bv && r < bi)) { bv = v; bi = r; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
            __syncthreads();
            if (tid == 0) {
                double v = red_v[0]; int i = red_i[0];
                for (int t = 1; t < GJ_THREADS / 32; ++t)
                    if (red_v[t] > v || (red_v[t] == v && red_i[t] < i)) { v = red_v[t]; i = red_i[t]; }
                if (!(v > 0.0)) { if (bad == 0) bad = k + 1; i = k; }
                sh_piv = i; piv[k] = i;
            }
            __syncthreads();
            const int pr = sh_piv;
            if (tid == 0 && !(fabs(A[(size_t)pr * p.lda + k]) > 0.0) && bad == 0) bad = k + 1;
            // swap rows k and pr; stage the scaled pivot row and column k
            const double d = 1.0 / A[(size_t)pr * p.lda + k];
            __syncthreads();
            for (int c = tid; c < n; c += GJ_THREADS) {
                const double vk = A[(size_t)k * p.lda + c], vp = A[(size_t)pr * p.lda + c];
                if (pr != k) A[(size_t)pr * p.lda + c] = vk;
                rowk[c] = (c == k ? 1.0 : vp) * d;
            }
            __syncthreads();
            for (int r = tid; r < n; r += GJ_THREADS) colk[r] = (r == k) ? 0.0 : A[(size_t)r * p.lda + k];
            __syncthreads();
            for (size_t idx = tid; idx < (size_t)n * n; idx += GJ_THREADS) {
                const int r = (int)(idx / n), c = (int)(idx - (size_t)r * n);
                double* a = A + (size_t)r * p.lda + c;
                if (r == k) *a = rowk[c];
                else { const double base = (c == k) ? 0.0 : *a; *a = base - colk[r] * rowk[c]; }
            }
            __syncthreads();
        }
        // undo the row interchanges as column interchanges, last first
        for (int k = n - 1; k >= 0; --k) {
            const int pr = piv[k];
            if (pr != k) {
                for (int r = tid; r < n; r += GJ_THREADS) {
                    double* row = A + (size_t)r * p.lda;
                    const double t = row[k]; row[k] = row[pr]; row[pr] = t;
                }
            }
            __syncthreads();
        }
        if (tid == 0 && p.info) p.info[b] = bad;
        __syncthreads();
    }
}

}  // namespace kmpx

using namespace kmpx;

extern "C" long long kmpx_inverse_workspace_bytes(int n_max, int batch) {
    return (long long)n_max * batch * (long long)sizeof(int);
}

extern "C" int kmpx_inverse(double* A, int lda, long long strideA, const int* n_sys, int n_max, int batch,
                            int* info, void* ws, long long ws_bytes, void* stream) {
    if (!A) return fail_arg(1, "A is NULL");
    if (n_max <= 0 || n_max > 4096) return fail_arg(5, "n_max must be in 1..4096 for the general path");
    if (lda < n_max) return fail_arg(2, "lda < n_max");
    if (batch <= 0) return fail_arg(6, "batch must be > 0");
    if (!ws || ws_bytes < kmpx_inverse_workspace_bytes(n_max, batch)) return fail_arg(8, "workspace too small");
    GjArgs p{A, lda, strideA, n_sys, n_max, batch, info, (int*)ws};
    const size_t smem = (size_t)2 * n_max * sizeof(double);
    const int grid = batch < num_sms() ? batch : num_sms();
    k_inverse_gj<<<grid, GJ_THREADS, smem, (cudaStream_t)stream>>>(p);
    return check_launch("k_inverse_gj");
}
