// K7: k-means kernels.
//  sklearn mode  - the initialisation of the GMM (sklearn/cluster/_kmeans.py:180-278 k-means++,
//                  :630-758 Lloyd, _k_means_lloyd.pyx:180-230 assignment).  RandomState stream, cumulative
//                  sum and searchsorted stay on the host; everything that touches the n samples is here.
//  reference mode - kmp.cluster.KMeans (Cluster.py:88-102), bit-exact: no FMA contraction, numpy's
//                  summation orders, numpy's NaN-aware argmin.
#include "common.cuh"

namespace kmpx {

// X[i][:] -= mean ; x_sq[i] = |X[i]|^2     (sklearn/cluster/_kmeans.py:1486-1496)
__global__ void k_rows_center_sqnorm(double* X, long long n, int d, const double* mean, double* x_sq) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double acc = 0.0;
    for (int t = 0; t < d; ++t) {
        const double v = X[i * d + t] - (mean ? mean[t] : 0.0);
        X[i * d + t] = v;
        acc += v * v;
    }
    if (x_sq) x_sq[i] = acc;
}

constexpr int KPP_MAXT = 8;
constexpr int KPP_THREADS = 256;

// dist[t][i] = min(closest[i], max(0, |c_t|^2 - 2 c_t.x_i + |x_i|^2));  partial potentials per CTA
__global__ void __launch_bounds__(KPP_THREADS) k_kpp_candidates(const double* X, const double* x_sq, long long n, int d,
                                                                const int* cand, int T, const double* closest,
                                                                double* dist, double* part) {
    __shared__ double c_sh[KPP_MAXT * 16];
    __shared__ double csq_sh[KPP_MAXT];
    __shared__ double red[KPP_MAXT][KPP_THREADS / 32];
    const int tid = threadIdx.x;
    for (int idx = tid; idx < T * d; idx += KPP_THREADS) { const int t = idx / d; c_sh[idx] = X[(long long)cand[t] * d + (idx - t * d)]; }
    __syncthreads();
    if (tid < T) { double a = 0.0; for (int q = 0; q < d; ++q) a += c_sh[tid * d + q] * c_sh[tid * d + q]; csq_sh[tid] = a; }
    __syncthreads();
    double pot[KPP_MAXT];
#pragma unroll
    for (int t = 0; t < KPP_MAXT; ++t) pot[t] = 0.0;
    for (long long i = blockIdx.x * (long long)KPP_THREADS + tid; i < n; i += (long long)gridDim.x * KPP_THREADS) {
        double x[16];
        for (int q = 0; q < d; ++q) x[q] = X[i * d + q];
        const double xs = x_sq[i];
        const double cl = closest ? closest[i] : INFINITY;
#pragma unroll
        for (int t = 0; t < KPP_MAXT; ++t) {
            if (t < T) {
                double dot = 0.0;
                for (int q = 0; q < d; ++q) dot += c_sh[t * d + q] * x[q];
                double v = -2.0 * dot + csq_sh[t] + xs;
                v = v > 0.0 ? v : 0.0;
                v = v < cl ? v : cl;
                dist[(long long)t * n + i] = v;
                pot[t] += v;
            }
        }
    }
#pragma unroll
    for (int t = 0; t < KPP_MAXT; ++t) {
        const double s = warp_sum(pot[t]);
        if ((tid & 31) == 0) red[t][tid >> 5] = s;
    }
    __syncthreads();
    if (tid < T) { double s = 0.0; for (int w = 0; w < KPP_THREADS / 32; ++w) s += red[tid][w]; part[(size_t)blockIdx.x * KPP_MAXT + tid] = s; }
}
__global__ void k_kpp_reduce(const double* part, int nctas, int T, double* pot) {
    const int t = threadIdx.x;
    if (t < T) { double s = 0.0; for (int c = 0; c < nctas; ++c) s += part[(size_t)c * KPP_MAXT + t]; pot[t] = s; }
}

// labels[i] = argmin_j(|c_j|^2 - 2 x_i.c_j), first minimum (_k_means_lloyd.pyx:215-224)
__global__ void __launch_bounds__(256) k_kmeans_assign(const double* X, long long n, int d, int K, const double* centers,
                                                       const int* labels_prev, int* labels, int* n_changed) {
    extern __shared__ double sh[];
    double* c_sh = sh;              // [K][d]
    double* csq = sh + (size_t)K * d;
    for (int idx = threadIdx.x; idx < K * d; idx += blockDim.x) c_sh[idx] = centers[idx];
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) { double a = 0.0; for (int q = 0; q < d; ++q) a += c_sh[k * d + q] * c_sh[k * d + q]; csq[k] = a; }
    __syncthreads();
    int changed = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double x[16];
        for (int q = 0; q < d; ++q) x[q] = X[i * d + q];
        double best = 0.0; int bl = 0;
        for (int k = 0; k < K; ++k) {
            double dot = 0.0;
            for (int q = 0; q < d; ++q) dot += x[q] * c_sh[k * d + q];
            const double v = csq[k] - 2.0 * dot;
            if (k == 0 || v < best) { best = v; bl = k; }
        }
        labels[i] = bl;
        if (labels_prev && labels_prev[i] != bl) ++changed;
    }
    if (n_changed && changed) atomicAdd(n_changed, changed);
}

// ---------------------------------------------------------------- reference mode (bit-exact)
// distances: np.linalg.norm(X.T - c, axis=1) = sqrt(left-to-right sum of separately rounded squares);
// np.argmin over the K distances: first minimum, a NaN wins over everything (first NaN).
__global__ void __launch_bounds__(256) k_refkmeans_assign(const double* X, long long n, int d, int K, const double* cen,
                                                          long long* labels) {
    extern __shared__ double c_sh[];
    for (int idx = threadIdx.x; idx < K * d; idx += blockDim.x) c_sh[idx] = cen[idx];
    __syncthreads();
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double best = 0.0; int bl = 0; bool have_nan = false;
    for (int k = 0; k < K; ++k) {
        double acc = 0.0;
        for (int q = 0; q < d; ++q) {
            const double df = __dsub_rn(X[(long long)q * n + i], c_sh[k * d + q]);
            const double sq = __dmul_rn(df, df);
            acc = (q == 0) ? sq : __dadd_rn(acc, sq);
        }
        const double dist = __dsqrt_rn(acc);
        if (have_nan) continue;
        if (dist != dist) { have_nan = true; bl = k; continue; }
        if (k == 0 || dist < best) { best = dist; bl = k; }
    }
    labels[i] = bl;
}
// centroid[k][q] = (sequential sum over members in ascending sample index, from 0.0) / count ; empty -> NaN
__global__ void k_refkmeans_update(const double* X, long long n, int d, int K, const long long* labels, double* cen) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= K * d) return;
    const int k = idx / d, q = idx - k * d;
    double acc = 0.0; long long cnt = 0;
    const double* row = X + (long long)q * n;
    for (long long i = 0; i < n; ++i) if (labels[i] == k) { acc = __dadd_rn(acc, row[i]); ++cnt; }
    cen[idx] = cnt > 0 ? __ddiv_rn(acc, (double)cnt) : __longlong_as_double(0x7ff8000000000000LL);
}

}  // namespace kmpx

using namespace kmpx;

extern "C" int kmpx_rows_center_sqnorm(double* X, long long n, int d, const double* mean, double* x_sq, void* stream) {
    if (!X) return fail_arg(1, "X is NULL");
    if (n <= 0) return fail_arg(2, "n must be > 0");
    if (d < 1) return fail_arg(3, "d must be > 0");
    k_rows_center_sqnorm<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(X, n, d, mean, x_sq);
    return check_launch("k_rows_center_sqnorm");
}

extern "C" int kmpx_kpp_candidates(const double* X, const double* x_sq, long long n, int d, const int* cand, int T,
                                   const double* closest, double* dist, double* pot, void* ws, long long ws_bytes,
                                   void* stream) {
    if (!X || !x_sq) return fail_arg(1, "X / x_sq is NULL");
    if (n <= 0) return fail_arg(3, "n must be > 0");
    if (d < 1 || d > 16) return fail_arg(4, "d must be in 1..16");
    if (!cand) return fail_arg(5, "cand is NULL");
    if (T < 1 || T > KPP_MAXT) return fail_arg(6, "T must be in 1..8");
    if (!dist || !pot) return fail_arg(8, "dist / pot is NULL");
    const int grid = (int)((n + KPP_THREADS - 1) / KPP_THREADS < num_sms() * 4 ? (n + KPP_THREADS - 1) / KPP_THREADS : num_sms() * 4);
    if (!ws || ws_bytes < (long long)grid * KPP_MAXT * (long long)sizeof(double)) return fail_arg(10, "workspace too small (148*4*8 doubles)");
    k_kpp_candidates<<<grid, KPP_THREADS, 0, (cudaStream_t)stream>>>(X, x_sq, n, d, cand, T, closest, dist, (double*)ws);
    KMPX_TRY(check_launch("k_kpp_candidates"));
    k_kpp_reduce<<<1, 32, 0, (cudaStream_t)stream>>>((const double*)ws, grid, T, pot);
    return check_launch("k_kpp_reduce");
}

extern "C" int kmpx_kmeans_assign(const double* X, long long n, int d, int K, const double* centers,
                                  const int* labels_prev, int* labels, int* n_changed, void* stream) {
    if (!X) return fail_arg(1, "X is NULL");
    if (n <= 0) return fail_arg(2, "n must be > 0");
    if (d < 1 || d > 16) return fail_arg(3, "d must be in 1..16");
    if (K < 1 || K > 1024) return fail_arg(4, "K must be in 1..1024");
    if (!centers) return fail_arg(5, "centers is NULL");
    if (!labels) return fail_arg(7, "labels is NULL");
    const long long blocks = (n + 255) / 256;
    const int grid = (int)(blocks < num_sms() * 8 ? blocks : num_sms() * 8);
    const size_t smem = ((size_t)K * d + K) * sizeof(double);
    k_kmeans_assign<<<grid, 256, smem, (cudaStream_t)stream>>>(X, n, d, K, centers, labels_prev, labels, n_changed);
    return check_launch("k_kmeans_assign");
}

extern "C" int kmpx_refkmeans_assign(const double* X, long long n, int d, int K, const double* centroids,
                                     long long* labels, void* stream) {
    if (!X) return fail_arg(1, "X is NULL");
    if (n <= 0) return fail_arg(2, "n must be > 0");
    if (d < 1) return fail_arg(3, "d must be > 0");
    if (K < 1) return fail_arg(4, "K must be > 0");
    if ((size_t)K * d * sizeof(double) > 48 * 1024) return fail_arg(4, "K*d too large");
    if (!centroids || !labels) return fail_arg(5, "centroids / labels is NULL");
    k_refkmeans_assign<<<ceil_div(n, 256), 256, (size_t)K * d * sizeof(double), (cudaStream_t)stream>>>(X, n, d, K, centroids, labels);
    return check_launch("k_refkmeans_assign");
}

extern "C" int kmpx_refkmeans_update(const double* X, long long n, int d, int K, const long long* labels,
                                     double* centroids, void* stream) {
    if (!X || !labels || !centroids) return fail_arg(1, "NULL argument");
    if (n <= 0) return fail_arg(2, "n must be > 0");
    k_refkmeans_update<<<ceil_div((long long)K * d, 64), 64, 0, (cudaStream_t)stream>>>(X, n, d, K, labels, centroids);
    return check_launch("k_refkmeans_update");
}
