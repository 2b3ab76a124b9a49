// K8: Gaussian mixture regression  (GaussianMixtureModel.predict, GaussianMixtureModel.py:59-107)
//
//   h_k = pi_k N(x; mu_k^I, S_k^II);  h /= sum_k (h_k + REALMIN)                        (:85-90)
//   m_k = mu_k^O + S_k^OI (S_k^II)^-1 (x - mu_k^I);  mean = sum h_k m_k                  (:92-96)
//   cov = sum h_k (S_k^OO - S_k^OI (S_k^II)^-1 S_k^IO + m_k m_k^T) + reg I - mean mean^T (:98-105)
//
// Per-component constants (gain G_k, conditional covariance, log-normaliser, chol(S^II)^-1) are hoisted
// into shared memory by the first K threads of every CTA; then one thread per query.  Inputs x [I][M] and
// outputs mu [O][M], sigma [O][O][M] keep the reference's query-last layout, so every access is coalesced.
// Bytes per query: 8 (I + O + O^2); flops ~ K (30 + 2 I O + 2 O + 3 O (O+1)): at the ridge for K=8, O=2.
#include "common.cuh"

namespace kmpx {

constexpr int GMR_MAXI = 4;
constexpr double kRealMin = 2.2250738585072014e-308;
constexpr double kLog2PiG = 1.8378770664093454835606594728112;

struct GmrArgs {
    const double* priors; const double* means; const double* covs; int K, I, O;
    const double* x; long long M; double reg; double* mu; double* sigma;
};

// shared layout per component: [prior, logc, muI (I), Linv (I*I), muO (O), G (O*I), Sc (O*O)]
__host__ __device__ inline int gmr_stride(int I, int O) { return 2 + I + I * I + O + O * I + O * O; }

template <int O, bool SMALLK>       // SMALLK: K <= 8, the scalar-input path keeps the exponents in registers
__global__ void __launch_bounds__(128) k_gmr(GmrArgs p) {
    extern __shared__ double sh[];
    const int I = p.I, K = p.K, d = I + O;
    const int stride = gmr_stride(I, O);
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        double* c = sh + (size_t)k * stride;
        double S[GMR_MAXI][GMR_MAXI], L[GMR_MAXI][GMR_MAXI], Li[GMR_MAXI][GMR_MAXI], Sinv[GMR_MAXI][GMR_MAXI];
        for (int i = 0; i < I; ++i) for (int j = 0; j < I; ++j) S[i][j] = p.covs[((size_t)i * d + j) * K + k];
        // Cholesky of S^II and its inverse
        for (int j = 0; j < I; ++j) {
            double s = S[j][j];
            for (int t = 0; t < j; ++t) s -= L[j][t] * L[j][t];
            const double dj = sqrt(s);
            L[j][j] = dj;
            for (int i = j + 1; i < I; ++i) {
                double v = S[i][j];
                for (int t = 0; t < j; ++t) v -= L[i][t] * L[j][t];
                L[i][j] = v / dj;
            }
        }
        double logdet = 0.0;
        for (int cc = 0; cc < I; ++cc) {
            for (int i = 0; i < cc; ++i) Li[i][cc] = 0.0;
            Li[cc][cc] = 1.0 / L[cc][cc];
            for (int i = cc + 1; i < I; ++i) {
                double s = 0.0;
                for (int t = cc; t < i; ++t) s += L[i][t] * Li[t][cc];
                Li[i][cc] = -s / L[i][i];
            }
            logdet += log(L[cc][cc]);
        }
        for (int i = 0; i < I; ++i)
            for (int j = 0; j < I; ++j) {
                double s = 0.0;
                for (int t = 0; t < I; ++t) s += Li[t][i] * Li[t][j];
                Sinv[i][j] = (I == 1) ? 1.0 / S[0][0] : s;
            }
        c[0] = p.priors[k];
        c[1] = -0.5 * I * kLog2PiG - logdet;
        double* muI = c + 2; double* Lv = muI + I; double* muO = Lv + I * I; double* G = muO + O; double* Sc = G + O * I;
        for (int i = 0; i < I; ++i) muI[i] = p.means[(size_t)i * K + k];
        for (int i = 0; i < I; ++i) for (int j = 0; j < I; ++j) Lv[i * I + j] = Li[i][j];
        for (int a = 0; a < O; ++a) muO[a] = p.means[(size_t)(I + a) * K + k];
        for (int a = 0; a < O; ++a)
            for (int j = 0; j < I; ++j) {
                double s = 0.0;
                for (int t = 0; t < I; ++t) s += p.covs[((size_t)(I + a) * d + t) * K + k] * Sinv[t][j];
                G[a * I + j] = s;
            }
        for (int a = 0; a < O; ++a)
            for (int b = 0; b < O; ++b) {
                double s = 0.0;
                for (int t = 0; t < I; ++t) s += G[a * I + t] * p.covs[((size_t)t * d + I + b) * K + k];
                Sc[a * O + b] = p.covs[((size_t)(I + a) * d + I + b) * K + k] - s;
            }
    }
    __syncthreads();
    if (I == 1) {
        // Scalar input (every demo of the reference).  The kernel is bound by FP64 issue, not by HBM, unless the work per
        // query shrinks (56 bytes against ~40 FP64 operations per component), so:
        //   * pass 1 evaluates only the exponents  e_k = log(pi_k) + logc_k - z_k^2/2  (2 FMA + 1 max per component) and
        //     their maximum; pass 2 recomputes them and evaluates exp / gain / second moments only for components within
        //     60 ln 2 of that maximum: anything smaller contributes less than 2^-60 of the leading activation - far below
        //     half an ulp of every output - and is accounted as exactly 0 (+ REALMIN in the normaliser, like every
        //     component, GaussianMixtureModel.py:90).  On a time grid neighbouring queries activate the same 2-3 components,
        //     so the branch is warp-uniform; the per-lane select keeps a query's result independent of its neighbours;
        //   * two queries per thread (m, m + half a CTA-stride apart: both accesses stay coalesced) run as independent
        //     dependency chains, and the covariance accumulates its upper triangle only.
        // Per-component record for this path (overlay on the generic one, rebuilt here):
        //   [a = Linv/sqrt2, b = -a muI, e0 = logc + log prior, aO (O) = muO - G muI, G (O), Sc upper (O(O+1)/2)]
        constexpr int NP = O * (O + 1) / 2;
        constexpr int RS = 3 + 2 * O + NP;
        double* rec = sh + (size_t)K * stride;
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            const double* c = sh + (size_t)k * stride;
            const double* muO = c + 4; const double* G = muO + O; const double* Sc = G + O;
            double* r = rec + (size_t)k * RS;
            const double a = c[3] * 0.70710678118654752440;
            r[0] = a; r[1] = -a * c[2]; r[2] = c[1] + log(c[0]);
            for (int o = 0; o < O; ++o) { r[3 + o] = muO[o] - G[o] * c[2]; r[3 + O + o] = G[o]; }
            int t = 0;
            for (int o = 0; o < O; ++o) for (int q = o; q < O; ++q) r[3 + 2 * O + t++] = Sc[o * O + q];
        }
        __syncthreads();
        constexpr double kSkip = 41.58883083359672;      // 60 ln 2
        constexpr int CB = 4;                            // components per block: CB x 2 queries = 8 independent exp chains
        double* etab = rec + (size_t)K * RS;
        fill_exp2_table(etab, threadIdx.x, blockDim.x);
        __syncthreads();
        // a thread takes two ADJACENT queries (a warp 64 consecutive ones): on a sorted query grid the lanes of a warp then
        // share their active components and whole blocks of components are skipped; 16-byte loads / stores when M is even
        const long long span = (long long)gridDim.x * blockDim.x;
        const bool vec2 = (p.M & 1) == 0 && (reinterpret_cast<uintptr_t>(p.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.mu) & 15) == 0 &&
                          (reinterpret_cast<uintptr_t>(p.sigma) & 15) == 0;
        for (long long m0 = 2 * (blockIdx.x * (long long)blockDim.x + threadIdx.x); m0 < p.M; m0 += 2 * span) {
            const long long m1 = m0 + 1;
            const bool two = m1 < p.M;
            double x[2];
            if (vec2) { const double2 xv = *reinterpret_cast<const double2*>(p.x + m0); x[0] = xv.x; x[1] = xv.y; }
            else { x[0] = p.x[m0]; x[1] = two ? p.x[m1] : 0.0; }
            double mx[2] = {-INFINITY, -INFINITY};
            // K <= 8 (every mixture of the reference's demos): the exponents stay in registers between the two passes
            constexpr int KR = 8;
            constexpr bool keep = SMALLK;
            double er[SMALLK ? KR : 1][2];
            if constexpr (keep) {
#pragma unroll
                for (int k = 0; k < KR; ++k) {
                    const double* r = rec + (size_t)min(k, K - 1) * RS;
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const double z = fma(r[0], x[u], r[1]);
                        er[k][u] = fma(-z, z, r[2]);
                        mx[u] = fmax(mx[u], er[k][u]);
                    }
                }
            } else {
                for (int k = 0; k < K; ++k) {
                    const double* r = rec + (size_t)k * RS;
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const double z = fma(r[0], x[u], r[1]);
                        mx[u] = fmax(mx[u], fma(-z, z, r[2]));
                    }
                }
            }
            const double thr[2] = {mx[0] - kSkip, mx[1] - kSkip};
            double mean[2][O], sec[2][NP], hsum[2] = {0.0, 0.0};
#pragma unroll
            for (int u = 0; u < 2; ++u) {
#pragma unroll
                for (int o = 0; o < O; ++o) mean[u][o] = 0.0;
#pragma unroll
                for (int t = 0; t < NP; ++t) sec[u][t] = 0.0;
            }
            auto block = [&](int k0, const double (&e)[CB][2]) {
                bool act[CB][2]; bool any = false;
#pragma unroll
                for (int j = 0; j < CB; ++j)
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        // a NaN query stays active and propagates, as in the reference
                        act[j][u] = (k0 + j < K) && !(e[j][u] < thr[u]);
                        any = any || act[j][u];
                    }
                if (!__any_sync(0xffffffffu, any)) return;
                double h[CB][2];
#pragma unroll
                for (int j = 0; j < CB; ++j)
#pragma unroll
                    for (int u = 0; u < 2; ++u) h[j][u] = exp_tab(e[j][u], etab);
#pragma unroll
                for (int j = 0; j < CB; ++j) {
                    if (k0 + j < K) {
                        const double* r = rec + (size_t)(k0 + j) * RS;
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const double hh = act[j][u] ? h[j][u] : 0.0;
                            hsum[u] += hh;
                            double mk[O];
#pragma unroll
                            for (int o = 0; o < O; ++o) { mk[o] = fma(r[3 + O + o], x[u], r[3 + o]); mean[u][o] = fma(hh, mk[o], mean[u][o]); }
                            int t = 0;
#pragma unroll
                            for (int o = 0; o < O; ++o)
#pragma unroll
                                for (int q = o; q < O; ++q) { sec[u][t] = fma(hh, fma(mk[o], mk[q], r[3 + 2 * O + t]), sec[u][t]); ++t; }
                        }
                    }
                }
            };
            if constexpr (keep) {
#pragma unroll
                for (int k0 = 0; k0 < KR; k0 += CB) {
                    if (k0 < K) {
                        double e[CB][2];
#pragma unroll
                        for (int j = 0; j < CB; ++j) { e[j][0] = er[k0 + j][0]; e[j][1] = er[k0 + j][1]; }
                        block(k0, e);
                    }
                }
            } else {
                for (int k0 = 0; k0 < K; k0 += CB) {
                    double e[CB][2];
#pragma unroll
                    for (int j = 0; j < CB; ++j) {
                        const double* r = rec + (size_t)min(k0 + j, K - 1) * RS;
#pragma unroll
                        for (int u = 0; u < 2; ++u) { const double z = fma(r[0], x[u], r[1]); e[j][u] = fma(-z, z, r[2]); }
                    }
                    block(k0, e);
                }
            }
            // every component adds REALMIN to the normaliser (GaussianMixtureModel.py:90); h + REALMIN == h unless h < 2^-970
            hsum[0] += K * kRealMin; hsum[1] += K * kRealMin;
            double cv[2][NP];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const double inv = 1.0 / hsum[u];
#pragma unroll
                for (int o = 0; o < O; ++o) mean[u][o] *= inv;
                int t = 0;
#pragma unroll
                for (int o = 0; o < O; ++o)
#pragma unroll
                    for (int q = o; q < O; ++q) { cv[u][t] = sec[u][t] * inv + (o == q ? p.reg : 0.0) - mean[u][o] * mean[u][q]; ++t; }
            }
            if (vec2) {
                int t = 0;
#pragma unroll
                for (int o = 0; o < O; ++o) {
                    *reinterpret_cast<double2*>(p.mu + (size_t)o * p.M + m0) = make_double2(mean[0][o], mean[1][o]);
#pragma unroll
                    for (int q = o; q < O; ++q) {
                        const double2 v = make_double2(cv[0][t], cv[1][t]);
                        ++t;
                        *reinterpret_cast<double2*>(p.sigma + ((size_t)o * O + q) * p.M + m0) = v;
                        if (q != o) *reinterpret_cast<double2*>(p.sigma + ((size_t)q * O + o) * p.M + m0) = v;
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (u == 1 && !two) break;
                    const long long m = u ? m1 : m0;
                    int t = 0;
#pragma unroll
                    for (int o = 0; o < O; ++o) {
                        p.mu[(size_t)o * p.M + m] = mean[u][o];
#pragma unroll
                        for (int q = o; q < O; ++q) {
                            const double v = cv[u][t++];
                            p.sigma[((size_t)o * O + q) * p.M + m] = v;
                            if (q != o) p.sigma[((size_t)q * O + o) * p.M + m] = v;
                        }
                    }
                }
            }
        }
        return;
    }
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < p.M; m += (long long)gridDim.x * blockDim.x) {
        double x[GMR_MAXI];
        for (int i = 0; i < I; ++i) x[i] = p.x[(size_t)i * p.M + m];
        double mean[O], sec[O][O];
#pragma unroll
        for (int a = 0; a < O; ++a) { mean[a] = 0.0;
#pragma unroll
            for (int b = 0; b < O; ++b) sec[a][b] = 0.0; }
        double hsum = 0.0;
        for (int k = 0; k < K; ++k) {
            const double* c = sh + (size_t)k * stride;
            const double* muI = c + 2; const double* Lv = muI + I; const double* muO = Lv + I * I;
            const double* G = muO + O; const double* Sc = G + O * I;
            double dev[GMR_MAXI], q = 0.0;
            for (int i = 0; i < I; ++i) dev[i] = x[i] - muI[i];
            for (int i = 0; i < I; ++i) { double z = 0.0; for (int t = 0; t <= i; ++t) z += Lv[i * I + t] * dev[t]; q += z * z; }
            const double h = c[0] * exp(c[1] - 0.5 * q);
            hsum += h + kRealMin;
            double mk[O];
#pragma unroll
            for (int a = 0; a < O; ++a) { double s = muO[a]; for (int t = 0; t < I; ++t) s += G[a * I + t] * dev[t]; mk[a] = s; mean[a] += h * s; }
#pragma unroll
            for (int a = 0; a < O; ++a)
#pragma unroll
                for (int b = 0; b < O; ++b) sec[a][b] += h * (Sc[a * O + b] + mk[a] * mk[b]);
        }
        const double inv = 1.0 / hsum;
#pragma unroll
        for (int a = 0; a < O; ++a) mean[a] *= inv;
#pragma unroll
        for (int a = 0; a < O; ++a) {
            p.mu[(size_t)a * p.M + m] = mean[a];
#pragma unroll
            for (int b = 0; b < O; ++b)
                p.sigma[((size_t)a * O + b) * p.M + m] = sec[a][b] * inv + (a == b ? p.reg : 0.0) - mean[a] * mean[b];
        }
    }
}

template <int O>
static int launch_gmr(const GmrArgs& p, cudaStream_t st) {
    const size_t smem = ((size_t)p.K * (gmr_stride(p.I, O) + 3 + 2 * O + O * (O + 1) / 2) + kExpTabN) * sizeof(double);
    if (smem > 200 * 1024) { snprintf(g_err, sizeof(g_err), "kmpx: too many mixture components for k_gmr"); return -4; }
    if (p.K <= 8 && p.I == 1) {
        KMPX_TRY(check_cuda(cudaFuncSetAttribute(k_gmr<O, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "attr k_gmr"));
        const long long blocks8 = (p.M + 255) / 256;
        k_gmr<O, true><<<(int)(blocks8 < num_sms() * 16 ? blocks8 : num_sms() * 16), 128, smem, st>>>(p);
        return check_launch("k_gmr");
    }
    KMPX_TRY(check_cuda(cudaFuncSetAttribute(k_gmr<O, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "attr k_gmr"));
    const long long blocks = (p.M + 255) / 256;             // two queries per thread on the scalar-input path
    const int grid = (int)(blocks < num_sms() * 16 ? blocks : num_sms() * 16);
    k_gmr<O, false><<<grid, 128, smem, st>>>(p);
    return check_launch("k_gmr");
}

}  // namespace kmpx

using namespace kmpx;

extern "C" int kmpx_gmr(const double* priors, const double* means, const double* covs, int K, int I, int O,
                        const double* x, int M, double reg, double* mu, double* sigma_out, void* stream) {
    if (!priors || !means || !covs) return fail_arg(1, "mixture parameters are NULL");
    if (K < 1) return fail_arg(4, "K must be > 0");
    if (I < 1 || I > GMR_MAXI) return fail_arg(5, "I must be in 1..4");
    if (O < 1 || O > 8) return fail_arg(6, "O must be in 1..8");
    if (!x) return fail_arg(7, "x is NULL");
    if (M <= 0) return fail_arg(8, "M must be > 0");
    if (!mu || !sigma_out) return fail_arg(10, "outputs are NULL");
    GmrArgs p{priors, means, covs, K, I, O, x, M, reg, mu, sigma_out};
    cudaStream_t st = (cudaStream_t)stream;
    switch (O) {
        case 1: return launch_gmr<1>(p, st);
        case 2: return launch_gmr<2>(p, st);
        case 3: return launch_gmr<3>(p, st);
        case 4: return launch_gmr<4>(p, st);
        case 5: return launch_gmr<5>(p, st);
        case 6: return launch_gmr<6>(p, st);
        case 7: return launch_gmr<7>(p, st);
        default: return launch_gmr<8>(p, st);
    }
}
