// K5/K6: fused EM pass of the full-covariance Gaussian mixture behind GaussianMixtureModel.fit
// (GaussianMixtureModel.py:53 -> sklearn/mixture/_base.py:265-278; E-step _gaussian_mixture.py:490-553,
// M-step :282-320) and the M-step finalisation (:312-320, :358-370, :883-901).
//
// One pass over the samples produces the sufficient statistics {sum r, sum r x', sum r x' x'^T} per
// component and sum log p(x); responsibilities are never written to HBM.
//   E-step : log p_k(x) = Phi(x) . theta_k with the same feature row Phi = [1, x', vech(x' x'^T)] the M-step contracts and
//            theta_k the coefficients of the expanded quadratic form (built per CTA from mu_k and the precision factor
//            P_k): one (64 x FP) x (FP x K) product per tile on DMMA.  Warp w owns the samples 8 w .. 8 w + 7 and ALL
//            components, so the log-sum-exp, the arg-max and the normalisation need quad shuffles only - no shared-
//            memory exchange, no CTA barrier inside the E-step.  x' = x - shift (the data mean) keeps the cancellation
//            of the expanded form at ~1e-11 in log p, three orders below what parity needs.
//   M-step : stats[k][f] += sum_s R[s][k] * Phi[s][f] with Phi = [1, x', vech(x' x'^T)]  -- a
//            (K x TS) x (TS x F) contraction per tile on DMMA, accumulators persistent in registers.
//   Per-CTA partial sums are combined in CTA order by a second kernel => bitwise reproducible.
//
// Roofline: 8*n*d bytes per pass but ~180*n*K flops => FP64-bound for K >= 2.
#include "dmma_tile.cuh"

namespace kmpx {

constexpr int ETS = 64;           // samples per tile
constexpr int EPARTS = 4;         // threads per sample in the E-step: each owns a quarter of the components
constexpr int ETHREADS = 256;
constexpr double kLog2Pi = 1.8378770664093454835606594728112;

__host__ __device__ inline int gmm_F(int d) { return 1 + d + d * (d + 1) / 2; }
__host__ __device__ inline int pad8(int v) { return (v + 7) & ~7; }
// leading dimensions = 4 or 12 (mod 16) for conflict-free k-major fragment reads
__host__ __device__ inline int ld_conflict_free(int v) { int l = v + 4; while ((l & 15) != 4 && (l & 15) != 12) ++l; return l; }

struct EmArgs {
    const double* X; long long n; int d, K; const double* shift;
    const double* log_weights; const double* means; const double* prec_chol; const double* log_det;
    const int* hard_labels;           // != NULL: one-hot responsibilities from labels (no E-step)
    double* partials; int* labels; double* lse;
    int first_moments_only;           // hard labels for Lloyd: only [count, sum x] are accumulated (the other statistics are 0)
};

template <int D, int KB>   // KB = 8-component blocks per warp (K_pad <= 64*KB)
__global__ void __launch_bounds__(ETHREADS, 2) k_gmm_em(EmArgs p) {
    constexpr int F = 1 + D + D * (D + 1) / 2;
    constexpr int FP = (F + 7) & ~7;
    constexpr int NF = FP / 8;
    constexpr int TRI = D * (D + 1) / 2;
    extern __shared__ __align__(16) double smem[];
    const int K = p.K, KP = pad8(K);
    const int ldr = ld_conflict_free(KP), ldf = ld_conflict_free(FP);
    const int ldt = ld_conflict_free(KP);
    double* s_theta = smem;                       // [FP][ldt]: k-major B operand of the E-step product (columns >= K are zero)
    double* s_shift = s_theta + (size_t)FP * ldt; // [D] (+1 pad: what follows stays 16-byte aligned)
    double* s_x = s_shift + ((D + 1) & ~1);       // [ETS][D]
    double* s_R = s_x + (size_t)ETS * D;          // [ETS][ldr]
    double* s_Phi = s_R + (size_t)ETS * ldr;      // [ETS][ldf]
    double* s_red = s_Phi + (size_t)ETS * ldf;    // [EPARTS][ETS] max / sum exchange
    int* s_arg = reinterpret_cast<int*>(s_red + EPARTS * ETS);   // [EPARTS][ETS]
    double* s_etab = s_red + EPARTS * ETS + EPARTS * ETS / 2;     // [32] table of exp_tab
    fill_exp2_table(s_etab, threadIdx.x, ETHREADS);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const bool hard = p.hard_labels != nullptr;
    const int nf_use = p.first_moments_only ? (1 + D + 7) / 8 : NF;      // feature tiles the M-step contracts

    if (!hard) {
        // theta_k from (mu_k, P_k): Lambda = P P^T, mu' = mu - shift,
        //   log p = [c - mu'^T Lambda mu' / 2] + (Lambda mu')^T x' - sum_i Lambda_ii x'_i^2 / 2 - sum_{i<j} Lambda_ij x'_i x'_j
        for (int idx = tid; idx < FP * ldt; idx += ETHREADS) s_theta[idx] = 0.0;
        __syncthreads();
        for (int k = tid; k < K; k += ETHREADS) {
            double P[D][D], Lm[D][D], mu[D], lmu[D];
#pragma unroll
            for (int i = 0; i < D; ++i) {
                mu[i] = p.means[(size_t)k * D + i] - (p.shift ? p.shift[i] : 0.0);
#pragma unroll
                for (int j = 0; j < D; ++j) P[i][j] = j >= i ? p.prec_chol[((size_t)k * D + i) * D + j] : 0.0;
            }
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    double a = 0.0;
#pragma unroll
                    for (int c = 0; c < D; ++c) a += P[i][c] * P[j][c];
                    Lm[i][j] = a;
                }
            double quad = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                double a = 0.0;
#pragma unroll
                for (int j = 0; j < D; ++j) a += Lm[i][j] * mu[j];
                lmu[i] = a; quad += a * mu[i];
            }
            s_theta[k] = p.log_det[k] + p.log_weights[k] - 0.5 * D * kLog2Pi - 0.5 * quad;
#pragma unroll
            for (int i = 0; i < D; ++i) s_theta[(size_t)(1 + i) * ldt + k] = lmu[i];
            int t = 1 + D;
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = i; j < D; ++j) { s_theta[(size_t)t * ldt + k] = (i == j) ? -0.5 * Lm[i][i] : -Lm[i][j]; ++t; }
        }
    }
    for (int i = tid; i < D; i += ETHREADS) s_shift[i] = p.shift ? p.shift[i] : 0.0;
    __syncthreads();

    double acc[KB][NF][2];
#pragma unroll
    for (int q = 0; q < KB; ++q)
#pragma unroll
        for (int j = 0; j < NF; ++j) { acc[q][j][0] = 0.0; acc[q][j][1] = 0.0; }
    double ll_acc = 0.0;

    const int s_loc = tid % ETS, half = tid / ETS;            // sample in tile, component quarter ("half" for short)
    const int kper = (K + EPARTS - 1) / EPARTS;
    const int kbeg = min(K, half * kper), kend = min(K, kbeg + kper);
    const long long ntiles = (p.n + ETS - 1) / ETS;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long base = tile * ETS;
        const int cnt = (int)((p.n - base) < ETS ? (p.n - base) : ETS);
        __syncthreads();                                        // previous tile's M-step is done with s_R / s_Phi
        for (int idx = tid; idx < ETS * D; idx += ETHREADS) s_x[idx] = idx < cnt * D ? p.X[base * D + idx] : 0.0;
        __syncthreads();
        const bool live = s_loc < cnt;
        double x[D];
#pragma unroll
        for (int i = 0; i < D; ++i) x[i] = s_x[s_loc * D + i];
        {                                                        // Phi row = [1, x', vech(x' x'^T)]: the four threads of a
            double xs[D];                                        // sample write every fourth feature
#pragma unroll
            for (int i = 0; i < D; ++i) xs[i] = x[i] - s_shift[i];
            double* ph = s_Phi + (size_t)s_loc * ldf;
            if (half == 0) ph[0] = live ? 1.0 : 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) if (((1 + i) & 3) == half) ph[1 + i] = live ? xs[i] : 0.0;
            int t = 1 + D;
            if (!p.first_moments_only) {
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int j = i; j < D; ++j) { if ((t & 3) == half) ph[t] = live ? xs[i] * xs[j] : 0.0; ++t; }
            }
#pragma unroll
            for (int tt = 1 + D; tt < FP; ++tt) if (tt >= t && (tt & 3) == half) ph[tt] = 0.0;
        }
        double* Rrow = s_R + (size_t)s_loc * ldr;
        if (hard) {
            const int lab = live ? p.hard_labels[base + s_loc] : -1;
            for (int k = kbeg; k < kend; ++k) Rrow[k] = (k == lab) ? 1.0 : 0.0;
            if (half == EPARTS - 1) for (int k = K; k < KP; ++k) Rrow[k] = 0.0;
        } else {
            __syncthreads();                                     // the Phi rows of this tile are complete
            // ---- E-step on the tensor pipe: warp w, samples 8 w + g, all components.  lp[j][e] = weighted log-probability
            //      of component 8 j + 2 t4 + e (accumulator layout of an 8 x 8 DMMA tile).
            constexpr int NTE = 8 * KB;                          // component tiles of 8 (K_pad <= 64 KB)
            const int nte = KP >> 3;
            double lp[NTE][2];
#pragma unroll
            for (int j = 0; j < NTE; ++j) { lp[j][0] = 0.0; lp[j][1] = 0.0; }
            const double* arow = s_Phi + (size_t)(8 * warp + g) * ldf + t4;
            const double* brow = s_theta + (size_t)t4 * ldt + g;
#pragma unroll
            for (int ks = 0; ks < FP / 4; ++ks) {
                const double a = arow[4 * ks];
#pragma unroll
                for (int j = 0; j < NTE; ++j)
                    if (j < nte) dmma884(lp[j], a, brow[(size_t)4 * ks * ldt + 8 * j]);
            }
            // first maximum over the components (np.argmax rule), quad-wide
            double mx = -INFINITY; int am = 0;
#pragma unroll
            for (int j = 0; j < NTE; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = 8 * j + 2 * t4 + e;
                    if (k >= K) lp[j][e] = -INFINITY;
                    if (lp[j][e] > mx) { mx = lp[j][e]; am = k; }
                }
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                const double mo = __shfl_xor_sync(0xffffffffu, mx, o);
                const int ao = __shfl_xor_sync(0xffffffffu, am, o);
                if (mo > mx || (mo == mx && ao < am)) { mx = mo; am = ao; }
            }
            // exponentials as four interleaved branch-free chains (arguments <= 0; -inf beyond K gives 0)
            double se = 0.0;
#pragma unroll
            for (int j = 0; j < NTE; j += 2) {
                if (j < nte) {
                    double e4[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) e4[u] = exp_tab(lp[j + (u >> 1)][u & 1] - mx, s_etab);
#pragma unroll
                    for (int u = 0; u < 4; ++u) { lp[j + (u >> 1)][u & 1] = e4[u]; se += e4[u]; }
                }
            }
            se += __shfl_xor_sync(0xffffffffu, se, 1);
            se += __shfl_xor_sync(0xffffffffu, se, 2);
            const int srow = 8 * warp + g;
            const bool live_r = srow < cnt;
            const double inv = live_r ? 1.0 / se : 0.0;
            double* Rr = s_R + (size_t)srow * ldr + 2 * t4;
#pragma unroll
            for (int j = 0; j < NTE; ++j)
                if (j < nte) *reinterpret_cast<double2*>(Rr + 8 * j) = make_double2(lp[j][0] * inv, lp[j][1] * inv);
            if (t4 == 0 && live_r) {
                const double l = mx + log(se);
                ll_acc += l;
                if (p.lse) p.lse[base + srow] = l;
                if (p.labels) p.labels[base + srow] = am;
            }
        }
        __syncthreads();
        // M-step tile: stats[k][f] += sum_s R[s][k] Phi[s][f]
#pragma unroll
        for (int q = 0; q < KB; ++q) {
            const int kb = warp + q * 8;                          // 8-component block of this warp
            if (kb * 8 < KP) {
                for (int s0 = 0; s0 < ETS; s0 += 4) {
                    const double a = s_R[(size_t)(s0 + t4) * ldr + kb * 8 + g];
#pragma unroll
                    for (int j = 0; j < NF; ++j) {
                        if (j < nf_use) {
                            const double bv = s_Phi[(size_t)(s0 + t4) * ldf + j * 8 + g];
                            dmma884(acc[q][j], a, bv);
                        }
                    }
                }
            }
        }
    }
    // per-CTA partials: [KP][FP] then the log-likelihood
    double* part = p.partials + (size_t)blockIdx.x * ((size_t)KP * FP + 1);
#pragma unroll
    for (int q = 0; q < KB; ++q) {
        const int kb = warp + q * 8;
        if (kb * 8 < KP) {
#pragma unroll
            for (int j = 0; j < NF; ++j) {
                const int k = kb * 8 + g, f = j * 8 + 2 * t4;
                part[(size_t)k * FP + f] = acc[q][j][0];
                part[(size_t)k * FP + f + 1] = acc[q][j][1];
            }
        }
    }
    __syncthreads();
    if (t4 == 0) s_red[8 * warp + g] = ll_acc;                // one accumulator per sample slot of the tile, summed in slot order
    __syncthreads();
    if (tid == 0) { double t = 0.0; for (int i = 0; i < ETS; ++i) t += s_red[i]; part[(size_t)KP * FP] = t; }
}

// stats[k*F + f] = sum over CTAs (fixed order) of partials; stats[K*F] = log-likelihood sum
__global__ void k_gmm_reduce(const double* partials, int nctas, int K, int KP, int F, int FP, double* stats) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)KP * FP + 1;
    if (idx < K * F) {
        const int k = idx / F, f = idx - k * F;
        double s = 0.0;
        for (int c = 0; c < nctas; ++c) s += partials[c * stride + (size_t)k * FP + f];
        stats[idx] = s;
    } else if (idx == K * F) {
        double s = 0.0;
        for (int c = 0; c < nctas; ++c) s += partials[c * stride + (size_t)KP * FP];
        stats[idx] = s;
    }
}

struct FinArgs {
    const double* stats; double n_total; int first, d, K; double reg; const double* shift;
    double *weights, *log_weights, *means, *covs, *prec_chol, *log_det; int* info;
    double* em_state; double tol; int max_iter;   // optional device-side loop control, see kmpx_gmm_finalize_check
};

// One thread per component (K <= 1024, d <= 16).
__global__ void k_gmm_finalize(FinArgs p) {
    __shared__ double s_nk[1024];
    const int k = threadIdx.x, d = p.d, K = p.K, F = gmm_F(d);
    // sklearn's loop (mixture/_base.py:265-278) on the device: em_state = {done, n_iter, lower, converged}.  Once done,
    // the M-step is not applied any more, so the host may enqueue passes in batches and read the state once per batch.
    if (p.em_state) {
        if (p.em_state[0] != 0.0) return;
        __syncthreads();                       // every thread has read `done` before thread 0 may set it
        if (k == 0) {
            const double lower = p.stats[(size_t)K * F] / p.n_total, prev = p.em_state[2];
            const double it = p.em_state[1] + 1.0;
            p.em_state[1] = it; p.em_state[2] = lower;
            if (fabs(lower - prev) < p.tol) { p.em_state[0] = 1.0; p.em_state[3] = 1.0; }
            else if (it >= (double)p.max_iter) p.em_state[0] = 1.0;
        }
    }
    double nk = 0.0;
    if (k < K) { nk = p.stats[(size_t)k * F] + 10.0 * 2.220446049250313e-16; s_nk[k] = nk; }
    __syncthreads();
    if (k >= K) return;
    double tot = 0.0;
    if (p.first) tot = p.n_total; else for (int j = 0; j < K; ++j) tot += s_nk[j];
    const double wk = nk / tot;
    p.weights[k] = wk;
    p.log_weights[k] = log(wk);
    const double* st = p.stats + (size_t)k * F;
    double mu[16], cov[16][16], L[16][16], Li[16][16];
    for (int i = 0; i < d; ++i) mu[i] = st[1 + i] / nk;
    int t = 1 + d;
    for (int i = 0; i < d; ++i)
        for (int j = i; j < d; ++j) {
            double c = st[t++] / nk - mu[i] * mu[j];
            if (i == j) c += p.reg;
            cov[i][j] = c; cov[j][i] = c;
        }
    for (int i = 0; i < d; ++i) p.means[(size_t)k * d + i] = mu[i] + (p.shift ? p.shift[i] : 0.0);
    for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) p.covs[((size_t)k * d + i) * d + j] = cov[i][j];
    // Cholesky cov = L L^T
    int bad = 0;
    for (int j = 0; j < d; ++j) {
        double s = cov[j][j];
        for (int c = 0; c < j; ++c) s -= L[j][c] * L[j][c];
        if (!(s > 0.0) && bad == 0) bad = j + 1;
        const double dj = sqrt(s);
        L[j][j] = dj;
        for (int i = j + 1; i < d; ++i) {
            double v = cov[i][j];
            for (int c = 0; c < j; ++c) v -= L[i][c] * L[j][c];
            L[i][j] = v / dj;
        }
    }
    // Li = L^{-1} (lower); precisions_chol = Li^T (upper)
    double ld = 0.0;
    for (int c = 0; c < d; ++c) {
        for (int i = 0; i < c; ++i) Li[i][c] = 0.0;
        Li[c][c] = 1.0 / L[c][c];
        for (int i = c + 1; i < d; ++i) {
            double s = 0.0;
            for (int tt = c; tt < i; ++tt) s += L[i][tt] * Li[tt][c];
            Li[i][c] = -s / L[i][i];
        }
    }
    for (int i = 0; i < d; ++i) {
        for (int j = 0; j < d; ++j) p.prec_chol[((size_t)k * d + i) * d + j] = (j >= i) ? Li[j][i] : 0.0;
        ld += log(Li[i][i]);
    }
    p.log_det[k] = ld;
    if (p.info) p.info[k] = bad;
}

static int em_grid(long long n) {
    const long long tiles = (n + ETS - 1) / ETS;
    const long long g = tiles < 2 * num_sms() ? tiles : 2 * num_sms();      // two CTAs per SM
    return (int)(g < 1 ? 1 : g);
}
static size_t em_smem(int d, int K) {
    const int F = gmm_F(d), FP = pad8(F), KP = pad8(K), TRI = d * (d + 1) / 2;
    (void)TRI;
    size_t dbl = (size_t)FP * ld_conflict_free(KP) + ((d + 1) & ~1) + (size_t)ETS * d + (size_t)ETS * ld_conflict_free(KP) +
                 (size_t)ETS * ld_conflict_free(FP) + 2 * EPARTS * ETS + kExpTabN;
    return dbl * sizeof(double);
}

template <int D>
static int launch_em_d(const EmArgs& p, int grid, size_t smem, cudaStream_t st) {
    const int KP = pad8(p.K);
    if (KP <= 64) {
        KMPX_TRY(check_cuda(cudaFuncSetAttribute(k_gmm_em<D, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "attr k_gmm_em"));
        k_gmm_em<D, 1><<<grid, ETHREADS, smem, st>>>(p);
    } else {
        KMPX_TRY(check_cuda(cudaFuncSetAttribute(k_gmm_em<D, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "attr k_gmm_em"));
        k_gmm_em<D, 2><<<grid, ETHREADS, smem, st>>>(p);
    }
    return check_launch("k_gmm_em");
}

static int run_em(EmArgs p, double* stats, void* ws, long long ws_bytes, cudaStream_t st) {
    if (!p.X) return fail_arg(1, "X is NULL");
    if (p.n <= 0) return fail_arg(2, "n must be > 0");
    if (p.d < 1 || p.d > 8) return fail_arg(3, "d must be in 1..8");
    if (p.K < 1 || p.K > 128) return fail_arg(4, "K must be in 1..128");
    if (!stats) return fail_arg(10, "stats is NULL");
    const int F = gmm_F(p.d), FP = pad8(F), KP = pad8(p.K);
    const int grid = em_grid(p.n);
    const long long need = (long long)grid * ((long long)KP * FP + 1) * (long long)sizeof(double);
    if (!ws || ws_bytes < need) return fail_arg(13, "workspace too small (kmpx_gmm_workspace_bytes)");
    p.partials = (double*)ws;
    const size_t smem = em_smem(p.d, p.K);
    if (smem > 227 * 1024) return fail_arg(4, "K*d too large for the shared-memory tile");
    int rc;
    switch (p.d) {
        case 1: rc = launch_em_d<1>(p, grid, smem, st); break;
        case 2: rc = launch_em_d<2>(p, grid, smem, st); break;
        case 3: rc = launch_em_d<3>(p, grid, smem, st); break;
        case 4: rc = launch_em_d<4>(p, grid, smem, st); break;
        case 5: rc = launch_em_d<5>(p, grid, smem, st); break;
        case 6: rc = launch_em_d<6>(p, grid, smem, st); break;
        case 7: rc = launch_em_d<7>(p, grid, smem, st); break;
        default: rc = launch_em_d<8>(p, grid, smem, st); break;
    }
    KMPX_TRY(rc);
    k_gmm_reduce<<<ceil_div(p.K * F + 1, 128), 128, 0, st>>>(p.partials, grid, p.K, KP, F, FP, stats);
    return check_launch("k_gmm_reduce");
}

}  // namespace kmpx

using namespace kmpx;

extern "C" long long kmpx_gmm_workspace_bytes(int n, int d, int K) {
    (void)n;
    const int F = gmm_F(d), FP = pad8(F), KP = pad8(K);
    const int sms = num_sms() > kNumSMsB200 ? num_sms() : kNumSMsB200;
    return 2LL * sms * ((long long)KP * FP + 1) * (long long)sizeof(double);
}

extern "C" int kmpx_gmm_em_pass(const double* X, long long n, int d, int K, const double* shift,
                                const double* log_weights, const double* means, const double* prec_chol,
                                const double* log_det, double* stats, int* labels, double* lse,
                                void* ws, long long ws_bytes, void* stream) {
    if (!log_weights || !means || !prec_chol || !log_det) return fail_arg(6, "mixture parameters are NULL");
    EmArgs p{X, n, d, K, shift, log_weights, means, prec_chol, log_det, nullptr, nullptr, labels, lse, 0};
    return run_em(p, stats, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int kmpx_gmm_label_stats(const double* X, long long n, int d, int K, const double* shift,
                                    const int* labels, double* stats, void* ws, long long ws_bytes, void* stream) {
    if (!labels) return fail_arg(6, "labels is NULL");
    EmArgs p{X, n, d, K, shift, nullptr, nullptr, nullptr, nullptr, labels, nullptr, nullptr, nullptr, 0};
    return run_em(p, stats, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int kmpx_gmm_label_sums(const double* X, long long n, int d, int K, const double* shift,
                                   const int* labels, double* stats, void* ws, long long ws_bytes, void* stream) {
    if (!labels) return fail_arg(6, "labels is NULL");
    EmArgs p{X, n, d, K, shift, nullptr, nullptr, nullptr, nullptr, labels, nullptr, nullptr, nullptr, 1};
    return run_em(p, stats, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int kmpx_gmm_finalize(const double* stats, double n_total, int first, int d, int K, double reg,
                                 const double* shift, double* weights, double* log_weights, double* means,
                                 double* covs, double* prec_chol, double* log_det, int* info, void* stream) {
    if (!stats) return fail_arg(1, "stats is NULL");
    if (d < 1 || d > 16) return fail_arg(4, "d must be in 1..16");
    if (K < 1 || K > 1024) return fail_arg(5, "K must be in 1..1024");
    if (!weights || !log_weights || !means || !covs || !prec_chol || !log_det) return fail_arg(8, "output arrays are NULL");
    FinArgs p{stats, n_total, first, d, K, reg, shift, weights, log_weights, means, covs, prec_chol, log_det, info, nullptr, 0.0, 0};
    k_gmm_finalize<<<1, ((K + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(p);
    return check_launch("k_gmm_finalize");
}

extern "C" int kmpx_gmm_finalize_check(const double* stats, double n_total, int d, int K, double reg,
                                       const double* shift, double* weights, double* log_weights, double* means,
                                       double* covs, double* prec_chol, double* log_det, int* info,
                                       double tol, int max_iter, double* em_state, void* stream) {
    if (!stats || !em_state) return fail_arg(1, "stats / em_state is NULL");
    if (d < 1 || d > 16) return fail_arg(4, "d must be in 1..16");
    if (K < 1 || K > 1024) return fail_arg(5, "K must be in 1..1024");
    if (!weights || !log_weights || !means || !covs || !prec_chol || !log_det) return fail_arg(8, "output arrays are NULL");
    FinArgs p{stats, n_total, 0, d, K, reg, shift, weights, log_weights, means, covs, prec_chol, log_det, info, em_state, tol, max_iter};
    k_gmm_finalize<<<1, ((K + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(p);
    return check_launch("k_gmm_finalize");
}
