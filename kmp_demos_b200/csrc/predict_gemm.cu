// K4 for one LARGE system (cfg 4 of BASELINE.json: N = 8192, n = 16384, 100 000 queries): the posterior covariance as
// plain DMMA GEMMs instead of the fused streaming kernel.
//   kappaT[m][j] = k(s*_m, s_j)                             (Mc x N, query-major, once per chunk of Mc queries)
//   V_a = Linv[a Ns :, a Ns : a Ns + N] * kappaT^T          (NT product on the TMA-fed k_dgemm_nt_tma, tri = 2: the
//                                                            operand is lower triangular, k-tiles right of the diagonal
//                                                            tile are skipped)
//   xi, Sigma: column reductions of V_a (.) V_b and V_a (.) w over the rows, two passes with a fixed order
// The fused kernel (predict_stream.cu) regenerates the kappa tile for every 128-row tile of Linv and hands every
// 16 columns over through a CTA barrier; here the same O^2 Ns^2 M flops run at the rate of k_dgemm (128 x 128 x 32
// tiles, ~30 TFLOP/s) and the extra HBM traffic - V written and read once, 8 (n + Ns) Mc bytes - costs < 1 ms per
// 8192 queries.  Precondition: the strict upper triangle of F holds zeros (kmpx_trtri guarantees it for n > 512).
#include "kmpx_internal.cuh"

namespace kmpx {

constexpr int PG_RB = 64;       // row blocks of the first reduction pass
constexpr int PG_T = 128;       // queries per CTA of the reduction kernels

// kappaT[m][j], m < Mc (rows beyond the M live queries are zero), j < ldk (columns beyond N are zero)
__global__ void __launch_bounds__(256) k_pg_kappa(const double* s, int N, int I, const double* sq, int M, int Mc, int ldk, double sigma_f, double* kappa) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)Mc * ldk) return;
    const int m = (int)(idx / ldk), j = (int)(idx - (long long)m * ldk);
    kappa[idx] = (m < M && j < N) ? rbf(sq + (size_t)m * I, s + (size_t)j * I, I, sigma_f, 0.0, 0.0) : 0.0;
}

// partial[rb][t][m] over the rows of block rb:  t = q00, (q01, q11,) mu0 (, mu1)
template <int O>
__global__ void __launch_bounds__(PG_T) k_pg_reduce(const double* V0, const double* V1, const double* w, int n, int Ns, int ldv,
                                                    int rows_per_block, double* partial) {
    constexpr int NT = O * (O + 1) / 2 + O;
    const int m = blockIdx.x * PG_T + threadIdx.x, rb = blockIdx.y;
    if (m >= ldv) return;
    const int r0 = rb * rows_per_block, r1 = min(n, r0 + rows_per_block);
    double q00 = 0.0, q01 = 0.0, q11 = 0.0, mu0 = 0.0, mu1 = 0.0;
    for (int r = r0; r < r1; ++r) {
        const double v0 = V0[(size_t)r * ldv + m], wr = w[r];
        q00 = fma(v0, v0, q00); mu0 = fma(v0, wr, mu0);
        if (O == 2 && r >= Ns) {
            const double v1 = V1[(size_t)(r - Ns) * ldv + m];
            q01 = fma(v0, v1, q01); q11 = fma(v1, v1, q11); mu1 = fma(v1, wr, mu1);
        }
    }
    double* out = partial + (size_t)rb * NT * ldv + m;
    if (O == 1) { out[0] = q00; out[ldv] = mu0; }
    else { out[0] = q00; out[ldv] = q01; out[2 * (size_t)ldv] = q11; out[3 * (size_t)ldv] = mu0; out[4 * (size_t)ldv] = mu1; }
}

template <int O>
__global__ void __launch_bounds__(PG_T) k_pg_finish(const double* partial, int nrb, int ldv, int M, int m_off, int M_total, double alpha_kmp,
                                                    double* xi, double* sigma_out) {
    constexpr int NT = O * (O + 1) / 2 + O, NPAIR = O * (O + 1) / 2;
    const int m = blockIdx.x * PG_T + threadIdx.x;
    if (m >= M) return;
    for (int t = 0; t < NT; ++t) {
        double v = 0.0;
        for (int rb = 0; rb < nrb; ++rb) v += partial[((size_t)rb * NT + t) * ldv + m];      // fixed order
        if (t >= NPAIR) xi[(size_t)(t - NPAIR) * M_total + m_off + m] = v;
        else if (sigma_out) {
            const int a = (O == 2 && t == 2) ? 1 : 0, bb = (O == 2 && t >= 1) ? 1 : 0;
            const double val = alpha_kmp * ((a == bb ? 1.0 : 0.0) - v);
            sigma_out[((size_t)a * O + bb) * M_total + m_off + m] = val;
            if (a != bb) sigma_out[((size_t)bb * O + a) * M_total + m_off + m] = val;
        }
    }
}

static long long pg_ws_doubles(int N, int O, int Mc) {
    const long long Ns = comp_stride(N), n = (long long)O * Ns, nt = O * (O + 1) / 2 + O;
    long long v = 0;
    for (int a = 0; a < O; ++a) v += (n - a * Ns) * Mc;
    return (long long)comp_stride(N) * Mc + v + (long long)PG_RB * nt * Mc;
}

}  // namespace kmpx

using namespace kmpx;

extern "C" long long kmpx_kmp_predict_gemm_workspace_bytes(int N, int O, int Mc) {
    return pg_ws_doubles(N, O, (int)round_up(Mc, PG_T)) * (long long)sizeof(double);
}

extern "C" int kmpx_kmp_predict_gemm(const double* F, int lda, const double* w, const double* s, int N, int I, int O,
                                     const double* sq, int M, double sigma_f, double alpha_kmp, double* ws, long long ws_bytes,
                                     double* xi, double* sigma_out, void* stream) {
    if (!F || !w || !s || !sq || !ws || !xi) return fail_arg(1, "null pointer");
    if (O < 1 || O > 2) return fail_arg(7, "the GEMM route supports O <= 2");
    if (N <= 0 || I <= 0 || M <= 0) return fail_arg(5, "sizes must be positive");
    if ((lda & 1) || (reinterpret_cast<uintptr_t>(F) & 15)) return fail_arg(2, "F must be 16-byte aligned with an even leading dimension");
    cudaStream_t st = (cudaStream_t)stream;
    const int Ns = comp_stride(N), n = O * Ns;
    // largest chunk of queries (multiple of 128) the workspace holds
    int Mc = (int)round_up(M, PG_T);
    while (Mc > PG_T && pg_ws_doubles(N, O, Mc) * (long long)sizeof(double) > ws_bytes) Mc -= PG_T;
    if (pg_ws_doubles(N, O, Mc) * (long long)sizeof(double) > ws_bytes) return fail_arg(13, "workspace too small (kmpx_kmp_predict_gemm_workspace_bytes)");
    double* kappa = ws;
    const int ldk = comp_stride(N);
    double* V[2] = {kappa + (size_t)ldk * Mc, nullptr};
    if (O == 2) V[1] = V[0] + (size_t)n * Mc;
    double* partial = (O == 2 ? V[1] + (size_t)Ns * Mc : V[0] + (size_t)n * Mc);
    const int rows_per_block = ceil_div(n, PG_RB);
    for (int m_off = 0; m_off < M; m_off += Mc) {
        const int mc = M - m_off < Mc ? M - m_off : Mc;
        k_pg_kappa<<<ceil_div((long long)ldk * Mc, 256), 256, 0, st>>>(s, N, I, sq + (size_t)m_off * I, mc, Mc, ldk, sigma_f, kappa);
        KMPX_TRY(check_launch("k_pg_kappa"));
        for (int a = 0; a < O; ++a) {
            GemmArgs g{n - a * Ns, Mc, N, 1.0, F + (size_t)a * Ns * lda + (size_t)a * Ns, lda, 0, kappa, ldk, 0, 0.0, V[a], Mc, 0, 2};
            KMPX_TRY(launch_dgemm(0, 1, g, 1, st));
        }
        dim3 grid(Mc / PG_T, PG_RB);
        if (O == 1) k_pg_reduce<1><<<grid, PG_T, 0, st>>>(V[0], nullptr, w, n, Ns, Mc, rows_per_block, partial);
        else k_pg_reduce<2><<<grid, PG_T, 0, st>>>(V[0], V[1], w, n, Ns, Mc, rows_per_block, partial);
        KMPX_TRY(check_launch("k_pg_reduce"));
        if (O == 1) k_pg_finish<1><<<ceil_div(mc, PG_T), PG_T, 0, st>>>(partial, PG_RB, Mc, mc, m_off, M, alpha_kmp, xi, sigma_out);
        else k_pg_finish<2><<<ceil_div(mc, PG_T), PG_T, 0, st>>>(partial, PG_RB, Mc, mc, m_off, M, alpha_kmp, xi, sigma_out);
        KMPX_TRY(check_launch("k_pg_finish"));
    }
    return 0;
}
