// K4 for large systems (plain RBF kernel, O <= 2, symmetric factor; cfg 4 of BASELINE.json: N = 8192, n = 16384,
// 100 000 queries).  The 64 x N kernel panel of predict_fast.cu no longer fits shared memory, so the B operand is
// generated inside the pipeline instead:
//   * a CTA owns 64 queries and walks ALL row tiles (128 rows) of Linv; CTAs advance through Linv roughly in step,
//     so the 1 GB factor is served from L2 most of the time;
//   * per k-block of 16 stored points the ring stage holds the Linv tile of output 0 (columns j), the Linv tile of
//     output 1 (columns Ns + j) and ONE 64 x 16 kappa tile shared by both outputs (one branch-free exp per element, four
//     interleaved chains per thread, evaluated two stages ahead and stored after the MMAs of the current stage);
//   * warp tile 16 rows x 64 queries x both outputs, Gram / mean reduction from the accumulator registers exactly
//     as in predict_fast.cu; fixed reduction order => bitwise reproducible.
// Algorithmic flops: O^2 Ns^2 M; exp evaluations: N * M * (n / 128) (about 13% of the FP64 work at O = 2).
#include "dmma_tile.cuh"
#include "kmpx_internal.cuh"
#include <type_traits>

namespace kmpx {

constexpr int SBM = 128;        // rows of Linv per tile (8 warps x 16)
constexpr int SQT = 64;         // queries per CTA
constexpr int SBK = 16;         // stored points per pipeline stage
constexpr int SLD = 20;         // tile leading dimension (4 mod 16)
constexpr int SST = 3;          // pipeline stages
constexpr int STHREADS = 256;
constexpr int SWARPS = 8;

template <int O>
__host__ __device__ inline size_t stream_smem_bytes(int I) {
    const int nt = O * (O + 1) / 2 + O;
    size_t d = (size_t)SST * (O * SBM * SLD + SQT * SLD) + (size_t)SWARPS * nt * SQT + SBM + (size_t)SQT * I;
    return d * sizeof(double);
}

template <int O>
__global__ void __launch_bounds__(STHREADS, 1) k_kmp_predict_stream(PredictArgs p) {
    constexpr int NT = O * (O + 1) / 2 + O;
    constexpr int STAGE = O * SBM * SLD + SQT * SLD;      // doubles per stage: A tiles (per output) then the kappa tile
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int b = blockIdx.y, m0 = blockIdx.x * SQT;
    const int I = p.I;
    const int N = p.n_pts ? p.n_pts[b] : p.n_max;
    const int Ns = comp_stride(N);
    const int n = O * Ns;
    const bool one_dim = I == 1;

    double* stages = smem;
    double* racc = stages + (size_t)SST * STAGE;              // [SWARPS][NT][SQT]
    double* wsh = racc + (size_t)SWARPS * NT * SQT;            // [SBM]
    double* sqs = wsh + SBM;                                   // [SQT][I]
    double* my_racc = racc + (size_t)warp * NT * SQT;

    const double* F = p.F + (size_t)b * p.strideA;
    const double* s = p.s + (size_t)b * p.n_max * I;
    const double* sq = p.sq + (size_t)b * p.stride_sq;
    const double* w = p.w + (size_t)b * p.ldw;

    for (int idx = tid; idx < SQT * I; idx += STHREADS) {
        const int m = idx / I;
        sqs[idx] = (m0 + m < p.M) ? sq[(size_t)(m0 + m) * I + (idx - m * I)] : 0.0;
    }
    for (int idx = tid; idx < SWARPS * NT * SQT; idx += STHREADS) racc[idx] = 0.0;
    __syncthreads();

    const int ntiles = (n + SBM - 1) / SBM;
    for (int rt = 0; rt < ntiles; ++rt) {
        const int r0 = rt * SBM;
        const int rows_valid = min(SBM, n - r0);
        // stored points needed by this row tile: output 0 -> columns j <= row, output 1 -> columns Ns + j <= row
        int jmax0 = r0 + SBM; jmax0 = jmax0 > N ? N : jmax0;
        int jmax1 = 0;
        if (O == 2) { jmax1 = r0 + SBM - Ns; jmax1 = jmax1 < 0 ? 0 : (jmax1 > N ? N : jmax1); }
        const int nk = (jmax0 + SBK - 1) / SBK;
        if (tid < SBM) wsh[tid] = (tid < rows_valid) ? w[r0 + tid] : 0.0;

        auto issue = [&](int kt) {
            double* st = stages + (size_t)(kt % SST) * STAGE;
            const int j0 = kt * SBK;
            load_tile_async<SBK, SLD, true>(st, F + (size_t)r0 * p.lda + j0, p.lda, SBM, rows_valid, r0, j0, jmax0, tid, STHREADS);
            if (O == 2 && j0 < jmax1)
                load_tile_async<SBK, SLD, true>(st + SBM * SLD, F + (size_t)r0 * p.lda + Ns + j0, p.lda, SBM, rows_valid, r0,
                                                Ns + j0, Ns + jmax1, tid, STHREADS);
        };
        // kappa tile of stage kt: SQT * SBK / STHREADS = 4 coefficients per thread, evaluated as four interleaved branch-free
        // chains into registers (the library exp serialises them and its latency would sit in front of the DMMAs of
        // the same warp); they are stored after the MMAs of the current stage
        constexpr int KPT = SQT * SBK / STHREADS;
        auto kappa_eval_t = [&](int kt, double (&kv)[KPT], auto od) {       // the I == 1 test stays outside the chains
            constexpr bool OD = decltype(od)::value;
            const int j0 = kt * SBK;
#pragma unroll
            for (int u = 0; u < KPT; ++u) {
                const int idx = tid + u * STHREADS;
                const int m = idx / SBK, j = j0 + (idx - m * SBK);
                const double* sp = s + (size_t)min(j, N - 1) * I;
                double e;
                if (OD) { const double d = fabs(sqs[m] - sp[0]); e = exp_nonpos(-p.sigma_f * (d * d)); }
                else e = rbf_fast(sqs + m * I, sp, I, p.sigma_f);
                kv[u] = (j < N && m0 + m < p.M) ? e : 0.0;
            }
        };
        auto kappa_eval = [&](int kt, double (&kv)[KPT]) {
            if (one_dim) kappa_eval_t(kt, kv, std::true_type{}); else kappa_eval_t(kt, kv, std::false_type{});
        };
        auto kappa_store = [&](int kt, const double (&kv)[KPT]) {
            double* bt = stages + (size_t)(kt % SST) * STAGE + O * SBM * SLD;
#pragma unroll
            for (int u = 0; u < KPT; ++u) {
                const int idx = tid + u * STHREADS;
                const int m = idx / SBK;
                bt[m * SLD + (idx - m * SBK)] = kv[u];
            }
        };

        double acc0[2][8][2], acc1[2][8][2];
        zero_acc(acc0);
        zero_acc(acc1);
#pragma unroll
        for (int st = 0; st < SST - 1; ++st) {
            if (st < nk) { issue(st); double kv[KPT]; kappa_eval(st, kv); kappa_store(st, kv); }
            cp_async_commit();
        }
        for (int kt = 0; kt < nk; ++kt) {
            cp_async_wait<SST - 2>();
            __syncthreads();
            const bool more = kt + SST - 1 < nk;
            double kv[KPT];
            if (more) { issue(kt + SST - 1); kappa_eval(kt + SST - 1, kv); }
            cp_async_commit();
            const double* st = stages + (size_t)(kt % SST) * STAGE;
            const double* bt = st + O * SBM * SLD;
#pragma unroll
            for (int kk = 0; kk < SBK; kk += 4) warp_mma_k4<2, 8, false>(acc0, st + (warp * 16) * SLD + kk, SLD, bt + kk, SLD, lane);
            if (O == 2 && kt * SBK < jmax1) {
#pragma unroll
                for (int kk = 0; kk < SBK; kk += 4)
                    warp_mma_k4<2, 8, false>(acc1, st + SBM * SLD + (warp * 16) * SLD + kk, SLD, bt + kk, SLD, lane);
            }
            if (more) kappa_store(kt + SST - 1, kv);
        }
        cp_async_wait<0>();
        // reduce this warp's 16 rows of V straight from the accumulators
        double wv[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) wv[i] = wsh[warp * 16 + i * 8 + g];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            double red[NT][2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double q00 = 0.0, q01 = 0.0, q11 = 0.0, mu0 = 0.0, mu1 = 0.0;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const double v0 = acc0[i][j][e];
                    q00 += v0 * v0; mu0 += v0 * wv[i];
                    if (O == 2) { const double v1 = acc1[i][j][e]; q01 += v0 * v1; q11 += v1 * v1; mu1 += v1 * wv[i]; }
                }
                if (O == 1) { red[0][e] = q00; red[1][e] = mu0; }
                else { red[0][e] = q00; red[1][e] = q01; red[2][e] = q11; red[3][e] = mu0; red[NT - 1][e] = mu1; }
            }
#pragma unroll
            for (int t = 0; t < NT; ++t)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double v = red[t][e];
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    if (g == 0) my_racc[t * SQT + j * 8 + 2 * t4 + e] += v;
                }
        }
        __syncthreads();      // wsh and the ring are reused by the next row tile
    }

    double* xi = p.xi + (size_t)b * O * p.M;
    double* sg_out = p.sigma_out ? p.sigma_out + (size_t)b * O * O * p.M : nullptr;
    constexpr int NPAIR = O * (O + 1) / 2;
    for (int idx = tid; idx < NT * SQT; idx += STHREADS) {
        const int t = idx / SQT, m = idx - t * SQT;
        if (m0 + m >= p.M) continue;
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < SWARPS; ++q) v += racc[((size_t)q * NT + t) * SQT + m];
        if (t >= NPAIR) {
            xi[(size_t)(t - NPAIR) * p.M + m0 + m] = v;
        } else if (sg_out) {
            const int a = (O == 2 && t == 2) ? 1 : 0;
            const int bb = (O == 2 && t >= 1) ? 1 : 0;
            const double val = p.alpha_kmp * ((a == bb ? 1.0 : 0.0) - v);
            sg_out[((size_t)a * O + bb) * p.M + m0 + m] = val;
            if (a != bb) sg_out[((size_t)bb * O + a) * p.M + m0 + m] = val;
        }
    }
}

bool predict_stream_applicable(const PredictArgs& p) {
    return !p.time_driven && !p.general && p.O <= 2;
}

int launch_predict_stream(PredictArgs& p, cudaStream_t st) {
    dim3 grid(ceil_div(p.M, SQT), p.batch);
    if (p.O == 1) {
        const size_t smem = stream_smem_bytes<1>(p.I);
        static SmemCache set1;
        KMPX_TRY(ensure_dyn_smem(k_kmp_predict_stream<1>, smem, set1, "k_kmp_predict_stream<1>"));
        k_kmp_predict_stream<1><<<grid, STHREADS, smem, st>>>(p);
    } else {
        const size_t smem = stream_smem_bytes<2>(p.I);
        static SmemCache set2;
        KMPX_TRY(ensure_dyn_smem(k_kmp_predict_stream<2>, smem, set2, "k_kmp_predict_stream<2>"));
        k_kmp_predict_stream<2><<<grid, STHREADS, smem, st>>>(p);
    }
    return check_launch("k_kmp_predict_stream");
}

}  // namespace kmpx
