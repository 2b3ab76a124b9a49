// K4: KMP predict  (KMP.py:160-194)
//
//   xi*(m)    = k*(m) A^{-1} Y                       sigma*(m) = alpha (k(s*,s*) - k*(m) A^{-1} k*(m)^T)
//
// With A = L L^T factored in the component-major order (row = a*Ns + i) and F = L^{-1}:
//   V_a = F[:, class-a columns] * kappa          (n x M, per output a; "kappa" = coefficient matrix N x M)
//   k* A^{-1} k*^T [a][b] = sum_r V_a[r] V_b[r]     k* A^{-1} Y [a] = sum_r V_a[r] w[r],  w = F y'
// V is a triangular GEMM (F lower) on DMMA whose result is never stored: each CTA owns MT queries of one
// problem, walks the row tiles of F, and reduces the V tile over its rows straight away.
//
//   plain kernel (C=1): the B operand kappa[m][j] (MT x N) is computed once per CTA and stays in shared memory
//   otherwise         : the 8 x MT B tile is generated inside the cp.async ring, one exp per element
//
// factor_kind 1 (general, non-symmetric A): F = A^{-1} (full), U_b = F k*_b^T and the
// reduction uses k*_a[r] U_b[r]; w = alpha' = F y'.
//
// Algorithmic flops: O^2 N^2 M (triangular) for the symmetric path; bytes: F is streamed once per CTA from
// L2/HBM (4 n^2 bytes, lower half), outputs 8 (O + O^2) M bytes.
#include "dmma_tile.cuh"
#include "kmpx_internal.cuh"

namespace kmpx {

constexpr int PBM = 64;        // rows of F per tile
constexpr int PBK = 8;
constexpr int PLDT = 12;
constexpr int PSTAGES = 3;
constexpr int PTHREADS = 256;
constexpr int PMAXSEG = 16;    // O*C <= 16

template <int MT>
struct PredictSmem {
    // all offsets in doubles
    static __host__ __device__ size_t vbuf(int O) { return (size_t)O * PBM * (MT + 4); }
    static __host__ __device__ size_t stageA() { return (size_t)PBM * PLDT; }
    static __host__ __device__ size_t stageB() { return (size_t)MT * PLDT; }
    static __host__ __device__ size_t total(int O, int I, int resident, int ldp) {
        const int pairs = O * O + O;
        size_t d = vbuf(O) + PSTAGES * (stageA() + (resident ? 0 : stageB())) + (size_t)pairs * MT + PBM + (size_t)MT * I;
        if (resident) d += (size_t)MT * ldp;
        return d * sizeof(double);
    }
};

__device__ __forceinline__ double coef_at(const double* sq, const double* si, int I, double sigma_f, bool td, int ca, int cb) {
    if (!td) return rbf(sq, si, I, sigma_f, 0.0, 0.0);
    double co[4];
    kernel_coeffs(sq, si, I, sigma_f, true, co);
    return co[ca * 2 + cb];
}

template <int MT>
__global__ void __launch_bounds__(PTHREADS) k_kmp_predict(PredictArgs p) {
    constexpr int NI = MT / 32;              // warp tile 32 rows x MT/4 queries; 8 warps = 2 (rows) x 4 (queries)
    constexpr int LDV = MT + 4;
    extern __shared__ __align__(16) double smem[];
    __shared__ int seg_start[PMAXSEG + 1], seg_col0[PMAXSEG], seg_jmax[PMAXSEG];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int wr = warp >> 2, wc = warp & 3;         // warp row (0..1), warp column (0..3)
    const int b = blockIdx.y;
    const int m0 = blockIdx.x * MT;
    const int O = p.O, I = p.I;
    const bool td = p.time_driven != 0;
    const int C = td ? 2 : 1, P = O / C;
    const int N = p.n_pts ? p.n_pts[b] : p.n_max;
    const int Ns = comp_stride(N);
    const int n = O * Ns;
    const int npair = p.general ? O * O : O * (O + 1) / 2;
    const int ntask = npair + O;

    double* Vbuf = smem;
    double* stages = Vbuf + PredictSmem<MT>::vbuf(O);
    const size_t stage_elems = PredictSmem<MT>::stageA() + (p.resident ? 0 : PredictSmem<MT>::stageB());
    double* racc = stages + PSTAGES * stage_elems;          // [ntask][MT]
    double* wbuf = racc + (size_t)(O * O + O) * MT;           // [PBM]
    double* sqs = wbuf + PBM;                                 // [MT][I]
    double* panel = sqs + (size_t)MT * I;                     // [MT][ldp]  (resident only)

    const double* F = p.F + (size_t)b * p.strideA;
    const double* s = p.s + (size_t)b * p.n_max * I;
    const double* sq = p.sq + (size_t)b * p.stride_sq;
    const double* w = p.w + (size_t)b * p.ldw;

    for (int idx = tid; idx < MT * I; idx += PTHREADS) {
        const int m = idx / I;
        sqs[idx] = (m0 + m < p.M) ? sq[(size_t)(m0 + m) * I + (idx - m * I)] : 0.0;
    }
    for (int idx = tid; idx < ntask * MT; idx += PTHREADS) racc[idx] = 0.0;
    __syncthreads();
    if (p.resident) {
        for (int idx = tid; idx < MT * p.ldp; idx += PTHREADS) {
            const int m = idx / p.ldp, j = idx - m * p.ldp;
            double v = 0.0;
            if (j < N && m0 + m < p.M) v = rbf(sqs + m * I, s + (size_t)j * I, I, p.sigma_f, 0.0, 0.0);
            panel[idx] = v;
        }
    }
    __syncthreads();

    const int ntiles = (n + PBM - 1) / PBM;
    for (int rt = 0; rt < ntiles; ++rt) {
        const int r0 = rt * PBM;
        const int rows_valid = min(PBM, n - r0);
        if (tid == 0) {
            int tot = 0;
            for (int a = 0; a < O; ++a) {
                for (int cb = 0; cb < C; ++cb) {
                    const int comp = cb * P + (a % P);
                    const int col0 = comp * Ns;
                    int jmax = p.general ? N : (r0 + PBM - col0);
                    jmax = jmax < 0 ? 0 : (jmax > N ? N : jmax);
                    const int sg = a * C + cb;
                    seg_start[sg] = tot; seg_col0[sg] = col0; seg_jmax[sg] = jmax;
                    tot += (jmax + PBK - 1) / PBK;
                }
            }
            seg_start[O * C] = tot;
        }
        if (tid < PBM) wbuf[tid] = (tid < rows_valid) ? w[r0 + tid] : 0.0;
        __syncthreads();
        const int nseg = O * C;
        const int W = seg_start[nseg];

        auto locate = [&](int wi, int& sg, int& jb) {
            sg = 0;
            while (sg + 1 < nseg && seg_start[sg + 1] <= wi) ++sg;
            jb = wi - seg_start[sg];
        };
        auto issue = [&](int wi) {
            int sg, jb;
            locate(wi, sg, jb);
            double* st = stages + (wi % PSTAGES) * stage_elems;
            const int col_g0 = seg_col0[sg] + jb * PBK;
            const int col_end = seg_col0[sg] + seg_jmax[sg];
            const double* gsrc = F + (size_t)r0 * p.lda + col_g0;
            if (p.general)
                load_tile_async<PBK, PLDT, false>(st, gsrc, p.lda, PBM, rows_valid, r0, col_g0, col_end, tid, PTHREADS);
            else
                load_tile_async<PBK, PLDT, true>(st, gsrc, p.lda, PBM, rows_valid, r0, col_g0, col_end, tid, PTHREADS);
            if (!p.resident) {
                double* bt = st + PredictSmem<MT>::stageA();
                const int a = sg / C, cb = sg - a * C, ca = a / P;
                for (int idx = tid; idx < MT * PBK; idx += PTHREADS) {
                    const int m = idx / PBK, kk = idx - m * PBK;
                    const int j = jb * PBK + kk;
                    double v = 0.0;
                    if (j < N && m0 + m < p.M) v = coef_at(sqs + m * I, s + (size_t)j * I, I, p.sigma_f, td, ca, cb);
                    bt[m * PLDT + kk] = v;
                }
            }
        };

        double acc[4][NI][2];
        zero_acc(acc);
#pragma unroll
        for (int sidx = 0; sidx < PSTAGES - 1; ++sidx) { if (sidx < W) issue(sidx); cp_async_commit(); }
        int prev_a = -1;
        for (int wi = 0; wi < W; ++wi) {
            cp_async_wait<PSTAGES - 2>();
            __syncthreads();
            if (wi + PSTAGES - 1 < W) issue(wi + PSTAGES - 1);
            cp_async_commit();
            int sg, jb;
            locate(wi, sg, jb);
            const int a = sg / C;
            if (a != prev_a) { zero_acc(acc); prev_a = a; }
            const double* tile = stages + (wi % PSTAGES) * stage_elems;
            const double* Bs;
            int ldb;
            if (p.resident) { Bs = panel + (size_t)(wc * (MT / 4)) * p.ldp + jb * PBK; ldb = p.ldp; }
            else { Bs = tile + PredictSmem<MT>::stageA() + (wc * (MT / 4)) * PLDT; ldb = PLDT; }
#pragma unroll
            for (int kk = 0; kk < PBK; kk += 4)
                warp_mma_k4<4, NI, false>(acc, tile + (wr * 32) * PLDT + kk, PLDT, Bs + kk, ldb, lane);
            // last item of this output a?  -> flush the V_a tile
            bool last = (wi + 1 == W);
            if (!last) { int sg2, jb2; locate(wi + 1, sg2, jb2); last = (sg2 / C != a); }
            if (last) {
                double* V = Vbuf + (size_t)a * PBM * LDV;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < NI; ++j) {
                        const int r = wr * 32 + i * 8 + g, c = wc * (MT / 4) + j * 8 + 2 * t4;
                        V[r * LDV + c] = acc[i][j][0];
                        V[r * LDV + c + 1] = acc[i][j][1];
                    }
            }
        }
        cp_async_wait<0>();
        __syncthreads();
        // outputs a that had no work in this row tile: V_a = 0
        for (int a = 0; a < O; ++a) {
            if (seg_start[(a + 1) * C] == seg_start[a * C]) {
                double* V = Vbuf + (size_t)a * PBM * LDV;
                for (int idx = tid; idx < PBM * LDV; idx += PTHREADS) V[idx] = 0.0;
            }
        }
        __syncthreads();
        // reduce the V tile over its rows
        for (int item = tid; item < ntask * MT; item += PTHREADS) {
            const int task = item / MT, m = item - task * MT;
            double sum = 0.0;
            if (!p.general) {
                if (task < npair) {
                    int a = 0, rem = task;                 // task -> (a <= bb), row-major over the upper triangle
                    while (rem >= O - a) { rem -= O - a; ++a; }
                    const int bb = a + rem;
                    const double* Va = Vbuf + (size_t)a * PBM * LDV + m;
                    const double* Vb = Vbuf + (size_t)bb * PBM * LDV + m;
#pragma unroll 8
                    for (int r = 0; r < PBM; ++r) sum += Va[r * LDV] * Vb[r * LDV];
                } else {
                    const double* Va = Vbuf + (size_t)(task - npair) * PBM * LDV + m;
#pragma unroll 8
                    for (int r = 0; r < PBM; ++r) sum += Va[r * LDV] * wbuf[r];
                }
            } else {
                // general: sum_r k*_a[r](m) * X[r],  X = U_b tile (covariance) or alpha' (mean)
                const bool is_mean = task >= npair;
                const int a = is_mean ? task - npair : task / O;
                const int bb = is_mean ? 0 : task - a * O;
                const double* Ub = Vbuf + (size_t)bb * PBM * LDV + m;
                if (m0 + m < p.M) {
                    for (int r = 0; r < rows_valid; ++r) {
                        const int gr = r0 + r;
                        const int a2 = gr / Ns, i = gr - a2 * Ns;
                        if (i >= N || (a2 % P) != (a % P)) continue;
                        const double kv = coef_at(sqs + m * I, s + (size_t)i * I, I, p.sigma_f, td, a / P, a2 / P);
                        sum += kv * (is_mean ? wbuf[r] : Ub[r * LDV]);
                    }
                }
            }
            racc[item] += sum;
        }
        __syncthreads();
    }

    // outputs, reference layout: xi [O][M], sigma [O][O][M]
    double* xi = p.xi + (size_t)b * O * p.M;
    double* sg_out = p.sigma_out ? p.sigma_out + (size_t)b * O * O * p.M : nullptr;
    for (int idx = tid; idx < O * MT; idx += PTHREADS) {
        const int a = idx / MT, m = idx - a * MT;
        if (m0 + m < p.M) xi[(size_t)a * p.M + m0 + m] = racc[(npair + a) * MT + m];
    }
    if (sg_out) {
        for (int idx = tid; idx < O * O * MT; idx += PTHREADS) {
            const int ab = idx / MT, m = idx - ab * MT;
            const int a = ab / O, bb = ab - a * O;
            if (m0 + m >= p.M) continue;
            double quad;
            if (p.general) quad = racc[(a * O + bb) * MT + m];
            else {
                const int lo = a < bb ? a : bb, hi = a < bb ? bb : a;
                const int task = lo * O - lo * (lo - 1) / 2 + (hi - lo);
                quad = racc[task * MT + m];
            }
            double kss = 0.0;
            if (!td) kss = (a == bb) ? 1.0 : 0.0;      // exp(-sigma_f*0) = 1 exactly (KMP.py:189)
            else if ((a % P) == (bb % P)) kss = coef_at(sqs + m * I, sqs + m * I, I, p.sigma_f, true, a / P, bb / P);
            sg_out[((size_t)a * O + bb) * p.M + m0 + m] = p.alpha_kmp * (kss - quad);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// mean-only predict from alpha:  xi[a][m] = sum_i sum_cb coef[a/P][cb](s*_m, s_i) alpha[i][cb*P + a%P]
struct MeanArgs {
    const double* alpha; const double* s; const int* n_pts; int n_max, I, O, batch;
    const double* sq; long long stride_sq; int M; double sigma_f; int time_driven; double* xi;
};
constexpr int MEAN_CHUNK = 256;

__global__ void __launch_bounds__(128) k_kmp_predict_mean(MeanArgs p) {
    __shared__ double s_sh[MEAN_CHUNK * 4];
    __shared__ double a_sh[MEAN_CHUNK * 8];
    const int b = blockIdx.y, m = blockIdx.x * blockDim.x + threadIdx.x;
    const int O = p.O, I = p.I;
    const bool td = p.time_driven != 0;
    const int C = td ? 2 : 1, P = O / C;
    const int N = p.n_pts ? p.n_pts[b] : p.n_max;
    const double* s = p.s + (size_t)b * p.n_max * I;
    const double* al = p.alpha + (size_t)b * p.n_max * O;
    const double* sq = p.sq + (size_t)b * p.stride_sq;
    double q[4];
    for (int t = 0; t < I; ++t) q[t] = (m < p.M) ? sq[(size_t)m * I + t] : 0.0;
    double acc[8];
    for (int a = 0; a < O; ++a) acc[a] = 0.0;
    for (int i0 = 0; i0 < N; i0 += MEAN_CHUNK) {
        const int cnt = min(MEAN_CHUNK, N - i0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < cnt * I; idx += blockDim.x) s_sh[idx] = s[(size_t)i0 * I + idx];
        for (int idx = threadIdx.x; idx < cnt * O; idx += blockDim.x) a_sh[idx] = al[(size_t)i0 * O + idx];
        __syncthreads();
        if (m < p.M) {
            for (int i = 0; i < cnt; ++i) {
                double co[4];
                kernel_coeffs(q, s_sh + i * I, I, p.sigma_f, td, co);
                for (int a = 0; a < O; ++a) {
                    const int ca = a / P, pa = a - ca * P;
                    double v = co[ca * 2] * a_sh[i * O + pa];
                    if (td) v += co[ca * 2 + 1] * a_sh[i * O + P + pa];
                    acc[a] += v;
                }
            }
        }
    }
    if (m < p.M) for (int a = 0; a < O; ++a) p.xi[((size_t)b * O + a) * p.M + m] = acc[a];
}

// Plain kernel: CTA = 32 queries (lane = query), eight warps split the stored points of every staged chunk; the
// coefficients are four interleaved branch-free exponentials per lane (the loop above evaluates one library exp at a
// time with one thread per query: 19.6 G exp/s at cfg 4); the eight partial sums are added in warp order (deterministic).
constexpr int MEANP_WARPS = 8;
__global__ void __launch_bounds__(MEANP_WARPS * 32) k_kmp_predict_mean_plain(MeanArgs p) {
    __shared__ double s_sh[MEAN_CHUNK * 4];
    __shared__ double a_sh[MEAN_CHUNK * 8];
    __shared__ double red[MEANP_WARPS][8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, m = blockIdx.x * 32 + lane;
    const int O = p.O, I = p.I;
    const int N = p.n_pts ? p.n_pts[b] : p.n_max;
    const double* s = p.s + (size_t)b * p.n_max * I;
    const double* al = p.alpha + (size_t)b * p.n_max * O;
    const double* sq = p.sq + (size_t)b * p.stride_sq;
    double q[4];
    for (int t = 0; t < I; ++t) q[t] = sq[(size_t)min(m, p.M - 1) * I + t];
    double acc[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) acc[a] = 0.0;
    constexpr int PER = MEAN_CHUNK / MEANP_WARPS;       // 32 points per warp and chunk
    for (int i0 = 0; i0 < N; i0 += MEAN_CHUNK) {
        const int cnt = min(MEAN_CHUNK, N - i0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < MEAN_CHUNK * I; idx += blockDim.x) s_sh[idx] = idx < cnt * I ? s[(size_t)i0 * I + idx] : 0.0;
        for (int idx = threadIdx.x; idx < MEAN_CHUNK * O; idx += blockDim.x) a_sh[idx] = idx < cnt * O ? al[(size_t)i0 * O + idx] : 0.0;   // alpha = 0 beyond N
        __syncthreads();
        if (warp * PER >= cnt) continue;
        for (int i = warp * PER; i < (warp + 1) * PER; i += 4) {
            double e[4];
            if (I == 1) {
#pragma unroll
                for (int u = 0; u < 4; ++u) { const double d = fabs(q[0] - s_sh[i + u]); e[u] = exp_nonpos(-p.sigma_f * (d * d)); }
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) e[u] = rbf_fast(q, s_sh + (i + u) * I, I, p.sigma_f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int a = 0; a < 8; ++a) if (a < O) acc[a] = fma(e[u], a_sh[(i + u) * O + a], acc[a]);
        }
    }
#pragma unroll
    for (int a = 0; a < 8; ++a) if (a < O) red[warp][a][lane] = acc[a];
    __syncthreads();
    if (warp == 0 && m < p.M) {
        for (int a = 0; a < O; ++a) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < MEANP_WARPS; ++w) v += red[w][a][lane];
            p.xi[((size_t)b * O + a) * p.M + m] = v;
        }
    }
}

template <int MT>
static int launch_predict(PredictArgs& p, cudaStream_t st) {
    // resident coefficient panel only for the plain kernel and when it fits
    const int ldp = (int)round_up(p.n_max, 16) + 4;
    size_t sm_res = PredictSmem<MT>::total(p.O, p.I, 1, ldp);
    p.resident = (!p.time_driven && sm_res <= 220 * 1024) ? 1 : 0;
    p.ldp = ldp;
    const size_t smem = PredictSmem<MT>::total(p.O, p.I, p.resident, ldp);
    if (smem > 226 * 1024) { snprintf(g_err, sizeof(g_err), "kmpx: predict tile does not fit shared memory (O=%d)", p.O); return -11; }
    static SmemCache smem_set;
    KMPX_TRY(ensure_dyn_smem(k_kmp_predict<MT>, smem, smem_set, "k_kmp_predict"));
    dim3 grid(ceil_div(p.M, MT), p.batch);
    k_kmp_predict<MT><<<grid, PTHREADS, smem, st>>>(p);
    return check_launch("k_kmp_predict");
}

}  // namespace kmpx

using namespace kmpx;

static int predict_dispatch(const double* F, int lda, long long strideA, int factor_kind,
                            const double* w, int ldw, const double* s, const int* n_pts, int n_max, int I, int O,
                            int batch, const double* sq, long long stride_sq, int M,
                            double sigma_f, double alpha_kmp, int time_driven,
                            double* xi, double* sigma_out, int persistent, void* stream) {
    if (!F) return fail_arg(1, "F is NULL");
    if (lda & 1) return fail_arg(2, "lda must be even");
    if (strideA & 1) return fail_arg(3, "strideA must be even");
    if (factor_kind != 0 && factor_kind != 1) return fail_arg(4, "factor_kind must be 0 or 1");
    if (!w) return fail_arg(5, "w is NULL");
    if (!s) return fail_arg(7, "s is NULL");
    if (n_max <= 0) return fail_arg(9, "n_max must be > 0");
    if (I <= 0 || I > 4) return fail_arg(10, "I must be in 1..4");
    if (O <= 0 || O > 8) return fail_arg(11, "O must be in 1..8");
    if (time_driven && (O & 1)) return fail_arg(11, "time-driven kernel needs an even O");
    if (batch <= 0 || batch > 65535) return fail_arg(12, "batch must be in 1..65535");
    if (!sq) return fail_arg(13, "sq is NULL");
    if (M <= 0) return fail_arg(15, "M must be > 0");
    if (!xi) return fail_arg(19, "xi is NULL");
    if ((reinterpret_cast<uintptr_t>(F) & 15) != 0) return fail_arg(1, "F must be 16-byte aligned");
    PredictArgs p{F, lda, strideA, factor_kind, w, ldw, s, n_pts, n_max, I, O, batch, sq, stride_sq, M,
                  sigma_f, alpha_kmp, time_driven, xi, sigma_out, 0, 0, g_host_dbg, g_predict_kskip};
    if (persistent && predict_pers_applicable(p)) return launch_predict_pers(p, (cudaStream_t)stream);
    if (predict_tab_applicable(p)) return launch_predict_tab(p, (cudaStream_t)stream);
    if (predict_fast_applicable(p)) return launch_predict_fast(p, (cudaStream_t)stream);
    if (predict_stream_applicable(p)) return launch_predict_stream(p, (cudaStream_t)stream);
    if (O <= 2) return launch_predict<64>(p, (cudaStream_t)stream);
    return launch_predict<32>(p, (cudaStream_t)stream);
}

extern "C" int kmpx_predict_skip_threshold(double threshold) {
    if (!(threshold >= 0.0) || threshold > 1e-12) return fail_arg(1, "threshold must be in [0, 1e-12]");
    g_predict_kskip = threshold;
    return 0;
}

extern "C" int kmpx_predict_tile_queries(int queries) {
    if (queries != 0 && queries != 32 && queries != 64) return fail_arg(1, "queries per CTA must be 0 (automatic), 32 or 64");
    g_predict_tile_queries = queries;
    return 0;
}

extern "C" int kmpx_kmp_predict(const double* F, int lda, long long strideA, int factor_kind,
                                const double* w, int ldw, const double* s, const int* n_pts, int n_max, int I, int O,
                                int batch, const double* sq, long long stride_sq, int M,
                                double sigma_f, double alpha_kmp, int time_driven,
                                double* xi, double* sigma_out, void* stream) {
    return predict_dispatch(F, lda, strideA, factor_kind, w, ldw, s, n_pts, n_max, I, O, batch, sq, stride_sq, M, sigma_f, alpha_kmp,
                            time_driven, xi, sigma_out, 0, stream);
}
extern "C" int kmpx_kmp_predict_persistent(const double* F, int lda, long long strideA, int factor_kind,
                                           const double* w, int ldw, const double* s, const int* n_pts, int n_max, int I, int O,
                                           int batch, const double* sq, long long stride_sq, int M,
                                           double sigma_f, double alpha_kmp, int time_driven,
                                           double* xi, double* sigma_out, void* stream) {
    return predict_dispatch(F, lda, strideA, factor_kind, w, ldw, s, n_pts, n_max, I, O, batch, sq, stride_sq, M, sigma_f, alpha_kmp,
                            time_driven, xi, sigma_out, 1, stream);
}

extern "C" int kmpx_kmp_predict_mean(const double* alpha, const double* s, const int* n_pts, int n_max, int I, int O,
                                     int batch, const double* sq, long long stride_sq, int M,
                                     double sigma_f, int time_driven, double* xi, void* stream) {
    if (!alpha) return fail_arg(1, "alpha is NULL");
    if (!s) return fail_arg(2, "s is NULL");
    if (I <= 0 || I > 4) return fail_arg(5, "I must be in 1..4");
    if (O <= 0 || O > 8) return fail_arg(6, "O must be in 1..8");
    if (batch <= 0 || batch > 65535) return fail_arg(7, "batch must be in 1..65535");
    if (!sq) return fail_arg(8, "sq is NULL");
    if (M <= 0) return fail_arg(10, "M must be > 0");
    if (!xi) return fail_arg(13, "xi is NULL");
    MeanArgs p{alpha, s, n_pts, n_max, I, O, batch, sq, stride_sq, M, sigma_f, time_driven, xi};
    if (!time_driven) {
        dim3 gridp(ceil_div(M, 32), batch);
        k_kmp_predict_mean_plain<<<gridp, MEANP_WARPS * 32, 0, (cudaStream_t)stream>>>(p);
        return check_launch("k_kmp_predict_mean_plain");
    }
    dim3 grid(ceil_div(M, 128), batch);
    k_kmp_predict_mean<<<grid, 128, 0, (cudaStream_t)stream>>>(p);
    return check_launch("k_kmp_predict_mean");
}
