// K4 fast path: plain RBF kernel, O <= 2, symmetric factor (the cfg-1/3/4 shape of BASELINE.json).
//
// Same mathematics as predict.cu (V_a = Linv[:, block a] * kappa; Gram / mean reductions over the rows of V),
// organised so that the FP64 tensor pipe never waits on a CTA-wide barrier:
//   * a CTA owns 64 queries of one problem; kappa[m][j] (64 x N) is evaluated once and stays in shared memory;
//   * each of the 8 warps owns a static list of 16-row chunks of Linv (the triangular costs grow with the row
//     index, so the chunks are dealt from the bottom up in boustrophedon order) and streams them through a
//     PRIVATE 2-stage cp.async ring (32 columns per stage) - only __syncwarp inside the main loop;
//   * warp tile = 16 rows x 64 queries x both outputs: 16 independent DMMA.8x8x4 per k-step of 4 and output,
//     10 LDS.64 per 16 DMMA;
//   * the V tile is reduced straight from the accumulator registers (products, a 3-step reduce-scatter over the
//     8 row lanes) into a per-warp slice of shared memory; slices are summed in warp order at the end, so results are
//     bitwise reproducible and independent of the batch a problem sits in.
// Algorithmic flops per problem: O^2 Ns^2 M (+ lower-order reductions); F is streamed from L2 once per CTA.
#include "dmma_tile.cuh"
#include "kmpx_internal.cuh"

namespace kmpx {

constexpr int QT = 64;          // queries per CTA
constexpr int RC = 16;          // rows of Linv per chunk
constexpr int QK = 32;          // k per stage: 8 k-steps of 4 = 128 DMMA per warp between two ring hand-overs
constexpr int QLD = 36;         // ring tile leading dimension (4 mod 16)
constexpr int QST = 2;          // ring stages per warp (one tile in flight while one is consumed)
constexpr int QWARPS = 8;
constexpr int QTHREADS = QWARPS * 32;
constexpr int QMAXCH = 80;      // chunks per problem: n_sys <= 1280

template <int O> struct FastCfg { static constexpr int NT = O * (O + 1) / 2 + O; };

__host__ __device__ inline size_t fast_smem_bytes(int O, int I, int n_max_pts, int ldp) {
    const int nt = O * (O + 1) / 2 + O;
    const int n_sys_max = O * comp_stride(n_max_pts);
    size_t d = (size_t)QT * ldp + (size_t)QWARPS * QST * RC * QLD + (size_t)QWARPS * nt * QT + round_up(n_sys_max, 2) + (size_t)QT * I;
    return d * sizeof(double);
}

struct Cursor {   // position of a warp in its flattened list of (chunk, output a, k-block) items
    int ci, a, kb, nk, r0, col0, jmax;
    bool valid;
};

template <int O>
__global__ void __launch_bounds__(QTHREADS, 1) k_kmp_predict_fast(PredictArgs p) {
    constexpr int NT = FastCfg<O>::NT;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int b = blockIdx.y, m0 = blockIdx.x * QT;
    const int I = p.I;
    const int N = p.n_pts ? p.n_pts[b] : p.n_max;
    const int Ns = comp_stride(N);
    const int n = O * Ns;
    const int ldp = p.ldp;
    const int n_sys_max = O * comp_stride(p.n_max);

    double* panel = smem;                                               // [QT][ldp]
    double* rings = panel + (size_t)QT * ldp;                           // [QWARPS][QST][RC*QLD]
    double* racc = rings + (size_t)QWARPS * QST * RC * QLD;              // [QWARPS][NT][QT]
    double* wsh = racc + (size_t)QWARPS * NT * QT;                       // [n_sys_max]
    double* sqs = wsh + round_up(n_sys_max, 2);                          // [QT][I]

    const double* F = p.F + (size_t)b * p.strideA;
    const double* s = p.s + (size_t)b * p.n_max * I;
    const double* sq = p.sq + (size_t)b * p.stride_sq;
    const double* w = p.w + (size_t)b * p.ldw;

    for (int idx = tid; idx < QT * I; idx += QTHREADS) {
        const int m = idx / I;
        sqs[idx] = (m0 + m < p.M) ? sq[(size_t)(m0 + m) * I + (idx - m * I)] : 0.0;
    }
    for (int idx = tid; idx < n_sys_max; idx += QTHREADS) wsh[idx] = idx < n ? w[idx] : 0.0;
    for (int idx = tid; idx < QWARPS * NT * QT; idx += QTHREADS) racc[idx] = 0.0;
    const int nchunks = (n + RC - 1) / RC;
    __syncthreads();

    // ---------------------------------------------------------------- per-warp streaming loop
    double* ring = rings + (size_t)warp * QST * RC * QLD;
    double* my_racc = racc + (size_t)warp * NT * QT;
    // k-th chunk of this warp: position k*8 + (k even ? warp : 7 - warp) in the list of chunks sorted by
    // descending cost (= descending row index)
    auto chunk_of = [&](int k) { const int pos = k * QWARPS + ((k & 1) ? (QWARPS - 1 - warp) : warp); return nchunks - 1 - pos; };

    auto seek = [&](Cursor& c) {          // make (ci, a) point at the next segment with work, or invalidate
        while (chunk_of(c.ci) >= 0) {
            c.r0 = chunk_of(c.ci) * RC;
            while (c.a < O) {
                c.col0 = c.a * Ns;
                int jm = c.r0 + RC - c.col0;
                jm = jm < 0 ? 0 : (jm > N ? N : jm);
                c.jmax = jm;
                c.nk = (jm + QK - 1) / QK;
                if (c.nk > 0) { c.valid = true; return; }
                ++c.a;
            }
            ++c.ci; c.a = 0;
        }
        c.valid = false;
    };
    auto advance = [&](Cursor& c) {
        if (++c.kb >= c.nk) { c.kb = 0; ++c.a; seek(c); }
    };
    auto issue = [&](const Cursor& c, int slot) {
        const int cg0 = c.col0 + c.kb * QK;
        int rows_valid = n - c.r0; rows_valid = rows_valid > RC ? RC : rows_valid;
        double* dst = ring + slot * RC * QLD;
        const double* src = F + (size_t)c.r0 * p.lda + cg0;
        if (p.dbg & 2) return;
        if (rows_valid == RC && cg0 + QK <= c.col0 + c.jmax && cg0 + QK <= c.r0 + 1) {
            // interior tile (entirely below the diagonal, inside the block): 8 unconditional 16-byte copies per lane
            const int q = lane & 15, rb = lane >> 4;
#pragma unroll
            for (int i = 0; i < RC / 2; ++i) {
                const int r = rb + 2 * i;
                cp_async16(dst + r * QLD + 2 * q, src + (size_t)r * p.lda + 2 * q, 16);
            }
        } else {
            load_tile_async<QK, QLD, true>(dst, src, p.lda, RC, rows_valid, c.r0, cg0, c.col0 + c.jmax, lane, 32);
        }
    };

    Cursor prod{0, 0, 0, 0, 0, 0, 0, false}, cons{0, 0, 0, 0, 0, 0, 0, false};
    seek(prod);
    seek(cons);
    int issued = 0, consumed = 0;
#pragma unroll
    for (int st = 0; st < QST - 1; ++st) {
        if (prod.valid) { issue(prod, issued % QST); advance(prod); ++issued; }
        cp_async_commit();
    }
    // kappa panel (the B operand), evaluated while the first Linv tiles are in flight: warp-strided over the
    // queries, lane-strided over the stored points
    for (int m = warp; m < QT; m += QWARPS) {
        const bool live = m0 + m < p.M;
        for (int j = lane; j < ldp; j += 64) {          // two independent exp chains per iteration
            const int j2 = j + 32;
            const bool c1 = j < N && live && !(p.dbg & 4), c2 = j2 < N && live && !(p.dbg & 4);
            // branch-free (indices clamped, results selected) so that the two chains are interleaved by the scheduler
            const double e1 = rbf(sqs + m * I, s + (size_t)min(j, N - 1) * I, I, p.sigma_f, 0.0, 0.0);
            const double e2 = rbf(sqs + m * I, s + (size_t)min(j2, N - 1) * I, I, p.sigma_f, 0.0, 0.0);
            const double v = c1 ? e1 : 0.0, v2 = c2 ? e2 : 0.0;
            panel[(size_t)m * ldp + j] = v;
            if (j2 < ldp) panel[(size_t)m * ldp + j2] = v2;
        }
    }
    __syncthreads();
    double acc0[2][8][2], acc1[2][8][2];
    zero_acc(acc0);
    zero_acc(acc1);
    while (cons.valid) {
        cp_async_wait<QST - 2>();
        __syncwarp();
        if (prod.valid) { issue(prod, issued % QST); advance(prod); ++issued; }
        cp_async_commit();
        const double* tile = ring + (consumed % QST) * RC * QLD;
        const double* Bs = panel + cons.kb * QK;
        const int kvalid = (p.dbg & 1) ? 0 : cons.jmax - cons.kb * QK;      // columns of this k-block that carry data
        if (kvalid >= QK) {
            if (O == 1 || cons.a == 0) {
#pragma unroll
                for (int kk = 0; kk < QK; kk += 4) warp_mma_k4<2, 8, false>(acc0, tile + kk, QLD, Bs + kk, ldp, lane);
            } else {
#pragma unroll
                for (int kk = 0; kk < QK; kk += 4) warp_mma_k4<2, 8, false>(acc1, tile + kk, QLD, Bs + kk, ldp, lane);
            }
        } else {                                          // last k-block of a segment: skip the zero-filled tail
            if (O == 1 || cons.a == 0) {
#pragma unroll 1
                for (int kk = 0; kk < kvalid; kk += 4) warp_mma_k4<2, 8, false>(acc0, tile + kk, QLD, Bs + kk, ldp, lane);
            } else {
#pragma unroll 1
                for (int kk = 0; kk < kvalid; kk += 4) warp_mma_k4<2, 8, false>(acc1, tile + kk, QLD, Bs + kk, ldp, lane);
            }
        }
        ++consumed;
        const int cur_ci = cons.ci, cur_r0 = cons.r0;
        advance(cons);
        if ((!cons.valid || cons.ci != cur_ci) && !(p.dbg & 8)) {
            // chunk complete: reduce the V tile over its 16 rows straight from the accumulators
            double wv[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) { const int r = cur_r0 + i * 8 + g; wv[i] = r < n ? wsh[r] : 0.0; }
            // per-thread partial sums over its two row fragments: val[t][j][e] for column j*8 + 2*t4 + e
            double val[NT][8][2];
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double q00 = 0.0, q01 = 0.0, q11 = 0.0, mu0 = 0.0, mu1 = 0.0;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const double v0 = acc0[i][j][e];
                        q00 += v0 * v0; mu0 += v0 * wv[i];
                        if (O == 2) { const double v1 = acc1[i][j][e]; q01 += v0 * v1; q11 += v1 * v1; mu1 += v1 * wv[i]; }
                    }
                    if (O == 1) { val[0][j][e] = q00; val[1][j][e] = mu0; }
                    else { val[0][j][e] = q00; val[1][j][e] = q01; val[2][j][e] = q11; val[3][j][e] = mu0; val[NT - 1][j][e] = mu1; }
                }
            // reduce-scatter over the 8 row lanes (lane bits 4,3,2 = g): each step halves the number of values a lane
            // keeps, so 7 shuffles per (t, e) instead of 24; lane g ends up owning the column block j == g
            const bool g2 = (lane & 16) != 0, g1 = (lane & 8) != 0, g0 = (lane & 4) != 0;
#pragma unroll
            for (int t = 0; t < NT; ++t)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double s4[4], s2[2];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const double keep = g2 ? val[t][jj + 4][e] : val[t][jj][e];
                        const double send = g2 ? val[t][jj][e] : val[t][jj + 4][e];
                        s4[jj] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        const double keep = g1 ? s4[jj + 2] : s4[jj];
                        const double send = g1 ? s4[jj] : s4[jj + 2];
                        s2[jj] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    }
                    const double keep = g0 ? s2[1] : s2[0];
                    const double send = g0 ? s2[0] : s2[1];
                    const double tot = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                    my_racc[t * QT + g * 8 + 2 * t4 + e] += tot;
                }
            zero_acc(acc0);
            zero_acc(acc1);
        }
    }
    cp_async_wait<0>();
    __syncthreads();

    // ---------------------------------------------------------------- outputs (reference layout, query last)
    double* xi = p.xi + (size_t)b * O * p.M;
    double* sg_out = p.sigma_out ? p.sigma_out + (size_t)b * O * O * p.M : nullptr;
    constexpr int NPAIR = O * (O + 1) / 2;
    for (int idx = tid; idx < NT * QT; idx += QTHREADS) {
        const int t = idx / QT, m = idx - t * QT;
        if (m0 + m >= p.M) continue;
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < QWARPS; ++q) v += racc[((size_t)q * NT + t) * QT + m];
        if (t >= NPAIR) {
            xi[(size_t)(t - NPAIR) * p.M + m0 + m] = v;
        } else if (sg_out) {
            // pair index -> (a <= bb): O == 1: (0,0);  O == 2: 0 -> (0,0), 1 -> (0,1), 2 -> (1,1)
            const int a = (O == 2 && t == 2) ? 1 : 0;
            const int bb = (O == 2 && t >= 1) ? 1 : 0;
            const double val = p.alpha_kmp * ((a == bb ? 1.0 : 0.0) - v);
            sg_out[((size_t)a * O + bb) * p.M + m0 + m] = val;
            if (a != bb) sg_out[((size_t)bb * O + a) * p.M + m0 + m] = val;
        }
    }
}

bool predict_fast_applicable(const PredictArgs& p) {
    if (p.time_driven || p.general || p.O > 2) return false;
    const int ldp = (int)round_up(p.n_max, QK) + 4;
    if ((long long)p.O * comp_stride(p.n_max) > (long long)QMAXCH * RC) return false;
    return fast_smem_bytes(p.O, p.I, p.n_max, ldp) <= 226 * 1024;
}

int launch_predict_fast(PredictArgs& p, cudaStream_t st) {
    p.ldp = (int)round_up(p.n_max, QK) + 4;      // 4 (mod 16) and long enough for the last k-block of a row
    p.resident = 1;
    const size_t smem = fast_smem_bytes(p.O, p.I, p.n_max, p.ldp);
    dim3 grid(ceil_div(p.M, QT), p.batch);
    if (p.O == 1) {
        static size_t set1 = 48 * 1024;
        KMPX_TRY(ensure_dyn_smem(k_kmp_predict_fast<1>, smem, set1, "k_kmp_predict_fast<1>"));
        k_kmp_predict_fast<1><<<grid, QTHREADS, smem, st>>>(p);
    } else {
        static size_t set2 = 48 * 1024;
        KMPX_TRY(ensure_dyn_smem(k_kmp_predict_fast<2>, smem, set2, "k_kmp_predict_fast<2>"));
        k_kmp_predict_fast<2><<<grid, QTHREADS, smem, st>>>(p);
    }
    return check_launch("k_kmp_predict_fast");
}

}  // namespace kmpx
