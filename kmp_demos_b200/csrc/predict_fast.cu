// K4 fast path: plain RBF kernel, O <= 2, symmetric factor (the cfg-1/3/4 shape of BASELINE.json).
//
// Same mathematics as predict.cu (V_a = Linv[:, block a] * kappa; Gram / mean reductions over the rows of V),
// organised so that the FP64 tensor pipe never waits on a CTA-wide barrier and its warps issue nothing but
// fragment loads and DMMAs:
//   * a CTA owns 64 queries of one problem; kappa[m][j] (64 x N) is evaluated once and stays in shared memory;
//   * each of the 8 MMA warps owns a static list of 16-row chunks of Linv (the triangular costs grow with the row
//     index, so the chunks are dealt from the bottom up in boustrophedon order) and consumes them from a PRIVATE
//     2-stage ring of 16 x 32 tiles;
//   * a producer warpgroup feeds the eight rings through the TMA unit (cp.async.bulk.tensor, boxes of 16 rows x 8 columns,
//     64-byte swizzle, one mbarrier pair per stage): one producer lane per MMA warp runs that warp's tile cursor, so the
//     MMA warps carry no address arithmetic at all; the register file is split 232 / 40 between the roles;
//   * TMA tiles carry no triangular mask: every A fragment is masked in registers against the last valid column of
//     its row (columns right of the diagonal, rows beyond the system and columns beyond the stored points may hold
//     anything, NaN included);
//   * warp tile = 16 rows x 64 queries x both outputs: 16 independent DMMA.8x8x4 per k-step of 4 and output,
//     10 LDS.64 per 16 DMMA; a k-step uses the columns {0,1,4,5} / {2,3,6,7} of an 8-column box (conflict-free reads
//     of the swizzled tile), the kappa panel stores its columns in the same order;
//   * the V tile is reduced straight from the accumulator registers (products, a 3-step reduce-scatter over the
//     8 row lanes) into a per-warp slice of shared memory; slices are summed in warp order at the end, so results are
//     bitwise reproducible and independent of the batch a problem sits in.
// Algorithmic flops per problem: O^2 Ns^2 M (+ lower-order reductions); F is streamed from L2 once per CTA.
#include "dmma_tile.cuh"
#include "kmpx_internal.cuh"
#include "tma_util.cuh"
#include <type_traits>

namespace kmpx {

constexpr int QT = 64;          // queries per CTA
constexpr int RC = 16;          // rows of Linv per chunk
constexpr int QK = 32;          // k per stage: 8 k-steps of 4 = 128 DMMA per warp between two ring hand-overs
constexpr int QST = 2;          // ring stages per warp (one tile in flight while one is consumed)
constexpr int QWARPS = 8;       // MMA warps
constexpr int QTHREADS = QWARPS * 32;
constexpr int PTHREADS = QTHREADS + 128;      // + producer warpgroup (setmaxnreg works on groups of four warps)
constexpr int QMAXCH = 80;      // chunks per problem: n_sys <= 1280
constexpr int QSTAGE_BYTES = RC * QK * 8;     // 4 boxes of 16 rows x 64 B
constexpr int QSTAGE_ELEMS = RC * QK;

template <int O> struct FastCfg { static constexpr int NT = O * (O + 1) / 2 + O; };

struct FastSmem { size_t rings, panel, racc, wsh, sqs, etab, bars, total; };     // byte offsets from the 1 KB aligned base
__host__ __device__ inline FastSmem fast_smem_layout(int O, int I, int n_max_pts, int ldp) {
    const int nt = O * (O + 1) / 2 + O;
    const int n_sys_max = O * comp_stride(n_max_pts);
    FastSmem L;
    L.rings = 0;
    L.panel = (size_t)QWARPS * QST * QSTAGE_BYTES;
    L.racc = L.panel + (size_t)QT * ldp * 8;
    L.wsh = L.racc + (size_t)QWARPS * nt * QT * 8;
    L.sqs = L.wsh + (size_t)round_up(n_sys_max, 2) * 8;
    L.etab = L.sqs + (size_t)round_up(QT * I, 2) * 8;
    L.bars = L.etab + (size_t)kExpTabN * 8;
    L.total = L.bars + (size_t)2 * QWARPS * QST * 8 + 1024 /* alignment slack */;
    return L;
}

struct Cursor {   // position of a warp in its flattened list of (chunk, output a, k-block) items
    int ci, a, kb, nk, r0, col0, jmax;
    bool valid;
};
// k-th chunk of MMA warp w: position k*8 + (k even ? w : 7 - w) in the list of chunks sorted by descending cost
// (= descending row index)
__device__ __forceinline__ int fast_chunk_of(int k, int w, int nchunks) {
    const int pos = k * QWARPS + ((k & 1) ? (QWARPS - 1 - w) : w);
    return nchunks - 1 - pos;
}
template <int O>
__device__ __forceinline__ void fast_seek(Cursor& c, int w, int nchunks, int N, int Ns) {   // next segment with work, or invalidate
    while (fast_chunk_of(c.ci, w, nchunks) >= 0) {
        c.r0 = fast_chunk_of(c.ci, w, nchunks) * RC;
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
}
template <int O>
__device__ __forceinline__ void fast_advance(Cursor& c, int w, int nchunks, int N, int Ns) {
    if (++c.kb >= c.nk) { c.kb = 0; ++c.a; fast_seek<O>(c, w, nchunks, N, Ns); }
}

template <int O>
__global__ void __launch_bounds__(PTHREADS, 1) k_kmp_predict_fast(const __grid_constant__ CUtensorMap tmap, PredictArgs p, FastSmem lay) {
    constexpr int NT = FastCfg<O>::NT;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int b = blockIdx.y, m0 = blockIdx.x * QT;
    const int I = p.I;
    const int N = p.n_pts ? p.n_pts[b] : p.n_max;
    const int Ns = comp_stride(N);
    const int n = O * Ns;
    const int ldp = p.ldp;
    const int n_sys_max = O * comp_stride(p.n_max);
    const int nchunks = (n + RC - 1) / RC;

    unsigned char* base = smem_align1k(smem_raw);
    double* rings = reinterpret_cast<double*>(base + lay.rings);        // [QWARPS][QST][16 x 32 swizzled]
    double* panel = reinterpret_cast<double*>(base + lay.panel);        // [QT][ldp], columns in k-step order
    double* racc = reinterpret_cast<double*>(base + lay.racc);          // [QWARPS][NT][QT]
    double* wsh = reinterpret_cast<double*>(base + lay.wsh);            // [n_sys_max]
    double* sqs = reinterpret_cast<double*>(base + lay.sqs);            // [QT][I]
    double* etab = reinterpret_cast<double*>(base + lay.etab);          // [32] table of exp_tab
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + lay.bars);      // full[QWARPS][QST], empty[QWARPS][QST]
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + QWARPS * QST);

    if (tid == 0) {
        for (int i = 0; i < QWARPS * QST; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    // ---------------------------------------------------------------- prologue: ALL twelve warps (the register split
    //      between the roles is made afterwards, so the producer warps' lanes work on the kappa panel too)
    const double* s = p.s + (size_t)b * p.n_max * I;
    const double* sq = p.sq + (size_t)b * p.stride_sq;
    const double* w = p.w + (size_t)b * p.ldw;
    for (int idx = tid; idx < QT * I; idx += PTHREADS) {
        const int m = idx / I;
        sqs[idx] = (m0 + m < p.M) ? sq[(size_t)(m0 + m) * I + (idx - m * I)] : 0.0;
    }
    for (int idx = tid; idx < n_sys_max; idx += PTHREADS) wsh[idx] = idx < n ? w[idx] : 0.0;
    for (int idx = tid; idx < QWARPS * NT * QT; idx += PTHREADS) racc[idx] = 0.0;
    fill_exp2_table(etab, tid, PTHREADS);
    // producer cursors (lanes 0, 1 of the four producer warps: two streams each); the first tiles are requested now
    // so that they land while the panel is evaluated
    const bool is_producer = warp >= QWARPS && lane < 2;
    const int cw = is_producer ? (warp - QWARPS) * 2 + lane : 0;
    Cursor prod{0, 0, 0, 0, 0, 0, 0, false};
    int pstage = 0;
    uint32_t ppar = 0;
    bool pwrapped = false;
    auto produce_one = [&]() {
        const uint32_t bi = 8 * (cw * QST + pstage);
        const uint32_t bar = full0 + bi;
        mbar_expect_tx(bar, QSTAGE_BYTES);
        const uint32_t dst = smem_u32(rings + (size_t)(cw * QST + pstage) * QSTAGE_ELEMS);
        const int col = prod.col0 + prod.kb * QK;
#pragma unroll
        for (int q = 0; q < QK / 8; ++q) tma_load_box(dst + q * (RC * 64), &tmap, col + 8 * q, prod.r0, b, bar);
        fast_advance<O>(prod, cw, nchunks, N, Ns);
        if (++pstage == QST) { pstage = 0; ppar ^= 1; pwrapped = true; }
    };
    if (is_producer) {
        fast_seek<O>(prod, cw, nchunks, N, Ns);
        for (int t = 0; t < QST && prod.valid; ++t) produce_one();
    }
    __syncthreads();          // sqs is complete
    // kappa panel (the B operand), evaluated while the first Linv tiles are in flight: warp-strided over the queries,
    // lane-strided over the stored points; column j goes to position (j & ~7) | swap_bits_1_2(j & 7)  (k-step order).
    // A dependent FP64 operation costs 40-60 cycles here and an exp is a chain of ~25 of them, so every lane keeps
    // FILL_J independent chains in flight (seven columns of one query; fourteen were measured slower); indices are
    // clamped and the results selected, and the exponential itself is branch-free (exp_nonpos), so the chains interleave.
    constexpr int FILL_J = 7, FILL_WARPS = PTHREADS / 32;
    auto fill = [&](auto one_dim) {         // the I == 1 test is hoisted: a branch inside the chain body would serialise the chains
        constexpr bool ONE = decltype(one_dim)::value;
        for (int m = warp; m < QT && !(p.dbg & 4); m += FILL_WARPS) {
            const bool live = m0 + m < p.M;
            for (int j0 = lane; j0 < ldp - 4; j0 += 32 * FILL_J) {       // ldp - 4 is a multiple of 32
                double e[FILL_J];
#pragma unroll
                for (int u = 0; u < FILL_J; ++u) {
                    const double* sp = s + (size_t)min(j0 + 32 * u, N - 1) * I;
                    if (ONE) { const double d = fabs(sqs[m] - sp[0]); e[u] = exp_tab(-p.sigma_f * (d * d), etab); }
                    else e[u] = rbf_fast(sqs + m * I, sp, I, p.sigma_f);
                }
                const int pj = (j0 & ~6) | ((j0 & 2) << 1) | ((j0 & 4) >> 1);       // the same permutation applies to j0 + 32 u
#pragma unroll
                for (int u = 0; u < FILL_J; ++u) {
                    const int j = j0 + 32 * u;
                    if (j < ldp - 4) panel[(size_t)m * ldp + pj + 32 * u] = (j < N && live) ? e[u] : 0.0;
                }
            }
        }
    };
    if (I == 1) fill(std::true_type{}); else fill(std::false_type{});
    __syncthreads();          // panel, wsh, racc are complete

    if (warp >= QWARPS) {
        // ---- producers: one lane per MMA warp runs that warp's tile cursor and issues its TMA boxes
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
        if (lane >= 2) return;
        while (prod.valid) {
            const uint32_t bi = 8 * (cw * QST + pstage);
            if (pwrapped && !mbar_test(empty0 + bi, ppar ^ 1)) { __nanosleep(100); continue; }   // polling must not eat the MMA warps' issue slots
            produce_one();
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");

    // ---------------------------------------------------------------- per-warp streaming loop
    const double* ring = rings + (size_t)warp * QST * QSTAGE_ELEMS;
    double* my_racc = racc + (size_t)warp * NT * QT;
    // per-lane constants of the swizzled 8-column boxes: element offset of (row g, k-step h) and the logical column
    const int sw = (g >> 1) & 3, c0 = (t4 >> 1) << 1;
    const int o0 = g * 8 + ((c0 ^ sw) << 1) + (t4 & 1);
    const int o1 = g * 8 + (((c0 | 1) ^ sw) << 1) + (t4 & 1);
    const int cm0 = (t4 & 1) | ((t4 & 2) << 1);          // k-step 0 uses the columns {0,1,4,5} of a box, k-step 1 {2,3,6,7}
    Cursor cons{0, 0, 0, 0, 0, 0, 0, false};
    fast_seek<O>(cons, warp, nchunks, N, Ns);
    int stage = 0;
    uint32_t par = 0;
    double acc0[2][8][2], acc1[2][8][2];
    zero_acc(acc0);
    zero_acc(acc1);
    while (cons.valid) {
        const uint32_t bi = 8 * (warp * QST + stage);
        mbar_wait(full0 + bi, par);
        const double* tile = ring + stage * QSTAGE_ELEMS;
        const double* Bs = panel + cons.kb * QK;
        // last valid point index of the two rows this lane carries in this segment (-1: row outside the system)
        int lim[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = cons.r0 + i * 8 + g;
            lim[i] = r < n ? min(N - 1, r - cons.col0) - cons.kb * QK : -1;
        }
        const int kvalid = (p.dbg & 1) ? 0 : cons.jmax - cons.kb * QK;      // columns of this k-block that carry data
        const int nst = kvalid >= QK ? 8 : 2 * ((kvalid + 7) >> 3);         // k-steps of 4 (whole boxes)
        const bool first = (O == 1 || cons.a == 0);
        auto kstep = [&](int st, double (&acc)[2][8][2]) {
            const int q = st >> 1, h = st & 1;
            const int jl = 8 * q + 2 * h + cm0;                             // logical column of this lane inside the tile
            double a[2], bf[8];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double v = tile[q * (RC * 8) + i * 64 + (h ? o1 : o0)];
                a[i] = jl <= lim[i] ? v : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) bf[j] = Bs[(size_t)(j * 8 + g) * ldp + 4 * st + t4];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma884(acc[i][j], a[i], bf[j]);
        };
        if (nst == 8) {
            if (first) {
#pragma unroll
                for (int st = 0; st < 8; ++st) kstep(st, acc0);
            } else {
#pragma unroll
                for (int st = 0; st < 8; ++st) kstep(st, acc1);
            }
        } else {                                          // last k-block of a segment: skip the boxes without data
            if (first) {
#pragma unroll 1
                for (int st = 0; st < nst; ++st) kstep(st, acc0);
            } else {
#pragma unroll 1
                for (int st = 0; st < nst; ++st) kstep(st, acc1);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + bi);
        if (++stage == QST) { stage = 0; par ^= 1; }
        const int cur_ci = cons.ci, cur_r0 = cons.r0;
        fast_advance<O>(cons, warp, nchunks, N, Ns);
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
    asm volatile("bar.sync 1, 256;\n" ::: "memory");

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
    if ((p.lda & 1) || (p.batch > 1 && (p.strideA & 1)) || (reinterpret_cast<uintptr_t>(p.F) & 15)) return false;   // TMA
    const int ldp = (int)round_up(p.n_max, QK) + 4;
    if ((long long)p.O * comp_stride(p.n_max) > (long long)QMAXCH * RC) return false;
    return fast_smem_layout(p.O, p.I, p.n_max, ldp).total <= 226 * 1024;
}

int launch_predict_fast(PredictArgs& p, cudaStream_t st) {
    p.ldp = (int)round_up(p.n_max, QK) + 4;      // 4 (mod 16) and long enough for the last k-block of a row
    p.resident = 1;
    const FastSmem lay = fast_smem_layout(p.O, p.I, p.n_max, p.ldp);
    const int n_sys_max = p.O * comp_stride(p.n_max);
    CUtensorMap map;
    KMPX_TRY(make_batch_map(&map, p.F, p.lda, p.strideA, n_sys_max, p.batch, RC));
    dim3 grid(ceil_div(p.M, QT), p.batch);
    if (p.O == 1) {
        static SmemCache set1;
        KMPX_TRY(ensure_dyn_smem(k_kmp_predict_fast<1>, lay.total, set1, "k_kmp_predict_fast<1>"));
        k_kmp_predict_fast<1><<<grid, PTHREADS, lay.total, st>>>(map, p, lay);
    } else {
        static SmemCache set2;
        KMPX_TRY(ensure_dyn_smem(k_kmp_predict_fast<2>, lay.total, set2, "k_kmp_predict_fast<2>"));
        k_kmp_predict_fast<2><<<grid, PTHREADS, lay.total, st>>>(map, p, lay);
    }
    return check_launch("k_kmp_predict_fast");
}

}  // namespace kmpx
