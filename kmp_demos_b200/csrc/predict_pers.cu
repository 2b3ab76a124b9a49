// K4 persistent fast path: plain RBF kernel, O <= 2, symmetric factor, batches with many (problem, query tile) pairs.
//
// Same mathematics, warp tile and TMA rings as k_kmp_predict_fast (predict_fast.cu); what changes is what happens
// AROUND the streaming loop.  There a CTA owns one tile of 64 queries: all twelve warps evaluate the kappa panel
// (12 k of its 132 k cycles), the eight MMA warps contract their static chunk lists (5 % imbalance, the slowest warp
// sets the time) and a CTA-wide epilogue writes the results - prologue, imbalance and epilogue are all exposed,
// because the 117 KB panel allows one CTA per SM.  Here the CTA is persistent over tiles and the roles never meet
// at a CTA barrier:
//   * tiles of QT = 40 queries, so that TWO kappa panels fit: while the MMA warps contract tile i, the four fill warps
//     evaluate the panel of tile i+1 (and the query / w vectors that go with it) and then reduce + write the results of
//     tile i-1; hand-overs through mbarriers (panel full / empty, partial sums full / empty);
//   * an MMA warp that is done with its chunks of tile i starts on tile i+1 at once; the chunk lists rotate with the
//     tile index (list (w + tile) mod 8 of the boustrophedon deal), so that the long and the short lists even out
//     over the tiles of a CTA instead of costing their difference on every tile.  The rotation depends on the
//     query-tile index only, hence the summation order of a problem's results - and with it every bit of them - is
//     independent of the batch the problem sits in and of the grid;
//   * every MMA warp issues the TMA boxes of its own ring (its lane 0 runs a second cursor PST - 1 stages ahead, across
//     tile boundaries): no producer warp, no empty barriers, no polling - a stage is re-armed the moment it is released.
//     (A single producer warp whose lanes poll eight rings was measured to bound the kernel at 3 TB/s of L2 traffic.)
// Algorithmic flops per problem: O^2 Ns^2 M (+ lower-order reductions); F is streamed from L2 once per tile.
#include "dmma_tile.cuh"
#include "kmpx_internal.cuh"
#include "tma_util.cuh"
#include <type_traits>

namespace kmpx {

constexpr int PRC = 16;          // rows of Linv per chunk
constexpr int PWARPS = 8;        // MMA warps
constexpr int PST = 2;           // ring stages per MMA warp
constexpr int PFILL = 128;       // threads of the four fill warps
constexpr int PALL = 384;        // 8 MMA warps + 4 fill warps
constexpr int PMAXCH = 80;       // chunks per problem: n_sys <= 1280

struct PersSmem { size_t rings, panel, racc, wsh, sqs, etab, bars, total; int ldp, wsh_elems, sqs_elems; };
template <int QT, int QK>
__host__ __device__ inline PersSmem pers_smem_layout(int O, int I, int n_max_pts) {
    const int nt = O * (O + 1) / 2 + O;
    PersSmem L;
    L.ldp = (int)round_up(n_max_pts, 8) + 4;              // 4 or 12 (mod 16): conflict-free B fragment reads
    L.wsh_elems = (int)round_up(O * comp_stride(n_max_pts), 2);
    L.sqs_elems = (int)round_up(QT * I, 2);
    L.rings = 0;
    L.panel = (size_t)PWARPS * PST * PRC * QK * 8;
    L.racc = L.panel + (size_t)2 * QT * L.ldp * 8;
    L.wsh = L.racc + (size_t)PWARPS * nt * QT * 8;
    L.sqs = L.wsh + (size_t)2 * L.wsh_elems * 8;
    L.etab = L.sqs + (size_t)2 * L.sqs_elems * 8;
    L.bars = L.etab + (size_t)kExpTabN * 8;
    L.total = L.bars + (size_t)(PWARPS * PST + 6) * 8 + 1024 /* alignment slack */;
    return L;
}

struct PCursor {   // position of a warp in its flattened list of (chunk, output a, k-block) items of one tile
    int ci, a, kb, nk, r0, col0, jmax;
    bool valid;
};
// k-th chunk of list l: position k*8 + (k even ? l : 7 - l) in the list of chunks sorted by descending cost (= descending row index)
__device__ __forceinline__ int pers_chunk_of(int k, int l, int nchunks) {
    const int pos = k * PWARPS + ((k & 1) ? (PWARPS - 1 - l) : l);
    return nchunks - 1 - pos;
}
template <int O, int QK>
__device__ __forceinline__ void pers_seek(PCursor& c, int l, int nchunks, int N, int Ns) {   // next segment with work, or invalidate
    while (pers_chunk_of(c.ci, l, nchunks) >= 0) {
        c.r0 = pers_chunk_of(c.ci, l, nchunks) * PRC;
        while (c.a < O) {
            c.col0 = c.a * Ns;
            int jm = c.r0 + PRC - c.col0;
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
template <int O, int QK>
__device__ __forceinline__ void pers_advance(PCursor& c, int l, int nchunks, int N, int Ns) {
    if (++c.kb >= c.nk) { c.kb = 0; ++c.a; pers_seek<O, QK>(c, l, nchunks, N, Ns); }
}
__device__ __forceinline__ void pers_cursor_start(PCursor& c) { c.ci = 0; c.a = 0; c.kb = 0; c.nk = 0; c.r0 = 0; c.col0 = 0; c.jmax = 0; c.valid = false; }

template <int O, int QT, int QK>
__global__ void __launch_bounds__(PALL, 1) k_kmp_predict_pers(const __grid_constant__ CUtensorMap tmap, PredictArgs p, PersSmem lay) {
    constexpr int NT = O * (O + 1) / 2 + O;
    constexpr int NJ = QT / 8;                          // n-tiles of 8 queries
    constexpr int STAGE_ELEMS = PRC * QK;
    constexpr int STAGE_BYTES = STAGE_ELEMS * 8;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int I = p.I, ldp = lay.ldp;
    const int ntile = (p.M + QT - 1) / QT;
    const long long total = (long long)p.batch * ntile;

    unsigned char* base = smem_align1k(smem_raw);
    double* rings = reinterpret_cast<double*>(base + lay.rings);        // [PWARPS][PST][16 x QK swizzled]
    double* panel0 = reinterpret_cast<double*>(base + lay.panel);       // [2][QT][ldp], columns in k-step order
    double* racc = reinterpret_cast<double*>(base + lay.racc);          // [PWARPS][NT][QT]
    double* wsh0 = reinterpret_cast<double*>(base + lay.wsh);           // [2][wsh_elems]
    double* sqs0 = reinterpret_cast<double*>(base + lay.sqs);           // [2][QT][I]
    double* etab = reinterpret_cast<double*>(base + lay.etab);          // [32] table of exp_tab
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + lay.bars);
    const uint32_t full0 = smem_u32(bars);
    const uint32_t pfull0 = smem_u32(bars + PWARPS * PST), pempty0 = pfull0 + 16;      // panel [2] each
    const uint32_t rfull = pfull0 + 32, rempty = pfull0 + 40;                             // partial sums

    if (tid == 0) {
        for (int i = 0; i < PWARPS * PST; ++i) mbar_init(full0 + 8 * i, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(pfull0 + 8 * i, PFILL); mbar_init(pempty0 + 8 * i, PWARPS); }
        mbar_init(rfull, PWARPS);
        mbar_init(rempty, PFILL);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int idx = tid; idx < PWARPS * NT * QT; idx += PALL) racc[idx] = 0.0;
    fill_exp2_table(etab, tid, PALL);
    __syncthreads();

    if (warp >= PWARPS) {
        // ---------------------------------------------------------------- fill warps: kappa panel of tile i+1, results of tile i-1
        asm volatile("setmaxnreg.dec.sync.aligned.u32 136;\n");
        const int ft = tid - PWARPS * 32, fw = ft >> 5;
        constexpr int NPAIR = O * (O + 1) / 2;
        auto output = [&](long long T, int it) {
            const int b = (int)(T / ntile), m0 = (int)(T - (long long)b * ntile) * QT;
            mbar_wait(rfull, it & 1);
            double* xi = p.xi + (size_t)b * O * p.M;
            double* sg_out = p.sigma_out ? p.sigma_out + (size_t)b * O * O * p.M : nullptr;
            for (int idx = ft; idx < NT * QT; idx += PFILL) {
                const int t = idx / QT, m = idx - t * QT;
                double v = 0.0;
#pragma unroll
                for (int q = 0; q < PWARPS; ++q) { v += racc[((size_t)q * NT + t) * QT + m]; racc[((size_t)q * NT + t) * QT + m] = 0.0; }
                if (m0 + m >= p.M) continue;
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
            mbar_arrive(rempty);
        };
        int it = 0;
        long long Tprev = -1;
        for (long long T = blockIdx.x; T < total; T += gridDim.x, ++it) {
            const int buf = it & 1;
            const int b = (int)(T / ntile), m0 = (int)(T - (long long)b * ntile) * QT;
            const int N = p.n_pts ? p.n_pts[b] : p.n_max;
            const int n = O * comp_stride(N);
            if (it >= 2) mbar_wait(pempty0 + 8 * buf, ((it >> 1) - 1) & 1);
            double* panel = panel0 + (size_t)buf * QT * ldp;
            double* wsh = wsh0 + (size_t)buf * lay.wsh_elems;
            double* sqs = sqs0 + (size_t)buf * lay.sqs_elems;
            const double* s = p.s + (size_t)b * p.n_max * I;
            const double* sq = p.sq + (size_t)b * p.stride_sq;
            const double* w = p.w + (size_t)b * p.ldw;
            for (int idx = ft; idx < QT * I; idx += PFILL) {
                const int m = idx / I;
                sqs[idx] = (m0 + m < p.M) ? sq[(size_t)(m0 + m) * I + (idx - m * I)] : 0.0;
            }
            for (int idx = ft; idx < lay.wsh_elems; idx += PFILL) wsh[idx] = idx < n ? w[idx] : 0.0;
            asm volatile("bar.sync 2, 128;\n" ::: "memory");          // sqs is complete
            // column j goes to position (j & ~7) | swap_bits_1_2(j & 7)  (k-step order); FILL_J independent chains per lane
            constexpr int FILL_J = 7;
            auto fill = [&](auto one_dim) {
                constexpr bool ONE = decltype(one_dim)::value;
                for (int m = fw; m < QT && !(p.dbg & 4); m += 4) {
                    const bool live = m0 + m < p.M;
                    for (int j0 = lane; j0 < ldp - 4; j0 += 32 * FILL_J) {
                        double e[FILL_J];
#pragma unroll
                        for (int u = 0; u < FILL_J; ++u) {
                            const double* sp = s + (size_t)min(j0 + 32 * u, N - 1) * I;
                            if (ONE) { const double d = fabs(sqs[m] - sp[0]); e[u] = exp_tab(-p.sigma_f * (d * d), etab); }
                            else e[u] = rbf_tab(sqs + m * I, sp, I, p.sigma_f, etab);
                        }
                        const int pj = (j0 & ~6) | ((j0 & 2) << 1) | ((j0 & 4) >> 1);
#pragma unroll
                        for (int u = 0; u < FILL_J; ++u) {
                            const int j = j0 + 32 * u;
                            if (j < ldp - 4) panel[(size_t)m * ldp + pj + 32 * u] = (j < N && live) ? e[u] : 0.0;
                        }
                    }
                }
            };
            if (I == 1) fill(std::true_type{}); else fill(std::false_type{});
            mbar_arrive(pfull0 + 8 * buf);
            if (Tprev >= 0) output(Tprev, it - 1);
            Tprev = T;
        }
        if (Tprev >= 0) output(Tprev, it - 1);
        return;
    }

    // ---------------------------------------------------------------- MMA warps
    // register split: the fill code needs 130 registers (measured by compiling the role alone), the MMA warps take what the
    // other warpgroup gives back: (168 - 136) * 128 = (184 - 168) * 256 - a larger request would wait forever
    asm volatile("setmaxnreg.inc.sync.aligned.u32 184;\n");
    const double* ring = rings + (size_t)warp * PST * STAGE_ELEMS;
    double* my_racc = racc + (size_t)warp * NT * QT;
    // per-lane constants of the swizzled 8-column boxes: element offset of (row g, k-step h) and the logical column
    const int sw = (g >> 1) & 3, c0 = (t4 >> 1) << 1;
    const int o0 = g * 8 + ((c0 ^ sw) << 1) + (t4 & 1);
    const int o1 = g * 8 + (((c0 | 1) ^ sw) << 1) + (t4 & 1);
    const int cm0 = (t4 & 1) | ((t4 & 2) << 1);          // k-step 0 uses the columns {0,1,4,5} of a box, k-step 1 {2,3,6,7}
    // ---- producer side of this warp's ring: a second cursor that runs up to PST stages ahead of the consumer, across tiles
    uint32_t issued = 0, consumed = 0;
    long long pT = blockIdx.x;
    int pb = 0, pN = 0, pNs = 0, pnchunks = 0, plist = 0;
    PCursor prod;
    pers_cursor_start(prod);
    auto prod_open_tile = [&]() {       // positions the producer cursor on the first item of tile pT (or of a later tile)
        while (pT < total) {
            pb = (int)(pT / ntile);
            const int mt = (int)(pT - (long long)pb * ntile);
            pN = p.n_pts ? p.n_pts[pb] : p.n_max;
            pNs = comp_stride(pN);
            pnchunks = (O * pNs + PRC - 1) / PRC;
            plist = (warp + mt) & (PWARPS - 1);
            pers_cursor_start(prod);
            pers_seek<O, QK>(prod, plist, pnchunks, pN, pNs);
            if (prod.valid) return;
            pT += gridDim.x;
        }
    };
    prod_open_tile();
    auto issue = [&]() {
        while (prod.valid && issued - consumed < (uint32_t)PST) {
            if (lane == 0) {
                const uint32_t sidx = warp * PST + issued % PST;
                const uint32_t bar = full0 + 8 * sidx;
                mbar_expect_tx(bar, STAGE_BYTES);
                const uint32_t dst = smem_u32(rings + (size_t)sidx * STAGE_ELEMS);
                const int col = prod.col0 + prod.kb * QK;
#pragma unroll
                for (int q = 0; q < QK / 8; ++q) tma_load_box(dst + q * (PRC * 64), &tmap, col + 8 * q, prod.r0, pb, bar);
            }
            ++issued;
            pers_advance<O, QK>(prod, plist, pnchunks, pN, pNs);
            if (!prod.valid) { pT += gridDim.x; prod_open_tile(); }
        }
    };
    issue();
    int it = 0;
    for (long long T = blockIdx.x; T < total; T += gridDim.x, ++it) {
        const int buf = it & 1;
        const int b = (int)(T / ntile), mt = (int)(T - (long long)b * ntile);
        const int N = p.n_pts ? p.n_pts[b] : p.n_max;
        const int Ns = comp_stride(N);
        const int n = O * Ns;
        const int nchunks = (n + PRC - 1) / PRC;
        const int list = (warp + mt) & (PWARPS - 1);
        const double* panel = panel0 + (size_t)buf * QT * ldp;
        const double* wsh = wsh0 + (size_t)buf * lay.wsh_elems;
        mbar_wait(pfull0 + 8 * buf, (it >> 1) & 1);
        bool racc_free = it == 0;
        PCursor cons;
        pers_cursor_start(cons);
        pers_seek<O, QK>(cons, list, nchunks, N, Ns);
        double acc0[2][NJ][2], acc1[2][NJ][2];
        zero_acc(acc0);
        zero_acc(acc1);
        while (cons.valid) {
            const uint32_t stage = consumed % PST;
            mbar_wait(full0 + 8 * (warp * PST + stage), (consumed / PST) & 1);
            const double* tile = ring + stage * STAGE_ELEMS;
            const double* Bs = panel + cons.kb * QK;
            // last valid point index of the two rows this lane carries in this segment (-1: row outside the system)
            int lim[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int r = cons.r0 + i * 8 + g;
                lim[i] = r < n ? min(N - 1, r - cons.col0) - cons.kb * QK : -1;
            }
            const int kvalid = (p.dbg & 1) ? 0 : cons.jmax - cons.kb * QK;      // columns of this k-block that carry data
            constexpr int KST = QK / 4;
            const int nst = kvalid >= QK ? KST : 2 * ((kvalid + 7) >> 3);       // k-steps of 4 (whole boxes)
            const bool first = (O == 1 || cons.a == 0);
            auto kstep = [&](int st, double (&acc)[2][NJ][2]) {
                const int q = st >> 1, h = st & 1;
                const int jl = 8 * q + 2 * h + cm0;                             // logical column of this lane inside the tile
                double a[2], bf[NJ];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const double v = tile[q * (PRC * 8) + i * 64 + (h ? o1 : o0)];
                    a[i] = jl <= lim[i] ? v : 0.0;
                }
#pragma unroll
                for (int j = 0; j < NJ; ++j) bf[j] = Bs[(size_t)(j * 8 + g) * ldp + 4 * st + t4];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) dmma884(acc[i][j], a[i], bf[j]);
            };
            if (nst == KST) {
                if (first) {
#pragma unroll
                    for (int st = 0; st < KST; ++st) kstep(st, acc0);
                } else {
#pragma unroll
                    for (int st = 0; st < KST; ++st) kstep(st, acc1);
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
            ++consumed;
            issue();          // re-arm the stage just released
            const int cur_ci = cons.ci, cur_r0 = cons.r0;
            pers_advance<O, QK>(cons, list, nchunks, N, Ns);
            if ((!cons.valid || cons.ci != cur_ci) && !(p.dbg & 8)) {
                // chunk complete: reduce the V tile over its 16 rows straight from the accumulators
                if (!racc_free) { mbar_wait(rempty, (it - 1) & 1); racc_free = true; }    // the fill warps have consumed the previous tile's sums
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
                        if (j < NJ) {
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                const double v0 = acc0[i][j < NJ ? j : 0][e];
                                q00 += v0 * v0; mu0 += v0 * wv[i];
                                if (O == 2) { const double v1 = acc1[i][j < NJ ? j : 0][e]; q01 += v0 * v1; q11 += v1 * v1; mu1 += v1 * wv[i]; }
                            }
                        }
                        if (O == 1) { val[0][j][e] = q00; val[1][j][e] = mu0; }
                        else { val[0][j][e] = q00; val[1][j][e] = q01; val[2][j][e] = q11; val[3][j][e] = mu0; val[NT - 1][j][e] = mu1; }
                    }
                // reduce-scatter over the 8 row lanes (lane bits 4,3,2 = g): lane g ends up owning the column block j == g
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
                        if (g < NJ) my_racc[t * QT + g * 8 + 2 * t4 + e] += tot;
                    }
                zero_acc(acc0);
                zero_acc(acc1);
            }
        }
        if (!racc_free) mbar_wait(rempty, (it - 1) & 1);          // (a warp without chunks: keeps the phases in step)
        __syncwarp();
        if (lane == 0) { mbar_arrive(rfull); mbar_arrive(pempty0 + 8 * buf); }
    }
}

template <int QT, int QK>
static bool pers_fits(const PredictArgs& p) {
    return pers_smem_layout<QT, QK>(p.O, p.I, p.n_max).total <= 226 * 1024;
}
constexpr int PERS_QT = 40, PERS_QK = 32;

bool predict_pers_applicable(const PredictArgs& p) {
    if (p.time_driven || p.general || p.O > 2) return false;
    if ((p.lda & 1) || (p.batch > 1 && (p.strideA & 1)) || (reinterpret_cast<uintptr_t>(p.F) & 15)) return false;   // TMA
    if ((long long)p.O * comp_stride(p.n_max) > (long long)PMAXCH * PRC) return false;
    return pers_fits<PERS_QT, PERS_QK>(p);
}

template <int O, int QT, int QK>
static int launch_pers(const CUtensorMap& map, const PredictArgs& p, const PersSmem& lay, cudaStream_t st) {
    static SmemCache set;
    KMPX_TRY(ensure_dyn_smem(k_kmp_predict_pers<O, QT, QK>, lay.total, set, "k_kmp_predict_pers"));
    const long long total = (long long)p.batch * ceil_div(p.M, QT);
    const int grid = total < num_sms() ? (int)total : num_sms();
    k_kmp_predict_pers<O, QT, QK><<<grid, PALL, lay.total, st>>>(map, p, lay);
    return check_launch("k_kmp_predict_pers");
}

int launch_predict_pers(PredictArgs& p, cudaStream_t st) {
    const PersSmem lay = pers_smem_layout<PERS_QT, PERS_QK>(p.O, p.I, p.n_max);
    p.ldp = lay.ldp;
    p.resident = 1;
    const int n_sys_max = p.O * comp_stride(p.n_max);
    CUtensorMap map;
    KMPX_TRY(make_batch_map(&map, p.F, p.lda, p.strideA, n_sys_max, p.batch, PRC));
    if (p.O == 1) return launch_pers<1, PERS_QT, PERS_QK>(map, p, lay, st);
    return launch_pers<2, PERS_QT, PERS_QK>(map, p, lay, st);
}

}  // namespace kmpx
