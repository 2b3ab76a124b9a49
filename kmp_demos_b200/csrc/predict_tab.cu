// K4 fast path, table-driven form: plain RBF kernel, O <= 2, symmetric factor (the cfg-1/3 shape of BASELINE.json).
//
// Same mathematics, CTA structure (a tile of queries of one problem, kappa panel resident in shared memory, MMA warps with
// private 2-stage TMA rings fed by a producer warpgroup), warp tile and reductions as k_kmp_predict_fast (predict_fast.cu),
// in two shapes: 64 queries / eight MMA warps / one CTA per SM, and 32 queries / four MMA warps / two CTAs per SM (the
// default when kappa boxes are skipped: one CTA's set-up, panel and epilogue run under the other's contraction).
// What changes is everything an MMA warp does BETWEEN its DMMAs.  Measured with clock64 on one MMA warp of that kernel
// (112 k cycles in the streaming loop): 81 k in full stages - at the pipe's rate, two warps per scheduler saturate it - but
// 13 k in the cursor arithmetic between two stages (which chunk, which output, which k-block, how many columns are valid:
// ~570 cycles of dependent integer code per stage), 9 k in the per-chunk reductions, 5 k waiting for tiles and 5 k in the
// rolled loop of the short last k-blocks; a single warp reaches only ~57 % of its scheduler's DMMA rate, so whatever one warp
// spends outside its DMMA loop is only half covered by its partner.  Here
//   * the stage sequence of every warp - and the assignment of the 16-row chunks to the warps - is a pure function of
//     (O, N): the host tabulates it once per (device, O, n_max) for every N (8 bytes per stage: row, column, k-steps, flags),
//     the CTA copies its row of the table into shared memory, and producer lanes and MMA warps just walk it;
//   * the assignment is longest-first by a cycle model of stages and reductions, refined by moves and swaps off the most
//     loaded warp (the boustrophedon deal left 5 % between the first and the last warp);
//   * chunks of the first component block (V_1 = 0) reduce only the two sums that are not identically zero;
//   * the short last k-block runs its boxes (two k-steps each) unrolled;
//   * RBF locality: boxes of 8 stored points that are further than sqrt(-ln(2^-68) / sigma_f) from the bounding
//     box of the tile's queries have kappa < 2^-68 for the whole tile and are skipped - stages without a live box are never
//     requested, partly live stages request and run their live boxes only; at the reference's sigma_f = 200 half of the
//     contraction goes, and what is dropped is below the rounding error of what is kept (kmpx_predict_skip_threshold moves
//     or disables the threshold).
// Results are bitwise reproducible and independent of the batch a problem sits in (the schedule depends on (O, N) only).
// Algorithmic flops per problem: O^2 Ns^2 M (+ lower-order reductions); F is streamed from L2 once per CTA.
#include "dmma_tile.cuh"
#include "kmpx_internal.cuh"
#include "tma_util.cuh"
#include <algorithm>
#include <cmath>
#include <map>
#include <mutex>
#include <tuple>
#include <type_traits>
#include <vector>

namespace kmpx {

__device__ unsigned long long g_pstats[2];      // k-steps executed / scheduled by k_kmp_predict_tab, counted under kmpx_debug_set(32768)

constexpr int TRC = 16;          // rows of Linv per chunk
constexpr int TQK = 32;          // k per stage: 8 k-steps of 4 = 128 DMMA per warp between two ring hand-overs
constexpr int TST = 2;           // ring stages per warp
constexpr int TWST = 384;        // MMA warps x stages per warp the table holds (8 x 48 or 4 x 96: n_sys <= ~600; larger systems take k_kmp_predict_fast)
constexpr int TSTAGE_BYTES = TRC * TQK * 8;   // 4 boxes of 16 rows x 64 B
constexpr int TSTAGE_ELEMS = TRC * TQK;
constexpr int TROW = 8 + TWST * 2;             // ints per table row: stage counts per warp, then [warp][stage]{d0, d1}

// stage descriptor: d0 = r0 | a << 12 | chunk_end << 13 | has_v1 << 14 | nst << 16;  d1 = k offset in the panel | col0 << 16
__host__ __device__ inline int tdesc0(int r0, int a, int chunk_end, int has_v1, int nst) { return r0 | (a << 12) | (chunk_end << 13) | (has_v1 << 14) | (nst << 16); }

struct TabSmem { size_t rings, panel, racc, wsh, sqs, etab, sched, clist, bars, total; };
__host__ __device__ inline TabSmem tab_smem_layout(int O, int I, int n_max_pts, int ldp, int TQT, int TWARPS) {
    const int nt = O * (O + 1) / 2 + O;
    const int n_sys_max = O * comp_stride(n_max_pts);
    TabSmem L;
    L.rings = 0;
    L.panel = (size_t)TWARPS * TST * TSTAGE_BYTES;
    L.racc = L.panel + (size_t)TQT * ldp * 8;
    L.wsh = L.racc + (size_t)TWARPS * nt * TQT * 8;
    L.sqs = L.wsh + (size_t)round_up(n_sys_max, 2) * 8;
    L.etab = L.sqs + (size_t)round_up(TQT * I, 2) * 8;
    L.sched = L.etab + (size_t)kExpTabN * 8;
    L.clist = L.sched + (size_t)(TROW + 8 + 16) * 4;    // + the kappa box masks of the tile and the corners of its query box
    L.bars = L.clist + (size_t)TWST * 2 * 4;            // the live stages of every MMA warp, compacted per tile
    L.total = L.bars + (size_t)2 * TWARPS * TST * 8 + 1024 /* alignment slack */;
    return L;
}

// ---- host: the schedule of one (O, N).  Returns false when a warp would need more than TMAXST stages.
static bool tab_schedule_host(int O, int N, int* row, int TWARPS) {
    const int TMAXST = TWST / TWARPS;
    const int Ns = comp_stride(N), n = O * Ns;
    const int nchunks = (n + TRC - 1) / TRC;
    struct Stage { int d0, d1; };
    std::vector<std::vector<Stage>> st(nchunks);
    std::vector<int> cost(nchunks), order(nchunks), owner(nchunks);
    for (int c = 0; c < nchunks; ++c) {
        const int r0 = c * TRC;
        int cyc = 0, has_v1 = 0;
        for (int a = 0; a < O; ++a) {
            int jm = r0 + TRC - a * Ns;
            jm = jm < 0 ? 0 : (jm > N ? N : jm);
            if (jm > 0 && a == 1) has_v1 = 1;
            for (int kb = 0; kb * TQK < jm; ++kb) {
                const int kv = jm - kb * TQK;
                const int nst = kv >= TQK ? TQK / 4 : 2 * ((kv + 7) >> 3);
                st[c].push_back({tdesc0(r0, a, 0, 0, nst), (kb * TQK) | ((a * Ns) << 16)});
                cyc += nst * 32 * 16 / 16 + 120;              // 16 DMMAs per k-step at one per 32 cycles and warp; hand-over
            }
        }
        for (auto& sgm : st[c]) sgm.d0 |= has_v1 << 14;
        if (!st[c].empty()) st[c].back().d0 |= 1 << 13;
        cost[c] = cyc + (has_v1 ? 700 : 330);
        order[c] = c;
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    long long load[8] = {0};
    for (int k = 0; k < nchunks; ++k) {
        int best = 0;
        for (int w = 1; w < TWARPS; ++w) if (load[w] < load[best]) best = w;
        owner[order[k]] = best;
        load[best] += cost[order[k]];
    }
    // local search: move a chunk off, or swap a chunk of, the most loaded warp while that lowers the maximum
    for (int iter = 0; iter < 400; ++iter) {
        int hi = 0;
        for (int w = 1; w < TWARPS; ++w) if (load[w] > load[hi]) hi = w;
        long long best_gain = 0;
        int bi = -1, bj = -1, bw = -1;
        for (int i = 0; i < nchunks; ++i) {
            if (owner[i] != hi) continue;
            for (int w = 0; w < TWARPS; ++w) {
                if (w == hi) continue;
                const long long after = std::max(load[hi] - cost[i], load[w] + cost[i]);
                if (load[hi] - after > best_gain) { best_gain = load[hi] - after; bi = i; bj = -1; bw = w; }
                for (int j = 0; j < nchunks; ++j) {
                    if (owner[j] != w || cost[j] >= cost[i]) continue;
                    const long long after2 = std::max(load[hi] - cost[i] + cost[j], load[w] + cost[i] - cost[j]);
                    if (load[hi] - after2 > best_gain) { best_gain = load[hi] - after2; bi = i; bj = j; bw = w; }
                }
            }
        }
        if (bi < 0) break;
        load[hi] -= cost[bi]; load[bw] += cost[bi]; owner[bi] = bw;
        if (bj >= 0) { load[bw] -= cost[bj]; load[hi] += cost[bj]; owner[bj] = hi; }
    }
    int cnt[8] = {0};
    for (int k = 0; k < nchunks; ++k) {             // every warp takes its chunks longest first
        const int c = order[k], w = owner[c];
        for (const auto& sgm : st[c]) {
            if (cnt[w] >= TMAXST) return false;
            row[8 + (w * TMAXST + cnt[w]) * 2] = sgm.d0;
            row[8 + (w * TMAXST + cnt[w]) * 2 + 1] = sgm.d1;
            ++cnt[w];
        }
    }
    for (int w = 0; w < TWARPS; ++w) row[w] = cnt[w];
    return true;
}
// device copy of the table for (O, n_max), built on first use (per device; never freed: 3 KB per N).  *out = nullptr when
// some N <= n_max does not fit the table (the caller then takes k_kmp_predict_fast).
static int tab_sched_table(int O, int n_max, int NW, const int** out) {
    static std::mutex mu;
    static std::map<std::tuple<int, int, int, int>, int*> cache;
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_tuple(current_device(), O, n_max, NW);
    auto it = cache.find(key);
    if (it == cache.end()) {
        std::vector<int> host((size_t)(n_max + 1) * TROW, 0);
        bool ok = true;
        for (int N = 1; N <= n_max && ok; ++N) ok = tab_schedule_host(O, N, host.data() + (size_t)N * TROW, NW);
        int* dptr = nullptr;
        if (ok) {
            KMPX_TRY(check_cuda(cudaMalloc(&dptr, host.size() * sizeof(int)), "cudaMalloc (predict schedule table)"));
            KMPX_TRY(check_cuda(cudaMemcpy(dptr, host.data(), host.size() * sizeof(int), cudaMemcpyHostToDevice), "cudaMemcpy (predict schedule table)"));
        }
        it = cache.emplace(key, dptr).first;
    }
    *out = it->second;
    return 0;
}

// Two CTA shapes: <64 queries, 8 MMA warps> - one CTA of 384 threads per SM - and <32 queries, 4 MMA warps> - two CTAs of 256
// threads per SM, whose set-up, kappa panel and epilogue overlap the other CTA's DMMA loop.
template <int O, int TQT, int TWARPS>
__global__ void __launch_bounds__(TWARPS * 32 + 128, TWARPS == 4 ? 2 : 1) k_kmp_predict_tab(const __grid_constant__ CUtensorMap tmap, PredictArgs p, TabSmem lay, const int* sched_tab) {
    constexpr int TTHR = TWARPS * 32;          // MMA threads
    constexpr int TALL = TTHR + 128;           // + producer warpgroup (setmaxnreg works on groups of four warps)
    constexpr int TMAXST = TWST / TWARPS;
    constexpr int NJ = TQT / 8;                // n-tiles of 8 queries per warp tile
    constexpr int NT = O * (O + 1) / 2 + O;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int b = blockIdx.y, m0 = blockIdx.x * TQT;
    const int I = p.I;
    const int N = p.n_pts ? p.n_pts[b] : p.n_max;
    const int Ns = comp_stride(N);
    const int n = O * Ns;
    const int ldp = p.ldp;
    const int n_sys_max = O * comp_stride(p.n_max);

    unsigned char* base = smem_align1k(smem_raw);
    double* rings = reinterpret_cast<double*>(base + lay.rings);        // [TWARPS][TST][16 x 32 swizzled]
    double* panel = reinterpret_cast<double*>(base + lay.panel);        // [TQT][ldp], columns in k-step order
    double* racc = reinterpret_cast<double*>(base + lay.racc);          // [TWARPS][NT][TQT]
    double* wsh = reinterpret_cast<double*>(base + lay.wsh);            // [n_sys_max]
    double* sqs = reinterpret_cast<double*>(base + lay.sqs);            // [TQT][I]
    double* etab = reinterpret_cast<double*>(base + lay.etab);          // [32] table of exp_tab
    int* sched = reinterpret_cast<int*>(base + lay.sched);                // this N's row of the schedule table
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + lay.bars);      // full[TWARPS][TST], empty[TWARPS][TST]
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + TWARPS * TST);

    unsigned* kmask = reinterpret_cast<unsigned*>(sched + TROW);          // bit q of word q / 32: the columns 8 q .. 8 q + 7 of the kappa panel carry data
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) kmask[i] = 0u;
        for (int i = 0; i < TWARPS * TST; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    // ---------------------------------------------------------------- prologue: ALL twelve warps (the register split
    //      between the roles is made afterwards, so the producer warps' lanes work on the kappa panel too)
    const double* s = p.s + (size_t)b * p.n_max * I;
    const double* sq = p.sq + (size_t)b * p.stride_sq;
    const double* w = p.w + (size_t)b * p.ldw;
    // scalar inputs, N <= 224: every lane keeps "its" FILL_J stored points in registers from here on - the loads are in flight
    // while the queries, w and the schedule arrive, and both the box mask and the kappa panel are evaluated from them
    constexpr int FILL_J = 7, FILL_WARPS = TALL / 32;
    const bool s_in_regs = I == 1 && ldp - 4 <= 32 * FILL_J;
    double sj[FILL_J];
#pragma unroll
    for (int u = 0; u < FILL_J; ++u) sj[u] = s_in_regs ? s[min(lane + 32 * u, N - 1)] : 0.0;
    for (int idx = tid; idx < TQT * I; idx += TALL) {
        const int m = idx / I;
        sqs[idx] = (m0 + m < p.M) ? sq[(size_t)(m0 + m) * I + (idx - m * I)] : 0.0;
    }
    for (int idx = tid; idx < n_sys_max; idx += TALL) wsh[idx] = idx < n ? w[idx] : 0.0;
    for (int idx = tid; idx < TWARPS * NT * TQT; idx += TALL) racc[idx] = 0.0;
    fill_exp2_table(etab, tid, TALL);
    for (int idx = tid; idx < TROW; idx += TALL) sched[idx] = sched_tab[(size_t)N * TROW + idx];
    __syncthreads();          // the schedule and the queries are complete
    // RBF locality.  k(q, s_j) = exp(-sigma_f |q - s_j|^2) < kskip (2^-68 by default) as soon as s_j is further
    // than d = sqrt(-ln(kskip) / sigma_f) from the bounding box of the tile's live queries.  A box of 8 stored points that are
    // all that far contributes at most 2^-68 |Linv row|_1 <= 2^-60 max|Linv row| to an entry of V (n <= 256 terms) - 1/128 of
    // ONE rounding error of a retained term of that size, far inside the error of the dense sum itself - so it is dead: its
    // stages are never requested, its panel entries never read.  Bit q of kmask[q / 32]: box q (columns 8 q .. 8 q + 7) is live.
    const bool keep_all = (p.dbg & 8192) != 0 || !(p.kskip > 0.0);
    if (!keep_all && s_in_regs) {
        // one warp: the interval of the tile's live queries, then the 4 boxes of each group of 32 stored points from one ballot
        if (warp == 0) {
            double lo = INFINITY, hi = -INFINITY;
            bool odd = false;           // a NaN / infinite query: the reference propagates it, so this tile keeps every box
            for (int m = lane; m < TQT; m += 32)
                if (m0 + m < p.M) { const double v = sqs[m]; lo = fmin(lo, v); hi = fmax(hi, v); odd = odd || !(fabs(v) < INFINITY); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
            unsigned word = 0u;
#pragma unroll
            for (int u = 0; u < FILL_J; ++u) {
                const double dd = fmax(fmax(lo - sj[u], sj[u] - hi), 0.0);
                const unsigned near = __ballot_sync(0xffffffffu, lane + 32 * u < N && !(dd * dd > p.kskip_d2));      // (a NaN coordinate keeps its box)
#pragma unroll
                for (int k = 0; k < 4; ++k) if ((near >> (8 * k)) & 0xffu) word |= 1u << (4 * u + k);
            }
            if (__any_sync(0xffffffffu, odd)) { if (lane < 8) kmask[lane] = ~0u; }
            else if (lane == 0) kmask[0] = word;
        }
        __syncthreads();
    } else if (!keep_all) {
        double* qbox = reinterpret_cast<double*>(kmask + 8) ;          // [I] lower and [I] upper corner of the live queries
        if (warp == 0) {
            bool odd = false;           // a NaN / infinite query coordinate: the reference propagates it, so this tile keeps every box
            for (int i = 0; i < I; ++i) {
                double lo = INFINITY, hi = -INFINITY;
                for (int m = lane; m < TQT; m += 32)
                    if (m0 + m < p.M) { const double v = sqs[m * I + i]; lo = fmin(lo, v); hi = fmax(hi, v); odd = odd || !(fabs(v) < INFINITY); }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
                if (lane == 0) { qbox[i] = lo; qbox[I + i] = hi; }
            }
            if (__any_sync(0xffffffffu, odd) && lane < 8) kmask[lane] = ~0u;
        }
        __syncthreads();
        const double lim2 = p.kskip_d2;      // squared distance beyond which k < kskip
        for (int j0 = warp * 32; j0 < N; j0 += TALL) {
            const int j = j0 + lane;
            double d2 = 0.0;
            if (j < N)
                for (int i = 0; i < I; ++i) {
                    const double v = s[(size_t)j * I + i];
                    const double dd = fmax(fmax(qbox[i] - v, v - qbox[I + i]), 0.0);       // distance to the box along axis i
                    d2 = fma(dd, dd, d2);
                }
            const unsigned near = __ballot_sync(0xffffffffu, j < N && !(d2 > lim2));      // (a NaN coordinate keeps its box)
            if (lane < 4 && ((near >> (8 * lane)) & 0xffu)) atomicOr(kmask + (j0 >> 8), 1u << (((j0 >> 3) + lane) & 31));
        }
        __syncthreads();
    }
    auto live_boxes = [&](int kcol, int nst) -> unsigned {      // live boxes among the nst / 2 boxes with data of the stage at column kcol
        const unsigned have = (1u << (nst >> 1)) - 1u;
        if (keep_all) return have;
        const int q = kcol >> 3;                                  // kcol is a multiple of 32: the four boxes sit in one word
        return (kmask[q >> 5] >> (q & 31)) & have;
    };
    // Every MMA warp compacts ITS stage list under the tile's mask once: {d0 | live boxes << 20, d1} of the live stages only, the
    // chunk-end flag moved to the last live stage of each chunk - so the streaming loop neither visits dead stages nor looks the
    // mask up between two stages.  (The producer lanes walk the full list: they run ahead of the MMA warps anyway.)
    int* clist = reinterpret_cast<int*>(base + lay.clist) + (warp < TWARPS ? warp : 0) * TMAXST * 2;
    int n_live = 0, ks_all = 0;
    if (warp < TWARPS) {
        const int* cd = sched + 8 + warp * TMAXST * 2;
        const int cnt = sched[warp];
        for (int i0 = 0; i0 < cnt; i0 += 32) {
            const int i = i0 + lane;
            int d0 = 0, d1 = 0;
            unsigned lv = 0u;
            if (i < cnt) { d0 = cd[2 * i]; d1 = cd[2 * i + 1]; lv = live_boxes(d1 & 0xffff, d0 >> 16); ks_all += d0 >> 16; }
            const unsigned bal = __ballot_sync(0xffffffffu, lv != 0u);
            if (lv != 0u) {
                const int pos = n_live + __popc(bal & ((1u << lane) - 1u));
                clist[2 * pos] = (d0 & ~(1 << 13)) | (int)(lv << 20);
                clist[2 * pos + 1] = d1;
            }
            n_live += __popc(bal);
        }
        __syncwarp();
        for (int j = lane; j < n_live; j += 32)
            if (j == n_live - 1 || ((clist[2 * (j + 1)] ^ clist[2 * j]) & 0xfff) != 0) atomicOr(clist + 2 * j, 1 << 13);
        __syncwarp();
    }
    // producer cursors (lanes 0, 1 of the four producer warps: two streams each); the first tiles are requested now
    // so that they land while the panel is evaluated
    constexpr int PL = TWARPS / 4;             // producer lanes per warp of the producer warpgroup
    const bool is_producer = warp >= TWARPS && lane < PL;
    const int cw = is_producer ? (warp - TWARPS) * PL + lane : 0;
    const int* pdesc = sched + 8 + cw * TMAXST * 2;
    const int pcount = is_producer ? sched[cw] : 0;
    int pk = 0;
    int pstage = 0;
    uint32_t ppar = 0;
    bool pwrapped = false;
    auto produce_one = [&](unsigned live) {          // only the live boxes of the stage travel
        const uint32_t bi = 8 * (cw * TST + pstage);
        const uint32_t bar = full0 + bi;
        mbar_expect_tx(bar, __popc(live) * (TRC * 64));
        const uint32_t dst = smem_u32(rings + (size_t)(cw * TST + pstage) * TSTAGE_ELEMS);
        const int d0 = pdesc[2 * pk], d1 = pdesc[2 * pk + 1];
        const int col = (d1 >> 16) + (d1 & 0xffff), prow = d0 & 0xfff;
#pragma unroll
        for (int q = 0; q < TQK / 8; ++q)
            if (live & (1u << q)) tma_load_box(dst + q * (TRC * 64), &tmap, col + 8 * q, prow, b, bar);
        ++pk;
        if (++pstage == TST) { pstage = 0; ppar ^= 1; pwrapped = true; }
    };
    if (is_producer) {
        for (int t = 0; t < TST && pk < pcount;) {
            const unsigned lv = live_boxes(pdesc[2 * pk + 1] & 0xffff, pdesc[2 * pk] >> 16);
            if (lv == 0u) { ++pk; continue; }
            produce_one(lv);
            ++t;
        }
    }
    // kappa panel (the B operand), evaluated while the first Linv tiles are in flight: warp-strided over the queries,
    // lane-strided over the stored points; column j goes to position (j & ~7) | swap_bits_1_2(j & 7)  (k-step order).
    // A dependent FP64 operation costs 40-60 cycles here and an exp is a chain of ~25 of them, so every lane keeps
    // FILL_J independent chains in flight (seven columns of one query; fourteen were measured slower); indices are
    // clamped and the results selected, and the exponential itself is branch-free (exp_nonpos), so the chains interleave.
    auto fill = [&](auto one_dim) {         // the I == 1 test is hoisted: a branch inside the chain body would serialise the chains
        constexpr bool ONE = decltype(one_dim)::value;
        if (ONE && s_in_regs && !keep_all && TQT == 4 * FILL_WARPS) {
            // 32-query tile under the box mask: a warp owns four queries, and only the groups of 32 stored points (four boxes)
            // with a live box are evaluated - the panel entries of dead boxes are never read.  Four independent chains per group;
            // the FP64 pipe this shares with the partner CTA's DMMAs is what bounds the panel, so the work saved is time saved.
            const int pj = (lane & ~6) | ((lane & 2) << 1) | ((lane & 4) >> 1);
            const unsigned kword = kmask[0];
            double qv[4];
            bool lq[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { const int m = warp + k * FILL_WARPS; qv[k] = sqs[m]; lq[k] = m0 + m < p.M; }
            if (p.dbg & 4) return;
#pragma unroll
            for (int u = 0; u < FILL_J; ++u) {
                if (((kword >> (4 * u)) & 0xfu) == 0u) continue;          // warp-uniform
                double e[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { const double d = fabs(qv[k] - sj[u]); e[k] = exp_tab(-p.sigma_f * (d * d), etab); }
                const int j = lane + 32 * u;
                if (j < ldp - 4) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) panel[(size_t)(warp + k * FILL_WARPS) * ldp + pj + 32 * u] = (j < N && lq[k]) ? e[k] : 0.0;
                }
            }
            return;
        }
        if (ONE && s_in_regs) {
            // scalar inputs, N <= 224: a lane owns the same FILL_J stored points for every query, so they are loaded once, and
            // two queries are evaluated per batch: 2 FILL_J independent chains per lane (the panel is latency bound: each warp
            // walks ~1000 dependent-chain instructions whatever the other warps do)
            const int pj = (lane & ~6) | ((lane & 2) << 1) | ((lane & 4) >> 1);
            for (int ma = warp; ma < TQT && !(p.dbg & 4); ma += 2 * FILL_WARPS) {
                const int mb = ma + FILL_WARPS;
                const double qa = sqs[ma], qb = sqs[min(mb, TQT - 1)];
                double ea[FILL_J], eb[FILL_J];
#pragma unroll
                for (int u = 0; u < FILL_J; ++u) {
                    const double da = fabs(qa - sj[u]), db = fabs(qb - sj[u]);
                    ea[u] = exp_tab(-p.sigma_f * (da * da), etab);
                    eb[u] = exp_tab(-p.sigma_f * (db * db), etab);
                }
                const bool la = m0 + ma < p.M, lb = mb < TQT && m0 + mb < p.M;
#pragma unroll
                for (int u = 0; u < FILL_J; ++u) {
                    const int j = lane + 32 * u;
                    if (j < ldp - 4) {
                        const double va = (j < N && la) ? ea[u] : 0.0, vb = (j < N && lb) ? eb[u] : 0.0;
                        panel[(size_t)ma * ldp + pj + 32 * u] = va;
                        if (mb < TQT) panel[(size_t)mb * ldp + pj + 32 * u] = vb;
                    }
                }
            }
            return;
        }
        for (int m = warp; m < TQT && !(p.dbg & 4); m += FILL_WARPS) {
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
                    if (j < ldp - 4) {
                        const double v = (j < N && live) ? e[u] : 0.0;
                        panel[(size_t)m * ldp + pj + 32 * u] = v;
                    }
                }
            }
        }
    };
    if (I == 1) fill(std::true_type{}); else fill(std::false_type{});
    __syncthreads();          // panel, wsh, racc are complete

    if (warp >= TWARPS) {
        // ---- producers: one lane per MMA warp runs that warp's tile cursor and issues its TMA boxes
        // register split: one CTA of 384 threads at 168 -> 40 / 232;  two CTAs of 256 threads at 128 -> 48 / 208
        if (TWARPS == 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n"); else asm volatile("setmaxnreg.dec.sync.aligned.u32 48;\n");
        if (lane >= PL) return;
        while (pk < pcount) {
            const unsigned lv = live_boxes(pdesc[2 * pk + 1] & 0xffff, pdesc[2 * pk] >> 16);
            if (lv == 0u) { ++pk; continue; }
            const uint32_t bi = 8 * (cw * TST + pstage);
            if (PL == 1) { if (pwrapped) mbar_wait(empty0 + bi, ppar ^ 1); }      // one stream per producer warp: the hardware-suspended wait
            else if (pwrapped && !mbar_test(empty0 + bi, ppar ^ 1)) { __nanosleep(100); continue; }   // polling must not eat the MMA warps' issue slots
            produce_one(lv);
        }
        return;
    }
    if (TWARPS == 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n"); else asm volatile("setmaxnreg.inc.sync.aligned.u32 208;\n");

    // ---------------------------------------------------------------- per-warp streaming loop
    const double* ring = rings + (size_t)warp * TST * TSTAGE_ELEMS;
    double* my_racc = racc + (size_t)warp * NT * TQT;
    // per-lane constants of the swizzled 8-column boxes: element offset of (row g, k-step h) and the logical column
    const int sw = (g >> 1) & 3, c0 = (t4 >> 1) << 1;
    const int o0 = g * 8 + ((c0 ^ sw) << 1) + (t4 & 1);
    const int o1 = g * 8 + (((c0 | 1) ^ sw) << 1) + (t4 & 1);
    const int cm0 = (t4 & 1) | ((t4 & 2) << 1);          // k-step 0 uses the columns {0,1,4,5} of a box, k-step 1 {2,3,6,7}
    const int* cdesc = clist;
    const int ccount = n_live;
    int stage = 0;
    uint32_t par = 0;
    double acc0[2][NJ][2], acc1[2][NJ][2];
    zero_acc(acc0);
    zero_acc(acc1);
    // chunk complete: reduce the V tile over its 16 rows straight from the accumulators (products, then a 3-step reduce-scatter
    // over the 8 row lanes: lane g ends up owning the column block j == g).  Sums t: O == 1: 0 = q00, 1 = mu0;
    // O == 2: 0 = q00, 1 = q01, 2 = q11, 3 = mu0, 4 = mu1.  HAS_V1 = false (first component block, V_1 = 0): the sums with V_1 are skipped.
    auto reduce = [&](int row0, auto has_v1) {
        constexpr bool HAS_V1 = decltype(has_v1)::value;
        double wv[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) { const int r = row0 + i * 8 + g; wv[i] = r < n ? wsh[r] : 0.0; }
        const bool g2 = (lane & 16) != 0, g1 = (lane & 8) != 0, g0 = (lane & 4) != 0;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const bool needs_v1 = O == 2 && (t == 1 || t == 2 || t == 4);
            if (needs_v1 && !HAS_V1) continue;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double val[NJ];
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    double sum = 0.0;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const double v0 = acc0[i][j][e];
                        const double v1 = HAS_V1 ? acc1[i][j][e] : 0.0;
                        if (O == 1) sum += t == 0 ? v0 * v0 : v0 * wv[i];
                        else sum += t == 0 ? v0 * v0 : (t == 1 ? v0 * v1 : (t == 2 ? v1 * v1 : (t == 3 ? v0 * wv[i] : v1 * wv[i])));
                    }
                    val[j] = sum;
                }
                if constexpr (NJ == 8) {
                    double s4[4], s2[2];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const double keep = g2 ? val[jj + 4] : val[jj];
                        const double send = g2 ? val[jj] : val[jj + 4];
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
                    my_racc[t * TQT + g * 8 + 2 * t4 + e] += tot;
                } else {
                    // four column blocks: two scatter steps (lane pair (g2, g1) owns block 2 g2 + g1), then a plain sum over g0
                    double s2[2];
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        const double keep = g2 ? val[jj + 2] : val[jj];
                        const double send = g2 ? val[jj] : val[jj + 2];
                        s2[jj] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
                    const double keep = g1 ? s2[1] : s2[0];
                    const double send = g1 ? s2[0] : s2[1];
                    double tot = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    tot += __shfl_xor_sync(0xffffffffu, tot, 4);
                    if (!g0) my_racc[t * TQT + (g >> 1) * 8 + 2 * t4 + e] += tot;
                }
            }
        }
    };
    int ks_run = 0;
    for (int ck = 0; ck < ccount; ++ck) {
        const int d0 = cdesc[2 * ck], d1 = cdesc[2 * ck + 1];
        const int c_r0 = d0 & 0xfff, c_kcol = d1 & 0xffff, c_col0 = d1 >> 16;
        const unsigned live = (unsigned)(d0 >> 20) & 0xfu;
        ks_run += 2 * __popc(live);
        const uint32_t bi = 8 * (warp * TST + stage);
        mbar_wait(full0 + bi, par);
        const double* tile = ring + stage * TSTAGE_ELEMS;
        const double* Bs = panel + c_kcol;
        // last valid point index of the two rows this lane carries in this segment (-1: row outside the system)
        int lim[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = c_r0 + i * 8 + g;
            lim[i] = r < n ? min(N - 1, r - c_col0) - c_kcol : -1;
        }
        const int nst = (p.dbg & 1) ? 0 : ((d0 >> 16) & 0xf);                  // k-steps of 4 that carry data (whole boxes)
        const bool first = (O == 1 || !(d0 & (1 << 12)));
        auto kstep = [&](int st, double (&acc)[2][NJ][2]) {
            const int q = st >> 1, h = st & 1;
            const int jl = 8 * q + 2 * h + cm0;                             // logical column of this lane inside the tile
            double a[2], bf[NJ];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double v = tile[q * (TRC * 8) + i * 64 + (h ? o1 : o0)];
                a[i] = jl <= lim[i] ? v : 0.0;
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) bf[j] = Bs[(size_t)(j * 8 + g) * ldp + 4 * st + t4];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884(acc[i][j], a[i], bf[j]);
        };
        if (live == 0xfu && nst == 8) {
            if (first) {
#pragma unroll
                for (int st = 0; st < 8; ++st) kstep(st, acc0);
            } else {
#pragma unroll
                for (int st = 0; st < 8; ++st) kstep(st, acc1);
            }
        } else {                                          // short last k-block / partly live stage: only the live boxes (two k-steps
            const unsigned lv = (p.dbg & 1) ? 0u : live;  // each), every box with static offsets (live is already cut to the boxes with data)
            if (first) {
#pragma unroll
                for (int q = 0; q < 4; ++q) if (lv & (1u << q)) { kstep(2 * q, acc0); kstep(2 * q + 1, acc0); }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) if (lv & (1u << q)) { kstep(2 * q, acc1); kstep(2 * q + 1, acc1); }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + bi);
        if (++stage == TST) { stage = 0; par ^= 1; }
        if ((d0 & (1 << 13)) && !(p.dbg & 8)) {
            // chunk complete (its last live stage): reduce the V tile over its 16 rows straight from the accumulators
            if (O == 2 && (d0 & (1 << 14))) reduce(c_r0, std::true_type{});
            else reduce(c_r0, std::false_type{});
            zero_acc(acc0);
            zero_acc(acc1);
        }
    }
    if (p.dbg & 32768) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ks_all += __shfl_xor_sync(0xffffffffu, ks_all, o);
        if (lane == 0) { atomicAdd(&g_pstats[0], (unsigned long long)ks_run); atomicAdd(&g_pstats[1], (unsigned long long)ks_all); }
    }
    if (TWARPS == 8) asm volatile("bar.sync 1, 256;\n" ::: "memory"); else asm volatile("bar.sync 1, 128;\n" ::: "memory");

    // ---------------------------------------------------------------- outputs (reference layout, query last)
    double* xi = p.xi + (size_t)b * O * p.M;
    double* sg_out = p.sigma_out ? p.sigma_out + (size_t)b * O * O * p.M : nullptr;
    constexpr int NPAIR = O * (O + 1) / 2;
    for (int idx = tid; idx < NT * TQT; idx += TTHR) {
        const int t = idx / TQT, m = idx - t * TQT;
        if (m0 + m >= p.M) continue;
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < TWARPS; ++q) v += racc[((size_t)q * NT + t) * TQT + m];
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

int debug_predict_stats(long long* out, int reset) {
    unsigned long long v[2];
    KMPX_TRY(check_cuda(cudaMemcpyFromSymbol(v, g_pstats, sizeof(v)), "cudaMemcpyFromSymbol"));
    out[0] = (long long)v[0]; out[1] = (long long)v[1];
    if (reset) { unsigned long long z[2] = {0, 0}; KMPX_TRY(check_cuda(cudaMemcpyToSymbol(g_pstats, z, sizeof(z)), "cudaMemcpyToSymbol")); }
    return 0;
}

// The shape of the CTA: kmpx_predict_tile_queries(32 | 64) or, by default, 32 queries (two CTAs per SM) when kappa boxes are skipped
// and 64 (one CTA per SM) when everything is kept - a function of the process-wide options only, never of the batch.
static bool tab_small_shape(const PredictArgs& p) {
    if (g_predict_tile_queries == 32) return true;
    if (g_predict_tile_queries == 64) return false;
    return p.kskip > 0.0 && !(g_host_dbg & 8192);
}

bool predict_tab_applicable(const PredictArgs& p) {
    if (g_host_dbg & 4096) return false;
    if (p.time_driven || p.general || p.O > 2) return false;
    if ((p.lda & 1) || (p.batch > 1 && (p.strideA & 1)) || (reinterpret_cast<uintptr_t>(p.F) & 15)) return false;   // TMA
    if ((long long)p.O * comp_stride(p.n_max) > 0xfff || p.n_max > 1024) return false;                                // descriptor fields, 32-bit k-block mask
    const int ldp = (int)round_up(p.n_max, TQK) + 4;
    const bool sm = tab_small_shape(p);
    if (tab_smem_layout(p.O, p.I, p.n_max, ldp, sm ? 32 : 64, sm ? 4 : 8).total > (size_t)(sm ? 110 : 226) * 1024) return false;
    const int* tab = nullptr;
    if (tab_sched_table(p.O, p.n_max, sm ? 4 : 8, &tab) != 0) return false;
    return tab != nullptr;
}

template <int O, int QTT, int NW>
static int launch_tab(const CUtensorMap& map, const PredictArgs& p, const TabSmem& lay, const int* tab, cudaStream_t st) {
    static SmemCache set;
    KMPX_TRY(ensure_dyn_smem(k_kmp_predict_tab<O, QTT, NW>, lay.total, set, "k_kmp_predict_tab"));
    dim3 grid(ceil_div(p.M, QTT), p.batch);
    k_kmp_predict_tab<O, QTT, NW><<<grid, NW * 32 + 128, lay.total, st>>>(map, p, lay, tab);
    return check_launch("k_kmp_predict_tab");
}

int launch_predict_tab(PredictArgs& p, cudaStream_t st) {
    p.ldp = (int)round_up(p.n_max, TQK) + 4;      // 4 (mod 16) and long enough for the last k-block of a row
    p.resident = 1;
    p.kskip_d2 = p.kskip > 0.0 ? -std::log(p.kskip) / p.sigma_f * (1.0 + 1e-9) : INFINITY;     // squared distance beyond which k < kskip
    const bool sm = tab_small_shape(p);
    const TabSmem lay = tab_smem_layout(p.O, p.I, p.n_max, p.ldp, sm ? 32 : 64, sm ? 4 : 8);
    const int n_sys_max = p.O * comp_stride(p.n_max);
    CUtensorMap map;
    KMPX_TRY(make_batch_map(&map, p.F, p.lda, p.strideA, n_sys_max, p.batch, TRC));
    const int* tab = nullptr;
    KMPX_TRY(tab_sched_table(p.O, p.n_max, sm ? 4 : 8, &tab));
    if (sm) return p.O == 1 ? launch_tab<1, 32, 4>(map, p, lay, tab, st) : launch_tab<2, 32, 4>(map, p, lay, tab, st);
    return p.O == 1 ? launch_tab<1, 64, 8>(map, p, lay, tab, st) : launch_tab<2, 64, 8>(map, p, lay, tab, st);
}

}  // namespace kmpx

extern "C" int kmpx_debug_predict_stats(long long* out_host, int reset) { return kmpx::debug_predict_stats(out_host, reset); }

// Host only (no device call): the stage schedule of the table-driven predict kernel for one (O, N) and `warps` MMA warps (4 or 8).
// row_out[0 .. warps): stages per warp;  row_out[8 + 2 (w * (768 / 2 / warps) + i)], [... + 1]: descriptor {d0, d1} of stage i of warp w
// (d0 = r0 | a << 12 | chunk_end << 13 | has_v1 << 14 | k-steps << 16;  d1 = k offset | first column of the block << 16).
// Returns the ints per row (776), 0 when a warp would need more stages than the table holds, < 0 on a bad argument.
extern "C" int kmpx_debug_predict_schedule(int O, int N, int warps, int* row_out, int row_cap) {
    using namespace kmpx;
    if (O < 1 || O > 2 || N < 1 || N > 1024) return fail_arg(1, "O must be 1 or 2, N in [1, 1024]");
    if (warps != 4 && warps != 8) return fail_arg(3, "warps must be 4 or 8");
    if (!row_out || row_cap < TROW) return fail_arg(4, "row_out must hold 776 ints");
    std::vector<int> row(TROW, 0);
    const bool ok = tab_schedule_host(O, N, row.data(), warps);
    std::copy(row.begin(), row.end(), row_out);
    return ok ? TROW : 0;
}
