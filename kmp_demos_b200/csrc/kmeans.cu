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

// ---------------------------------------------------------------- k-means++ sampling on the device
// sklearn draws a centre candidate as  searchsorted(cumsum(closest), u * pot)  (_kmeans.py:227-231) with numpy's
// SEQUENTIAL fp64 cumulative sum.  A parallel scan rounds differently, so the device works with (nearly) exact
// double-double prefix sums S_i and decides only when the decision is provably the one numpy takes: the sequential sum
// c_i of non-negative terms satisfies |c_i - S_i| <= i u S_i, and the potential numpy multiplies the uniform draw with
// (a BLAS dot product, any summation order) is within n u pot of the exact one.  A draw whose target lies further than
// that from both neighbouring prefix sums has the same index under every rounding; anything closer is FLAGGED and the
// host repeats that single draw with numpy's own expressions.
struct dd { double hi, lo; };
__device__ __forceinline__ dd dd_make(double a) { return dd{a, 0.0}; }
__device__ __forceinline__ dd dd_add_d(dd a, double b) {
    const double s = a.hi + b, bb = s - a.hi;
    const double e = (a.hi - (s - bb)) + (b - bb) + a.lo;
    const double hi = s + e;
    return dd{hi, e - (hi - s)};
}
__device__ __forceinline__ dd dd_add(dd a, dd b) {
    const double s = a.hi + b.hi, bb = s - a.hi;
    const double e = (a.hi - (s - bb)) + (b.hi - bb) + a.lo + b.lo;
    const double hi = s + e;
    return dd{hi, e - (hi - s)};
}
__device__ __forceinline__ dd dd_sub(dd a, dd b) { return dd_add(a, dd{-b.hi, -b.lo}); }
__device__ __forceinline__ bool dd_less(dd a, dd b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
__device__ __forceinline__ dd dd_shfl(dd v, int src) { return dd{__shfl_sync(0xffffffffu, v.hi, src), __shfl_sync(0xffffffffu, v.lo, src)}; }

constexpr int KPS_BS = 4096;         // elements per block sum
constexpr int KPS_THREADS = 256;

// bs[b] = double-double sum of closest[b*4096 .. (b+1)*4096)
__global__ void __launch_bounds__(KPS_THREADS) k_kpp_blocksums(const double* closest, long long n, double* bs) {
    __shared__ double sh_hi[KPS_THREADS], sh_lo[KPS_THREADS];
    const int tid = threadIdx.x;
    const long long base = (long long)blockIdx.x * KPS_BS;
    dd acc = dd_make(0.0);
#pragma unroll 4
    for (int q = 0; q < KPS_BS / KPS_THREADS; ++q) {
        const long long i = base + tid + (long long)q * KPS_THREADS;
        if (i < n) acc = dd_add_d(acc, closest[i]);
    }
    sh_hi[tid] = acc.hi; sh_lo[tid] = acc.lo;
    __syncthreads();
    for (int off = KPS_THREADS / 2; off > 0; off >>= 1) {
        if (tid < off) {
            const dd v = dd_add(dd{sh_hi[tid], sh_lo[tid]}, dd{sh_hi[tid + off], sh_lo[tid + off]});
            sh_hi[tid] = v.hi; sh_lo[tid] = v.lo;
        }
        __syncthreads();
    }
    if (tid == 0) { bs[2 * (size_t)blockIdx.x] = sh_hi[0]; bs[2 * (size_t)blockIdx.x + 1] = sh_lo[0]; }
}

struct KppPickArgs {
    const double* closest; long long n; int nb; double* bs;           // bs: [nb][2] block sums in, exclusive prefixes out
    const double* u; int T; const double* pot;                       // uniform draws (device), current potential (device scalar)
    const double* before; long long n_before, n_total; int last_rank; // dd potential / row count of the ranks in front
    long long* cand; int* flag; double* local_total;                  // outputs
};
// One CTA: exclusive double-double scan of the block sums, then warp t resolves draw t.
__global__ void __launch_bounds__(KPS_THREADS) k_kpp_pick(KppPickArgs p) {
    __shared__ double ph[KPS_THREADS], pl[KPS_THREADS];
    __shared__ double tot_sh[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (p.nb + KPS_THREADS - 1) / KPS_THREADS;
    const int b0 = min(p.nb, tid * per), b1 = min(p.nb, b0 + per);
    dd acc = dd_make(0.0);
    for (int b = b0; b < b1; ++b) acc = dd_add(acc, dd{p.bs[2 * (size_t)b], p.bs[2 * (size_t)b + 1]});
    ph[tid] = acc.hi; pl[tid] = acc.lo;
    __syncthreads();
    if (tid == 0) {
        dd run = dd_make(0.0);
        for (int t = 0; t < KPS_THREADS; ++t) { const dd v{ph[t], pl[t]}; ph[t] = run.hi; pl[t] = run.lo; run = dd_add(run, v); }
        tot_sh[0] = run.hi; tot_sh[1] = run.lo;
        if (p.local_total) { p.local_total[0] = run.hi; p.local_total[1] = run.lo; }
    }
    __syncthreads();
    dd run{ph[tid], pl[tid]};
    for (int b = b0; b < b1; ++b) {
        const dd v{p.bs[2 * (size_t)b], p.bs[2 * (size_t)b + 1]};
        p.bs[2 * (size_t)b] = run.hi; p.bs[2 * (size_t)b + 1] = run.lo;
        run = dd_add(run, v);
    }
    __syncthreads();
    if (warp >= p.T) return;
    const int t = warp;
    const dd total{tot_sh[0], tot_sh[1]};
    const double pot = *p.pot;
    const double r = p.u[t] * pot;                                    // numpy: uniform(size=T) * current_pot
    const dd before = p.before ? dd{p.before[0], p.before[1]} : dd_make(0.0);
    const dd target = dd_sub(dd_make(r), before);                    // position inside this rank's rows
    // rigorous slack: cumulative-sum rounding (<= n u S) + potential rounding (<= n u pot) + the rounding of r itself
    const double slack = 2.0 * ((double)p.n_total + 8.0) * 1.1102230246251565e-16 * fabs(pot);
    long long cand = -1; int flag = 0;
    const bool below = !dd_less(dd_make(0.0), target);               // target <= 0: an earlier rank owns the draw
    const bool above = dd_less(total, target);                       // target > local total: a later rank (or the clip)
    if (below) { if (target.hi > -slack && p.n_before > 0) flag = 1; if (p.n_before == 0) { cand = 0; flag = 1; } }
    else if (above) {
        if (p.last_rank) { cand = p.n - 1; if (dd_sub(target, total).hi <= slack) flag = 1; }     // numpy clips to n-1
        else if (dd_sub(target, total).hi <= slack) flag = 1;
    } else {
        // block with prefix[b] < target <= prefix[b+1]: last b whose exclusive prefix is < target
        int lo_b = 0, hi_b = p.nb - 1;
        while (lo_b < hi_b) {
            const int mid = (lo_b + hi_b + 1) >> 1;
            if (dd_less(dd{p.bs[2 * (size_t)mid], p.bs[2 * (size_t)mid + 1]}, target)) lo_b = mid; else hi_b = mid - 1;
        }
        const long long e0 = (long long)lo_b * KPS_BS;
        constexpr int PER_LANE = KPS_BS / 32;
        dd mine = dd_make(0.0);
        for (int q = 0; q < PER_LANE; ++q) { const long long i = e0 + (long long)lane * PER_LANE + q; if (i < p.n) mine = dd_add_d(mine, p.closest[i]); }
        dd excl = dd{p.bs[2 * (size_t)lo_b], p.bs[2 * (size_t)lo_b + 1]};
        int owner = 31; dd owner_excl = excl;
        bool found = false;
        for (int l = 0; l < 32; ++l) {
            const dd v = dd_shfl(mine, l);
            const dd incl = dd_add(excl, v);
            if (!found && !dd_less(incl, target)) { owner = l; owner_excl = excl; found = true; }
            excl = incl;
        }
        if (!found) { owner = 31; flag = 1; }                        // (only through double-double rounding at a block edge)
        if (lane == owner) {
            dd s = owner_excl;
            long long pick = -1; double m_lo = 0.0, m_hi = 0.0;
            for (int q = 0; q < PER_LANE; ++q) {
                const long long i = e0 + (long long)lane * PER_LANE + q;
                if (i >= p.n) break;
                const dd nxt = dd_add_d(s, p.closest[i]);
                if (!dd_less(nxt, target)) { pick = i; m_lo = dd_sub(target, s).hi; m_hi = dd_sub(nxt, target).hi; break; }
                s = nxt;
            }
            if (pick < 0) { pick = min(p.n - 1, e0 + (long long)lane * PER_LANE + PER_LANE - 1); flag = 1; }
            else if (!(m_lo > slack) || !(m_hi > slack)) flag = 1;
            cand = pick;
        }
        cand = __shfl_sync(0xffffffffu, cand, owner);
        flag = __shfl_sync(0xffffffffu, flag, owner) | flag;
    }
    if (lane == 0) { p.cand[t] = cand; p.flag[t] = flag; }
}

// rows[t] = [X[cand[t]][0..d), 1, n_before + cand[t]] for the draws this rank owns (cand >= 0), zeros otherwise: summed
// over the ranks (all-reduce) every rank holds every candidate row, how many ranks claimed it, and its global index
__global__ void k_kpp_gather_rows(const double* X, int d, const long long* cand, int T, long long n_before, double* rows) {
    const int t = blockIdx.x, q = threadIdx.x;
    if (t >= T || q >= d + 2) return;
    const long long c = cand[t];
    double v = 0.0;
    if (c >= 0) v = q < d ? X[c * d + q] : (q == d ? 1.0 : (double)(n_before + c));
    rows[(size_t)t * (d + 2) + q] = v;
}

// dist[t][i] = min(closest[i], max(0, |c_t|^2 - 2 c_t.x_i + |x_i|^2)) for candidate ROWS (rows [T][ldr]); per-CTA
// double-double partial potentials, combined in CTA order: pot[t] = {hi, lo}
__global__ void __launch_bounds__(KPP_THREADS) k_kpp_candidates_rows(const double* X, const double* x_sq, long long n, int d,
                                                                     const double* rows, int ldr, int T, const double* closest,
                                                                     double* dist, double* part) {
    __shared__ double c_sh[KPP_MAXT * 16];
    __shared__ double csq_sh[KPP_MAXT];
    __shared__ double red[KPP_MAXT][KPP_THREADS / 32][2];
    const int tid = threadIdx.x;
    for (int idx = tid; idx < T * d; idx += KPP_THREADS) { const int t = idx / d; c_sh[idx] = rows[(size_t)t * ldr + (idx - t * d)]; }
    __syncthreads();
    if (tid < T) { double a = 0.0; for (int q = 0; q < d; ++q) a += c_sh[tid * d + q] * c_sh[tid * d + q]; csq_sh[tid] = a; }
    __syncthreads();
    dd pot[KPP_MAXT];
#pragma unroll
    for (int t = 0; t < KPP_MAXT; ++t) pot[t] = dd_make(0.0);
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
                pot[t] = dd_add_d(pot[t], v);
            }
        }
    }
#pragma unroll
    for (int t = 0; t < KPP_MAXT; ++t) {
        dd s = pot[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s = dd_add(s, dd{__shfl_xor_sync(0xffffffffu, s.hi, o), __shfl_xor_sync(0xffffffffu, s.lo, o)});
        if ((tid & 31) == 0) { red[t][tid >> 5][0] = s.hi; red[t][tid >> 5][1] = s.lo; }
    }
    __syncthreads();
    if (tid < T) {
        dd s = dd_make(0.0);
        for (int w = 0; w < KPP_THREADS / 32; ++w) s = dd_add(s, dd{red[tid][w][0], red[tid][w][1]});
        part[((size_t)blockIdx.x * KPP_MAXT + tid) * 2] = s.hi; part[((size_t)blockIdx.x * KPP_MAXT + tid) * 2 + 1] = s.lo;
    }
}
__global__ void k_kpp_reduce_dd(const double* part, int nctas, int T, double* pot) {
    const int t = threadIdx.x;
    if (t < T) {
        dd s = dd_make(0.0);
        for (int c = 0; c < nctas; ++c) s = dd_add(s, dd{part[((size_t)c * KPP_MAXT + t) * 2], part[((size_t)c * KPP_MAXT + t) * 2 + 1]});
        pot[2 * t] = s.hi; pot[2 * t + 1] = s.lo;
    }
}

// The greedy choice of _kmeans.py:245-251 on the device: pots [R][T][2] = per-rank double-double potentials of the T
// candidates; total[t] = sum over the ranks in rank order; best = first minimum (np.argmin) unless force_best >= 0.
// Writes the new potential, the double-double potential of the ranks in front of `rank` for the next draw, the chosen
// row and its global index; *flag = 1 when a DIFFERENT candidate's potential is within the rounding slack of the best
// one (numpy's BLAS sum could order them either way - the host then decides with numpy's own product).
__global__ void k_kpp_choose(const double* pots, int R, int T, int rank, const double* rows, int d, int c, long long n_total,
                             int force_best, double* pot_out, double* before_out, double* centers, long long* center_ids,
                             int* best_out, int* flag) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int best = 0; double bv = 0.0; double tot[KPP_MAXT];
    for (int t = 0; t < T; ++t) {
        dd s = dd_make(0.0);
        for (int r = 0; r < R; ++r) s = dd_add(s, dd{pots[((size_t)r * T + t) * 2], pots[((size_t)r * T + t) * 2 + 1]});
        tot[t] = s.hi;
        if (t == 0 || s.hi < bv) { bv = s.hi; best = t; }
    }
    int fl = 0;
    const double slack = 2.0 * ((double)n_total + 8.0) * 1.1102230246251565e-16 * fabs(bv);
    for (int t = 0; t < T; ++t)
        if (t != best && rows[(size_t)t * (d + 2) + d + 1] != rows[(size_t)best * (d + 2) + d + 1] && tot[t] - bv <= slack) fl = 1;
    if (force_best >= 0) { best = force_best; bv = tot[best]; fl = 0; }
    *pot_out = bv;
    *best_out = best;
    if (flag) *flag = fl;
    dd before = dd_make(0.0);
    for (int r = 0; r < rank; ++r) before = dd_add(before, dd{pots[((size_t)r * T + best) * 2], pots[((size_t)r * T + best) * 2 + 1]});
    before_out[0] = before.hi; before_out[1] = before.lo;
    for (int q = 0; q < d; ++q) centers[(size_t)c * d + q] = rows[(size_t)best * (d + 2) + q];
    center_ids[c] = (long long)rows[(size_t)best * (d + 2) + d + 1];
}
// closest[i] = dist[*best][i]
__global__ void k_kpp_take_row(const double* dist, const int* best, long long n, double* closest) {
    const double* src = dist + (size_t)(*best) * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) closest[i] = src[i];
}

// labels[i] = argmin_j(|c_j|^2 - 2 x_i.c_j), first minimum (_k_means_lloyd.pyx:215-224)
template <int D>      // D = number of features (1..8: loops unrolled, samples in registers), 0 = run-time d <= 16
__global__ void __launch_bounds__(256) k_kmeans_assign(const double* X, long long n, int d_rt, int K, const double* centers,
                                                       const int* labels_prev, int* labels, int* n_changed) {
    const int d = D > 0 ? D : d_rt;
    extern __shared__ double sh[];
    double* c_sh = sh;              // [K][d]
    double* csq = sh + (size_t)K * d;
    for (int idx = threadIdx.x; idx < K * d; idx += blockDim.x) c_sh[idx] = centers[idx];
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) { double a = 0.0; for (int q = 0; q < d; ++q) a += c_sh[k * d + q] * c_sh[k * d + q]; csq[k] = a; }
    __syncthreads();
    int changed = 0;
    // four samples per thread: a centre coordinate is loaded from shared memory once and used four times (with one
    // sample per thread the loop issues one broadcast load per FMA and is bound by the shared-memory pipe).  The
    // arithmetic of a sample is unchanged (same FMA chain over the features, same first-minimum rule).
    constexpr int SPT = 4;
    const long long span = (long long)gridDim.x * blockDim.x;
    for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < n; i0 += SPT * span) {
        double x[SPT][D > 0 ? D : 16];
#pragma unroll
        for (int u = 0; u < SPT; ++u) {
            const long long i = i0 + u * span < n ? i0 + u * span : i0;
#pragma unroll
            for (int q = 0; q < (D > 0 ? D : 16); ++q) if (q < d) x[u][q] = X[i * d + q];
        }
        double best[SPT]; int bl[SPT];
#pragma unroll
        for (int u = 0; u < SPT; ++u) { best[u] = 0.0; bl[u] = 0; }
        for (int k = 0; k < K; ++k) {
            double dot[SPT];
#pragma unroll
            for (int u = 0; u < SPT; ++u) dot[u] = 0.0;
#pragma unroll
            for (int q = 0; q < (D > 0 ? D : 16); ++q) {
                if (q < d) {
                    const double cq = c_sh[k * d + q];
#pragma unroll
                    for (int u = 0; u < SPT; ++u) dot[u] += x[u][q] * cq;
                }
            }
            const double ck = csq[k];
#pragma unroll
            for (int u = 0; u < SPT; ++u) {
                const double v = ck - 2.0 * dot[u];
                if (k == 0 || v < best[u]) { best[u] = v; bl[u] = k; }
            }
        }
#pragma unroll
        for (int u = 0; u < SPT; ++u) {
            const long long i = i0 + u * span;
            if (i < n) {
                labels[i] = bl[u];
                if (labels_prev && labels_prev[i] != bl[u]) ++changed;
            }
        }
    }
    if (n_changed && changed) atomicAdd(n_changed, changed);
}

// One Lloyd iteration's centre update and stopping rules on the device (_kmeans.py:700-737, _k_means_common.pyx
// _average_centers / _center_shift): stats = [count, sum x (d), ...] per cluster (the packed statistics of
// k_gmm_em in hard-label mode, already summed over the ranks), state = {done, n_iter, strict, empty}.
// Once `done` is set the kernel changes nothing, so the host may enqueue iterations in batches and read `state` once
// per batch: the assignment and statistics kernels of the surplus iterations recompute the same values.
__global__ void k_lloyd_update(const double* stats, int K, int d, int F, double* centers, const double* tol_abs,
                               const double* n_changed_total, int* n_changed_local, int max_iter, int* state) {
    __shared__ double shift_sh[1024];
    __shared__ int empty_sh;
    const int k = threadIdx.x;
    if (k == 0) empty_sh = 0;
    __syncthreads();
    if (state[0] || state[3]) { if (k == 0 && n_changed_local) *n_changed_local = 0; return; }
    double sh = 0.0;
    double nw[16];
    if (k < K) {
        const double cnt = stats[(size_t)k * F];
        if (!(cnt > 0.0)) atomicExch(&empty_sh, 1);
        const double alpha = 1.0 / cnt;
        for (int q = 0; q < d; ++q) {
            nw[q] = stats[(size_t)k * F + 1 + q] * alpha;
            const double df = nw[q] - centers[(size_t)k * d + q];
            sh += df * df;
        }
    }
    shift_sh[k] = sh;
    __syncthreads();
    if (empty_sh) {            // an empty cluster: the host relocates (rare), nothing is updated here
        if (k == 0) { state[3] = 1; if (n_changed_local) *n_changed_local = 0; }
        return;
    }
    if (k < K) for (int q = 0; q < d; ++q) centers[(size_t)k * d + q] = nw[q];
    if (k == 0) {
        double tot = 0.0;
        for (int j = 0; j < K; ++j) tot += shift_sh[j];
        state[1] += 1;
        if (*n_changed_total == 0.0) { state[2] = 1; state[0] = 1; }
        else if (tot <= *tol_abs) state[0] = 1;
        else if (state[1] >= max_iter) state[0] = 1;
        if (n_changed_local) *n_changed_local = 0;
    }
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
    const long long blocks = (n + 1023) / 1024;                   // four samples per thread
    const int grid = (int)(blocks < num_sms() * 8 ? (blocks < 1 ? 1 : blocks) : num_sms() * 8);
    const size_t smem = ((size_t)K * d + K) * sizeof(double);
    cudaStream_t st = (cudaStream_t)stream;
#define KMPX_ASSIGN(DD) k_kmeans_assign<DD><<<grid, 256, smem, st>>>(X, n, d, K, centers, labels_prev, labels, n_changed)
    switch (d) {
        case 1: KMPX_ASSIGN(1); break; case 2: KMPX_ASSIGN(2); break; case 3: KMPX_ASSIGN(3); break; case 4: KMPX_ASSIGN(4); break;
        case 5: KMPX_ASSIGN(5); break; case 6: KMPX_ASSIGN(6); break; case 7: KMPX_ASSIGN(7); break; case 8: KMPX_ASSIGN(8); break;
        default: KMPX_ASSIGN(0); break;
    }
#undef KMPX_ASSIGN
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

extern "C" long long kmpx_kpp_sample_workspace_bytes(long long n) {
    const long long nb = (n + KPS_BS - 1) / KPS_BS;
    return (2 * nb + 8) * (long long)sizeof(double);
}

extern "C" int kmpx_kpp_sample(const double* closest, long long n, const double* u, int T, const double* pot,
                               const double* before, long long n_before, long long n_total, int last_rank,
                               long long* cand, int* flag, void* ws, long long ws_bytes, void* stream) {
    if (!closest || !u || !pot || !cand || !flag) return fail_arg(1, "NULL argument");
    if (n <= 0 || n_total < n) return fail_arg(2, "n must be > 0 and <= n_total");
    if (T < 1 || T > KPP_MAXT) return fail_arg(4, "T must be in 1..8");
    if (!ws || ws_bytes < kmpx_kpp_sample_workspace_bytes(n)) return fail_arg(12, "workspace too small (kmpx_kpp_sample_workspace_bytes)");
    const long long nb = (n + KPS_BS - 1) / KPS_BS;
    if (nb > 0x7fffffffLL) return fail_arg(2, "n too large");
    cudaStream_t st = (cudaStream_t)stream;
    k_kpp_blocksums<<<(int)nb, KPS_THREADS, 0, st>>>(closest, n, (double*)ws);
    KMPX_TRY(check_launch("k_kpp_blocksums"));
    KppPickArgs a{closest, n, (int)nb, (double*)ws, u, T, pot, before, n_before, n_total, last_rank, cand, flag, nullptr};
    k_kpp_pick<<<1, KPS_THREADS, 0, st>>>(a);
    return check_launch("k_kpp_pick");
}

extern "C" int kmpx_kpp_gather_rows(const double* X, int d, const long long* cand, int T, long long n_before, double* rows,
                                    void* stream) {
    if (!X || !cand || !rows) return fail_arg(1, "NULL argument");
    if (d < 1 || d > 16) return fail_arg(2, "d must be in 1..16");
    if (T < 1 || T > KPP_MAXT) return fail_arg(4, "T must be in 1..8");
    k_kpp_gather_rows<<<T, 32, 0, (cudaStream_t)stream>>>(X, d, cand, T, n_before, rows);
    return check_launch("k_kpp_gather_rows");
}

extern "C" int kmpx_kpp_candidates_rows(const double* X, const double* x_sq, long long n, int d, const double* rows, int T,
                                        const double* closest, double* dist, double* pot, void* ws, long long ws_bytes,
                                        void* stream) {
    if (!X || !x_sq || !rows) return fail_arg(1, "X / x_sq / rows is NULL");
    if (n <= 0) return fail_arg(3, "n must be > 0");
    if (d < 1 || d > 16) return fail_arg(4, "d must be in 1..16");
    if (T < 1 || T > KPP_MAXT) return fail_arg(6, "T must be in 1..8");
    if (!dist || !pot) return fail_arg(8, "dist / pot is NULL");
    const int grid = (int)((n + KPP_THREADS - 1) / KPP_THREADS < num_sms() * 4 ? (n + KPP_THREADS - 1) / KPP_THREADS : num_sms() * 4);
    if (!ws || ws_bytes < (long long)grid * KPP_MAXT * 2 * (long long)sizeof(double)) return fail_arg(10, "workspace too small (148*4*16 doubles)");
    k_kpp_candidates_rows<<<grid, KPP_THREADS, 0, (cudaStream_t)stream>>>(X, x_sq, n, d, rows, d + 2, T, closest, dist, (double*)ws);
    KMPX_TRY(check_launch("k_kpp_candidates_rows"));
    k_kpp_reduce_dd<<<1, 32, 0, (cudaStream_t)stream>>>((const double*)ws, grid, T, pot);
    return check_launch("k_kpp_reduce_dd");
}

extern "C" int kmpx_kpp_choose(const double* pots, int R, int T, int rank, const double* rows, int d, int c, long long n_total,
                               int force_best, const double* dist, long long n, double* pot_out, double* before_out,
                               double* centers, long long* center_ids, int* best_out, int* flag, double* closest, void* stream) {
    if (!pots || !rows || !pot_out || !before_out || !centers || !center_ids || !best_out) return fail_arg(1, "NULL argument");
    if (R < 1 || rank < 0 || rank >= R) return fail_arg(2, "bad rank / world");
    if (T < 1 || T > KPP_MAXT || force_best >= T) return fail_arg(3, "T must be in 1..8");
    cudaStream_t st = (cudaStream_t)stream;
    k_kpp_choose<<<1, 32, 0, st>>>(pots, R, T, rank, rows, d, c, n_total, force_best, pot_out, before_out, centers, center_ids, best_out, flag);
    KMPX_TRY(check_launch("k_kpp_choose"));
    if (dist && closest) {
        const long long blocks = (n + 255) / 256;
        k_kpp_take_row<<<(int)(blocks < num_sms() * 8 ? blocks : num_sms() * 8), 256, 0, st>>>(dist, best_out, n, closest);
        KMPX_TRY(check_launch("k_kpp_take_row"));
    }
    return 0;
}

extern "C" int kmpx_lloyd_update(const double* stats, int K, int d, double* centers, const double* tol_abs,
                                 const double* n_changed_total, int* n_changed_local, int max_iter, int* state, void* stream) {
    if (!stats || !centers || !tol_abs || !n_changed_total || !state) return fail_arg(1, "NULL argument");
    if (K < 1 || K > 1024) return fail_arg(2, "K must be in 1..1024");
    if (d < 1 || d > 16) return fail_arg(3, "d must be in 1..16");
    const int F = 1 + d + d * (d + 1) / 2;
    k_lloyd_update<<<1, ((K + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(stats, K, d, F, centers, tol_abs, n_changed_total, n_changed_local, max_iter, state);
    return check_launch("k_lloyd_update");
}
