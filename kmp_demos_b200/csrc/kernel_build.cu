// K1: kernel-matrix builder  A = K (x) I_O + lambda * blockdiag(Sigma)   (KMP.py:93-128,146-156)
// and the batched set_waypoint database update (KMP.py:72-89).
//
// Roofline: pure HBM write, 8*n^2 bytes per problem (4*n*(n+1) for KMPX_UPLO_LOWER); inputs are
// 8*N*(I+O^2) bytes and stay in L1/L2.  One thread produces two adjacent columns of one row and issues a
// single 16-byte store; a warp covers 512 contiguous bytes of a row, a CTA a 32x64 tile.
#include "common.cuh"
#include "kmpx_internal.cuh"

namespace kmpx {

struct BuildArgs {
    const double* s; const double* sigma; const int* n_pts; double* A;
    int n_max, I, O, lda; long long strideA;
    double sigma_f, lambda; int time_driven, layout, uplo;
};

__device__ __forceinline__ double build_element(const BuildArgs& p, const double* s, const double* sig,
                                                int N, int r, int c) {
    const int O = p.O;
    int a, i, b, j;
    if (p.layout == KMPX_LAYOUT_POINT) { i = r / O; a = r - i * O; j = c / O; b = c - j * O; }
    else {
        const int Ns = comp_stride(N);
        a = r / Ns; i = r - a * Ns; b = c / Ns; j = c - b * Ns;
        if (i >= N || j >= N) return (r == c) ? 1.0 : 0.0;      // identity padding
    }
    double v = 0.0;
    if (!p.time_driven) {
        if (a == b) v = rbf(s + (size_t)i * p.I, s + (size_t)j * p.I, p.I, p.sigma_f, 0.0, 0.0);
    } else {
        const int P = O >> 1;
        const int ca = a / P, cb = b / P;
        if (a - ca * P == b - cb * P) {
            double co[4];
            kernel_coeffs(s + (size_t)i * p.I, s + (size_t)j * p.I, p.I, p.sigma_f, true, co);
            v = co[ca * 2 + cb];
        }
    }
    if (i == j) v += p.lambda * sig[((size_t)i * O + a) * O + b];
    return v;
}

constexpr int kBuildRows = 32, kBuildCols = 64;

__global__ void __launch_bounds__(256) k_kernel_build(BuildArgs p) {
    const int bidx = blockIdx.z;
    const int N = p.n_pts ? p.n_pts[bidx] : p.n_max;
    const int n = sys_size(N, p.O, p.layout);
    const int c0 = blockIdx.x * kBuildCols + (threadIdx.x & 31) * 2;
    const int rbase = blockIdx.y * kBuildRows;
    if (rbase >= n || blockIdx.x * kBuildCols >= n) return;
    if (p.uplo == KMPX_UPLO_LOWER && blockIdx.x * kBuildCols > rbase + kBuildRows - 1) return;
    const double* s = p.s + (size_t)bidx * p.n_max * p.I;
    const double* sig = p.sigma + (size_t)bidx * p.n_max * p.O * p.O;
    double* A = p.A + (size_t)bidx * p.strideA;
    const bool vec_ok = ((p.lda & 1) == 0) && ((p.strideA & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
#pragma unroll 1
    for (int rr = (threadIdx.x >> 5); rr < kBuildRows; rr += 8) {
        const int r = rbase + rr;
        if (r >= n || c0 >= n) continue;
        if (p.uplo == KMPX_UPLO_LOWER && c0 > r) continue;
        const bool second = (c0 + 1 < n) && !(p.uplo == KMPX_UPLO_LOWER && c0 + 1 > r);
        double v0 = build_element(p, s, sig, N, r, c0);
        double v1 = second ? build_element(p, s, sig, N, r, c0 + 1) : 0.0;
        double* dst = A + (size_t)r * p.lda + c0;
        if (second && vec_ok) {
            *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
        } else {
            dst[0] = v0;
            if (second) dst[1] = v1;
        }
    }
}


// Component-major fast path (what the factorisation consumes): a thread owns the point pair (i, j..j+1), evaluates
// the kernel coefficients ONCE and writes them into every (a, b) block - n^2/O^2 exp instead of n^2/O, no integer
// division.  Each store is 16 bytes; for fixed (i, a, b) a warp writes 512 contiguous bytes.
// OC: compile-time number of outputs (0: read p.O) - the loops over the (a, b) blocks unroll and their predicates fold, which matters on
// small systems, where the kernel is bound by instruction issue
template <bool TD, bool BULK, int OC>
__global__ void __launch_bounds__(256) k_kernel_build_comp(BuildArgs p) {
    __shared__ double etab[kExpTabN];
    if (!TD) { fill_exp2_table(etab, threadIdx.x, 256); __syncthreads(); }
    const int bidx = blockIdx.z;
    const int N = p.n_pts ? p.n_pts[bidx] : p.n_max;
    const int Ns = comp_stride(N);
    const int O = OC ? OC : p.O, P = TD ? (O >> 1) : O;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j0 = blockIdx.x * kBuildCols + lane * 2;
    const int ibase = blockIdx.y * kBuildRows;
    if (ibase >= Ns || blockIdx.x * kBuildCols >= Ns) return;
    const bool lower = p.uplo == KMPX_UPLO_LOWER;
    if (lower && O == 1 && blockIdx.x * kBuildCols > ibase + kBuildRows - 1) return;
    const double* s = p.s + (size_t)bidx * p.n_max * p.I;
    const double* sig = p.sigma + (size_t)bidx * p.n_max * O * O;
    double* A = p.A + (size_t)bidx * p.strideA;
    if (j0 >= Ns) return;
    if (!TD && lower && blockIdx.x * kBuildCols > ibase + kBuildRows - 1) {
        // plain kernel, lower-only, tile strictly above the point diagonal: nothing to evaluate - the only entries of the
        // lower triangle in these columns are the zeros of the off-diagonal component blocks (a > b)
        for (int ii = warp; ii < kBuildRows; ii += 8) {
            const int i = ibase + ii;
            if (i >= Ns) break;
            for (int a = 1; a < O; ++a)
                for (int b = 0; b < a; ++b)
                    *reinterpret_cast<double2*>(A + (size_t)(a * Ns + i) * p.lda + b * Ns + j0) = make_double2(0.0, 0.0);
        }
        return;
    }
    // BULK (plain kernel, large systems): the eight coefficients a lane needs for its four rows are evaluated up front as
    // eight interleaved branch-free chains (indices clamped, results selected): cfg 4 full 5.6 -> 6.1 TB/s, lower
    // 3.2 -> 4.7 TB/s.  On small systems the diagonal and ragged tiles dominate and the wasted evaluations cost 20 %,
    // so those keep the per-row form.
    double pre[kBuildRows / 8][2];
    if (BULK) {
#pragma unroll
        for (int r = 0; r < kBuildRows / 8; ++r) {
            const double* si = s + (size_t)min(ibase + warp + 8 * r, N - 1) * p.I;
            pre[r][0] = rbf_tab(si, s + (size_t)min(j0, N - 1) * p.I, p.I, p.sigma_f, etab);
            pre[r][1] = rbf_tab(si, s + (size_t)min(j0 + 1, N - 1) * p.I, p.I, p.sigma_f, etab);
        }
    }
#pragma unroll (BULK ? kBuildRows / 8 : 1)
    for (int r = 0; r < kBuildRows / 8; ++r) {
        const int ii = warp + 8 * r;
        const int i = ibase + ii;
        if (i >= Ns) break;
        double co[2][4];
#pragma unroll
        for (int e = 0; e < 2; ++e) co[e][0] = co[e][1] = co[e][2] = co[e][3] = 0.0;
        if (TD) {
#pragma unroll
            for (int e = 0; e < 2; ++e)
                if (i < N && j0 + e < N) kernel_coeffs(s + (size_t)i * p.I, s + (size_t)(j0 + e) * p.I, p.I, p.sigma_f, true, co[e]);
        } else {
            // plain kernel: both coefficients of the column pair as two interleaved branch-free chains (indices clamped,
            // results selected; the library exp would serialise them and the kernel becomes latency bound on small
            // systems).  Lower-only: coefficients are only used in the diagonal (a == b) blocks, i.e. for j <= i.
            const bool any = i < N && j0 < N && (!lower || j0 <= i);
            if (any) {
                double e0, e1;
                if (BULK) { e0 = pre[r][0]; e1 = pre[r][1]; }
                else {
                    const double* si = s + (size_t)i * p.I;
                    e0 = rbf_tab(si, s + (size_t)j0 * p.I, p.I, p.sigma_f, etab);
                    e1 = rbf_tab(si, s + (size_t)min(j0 + 1, N - 1) * p.I, p.I, p.sigma_f, etab);
                }
                co[0][0] = e0;
                co[1][0] = (j0 + 1 < N && (!lower || j0 + 1 <= i)) ? e1 : 0.0;
            }
        }
        const bool diag0 = (i == j0), diag1 = (i == j0 + 1);
        for (int a = 0; a < O; ++a) {
            double* row = A + (size_t)(a * Ns + i) * p.lda;
            for (int b = 0; b < O; ++b) {
                if (lower && b > a) continue;
                // compile-time O: lambda * Sigma_i[a][b] once per (row, block), a warp-uniform load outside the element loop
                const double lsig = (OC && i < N) ? p.lambda * sig[((size_t)i * O + a) * O + b] : 0.0;
                double v[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const bool dg = e ? diag1 : diag0;
                    double x;
                    if (i >= N || j0 + e >= N) x = (a == b && dg) ? 1.0 : 0.0;           // identity padding
                    else {
                        x = TD ? (((a % P) == (b % P)) ? co[e][(a / P) * 2 + (b / P)] : 0.0) : ((a == b) ? co[e][0] : 0.0);
                        if (OC) x = dg ? x + lsig : x;          // (the same sum: lambda * Sigma is added to the diagonal lane only)
                        else if (dg) x += p.lambda * sig[((size_t)i * O + a) * O + b];
                    }
                    v[e] = x;
                }
                double* dst = row + b * Ns + j0;
                if (lower && a == b) {
                    if (j0 > i) continue;
                    if (j0 + 1 > i) { dst[0] = v[0]; continue; }
                }
                *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// set_waypoint database update: one thread per problem (the search is over <= a few hundred points and
// the via-point count is tiny); a warp-wide copy of the base trajectory precedes it.
struct WaypointArgs {
    const double *s_base, *y_base, *sigma_base; const int* n_base; int n_base_max; const int* base_of;
    const double *via_s, *via_y, *via_sigma; int P; double tol;
    double *s, *y, *sigma; int* n_pts; int n_max, I, O, batch;
};

__global__ void __launch_bounds__(128) k_waypoint_insert(WaypointArgs p) {
    const int b = blockIdx.x;
    if (b >= p.batch) return;
    const int base = p.base_of ? p.base_of[b] : 0;
    const int N0 = p.n_base ? p.n_base[base] : p.n_base_max;
    const int I = p.I, O = p.O;
    const double* sb = p.s_base + (size_t)base * p.n_base_max * I;
    const double* yb = p.y_base + (size_t)base * p.n_base_max * O;
    const double* gb = p.sigma_base + (size_t)base * p.n_base_max * O * O;
    double* s = p.s + (size_t)b * p.n_max * I;
    double* y = p.y + (size_t)b * p.n_max * O;
    double* g = p.sigma + (size_t)b * p.n_max * O * O;
    for (int t = threadIdx.x; t < N0 * I; t += blockDim.x) s[t] = sb[t];
    for (int t = threadIdx.x; t < N0 * O; t += blockDim.x) y[t] = yb[t];
    for (int t = threadIdx.x; t < N0 * O * O; t += blockDim.x) g[t] = gb[t];
    __syncthreads();
    __shared__ double sh_d[128];
    __shared__ int sh_i[128];
    __shared__ int sh_N;
    if (threadIdx.x == 0) sh_N = N0;
    __syncthreads();
    for (int j = 0; j < p.P; ++j) {
        const int N = N0;   // the scan covers the points of the last fit only (self.N, KMP.py:74)
        const double* vs = p.via_s + ((size_t)b * p.P + j) * I;
        // nearest stored input, first minimum wins (KMP.py:75-79: strict '<' while scanning upwards)
        double best = INFINITY; int besti = -1;
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            double acc = 0.0;
            for (int t = 0; t < I; ++t) { double d = s[(size_t)i * I + t] - vs[t]; acc += d * d; }
            double dist = (I == 1) ? fabs(s[i] - vs[0]) : sqrt(acc);
            if (dist < best) { best = dist; besti = i; }
        }
        sh_d[threadIdx.x] = best; sh_i[threadIdx.x] = besti;
        __syncthreads();
        if (threadIdx.x == 0) {
            double bd = INFINITY; int bi = -1;
            for (int t = 0; t < blockDim.x; ++t) {
                if (sh_i[t] < 0) continue;
                if (sh_d[t] < bd || (sh_d[t] == bd && sh_i[t] < bi)) { bd = sh_d[t]; bi = sh_i[t]; }
            }
            int dst;
            if (bd < p.tol) dst = bi;                       // overwrite (KMP.py:80-84)
            else { dst = sh_N; if (dst < p.n_max) sh_N = dst + 1; else dst = -1; }   // append (KMP.py:85-89)
            sh_i[0] = dst;
        }
        __syncthreads();
        const int dst = sh_i[0];
        if (dst >= 0) {
            const double* vy = p.via_y + ((size_t)b * p.P + j) * O;
            const double* vg = p.via_sigma + ((size_t)b * p.P + j) * O * O;
            for (int t = threadIdx.x; t < I; t += blockDim.x) s[(size_t)dst * I + t] = vs[t];
            for (int t = threadIdx.x; t < O; t += blockDim.x) y[(size_t)dst * O + t] = vy[t];
            for (int t = threadIdx.x; t < O * O; t += blockDim.x) g[(size_t)dst * O * O + t] = vg[t];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) p.n_pts[b] = sh_N;
}

}  // namespace kmpx

using namespace kmpx;

extern "C" int kmpx_kernel_build(const double* s, const double* sigma, const int* n_pts, double* A,
                                 int batch, int n_max, int I, int O, int lda, long long strideA,
                                 double sigma_f, double lambda, int time_driven, int layout, int uplo, void* stream) {
    if (!s) return fail_arg(1, "s is NULL");
    if (!sigma) return fail_arg(2, "sigma is NULL");
    if (!A) return fail_arg(4, "A is NULL");
    if (batch <= 0) return fail_arg(5, "batch must be > 0");
    if (batch > 65535) return fail_arg(5, "batch must be <= 65535 per call");
    if (n_max <= 0) return fail_arg(6, "n_max must be > 0");
    if (I <= 0) return fail_arg(7, "I must be > 0");
    if (O <= 0) return fail_arg(8, "O must be > 0");
    if (time_driven && (O & 1)) return fail_arg(8, "time-driven kernel needs an even O (KMP.py:125)");
    const long long n = sys_size(n_max, O, layout);
    if (lda < n) return fail_arg(9, "lda < n_max*O");
    if (batch > 1 && strideA < (long long)lda * n) return fail_arg(10, "strideA too small");
    if (layout != KMPX_LAYOUT_POINT && layout != KMPX_LAYOUT_COMPONENT) return fail_arg(14, "unknown layout");
    BuildArgs p{s, sigma, n_pts, A, n_max, I, O, lda, strideA, sigma_f, lambda, time_driven, layout, uplo};
    const bool vec_ok = ((lda & 1) == 0) && ((strideA & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    if (layout == KMPX_LAYOUT_COMPONENT && vec_ok) {
        const int Ns = comp_stride(n_max);
        dim3 grid(ceil_div(Ns, kBuildCols), ceil_div(Ns, kBuildRows), batch);
        if (grid.y > 65535) return fail_arg(6, "n_max too large for one launch");
        if (time_driven) k_kernel_build_comp<true, false, 0><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
        else if (Ns >= 1024) k_kernel_build_comp<false, true, 0><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
        else if (O == 2 && !(g_host_dbg & 131072)) k_kernel_build_comp<false, false, 2><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
        else if (O == 1 && !(g_host_dbg & 131072)) k_kernel_build_comp<false, false, 1><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
        else k_kernel_build_comp<false, false, 0><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
        return check_launch("k_kernel_build_comp");
    }
    dim3 grid(ceil_div(n, kBuildCols), ceil_div(n, kBuildRows), batch);
    if (grid.y > 65535) return fail_arg(6, "n_max*O too large for one launch");
    k_kernel_build<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("k_kernel_build");
}

extern "C" int kmpx_waypoint_insert(const double* s_base, const double* y_base, const double* sigma_base,
                                    const int* n_base, int n_base_max, const int* base_of,
                                    const double* via_s, const double* via_y, const double* via_sigma, int P, double tol,
                                    double* s, double* y, double* sigma, int* n_pts, int n_max, int I, int O, int batch,
                                    void* stream) {
    if (!s_base || !y_base || !sigma_base) return fail_arg(1, "base trajectory is NULL");
    if (P < 0) return fail_arg(10, "P must be >= 0");
    if (P > 0 && (!via_s || !via_y || !via_sigma)) return fail_arg(7, "via-point arrays are NULL");
    if (!s || !y || !sigma || !n_pts) return fail_arg(12, "output arrays are NULL");
    if (n_max < n_base_max) return fail_arg(16, "n_max < n_base_max");
    if (batch <= 0) return fail_arg(19, "batch must be > 0");
    WaypointArgs p{s_base, y_base, sigma_base, n_base, n_base_max, base_of, via_s, via_y, via_sigma, P, tol,
                   s, y, sigma, n_pts, n_max, I, O, batch};
    k_waypoint_insert<<<batch, 128, 0, (cudaStream_t)stream>>>(p);
    return check_launch("k_waypoint_insert");
}
