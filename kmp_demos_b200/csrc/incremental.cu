// incremental adaptation.  KMP.set_waypoint refits the whole system for every via-point set
// (KMP.py:91); for a batch of problems that share a base trajectory the refit is replaced by an exact low-rank
// correction of the BASE solution:
//   * a via point that overwrites a stored point (KMP.py:80-84) first DELETES that point's O unknowns;
//   * every via point is then APPENDED as a new point (bordered system), whether it replaced one or not.
// With G = A_b^-1 (base), d = deleted unknowns, Cb = kernel columns of the appended points (rows of deleted points
// zeroed), H = (G_dd)^-1, W = [G[:,d], G Cb] and, per query, v = k_b(s*), u = v W = [p, q]:
//   N1 = H (G Cb)_d            S = D - (Cb^T G Cb - (G Cb)_d^T H (G Cb)_d)        z = k_a - q + p N1
//   k A'^-1 k^T = v G v^T - p H p^T + z S^-1 z^T
//   k A'^-1 Y'  = v G y_b - p H alpha_d + z S^-1 (y_a - Cb^T alpha + (G Cb)_d^T H alpha_d)
// (block-inverse / Schur identities; mathematically the direct solution, verified against it to <= 1e-9).
// v G v^T and v G y_b are the base predictions, shared by all problems of a base.  Work per problem drops from
// n^3/3 + O^2 Ns^2 M to O(n^2 + r n M), r = number of touched unknowns <= 2 O P.
//
// Kernels: k_inc_alpha (alpha = G y per base), k_inc_kappa (kappa[base][m][j], shared by the problems of a base),
// k_inc_setup (one CTA per problem: classification of the via points exactly as k_waypoint_insert, W, the small
// matrices; in the grouped form the appended columns of W come from one GEMM per base and component instead), k_inc_predict (one thread per query, CTA = 128 queries of one problem: u = kappa^T W with W broadcast from
// shared memory, per-query algebra in registers, outputs in the reference layouts).  Limits: O <= 2, P <= 2 (r <= 8), plain kernel.
#include "kmpx_internal.cuh"

namespace kmpx {

constexpr int IMAXO = 2, IMAXP = 2, IR = IMAXO * IMAXP;    // at most IR deleted and IR appended unknowns
constexpr int IREC = 64;                                    // doubles per problem record: H[16] N1[16] Sinv[16] a_p[4] a_z[4]
constexpr int IIREC = 8;                                    // ints per problem: nd, napp, del_idx[2], app_via[2]
// W keeps eight FIXED column slots per problem so that the appended columns of all problems of a base can be produced
// by one GEMM with a uniform row stride:  deleted (k, bc) -> slot 4k + bc,  appended (q, bc) -> slot 4q + 2 + bc.
__host__ __device__ inline int slot_del(int k, int bc) { return 4 * k + bc; }
__host__ __device__ inline int slot_app(int q, int bc) { return 4 * q + 2 + bc; }
constexpr int IQT = 128;                                    // threads per CTA of k_inc_predict
constexpr int IQPT = 2;                                     // queries per thread (four were measured slower: 208 registers, half the CTAs)

// alpha_sys[base][r] = sum_c G[r][c] y_sys[c]   (system = component-major order, padding rows carry y = 0)
__global__ void __launch_bounds__(256) k_inc_alpha(const double* G, int ldg, long long strideG, const double* y, int N0, int O,
                                                   double* alpha_sys) {
    extern __shared__ double ys[];
    const int l = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Ns = comp_stride(N0), n0 = O * Ns;
    for (int r = tid; r < n0; r += blockDim.x) { const int a = r / Ns, i = r - a * Ns; ys[r] = i < N0 ? y[((size_t)l * N0 + i) * O + a] : 0.0; }
    __syncthreads();
    const double* Gl = G + (size_t)l * strideG;
    for (int r = warp; r < n0; r += 8) {
        double acc = 0.0;
        for (int c = lane; c < n0; c += 32) acc += Gl[(size_t)r * ldg + c] * ys[c];
        acc = warp_sum(acc);
        if (lane == 0) alpha_sys[(size_t)l * n0 + r] = acc;
    }
}

// kappa[base][j][m] = k(s*_m, s_j)   (stored point major: a thread per query reads it coalesced)
__global__ void __launch_bounds__(256) k_inc_kappa(const double* s, int N0, int I, const double* sq, int M, double sigma_f, double* kappa) {
    const int l = blockIdx.y;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)M * N0) return;
    const int j = (int)(idx / M), m = (int)(idx - (long long)j * M);
    kappa[((size_t)l * N0 + j) * M + m] = rbf(sq + (size_t)m * I, s + ((size_t)l * N0 + j) * I, I, sigma_f, 0.0, 0.0);
}

// r x r inverse (r <= 4) of a symmetric positive definite matrix by Gauss-Jordan, one thread
__device__ inline void small_inverse(const double (&A)[IR][IR], int r, double (&X)[IR][IR]) {
    double a[IR][IR];
    for (int i = 0; i < IR; ++i)
        for (int j = 0; j < IR; ++j) { a[i][j] = (i < r && j < r) ? A[i][j] : (i == j ? 1.0 : 0.0); X[i][j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < IR; ++c) {
        const double piv = 1.0 / a[c][c];
        for (int j = 0; j < IR; ++j) { a[c][j] *= piv; X[c][j] *= piv; }
        for (int i = 0; i < IR; ++i) {
            if (i == c) continue;
            const double f = a[i][c];
            for (int j = 0; j < IR; ++j) { a[i][j] -= f * a[c][j]; X[i][j] -= f * X[c][j]; }
        }
    }
}

struct IncSetupArgs {
    const double *s_base, *G, *alpha_sys; int ldg; long long strideG; int N0, I, O;
    const int* base_of; const double *via_s, *via_y, *via_sigma; int P; double tol, sigma_f, lam; int batch;
    double* W; double* rec; int* irec;
    double* kcg; int mode;      // mode 0: everything; 1: classification + kernel columns -> kcg [b][IMAXP][Ns]; 2: the rest (W appended columns given)
};

__global__ void __launch_bounds__(128) k_inc_setup(IncSetupArgs p) {
    extern __shared__ double sm[];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N0 = p.N0, I = p.I, O = p.O, Ns = comp_stride(N0), n0 = O * Ns;
    const int base = p.base_of ? p.base_of[b] : 0;
    const double* sb = p.s_base + (size_t)base * N0 * I;
    const double* Gb = p.G + (size_t)base * p.strideG;
    const double* al = p.alpha_sys + (size_t)base * n0;
    double* kc = sm;                          // [IMAXP][N0]
    double* ctgc = kc + IMAXP * N0;           // [IR][IR]   Cb^T G Cb
    double* cta = ctgc + IR * IR;             // [IR]       Cb^T alpha
    __shared__ double sh_d[128];
    __shared__ int sh_i[128];
    __shared__ int nd, napp, del_idx[IMAXP], del_owner[IMAXP], app_via[IMAXP];
    if (tid == 0) { nd = 0; napp = 0; }
    __syncthreads();
    if (p.mode == 2) {           // classification and kernel columns were produced by the mode-1 pass
        if (tid == 0) {
            const int* ir = p.irec + (size_t)b * IIREC;
            nd = ir[0]; napp = ir[1];
            for (int k = 0; k < IMAXP; ++k) { del_idx[k] = ir[2 + k]; app_via[k] = ir[4 + k]; }
        }
        for (int idx = tid; idx < IMAXP * N0; idx += blockDim.x) kc[idx] = p.kcg[((size_t)b * IMAXP + idx / N0) * Ns + idx % N0];
        __syncthreads();
    }
    // ---- classification of the via points: nearest stored input of the BASE points, strict '< tol', first minimum
    //      (KMP.py:74-80, same scan as k_waypoint_insert); a later via point that hits the same stored point
    //      supersedes the earlier one
    for (int j = 0; j < (p.mode == 2 ? 0 : p.P); ++j) {
        const double* vs = p.via_s + ((size_t)b * p.P + j) * I;
        double best = INFINITY; int besti = -1;
        for (int i = tid; i < N0; i += blockDim.x) {
            // the scan sees the database as the earlier via points of this call left it (KMP.py:80-84 overwrites in
            // place): a stored point that was replaced carries the input of the via point that replaced it
            const double* si = sb + (size_t)i * I;
            for (int k = 0; k < nd; ++k)
                if (del_idx[k] == i) si = p.via_s + ((size_t)b * p.P + app_via[del_owner[k]]) * I;
            double acc = 0.0;
            for (int t = 0; t < I; ++t) { const double d = si[t] - vs[t]; acc += d * d; }
            const double dist = (I == 1) ? fabs(si[0] - vs[0]) : sqrt(acc);
            if (dist < best) { best = dist; besti = i; }
        }
        sh_d[tid] = best; sh_i[tid] = besti;
        __syncthreads();
        if (tid == 0) {
            double bd = INFINITY; int bi = -1;
            for (int t = 0; t < (int)blockDim.x; ++t) {
                if (sh_i[t] < 0) continue;
                if (sh_d[t] < bd || (sh_d[t] == bd && sh_i[t] < bi)) { bd = sh_d[t]; bi = sh_i[t]; }
            }
            if (bd < p.tol) {
                int k = 0;
                while (k < nd && del_idx[k] != bi) ++k;
                if (k < nd) {                    // the stored point was already replaced by an earlier via point: drop that one
                    const int pos = del_owner[k];
                    for (int q = pos; q + 1 < napp; ++q) app_via[q] = app_via[q + 1];
                    --napp;
                    for (int k2 = 0; k2 < nd; ++k2) if (del_owner[k2] > pos) --del_owner[k2];
                    del_owner[k] = napp;
                } else { del_idx[nd] = bi; del_owner[nd] = napp; ++nd; }
            }
            app_via[napp++] = j;
        }
        __syncthreads();
    }
    const int r1 = O * nd, r2 = O * napp;
    // ---- kernel columns of the appended points against the kept base points
    if (p.mode != 2) {
        for (int idx = tid; idx < napp * N0; idx += blockDim.x) {
            const int q = idx / N0, j = idx - q * N0;
            const double* vs = p.via_s + ((size_t)b * p.P + app_via[q]) * I;
            bool gone = false;
            for (int k = 0; k < nd; ++k) gone |= (del_idx[k] == j);
            kc[q * N0 + j] = gone ? 0.0 : rbf(sb + (size_t)j * I, vs, I, p.sigma_f, 0.0, 0.0);
        }
        __syncthreads();
    }
    if (p.mode == 1) {           // hand the kernel columns (zero rows for absent via points) and the classification over
        for (int idx = tid; idx < IMAXP * Ns; idx += blockDim.x) {
            const int q = idx / Ns, j = idx - q * Ns;
            p.kcg[((size_t)b * IMAXP + q) * Ns + j] = (q < napp && j < N0) ? kc[q * N0 + j] : 0.0;
        }
        if (tid == 0) {
            int* ir = p.irec + (size_t)b * IIREC;
            ir[0] = nd; ir[1] = napp;
            for (int k = 0; k < IMAXP; ++k) { ir[2 + k] = k < nd ? del_idx[k] : -1; ir[4 + k] = k < napp ? app_via[k] : -1; }
        }
        return;
    }
    // ---- W: deleted columns = rows of the symmetric G; appended columns = G Cb
    double* W = p.W + (size_t)b * 2 * IR * n0;
    for (int idx = tid; idx < r1 * n0; idx += blockDim.x) {
        const int t = idx / n0, row = idx - t * n0;
        const int k = t / O, bc = t - k * O;
        W[(size_t)slot_del(k, bc) * n0 + row] = Gb[(size_t)(bc * Ns + del_idx[k]) * p.ldg + row];
    }
    // (G is symmetric: column `col` of the product is accumulated by one thread that walks DOWN the rows of G, so a warp
    //  reads 256 contiguous bytes per step and there is no cross-lane reduction; the kernel column is broadcast from
    //  shared memory)
    for (int col = tid; col < (p.mode == 2 ? 0 : n0); col += blockDim.x) {
        double acc[IMAXP][IMAXO];
        for (int q = 0; q < IMAXP; ++q) for (int bc = 0; bc < IMAXO; ++bc) acc[q][bc] = 0.0;
        for (int bc = 0; bc < O; ++bc) {
            const double* gcol = Gb + (size_t)bc * Ns * p.ldg + col;
            int j = 0;
            for (; j + 16 <= N0; j += 16) {          // sixteen loads in flight per thread: the loop is L2-latency bound
                double gv[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) gv[e] = __ldg(gcol + (size_t)(j + e) * p.ldg);
#pragma unroll
                for (int e = 0; e < 16; ++e)
#pragma unroll
                    for (int q = 0; q < IMAXP; ++q) if (q < napp) acc[q][bc] = fma(gv[e], kc[q * N0 + j + e], acc[q][bc]);
            }
            for (; j < N0; ++j) {
                const double gv = gcol[(size_t)j * p.ldg];
                for (int q = 0; q < napp; ++q) acc[q][bc] = fma(gv, kc[q * N0 + j], acc[q][bc]);
            }
        }
        for (int q = 0; q < napp; ++q)
            for (int bc = 0; bc < O; ++bc) W[(size_t)slot_app(q, bc) * n0 + col] = acc[q][bc];
    }
    __syncthreads();          // W of this problem is complete (CTA scope)
    // ---- Cb^T (G Cb) and Cb^T alpha: dot products over the stored points, one warp each
    for (int e = warp; e < r2 * r2 + r2; e += 4) {
        double acc = 0.0;
        if (e < r2 * r2) {
            const int t = e / r2, t2 = e - t * r2;               // t = (q, bc) row, t2 = (q', bc') column
            const int q = t / O, bc = t - q * O;
            for (int j = lane; j < N0; j += 32) acc += kc[q * N0 + j] * W[(size_t)slot_app(t2 / O, t2 % O) * n0 + bc * Ns + j];
            acc = warp_sum(acc);
            if (lane == 0) ctgc[t * IR + t2] = acc;
        } else {
            const int t = e - r2 * r2, q = t / O, bc = t - q * O;
            for (int j = lane; j < N0; j += 32) acc += kc[q * N0 + j] * al[bc * Ns + j];
            acc = warp_sum(acc);
            if (lane == 0) cta[t] = acc;
        }
    }
    __syncthreads();
    if (tid == 0) {
        double Gdd[IR][IR], H[IR][IR], GCd[IR][IR], N1[IR][IR], S[IR][IR], Sinv[IR][IR], ad[IR], ya[IR], hv[IR];
        for (int t = 0; t < IR; ++t) for (int t2 = 0; t2 < IR; ++t2) { Gdd[t][t2] = 0.0; GCd[t][t2] = 0.0; N1[t][t2] = 0.0; S[t][t2] = 0.0; H[t][t2] = 0.0; }
        for (int t = 0; t < r1; ++t) {
            const int k = t / O, bc = t - k * O, row = bc * Ns + del_idx[k];
            ad[t] = al[row];
            for (int t2 = 0; t2 < r1; ++t2) { const int k2 = t2 / O, bc2 = t2 - k2 * O; Gdd[t][t2] = Gb[(size_t)row * p.ldg + bc2 * Ns + del_idx[k2]]; }
            for (int t2 = 0; t2 < r2; ++t2) GCd[t][t2] = W[(size_t)slot_app(t2 / O, t2 % O) * n0 + row];
        }
        if (r1 > 0) small_inverse(Gdd, r1, H);
        for (int t = 0; t < r1; ++t)
            for (int t2 = 0; t2 < r2; ++t2) { double a = 0.0; for (int u = 0; u < r1; ++u) a += H[t][u] * GCd[u][t2]; N1[t][t2] = a; }
        for (int t = 0; t < r1; ++t) { double a = 0.0; for (int u = 0; u < r1; ++u) a += H[t][u] * ad[u]; hv[t] = a; }     // H alpha_d
        // D: kernel between the appended points + lambda * Sigma on the diagonal blocks
        for (int t = 0; t < r2; ++t) {
            const int q = t / O, bc = t - q * O;
            const double* vq = p.via_s + ((size_t)b * p.P + app_via[q]) * I;
            ya[t] = p.via_y[((size_t)b * p.P + app_via[q]) * O + bc];
            for (int t2 = 0; t2 < r2; ++t2) {
                const int q2 = t2 / O, bc2 = t2 - q2 * O;
                const double* vq2 = p.via_s + ((size_t)b * p.P + app_via[q2]) * I;
                double dv = (bc == bc2) ? rbf(vq, vq2, I, p.sigma_f, 0.0, 0.0) : 0.0;
                if (q == q2) dv += p.lam * p.via_sigma[((size_t)b * p.P + app_via[q]) * O * O + bc * O + bc2];
                double corr = 0.0;
                for (int u = 0; u < r1; ++u) corr += GCd[u][t] * N1[u][t2];          // (G Cb)_d^T H (G Cb)_d
                S[t][t2] = dv - ctgc[t * IR + t2] + corr;
            }
        }
        small_inverse(S, r2, Sinv);
        double* rec = p.rec + (size_t)b * IREC;
        for (int t = 0; t < IR; ++t)
            for (int t2 = 0; t2 < IR; ++t2) {
                rec[t * IR + t2] = (t < r1 && t2 < r1) ? H[t][t2] : 0.0;
                rec[16 + t * IR + t2] = (t < r1 && t2 < r2) ? N1[t][t2] : 0.0;
                rec[32 + t * IR + t2] = (t < r2 && t2 < r2) ? Sinv[t][t2] : 0.0;
            }
        for (int t = 0; t < IR; ++t) rec[48 + t] = t < r1 ? -hv[t] : 0.0;                // a_p = -H alpha_d
        for (int t = 0; t < IR; ++t) {                                                   // a_z = Sinv (y_a - rho)
            double a = 0.0;
            for (int t2 = 0; t2 < r2; ++t2) {
                double rho = cta[t2];
                for (int u = 0; u < r1; ++u) rho -= GCd[u][t2] * hv[u];
                a += Sinv[t][t2] * (ya[t2] - rho);
            }
            rec[52 + t] = t < r2 ? a : 0.0;
        }
        int* ir = p.irec + (size_t)b * IIREC;
        ir[0] = nd; ir[1] = napp;
        for (int k = 0; k < IMAXP; ++k) { ir[2 + k] = k < nd ? del_idx[k] : -1; ir[4 + k] = k < napp ? app_via[k] : -1; }
    }
}

struct IncPredictArgs {
    const double *kappa, *W, *rec; const int* irec; const int* base_of;
    const double *via_s, *sq; int P, I, O, N0, M, batch; double sigma_f, alpha_kmp;
    const double *xi_base, *sigma_base;      // [base][O][M], [base][O][O][M]
    double *xi, *sigma_out;
};

// per-query algebra of k_inc_predict (u = v W for this query, in registers)
__device__ __forceinline__ void inc_combine(const IncPredictArgs& p, const double (&u)[IMAXO * 2 * IR], const double* rec, const int* ir,
                                            int b, int base, int mg, int r1, int r2, int napp) {
    const int O = p.O, I = p.I, M = p.M;
    const double* H = rec, *N1 = rec + 16, *Sinv = rec + 32, *a_p = rec + 48, *a_z = rec + 52;
    double z[IMAXO][IR];
#pragma unroll
    for (int a = 0; a < IMAXO; ++a)
#pragma unroll
        for (int t2 = 0; t2 < IR; ++t2) z[a][t2] = 0.0;
    double ka[IMAXP];
#pragma unroll
    for (int q = 0; q < IMAXP; ++q)
        ka[q] = q < napp ? rbf(p.sq + (size_t)mg * I, p.via_s + ((size_t)b * p.P + ir[4 + q]) * I, I, p.sigma_f, 0.0, 0.0) : 0.0;
    // compact views of the fixed column slots (no dynamic register indexing): up[a][t] deleted unknown t = k*O + bc,
    // uq[a][t2] appended unknown t2 = q*O + bc
    double up[IMAXO][IR], uq[IMAXO][IR];
#pragma unroll
    for (int a = 0; a < IMAXO; ++a)
#pragma unroll
        for (int t = 0; t < IR; ++t) {
            const int sd = slot_del(t / O, t % O), sa = slot_app(t / O, t % O);
            double vd = 0.0, va = 0.0;
#pragma unroll
            for (int c = 0; c < 2 * IR; ++c) { if (c == sd) vd = u[a * 2 * IR + c]; if (c == sa) va = u[a * 2 * IR + c]; }
            up[a][t] = vd; uq[a][t] = va;
        }
#pragma unroll
    for (int a = 0; a < IMAXO; ++a)
#pragma unroll
        for (int t2 = 0; t2 < IR; ++t2) {
            if (a < O && t2 < r2) {
                const int q = t2 / O, bc = t2 - q * O;
                double kav = 0.0;
#pragma unroll
                for (int qq = 0; qq < IMAXP; ++qq) if (qq == q && bc == a) kav = ka[qq];
                double v = kav - uq[a][t2];
#pragma unroll
                for (int t = 0; t < IR; ++t) if (t < r1) v = fma(up[a][t], N1[t * IR + t2], v);
                z[a][t2] = v;
            }
        }
#pragma unroll
    for (int a = 0; a < IMAXO; ++a) {
        if (a >= O) continue;
        double mean = p.xi_base[((size_t)base * O + a) * M + mg];
#pragma unroll
        for (int t = 0; t < IR; ++t) if (t < r1) mean = fma(up[a][t], a_p[t], mean);
#pragma unroll
        for (int t2 = 0; t2 < IR; ++t2) mean = fma(z[a][t2], a_z[t2], mean);
        p.xi[((size_t)b * O + a) * M + mg] = mean;
        if (p.sigma_out) {
#pragma unroll
            for (int a2 = 0; a2 < IMAXO; ++a2) {
                if (a2 >= O) continue;
                double php = 0.0, zsz = 0.0;
#pragma unroll
                for (int t = 0; t < IR; ++t) {
                    double h = 0.0, g = 0.0;
#pragma unroll
                    for (int t2 = 0; t2 < IR; ++t2) {
                        if (t < r1 && t2 < r1) h = fma(H[t * IR + t2], up[a2][t2], h);
                        g = fma(Sinv[t * IR + t2], z[a2][t2], g);
                    }
                    if (t < r1) php = fma(up[a][t], h, php);
                    zsz = fma(z[a][t], g, zsz);
                }
                p.sigma_out[(((size_t)b * O + a) * O + a2) * M + mg] =
                    p.sigma_base[(((size_t)base * O + a) * O + a2) * M + mg] + p.alpha_kmp * (php - zsz);
            }
        }
    }
}

__global__ void __launch_bounds__(IQT) k_inc_predict(IncPredictArgs p) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x;
    const int b = blockIdx.y, mq0 = blockIdx.x * (IQPT * IQT) + tid;                   // IQPT queries per thread: mq0 + q * IQT
    const int N0 = p.N0, O = p.O, Ns = comp_stride(N0), n0 = O * Ns, M = p.M;
    const int base = p.base_of ? p.base_of[b] : 0;
    constexpr int NC = IMAXO * 2 * IR;                           // 16 columns per stored point: [component][touched unknown]
    double* Wt = sm;                                            // [N0][NC]   Wt[j][a*8 + t] = W[t][(a, j)]
    double* rec = Wt + (size_t)N0 * NC;                         // [IREC]
    const int* ir = p.irec + (size_t)b * IIREC;
    const int nd = ir[0], napp = ir[1], r1 = O * nd, r2 = O * napp;
    const double* Wg = p.W + (size_t)b * 2 * IR * n0;
    for (int idx = tid; idx < N0 * NC; idx += IQT) {
        const int j = idx / NC, c = idx - j * NC, a = c / (2 * IR), t = c - a * 2 * IR;
        const bool used = a < O && (t & 1) < O && ((t & 2) ? (t >> 2) < napp : (t >> 2) < nd);
        Wt[idx] = used ? Wg[(size_t)t * n0 + a * Ns + j] : 0.0;
    }
    if (tid < IREC) rec[tid] = p.rec[(size_t)b * IREC + tid];
    __syncthreads();
    if (mq0 >= M) return;
    // ---- u[a][t] = sum_j kappa[j][m] W[t][(a, j)]: IQPT queries per thread, so that every 16-byte broadcast load of W
    //      feeds 2 * IQPT FMAs, IQPT x 16
    //      independent accumulators, kappa read coalesced (stored point major)
    double u[IQPT][NC];
    const double* kp[IQPT];
#pragma unroll
    for (int q = 0; q < IQPT; ++q) {
#pragma unroll
        for (int c = 0; c < NC; ++c) u[q][c] = 0.0;
        kp[q] = p.kappa + (size_t)base * N0 * M + min(mq0 + q * IQT, M - 1);
    }
    for (int j = 0; j < N0; ++j) {
        double kv[IQPT];
#pragma unroll
        for (int q = 0; q < IQPT; ++q) kv[q] = __ldg(kp[q] + (size_t)j * M);
        const double2* wr = reinterpret_cast<const double2*>(Wt + (size_t)j * NC);
#pragma unroll
        for (int c = 0; c < NC / 2; ++c) {
            const double2 w2 = wr[c];
#pragma unroll
            for (int q = 0; q < IQPT; ++q) { u[q][2 * c] = fma(kv[q], w2.x, u[q][2 * c]); u[q][2 * c + 1] = fma(kv[q], w2.y, u[q][2 * c + 1]); }
        }
    }
#pragma unroll
    for (int q = 0; q < IQPT; ++q)
        if (mq0 + q * IQT < M) inc_combine(p, u[q], rec, ir, b, base, mq0 + q * IQT, r1, r2, napp);
}

}  // namespace kmpx

using namespace kmpx;

static SmemCache g_inc_setup_smem, g_inc_predict_smem;    // one per kernel: both entry points launch the same two kernels

extern "C" long long kmpx_inc_workspace_doubles(int N0, int O, int batch) {
    const long long n0 = (long long)O * comp_stride(N0);
    const long long d = (long long)batch * (2 * IR * n0 + IREC) + ((long long)batch * IIREC + 1) / 2;
    return (d + 1) & ~1LL;       // even: what follows the record block stays 16-byte aligned
}

extern "C" int kmpx_inc_base(const double* G, int ldg, long long strideG, const double* y_base, const double* s_base, int N0, int I,
                             int O, int n_bases, const double* sq, int M, double sigma_f, double* alpha_sys, double* kappa,
                             void* stream) {
    if (!G || !y_base || !s_base || !sq || !alpha_sys || !kappa) return fail_arg(1, "null pointer");
    if (O < 1 || O > IMAXO) return fail_arg(8, "incremental adaptation supports O <= 2");
    if (N0 <= 0 || I <= 0 || n_bases <= 0 || M <= 0) return fail_arg(6, "sizes must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    const int n0 = O * comp_stride(N0);
    k_inc_alpha<<<n_bases, 256, n0 * sizeof(double), st>>>(G, ldg, strideG, y_base, N0, O, alpha_sys);
    KMPX_TRY(check_launch("k_inc_alpha"));
    dim3 grid(ceil_div((long long)M * N0, 256), n_bases);
    k_inc_kappa<<<grid, 256, 0, st>>>(s_base, N0, I, sq, M, sigma_f, kappa);
    return check_launch("k_inc_kappa");
}

// Grouped form: the problems [0, batch) are sorted by base; group g holds group_count[g] consecutive problems of base
// group_base[g] (host arrays).  The appended columns G Cb of a whole group come from one DMMA GEMM per component
// (rows = (problem, via point), K = stored points) instead of one pass over G per problem.
extern "C" long long kmpx_inc_grouped_workspace_doubles(int N0, int O, int batch) {
    return kmpx_inc_workspace_doubles(N0, O, batch) + (long long)batch * IMAXP * comp_stride(N0);
}

extern "C" int kmpx_inc_adapt_grouped(const double* s_base, const double* G, int ldg, long long strideG, const double* alpha_sys,
                                      const double* kappa, const double* xi_base, const double* sigma_base, int N0, int I, int O,
                                      const int* base_of, const int* group_base, const int* group_count, int n_groups,
                                      const double* via_s, const double* via_y, const double* via_sigma, int P,
                                      double tol, double sigma_f, double lam, double alpha_kmp, int batch, const double* sq, int M,
                                      double* ws, long long ws_doubles, double* xi, double* sigma_out, void* stream) {
    if (!s_base || !G || !alpha_sys || !kappa || !xi_base || !via_s || !via_y || !via_sigma || !sq || !ws || !xi || !base_of)
        return fail_arg(1, "null pointer");
    if (!group_base || !group_count || n_groups <= 0) return fail_arg(13, "group description missing");
    if (sigma_out && !sigma_base) return fail_arg(8, "sigma_base is required for the covariance");
    if (O < 1 || O > IMAXO) return fail_arg(11, "incremental adaptation supports O <= 2");
    if (P < 1 || P > IMAXP) return fail_arg(19, "incremental adaptation supports P <= 2 via points per problem");
    if (batch <= 0 || M <= 0 || N0 <= 0 || I <= 0) return fail_arg(24, "sizes must be positive");
    if ((ldg & 1) || (strideG & 1) || (reinterpret_cast<uintptr_t>(G) & 15)) return fail_arg(2, "G must be 16-byte aligned with even strides");
    if (ws_doubles < kmpx_inc_grouped_workspace_doubles(N0, O, batch)) return fail_arg(28, "workspace too small (kmpx_inc_grouped_workspace_doubles)");
    long long total = 0;
    for (int g = 0; g < n_groups; ++g) total += group_count[g];
    if (total != batch) return fail_arg(15, "group counts do not add up to the batch");
    cudaStream_t st = (cudaStream_t)stream;
    const int Ns = comp_stride(N0), n0 = O * Ns;
    double* W = ws;
    double* rec = W + (size_t)batch * 2 * IR * n0;
    int* irec = reinterpret_cast<int*>(rec + (size_t)batch * IREC);
    double* kcg = ws + kmpx_inc_workspace_doubles(N0, O, batch);
    IncSetupArgs a{s_base, G, alpha_sys, ldg, strideG, N0, I, O, base_of, via_s, via_y, via_sigma, P, tol, sigma_f, lam, batch, W, rec, irec, kcg, 1};
    const size_t sm1 = ((size_t)IMAXP * N0 + IR * IR + IR) * sizeof(double);
    KMPX_TRY(ensure_dyn_smem(k_inc_setup, sm1, g_inc_setup_smem, "k_inc_setup"));
    k_inc_setup<<<batch, 128, sm1, st>>>(a);                                   // classification + kernel columns
    KMPX_TRY(check_launch("k_inc_setup"));
    long long off = 0;
    for (int g = 0; g < n_groups; ++g) {
        const int cnt = group_count[g];
        if (cnt <= 0) continue;
        const double* Gb = G + (size_t)group_base[g] * strideG;
        for (int bc = 0; bc < O; ++bc) {
            // rows (problem, via point) x columns (system index):  W[.., slot_app(q, bc), :] = kc(problem, q) * G[(bc, :), :]
            GemmArgs ga{IMAXP * cnt, n0, N0, 1.0, kcg + (size_t)off * IMAXP * Ns, Ns, 0, Gb + (size_t)bc * Ns * ldg, ldg, 0, 0.0,
                        W + (size_t)off * 2 * IR * n0 + (size_t)slot_app(0, bc) * n0, 4 * n0, 0, 0};
            KMPX_TRY(launch_dgemm(0, 0, ga, 1, st));
        }
        off += cnt;
    }
    a.mode = 2;
    k_inc_setup<<<batch, 128, sm1, st>>>(a);                                   // deleted columns, small matrices
    KMPX_TRY(check_launch("k_inc_setup"));
    IncPredictArgs q{kappa, W, rec, irec, base_of, via_s, sq, P, I, O, N0, M, batch, sigma_f, alpha_kmp, xi_base, sigma_base, xi, sigma_out};
    const size_t sm2 = ((size_t)N0 * IMAXO * 2 * IR + IREC) * sizeof(double);
    KMPX_TRY(ensure_dyn_smem(k_inc_predict, sm2, g_inc_predict_smem, "k_inc_predict"));
    dim3 grid(ceil_div(M, IQPT * IQT), batch);
    k_inc_predict<<<grid, IQT, sm2, st>>>(q);
    return check_launch("k_inc_predict");
}

extern "C" int kmpx_inc_adapt(const double* s_base, const double* G, int ldg, long long strideG, const double* alpha_sys,
                              const double* kappa, const double* xi_base, const double* sigma_base, int N0, int I, int O,
                              const int* base_of, const double* via_s, const double* via_y, const double* via_sigma, int P,
                              double tol, double sigma_f, double lam, double alpha_kmp, int batch, const double* sq, int M,
                              double* ws, long long ws_doubles, double* xi, double* sigma_out, void* stream) {
    if (!s_base || !G || !alpha_sys || !kappa || !xi_base || !via_s || !via_y || !via_sigma || !sq || !ws || !xi)
        return fail_arg(1, "null pointer");
    if (sigma_out && !sigma_base) return fail_arg(8, "sigma_base is required for the covariance");
    if (O < 1 || O > IMAXO) return fail_arg(11, "incremental adaptation supports O <= 2");
    if (P < 1 || P > IMAXP) return fail_arg(16, "incremental adaptation supports P <= 2 via points per problem");
    if (batch <= 0 || M <= 0 || N0 <= 0 || I <= 0) return fail_arg(21, "sizes must be positive");
    if (ws_doubles < kmpx_inc_workspace_doubles(N0, O, batch)) return fail_arg(25, "workspace too small (kmpx_inc_workspace_doubles)");
    cudaStream_t st = (cudaStream_t)stream;
    const int n0 = O * comp_stride(N0);
    double* W = ws;
    double* rec = W + (size_t)batch * 2 * IR * n0;
    int* irec = reinterpret_cast<int*>(rec + (size_t)batch * IREC);
    IncSetupArgs a{s_base, G, alpha_sys, ldg, strideG, N0, I, O, base_of, via_s, via_y, via_sigma, P, tol, sigma_f, lam, batch, W, rec, irec};
    const size_t sm1 = ((size_t)IMAXP * N0 + IR * IR + IR) * sizeof(double);
    KMPX_TRY(ensure_dyn_smem(k_inc_setup, sm1, g_inc_setup_smem, "k_inc_setup"));
    k_inc_setup<<<batch, 128, sm1, st>>>(a);
    KMPX_TRY(check_launch("k_inc_setup"));
    IncPredictArgs q{kappa, W, rec, irec, base_of, via_s, sq, P, I, O, N0, M, batch, sigma_f, alpha_kmp, xi_base, sigma_base, xi, sigma_out};
    const size_t sm2 = ((size_t)N0 * IMAXO * 2 * IR + IREC) * sizeof(double);
    KMPX_TRY(ensure_dyn_smem(k_inc_predict, sm2, g_inc_predict_smem, "k_inc_predict"));
    dim3 grid(ceil_div(M, IQPT * IQT), batch);
    k_inc_predict<<<grid, IQT, sm2, st>>>(q);
    return check_launch("k_inc_predict");
}
