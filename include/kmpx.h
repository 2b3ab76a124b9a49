/*
 * kmpx.h - C ABI of libkmpx.so: the sm_100a (B200) CUDA implementation of the KMP learn-and-adapt
 * hot path of lbusellato/KMP_demos.
 *
 * The reference has no FFI of its own (it is pure Python); the boundary it offers is its Python class
 * API.  Each entry point below replaces the numerical body of one reference method and is what a
 * ctypes binding placed inside that method would call (INTEGRATION.md shows the stubs).  Citations are
 * file:line in /root/reference.
 *
 * Conventions
 *  - every function returns 0 on success, <0 for a bad argument (-k = k-th argument), >0 for a CUDA
 *    error code; kmpx_last_error() returns a thread-local message.  Nothing throws across the boundary.
 *  - all array arguments are DEVICE pointers to caller-owned buffers unless the name ends in _host;
 *    fp64 everywhere, labels int32 (sklearn path) or int64 (reference KMeans path).
 *  - the library allocates nothing persistent; scratch is passed in (`ws`, `ws_bytes`, sized by the
 *    matching *_workspace_bytes function).  No call synchronises the host.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *  - batched problems are ragged: `n_pts[b]` (device int32, may be NULL => n_max for every problem) is
 *    the number of stored reference points of problem b; buffers are strided by the *_max sizes.
 *
 * KMP data layouts on the device ("point-major" = the reference's vec order, KMP.py:151,185-186)
 *    s      [batch][n_max][I]        inputs            (reference keeps (I,N); host transposes)
 *    y      [batch][n_max][O]        outputs mu        (reference (O,N))
 *    sigma  [batch][n_max][O][O]     covariances       (reference (O,O,N))
 *  System matrices A (n_b x n_b, n_b = n_pts[b]*O, row-major, leading dimension lda, batch stride
 *  strideA elements) use one of two index orders:
 *    KMPX_LAYOUT_POINT      row = i*O + a   - the reference's order (KMP.py:151)
 *    KMPX_LAYOUT_COMPONENT  row = a*N_b + i - component-major; what the factorisation / predict
 *                                             kernels consume (every GEMM operand is a contiguous panel)
 */
#ifndef KMPX_H
#define KMPX_H

#ifdef __cplusplus
extern "C" {
#endif

#define KMPX_VERSION 100
#define KMPX_LAYOUT_POINT 0
#define KMPX_LAYOUT_COMPONENT 1
#define KMPX_UPLO_FULL 0
#define KMPX_UPLO_LOWER 1
#define KMPX_SMALL_MAX 512 /* largest system handled by the one-CTA-per-problem kernels */

int kmpx_version(void);
const char* kmpx_last_error(void);
/* Number of kernels launched by this library in the calling process since load (bench.py's gpu_launches). */
long long kmpx_launch_count(void);
/* Profiling aid: cycle counters per phase of k_potrf_small / k_trtri_small accumulated by CTA 0 (16 host int64). */
int kmpx_debug_phases(long long* out_host, int reset);
/* Profiling experiments (tools/ only; 0 = normal operation).  Results are WRONG with bits 0-3 set.
 *   bit 0 (1)  skip the DMMAs of the contractions (factor kernels, fused predict)
 *   bit 1 (2)  skip the cp.async loads of k_potrf_small / k_trtri_small
 *   bit 2 (4)  skip the kappa panel fill of the fused predict;  bit 3 (8) skip its per-chunk reductions
 *   bit 4 (16) force the cp.async factor kernels instead of the TMA ones;  bit 5 (32) no L2 prefetch in the TMA factor kernels
 *   bit 6 (64) per-category timing printout of the blocked large-n Cholesky (host side)
 *   bit 7 (128) route NT products through the cp.async GEMM instead of the TMA-fed persistent one (results stay correct) */
int kmpx_debug_set(int flags);
/* k-steps executed / scheduled by the table-driven predict kernel since the last reset (out_host[2]), counted while
 * kmpx_debug_set(32768) is on: how much of the dense contraction the 2^-60 kappa-block skipping left. */
int kmpx_debug_predict_stats(long long* out_host, int reset);
/* Host only: the (O, N) row of the predict kernel's stage table for 4 or 8 MMA warps (layout in predict_tab.cu); returns the ints per
 * row (776), 0 when the system is too large for the table (the kernel is then not used), < 0 on a bad argument. */
int kmpx_debug_predict_schedule(int O, int N, int warps, int* row_out, int row_cap);
/* y[i] = exp(x[i]) through the branch-free exponentials of the hot kernels (mode 0: exp_nonpos, 1: table-driven exp_tab);
 * accuracy probe for the tests, x and y are device arrays. */
int kmpx_debug_exp(const double* x, double* y, long long n, int mode, void* stream);

/* ---- K1: kernel-matrix builder.  Replaces the N^2 Python loop of KMP.fit, KMP.py:146-156, with
 *      KMP.__kernel_matrix (KMP.py:93-128) evaluated in-kernel (plain RBF block or the finite-difference
 *      time-driven 2x2 block form).  A = K (x) I + lambda*blockdiag(Sigma), written once, 16-byte stores.
 *      uplo = KMPX_UPLO_LOWER writes only row >= col (what potrf consumes). */
int kmpx_kernel_build(const double* s, const double* sigma, const int* n_pts, double* A,
                      int batch, int n_max, int I, int O, int lda, long long strideA,
                      double sigma_f, double lambda, int time_driven, int layout, int uplo, void* stream);

/* ---- K2: batched fp64 Cholesky A = L L^T, lower, in place (replaces numpy.linalg.inv at KMP.py:157
 *      for symmetric systems).  n_sys[b] = system size (NULL => n_max).  info[b] = 0 or the 1-based
 *      index of the first non-positive pivot (LAPACK convention).  n_max <= KMPX_SMALL_MAX: one CTA per
 *      problem; larger: blocked right-looking on DMMA GEMMs (needs ws). */
long long kmpx_potrf_workspace_bytes(int n_max, int batch);
int kmpx_potrf(double* A, int lda, long long strideA, const int* n_sys, int n_max, int batch,
               int* info, void* ws, long long ws_bytes, void* stream);

/* ---- K3: in-place inverse of the lower-triangular factor, L <- L^{-1} (same dispatch as potrf).
 *      Only the lower triangle of the result is defined.  What the call may write besides it: the strict upper
 *      triangle (zeros; all of it for n > KMPX_SMALL_MAX, the 32 x 32 diagonal blocks otherwise) and, in a ragged
 *      batch, the columns n_sys[b] .. n_max-1 of the rows of problem b (zeros; scratch inside the problem's own
 *      n_max x lda slot).  Rows >= n_sys[b] are never touched. */
long long kmpx_trtri_workspace_bytes(int n_max, int batch);
int kmpx_trtri(double* L, int lda, long long strideA, const int* n_sys, int n_max, int batch,
               void* ws, long long ws_bytes, void* stream);

/* ---- K2 / K3, warp-per-problem variants for LARGE batches of small systems (64 <= n_max <= KMPX_SMALL_MAX; cfg 3).
 *      Same mathematics and result layout as kmpx_potrf / kmpx_trtri (KMP.py:157), but a warp owns a problem from the
 *      first block column to the last (factor_wpp.cuh): twelve independent problems per SM hide each other's latency,
 *      which pays once the batch holds several problems per warp slot (batch >~ 4 * 148).  kmpx_trtri_wpp writes the
 *      lower triangle only (plus explicit zeros above the diagonal inside the 32 x 32 diagonal blocks); nothing else of
 *      a problem's n_max x lda slot is touched.
 *      struct_o: 0, or 2 for component-major KMP systems with two outputs and the plain kernel (kmpx_kernel_build with
 *      O = 2, KMPX_LAYOUT_COMPONENT: A = [[A00, D], [D, A11]], D diagonal, component stride n_sys[b] / 2) - blocks of
 *      the factor that are exactly zero are then skipped; the results are bit-identical to struct_o = 0. */
int kmpx_factor_wpp_applicable(const double* A, int lda, long long strideA, int n_max, int batch);
int kmpx_potrf_wpp(double* A, int lda, long long strideA, const int* n_sys, int n_max, int batch, int* info,
                   int struct_o, void* stream);
int kmpx_trtri_wpp(double* L, int lda, long long strideA, const int* n_sys, int n_max, int batch, int struct_o, void* stream);
/* kmpx_trtri_wpp that also leaves w = Linv y' (the w of kmpx_solve_alpha, system order, [batch][ldw]) - formed from the
 * blocks of the inverse while they are on chip, which saves the extra pass over Linv.  Component-major systems only:
 * n_max = O * (n_max_pts rounded up to even); y point-major [batch][n_max_pts][O], n_pts [batch] (NULL with n_sys NULL). */
int kmpx_trtri_wpp_solve(double* L, int lda, long long strideA, const int* n_sys, int n_max, int batch, int struct_o,
                         const double* y, const int* n_pts, int n_max_pts, int O, double* w, int ldw, void* stream);

/* ---- K1 + K2 + K3 fused for the small plain-kernel systems (O <= 2, N <= ~204: cfg 1 / cfg 3).  Replaces KMP.fit's
 *      N^2 loop and numpy.linalg.inv, KMP.py:146-157, in ONE kernel: the matrix is generated from (s, sigma) inside shared
 *      memory and factorised / inverted there exploiting that the cross-component blocks lambda * Sigma_ab are diagonal
 *      (factor_struct.cuh).  F receives L^-1 in the component-major order exactly as kmpx_potrf + kmpx_trtri leave it
 *      (lower triangle; the strict upper triangle is scratch).  info[b] = 0 or the 1-based index of the first non-positive
 *      pivot.  sigma blocks must be symmetric (the caller routes non-symmetric ones to the general path). */
int kmpx_factor_struct_applicable(int n_max_pts, int I, int O);
int kmpx_factor_struct(const double* s, const double* sigma, const int* n_pts, double* F, int lda, long long strideF,
                       int n_max_pts, int I, int O, int batch, double sigma_f, double lambda, int* info, void* stream);

/* ---- K3a: w = Linv * P y and alpha = P^T Linv^T w  (= (K+lambda Sigma)^{-1} mu, the reference's
 *      `_estimator @ Y`, KMP.py:185-187).  y, alpha are point-major [batch][n_max][O]; P is the permutation
 *      to `layout`; w [batch][ldw] is kept in system order for the predict kernel. */
int kmpx_solve_alpha(const double* Linv, int lda, long long strideA, const int* n_pts, int n_max, int O,
                     int batch, int layout, const double* y, double* w, int ldw, double* alpha, void* stream);

/* ---- general (non-symmetric) path: in-place Gauss-Jordan inverse with partial pivoting, one CTA per
 *      problem.  Needed because demos.py:251,285 hands set_waypoint an un-listed covariance, which makes the
 *      stored block non-symmetric; the reference's LU inverse digests it.  info[b] > 0 =>
 *      singular (the reference raises LinAlgError).  ws: batch*n_max int32 (pivot rows). */
long long kmpx_inverse_workspace_bytes(int n_max, int batch);
int kmpx_inverse(double* A, int lda, long long strideA, const int* n_sys, int n_max, int batch,
                 int* info, void* ws, long long ws_bytes, void* stream);
/* alpha' = Ainv * P y for the general path (same argument meaning as kmpx_solve_alpha); here w receives
 * alpha' in system order (what kmpx_kmp_predict with factor_kind 1 expects) and alpha the point-major copy. */
int kmpx_apply_inverse(const double* Ainv, int lda, long long strideA, const int* n_pts, int n_max, int O,
                       int batch, int layout, const double* y, double* w, int ldw, double* alpha, void* stream);

/* Size of the linear system for N stored points: N*O (point-major) or O*roundup(N,2) (component-major:
 * each component block is padded to an even length with identity rows so that it starts 16-byte aligned).
 * kmpx_sys_sizes fills a device array for a ragged batch. */
int kmpx_sys_size(int N, int O, int layout);
int kmpx_sys_sizes(const int* n_pts, int O, int layout, int batch, int* n_sys, void* stream);

/* ---- K4: predict.  Replaces the per-query loop of KMP.predict, KMP.py:178-190.
 *      xi    [batch][O][M]     = k* (K+lambda Sigma)^{-1} Y
 *      sigma [batch][O][O][M]  = alpha_kmp * (kernel(s*,s*) - k* (K+lambda Sigma)^{-1} k*^T)
 *      (the reference's output layout, query index last).
 *      factor_kind 0: F = Linv (component-major, from potrf+trtri), w from kmpx_solve_alpha:
 *                     V = Linv k*^T on DMMA, xi = V^T w, sigma = alpha (k** - V^T V), fused, V never stored.
 *      factor_kind 1: F = Ainv (component-major, general path): U = Ainv k*^T, xi = U^T y', sigma uses k* U.
 *      queries sq [batch or 1][M][I]; stride_sq = 0 shares one query grid between all problems.
 *      want_cov = 0 skips the covariance (mean only, from alpha: cheap path). */
int kmpx_kmp_predict(const double* F, int lda, long long strideA, int factor_kind,
                     const double* w, int ldw, const double* s, const int* n_pts, int n_max, int I, int O,
                     int batch, const double* sq, long long stride_sq, int M,
                     double sigma_f, double alpha_kmp, int time_driven,
                     double* xi, double* sigma_out, void* stream);
/* RBF locality in kmpx_kmp_predict (plain kernel, O <= 2, the table-driven kernel): a box of 8 stored points whose kernel
 * coefficients k(s*, s_j) are below `threshold` for ALL queries of a tile (32 or 64, below) is treated as zero - its share of the
 * contraction is skipped.  Default 2^-68: an entry of V = Linv k*^T then loses at most 2^-60 max|Linv row|, 1/128 of one
 * rounding error of a retained term; with the reference's sigma_f = 200 about half of the contraction of a
 * cfg-3 problem goes.  threshold = 0 keeps every term.  Process-wide setting, read at launch time. */
int kmpx_predict_skip_threshold(double threshold);
/* Queries per CTA of the same kernel: 64 (one CTA of eight MMA warps per SM) or 32 (two CTAs of four MMA warps per SM, whose
 * set-up and epilogue overlap the other CTA's contraction).  0 (default) chooses 32 when boxes are skipped (threshold > 0)
 * and 64 when every term is kept - measured at cfg 3: 116 vs 120 ms per 26 624 problems with skipping, 185 vs 179 ms
 * without.  The two shapes sum in different orders (results differ by ~1e-16 norm-wise); the choice never depends on the
 * batch, so a problem's bits do not depend on its neighbours.  Process-wide setting, read at launch time. */
int kmpx_predict_tile_queries(int queries);

/* kmpx_kmp_predict for LARGE batches (many (problem, query tile) pairs per SM): same arguments and mathematics; where
 * the fast path applies (plain kernel, O <= 2, factor_kind 0) persistent CTAs walk tiles of 40 queries with the kappa panel
 * of the next tile evaluated, and the results of the previous one written, while the tensor pipe contracts the current
 * one (predict_pers.cu).  Results agree with kmpx_kmp_predict to rounding (the summation order over the rows differs);
 * a problem's bits do not depend on the batch it sits in.  Falls back to kmpx_kmp_predict's kernels elsewhere.
 * Measured 6 % slower than kmpx_kmp_predict on the cfg-3 shape (the smaller tile costs more in the streaming loop than
 * the hidden prologue saves): an experiment kept for reference, not what KMPBatch calls by default. */
int kmpx_kmp_predict_persistent(const double* F, int lda, long long strideA, int factor_kind,
                                const double* w, int ldw, const double* s, const int* n_pts, int n_max, int I, int O,
                                int batch, const double* sq, long long stride_sq, int M,
                                double sigma_f, double alpha_kmp, int time_driven,
                                double* xi, double* sigma_out, void* stream);
/* The same prediction for ONE large system (n > KMPX_SMALL_MAX, plain kernel, O <= 2, symmetric factor) as DMMA GEMMs:
 * kappa^T (Mc x N) once per chunk of queries, V_a = Linv[a Ns:, a Ns: a Ns + N] kappa with kmpx_dgemm's triangular mode,
 * column reductions for xi / sigma.  F as left by kmpx_trtri for n > KMPX_SMALL_MAX (strict upper triangle zero).
 * The workspace decides the chunk size Mc (a multiple of 128; 8192 queries of cfg 4 need 2.2 GB). */
long long kmpx_kmp_predict_gemm_workspace_bytes(int N, int O, int Mc);
int kmpx_kmp_predict_gemm(const double* F, int lda, const double* w, const double* s, int N, int I, int O,
                          const double* sq, int M, double sigma_f, double alpha_kmp, double* ws, long long ws_bytes,
                          double* xi, double* sigma_out, void* stream);
/* mean-only predict from alpha (point-major [batch][n_max][O]): xi[a][m] = sum_i k(s*_m,s_i) alpha_i. */
int kmpx_kmp_predict_mean(const double* alpha, const double* s, const int* n_pts, int n_max, int I, int O,
                          int batch, const double* sq, long long stride_sq, int M,
                          double sigma_f, int time_driven, double* xi, void* stream);

/* ---- a4: KMP.set_waypoint's database update (KMP.py:72-89) for a batch of problems that share base
 *      trajectories.  For via-point j = 0..P-1 in order: nearest stored input by L2, strict `< tol` =>
 *      overwrite in place, else append.  base_of[b] selects the base trajectory of problem b. */
int kmpx_waypoint_insert(const double* s_base, const double* y_base, const double* sigma_base,
                         const int* n_base, int n_base_max, const int* base_of,
                         const double* via_s, const double* via_y, const double* via_sigma, int P, double tol,
                         double* s, double* y, double* sigma, int* n_pts, int n_max, int I, int O, int batch,
                         void* stream);

/* ---- incremental adaptation for batches of problems that share base trajectories.  Replaces the
 *      full refit of KMP.set_waypoint (KMP.py:91) + KMP.predict (KMP.py:160-194) by an exact low-rank correction of
 *      the base solution (deleted + appended points, block-inverse identities; see csrc/incremental.cu).
 *      Per base, once:  G = A_b^-1 [n_bases][n0][ldg] in component-major order (n0 = kmpx_sys_size(N0, O, COMPONENT)),
 *      the base predictions xi_base [n_bases][O][M], sigma_base [n_bases][O][O][M] (kmpx_kmp_predict), and
 *      kmpx_inc_base -> alpha_sys [n_bases][n0] = G y, kappa [n_bases][N0][M] = k(s*_m, s_j).
 *      Per batch: kmpx_inc_adapt.  via_* as in kmpx_waypoint_insert (same nearest-point / tolerance rule);
 *      limits O <= 2, P <= 2, plain kernel, all bases with N0 points.  ws: kmpx_inc_workspace_doubles() doubles. */
long long kmpx_inc_workspace_doubles(int N0, int O, int batch);
int kmpx_inc_base(const double* G, int ldg, long long strideG, const double* y_base, const double* s_base, int N0, int I,
                  int O, int n_bases, const double* sq, int M, double sigma_f, double* alpha_sys, double* kappa,
                  void* stream);
int kmpx_inc_adapt(const double* s_base, const double* G, int ldg, long long strideG, const double* alpha_sys,
                   const double* kappa, const double* xi_base, const double* sigma_base, int N0, int I, int O,
                   const int* base_of, const double* via_s, const double* via_y, const double* via_sigma, int P,
                   double tol, double sigma_f, double lam, double alpha_kmp, int batch, const double* sq, int M,
                   double* ws, long long ws_doubles, double* xi, double* sigma_out, void* stream);

/* The same with the problems sorted by base: group g = group_count[g] consecutive problems of base group_base[g]
 * (HOST arrays).  The appended columns of a whole group come from one kmpx_dgemm-class GEMM per output component. */
long long kmpx_inc_grouped_workspace_doubles(int N0, int O, int batch);
int kmpx_inc_adapt_grouped(const double* s_base, const double* G, int ldg, long long strideG, const double* alpha_sys,
                           const double* kappa, const double* xi_base, const double* sigma_base, int N0, int I, int O,
                           const int* base_of, const int* group_base, const int* group_count, int n_groups,
                           const double* via_s, const double* via_y, const double* via_sigma, int P,
                           double tol, double sigma_f, double lam, double alpha_kmp, int batch, const double* sq, int M,
                           double* ws, long long ws_doubles, double* xi, double* sigma_out, void* stream);

/* ---- K8: GMR conditioning.  Replaces GaussianMixtureModel.predict, GaussianMixtureModel.py:83-105.
 *      priors [K], means [d][K], covs [d][d][K] (the reference's feature-major attributes), x [I][M];
 *      out mu [O][M], sigma [O][O][M].  Keeps the REALMIN normaliser (:90) and reg*I - mu mu^T (:104-105). */
int kmpx_gmr(const double* priors, const double* means, const double* covs, int K, int I, int O,
             const double* x, int M, double reg, double* mu, double* sigma_out, void* stream);

/* ---- K5/K6: one fused EM pass.  Replaces sklearn's _e_step + _m_step
 *      (sklearn/mixture/_base.py:265-278, _gaussian_mixture.py:282-320,490-553) behind
 *      GaussianMixtureModel.fit, GaussianMixtureModel.py:53.
 *      X [n][d] row-major samples; `shift` [d] is subtracted from every sample before the moments are
 *      accumulated (EM is shift-invariant).  Parameters: log_weights [K], means [K][d],
 *      prec_chol [K][d][d] (upper-triangular P_k = chol(Sigma_k)^{-T}), log_det [K].
 *      stats [K*F + 1], F = 1 + d + d(d+1)/2: per component {sum r, sum r x', sum r x' x'^T (upper, row order)}
 *      with x' = x - shift, then sum_n log p(x_n).  This is the buffer the NCCL all-reduce sums when the
 *      samples are partitioned across GPUs.  labels (int32 [n], argmax responsibility) and lse [n] optional. */
long long kmpx_gmm_workspace_bytes(int n, int d, int K);
int kmpx_gmm_em_pass(const double* X, long long n, int d, int K, const double* shift,
                     const double* log_weights, const double* means, const double* prec_chol,
                     const double* log_det, double* stats, int* labels, double* lse,
                     void* ws, long long ws_bytes, void* stream);
/* Same sufficient statistics from hard labels (one-hot responsibilities): the initial M-step after the
 * k-means initialisation (sklearn/mixture/_base.py:119-128) and the Lloyd centre update. */
int kmpx_gmm_label_stats(const double* X, long long n, int d, int K, const double* shift,
                         const int* labels, double* stats, void* ws, long long ws_bytes, void* stream);
/* The same with only the count and the first moments accumulated (Lloyd iterations need nothing else; the second-moment
 * entries of stats are written as 0). */
int kmpx_gmm_label_sums(const double* X, long long n, int d, int K, const double* shift,
                         const int* labels, double* stats, void* ws, long long ws_bytes, void* stream);
/* M-step finalisation from (all-reduced) statistics: nk = sum r + 10 eps, means, covariances + reg*I,
 * weights = nk / n_total (first = 1) or nk / sum nk, precision Cholesky, log-dets
 * (sklearn/mixture/_gaussian_mixture.py:312-320,358-370,883-901).  info[k] != 0 => covariance not SPD. */
int kmpx_gmm_finalize(const double* stats, double n_total, int first, int d, int K, double reg,
                      const double* shift, double* weights, double* log_weights, double* means,
                      double* covs, double* prec_chol, double* log_det, int* info, void* stream);

/* kmpx_gmm_finalize (first = 0) with sklearn's stopping rule on the device (mixture/_base.py:265-278):
 * em_state = {done, n_iter, lower, converged} (doubles, device); lower = stats[K*F] / n_total is compared with the stored
 * one, |delta| < tol or n_iter == max_iter sets done; once done the call leaves the parameters untouched, so passes can
 * be enqueued in batches with one read of em_state per batch. */
int kmpx_gmm_finalize_check(const double* stats, double n_total, int d, int K, double reg,
                            const double* shift, double* weights, double* log_weights, double* means,
                            double* covs, double* prec_chol, double* log_det, int* info,
                            double tol, int max_iter, double* em_state, void* stream);

/* ---- K7 (sklearn mode): the k-means initialisation of the GMM
 *      (sklearn/cluster/_kmeans.py:180-278,630-758, _k_means_lloyd.pyx:23-230).  Only the RandomState stream stays on
 *      the host; sampling, distances, minima, potentials, the greedy choice, assignment, centre sums, centre update and
 *      the stopping rules are device kernels.  Rows may be partitioned over ranks (n_before rows in front of this
 *      rank's n, n_total in all); the exchanges between the kernels are the caller's (torch.distributed / NCCL). */
int kmpx_rows_center_sqnorm(double* X, long long n, int d, const double* mean, double* x_sq, void* stream);
/* dist[t][i] = min(closest[i], max(0, |c_t|^2 - 2 c_t.x_i + |x_i|^2)), pot[t] = sum_i dist[t][i];
 * closest may be NULL (first centre).  cand [T] are row indices into X. */
int kmpx_kpp_candidates(const double* X, const double* x_sq, long long n, int d, const int* cand, int T,
                        const double* closest, double* dist, double* pot, void* ws, long long ws_bytes,
                        void* stream);
/* k-means++ draw, replaces  searchsorted(cumsum(closest), uniform(T) * pot)  (_kmeans.py:227-231): u [T] uniform draws,
 * pot the current potential (device scalar), before {hi, lo} the double-double potential of the ranks in front (NULL: 0).
 * cand[t] = LOCAL row index of draw t or -1 when another rank owns it; flag[t] = 1 when the draw lies within the
 * rounding uncertainty of numpy's sequential cumulative sum / BLAS potential (the caller repeats it with numpy). */
long long kmpx_kpp_sample_workspace_bytes(long long n);
int kmpx_kpp_sample(const double* closest, long long n, const double* u, int T, const double* pot,
                    const double* before, long long n_before, long long n_total, int last_rank,
                    long long* cand, int* flag, void* ws, long long ws_bytes, void* stream);
/* rows [T][d+2] = [X[cand[t]], 1, n_before + cand[t]] for owned draws, zeros otherwise (summed over the ranks by the caller) */
int kmpx_kpp_gather_rows(const double* X, int d, const long long* cand, int T, long long n_before, double* rows,
                         void* stream);
/* kmpx_kpp_candidates for candidate ROWS [T][d+2]; pot [T][2] = this rank's potentials in double-double */
int kmpx_kpp_candidates_rows(const double* X, const double* x_sq, long long n, int d, const double* rows, int T,
                             const double* closest, double* dist, double* pot, void* ws, long long ws_bytes,
                             void* stream);
/* greedy choice (_kmeans.py:245-251): pots [R][T][2] per-rank potentials; best = first minimum of their rank-ordered
 * sums (force_best >= 0 overrides); writes pot_out, before_out {hi, lo} for the next draw, centers[c], center_ids[c],
 * best_out, flag (two different candidates within rounding slack), and closest[i] = dist[best][i] when dist != NULL. */
int kmpx_kpp_choose(const double* pots, int R, int T, int rank, const double* rows, int d, int c, long long n_total,
                    int force_best, const double* dist, long long n, double* pot_out, double* before_out,
                    double* centers, long long* center_ids, int* best_out, int* flag, double* closest, void* stream);
/* Lloyd centre update + stopping rules (_kmeans.py:700-737): stats = packed hard-label statistics (kmpx_gmm_label_stats,
 * summed over the ranks), n_changed_total device scalar (double), state = {done, n_iter, strict, empty}; a no-op once
 * done or when a cluster is empty (state[3], the host relocates). */
int kmpx_lloyd_update(const double* stats, int K, int d, double* centers, const double* tol_abs,
                      const double* n_changed_total, int* n_changed_local, int max_iter, int* state, void* stream);
/* labels[i] = argmin_j(|c_j|^2 - 2 x_i.c_j) (first minimum); n_changed += (labels != labels_prev). */
int kmpx_kmeans_assign(const double* X, long long n, int d, int K, const double* centers,
                       const int* labels_prev, int* labels, int* n_changed, void* stream);

/* ---- K7 (reference mode): kmp.cluster.KMeans, Cluster.py:88-102, bit-exact: feature-major X [d][n],
 *      distances = sqrt of a left-to-right sum of separately rounded squares (no FMA), first-index argmin
 *      with numpy's NaN rule, labels int64; update = sequential sum in sample order then one division,
 *      NaN for an empty cluster. */
int kmpx_refkmeans_assign(const double* X, long long n, int d, int K, const double* centroids,
                          long long* labels, void* stream);
int kmpx_refkmeans_update(const double* X, long long n, int d, int K, const long long* labels,
                          double* centroids, void* stream);

/* ---- K9: quaternion maps, kmp/types/quaternion.py (scalar-first [v,u0,u1,u2], every result re-normalised
 *      as the reference constructor does, :18-21).  M elements; for kmpx_quat_mul the second operand is
 *      broadcast when q_bcast != 0. */
int kmpx_quat_normalize(const double* q, double* out, long long M, void* stream);
int kmpx_quat_exp(const double* w, double* q, long long M, void* stream);            /* :60-80  */
int kmpx_quat_log(const double* q, double* w, long long M, void* stream);            /* :82-96  */
int kmpx_quat_mul(const double* p, const double* q, int q_bcast, int conj_q, double* out, long long M,
                  void* stream);                                                     /* :182-202, :219-227 */
int kmpx_quat_from_rotvec(const double* r, double* q, long long M, void* stream);    /* :23-42  */

/* ---- dense fp64 GEMM on DMMA (the building block of the blocked factorisations; exported for tests and
 *      for the lazily materialised `_estimator`).  C = alpha*op(A)*op(B) + beta*C, row-major, strided batch.
 *      tri: 0 none; 1 compute only tiles with row >= col of C (SYRK-like); 2 op(A) is lower-triangular
 *      (skip k > row); 3 op(B) is lower-triangular (skip k < col); 4 (NT form only) B [n][k] is lower-triangular with
 *      explicit zeros above the diagonal (skip k > col).  The NT form runs on a TMA-fed persistent kernel when rows and
 *      batches are 16-byte aligned (dgemm_tma.cu), everything else on the cp.async kernel (dgemm.cu).  */
int kmpx_dgemm(int transA, int transB, int m, int n, int k, double alpha, const double* A, int lda,
               long long strideA, const double* B, int ldb, long long strideB, double beta, double* C,
               int ldc, long long strideC, int batch, int tri, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KMPX_H */
