// K2/K3 for large systems (n > KMPX_SMALL_MAX): blocked algorithms whose flops all run in k_dgemm (DMMA),
// plus the public kmpx_potrf / kmpx_trtri / kmpx_solve_alpha / kmpx_apply_inverse dispatchers.
//
//   potrf  right-looking, NB = 256:  diag block -> k_potrf_small (one CTA);  its inverse -> workspace;
//          panel  L21 = A21 * Dinv^T  (GEMM);  trailing  A22 -= L21 L21^T  (GEMM, lower tiles only).
//   trtri  right to left, NB = 256:  T = L21 * Dinv (GEMM);  L21 <- -Linv22 * T  (GEMM, triangular k-range).
// cfg 4 (n = 16384): potrf = n^3/3 = 1.47e12 flop, 64 block steps.
#include "dmma_tile.cuh"
#include "kmpx_internal.cuh"
#include <vector>

namespace kmpx {

constexpr int LNB = 256;


// dst (clean, ld_dst) <- lower triangle of src block, zeros above the diagonal
__global__ void k_copy_lower(const double* src, int ld_src, double* dst, int ld_dst, int nb) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c < nb && r < nb) dst[(size_t)r * ld_dst + c] = (c <= r) ? src[(size_t)r * ld_src + c] : 0.0;
}
__global__ void k_copy_block(const double* src, int ld_src, double* dst, int ld_dst, int rows, int cols) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (c < cols && r < rows) dst[(size_t)r * ld_dst + c] = src[(size_t)r * ld_src + c];
}
__global__ void k_zero_upper(double* A, int lda, int n) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (r < n && c < n && c > r) A[(size_t)r * lda + c] = 0.0;
}
__global__ void k_set_int(int* p, int v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

static int copy_block(const double* src, int ld_src, double* dst, int ld_dst, int rows, int cols, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return 0;
    dim3 blk(64, 4), grid(ceil_div(cols, 64), ceil_div(rows, 4));
    k_copy_block<<<grid, blk, 0, st>>>(src, ld_src, dst, ld_dst, rows, cols);
    return check_launch("k_copy_block");
}

// workspace layout (doubles): Dinv[2] [LNB][LNB] | panel[2] [n][LNB]
static long long large_ws_doubles(int n) { return 2LL * LNB * LNB + 2LL * round_up(n, 2) * LNB; }

struct SideStream {
    cudaStream_t s = nullptr; cudaEvent_t ev_main = nullptr, ev_side = nullptr;
    int ensure() {
        if (s) return 0;
        // highest priority: its single-CTA kernels must be dispatched as soon as any SM drains, ahead of the
        // thousands of pending CTAs of the trailing update on the caller's stream
        int least = 0, greatest = 0;
        KMPX_TRY(check_cuda(cudaDeviceGetStreamPriorityRange(&least, &greatest), "cudaDeviceGetStreamPriorityRange"));
        KMPX_TRY(check_cuda(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, greatest), "cudaStreamCreateWithPriority"));
        KMPX_TRY(check_cuda(cudaEventCreateWithFlags(&ev_main, cudaEventDisableTiming), "cudaEventCreate"));
        KMPX_TRY(check_cuda(cudaEventCreateWithFlags(&ev_side, cudaEventDisableTiming), "cudaEventCreate"));
        return 0;
    }
};
static SideStream g_side_dev[kMaxDevices];       // streams and events belong to a device
#define g_side (g_side_dev[current_device()])

static int diag_factor(double* D, int lda, int nb, double* Dinv, int* info, int k, cudaStream_t st) {
    KMPX_TRY(launch_potrf_small(D, lda, 0, nullptr, nb, 1, info, k, 1, st));
    k_copy_lower<<<dim3(ceil_div(nb, 128), nb), 128, 0, st>>>(D, lda, Dinv, LNB, nb);
    KMPX_TRY(check_launch("k_copy_lower"));
    return launch_trtri_small(Dinv, LNB, 0, nullptr, nb, 1, st);
}

// Right-looking blocked Cholesky with one step of look-ahead, two streams:
//   main stream   look-ahead(s): block column s+1 of the trailing matrix -= P_s P_s^T      (m x 256, K = 256)
//                 trailing(s):   the rest of the trailing matrix -= P_s P_s^T               (lower tiles, persistent GEMM
//                                that leaves a few SMs to the side stream)
//   side stream   after look-ahead(s): factor + invert the diagonal block s+1 (single-CTA kernels), then the panel
//                 P_{s+1} = A21 * Dinv^T (into the other of two dense panel buffers) and the write-back of L21 into A.
//                 The panel GEMM's CTAs are scheduled as the trailing update's persistent CTAs run out of tiles, i.e. it
//                 fills the tail of trailing(s) instead of following it.
// look-ahead(s+1) on the main stream waits for the side stream.
static int potrf_large(double* A, int lda, int n, int* info, double* ws, cudaStream_t st) {
    double* Dinv[2] = {ws, ws + (size_t)LNB * LNB};
    double* panel[2] = {ws + 2 * (size_t)LNB * LNB, ws + 2 * (size_t)LNB * LNB + (size_t)round_up(n, 2) * LNB};
    KMPX_TRY(g_side.ensure());
    if (info) { k_set_int<<<1, 32, 0, st>>>(info, 0, 1); KMPX_TRY(check_launch("k_set_int")); }
    // profiling aid (kmpx_debug_set bit 6): per-category device time on the caller's stream
    const bool prof = (g_host_dbg & 64) != 0;
    std::vector<cudaEvent_t> evs;
    auto mark = [&]() { if (prof) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); evs.push_back(e); } };
    // diagonal block k, its inverse, the panel below it and the write-back of L21: everything step `step` needs before its
    // look-ahead, on stream `s`
    auto prepare = [&](int step, cudaStream_t s) -> int {
        const int k = step * LNB;
        const int nb = n - k < LNB ? n - k : LNB;
        KMPX_TRY(diag_factor(A + (size_t)k * lda + k, lda, nb, Dinv[step & 1], info, k, s));
        const int m = n - k - nb;
        if (m <= 0) return 0;
        double* A21 = A + (size_t)(k + nb) * lda + k;
        GemmArgs g1{m, nb, nb, 1.0, A21, lda, 0, Dinv[step & 1], LNB, 0, 0.0, panel[step & 1], LNB, 0, 4, 0};   // Dinv: zeros above the diagonal
        KMPX_TRY(launch_dgemm(0, 1, g1, 1, s));                                // P = A21 * Dinv^T (dense copy)
        return copy_block(panel[step & 1], LNB, A21, lda, m, nb, s);          // L21 in place (nothing reads it again)
    };
    KMPX_TRY(prepare(0, st));
    int step = 0;
    for (int k = 0; k < n; k += LNB, ++step) {
        const int nb = n - k < LNB ? n - k : LNB;
        const int m = n - k - nb;
        if (m <= 0) break;
        double* P = panel[step & 1];
        double* A22 = A + (size_t)(k + nb) * lda + k + nb;
        const int nb2 = m < LNB ? m : LNB;
        mark();
        GemmArgs g2{m, nb2, nb, -1.0, P, LNB, 0, P, LNB, 0, 1.0, A22, lda, 0, 1, 4};
        KMPX_TRY(launch_dgemm(0, 1, g2, 1, st));                               // look-ahead: next block column
        mark();
        KMPX_TRY(check_cuda(cudaEventRecord(g_side.ev_main, st), "cudaEventRecord"));
        KMPX_TRY(check_cuda(cudaStreamWaitEvent(g_side.s, g_side.ev_main, 0), "cudaStreamWaitEvent"));
        KMPX_TRY(prepare(step + 1, g_side.s));
        KMPX_TRY(check_cuda(cudaEventRecord(g_side.ev_side, g_side.s), "cudaEventRecord"));
        if (m > nb2) {
            const int m2 = m - nb2;
            double* P2 = P + (size_t)nb2 * LNB;
            GemmArgs g3{m2, m2, nb, -1.0, P2, LNB, 0, P2, LNB, 0, 1.0, A22 + (size_t)nb2 * lda + nb2, lda, 0, 1, 4};
            KMPX_TRY(launch_dgemm(0, 1, g3, 1, st));                           // rest of the trailing update (lower tiles)
        }
        mark();
        KMPX_TRY(check_cuda(cudaStreamWaitEvent(st, g_side.ev_side, 0), "cudaStreamWaitEvent"));
        mark();
    }
    if (prof && !evs.empty()) {
        cudaStreamSynchronize(st);
        double tot[4] = {0, 0, 0, 0};       // look-ahead gemm, trailing gemm, wait for the side stream, (between steps)
        for (size_t i = 0; i + 1 < evs.size(); ++i) { float ms = 0; cudaEventElapsedTime(&ms, evs[i], evs[i + 1]); tot[i % 4] += ms; }
        fprintf(stderr, "potrf_large n=%d: look-ahead %.2f ms, trailing %.2f, side wait (diag + panel) %.2f, step gap %.2f\n", n,
                tot[0], tot[1], tot[2], tot[3]);
        for (cudaEvent_t e : evs) cudaEventDestroy(e);
    }
    return 0;
}

// In-place inverse of the lower factor by recursive doubling over 256-blocks:
//   [[A, 0], [B, C]]^{-1} = [[A^{-1}, 0], [-C^{-1} B A^{-1}, C^{-1}]]
// level 0 inverts all diagonal 256-blocks in ONE batched launch of k_trtri_small; level S (S = 1, 2, 4, ... blocks)
// merges neighbouring inverted spans with two GEMMs per pair, T = B * A^{-1} and B <- -C^{-1} * T, and all pairs of a
// level run as one strided batch.  Every flop beyond level 0 is a large DMMA GEMM with thousands of CTAs (the
// block-column sweep it replaces had only (n - k)/128 x 2 CTAs per step).  Workspace: (n/2)^2 doubles for T.
static long long trtri_ws_doubles(int n) {
    const long long half = round_up((long long)(n + 1) / 2, LNB);
    return half * half + 2LL * LNB * LNB;
}

static int trtri_large(double* L, int lda, int n, double* ws, cudaStream_t st) {
    double* T = ws;
    {
        dim3 blk(64, 4), grid(ceil_div(n, 64), ceil_div(n, 4));
        k_zero_upper<<<grid, blk, 0, st>>>(L, lda, n);
        KMPX_TRY(check_launch("k_zero_upper"));
    }
    const int nblk = (n + LNB - 1) / LNB;
    const long long dstride = (long long)LNB * lda + LNB;                // from one diagonal block to the next
    const int nfull = n / LNB;
    if (nfull > 0) KMPX_TRY(launch_trtri_small(L, lda, dstride, nullptr, LNB, nfull, st));
    if (nblk > nfull) KMPX_TRY(launch_trtri_small(L + (size_t)nfull * dstride, lda, 0, nullptr, n - nfull * LNB, 1, st));
    for (int S = 1; S < nblk; S *= 2) {
        const int span = S * LNB;                                        // rows/cols of an inverted span
        // groups [g, g + 2S): A = [g, g+S), C = [g+S, min(g+2S, nblk)); full groups have identical shapes
        int nfullgrp = 0;
        while ((long long)(nfullgrp + 1) * 2 * span <= n) ++nfullgrp;
        const long long gstride = 2LL * span * lda + 2LL * span;
        if (nfullgrp > 0) {
            double* B0 = L + (size_t)span * lda;                         // B of group 0: rows [span, 2 span), cols [0, span)
            double* A0 = L;
            double* C0 = L + (size_t)span * lda + span;
            GemmArgs g1{span, span, span, 1.0, B0, lda, gstride, A0, lda, gstride, 0.0, T, span, (long long)span * span, 3};
            KMPX_TRY(launch_dgemm(0, 0, g1, nfullgrp, st));              // T = B * Ainv   (Ainv lower: k >= col)
            GemmArgs g2{span, span, span, -1.0, C0, lda, gstride, T, span, (long long)span * span, 0.0, B0, lda, gstride, 2};
            KMPX_TRY(launch_dgemm(0, 0, g2, nfullgrp, st));              // B <- -Cinv * T (Cinv lower: k <= row)
        }
        const long long g0 = (long long)nfullgrp * 2 * span;             // first row of the partial group, if any
        const int mc = n - (int)g0 - span;                               // rows of its C part
        if (mc > 0) {
            double* A1 = L + (size_t)g0 * lda + g0;
            double* B1 = L + (size_t)(g0 + span) * lda + g0;
            double* C1 = L + (size_t)(g0 + span) * lda + g0 + span;
            GemmArgs g1{mc, span, span, 1.0, B1, lda, 0, A1, lda, 0, 0.0, T, span, 0, 3};
            KMPX_TRY(launch_dgemm(0, 0, g1, 1, st));
            GemmArgs g2{mc, span, mc, -1.0, C1, lda, 0, T, span, 0, 0.0, B1, lda, 0, 2};
            KMPX_TRY(launch_dgemm(0, 0, g2, 1, st));
        }
    }
    return 0;
}

// ---- large-n alpha solve: two bandwidth-bound passes over F with many CTAs -----------------------------
__device__ __forceinline__ double y_sys(const double* y, int r, int N, int O, int layout) {
    if (layout == KMPX_LAYOUT_POINT) return y[r];
    const int Ns = comp_stride(N);
    const int a = r / Ns, i = r - a * Ns;
    return i < N ? y[(size_t)i * O + a] : 0.0;
}
// w[r] = sum_{c <= r (or all c)} F[r][c] y'[c], one warp per row
__global__ void __launch_bounds__(256) k_trmv_rows(const double* F, int lda, int n, int N, int O, int layout, int general,
                                                   const double* y, double* w) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n) return;
    const double* row = F + (size_t)r * lda;
    const int cend = general ? n : r + 1;
    double acc = 0.0;
    for (int c = lane; c < cend; c += 32) acc += row[c] * y_sys(y, c, N, O, layout);
    acc = warp_sum(acc);
    if (lane == 0) w[r] = acc;
}
// alpha[i][a] = sum_{r >= c} F[r][c] w[r]; a CTA owns 32 columns, its 8 warps split the rows
__global__ void __launch_bounds__(256) k_trmv_cols(const double* F, int lda, int n, int N, int O, int layout,
                                                   const double* w, double* alpha) {
    __shared__ double part[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const int c0 = blockIdx.x * 32;
    double acc = 0.0;
    if (c < n) for (int r = c0 + warp; r < n; r += 8) if (r >= c) acc += F[(size_t)r * lda + c] * w[r];
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && c < n) {
        double s = 0.0;
        for (int t = 0; t < 8; ++t) s += part[t][lane];
        if (layout == KMPX_LAYOUT_POINT) alpha[c] = s;
        else { const int Ns = comp_stride(N); const int a = c / Ns, i = c - a * Ns; if (i < N) alpha[(size_t)i * O + a] = s; }
    }
}
__global__ void k_scatter_alpha(const double* ap, int n, int N, int O, int layout, double* alpha) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    if (layout == KMPX_LAYOUT_POINT) alpha[c] = ap[c];
    else { const int Ns = comp_stride(N); const int a = c / Ns, i = c - a * Ns; if (i < N) alpha[(size_t)i * O + a] = ap[c]; }
}
__global__ void k_sys_sizes(const int* n_pts, int O, int layout, int batch, int* n_sys) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch) n_sys[b] = layout == KMPX_LAYOUT_POINT ? n_pts[b] * O : comp_stride(n_pts[b]) * O;
}

}  // namespace kmpx

using namespace kmpx;

extern "C" long long kmpx_potrf_workspace_bytes(int n_max, int batch) {
    (void)batch;
    return n_max <= KMPX_SMALL_MAX ? 0 : large_ws_doubles(n_max) * (long long)sizeof(double);
}
extern "C" long long kmpx_trtri_workspace_bytes(int n_max, int batch) {
    (void)batch;
    return n_max <= KMPX_SMALL_MAX ? 0 : trtri_ws_doubles(n_max) * (long long)sizeof(double);
}

static int check_factor_args(double* A, int lda, long long strideA, int n_max, int batch) {
    if (!A) return fail_arg(1, "matrix is NULL");
    if (n_max <= 0) return fail_arg(5, "n_max must be > 0");
    if (lda < n_max) return fail_arg(2, "lda < n_max");
    if (lda & 1) return fail_arg(2, "lda must be even (16-byte cp.async rows)");
    if (strideA & 1) return fail_arg(3, "strideA must be even");
    if (reinterpret_cast<uintptr_t>(A) & 15) return fail_arg(1, "matrix must be 16-byte aligned");
    if (batch <= 0) return fail_arg(6, "batch must be > 0");
    return 0;
}

extern "C" int kmpx_potrf(double* A, int lda, long long strideA, const int* n_sys, int n_max, int batch,
                          int* info, void* ws, long long ws_bytes, void* stream) {
    KMPX_TRY(check_factor_args(A, lda, strideA, n_max, batch));
    cudaStream_t st = (cudaStream_t)stream;
    if (n_max <= KMPX_SMALL_MAX) return launch_potrf_small(A, lda, strideA, n_sys, n_max, batch, info, 0, 0, st);
    if (n_sys) return fail_arg(4, "ragged batches are only supported for n_max <= KMPX_SMALL_MAX");
    if (!ws || ws_bytes < kmpx_potrf_workspace_bytes(n_max, batch)) return fail_arg(8, "workspace too small");
    for (int b = 0; b < batch; ++b)
        KMPX_TRY(potrf_large(A + (size_t)b * strideA, lda, n_max, info ? info + b : nullptr, (double*)ws, st));
    return 0;
}

extern "C" int kmpx_trtri(double* L, int lda, long long strideA, const int* n_sys, int n_max, int batch,
                          void* ws, long long ws_bytes, void* stream) {
    KMPX_TRY(check_factor_args(L, lda, strideA, n_max, batch));
    cudaStream_t st = (cudaStream_t)stream;
    if (n_max <= KMPX_SMALL_MAX) return launch_trtri_small(L, lda, strideA, n_sys, n_max, batch, st);
    if (n_sys) return fail_arg(4, "ragged batches are only supported for n_max <= KMPX_SMALL_MAX");
    if (!ws || ws_bytes < kmpx_trtri_workspace_bytes(n_max, batch)) return fail_arg(7, "workspace too small");
    for (int b = 0; b < batch; ++b) KMPX_TRY(trtri_large(L + (size_t)b * strideA, lda, n_max, (double*)ws, st));
    return 0;
}

static int solve_common(const double* F, int lda, long long strideA, const int* n_pts, int n_max, int O, int batch,
                        int layout, const double* y, double* w, int ldw, double* alpha, int general, void* stream) {
    if (!F) return fail_arg(1, "factor is NULL");
    if (n_max <= 0) return fail_arg(5, "n_max must be > 0");
    if (O <= 0) return fail_arg(6, "O must be > 0");
    if (batch <= 0) return fail_arg(7, "batch must be > 0");
    if (layout != KMPX_LAYOUT_POINT && layout != KMPX_LAYOUT_COMPONENT) return fail_arg(8, "unknown layout");
    if (!y) return fail_arg(9, "y is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n_sys_max = layout == KMPX_LAYOUT_POINT ? (long long)n_max * O : (long long)comp_stride(n_max) * O;
    if (2 * n_sys_max * (long long)sizeof(double) <= 200 * 1024) {
        AlphaArgs p{F, lda, strideA, n_pts, n_max, O, batch, layout, y, w, ldw, alpha, general};
        return launch_solve_alpha(p, st);
    }
    if (n_pts) return fail_arg(4, "ragged batches are only supported for small systems");
    if (!w) return fail_arg(10, "w is required for large systems");
    const int n = (int)n_sys_max;
    for (int b = 0; b < batch; ++b) {
        const double* Fb = F + (size_t)b * strideA;
        const double* yb = y + (size_t)b * n_max * O;
        double* wb = w + (size_t)b * ldw;
        double* ab = alpha ? alpha + (size_t)b * n_max * O : nullptr;
        k_trmv_rows<<<ceil_div(n, 8), 256, 0, st>>>(Fb, lda, n, n_max, O, layout, general, yb, wb);
        KMPX_TRY(check_launch("k_trmv_rows"));
        if (!ab) continue;
        if (general) { k_scatter_alpha<<<ceil_div(n, 256), 256, 0, st>>>(wb, n, n_max, O, layout, ab); KMPX_TRY(check_launch("k_scatter_alpha")); }
        else { k_trmv_cols<<<ceil_div(n, 32), 256, 0, st>>>(Fb, lda, n, n_max, O, layout, wb, ab); KMPX_TRY(check_launch("k_trmv_cols")); }
    }
    return 0;
}

extern "C" int kmpx_solve_alpha(const double* Linv, int lda, long long strideA, const int* n_pts, int n_max, int O,
                                int batch, int layout, const double* y, double* w, int ldw, double* alpha, void* stream) {
    return solve_common(Linv, lda, strideA, n_pts, n_max, O, batch, layout, y, w, ldw, alpha, 0, stream);
}

extern "C" int kmpx_apply_inverse(const double* Ainv, int lda, long long strideA, const int* n_pts, int n_max, int O,
                                  int batch, int layout, const double* y, double* w, int ldw, double* alpha, void* stream) {
    return solve_common(Ainv, lda, strideA, n_pts, n_max, O, batch, layout, y, w, ldw, alpha, 1, stream);
}

extern "C" int kmpx_sys_size(int N, int O, int layout) {
    return layout == KMPX_LAYOUT_POINT ? N * O : comp_stride(N) * O;
}

extern "C" int kmpx_sys_sizes(const int* n_pts, int O, int layout, int batch, int* n_sys, void* stream) {
    if (!n_pts) return fail_arg(1, "n_pts is NULL");
    if (!n_sys) return fail_arg(5, "n_sys is NULL");
    if (batch <= 0) return fail_arg(4, "batch must be > 0");
    k_sys_sizes<<<ceil_div(batch, 256), 256, 0, (cudaStream_t)stream>>>(n_pts, O, layout, batch, n_sys);
    return check_launch("k_sys_sizes");
}
