// K2/K3 for small systems (n <= KMPX_SMALL_MAX): one CTA per problem, persistent over the batch.
//
// This file holds the cp.async variants (taken for n < 64, unaligned operands or when the buffers of the TMA kernels
// do not fit) and the 32 x 32 block helpers; the TMA-fed kernels with separate producer / auxiliary warps that serve
// the benchmark sizes are in factor_tma.cuh, which is included at the end of this translation unit.
//
//   k_potrf_small   left-looking blocked Cholesky, block columns of 32.  The update of block column k0
//                   (rows k0..n-1) by all previous columns is the only dense contraction:
//                   C(m x 32) += L[k0:n, 0:k0] * L[k0:k0+32, 0:k0]^T  on DMMA, the operand streamed once
//                   per block column through a 4-stage cp.async ring, two k-tiles per barrier (the B operand is
//                   the first 32 rows of the same tile).  Diagonal block: cta_potf2_b8 (8-column panels) and
//                   cta_trtri32; panel solve = DMMA multiply by the inverse of the diagonal block.
//   k_trtri_small   in-place inverse of the lower factor, right to left:
//                   X = -Linv22 * (L21 * Dinv)   with the (m x m, lower) Linv22 streamed on DMMA.
//   k_solve_alpha   w = Linv * y' (four rows in flight per warp), optionally alpha' = Linv^T * w.
//
// Algorithmic flops: n^3/3 each for potrf and trtri; bytes: the matrix is read and written once.
// kmpx_debug_set / kmpx_debug_phases: profiling aids (skip the MMAs of the K-loops, force the cp.async variants,
// per-phase cycle counters of CTA 0) used by tools/factor_phases.py; they do not change results when unset.
#include "dmma_tile.cuh"
#include "kmpx_internal.cuh"

namespace kmpx {

constexpr int NB = 32;       // block-column width
constexpr int FBK = 8;       // k-step of the streamed operand
constexpr int FLDT = 12;     // leading dimension of a stage tile [rows][8]   (12 = conflict-free, see dmma_tile.cuh)
constexpr int FLDP = 36;     // leading dimension of the panel / diagonal blocks [rows][32]  (4 mod 16)
constexpr int FLDB = 36;     // leading dimension of the k-major B stage tile [8][32]
constexpr int FSTAGES = 4;      // ring depth of k_potrf_small
constexpr int TSTAGES = 3;      // ring depth of k_trtri_small (its stage also carries the B tile)
constexpr int FTHREADS = 256;
constexpr int FWARPS = FTHREADS / 32;

// Cycle counters per phase, accumulated by thread 0 of CTA 0 (read with kmpx_debug_phases; profiling aid only).
__device__ long long g_phase[16];
__device__ int g_dbg;   // profiling experiments: bit 0 = skip the MMAs of the K-loops, bit 1 = skip their cp.async loads
#define PHASE_MARK(slot) do { if (blockIdx.x == 0 && tid == 0) { const long long _t = clock64(); g_phase[slot] += _t - t_phase; t_phase = _t; } } while (0)

__host__ __device__ inline size_t potrf_small_smem(int n_max) {
    const size_t rows = (size_t)((n_max + 31) / 32) * 32;
    const size_t stages = (size_t)FSTAGES * rows * FLDT * sizeof(double);
    const size_t panel = rows * FLDP * sizeof(double);
    return (stages > panel ? stages : panel) + (size_t)2 * 32 * FLDP * sizeof(double);
}
__host__ __device__ inline size_t trtri_small_smem(int n_max) {
    const size_t rows = (size_t)((n_max + 31) / 32) * 32;
    return (size_t)TSTAGES * (rows * FLDT + FBK * FLDB) * sizeof(double) + 2 * 32 * FLDP * sizeof(double);
}

// ---- 32 x 32 diagonal blocks, whole CTA, shared memory, small loops (fully unrolled single-warp versions were
//      measured to be instruction-cache bound, barrier-per-column versions latency bound) --------------------

// Barrier used inside the 32 x 32 block helpers: the whole CTA, or (NAMED) only the 256 threads of the MMA warps
// of the TMA kernels, whose producer warp never joins.
template <bool NAMED>
__device__ __forceinline__ void block_sync() {
    if (NAMED) asm volatile("bar.sync 1, 256;\n" ::: "memory");
    else __syncthreads();
}

// Reciprocal square root on the latency-critical pivot path: single-precision seed + two Newton steps in fp64
// (relative error ~1e-16) instead of the much longer library sequence; out-of-range pivots take the library call.
__device__ __forceinline__ double fast_rsqrt(double x) {
    if (!(x > 1e-30 && x < 1e30)) return rsqrt(x);
    double y = (double)rsqrtf((float)x);
    const double hx = 0.5 * x;
    y = y * (1.5 - hx * y * y);      // 22-bit seed -> ~44 bits
    y = y * (1.5 - hx * y * y);      // -> fp64 accuracy (a third step would change the last bit at most)
    return y;
}

// Blocked variant: the 32 x 32 block is processed in four 8-column panels.  The 8 x 8 diagonal block (and its
// inverse) is done by ONE thread entirely in registers (8 sequential pivots, no shared-memory round trips); the
// panel solve and the rank-8 trailing update are spread over all threads.  Three barriers per panel, 12 in total,
// instead of 32 latency-bound column steps.  S: >= 64 doubles of scratch (inverse of the 8 x 8 block).
template <bool NAMED = false>
__device__ __forceinline__ void cta_potf2_b8(double* P, double* Ld, double* S, int kb, int tid, int* bad) {
    if (tid == 0) *bad = 0;
    for (int idx = tid; idx < 32 * 32; idx += FTHREADS) Ld[(idx >> 5) * FLDP + (idx & 31)] = ((idx >> 5) == (idx & 31)) ? 1.0 : 0.0;
    block_sync<NAMED>();
    for (int base = 0; base < kb; base += 8) {
        const int w = min(8, kb - base);
        if (tid == 0) {
            double a[8][8], x[8][8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    a[i][j] = (i < w && j <= i) ? P[(base + i) * FLDP + base + j] : ((i == j) ? 1.0 : 0.0);
            int b = 0;
            double rsv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const double d = a[j][j];
                if (!(d > 0.0) && b == 0 && j < w) b = base + j + 1;
                const double rs = fast_rsqrt(d);
                rsv[j] = rs;                 // = 1 / L_jj: the inverse below needs no division
                a[j][j] = d * rs;
#pragma unroll
                for (int i = j + 1; i < 8; ++i) a[i][j] *= rs;
#pragma unroll
                for (int c = j + 1; c < 8; ++c)
#pragma unroll
                    for (int i = c; i < 8; ++i) a[i][c] -= a[i][j] * a[c][j];
            }
            if (b != 0 && *bad == 0) *bad = b;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i][c] = 0.0;
                x[c][c] = rsv[c];
#pragma unroll
                for (int i = c + 1; i < 8; ++i) {
                    double s = 0.0;
#pragma unroll
                    for (int t = c; t < i; ++t) s += a[i][t] * x[t][c];
                    x[i][c] = -s * rsv[i];
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    S[i * 8 + j] = x[i][j];
                    if (i < w && j <= i) Ld[(base + i) * FLDP + base + j] = a[i][j];
                }
        }
        block_sync<NAMED>();
        const int rb = kb - base - 8;        // rows below this 8 x 8 block
        if (rb > 0) {
            for (int idx = tid; idx < rb * 8; idx += FTHREADS) {
                const int r = idx >> 3, c = idx & 7;
                const double* arow = P + (base + 8 + r) * FLDP + base;
                double sum = 0.0;
#pragma unroll
                for (int t = 0; t < 8; ++t) if (t <= c) sum += arow[t] * S[c * 8 + t];
                Ld[(base + 8 + r) * FLDP + base + c] = sum;
            }
            block_sync<NAMED>();
            for (int idx = tid; idx < rb * rb; idx += FTHREADS) {
                const int i = idx / rb, j = idx - i * rb;
                if (j <= i) {
                    const double* li = Ld + (base + 8 + i) * FLDP + base;
                    const double* lj = Ld + (base + 8 + j) * FLDP + base;
                    double sum = 0.0;
#pragma unroll
                    for (int t = 0; t < 8; ++t) sum += li[t] * lj[t];
                    P[(base + 8 + i) * FLDP + base + 8 + j] -= sum;
                }
            }
            block_sync<NAMED>();
        }
    }
}

// X = L^{-1} for the 32 x 32 lower-triangular L (identity-padded), recursive 8 -> 16 -> 32 blocking:
// [[L00,0],[L10,L11]]^{-1} = [[X00,0],[-X11 L10 X00, X11]].  S: 256 doubles of scratch.  Five barriers.
template <bool NAMED = false>
__device__ __forceinline__ void cta_trtri32(const double* L, double* X, double* S, int tid) {
    for (int idx = tid; idx < 32 * 32; idx += FTHREADS) X[(idx >> 5) * FLDP + (idx & 31)] = 0.0;
    block_sync<NAMED>();
    if (tid < 32) {                                   // four 8x8 diagonal blocks, one thread per column
        const int base = (tid >> 3) * 8, c = tid & 7;
        double x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double v = 0.0;
            if (i == c) v = 1.0 / L[(base + i) * FLDP + base + i];
            else if (i > c) {
                double sum = 0.0;
#pragma unroll
                for (int t = 0; t < 8; ++t) if (t >= c && t < i) sum += L[(base + i) * FLDP + base + t] * x[t];
                v = -sum / L[(base + i) * FLDP + base + i];
            }
            x[i] = v;
            X[(base + i) * FLDP + base + c] = v;
        }
    }
    block_sync<NAMED>();
    if (tid < 128) {                                  // level 16: T = L10 X00 for both 16-blocks
        const int pb = (tid >> 6) * 16, r = (tid >> 3) & 7, c = tid & 7;
        double sum = 0.0;
#pragma unroll
        for (int t = 0; t < 8; ++t) sum += L[(pb + 8 + r) * FLDP + pb + t] * X[(pb + t) * FLDP + pb + c];
        S[tid] = sum;
    }
    block_sync<NAMED>();
    if (tid < 128) {                                  // X10 = -X11 T
        const int q = tid >> 6, pb = q * 16, r = (tid >> 3) & 7, c = tid & 7;
        double sum = 0.0;
#pragma unroll
        for (int t = 0; t < 8; ++t) sum += X[(pb + 8 + r) * FLDP + pb + 8 + t] * S[q * 64 + t * 8 + c];
        X[(pb + 8 + r) * FLDP + pb + c] = -sum;
    }
    block_sync<NAMED>();
    {                                                 // level 32: T = L10 X00 (16 x 16), one output per thread
        const int r = tid >> 4, c = tid & 15;
        double sum = 0.0;
#pragma unroll
        for (int t = 0; t < 16; ++t) sum += L[(16 + r) * FLDP + t] * X[t * FLDP + c];
        S[tid] = sum;
    }
    block_sync<NAMED>();
    {
        const int r = tid >> 4, c = tid & 15;
        double sum = 0.0;
#pragma unroll
        for (int t = 0; t < 16; ++t) sum += X[(16 + r) * FLDP + 16 + t] * S[t * 16 + c];
        X[(16 + r) * FLDP + c] = -sum;
    }
    block_sync<NAMED>();
}

// X(32 x 32) = A(32 x 32) * op(D) on DMMA for one warp.  A rows come from `src` (row-major, leading dimension
// lds, shared or global), D is a 32x32 block in shared memory (leading dimension FLDP); X goes to `dst`
// (may alias src: every A fragment is in registers before the first store).  Rows >= rows_valid are skipped.
//   D_TRANS: X = A * D^T   (B[k][n] = D[n][k])        else: X = A * D   (B[k][n] = D[k][n])
template <bool D_TRANS>
__device__ __forceinline__ void warp_chunk_mul32(const double* src, int lds, double* dst, int ldd, const double* D,
                                                 int rows_valid, double scale, int lane) {
    const int g = lane >> 2, t4 = lane & 3;
    double a[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = i * 8 + g;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) a[i][ks] = (r < rows_valid) ? src[(size_t)r * lds + ks * 4 + t4] : 0.0;
    }
    double acc[4][4][2];
    zero_acc(acc);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        double bf[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) bf[j] = D_TRANS ? D[(j * 8 + g) * FLDP + ks * 4 + t4] : D[(ks * 4 + t4) * FLDP + j * 8 + g];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma884(acc[i][j], a[i][ks], bf[j]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = i * 8 + g;
        if (r < rows_valid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                dst[(size_t)r * ldd + j * 8 + 2 * t4] = scale * acc[i][j][0];
                dst[(size_t)r * ldd + j * 8 + 2 * t4 + 1] = scale * acc[i][j][1];
            }
        }
    }
}

__global__ void __launch_bounds__(FTHREADS) k_potrf_small(FactorArgs p) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int rows_max = ((p.n_max + 31) / 32) * 32;
    const size_t stage_elems = (size_t)rows_max * FLDT;
    const size_t stages_bytes = (size_t)FSTAGES * stage_elems;
    const size_t panel_elems = (size_t)rows_max * FLDP;
    double* stage0 = smem;
    double* panel = smem;                                                     // aliases the stage ring
    double* Dinv = smem + (stages_bytes > panel_elems ? stages_bytes : panel_elems);    // [32][FLDP] inverse of the diagonal block
    double* Ld = Dinv + 32 * FLDP;                                                          // [32][FLDP] factor of the diagonal block
    __shared__ int sh_bad;
    __shared__ double scratch[256];

    for (int b = blockIdx.x; b < p.batch; b += gridDim.x) {
        const int n = p.n_sys ? p.n_sys[b] : p.n_max;
        double* A = p.A + (size_t)b * p.strideA;
        int bad_total = 0;
        long long t_phase = clock64();
        for (int k0 = 0; k0 < n; k0 += NB) {
            const int kb = min(NB, n - k0);
            const int m = n - k0;
            const int nch = (m + 31) >> 5;
            const int rows_t = nch * 32;
            double acc[2][4][4][2];
            zero_acc(acc[0]);
            zero_acc(acc[1]);
            const int nk = k0 / FBK;
            if (nk > 0) {
                const double* gA = A + (size_t)k0 * p.lda;
                // two k-tiles per barrier (nk is a multiple of 4): tiles kt, kt+1 are consumed while kt+2, kt+3 fly
                // every thread copies the same (row, 16-byte chunk) slots in each k-tile: row = tid/4 + 64*i, chunk = tid%4;
                // pointers are formed once per block column, the per-tile work is one add and one cp.async per slot
                const int lr = tid >> 2, lq = tid & 3;
                const double* lsrc = gA + (size_t)lr * p.lda + 2 * lq;
                const int ldst = lr * FLDT + 2 * lq;
                auto issue2 = [&](int kt) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (kt + u < nk && !(g_dbg & 2)) {
                            double* st = stage0 + ((kt + u) % FSTAGES) * stage_elems + ldst;
                            const double* src = lsrc + (kt + u) * FBK;
                            for (int r = lr; r < rows_t; r += FTHREADS / 4) {
                                const bool ok = r < m;
                                cp_async16(st, ok ? src : gA, ok ? 16 : 0);
                                st += (FTHREADS / 4) * FLDT;
                                src += (size_t)(FTHREADS / 4) * p.lda;
                            }
                        }
                    }
                    cp_async_commit();
                };
                issue2(0);
                for (int kt = 0; kt < nk; kt += 2) {
                    cp_async_wait<0>();
                    __syncthreads();
                    issue2(kt + 2);
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const double* tile = stage0 + ((kt + u) % FSTAGES) * stage_elems;
#pragma unroll
                        for (int kk = 0; kk < FBK; kk += 4) {
                            if (g_dbg & 1) continue;
                            if (warp < nch) warp_mma_k4<4, 4, false>(acc[0], tile + (warp * 32) * FLDT + kk, FLDT, tile + kk, FLDT, lane);
                            if (warp + FWARPS < nch)
                                warp_mma_k4<4, 4, false>(acc[1], tile + ((warp + FWARPS) * 32) * FLDT + kk, FLDT, tile + kk, FLDT, lane);
                        }
                    }
                }
                cp_async_wait<0>();
            }
            __syncthreads();   // the stage ring is dead from here on: the panel may overwrite it
            PHASE_MARK(0);
            // panel <- A[k0:n, k0:k0+kb] (coalesced 16-byte cp.async, all in flight at once), then panel -= acc
            load_tile_async<NB, FLDP, false>(panel, A + (size_t)k0 * p.lda + k0, p.lda, m, m, k0, k0, k0 + kb, tid, FTHREADS);
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int ch = warp + c * FWARPS;
                if (ch < nch) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = ch * 32 + i * 8 + g;
                        if (r < m) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                double* dst = panel + r * FLDP + j * 8 + 2 * t4;
                                dst[0] -= acc[c][i][j][0];
                                dst[1] -= acc[c][i][j][1];
                            }
                        }
                    }
                }
            }
            __syncthreads();
            PHASE_MARK(1);
            cta_potf2_b8(panel, Ld, scratch, kb, tid, &sh_bad);
            PHASE_MARK(2);
            if (m > NB) cta_trtri32(Ld, Dinv, scratch, tid);
            for (int idx = tid; idx < NB * NB; idx += FTHREADS) {
                const int r = idx >> 5, c = idx & 31;
                if (r < kb && c <= r) panel[r * FLDP + c] = Ld[r * FLDP + c];
            }
            __syncthreads();
            PHASE_MARK(3);
            if (sh_bad != 0 && bad_total == 0) bad_total = k0 + sh_bad;
            // panel rows below the diagonal block:  X = A21 * L_kk^{-T}   (only exists when kb == 32)
            if (m > NB) {
                for (int ch = 1 + warp; ch < nch; ch += FWARPS)
                    warp_chunk_mul32<true>(panel + (size_t)ch * 32 * FLDP, FLDP, panel + (size_t)ch * 32 * FLDP, FLDP, Dinv,
                                           m - ch * 32, 1.0, lane);
            }
            __syncthreads();
            PHASE_MARK(4);
            for (int idx = tid; idx < m * NB; idx += FTHREADS) {
                const int r = idx >> 5, c = idx & 31;
                if (c < kb && (r >= NB || c <= r)) A[(size_t)(k0 + r) * p.lda + k0 + c] = panel[r * FLDP + c];
            }
            __threadfence();   // the block column is read back through cp.async (L2) by later iterations
            __syncthreads();
            PHASE_MARK(5);
        }
        if (tid == 0 && p.info) {
            if (!p.info_accumulate) p.info[b] = bad_total;
            else if (bad_total != 0 && p.info[b] == 0) p.info[b] = p.info_offset + bad_total;
        }
    }
}

__global__ void __launch_bounds__(FTHREADS) k_trtri_small(FactorArgs p) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int rows_max = ((p.n_max + 31) / 32) * 32;
    const size_t stage_elems = (size_t)rows_max * FLDT + FBK * FLDB;
    double* stage0 = smem;
    double* Dsm = smem + TSTAGES * stage_elems;
    double* Dinv = Dsm + 32 * FLDP;
    __shared__ double scratch[256];

    for (int b = blockIdx.x; b < p.batch; b += gridDim.x) {
        const int n = p.n_sys ? p.n_sys[b] : p.n_max;
        double* L = p.A + (size_t)b * p.strideA;
        const int nblk = (n + NB - 1) / NB;
        long long t_phase = clock64();
        for (int jb = nblk - 1; jb >= 0; --jb) {
            const int k0 = jb * NB;
            const int kb = min(NB, n - k0);
            const int m = n - k0 - kb;            // rows below the diagonal block
            for (int idx = tid; idx < NB * NB; idx += FTHREADS) {
                const int r = idx >> 5, c = idx & 31;
                Dsm[r * FLDP + c] = (r < kb && c <= r) ? L[(size_t)(k0 + r) * p.lda + k0 + c] : ((r == c) ? 1.0 : 0.0);
            }
            __syncthreads();
            cta_trtri32(Dsm, Dinv, scratch, tid);
            PHASE_MARK(8);
            if (m > 0) {
                // T = L21 * Dinv, in place in global memory, one 32-row chunk per warp (kb == 32 here)
                {
                    const int nch_t = (m + 31) >> 5;
                    for (int ch = warp; ch < nch_t; ch += FWARPS) {
                        double* rows = L + (size_t)(k0 + NB + ch * 32) * p.lda + k0;
                        warp_chunk_mul32<false>(rows, p.lda, rows, p.lda, Dinv, m - ch * 32, 1.0, lane);
                    }
                }
                __threadfence();
                __syncthreads();
                PHASE_MARK(9);
                // X = -Linv22 * T   (Linv22 lower-triangular m x m, streamed; T streamed k-major)
                const int nch = (m + 31) >> 5;
                const int rows_t = nch * 32;
                const int nk = (m + FBK - 1) / FBK;
                const int r_g0 = k0 + NB;                                  // first trailing row / column
                const double* gA = L + (size_t)r_g0 * p.lda + r_g0;          // Linv22
                const double* gT = L + (size_t)r_g0 * p.lda + k0;            // T (m x 32)
                double acc[2][4][4][2];
                zero_acc(acc[0]);
                zero_acc(acc[1]);
                auto issue = [&](int kt) {
                    double* st = stage0 + (kt % TSTAGES) * stage_elems;
                    load_tile_async<FBK, FLDT, true>(st, gA + kt * FBK, p.lda, rows_t, m, r_g0, r_g0 + kt * FBK, r_g0 + m,
                                                     tid, FTHREADS);
                    // B tile: rows kt*8 .. +8 of T, 32 columns -> [8][FLDB]
                    double* bt = st + (size_t)rows_max * FLDT;
                    for (int idx = tid; idx < FBK * (NB / 2); idx += FTHREADS) {
                        const int r = idx / (NB / 2), q = idx - r * (NB / 2);
                        const int tr = kt * FBK + r;
                        cp_async16(bt + r * FLDB + 2 * q, tr < m ? gT + (size_t)tr * p.lda + 2 * q : gT, tr < m ? 16 : 0);
                    }
                };
#pragma unroll
                for (int s = 0; s < TSTAGES - 1; ++s) { if (s < nk) issue(s); cp_async_commit(); }
                for (int kt = 0; kt < nk; ++kt) {
                    cp_async_wait<TSTAGES - 2>();
                    __syncthreads();
                    if (kt + TSTAGES - 1 < nk) issue(kt + TSTAGES - 1);
                    cp_async_commit();
                    const double* tile = stage0 + (kt % TSTAGES) * stage_elems;
                    const double* bt = tile + (size_t)rows_max * FLDT;
#pragma unroll
                    for (int kk = 0; kk < FBK; kk += 4) {
                        // a 32-row chunk ch only needs k < 32*(ch+1)
                        if (warp < nch && kt * FBK + kk < 32 * (warp + 1))
                            warp_mma_k4<4, 4, true>(acc[0], tile + (warp * 32) * FLDT + kk, FLDT, bt + kk * FLDB, FLDB, lane);
                        if (warp + FWARPS < nch && kt * FBK + kk < 32 * (warp + FWARPS + 1))
                            warp_mma_k4<4, 4, true>(acc[1], tile + ((warp + FWARPS) * 32) * FLDT + kk, FLDT, bt + kk * FLDB, FLDB, lane);
                    }
                }
                cp_async_wait<0>();
                __syncthreads();   // every read of T is complete: X may overwrite it
                PHASE_MARK(10);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int ch = warp + c * FWARPS;
                    if (ch < nch) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int r = ch * 32 + i * 8 + g;
                            if (r < m) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    double2 v = make_double2(-acc[c][i][j][0], -acc[c][i][j][1]);
                                    *reinterpret_cast<double2*>(L + (size_t)(r_g0 + r) * p.lda + k0 + j * 8 + 2 * t4) = v;
                                }
                            }
                        }
                    }
                }
            }
            for (int idx = tid; idx < NB * NB; idx += FTHREADS) {
                const int r = idx >> 5, c = idx & 31;
                if (r < kb && c <= r) L[(size_t)(k0 + r) * p.lda + k0 + c] = Dinv[r * FLDP + c];
            }
            __threadfence();
            __syncthreads();
            PHASE_MARK(11);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// symmetric: w = Linv*y', alpha' = Linv^T*w.   general: alpha' = Ainv*y' (w untouched).
__global__ void __launch_bounds__(256) k_solve_alpha(AlphaArgs p) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nmax_sys = sys_size(p.n_max, p.O, p.layout);
    double* ys = smem;                 // y' in system order
    double* ws = smem + nmax_sys;      // w
    for (int b = blockIdx.x; b < p.batch; b += gridDim.x) {
        const int N = p.n_pts ? p.n_pts[b] : p.n_max;
        const int O = p.O;
        const int Ns = (p.layout == KMPX_LAYOUT_POINT) ? N : comp_stride(N);
        const int n = sys_size(N, O, p.layout);
        const double* F = p.F + (size_t)b * p.strideA;
        const double* y = p.y + (size_t)b * p.n_max * O;
        for (int r = tid; r < n; r += blockDim.x) ys[r] = 0.0;     // identity padding rows carry y' = 0
        __syncthreads();
        for (int idx = tid; idx < N * O; idx += blockDim.x) {
            const int i = idx / O, a = idx - i * O;
            const int r = (p.layout == KMPX_LAYOUT_POINT) ? idx : a * Ns + i;
            ys[r] = y[idx];
        }
        __syncthreads();
        // four rows per warp and pass: their loads are in flight together (one row at a time is latency bound)
        for (int r0 = warp * 4; r0 < n; r0 += 32) {
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            const int cmax = p.general ? n : min(n, r0 + 4);
            for (int c = lane; c < cmax; c += 32) {
                const double yv = ys[c];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int r = min(r0 + q, n - 1);
                    const double v = F[(size_t)r * p.lda + c];
                    if (p.general || c <= r0 + q) acc[q] += v * yv;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double tot = warp_sum(acc[q]);
                if (lane == 0 && r0 + q < n) ws[r0 + q] = tot;
            }
        }
        __syncthreads();
        if (!p.general) {
            if (p.w) for (int r = tid; r < n; r += blockDim.x) p.w[(size_t)b * p.ldw + r] = ws[r];
            if (p.alpha) {         // second pass only when alpha itself is wanted (the fused predict needs w only)
                for (int c = tid; c < n; c += blockDim.x) {
                    double acc = 0.0;
                    for (int r = c; r < n; ++r) acc += F[(size_t)r * p.lda + c] * ws[r];
                    ys[c] = acc;       // each thread owns entry c
                }
            }
            __syncthreads();
        }
        const double* res = p.general ? ws : ys;
        if (p.alpha) {
            double* al = p.alpha + (size_t)b * p.n_max * O;
            for (int idx = tid; idx < N * O; idx += blockDim.x) {
                const int i = idx / O, a = idx - i * O;
                const int r = (p.layout == KMPX_LAYOUT_POINT) ? idx : a * Ns + i;
                al[idx] = res[r];
            }
        }
        if (p.general && p.w) for (int r = tid; r < n; r += blockDim.x) p.w[(size_t)b * p.ldw + r] = ws[r];
        __syncthreads();
    }
}

}  // namespace kmpx
#include "factor_tma.cuh"
#include "factor_struct.cuh"
#include "factor_wpp.cuh"
namespace kmpx {

int debug_set(int flags) { g_host_dbg = flags; return check_cuda(cudaMemcpyToSymbol(g_dbg, &flags, sizeof(int)), "cudaMemcpyToSymbol"); }

int debug_phases(long long* out, int reset) {
    KMPX_TRY(check_cuda(cudaMemcpyFromSymbol(out, g_phase, sizeof(long long) * 16), "cudaMemcpyFromSymbol"));
    if (reset) { long long z[16] = {0}; KMPX_TRY(check_cuda(cudaMemcpyToSymbol(g_phase, z, sizeof(z)), "cudaMemcpyToSymbol")); }
    return 0;
}

int launch_potrf_small(double* A, int lda, long long strideA, const int* n_sys, int n_max, int batch, int* info,
                       int info_offset, int info_accumulate, cudaStream_t st) {
    if (tma_applicable(A, lda, strideA, n_max, batch, false)) {
        CUtensorMap map;
        KMPX_TRY(make_batch_map(&map, A, lda, strideA, n_max, batch, TBOX));
        const TmaSmem lay = tma_smem_layout(n_max, false, kMaxDynSmem);
        static SmemCache tma_smem_set;
        KMPX_TRY(ensure_dyn_smem(k_potrf_tma, lay.total, tma_smem_set, "k_potrf_tma"));
        FactorArgs p{A, lda, strideA, n_sys, n_max, batch, info, info_offset, info_accumulate};
        k_potrf_tma<<<batch < num_sms() ? batch : num_sms(), TTHREADS, lay.total, st>>>(map, p, lay);
        return check_launch("k_potrf_tma");
    }
    const size_t smem = potrf_small_smem(n_max);
    static SmemCache smem_set;
    KMPX_TRY(ensure_dyn_smem(k_potrf_small, smem, smem_set, "k_potrf_small"));
    FactorArgs p{A, lda, strideA, n_sys, n_max, batch, info, info_offset, info_accumulate};
    const int per_sm = smem <= 110 * 1024 ? 2 : 1;
    const int grid = batch < num_sms() * per_sm ? batch : num_sms() * per_sm;
    k_potrf_small<<<grid, FTHREADS, smem, st>>>(p);
    return check_launch("k_potrf_small");
}

int launch_trtri_small(double* L, int lda, long long strideA, const int* n_sys, int n_max, int batch, cudaStream_t st) {
    if (tma_applicable(L, lda, strideA, n_max, batch, true)) {
        CUtensorMap map;
        KMPX_TRY(make_batch_map(&map, L, lda, strideA, n_max, batch, TBOX));
        const TmaSmem lay = tma_smem_layout(n_max, true, kMaxDynSmem);
        static SmemCache tma_smem_set;
        KMPX_TRY(ensure_dyn_smem(k_trtri_tma, lay.total, tma_smem_set, "k_trtri_tma"));
        FactorArgs p{L, lda, strideA, n_sys, n_max, batch, nullptr, 0, 0};
        k_trtri_tma<<<batch < num_sms() ? batch : num_sms(), TTHREADS, lay.total, st>>>(map, p, lay);
        return check_launch("k_trtri_tma");
    }
    const size_t smem = trtri_small_smem(n_max);
    static SmemCache smem_set;
    KMPX_TRY(ensure_dyn_smem(k_trtri_small, smem, smem_set, "k_trtri_small"));
    FactorArgs p{L, lda, strideA, n_sys, n_max, batch, nullptr, 0, 0};
    const int per_sm = smem <= 110 * 1024 ? 2 : 1;
    const int grid = batch < num_sms() * per_sm ? batch : num_sms() * per_sm;
    k_trtri_small<<<grid, FTHREADS, smem, st>>>(p);
    return check_launch("k_trtri_small");
}

int launch_solve_alpha(const AlphaArgs& p, cudaStream_t st) {
    const size_t smem = (size_t)2 * sys_size(p.n_max, p.O, p.layout) * sizeof(double);
    static SmemCache smem_set;
    KMPX_TRY(ensure_dyn_smem(k_solve_alpha, smem, smem_set, "k_solve_alpha"));
    const int grid = p.batch < num_sms() * 4 ? p.batch : num_sms() * 4;
    k_solve_alpha<<<grid, 256, smem, st>>>(p);
    return check_launch("k_solve_alpha");
}

}  // namespace kmpx

extern "C" int kmpx_factor_struct_applicable(int n_max_pts, int I, int O) {
    return kmpx::factor_struct_applicable(n_max_pts, I, O, 2, 0, nullptr) ? 1 : 0;
}
extern "C" int kmpx_factor_struct(const double* s, const double* sigma, const int* n_pts, double* F, int lda, long long strideF,
                                  int n_max_pts, int I, int O, int batch, double sigma_f, double lambda, int* info, void* stream) {
    using namespace kmpx;
    if (!s || !sigma || !F) return fail_arg(1, "s / sigma / F is NULL");
    if (batch <= 0) return fail_arg(10, "batch must be > 0");
    if (!factor_struct_applicable(n_max_pts, I, O, lda, strideF, F))
        return fail_arg(7, "kmpx_factor_struct covers the plain kernel with O <= 2, aligned F and N <= ~204 (kmpx_factor_struct_applicable)");
    if (lda < O * comp_stride(n_max_pts)) return fail_arg(5, "lda < system size");
    StructArgs p{s, sigma, n_pts, F, lda, strideF, n_max_pts, I, O, batch, sigma_f, lambda, info, 0};
    return launch_factor_struct(p, (cudaStream_t)stream);
}
extern "C" int kmpx_factor_wpp_applicable(const double* A, int lda, long long strideA, int n_max, int batch) {
    return kmpx::wpp_applicable(A, lda, strideA, n_max, batch) ? 1 : 0;
}
extern "C" int kmpx_potrf_wpp(double* A, int lda, long long strideA, const int* n_sys, int n_max, int batch, int* info,
                              int struct_o, void* stream) {
    using namespace kmpx;
    if (!A) return fail_arg(1, "matrix is NULL");
    if (batch <= 0) return fail_arg(6, "batch must be > 0");
    if (lda < n_max) return fail_arg(2, "lda < n_max");
    if (struct_o != 0 && struct_o != 2) return fail_arg(8, "struct_o must be 0 or 2");
    if (!wpp_applicable(A, lda, strideA, n_max, batch))
        return fail_arg(5, "kmpx_potrf_wpp needs 64 <= n_max <= KMPX_SMALL_MAX, even lda / strideA and a 16-byte aligned matrix");
    return launch_potrf_wpp(A, lda, strideA, n_sys, n_max, batch, info, struct_o, (cudaStream_t)stream);
}
extern "C" int kmpx_trtri_wpp(double* L, int lda, long long strideA, const int* n_sys, int n_max, int batch, int struct_o, void* stream) {
    using namespace kmpx;
    if (!L) return fail_arg(1, "matrix is NULL");
    if (batch <= 0) return fail_arg(6, "batch must be > 0");
    if (lda < n_max) return fail_arg(2, "lda < n_max");
    if (struct_o != 0 && struct_o != 2) return fail_arg(7, "struct_o must be 0 or 2");
    if (!wpp_applicable(L, lda, strideA, n_max, batch))
        return fail_arg(5, "kmpx_trtri_wpp needs 64 <= n_max <= KMPX_SMALL_MAX, even lda / strideA and a 16-byte aligned matrix");
    return launch_trtri_wpp(L, lda, strideA, n_sys, n_max, batch, struct_o, WppSolve{nullptr, nullptr, 0, 0, nullptr, 0}, (cudaStream_t)stream);
}
extern "C" int kmpx_trtri_wpp_solve(double* L, int lda, long long strideA, const int* n_sys, int n_max, int batch, int struct_o,
                                    const double* y, const int* n_pts, int n_max_pts, int O, double* w, int ldw, void* stream) {
    using namespace kmpx;
    if (!L) return fail_arg(1, "matrix is NULL");
    if (batch <= 0) return fail_arg(6, "batch must be > 0");
    if (lda < n_max) return fail_arg(2, "lda < n_max");
    if (struct_o != 0 && struct_o != 2) return fail_arg(7, "struct_o must be 0 or 2");
    if (!y) return fail_arg(8, "y is NULL");
    if (O <= 0 || n_max_pts <= 0 || comp_stride(n_max_pts) * O != n_max) return fail_arg(10, "n_max must be the component-major system size of (n_max_pts, O)");
    if (!n_sys != !n_pts) return fail_arg(9, "n_sys and n_pts must both be given (ragged batch) or both be NULL");
    if (!w) return fail_arg(12, "w is NULL");
    if (ldw < n_max) return fail_arg(13, "ldw < n_max");
    if (!wpp_applicable(L, lda, strideA, n_max, batch))
        return fail_arg(5, "kmpx_trtri_wpp_solve needs 64 <= n_max <= KMPX_SMALL_MAX, even lda / strideA and a 16-byte aligned matrix");
    return launch_trtri_wpp(L, lda, strideA, n_sys, n_max, batch, struct_o, WppSolve{y, n_pts, n_max_pts, O, w, ldw}, (cudaStream_t)stream);
}
extern "C" int kmpx_debug_phases(long long* out_host, int reset) { return kmpx::debug_phases(out_host, reset); }
extern "C" int kmpx_debug_set(int flags) { return kmpx::debug_set(flags); }
