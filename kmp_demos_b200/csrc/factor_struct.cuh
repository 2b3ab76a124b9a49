// Structured, shared-memory-resident fit of the small KMP systems (textually included by factor_small.cu after
// factor_tma.cuh): kernel-matrix build + Cholesky + triangular inverse in ONE kernel for the plain RBF kernel with
// O <= 2 outputs and N <= ~204 stored points - the shape of cfg 1 / cfg 3 (KMP.py:146-157).
//
// In the component-major order the system is  A = [[A00, A10^T], [A10, A11]]  with  Aaa = K + lambda * D_aa  (N x N, SPD)
// and  A10 = lambda * D_10  DIAGONAL (the cross-covariances of a point couple only its own two outputs).  Its Cholesky
// factor and the inverse of that factor follow from N x N pieces:
//     L11 = chol(A00)        X11 = L11^-1                      (L^-1 top-left)
//     G   = X11^T X11 = A00^-1,   W = D_10' G  (D_10' = lambda D_10)
//     S   = A11 - W D_10'    L22 = chol(S)   X22 = L22^-1       (L^-1 bottom-right)
//     X21 = -X22 W                                              (L^-1 bottom-left)
// i.e. exactly the blocks of the same factorisation k_potrf_tma / k_trtri_tma compute (same ordering, same pivots up to
// rounding), at half the flops - and, what matters more for these sizes, with every sequentially dependent step on an
// N x N lower triangle that STAYS IN SHARED MEMORY (block-column packed, 190 KB at N = 202): no block column is written
// to global memory, fenced and streamed back through TMA between two steps, which is where the generic kernels spend
// two thirds of their time.  The matrix is generated in place from (s, Sigma): K1 is fused away.
//
// Storage T: block column c (32 columns) holds the rows 32 c .. Rs-1 of the lower triangle, row-major with 32 doubles per
// row; the 16-byte chunk q of storage row r lives at chunk q ^ ((r & 3) << 1), which makes every DMMA fragment read used
// here (rows x k, and the transposed k x rows) conflict-free.  Rows >= N and the strict upper part of the diagonal
// blocks hold zeros.  One CTA (8 warps) per problem, persistent over the batch; the 32 x 32 diagonal blocks are factored
// and inverted by the routines of factor_tma.cuh on warps 0-2.
#pragma once

namespace kmpx {

constexpr int SB = 32;                       // block size
constexpr int STHREADS = 256;
constexpr int SWARPS = STHREADS / 32;

struct StructArgs {
    const double* s; const double* sigma; const int* n_pts;     // [B][n_max][I], [B][n_max][O][O]
    double* F; int lda; long long strideF;                       // out: L^-1, component-major, (O Ns) x lda per problem
    int n_max, I, O, batch; double sigma_f, lambda;
    int* info;
    int Rs;                                                      // storage rows of block column 0 (n_max rounded up to even)
};

__host__ __device__ inline int st_rows_before(int c, int Rs) { return c * Rs - 16 * c * (c - 1); }     // sum_{c'<c} (Rs - 32 c')
__host__ __device__ inline size_t struct_smem_bytes(int Rs, int I) {
    const int nb = (Rs + SB - 1) / SB;
    size_t doubles = (size_t)st_rows_before(nb, Rs) * SB;                       // T (the last block column may be ragged)
    doubles += 3 * 32 * FLDP + 256;                                             // Pb, Ld, Xb, scratch
    doubles += (size_t)Rs * (I + 3) + kExpTabN + 8;                             // s, d10, d00, d11, exp table, flags
    return doubles * sizeof(double) + 16;
}

#define ST_MARK(slot) do { if (blockIdx.x == 0 && tid == 0) { const long long _t = clock64(); g_phase[slot] += _t - *x.t_phase; *x.t_phase = _t; } } while (0)

struct StructCtx {
    double* T; double* Pb; double* Ld; double* Xb; double* scratch;
    double* s_sh; double* d10; double* daa[2]; double* etab; int* bad;
    int Rs, N, nb, I;
    double sigma_f;
    long long* t_phase;          // profiling: last mark of thread 0 (kmpx_debug_phases)
};

__device__ __forceinline__ int st_swz(int rr, int cc) { return ((((cc >> 1) ^ ((rr & 3) << 1)) << 1) | (cc & 1)); }
// pointer to storage row rr (relative to the block column) of block column c
__device__ __forceinline__ double* st_row(const StructCtx& x, int c, int rr) { return x.T + ((size_t)st_rows_before(c, x.Rs) + rr) * SB; }
__device__ __forceinline__ int st_rows(const StructCtx& x, int c) { return x.Rs - SB * c; }

// ---- fragment loads from T.  `base` = st_row(c, r0) of the first row of a 32-row block, `avail` = storage rows from r0 on
//      (rows beyond are masked to zero, their address is clamped).  N form: element (row 8 tile + g, column 4 s + t4);
//      T form: element (row 4 s + t4, column 8 tile + g).
__device__ __forceinline__ double st_ldN(const double* base, int avail, int tile, int s, int g, int t4) {
    const int r = 8 * tile + g;
    const int rc = r < avail ? r : 0;
    const double v = base[rc * SB + ((((2 * s + (t4 >> 1)) ^ ((g & 3) << 1)) << 1) | (t4 & 1))];
    return r < avail ? v : 0.0;
}
__device__ __forceinline__ double st_ldT(const double* base, int avail, int tile, int s, int g, int t4) {
    const int r = 4 * s + t4;
    const int rc = r < avail ? r : 0;
    const int cc = 8 * tile + g;
    const double v = base[rc * SB + ((((cc >> 1) ^ (t4 << 1)) << 1) | (cc & 1))];
    return r < avail ? v : 0.0;
}
// Per-lane element offsets of the two fragment forms, computed once per thread, so that a fragment load from a FULL
// 32-row block is one LDS with an immediate offset:  N form  base[cs[s] + 256 tile],  T form  base[ct[tile] + 128 s].
struct StLane { int cs[8]; int ct[4]; };
__device__ __forceinline__ StLane st_lane_offsets(int lane) {
    const int g = lane >> 2, t4 = lane & 3;
    StLane L;
#pragma unroll
    for (int s = 0; s < 8; ++s) L.cs[s] = g * SB + ((((2 * s + (t4 >> 1)) ^ ((g & 3) << 1)) << 1) | (t4 & 1));
#pragma unroll
    for (int i = 0; i < 4; ++i) { const int cc = 8 * i + g; L.ct[i] = t4 * SB + ((((cc >> 1) ^ (t4 << 1)) << 1) | (cc & 1)); }
    return L;
}
// acc (32 x 32, accumulator layout) += sign * op(A) * op(B) for two 32 x 32 blocks of T.
//   TA = false: A[m][k] = Ablk(row m, col k);   TA = true: A[m][k] = Ablk(row k, col m)
//   TB = true : B[k][n] = Bblk(row n, col k);   TB = false: B[k][n] = Bblk(row k, col n)
// Blocks with all 32 storage rows present take the unmasked loads; ragged blocks (the last rows of a block column) mask.
template <bool TA, bool TB>
__device__ __forceinline__ void st_blk_mma(double (&acc)[4][4][2], const double* Ablk, int availA, const double* Bblk, int availB,
                                           double sign, int lane, const StLane& L) {
    const int g = lane >> 2, t4 = lane & 3;
    if (availA >= SB && availB >= SB) {
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sign * (TA ? Ablk[L.ct[i] + 4 * SB * s] : Ablk[L.cs[s] + 8 * SB * i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = TB ? Bblk[L.cs[s] + 8 * SB * j] : Bblk[L.ct[j] + 4 * SB * s];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j], a[i], b[j]);
        }
        return;
    }
#pragma unroll 1
    for (int s = 0; s < 8; ++s) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = sign * (TA ? st_ldT(Ablk, availA, i, s, g, t4) : st_ldN(Ablk, availA, i, s, g, t4));
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = TB ? st_ldN(Bblk, availB, j, s, g, t4) : st_ldT(Bblk, availB, j, s, g, t4);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma884(acc[i][j], a[i], b[j]);
    }
}
// a 32 x 32 block of T <-> accumulator layout (element (8 i + g, 8 j + 2 t4 + e)); rows beyond `avail` read as zero / are not stored
__device__ __forceinline__ void st_blk_load(double (&acc)[4][4][2], const double* blk, int avail, int lane) {
    const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = 8 * i + g;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double2 v = make_double2(0.0, 0.0);
            if (r < avail) v = *reinterpret_cast<const double2*>(blk + r * SB + (((4 * j + t4) ^ ((g & 3) << 1)) << 1));
            acc[i][j][0] = v.x; acc[i][j][1] = v.y;
        }
    }
}
__device__ __forceinline__ void st_blk_store(const double (&acc)[4][4][2], double* blk, int avail, int lane) {
    const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = 8 * i + g;
        if (r >= avail) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            *reinterpret_cast<double2*>(blk + r * SB + (((4 * j + t4) ^ ((g & 3) << 1)) << 1)) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
}

// diagonal block j of T -> Pb (FLDP layout, lower part, zeros elsewhere); whole CTA
__device__ __forceinline__ void st_diag_to_buf(const StructCtx& x, int j, double* dst, int kb, int tid) {
    const double* blk = st_row(x, j, 0);
    for (int idx = tid; idx < 32 * 32; idx += STHREADS) {
        const int r = idx >> 5, c = idx & 31;
        dst[r * FLDP + c] = (r < kb && c <= r) ? blk[r * SB + st_swz(r, c)] : 0.0;
    }
}
// FLDP buffer (lower triangle) -> diagonal block j of T (zeros above the diagonal and beyond kb)
__device__ __forceinline__ void st_buf_to_diag(const StructCtx& x, int j, const double* src, int kb, int tid) {
    double* blk = st_row(x, j, 0);
    const int avail = st_rows(x, j) < 32 ? st_rows(x, j) : 32;
    for (int idx = tid; idx < avail * 32; idx += STHREADS) {
        const int r = idx >> 5, c = idx & 31;
        blk[r * SB + st_swz(r, c)] = (r < kb && c <= r) ? src[r * FLDP + c] : 0.0;
    }
}

// strided variants for a subset of `nth` threads (t = 0 .. nth-1): diagonal block j <-> FLDP buffer
__device__ __forceinline__ void st_diag_to_buf_n(const StructCtx& x, int j, double* dst, int kb, int t, int nth) {
    const double* blk = st_row(x, j, 0);
    for (int idx = t; idx < 32 * 32; idx += nth) {
        const int r = idx >> 5, c = idx & 31;
        dst[r * FLDP + c] = (r < kb && c <= r) ? blk[r * SB + st_swz(r, c)] : 0.0;
    }
}
__device__ __forceinline__ void st_buf_to_diag_n(const StructCtx& x, int j, const double* src, int kb, int t, int nth) {
    double* blk = st_row(x, j, 0);
    const int avail = st_rows(x, j) < 32 ? st_rows(x, j) : 32;
    for (int idx = t; idx < avail * 32; idx += nth) {
        const int r = idx >> 5, c = idx & 31;
        blk[r * SB + st_swz(r, c)] = (r < kb && c <= r) ? src[r * FLDP + c] : 0.0;
    }
}
// Diagonal block j on the three auxiliary warps (t = 0..95): factor, invert, put L_jj back into T and the inverse X_jj
// (a final block of L^-1, and what st_trtri starts from) into global memory at xdiag (leading dimension lda).
__device__ __forceinline__ void st_diag_step(const StructCtx& x, int j, int t, double* xdiag, int lda, int* bad_total, int pivot_offset) {
    const int kb = min(SB, x.N - SB * j);
    st_diag_to_buf_n(x, j, x.Pb, kb, t, AUXT);
    aux_warps_sync();
    aux_potf2(x.Pb, x.Ld, x.Xb, x.scratch, kb, t, x.bad);
    aux_trtri32_finish(x.Ld, x.Xb, x.scratch, t);
    if (t == 0 && *x.bad != 0 && *bad_total == 0) *bad_total = pivot_offset + SB * j + *x.bad;
    st_buf_to_diag_n(x, j, x.Ld, kb, t, AUXT);
    for (int idx = t; idx < kb * 32; idx += AUXT) {
        const int r = idx >> 5, c = idx & 31;
        if (c <= r) xdiag[(size_t)(SB * j + r) * lda + SB * j + c] = x.Xb[r * FLDP + c];
    }
}

// In-place Cholesky of the N x N lower triangle in T (right-looking, block columns of 32).  Per step: panel, trailing
// update on all eight warps, then the next diagonal block (factor + inverse, a latency-bound chain of 32 dependent pivots,
// ~18 k cycles) on warps 0-2.  Whole CTA; CTA barriers between the phases.
__device__ __noinline__ void st_chol(const StructCtx& x, int tid, int* bad_total, int pivot_offset, double* xdiag, int lda) {
    const int lane = tid & 31, warp = tid >> 5;
    const StLane L = st_lane_offsets(lane);
    const int N = x.N, nb = x.nb;
    if (warp < 3) st_diag_step(x, 0, tid, xdiag, lda, bad_total, pivot_offset);
    __syncthreads();
    ST_MARK(1);
    for (int j = 0; j + 1 < nb; ++j) {
        const int m = N - SB * (j + 1);                    // rows below the (full) diagonal block j
        // ---- panel: rows below <- rows * L_jj^-T, 32-row chunks over the warps (a chunk is read and written by one warp)
        const int nch = (m + SB - 1) / SB;
        for (int ch = warp; ch < nch; ch += SWARPS) {
            double* blk = st_row(x, j, SB * (ch + 1));
            const int avail = st_rows(x, j) - SB * (ch + 1);
            double acc[4][4][2];
            st_blk_load(acc, blk, avail, lane);
            chunk_mul_regs<true>(acc, x.Xb, 1.0, lane);
            st_blk_store(acc, blk, avail, lane);
        }
        __syncthreads();
        ST_MARK(2);
        auto update = [&](int bi, int bj) {                // C(bi, bj) -= P(bi) P(bj)^T
            double* cblk = st_row(x, bj, SB * (bi - bj));
            const int availC = st_rows(x, bj) - SB * (bi - bj);
            double acc[4][4][2];
            st_blk_load(acc, cblk, availC, lane);
            st_blk_mma<false, true>(acc, st_row(x, j, SB * (bi - j)), st_rows(x, j) - SB * (bi - j),
                                    st_row(x, j, SB * (bj - j)), st_rows(x, j) - SB * (bj - j), -1.0, lane, L);
            if (bi == bj) {                                // keep the strict upper part of a diagonal block at zero
                const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                        for (int e = 0; e < 2; ++e) if (8 * jj + 2 * t4 + e > 8 * i + g) acc[i][jj][e] = 0.0;
            }
            st_blk_store(acc, cblk, availC, lane);
        };
        // ---- block column j+1 first (<= 6 blocks, one round)
        for (int bi = j + 1 + warp; bi < nb; bi += SWARPS) update(bi, j + 1);
        __syncthreads();
        ST_MARK(3);
        // ---- the remaining trailing blocks (bj >= j+2) on all warps, THEN the diagonal block j+1 on warps 0-2.  (Running
        //      the two concurrently was measured slower: the pivot chain of the diagonal block is a sequence of dependent
        //      scalar FP64 operations, and with DMMAs in flight on the same pipe each of them takes ~75 instead of ~40
        //      cycles - 31 k instead of 18 k cycles per block, more than the overlap saves.)
        {
            int t = 0;
            for (int bi = j + 2; bi < nb; ++bi)
                for (int bj = j + 2; bj <= bi; ++bj, ++t)
                    if (t % SWARPS == warp) update(bi, bj);
        }
        __syncthreads();
        ST_MARK(3);
        if (warp < 3) st_diag_step(x, j + 1, tid, xdiag, lda, bad_total, pivot_offset);
        __syncthreads();
        ST_MARK(1);
    }
}

// In-place inverse of the lower-triangular factor in T, block columns right to left:
//   X(j,j) = L(j,j)^-1 (read back from xdiag, where st_chol left it);  X(below, j) = -X(below, below) * (L(below, j) * X(j,j)).
// The triangular product gives chunk ch of the block column ch+1 block products: the warps without a chunk of their own
// take the first half of the longest ones and hand their partial sums over through the (now free) diagonal-block buffers.
__device__ __noinline__ void st_trtri(const StructCtx& x, int tid, const double* xdiag, int lda) {
    const int lane = tid & 31, warp = tid >> 5;
    const StLane L = st_lane_offsets(lane);
    const int N = x.N, nb = x.nb;
    double* pbuf[3] = {x.Pb, x.Ld, x.Xb};
    for (int j = nb - 1; j >= 0; --j) {
        const int kb = min(SB, N - SB * j);
        const int m = N - SB * j - kb;
        for (int idx = tid; idx < 32 * 32; idx += STHREADS) {
            const int r = idx >> 5, c = idx & 31;
            x.Xb[r * FLDP + c] = (r < kb && c <= r) ? xdiag[(size_t)(SB * j + r) * lda + SB * j + c] : 0.0;
        }
        __syncthreads();
        ST_MARK(4);
        st_buf_to_diag(x, j, x.Xb, kb, tid);
        const int nch = (m + SB - 1) / SB;                 // <= 6 chunks
        // Tm = L(below, j) * X(j,j), in place (a chunk belongs to one warp)
        for (int ch = warp; ch < nch; ch += SWARPS) {
            double* blk = st_row(x, j, SB * (ch + 1));
            const int avail = st_rows(x, j) - SB * (ch + 1);
            double acc[4][4][2];
            st_blk_load(acc, blk, avail, lane);
            chunk_mul_regs<false>(acc, x.Xb, 1.0, lane);
            st_blk_store(acc, blk, avail, lane);
        }
        __syncthreads();
        ST_MARK(5);
        // out(ch) = -sum_{k <= ch} X22(ch, k) Tm(k).  Owner of chunk ch: warp ch.  Helper h = warp - nch (h < 3) takes
        // k = 0 .. half-1 of chunk nch-1-h, the owner the rest; all chunks are formed in registers before any is stored.
        double out[4][4][2];
        const int nhelp = min(3, min(SWARPS - nch, nch));
        int ch = -1, k0 = 0, k1 = 0, hslot = -1;
        if (warp < nch) {
            ch = warp; k1 = ch + 1;
            const int h = nch - 1 - ch;                    // helper index of this chunk, if it has one
            if (h < nhelp) { k0 = (ch + 1) / 2; hslot = h; }
        } else if (warp - nch < nhelp) {
            hslot = warp - nch; ch = nch - 1 - hslot; k0 = 0; k1 = (ch + 1) / 2;
        }
        const bool owner = warp < nch, helper = !owner && ch >= 0;
        if (ch >= 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) { out[i][jj][0] = 0.0; out[i][jj][1] = 0.0; }
            for (int k = k0; k < k1; ++k) {
                const int bc = j + 1 + k;                  // block column of X22(ch, k)
                st_blk_mma<false, false>(out, st_row(x, bc, SB * (ch - k)), st_rows(x, bc) - SB * (ch - k),
                                         st_row(x, j, SB * (k + 1)), st_rows(x, j) - SB * (k + 1), -1.0, lane, L);
            }
        }
        __syncthreads();                                   // every product has read Tm and X(j,j) (Xb doubles as a hand-over buffer)
        if (helper && k1 > k0) {
            const int g = lane >> 2, t4 = lane & 3;
            double* hb = pbuf[hslot];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    *reinterpret_cast<double2*>(hb + (8 * i + g) * FLDP + 8 * jj + 2 * t4) = make_double2(out[i][jj][0], out[i][jj][1]);
        }
        __syncthreads();
        if (owner) {
            if (hslot >= 0 && k0 > 0) {
                const int g = lane >> 2, t4 = lane & 3;
                const double* hb = pbuf[hslot];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const double2 v = *reinterpret_cast<const double2*>(hb + (8 * i + g) * FLDP + 8 * jj + 2 * t4);
                        out[i][jj][0] += v.x; out[i][jj][1] += v.y;
                    }
            }
            st_blk_store(out, st_row(x, j, SB * (ch + 1)), st_rows(x, j) - SB * (ch + 1), lane);
        }
        __syncthreads();
        ST_MARK(8);
    }
}

// lower triangle of K + lambda * D (D = dvec on the diagonal) into T, zeros everywhere else.  The exponentials run as eight
// interleaved chains per thread (a dependent FP64 operation costs ~40 cycles here).
__device__ __noinline__ void st_build(const StructCtx& x, const double* dvec, double lambda, int tid) {
    const int N = x.N, I = x.I;
    for (int c = 0; c < x.nb; ++c) {
        double* blk = st_row(x, c, 0);
        const int total = st_rows(x, c) * SB;
        for (int idx0 = tid; idx0 < total; idx0 += 8 * STHREADS) {
            double e[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = idx0 + u * STHREADS;
                const int rr = min(idx, total - 1) >> 5, cc = idx & 31;
                const int i = min(SB * c + rr, N - 1), k = min(SB * c + cc, N - 1);
                e[u] = rbf_tab(x.s_sh + (size_t)i * I, x.s_sh + (size_t)k * I, I, x.sigma_f, x.etab);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = idx0 + u * STHREADS;
                if (idx >= total) continue;
                const int rr = idx >> 5, cc = idx & 31;
                const int i = SB * c + rr, k = SB * c + cc;
                double v = 0.0;
                if (i < N && k <= i) { v = e[u]; if (i == k) v += lambda * dvec[i]; }
                blk[rr * SB + st_swz(rr, cc)] = v;
            }
        }
    }
}

// lower triangle of T (rows < N) -> global block at dst (leading dimension lda): dst[i][k] = T(i, k), k <= i
__device__ __noinline__ void st_write_lower(const StructCtx& x, double* dst, int lda, int tid) {
    const int N = x.N;
    for (int c = 0; c < x.nb; ++c) {
        const double* blk = st_row(x, c, 0);
        const int rows = N - SB * c;
        for (int idx = tid; idx < rows * 16; idx += STHREADS) {
            const int rr = idx >> 4, q = idx & 15;
            const int i = SB * c + rr, k = SB * c + 2 * q;
            if (k > i) continue;
            const double2 v = *reinterpret_cast<const double2*>(blk + rr * SB + ((q ^ ((rr & 3) << 1)) << 1));
            if (k + 1 <= i) *reinterpret_cast<double2*>(dst + (size_t)i * lda + k) = v;
            else dst[(size_t)i * lda + k] = v.x;
        }
    }
}

__global__ void __launch_bounds__(STHREADS, 1) k_factor_struct(StructArgs p) {
    extern __shared__ __align__(16) double st_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const StLane L = st_lane_offsets(lane);
    StructCtx x;
    x.Rs = p.Rs; x.I = p.I; x.sigma_f = p.sigma_f;
    const int nb_max = (p.Rs + SB - 1) / SB;
    x.T = st_smem;
    x.Pb = x.T + (size_t)st_rows_before(nb_max, p.Rs) * SB;
    x.Ld = x.Pb + 32 * FLDP; x.Xb = x.Ld + 32 * FLDP; x.scratch = x.Xb + 32 * FLDP;
    x.s_sh = x.scratch + 256;
    x.d10 = x.s_sh + (size_t)p.Rs * p.I; x.daa[0] = x.d10 + p.Rs; x.daa[1] = x.daa[0] + p.Rs;
    x.etab = x.daa[1] + p.Rs;
    x.bad = reinterpret_cast<int*>(x.etab + kExpTabN);
    int* bad_total = x.bad + 1;
    __shared__ long long t_phase_sh;
    x.t_phase = &t_phase_sh;
    if (tid == 0) t_phase_sh = clock64();
    fill_exp2_table(x.etab, tid, STHREADS);
    const int O = p.O;
    for (int b = blockIdx.x; b < p.batch; b += gridDim.x) {
        const int N = p.n_pts ? p.n_pts[b] : p.n_max;
        const int Ns = comp_stride(N);
        x.N = N; x.nb = (N + SB - 1) / SB;
        const double* s = p.s + (size_t)b * p.n_max * p.I;
        const double* sg = p.sigma + (size_t)b * p.n_max * O * O;
        double* F = p.F + (size_t)b * p.strideF;
        __syncthreads();                                   // the previous problem is completely written out
        for (int idx = tid; idx < N * p.I; idx += STHREADS) x.s_sh[idx] = s[idx];
        for (int i = tid; i < N; i += STHREADS) {
            x.daa[0][i] = sg[(size_t)i * O * O];
            if (O == 2) { x.d10[i] = p.lambda * sg[(size_t)i * 4 + 2]; x.daa[1][i] = sg[(size_t)i * 4 + 3]; }
        }
        if (tid == 0) { *x.bad = 0; *bad_total = 0; }
        __syncthreads();
        // ---- component 0: L11, X11
        ST_MARK(9);
        st_build(x, x.daa[0], p.lambda, tid);
        __syncthreads();
        ST_MARK(0);
        st_chol(x, tid, bad_total, 0, F, p.lda);
        st_trtri(x, tid, F, p.lda);
        st_write_lower(x, F, p.lda, tid);
        ST_MARK(9);
        if (Ns > N) {                                      // identity padding row of component 0
            for (int k = tid; k <= N; k += STHREADS) F[(size_t)N * p.lda + k] = (k == N) ? 1.0 : 0.0;
        }
        if (O == 2) {
            double* Wg = F + Ns;                            // scratch: the (unused) upper-right block, W[i][k] at F[i][Ns + k]
            double* X21 = F + (size_t)Ns * p.lda;           // bottom-left block
            __syncthreads();
            // ---- G = X11^T X11 block row by block row; W = D10' G to global (both triangles), S = A11 - W D10' in place
            for (int bi = 0; bi < x.nb; ++bi) {
                const int bj = warp;
                double acc[4][4][2];
                if (bj <= bi) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) { acc[i][jj][0] = 0.0; acc[i][jj][1] = 0.0; }
                    for (int k = bi; k < x.nb; ++k)
                        st_blk_mma<true, false>(acc, st_row(x, bi, SB * (k - bi)), st_rows(x, bi) - SB * (k - bi),
                                                st_row(x, bj, SB * (k - bj)), st_rows(x, bj) - SB * (k - bj), 1.0, lane, L);
                }
                __syncthreads();                           // every product of this block row has read T before S overwrites it
                if (bj <= bi) {
                    double sblk[4][4][2];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = SB * bi + 8 * i + g;
                        const double di = r < N ? x.d10[r] : 0.0;
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            double kv[2];
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int c = SB * bj + 8 * jj + 2 * t4 + e;
                                kv[e] = rbf_tab(x.s_sh + (size_t)min(r, N - 1) * p.I, x.s_sh + (size_t)min(c, N - 1) * p.I, p.I, p.sigma_f, x.etab);
                            }
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int c = SB * bj + 8 * jj + 2 * t4 + e;
                                const double gv = acc[i][jj][e];
                                const bool ok = r < N && c < N;
                                // W is symmetric up to the row scaling: both positions come from ONE computed value (the
                                // lower one of a diagonal block), so the result does not depend on which lane stores last
                                if (ok && (bj < bi || c <= r)) {
                                    Wg[(size_t)r * p.lda + c] = di * gv;
                                    if (c != r) Wg[(size_t)c * p.lda + r] = x.d10[c] * gv;
                                }
                                double sv = 0.0;
                                if (ok && c <= r) {
                                    sv = kv[e] - di * gv * x.d10[c];
                                    if (c == r) sv += p.lambda * x.daa[1][r];
                                }
                                sblk[i][jj][e] = sv;
                            }
                        }
                    }
                    st_blk_store(sblk, st_row(x, bj, SB * (bi - bj)), st_rows(x, bj) - SB * (bi - bj), lane);
                }
                __syncthreads();
            }
            ST_MARK(10);
            // ---- component 1: L22, X22
            st_chol(x, tid, bad_total, Ns, F + (size_t)Ns * p.lda + Ns, p.lda);
            st_trtri(x, tid, F + (size_t)Ns * p.lda + Ns, p.lda);
            st_write_lower(x, F + (size_t)Ns * p.lda + Ns, p.lda, tid);
            if (Ns > N) {
                for (int k = tid; k <= Ns + N; k += STHREADS) F[(size_t)(Ns + N) * p.lda + k] = (k == Ns + N) ? 1.0 : 0.0;
            }
            __threadfence_block();
            __syncthreads();                               // W (written by this CTA) is visible to all its threads
            ST_MARK(9);
            // ---- X21 = -X22 W: 32 x 32 output blocks over the warps; X22 from T, W from global (L1/L2)
            const int nblk = x.nb * x.nb;
            for (int t = warp; t < nblk; t += SWARPS) {
                const int bi = t / x.nb, bj = t - bi * x.nb;
                double acc[4][4][2];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) { acc[i][jj][0] = 0.0; acc[i][jj][1] = 0.0; }
                for (int k = 0; k <= bi; ++k) {
                    const double* ablk = st_row(x, k, SB * (bi - k));
                    const int availA = st_rows(x, k) - SB * (bi - k);
                    const double* wblk = Wg + (size_t)(SB * k) * p.lda + SB * bj;
#pragma unroll
                    for (int s8 = 0; s8 < 8; ++s8) {
                        double a[4], bfr[4];
                        const int wr = SB * k + 4 * s8 + t4;
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int wc = SB * bj + 8 * jj + g;
                            bfr[jj] = (wr < N && wc < N) ? wblk[(size_t)(4 * s8 + t4) * p.lda + 8 * jj + g] : 0.0;
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) a[i] = availA >= SB ? -ablk[L.cs[s8] + 8 * SB * i] : -st_ldN(ablk, availA, i, s8, g, t4);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) dmma884(acc[i][jj], a[i], bfr[jj]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = SB * bi + 8 * i + g;
                    if (r >= N) continue;
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int c = SB * bj + 8 * jj + 2 * t4;
                        if (c + 1 < N) *reinterpret_cast<double2*>(X21 + (size_t)r * p.lda + c) = make_double2(acc[i][jj][0], acc[i][jj][1]);
                        else if (c < N) X21[(size_t)r * p.lda + c] = acc[i][jj][0];
                    }
                }
            }
            if (Ns > N) {                                  // padding column N of the bottom-left block, padding row of it
                for (int r = tid; r < N; r += STHREADS) X21[(size_t)r * p.lda + N] = 0.0;
                for (int k = tid; k < Ns; k += STHREADS) X21[(size_t)N * p.lda + k] = 0.0;
            }
        }
        __syncthreads();
        ST_MARK(11);
        if (tid == 0 && p.info) p.info[b] = *bad_total;
    }
}

inline bool factor_struct_applicable(int n_max_pts, int I, int O, int lda, long long strideF, const double* F) {
    if (g_host_dbg & 256) return false;
    if (O < 1 || O > 2 || I < 1 || n_max_pts < 1) return false;
    if ((lda & 1) || (strideF & 1) || (reinterpret_cast<uintptr_t>(F) & 15)) return false;
    const int Rs = comp_stride(n_max_pts);
    return struct_smem_bytes(Rs, I) <= 227 * 1024;
}

inline int launch_factor_struct(StructArgs p, cudaStream_t st) {
    p.Rs = comp_stride(p.n_max);
    const size_t smem = struct_smem_bytes(p.Rs, p.I);
    static SmemCache cache;
    KMPX_TRY(ensure_dyn_smem(k_factor_struct, smem, cache, "k_factor_struct"));
    k_factor_struct<<<p.batch < num_sms() ? p.batch : num_sms(), STHREADS, smem, st>>>(p);
    return check_launch("k_factor_struct");
}

}  // namespace kmpx
