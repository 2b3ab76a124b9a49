// Warp- and CTA-level building blocks around the FP64 tensor-core MMA (DMMA.8x8x4 on sm_100a).
//
// Shared-memory operand tiles:
//   A tile  [rows][k]   row-major, leading dimension LDA (doubles)
//   B tile  either  [n][k] ("n-major", k contiguous)  or  [k][n] ("k-major")
// Bank-conflict-free fragment reads need LD = 4 or 12 (mod 16) doubles: a half-warp reads
// (row or n = 0..3) x (k = 0..3) and 4*row + k then covers 16 distinct 8-byte banks.
#pragma once
#include "common.cuh"

namespace kmpx {

// acc[MI][NI][2] += A(8*MI x 4) * B(4 x 8*NI) for one k-step of 4.
//   As points at element (first row of the warp tile, k0);  Bs at (first n of the warp tile, k0).
template <int MI, int NI, bool B_KMAJOR>
__device__ __forceinline__ void warp_mma_k4(double (&acc)[MI][NI][2], const double* __restrict__ As, int lda,
                                            const double* __restrict__ Bs, int ldb, int lane) {
    const int g = lane >> 2, t = lane & 3;
    double a[MI], b[NI];
#pragma unroll
    for (int i = 0; i < MI; ++i) a[i] = As[(i * 8 + g) * lda + t];
#pragma unroll
    for (int j = 0; j < NI; ++j) b[j] = B_KMAJOR ? Bs[t * ldb + j * 8 + g] : Bs[(j * 8 + g) * ldb + t];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma884(acc[i][j], a[i], b[j]);
}

template <int MI, int NI>
__device__ __forceinline__ void zero_acc(double (&acc)[MI][NI][2]) {
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
}

// Asynchronously copies a (rows x BK) tile of a row-major global matrix into shared memory [rows][LD],
// 16 bytes (two columns) per cp.async, zero-filling what lies outside the valid region:
//   rows  >= rows_valid                      -> zeros
//   cols  >= col_end (global column index)    -> zeros
//   TRI:  global column > global row          -> zeros   (lower-triangular operand stored in a full array)
// g points at global element (row_g0, col_g0); both lda and col_g0 must be even and g 16-byte aligned.
template <int BK, int LD, bool TRI>
__device__ __forceinline__ void load_tile_async(double* __restrict__ dst, const double* __restrict__ g, int lda,
                                                int rows, int rows_valid, int row_g0, int col_g0, int col_end,
                                                int tid, int nthreads) {
    constexpr int CH = BK / 2;   // 16-byte chunks per row
    const int total = rows * CH;
    for (int idx = tid; idx < total; idx += nthreads) {
        const int r = idx / CH, q = idx - r * CH;
        const int gc = col_g0 + 2 * q;
        int nv = 0;
        if (r < rows_valid) {
            nv = col_end - gc;
            if (TRI) { const int lim = row_g0 + r - gc + 1; nv = nv < lim ? nv : lim; }
            nv = nv < 0 ? 0 : (nv > 2 ? 2 : nv);
        }
        const double* src = nv > 0 ? g + (size_t)r * lda + 2 * q : g;
        cp_async16(dst + r * LD + 2 * q, src, nv * 8);
    }
}

}  // namespace kmpx
