// Warp-per-problem variants of the small-system factor / inverse kernels (textually included by factor_small.cu).
//
// k_potrf_tma / k_trtri_tma give one problem to a CTA of twelve warps with three roles; two thirds of their time is the
// per-block-column skeleton (panel loads, the 32 x 32 solve, stores, fence, barrier, the diagonal-block chain) that all
// MMA warps walk in lock-step, so the FP64 tensor pipe idles 60 % of the time.  Here a WARP owns a
// problem from the first block column to the last and shares nothing with the other warps of its CTA: no CTA barrier,
// no hand-over between roles, no progress counters.  The warps of an SM run independent problems that drift apart, so
// that the latency-bound pieces of one (in-warp Cholesky / inverse of a 32 x 32 diagonal block, global loads of a
// panel, stores + fence) are covered by the contractions of the others.
//
//   * operands of the block-column contraction arrive through the TMA unit in a PRIVATE ring per warp: a stage holds
//     the k-tile (8 columns, 64-byte swizzle, the layout of factor_tma.cuh) of the 32 B rows and of U units of 32 A
//     rows (U = 2: the B rows are streamed once per 64 rows); the warp's own lane 0 issues the boxes - the tile
//     sequence of a problem is a pure function of its size - as far ahead as the ring allows and as the columns it
//     reads are final (written AND fenced by this warp);
//   * potrf, block column j: the diagonal block's history is a self product of its 32 rows; it is factored and
//     inverted inside the warp (lane = row for the Cholesky recurrence, lane = column for the inverse) on a 33-stride
//     block that borrows the (drained) ring; the inverse then lives in DMMA fragment order in 5 KB of shared memory;
//     every 32-row unit below is  X = (A - L[unit, :k0] L[k0:k0+32, :k0]^T) Dinv^T;
//   * trtri, block column jb (right to left):  X21 = -(X22 L21) Dinv  with the row units taken bottom-up: unit u needs the
//     rows k < 32 (u + 1) of L21, so what it overwrites is never read again in the step; L21 ITSELF is the k-major B
//     operand (boxes of 8 rows x 8 columns, whose 64-byte swizzle makes the k-major fragment reads conflict-free), the
//     accumulator is multiplied by Dinv from registers like potrf's panel - no T = L21 Dinv, no scratch, no transposition;
//   * struct_o == 2 declares the component-major KMP system of two outputs (A = [[A00, D], [D, A11]] with D diagonal,
//     component stride n / 2): L21 = D L11^-T is upper triangular, so k-tiles and whole units that are exactly
//     zero are skipped (the results are bit-identical to the unskipped run: only products with exact zeros are dropped).
//
// Per-problem HBM traffic is higher than in the CTA-per-problem kernels (> 1000 problems in flight do not fit the L2,
// and the B rows are streamed once per pass): 3 - 6 MB instead of ~1.4 MB at n = 404 - the price of hiding the latency.
#pragma once
#include "tma_util.cuh"

namespace kmpx {

constexpr int WPP_BOX_ELEMS = 32 * 8;              // one TMA box: 32 rows x 8 columns
constexpr int WPP_BOX_BYTES = WPP_BOX_ELEMS * 8;
constexpr int WLD = 33;                            // leading dimension of the transient diagonal block
constexpr int WPP_NFRAG = 20;                      // non-zero (n-tile, k-step) fragments of a triangular 32 x 32 B operand
constexpr int WPP_DFRAG_ELEMS = WPP_NFRAG * 32;

// Variants V: 0 = one unit per pass, 12 warps;  1 = two units per pass, 8 warps (255 registers);  2 = two units, 12 warps
// (168 registers), two stages.  (ptxas sizes the register file of sm_100a for CTAs in multiples of four warps.)
template <int V> struct WppCfg {
    static constexpr int U = V == 0 ? 1 : 2;                       // 32-row units per pass (the B rows are streamed once per pass)
    static constexpr int WARPS = V == 1 ? 8 : 12;                  // problems in flight per SM
    static constexpr int THREADS = WARPS * 32;
    static constexpr int S = V == 2 ? 2 : 3;                       // ring stages per warp
    static constexpr int STAGE_ELEMS = (1 + U) * WPP_BOX_ELEMS;    // [B rows | A rows of unit 0 | A rows of unit 1]
    static constexpr size_t RING_BYTES = (size_t)WARPS * S * STAGE_ELEMS * 8;
    static constexpr size_t DFRAG_BYTES = (size_t)WARPS * WPP_DFRAG_ELEMS * 8;
    static constexpr size_t SMEM = RING_BYTES + DFRAG_BYTES + (size_t)WARPS * S * 8 + 1024;
    static_assert(S * STAGE_ELEMS >= 32 * WLD, "the transient diagonal block borrows the ring");
    static_assert(SMEM <= 227 * 1024, "shared memory");
};

// acc[c] += (A rows of unit c) * (B rows)^T over the 8 columns of one stage (two k-steps).  SELF: unit 0 IS the B rows.
template <int U, bool SELF>
__device__ __forceinline__ void wpp_tile_mma(double (&acc)[U][4][4][2], const double* tile, int o0, int o1, bool live0, bool live1) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int o = s ? o1 : o0;
        double bf[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) bf[j] = tile[j * 64 + o];
#pragma unroll
        for (int c = 0; c < (SELF ? 1 : U); ++c) {
            if (c == 0 ? live0 : live1) {
                double a[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = SELF ? bf[i] : tile[(1 + c) * WPP_BOX_ELEMS + i * 64 + o];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[c][i][j], a[i], bf[j]);
            }
        }
    }
}

// index of the (n-tile j, k-step q) fragment in the fragment-order copy of a triangular block, -1 when it is zero
template <bool D_TRANS>
__host__ __device__ constexpr int wpp_frag_index(int j, int q) {
    int idx = 0;
    for (int jj = 0; jj < 4; ++jj)
        for (int qq = 0; qq < 8; ++qq) {
            const bool live = D_TRANS ? (4 * qq <= 8 * jj + 7) : (4 * qq + 3 >= 8 * jj);
            if (jj == j && qq == q) return live ? idx : -1;
            if (live) ++idx;
        }
    return -1;
}
// D (32 x 32, leading dimension WLD, lower triangular) -> fragment order: F[idx * 32 + lane] is the B fragment element
// of lane (g, t4) for  B = D^T (D_TRANS: B[k][n] = D[n][k])  or  B = D.
template <bool D_TRANS>
__device__ __forceinline__ void wpp_extract_frags(const double* D, double* F, int lane) {
    const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            constexpr bool T = D_TRANS;
            const int idx = wpp_frag_index<T>(j, q);
            if (idx >= 0) F[idx * 32 + lane] = T ? D[(j * 8 + g) * WLD + 4 * q + t4] : D[(4 * q + t4) * WLD + j * 8 + g];
        }
}
// P <- P * op(D) for a 32 x 32 chunk in accumulator layout, op(D) given in fragment order.
template <bool D_TRANS>
__device__ __forceinline__ void wpp_chunk_mul(double (&P)[4][4][2], const double* F, int lane) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double a[8];
        a[0] = c_to_a_frag<0>(P[i], lane); a[1] = c_to_a_frag<1>(P[i], lane);
        a[2] = c_to_a_frag<2>(P[i], lane); a[3] = c_to_a_frag<3>(P[i], lane);
        a[4] = c_to_a_frag<4>(P[i], lane); a[5] = c_to_a_frag<5>(P[i], lane);
        a[6] = c_to_a_frag<6>(P[i], lane); a[7] = c_to_a_frag<7>(P[i], lane);
        double out[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j) { out[j][0] = 0.0; out[j][1] = 0.0; }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                constexpr bool T = D_TRANS;
                const int idx = wpp_frag_index<T>(j, q);
                if (idx >= 0) dmma884(out[j], a[q], F[idx * 32 + lane]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { P[i][j][0] = out[j][0]; P[i][j][1] = out[j][1]; }
    }
}

// In-warp Cholesky factor of the 32 x 32 block D (lower triangle; rows / columns >= kb are identity), in place.
// Lane i owns row i: column j of the factor is  L[i][j] = (D[i][j] - sum_{p<j} L[i][p] L[j][p]) / L[j][j]  for all rows at
// once (row j is a broadcast read, the own row a conflict-free one at stride 33).  Returns 1 / L[lane][lane];
// bad = 1-based index of the first non-positive pivot among the first kb (0: none).
__device__ __forceinline__ double wpp_chol32(double* D, int kb, int lane, int& bad) {
    double rinv = 1.0;
    bad = 0;
    const double* my = D + lane * WLD;
    for (int j = 0; j < 32; ++j) {
        const double* rj = D + j * WLD;
        double s0 = my[j], s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int p = 0;
        for (; p + 4 <= j; p += 4) {
            s0 = fma(-my[p], rj[p], s0);
            s1 = fma(-my[p + 1], rj[p + 1], s1);
            s2 = fma(-my[p + 2], rj[p + 2], s2);
            s3 = fma(-my[p + 3], rj[p + 3], s3);
        }
        for (; p < j; ++p) s0 = fma(-my[p], rj[p], s0);
        const double s = (s0 + s1) + (s2 + s3);
        const double d = __shfl_sync(0xffffffffu, s, j);
        if (!(d > 0.0) && bad == 0 && j < kb) bad = j + 1;
        const double rs = rsqrt_chain(d);
        if (lane == j) rinv = rs;
        if (lane >= j) D[lane * WLD + j] = (lane == j) ? d * rs : s * rs;
        __syncwarp();
    }
    return rinv;
}

// In-warp inverse of the lower-triangular 32 x 32 block D, in place (zeros above the diagonal on entry and on exit).
// Lane c owns column c:  X[i][c] = (delta_ic - sum_{p<i} L[i][p] X[p][c]) / L[i][i], rows top to bottom; row i of L is
// read (broadcast) before row i of X overwrites it.  rinv: 1 / L[lane][lane].
__device__ __forceinline__ void wpp_trinv32(double* D, double rinv, int lane) {
    for (int i = 0; i < 32; ++i) {
        const double ri = __shfl_sync(0xffffffffu, rinv, i);
        const double* Li = D + i * WLD;
        double s0 = (i == lane) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int p = 0;
        for (; p + 4 <= i; p += 4) {
            s0 = fma(-Li[p], D[p * WLD + lane], s0);
            s1 = fma(-Li[p + 1], D[(p + 1) * WLD + lane], s1);
            s2 = fma(-Li[p + 2], D[(p + 2) * WLD + lane], s2);
            s3 = fma(-Li[p + 3], D[(p + 3) * WLD + lane], s3);
        }
        for (; p < i; ++p) s0 = fma(-Li[p], D[p * WLD + lane], s0);
        const double x = ((s0 + s1) + (s2 + s3)) * ri;
        __syncwarp();
        D[i * WLD + lane] = x;
        __syncwarp();
    }
}

// The per-warp ring: `issued` / `consumed` count tiles since the kernel started; tile t lives in stage t % S and
// completes phase (t / S) & 1 of that stage's mbarrier.
template <int V> struct WppRing {
    using C = WppCfg<V>;
    double* stage0; uint32_t full0; uint32_t issued, consumed;
    __device__ __forceinline__ const double* wait_tile() const {
        mbar_wait(full0 + 8 * (consumed % C::S), (consumed / C::S) & 1);
        return stage0 + (size_t)(consumed % C::S) * C::STAGE_ELEMS;
    }
    __device__ __forceinline__ void release_tile() { __syncwarp(); ++consumed; }
    __device__ __forceinline__ bool has_room() const { return issued - consumed < (uint32_t)C::S; }
    __device__ __forceinline__ bool drained() const { return issued == consumed; }
    // lane 0: arm the barrier of the next stage for `nboxes` boxes; returns the shared-memory address of the stage
    __device__ __forceinline__ uint32_t arm(int nboxes, uint32_t& bar) const {
        bar = full0 + 8 * (issued % C::S);
        mbar_expect_tx(bar, (uint32_t)nboxes * WPP_BOX_BYTES);
        return smem_u32(stage0 + (size_t)(issued % C::S) * C::STAGE_ELEMS);
    }
};
template <int V>
__device__ __forceinline__ WppRing<V> wpp_ring_setup(unsigned char* smem_raw, int warp, int lane, double*& F) {
    using C = WppCfg<V>;
    unsigned char* base = smem_align1k(smem_raw);
    WppRing<V> r;
    r.stage0 = reinterpret_cast<double*>(base) + (size_t)warp * C::S * C::STAGE_ELEMS;
    F = reinterpret_cast<double*>(base + C::RING_BYTES) + (size_t)warp * WPP_DFRAG_ELEMS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + C::RING_BYTES + C::DFRAG_BYTES) + warp * C::S;
    r.full0 = smem_u32(bars);
    r.issued = 0; r.consumed = 0;
    if (lane == 0) {
        for (int s = 0; s < C::S; ++s) mbar_init(r.full0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async;\n" ::: "memory");
    }
    __syncwarp();
    return r;
}
// the transient 32 x 33 block may be written by this warp only while no TMA box is in flight into the ring, and the
// next box may only be issued after the warp is done with it: generic-proxy accesses vs. async-proxy writes
__device__ __forceinline__ void wpp_fence_smem_to_tma() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// L2 prefetch of a 32-row x 32-column block (256 bytes per row: two or three 128-byte lines) that this warp will read
// from global memory after its next contraction
__device__ __forceinline__ void wpp_prefetch_block(const double* A, int lda, int row0, int col0, int n, int lane) {
    const int r = row0 + lane;
    if (r < n) {
        const char* q = reinterpret_cast<const char*>(A + (size_t)r * lda + col0);
        asm volatile("prefetch.global.L2 [%0];\n" :: "l"(q));
        asm volatile("prefetch.global.L2 [%0];\n" :: "l"(q + 128));
        asm volatile("prefetch.global.L2 [%0];\n" :: "l"(q + 248));
    }
}

// One pass of a contraction: the B rows, up to U units of A rows, the k-tiles [tlo, thi); unit 0 only takes part in the
// tiles below t0hi (trtri: its triangular k-range ends one unit earlier than that of unit 1).
struct WppPass { int brow, arow0, arow1, cbase, tlo, thi, t0hi; };

// first k-tile of the rows >= row0 that is not identically zero (struct_ns > 0: the entries L[r][k], r >= struct_ns,
// k < r - struct_ns vanish), clamped to the tile count of the contraction
__device__ __forceinline__ int wpp_first_tile(int row0, int struct_ns, int ntile) {
    const int z = struct_ns > 0 ? row0 - struct_ns : 0;
    return z > 0 ? min(z >> 3, ntile) : 0;
}
// potrf, block column j: pass q = -1 is the diagonal block (self product), q >= 0 the units U q .. U q + U - 1 below it.
// Returns false when the pass does not exist (no such units, or - struct_ns - all of its units are exactly zero).
template <int U>
__device__ __forceinline__ bool wpp_potrf_pass(int j, int q, int n, int struct_ns, WppPass& ps) {
    const int k0 = NB * j, ntile = 4 * j;
    ps.brow = k0; ps.cbase = 0; ps.thi = ntile; ps.t0hi = ntile; ps.arow1 = -1;
    if (q < 0) { ps.arow0 = -1; ps.tlo = wpp_first_tile(k0, struct_ns, ntile); return true; }
    const int row0 = k0 + NB + U * NB * q;
    if (row0 >= n) return false;
    if (struct_ns > 0 && row0 - struct_ns >= k0 + NB) return false;          // this unit and all below it are exactly zero
    ps.arow0 = row0;
    ps.tlo = wpp_first_tile(row0, struct_ns, ntile);
    if (U == 2) {
        const int row1 = row0 + NB;
        if (row1 < n && !(struct_ns > 0 && row1 - struct_ns >= k0 + NB)) ps.arow1 = row1;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------------------
// Left-looking blocked Cholesky, one warp per problem.
template <int V>
__global__ void __launch_bounds__(WppCfg<V>::THREADS, 1) k_potrf_wpp(const __grid_constant__ CUtensorMap tmap, FactorArgs p, int struct_o) {
    using C = WppCfg<V>;
    constexpr int U = C::U;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    double* F;
    WppRing<V> ring = wpp_ring_setup<V>(smem_raw, warp, lane, F);
    double* D = ring.stage0;                      // transient diagonal block (ring drained)
    int o0, o1;
    swz_offsets(lane, o0, o1);

    for (int b = warp * gridDim.x + blockIdx.x; b < p.batch; b += gridDim.x * C::WARPS) {
        const int n = p.n_sys ? p.n_sys[b] : p.n_max;
        const int struct_ns = struct_o == 2 ? n >> 1 : 0;      // component stride of this problem
        double* A = p.A + (size_t)b * p.strideA;
        const int nblk = (n + NB - 1) / NB;
        int bad_total = 0;
        // ---- producer cursor: pass (pj, pq) and k-tile pt; `gate`: the last pass the consumer allows (the diagonal block
        //      borrows the ring, so the tiles of the units below are held back until it is done)
        int pj = 1, pq = -1, pt = 0, cols_final = 0, gate = 0;
        WppPass pp;
        bool pvalid = nblk > 1 && wpp_potrf_pass<U>(1, -1, n, struct_ns, pp);
        if (pvalid) pt = pp.tlo;
        auto seek = [&]() {      // normalise the cursor to the next existing tile (or pvalid = false)
            while (pvalid && pt >= pp.thi) {
                ++pq;
                if (!wpp_potrf_pass<U>(pj, pq, n, struct_ns, pp)) {
                    ++pj; pq = -1;
                    if (pj >= nblk) { pvalid = false; break; }
                    wpp_potrf_pass<U>(pj, -1, n, struct_ns, pp);
                }
                pt = pp.tlo;
            }
        };
        seek();
        auto pump = [&]() {
            while (pvalid && ring.has_room() && 8 * pt + 8 <= cols_final && 2 * pj + (pq >= 0) <= gate) {
                if (lane == 0) {
                    uint32_t bar;
                    const uint32_t dst = ring.arm(1 + (pp.arow0 >= 0) + (pp.arow1 >= 0), bar);
                    tma_load_box(dst, &tmap, 8 * pt, pp.brow, b, bar);
                    if (pp.arow0 >= 0) tma_load_box(dst + WPP_BOX_BYTES, &tmap, 8 * pt, pp.arow0, b, bar);
                    if (pp.arow1 >= 0) tma_load_box(dst + 2 * WPP_BOX_BYTES, &tmap, 8 * pt, pp.arow1, b, bar);
                }
                ++ring.issued;
                ++pt;
                seek();
            }
        };

        for (int j = 0; j < nblk; ++j) {
            const int k0 = j * NB, kb = min(NB, n - k0);
            double acc[U][4][4][2];
            WppPass ps;
            // ---- diagonal block: history (self product of its rows), factor, inverse
            gate = 2 * j;
            wpp_prefetch_block(A, p.lda, k0, k0, n, lane);
            zero_acc(acc[0]);
            wpp_potrf_pass<U>(j, -1, n, struct_ns, ps);
            for (int t = ps.tlo; t < ps.thi; ++t) {
                pump();
                const double* tile = ring.wait_tile();
                if (!(g_dbg & 1)) wpp_tile_mma<U, true>(acc, tile, o0, o1, true, false);
                ring.release_tile();
            }
            // (the ring is drained here: the gate holds back the tiles of the units)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int r = i * 8 + g, c = jj * 8 + 2 * t4 + e;
                        double v = (r == c) ? 1.0 : 0.0;
                        if (r < kb && c < kb) v = c <= r ? A[(size_t)(k0 + r) * p.lda + k0 + c] - acc[0][i][jj][e] : 0.0;
                        D[r * WLD + c] = v;
                    }
            __syncwarp();
            int bad;
            const double rinv = wpp_chol32(D, kb, lane, bad);
            if (bad != 0 && bad_total == 0) bad_total = k0 + bad;
            for (int r = 0; r < kb; ++r)
                if (lane <= r) A[(size_t)(k0 + r) * p.lda + k0 + lane] = D[r * WLD + lane];
            if (k0 + NB < n) {
                wpp_trinv32(D, rinv, lane);
                wpp_extract_frags<true>(D, F, lane);
            }
            wpp_fence_smem_to_tma();
            __syncwarp();
            gate = 2 * j + 1;
            // ---- the 32-row units below the diagonal block, U at a time
            for (int q = 0; wpp_potrf_pass<U>(j, q, n, struct_ns, ps); ++q) {
                const bool live1 = ps.arow1 >= 0;
                pump();
                wpp_prefetch_block(A, p.lda, ps.arow0, k0, n, lane);
                if (live1) wpp_prefetch_block(A, p.lda, ps.arow1, k0, n, lane);
#pragma unroll
                for (int c = 0; c < U; ++c) zero_acc(acc[c]);
                for (int t = ps.tlo; t < ps.thi; ++t) {
                    pump();
                    const double* tile = ring.wait_tile();
                    if (!(g_dbg & 1)) wpp_tile_mma<U, false>(acc, tile, o0, o1, true, live1);
                    ring.release_tile();
                }
                pump();
#pragma unroll
                for (int c = 0; c < U; ++c) {
                    if (c == 1 && !live1) break;
                    const int row0 = c ? ps.arow1 : ps.arow0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int gr = min(row0 + i * 8 + g, n - 1);
                        const bool ok = row0 + i * 8 + g < n;
                        double2 v[4];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) v[jj] = *reinterpret_cast<const double2*>(A + (size_t)gr * p.lda + k0 + jj * 8 + 2 * t4);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            acc[c][i][jj][0] = ok ? v[jj].x - acc[c][i][jj][0] : 0.0;
                            acc[c][i][jj][1] = ok ? v[jj].y - acc[c][i][jj][1] : 0.0;
                        }
                    }
                    wpp_chunk_mul<true>(acc[c], F, lane);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int gr = row0 + i * 8 + g;
                        if (gr < n) {
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj)
                                *reinterpret_cast<double2*>(A + (size_t)gr * p.lda + k0 + jj * 8 + 2 * t4) = make_double2(acc[c][i][jj][0], acc[c][i][jj][1]);
                        }
                    }
                }
            }
            // the block column is read back by the TMA unit from the next step on
            fence_global_to_tma();
            __syncwarp();
            cols_final = k0 + NB;
        }
        if (lane == 0 && p.info) {
            if (!p.info_accumulate) p.info[b] = bad_total;
            else if (bad_total != 0 && p.info[b] == 0) p.info[b] = p.info_offset + bad_total;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// In-place inverse of the lower factor, one warp per problem, block columns right to left.
// trtri, step with first trailing row r0 and m rows below the diagonal block: pass q = units U q .. of the trailing rows;
// unit u needs the k-tiles t < 4 (u + 1) (X22 is lower triangular), at most ktot (rows of T; struct_ns: non-zero rows of T)
template <int U>
__device__ __forceinline__ bool wpp_trtri_pass(int k0, int q, int n, int ktot, WppPass& ps) {
    const int r0 = k0 + NB;
    const int u0 = U * q;
    const int row0 = r0 + u0 * NB;
    if (row0 >= n) return false;
    ps.brow = k0; ps.cbase = r0; ps.tlo = 0; ps.arow0 = row0; ps.arow1 = -1;
    ps.t0hi = min(4 * (u0 + 1), ktot);
    ps.thi = ps.t0hi;
    if (U == 2 && row0 + NB < n) { ps.arow1 = row0 + NB; ps.thi = min(4 * (u0 + 2), ktot); }
    return true;
}

// Optional by-product of the inverse:  w = Linv y'  (what kmpx_solve_alpha computes in a pass of its own), accumulated block
// column by block column from the pieces of X while they are in registers / shared memory.  y: point-major targets
// [batch][n_max_pts][O]; the systems are component-major (row = a Ns + i, padding rows carry y' = 0).  y == nullptr: off.
struct WppSolve { const double* y; const int* n_pts; int n_max_pts, O; double* w; int ldw; };
__device__ __forceinline__ double wpp_rhs(const WppSolve& sv, const double* yb, int N, int Ns, int c) {
    const int a = c / Ns, i = c - a * Ns;
    return (a < sv.O && i < N) ? yb[(size_t)i * sv.O + a] : 0.0;
}

// acc[c] += (A rows of unit c) * B over the 8 k of one stage, B given K-MAJOR: four boxes of 8 k-rows x 8 columns (the rows
// r0 + 8 t .. of L21 itself, 64-byte swizzle).  Lane (g, t4) reads B[k][8 j + g] with k = {0,1,4,5}[t4] (step 0) /
// {2,3,6,7}[t4] (step 1) - the k order of the A fragments; the swizzle puts the four rows of a half-warp into distinct banks.
// kvalid < 8 (last tile of a ragged / structurally short k-range): fragments of k >= kvalid are zeroed in registers.
template <int U>
__device__ __forceinline__ void wpp_tile_mma_kmajor(double (&acc)[U][4][4][2], const double* tile, int o0, int o1, int ob0, int ob1,
                                                    int kk0, bool live0, bool live1, int kvalid) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int o = s ? o1 : o0, ob = s ? ob1 : ob0;
        const bool kok = kk0 + 2 * s < kvalid;
        double bf[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { const double v = tile[j * 64 + ob]; bf[j] = kok ? v : 0.0; }
#pragma unroll
        for (int c = 0; c < U; ++c) {
            if (c == 0 ? live0 : live1) {
                double a[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { const double v = tile[(1 + c) * WPP_BOX_ELEMS + i * 64 + o]; a[i] = kok ? v : 0.0; }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[c][i][j], a[i], bf[j]);
            }
        }
    }
}

// tmap: boxes of 32 rows x 8 columns (A rows = rows of X22);  tmapb: boxes of 8 rows x 8 columns (k-major B = L21 itself).
// X21 = -(X22 L21) Dinv, the row units taken BOTTOM-UP: unit u needs the rows k < 32 (u + 1) of L21 - its own and those above -
// so what it overwrites is never read again in this step, and neither T = L21 Dinv nor any scratch outside the lower
// triangle is needed.
template <int V>
__global__ void __launch_bounds__(WppCfg<V>::THREADS, 1) k_trtri_wpp(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmapb,
                                                                     FactorArgs p, int struct_o, WppSolve sv) {
    using C = WppCfg<V>;
    constexpr int U = C::U;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    double* F;
    WppRing<V> ring = wpp_ring_setup<V>(smem_raw, warp, lane, F);
    double* D = ring.stage0;
    int o0, o1;
    swz_offsets(lane, o0, o1);
    // k index of this lane inside a k-tile (k-step 0; + 2 for k-step 1) and its offsets in a k-major box
    const int kk0 = (t4 & 1) + 4 * (t4 >> 1);
    const int ob0 = kk0 * 8 + (((g >> 1) ^ ((kk0 >> 1) & 3)) << 1) + (g & 1);
    const int ob1 = (kk0 + 2) * 8 + (((g >> 1) ^ (((kk0 + 2) >> 1) & 3)) << 1) + (g & 1);

    for (int b = warp * gridDim.x + blockIdx.x; b < p.batch; b += gridDim.x * C::WARPS) {
        const int n = p.n_sys ? p.n_sys[b] : p.n_max;
        const int struct_ns = struct_o == 2 ? n >> 1 : 0;
        double* L = p.A + (size_t)b * p.strideA;
        const int nblk = (n + NB - 1) / NB;
        const bool solve = sv.y != nullptr;
        const double* yb = solve ? sv.y + (size_t)b * sv.n_max_pts * sv.O : nullptr;
        double* wb = solve ? sv.w + (size_t)b * sv.ldw : nullptr;
        const int svN = solve ? (sv.n_pts ? sv.n_pts[b] : sv.n_max_pts) : 0;
        const int svNs = solve ? n / sv.O : 1;          // component stride of this problem
        for (int jb = nblk - 1; jb >= 0; --jb) {
            const int k0 = jb * NB, kb = min(NB, n - k0);
            const int r0 = k0 + NB;
            const int m = n - k0 - kb;                       // rows below the diagonal block
            // ---- inverse of the diagonal block (the ring is drained), written back with an explicit zero upper triangle
            for (int r = 0; r < 32; ++r) {
                double v = (r == lane) ? 1.0 : 0.0;
                if (r < kb && lane < kb) v = lane <= r ? L[(size_t)(k0 + r) * p.lda + k0 + lane] : 0.0;
                D[r * WLD + lane] = v;
            }
            __syncwarp();
            wpp_trinv32(D, rcp_chain(D[lane * WLD + lane]), lane);
            for (int r = 0; r < kb; ++r)
                if (lane < kb) L[(size_t)(k0 + r) * p.lda + k0 + lane] = D[r * WLD + lane];
            // w[k0 + r] = X_jj[r][:] y'[k0 ..]: the first contribution to these rows (later steps add those of the block columns to the left)
            double ylane = 0.0;          // y'[k0 + lane]
            if (solve) {
                ylane = (k0 + lane < n) ? wpp_rhs(sv, yb, svN, svNs, k0 + lane) : 0.0;
                double dot = 0.0;
                for (int c = 0; c < 32; ++c) dot = fma(D[lane * WLD + c], __shfl_sync(0xffffffffu, ylane, c), dot);
                if (lane < kb) wb[k0 + lane] = dot;
            }
            if (m > 0) {
                wpp_extract_frags<false>(D, F, lane);
                wpp_fence_smem_to_tma();
                __syncwarp();
                // rows of L21 beyond the component stride vanish (struct_ns: L[r][k0 .. k0+31] = 0 for r - struct_ns >= k0 + 32)
                const int mt = struct_ns > 0 ? min(m, struct_ns) : m;
                const int ktot = (mt + 7) >> 3;
                const int np = ((m + NB - 1) / NB + U - 1) / U;          // passes of U units
                double acc[U][4][4][2];
                int pq = np - 1, pt = 0;
                WppPass pp;
                bool pvalid = wpp_trtri_pass<U>(k0, pq, n, ktot, pp);
                auto pump = [&]() {
                    while (pvalid && ring.has_room()) {
                        if (lane == 0) {
                            const bool a0 = pt < pp.t0hi, a1 = pp.arow1 >= 0;
                            uint32_t bar;
                            const uint32_t dst = ring.arm(1 + a0 + a1, bar);
#pragma unroll
                            for (int jn = 0; jn < 4; ++jn) tma_load_box(dst + jn * 512, &tmapb, k0 + 8 * jn, r0 + 8 * pt, b, bar);
                            if (a0) tma_load_box(dst + WPP_BOX_BYTES, &tmap, r0 + 8 * pt, pp.arow0, b, bar);
                            if (a1) tma_load_box(dst + 2 * WPP_BOX_BYTES, &tmap, r0 + 8 * pt, pp.arow1, b, bar);
                        }
                        ++ring.issued;
                        if (++pt >= pp.thi) { pt = 0; --pq; pvalid = pq >= 0 && wpp_trtri_pass<U>(k0, pq, n, ktot, pp); }
                    }
                };
                WppPass ps;
                for (int q = np - 1; q >= 0; --q) {
                    wpp_trtri_pass<U>(k0, q, n, ktot, ps);
                    const bool live1 = ps.arow1 >= 0;
#pragma unroll
                    for (int c = 0; c < U; ++c) zero_acc(acc[c]);
                    for (int t = 0; t < ps.thi; ++t) {
                        pump();
                        const double* tile = ring.wait_tile();
                        if (!(g_dbg & 1)) wpp_tile_mma_kmajor<U>(acc, tile, o0, o1, ob0, ob1, kk0, t < ps.t0hi, live1, mt - 8 * t);
                        ring.release_tile();
                    }
                    pump();
#pragma unroll
                    for (int c = 0; c < U; ++c) {
                        if (c == 1 && !live1) break;
                        const int row0 = c ? ps.arow1 : ps.arow0;
                        wpp_chunk_mul<false>(acc[c], F, lane);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int r = row0 + i * 8 + g;
                            if (r < n) {
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj)
                                    *reinterpret_cast<double2*>(L + (size_t)r * p.lda + k0 + jj * 8 + 2 * t4) = make_double2(-acc[c][i][jj][0], -acc[c][i][jj][1]);
                            }
                            if (solve) {          // w[r] += X21[r][:] y'[k0 ..] = -(acc row) . y'
                                double part = 0.0;
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj) {
                                    part = fma(acc[c][i][jj][0], __shfl_sync(0xffffffffu, ylane, jj * 8 + 2 * t4), part);
                                    part = fma(acc[c][i][jj][1], __shfl_sync(0xffffffffu, ylane, jj * 8 + 2 * t4 + 1), part);
                                }
                                part += __shfl_xor_sync(0xffffffffu, part, 1);
                                part += __shfl_xor_sync(0xffffffffu, part, 2);
                                if (t4 == 0 && r < n) wb[r] -= part;
                            }
                        }
                    }
                }
            }
            fence_global_to_tma();
            __syncwarp();
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------
// Same operand requirements as the TMA kernels (16-byte aligned rows and problems); any 64 <= n_max <= KMPX_SMALL_MAX.
inline bool wpp_applicable(const double* A, int lda, long long strideA, int n_max, int batch) {
    if (n_max < 2 * NB || n_max > KMPX_SMALL_MAX) return false;
    if ((lda & 1) || (batch > 1 && (strideA & 1)) || (reinterpret_cast<uintptr_t>(A) & 15)) return false;
    return true;
}
template <int V> inline int wpp_grid(int batch) {
    const int ctas = (batch + WppCfg<V>::WARPS - 1) / WppCfg<V>::WARPS;
    return ctas < num_sms() ? ctas : num_sms();
}
constexpr int WPP_V_POTRF = 0, WPP_V_TRTRI = 2;       // default variants (kmpx_debug_set bits 128 / 256 add 1 / 2 modulo 3: experiments)

template <int V>
inline int launch_potrf_wpp_v(const CUtensorMap& map, const FactorArgs& p, int struct_o, cudaStream_t st) {
    static SmemCache smem_set;
    KMPX_TRY(ensure_dyn_smem(k_potrf_wpp<V>, WppCfg<V>::SMEM, smem_set, "k_potrf_wpp"));
    k_potrf_wpp<V><<<wpp_grid<V>(p.batch), WppCfg<V>::THREADS, WppCfg<V>::SMEM, st>>>(map, p, struct_o);
    return check_launch("k_potrf_wpp");
}
template <int V>
inline int launch_trtri_wpp_v(const CUtensorMap& map, const CUtensorMap& mapb, const FactorArgs& p, int struct_o, const WppSolve& sv, cudaStream_t st) {
    static SmemCache smem_set;
    KMPX_TRY(ensure_dyn_smem(k_trtri_wpp<V>, WppCfg<V>::SMEM, smem_set, "k_trtri_wpp"));
    k_trtri_wpp<V><<<wpp_grid<V>(p.batch), WppCfg<V>::THREADS, WppCfg<V>::SMEM, st>>>(map, mapb, p, struct_o, sv);
    return check_launch("k_trtri_wpp");
}
inline int wpp_variant(int dflt) { return (dflt + ((g_host_dbg >> 7) & 3)) % 3; }
inline int launch_potrf_wpp(double* A, int lda, long long strideA, const int* n_sys, int n_max, int batch, int* info,
                            int struct_o, cudaStream_t st) {
    CUtensorMap map;
    KMPX_TRY(make_batch_map(&map, A, lda, strideA, n_max, batch, NB));
    FactorArgs p{A, lda, strideA, n_sys, n_max, batch, info, 0, 0};
    switch (wpp_variant(WPP_V_POTRF)) {
        case 0: return launch_potrf_wpp_v<0>(map, p, struct_o, st);
        case 1: return launch_potrf_wpp_v<1>(map, p, struct_o, st);
        default: return launch_potrf_wpp_v<2>(map, p, struct_o, st);
    }
}
inline int launch_trtri_wpp(double* L, int lda, long long strideA, const int* n_sys, int n_max, int batch, int struct_o, const WppSolve& sv,
                            cudaStream_t st) {
    CUtensorMap map, mapb;
    KMPX_TRY(make_batch_map(&map, L, lda, strideA, n_max, batch, NB));
    KMPX_TRY(make_batch_map(&mapb, L, lda, strideA, n_max, batch, 8));
    FactorArgs p{L, lda, strideA, n_sys, n_max, batch, nullptr, 0, 0};
    switch (wpp_variant(WPP_V_TRTRI)) {
        case 0: return launch_trtri_wpp_v<0>(map, mapb, p, struct_o, sv, st);
        case 1: return launch_trtri_wpp_v<1>(map, mapb, p, struct_o, sv, st);
        default: return launch_trtri_wpp_v<2>(map, mapb, p, struct_o, sv, st);
    }
}

}  // namespace kmpx
