// TMA-fed variants of the small-system factor / inverse kernels (textually included by factor_small.cu).
//
// Same algorithms as k_potrf_small / k_trtri_small; what changes is how the streamed operand of the block-column
// contraction reaches the tensor pipe:
//   * the (rows x 8) k-tiles are fetched by the TMA unit (cp.async.bulk.tensor, one 3-D tensor map over
//     [problem][row][column], boxes of 64 rows x 8 columns, 64-byte swizzle) into a ring of dense shared-memory
//     stages; one thread issues them, completion is signalled on an mbarrier per stage (full) and the eight
//     warps hand a stage back through a second mbarrier (empty) - there is no block barrier in the K-loop and no
//     per-thread cp.async address arithmetic competing with the fragment loads for the load/store pipe;
//   * k-tiles that only depend on block columns that are already final are requested while the CTA is still busy
//     with the diagonal block of the previous column (look-ahead across block columns);
//   * the block column being produced never passes through shared memory: it is read from global memory straight
//     into the accumulator layout, multiplied by the inverse of the diagonal block with the A fragments formed
//     by warp shuffles, and stored from registers.
//
// Swizzled stage layout (64 B per row, 16-byte chunk q of row r stored at chunk q ^ ((r >> 1) & 3)): a DMMA k-step
// uses the columns {0,1,4,5} (step 0) or {2,3,6,7} (step 1) of the k-tile, which makes the 8-byte fragment reads of
// a half-warp (rows g = 0..3, four columns) hit sixteen distinct banks.  Both operands use the same column order,
// so the contraction is unchanged.
#pragma once
#include "tma_util.cuh"

namespace kmpx {

constexpr int TBOX = 64;                  // rows per TMA box
constexpr int TBOX_BYTES = TBOX * 64;     // 64 rows x 8 doubles
constexpr int TMAX_STAGES = 8;
constexpr size_t kMaxDynSmem = 227 * 1024 - 4096;   // 227 KB minus the static shared memory of the kernels

// The ring of k-tiles shared by both kernels: tile t lives in stage t % S, its barriers are in phase (t / S) & 1.
// Consumers (the eight MMA warps) and the producer warp each walk the ring with their own cursor.
struct TileRing {
    double* stage0; uint32_t full0, empty0; int S, stage_elems;
    int stage; uint32_t par; bool wrapped;
    __device__ __forceinline__ void advance() { if (++stage == S) { stage = 0; par ^= 1; wrapped = true; } }
    __device__ __forceinline__ const double* wait_tile() const {      // consumer: the next tile has landed
        mbar_wait(full0 + 8 * stage, par);
        return stage0 + (size_t)stage * stage_elems;
    }
    __device__ __forceinline__ void release_tile(int lane) {          // consumer: this warp is done with it
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * stage);
        advance();
    }
    // producer warp: waits until every MMA warp has released the previous content of the next stage, arms its barrier
    __device__ __forceinline__ uint32_t acquire_stage(uint32_t& bar, int nboxes, int lane) {
        if (wrapped) mbar_wait(empty0 + 8 * stage, par ^ 1);
        bar = full0 + 8 * stage;
        if (lane == 0) mbar_expect_tx(bar, (uint32_t)nboxes * TBOX_BYTES);
        const uint32_t dst = smem_u32(stage0 + (size_t)stage * stage_elems);
        advance();
        return dst;
    }
};
// barrier among the 96 threads of the three auxiliary warps
__device__ __forceinline__ void aux_warps_sync() { asm volatile("bar.sync 2, 96;\n" ::: "memory"); }
// barrier among the 256 threads of the MMA warps (the producer warp never joins it)
__device__ __forceinline__ void mma_warps_sync() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }
// progress counter in shared memory: written by one MMA thread after mma_warps_sync(), polled by the producer warp
__device__ __forceinline__ void progress_publish(int* ctr, int v) {
    asm volatile("st.release.cta.shared.s32 [%0], %1;\n" :: "r"(smem_u32(ctr)), "r"(v) : "memory");
}
__device__ __forceinline__ int progress_wait(int* ctr, int need, int seen) {
    while (seen < need) asm volatile("ld.acquire.cta.shared.s32 %0, [%1];\n" : "=r"(seen) : "r"(smem_u32(ctr)) : "memory");
    return seen;
}

// Per-lane offsets (in doubles, relative to the first row of an 8-row group) of the fragment element for the two
// k-steps of a swizzled k-tile.
__device__ __forceinline__ void swz_offsets(int lane, int& o0, int& o1) {
    const int g = lane >> 2, t4 = lane & 3;
    const int sw = (g >> 1) & 3, c0 = (t4 >> 1) << 1;
    o0 = g * 8 + ((c0 ^ sw) << 1) + (t4 & 1);
    o1 = g * 8 + (((c0 | 1) ^ sw) << 1) + (t4 & 1);
}

// A fragment (rows g, k-columns 4q..4q+3) of a 8 x 32 row tile held in accumulator layout: element
// (g, 4q + t4) lives in lane (g, 2(q&1) + t4/2), tile q/2, slot t4&1.
template <int Q>
__device__ __forceinline__ double c_to_a_frag(const double (&row_tile)[4][2], int lane) {
    const int t4 = lane & 3;
    const int src = (lane & ~3) | ((Q & 1) << 1) | (t4 >> 1);
    const double v0 = __shfl_sync(0xffffffffu, row_tile[Q >> 1][0], src);
    const double v1 = __shfl_sync(0xffffffffu, row_tile[Q >> 1][1], src);
    return (t4 & 1) ? v1 : v0;
}

// In place, for one warp and one 32 x 32 chunk in accumulator layout:  P <- scale * P * op(D), D a lower-triangular
// 32 x 32 block in shared memory (leading dimension FLDP).  D_TRANS: op(D) = D^T, else D.  Zero blocks of D are skipped.
template <bool D_TRANS>
__device__ __forceinline__ void chunk_mul_regs(double (&P)[4][4][2], const double* D, double scale, int lane) {
    const int g = lane >> 2, t4 = lane & 3;
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
                // D^T: B[k][n] = D[n][k], non-zero for k <= n  ->  4q <= 8j+7;   D: B[k][n] = D[k][n], k >= n -> 4q+3 >= 8j
                const bool live = D_TRANS ? (4 * q <= 8 * j + 7) : (4 * q + 3 >= 8 * j);
                if (live) {
                    const double b = D_TRANS ? D[(j * 8 + g) * FLDP + 4 * q + t4] : D[(4 * q + t4) * FLDP + j * 8 + g];
                    dmma884(out[j], a[q], b);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { P[i][j][0] = scale * out[j][0]; P[i][j][1] = scale * out[j][1]; }
    }
}

// acc[c] += (32 tile rows starting at row offset ro_c) times the B fragments, for the two k-steps of one swizzled k-tile.
//   B_FROM_TILE: B[k][n] = tile row n (the first 32 rows of the tile)         (potrf)
//   else:        B[k][n] = Bk[(4*step + t4) * FLDB + n]  (k-major block in shared memory, rows already in k-step order)
//   self (warp-uniform): acc[1] += R * R^T for the 32 tile rows R starting at ro1 (both operands are the same fragments)
template <bool B_FROM_TILE>
__device__ __forceinline__ void tile_mma(double (&acc)[2][4][4][2], const double* tile, const double* Bk, int o0, int o1,
                                         bool live0, int ro0, bool live1, int ro1, bool self, int lane) {
    const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int o = s ? o1 : o0;
        double bf[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) bf[j] = B_FROM_TILE ? tile[j * 64 + o] : Bk[(4 * s + t4) * FLDB + j * 8 + g];
        if (live0) {
            double a[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = tile[(ro0 + i * 8) * 8 + o];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[0][i][j], a[i], bf[j]);
        }
        if (live1 || self) {
            double a[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = tile[(ro1 + i * 8) * 8 + o];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[1][i][j], a[i], self ? a[j] : bf[j]);
        }
    }
}

struct TmaSmem {   // carve-up of the dynamic shared memory of both kernels
    int S, ring_rows, t_rows;
    size_t ring_bytes, t_bytes, blocks_bytes, total;
};
__host__ __device__ inline TmaSmem tma_smem_layout(int n_max, bool with_t, size_t budget) {   // with_t: trtri
    TmaSmem L;
    const int below = n_max > NB ? n_max - NB : 0;          // rows of the largest streamed tile
    L.ring_rows = (below + TBOX - 1) / TBOX * TBOX;
    L.t_rows = with_t ? (below + 7) / 8 * 8 : 0;
    L.t_bytes = (size_t)L.t_rows * FLDP * sizeof(double);
    L.blocks_bytes = (size_t)(with_t ? 3 : 5) * 32 * FLDP * sizeof(double);
    const size_t fixed = L.t_bytes + L.blocks_bytes + 2 * TMAX_STAGES * 8 + 16 /* progress counters */ + 1024 /* alignment slack */;
    const size_t stage = (size_t)L.ring_rows * 64;
    int S = stage == 0 || budget < fixed ? 0 : (int)((budget - fixed) / stage);
    L.S = S > TMAX_STAGES ? TMAX_STAGES : S;
    L.ring_bytes = (size_t)L.S * stage;
    L.total = L.ring_bytes + fixed;
    return L;
}

// ------------------------------------------------------------------------------------------------------------
// Roles inside a CTA of 384 threads:
//   warps 0-7   MMA warps: block-column contraction, panel solve, stores
//   warp  8     producer: issues the TMA boxes
//   warps 9-11  auxiliary warps: Cholesky factor and inverse of the 32 x 32 diagonal blocks, concurrently with the
//               contraction (the diagonal block of the NEXT column is finished one step early, see k_potrf_tma)
constexpr int TTHREADS = FTHREADS + 128;
constexpr int AUXT = 96;
// Register split between the roles (setmaxnreg works on warpgroups of four warps): each scheduler hosts two MMA
// warps and one warp of the producer/auxiliary group, 2 * 208 + 88 registers per lane fit its 16K-entry file
// (the auxiliary code needs 72; ptxas does not check a region against its setmaxnreg budget - see k_aux_regs_probe).
__device__ __forceinline__ void producer_regs() { asm volatile("setmaxnreg.dec.sync.aligned.u32 88;\n"); }
__device__ __forceinline__ void mma_regs() { asm volatile("setmaxnreg.inc.sync.aligned.u32 208;\n"); }

struct TmaCta {   // what both kernels set up identically
    TileRing ring; double* blocks; int* progress;     // progress[0]: columns/steps done, [1], [2]: hand-overs with the aux warps
};
__device__ __forceinline__ TmaCta tma_cta_setup(unsigned char* smem_raw, const TmaSmem& lay, int tid) {
    unsigned char* base = smem_align1k(smem_raw);
    TmaCta c;
    c.blocks = reinterpret_cast<double*>(base + lay.ring_bytes);        // [T rows (trtri)] + the 32 x FLDP blocks
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + lay.ring_bytes + lay.t_bytes + lay.blocks_bytes);
    c.progress = reinterpret_cast<int*>(bars + 2 * TMAX_STAGES);
    c.ring.stage0 = reinterpret_cast<double*>(base); c.ring.S = lay.S; c.ring.stage_elems = lay.ring_rows * 8;
    c.ring.full0 = smem_u32(bars); c.ring.empty0 = smem_u32(bars + TMAX_STAGES);
    c.ring.stage = 0; c.ring.par = 0; c.ring.wrapped = false;
    if (tid == 0) {
        for (int s = 0; s < lay.S; ++s) { mbar_init(c.ring.full0 + 8 * s, 1); mbar_init(c.ring.empty0 + 8 * s, FWARPS); }
        c.progress[0] = 0; c.progress[1] = 0; c.progress[2] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    return c;
}

// Reciprocal square root for the pivot chain of the 8 x 8 blocks, where every dependent FP64 operation costs its full
// pipeline latency: hardware seed (rsqrt.approx.f64, ~2^-22) + ONE third-order step
//   e = 1 - x y^2,   y' = y + y e (1/2 + 3/8 e)          (error O(e^3) < 2^-64)
// = 5 dependent operations instead of the 9 of a single-precision seed with two Newton steps.
__device__ __forceinline__ double rsqrt_chain(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double xy = x * y;
    const double e = fma(-xy, y, 1.0);
    const double pp = fma(0.375, e, 0.5);
    const double ye = y * e;
    return fma(ye, pp, y);
}

// One warp: Cholesky factor and inverse of the w x w (w <= 8, identity padded) diagonal block of P at (base, base).
// Lane r (mod 8) owns row r of the factor and column r of the inverse; the diagonal is replicated in every lane so that
// the pivot needs no broadcast; the dependent chain per pivot is rsqrt (5 ops), scale, diagonal update (7 FP64 ops).
// The inverse x[i] = -(sum_{q<i} L[i][q] x[q]) / L_ii advances one row per pivot, off the critical path.
// Outputs: Ld (lower triangle of the block), S[i*8+j] = Xblock[i][j], X diagonal block; *bad = first bad pivot (1-based).
__device__ __forceinline__ void chol8_inverse(const double* P, double* Ld, double* X, double* S, int base, int w, int lane, int* bad) {
    const int r = lane & 7;
    double a[8], dd[8], x[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        a[c] = (r < w && c < r) ? P[(base + r) * FLDP + base + c] : 0.0;
        dd[c] = c < w ? P[(base + c) * FLDP + base + c] : 1.0;
    }
    int bd = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        // column j, still unscaled, broadcast while the rsqrt of the pivot is in flight: no shuffle on the pivot chain
        double ac[8];
#pragma unroll
        for (int c = j + 1; c < 8; ++c) ac[c] = __shfl_sync(0xffffffffu, a[j], c);
        const double d = dd[j];
        if (!(d > 0.0) && bd == 0 && j < w) bd = base + j + 1;
        const double rs = rsqrt_chain(d);
        const double l = r > j ? a[j] * rs : 0.0;
        a[j] = (r == j) ? d * rs : l;
#pragma unroll
        for (int c = j + 1; c < 8; ++c) {
            const double lc = ac[c] * rs;
            dd[c] = fma(-lc, lc, dd[c]);
            if (r > c) a[c] = fma(-l, lc, a[c]);
        }
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < j; ++q) sum = fma(__shfl_sync(0xffffffffu, a[q], j), x[q], sum);
        x[j] = (j < r) ? 0.0 : ((j == r) ? rs : -sum * rs);
    }
    if (lane < 8) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (r < w && c <= r) Ld[(base + r) * FLDP + base + c] = a[c];
            S[c * 8 + r] = x[c];
            X[(base + c) * FLDP + base + r] = x[c];
        }
        if (lane == 0 && bd != 0 && *bad == 0) *bad = bd;
    }
}

// ---- diagonal 32 x 32 blocks on the three auxiliary warps (t = 0..95, at most 64 registers) ----------------------
// Cholesky factor of the kb x kb block in P -> Ld (identity padded) AND the inverses of its four 8 x 8 diagonal blocks
// -> X (everything else of X zeroed).  8-column panels; the 8 x 8 block itself is factored and inverted by eight
// lanes of one warp that hold one row / one column each and exchange entries by shuffles.
__device__ __forceinline__ void aux_potf2(double* P, double* Ld, double* X, double* S, int kb, int t, int* bad) {
    const int lane = t & 31;
    for (int idx = t; idx < 32 * 32; idx += AUXT) {
        const int r = idx >> 5, c = idx & 31;
        Ld[r * FLDP + c] = (r == c) ? 1.0 : 0.0;
        X[r * FLDP + c] = 0.0;
    }
    if (t == 0) *bad = 0;
    aux_warps_sync();
    for (int base = 0; base < 32; base += 8) {
        const int w = min(8, kb - base);      // <= 0: identity block
        long long t_in = clock64();
        if (t < 32) chol8_inverse(P, Ld, X, S, base, w, lane, bad);
        if (blockIdx.x == 0 && t == 0) { const long long t2 = clock64(); g_phase[6] += t2 - t_in; t_in = t2; }
        aux_warps_sync();
        const int rb = kb - base - 8;        // rows below this 8 x 8 block
        if (rb > 0) {
            // panel solve and rank-8 update on the tensor pipe, 8 x 8 tiles dealt to the three warps
            const int aw = t >> 5, g = lane >> 2, t4 = lane & 3;
            const int ntile = (rb + 7) >> 3;
            for (int rt = aw; rt < ntile; rt += 3) {         // Ld[rows, panel] = P[rows, panel] * Xblock^T
                const int row = base + 8 + 8 * rt + g;
                double c[2] = {0.0, 0.0};
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) dmma884(c, P[row * FLDP + base + 4 * ks + t4], S[g * 8 + 4 * ks + t4]);
                if (row < kb) { Ld[row * FLDP + base + 2 * t4] = c[0]; Ld[row * FLDP + base + 2 * t4 + 1] = c[1]; }
            }
            aux_warps_sync();
            int tile = 0;
            for (int ti = 0; ti < ntile; ++ti)
                for (int tj = 0; tj <= ti; ++tj, ++tile) {   // P[ti, tj] -= Lp[ti] * Lp[tj]^T
                    if (tile % 3 != aw) continue;
                    const int row = base + 8 + 8 * ti + g;
                    double* cp = P + row * FLDP + base + 8 + 8 * tj + 2 * t4;
                    double c[2] = {cp[0], cp[1]};
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        dmma884(c, -Ld[row * FLDP + base + 4 * ks + t4], Ld[(base + 8 + 8 * tj + g) * FLDP + base + 4 * ks + t4]);
                    cp[0] = c[0]; cp[1] = c[1];
                }
            aux_warps_sync();
        }
        if (blockIdx.x == 0 && t == 0) g_phase[7] += clock64() - t_in;
    }
}
// Branch-free reciprocal (hardware seed + two Newton steps, ~1 ulp).
__device__ __forceinline__ double rcp_chain(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
// Inverses of the four 8 x 8 diagonal blocks of the lower-triangular L -> X (rest of X zeroed), one thread per column.
// The reciprocals of the diagonal are formed up front (lane (block, c) holds entry c, shuffled to the block's lanes),
// so that a row costs two dependent operations (last product of the sum, scaling) instead of a division.
__device__ __forceinline__ void aux_diag8_inverse(const double* L, double* X, int t) {
    for (int idx = t; idx < 32 * 32; idx += AUXT) X[(idx >> 5) * FLDP + (idx & 31)] = 0.0;
    aux_warps_sync();
    if (t < 32) {
        const int base = (t >> 3) * 8, c = t & 7;
        double x[8];
        const double rc = rcp_chain(L[(base + c) * FLDP + base + c]);      // lane (block, c): reciprocal of diagonal entry c
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const double ri = __shfl_sync(0xffffffffu, rc, (t & 24) | i);
            double v = 0.0;
            if (i == c) v = ri;
            else if (i > c) {
                double sum = 0.0;
#pragma unroll
                for (int q = 0; q < 8; ++q) if (q >= c && q < i) sum += L[(base + i) * FLDP + base + q] * x[q];
                v = -sum * ri;
            }
            x[i] = v;
            X[(base + i) * FLDP + base + c] = v;
        }
    }
    aux_warps_sync();
}
// Completes X = L^{-1} (32 x 32) from the inverses of the 8 x 8 diagonal blocks: 8 -> 16 -> 32,
// [[L00,0],[L10,L11]]^{-1} = [[X00,0],[-X11 L10 X00, X11]].  S: 256 doubles of scratch.
__device__ __forceinline__ void aux_trtri32_finish(const double* L, double* X, double* S, int t) {
    for (int idx = t; idx < 128; idx += AUXT) {            // level 16: T = L10 X00 for both 16-blocks
        const int pb = (idx >> 6) * 16, r = (idx >> 3) & 7, c = idx & 7;
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) sum += L[(pb + 8 + r) * FLDP + pb + q] * X[(pb + q) * FLDP + pb + c];
        S[idx] = sum;
    }
    aux_warps_sync();
    for (int idx = t; idx < 128; idx += AUXT) {            // X10 = -X11 T
        const int qd = idx >> 6, pb = qd * 16, r = (idx >> 3) & 7, c = idx & 7;
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) sum += X[(pb + 8 + r) * FLDP + pb + 8 + q] * S[qd * 64 + q * 8 + c];
        X[(pb + 8 + r) * FLDP + pb + c] = -sum;
    }
    aux_warps_sync();
    for (int idx = t; idx < 256; idx += AUXT) {            // level 32: T = L10 X00 (16 x 16)
        const int r = idx >> 4, c = idx & 15;
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < 16; ++q) sum += L[(16 + r) * FLDP + q] * X[q * FLDP + c];
        S[idx] = sum;
    }
    aux_warps_sync();
    for (int idx = t; idx < 256; idx += AUXT) {
        const int r = idx >> 4, c = idx & 15;
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < 16; ++q) sum += X[(16 + r) * FLDP + 16 + q] * S[q * 16 + c];
        X[(16 + r) * FLDP + c] = -sum;
    }
    aux_warps_sync();
}

// L2 prefetch of rows [c0, n) x columns [c0, c0 + NB) of A by the 256 threads of the MMA warps (a row segment of
// 256 bytes touches two or three 128-byte lines depending on its alignment).
__device__ __forceinline__ void prefetch_block_column(const double* A, int lda, int c0, int n, int tid) {
    const int w = min(NB, n - c0) * (int)sizeof(double);
    for (int r = c0 + tid; r < n; r += FTHREADS) {
        const char* q = reinterpret_cast<const char*>(A + (size_t)r * lda + c0);
        asm volatile("prefetch.global.L2 [%0];\n" :: "l"(q));
        if (w > 128) asm volatile("prefetch.global.L2 [%0];\n" :: "l"(q + 128));
        asm volatile("prefetch.global.L2 [%0];\n" :: "l"(q + w - 8));
    }
}

// ------------------------------------------------------------------------------------------------------------
// Left-looking blocked Cholesky.  Step j (block column k0 = 32 j, m = n - k0 rows):
//   MMA warps   acc(unit u) = L[rows of unit u, 0:k0] * L[k0:k0+32, 0:k0]^T for the 32-row units below the diagonal
//               block (unit u -> warp u % 7), and - warp 7, second accumulator - the self product of unit 0, i.e. the
//               history of the NEXT diagonal block;   X(u) = (A - acc) * Dinv_j^T, stored from registers; warp 7 adds
//               A_{j+1,j+1} and -X(0) X(0)^T (X(0) handed over in shared memory) and publishes the next diagonal block.
//   aux warps   factor and invert the diagonal block j (published at the end of step j-1) while the MMA warps are in
//               the contraction of step j; they also store L_jj.
__global__ void __launch_bounds__(TTHREADS, 1) k_potrf_tma(const __grid_constant__ CUtensorMap tmap, FactorArgs p, TmaSmem lay) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    __shared__ int sh_bad;
    __shared__ double scratch[256];
    TmaCta cta = tma_cta_setup(smem_raw, lay, tid);
    TileRing& ring = cta.ring;
    double* Dblk = cta.blocks;            // updated diagonal block (warp 7 -> aux warps)
    double* Ld = Dblk + 32 * FLDP;
    double* Dinv2 = Ld + 32 * FLDP;       // aux warps -> MMA warps, two buffers alternating over the blocks of the CTA
    double* Xs = Dinv2 + 2 * 32 * FLDP;   // X(0) of the current step (warp 0 -> warp 7)
    int* diag_ready = cta.progress + 1;   // diagonal blocks published by warp 7
    int* dinv_ready = cta.progress + 2;   // diagonal blocks factored + inverted by the aux warps

    if (warp >= FWARPS) {
        producer_regs();
        if (warp == FWARPS) {
            // ---- producer warp: k-tile pkt (columns 8*pkt..) of the block column starting at pk0 needs the first
            //      pkt/4 + 1 block columns of this problem to be final; progress[0] counts finished block columns
            int done_base = 0, seen = 0;
            for (int b = blockIdx.x; b < p.batch; b += gridDim.x) {
                const int n = p.n_sys ? p.n_sys[b] : p.n_max;
                for (int pk0 = NB; pk0 + NB < n; pk0 += NB) {     // the last block column has no rows below its diagonal block
                    const int nboxes = (n - pk0 + TBOX - 1) / TBOX;
                    for (int pkt = 0; pkt < pk0 / 8; ++pkt) {
                        seen = progress_wait(cta.progress, done_base + (pkt >> 2) + 1, seen);
                        uint32_t bar;
                        const uint32_t dst = ring.acquire_stage(bar, nboxes, lane);
                        if (lane == 0)
                            for (int q = 0; q < nboxes; ++q) tma_load_box(dst + q * TBOX_BYTES, &tmap, 8 * pkt, pk0 + q * TBOX, b, bar);
                        __syncwarp();
                    }
                }
                done_base += (n + NB - 1) / NB;
            }
            return;
        }
        // ---- auxiliary warps
        const int t = tid - (FWARPS + 1) * 32;
        int blocks = 0, seen = 0;
        for (int b = blockIdx.x; b < p.batch; b += gridDim.x) {
            const int n = p.n_sys ? p.n_sys[b] : p.n_max;
            double* A = p.A + (size_t)b * p.strideA;
            for (int k0 = 0; k0 < n; k0 += NB) {
                const int kb = min(NB, n - k0);
                long long t_aux = clock64();
#define AUX_MARK(slot) do { if (blockIdx.x == 0 && t == 0) { const long long _t = clock64(); g_phase[slot] += _t - t_aux; t_aux = _t; } } while (0)
                double* Dinv = Dinv2 + (blocks & 1) * 32 * FLDP;
                seen = progress_wait(diag_ready, blocks + 1, seen);
                AUX_MARK(12);
                aux_potf2(Dblk, Ld, Dinv, scratch, kb, t, &sh_bad);
                AUX_MARK(13);
                if (n - k0 > NB) aux_trtri32_finish(Ld, Dinv, scratch, t);
                ++blocks;
                if (t == 0) progress_publish(dinv_ready, blocks);      // (both helpers end with a barrier of the aux warps)
                AUX_MARK(14);
                // L_jj itself is off the critical path (no later step streams it)
                for (int idx = t; idx < NB * NB; idx += AUXT) {
                    const int r = idx >> 5, c = idx & 31;
                    if (r < kb && c <= r) A[(size_t)(k0 + r) * p.lda + k0 + c] = Ld[r * FLDP + c];
                }
                aux_warps_sync();
                AUX_MARK(15);
            }
        }
        return;
    }

    mma_regs();
    int o0, o1;
    swz_offsets(lane, o0, o1);
    int done = 0, seen_dinv = 0;
    for (int b = blockIdx.x; b < p.batch; b += gridDim.x) {
        const int n = p.n_sys ? p.n_sys[b] : p.n_max;
        double* A = p.A + (size_t)b * p.strideA;
        int bad_total = 0;
        long long t_phase = clock64();
        double acc[2][4][4][2];
        zero_acc(acc[1]);
        for (int k0 = 0; k0 < n; k0 += NB) {
            const int kb = min(NB, n - k0);
            const int m = n - k0;
            const int nun = (m - 1) >> 5;                 // 32-row units below the diagonal block
            // unit u -> (warp u % 7, accumulator u / 7) for u < 14, unit 14 (n > 480 only) -> (warp 7, accumulator 0):
            // warp 7 normally carries nothing but the self product, whose tail (wait for X(0), rank-32 update, publish)
            // is then as long as the two panel products of the other warps instead of coming on top of one
            const int u0 = warp < FWARPS - 1 ? warp : 14, u1 = warp + (FWARPS - 1);
            const bool live0 = u0 < nun, live1 = warp < FWARPS - 1 && u1 < nun;
            const bool self = warp == FWARPS - 1 && nun > 0;          // warp 7's second accumulator: next diagonal block
            // ---- first block column of a problem: warp 7 publishes A_00 (later diagonal blocks are published one step early)
            if (warp == FWARPS - 1 && k0 == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int r = i * 8 + g, c = j * 8 + 2 * t4 + e;
                            Dblk[r * FLDP + c] = (r < n && c < n) ? A[(size_t)r * p.lda + c] : 0.0;
                        }
                __syncwarp();
                if (lane == 0) progress_publish(diag_ready, done + 1);
            }
            zero_acc(acc[0]);
            zero_acc(acc[1]);
            // ---- L2 prefetch of what the NEXT step reads straight from global memory (its block column of A, diagonal
            //      block included), and in the last two steps the first two block columns of the CTA's next problem:
            //      the per-SM share of the HBM bandwidth (~27 B/clk) would otherwise be paid as a burst in front of the
            //      panel solve (128 KB per step) while the memory system idles during the contraction
            if (!(g_dbg & 32)) {
                const int last = (n - 1) / NB * NB;                 // k0 of the last step
                const int bn = b + gridDim.x;
                const int nn = (k0 + 2 * NB > last && bn < p.batch) ? (p.n_sys ? p.n_sys[bn] : p.n_max) : 0;
                if (k0 + NB <= last) prefetch_block_column(A, p.lda, k0 + NB, n, tid);
                else if (nn > NB) prefetch_block_column(p.A + (size_t)bn * p.strideA, p.lda, NB, nn, tid);
                if (k0 + NB <= last && nn > 0) prefetch_block_column(p.A + (size_t)bn * p.strideA, p.lda, 0, nn, tid);
            }
            const int nk = nun > 0 ? k0 / 8 : 0;
            for (int kt = 0; kt < nk; ++kt) {
                const double* tile = ring.wait_tile();
                if (!(g_dbg & 1))
                    tile_mma<true>(acc, tile, nullptr, o0, o1, live0, 32 + u0 * 32, live1, self ? 32 : 32 + u1 * 32, self, lane);
                ring.release_tile(lane);
            }
            PHASE_MARK(0);
            // ---- P = A[rows, k0:k0+32] - acc straight from global memory into the accumulator layout (units exist only
            //      below full block columns, kb == 32): sixteen 16-byte loads per unit in flight, row index clamped.
            //      (Loading -A into the accumulators BEFORE the contraction instead was measured 2 % slower: the same
            //      transfer time then stalls the first DMMAs of every warp.)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int u = c ? u1 : u0;
                const bool nextdiag = c == 1 && self;
                if ((c ? live1 : live0) || nextdiag) {
                    const int row0 = k0 + NB + (nextdiag ? 0 : u * 32);
                    const int col0 = nextdiag ? k0 + NB : k0;
                    if (!nextdiag || n - (k0 + NB) >= NB) {
                        double2 v[4][4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int gr = min(row0 + i * 8 + g, n - 1);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                v[i][j] = *reinterpret_cast<const double2*>(A + (size_t)gr * p.lda + col0 + j * 8 + 2 * t4);
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const bool ok = row0 + i * 8 + g < n;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                acc[c][i][j][0] = ok ? v[i][j].x - acc[c][i][j][0] : 0.0;
                                acc[c][i][j][1] = ok ? v[i][j].y - acc[c][i][j][1] : 0.0;
                            }
                        }
                    } else {       // ragged last diagonal block: guarded element-wise loads
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j)
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const int gr = row0 + i * 8 + g, gc = col0 + j * 8 + 2 * t4 + e;
                                    const double v = (gr < n && gc < n) ? A[(size_t)gr * p.lda + gc] : 0.0;
                                    acc[c][i][j][e] = (gr < n && gc < n) ? v - acc[c][i][j][e] : 0.0;
                                }
                    }
                }
            }
            PHASE_MARK(1);
            // ---- wait for the factor / inverse of the diagonal block of this step
            seen_dinv = progress_wait(dinv_ready, done + 1, seen_dinv);
            PHASE_MARK(2);
            if (sh_bad != 0 && bad_total == 0) bad_total = k0 + sh_bad;
            const double* Dinv = Dinv2 + (done & 1) * 32 * FLDP;
            // ---- next diagonal block first: warp 7 takes X(0) from warp 0 through shared memory, subtracts X(0) X(0)^T from
            //      A_{j+1,j+1} - history (its second accumulator) and publishes the block, so that the aux warps factor
            //      it while the MMA warps are still storing this block column and contracting the next one
            if (self) {
                asm volatile("bar.sync 3, 64;\n" ::: "memory");          // warp 0 arrives below once Xs is written
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    double xf[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) xf[i] = Xs[(i * 8 + g) * FLDP + 4 * q + t4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) dmma884(acc[1][i][j], -xf[i], xf[j]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<double2*>(Dblk + (i * 8 + g) * FLDP + j * 8 + 2 * t4) = make_double2(acc[1][i][j][0], acc[1][i][j][1]);
                __syncwarp();
                if (lane == 0) progress_publish(diag_ready, done + 2);
            }
            // ---- X = P * L_kk^{-T}, stored from registers
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int u = c ? u1 : u0;
                if (c ? live1 : live0) {
                    chunk_mul_regs<true>(acc[c], Dinv, 1.0, lane);
                    if (u == 0) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                *reinterpret_cast<double2*>(Xs + (i * 8 + g) * FLDP + j * 8 + 2 * t4) = make_double2(acc[c][i][j][0], acc[c][i][j][1]);
                        asm volatile("bar.arrive 3, 64;\n" ::: "memory");
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int gr = k0 + NB + u * 32 + i * 8 + g;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const double2 v = make_double2(acc[c][i][j][0], acc[c][i][j][1]);
                            if (gr < n) *reinterpret_cast<double2*>(A + (size_t)gr * p.lda + k0 + j * 8 + 2 * t4) = v;
                        }
                    }
                }
            }
            PHASE_MARK(4);
            fence_global_to_tma();   // the block column is read back by the TMA unit in later iterations
            mma_warps_sync();
            ++done;
            if (tid == 0) progress_publish(cta.progress, done);
            PHASE_MARK(5);
        }
        if (tid == 0 && p.info) {
            if (!p.info_accumulate) p.info[b] = bad_total;
            else if (bad_total != 0 && p.info[b] == 0) p.info[b] = p.info_offset + bad_total;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// In-place inverse of the lower factor, block columns right to left:  X[below, jb] = -Linv22 * (L21 * Dinv).
// T = L21 * Dinv lives in shared memory (k-major B operand, rows stored in k-step order); Linv22 is streamed by TMA,
// older block columns first so that they can be requested before the previous step has been written back.
// The aux warps invert the diagonal block of step jb-1 (double-buffered) while the MMA warps work on step jb, and write
// it back with an explicit zero upper triangle (the TMA tiles carry no triangular mask).
// tile order of a step with nk k-tiles: the tiles of older block columns (kt >= 4) first, then kt = 0..3
__device__ __forceinline__ int trtri_tile_of(int pos, int nk) { return nk <= 4 ? pos : (pos < nk - 4 ? pos + 4 : pos - (nk - 4)); }

__global__ void __launch_bounds__(TTHREADS, 1) k_trtri_tma(const __grid_constant__ CUtensorMap tmap, FactorArgs p, TmaSmem lay) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    __shared__ double scratch[256];
    TmaCta cta = tma_cta_setup(smem_raw, lay, tid);
    TileRing& ring = cta.ring;
    double* Tb = cta.blocks;                                  // [t_rows][FLDB]
    double* Dsm = Tb + (size_t)lay.t_rows * FLDB;             // aux warps: diagonal block being inverted
    double* Dinv2 = Dsm + 32 * FLDP;                          // two buffers, alternating over the steps of the CTA
    int* dinv_ready = cta.progress + 1;                       // diagonal blocks inverted by the aux warps (count)
    int* steps_done = cta.progress;                           // steps finished by the MMA warps (count)

    if (warp >= FWARPS) {
        producer_regs();
        if (warp == FWARPS) {
            // ---- producer warp: k-tile kt of step pjb lies in block column pjb + 1 + kt/4, which is final once the steps
            //      nblk-1 .. that column are done
            int done_base = 0, seen = 0;
            for (int b = blockIdx.x; b < p.batch; b += gridDim.x) {
                const int n = p.n_sys ? p.n_sys[b] : p.n_max;
                const int nblk = (n + NB - 1) / NB;
                for (int pjb = nblk - 2; pjb >= 0; --pjb) {
                    const int r0 = (pjb + 1) * NB;               // first trailing row / column of step pjb
                    const int m = n - r0;
                    const int nk = (m + 7) / 8;
                    const int nboxes = (m + TBOX - 1) / TBOX;
                    for (int pos = 0; pos < nk; ++pos) {
                        const int kt = trtri_tile_of(pos, nk);
                        seen = progress_wait(steps_done, done_base + nblk - (pjb + 1 + (kt >> 2)), seen);
                        const int q0 = (kt >> 2) >> 1;           // rows above 32*(kt/4) never need this k-tile
                        uint32_t bar;
                        const uint32_t dst = ring.acquire_stage(bar, nboxes - q0, lane);
                        if (lane == 0)
                            for (int q = q0; q < nboxes; ++q) tma_load_box(dst + q * TBOX_BYTES, &tmap, r0 + 8 * kt, r0 + q * TBOX, b, bar);
                        __syncwarp();
                    }
                }
                done_base += nblk;
            }
            return;
        }
        // ---- auxiliary warps: a block may overwrite its buffer once the MMA warps have finished the step before the previous one
        const int t = tid - (FWARPS + 1) * 32;
        int blocks = 0, seen = 0;
        for (int b = blockIdx.x; b < p.batch; b += gridDim.x) {
            const int n = p.n_sys ? p.n_sys[b] : p.n_max;
            double* L = p.A + (size_t)b * p.strideA;
            const int nblk = (n + NB - 1) / NB;
            for (int jb = nblk - 1; jb >= 0; --jb) {
                const int k0 = jb * NB, kb = min(NB, n - k0);
                seen = progress_wait(steps_done, blocks - 1, seen);
                double* Dinv = Dinv2 + (blocks & 1) * 32 * FLDP;
                for (int idx = t; idx < NB * NB; idx += AUXT) {
                    const int r = idx >> 5, c = idx & 31;
                    Dsm[r * FLDP + c] = (r < kb && c <= r) ? L[(size_t)(k0 + r) * p.lda + k0 + c] : ((r == c) ? 1.0 : 0.0);
                }
                aux_warps_sync();
                aux_diag8_inverse(Dsm, Dinv, t);
                aux_trtri32_finish(Dsm, Dinv, scratch, t);
                for (int idx = t; idx < NB * NB; idx += AUXT) {
                    const int r = idx >> 5, c = idx & 31;
                    if (r < kb && c < kb) L[(size_t)(k0 + r) * p.lda + k0 + c] = c <= r ? Dinv[r * FLDP + c] : 0.0;
                }
                fence_global_to_tma();
                aux_warps_sync();
                ++blocks;
                if (t == 0) progress_publish(dinv_ready, blocks);
            }
        }
        return;
    }

    mma_regs();
    int o0, o1;
    swz_offsets(lane, o0, o1);
    int done = 0, seen_dinv = 0;
    for (int b = blockIdx.x; b < p.batch; b += gridDim.x) {
        const int n = p.n_sys ? p.n_sys[b] : p.n_max;
        double* L = p.A + (size_t)b * p.strideA;
        const int nblk = (n + NB - 1) / NB;
        long long t_phase = clock64();
        // ragged batch: the columns n .. n_max-1 of this problem are streamed with the last k-tile; make them finite
        if (n < p.n_max) {
            const int cw = min(p.n_max, (n + 7) / 8 * 8) - n;
            for (int idx = tid; idx < n * cw; idx += FTHREADS) L[(size_t)(idx / cw) * p.lda + n + idx % cw] = 0.0;
        }
        for (int jb = nblk - 1; jb >= 0; --jb) {
            const int k0 = jb * NB;
            const int kb = min(NB, n - k0);
            const int m = n - k0 - kb;            // rows below the diagonal block
            const double* Dinv = Dinv2 + (done & 1) * 32 * FLDP;
            bool waited = false;      // the inverse of the diagonal block is awaited with the loads of L21 already in flight
            // L2 prefetch of the block column the NEXT step (and the aux warps) read from global memory; in the last step
            // the last two block columns of the CTA's next problem (see k_potrf_tma)
            if (!(g_dbg & 32)) {
                if (jb > 0) prefetch_block_column(L, p.lda, k0 - NB, n, tid);
                else if (b + (int)gridDim.x < p.batch) {
                    const int bn = b + gridDim.x;
                    const int nn = p.n_sys ? p.n_sys[bn] : p.n_max;
                    const int cl = (nn - 1) / NB * NB;
                    if (nn > 0) prefetch_block_column(p.A + (size_t)bn * p.strideA, p.lda, cl, nn, tid);
                    if (cl >= NB) prefetch_block_column(p.A + (size_t)bn * p.strideA, p.lda, cl - NB, nn, tid);
                }
            }
            if (m > 0) {
                const int r_g0 = k0 + NB;
                const int nch = (m + 31) >> 5;
                const bool live0 = warp < nch, live1 = warp + FWARPS < nch;
                const int nk = (m + 7) / 8;
                double acc[2][4][4][2];
                // T = L21 * Dinv: L21 straight into the accumulator layout, product by shuffles, rows to shared memory
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int ch = warp + c * FWARPS;
                    if (ch < nch) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {      // row index clamped, result masked: all loads in flight at once
                            const int r = ch * 32 + i * 8 + g;
                            const double* src = L + (size_t)(r_g0 + min(r, m - 1)) * p.lda + k0 + 2 * t4;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const double2 v = *reinterpret_cast<const double2*>(src + j * 8);
                                acc[0][i][j][0] = r < m ? v.x : 0.0; acc[0][i][j][1] = r < m ? v.y : 0.0;
                            }
                        }
                        if (!waited) { seen_dinv = progress_wait(dinv_ready, done + 1, seen_dinv); waited = true; PHASE_MARK(8); }
                        chunk_mul_regs<false>(acc[0], Dinv, 1.0, lane);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int r = ch * 32 + i * 8 + g;
                            // k-step order inside each group of 8 rows: {0,1,4,5 | 2,3,6,7}  (bits 1 and 2 swapped)
                            const int rp = (r & ~7) | (r & 1) | ((r & 2) << 1) | ((r & 4) >> 1);
                            if (rp < lay.t_rows) {
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    *reinterpret_cast<double2*>(Tb + (size_t)rp * FLDB + j * 8 + 2 * t4) =
                                        make_double2(acc[0][i][j][0], acc[0][i][j][1]);
                            }
                        }
                    }
                }
                if (!waited) { seen_dinv = progress_wait(dinv_ready, done + 1, seen_dinv); waited = true; }
                mma_warps_sync();
                PHASE_MARK(9);
                zero_acc(acc[0]);
                zero_acc(acc[1]);
                for (int pos = 0; pos < nk; ++pos) {
                    const int kt = trtri_tile_of(pos, nk);
                    const double* tile = ring.wait_tile();
                    // a 32-row chunk ch only needs k < 32*(ch+1)
                    if (!(g_dbg & 1))
                        tile_mma<false>(acc, tile, Tb + (size_t)kt * 8 * FLDB, o0, o1, live0 && kt < 4 * (warp + 1), warp * 32,
                                        live1 && kt < 4 * (warp + FWARPS + 1), (warp + FWARPS) * 32, false, lane);
                    ring.release_tile(lane);
                }
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
                                for (int j = 0; j < 4; ++j)
                                    *reinterpret_cast<double2*>(L + (size_t)(r_g0 + r) * p.lda + k0 + j * 8 + 2 * t4) =
                                        make_double2(-acc[c][i][j][0], -acc[c][i][j][1]);
                            }
                        }
                    }
                }
            }
            // (the aux warps' write-back of the inverted diagonal block precedes dinv_ready: it must be awaited before
            //  steps_done lets the producer stream that block)
            if (!waited) seen_dinv = progress_wait(dinv_ready, done + 1, seen_dinv);
            fence_global_to_tma();
            mma_warps_sync();   // also: every warp is done with T and this Dinv buffer
            ++done;
            if (tid == 0) progress_publish(steps_done, done);
            PHASE_MARK(11);
        }
    }
}


// register-budget probe for the code that runs on the auxiliary warps (never launched)
__global__ void __launch_bounds__(AUXT) k_aux_regs_probe(double* P, double* Ld, double* X, double* S, int kb, int* bad) {
    aux_potf2(P, Ld, X, S, kb, threadIdx.x, bad);
    aux_diag8_inverse(Ld, X, threadIdx.x);
    aux_trtri32_finish(Ld, X, S, threadIdx.x);
}

// ---- host side ---------------------------------------------------------------------------------------------
// The TMA kernels need 16-byte aligned rows and problems, and at least three ring stages next to the fixed buffers.
inline bool tma_applicable(const double* A, int lda, long long strideA, int n_max, int batch, bool with_t) {
    if (g_host_dbg & 16) return false;
    if (n_max < 2 * NB || (lda & 1) || (batch > 1 && (strideA & 1)) || (reinterpret_cast<uintptr_t>(A) & 15)) return false;
    return tma_smem_layout(n_max, with_t, kMaxDynSmem).S >= 3;
}

}  // namespace kmpx
