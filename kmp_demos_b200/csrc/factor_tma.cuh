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
#include <cuda.h>

namespace kmpx {

constexpr int TBOX = 64;                  // rows per TMA box
constexpr int TBOX_BYTES = TBOX * 64;     // 64 rows x 8 doubles
constexpr int TMAX_STAGES = 8;
constexpr size_t kMaxDynSmem = 227 * 1024 - 4096;   // 227 KB minus the static shared memory of the kernels

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}
// box (8 columns x 64 rows) of problem `b` at (col, row) -> shared memory, completion on `bar`
__device__ __forceinline__ void tma_load_box(uint32_t dst, const CUtensorMap* map, int col, int row, int b, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
        :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(col), "r"(row), "r"(b), "r"(bar) : "memory");
}
// makes global-memory writes of this thread visible to later TMA (async proxy) reads
__device__ __forceinline__ void fence_global_to_tma() {
    __threadfence();
    asm volatile("fence.proxy.async;\n" ::: "memory");
}

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

// acc[c] += tile rows of chunk (warp + 8c) times the B fragments, for the two k-steps of one swizzled k-tile.
//   B_FROM_TILE: B[k][n] = tile row n (the first 32 rows of the tile)         (potrf)
//   else:        B[k][n] = Bk[(4*step + t4) * FLDB + n]  (k-major block in shared memory, rows already in k-step order)
template <bool B_FROM_TILE>
__device__ __forceinline__ void tile_mma(double (&acc)[2][4][4][2], const double* tile, const double* Bk, int o0, int o1,
                                         bool live0, bool live1, int warp, int lane) {
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
            for (int i = 0; i < 4; ++i) a[i] = tile[(warp * 32 + i * 8) * 8 + o];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[0][i][j], a[i], bf[j]);
        }
        if (live1) {
            double a[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = tile[((warp + FWARPS) * 32 + i * 8) * 8 + o];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[1][i][j], a[i], bf[j]);
        }
    }
}

struct TmaSmem {   // carve-up of the dynamic shared memory of both kernels
    int S, ring_rows, t_rows;
    size_t ring_bytes, t_bytes, blocks_bytes, total;
};
__host__ __device__ inline TmaSmem tma_smem_layout(int n_max, bool with_t, size_t budget) {
    TmaSmem L;
    const int below = n_max > NB ? n_max - NB : 0;          // rows of the largest streamed tile
    L.ring_rows = (below + TBOX - 1) / TBOX * TBOX;
    L.t_rows = with_t ? (below + 7) / 8 * 8 : 0;
    L.t_bytes = (size_t)L.t_rows * FLDP * sizeof(double);
    L.blocks_bytes = (size_t)3 * 32 * FLDP * sizeof(double);
    const size_t fixed = L.t_bytes + L.blocks_bytes + 2 * TMAX_STAGES * 8 + 16 /* progress counter */ + 1024 /* alignment slack */;
    const size_t stage = (size_t)L.ring_rows * 64;
    int S = stage == 0 || budget < fixed ? 0 : (int)((budget - fixed) / stage);
    L.S = S > TMAX_STAGES ? TMAX_STAGES : S;
    L.ring_bytes = (size_t)L.S * stage;
    L.total = L.ring_bytes + fixed;
    return L;
}

// ------------------------------------------------------------------------------------------------------------
constexpr int TTHREADS = FTHREADS + 128;   // eight MMA warps + one producer warpgroup (one of its warps works)
// Register split between the roles (setmaxnreg works on warpgroups of four warps): each scheduler hosts two MMA
// warps and one warp of the producer group, 2 * 232 + 40 registers per lane fit its 16K-entry file.
__device__ __forceinline__ void producer_regs() { asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n"); }
__device__ __forceinline__ void mma_regs() { asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n"); }

struct TmaCta {   // what both kernels set up identically
    TileRing ring; double* blocks; int* progress;
};
__device__ __forceinline__ TmaCta tma_cta_setup(unsigned char* smem_raw, const TmaSmem& lay, int tid) {
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    TmaCta c;
    c.blocks = reinterpret_cast<double*>(base + lay.ring_bytes);        // [T rows (trtri)] + three 32 x FLDP blocks
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + lay.ring_bytes + lay.t_bytes + lay.blocks_bytes);
    c.progress = reinterpret_cast<int*>(bars + 2 * TMAX_STAGES);
    c.ring.stage0 = reinterpret_cast<double*>(base); c.ring.S = lay.S; c.ring.stage_elems = lay.ring_rows * 8;
    c.ring.full0 = smem_u32(bars); c.ring.empty0 = smem_u32(bars + TMAX_STAGES);
    c.ring.stage = 0; c.ring.par = 0; c.ring.wrapped = false;
    if (tid == 0) {
        for (int s = 0; s < lay.S; ++s) { mbar_init(c.ring.full0 + 8 * s, 1); mbar_init(c.ring.empty0 + 8 * s, FWARPS); }
        *c.progress = 0;
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    return c;
}

__global__ void __launch_bounds__(TTHREADS, 1) k_potrf_tma(const __grid_constant__ CUtensorMap tmap, FactorArgs p, TmaSmem lay) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    __shared__ int sh_bad;
    __shared__ double scratch[256];
    TmaCta cta = tma_cta_setup(smem_raw, lay, tid);
    TileRing& ring = cta.ring;
    double* Dblk = cta.blocks;            // updated diagonal block
    double* Ld = Dblk + 32 * FLDP;
    double* Dinv = Ld + 32 * FLDP;

    if (warp >= FWARPS) {
        producer_regs();
        if (warp > FWARPS) return;
        // ---- producer warp: k-tile pkt (columns 8*pkt..) of the block column starting at pk0 needs the first
        //      pkt/4 + 1 block columns of this problem to be final; `done` counts finished block columns of the CTA
        int done_base = 0, seen = 0;
        for (int b = blockIdx.x; b < p.batch; b += gridDim.x) {
            const int n = p.n_sys ? p.n_sys[b] : p.n_max;
            for (int pk0 = NB; pk0 < n; pk0 += NB) {
                const int nboxes = (n - pk0 + TBOX - 1) / TBOX;
                for (int pkt = 0; pkt < pk0 / 8; ++pkt) {
                    seen = progress_wait(cta.progress, done_base + (pkt >> 2) + 1, seen);
                    uint32_t bar;
                    const uint32_t dst = ring.acquire_stage(bar, nboxes, lane);
                    if (lane == 0 && !(g_dbg & 2))
                        for (int q = 0; q < nboxes; ++q) tma_load_box(dst + q * TBOX_BYTES, &tmap, 8 * pkt, pk0 + q * TBOX, b, bar);
                    __syncwarp();
                }
            }
            done_base += (n + NB - 1) / NB;
        }
        return;
    }

    mma_regs();
    int o0, o1;
    swz_offsets(lane, o0, o1);
    int done = 0;
    for (int b = blockIdx.x; b < p.batch; b += gridDim.x) {
        const int n = p.n_sys ? p.n_sys[b] : p.n_max;
        double* A = p.A + (size_t)b * p.strideA;
        int bad_total = 0;
        long long t_phase = clock64();
        for (int k0 = 0; k0 < n; k0 += NB) {
            const int kb = min(NB, n - k0);
            const int m = n - k0;
            const int nch = (m + 31) >> 5;
            const bool live0 = warp < nch, live1 = warp + FWARPS < nch;
            double acc[2][4][4][2];
            zero_acc(acc[0]);
            zero_acc(acc[1]);
            const int nk = k0 / 8;
            for (int kt = 0; kt < nk; ++kt) {
                const double* tile = ring.wait_tile();
                if (!(g_dbg & 1)) tile_mma<true>(acc, tile, nullptr, o0, o1, live0, live1, warp, lane);
                ring.release_tile(lane);
            }
            PHASE_MARK(0);
            // P = A[k0:n, k0:k0+32] - acc, straight from global memory into the accumulator layout.  Full block
            // columns: the sixteen 16-byte loads of a chunk are issued back to back (row index clamped, result
            // masked) so that their latencies overlap; the ragged last block takes the guarded element-wise path.
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int ch = warp + c * FWARPS;
                if (ch < nch) {
                    if (kb == NB) {
                        double2 v[4][4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int gr = min(k0 + ch * 32 + i * 8 + g, n - 1);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                v[i][j] = *reinterpret_cast<const double2*>(A + (size_t)gr * p.lda + k0 + j * 8 + 2 * t4);
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const bool ok = k0 + ch * 32 + i * 8 + g < n;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                acc[c][i][j][0] = ok ? v[i][j].x - acc[c][i][j][0] : 0.0;
                                acc[c][i][j][1] = ok ? v[i][j].y - acc[c][i][j][1] : 0.0;
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int gr = k0 + ch * 32 + i * 8 + g;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int gc = k0 + j * 8 + 2 * t4;
                                double v0 = 0.0, v1 = 0.0;
                                if (gr < n && gc < n) v0 = A[(size_t)gr * p.lda + gc];
                                if (gr < n && gc + 1 < n) v1 = A[(size_t)gr * p.lda + gc + 1];
                                acc[c][i][j][0] = v0 - acc[c][i][j][0];
                                acc[c][i][j][1] = v1 - acc[c][i][j][1];
                            }
                        }
                    }
                }
            }
            if (warp == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        Dblk[(i * 8 + g) * FLDP + j * 8 + 2 * t4] = acc[0][i][j][0];
                        Dblk[(i * 8 + g) * FLDP + j * 8 + 2 * t4 + 1] = acc[0][i][j][1];
                    }
            }
            mma_warps_sync();
            PHASE_MARK(1);
            cta_potf2_b8<true>(Dblk, Ld, scratch, kb, tid, &sh_bad);
            PHASE_MARK(2);
            if (m > NB) cta_trtri32<true>(Ld, Dinv, scratch, tid);
            PHASE_MARK(3);
            if (sh_bad != 0 && bad_total == 0) bad_total = k0 + sh_bad;
            // rows below the diagonal block:  X = P * L_kk^{-T}, stored from registers   (only exists when kb == 32)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int ch = warp + c * FWARPS;
                if (ch >= 1 && ch < nch) {
                    chunk_mul_regs<true>(acc[c], Dinv, 1.0, lane);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int gr = k0 + ch * 32 + i * 8 + g;
                        if (gr < n) {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                *reinterpret_cast<double2*>(A + (size_t)gr * p.lda + k0 + j * 8 + 2 * t4) =
                                    make_double2(acc[c][i][j][0], acc[c][i][j][1]);
                        }
                    }
                }
            }
            PHASE_MARK(4);
            for (int idx = tid; idx < NB * NB; idx += FTHREADS) {
                const int r = idx >> 5, c = idx & 31;
                if (r < kb && c <= r) A[(size_t)(k0 + r) * p.lda + k0 + c] = Ld[r * FLDP + c];
            }
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
// Diagonal blocks are written back with an explicit zero upper triangle (the TMA tiles carry no triangular mask).
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
    double* Dsm = Tb + (size_t)lay.t_rows * FLDB;
    double* Dinv = Dsm + 32 * FLDP;

    if (warp >= FWARPS) {
        producer_regs();
        if (warp > FWARPS) return;
        // ---- producer warp: k-tile kt of step pjb lies in block column pjb + 1 + kt/4, which is final once the steps
        //      nblk-1 .. that column are done; `done` counts finished steps of the CTA
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
                    seen = progress_wait(cta.progress, done_base + nblk - (pjb + 1 + (kt >> 2)), seen);
                    const int q0 = (kt >> 2) >> 1;           // rows above 32*(kt/4) never need this k-tile
                    uint32_t bar;
                    const uint32_t dst = ring.acquire_stage(bar, nboxes - q0, lane);
                    if (lane == 0 && !(g_dbg & 2))
                        for (int q = q0; q < nboxes; ++q) tma_load_box(dst + q * TBOX_BYTES, &tmap, r0 + 8 * kt, r0 + q * TBOX, b, bar);
                    __syncwarp();
                }
            }
            done_base += nblk;
        }
        return;
    }

    mma_regs();
    int o0, o1;
    swz_offsets(lane, o0, o1);
    int done = 0;
    for (int b = blockIdx.x; b < p.batch; b += gridDim.x) {
        const int n = p.n_sys ? p.n_sys[b] : p.n_max;
        double* L = p.A + (size_t)b * p.strideA;
        const int nblk = (n + NB - 1) / NB;
        long long t_phase = clock64();
        // ragged batch: the columns n .. n_max-1 of this problem are streamed with the last k-tile; make them finite
        if (n < p.n_max) {
            const int cw = min(p.n_max, (n + 7) / 8 * 8) - n;
            for (int idx = tid; idx < n * cw; idx += FTHREADS) L[(size_t)(idx / cw) * p.lda + n + idx % cw] = 0.0;
            fence_global_to_tma();
        }
        for (int jb = nblk - 1; jb >= 0; --jb) {
            const int k0 = jb * NB;
            const int kb = min(NB, n - k0);
            const int m = n - k0 - kb;            // rows below the diagonal block
            for (int idx = tid; idx < NB * NB; idx += FTHREADS) {
                const int r = idx >> 5, c = idx & 31;
                Dsm[r * FLDP + c] = (r < kb && c <= r) ? L[(size_t)(k0 + r) * p.lda + k0 + c] : ((r == c) ? 1.0 : 0.0);
            }
            mma_warps_sync();
            cta_trtri32<true>(Dsm, Dinv, scratch, tid);
            PHASE_MARK(8);
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
                mma_warps_sync();
                PHASE_MARK(9);
                zero_acc(acc[0]);
                zero_acc(acc[1]);
                for (int pos = 0; pos < nk; ++pos) {
                    const int kt = trtri_tile_of(pos, nk);
                    const double* tile = ring.wait_tile();
                    // a 32-row chunk ch only needs k < 32*(ch+1)
                    if (!(g_dbg & 1))
                        tile_mma<false>(acc, tile, Tb + (size_t)kt * 8 * FLDB, o0, o1, live0 && kt < 4 * (warp + 1),
                                        live1 && kt < 4 * (warp + FWARPS + 1), warp, lane);
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
            for (int idx = tid; idx < NB * NB; idx += FTHREADS) {
                const int r = idx >> 5, c = idx & 31;
                if (r < kb && c < kb) L[(size_t)(k0 + r) * p.lda + k0 + c] = c <= r ? Dinv[r * FLDP + c] : 0.0;
            }
            fence_global_to_tma();
            mma_warps_sync();   // also: every warp is done with T before the next step rewrites it
            ++done;
            if (tid == 0) progress_publish(cta.progress, done);
            PHASE_MARK(11);
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TensorMapEncodeFn tensor_map_encoder() {
    static TensorMapEncodeFn fn = [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            ptr = nullptr;
        return reinterpret_cast<TensorMapEncodeFn>(ptr);
    }();
    return fn;
}

// 3-D map over [problem][row][column] of a batch of row-major n_max x n_max matrices, boxes of 64 rows x 8 columns.
inline int make_batch_map(CUtensorMap* map, const double* A, int lda, long long strideA, int n_max, int batch) {
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (!enc) { snprintf(g_err, sizeof(g_err), "kmpx: cuTensorMapEncodeTiled is not available from the driver"); return (int)cudaErrorNotSupported; }
    const cuuint64_t dims[3] = {(cuuint64_t)n_max, (cuuint64_t)n_max, (cuuint64_t)batch};
    const cuuint64_t strides[2] = {(cuuint64_t)lda * 8, (cuuint64_t)(batch > 1 ? strideA : (long long)lda * n_max) * 8};
    const cuuint32_t box[3] = {8, TBOX, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(A), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_err, sizeof(g_err), "kmpx: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return (int)cudaErrorInvalidValue; }
    return 0;
}

// The TMA kernels need 16-byte aligned rows and problems, and at least three ring stages next to the fixed buffers.
inline bool tma_applicable(const double* A, int lda, long long strideA, int n_max, int batch, bool with_t) {
    if (g_host_dbg & 16) return false;
    if (n_max < 2 * NB || (lda & 1) || (batch > 1 && (strideA & 1)) || (reinterpret_cast<uintptr_t>(A) & 15)) return false;
    return tma_smem_layout(n_max, with_t, kMaxDynSmem).S >= 3;
}

}  // namespace kmpx
