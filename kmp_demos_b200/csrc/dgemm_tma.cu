// TMA-fed persistent fp64 GEMM for the NT shape  C = alpha * A * B^T + beta * C  (A [m][k], B [n][k], row-major): the
// trailing update / panel product of the blocked large-n Cholesky (factor_large.cu) and V = Linv * kappa of the cfg-4
// predict (predict_gemm.cu).  Same 128 x 128 x 32 CTA tile and 64 x 32 warp tile as k_dgemm (dgemm.cu); what changes:
//   * operands arrive through the tensor memory accelerator: one 3-D tensor map per operand over [batch][row][k],
//     boxes of 128 rows x 8 k (64-byte rows, 64-byte swizzle - the layout of the small-system kernels, tma_util.cuh),
//     eight boxes per stage, completion on an mbarrier; a producer warp runs the ring, the eight MMA warps never meet a
//     CTA-wide barrier and carry no address arithmetic for the loads.  Out-of-range rows / k are zero-filled by the unit;
//   * CTAs are persistent (one per SM) and walk a static list of output tiles, so the ring keeps streaming the next
//     tile's operands while the warps are in the epilogue of the current one; the C tile of a beta != 0 update is
//     prefetched into L2 at the start of its main loop (at K = 256 the read-modify-write of C is a fifth of the tile's
//     DMMA time when it waits on HBM);
//   * tri = 1 (lower tiles only) enumerates just the tiles on or below the diagonal; tri = 2 (A lower triangular) cuts
//     the k range at the diagonal tile.
// A DMMA k-step uses the k columns {0,1,4,5} / {2,3,6,7} of an 8-column box for BOTH operands (conflict-free reads of the
// swizzled rows; the contraction is a sum over k, so a common permutation changes nothing).
#include "dmma_tile.cuh"
#include "kmpx_internal.cuh"
#include "tma_util.cuh"

namespace kmpx {

constexpr int HBM_ = 128, HBN = 128, HBK = 32, HSTAGES = 3;
constexpr int HWARPS = 8, HTHREADS = HWARPS * 32 + 128;        // + producer warpgroup (setmaxnreg works on four warps)
constexpr int HBOX_BYTES = 128 * 64;                            // 128 rows x 8 doubles
constexpr int HOPER_BYTES = (HBK / 8) * HBOX_BYTES;             // one operand of one stage: 32 KB
constexpr int HSTAGE_BYTES = 2 * HOPER_BYTES;
constexpr size_t HSMEM = (size_t)HSTAGES * HSTAGE_BYTES + 2 * HSTAGES * 8 + 1024;

struct TmaGemmArgs {
    int m, n, k; double alpha, beta; double* C; int ldc; long long strideC; int tri;
    int tiles_m, tiles_n, tiles_per_mat, tiles_total;
};

// linear tile index -> (tile row, tile column); tri = 1: only tiles with tn <= tm, row by row
__device__ __forceinline__ void tile_coords(const TmaGemmArgs& p, int t, int& tm, int& tn) {
    if (p.tri != 1) { tm = t / p.tiles_n; tn = t - tm * p.tiles_n; return; }
    // tiles of row r: min(r + 1, tiles_n); rows 0 .. tiles_n - 1 form the triangle, the rest are full rows
    const int tri_rows = min(p.tiles_m, p.tiles_n);
    const int tri_count = tri_rows * (tri_rows + 1) / 2;
    if (t < tri_count) {
        int r = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
        while (r * (r + 1) / 2 > t) --r;
        while ((r + 1) * (r + 2) / 2 <= t) ++r;
        tm = r; tn = t - r * (r + 1) / 2;
    } else {
        const int u = t - tri_count;
        tm = tri_rows + u / p.tiles_n; tn = u - (u / p.tiles_n) * p.tiles_n;
    }
}
__device__ __forceinline__ int tile_k_tiles(const TmaGemmArgs& p, int tm, int tn) {
    // tri = 2: A lower triangular (k <= row);  tri = 4: B lower triangular (B[n][k] = 0 for k > n, the transposed
    // inverse of a Cholesky block): the k range ends at the diagonal tile of the triangular operand
    const int kend = p.tri == 2 ? min(p.k, (tm + 1) * HBM_) : (p.tri == 4 ? min(p.k, (tn + 1) * HBN) : p.k);
    return (kend + HBK - 1) / HBK;
}

__global__ void __launch_bounds__(HTHREADS, 1) k_dgemm_nt_tma(const __grid_constant__ CUtensorMap mapA,
                                                             const __grid_constant__ CUtensorMap mapB, TmaGemmArgs p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = smem_align1k(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)HSTAGES * HSTAGE_BYTES);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + HSTAGES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < HSTAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, HWARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (warp >= HWARPS) {
        // ---------------------------------------------------------------- producer
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
        if (warp != HWARPS || lane != 0) return;
        int stage = 0; uint32_t par = 0; bool wrapped = false;
        for (int t = blockIdx.x; t < p.tiles_total; t += gridDim.x) {
            const int z = t / p.tiles_per_mat;
            int tm, tn;
            tile_coords(p, t - z * p.tiles_per_mat, tm, tn);
            const int nk = tile_k_tiles(p, tm, tn);
            for (int kt = 0; kt < nk; ++kt) {
                if (wrapped) mbar_wait(empty0 + 8 * stage, par ^ 1);
                const uint32_t bar = full0 + 8 * stage;
                mbar_expect_tx(bar, HSTAGE_BYTES);
                const uint32_t dst = smem_u32(base + (size_t)stage * HSTAGE_BYTES);
#pragma unroll
                for (int q = 0; q < HBK / 8; ++q) {
                    tma_load_box(dst + q * HBOX_BYTES, &mapA, kt * HBK + 8 * q, tm * HBM_, z, bar);
                    tma_load_box(dst + HOPER_BYTES + q * HBOX_BYTES, &mapB, kt * HBK + 8 * q, tn * HBN, z, bar);
                }
                if (++stage == HSTAGES) { stage = 0; par ^= 1; wrapped = true; }
            }
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;\n");

    // ---------------------------------------------------------------- MMA warps
    const int g = lane >> 2, t4 = lane & 3;
    const int wr = warp >> 2, wc = warp & 3;
    const int sw = (g >> 1) & 3, c0 = (t4 >> 1) << 1;
    const int o0 = g * 8 + ((c0 ^ sw) << 1) + (t4 & 1);              // k-step 0 of a box: columns {0,1,4,5}
    const int o1 = g * 8 + (((c0 | 1) ^ sw) << 1) + (t4 & 1);        // k-step 1: columns {2,3,6,7}
    const bool vec_ok = ((p.ldc & 1) == 0) && ((p.strideC & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    int stage = 0; uint32_t par = 0;
    for (int t = blockIdx.x; t < p.tiles_total; t += gridDim.x) {
        const int z = t / p.tiles_per_mat;
        int tm, tn;
        tile_coords(p, t - z * p.tiles_per_mat, tm, tn);
        const int m0 = tm * HBM_, n0 = tn * HBN;
        const int nk = tile_k_tiles(p, tm, tn);
        double* C = p.C + (size_t)z * p.strideC;
        const bool interior = vec_ok && m0 + HBM_ <= p.m && n0 + HBN <= p.n;
        if (p.beta != 0.0 && interior) {
            // the C tile of this update into L2 while the main loop runs: 128 rows x 1 KB, every thread touches four rows' lines
            const char* cb = reinterpret_cast<const char*>(C + (size_t)m0 * p.ldc + n0);
            for (int idx = tid; idx < HBM_ * 8; idx += HWARPS * 32) {
                const int r = idx >> 3, q = idx & 7;
                asm volatile("prefetch.global.L2 [%0];\n" :: "l"(cb + (size_t)r * p.ldc * 8 + q * 128));
            }
        }
        double acc[8][4][2];
        zero_acc(acc);
        for (int kt = 0; kt < nk; ++kt) {
            mbar_wait(full0 + 8 * stage, par);
            const double* sa = reinterpret_cast<const double*>(base + (size_t)stage * HSTAGE_BYTES);
            const double* sb = sa + HOPER_BYTES / 8;
#pragma unroll
            for (int st = 0; st < HBK / 4; ++st) {
                const int q = st >> 1;
                const int o = (st & 1) ? o1 : o0;
                double a[8], bf[4];
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = sa[q * (128 * 8) + (wr * 64 + i * 8) * 8 + o];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = sb[q * (128 * 8) + (wc * 32 + j * 8) * 8 + o];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j], a[i], bf[j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * stage);
            if (++stage == HSTAGES) { stage = 0; par ^= 1; }
        }
        // ---- epilogue
        if (interior) {
            double* cbase = C + (size_t)(m0 + wr * 64 + g) * p.ldc + n0 + wc * 32 + 2 * t4;
            if (p.beta != 0.0) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {                     // two halves of 16 loads in flight: bounded register use
                    double2 o[4][4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) o[i][j] = *reinterpret_cast<const double2*>(cbase + (size_t)(h * 4 + i) * 8 * p.ldc + j * 8);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            *reinterpret_cast<double2*>(cbase + (size_t)(h * 4 + i) * 8 * p.ldc + j * 8) =
                                make_double2(fma(p.beta, o[i][j].x, p.alpha * acc[h * 4 + i][j][0]),
                                             fma(p.beta, o[i][j].y, p.alpha * acc[h * 4 + i][j][1]));
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<double2*>(cbase + (size_t)i * 8 * p.ldc + j * 8) =
                            make_double2(p.alpha * acc[i][j][0], p.alpha * acc[i][j][1]);
            }
            continue;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = m0 + wr * 64 + i * 8 + g;
            if (r >= p.m) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = n0 + wc * 32 + j * 8 + 2 * t4;
                if (c >= p.n) continue;
                double* dst = C + (size_t)r * p.ldc + c;
                double v0 = p.alpha * acc[i][j][0], v1 = p.alpha * acc[i][j][1];
                if (c + 1 < p.n && vec_ok) {
                    if (p.beta != 0.0) { const double2 o = *reinterpret_cast<const double2*>(dst); v0 += p.beta * o.x; v1 += p.beta * o.y; }
                    *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
                } else {
                    if (p.beta != 0.0) v0 += p.beta * dst[0];
                    dst[0] = v0;
                    if (c + 1 < p.n) { if (p.beta != 0.0) v1 += p.beta * dst[1]; dst[1] = v1; }
                }
            }
        }
    }
}

// 3-D map over [batch][row][k] of a row-major (rows x cols) operand, boxes of 128 rows x 8 columns, 64-byte swizzle
static int make_operand_map(CUtensorMap* map, const double* A, int ld, long long stride, int rows, int cols, int batch) {
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (!enc) { snprintf(g_err, sizeof(g_err), "kmpx: cuTensorMapEncodeTiled is not available from the driver"); return (int)cudaErrorNotSupported; }
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * 8, (cuuint64_t)(batch > 1 ? stride : (long long)ld * rows) * 8};
    const cuuint32_t box[3] = {8, 128, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(A), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_err, sizeof(g_err), "kmpx: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return (int)cudaErrorInvalidValue; }
    return 0;
}

// The NT product through the TMA kernel when the operands qualify (16-byte aligned rows and batches, tri in {0,1,2});
// returns 1 when the caller has to take the cp.async kernel instead.  tri = 4 (B lower triangular) exists only here: the
// cp.async kernel treats it as a dense product (the skipped k-tiles hold zeros).
int launch_dgemm_nt_tma(const GemmArgs& g, int batch, cudaStream_t st) {
    if (g_host_dbg & 128) return 1;
    if (g.tri == 3 || g.k < 8 || g.m < 8 || g.n < 8) return 1;
    if ((g.lda & 1) || (g.ldb & 1) || (g.strideA & 1) || (g.strideB & 1)) return 1;
    if ((reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.B) & 15)) return 1;
    if (batch > 1 && (g.strideA == 0 || g.strideB == 0)) return 1;           // a tensor map needs a real batch stride
    if ((long long)g.lda * 8 >= (1LL << 40) || (long long)g.ldb * 8 >= (1LL << 40)) return 1;
    CUtensorMap mapA, mapB;
    KMPX_TRY(make_operand_map(&mapA, g.A, g.lda, g.strideA, g.m, g.k, batch));
    KMPX_TRY(make_operand_map(&mapB, g.B, g.ldb, g.strideB, g.n, g.k, batch));
    TmaGemmArgs p{g.m, g.n, g.k, g.alpha, g.beta, g.C, g.ldc, g.strideC, g.tri, 0, 0, 0, 0};
    p.tiles_m = ceil_div(g.m, HBM_); p.tiles_n = ceil_div(g.n, HBN);
    if (g.tri == 1) {
        const int tr = p.tiles_m < p.tiles_n ? p.tiles_m : p.tiles_n;
        p.tiles_per_mat = tr * (tr + 1) / 2 + (p.tiles_m - tr) * p.tiles_n;
    } else p.tiles_per_mat = p.tiles_m * p.tiles_n;
    const long long total = (long long)p.tiles_per_mat * batch;
    if (total > 0x7fffffffLL) return 1;
    p.tiles_total = (int)total;
    static SmemCache cache;
    KMPX_TRY(ensure_dyn_smem(k_dgemm_nt_tma, HSMEM, cache, "k_dgemm_nt_tma"));
    int sms = num_sms() - (g.reserve_sms > 0 ? g.reserve_sms : 0);
    if (sms < 1) sms = 1;
    const int grid = p.tiles_total < sms ? p.tiles_total : sms;
    k_dgemm_nt_tma<<<grid, HTHREADS, HSMEM, st>>>(mapA, mapB, p);
    return check_launch("k_dgemm_nt_tma");
}

}  // namespace kmpx
