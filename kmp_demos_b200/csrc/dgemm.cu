// Dense fp64 GEMM on DMMA: C = alpha * op(A) * op(B) + beta * C, row-major, strided batch.
// The building block of the blocked large-n Cholesky / triangular inverse (factor_large.cu).
//
// CTA tile 128 x 128 x 32, 256 threads = 8 warps as 2 (rows) x 4 (cols), warp tile 64 x 32
// (64 accumulators per thread), 3-stage cp.async ring.  Per k-step of 4 a warp issues 12 LDS.64 for 32
// DMMA.8x8x4, so the FP64 tensor pipe (16 cycles per DMMA per SM sub-partition) is the limiter.
//
// Operand tiles in shared memory keep the global orientation (no transposition on the way in):
//   op(A)=A   : global A[m][k]  -> smem [128][20]        op(A)=A^T : global A[k][m] -> smem [16][132]
//   op(B)=B^T : global B[n][k]  -> smem [128][20]        op(B)=B   : global B[k][n] -> smem [16][132]
// Leading dimensions 20 and 132 are 4 (mod 16): fragment reads are bank-conflict free.
#include "dmma_tile.cuh"
#include "kmpx_internal.cuh"

namespace kmpx {

constexpr int GBM = 128, GBN = 128, GBK = 32, GSTAGES = 3, GTHREADS = 256;
constexpr int GLDK = GBK + 4;      // [rows][k] tiles
constexpr int GLDM = GBM + 4;      // [k][rows] tiles
constexpr int GTILE = (GBM * GLDK > GBK * GLDM) ? GBM * GLDK : GBK * GLDM;   // doubles per operand tile

// Copies a (rows x cols) window of a row-major global matrix to smem [rows][LD]; cols is even, the window
// starts at an even column; elements outside (rows_valid, cols_valid) are zero-filled.
template <int LD>
__device__ __forceinline__ void load_window_async(double* dst, const double* g, int ld, int rows, int cols,
                                                  int rows_valid, int cols_valid, int tid) {
    const int ch = cols >> 1;
    for (int idx = tid; idx < rows * ch; idx += GTHREADS) {
        const int r = idx / ch, q = idx - r * ch;
        int nv = 0;
        if (r < rows_valid) { nv = cols_valid - 2 * q; nv = nv < 0 ? 0 : (nv > 2 ? 2 : nv); }
        cp_async16(dst + r * LD + 2 * q, nv > 0 ? g + (size_t)r * ld + 2 * q : g, nv * 8);
    }
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(GTHREADS) k_dgemm(GemmArgs p) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int wr = warp >> 2, wc = warp & 3;
    const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
    if (p.tri == 1 && n0 > m0 + GBM - 1) return;           // strictly-upper tile of a lower-only update
    const double* A = p.A + (size_t)blockIdx.z * p.strideA;
    const double* B = p.B + (size_t)blockIdx.z * p.strideB;
    double* C = p.C + (size_t)blockIdx.z * p.strideC;

    int kbeg = 0, kend = p.k;
    if (p.tri == 2) kend = min(p.k, m0 + GBM);               // op(A) lower-triangular: k <= row
    if (p.tri == 3) kbeg = (n0 / GBK) * GBK;                 // op(B) lower-triangular: k >= col
    const int nk = (kend - kbeg + GBK - 1) / GBK;

    auto issue = [&](int kt) {
        double* sa = smem + (size_t)(kt % GSTAGES) * 2 * GTILE;
        double* sb = sa + GTILE;
        const int k0 = kbeg + kt * GBK;
        if (!TA) load_window_async<GLDK>(sa, A + (size_t)m0 * p.lda + k0, p.lda, GBM, GBK, p.m - m0, kend - k0, tid);
        else     load_window_async<GLDM>(sa, A + (size_t)k0 * p.lda + m0, p.lda, GBK, GBM, kend - k0, p.m - m0, tid);
        if (TB)  load_window_async<GLDK>(sb, B + (size_t)n0 * p.ldb + k0, p.ldb, GBN, GBK, p.n - n0, kend - k0, tid);
        else     load_window_async<GLDM>(sb, B + (size_t)k0 * p.ldb + n0, p.ldb, GBK, GBN, kend - k0, p.n - n0, tid);
    };

    double acc[8][4][2];
    zero_acc(acc);
#pragma unroll
    for (int s = 0; s < GSTAGES - 1; ++s) { if (s < nk) issue(s); cp_async_commit(); }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<GSTAGES - 2>();
        __syncthreads();
        if (kt + GSTAGES - 1 < nk) issue(kt + GSTAGES - 1);
        cp_async_commit();
        const double* sa = smem + (size_t)(kt % GSTAGES) * 2 * GTILE;
        const double* sb = sa + GTILE;
#pragma unroll
        for (int kk = 0; kk < GBK; kk += 4) {
            double a[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                a[i] = TA ? sa[(kk + t4) * GLDM + wr * 64 + i * 8 + g] : sa[(wr * 64 + i * 8 + g) * GLDK + kk + t4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                bf[j] = TB ? sb[(wc * 32 + j * 8 + g) * GLDK + kk + t4] : sb[(kk + t4) * GLDM + wc * 32 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j], a[i], bf[j]);
        }
    }
    cp_async_wait<0>();
    const bool vec_ok = ((p.ldc & 1) == 0) && ((p.strideC & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    // Interior tile with beta != 0: all thirty-two 16-byte reads of C are issued before the first use, so their (HBM)
    // latency is paid once.  Left to the guarded loop below the reads are consumed one by one and the read-modify-write
    // costs a quarter of a K = 256 tile (measured 20.9 vs 28.0 TFLOP/s for beta = 1 vs 0).
    if (vec_ok && p.beta != 0.0 && m0 + GBM <= p.m && n0 + GBN <= p.n) {
        double* cbase = C + (size_t)(m0 + wr * 64 + g) * p.ldc + n0 + wc * 32 + 2 * t4;
        double2 o[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) o[i][j] = *reinterpret_cast<const double2*>(cbase + (size_t)i * 8 * p.ldc + j * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<double2*>(cbase + (size_t)i * 8 * p.ldc + j * 8) =
                    make_double2(fma(p.beta, o[i][j].x, p.alpha * acc[i][j][0]), fma(p.beta, o[i][j].y, p.alpha * acc[i][j][1]));
        return;
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

int launch_dgemm(int transA, int transB, const GemmArgs& p, int batch, cudaStream_t st) {
    if (p.m <= 0 || p.n <= 0) return 0;
    if (!transA && transB) {                       // NT: the TMA-fed persistent kernel when the operands qualify
        const int rc = launch_dgemm_nt_tma(p, batch, st);
        if (rc != 1) return rc;
    }
    constexpr size_t smem = (size_t)GSTAGES * 2 * GTILE * sizeof(double);
    static bool attr_done[kMaxDevices] = {};          // the attribute is per device
    const int dev = current_device();
    if (!attr_done[dev]) {
        KMPX_TRY(check_cuda(cudaFuncSetAttribute(k_dgemm<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "attr dgemm NN"));
        KMPX_TRY(check_cuda(cudaFuncSetAttribute(k_dgemm<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "attr dgemm NT"));
        KMPX_TRY(check_cuda(cudaFuncSetAttribute(k_dgemm<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "attr dgemm TN"));
        KMPX_TRY(check_cuda(cudaFuncSetAttribute(k_dgemm<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "attr dgemm TT"));
        attr_done[dev] = true;
    }
    dim3 grid(ceil_div(p.n, GBN), ceil_div(p.m, GBM), batch);
    if (!transA && !transB) k_dgemm<false, false><<<grid, GTHREADS, smem, st>>>(p);
    else if (!transA && transB) k_dgemm<false, true><<<grid, GTHREADS, smem, st>>>(p);
    else if (transA && !transB) k_dgemm<true, false><<<grid, GTHREADS, smem, st>>>(p);
    else k_dgemm<true, true><<<grid, GTHREADS, smem, st>>>(p);
    return check_launch("k_dgemm");
}

}  // namespace kmpx

using namespace kmpx;

extern "C" int kmpx_dgemm(int transA, int transB, int m, int n, int k, double alpha, const double* A, int lda,
                          long long strideA, const double* B, int ldb, long long strideB, double beta, double* C,
                          int ldc, long long strideC, int batch, int tri, void* stream) {
    if (m < 0) return fail_arg(3, "m < 0");
    if (n < 0) return fail_arg(4, "n < 0");
    if (k < 0) return fail_arg(5, "k < 0");
    if (!A) return fail_arg(7, "A is NULL");
    if (lda & 1) return fail_arg(8, "lda must be even");
    if (!B) return fail_arg(10, "B is NULL");
    if (ldb & 1) return fail_arg(11, "ldb must be even");
    if (!C) return fail_arg(14, "C is NULL");
    if ((strideA & 1) || (strideB & 1)) return fail_arg(9, "batch strides must be even");
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return fail_arg(7, "A and B must be 16-byte aligned");
    if (batch <= 0 || batch > 65535) return fail_arg(17, "batch must be in 1..65535");
    if (tri < 0 || tri > 4) return fail_arg(18, "tri must be in 0..4");
    if (tri == 4 && !(transB && !transA)) return fail_arg(18, "tri=4 needs the NT form");
    if (tri == 2 && transA) return fail_arg(18, "tri=2 needs op(A)=A");
    if (tri == 3 && transB) return fail_arg(18, "tri=3 needs op(B)=B");
    GemmArgs p{m, n, k, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC, tri, 0};
    return launch_dgemm(transA, transB, p, batch, (cudaStream_t)stream);
}
