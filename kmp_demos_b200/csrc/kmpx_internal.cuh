// Argument structs and launchers shared between the translation units of libkmpx.
#pragma once
#include "common.cuh"

namespace kmpx {

// Component-major system order: row = a*Ns + i with Ns = N rounded up to even, so that every component
// block starts on a 16-byte boundary (cp.async granularity).  Padding rows/columns (i >= N) are identity.
__host__ __device__ inline int comp_stride(int N) { return (N + 1) & ~1; }
__host__ __device__ inline int sys_size(int N, int O, int layout) {
    return layout == KMPX_LAYOUT_POINT ? N * O : comp_stride(N) * O;
}

struct FactorArgs {
    double* A; int lda; long long strideA; const int* n_sys; int n_max; int batch; int* info;
    int info_offset;       // added to a non-zero local pivot index (blocked large-n driver)
    int info_accumulate;   // keep an already recorded failure instead of overwriting it
};
struct GemmArgs {
    int m, n, k; double alpha; const double* A; int lda; long long strideA;
    const double* B; int ldb; long long strideB; double beta; double* C; int ldc; long long strideC; int tri;
    int reserve_sms;           // persistent kernels leave this many SMs to concurrent streams (0 when omitted)
};
struct AlphaArgs {
    const double* F; int lda; long long strideA; const int* n_pts; int n_max, O, batch, layout;
    const double* y; double* w; int ldw; double* alpha; int general;
};

struct PredictArgs {
    const double* F; int lda; long long strideA; int general;
    const double* w; int ldw; const double* s; const int* n_pts; int n_max, I, O, batch;
    const double* sq; long long stride_sq; int M;
    double sigma_f, alpha_kmp; int time_driven;
    double* xi; double* sigma_out;
    int resident; int ldp;   // panel leading dimension (resident mode)
    int dbg;                 // profiling experiments: bit 0 skip MMAs, bit 1 skip loads
    double kskip;            // kernel coefficients below this for a whole query tile are treated as zero (0: keep everything)
    double kskip_d2;         // -ln(kskip) / sigma_f: the squared distance beyond which k < kskip (set by the launcher)
};
bool predict_fast_applicable(const PredictArgs& p);
int launch_predict_fast(PredictArgs& p, cudaStream_t st);
extern double g_predict_kskip;   // kmpx_predict_skip_threshold
extern int g_predict_tile_queries;   // kmpx_predict_tile_queries
bool predict_tab_applicable(const PredictArgs& p);
int launch_predict_tab(PredictArgs& p, cudaStream_t st);
bool predict_pers_applicable(const PredictArgs& p);
int launch_predict_pers(PredictArgs& p, cudaStream_t st);
bool predict_stream_applicable(const PredictArgs& p);
int launch_predict_stream(PredictArgs& p, cudaStream_t st);

int launch_dgemm(int transA, int transB, const GemmArgs& p, int batch, cudaStream_t st);
int launch_dgemm_nt_tma(const GemmArgs& p, int batch, cudaStream_t st);   // 0 launched, 1 not applicable, else error
int launch_potrf_small(double* A, int lda, long long strideA, const int* n_sys, int n_max, int batch, int* info,
                       int info_offset, int info_accumulate, cudaStream_t st);
int launch_trtri_small(double* L, int lda, long long strideA, const int* n_sys, int n_max, int batch, cudaStream_t st);
int launch_solve_alpha(const AlphaArgs& p, cudaStream_t st);

}  // namespace kmpx
