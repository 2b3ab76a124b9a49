// TMA / mbarrier building blocks shared by the kernels that stream matrix tiles through the tensor memory accelerator.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace kmpx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// first 1 KB aligned address inside a shared-memory buffer; plain pointer arithmetic on the __shared__ array, so that
// the compiler keeps the address space (a round trip through uintptr_t turns every access into a generic LD/ST)
__device__ __forceinline__ unsigned char* smem_align1k(unsigned char* p) { return p + ((1024u - (smem_u32(p) & 1023u)) & 1023u); }
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

// non-blocking probe of an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
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

// 3-D map over [problem][row][column] of a batch of row-major n_max x n_max matrices, boxes of box_rows rows x 8 columns
// (64 bytes per row, 64-byte swizzle: the 16-byte chunk q of box row r lands at chunk q ^ ((r >> 1) & 3)).
inline int make_batch_map(CUtensorMap* map, const double* A, int lda, long long strideA, int n_max, int batch, int box_rows) {
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (!enc) { snprintf(g_err, sizeof(g_err), "kmpx: cuTensorMapEncodeTiled is not available from the driver"); return (int)cudaErrorNotSupported; }
    const cuuint64_t dims[3] = {(cuuint64_t)n_max, (cuuint64_t)n_max, (cuuint64_t)batch};
    const cuuint64_t strides[2] = {(cuuint64_t)lda * 8, (cuuint64_t)(batch > 1 ? strideA : (long long)lda * n_max) * 8};
    const cuuint32_t box[3] = {8, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(A), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_err, sizeof(g_err), "kmpx: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return (int)cudaErrorInvalidValue; }
    return 0;
}


}  // namespace kmpx
