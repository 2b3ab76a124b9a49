// K9: batched quaternion maps (kmp/types/quaternion.py).  Quaternions are scalar-first [v, u0, u1, u2];
// every result goes through the reference constructor's normalisation (quaternion.py:18-21,121-129).
// Elementwise and HBM-bound: 56 bytes per element for exp/log, 64-96 for the product.
#include "common.cuh"

namespace kmpx {

struct Q4 { double v, x, y, z; };

__device__ __forceinline__ Q4 q_normalize(Q4 q) {
    // np.linalg.norm of the 4-vector: sqrt of the left-to-right sum of squares
    const double n = sqrt(q.v * q.v + q.x * q.x + q.y * q.y + q.z * q.z);
    return Q4{q.v / n, q.x / n, q.y / n, q.z / n};
}
__device__ __forceinline__ Q4 q_load(const double* p) {
    const double2 a = *reinterpret_cast<const double2*>(p);
    const double2 b = *reinterpret_cast<const double2*>(p + 2);
    return Q4{a.x, a.y, b.x, b.y};
}
__device__ __forceinline__ void q_store(double* p, Q4 q) {
    *reinterpret_cast<double2*>(p) = make_double2(q.v, q.x);
    *reinterpret_cast<double2*>(p + 2) = make_double2(q.y, q.z);
}
// np.allclose(a, 0): |a_i| <= atol (1e-8) for every component (rtol*|0| = 0); NaN is never close
__device__ __forceinline__ bool allclose0(double a, double b, double c) {
    return fabs(a) <= 1e-8 && fabs(b) <= 1e-8 && fabs(c) <= 1e-8;
}
__device__ __forceinline__ double sign_of(double v) { return v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : (v == 0.0 ? 0.0 : v)); }

__global__ void k_quat_normalize(const double* q, double* out, long long M) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < M) q_store(out + 4 * i, q_normalize(q_load(q + 4 * i)));
}

// quaternion.exp (quaternion.py:60-80)
__global__ void k_quat_exp(const double* w, double* q, long long M) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= M) return;
    const double a = w[3 * i], b = w[3 * i + 1], c = w[3 * i + 2];
    Q4 r{1.0, 0.0, 0.0, 0.0};
    if (!allclose0(a, b, c)) {
        const double n = sqrt(a * a + b * b + c * c);
        const double sn = sin(n);
        r = Q4{cos(n), sn * a / n, sn * b / n, sn * c / n};
    }
    q_store(q + 4 * i, q_normalize(r));
}

// quaternion.log (quaternion.py:82-96)
__global__ void k_quat_log(const double* q, double* w, long long M) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= M) return;
    const Q4 t = q_load(q + 4 * i);
    const double n = sqrt(t.v * t.v + t.x * t.x + t.y * t.y + t.z * t.z);
    const double sg = sign_of(t.v);
    const double ux = sg * t.x / n, uy = sg * t.y / n, uz = sg * t.z / n, v = sg * t.v / n;
    double o0 = 0.0, o1 = 0.0, o2 = 0.0;
    if (!allclose0(ux, uy, uz)) {
        const double un = sqrt(ux * ux + uy * uy + uz * uz);
        const double ac = acos(v);
        o0 = ac * ux / un; o1 = ac * uy / un; o2 = ac * uz / un;
    }
    w[3 * i] = o0; w[3 * i + 1] = o1; w[3 * i + 2] = o2;
}

// Hamilton product p*q (quaternion.py:182-202), optionally with q conjugated first (~q, :219-227, which
// re-normalises q before the product exactly as the reference's temporary object does)
__global__ void k_quat_mul(const double* p, const double* q, int q_bcast, int conj_q, double* out, long long M) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= M) return;
    const Q4 a = q_load(p + 4 * i);
    Q4 b = q_load(q + (q_bcast ? 0 : 4 * i));
    if (conj_q) b = q_normalize(Q4{b.v, -b.x, -b.y, -b.z});
    Q4 r;
    r.x = a.v * b.x + b.v * a.x + (a.y * b.z - a.z * b.y);
    r.y = a.v * b.y + b.v * a.y + (a.z * b.x - a.x * b.z);
    r.z = a.v * b.z + b.v * a.z + (a.x * b.y - a.y * b.x);
    r.v = a.v * b.v - (a.x * b.x + a.y * b.y + a.z * b.z);
    q_store(out + 4 * i, q_normalize(r));
}

// quaternion.from_rotation_vector incl. the abs(cos) quirk (quaternion.py:23-42); a zero vector gives NaNs
__global__ void k_quat_from_rotvec(const double* r, double* q, long long M) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= M) return;
    const double a = r[3 * i], b = r[3 * i + 1], c = r[3 * i + 2];
    const double ang = sqrt(a * a + b * b + c * c);
    const double sh = sin(ang / 2);
    q_store(q + 4 * i, q_normalize(Q4{fabs(cos(ang / 2)), a / ang * sh, b / ang * sh, c / ang * sh}));
}

}  // namespace kmpx

using namespace kmpx;

#define QUAT_CHECK(inp, outp)                                                         \
    if (!(inp)) return fail_arg(1, "input is NULL");                                  \
    if (!(outp)) return fail_arg(2, "output is NULL");                                \
    if (M <= 0) return fail_arg(3, "M must be > 0");

extern "C" int kmpx_quat_normalize(const double* q, double* out, long long M, void* stream) {
    QUAT_CHECK(q, out)
    if ((reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return fail_arg(1, "quaternion arrays must be 16-byte aligned");
    k_quat_normalize<<<ceil_div(M, 256), 256, 0, (cudaStream_t)stream>>>(q, out, M);
    return check_launch("k_quat_normalize");
}
extern "C" int kmpx_quat_exp(const double* w, double* q, long long M, void* stream) {
    QUAT_CHECK(w, q)
    if (reinterpret_cast<uintptr_t>(q) & 15) return fail_arg(2, "quaternion arrays must be 16-byte aligned");
    k_quat_exp<<<ceil_div(M, 256), 256, 0, (cudaStream_t)stream>>>(w, q, M);
    return check_launch("k_quat_exp");
}
extern "C" int kmpx_quat_log(const double* q, double* w, long long M, void* stream) {
    QUAT_CHECK(q, w)
    if (reinterpret_cast<uintptr_t>(q) & 15) return fail_arg(1, "quaternion arrays must be 16-byte aligned");
    k_quat_log<<<ceil_div(M, 256), 256, 0, (cudaStream_t)stream>>>(q, w, M);
    return check_launch("k_quat_log");
}
extern "C" int kmpx_quat_mul(const double* p, const double* q, int q_bcast, int conj_q, double* out, long long M,
                             void* stream) {
    if (!p || !q) return fail_arg(1, "input is NULL");
    if (!out) return fail_arg(5, "output is NULL");
    if (M <= 0) return fail_arg(6, "M must be > 0");
    if ((reinterpret_cast<uintptr_t>(p) & 15) || (reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
        return fail_arg(1, "quaternion arrays must be 16-byte aligned");
    k_quat_mul<<<ceil_div(M, 256), 256, 0, (cudaStream_t)stream>>>(p, q, q_bcast, conj_q, out, M);
    return check_launch("k_quat_mul");
}
extern "C" int kmpx_quat_from_rotvec(const double* r, double* q, long long M, void* stream) {
    QUAT_CHECK(r, q)
    if (reinterpret_cast<uintptr_t>(q) & 15) return fail_arg(2, "quaternion arrays must be 16-byte aligned");
    k_quat_from_rotvec<<<ceil_div(M, 256), 256, 0, (cudaStream_t)stream>>>(r, q, M);
    return check_launch("k_quat_from_rotvec");
}
