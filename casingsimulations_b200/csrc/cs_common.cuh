// Common device/host definitions for the casing_b200 solver library (sm_100a).
//
// Vector layouts follow discretize.CylMesh as used by the reference
// (casingSimulations/utils.py:41-47, 63-67): i = r fastest, j = theta, k = z,
// Fortran order, concatenated [Fx|Fy|Fz] and [Ex|Ey|Ez]; a cylindrically
// symmetric mesh (nth == 1) has only [Fx|Fz] faces and Ey edges.  The 3-D Ez
// block holds, per z-cell level k, the single shared axis edge first and then
// the nx*nth edges at r_{i+1}, theta node j.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#define CS_OK 0
#define CS_ERR 1

namespace cs {

void set_error(const std::string& msg);

#define CS_CUDA(call)                                                                  \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            char buf__[512];                                                           \
            snprintf(buf__, sizeof buf__, "%s:%d CUDA error %s: %s", __FILE__, __LINE__, \
                     cudaGetErrorName(e__), cudaGetErrorString(e__));                  \
            cs::set_error(buf__);                                                      \
            return CS_ERR;                                                             \
        }                                                                              \
    } while (0)

#define CS_CHECK(cond, msg)                                                            \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            char buf__[512];                                                           \
            snprintf(buf__, sizeof buf__, "%s:%d %s", __FILE__, __LINE__, msg);        \
            cs::set_error(buf__);                                                      \
            return CS_ERR;                                                             \
        }                                                                              \
    } while (0)

#define CS_TRY(call)                                                                   \
    do {                                                                               \
        int r__ = (call);                                                              \
        if (r__ != CS_OK) return r__;                                                  \
    } while (0)

// ---- complex128 on the device: plain double2 with explicit helpers ----------
typedef double2 cplx;

__host__ __device__ __forceinline__ cplx mk(double re, double im) { return make_double2(re, im); }
__host__ __device__ __forceinline__ cplx operator+(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx operator-(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx operator-(cplx a) { return mk(-a.x, -a.y); }
__host__ __device__ __forceinline__ cplx operator*(cplx a, cplx b) {
    return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cplx operator*(double a, cplx b) { return mk(a * b.x, a * b.y); }
__host__ __device__ __forceinline__ cplx operator*(cplx b, double a) { return mk(a * b.x, a * b.y); }
__host__ __device__ __forceinline__ cplx& operator+=(cplx& a, cplx b) { a.x += b.x; a.y += b.y; return a; }
__host__ __device__ __forceinline__ cplx& operator-=(cplx& a, cplx b) { a.x -= b.x; a.y -= b.y; return a; }
__host__ __device__ __forceinline__ cplx cinv(cplx a) {
    // Smith-free reciprocal: operands here are far from overflow
    double d = 1.0 / (a.x * a.x + a.y * a.y);
    return mk(a.x * d, -a.y * d);
}
__host__ __device__ __forceinline__ cplx operator/(cplx a, cplx b) { return a * cinv(b); }
// fused multiply-accumulate forms: 4 (complex) / 1 (real) DFMA instead of the mul + sub sequence
// the operator forms compile to (6 FP64 instructions per complex term)
__device__ __forceinline__ void fmacc(cplx& acc, cplx a, cplx b) {        // acc += a * b
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
__device__ __forceinline__ void fmsub(cplx& acc, cplx a, cplx b) {        // acc -= a * b
    acc.x = fma(-a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(-a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}
__device__ __forceinline__ void fmacc(double& acc, double a, double b) { acc = fma(a, b, acc); }
__device__ __forceinline__ void fmsub(double& acc, double a, double b) { acc = fma(-a, b, acc); }
__host__ __device__ __forceinline__ cplx conjc(cplx a) { return mk(a.x, -a.y); }

// 1/d without the IEEE slow path: hardware seed (2^-23) + two Newton steps; d normal and nonzero
__device__ __forceinline__ double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    return r;
}
__device__ __forceinline__ cplx fast_inv(cplx a) {
    double d = fast_rcp(fma(a.x, a.x, a.y * a.y));
    return mk(a.x * d, -a.y * d);
}
__device__ __forceinline__ double fast_inv(double a) { return fast_rcp(a); }

// scalar traits so kernels can be written once for double and cplx
template <typename T> struct Sc;
template <> struct Sc<double> {
    __host__ __device__ static __forceinline__ double zero() { return 0.0; }
    __host__ __device__ static __forceinline__ double one() { return 1.0; }
    __host__ __device__ static __forceinline__ double from(double re, double) { return re; }
    __host__ __device__ static __forceinline__ double inv(double a) { return 1.0 / a; }
    __host__ __device__ static __forceinline__ double conj(double a) { return a; }
    __host__ __device__ static __forceinline__ double abs2(double a) { return a * a; }
    __host__ __device__ static __forceinline__ double re(double a) { return a; }
    __host__ __device__ static __forceinline__ double im(double) { return 0.0; }
};
template <> struct Sc<cplx> {
    __host__ __device__ static __forceinline__ cplx zero() { return mk(0.0, 0.0); }
    __host__ __device__ static __forceinline__ cplx one() { return mk(1.0, 0.0); }
    __host__ __device__ static __forceinline__ cplx from(double re, double im) { return mk(re, im); }
    __host__ __device__ static __forceinline__ cplx inv(cplx a) { return cinv(a); }
    __host__ __device__ static __forceinline__ cplx conj(cplx a) { return conjc(a); }
    __host__ __device__ static __forceinline__ double abs2(cplx a) { return a.x * a.x + a.y * a.y; }
    __host__ __device__ static __forceinline__ double re(cplx a) { return a.x; }
    __host__ __device__ static __forceinline__ double im(cplx a) { return a.y; }
};

// ---- mesh tables visible to kernels (passed by value) ------------------------
struct MeshD {
    int nx, nth, nz, sym;          // sym == 1 <=> nth == 1 (cylindrically symmetric)
    const double* hx;              // [nx]   radial widths
    const double* rn;              // [nx+1] radial nodes, rn[0] = 0
    const double* ring;            // [nx]   0.5*(rn[i+1]^2 - rn[i]^2)
    const double* hth;             // [nth]  azimuthal widths (sum 2*pi)
    const double* hz;              // [nz]   vertical widths
    double w_axis_edge, w_axis_face;
    long long nC, nFx, nFy, nFz, nEx, nEy, nEz, nF, nE;
    long long nxy;                 // nx*nth
    long long ezp;                 // nx*nth + 1 : Ez entries per z level (3-D)

    __host__ __device__ __forceinline__ int jm(int j) const { return j == 0 ? nth - 1 : j - 1; }
    __host__ __device__ __forceinline__ int jp(int j) const { return j == nth - 1 ? 0 : j + 1; }
    // flattened indices inside each block
    __host__ __device__ __forceinline__ long long cidx(int i, int j, int k) const { return i + (long long)nx * (j + (long long)nth * k); }
    __host__ __device__ __forceinline__ long long ez(int i, int j, int k) const { return k * ezp + 1 + i + (long long)nx * j; }
    __host__ __device__ __forceinline__ long long ezaxis(int k) const { return k * ezp; }
#ifdef __CUDACC__
    __device__ __forceinline__ double aFx(int i, int j, int k) const { return __ldg(rn + i + 1) * __ldg(hth + j) * __ldg(hz + k); }
    __device__ __forceinline__ double aFy(int i, int k) const { return __ldg(hx + i) * __ldg(hz + k); }
    __device__ __forceinline__ double aFz(int i, int j) const { return __ldg(ring + i) * __ldg(hth + j); }
    __device__ __forceinline__ double vol(int i, int j, int k) const { return __ldg(ring + i) * __ldg(hth + j) * __ldg(hz + k); }
    __device__ __forceinline__ double lEx(int i) const { return __ldg(hx + i); }
    __device__ __forceinline__ double lEy(int i, int j) const { return __ldg(rn + i + 1) * __ldg(hth + j); }
    __device__ __forceinline__ double lEz(int k) const { return __ldg(hz + k); }
#endif
};

static inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

}  // namespace cs
