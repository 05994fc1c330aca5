// Extended-precision (double-double) residual of the edge-form operator
//     r = b - ( C^T diag(alpha) C x + s beta .* x ),      C = diag(1/area) C_topo diag(len).
//
// Why: on a casing model the face weights alpha / area^2 span 1e-7 (steel) .. 1e10 (2.5 mm air cells).  In an fp64
// evaluation the circulation  sum +- len_e x_e  around an air face is a difference of O(|len x|) terms with a rounding
// error of eps |len x|, which the 1e10 weight turns into a residual error that is huge on the scale of the soft
// modes (current loops inside the steel wall, stiffness 1e-3): the converged fp64 iterate is then only accurate to
// 1e-5 .. 1e-4 in j at 1 Hz -- the reference's own fp64 direct solve is ten times worse.  With exact products
// (FMA two-product) and double-double sums the residual of the stored iterate is exact to 1e-32 relative, and a few
// steps of iterative refinement  x += M^-1 r  through the banded factors reach the exact solution of the discrete
// system to fp64 representation accuracy (tests/test_gpu_parity.py compares with oracle.exact_mesh).
//
// Two simple one-thread-per-slot passes (a handful of launches per solve, not a hot kernel):
//   faces:  t_f = w_f * sum_{e in f} +- len_e x_e           (double-double, stored as hi / lo arrays)
//   edges:  r_e = b_e - ( len_e * sum_{f around e} +- t_f + ca beta_e xa_e + cb beta_e xb_e )
// Complex vectors are handled as their real and imaginary parts (the curl-curl part is a real operator):
//   y_re = K x_re + beta (s_re x_re - s_im x_im),   y_im = K x_im + beta (s_re x_im + s_im x_re).
#include "cs_internal.h"
#include "cs_stencil.cuh"

namespace cs {

struct DD { double hi, lo; };
__device__ __forceinline__ DD two_sum(double a, double b) {
    double s = __dadd_rn(a, b), bb = __dadd_rn(s, -a);
    return {s, __dadd_rn(__dadd_rn(a, -__dadd_rn(s, -bb)), __dadd_rn(b, -bb))};
}
__device__ __forceinline__ DD quick_two_sum(double a, double b) {
    double s = __dadd_rn(a, b);
    return {s, __dadd_rn(b, -__dadd_rn(s, -a))};
}
__device__ __forceinline__ DD two_prod(double a, double b) {
    double p = __dmul_rn(a, b);
    return {p, __fma_rn(a, b, -p)};
}
__device__ __forceinline__ DD dd_add(DD a, DD b) {
    DD s = two_sum(a.hi, b.hi), t = two_sum(a.lo, b.lo);
    s.lo = __dadd_rn(s.lo, t.hi);
    s = quick_two_sum(s.hi, s.lo);
    s.lo = __dadd_rn(s.lo, t.lo);
    return quick_two_sum(s.hi, s.lo);
}
__device__ __forceinline__ DD dd_sub(DD a, DD b) { return dd_add(a, DD{-b.hi, -b.lo}); }
__device__ __forceinline__ DD dd_mul_d(DD a, double b) {
    DD p = two_prod(a.hi, b);
    p.lo = __fma_rn(a.lo, b, p.lo);
    return quick_two_sum(p.hi, p.lo);
}
// acc + sgn * a * b with the product exact
__device__ __forceinline__ DD dd_fma(DD acc, double a, double b) { return dd_add(acc, two_prod(a, b)); }

// strided read of one real component (xs = 2: real / imaginary part of an interleaved complex vector)
struct XV {
    const double* x; int xs; long long oy, oz;      // offsets of the Ey and Ez blocks
    __device__ __forceinline__ double ex(long long c) const { return x[c * xs]; }
    __device__ __forceinline__ double ey(long long c) const { return x[(oy + c) * xs]; }
    __device__ __forceinline__ double ez(long long c) const { return x[(oz + c) * xs]; }
};

__global__ void k_dd_faces(MeshD m, const double* __restrict__ wF, const double* __restrict__ x, int xs,
                           double* __restrict__ thi, double* __restrict__ tlo) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    Slot s;
    if (!slot_of(m, t, m.nz + 1, s)) return;
    const int i = s.i, j = s.j, k = s.k;
    XV v{x, xs, m.nEx, m.nEx + m.nEy};
    const long long c = m.cidx(i, j, k);
    const DD zero{0.0, 0.0};
    if (k < m.nz) {
        // Fx: -lEy (Ey[k+1] - Ey[k]) + lEz (Ez[j+1] - Ez[j])
        const long long c1 = m.cidx(i, j, k + 1);
        double ly = m.lEy(i, j);
        DD a = dd_fma(zero, -ly, v.ey(c1));
        a = dd_fma(a, ly, v.ey(c));
        if (!m.sym) {
            double lz = m.lEz(k);
            a = dd_fma(a, lz, v.ez(m.ez(i, m.jp(j), k)));
            a = dd_fma(a, -lz, v.ez(m.ez(i, j, k)));
        }
        a = dd_mul_d(a, wF[c]);
        thi[c] = a.hi; tlo[c] = a.lo;
        if (!m.sym) {
            // Fy: lEx (Ex[k+1] - Ex[k]) - lEz (Ez[i] - Ez[i-1 | axis])
            double lx = m.lEx(i), lz = m.lEz(k);
            double zin = (i > 0) ? v.ez(m.ez(i - 1, j, k)) : v.ez(m.ezaxis(k));
            DD b = dd_fma(zero, lx, v.ex(c1));
            b = dd_fma(b, -lx, v.ex(c));
            b = dd_fma(b, -lz, v.ez(m.ez(i, j, k)));
            b = dd_fma(b, lz, zin);
            b = dd_mul_d(b, wF[m.nFx + c]);
            thi[m.nFx + c] = b.hi; tlo[m.nFx + c] = b.lo;
        }
    }
    // Fz: lEy(i) Ey(i) - lEy(i-1) Ey(i-1) - lEx (Ex[j+1] - Ex[j])
    DD z = dd_fma(zero, m.lEy(i, j), v.ey(c));
    if (i > 0) z = dd_fma(z, -m.lEy(i - 1, j), v.ey(c - 1));
    if (!m.sym) {
        double lx = m.lEx(i);
        z = dd_fma(z, -lx, v.ex(m.cidx(i, m.jp(j), k)));
        z = dd_fma(z, lx, v.ex(c));
    }
    const long long fz = m.nFx + m.nFy + c;
    z = dd_mul_d(z, wF[fz]);
    thi[fz] = z.hi; tlo[fz] = z.lo;
}

struct TF {
    const MeshD& m; const double* hi; const double* lo;
    __device__ __forceinline__ DD operator()(int kind, int i, int j, int k) const {
        long long f = m.cidx(i, j, k) + (kind == 0 ? 0 : (kind == 1 ? m.nFx : m.nFx + m.nFy));
        return DD{hi[f], lo[f]};
    }
};

// r = b - (len * gather + ca beta xa + cb beta xb), one real component; rs / bs: strides of r and b
__device__ __forceinline__ void dd_finish(DD g, double len, double beta, double xa, double ca, double xb, double cb,
                                          const double* b, double* r, long long e, int vs) {
    DD y = dd_mul_d(g, len);
    y = dd_add(y, dd_mul_d(two_prod(beta, xa), ca));
    if (cb != 0.0) y = dd_add(y, dd_mul_d(two_prod(beta, xb), cb));
    DD d = dd_sub(DD{b[e * vs], 0.0}, y);
    r[e * vs] = d.hi + d.lo;
}

__global__ void k_dd_edges(MeshD m, const double* __restrict__ beta, const double* __restrict__ thi, const double* __restrict__ tlo,
                           const double* __restrict__ b, const double* __restrict__ xa, double ca,
                           const double* __restrict__ xb, double cb, int vs, double* __restrict__ r) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    Slot s;
    if (!slot_of(m, t, m.nz + 1, s)) return;
    const int i = s.i, j = s.j, k = s.k;
    TF f{m, thi, tlo};
    const long long c = m.cidx(i, j, k);
    {   // Ey: Fz(i) - Fz(i+1) + Fx(k) - Fx(k-1)
        DD g = f(2, i, j, k);
        if (i + 1 < m.nx) g = dd_sub(g, f(2, i + 1, j, k));
        if (k < m.nz) g = dd_add(g, f(0, i, j, k));
        if (k > 0) g = dd_sub(g, f(0, i, j, k - 1));
        long long e = m.nEx + c;
        dd_finish(g, m.lEy(i, j), beta[e], xa[e * vs], ca, xb[e * vs], cb, b, r, e, vs);
    }
    if (m.sym) return;
    {   // Ex: Fz(j) - Fz(j-1) - Fy(k) + Fy(k-1)
        DD g = dd_sub(f(2, i, j, k), f(2, i, m.jm(j), k));
        if (k < m.nz) g = dd_sub(g, f(1, i, j, k));
        if (k > 0) g = dd_add(g, f(1, i, j, k - 1));
        dd_finish(g, m.lEx(i), beta[c], xa[c * vs], ca, xb[c * vs], cb, b, r, c, vs);
    }
    if (k < m.nz) {
        // Ez: Fx(j-1) - Fx(j) - Fy(i) + Fy(i+1)
        DD g = dd_sub(f(0, i, m.jm(j), k), f(0, i, j, k));
        g = dd_sub(g, f(1, i, j, k));
        if (i + 1 < m.nx) g = dd_add(g, f(1, i + 1, j, k));
        long long e = m.nEx + m.nEy + m.ez(i, j, k);
        dd_finish(g, m.lEz(k), beta[e], xa[e * vs], ca, xb[e * vs], cb, b, r, e, vs);
        if (i == 0 && j == 0) {
            // the shared axis edge: sum over all theta of Fy(0, j, k)
            DD ga{0.0, 0.0};
            for (int jj = 0; jj < m.nth; ++jj) ga = dd_add(ga, f(1, 0, jj, k));
            long long a = m.nEx + m.nEy + m.ezaxis(k);
            dd_finish(ga, m.lEz(k), beta[a], xa[a * vs], ca, xb[a * vs], cb, b, r, a, vs);
        }
    }
}

// r[sys] = b[sys] - A(s_sys) x[sys] for nsys vectors; s_h: 2 doubles per system (host); work: 2 * nF doubles
int launch_edge_resid_dd(const MeshD& m, const double* wF, const double* beta, const double* s_h, const void* x, const void* b,
                         void* r, int nsys, bool cplx_vec, double* work, cudaStream_t st) {
    const long long n = m.nxy * (m.nz + 1);
    const unsigned nblk = (unsigned)cdiv(n, 256);
    double* thi = work; double* tlo = work + m.nF;
    const int vs = cplx_vec ? 2 : 1;
    for (int sys = 0; sys < nsys; ++sys) {
        const double* xs = (const double*)x + (size_t)sys * m.nE * vs;
        const double* bs = (const double*)b + (size_t)sys * m.nE * vs;
        double* rs = (double*)r + (size_t)sys * m.nE * vs;
        const double sr = s_h[2 * sys], si = s_h[2 * sys + 1];
        for (int part = 0; part < vs; ++part) {
            const double* xa = xs + part;                   // own component
            const double* xb = xs + (vs == 2 ? 1 - part : 0);
            const double cb = (vs == 2) ? (part == 0 ? -si : si) : 0.0;
            k_dd_faces<<<nblk, 256, 0, st>>>(m, wF, xa, vs, thi, tlo);
            k_dd_edges<<<nblk, 256, 0, st>>>(m, beta, thi, tlo, bs + part, xa, sr, xb, cb, vs, rs + part);
        }
    }
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

}  // namespace cs
