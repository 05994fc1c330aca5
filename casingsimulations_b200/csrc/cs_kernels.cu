// K1-K4, K8: model painting, inner-product assembly, and matrix-free stencil
// applies on the (r, theta, z) mesh.  All arithmetic fp64 / complex128; every
// kernel is HBM-bound (DESIGN.md gives bytes per cell) and uses one thread per
// (i, j, k) slot with i fastest => fully coalesced streams of the big arrays,
// neighbour values come from L1/L2 (a z-plane is nx*nth*16 B <= 223 KB).
#include "cs_internal.h"
#include <cstdlib>
#include "cs_stencil.cuh"
#include <cuda_pipeline.h>

namespace cs {

static const int TPB = 256;
static inline dim3 grid1(long long n, int nsys = 1) { return dim3((unsigned)cdiv(n, TPB), (unsigned)nsys, 1); }
#define MU_0 1.25663706143591729539e-06

// ============================ K1: paint ========================================
// params (doubles): 0 sigma_back 1 has_air 2 sigma_air 3 surface_z 4 has_casing
// 5 sigma_casing 6 sigma_inside 7 casing_a 8 casing_b 9 casing_z0 10 casing_z1
// 11 mur_back 12 mur_casing 13 mur_inside 14 has_layer 15 sigma_layer
// 16 layer_z0 17 layer_z1.  Predicates are the strict inequalities of
// model.py:249 (air), :290-293 (layer), :514-529, :538 (casing / inside);
// paint order background -> air -> layer -> casing -> inside (model.py:258-260,
// 301-303, 566-567).
__global__ void k_paint(MeshD m, const double* __restrict__ zn, const double* __restrict__ par,
                        double* __restrict__ sigma, double* __restrict__ mu) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    Slot s;
    if (!slot_of(m, t, m.nz, s)) return;
    double rc = 0.5 * (m.rn[s.i + 1] + m.rn[s.i]);
    double zc = 0.5 * (zn[s.k + 1] + zn[s.k]);
    double sg = par[0], mr = par[11];
    if (par[1] != 0.0 && zc > par[3]) sg = par[2];
    if (par[14] != 0.0 && zc < par[17] && zc > par[16]) sg = par[15];
    if (par[4] != 0.0) {
        bool inz = (zc > par[9]) && (zc < par[10]);
        if (inz && rc > par[7] && rc < par[8]) { sg = par[5]; mr = par[12]; }
        if (inz && rc < par[7]) { sg = par[6]; mr = par[13]; }
    }
    sigma[t] = sg;
    mu[t] = MU_0 * mr;
}

int launch_paint(const MeshD& m, const double* zn_d, const double* par_d, double* sigma, double* mu, cudaStream_t st) {
    k_paint<<<grid1(m.nC), TPB, 0, st>>>(m, zn_d, par_d, sigma, mu);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// ============================ K2: inner products ================================
// Mf(p)_f = sum over the (<=2) cells touching f of w * vol_c p_c, w = 1/2 except
// the innermost cell's share of the face at r_1 (w_axis_face); Me(p)_e likewise
// with 4 cells and w = 1/4 (radial factor of the Ey edge at r_1: w_axis_edge);
// the shared axis Ez edge collects 1/2 vol p from every innermost cell.
__device__ __forceinline__ double vp(const MeshD& m, const double* __restrict__ p, int inv, int i, int j, int k) {
    double v = p ? p[m.cidx(i, j, k)] : 1.0;     // p == nullptr: unit property (Mf, Me)
    return m.vol(i, j, k) * (inv ? 1.0 / v : v);
}

__global__ void k_face_ip(MeshD m, const double* __restrict__ p, int invProp, int invMat, double* __restrict__ out) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    Slot s;
    if (!slot_of(m, t, m.nz + 1, s)) return;
    int i = s.i, j = s.j, k = s.k;
    if (k < m.nz) {
        double a = (i == 0 ? m.w_axis_face : 0.5) * vp(m, p, invProp, i, j, k);
        if (i + 1 < m.nx) a += 0.5 * vp(m, p, invProp, i + 1, j, k);
        out[m.cidx(i, j, k)] = invMat ? 1.0 / a : a;
        if (!m.sym) {
            double b = 0.5 * (vp(m, p, invProp, i, m.jm(j), k) + vp(m, p, invProp, i, j, k));
            out[m.nFx + m.cidx(i, j, k)] = invMat ? 1.0 / b : b;
        }
    }
    double c = 0.0;
    if (k > 0) c += 0.5 * vp(m, p, invProp, i, j, k - 1);
    if (k < m.nz) c += 0.5 * vp(m, p, invProp, i, j, k);
    out[m.nFx + m.nFy + m.cidx(i, j, k)] = invMat ? 1.0 / c : c;
}

__global__ void k_edge_ip(MeshD m, const double* __restrict__ p, int invProp, int invMat, double* __restrict__ out) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    Slot s;
    if (!slot_of(m, t, m.nz + 1, s)) return;
    int i = s.i, j = s.j, k = s.k;
    double wr = (i == 0) ? m.w_axis_edge : 0.5;
    // Ey (r_{i+1}, cell j, node k): cells (i | i+1, j, k-1 | k)
    double ey = 0.0;
    for (int kk = k - 1; kk <= k; ++kk) {
        if (kk < 0 || kk >= m.nz) continue;
        ey += 0.5 * wr * vp(m, p, invProp, i, j, kk);
        if (i + 1 < m.nx) ey += 0.25 * vp(m, p, invProp, i + 1, j, kk);
    }
    out[m.nEx + m.cidx(i, j, k)] = invMat ? 1.0 / ey : ey;
    if (m.sym) return;
    // Ex (cell i, node j, node k): cells (i, j-1 | j, k-1 | k)
    double ex = 0.0;
    for (int kk = k - 1; kk <= k; ++kk) {
        if (kk < 0 || kk >= m.nz) continue;
        ex += 0.25 * (vp(m, p, invProp, i, m.jm(j), kk) + vp(m, p, invProp, i, j, kk));
    }
    out[m.cidx(i, j, k)] = invMat ? 1.0 / ex : ex;
    if (k < m.nz) {
        // Ez (r_{i+1}, node j, cell k): cells (i | i+1, j-1 | j, k)
        double ez = 0.25 * (vp(m, p, invProp, i, m.jm(j), k) + vp(m, p, invProp, i, j, k));
        if (i + 1 < m.nx) ez += 0.25 * (vp(m, p, invProp, i + 1, m.jm(j), k) + vp(m, p, invProp, i + 1, j, k));
        out[m.nEx + m.nEy + m.ez(i, j, k)] = invMat ? 1.0 / ez : ez;
        if (i == 0 && j == 0) {
            double ax = 0.0;
            for (int jj = 0; jj < m.nth; ++jj) ax += 0.5 * vp(m, p, invProp, 0, jj, k);
            out[m.nEx + m.nEy + m.ezaxis(k)] = invMat ? 1.0 / ax : ax;
        }
    }
}

int launch_face_ip(const MeshD& m, const double* prop, int invProp, int invMat, double* out, cudaStream_t st) {
    k_face_ip<<<grid1(m.nxy * (m.nz + 1)), TPB, 0, st>>>(m, prop, invProp, invMat, out);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}
int launch_edge_ip(const MeshD& m, const double* prop, int invProp, int invMat, double* out, cudaStream_t st) {
    k_edge_ip<<<grid1(m.nxy * (m.nz + 1)), TPB, 0, st>>>(m, prop, invProp, invMat, out);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

__global__ void k_face_scale(MeshD m, const double* __restrict__ in, double* __restrict__ out, int mode) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    Slot s;
    if (!slot_of(m, t, m.nz + 1, s)) return;
    int i = s.i, j = s.j, k = s.k;
    auto f = [&](double v, double a) -> double {
        switch (mode) {
            case 0: return a * a / v;
            case 1: return v / (a * a);
            case 2: return v / a;
            default: return v * a;
        }
    };
    if (k < m.nz) {
        long long c = m.cidx(i, j, k);
        out[c] = f(in[c], m.aFx(i, j, k));
        if (!m.sym) out[m.nFx + c] = f(in[m.nFx + c], m.aFy(i, k));
    }
    long long c = m.nFx + m.nFy + m.cidx(i, j, k);
    out[c] = f(in[c], m.aFz(i, j));
}
__global__ void k_edge_scale(MeshD m, const double* __restrict__ in, double* __restrict__ out, int mode) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    Slot s;
    if (!slot_of(m, t, m.nz + 1, s)) return;
    int i = s.i, j = s.j, k = s.k;
    auto f = [&](double v, double l) -> double { return mode == 0 ? l * l * v : (mode == 1 ? l * l / v : v / (l * l)); };
    long long c = m.cidx(i, j, k);
    out[m.nEx + c] = f(in[m.nEx + c], m.lEy(i, j));
    if (m.sym) return;
    out[c] = f(in[c], m.lEx(i));
    if (k < m.nz) {
        long long e = m.nEx + m.nEy + m.ez(i, j, k);
        out[e] = f(in[e], m.lEz(k));
        if (i == 0 && j == 0) {
            long long a = m.nEx + m.nEy + m.ezaxis(k);
            out[a] = f(in[a], m.lEz(k));
        }
    }
}
int launch_face_scale(const MeshD& m, const double* in, double* out, int mode, cudaStream_t st) {
    k_face_scale<<<grid1(m.nxy * (m.nz + 1)), TPB, 0, st>>>(m, in, out, mode);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}
int launch_edge_scale(const MeshD& m, const double* in, double* out, int mode, cudaStream_t st) {
    k_edge_scale<<<grid1(m.nxy * (m.nz + 1)), TPB, 0, st>>>(m, in, out, mode);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// ============================ K3: DC apply =======================================
// y_c = sum_f g_f (x_c - x_nbr(f)); faces on the outer r / bottom / top boundary
// have no neighbour (Dirichlet phi = 0 outside; SimPEG Problem3D_CC, run.py:339-344)
// g = area^2 / MfRho laid out like a face vector.  40 B/cell (3-D), 32 B/cell (2-D).
// One thread per cell, i fastest: 55 us at 4.86 M cells = 3.5 TB/s = 54 % of the measured copy peak.
// Two alternatives were built and measured on B200 and did not beat it: a register z-marching
// variant (63-97 us: fewer, longer threads expose less memory-level parallelism) and a K4-style
// tiled kernel with cp.async-staged x planes (55 us: DRAM traffic and L2 traffic drop, but at one
// cell per thread per level the staging / addressing overhead is 226 instructions per cell and the
// kernel becomes issue-bound, ncu issue-active 63 %).  Next step: several cells per thread.
template <typename T>
__global__ void k_dc_apply(MeshD m, const double* __restrict__ g, double shift0,
                           const T* __restrict__ x, T* __restrict__ y) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    Slot s;
    if (!slot_of(m, t, m.nz, s)) return;
    x += blockIdx.y * m.nC;
    y += blockIdx.y * m.nC;
    int i = s.i, j = s.j, k = s.k;
    const double* gx = g;
    const double* gy = g + m.nFx;
    const double* gz = g + m.nFx + m.nFy;
    T xc = x[t];
    T acc = gx[t] * ((i + 1 < m.nx) ? (xc - x[t + 1]) : xc);
    if (i > 0) acc += gx[t - 1] * (xc - x[t - 1]);
    if (!m.sym) {
        long long cm = m.cidx(i, m.jm(j), k), cp = m.cidx(i, m.jp(j), k);
        acc += gy[t] * (xc - x[cm]) + gy[cp] * (xc - x[cp]);
    }
    acc += gz[t] * ((k > 0) ? (xc - x[t - m.nxy]) : xc);
    acc += gz[t + m.nxy] * ((k + 1 < m.nz) ? (xc - x[t + m.nxy]) : xc);
    if (t == 0) acc += shift0 * xc;
    y[t] = acc;
}

// Variant for large 3-D meshes: a thread owns one (r, theta) column over ZC consecutive z levels and keeps the three x planes
// and the lower z weight in registers while it marches, so the z neighbours and gz come from HBM/L2 once instead of three /
// two times (the one-cell kernel moves ~2.5x the algorithmic bytes through L2).  ZC = 8 keeps 600 k threads in flight at
// 4.86 M cells (a full-height march had too few threads, see above).
template <typename T, int ZC>
__global__ void __launch_bounds__(256)
k_dc_apply_zc(MeshD m, const double* __restrict__ g, double shift0, const T* __restrict__ x, T* __restrict__ y) {
    const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;      // (i, j) column
    if (col >= m.nxy) return;
    const int k0 = blockIdx.y * ZC;
    x += blockIdx.z * m.nC;
    y += blockIdx.z * m.nC;
    const int i = (int)(col % m.nx), j = (int)(col / m.nx);
    const long long cm = (long long)i + (long long)m.nx * m.jm(j) - col, cp = (long long)i + (long long)m.nx * m.jp(j) - col;   // theta neighbours, relative
    const double* gx = g;
    const double* gy = g + m.nFx;
    const double* gz = g + m.nFx + m.nFy;
    long long t = col + (long long)k0 * m.nxy;
    const int kend = min(k0 + ZC, m.nz);
    T xb = (k0 > 0) ? x[t - m.nxy] : Sc<T>::zero(), xc = x[t];
    double gzl = gz[t];
    for (int k = k0; k < kend; ++k, t += m.nxy) {
        const bool top = (k + 1 >= m.nz);
        T xa = top ? Sc<T>::zero() : x[t + m.nxy];
        double gzu = gz[t + m.nxy];
        T acc = gx[t] * ((i + 1 < m.nx) ? (xc - x[t + 1]) : xc);
        if (i > 0) acc += gx[t - 1] * (xc - x[t - 1]);
        acc += gy[t] * (xc - x[t + cm]) + gy[t + cp] * (xc - x[t + cp]);
        acc += gzl * ((k > 0) ? (xc - xb) : xc);
        acc += gzu * (top ? xc : (xc - xa));
        if (t == 0) acc += shift0 * xc;
        y[t] = acc;
        xb = xc; xc = xa; gzl = gzu;
    }
}

int launch_dc_apply(const MeshD& m, const double* g, double shift0, const void* x, void* y, int nsys, bool cplx_vec, cudaStream_t st) {
    static const int zc_on = [] { const char* e = getenv("CS_DC_ZC"); return e ? atoi(e) : 1; }();
    if (zc_on && !m.sym && m.nC >= (1 << 20) && m.nz >= 64) {
        const int ZC = 8;
        dim3 grd((unsigned)cdiv(m.nxy, 256), (unsigned)((m.nz + ZC - 1) / ZC), (unsigned)nsys);
        if (cplx_vec) k_dc_apply_zc<cplx, ZC><<<grd, 256, 0, st>>>(m, g, shift0, (const cplx*)x, (cplx*)y);
        else k_dc_apply_zc<double, ZC><<<grd, 256, 0, st>>>(m, g, shift0, (const double*)x, (double*)y);
        CS_CUDA(cudaGetLastError());
        return CS_OK;
    }
    if (cplx_vec) k_dc_apply<cplx><<<grid1(m.nC, nsys), TPB, 0, st>>>(m, g, shift0, (const cplx*)x, (cplx*)y);
    else k_dc_apply<double><<<grid1(m.nC, nsys), TPB, 0, st>>>(m, g, shift0, (const double*)x, (double*)y);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// Dt = vol*faceDiv = D_topo diag(area): faces -> cells, and its transpose
template <typename T>
__global__ void k_div(MeshD m, const T* __restrict__ xF, T* __restrict__ yC) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    Slot s;
    if (!slot_of(m, t, m.nz, s)) return;
    int i = s.i, j = s.j, k = s.k;
    FaceV<T> f(m, xF);
    T acc = m.aFx(i, j, k) * f.x[t];
    if (i > 0) acc -= m.aFx(i - 1, j, k) * f.x[t - 1];
    if (!m.sym) acc += m.aFy(i, k) * (f.y[m.cidx(i, m.jp(j), k)] - f.y[t]);
    acc += m.aFz(i, j) * (f.z[t + m.nxy] - f.z[t]);
    yC[t] = acc;
}
template <typename T>
__global__ void k_divT(MeshD m, const T* __restrict__ xC, T* __restrict__ yF) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    Slot s;
    if (!slot_of(m, t, m.nz + 1, s)) return;
    int i = s.i, j = s.j, k = s.k;
    long long c = m.cidx(i, j, k);
    if (k < m.nz) {
        T xc = xC[c];
        yF[c] = m.aFx(i, j, k) * ((i + 1 < m.nx) ? (xc - xC[c + 1]) : xc);
        if (!m.sym) yF[m.nFx + c] = m.aFy(i, k) * (xC[m.cidx(i, m.jm(j), k)] - xc);
    }
    T lo = (k > 0) ? xC[c - m.nxy] : Sc<T>::zero();
    T hi = (k < m.nz) ? xC[c] : Sc<T>::zero();
    yF[m.nFx + m.nFy + c] = m.aFz(i, j) * (lo - hi);
}
int launch_div(const MeshD& m, const void* xF, void* yC, bool cplx_vec, cudaStream_t st) {
    if (cplx_vec) k_div<cplx><<<grid1(m.nC), TPB, 0, st>>>(m, (const cplx*)xF, (cplx*)yC);
    else k_div<double><<<grid1(m.nC), TPB, 0, st>>>(m, (const double*)xF, (double*)yC);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}
int launch_divT(const MeshD& m, const void* xC, void* yF, bool cplx_vec, cudaStream_t st) {
    long long n = m.nxy * (m.nz + 1);
    if (cplx_vec) k_divT<cplx><<<grid1(n), TPB, 0, st>>>(m, (const cplx*)xC, (cplx*)yF);
    else k_divT<double><<<grid1(n), TPB, 0, st>>>(m, (const double*)xC, (double*)yF);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// ============================ curl / curl^T ======================================
template <typename T>
__global__ void k_curl(MeshD m, const T* __restrict__ xE, T* __restrict__ yF) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    Slot s;
    if (!slot_of(m, t, m.nz + 1, s)) return;
    int i = s.i, j = s.j, k = s.k;
    EdgeV<T> e(m, xE);
    long long c = m.cidx(i, j, k);
    if (k < m.nz) {
        yF[c] = (1.0 / m.aFx(i, j, k)) * circFx(m, e, i, j, k);
        if (!m.sym) yF[m.nFx + c] = (1.0 / m.aFy(i, k)) * circFy(m, e, i, j, k);
    }
    yF[m.nFx + m.nFy + c] = (1.0 / m.aFz(i, j)) * circFz(m, e, i, j, k);
}

template <typename T> struct PlainFace {
    const MeshD& m; FaceV<T> f;
    __device__ PlainFace(const MeshD& m_, const T* v) : m(m_), f(m_, v) {}
    // value / area of face (kind, i, j, k)
    __device__ __forceinline__ T operator()(int kind, int i, int j, int k) const {
        long long c = m.cidx(i, j, k);
        if (kind == 0) return (1.0 / m.aFx(i, j, k)) * f.x[c];
        if (kind == 1) return (1.0 / m.aFy(i, k)) * f.y[c];
        return (1.0 / m.aFz(i, j)) * f.z[c];
    }
};

template <typename T>
__global__ void k_curlT(MeshD m, const T* __restrict__ xF, T* __restrict__ yE) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    Slot s;
    if (!slot_of(m, t, m.nz + 1, s)) return;
    int i = s.i, j = s.j, k = s.k;
    PlainFace<T> f(m, xF);
    long long c = m.cidx(i, j, k);
    yE[m.nEx + c] = m.lEy(i, j) * gathEy<T>(m, f, i, j, k);
    if (m.sym) return;
    yE[c] = m.lEx(i) * gathEx<T>(m, f, i, j, k);
    if (k < m.nz) {
        yE[m.nEx + m.nEy + m.ez(i, j, k)] = m.lEz(k) * gathEz<T>(m, f, i, j, k);
        if (i == 0 && j == 0) yE[m.nEx + m.nEy + m.ezaxis(k)] = m.lEz(k) * gathEzAxis<T>(m, f, k);
    }
}
int launch_curl(const MeshD& m, const void* xE, void* yF, bool cplx_vec, cudaStream_t st) {
    long long n = m.nxy * (m.nz + 1);
    if (cplx_vec) k_curl<cplx><<<grid1(n), TPB, 0, st>>>(m, (const cplx*)xE, (cplx*)yF);
    else k_curl<double><<<grid1(n), TPB, 0, st>>>(m, (const double*)xE, (double*)yF);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}
int launch_curlT(const MeshD& m, const void* xF, void* yE, bool cplx_vec, cudaStream_t st) {
    long long n = m.nxy * (m.nz + 1);
    if (cplx_vec) k_curlT<cplx><<<grid1(n), TPB, 0, st>>>(m, (const cplx*)xF, (cplx*)yE);
    else k_curlT<double><<<grid1(n), TPB, 0, st>>>(m, (const double*)xF, (double*)yE);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// ============================ nodal gradient and nodal Laplacian (3-D) =============
// discretize mesh.nodalGrad = diag(1/len) G_topo on the deflated node set: per z-node level the single
// axis node first, then the nx*nth nodes at (r_{i}, theta_j), i = 1..nx.  Needed by the E-B time stepping: the DC initial
// state of a grounded source, (G^T MeSigma G) phi = G^T s_e, e0 = -G phi, and the per-step cleaning of the
// gradient null space.  npl = nx*nth + 1 nodes per level.
__device__ __forceinline__ long long nodeidx(const MeshD& m, int i /*0 = axis, 1..nx*/, int j, int k) {
    return (long long)k * m.ezp + (i == 0 ? 0 : 1 + (i - 1) + (long long)m.nx * j);
}
// y_E = scale_E .* (G_topo x_N)      (scale = 1/len gives nodalGrad; scale may be null = topological gradient)
template <typename T>
__global__ void k_node_grad(MeshD m, const T* __restrict__ xN, T* __restrict__ yE, int with_len) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    Slot s;
    if (!slot_of(m, t, m.nz + 1, s)) return;
    const int i = s.i, j = s.j, k = s.k;                 // edge radial index i <-> node radial index i + 1
    const long long c = m.cidx(i, j, k);
    const T xc = xN[nodeidx(m, i + 1, j, k)];
    const double sx = with_len ? 1.0 / m.lEx(i) : 1.0, sy = with_len ? 1.0 / m.lEy(i, j) : 1.0;
    yE[c] = sx * (xc - xN[nodeidx(m, i, j, k)]);                                   // Ex: r_i -> r_{i+1}
    yE[m.nEx + c] = sy * (xN[nodeidx(m, i + 1, m.jp(j), k)] - xc);                 // Ey: theta_j -> theta_{j+1}
    if (k < m.nz) {
        const double sz = with_len ? 1.0 / m.lEz(k) : 1.0;
        yE[m.nEx + m.nEy + m.ez(i, j, k)] = sz * (xN[nodeidx(m, i + 1, j, k + 1)] - xc);
        if (i == 0 && j == 0) yE[m.nEx + m.nEy + m.ezaxis(k)] = sz * (xN[nodeidx(m, 0, 0, k + 1)] - xN[nodeidx(m, 0, 0, k)]);
    }
}
// y_N = G_topo^T (w_E .* (G_topo x_N)) + shift0 * x[0] on row 0 ; w == null: y_N = G_topo^T xin_E (plain transpose, xin = x)
template <typename T, bool LAPLACE>
__global__ void k_node_apply(MeshD m, const double* __restrict__ w, double shift0, const T* __restrict__ x, T* __restrict__ y) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    Slot s;
    if (!slot_of(m, t, m.nz + 1, s)) return;
    const int i = s.i, j = s.j, k = s.k;                 // node (i + 1, j, k)
    const long long c = m.cidx(i, j, k);
    const long long oy = m.nEx, oz = m.nEx + m.nEy;
    const int jm = m.jm(j), jp = m.jp(j);
    T acc = Sc<T>::zero();
    if (LAPLACE) {
        const T xc = x[nodeidx(m, i + 1, j, k)];
        acc += w[c] * (xc - x[nodeidx(m, i, j, k)]);                                             // Ex(i): towards r_i / the axis
        if (i + 1 < m.nx) acc += w[c + 1] * (xc - x[nodeidx(m, i + 2, j, k)]);                   // Ex(i+1): towards r_{i+2}
        acc += w[oy + c] * (xc - x[nodeidx(m, i + 1, jp, k)]);                                   // Ey(i, j)
        acc += w[oy + m.cidx(i, jm, k)] * (xc - x[nodeidx(m, i + 1, jm, k)]);                    // Ey(i, j-1)
        if (k < m.nz) acc += w[oz + m.ez(i, j, k)] * (xc - x[nodeidx(m, i + 1, j, k + 1)]);
        if (k > 0) acc += w[oz + m.ez(i, j, k - 1)] * (xc - x[nodeidx(m, i + 1, j, k - 1)]);
    } else {
        // G_topo^T: + for the edge's head node, - for its tail node
        acc += x[c];                                                                             // Ex(i) ends here
        if (i + 1 < m.nx) acc -= x[c + 1];                                                       // Ex(i+1) starts here
        acc -= x[oy + c];                                                                        // Ey(i, j) starts here
        acc += x[oy + m.cidx(i, jm, k)];                                                         // Ey(i, j-1) ends here
        if (k < m.nz) acc -= x[oz + m.ez(i, j, k)];
        if (k > 0) acc += x[oz + m.ez(i, j, k - 1)];
    }
    y[nodeidx(m, i + 1, j, k)] = acc;
}
// the axis node of every level: Ex(0, j, k) for all j + the axis Ez edges above / below
template <typename T, bool LAPLACE>
__global__ void k_node_apply_axis(MeshD m, const double* __restrict__ w, double shift0, const T* __restrict__ x, T* __restrict__ y) {
    int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (k > m.nz) return;
    const long long oz = m.nEx + m.nEy;
    double re = 0.0, im = 0.0;
    T xa = LAPLACE ? x[nodeidx(m, 0, 0, k)] : Sc<T>::zero();
    for (int j = lane; j < m.nth; j += 32) {
        long long c = m.cidx(0, j, k);
        T v = LAPLACE ? w[c] * (xa - x[nodeidx(m, 1, j, k)]) : Sc<T>::zero() - x[c];            // Ex(0, j, k) starts at the axis
        re += Sc<T>::re(v); im += Sc<T>::im(v);
    }
    for (int o = 16; o > 0; o >>= 1) { re += __shfl_down_sync(0xffffffffu, re, o); im += __shfl_down_sync(0xffffffffu, im, o); }
    if (lane == 0) {
        T acc = Sc<T>::from(re, im);
        if (LAPLACE) {
            if (k < m.nz) acc += w[oz + m.ezaxis(k)] * (xa - x[nodeidx(m, 0, 0, k + 1)]);
            if (k > 0) acc += w[oz + m.ezaxis(k - 1)] * (xa - x[nodeidx(m, 0, 0, k - 1)]);
            if (k == 0) acc += shift0 * xa;                                                      // node 0 = the axis node of level 0
        } else {
            if (k < m.nz) acc -= x[oz + m.ezaxis(k)];
            if (k > 0) acc += x[oz + m.ezaxis(k - 1)];
        }
        y[nodeidx(m, 0, 0, k)] = acc;
    }
}
int launch_node_grad(const MeshD& m, const void* xN, void* yE, bool cplx_vec, bool with_len, cudaStream_t st) {
    CS_CHECK(!m.sym, "nodal operators need a 3-D mesh (a cylindrically symmetric mesh has no nodes)");
    long long n = m.nxy * (m.nz + 1);
    if (cplx_vec) k_node_grad<cplx><<<grid1(n), TPB, 0, st>>>(m, (const cplx*)xN, (cplx*)yE, with_len ? 1 : 0);
    else k_node_grad<double><<<grid1(n), TPB, 0, st>>>(m, (const double*)xN, (double*)yE, with_len ? 1 : 0);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}
// w != null: y = G_topo^T diag(w) G_topo x (+ shift0 x[0]), x, y nodal, nsys vectors;  w == null: y_N = G_topo^T x_E (one vector)
int launch_node_apply(const MeshD& m, const double* w, double shift0, const void* x, void* y, int nsys, bool cplx_vec, cudaStream_t st) {
    CS_CHECK(!m.sym, "nodal operators need a 3-D mesh (a cylindrically symmetric mesh has no nodes)");
    const long long n = m.nxy * (m.nz + 1), nN = (long long)(m.nz + 1) * m.ezp;
    const long long nin = w ? nN : m.nE;
    for (int sidx = 0; sidx < nsys; ++sidx) {
        const size_t es = cplx_vec ? sizeof(cplx) : sizeof(double);
        const void* xs = (const char*)x + (size_t)sidx * nin * es;
        void* ys = (char*)y + (size_t)sidx * nN * es;
        dim3 ga((m.nz + 1 + 7) / 8);
        if (w) {
            if (cplx_vec) { k_node_apply<cplx, true><<<grid1(n), TPB, 0, st>>>(m, w, shift0, (const cplx*)xs, (cplx*)ys);
                            k_node_apply_axis<cplx, true><<<ga, 256, 0, st>>>(m, w, shift0, (const cplx*)xs, (cplx*)ys); }
            else { k_node_apply<double, true><<<grid1(n), TPB, 0, st>>>(m, w, shift0, (const double*)xs, (double*)ys);
                   k_node_apply_axis<double, true><<<ga, 256, 0, st>>>(m, w, shift0, (const double*)xs, (double*)ys); }
        } else {
            if (cplx_vec) { k_node_apply<cplx, false><<<grid1(n), TPB, 0, st>>>(m, w, 0.0, (const cplx*)xs, (cplx*)ys);
                            k_node_apply_axis<cplx, false><<<ga, 256, 0, st>>>(m, w, 0.0, (const cplx*)xs, (cplx*)ys); }
            else { k_node_apply<double, false><<<grid1(n), TPB, 0, st>>>(m, w, 0.0, (const double*)xs, (double*)ys);
                   k_node_apply_axis<double, false><<<ga, 256, 0, st>>>(m, w, 0.0, (const double*)xs, (double*)ys); }
        }
    }
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// ============================ face / edge -> cell-centre vector averages ==========
// discretize mesh.aveF2CCV / aveE2CCV as used by the viewer's re-gridding (view.py:300-308): per cell the average of the
// faces (2) / edges (4) of each component around it, output blocks [x | y | z] of nC (symmetric mesh: [x | z] for faces,
// [y] for edges).  Same radial axis weights as the inner products.
template <typename T>
__global__ void k_ave_f2ccv(MeshD m, const T* __restrict__ xF, T* __restrict__ y) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    Slot s;
    if (!slot_of(m, t, m.nz, s)) return;
    const int i = s.i, j = s.j, k = s.k;
    const long long c = m.cidx(i, j, k);
    const T* fx = xF; const T* fy = xF + m.nFx; const T* fz = xF + m.nFx + m.nFy;
    y[c] = (i > 0) ? 0.5 * (fx[c - 1] + fx[c]) : m.w_axis_face * fx[c];
    long long o = m.nC;
    if (!m.sym) { y[o + c] = 0.5 * (fy[c] + fy[m.cidx(i, m.jp(j), k)]); o += m.nC; }
    y[o + c] = 0.5 * (fz[c] + fz[c + m.nxy]);
}
template <typename T>
__global__ void k_ave_e2ccv(MeshD m, const T* __restrict__ xE, T* __restrict__ y) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    Slot s;
    if (!slot_of(m, t, m.nz, s)) return;
    const int i = s.i, j = s.j, k = s.k;
    const long long c = m.cidx(i, j, k), cu = c + m.nxy;      // node levels k and k + 1
    const T* ey = xE + m.nEx;
    T vy = (i > 0) ? 0.25 * (ey[c - 1] + ey[c] + ey[cu - 1] + ey[cu]) : (0.5 * m.w_axis_edge) * (ey[c] + ey[cu]);
    if (m.sym) { y[c] = vy; return; }
    const T* ex = xE; const T* ez = xE + m.nEx + m.nEy;
    const int jp = m.jp(j);
    const long long cj = m.cidx(i, jp, k);
    y[c] = 0.25 * (ex[c] + ex[cj] + ex[cu] + ex[cj + m.nxy]);
    y[m.nC + c] = vy;
    T zin0 = (i > 0) ? ez[m.ez(i - 1, j, k)] : ez[m.ezaxis(k)], zin1 = (i > 0) ? ez[m.ez(i - 1, jp, k)] : ez[m.ezaxis(k)];
    y[2 * m.nC + c] = 0.25 * (zin0 + ez[m.ez(i, j, k)] + zin1 + ez[m.ez(i, jp, k)]);
}
int launch_ave_ccv(const MeshD& m, bool from_faces, const void* x, void* y, bool cplx_vec, cudaStream_t st) {
    if (from_faces) {
        if (cplx_vec) k_ave_f2ccv<cplx><<<grid1(m.nC), TPB, 0, st>>>(m, (const cplx*)x, (cplx*)y);
        else k_ave_f2ccv<double><<<grid1(m.nC), TPB, 0, st>>>(m, (const double*)x, (double*)y);
    } else {
        if (cplx_vec) k_ave_e2ccv<cplx><<<grid1(m.nC), TPB, 0, st>>>(m, (const cplx*)x, (cplx*)y);
        else k_ave_e2ccv<double><<<grid1(m.nC), TPB, 0, st>>>(m, (const double*)x, (double*)y);
    }
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// ============================ K4: fused edge-form operator =======================
// y = C^T diag(alpha) C x + s * beta .* x  with C = diag(1/area) C_topo diag(len):
//   y_e = len_e * sum_{f around e} +-[ wF_f * circ_f(len .* x) ] + s beta_e x_e,
// wF = alpha / area^2.  Serves FDEM-h (alpha = MfRho, beta = MeMu, s = i w),
// FDEM-e / TDEM-e (alpha = MfMui, beta = MeSigma, s = i w | 1/dt).  3-D complex:
// x 48 + y 48 + wF 24 + beta 24 = 144 B/cell; 2-D symmetric: 56 B/cell.
// Implementation: each CTA owns a tile of 128 radial cells x TJ azimuthal cells (+ one halo
// row j0-1, + one halo lane when the tile is not the last in r) and marches EDGE_KZ node
// levels in z.  The x values of a level (a (TJ+2) x 129 tile of each of Ex, Ey at node
// level s and of Ez at cell level s, halos included) are staged in shared memory by
// cp.async (LDGSTS) three levels ahead through a 4-slot ring, so the DRAM latency is
// decoupled from the arithmetic.  Per level every thread computes the three faces it owns
// (Fx, Fy of cell level k, Fz of node level k) exactly once, publishes w*circ in shared
// memory and, after a barrier, gathers the faces around its three edges; the faces of the
// level below stay in registers; weights are register-prefetched one level ahead.
static const int EDGE_NS = 4;             // ring slots

template <typename T, bool SYM, int TI>
__global__ void __launch_bounds__(TI * 5, (TI <= 64 ? 2 : 1))
k_edge_op(MeshD m, const double* __restrict__ wF, const double* __restrict__ beta,
          const double* __restrict__ s_d, const T* __restrict__ x, T* __restrict__ y, int tj, int nchunk, int kz, int istride,
          int jtile0, int jtiles) {
    constexpr int TW = TI + 1;                         // staged tile width: lanes i0-1 .. i0+TI-1
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int rows = blockDim.y;                       // face rows: (halo j0-1) + tj   [SYM: 1]
    const int trows = SYM ? 1 : rows + 1;              // staged x rows: j0-1 .. j0+tj
    const int tsz = trows * TW;                        // elements per staged array
    const int ncomp = SYM ? 1 : 3;
    T* stage = reinterpret_cast<T*>(smem_raw);         // [EDGE_NS][ncomp][trows][TW]
    T* fxb = stage + (size_t)EDGE_NS * ncomp * tsz;    // [2][3][rows][TI] face exchange (double buffer)
    const int li = threadIdx.x, row = threadIdx.y;
    const int tid = row * TI + li, nthr = TI * rows;
    const int i0 = blockIdx.x * istride;
    const int i = i0 + li;
    const int sys = blockIdx.z / nchunk, chunk = blockIdx.z % nchunk;
    const int j0 = (((int)blockIdx.y + jtile0) % jtiles) * tj;      // (probing: only the theta tiles that carry the response)
    const int jraw = SYM ? 0 : j0 + row - 1;
    const bool valid = (i < m.nx) && (SYM || jraw < m.nth);
    const int j = SYM ? 0 : (jraw < 0 ? m.nth - 1 : (jraw >= m.nth ? 0 : jraw));
    const bool out_ok = valid && (SYM || row >= 1) && (li < istride || i + 1 >= m.nx);
    const bool has_ip = (i + 1 < m.nx);
    x += sys * m.nE;
    y += sys * m.nE;
    const T s = Sc<T>::from(s_d[2 * sys], s_d[2 * sys + 1]);
    const T* X = x; const T* Y = x + m.nEx; const T* Z = x + m.nEx + m.nEy;
    T* YX = y; T* YY = y + m.nEx; T* YZ = y + m.nEx + m.nEy;
    const double* wx = wF; const double* wy = wF + m.nFx; const double* wz = wF + m.nFx + m.nFy;
    const double* bx = beta; const double* by = beta + m.nEx; const double* bz = beta + m.nEx + m.nEy;
    const int nlev = m.nz + 1;
    const int k0 = chunk * kz, k1 = min(nlev, k0 + kz);
    const int kbeg = (k0 > 0) ? k0 - 1 : 0;            // one extra level below: its faces feed level k0
    const int ii = valid ? i : 0;
    const double lex = SYM ? 0.0 : m.lEx(ii), ley = m.lEy(ii, j), leym = (ii > 0) ? m.lEy(ii - 1, j) : 0.0;
    const long long col = (long long)ii + (long long)m.nx * j;
    const long long ezc = 1 + (long long)ii + (long long)m.nx * j;       // Ez offset inside a level block
    const T zero = Sc<T>::zero();

    // ---- staging descriptors: which tile elements this thread copies every level ----------
    constexpr int NQ = SYM ? 2 : 2;                     // ceil(tsz / nthr) <= 2 for every configuration used (checked on the host)
    int eoff[NQ]; long long cb[NQ]; long long zb[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        int e = tid + q * nthr;
        eoff[q] = -1; cb[q] = -1; zb[q] = -1;
        if (e < tsz) {
            eoff[q] = e;
            int rr = e / TW, ll = e - rr * TW;
            int gi = i0 - 1 + ll, gj = 0;
            bool jok = true;
            if (!SYM) {
                int jr = j0 - 1 + rr;
                jok = jr <= m.nth;                      // one wrapped row past the end is still needed (jp)
                gj = jr < 0 ? m.nth - 1 : (jr >= m.nth ? jr - m.nth : jr);
            }
            if (jok && gi >= 0 && gi < m.nx) { cb[q] = (long long)gi + (long long)m.nx * gj; zb[q] = 1 + cb[q]; }
            else if (jok && gi == -1) zb[q] = 0;        // the axis edge stands in for Ez(r_0)
        }
    }
    auto load_stage = [&](int sl) {
        if (sl > m.nz) return;                          // beyond the top node level: nothing to load
        T* st = stage + (size_t)(sl % EDGE_NS) * ncomp * tsz;
        const long long onode = m.nxy * sl, ocell = m.ezp * sl;
        const bool zlev = (!SYM) && sl < m.nz;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int e = eoff[q];
            if (e < 0) continue;
            if (cb[q] >= 0) {
                __pipeline_memcpy_async(st + e, Y + cb[q] + onode, sizeof(T));
                if (!SYM) __pipeline_memcpy_async(st + tsz + e, X + cb[q] + onode, sizeof(T));
            } else {
                st[e] = zero;
                if (!SYM) st[tsz + e] = zero;
            }
            if (zlev) {
                if (zb[q] >= 0) __pipeline_memcpy_async(st + 2 * tsz + e, Z + ocell + zb[q], sizeof(T));
                else st[2 * tsz + e] = zero;
            }
        }
    };
    load_stage(kbeg);     __pipeline_commit();
    load_stage(kbeg + 1); __pipeline_commit();
    load_stage(kbeg + 2); __pipeline_commit();

    T tFxp = zero, tFyp = zero;
    double wzk = 0.0, wxk = 0.0, wyk = 0.0, byk = 0.0, bxk = 0.0, bzk = 0.0;
    if (valid) {
        long long c = col + m.nxy * kbeg;
        wzk = wz[c];
        byk = by[c];
        if (!SYM) bxk = bx[c];
        if (kbeg < m.nz) { wxk = wx[c]; if (!SYM) { wyk = wy[c]; bzk = bz[kbeg * m.ezp + ezc]; } }
    }
    __pipeline_wait_prior(1);                            // stages kbeg, kbeg+1 landed
    __syncthreads();
    const int so = (SYM ? 0 : row) * TW + li;            // (row, lane-1) inside a staged array; +1 = own lane
    long long cnode = col + m.nxy * kbeg, ccell = (long long)kbeg * m.ezp + ezc;   // running offsets of level k
    for (int k = kbeg; k < k1; ++k, cnode += m.nxy, ccell += m.ezp) {
        const bool cell = (k < m.nz);
        load_stage(k + 3);
        __pipeline_commit();
        // prefetch next level's weights into registers
        double wzn = 0.0, wxn = 0.0, wyn = 0.0, byn = 0.0, bxn = 0.0, bzn = 0.0;
        if (valid && k + 1 < k1) {
            const long long c = cnode + m.nxy;
            wzn = wz[c];
            byn = by[c];
            if (!SYM) bxn = bx[c];
            if (k + 1 < m.nz) { wxn = wx[c]; if (!SYM) { wyn = wy[c]; bzn = bz[ccell + m.ezp]; } }
        }
        const T* sk = stage + (size_t)(k % EDGE_NS) * ncomp * tsz;
        const T* sk1 = stage + (size_t)((k + 1) % EDGE_NS) * ncomp * tsz;
        T* fx = fxb + (size_t)(k & 1) * 3 * rows * TI;
        T tFx = zero, tFy = zero, tFz = zero, Xk = zero, Zk = zero;
        const double lez = (!SYM && cell) ? m.lEz(k) : 0.0;
        T Yk = sk[so + 1];
        if (valid) {
            T Yim = sk[so];
            T cfz = ley * Yk - leym * Yim;
            if (!SYM) {
                Xk = sk[tsz + so + 1];
                cfz -= lex * (sk[tsz + so + TW + 1] - Xk);
            }
            tFz = wzk * cfz;
            if (cell) {
                T Yk1 = sk1[so + 1];
                T cfx = (-ley) * (Yk1 - Yk);
                if (!SYM) {
                    T Xk1 = sk1[tsz + so + 1];
                    Zk = sk[2 * tsz + so + 1];
                    T Zjp = sk[2 * tsz + so + TW + 1];
                    T Zim = sk[2 * tsz + so];
                    cfx += lez * (Zjp - Zk);
                    tFy = wyk * (lex * (Xk1 - Xk) - lez * (Zk - Zim));
                }
                tFx = wxk * cfx;
            }
        }
        fx[(0 * rows + row) * TI + li] = tFx;
        fx[(1 * rows + row) * TI + li] = tFy;
        fx[(2 * rows + row) * TI + li] = tFz;
        // the one barrier of the level: publishes the faces AND (thanks to the wait) the
        // staged tiles that the next level reads
        __pipeline_wait_prior(1);                        // stages <= k+2 landed
        __syncthreads();
        if (out_ok && k >= k0) {
            const long long c = cnode;
            T tFz_ip = has_ip ? fx[(2 * rows + row) * TI + li + 1] : zero;
            {
                T r = tFx - tFxp + tFz - tFz_ip;
                YY[c] = ley * r + s * (byk * Yk);
            }
            if (!SYM) {
                T tFz_jm = fx[(2 * rows + row - 1) * TI + li];
                T r = tFyp - tFy + tFz - tFz_jm;
                YX[c] = lex * r + s * (bxk * Xk);
                if (cell) {
                    T tFx_jm = fx[(0 * rows + row - 1) * TI + li];
                    T tFy_ip = has_ip ? fx[(1 * rows + row) * TI + li + 1] : zero;
                    T r2 = tFx_jm - tFx - tFy + tFy_ip;
                    YZ[ccell] = lez * r2 + s * (bzk * Zk);
                }
            }
        }
        tFxp = tFx; tFyp = tFy;
        wzk = wzn; wxk = wxn; wyk = wyn; byk = byn; bxk = bxn; bzk = bzn;
    }
    __pipeline_wait_prior(0);
}

// the single shared axis edge per z level: lEz * sum_j wFy(0,j,k) circFy(0,j,k) + s beta x
template <typename T>
__global__ void k_edge_op_axis(MeshD m, const double* __restrict__ wF, const double* __restrict__ beta,
                               const double* __restrict__ s_d, const T* __restrict__ x, T* __restrict__ y) {
    int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31, sys = blockIdx.y;
    if (k >= m.nz) return;
    x += sys * m.nE;
    y += sys * m.nE;
    const T* X = x; const T* Z = x + m.nEx + m.nEy;
    const double* wy = wF + m.nFx;
    T az = Z[m.ezaxis(k)];
    double re = 0.0, im = 0.0;
    for (int j = lane; j < m.nth; j += 32) {
        long long c = m.cidx(0, j, k);
        T v = wy[c] * (m.lEx(0) * (X[c + m.nxy] - X[c]) - m.lEz(k) * (Z[m.ez(0, j, k)] - az));
        re += Sc<T>::re(v); im += Sc<T>::im(v);
    }
    for (int o = 16; o > 0; o >>= 1) { re += __shfl_down_sync(0xffffffffu, re, o); im += __shfl_down_sync(0xffffffffu, im, o); }
    if (lane == 0) {
        T s = Sc<T>::from(s_d[2 * sys], s_d[2 * sys + 1]);
        long long a = m.nEx + m.nEy + m.ezaxis(k);
        y[a] = m.lEz(k) * Sc<T>::from(re, im) + s * (beta[a] * az);
    }
}

template <typename T, bool SYM, int TI>
static int launch_edge_op_cfg(const MeshD& m, const double* wF, const double* beta, const double* s_d,
                              const T* x, T* y, int nsys, cudaStream_t st, int jc0 = 0, int jcn = 0) {
    static int nsm = 0;
    if (!nsm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev); if (nsm <= 0) nsm = 148; }
    int nlev = m.nz + 1;
    int itiles = (m.nx <= TI) ? 1 : (m.nx + TI - 2) / (TI - 1);
    int istride = (itiles == 1) ? TI : TI - 1;
    int tj = SYM ? 1 : (m.nth < 4 ? m.nth : 4);
    int jtiles = SYM ? 1 : (m.nth + tj - 1) / tj;
    // outputs wanted for the theta columns jc0 .. jc0 + jcn - 1 only (jcn <= 0: all): launch the tiles that hold them
    int jt0 = 0, jtn = jtiles;
    if (!SYM && jcn > 0 && jcn < m.nth) {
        int a = ((jc0 % m.nth) + m.nth) % m.nth;
        jt0 = a / tj;
        jtn = std::min(jtiles, (a + jcn - 1) / tj - jt0 + 1);
    }
    // z chunks: ~2 waves of resident CTAs when the mesh is large enough, at least ~16 levels
    // per chunk so the extra priming level stays cheap
    int resident = (TI <= 64) ? 2 : 1;
    long long plane_ctas = (long long)itiles * jtn * nsys;
    int nchunk = (int)((2LL * nsm * resident + plane_ctas - 1) / plane_ctas);
    int maxchunk = (nlev + 15) / 16;
    if (nchunk > maxchunk) nchunk = maxchunk;
    if (nchunk < 1) nchunk = 1;
    int kz = (nlev + nchunk - 1) / nchunk;
    nchunk = (nlev + kz - 1) / kz;
    int rows = SYM ? 1 : tj + 1;
    int trows = SYM ? 1 : rows + 1, ncomp = SYM ? 1 : 3;
    size_t smb = ((size_t)EDGE_NS * ncomp * trows * (TI + 1) + 2 * 3 * (size_t)rows * TI) * sizeof(T);
    CS_CHECK(trows * (TI + 1) <= 2 * TI * rows, "staging descriptor count too small");
    dim3 blk(TI, rows), grd(itiles, jtn, nchunk * nsys);
    // the attribute is per device and not capturable into a CUDA graph: set it once per (instantiation, device)
    static bool attr_set[64] = {false};
    int dev = 0;
    CS_CUDA(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        size_t mx = ((size_t)EDGE_NS * 3 * 6 * (TI + 1) + 2 * 3 * 5 * TI) * sizeof(T);
        CS_CUDA(cudaFuncSetAttribute(k_edge_op<T, SYM, TI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx));
        if (dev < 64) attr_set[dev] = true;
    }
    k_edge_op<T, SYM, TI><<<grd, blk, smb, st>>>(m, wF, beta, s_d, x, y, tj, nchunk, kz, istride, jt0, jtiles);
    return CS_OK;
}

template <typename T>
static int launch_edge_op_t(const MeshD& m, const double* wF, const double* beta, const double* s_d,
                            const T* x, T* y, int nsys, cudaStream_t st, int jc0, int jcn) {
    if (m.sym) {
        CS_TRY((launch_edge_op_cfg<T, true, 128>(m, wF, beta, s_d, x, y, nsys, st)));
    } else {
        CS_TRY((launch_edge_op_cfg<T, false, 64>(m, wF, beta, s_d, x, y, nsys, st, jc0, jcn)));
        dim3 ga((m.nz + 7) / 8, nsys);
        k_edge_op_axis<T><<<ga, 256, 0, st>>>(m, wF, beta, s_d, x, y);
    }
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

int launch_edge_op(const MeshD& m, const double* wF, const double* beta, const double* s_d,
                   const void* x, void* y, int nsys, bool cplx_vec, cudaStream_t st, int jc0, int jcn) {
    if (cplx_vec) return launch_edge_op_t<cplx>(m, wF, beta, s_d, (const cplx*)x, (cplx*)y, nsys, st, jc0, jcn);
    return launch_edge_op_t<double>(m, wF, beta, s_d, (const double*)x, (double*)y, nsys, st, jc0, jcn);
}

// ============================ K4': fused face-form operator ======================
// y = alpha .* ( C diag(gamma) C^T (alpha .* x) + s x ),
//   y_f = uF_f * circ_f( gE .* d ) + s * uF_f * area_f * x_f,
//   d_e = sum_{f' around e} +- uF_f' x_f',  uF = alpha/area, gE = len^2 * gamma.
// Serves FDEM-j / TDEM-j (alpha = MfRho, gamma = MeMuI) and the b forms
// (alpha = MfMui, gamma = MeSigmaI).
template <typename T> struct UFace {
    const MeshD& m; FaceV<T> f; const double *ux, *uy, *uz;
    __device__ UFace(const MeshD& m_, const T* v, const double* u) : m(m_), f(m_, v) {
        ux = u; uy = u + m_.nFx; uz = u + m_.nFx + m_.nFy;
    }
    __device__ __forceinline__ T operator()(int kind, int i, int j, int k) const {
        long long c = m.cidx(i, j, k);
        if (kind == 0) return ux[c] * f.x[c];
        if (kind == 1) return uy[c] * f.y[c];
        return uz[c] * f.z[c];
    }
};
// gE_e * d_e for each edge kind, evaluated on the fly
template <typename T> struct GD {
    const MeshD& m; UFace<T> u; const double *gx, *gy, *gz;
    __device__ GD(const MeshD& m_, const T* v, const double* uF, const double* gE) : m(m_), u(m_, v, uF) {
        gx = gE; gy = gE + m_.nEx; gz = gE + m_.nEx + m_.nEy;
    }
    __device__ __forceinline__ T ex(int i, int j, int k) const { return gx[m.cidx(i, j, k)] * gathEx<T>(m, u, i, j, k); }
    __device__ __forceinline__ T ey(int i, int j, int k) const { return gy[m.cidx(i, j, k)] * gathEy<T>(m, u, i, j, k); }
    __device__ __forceinline__ T ez(int i, int j, int k) const { return gz[m.ez(i, j, k)] * gathEz<T>(m, u, i, j, k); }
    __device__ __forceinline__ T ezax(int k) const { return gz[m.ezaxis(k)] * gathEzAxis<T>(m, u, k); }
};

// edge values read back from an edge vector (second pass of the two-pass variant)
template <typename T> struct EdgeRead {
    const MeshD& m; const T *x, *y, *z;
    __device__ EdgeRead(const MeshD& m_, const T* e) : m(m_) { x = e; y = e + m_.nEx; z = e + m_.nEx + m_.nEy; }
    __device__ __forceinline__ T ex(int i, int j, int k) const { return x[m.cidx(i, j, k)]; }
    __device__ __forceinline__ T ey(int i, int j, int k) const { return y[m.cidx(i, j, k)]; }
    __device__ __forceinline__ T ez(int i, int j, int k) const { return z[m.ez(i, j, k)]; }
    __device__ __forceinline__ T ezax(int k) const { return z[m.ezaxis(k)]; }
};

// faces of slot (i, j, k) from the edge values d:  y_f = uF_f circ_f(d) + s uF_f area_f x_f
template <typename T, typename D>
__device__ __forceinline__ void face_op_slot(const MeshD& m, const D& d, const double* __restrict__ uF, T s,
                                             const T* __restrict__ x, T* __restrict__ y, int i, int j, int k) {
    long long c = m.cidx(i, j, k);
    if (k < m.nz) {
        // Fx: + Ez(j+1) - Ez(j) - Ey(k+1) + Ey(k)
        T r = d.ey(i, j, k) - d.ey(i, j, k + 1);
        if (!m.sym) r += d.ez(i, m.jp(j), k) - d.ez(i, j, k);
        double u = uF[c];
        y[c] = u * r + s * ((u * m.aFx(i, j, k)) * x[c]);
        if (!m.sym) {
            // Fy: + Ex(k+1) - Ex(k) - Ez(i) + Ez(i-1 | axis)
            T zin = (i > 0) ? d.ez(i - 1, j, k) : d.ezax(k);
            T r2 = d.ex(i, j, k + 1) - d.ex(i, j, k) - d.ez(i, j, k) + zin;
            double u2 = uF[m.nFx + c];
            y[m.nFx + c] = u2 * r2 + s * ((u2 * m.aFy(i, k)) * x[m.nFx + c]);
        }
    }
    // Fz: + Ey(i) - Ey(i-1) - Ex(j+1) + Ex(j)
    T r3 = d.ey(i, j, k);
    if (i > 0) r3 -= d.ey(i - 1, j, k);
    if (!m.sym) r3 -= d.ex(i, m.jp(j), k) - d.ex(i, j, k);
    long long fz = m.nFx + m.nFy + c;
    double u3 = uF[fz];
    y[fz] = u3 * r3 + s * ((u3 * m.aFz(i, j)) * x[fz]);
}

// single pass: every edge value is recomputed by each of the faces around it (12 edge evaluations of 4
// face loads each per slot) -- no workspace
template <typename T>
__global__ void __launch_bounds__(TPB)
k_face_op(MeshD m, const double* __restrict__ uF, const double* __restrict__ gE,
          const double* __restrict__ s_d, const T* __restrict__ x, T* __restrict__ y) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    Slot sl;
    if (!slot_of(m, t, m.nz + 1, sl)) return;
    int sys = blockIdx.y;
    x += sys * m.nF;
    y += sys * m.nF;
    T s = Sc<T>::from(s_d[2 * sys], s_d[2 * sys + 1]);
    GD<T> d(m, x, uF, gE);
    face_op_slot<T, GD<T>>(m, d, uF, s, x, y, sl.i, sl.j, sl.k);
}

// two passes through an edge workspace: (1) e = gE .* C^T (uF .* x), every edge once;
// (2) the faces from e.  3 + 12 gathers per slot instead of 48, for 48 B/cell (real) of extra traffic.
template <typename T>
__global__ void __launch_bounds__(TPB)
k_face_op_edges(MeshD m, const double* __restrict__ uF, const double* __restrict__ gE,
                const T* __restrict__ x, T* __restrict__ e) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    Slot sl;
    if (!slot_of(m, t, m.nz + 1, sl)) return;
    int sys = blockIdx.y;
    x += sys * m.nF;
    e += sys * m.nE;
    int i = sl.i, j = sl.j, k = sl.k;
    GD<T> d(m, x, uF, gE);
    long long c = m.cidx(i, j, k);
    e[m.nEx + c] = d.ey(i, j, k);
    if (!m.sym) {
        e[c] = d.ex(i, j, k);
        if (k < m.nz) {
            e[m.nEx + m.nEy + m.ez(i, j, k)] = d.ez(i, j, k);
            if (i == 0 && j == 0) e[m.nEx + m.nEy + m.ezaxis(k)] = d.ezax(k);
        }
    }
}
template <typename T>
__global__ void __launch_bounds__(TPB)
k_face_op_faces(MeshD m, const double* __restrict__ uF, const double* __restrict__ s_d,
                const T* __restrict__ x, const T* __restrict__ e, T* __restrict__ y) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    Slot sl;
    if (!slot_of(m, t, m.nz + 1, sl)) return;
    int sys = blockIdx.y;
    x += sys * m.nF;
    y += sys * m.nF;
    e += sys * m.nE;
    T s = Sc<T>::from(s_d[2 * sys], s_d[2 * sys + 1]);
    EdgeRead<T> d(m, e);
    face_op_slot<T, EdgeRead<T>>(m, d, uF, s, x, y, sl.i, sl.j, sl.k);
}

// ---- single pass, full theta ring per CTA, z march ("ring" kernel; 3-D meshes with nth * 16 <= 512) ----
// The two-pass variant moves 192 B/cell (real) for 96 algorithmic ones.  Here a CTA owns TW - 2 = 14 radial cells x ALL nth azimuthal
// cells (so theta needs no halo: the neighbours j +- 1 wrap inside the tile; the first and last lane are the radial halo) and marches
// ZC node levels.  Per level every thread loads its own three faces / weights once (register-prefetched one level
// ahead), publishes u = uF x in shared memory, the edge values gE * (signed sum of the four u around the edge) are
// computed once per edge and published, and the faces are gathered from them; the values of the level below that the
// Fx / Fy faces need (u, the node-level edges, the Ez edges of the cell level) stay in registers.  Two barriers per level,
// no workspace in HBM.  The shared axis edge (sum over all theta of the innermost Fy faces) comes from a one-warp-per-level
// pre-kernel.
template <typename T> __device__ __forceinline__ T shfl_down_t(T v, int o);
template <> __device__ __forceinline__ double shfl_down_t<double>(double v, int o) { return __shfl_down_sync(0xffffffffu, v, o); }
template <> __device__ __forceinline__ cplx shfl_down_t<cplx>(cplx v, int o) {
    return mk(__shfl_down_sync(0xffffffffu, v.x, o), __shfl_down_sync(0xffffffffu, v.y, o));
}
template <typename T>
__global__ void k_face_axis(MeshD m, const double* __restrict__ uF, const double* __restrict__ gE, const T* __restrict__ x, T* __restrict__ axis) {
    const int k = blockIdx.x, sys = blockIdx.y;
    x += (size_t)sys * m.nF;
    const double* uy = uF + m.nFx;
    T r = Sc<T>::zero();
    for (int j = threadIdx.x; j < m.nth; j += 32) {
        long long c = m.cidx(0, j, k);
        r += uy[c] * x[m.nFx + c];
    }
    for (int o = 16; o; o >>= 1) r += shfl_down_t<T>(r, o);
    if (threadIdx.x == 0) axis[(size_t)sys * m.nz + k] = gE[m.nEx + m.nEy + m.ezaxis(k)] * r;
}

template <typename T, int ZC, int TW>
__global__ void __launch_bounds__(512)
k_face_op_ring(MeshD m, const double* __restrict__ uF, const double* __restrict__ gE, const double* __restrict__ s_d,
               const T* __restrict__ x, T* __restrict__ y, const T* __restrict__ axis) {
    extern __shared__ __align__(16) unsigned char face_sm[];
    const int nth = blockDim.y;
    const int tsz = nth * TW;
    T* U = reinterpret_cast<T*>(face_sm);          // [3][nth][TW]  u of Fx, Fy (cell level n), Fz (node level n)
    T* E = U + 3 * tsz;                            // [3][nth][TW]  gE * d of Ex, Ey (node level n), Ez (cell level n)
    const int l = threadIdx.x, j = threadIdx.y;
    const int i = (int)blockIdx.x * (TW - 2) - 1 + l;
    const bool in = (i >= 0 && i < m.nx);
    const bool edge_ok = in && l <= TW - 2;
    const bool out_ok = in && l >= 1 && l <= TW - 2;
    const bool has_ip = (i + 1 < m.nx);
    const int sys = blockIdx.z;
    x += (size_t)sys * m.nF;
    y += (size_t)sys * m.nF;
    axis += (size_t)sys * m.nz;
    const T s = Sc<T>::from(s_d[2 * sys], s_d[2 * sys + 1]);
    const T zero = Sc<T>::zero();
    const int k0 = (int)blockIdx.y * ZC, k1 = min(m.nz, k0 + ZC);
    const T* X = x; const T* Yf = x + m.nFx; const T* Z = x + m.nFx + m.nFy;
    T* OX = y; T* OY = y + m.nFx; T* OZ = y + m.nFx + m.nFy;
    const double* wxp_ = uF; const double* wyp_ = uF + m.nFx; const double* wzp_ = uF + m.nFx + m.nFy;
    const double* gx = gE; const double* gy = gE + m.nEx; const double* gz = gE + m.nEx + m.nEy;
    const long long col = (long long)(in ? i : 0) + (long long)m.nx * j;
    const int me = j * TW + l, mjm = ((j == 0) ? nth - 1 : j - 1) * TW + l, mjp = ((j + 1 == nth) ? 0 : j + 1) * TW + l;
    const double afz = in ? m.aFz(i, j) : 0.0;

    // values of the level below (cell level n-1 / node level n-1)
    T uxp = zero, uyp = zero, xxp = zero, xyp = zero, exq = zero, eyq = zero, ezq = zero, ezjp_q = zero, ezim_q = zero;
    double wxq = 0.0, wyq = 0.0;
    if (in && k0 > 0) {
        const long long c = col + m.nxy * (long long)(k0 - 1);
        uxp = wxp_[c] * X[c];
        uyp = wyp_[c] * Yf[c];
    }
    // register prefetch of level n
    T cxx = zero, cxy = zero, cxz = zero;
    double cwx = 0.0, cwy = 0.0, cwz = 0.0, cgx = 0.0, cgy = 0.0, cgz = 0.0;
    auto fetch = [&](int n, T& fx_, T& fy_, T& fz_, double& wx_, double& wy_, double& wz_, double& gx_, double& gy_, double& gz_) {
        fx_ = zero; fy_ = zero; fz_ = zero; wx_ = wy_ = wz_ = gx_ = gy_ = gz_ = 0.0;
        if (!in || n > k1) return;
        const long long c = col + m.nxy * (long long)n;
        fz_ = Z[c]; wz_ = wzp_[c]; gx_ = gx[c]; gy_ = gy[c];
        if (n < m.nz) { fx_ = X[c]; fy_ = Yf[c]; wx_ = wxp_[c]; wy_ = wyp_[c]; gz_ = gz[(long long)n * m.ezp + 1 + col]; }
    };
    fetch(k0, cxx, cxy, cxz, cwx, cwy, cwz, cgx, cgy, cgz);
    for (int n = k0; n <= k1; ++n) {
        const bool cell = n < m.nz;
        T nxx, nxy_, nxz; double nwx, nwy, nwz, ngx, ngy, ngz;
        fetch(n + 1, nxx, nxy_, nxz, nwx, nwy, nwz, ngx, ngy, ngz);
        const T ux = cwx * cxx, uy = cwy * cxy, uz = cwz * cxz;        // (zero above the top cell level / outside the mesh)
        U[me] = ux; U[tsz + me] = uy; U[2 * tsz + me] = uz;
        __syncthreads();
        T ex = zero, ey = zero, ez = zero;
        if (edge_ok) {
            const T uz_ip = has_ip ? U[2 * tsz + me + 1] : zero;
            ex = cgx * (uz - U[2 * tsz + mjm] - uy + uyp);
            ey = cgy * (uz - uz_ip + ux - uxp);
            if (cell) {
                const T uy_ip = has_ip ? U[tsz + me + 1] : zero;
                ez = cgz * (U[mjm] - ux - uy + uy_ip);
            }
        }
        E[me] = ex; E[tsz + me] = ey; E[2 * tsz + me] = ez;
        __syncthreads();
        if (out_ok) {
            const long long c = col + m.nxy * (long long)n;
            if (n < k1 || n == m.nz) {
                // Fz of node level n
                T r3 = ey - E[mjp] + ex;
                if (i > 0) r3 -= E[tsz + me - 1];
                OZ[c] = cwz * r3 + s * ((cwz * afz) * cxz);
            }
            if (n > k0) {
                // Fx, Fy of cell level n - 1
                const long long cb = c - m.nxy;
                const T r = eyq - ey + ezjp_q - ezq;
                OX[cb] = wxq * r + s * ((wxq * m.aFx(i, j, n - 1)) * xxp);
                const T r2 = ex - exq - ezq + ezim_q;
                OY[cb] = wyq * r2 + s * ((wyq * m.aFy(i, n - 1)) * xyp);
            }
            if (cell) {
                ezjp_q = E[2 * tsz + mjp];
                ezim_q = (i > 0) ? E[2 * tsz + me - 1] : axis[n];
            }
        }
        uxp = ux; uyp = uy; xxp = cxx; xyp = cxy; wxq = cwx; wyq = cwy; exq = ex; eyq = ey; ezq = ez;
        cxx = nxx; cxy = nxy_; cxz = nxz; cwx = nwx; cwy = nwy; cwz = nwz; cgx = ngx; cgy = ngy; cgz = ngz;
    }
}

// ework: nsys * nE elements of workspace, or nullptr for the single-pass kernel
int launch_face_op(const MeshD& m, const double* uF, const double* gE, const double* s_d,
                   const void* x, void* y, int nsys, bool cplx_vec, void* ework, cudaStream_t st) {
    long long n = m.nxy * (m.nz + 1);
    // large 3-D meshes whose theta ring fits one CTA: the single-pass ring kernel (CS_FACE_OP=twopass: the two-pass one)
    static const bool twopass = [] { const char* e = getenv("CS_FACE_OP"); return e && !strcmp(e, "twopass"); }();
    if (ework && !twopass && !m.sym && m.nth >= 4 && m.nth * 16 <= 512 && m.nz >= 32) {
        constexpr int TW = 16;
        static const int zc = [] { const char* e = getenv("CS_FACE_ZC"); int v = e ? atoi(e) : 16; return (v == 8 || v == 32) ? v : 16; }();
        const size_t es = cplx_vec ? sizeof(cplx) : sizeof(double);
        const size_t sm = (size_t)6 * m.nth * TW * es;
        dim3 blk(TW, m.nth), grd((unsigned)((m.nx + TW - 3) / (TW - 2)), (unsigned)((m.nz + zc - 1) / zc), (unsigned)nsys), ga((unsigned)m.nz, (unsigned)nsys);
        auto run = [&](auto kern_c, auto kern_d) -> int {
            if (cplx_vec) {
                CS_CUDA(cudaFuncSetAttribute(kern_c, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                k_face_axis<cplx><<<ga, 32, 0, st>>>(m, uF, gE, (const cplx*)x, (cplx*)ework);
                kern_c<<<grd, blk, sm, st>>>(m, uF, gE, s_d, (const cplx*)x, (cplx*)y, (const cplx*)ework);
            } else {
                CS_CUDA(cudaFuncSetAttribute(kern_d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                k_face_axis<double><<<ga, 32, 0, st>>>(m, uF, gE, (const double*)x, (double*)ework);
                kern_d<<<grd, blk, sm, st>>>(m, uF, gE, s_d, (const double*)x, (double*)y, (const double*)ework);
            }
            return CS_OK;
        };
        if (zc == 8) CS_TRY(run(k_face_op_ring<cplx, 8, TW>, k_face_op_ring<double, 8, TW>));
        else if (zc == 32) CS_TRY(run(k_face_op_ring<cplx, 32, TW>, k_face_op_ring<double, 32, TW>));
        else CS_TRY(run(k_face_op_ring<cplx, 16, TW>, k_face_op_ring<double, 16, TW>));
        CS_CUDA(cudaGetLastError());
        return CS_OK;
    }
    if (ework) {
        if (cplx_vec) {
            k_face_op_edges<cplx><<<grid1(n, nsys), TPB, 0, st>>>(m, uF, gE, (const cplx*)x, (cplx*)ework);
            k_face_op_faces<cplx><<<grid1(n, nsys), TPB, 0, st>>>(m, uF, s_d, (const cplx*)x, (const cplx*)ework, (cplx*)y);
        } else {
            k_face_op_edges<double><<<grid1(n, nsys), TPB, 0, st>>>(m, uF, gE, (const double*)x, (double*)ework);
            k_face_op_faces<double><<<grid1(n, nsys), TPB, 0, st>>>(m, uF, s_d, (const double*)x, (const double*)ework, (double*)y);
        }
    } else if (cplx_vec) k_face_op<cplx><<<grid1(n, nsys), TPB, 0, st>>>(m, uF, gE, s_d, (const cplx*)x, (cplx*)y);
    else k_face_op<double><<<grid1(n, nsys), TPB, 0, st>>>(m, uF, gE, s_d, (const double*)x, (double*)y);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// ============================ CSR SpMV (Solver(A) plugin path) =====================
// y = A x for an assembled matrix handed over through the SimPEG solver protocol
// (run.py:246, 294, 343).  One thread per row: the systems have <= 33 nonzeros per row.
template <typename TA, typename TX>
__global__ void k_csr_spmv(long long n, const long long* __restrict__ indptr, const int* __restrict__ indices,
                           const TA* __restrict__ vals, const TX* __restrict__ x, TX* __restrict__ y) {
    long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r >= n) return;
    x += blockIdx.y * n;
    y += blockIdx.y * n;
    TX acc = Sc<TX>::zero();
    for (long long e = indptr[r]; e < indptr[r + 1]; ++e) acc += vals[e] * x[indices[e]];
    y[r] = acc;
}
int launch_csr_spmv(long long n, const long long* indptr, const int* indices, const void* vals, bool a_cplx,
                    const void* x, void* y, int nsys, bool x_cplx, cudaStream_t st) {
    dim3 g((unsigned)cdiv(n, TPB), nsys);
    if (a_cplx && x_cplx) k_csr_spmv<cplx, cplx><<<g, TPB, 0, st>>>(n, indptr, indices, (const cplx*)vals, (const cplx*)x, (cplx*)y);
    else if (!a_cplx && x_cplx) k_csr_spmv<double, cplx><<<g, TPB, 0, st>>>(n, indptr, indices, (const double*)vals, (const cplx*)x, (cplx*)y);
    else if (!a_cplx && !x_cplx) k_csr_spmv<double, double><<<g, TPB, 0, st>>>(n, indptr, indices, (const double*)vals, (const double*)x, (double*)y);
    else CS_CHECK(false, "complex matrix needs complex vectors");
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

}  // namespace cs
