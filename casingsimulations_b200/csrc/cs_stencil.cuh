// Matrix-free cylindrical finite-volume stencils (K1-K4, K8 of DESIGN.md).
//
// Restates discretize.CylMesh.{faceDiv, edgeCurl, getFaceInnerProduct,
// getEdgeInnerProduct} (reference call sites: casingSimulations/mesh.py:518,563;
// run.py:240-248, 287-296, 339-344) as fused per-point kernels.  One thread per
// (i, j, k) slot, i fastest, so every load of a warp is a contiguous run of r.
#pragma once
#include "cs_common.cuh"

namespace cs {

// ---------------------------------------------------------------- helpers ---
struct Slot { int i, j, k; };

__device__ __forceinline__ bool slot_of(const MeshD& m, long long t, int nlev, Slot& s) {
    long long tot = m.nxy * nlev;
    if (t >= tot) return false;
    s.i = (int)(t % m.nx);
    long long q = t / m.nx;
    s.j = (int)(q % m.nth);
    s.k = (int)(q / m.nth);
    return true;
}

// views of the three edge / face blocks of one vector
template <typename T> struct EdgeV {
    const T *x, *y, *z;   // Ex, Ey, Ez blocks (x, z null on a symmetric mesh)
    __device__ __forceinline__ EdgeV(const MeshD& m, const T* v) {
        if (m.sym) { x = nullptr; y = v; z = nullptr; }
        else { x = v; y = v + m.nEx; z = v + m.nEx + m.nEy; }
    }
};
template <typename T> struct FaceV {
    const T *x, *y, *z;   // Fx, Fy, Fz blocks (y null on a symmetric mesh)
    __device__ __forceinline__ FaceV(const MeshD& m, const T* v) {
        x = v; y = m.sym ? nullptr : v + m.nFx; z = v + m.nFx + m.nFy;
    }
};

// ---- circulation of (len * e) around each face: (C_topo diag(len) e)_f ------
template <typename T>
__device__ __forceinline__ T circFx(const MeshD& m, const EdgeV<T>& e, int i, int j, int k) {
    long long c0 = m.cidx(i, j, k), c1 = m.cidx(i, j, k + 1);
    T r = (-m.lEy(i, j)) * (e.y[c1] - e.y[c0]);
    if (!m.sym) r += m.lEz(k) * (e.z[m.ez(i, m.jp(j), k)] - e.z[m.ez(i, j, k)]);
    return r;
}
template <typename T>
__device__ __forceinline__ T circFy(const MeshD& m, const EdgeV<T>& e, int i, int j, int k) {
    T zin = (i > 0) ? e.z[m.ez(i - 1, j, k)] : e.z[m.ezaxis(k)];
    return m.lEx(i) * (e.x[m.cidx(i, j, k + 1)] - e.x[m.cidx(i, j, k)]) - m.lEz(k) * (e.z[m.ez(i, j, k)] - zin);
}
template <typename T>
__device__ __forceinline__ T circFz(const MeshD& m, const EdgeV<T>& e, int i, int j, int k) {
    long long c = m.cidx(i, j, k);
    T r = m.lEy(i, j) * e.y[c];
    if (i > 0) r -= m.lEy(i - 1, j) * e.y[c - 1];
    if (!m.sym) r -= m.lEx(i) * (e.x[m.cidx(i, m.jp(j), k)] - e.x[c]);
    return r;
}

// ---- (C_topo^T f)_e : signed sum of the faces around each edge -------------
// f is accessed through a functor F(kind, i, j, k) so the same topology code
// serves plain vectors and fused "weight * circulation" expressions.
template <typename T, typename F>
__device__ __forceinline__ T gathEx(const MeshD& m, F f, int i, int j, int k) {
    T r = f(2, i, j, k) - f(2, i, m.jm(j), k);
    if (k < m.nz) r -= f(1, i, j, k);
    if (k > 0) r += f(1, i, j, k - 1);
    return r;
}
template <typename T, typename F>
__device__ __forceinline__ T gathEy(const MeshD& m, F f, int i, int j, int k) {
    T r = f(2, i, j, k);
    if (i + 1 < m.nx) r -= f(2, i + 1, j, k);
    if (k < m.nz) r += f(0, i, j, k);
    if (k > 0) r -= f(0, i, j, k - 1);
    return r;
}
template <typename T, typename F>
__device__ __forceinline__ T gathEz(const MeshD& m, F f, int i, int j, int k) {
    T r = f(0, i, m.jm(j), k) - f(0, i, j, k) - f(1, i, j, k);
    if (i + 1 < m.nx) r += f(1, i + 1, j, k);
    return r;
}
template <typename T, typename F>
__device__ __forceinline__ T gathEzAxis(const MeshD& m, F f, int k) {
    T r = Sc<T>::zero();
    for (int j = 0; j < m.nth; ++j) r += f(1, 0, j, k);
    return r;
}

}  // namespace cs
