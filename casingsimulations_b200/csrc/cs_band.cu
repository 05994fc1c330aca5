// K5: theta-DFT + batched banded direct solves -- the preconditioner.
//
// For a theta-invariant model on uniform h_theta every operator of the path is
// block-circulant in theta, so a DFT along j decouples it into nth independent
// 2-D (r, z) systems.  In "band ordering" (z level slowest, r next, component
// fastest) each is banded with half bandwidth p ~ (#components * nx); it is
// factored exactly by a blocked right-looking LU without pivoting (the systems
// are symmetric positive definite, or complex symmetric with a definite
// imaginary part) and reused for every right-hand side / time step that shares
// the shift s.  The single shared axis Ez edge is theta-independent: it lives in
// the mode-0 system only.
//
// Band storage is LAPACK-style column-major: A(i,j) -> AB[(pw + i - j) + j*ldab],
// |i-j| <= pw = p + nb.  Seen through base = AB + pw with leading dimension
// ld = ldab - 1 every block inside the band is a plain dense column-major block,
// which is what the panel / update / solve kernels use.
#include "cs_internal.h"
#include <cuda_pipeline.h>
#include <vector>
#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>
#include <cstdlib>
#include <cstdio>
#include <cstring>

namespace cs {

template <typename T> __device__ __forceinline__ T shfl_t(T v, int src);
template <> __device__ __forceinline__ double shfl_t<double>(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <> __device__ __forceinline__ cplx shfl_t<cplx>(cplx v, int src) {
    return mk(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
// broadcast within each half-warp group of 16 lanes (lane r <-> row r of a 16 x 16 block)
template <typename T> __device__ __forceinline__ T shfl_row(T v, int src) { return shfl_t<T>(v, src); }

static const int TPB = 128;

// Host-side state shared by all contexts (graph cache, capture streams, scratch buffers) is guarded
// by one mutex; the small device scratch buffers are per (purpose, device, stream), so independent
// contexts (one stream each) never share one.
static std::mutex g_band_mu;
struct BandScratchBuf { void* p = nullptr; size_t bytes = 0; };
static std::map<std::tuple<int, int, cudaStream_t>, BandScratchBuf>& band_scratch_map() {
    static std::map<std::tuple<int, int, cudaStream_t>, BandScratchBuf> bufs;
    return bufs;
}
static int band_scratch(int which, cudaStream_t st, size_t need, bool zero_new, void** out) {
    int dev = 0;
    CS_CUDA(cudaGetDevice(&dev));
    BandScratchBuf& b = band_scratch_map()[std::make_tuple(which, dev, st)];
    if (need > b.bytes) {
        if (b.p) cudaFree(b.p);                  // (synchronises with everything still using it)
        b.p = nullptr; b.bytes = 0;
        if (need < (4u << 20)) need = 4u << 20;
        CS_CUDA(cudaMalloc(&b.p, need));
        if (zero_new) CS_CUDA(cudaMemset(b.p, 0, need));
        b.bytes = need;
    }
    *out = b.p;
    return CS_OK;
}
// a context that goes away takes the scratch buffers of its stream with it
void band_release_stream(cudaStream_t st) {
    std::lock_guard<std::mutex> lock(g_band_mu);
    auto& bufs = band_scratch_map();
    for (auto it = bufs.begin(); it != bufs.end();) {
        if (std::get<2>(it->first) == st) { if (it->second.p) cudaFree(it->second.p); it = bufs.erase(it); }
        else ++it;
    }
}

// ------------------------------------------------------------------ DFT -----
// Rotated theta-mode basis.  For mode m (phi = 2 pi m / nth) the mode matrix K_m of a theta-invariant operator is
// Hermitian; the only complex entries couple the component that is staggered by half a cell in theta (Ey against
// Ex / Ez, Fy against Fx / Fz) to the others, with phase (exp(+-i phi) - 1) = +-2i sin(phi/2) exp(+-i phi/2).
// With T_m = diag(1 | i exp(i hs phi / 2)) (hs = +-1: the stagger in half cells)
//     K''_m = T_m^* K_m T_m   is REAL symmetric   and   K''_{nth-m} = K''_m
// (checked in numpy: tools/proto/mode_similarity.py), so only modes 0 .. nth/2 are stored and factored, in real
// arithmetic when the shift is real (DC, TDEM) -- half the systems, half the bytes per entry.  The rotation costs
// nothing: it turns the DFT twiddle exp(-2 pi i j m / nth) into (-i) exp(-i pi m (2 j + hs) / nth).
//   gather : xm[v][m][q]  = conj(T_m[q]) sum_j x3d[v][map3d[q] + j*jstr[q]] exp(-2 pi i j m / nth)
//   scatter: x3d[..j..]   = (1/nth) sum_m T_m[q] xm[v][m][q] exp(+2 pi i j m / nth)
// mode_sel 0: all nth modes, complex (complex factors; mode m is solved with the factor of min(m, nth - m));
//          1: modes 0 .. nmodes-1, complex (probing);
//          2: modes 0 .. nmodes-1 of a REAL 3-D vector split into (Re, Im) real vectors (real factors; the other
//             modes are complex conjugates in the rotated basis as well).
// A block handles DFT_QB unknowns x all modes: the theta line of every unknown is staged once in shared memory,
// the 2 nth twiddles exp(-+ i pi t / nth) are tabulated per block (no sincospi in the inner loop).
static const int DFT_QB = 32, DFT_MY = 8;
__device__ __forceinline__ cplx to_c(double v) { return mk(v, 0.0); }
__device__ __forceinline__ cplx to_c(cplx v) { return v; }
template <typename T> __device__ __forceinline__ T from_c(cplx v);
template <> __device__ __forceinline__ double from_c<double>(cplx v) { return v.x; }
template <> __device__ __forceinline__ cplx from_c<cplx>(cplx v) { return v; }

template <typename TX, typename TM>
__global__ void __launch_bounds__(DFT_QB * DFT_MY)
k_dft_gather(const long long* __restrict__ map3d, const int* __restrict__ jstr, const signed char* __restrict__ hs,
             long long n2, int nth, int nfac, int mode_sel, const TX* __restrict__ x3d, long long n3d,
             TM* __restrict__ xm, int jshift, int thin) {
    extern __shared__ __align__(16) unsigned char dft_sm[];
    cplx* tw = reinterpret_cast<cplx*>(dft_sm);        // [2 nth]: exp(-i pi t / nth)
    cplx* xs = tw + 2 * nth;                           // [nj][DFT_QB]
    const int tq = threadIdx.x, ty = threadIdx.y;
    const int nj = thin ? 3 : nth;
    const int n2t = 2 * nth;
    for (int t = ty * DFT_QB + tq; t < n2t; t += DFT_QB * DFT_MY) {
        double sn, cs_;
        sincospi(-(double)t / (double)nth, &sn, &cs_);
        tw[t] = mk(cs_, sn);
    }
    const long long q = blockIdx.x * (long long)DFT_QB + tq;
    const int v = blockIdx.y;
    x3d += (size_t)v * n3d;
    const int nmv = mode_sel == 0 ? nth : (mode_sel == 1 ? nfac : 2 * nfac);
    xm += (size_t)v * nmv * n2;
    long long b = -1; int st = 0, h2 = 0;
    if (q < n2) { b = map3d[q]; st = jstr[q]; h2 = hs[q]; }
    for (int jj = ty; jj < nj; jj += DFT_MY) {
        cplx val = mk(0.0, 0.0);
        if (b >= 0 && st != 0) {
            // jshift: theta column that plays the role of column 0; thin: only columns jshift-1 .. jshift+1 carry data
            int col = (jshift + jj - (thin ? 1 : 0)) % nth;
            if (col < 0) col += nth;
            val = to_c(x3d[b + (long long)col * st]);
        }
        xs[jj * DFT_QB + tq] = val;
    }
    __syncthreads();
    if (q >= n2) return;
    const int nm = mode_sel == 0 ? nth : nfac;
    for (int m = ty; m < nm; m += DFT_MY) {
        cplx acc = mk(0.0, 0.0);
        if (b >= 0) {
            // the shared axis unknown lives in mode 0 only.  Its row of the mode-0 system sums nth theta columns and its column
            // reaches nth of them: scaled by sqrt(nth) here and by 1 / sqrt(nth) in the scatter the system stays SYMMETRIC
            if (st == 0) { if (m == 0) acc = sqrt((double)nth) * to_c(x3d[b]); }
            else {
                // entry jj sits at theta position p = jj (thin: jj - 1) + hs/2: twiddle index m (2 p + hs) mod 2 nth
                int idx = ((m * (h2 - (thin ? 2 : 0))) % n2t + n2t) % n2t;
                const int stp = (2 * m) % n2t;
                for (int jj = 0; jj < nj; ++jj) {
                    fmacc(acc, xs[jj * DFT_QB + tq], tw[idx]);
                    idx += stp; if (idx >= n2t) idx -= n2t;
                }
                if (h2 != 0) acc = mk(acc.y, -acc.x);              // conj(T): times -i
            }
        }
        if constexpr (sizeof(TM) == sizeof(cplx)) xm[(size_t)m * n2 + q] = acc;
        else { xm[(size_t)(2 * m) * n2 + q] = acc.x; xm[(size_t)(2 * m + 1) * n2 + q] = acc.y; }
    }
}

template <typename TX, typename TM>
__global__ void __launch_bounds__(DFT_QB * DFT_MY)
k_dft_scatter(const long long* __restrict__ map3d, const int* __restrict__ jstr, const signed char* __restrict__ hs,
              long long n2, int nth, int nfac, int mode_sel, const TM* __restrict__ xm, TX* __restrict__ x3d, long long n3d) {
    extern __shared__ __align__(16) unsigned char dft_sm[];
    cplx* tw = reinterpret_cast<cplx*>(dft_sm);        // [2 nth]: exp(+i pi t / nth)
    cplx* ys = tw + 2 * nth;                           // [nm][DFT_QB]
    const int tq = threadIdx.x, ty = threadIdx.y;
    const int n2t = 2 * nth;
    for (int t = ty * DFT_QB + tq; t < n2t; t += DFT_QB * DFT_MY) {
        double sn, cs_;
        sincospi((double)t / (double)nth, &sn, &cs_);
        tw[t] = mk(cs_, sn);
    }
    const long long q = blockIdx.x * (long long)DFT_QB + tq;
    const int v = blockIdx.y;
    x3d += (size_t)v * n3d;
    const int nm = mode_sel == 0 ? nth : nfac;
    const int nmv = mode_sel == 0 ? nth : 2 * nfac;
    xm += (size_t)v * nmv * n2;
    long long b = -1; int st = 0, h2 = 0;
    if (q < n2) { b = map3d[q]; st = jstr[q]; h2 = hs[q]; }
    for (int m = ty; m < nm; m += DFT_MY) {
        cplx val = mk(0.0, 0.0);
        if (b >= 0) {
            if constexpr (sizeof(TM) == sizeof(cplx)) val = to_c(xm[(size_t)m * n2 + q]);
            else {
                // real factors: (Re, Im) of modes 0 .. nth/2; modes m and nth - m are conjugates, so the sum over all
                // modes is 2 Re(sum over the stored ones) with weight 1/2 for the self-conjugate modes 0 and nth/2
                double wgt = (m == 0 || 2 * m == nth) ? 1.0 : 2.0;
                val = mk(wgt * (double)xm[(size_t)(2 * m) * n2 + q], wgt * (double)xm[(size_t)(2 * m + 1) * n2 + q]);
            }
        }
        ys[m * DFT_QB + tq] = val;
    }
    __syncthreads();
    if (q >= n2 || b < 0) return;
    const double inv = 1.0 / (double)nth;
    if (st == 0) {
        if (ty == 0) {
            cplx a = ys[tq];                                       // mode 0, T = 1 (weight 1 above)
            x3d[b] = from_c<TX>((1.0 / sqrt((double)nth)) * a);            // (the symmetric scaling of the axis unknown, see the gather)
        }
        return;
    }
    for (int j = ty; j < nth; j += DFT_MY) {
        cplx acc = mk(0.0, 0.0);
        const int stp = ((2 * j + h2) % n2t + n2t) % n2t;
        int idx = 0;
        for (int m = 0; m < nm; ++m) {
            fmacc(acc, ys[m * DFT_QB + tq], tw[idx]);
            idx += stp; if (idx >= n2t) idx -= n2t;
        }
        if (h2 != 0) acc = mk(-acc.y, acc.x);                      // T: times i
        x3d[b + (long long)j * st] = from_c<TX>(mk(acc.x * inv, acc.y * inv));   // (real factors: from_c<double> takes Re)
    }
}

// cylindrically symmetric mesh: one mode, no rotation -- a gather / scatter through map3d
template <typename TX, typename TM>
__global__ void k_mode_copy(const long long* __restrict__ map3d, long long n2, TX* __restrict__ x3d, long long n3d, TM* __restrict__ xm, int to_modes) {
    long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q >= n2) return;
    x3d += (size_t)blockIdx.y * n3d;
    xm += (size_t)blockIdx.y * n2;
    long long b = map3d[q];
    if (to_modes) xm[q] = (b >= 0) ? from_c<TM>(to_c(x3d[b])) : Sc<TM>::zero();
    else if (b >= 0) x3d[b] = from_c<TX>(to_c(xm[q]));
}

static size_t dft_smem(int nth, int rows) { return (size_t)(2 * nth + rows * DFT_QB) * sizeof(cplx); }
template <typename K> static int dft_smem_attr(K kern, size_t sm) {
    if (sm > 48 * 1024) CS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    return CS_OK;
}

int band_dft_gather(const BandPlan& bp, const void* x3d, bool x_cplx, void* xm, bool xm_cplx, long long n3d, int nvec, cudaStream_t st) {
    CS_CHECK(!(x_cplx && !xm_cplx), "complex vectors need complex factors (complex shift)");
    if (bp.nth == 1) {
        dim3 g((unsigned)cdiv(bp.n2, 256), nvec);
        if (x_cplx) k_mode_copy<cplx, cplx><<<g, 256, 0, st>>>(bp.map3d, bp.n2, (cplx*)x3d, n3d, (cplx*)xm, 1);
        else if (xm_cplx) k_mode_copy<double, cplx><<<g, 256, 0, st>>>(bp.map3d, bp.n2, (double*)x3d, n3d, (cplx*)xm, 1);
        else k_mode_copy<double, double><<<g, 256, 0, st>>>(bp.map3d, bp.n2, (double*)x3d, n3d, (double*)xm, 1);
        CS_CUDA(cudaGetLastError());
        return CS_OK;
    }
    dim3 g((unsigned)cdiv(bp.n2, DFT_QB), nvec), blk(DFT_QB, DFT_MY);
    size_t sm = dft_smem(bp.nth, bp.nth);
    const int sel = xm_cplx ? 0 : 2;
    if (x_cplx) {
        CS_TRY(dft_smem_attr(k_dft_gather<cplx, cplx>, sm));
        k_dft_gather<cplx, cplx><<<g, blk, sm, st>>>(bp.map3d, bp.jstr, bp.hs, bp.n2, bp.nth, bp.nmodes, sel, (const cplx*)x3d, n3d, (cplx*)xm, 0, 0);
    } else if (xm_cplx) {
        CS_TRY(dft_smem_attr(k_dft_gather<double, cplx>, sm));
        k_dft_gather<double, cplx><<<g, blk, sm, st>>>(bp.map3d, bp.jstr, bp.hs, bp.n2, bp.nth, bp.nmodes, sel, (const double*)x3d, n3d, (cplx*)xm, 0, 0);
    } else {
        CS_TRY(dft_smem_attr(k_dft_gather<double, double>, sm));
        k_dft_gather<double, double><<<g, blk, sm, st>>>(bp.map3d, bp.jstr, bp.hs, bp.n2, bp.nth, bp.nmodes, sel, (const double*)x3d, n3d, (double*)xm, 0, 0);
    }
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}
int band_dft_scatter(const BandPlan& bp, const void* xm, bool xm_cplx, void* x3d, bool x_cplx, long long n3d, int nvec, cudaStream_t st) {
    CS_CHECK(!(x_cplx && !xm_cplx), "complex vectors need complex factors (complex shift)");
    if (bp.nth == 1) {
        dim3 g((unsigned)cdiv(bp.n2, 256), nvec);
        if (x_cplx) k_mode_copy<cplx, cplx><<<g, 256, 0, st>>>(bp.map3d, bp.n2, (cplx*)x3d, n3d, (cplx*)xm, 0);
        else if (xm_cplx) k_mode_copy<double, cplx><<<g, 256, 0, st>>>(bp.map3d, bp.n2, (double*)x3d, n3d, (cplx*)xm, 0);
        else k_mode_copy<double, double><<<g, 256, 0, st>>>(bp.map3d, bp.n2, (double*)x3d, n3d, (double*)xm, 0);
        CS_CUDA(cudaGetLastError());
        return CS_OK;
    }
    dim3 g((unsigned)cdiv(bp.n2, DFT_QB), nvec), blk(DFT_QB, DFT_MY);
    const int sel = xm_cplx ? 0 : 2;
    size_t sm = dft_smem(bp.nth, xm_cplx ? bp.nth : bp.nmodes);
    if (x_cplx) {
        CS_TRY(dft_smem_attr(k_dft_scatter<cplx, cplx>, sm));
        k_dft_scatter<cplx, cplx><<<g, blk, sm, st>>>(bp.map3d, bp.jstr, bp.hs, bp.n2, bp.nth, bp.nmodes, sel, (const cplx*)xm, (cplx*)x3d, n3d);
    } else if (xm_cplx) {
        CS_TRY(dft_smem_attr(k_dft_scatter<double, cplx>, sm));
        k_dft_scatter<double, cplx><<<g, blk, sm, st>>>(bp.map3d, bp.jstr, bp.hs, bp.n2, bp.nth, bp.nmodes, sel, (const cplx*)xm, (double*)x3d, n3d);
    } else {
        CS_TRY(dft_smem_attr(k_dft_scatter<double, double>, sm));
        k_dft_scatter<double, double><<<g, blk, sm, st>>>(bp.map3d, bp.jstr, bp.hs, bp.n2, bp.nth, bp.nmodes, sel, (const double*)xm, (double*)x3d, n3d);
    }
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// ------------------------------------------------------------- probing ------
// Columns of the s-independent part K are obtained by applying the *same*
// stencil kernel that the Krylov loop uses to 2p+1 "coloured" probe vectors
// (ones at theta column jprobe for every band index q == color mod (2p+1)).  The stencils couple
// neighbouring theta columns only, so the response lives in columns jprobe-1 .. jprobe+1 and its
// three-term DFT holds column q of every mode matrix at once (cost independent of nth).
template <typename TX>
__global__ void k_probe_fill(const long long* __restrict__ map3d, const int* __restrict__ jstr, long long n2, int ncol, int color0,
                             long long n3d, int jprobe, TX* __restrict__ x3d, int set) {
    // probing vector blockIdx.y carries colour color0 + blockIdx.y; the ones sit in theta column jprobe (set = 0: take them out again)
    long long q = color0 + blockIdx.y + (blockIdx.x * (long long)blockDim.x + threadIdx.x) * ncol;
    if (q >= n2) return;
    long long b = map3d[q];
    if (b >= 0) x3d[(size_t)blockIdx.y * n3d + b + (long long)jprobe * jstr[q]] = set ? Sc<TX>::one() : Sc<TX>::zero();
}
// nb probing vectors (colours color .. color + nb - 1) at once
// mode 0: zero the vectors and set the ones (self-contained); 1: set the ones of these colours in vectors that are all zero;
// 2: take them out again (a full-vector memset per colour is 117 MB at 4.86 M cells)
int band_probe_fill(const BandPlan& bp, int color, int nb, void* x3d, long long n3d, bool x_cplx, cudaStream_t st, int mode) {
    int ncol = 2 * bp.p + 1;
    if (mode == 0) CS_CUDA(cudaMemsetAsync(x3d, 0, (size_t)nb * n3d * (x_cplx ? sizeof(cplx) : sizeof(double)), st));
    long long cnt = (bp.n2 + ncol - 1) / ncol;
    dim3 g((unsigned)cdiv(cnt, TPB), nb);
    const int set = mode == 2 ? 0 : 1;
    if (x_cplx) k_probe_fill<cplx><<<g, TPB, 0, st>>>(bp.map3d, bp.jstr, bp.n2, ncol, color, n3d, bp.jprobe, (cplx*)x3d, set);
    else k_probe_fill<double><<<g, TPB, 0, st>>>(bp.map3d, bp.jstr, bp.n2, ncol, color, n3d, bp.jprobe, (double*)x3d, set);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}
int band_probe_gather(const BandPlan& bp, const void* y3d, bool y_cplx, void* ym, long long n3d, int nb, cudaStream_t st) {
    if (bp.nth == 1) {
        dim3 g((unsigned)cdiv(bp.n2, 256), nb);
        if (y_cplx) k_mode_copy<cplx, cplx><<<g, 256, 0, st>>>(bp.map3d, bp.n2, (cplx*)y3d, n3d, (cplx*)ym, 1);
        else k_mode_copy<double, cplx><<<g, 256, 0, st>>>(bp.map3d, bp.n2, (double*)y3d, n3d, (cplx*)ym, 1);
        CS_CUDA(cudaGetLastError());
        return CS_OK;
    }
    const int thin = (bp.nth >= 3 && bp.thin_probe) ? 1 : 0;
    dim3 g((unsigned)cdiv(bp.n2, DFT_QB), nb), blk(DFT_QB, DFT_MY);
    size_t sm = dft_smem(bp.nth, thin ? 3 : bp.nth);
    if (y_cplx) {
        CS_TRY(dft_smem_attr(k_dft_gather<cplx, cplx>, sm));
        k_dft_gather<cplx, cplx><<<g, blk, sm, st>>>(bp.map3d, bp.jstr, bp.hs, bp.n2, bp.nth, bp.nmodes, 1, (const cplx*)y3d, n3d, (cplx*)ym, bp.jprobe, thin);
    } else {
        CS_TRY(dft_smem_attr(k_dft_gather<double, cplx>, sm));
        k_dft_gather<double, cplx><<<g, blk, sm, st>>>(bp.map3d, bp.jstr, bp.hs, bp.n2, bp.nth, bp.nmodes, 1, (const double*)y3d, n3d, (cplx*)ym, bp.jprobe, thin);
    }
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// K''_m[r, q] = conj(T_m[r]) K_m[r, q] T_m[q]: the gather delivered conj(T_m[r]) K_m[r, q]; the column factor is applied here.
// Real band storage keeps the real part (the imaginary part is rounding noise for a theta-invariant operator; for a
// theta-varying one the circulant preconditioner is symmetrised by dropping it).
template <typename TB, typename TX>
__global__ void k_probe_store(const long long* __restrict__ map3d, const int* __restrict__ jstr, const signed char* __restrict__ hs,
                              long long n2, int p, int pw, int ldab, int color0, int nmodes, int nth, int jprobe,
                              const cplx* __restrict__ ym, const TX* __restrict__ y3d, long long n3d, TB* __restrict__ K) {
    long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int m = blockIdx.y;
    const int color = color0 + blockIdx.z;                   // response of probing vector blockIdx.z
    ym += (size_t)blockIdx.z * nmodes * n2;
    y3d += (size_t)blockIdx.z * n3d;
    if (r >= n2) return;
    if (map3d[r] < 0) return;
    if (jstr[r] == 0 && m > 0) return;
    int ncol = 2 * p + 1;
    long long lo = r - p;
    long long off = ((color - lo) % ncol + ncol) % ncol;
    long long q = lo + off;
    if (q < 0 || q >= n2) return;
    if (map3d[q] < 0) return;
    if (jstr[q] == 0 && m > 0) return;
    cplx val = ym[(size_t)m * n2 + r];
    if (jstr[q] == 0 && jstr[r] != 0) {
        // column of the shared axis unknown: it couples to EVERY theta column with the same entry B, the
        // mode-0 entry is the sum over all nth of them (the three-term gather saw only three); the axis unknown itself
        // is carried as sqrt(nth) x_axis (see k_dft_gather), which leaves sqrt(nth) B -- the transpose of its row
        val = sqrt((double)nth) * to_c(y3d[map3d[r] + (long long)jprobe * jstr[r]]);
        if (hs[r] != 0) val = mk(val.y, -val.x);
    } else if (jstr[q] == 0) {
        val = (1.0 / sqrt((double)nth)) * val;          // axis row (scaled by the gather) of an axis column: unscaled entry
    }
    const int h2 = hs[q];
    if (h2 != 0) {
        double sn, cs_;
        sincospi((double)(h2 * m) / (double)nth, &sn, &cs_);   // T_m[q] = i exp(i hs pi m / nth)
        val = val * mk(-sn, cs_);
    }
    K[(size_t)m * n2 * ldab + (size_t)(pw + r - q) + (size_t)q * ldab] = from_c<TB>(val);
}
// Thin probing, gather and store in one kernel: the response to a probe in theta column jprobe lives in the columns
// jprobe - 1 .. jprobe + 1, so entry [r, q] of every mode matrix is a 3-term sum that is formed right here from the 3-D
// response -- the mode vectors (17 x n2 complex numbers per probing vector: 124 MB written and read back at 4.86 M cells)
// never exist.  One thread per band row r, all stored modes; writes run down column q of the band (coalesced).
template <typename TB, typename TX>
__global__ void __launch_bounds__(128)
k_probe_store_thin(const long long* __restrict__ map3d, const int* __restrict__ jstr, const signed char* __restrict__ hs,
                   long long n2, int p, int pw, int ldab, int color0, int nmodes, int nth, int jprobe,
                   const TX* __restrict__ y3d, long long n3d, TB* __restrict__ K) {
    extern __shared__ __align__(16) unsigned char pst_sm[];
    cplx* tw = reinterpret_cast<cplx*>(pst_sm);        // [2 nth]: exp(-i pi t / nth)
    const int n2t = 2 * nth;
    for (int t = threadIdx.x; t < n2t; t += blockDim.x) {
        double sn, cs_;
        sincospi(-(double)t / (double)nth, &sn, &cs_);
        tw[t] = mk(cs_, sn);
    }
    __syncthreads();
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int color = color0 + blockIdx.y;
    y3d += (size_t)blockIdx.y * n3d;
    if (r >= n2) return;
    const long long b = map3d[r];
    if (b < 0) return;
    const int ncol = 2 * p + 1;
    const long long lo = r - p;
    const long long off = ((color - lo) % ncol + ncol) % ncol;
    const long long q = lo + off;
    if (q < 0 || q >= n2) return;
    if (map3d[q] < 0) return;
    const int st = jstr[r], stq = jstr[q], h2 = hs[r], hq = hs[q];
    const double rt = sqrt((double)nth);
    TB* Kc = K + (size_t)(pw + r - q) + (size_t)q * ldab;
    const size_t nel = (size_t)n2 * ldab;
    if (st == 0) {
        // axis row (mode 0 only), carried as sqrt(nth) x: scaled row; an axis column gives the unscaled entry back
        cplx val = rt * to_c(y3d[b]);
        if (stq == 0) val = (1.0 / rt) * val;
        if (hq != 0) val = mk(-val.y, val.x);          // T_0[q] = i for a staggered column
        Kc[0] = from_c<TB>(val);
        return;
    }
    cplx x0, x1, x2;
    {
        int c0 = (jprobe - 1 + nth) % nth, c2 = (jprobe + 1) % nth;
        x0 = to_c(y3d[b + (long long)c0 * st]); x1 = to_c(y3d[b + (long long)jprobe * st]); x2 = to_c(y3d[b + (long long)c2 * st]);
    }
    if (stq == 0) {
        // column of the shared axis unknown (mode 0 only): sqrt(nth) times the single-column response
        cplx val = rt * x1;
        if (h2 != 0) val = mk(val.y, -val.x);
        Kc[0] = from_c<TB>(val);
        return;
    }
    for (int m = 0; m < nmodes; ++m) {
        // sum_p x_p exp(-i pi m (2 p + hs_r) / nth), p = -1, 0, 1;  times conj(T_m[r]) = -i for a staggered row
        int idx = ((m * (h2 - 2)) % n2t + n2t) % n2t;
        const int stp = (2 * m) % n2t;
        cplx acc = mk(0.0, 0.0);
        fmacc(acc, x0, tw[idx]); idx += stp; if (idx >= n2t) idx -= n2t;
        fmacc(acc, x1, tw[idx]); idx += stp; if (idx >= n2t) idx -= n2t;
        fmacc(acc, x2, tw[idx]);
        if (h2 != 0) acc = mk(acc.y, -acc.x);
        if (hq != 0) {
            // T_m[q] = i exp(i hq pi m / nth) = i conj(tw[hq m])
            const cplx e = tw[((hq * m) % n2t + n2t) % n2t];
            acc = acc * mk(e.y, e.x);                  // i * conj(e) = i (e.x - i e.y) = e.y + i e.x
        }
        Kc[(size_t)m * nel] = from_c<TB>(acc);
    }
}
int band_probe_store_thin(const BandPlan& bp, int color, int nb, const void* y3d, bool y_cplx, long long n3d, cudaStream_t st) {
    CS_CHECK(bp.cplx_band == y_cplx, "probing vectors must have the scalar type of the band storage");
    CS_CHECK(bp.nth >= 3 && bp.thin_probe, "thin probing needs a three-column response");
    dim3 g((unsigned)cdiv(bp.n2, 128), nb);
    const size_t sm = (size_t)2 * bp.nth * sizeof(cplx);
    if (bp.cplx_band) k_probe_store_thin<cplx, cplx><<<g, 128, sm, st>>>(bp.map3d, bp.jstr, bp.hs, bp.n2, bp.p, bp.pw, bp.ldab, color, bp.nmodes, bp.nth, bp.jprobe, (const cplx*)y3d, n3d, (cplx*)bp.Kband);
    else k_probe_store_thin<double, double><<<g, 128, sm, st>>>(bp.map3d, bp.jstr, bp.hs, bp.n2, bp.p, bp.pw, bp.ldab, color, bp.nmodes, bp.nth, bp.jprobe, (const double*)y3d, n3d, (double*)bp.Kband);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}
int band_probe_store(const BandPlan& bp, int color, int nb, const void* ym, const void* y3d, bool y_cplx, long long n3d, cudaStream_t st) {
    dim3 g((unsigned)cdiv(bp.n2, TPB), bp.nmodes, nb);
    CS_CHECK(bp.cplx_band == y_cplx, "probing vectors must have the scalar type of the band storage");
    if (bp.cplx_band) k_probe_store<cplx, cplx><<<g, TPB, 0, st>>>(bp.map3d, bp.jstr, bp.hs, bp.n2, bp.p, bp.pw, bp.ldab, color, bp.nmodes, bp.nth, bp.jprobe, (const cplx*)ym, (const cplx*)y3d, n3d, (cplx*)bp.Kband);
    else k_probe_store<double, double><<<g, TPB, 0, st>>>(bp.map3d, bp.jstr, bp.hs, bp.n2, bp.p, bp.pw, bp.ldab, color, bp.nmodes, bp.nth, bp.jprobe, (const cplx*)ym, (const double*)y3d, n3d, (double*)bp.Kband);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// LU[f][m] = K[m] + s[f]*diag(mdiag); unknowns that do not exist in a mode
// (dummies; the axis edge for m > 0) get a unit diagonal.
template <typename TB, typename TL>
__global__ void k_form_shifted(const long long* __restrict__ map3d, const int* __restrict__ jstr, const double* __restrict__ mdiag,
                               long long n2, int pw, int ldab, int nmodes, const TB* __restrict__ K,
                               const double* __restrict__ s_d, TL* __restrict__ LU) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // element inside one band matrix
    long long nel = n2 * (long long)ldab;
    if (e >= nel) return;
    int m = blockIdx.y, f = blockIdx.z;
    long long q = e / ldab;
    int d = (int)(e - q * ldab);
    TB kv = K[(size_t)m * nel + e];
    TL v = Sc<TL>::from(Sc<TB>::re(kv), Sc<TB>::im(kv));
    if (d == pw) {
        bool dummy = (map3d[q] < 0) || (jstr[q] == 0 && m > 0);
        if (dummy) v = Sc<TL>::one();
        else v = v + Sc<TL>::from(s_d[2 * f], s_d[2 * f + 1]) * Sc<TL>::from(mdiag[q], 0.0);
    }
    LU[((size_t)f * nmodes + m) * nel + e] = v;
}
int band_form_shifted(const BandPlan& bp, const double* s_d, int nshift, void* LU, bool lu_cplx, cudaStream_t st) {
    long long nel = bp.n2 * (long long)bp.ldab;
    dim3 g((unsigned)cdiv(nel, 256), bp.nmodes, nshift);
    if (bp.cplx_band) {
        CS_CHECK(lu_cplx, "complex band needs complex LU");
        k_form_shifted<cplx, cplx><<<g, 256, 0, st>>>(bp.map3d, bp.jstr, bp.mdiag, bp.n2, bp.pw, bp.ldab, bp.nmodes, (const cplx*)bp.Kband, s_d, (cplx*)LU);
    } else if (lu_cplx) {
        k_form_shifted<double, cplx><<<g, 256, 0, st>>>(bp.map3d, bp.jstr, bp.mdiag, bp.n2, bp.pw, bp.ldab, bp.nmodes, (const double*)bp.Kband, s_d, (cplx*)LU);
    } else {
        k_form_shifted<double, double><<<g, 256, 0, st>>>(bp.map3d, bp.jstr, bp.mdiag, bp.n2, bp.pw, bp.ldab, bp.nmodes, (const double*)bp.Kband, s_d, (double*)LU);
    }
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// one mode matrix (Kband = that mode's real band) -> A(s) in complex storage; `mode` is its true theta-mode index
// (the shared axis unknown is a dummy in every mode but 0).  Used by the mode-sharded factorisation of the z-slab solve.
__global__ void k_form_shifted_one(const long long* __restrict__ map3d, const int* __restrict__ jstr, const double* __restrict__ mdiag,
                                   long long n2, int pw, int ldab, int mode, const double* __restrict__ K,
                                   const double* __restrict__ s_d, cplx* __restrict__ LU) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= n2 * (long long)ldab) return;
    long long q = e / ldab;
    int d = (int)(e - q * ldab);
    cplx v = mk(K[e], 0.0);
    if (d == pw) {
        bool dummy = (map3d[q] < 0) || (jstr[q] == 0 && mode > 0);
        v = dummy ? mk(1.0, 0.0) : mk(v.x + s_d[0] * mdiag[q], s_d[1] * mdiag[q]);
    }
    LU[e] = v;
}
int band_form_shifted_mode(const BandPlan& bp, int mode, const double* s_d, void* LU, cudaStream_t st) {
    long long nel = bp.n2 * (long long)bp.ldab;
    k_form_shifted_one<<<(unsigned)cdiv(nel, 256), 256, 0, st>>>(bp.map3d, bp.jstr, bp.mdiag, bp.n2, bp.pw, bp.ldab, mode,
                                                                  (const double*)bp.Kband, s_d, (cplx*)LU);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// (the diagonal is invariant under the rotation; modes m and nth - m share it)
// max |K(i,j) - K(j,i)| and max |K(i,j)| over the stored band of all modes (non-negative doubles order like their bit patterns)
template <typename TB>
__global__ void k_band_asym(const TB* __restrict__ K, long long n2, int p, int pw, int ldab, size_t nel, int stride, unsigned long long* __restrict__ out) {
    const long long j = blockIdx.x * (long long)stride;      // one column per CTA, threads across the sub-diagonals
    const TB* A = K + (size_t)blockIdx.y * nel;
    double da = 0.0, mx = 0.0;
    if (threadIdx.x == 0) mx = sqrt(Sc<TB>::abs2(A[pw + j * (long long)ldab]));
    for (int d = 1 + threadIdx.x; d <= p && j + d < n2; d += blockDim.x) {
        TB lo = A[(pw + d) + j * (long long)ldab], up = A[(pw - d) + (j + d) * (long long)ldab];
        da = fmax(da, sqrt(Sc<TB>::abs2(lo - up)));
        mx = fmax(mx, fmax(sqrt(Sc<TB>::abs2(lo)), sqrt(Sc<TB>::abs2(up))));
    }
    for (int o = 16; o; o >>= 1) { da = fmax(da, __shfl_xor_sync(0xffffffffu, da, o)); mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMax(out, (unsigned long long)__double_as_longlong(da)); atomicMax(out + 1, (unsigned long long)__double_as_longlong(mx)); }
}
// stride > 1: every stride-th column only (the transposed reads are uncoalesced: 60 ms for all columns of the 109x32x1394
// edge systems).  The operators of the path are symmetric by construction, for them the sampled check is a guard; an
// assembled CSR matrix is checked in full.
int band_check_symmetric(const BandPlan& bp, int stride, double* asym_rel, cudaStream_t st) {
    void* buf = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_band_mu);
        CS_TRY(band_scratch(3, st, 2 * sizeof(unsigned long long), false, &buf));
    }
    CS_CUDA(cudaMemsetAsync(buf, 0, 2 * sizeof(unsigned long long), st));
    if (stride < 1) stride = 1;
    dim3 g((unsigned)((bp.n2 + stride - 1) / stride), bp.nmodes);
    if (bp.cplx_band) k_band_asym<cplx><<<g, 128, 0, st>>>((const cplx*)bp.Kband, bp.n2, bp.p, bp.pw, bp.ldab, bp.band_elems(), stride, (unsigned long long*)buf);
    else k_band_asym<double><<<g, 128, 0, st>>>((const double*)bp.Kband, bp.n2, bp.p, bp.pw, bp.ldab, bp.band_elems(), stride, (unsigned long long*)buf);
    CS_CUDA(cudaGetLastError());
    unsigned long long h[2];
    CS_CUDA(cudaMemcpyAsync(h, buf, sizeof h, cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaStreamSynchronize(st));
    double da, mx;
    memcpy(&da, &h[0], 8); memcpy(&mx, &h[1], 8);
    *asym_rel = mx > 0 ? da / mx : 0.0;
    return CS_OK;
}

template <typename TB>
__global__ void k_diag_weights(const long long* __restrict__ map3d, const int* __restrict__ jstr, const double* __restrict__ mdiag,
                               long long n2, int pw, int ldab, int nmodes, int nth, const TB* __restrict__ K,
                               const double* __restrict__ s_d, long long n3d, double* __restrict__ w) {
    long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int f = blockIdx.y;
    if (q >= n2) return;
    long long b = map3d[q];
    if (b < 0) return;
    int st = jstr[q];
    size_t nel = (size_t)n2 * ldab;
    double re = 0.0, im = 0.0;
    int nm = (st == 0) ? 1 : nmodes;
    double cnt = 0.0;
    for (int m = 0; m < nm; ++m) {
        TB kv = K[(size_t)m * nel + (size_t)pw + (size_t)q * ldab];
        double wgt = (nth == 1 || st == 0 || m == 0 || 2 * m == nth) ? 1.0 : 2.0;
        re += wgt * Sc<TB>::re(kv); im += wgt * Sc<TB>::im(kv); cnt += wgt;
    }
    re /= cnt; im /= cnt;
    re += s_d[2 * f] * mdiag[q];
    im += s_d[2 * f + 1] * mdiag[q];
    double ww = 1.0 / sqrt(re * re + im * im);
    double* wf = w + (size_t)f * n3d;
    if (st == 0) { wf[b] = ww; return; }
    for (int j = 0; j < nth; ++j) wf[b + (long long)j * st] = ww;
}
int band_diag_weights(const BandPlan& bp, const double* s_d, int nshift, long long n3d, double* w, cudaStream_t st) {
    dim3 g((unsigned)cdiv(bp.n2, TPB), nshift);
    if (bp.cplx_band) k_diag_weights<cplx><<<g, TPB, 0, st>>>(bp.map3d, bp.jstr, bp.mdiag, bp.n2, bp.pw, bp.ldab, bp.nmodes, bp.nth, (const cplx*)bp.Kband, s_d, n3d, w);
    else k_diag_weights<double><<<g, TPB, 0, st>>>(bp.map3d, bp.jstr, bp.mdiag, bp.n2, bp.pw, bp.ldab, bp.nmodes, bp.nth, (const double*)bp.Kband, s_d, n3d, w);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// ---------------------------------------------------------- blocked LU ------
// Step t factors block column [t, t+nbe): panel kernel (one CTA per system)
//   A11 = L11 U11,  L21 = A21 U11^-1 (R x nbe),  U12 = L11^-1 A12 (nbe x R)
// then the update kernel (32x32 tiles x systems)  A22 -= L21 U12  (R x R).
// NB is a compile-time constant so the per-thread row / column solves live in
// registers; a short last block is padded with an identity.

// grid = (chunks, systems), 128 threads: every CTA factors the NB x NB diagonal block
// itself (one warp, ~1 us, redundant across chunks) and then solves 64 rows of A21 and
// 64 columns of A12.  A11 is NOT overwritten here (other chunks may still read it):
// chunk 0 parks the factored block in `scratch`, the update kernel copies it in place.
template <typename T, int NB>
__global__ void __launch_bounds__(160)
k_lu_panel(T* __restrict__ LU, size_t nel, long long n2, int p, int pw, int ld, long long t, T* __restrict__ scratch, int sym, int dahead) {
    __shared__ T D[NB][NB + 1];                      // D[r][c]
    T* A = LU + (size_t)blockIdx.y * nel + pw;       // dense view base
    int nbe = (int)min((long long)NB, n2 - t);
    int R = (int)min((long long)p, n2 - (t + nbe));
    int tid = threadIdx.x;
    // symmetric matrices: 128 rows of A21 per CTA and no column solves -- U12 = diag(U11) L21^T is written by the row threads
    int idx = sym ? blockIdx.x * 128 + (tid & 127) : blockIdx.x * 64 + (tid & 63);          // row / column handled by this thread
    bool isrow = (tid < (sym ? 128 : 64)) && idx < R, iscol = !sym && (tid >= 64 && tid < 128) && idx < R;
    // threads 0..127: 64 rows of A21 + 64 columns of A12; warp 4: the LU of the diagonal block.
    // The panel loads are issued first so that they overlap the A11 factorisation.
    T v[NB];
    long long row = t + nbe + idx, col = t + nbe + idx;
    if (isrow) {
#pragma unroll
        for (int c = 0; c < NB; ++c) v[c] = (c < nbe) ? A[row + (t + c) * (long long)ld] : Sc<T>::zero();
    } else if (iscol) {
        const T* cp = A + col * (long long)ld + t;
#pragma unroll
        for (int r = 0; r < NB; ++r) v[r] = (r < nbe) ? cp[r] : Sc<T>::zero();
    }
    // dahead: the pivot block was factored (and parked in `scratch`) by the trailing update of the previous step, off the
    // critical path -- see k_lu_update
    const T* Sin = scratch + (size_t)blockIdx.y * NB * NB;
    for (int e = tid; e < NB * NB; e += blockDim.x) {
        int r = e % NB, c = e / NB;
        T val = (r == c) ? Sc<T>::one() : Sc<T>::zero();
        if (dahead) val = Sin[e];
        else if (r < nbe && c < nbe) val = A[(t + r) + (t + c) * (long long)ld];
        D[r][c] = val;
    }
    __syncthreads();
    // unblocked LU of the NB x NB diagonal block by one warp (lane r owns row r)
    // (the LU warp reuses v[] for its row: it holds no panel data)
    if (tid >= 128 && !dahead) {
        int lane = tid - 128;
        int r = lane & (NB - 1);
#pragma unroll
        for (int c = 0; c < NB; ++c) v[c] = D[r][c];
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            T piv = Sc<T>::inv(shfl_row(v[c], c));       // pivot row c is final now
            T l = v[c] * piv;
            if (r > c) v[c] = l;
#pragma unroll
            for (int cc = c + 1; cc < NB; ++cc) {
                T u = shfl_row(v[cc], c);
                if (r > c) v[cc] -= l * u;
            }
        }
        if (lane < NB) {
#pragma unroll
            for (int c = 0; c < NB; ++c) D[r][c] = v[c];
        }
    }
    __syncthreads();
    if (blockIdx.x == 0 && !dahead) {
        T* S = scratch + (size_t)blockIdx.y * NB * NB;
        for (int e = tid; e < NB * NB; e += blockDim.x) S[e] = D[e % NB][e / NB];
    }
    if (isrow) {
        // row a of A21:  v U11 = row   (two interleaved accumulators shorten the fp64 dependency chain)
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            T acc0 = v[c], acc1 = Sc<T>::zero();
#pragma unroll
            for (int c2 = 0; c2 < c; ++c2) { if (c2 & 1) acc1 -= v[c2] * D[c2][c]; else acc0 -= v[c2] * D[c2][c]; }
            v[c] = (acc0 + acc1) * Sc<T>::inv(D[c][c]);
        }
#pragma unroll
        for (int c = 0; c < NB; ++c) if (c < nbe) A[row + (t + c) * (long long)ld] = v[c];
        if (sym) {
            T* cp = A + row * (long long)ld + t;         // column `row` of U12
#pragma unroll
            for (int r = 0; r < NB; ++r) if (r < nbe) cp[r] = D[r][r] * v[r];
        }
    } else if (iscol) {
        // column b of A12:  L11 w = col (unit lower)
#pragma unroll
        for (int r = 0; r < NB; ++r) {
            T acc0 = v[r], acc1 = Sc<T>::zero();
#pragma unroll
            for (int r2 = 0; r2 < r; ++r2) { if (r2 & 1) acc1 -= D[r][r2] * v[r2]; else acc0 -= D[r][r2] * v[r2]; }
            v[r] = acc0 + acc1;
        }
        T* cp = A + col * (long long)ld + t;
#pragma unroll
        for (int r = 0; r < NB; ++r) if (r < nbe) cp[r] = v[r];
    }
}

// Panel of a step whose pivot block is already factored (parked in `scratch` by the previous update, see k_lu_update):
// four threads per row of A21.  One thread per row runs 136 dependent-ish complex FMAs -- ~1 500 instructions issued one
// per ~10 cycles by a lone warp per scheduler, ~8 us of the step.  Here the row solve  v U11 = a  is right-looking over a
// quad: the owner of column c2 finishes v[c2] = a[c2] / U[c2][c2], one shuffle hands it to the quad, every thread updates
// the (at most four) later columns it owns -- a chain of 16 x (shuffle + complex FMA + complex multiply).  Symmetric
// matrices get U12 = diag(U11) L21^T from the same registers (16 contiguous elements per quad); general ones solve the
// columns of A12 (L11 w = a, unit lower) in the second half of the grid.
template <typename T>
__global__ void __launch_bounds__(256)
k_lu_panel4(T* __restrict__ LU, size_t nel, long long n2, int p, int pw, int ld, long long t, const T* __restrict__ scratch, int sym, int chunks) {
    constexpr int NB = 16;
    __shared__ T D[NB][NB + 1];
    __shared__ T invd[NB];
    T* A = LU + (size_t)blockIdx.y * nel + pw;
    const int nbe = (int)min((long long)NB, n2 - t);
    const int R = (int)min((long long)p, n2 - (t + nbe));
    const int tid = threadIdx.x, q = tid & 3, lane = tid & 31, qbase = lane & ~3;
    const bool colmode = (int)blockIdx.x >= chunks;
    const int idx = (colmode ? (int)blockIdx.x - chunks : (int)blockIdx.x) * 64 + (tid >> 2);
    const bool act = idx < R;
    const long long rc = t + nbe + idx;                 // the row of A21 (column of A12) of this quad
    const T zero = Sc<T>::zero();
    T a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = q + 4 * i;
        a[i] = zero;
        if (act && c < nbe) a[i] = colmode ? A[rc * (long long)ld + t + c] : A[rc + (t + c) * (long long)ld];
    }
    D[tid % NB][tid / NB] = scratch[(size_t)blockIdx.y * NB * NB + tid];
    __syncthreads();
    if (tid < NB) invd[tid] = Sc<T>::inv(D[tid][tid]);
    __syncthreads();
    if (!colmode) {
#pragma unroll
        for (int c2 = 0; c2 < NB; ++c2) {
            const int io = c2 >> 2, qo = c2 & 3;
            const T v = shfl_t<T>(a[io] * invd[c2], qbase | qo);
            if (q == qo) a[io] = v;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i < io) continue;
                if (i > io || q > qo) a[i] -= v * D[c2][q + 4 * i];
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = q + 4 * i;
            if (act && c < nbe) {
                A[rc + (t + c) * (long long)ld] = a[i];
                if (sym) A[(t + c) + rc * (long long)ld] = D[c][c] * a[i];
            }
        }
    } else {
#pragma unroll
        for (int r2 = 0; r2 < NB; ++r2) {
            const int io = r2 >> 2, qo = r2 & 3;
            const T w = shfl_t<T>(a[io], qbase | qo);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i < io) continue;
                if (i > io || q > qo) a[i] -= D[q + 4 * i][r2] * w;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = q + 4 * i;
            if (act && r < nbe) A[rc * (long long)ld + t + r] = a[i];
        }
    }
}

// (occupancy experiments: 5 CTAs/SM at 48 registers 119 ms, 6 at 40 registers 131 ms, this one -- 64 registers, 4 CTAs/SM -- 115 ms)
template <typename T, int NB>
__global__ void __launch_bounds__(256)
k_lu_update(T* __restrict__ LU, size_t nel, long long n2, int p, int pw, int ld, long long t, T* __restrict__ scratch, int part, int tiles, int dahead) {
    __shared__ T Ls[NB][32];
    __shared__ T Us[NB][33];
    T* A = LU + (size_t)blockIdx.z * nel + pw;
    int nbe = (int)min((long long)NB, n2 - t);
    int R = (int)min((long long)p, n2 - (t + nbe));
    // part 0: the whole R x R trailing block (grid tiles x tiles).  Look-ahead split: part 1 = the first tile row and tile
    // column (all the NEXT block step reads; 1-D grid of 2 tiles - 1), part 2 = the rest (grid (tiles-1) x (tiles-1))
    int bx = blockIdx.x, by = blockIdx.y;
    if (part == 1) { if (bx < tiles) by = 0; else { by = bx - tiles + 1; bx = 0; } }
    else if (part == 2) { bx += 1; by += 1; }
    else if (part == 3) {
        // symmetric matrices: the tiles on and below the diagonal only (1-D grid of tiles (tiles + 1) / 2, column by column);
        // the upper triangle of the window is never read again -- the panel mirrors the pivot rows from L21
        int L = bx, rem = tiles;
        by = 0;
        while (L >= rem) { L -= rem; --rem; ++by; }
        bx = by + L;
    }
    int a0 = bx * 32, b0 = by * 32;
    int tid = threadIdx.x;
    if (bx == 0 && by == 0) {
        // put the factored diagonal block (parked by the panel kernel) in place
        const T* S = scratch + (size_t)blockIdx.z * NB * NB;
        for (int e = tid; e < nbe * nbe; e += 256) {
            int r = e % nbe, c = e / nbe;
            A[(t + r) + (t + c) * (long long)ld] = S[r + c * NB];
        }
    }
    if (a0 >= R || b0 >= R) return;
    long long r0 = t + nbe + a0, c0 = t + nbe + b0;
    // thread -> rows (ta, ta+16), cols (tb, tb+16); start the C loads early
    int ta = tid % 16, tb = tid / 16;
    bool m0 = (a0 + ta < R), m1 = (a0 + ta + 16 < R), n0 = (b0 + tb < R), n1 = (b0 + tb + 16 < R);
    T* c00 = A + (r0 + ta) + (c0 + tb) * (long long)ld;
    T* c01 = A + (r0 + ta) + (c0 + tb + 16) * (long long)ld;
    T old00 = (m0 && n0) ? c00[0] : Sc<T>::zero(), old10 = (m1 && n0) ? c00[16] : Sc<T>::zero();
    T old01 = (m0 && n1) ? c01[0] : Sc<T>::zero(), old11 = (m1 && n1) ? c01[16] : Sc<T>::zero();
    for (int e = tid; e < NB * 32; e += 256) {
        int a = e % 32, c = e / 32;
        Ls[c][a] = (a0 + a < R && c < nbe) ? A[(r0 + a) + (t + c) * (long long)ld] : Sc<T>::zero();
    }
    for (int e = tid; e < NB * 32; e += 256) {
        int c = e % NB, b = e / NB;
        Us[c][b] = (b0 + b < R && c < nbe) ? A[(t + c) + (c0 + b) * (long long)ld] : Sc<T>::zero();
    }
    __syncthreads();
    T acc00 = Sc<T>::zero(), acc01 = Sc<T>::zero(), acc10 = Sc<T>::zero(), acc11 = Sc<T>::zero();
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        T l0 = Ls[c][ta], l1 = Ls[c][ta + 16];
        T u0 = Us[c][tb], u1 = Us[c][tb + 16];
        fmacc(acc00, l0, u0); fmacc(acc01, l0, u1); fmacc(acc10, l1, u0); fmacc(acc11, l1, u1);
    }
    if (m0 && n0) c00[0] = old00 - acc00;
    if (m1 && n0) c00[16] = old10 - acc10;
    if (m0 && n1) c01[0] = old01 - acc01;
    if (m1 && n1) c01[16] = old11 - acc11;
    // dahead: the CTA of the first diagonal tile also factors the NEXT pivot block (the leading NB x NB of its tile, still in
    // registers) and parks it in `scratch`, while the rest of the grid is still updating: the scalar LU of the pivot block
    // (NB dependent complex reciprocals, ~4 us) leaves the critical path of the panel kernel
    if (dahead && bx == 0 && by == 0 && t + nbe < n2) {
        // (the staged pivot rows are dead after the k loop of every thread: the block lives in the memory of Us)
        __syncthreads();
        T (*Dn)[NB + 1] = reinterpret_cast<T (*)[NB + 1]>(&Us[0][0]);
        const int nbe2 = (int)min((long long)NB, n2 - (t + nbe));
        auto put = [&](int r, int c, bool inwin, T fresh, const T* gp) {
            if (r >= NB || c >= NB) return;
            T val = (r == c) ? Sc<T>::one() : Sc<T>::zero();
            if (r < nbe2 && c < nbe2) val = inwin ? fresh : *gp;      // (narrow bands: beyond this step's reach, unchanged)
            Dn[r][c] = val;
        };
        put(ta, tb, m0 && n0, old00 - acc00, c00);
        put(ta + 16, tb, m1 && n0, old10 - acc10, c00 + 16);
        put(ta, tb + 16, m0 && n1, old01 - acc01, c01);
        put(ta + 16, tb + 16, m1 && n1, old11 - acc11, c01 + 16);
        __syncthreads();
        // right-looking scalar LU out of shared memory (no per-thread row in registers: the tile update's register
        // budget, and with it the occupancy of the whole grid, stays what it was)
        for (int c = 0; c < NB - 1; ++c) {
            if (tid > c && tid < NB) Dn[tid][c] = Dn[tid][c] * Sc<T>::inv(Dn[c][c]);
            __syncthreads();
#pragma unroll
            for (int q = 0; q < (NB * NB) / 256; ++q) {
                const int r = ta + 16 * (q & 1), cc = tb + 16 * (q >> 1);
                if (r > c && cc > c) Dn[r][cc] -= Dn[r][c] * Dn[c][cc];
            }
            __syncthreads();
        }
        T* S = scratch + (size_t)blockIdx.z * NB * NB;
        for (int e = tid; e < NB * NB; e += 256) S[e] = Dn[e % NB][e / NB];
    }
}

// ---- fused step: trailing update of step t + panel of step t+1 in ONE launch ("last block done") ----
// The two-kernel schedule pays a fixed ~12.5 us per block step for two dependent launches, most of it the latency-bound
// panel (102 CTAs).  Here the update kernel's CTAs count themselves off per system as they finish the first tile row /
// column of the trailing block -- everything the next panel reads -- and the CTA that completes the count runs that
// panel (pivot-block LU by one warp, then the row / column triangular solves in sweeps of 128 + 128) while the rest
// of the grid is still updating: no second launch, no waiting (nobody spins, so nothing can deadlock).  Grid =
// (systems, tiles^2): the system index is the FASTEST block index and the first-row/column tiles come first, so the panels
// of all systems overlap the bulk of the update.  Loads of the tail bypass L1 (__ldcg): the data were written by other SMs.
template <typename T> __device__ __forceinline__ T ldcg_t(const T* p);
template <> __device__ __forceinline__ double ldcg_t<double>(const double* p) { return __ldcg(p); }
template <> __device__ __forceinline__ cplx ldcg_t<cplx>(const cplx* p) { return __ldcg(p); }

template <typename T, int NB>
__device__ __noinline__ void lu_panel_tail(T* __restrict__ A, long long n2, int p, int ld, long long t, T* __restrict__ S, T (*D)[NB + 1]) {
    const int tid = threadIdx.x;
    const int nbe = (int)min((long long)NB, n2 - t);
    const int R = (int)min((long long)p, n2 - (t + nbe));
    for (int e = tid; e < NB * NB; e += 256) {
        int r = e % NB, c = e / NB;
        T val = (r == c) ? Sc<T>::one() : Sc<T>::zero();
        if (r < nbe && c < nbe) val = ldcg_t<T>(A + (t + r) + (t + c) * (long long)ld);
        D[r][c] = val;
    }
    __syncthreads();
    if (tid < 32) {
        const int r = tid & (NB - 1);
        T v[NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) v[c] = D[r][c];
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            T piv = Sc<T>::inv(shfl_row(v[c], c));
            T l = v[c] * piv;
            if (r > c) v[c] = l;
#pragma unroll
            for (int cc = c + 1; cc < NB; ++cc) {
                T u = shfl_row(v[cc], c);
                if (r > c) v[cc] -= l * u;
            }
        }
        if (tid < NB) {
#pragma unroll
            for (int c = 0; c < NB; ++c) D[r][c] = v[c];
        }
    }
    __syncthreads();
    for (int e = tid; e < NB * NB; e += 256) S[e] = D[e % NB][e / NB];
    for (int base = 0; base < R; base += 128) {
        const int idx = base + (tid & 127);
        const bool isrow = tid < 128;
        if (idx >= R) continue;
        T v[NB];
        if (isrow) {
            const long long row = t + nbe + idx;
#pragma unroll
            for (int c = 0; c < NB; ++c) v[c] = (c < nbe) ? ldcg_t<T>(A + row + (t + c) * (long long)ld) : Sc<T>::zero();
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                T acc0 = v[c], acc1 = Sc<T>::zero();
#pragma unroll
                for (int c2 = 0; c2 < c; ++c2) { if (c2 & 1) acc1 -= v[c2] * D[c2][c]; else acc0 -= v[c2] * D[c2][c]; }
                v[c] = (acc0 + acc1) * Sc<T>::inv(D[c][c]);
            }
#pragma unroll
            for (int c = 0; c < NB; ++c) if (c < nbe) A[row + (t + c) * (long long)ld] = v[c];
        } else {
            T* cp = A + (t + nbe + idx) * (long long)ld + t;
#pragma unroll
            for (int r = 0; r < NB; ++r) v[r] = (r < nbe) ? ldcg_t<T>(cp + r) : Sc<T>::zero();
#pragma unroll
            for (int r = 0; r < NB; ++r) {
                T acc0 = v[r], acc1 = Sc<T>::zero();
#pragma unroll
                for (int r2 = 0; r2 < r; ++r2) { if (r2 & 1) acc1 -= D[r][r2] * v[r2]; else acc0 -= D[r][r2] * v[r2]; }
                v[r] = acc0 + acc1;
            }
#pragma unroll
            for (int r = 0; r < NB; ++r) if (r < nbe) cp[r] = v[r];
        }
    }
}

template <typename T, int NB>
__global__ void __launch_bounds__(256, 2)
k_lu_update_fused(T* __restrict__ LU, size_t nel, long long n2, int p, int pw, int ld, long long t, T* __restrict__ scratch,
                  int tiles, int* __restrict__ cnt, int target) {
    __shared__ T Ls[NB][32];
    __shared__ T Us[NB][33];
    __shared__ int s_last;
    const int sys = blockIdx.x;
    T* A = LU + (size_t)sys * nel + pw;
    const int nbe = (int)min((long long)NB, n2 - t);
    const int R = (int)min((long long)p, n2 - (t + nbe));
    // tile order: the 2 tiles - 1 tiles of the first tile row / column first, then the rest
    const int nu1 = 2 * tiles - 1;
    int bx, by;
    const int L = blockIdx.y;
    if (L < nu1) { if (L < tiles) { bx = L; by = 0; } else { bx = 0; by = L - tiles + 1; } }
    else { int l2 = L - nu1; bx = 1 + l2 % (tiles - 1); by = 1 + l2 / (tiles - 1); }
    const int a0 = bx * 32, b0 = by * 32;
    const int tid = threadIdx.x;
    if (bx == 0 && by == 0) {
        // put the factored diagonal block (parked by the panel of this step) in place
        const T* S = scratch + (size_t)sys * NB * NB;
        for (int e = tid; e < nbe * nbe; e += 256) {
            int r = e % nbe, c = e / nbe;
            A[(t + r) + (t + c) * (long long)ld] = S[r + c * NB];
        }
    }
    if (a0 < R && b0 < R) {
        const long long r0 = t + nbe + a0, c0 = t + nbe + b0;
        const int ta = tid % 16, tb = tid / 16;
        const bool m0 = (a0 + ta < R), m1 = (a0 + ta + 16 < R), n0 = (b0 + tb < R), n1 = (b0 + tb + 16 < R);
        T* c00 = A + (r0 + ta) + (c0 + tb) * (long long)ld;
        T* c01 = A + (r0 + ta) + (c0 + tb + 16) * (long long)ld;
        T old00 = (m0 && n0) ? c00[0] : Sc<T>::zero(), old10 = (m1 && n0) ? c00[16] : Sc<T>::zero();
        T old01 = (m0 && n1) ? c01[0] : Sc<T>::zero(), old11 = (m1 && n1) ? c01[16] : Sc<T>::zero();
        for (int e = tid; e < NB * 32; e += 256) {
            int a = e % 32, c = e / 32;
            Ls[c][a] = (a0 + a < R && c < nbe) ? A[(r0 + a) + (t + c) * (long long)ld] : Sc<T>::zero();
        }
        for (int e = tid; e < NB * 32; e += 256) {
            int c = e % NB, b = e / NB;
            Us[c][b] = (b0 + b < R && c < nbe) ? A[(t + c) + (c0 + b) * (long long)ld] : Sc<T>::zero();
        }
        __syncthreads();
        T acc00 = Sc<T>::zero(), acc01 = Sc<T>::zero(), acc10 = Sc<T>::zero(), acc11 = Sc<T>::zero();
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            T l0 = Ls[c][ta], l1 = Ls[c][ta + 16];
            T u0 = Us[c][tb], u1 = Us[c][tb + 16];
            fmacc(acc00, l0, u0); fmacc(acc01, l0, u1); fmacc(acc10, l1, u0); fmacc(acc11, l1, u1);
        }
        if (m0 && n0) c00[0] = old00 - acc00;
        if (m1 && n0) c00[16] = old10 - acc10;
        if (m0 && n1) c01[0] = old01 - acc01;
        if (m1 && n1) c01[16] = old11 - acc11;
    }
    if (L >= nu1) return;
    // ---- tail: count this first-row/column tile off; the last one of the system factors the next panel ----
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(cnt + sys, 1) == target - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const long long tn = t + nbe;
    if (tn >= n2) return;
    lu_panel_tail<T, NB>(A, n2, p, ld, tn, scratch + (size_t)sys * NB * NB, reinterpret_cast<T (*)[NB + 1]>(&Ls[0][0]));
}

// Wide bands (p >= 128: the 3-D edge / face forms): 64 x 64 tiles with a 4 x 4 register block per thread.  A CTA of the 32 x 32
// kernel lives for 16 k-steps of 4 complex FMAs per thread -- its time is the load latency of its prologue; four times
// the work per CTA and per staged L / U element amortises it.  MEASURED: no gain (update chain 168 ms with either tile shape
// on 109x32x460; the chain costs 12.5 us per step + 0.95 us per system regardless of the tile) -- kept behind CS_LU_TILE=64.
template <typename T, int NB>
__global__ void __launch_bounds__(256, 2)
k_lu_update64(T* __restrict__ LU, size_t nel, long long n2, int p, int pw, int ld, long long t, const T* __restrict__ scratch, int part, int tiles) {
    __shared__ T Ls[NB][64];
    __shared__ T Us[NB][65];
    T* A = LU + (size_t)blockIdx.z * nel + pw;
    const int nbe = (int)min((long long)NB, n2 - t);
    const int R = (int)min((long long)p, n2 - (t + nbe));
    int bx = blockIdx.x, by = blockIdx.y;
    if (part == 1) { if (bx < tiles) by = 0; else { by = bx - tiles + 1; bx = 0; } }
    else if (part == 2) { bx += 1; by += 1; }
    const int a0 = bx * 64, b0 = by * 64;
    const int tid = threadIdx.x;
    if (bx == 0 && by == 0) {
        // put the factored diagonal block (parked by the panel kernel) in place
        const T* S = scratch + (size_t)blockIdx.z * NB * NB;
        for (int e = tid; e < nbe * nbe; e += 256) {
            int r = e % nbe, c = e / nbe;
            A[(t + r) + (t + c) * (long long)ld] = S[r + c * NB];
        }
    }
    if (a0 >= R || b0 >= R) return;
    const long long r0 = t + nbe + a0, c0 = t + nbe + b0;
    for (int e = tid; e < NB * 64; e += 256) {
        int a = e % 64, c = e / 64;
        Ls[c][a] = (a0 + a < R && c < nbe) ? A[(r0 + a) + (t + c) * (long long)ld] : Sc<T>::zero();
    }
    for (int e = tid; e < NB * 64; e += 256) {
        int c = e % NB, b = e / NB;
        Us[c][b] = (b0 + b < R && c < nbe) ? A[(t + c) + (c0 + b) * (long long)ld] : Sc<T>::zero();
    }
    const int ta = tid % 16, tb = tid / 16;           // rows ta + 16 i, columns tb + 16 j
    __syncthreads();
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = Sc<T>::zero();
#pragma unroll 4
    for (int c = 0; c < NB; ++c) {
        T l[4], u[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) l[i] = Ls[c][ta + 16 * i];
#pragma unroll
        for (int j = 0; j < 4; ++j) u[j] = Us[c][tb + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) fmacc(acc[i][j], l[i], u[j]);
    }
    // read-modify-write of the tile, one column of the register block at a time (the other resident CTA covers the latency;
    // holding all 16 old values through the k loop would spill at the 128-register cap of two CTAs per SM)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int b = b0 + tb + 16 * j;
        T* col = A + r0 + ta + (c0 + tb + 16 * j) * (long long)ld;
        T o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = (a0 + ta + 16 * i < R && b < R) ? col[16 * i] : Sc<T>::zero();
#pragma unroll
        for (int i = 0; i < 4; ++i) if (a0 + ta + 16 * i < R && b < R) col[16 * i] = o[i] - acc[i][j];
    }
}

// After the factorisation every NB x NB diagonal block holds L11 \ U11.  The triangular solves
// with those blocks are the sequential heart of the substitution kernels, so they are replaced
// once, here (fully parallel: one warp per block), by the explicit inverses
//   strictly lower part <- inv(L11) (unit diagonal implicit),  upper part <- inv(U11),
// which turns every block step of k_band_solve into two small dependency-free mat-vecs.
template <typename T, int NB>
__global__ void k_invert_diag(T* __restrict__ LU, size_t nel, long long n2, int pw, int ld) {
    __shared__ T D[NB][NB + 1];
    T* A = LU + (size_t)blockIdx.y * nel + pw;
    long long t = (long long)blockIdx.x * NB;
    int nbe = (int)min((long long)NB, n2 - t);
    int lane = threadIdx.x;
    for (int e = lane; e < NB * NB; e += 32) {
        int r = e % NB, c = e / NB;
        T val = (r == c) ? Sc<T>::one() : Sc<T>::zero();
        if (r < nbe && c < nbe) val = A[(t + r) + (t + c) * (long long)ld];
        D[r][c] = val;
    }
    __syncwarp();
    int j = lane & (NB - 1);
    const bool wr = lane < NB && j < nbe;
    {   // column j of inv(L), L unit lower  -> strictly lower part
        T xl[NB];
#pragma unroll
        for (int r = 0; r < NB; ++r) {
            T acc = (r == j) ? Sc<T>::one() : Sc<T>::zero();
#pragma unroll
            for (int k = 0; k < r; ++k) if (k >= j) acc -= D[r][k] * xl[k];
            xl[r] = (r >= j) ? acc : Sc<T>::zero();
        }
        if (wr) {
#pragma unroll
            for (int r = 0; r < NB; ++r) if (r < nbe && r > j) A[(t + r) + (t + j) * (long long)ld] = xl[r];
        }
    }
    {   // column j of inv(U), U upper with diagonal -> upper part (D still holds the factors)
        T xu[NB];
#pragma unroll
        for (int r = NB - 1; r >= 0; --r) {
            T acc = (r == j) ? Sc<T>::one() : Sc<T>::zero();
#pragma unroll
            for (int k = r + 1; k < NB; ++k) if (k <= j) acc -= D[r][k] * xu[k];
            xu[r] = (r <= j) ? acc * Sc<T>::inv(D[r][r]) : Sc<T>::zero();
        }
        if (wr) {
#pragma unroll
            for (int r = 0; r < NB; ++r) if (r < nbe && r <= j) A[(t + r) + (t + j) * (long long)ld] = xu[r];
        }
    }
}

// The factorisation is ~2 * n2 / NB tiny dependent launches (9 784 for the C2 sweep).  Issued
// one by one they make the solve hostage to host-thread jitter, so the whole sequence is
// captured once into a CUDA graph per (buffer, shape) and replayed with a single launch.
struct FactorGraphKey {
    void* LU; void* scratch; long long n2; int p, nsys, cplx, dev;
    bool operator==(const FactorGraphKey& o) const {
        return LU == o.LU && scratch == o.scratch && n2 == o.n2 && p == o.p && nsys == o.nsys && cplx == o.cplx && dev == o.dev;
    }
};
struct FactorGraph { FactorGraphKey key; cudaGraphExec_t exec; unsigned long long stamp; };
static std::vector<FactorGraph> g_graphs;
static unsigned long long g_stamp = 0;

// stB == nullptr: one stream, panel -> update per step.  Otherwise (inside stream capture) the look-ahead schedule:
//   stA:  P(t) -> [wait U2(t-1)] -> U1(t) -> P(t+1) ...        stB:  [wait P(t)] -> U2(t) ...
// U1(t) = the first tile row / column of the trailing update (everything panel t+1 reads), U2(t) = the rest, so the
// latency-bound panel of step t+1 runs concurrently with the throughput-bound bulk update of step t.
template <typename T, int NB>
static int band_factor_enqueue(const BandPlan& bp, T* LU, int nsys, T* scratch, cudaStream_t st, cudaStream_t stB = nullptr,
                               cudaEvent_t evP = nullptr, cudaEvent_t evU = nullptr, int* cnt = nullptr,
                               long long t_begin = 0, long long t_end = -1, size_t stride = 0, long long invert_n2 = -1) {
    // t_begin / t_end: block steps [t_begin, t_end) only (both multiples of NB or the end); stride: elements between
    // consecutive systems (0: band_elems()); invert_n2: k_invert_diag on the diagonal blocks below this row (-1: all, 0: none)
    if (t_end < 0) t_end = bp.n2;
    if (invert_n2 < 0) invert_n2 = bp.n2;
    size_t nel = stride ? stride : bp.band_elems();
    int ld = bp.ldab - 1;
    static const int tile_env = [] { const char* e = getenv("CS_LU_TILE"); return e ? atoi(e) : 0; }();
    const bool big = tile_env == 64 && NB == 16;      // opt-in: measured identical (273 ms vs 273 ms on 109x32x460), see below
    int tiles = big ? (bp.p + 63) / 64 : (bp.p + 31) / 32;
    int chunks = (bp.p + 63) / 64;
    // (timing experiments only: CS_LU_SKIP=panel|update drops one of the two kernels of every step -- wrong factors)
    static const int skip = [] { const char* e = getenv("CS_LU_SKIP"); return !e ? 0 : (!strcmp(e, "panel") ? 1 : (!strcmp(e, "update") ? 2 : 0)); }();
    const bool la = stB != nullptr && tiles >= 2;
    static const bool sym_env = [] { const char* e = getenv("CS_LU_SYM"); return !(e && atoi(e) == 0); }();
    const bool sym = bp.sym && sym_env && !la && !big;
    if (sym) chunks = (bp.p + 127) / 128;
    // pivot block of step t+1 factored by the update of step t (CS_LU_DAHEAD=0: by the panel kernel itself, as before)
    static const bool dahead_env = [] { const char* e = getenv("CS_LU_DAHEAD"); return !(e && atoi(e) == 0); }();
    const int dahead = (dahead_env && !la && !big) ? 1 : 0;      // (also under CS_LU_SKIP: timing experiments on garbage)
    static const bool panel4 = [] { const char* e = getenv("CS_LU_PANEL4"); return !(e && atoi(e) == 0); }();
    // MEASURED (109x32x460, 17 systems): 356 ms (443 ms at 3 CTAs/SM) vs 273 ms for the two-kernel schedule -- the one-CTA
    // panel tail of the last system is exposed and the lower occupancy slows the tiles; opt-in only (CS_LU_FUSED=1)
    static const bool fused_on = [] { const char* e = getenv("CS_LU_FUSED"); return e && atoi(e) == 1; }();
    if (fused_on && !la && !big && !sym && NB == 16 && tiles >= 2 && skip == 0 && cnt && t_begin == 0 && t_end == bp.n2 && invert_n2 == bp.n2) {
        // one launch per block step: update of step t + panel of step t+1 (k_lu_update_fused); the first panel stands alone
        if (cudaMemsetAsync(cnt, 0, (size_t)nsys * sizeof(int), st) != cudaSuccess) return CS_ERR;
        dim3 gp(chunks, nsys);
        k_lu_panel<T, NB><<<gp, 160, 0, st>>>(LU, nel, bp.n2, bp.p, bp.pw, ld, 0, scratch, 0, 0);
        int step = 0;
        for (long long t = 0; t < bp.n2; t += NB, ++step) {
            dim3 g(nsys, tiles * tiles, 1);
            k_lu_update_fused<T, NB><<<g, 256, 0, st>>>(LU, nel, bp.n2, bp.p, bp.pw, ld, t, scratch, tiles, cnt, (step + 1) * (2 * tiles - 1));
        }
        dim3 gi((unsigned)((bp.n2 + NB - 1) / NB), nsys);
        k_invert_diag<T, NB><<<gi, 32, 0, st>>>(LU, nel, bp.n2, bp.pw, ld);
        return CS_OK;
    }
    bool have_u2 = false;
    for (long long t = t_begin; t < t_end; t += NB) {
        dim3 gp(chunks, nsys);
        if (skip != 1) {
            if (NB == 16 && dahead && t > t_begin && panel4) {
                const int ch4 = (bp.p + 63) / 64;
                dim3 g4(sym ? ch4 : 2 * ch4, nsys);
                k_lu_panel4<T><<<g4, 256, 0, st>>>(LU, nel, bp.n2, bp.p, bp.pw, ld, t, (const T*)scratch, sym ? 1 : 0, ch4);
            } else k_lu_panel<T, NB><<<gp, 160, 0, st>>>(LU, nel, bp.n2, bp.p, bp.pw, ld, t, scratch, sym ? 1 : 0, (dahead && t > t_begin) ? 1 : 0);
        }
        if (skip == 2) continue;
        if (sym) {
            dim3 g(tiles * (tiles + 1) / 2, 1, nsys);
            k_lu_update<T, NB><<<g, 256, 0, st>>>(LU, nel, bp.n2, bp.p, bp.pw, ld, t, scratch, 3, tiles, dahead);
            continue;
        }
        if (!la) {
            dim3 g(tiles, tiles, nsys);
            if (big) k_lu_update64<T, 16><<<g, 256, 0, st>>>(LU, nel, bp.n2, bp.p, bp.pw, ld, t, (const T*)scratch, 0, tiles);
            else k_lu_update<T, NB><<<g, 256, 0, st>>>(LU, nel, bp.n2, bp.p, bp.pw, ld, t, scratch, 0, tiles, dahead);
            continue;
        }
        if (cudaEventRecord(evP, st) != cudaSuccess) return CS_ERR;
        if (have_u2 && cudaStreamWaitEvent(st, evU, 0) != cudaSuccess) return CS_ERR;       // U1(t) overlaps the region U2(t-1) wrote
        dim3 g1(2 * tiles - 1, 1, nsys);
        if (big) k_lu_update64<T, 16><<<g1, 256, 0, st>>>(LU, nel, bp.n2, bp.p, bp.pw, ld, t, (const T*)scratch, 1, tiles);
        else k_lu_update<T, NB><<<g1, 256, 0, st>>>(LU, nel, bp.n2, bp.p, bp.pw, ld, t, scratch, 1, tiles, 0);
        if (cudaStreamWaitEvent(stB, evP, 0) != cudaSuccess) return CS_ERR;
        dim3 g2(tiles - 1, tiles - 1, nsys);
        if (big) k_lu_update64<T, 16><<<g2, 256, 0, stB>>>(LU, nel, bp.n2, bp.p, bp.pw, ld, t, (const T*)scratch, 2, tiles);
        else k_lu_update<T, NB><<<g2, 256, 0, stB>>>(LU, nel, bp.n2, bp.p, bp.pw, ld, t, scratch, 2, tiles, 0);
        if (cudaEventRecord(evU, stB) != cudaSuccess) return CS_ERR;
        have_u2 = true;
    }
    if (la && have_u2 && cudaStreamWaitEvent(st, evU, 0) != cudaSuccess) return CS_ERR;      // join
    // diagonal blocks in the format of the substitution kernel (ITS block size, which may be smaller than the steps')
    if (invert_n2 > 0) {
        if (bp.nb == 16) {
            dim3 gi((unsigned)((invert_n2 + 15) / 16), nsys);
            k_invert_diag<T, 16><<<gi, 32, 0, st>>>(LU, nel, invert_n2, bp.pw, ld);
        } else {
            dim3 gi((unsigned)((invert_n2 + NB - 1) / NB), nsys);
            k_invert_diag<T, NB><<<gi, 32, 0, st>>>(LU, nel, invert_n2, bp.pw, ld);
        }
    }
    return CS_OK;
}

// ---- two-front ("twisted") elimination (BandPlan::twist) ----
// original index of entry i of half system `half` (-1: identity dummy)
// (padT / padB leading dummies; a trailing dummy pads the order to an even number when real factors are TMA-staged)
__device__ __forceinline__ long long tw_orig(int half, long long i, long long n2, long long padT, long long padB, long long nt, long long w) {
    if (half == 0) { const long long o = i - padT; return (o < 0 || o >= nt + w) ? -1 : o; }
    const long long o = n2 - 1 - (i - padB);
    return (i < padB || o < nt) ? -1 : o;
}
// LU2[f][m][half] = twisted halves of K[m] + s[f] diag(mdiag)  (see BandPlan::twist; the separator block goes to T, B's is zero)
template <typename TB, typename TL>
__global__ void k_form_twisted(const long long* __restrict__ map3d, const int* __restrict__ jstr, const double* __restrict__ mdiag,
                               long long n2, int pw, int ldab, int nmodes, long long n2h, long long nt, long long w, long long padT, long long padB,
                               const TB* __restrict__ K, const double* __restrict__ s_d, TL* __restrict__ LU, int mode_base) {
    // mode_base: true theta-mode index of stored mode 0 (the mode-sharded slab solve forms one mode at a time; the shared
    // axis unknown is a dummy in every mode but 0)
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long nelh = n2h * (long long)ldab;
    if (e >= nelh) return;
    const int m = blockIdx.y >> 1, half = blockIdx.y & 1, f = blockIdx.z;
    const long long q = e / ldab;
    const int d = (int)(e - q * ldab);
    const long long i = q + d - pw;
    TL v = Sc<TL>::zero();
    const long long oq = tw_orig(half, q, n2, padT, padB, nt, w);
    if (oq < 0) { if (d == pw) v = Sc<TL>::one(); }
    else if (i >= 0 && i < n2h) {
        const long long oi = tw_orig(half, i, n2, padT, padB, nt, w);
        const bool sq = oq >= nt && oq < nt + w, si = oi >= nt && oi < nt + w;
        if (oi >= 0 && !(half == 1 && sq && si)) {
            const TB kv = K[(size_t)m * ((size_t)n2 * ldab) + (size_t)(pw + oi - oq) + (size_t)oq * ldab];
            v = Sc<TL>::from(Sc<TB>::re(kv), Sc<TB>::im(kv));
            if (oi == oq) {
                const bool dummy = (map3d[oq] < 0) || (jstr[oq] == 0 && mode_base + m > 0);
                if (dummy) v = Sc<TL>::one();
                else v = v + Sc<TL>::from(s_d[2 * f], s_d[2 * f + 1]) * Sc<TL>::from(mdiag[oq], 0.0);
            }
        }
    }
    LU[(((size_t)f * nmodes + m) * 2 + half) * (size_t)nelh + e] = v;
}
int band_form_twisted(const BandPlan& bp, const double* s_d, int nshift, void* LU, bool lu_cplx, cudaStream_t st, int mode_base) {
    const long long nelh = (long long)bp.band_elems_h();
    dim3 g((unsigned)cdiv(nelh, 256), 2 * bp.nmodes, nshift);
#define CS_FORM_TW(TB, TL) k_form_twisted<TB, TL><<<g, 256, 0, st>>>(bp.map3d, bp.jstr, bp.mdiag, bp.n2, bp.pw, bp.ldab, bp.nmodes, bp.tw_n2h, bp.tw_nt, bp.tw_w, \
                                                                    bp.tw_padT, bp.tw_padB, (const TB*)bp.Kband, s_d, (TL*)LU, mode_base)
    if (bp.cplx_band) { CS_CHECK(lu_cplx, "complex band needs complex LU"); CS_FORM_TW(cplx, cplx); }
    else if (lu_cplx) CS_FORM_TW(double, cplx);
    else CS_FORM_TW(double, double);
#undef CS_FORM_TW
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}
// separator block of T += the Schur complement the bottom front left in B's separator block (reversed ordering).  Symmetric
// step LU: only the lower triangles (and the diagonal tiles) of the trailing blocks are up to date -- read B through its
// lower triangle and write both triangles of T.
template <typename T>
__global__ void k_twist_combine(T* __restrict__ LU, size_t nelh, long long M, long long w, int pw, int ld, int sym) {
    T* AT = LU + (size_t)blockIdx.y * 2 * nelh + pw;
    const T* AB = AT + nelh;
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= w * w) return;
    const long long i = e % w, j = e / w;
    if (sym && i < j) return;
    // B(M + w-1-i, M + w-1-j); symmetric: for i > j that is an upper entry of B -- take its mirror image
    long long bi = M + w - 1 - i, bj = M + w - 1 - j;
    if (sym && bi < bj) { long long t = bi; bi = bj; bj = t; }
    const T val = AT[(M + i) + (M + j) * (long long)ld] + AB[bi + bj * (long long)ld];
    AT[(M + i) + (M + j) * (long long)ld] = val;
    if (sym && i != j) AT[(M + j) + (M + i) * (long long)ld] = val;
}
template <typename T, int NB>
static int band_factor_twisted_enqueue(const BandPlan& bp, T* LU2, int nsys, T* scratch, cudaStream_t st) {
    BandPlan bh = bp;
    bh.n2 = bp.tw_n2h;
    const size_t nelh = bp.band_elems_h();
    const int ld = bp.ldab - 1;
    static const bool sym_env = [] { const char* e = getenv("CS_LU_SYM"); return !(e && atoi(e) == 0); }();
    // both fronts of every system: 2 nsys half systems, block steps [0, M)
    CS_TRY((band_factor_enqueue<T, NB>(bh, LU2, 2 * nsys, scratch, st, nullptr, nullptr, nullptr, nullptr, 0, bp.tw_M, 0, 0)));
    dim3 gc((unsigned)cdiv(bp.tw_w * bp.tw_w, 256), nsys);
    k_twist_combine<T><<<gc, 256, 0, st>>>(LU2, nelh, bp.tw_M, bp.tw_w, bp.pw, ld, (bp.sym && sym_env) ? 1 : 0);
    // the separator: the remaining steps of the T halves alone (stride of two half systems), then the diagonal blocks of T
    CS_TRY((band_factor_enqueue<T, NB>(bh, LU2, nsys, scratch, st, nullptr, nullptr, nullptr, nullptr, bp.tw_M, bp.tw_n2h, 2 * nelh, bp.tw_n2h)));
    // diagonal blocks of the eliminated part of the B halves (their separator block is never factored)
    CS_TRY((band_factor_enqueue<T, NB>(bh, LU2 + nelh, nsys, scratch, st, nullptr, nullptr, nullptr, nullptr, 0, 0, 2 * nelh, bp.tw_M)));
    return CS_OK;
}

template <typename T, int NB>
static int band_factor_t(const BandPlan& bp, T* LU, int nsys, cudaStream_t st) {
    CS_CHECK(bp.nbf == NB, "band plan step size must equal the instantiated NB");
    std::lock_guard<std::mutex> lock(g_band_mu);
    void* scratch = nullptr;
    const int tw = bp.tw_active ? 2 : 1;              // twisted: two half systems per system
    CS_TRY(band_scratch(0, st, (size_t)tw * nsys * NB * NB * sizeof(T), false, &scratch));
    void* cntv = nullptr;
    CS_TRY(band_scratch(2, st, (size_t)nsys * sizeof(int), true, &cntv));
    int* cnt = (int*)cntv;
    int dev_key = 0;
    CS_CUDA(cudaGetDevice(&dev_key));
    int groups_key = 0;
    { const char* e = getenv("CS_BAND_GROUPS"); if (e) groups_key = atoi(e); }
    { const char* e = getenv("CS_LU_LOOKAHEAD"); if (e && atoi(e) == 1) groups_key += 100; }
    { const char* e = getenv("CS_LU_TILE"); if (e) groups_key += 10000 * atoi(e); }
    { const char* e = getenv("CS_LU_FUSED"); if (e && atoi(e) == 1) groups_key += 1000000; }
    { const char* e = getenv("CS_LU_SYM"); if (bp.sym && !(e && atoi(e) == 0)) groups_key += 50; }
    { const char* e = getenv("CS_LU_DAHEAD"); if (e && atoi(e) == 0) groups_key += 25; }
    { const char* e = getenv("CS_LU_PANEL4"); if (e && atoi(e) == 0) groups_key += 12; }
    { const char* e = getenv("CS_LU_NBF"); if (e) groups_key += 6; }
    FactorGraphKey key{(void*)LU, scratch, bp.n2, bp.p, nsys, (int)(sizeof(T) == sizeof(cplx)) + 2 * NB + 128 * (tw - 1) + 1000 * groups_key, dev_key};
    for (auto& g : g_graphs)
        if (g.key == key) {
            g.stamp = ++g_stamp;
            CS_CUDA(cudaGraphLaunch(g.exec, st));
            return CS_OK;
        }
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool captured = false;
    // The chain is latency-bound and barely depends on the number of systems per launch, so the
    // systems are split into up to 4 groups whose chains run concurrently on forked streams
    // (inside the captured graph: parallel branches) and hide each other's latency.
    // (capture streams / events belong to a device: one set per device, created on first use under g_band_mu)
    static const int MAXG = 32;
    struct AuxSet { cudaStream_t aux[MAXG - 1] = {}, auxB[MAXG] = {}; cudaEvent_t ev_fork = nullptr, ev_join[MAXG - 1] = {}, evP[MAXG] = {}, evU[MAXG] = {}; };
    static std::map<int, AuxSet> aux_sets;
    int dev_now = 0;
    CS_CUDA(cudaGetDevice(&dev_now));
    AuxSet& as = aux_sets[dev_now];
    if (!as.ev_fork) {
        cudaEventCreateWithFlags(&as.ev_fork, cudaEventDisableTiming);
        for (int g = 0; g < MAXG - 1; ++g) { cudaStreamCreateWithFlags(&as.aux[g], cudaStreamNonBlocking); cudaEventCreateWithFlags(&as.ev_join[g], cudaEventDisableTiming); }
        for (int g = 0; g < MAXG; ++g) {
            cudaStreamCreateWithFlags(&as.auxB[g], cudaStreamNonBlocking);
            cudaEventCreateWithFlags(&as.evP[g], cudaEventDisableTiming); cudaEventCreateWithFlags(&as.evU[g], cudaEventDisableTiming);
        }
    }
    // measured on the 109x32x460 mesh (17 systems, p = 330): 271-273 ms with and without the look-ahead schedule, for 1, 2
    // and 4 stream groups -- the step time is the sum of the panel chain (11.6 us) and the update chain (17.8 us) either
    // way, so the simpler two-kernel schedule stays the default (CS_LU_LOOKAHEAD=1 opts in)
    static const bool lookahead = [] { const char* e = getenv("CS_LU_LOOKAHEAD"); return e && atoi(e) == 1; }();
    cudaStream_t* aux = as.aux; cudaEvent_t ev_fork = as.ev_fork; cudaEvent_t* ev_join = as.ev_join;
    int groups = (nsys >= 16) ? 4 : (nsys >= 6 ? 2 : 1);
    { const char* e = getenv("CS_BAND_GROUPS"); if (e && atoi(e) >= 1) groups = std::min(std::min(atoi(e), nsys), MAXG); }
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        bool okq = true;
        if (groups > 1) {
            okq = okq && cudaEventRecord(ev_fork, st) == cudaSuccess;
            for (int g = 1; g < groups; ++g) okq = okq && cudaStreamWaitEvent(aux[g - 1], ev_fork, 0) == cudaSuccess;
        }
        for (int g = 0; g < groups && okq; ++g) {
            int s0 = (int)((long long)nsys * g / groups), s1 = (int)((long long)nsys * (g + 1) / groups);
            cudaStream_t sg = (g == 0) ? st : aux[g - 1];
            if (bp.tw_active) okq = okq && band_factor_twisted_enqueue<T, NB>(bp, LU + (size_t)s0 * 2 * bp.band_elems_h(), s1 - s0, (T*)scratch + (size_t)s0 * 2 * NB * NB, sg) == CS_OK;
            else
            okq = okq && band_factor_enqueue<T, NB>(bp, LU + (size_t)s0 * bp.band_elems(), s1 - s0, (T*)scratch + (size_t)s0 * NB * NB, sg,
                                                    lookahead ? as.auxB[g] : nullptr, as.evP[g], as.evU[g], cnt + s0) == CS_OK;
        }
        for (int g = 1; g < groups; ++g) {
            okq = okq && cudaEventRecord(ev_join[g - 1], aux[g - 1]) == cudaSuccess;
            okq = okq && cudaStreamWaitEvent(st, ev_join[g - 1], 0) == cudaSuccess;
        }
        if (cudaStreamEndCapture(st, &graph) == cudaSuccess && graph && okq &&
            cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) captured = true;
        if (graph) cudaGraphDestroy(graph);
    }
    if (!captured) {
        cudaGetLastError();                 // capture unavailable: plain launches (same kernels)
        if (bp.tw_active) CS_TRY((band_factor_twisted_enqueue<T, NB>(bp, LU, nsys, (T*)scratch, st)));
        else
        CS_TRY((band_factor_enqueue<T, NB>(bp, LU, nsys, (T*)scratch, st, nullptr, nullptr, nullptr, cnt)));
        CS_CUDA(cudaGetLastError());
        return CS_OK;
    }
    if (g_graphs.size() >= 6) {             // keep a handful of shapes
        size_t old = 0;
        for (size_t i = 1; i < g_graphs.size(); ++i) if (g_graphs[i].stamp < g_graphs[old].stamp) old = i;
        cudaGraphExecDestroy(g_graphs[old].exec);
        g_graphs.erase(g_graphs.begin() + old);
    }
    g_graphs.push_back({key, exec, ++g_stamp});
    CS_CUDA(cudaGraphLaunch(exec, st));
    return CS_OK;
}
// ------------------------------------------------- cluster-resident LU ------
// The step-per-launch chain above is bound by launch boundaries and cold global round trips
// (~17 us per 16-column block step).  When the active window fits, the whole factorisation of
// one system is instead done by ONE thread-block cluster that keeps the window in (distributed)
// shared memory for the entire sweep:
//   * block column jb (NB columns = one contiguous NB*ldab chunk of the band array) is owned by
//     CTA jb % C; it is loaded ONCE (cp.async) when it enters the window, updated in shared
//     memory by the ~p/NB panels that touch it, factored by its owner and stored ONCE;
//   * the factorisation is the block LDU variant: the NB x NB pivot block is replaced by its full
//     inverse (in-place Gauss-Jordan, one element per thread), L21 = A21 inv(D), U12 = A12 stays
//     as it is -- so no triangular solves sit on the critical path and the trailing update reads
//     the pivot rows in place (BandPlan::diag_full tells k_band_solve about the format);
//   * the factored panel L21 is published through a small global (L2) buffer and a split cluster
//     barrier (arrive.release after publishing / wait.acquire after the CTA's own updates), so
//     the next owner's  update -> invert -> publish  chain is the only critical path and the
//     other CTAs' trailing updates overlap it (look-ahead).
// HBM traffic: the band array is read once and written once.
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cp_async16_cg(void* dst_smem, const void* src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8_ca(void* dst_smem, const void* src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_elem(cplx* d, const cplx* s) { cp_async16_cg(d, s); }
__device__ __forceinline__ void cp_async_elem(double* d, const double* s) { cp_async8_ca(d, s); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

static const int CL_THREADS = 256;
#ifndef CL_INV
#define CL_INV fast_inv
#endif
// optional phase timers (CS_BAND_PROF=1): thread 0 of every CTA accumulates clock64 deltas
#define CL_TICK(i) do { if (prof && threadIdx.x == 0) { long long t_ = clock64(); prof_s[i] += t_ - *tprev; *tprev = t_; } } while (0)

struct ClusterLuGeom {
    long long n2; int p, pw, ldab;
    int nblk;     // number of NB-column blocks
    int nbw;      // blocks a panel reaches to its right: ceil(p / NB)
    int sl;       // chunk slots per CTA: ceil((nbw + 1) / C)
    int prs;      // panel leading dimension: roundup(p, 4)
};

// One thread's TR x 4 register tile of the trailing update  A[i][j] -= L21[i][:] A12[:][j]  by panel kb.
// Lanes run along rows and a thread's TR rows are `ngr` apart, so every L / accumulator access of a
// warp is one contiguous run (no bank conflicts); the A12 rows are read in place from the chunk
// (same address across the warp: broadcast).
template <typename T, int NB, int TR>
struct ClTile {
    T* q; const T* uq; const T* lp;
    int ngr, ld, prs, rleft;      // rleft: rows left below this thread's first row
    bool cv[4];
    T acc[TR][4];
    // item `it` of the update of rows [rbeg, R) of the own blocks first, first + C, ...
    __device__ __forceinline__ void setup(const ClusterLuGeom& g, T* chunks, const T* Lb, int kb, int R, int rbeg,
                                          int first, int C, int it) {
        ld = g.ldab - 1; prs = g.prs;
        const size_t CH = (size_t)NB * g.ldab;
        ngr = (R - rbeg + TR - 1) / TR;
        int rg = it % ngr, cg = it / ngr;
        int ci0 = cg * 4, b = ci0 / NB, c0 = ci0 % NB;
        int jb = first + C * b;
        T* blk = chunks + (size_t)((jb / C) % g.sl) * CH + g.pw + (long long)c0 * ld;
        q = blk + (kb + 1 - jb) * NB + rbeg + rg;           // rows below the pivot block
        uq = blk + (kb - jb) * NB;                          // the pivot block's rows of these columns
        lp = Lb + rbeg + rg;
        rleft = R - rbeg - rg;
        long long gc0 = (long long)jb * NB + c0, lim = (long long)(kb + 1) * NB + R;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) cv[cc] = gc0 + cc < lim;   // columns the panel reaches
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
#pragma unroll
            for (int rr = 0; rr < TR; ++rr)
                acc[rr][cc] = (rr * ngr < rleft && cv[cc]) ? q[rr * ngr + cc * ld] : Sc<T>::zero();
    }
    __device__ __forceinline__ void step(int k) {
        T l[TR], u[4];
#pragma unroll
        for (int rr = 0; rr < TR; ++rr) l[rr] = lp[rr * ngr + k * prs];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) u[cc] = cv[cc] ? uq[k + cc * ld] : Sc<T>::zero();
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
#pragma unroll
            for (int rr = 0; rr < TR; ++rr) fmsub(acc[rr][cc], l[rr], u[cc]);
    }
    __device__ __forceinline__ void store() {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
#pragma unroll
            for (int rr = 0; rr < TR; ++rr)
                if (rr * ngr < rleft && cv[cc]) q[rr * ngr + cc * ld] = acc[rr][cc];
    }
};

// rows [rbeg, R) of `nown` own blocks by the thread group (tix of nthr)
template <typename T, int NB, int TR>
__device__ __forceinline__ void cl_trailing(const ClusterLuGeom& g, T* chunks, const T* Lb,
                                            int kb, int R, int rbeg, int first, int nown, int C, int tix, int nthr) {
    const int ngr = (R - rbeg + TR - 1) / TR, ngc = nown * (NB / 4);
    for (int it = tix; it < ngr * ngc; it += nthr) {
        ClTile<T, NB, TR> t;
        t.setup(g, chunks, Lb, kb, R, rbeg, first, C, it);
#pragma unroll 4
        for (int k = 0; k < NB; ++k) t.step(k);
        t.store();
    }
}

template <typename T, int NB>
__device__ __forceinline__ void cl_process(const ClusterLuGeom& g, T* chunks, const T* Lb,
                                           int kb, int R, int first, int nown, int C) {
    if (nown <= 0) return;                                   // (uniform per CTA)
    // few work items: the narrower register tile keeps all warps busy
    int items4 = ((R + 3) / 4) * (nown * NB / 4);
    if (items4 <= CL_THREADS / 2) cl_trailing<T, NB, 2>(g, chunks, Lb, kb, R, 0, first, nown, C, threadIdx.x, CL_THREADS);
    else cl_trailing<T, NB, 4>(g, chunks, Lb, kb, R, 0, first, nown, C, threadIdx.x, CL_THREADS);
}

// Factor block column kb (resident in `chunk`), optionally bringing it up to date with panel kb-1
// (in Lb) first -- the look-ahead step that is the critical path of the whole sweep:
//   1. the NB x NB pivot block gets its panel update (one element per thread);
//   2. warps 0-1: D <- inv(D) by an in-place Gauss-Jordan sweep without pivoting (4 elements per
//      thread, one 64-thread named barrier per pivot).  The sweep is a pure latency chain
//      (reciprocal -> scale -> eliminate per pivot), so meanwhile the other warps apply the panel
//      update to the rows BELOW the pivot block;
//   3. L21 <- A21 inv(D), written in place and published to P.
template <typename T, int NB>
__device__ __forceinline__ void cl_factor(const ClusterLuGeom& g, T* chunks, T* chunk, const T* Lb, bool update,
                                          T* __restrict__ P, T (*Di)[NB + 1], T (*rowb)[NB],
                                          T (*colb)[NB], int kb, int C, long long* prof, long long* prof_s, long long* tprev) {
    static_assert(NB == 16 && CL_THREADS == 256, "thread <-> element mappings below");
    const int ld = g.ldab - 1;
    T* sv = chunk + g.pw;
    const int tid = threadIdx.x;
    const int nbe = (int)min((long long)NB, g.n2 - (long long)kb * NB);
    const int R = (int)min((long long)g.p, g.n2 - ((long long)kb * NB + nbe));
    // rows of panel kb-1 that reach this block: its own NB rows + Rp - NB below
    const int Rp = update ? (int)min((long long)g.p, g.n2 - (long long)kb * NB) : 0;
    {
        const int r = tid % NB, c = tid / NB;
        T a = (r == c) ? Sc<T>::one() : Sc<T>::zero();       // a short last block is padded with the identity
        if (r < nbe && c < nbe) {
            a = sv[r + (long long)c * ld];
            if (update) {
                const T* uq = sv - NB + (long long)c * ld;   // rows of pivot block kb-1 in this column
#pragma unroll
                for (int k = 0; k < NB; ++k) fmsub(a, Lb[r + k * g.prs], uq[k]);
            }
        }
        Di[r][c] = a;
    }
    __syncthreads();
    CL_TICK(1);
    if (tid < 64) {
        // in-place Gauss-Jordan, thread (r, cq) holds row r, columns 4 cq .. 4 cq + 3.  Pivot k:
        //   a_kc <- a_kc / p;  a_rc -= (a_rk / p) a_kc;  a_rk <- -a_rk / p;  a_kk <- 1 / p
        // (a 2 x 2 block-pivot variant was tried: 8 instead of 16 dependent steps, but each step
        //  is so much longer for a lone warp that the sweep only went from 7.7k to 6.9k cycles.
        //  Also tried: L21 = A21 inv(D) by replaying the sweep's column operations on the A21 rows in
        //  warps 2..7 while the sweep runs -- the sweep slowed to 8.7k cycles and the replay still
        //  finished 4.1k cycles after it: the shadow of the sweep is too short for tile update + L21.)
        const int r = tid % NB, cq = tid / NB;
        T a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = Di[r][cq * 4 + j];
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            const int buf = k & 1, kq = k / 4, kj = k % 4;
            if (r == k) {
#pragma unroll
                for (int j = 0; j < 4; ++j) rowb[buf][cq * 4 + j] = a[j];
            }
            if (cq == kq) colb[buf][r] = a[kj];
            asm volatile("bar.sync 1, 64;" ::: "memory");
            T pv = CL_INV(rowb[buf][k]);
            // multiplier of the (unscaled) pivot row: rows != k eliminate with -a_rk / p, row k scales with 1 / p
            T t = (r == k) ? pv : Sc<T>::zero() - colb[buf][r] * pv;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = cq * 4 + j;
                if (col == k) a[j] = t;
                else {
                    T v = (r == k) ? Sc<T>::zero() : a[j];
                    fmacc(v, t, rowb[buf][col]);
                    a[j] = v;
                }
            }
        }
        asm volatile("bar.sync 1, 64;" ::: "memory");        // all reads of Di (pivot values) are long done; publish
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = cq * 4 + j;
            Di[r][col] = a[j];
            if (r < nbe && col < nbe) sv[r + (long long)col * ld] = a[j];
        }
        CL_TICK(3);
    } else if (Rp > NB) {
        // (the Gauss-Jordan warps 0 and 1 sit on SM sub-partitions 0 and 1; restricting this update to warps
        //  2, 3, 6, 7 -- sub-partitions 2, 3 -- made the sweep 5.5k instead of 6.6k cycles but the update
        //  7.2k instead of 4.6k: all six warps it is)
        long long tt0 = 0;
        if (prof && tid == 64) tt0 = clock64();
        cl_trailing<T, NB, 2>(g, chunks, Lb, kb - 1, Rp, NB, kb, 1, C, tid - 64, CL_THREADS - 64);
        if (prof && tid == 64) prof_s[9] += clock64() - tt0;
    }
    __syncthreads();
    CL_TICK(2);
    // L21 = A21 inv(D): thread -> (row, half of the columns); read everything, barrier, write in place
    const int HALF = NB / 2;
    for (int a0 = 0; a0 < R; a0 += CL_THREADS / 2) {
        int ar = a0 + (tid % (CL_THREADS / 2)), h = tid / (CL_THREADS / 2);
        T out[HALF];
        if (ar < R) {
            T row[NB];
#pragma unroll
            for (int c2 = 0; c2 < NB; ++c2) row[c2] = sv[NB + ar + (long long)c2 * ld];
#pragma unroll
            for (int cc = 0; cc < HALF; ++cc) {
                T acc = Sc<T>::zero();
#pragma unroll
                for (int c2 = 0; c2 < NB; ++c2) fmacc(acc, row[c2], Di[c2][h * HALF + cc]);
                out[cc] = acc;
            }
        }
        __syncthreads();
        if (ar < R) {
#pragma unroll
            for (int cc = 0; cc < HALF; ++cc) {
                int cq = h * HALF + cc;
                sv[NB + ar + (long long)cq * ld] = out[cc];
                P[ar + cq * g.prs] = out[cc];
            }
        }
    }
    __syncthreads();
    CL_TICK(4);
}

// where block columns come from: the already formed A(s) in the LU buffer (K == nullptr), or -- fused
// assembly -- the shift-independent band K[m] plus s_f * diag(mdiag) on the fly (k_form_shifted's rule),
// which saves writing and re-reading the nshift * nmodes shifted copies (10 GB each way for C2)
struct ClusterLuSource {
    const void* K; const double* s_d; const long long* map3d; const int* jstr; const double* mdiag; int nmodes;
};
template <typename T, typename TB, int NB>
__device__ __forceinline__ void cl_load_chunk(const ClusterLuGeom& g, T* chunk, const T* __restrict__ AB,
                                              const ClusterLuSource& src, int sys, int jb) {
    if (jb >= g.nblk) return;
    int ncols = (int)min((long long)NB, g.n2 - (long long)jb * NB);
    if (!src.K) {
        const T* a = AB + (size_t)jb * NB * g.ldab;
        for (int e = threadIdx.x; e < ncols * g.ldab; e += CL_THREADS) cp_async_elem(chunk + e, a + e);
        return;
    }
    const int m = sys % src.nmodes, f = sys / src.nmodes;
    const TB* kb = reinterpret_cast<const TB*>(src.K) + (size_t)m * g.n2 * g.ldab + (size_t)jb * NB * g.ldab;
    const T sv = Sc<T>::from(src.s_d[2 * f], src.s_d[2 * f + 1]);
    // all NB loads of a thread (one band offset d, every column) are issued before the first use
    for (int d = threadIdx.x; d < g.ldab; d += CL_THREADS) {
        TB kv[NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) if (c < ncols) kv[c] = kb[(size_t)c * g.ldab + d];
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            if (c < ncols) {
                T v = Sc<T>::from(Sc<TB>::re(kv[c]), Sc<TB>::im(kv[c]));
                if (d == g.pw) {
                    long long q = (long long)jb * NB + c;
                    bool dummy = (src.map3d[q] < 0) || (src.jstr[q] == 0 && m > 0);
                    v = dummy ? Sc<T>::one() : v + sv * Sc<T>::from(src.mdiag[q], 0.0);
                }
                chunk[(size_t)c * g.ldab + d] = v;
            }
        }
    }
}
// fused source, asynchronous: the raw K chunk travels with cp.async into `stg` (into the chunk itself when
// K already has the factor's scalar type); cl_finish_chunk converts it and adds the shift once it has landed
template <typename T, typename TB, int NB>
__device__ __forceinline__ void cl_load_chunk_async(const ClusterLuGeom& g, T* chunk, TB* stg, const ClusterLuSource& src, int sys, int jb) {
    int ncols = (int)min((long long)NB, g.n2 - (long long)jb * NB);
    const int m = sys % src.nmodes;
    const TB* kb = reinterpret_cast<const TB*>(src.K) + (size_t)m * g.n2 * g.ldab + (size_t)jb * NB * g.ldab;
    TB* dst = (sizeof(TB) == sizeof(T)) ? reinterpret_cast<TB*>(chunk) : stg;
    for (int e = threadIdx.x; e < ncols * g.ldab; e += CL_THREADS) cp_async_elem(dst + e, kb + e);
}
template <typename T, typename TB, int NB>
__device__ __forceinline__ void cl_finish_chunk(const ClusterLuGeom& g, T* chunk, const TB* stg, const ClusterLuSource& src, int sys, int jb) {
    int ncols = (int)min((long long)NB, g.n2 - (long long)jb * NB);
    const int m = sys % src.nmodes, f = sys / src.nmodes;
    const T sv = Sc<T>::from(src.s_d[2 * f], src.s_d[2 * f + 1]);
    if (sizeof(TB) != sizeof(T)) {
        for (int e = threadIdx.x; e < ncols * g.ldab; e += CL_THREADS) chunk[e] = Sc<T>::from(Sc<TB>::re(stg[e]), Sc<TB>::im(stg[e]));
        __syncthreads();
    }
    if ((int)threadIdx.x < ncols) {
        const int c = threadIdx.x;
        long long q = (long long)jb * NB + c;
        bool dummy = (src.map3d[q] < 0) || (src.jstr[q] == 0 && m > 0);
        T* dg = chunk + (size_t)c * g.ldab + g.pw;
        *dg = dummy ? Sc<T>::one() : *dg + sv * Sc<T>::from(src.mdiag[q], 0.0);
    }
}
template <typename T, int NB>
__device__ __forceinline__ void cl_store_chunk(const ClusterLuGeom& g, const T* chunk, T* __restrict__ AB, int jb) {
    int ncols = (int)min((long long)NB, g.n2 - (long long)jb * NB);
    T* dst = AB + (size_t)jb * NB * g.ldab;
    for (int e = threadIdx.x; e < ncols * g.ldab; e += CL_THREADS) dst[e] = chunk[e];
}

// grid = C * nsys CTAs in clusters of C; panels: per system 2 slots of prs * NB elements
template <typename T, typename TB, int NB>
__global__ void __launch_bounds__(CL_THREADS, 1)
k_band_lu_cluster(T* __restrict__ LU, size_t nel, ClusterLuGeom g, T* __restrict__ panels, ClusterLuSource src, long long* prof) {
    __shared__ long long prof_s[12];
    __shared__ long long tprev_s;
    long long* tprev = &tprev_s;
    if (threadIdx.x == 0) { for (int i = 0; i < 12; ++i) prof_s[i] = 0; tprev_s = clock64(); }
    extern __shared__ __align__(16) unsigned char cl_smem[];
    const int C = (int)cluster_nctarank(), rank = (int)cluster_ctarank();
    const int sys = blockIdx.x / C;
    T* AB = LU + (size_t)sys * nel;
    const size_t CH = (size_t)NB * g.ldab;
    T* chunks = reinterpret_cast<T*>(cl_smem);               // sl chunks (NB * ldab elements: 16-byte multiples)
    T* Lb = chunks + (size_t)g.sl * CH;                      // panel L21: prs x NB
    T (*Di)[NB + 1] = reinterpret_cast<T (*)[NB + 1]>(Lb + (size_t)g.prs * NB);
    T (*rowb)[NB] = reinterpret_cast<T (*)[NB]>(Di + NB);
    T (*colb)[NB] = rowb + 2;
    TB* stg = reinterpret_cast<TB*>(colb + 2);               // raw K chunk in flight (fused source with TB != T only)
    int pend_jb = -1;                                        // block column whose fused load still needs cl_finish_chunk
    T* pend_chunk = nullptr;
    T* P = panels + (size_t)sys * 2 * g.prs * NB;
    const int psz16 = (int)((size_t)g.prs * NB * sizeof(T) / 16);

    // prologue: the first window
    for (int jb = rank; jb < g.sl * C && jb < g.nblk; jb += C) cl_load_chunk<T, TB, NB>(g, chunks + (size_t)((jb / C) % g.sl) * CH, AB, src, sys, jb);
    cp_async_wait_all();
    __syncthreads();
    if (rank == 0) {
        cl_factor<T, NB>(g, chunks, chunks, Lb, false, P, Di, rowb, colb, 0, C, nullptr, prof_s, tprev);
        cl_store_chunk<T, NB>(g, chunks, AB, 0);
        __syncthreads();
        if (src.K) { if (g.sl * C < g.nblk) { cl_load_chunk_async<T, TB, NB>(g, chunks, stg, src, sys, g.sl * C); pend_jb = g.sl * C; pend_chunk = chunks; } }
        else cl_load_chunk<T, TB, NB>(g, chunks, AB, src, sys, g.sl * C);
    }
    cluster_arrive();
    cluster_wait();
    if (threadIdx.x == 0) tprev_s = clock64();

    for (int kb = 0; kb + 1 < g.nblk; ++kb) {
        // panel kb -> shared (L2 only: it was written by another SM)
        {
            const uint4* src = reinterpret_cast<const uint4*>(P + (size_t)(kb & 1) * g.prs * NB);
            uint4* dst = reinterpret_cast<uint4*>(Lb);
            for (int e = threadIdx.x; e < psz16; e += CL_THREADS) cp_async16_cg(dst + e, src + e);
        }
        cp_async_wait_all();                                 // also the chunk that entered during the previous step
        __syncthreads();
        if (pend_jb >= 0) {                                  // (uniform per CTA)
            cl_finish_chunk<T, TB, NB>(g, pend_chunk, stg, src, sys, pend_jb);
            pend_jb = -1;
            __syncthreads();
        }
        CL_TICK(0);
        const int R = (int)min((long long)g.p, g.n2 - (long long)(kb + 1) * NB);
        const int last = min(kb + g.nbw, g.nblk - 1);        // right-most block panel kb reaches
        int first = kb + 1 + (((rank - (kb + 1)) % C) + C) % C;
        int nown = (first <= last) ? (last - first) / C + 1 : 0;
        if (rank == (kb + 1) % C) {
            // look-ahead: bring block kb+1 up to date, factor and publish it, then do the rest
            T* chunk = chunks + (size_t)(((kb + 1) / C) % g.sl) * CH;
            cl_factor<T, NB>(g, chunks, chunk, Lb, true, P + (size_t)((kb + 1) & 1) * g.prs * NB, Di, rowb, colb, kb + 1, C, prof, prof_s, tprev);
            cluster_arrive();
            cl_process<T, NB>(g, chunks, Lb, kb, R, first + C, nown - 1, C);
            __syncthreads();
            CL_TICK(5);
            cl_store_chunk<T, NB>(g, chunk, AB, kb + 1);
            __syncthreads();
            if (src.K) { if (kb + 1 + g.sl * C < g.nblk) { cl_load_chunk_async<T, TB, NB>(g, chunk, stg, src, sys, kb + 1 + g.sl * C); pend_jb = kb + 1 + g.sl * C; pend_chunk = chunk; } }
            else cl_load_chunk<T, TB, NB>(g, chunk, AB, src, sys, kb + 1 + g.sl * C);
            CL_TICK(6);
        } else {
            cluster_arrive();
            cl_process<T, NB>(g, chunks, Lb, kb, R, first, nown, C);
            __syncthreads();
            CL_TICK(7);
        }
        __syncthreads();
        cluster_wait();
        CL_TICK(8);
    }
    if (prof && threadIdx.x == 0) for (int i = 0; i < 12; ++i) prof[(size_t)blockIdx.x * 12 + i] = prof_s[i];
}

struct ClusterLuPlan { bool ok; int C; ClusterLuGeom g; size_t smem; };
// stage_bytes: size of one element of the raw K band if the fused source needs a conversion buffer, else 0
template <typename T, int NB>
static ClusterLuPlan cluster_lu_plan(const BandPlan& bp, size_t stage_bytes = 0) {
    ClusterLuPlan pl{};
    pl.ok = false;
    const char* e = getenv("CS_BAND_LU");
    if (e && !strcmp(e, "steps")) return pl;
    // block LDU with explicitly inverted pivot blocks is only as accurate as those blocks are well
    // conditioned (3-D FDEM-h test: cond ~ 1e10 curl-curl blocks -> backward error 3e-8): scalar fields only
    if (!bp.scalar_field) return pl;
    ClusterLuGeom& g = pl.g;
    g.n2 = bp.n2; g.p = bp.p; g.pw = bp.pw; g.ldab = bp.ldab;
    g.nblk = (int)((bp.n2 + NB - 1) / NB);
    g.nbw = (bp.p + NB - 1) / NB;
    g.prs = ((bp.p + 3) / 4) * 4;
    if (bp.pw < g.nbw * NB || bp.p < 1) return pl;
    const int cands[3] = {4, 8, 2};
    for (int ci = 0; ci < 3; ++ci) {
        int C = cands[ci];
        if (C == 2 && g.nbw + 1 > 2) continue;
        if (C != 2 && g.nbw + 1 <= 2) continue;
        int sl = (g.nbw + 1 + C - 1) / C;
        if (sl > 4) continue;
        size_t CH = (size_t)NB * bp.ldab;
        size_t el = (size_t)sl * CH + (size_t)g.prs * NB + (size_t)NB * (NB + 1) + 4 * (size_t)NB;
        size_t bytes = el * sizeof(T) + CH * stage_bytes;
        if (bytes > 216 * 1024) continue;
        pl.ok = true; pl.C = C; g.sl = sl; pl.smem = bytes;
        return pl;
    }
    return pl;
}

template <typename T, typename TB, int NB>
static int band_factor_cluster(const BandPlan& bp, const ClusterLuPlan& pl, T* LU, int nsys, const ClusterLuSource& src, cudaStream_t st) {
    std::lock_guard<std::mutex> lock(g_band_mu);
    void* panels = nullptr;
    CS_TRY(band_scratch(1, st, (size_t)nsys * 2 * pl.g.prs * NB * sizeof(T), true, &panels));
    CS_CUDA(cudaFuncSetAttribute(k_band_lu_cluster<T, TB, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nsys * pl.C));
    cfg.blockDim = dim3(CL_THREADS);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = pl.C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    long long* prof = nullptr;
    const char* pe = getenv("CS_BAND_PROF");
    if (pe && atoi(pe) > 0) CS_CUDA(cudaMalloc(&prof, (size_t)nsys * pl.C * 12 * sizeof(long long)));
    CS_CUDA(cudaLaunchKernelEx(&cfg, k_band_lu_cluster<T, TB, NB>, LU, bp.band_elems(), pl.g, (T*)panels, src, prof));
    if (prof) {
        // debugging aid: average cycles per block step and phase (0 panel load, 1 pivot-block update, 3 Gauss-Jordan,
        // 2 wait for the row update, 4 L21 + publish, 5 owner's remaining update, 6 chunk store/load, 7 non-owner update, 8 barrier wait)
        std::vector<long long> h((size_t)nsys * pl.C * 12);
        CS_CUDA(cudaStreamSynchronize(st));
        CS_CUDA(cudaMemcpy(h.data(), prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(prof);
        double acc[12] = {0};
        for (size_t i = 0; i < h.size(); ++i) acc[i % 12] += (double)h[i];
        fprintf(stderr, "[cluster LU] C=%d sl=%d nblk=%d cycles/step (summed over the cluster's CTAs):", pl.C, pl.g.sl, pl.g.nblk);
        for (int i = 0; i < 10; ++i) fprintf(stderr, " p%d=%.0f", i, acc[i] / nsys / pl.g.nblk);
        fprintf(stderr, "\n");
    }
    return CS_OK;
}

template <typename T>
static int band_factor_any(const BandPlan& bp, T* LU, int nsys, cudaStream_t st) {
    if (bp.nb == 16) {
        ClusterLuPlan pl = cluster_lu_plan<T, 16>(bp);
        if (pl.ok) { bp.diag_full = 1; ClusterLuSource none{}; return band_factor_cluster<T, T, 16>(bp, pl, LU, nsys, none, st); }
        bp.diag_full = 0;
        if (bp.nbf == 32) return band_factor_t<T, 32>(bp, LU, nsys, st);
        return band_factor_t<T, 16>(bp, LU, nsys, st);
    }
    CS_CHECK(bp.nb == 32, "band plan block size must be 16 or 32");
    bp.diag_full = 0;
    return band_factor_t<T, 32>(bp, LU, nsys, st);
}
// A(s_f) = K + s_f diag(mdiag) formed and factored for nshift shifts.  Scalar fields: fused into the cluster
// kernel's chunk loads (no shifted copies in HBM); otherwise k_form_shifted + the step kernels.
// diagonal of every (half) system: finite?
template <typename T>
__global__ void k_diag_finite(const T* __restrict__ LU, size_t nel, long long n2, int pw, int ldab, int* __restrict__ bad) {
    const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q >= n2) return;
    const T v = LU[(size_t)blockIdx.y * nel + pw + q * (long long)ldab];
    if (!isfinite(Sc<T>::abs2(v))) *bad = 1;
}
int band_factor_finite(const BandPlan& bp, const void* LU, bool lu_cplx, int nshift, bool* ok, cudaStream_t st) {
    void* buf = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_band_mu);
        CS_TRY(band_scratch(3, st, 2 * sizeof(unsigned long long), false, &buf));
    }
    CS_CUDA(cudaMemsetAsync(buf, 0, sizeof(int), st));
    const long long n2 = bp.tw_active ? bp.tw_n2h : bp.n2;
    const size_t nel = bp.tw_active ? bp.band_elems_h() : bp.band_elems();
    dim3 g((unsigned)cdiv(n2, 256), (unsigned)(nshift * bp.nmodes * (bp.tw_active ? 2 : 1)));
    if (lu_cplx) k_diag_finite<cplx><<<g, 256, 0, st>>>((const cplx*)LU, nel, n2, bp.pw, bp.ldab, (int*)buf);
    else k_diag_finite<double><<<g, 256, 0, st>>>((const double*)LU, nel, n2, bp.pw, bp.ldab, (int*)buf);
    CS_CUDA(cudaGetLastError());
    int h = 0;
    CS_CUDA(cudaMemcpyAsync(&h, buf, sizeof h, cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaStreamSynchronize(st));
    *ok = (h == 0);
    return CS_OK;
}

int band_form_and_factor(const BandPlan& bp, const double* s_d, int nshift, void* LU, bool lu_cplx, cudaStream_t st, bool allow_twist) {
    bp.tw_active = bp.twist && allow_twist;
    if (bp.tw_active) {
        CS_TRY(band_form_twisted(bp, s_d, nshift, LU, lu_cplx, st, 0));
        bp.diag_full = 0;
        if (lu_cplx) return band_factor_t<cplx, 16>(bp, (cplx*)LU, nshift * bp.nmodes, st);
        return band_factor_t<double, 16>(bp, (double*)LU, nshift * bp.nmodes, st);
    }
    const char* e = getenv("CS_BAND_FUSE");
    bool fuse = !(e && atoi(e) == 0) && bp.nb == 16;
    if (fuse) {
        ClusterLuSource src{bp.Kband, s_d, bp.map3d, bp.jstr, bp.mdiag, bp.nmodes};
        const int nsys = nshift * bp.nmodes;
        if (lu_cplx) {
            ClusterLuPlan pl = cluster_lu_plan<cplx, 16>(bp, bp.cplx_band ? 0 : sizeof(double));
            if (pl.ok) {
                bp.diag_full = 1;
                if (bp.cplx_band) return band_factor_cluster<cplx, cplx, 16>(bp, pl, (cplx*)LU, nsys, src, st);
                return band_factor_cluster<cplx, double, 16>(bp, pl, (cplx*)LU, nsys, src, st);
            }
        } else if (!bp.cplx_band) {
            ClusterLuPlan pl = cluster_lu_plan<double, 16>(bp);
            if (pl.ok) { bp.diag_full = 1; return band_factor_cluster<double, double, 16>(bp, pl, (double*)LU, nsys, src, st); }
        }
    }
    CS_TRY(band_form_shifted(bp, s_d, nshift, LU, lu_cplx, st));
    return band_factor(bp, LU, lu_cplx, nshift * bp.nmodes, st);
}
int band_factor_launches(const BandPlan& bp, bool lu_cplx) {
    if (bp.nb == 16) {
        bool ok = lu_cplx ? cluster_lu_plan<cplx, 16>(bp).ok : cluster_lu_plan<double, 16>(bp).ok;
        if (ok) return 1;
    }
    if (bp.tw_active) return 4 + 2 * (int)((bp.tw_n2h + bp.nbf - 1) / bp.nbf);      // both fronts share the launches of a step
    return 1 + 2 * (int)((bp.n2 + bp.nbf - 1) / bp.nbf);
}
int band_factor(const BandPlan& bp, void* LU, bool lu_cplx, int nsys_tot, cudaStream_t st) {
    if (lu_cplx) return band_factor_any<cplx>(bp, (cplx*)LU, nsys_tot, st);
    return band_factor_any<double>(bp, (double*)LU, nsys_tot, st);
}

// ---------------------------------------------------------- banded solve ----
// One CTA per (vector, mode): forward then backward substitution in the
// column-oriented (axpy) form.  The block column of L / U needed by step s+1 is
// prefetched into shared memory with cp.async while step s computes (double
// buffer); the active window of the right-hand side lives in a circular
// shared-memory buffer, so the sequential critical path touches no global memory.
// Thread layout of k_band_solve: 4 threads per staged row (quarter q = tid & 3 owns block
// columns q, q+4, q+8, q+12), so a block step is 4 cp.async + 4 complex FMAs + a 2-stage
// shuffle reduction per thread instead of 16 + 16 sequential ones, and ~14 warps share the SM.
template <typename TL, int TPR, int NB>
__device__ __forceinline__ void prefetch_cols(TL* __restrict__ dst, int ldd, const TL* __restrict__ A, int ld,
                                              long long row0, int nrows, long long col0, int ncols) {
    // dst[r + c*ldd] = A[(row0 + r) + (col0 + c)*ld]
    const int r = threadIdx.x / TPR, q = threadIdx.x % TPR;
    if (r < nrows) {
        const TL* src = A + (row0 + r) + col0 * (long long)ld;
        TL* d = dst + r;
#pragma unroll
        for (int i = 0; i < NB / TPR; ++i) {
            int c = q + TPR * i;
            if (c < ncols) __pipeline_memcpy_async(d + c * ldd, src + c * (long long)ld, sizeof(TL));
        }
    }
}

template <typename T> __device__ __forceinline__ T shfl_xor_t(T v, int m);
template <> __device__ __forceinline__ double shfl_xor_t<double>(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <> __device__ __forceinline__ cplx shfl_xor_t<cplx>(cplx v, int m) {
    return mk(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}

// ---- producer warps + mbarrier (block columns of the factors -> shared memory ring) ----
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src, unsigned bytes, unsigned long long* bar) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(src), "r"(bytes), "r"(b) : "memory");
}
// TMA staging by the producer warp: column c is copied by producer thread c (complex only: 16-byte
// elements, so every column segment satisfies the alignment / size rules of cp.async.bulk).
// Measured on C2 (16 copies of 2 KB per step): one warp, one copy per lane 14.6 ms per solve;
// lane 0 of 2 or 4 warps 16.1 ms; one thread issuing all 16 copies 20.4 ms; cp.async by two
// producer warps 25 ms; cp.async spread over all compute warps 17.8 ms.
template <typename TL>
__device__ __forceinline__ void tma_stage_cols(TL* dst, int ldd, const TL* A, int ld, long long row0, int nrows,
                                               long long col0, int ncols, unsigned long long* bar, int tix, int nthr) {
    const unsigned bytes = (unsigned)(nrows * sizeof(TL));
    if (tix == 0) mbar_expect_tx(bar, bytes * ncols);
    __syncwarp();
    for (int c = tix; c < ncols; c += nthr) tma_bulk_g2s(dst + (size_t)c * ldd, A + row0 + (col0 + c) * (long long)ld, bytes, bar);
}
// the executing thread arrives on `bar` once all cp.async it has issued so far have landed
__device__ __forceinline__ void cp_async_mbar_arrive(unsigned long long* bar) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(a) : "memory");
}
// a thread group (tix of nthr) stages ncols columns of nrows contiguous elements: dst[r + c*ldd] = A[(row0 + r) + (col0 + c)*ld]
template <typename TL>
__device__ __forceinline__ void stage_cols_group(TL* dst, int ldd, const TL* __restrict__ A, int ld, long long row0, int nrows,
                                                 long long col0, int ncols, int tix, int nthr) {
    for (int c = 0; c < ncols; ++c) {
        const TL* src = A + row0 + (col0 + c) * (long long)ld;
        TL* d = dst + (size_t)c * ldd;
        for (int r = tix; r < nrows; r += nthr) cp_async_elem(d + r, src + r);
    }
}

__device__ __forceinline__ void bar_sync0() { asm volatile("bar.sync 0;" ::: "memory"); }

// HP: a dedicated producer warp (threads prod0 .. prod0 + 31, no compute) stages the factors.
//   It issues the copies of stage s + nst - 1 while the other warps compute step s and the compute
//   warps wait on the stage's mbarrier.  Complex systems stage with TMA bulk copies (one 2 KB column
//   segment per producer lane, expect_tx / complete_tx); real systems (8-byte elements: segments are
//   not 16-byte aligned) with cp.async + cp.async.mbarrier.arrive.
//   The producer runs its own loop; both sides meet at `bar.sync 0` once per block step.
// !HP: every thread issues its share of the stage with cp.async (commit / wait groups): the 64 LDGSTS
//   warp instructions of a stage (8 LSU cycles each) then sit right in front of the latency-critical
//   shared-memory loads of the step.
template <typename TL, typename TV, int TPR, int NB, bool HP>
__global__ void __launch_bounds__(TPR == 1 ? 1024 : (TPR == 2 ? 768 : 640))
k_band_solve(const TL* __restrict__ LU, size_t nel, long long n2, int p, int pw, int ld,
                             int nmodes, int nmv, int fold, int nth, const int* __restrict__ fmap, const int* __restrict__ shift_of_vec, TV* __restrict__ xm, int wsz, int nst, int diag_full, int prod0, int use_tma,
                             int tw, int half0, int fs0, int fs1, int bs1, int given, int flush) {
    // tw: twisted storage -- two half systems per (shift, mode) and per mode vector, this CTA works on half0 + blockIdx.z.
    // Forward block steps [fs0, fs1), then the first bs1 backward steps (counted from the end).  `given`: in the first
    // `given` backward steps the block solution is taken from the vector as it is (the separator solution computed by the
    // other half) and only carried to the rows above.  `flush`: after the forward sweep the partial sums of the p rows
    // below its last block go back to the vector (the separator rows of a twisted half).  One-front call: 0, 0, 0, ns, ns, 0, 0.
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // TMA bulk copies need 16-byte aligned column segments: always true for complex factors, for real ones when the plan
    // has an even half bandwidth and an even system size (BandPlan::tma_real); otherwise cp.async
    const bool USE_TMA = use_tma != 0;
    __shared__ unsigned long long full_bar[8];
    __shared__ TV yw[32][NB];                                 // per-warp copy of the block solution
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (HP) {
        if (tid == 0) {
            for (int i = 0; i < 8; ++i) mbar_init(&full_bar[i], USE_TMA ? 1u : 32u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        bar_sync0();
    }
    // mode vector mm of 3-D vector v is solved with the stored factor h: all nth modes of a complex vector fold onto
    // min(mm, nth - mm) (K''_{nth-m} = K''_m), the (Re, Im) pairs of a real vector's modes share factor mm / 2
    const int v = blockIdx.y, mm = blockIdx.x;
    const int m = fold == 0 ? mm : (fold == 1 ? min(mm, nth - mm) : (fold == 2 ? (mm >> 1) : fmap[mm]));
    const int f = shift_of_vec ? shift_of_vec[v] : 0;
    const int hf = half0 + (int)blockIdx.z;
    const TL* A = tw ? LU + (((size_t)f * nmodes + m) * 2 + hf) * nel + pw : LU + ((size_t)f * nmodes + m) * nel + pw;
    TV* b = tw ? xm + (((size_t)v * nmv + mm) * 2 + hf) * n2 : xm + ((size_t)v * nmv + mm) * n2;
    // rows per staged block column, padded so that the TPR column groups of a quarter warp fall into
    // distinct shared-memory banks (TPR = 4: ldd = 2 mod 8, TPR = 2: ldd = 4 mod 8)
    int ldd = NB + p;
    if (TPR == 4) while ((ldd & 7) != 2) ++ldd;
    if (TPR == 2) while ((ldd & 7) != 4) ++ldd;
    const size_t ssz = (size_t)ldd * NB;                      // elements per stage
    TL* ring = reinterpret_cast<TL*>(smem_raw);               // [nst][NB][ldd]
    TV* win = reinterpret_cast<TV*>(ring + (size_t)nst * ssz);   // circular window, wsz = power of two >= p + 12 NB
    const int wm = wsz - 1;
    const int ns = (int)((n2 + NB - 1) / NB);
    const long long tlast = (long long)(ns - 1) * NB;

    // stage of forward step s: rows [t, t + nbe + R) of columns [t, t + nbe);  backward: rows [t - Ra, t + nbe)
    auto stage_geom = [&](bool fwd, int s, long long& row0, int& nrows, long long& col0, int& ncols) {
        long long t2 = fwd ? (long long)s * NB : tlast - (long long)s * NB;
        int nbe2 = (int)min((long long)NB, n2 - t2);
        col0 = t2; ncols = nbe2;
        if (fwd) { row0 = t2; nrows = nbe2 + (int)min((long long)p, n2 - (t2 + nbe2)); }
        else { int Ra2 = (int)min((long long)p, t2); row0 = t2 - Ra2; nrows = Ra2 + nbe2; }
    };

    if (HP && tid >= prod0) {
        // ------------------------------ producer warp ------------------------------
        auto issue = [&](bool fwd, int s) {
            if (s >= (fwd ? fs1 : bs1)) return;
            long long row0, col0; int nrows, ncols;
            stage_geom(fwd, s, row0, nrows, col0, ncols);
            TL* dst = ring + (size_t)(s % nst) * ssz;
            if (USE_TMA) tma_stage_cols<TL>(dst, ldd, A, ld, row0, nrows, col0, ncols, &full_bar[s % nst], lane, 32);
            else {
                stage_cols_group<TL>(dst, ldd, A, ld, row0, nrows, col0, ncols, lane, 32);
                cp_async_mbar_arrive(&full_bar[s % nst]);
            }
        };
        for (int q = 0; q < nst - 1; ++q) issue(true, fs0 + q);
        for (int s = fs0; s < fs1; ++s) { bar_sync0(); issue(true, s + nst - 1); if (!diag_full) bar_sync0(); }
        bar_sync0();                                          // forward done: every ring slot is free again
        if (flush) bar_sync0();
        for (int q = 0; q < nst - 1; ++q) issue(false, q);
        for (int s = 0; s < bs1; ++s) { bar_sync0(); issue(false, s + nst - 1); bar_sync0(); }   // (two barriers per backward step)
        return;
    }

    // ------------------------------ compute warps ------------------------------
    constexpr int HALVES = 32 / NB;                           // NB = 16: two lanes share a row (column parity)
    const int r16 = lane % NB, half = lane / NB;              // block mat-vec: lane <-> (row, column parity)
    const int urow = tid / TPR, uq = tid % TPR;               // window update: thread <-> (row, column group)
    const TV zero = Sc<TV>::zero();
    unsigned phasebits = 0;
    // the NB right-hand-side rows that enter the window at step s (first touched at step s + 2) are
    // fetched with cp.async by warp 0, nst - 1 steps ahead: no register-staged global load on the critical path
    auto rows_prefetch = [&](bool fwd, int s) {
        if (s < ns && tid < NB) {
            long long e = fwd ? (long long)s * NB + 3 * NB + p + tid : n2 - (long long)(5 + s) * NB - p + tid;
            if (e >= 0 && e < n2) cp_async_elem(&win[e & wm], &b[e]);
        }
    };
    auto prefetch = [&](bool fwd, int s) {
        if (!HP && s < (fwd ? fs1 : bs1)) {
            long long row0, col0; int nrows, ncols;
            stage_geom(fwd, s, row0, nrows, col0, ncols);
            prefetch_cols<TL, TPR, NB>(ring + (size_t)(s % nst) * ssz, ldd, A, ld, row0, nrows, col0, ncols);
        }
        if (!HP || warp == 0) { rows_prefetch(fwd, s); __pipeline_commit(); }
    };
    auto stage_wait = [&](int slot) {
        if (HP) { mbar_wait(&full_bar[slot], (phasebits >> slot) & 1u); phasebits ^= 1u << slot; }
        if (!HP || warp == 0) __pipeline_wait_prior(nst - 2);
    };

    // ================= forward: L y = b (unit lower; diagonal blocks hold inv(L11), or nothing to apply) ==========
    if (fs1 > fs0) {
        const long long t0 = (long long)fs0 * NB;
        for (long long e = t0 + tid; e < t0 + 3 * NB + p && e < n2; e += (HP ? prod0 : blockDim.x)) win[e & wm] = b[e];
    }
    for (int q = 0; q < nst - 1; ++q) prefetch(true, fs0 + q);
    {
        int slot = fs0 % nst;
        for (int s = fs0; s < fs1; ++s) {
            const long long t = (long long)s * NB;
            const int nbe = (int)min((long long)NB, n2 - t);
            const int R = (int)min((long long)p, n2 - (t + nbe));
            const TL* cur = ring + (size_t)slot * ssz;
            stage_wait(slot);                                 // stage s has landed
            slot = (slot + 1 == nst) ? 0 : slot + 1;
            bar_sync0();                                      // ... and is visible; window of step s-1 complete
            prefetch(true, s + nst - 1);                      // refills the slot step s-1 just released
            // block LDU factors (diag_full) have a unit block diagonal in L: y1 is the window itself.  Scalar-LU factors need y1 = inv(L11) b1: warp 0 computes it once and hands it over
            // through shared memory + one more barrier (see the backward pass).
            const TL* crow = cur + nbe + urow;
            TV acc = zero;
            if (diag_full) {
                // y1 is the (finished) window itself: read it in place
                if (warp == 0 && lane < nbe) b[t + lane] = win[(t + lane) & wm];
                if (urow < R) {
#pragma unroll
                    for (int i = 0; i < NB / TPR; ++i) {
                        int c = uq + TPR * i;
                        if (c < nbe) fmacc(acc, crow[c * ldd], win[(t + c) & wm]);
                    }
                }
            } else {
                TL lreg[NB / TPR];
#pragma unroll
                for (int i = 0; i < NB / TPR; ++i) {
                    int c = uq + TPR * i;
                    lreg[i] = (urow < R && c < nbe) ? crow[c * ldd] : Sc<TL>::zero();
                }
                if (warp == 0) {
                    TV yr = (half == 0) ? win[(t + r16) & wm] : zero;
#pragma unroll
                    for (int i = 0; i < NB / HALVES; ++i) {
                        int c = HALVES * i + half;
                        if (c < r16) fmacc(yr, cur[r16 + c * ldd], win[(t + c) & wm]);
                    }
                    if (HALVES == 2) yr += shfl_xor_t<TV>(yr, 16);
                    if (lane < NB) yw[0][lane] = yr;
                    if (lane < nbe) b[t + lane] = yr;
                }
                bar_sync0();
#pragma unroll
                for (int i = 0; i < NB / TPR; ++i) {
                    int c = uq + TPR * i;
                    fmacc(acc, lreg[i], yw[0][c]);
                }
            }
            if (TPR >= 2) acc += shfl_xor_t<TV>(acc, 1);
            if (TPR >= 4) acc += shfl_xor_t<TV>(acc, 2);
            if (uq == 0 && urow < R) win[(t + nbe + urow) & wm] -= acc;
        }
    }
    __pipeline_wait_prior(0);
    bar_sync0();
    if (flush) {
        // twisted half: the rows below the last forward block (the separator) hold partial sums -- back to the vector
        const long long t1 = (long long)fs1 * NB;
        for (long long e = t1 + tid; e < t1 + p && e < n2; e += (HP ? prod0 : blockDim.x)) b[e] = win[e & wm];
        bar_sync0();
    }

    // ================= backward: U x = y (diagonal blocks hold inv(U11), or the full pivot-block inverse) =========
    // window now slides downwards; (re)load the last p + 4 NB rows of y
    if (bs1 > 0)
        for (long long e = max(0LL, n2 - (4 * NB + p)) + tid; e < n2; e += (HP ? prod0 : blockDim.x)) win[e & wm] = b[e];
    for (int q = 0; q < nst - 1; ++q) prefetch(false, q);
    {
        int slot = 0;
        for (int s = 0; s < bs1; ++s) {
            const long long t = tlast - (long long)s * NB;
            const int nbe = (int)min((long long)NB, n2 - t);
            const int Ra = (int)min((long long)p, t);
            const TL* cur = ring + (size_t)slot * ssz;        // rows [t - Ra, t + nbe), local row = row - (t - Ra)
            stage_wait(slot);
            slot = (slot + 1 == nst) ? 0 : slot + 1;
            bar_sync0();
            prefetch(false, s + nst - 1);
            // the block solution x1 = inv(U11) w (or inv(D) w) is computed by warp 0 alone and handed over through
            // shared memory + one more barrier: done redundantly by every warp (as the forward pass does with its
            // cheaper y1) the 16 x 16 block costs 24 LDS.128 per warp, ~1 500 LSU cycles per step for 16 warps
            TL lreg[NB / TPR];
#pragma unroll
            for (int i = 0; i < NB / TPR; ++i) {
                int c = uq + TPR * i;
                lreg[i] = (urow < Ra && c < nbe) ? cur[urow + c * ldd] : Sc<TL>::zero();
            }
            if (warp == 0) {
                TV xr = zero;
#pragma unroll
                for (int i = 0; i < NB / HALVES; ++i) {
                    int c = HALVES * i + half;
                    if ((diag_full || c >= r16) && c < nbe && r16 < nbe) fmacc(xr, cur[(Ra + r16) + c * ldd], win[(t + c) & wm]);
                }
                if (HALVES == 2) xr += shfl_xor_t<TV>(xr, 16);
                if (s < given) xr = (lane < nbe) ? b[t + lane] : zero;       // the separator solution, as given
                if (lane < NB) yw[0][lane] = xr;
                if (lane < nbe && s >= given) b[t + lane] = xr;
            }
            bar_sync0();
            TV acc = zero;
#pragma unroll
            for (int i = 0; i < NB / TPR; ++i) {
                int c = uq + TPR * i;
                fmacc(acc, lreg[i], yw[0][c]);
            }
            if (TPR >= 2) acc += shfl_xor_t<TV>(acc, 1);
            if (TPR >= 4) acc += shfl_xor_t<TV>(acc, 2);
            if (uq == 0 && urow < Ra) win[(t - Ra + urow) & wm] -= acc;
        }
    }
    __pipeline_wait_prior(0);
}

struct SolveRange { int tw = 0, half0 = 0, nhz = 1, fs0 = 0, fs1 = -1, bs1 = -1, given = 0, flush = 0; };

template <typename T, int TPR, int NB>
static int band_solve_launch(const BandPlan& bp, const T* LU, T* xm, int nvec, const int* sov, int th, size_t sm, int wsz, int nst, cudaStream_t st,
                             SolveRange rg = SolveRange()) {
    // one extra warp that only stages the factors, if the launch bounds leave room (CS_SOLVE_PROD=0: never).
    // Complex systems only: the producer of a real system has to use cp.async (no 16-byte alignment for TMA),
    // and one warp issuing all of a stage's cp.async is much slower than all warps sharing them
    // (C1 DC, 4 band solves: 141 ms with a producer warp vs 54 ms without).
    const int bound = TPR == 1 ? 1024 : (TPR == 2 ? 768 : 640);
    const char* e = getenv("CS_SOLVE_PROD");
    const bool tma = sizeof(T) == 16 || bp.tma_real;
    const bool hp = (tma || (e && atoi(e) > 0)) && (th + 32 <= bound) && !(e && atoi(e) == 0);
    if (hp) CS_CUDA(cudaFuncSetAttribute(k_band_solve<T, T, TPR, NB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 204 * 1024));
    else CS_CUDA(cudaFuncSetAttribute(k_band_solve<T, T, TPR, NB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 204 * 1024));
    const bool xc = sizeof(T) == sizeof(cplx);
    const int nmv = bp.nmv(xc), fold = bp.fold(xc);
    const long long n2 = rg.tw ? bp.tw_n2h : bp.n2;
    const size_t nel = rg.tw ? bp.band_elems_h() : bp.band_elems();
    const int ns = (int)((n2 + NB - 1) / NB);
    if (rg.fs1 < 0) rg.fs1 = ns;
    if (rg.bs1 < 0) rg.bs1 = ns;
    dim3 g(nmv, nvec, rg.nhz);
    if (hp) k_band_solve<T, T, TPR, NB, true><<<g, th + 32, sm, st>>>(LU, nel, n2, bp.p, bp.pw, bp.ldab - 1, bp.nmodes, nmv, fold, bp.nth, bp.fmap, sov, xm, wsz, nst, bp.diag_full, th, tma ? 1 : 0,
                                                                    rg.tw, rg.half0, rg.fs0, rg.fs1, rg.bs1, rg.given, rg.flush);
    else k_band_solve<T, T, TPR, NB, false><<<g, th, sm, st>>>(LU, nel, n2, bp.p, bp.pw, bp.ldab - 1, bp.nmodes, nmv, fold, bp.nth, bp.fmap, sov, xm, wsz, nst, bp.diag_full, 0, 0,
                                                             rg.tw, rg.half0, rg.fs0, rg.fs1, rg.bs1, rg.given, rg.flush);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// ---- twisted substitution: vectors between the one-front ordering xm [vec][mv][n2] and the halves xh [vec][mv][2][n2h] ----
template <typename T>
__global__ void k_twist_gather(const T* __restrict__ xm, T* __restrict__ xh, long long n2, long long n2h, long long M, long long padT, long long padB, long long nt, long long w) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n2h) return;
    const size_t vm = blockIdx.y;                       // (vector, mode vector) pair
    const int half = blockIdx.z;
    const long long o = tw_orig(half, i, n2, padT, padB, nt, w);
    T v = Sc<T>::zero();
    if (o >= 0 && !(half == 1 && i >= M)) v = xm[vm * n2 + o];        // (the separator right-hand side travels with T)
    xh[(vm * 2 + half) * n2h + i] = v;
}
template <typename T>
__global__ void k_twist_scatter(T* __restrict__ xm, const T* __restrict__ xh, long long n2, long long n2h, long long M, long long padT, long long padB, long long nt, long long w) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n2h) return;
    const size_t vm = blockIdx.y;
    const int half = blockIdx.z;
    if (half == 1 && i >= M) return;                    // the separator is written from T
    const long long o = tw_orig(half, i, n2, padT, padB, nt, w);
    if (o >= 0) xm[vm * n2 + o] = xh[(vm * 2 + half) * n2h + i];
}
// separator rows: dir 0: T += reversed B (partial sums of the two forward sweeps); dir 1: B = reversed T (the separator solution)
template <typename T>
__global__ void k_twist_sep(T* __restrict__ xh, long long n2h, long long M, long long w, int dir) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= w) return;
    T* t = xh + ((size_t)blockIdx.y * 2) * n2h + M;
    T* b = t + n2h;
    if (dir == 0) t[i] = t[i] + b[w - 1 - i];
    else b[w - 1 - i] = t[i];
}

template <typename T, int NB>
static int band_solve_t(const BandPlan& bp, const T* LU, T* xm, int nvec, const int* sov, cudaStream_t st) {
    int rowsn = bp.p + NB;
    int tpr = (4 * rowsn <= 640) ? 4 : (2 * rowsn <= 768 ? 2 : 1);   // C2: 4 -> 11.2 ms, 2 -> 11.4 ms, 1 -> 13.8 ms per solve
    { const char* e = getenv("CS_SOLVE_TPR"); if (e && (atoi(e) == 1 || atoi(e) == 2) && atoi(e) * rowsn <= 768) tpr = atoi(e); }
    int th = ((tpr * rowsn + 31) / 32) * 32;
    if (th < 64) th = 64;
    CS_CHECK(th <= 1024, "half bandwidth too large for the solve kernel (p + NB <= 1024)");
    // circular window: p + 4 NB live rows + the rows of up to 6 in-flight stages that land early
    int wsz = 64;
    while (wsz < bp.p + 12 * NB) wsz <<= 1;
    int ldd = rowsn;                             // padded leading dimension of a stage (see k_band_solve)
    if (tpr == 4) while ((ldd & 7) != 2) ++ldd;
    if (tpr == 2) while ((ldd & 7) != 4) ++ldd;
    size_t stage_b = (size_t)ldd * NB * sizeof(T);
    int nst = (int)((200 * 1024 - (size_t)wsz * sizeof(T)) / stage_b);   // ring depth: as deep as shared memory allows
    if (nst > 6) nst = 6;
    CS_CHECK(nst >= 2, "half bandwidth too large for the shared-memory staged solve");
    size_t sm = (size_t)nst * stage_b + (size_t)wsz * sizeof(T);
    auto launch = [&](T* x, SolveRange rg) -> int {
        if (tpr == 4) return band_solve_launch<T, 4, NB>(bp, LU, x, nvec, sov, th, sm, wsz, nst, st, rg);
        if (tpr == 2) return band_solve_launch<T, 2, NB>(bp, LU, x, nvec, sov, th, sm, wsz, nst, st, rg);
        return band_solve_launch<T, 1, NB>(bp, LU, x, nvec, sov, th, sm, wsz, nst, st, rg);
    };
    if (!bp.tw_active) return launch(xm, SolveRange());
    // two fronts: both halves sweep forward at once, the separator is solved on T, both halves sweep backward at once
    const bool xc = sizeof(T) == sizeof(cplx);
    const size_t nvm = (size_t)nvec * bp.nmv(xc);
    void* xhv = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_band_mu);
        CS_TRY(band_scratch(4, st, nvm * 2 * bp.tw_n2h * sizeof(T), false, &xhv));
    }
    T* xh = (T*)xhv;
    const long long n2h = bp.tw_n2h, M = bp.tw_M, w = bp.tw_w;
    const int ns = (int)((n2h + NB - 1) / NB), S1 = (int)(M / NB);
    dim3 gg((unsigned)cdiv(n2h, 256), (unsigned)nvm, 2), gs((unsigned)cdiv(w, 256), (unsigned)nvm);
    k_twist_gather<T><<<gg, 256, 0, st>>>(xm, xh, bp.n2, n2h, M, bp.tw_padT, bp.tw_padB, bp.tw_nt, w);
    SolveRange r1; r1.tw = 1; r1.nhz = 2; r1.fs0 = 0; r1.fs1 = S1; r1.bs1 = 0; r1.flush = 1;
    CS_TRY(launch(xh, r1));
    k_twist_sep<T><<<gs, 256, 0, st>>>(xh, n2h, M, w, 0);
    SolveRange r2; r2.tw = 1; r2.nhz = 1; r2.fs0 = S1; r2.fs1 = ns; r2.bs1 = ns - S1;
    CS_TRY(launch(xh, r2));
    k_twist_sep<T><<<gs, 256, 0, st>>>(xh, n2h, M, w, 1);
    SolveRange r3; r3.tw = 1; r3.nhz = 2; r3.fs0 = 0; r3.fs1 = 0; r3.bs1 = ns; r3.given = ns - S1;
    CS_TRY(launch(xh, r3));
    k_twist_scatter<T><<<gg, 256, 0, st>>>(xm, xh, bp.n2, n2h, M, bp.tw_padT, bp.tw_padB, bp.tw_nt, w);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// block size of a plan: 32 halves the length of the dependent step chains, but the staged block
// column (p + NB) x NB x 16 B must fit the shared-memory ring at least twice
int band_choose_nb(int p) {
    const char* e = getenv("CS_BAND_NB");
    // 32-column block steps halve the chain length but each step costs 3x (measured on C2:
    // factor 137 ms vs 84 ms, solve 25 vs 22 ms), so 16 is the default; CS_BAND_NB=32 opts in.
    size_t stage32 = (size_t)(p + 32) * 32 * sizeof(cplx);
    if (e && atoi(e) == 32 && 2 * stage32 + 16 * 1024 <= 200 * 1024) return 32;
    return 16;
}

void band_plan_twist(BandPlan& bp) {
    bp.twist = false;
    const char* tw_env = getenv("CS_TWIST");          // (read per plan: CS_TWIST=0 keeps the one-front elimination)
    const int on = tw_env ? atoi(tw_env) : 1;
    if (!on || bp.Lrow <= 0 || bp.nb != 16 || bp.nbf != 16) return;
    if (cluster_lu_plan<cplx, 16>(bp).ok || cluster_lu_plan<double, 16>(bp).ok) return;      // narrow bands: the cluster-resident LU
    const long long levels = bp.n2 / bp.Lrow;
    if (levels < 8 || bp.p > bp.Lrow + 16 || bp.Lrow < 16) return;                           // (the separator must decouple the two sides)
    const long long ks = levels / 2;
    bp.tw_nt = bp.Lrow * ks;
    bp.tw_w = bp.Lrow;
    const long long nbot = bp.n2 - bp.tw_nt - bp.tw_w;
    const long long mx = std::max(bp.tw_nt, nbot);
    bp.tw_M = (mx + 15) / 16 * 16;
    bp.tw_padT = bp.tw_M - bp.tw_nt;
    bp.tw_padB = bp.tw_M - nbot;
    bp.tw_n2h = bp.tw_M + bp.tw_w;
    if (bp.tma_real && (bp.tw_n2h & 1)) bp.tw_n2h += 1;                                      // (TMA staging of real factors needs an even order: one trailing dummy)
    bp.twist = true;
}

// Factorisation steps of 32 columns with the substitution's 16-column blocks (CS_LU_NBF=32): half as many dependent launch
// pairs.  MEASURED on 109x32x460 (17 systems, p = 330, symmetric steps): 215 ms vs 177 ms with 16-column steps -- a 32-column
// step costs 47.7 us against 2 x 19.6 us: the row solves of the panel (4 x the dependent arithmetic per row) sit on the
// critical path, not the launches.  Opt-in only.
int band_choose_nbf(int p, int nb) {
    const char* e = getenv("CS_LU_NBF");
    if (nb != 16) return nb;
    if (e && atoi(e) == 32 && p >= 64) return 32;
    return 16;
}

int band_solve(const BandPlan& bp, const void* LU, bool lu_cplx, void* xm, bool xm_cplx, int nvec, const int* shift_of_vec_d, cudaStream_t st) {
    CS_CHECK(lu_cplx == xm_cplx, "LU and mode vectors must have the same scalar type");
    if (bp.nb == 32) {
        if (lu_cplx) return band_solve_t<cplx, 32>(bp, (const cplx*)LU, (cplx*)xm, nvec, shift_of_vec_d, st);
        return band_solve_t<double, 32>(bp, (const double*)LU, (double*)xm, nvec, shift_of_vec_d, st);
    }
    CS_CHECK(bp.nb == 16, "band plan block size must be 16 or 32");
    if (lu_cplx) return band_solve_t<cplx, 16>(bp, (const cplx*)LU, (cplx*)xm, nvec, shift_of_vec_d, st);
    return band_solve_t<double, 16>(bp, (const double*)LU, (double*)xm, nvec, shift_of_vec_d, st);
}

}  // namespace cs
