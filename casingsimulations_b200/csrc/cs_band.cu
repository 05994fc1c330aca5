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
#include <cstdlib>

namespace cs {

template <typename T> __device__ __forceinline__ T shfl_t(T v, int src);
template <> __device__ __forceinline__ double shfl_t<double>(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <> __device__ __forceinline__ cplx shfl_t<cplx>(cplx v, int src) {
    return mk(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
// broadcast within each half-warp group of 16 lanes (lane r <-> row r of a 16 x 16 block)
template <typename T> __device__ __forceinline__ T shfl_row(T v, int src) { return shfl_t<T>(v, src); }

static const int TPB = 128;
#define TWO_PI 6.28318530717958647692

// ------------------------------------------------------------------ DFT -----
// xm[v][m][q] = sum_j x3d[v][map3d[q] + j*jstr[q]] * exp(-2 pi i j m / nth)
template <typename TX, typename TM>
__global__ void k_dft_gather(const long long* __restrict__ map3d, const int* __restrict__ jstr, long long n2, int nth,
                             const TX* __restrict__ x3d, long long n3d, TM* __restrict__ xm) {
    long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int v = blockIdx.y;
    if (q >= n2) return;
    x3d += v * n3d;
    xm += (long long)v * nth * n2;
    long long b = map3d[q];
    int st = jstr[q];
    if (b < 0) {
        for (int m = 0; m < nth; ++m) xm[m * n2 + q] = Sc<TM>::zero();
        return;
    }
    if (st == 0 || nth == 1) {
        TX val = x3d[b];
        xm[q] = Sc<TM>::from(Sc<TX>::re(val), Sc<TX>::im(val));
        for (int m = 1; m < nth; ++m) xm[m * n2 + q] = Sc<TM>::zero();
        return;
    }
    for (int m = 0; m < nth; ++m) {
        double re = 0.0, im = 0.0;
        for (int j = 0; j < nth; ++j) {
            TX val = x3d[b + (long long)j * st];
            double sn, cs_;
            sincospi(-2.0 * (double)((j * m) % nth) / (double)nth, &sn, &cs_);
            double vr = Sc<TX>::re(val), vi = Sc<TX>::im(val);
            re += vr * cs_ - vi * sn;
            im += vr * sn + vi * cs_;
        }
        xm[m * n2 + q] = Sc<TM>::from(re, im);
    }
}

// x3d[v][map3d[q] + j*jstr[q]] = (1/nth) sum_m xm[v][m][q] exp(+2 pi i j m / nth)
template <typename TX, typename TM>
__global__ void k_dft_scatter(const long long* __restrict__ map3d, const int* __restrict__ jstr, long long n2, int nth,
                              const TM* __restrict__ xm, TX* __restrict__ x3d, long long n3d) {
    long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int v = blockIdx.y;
    if (q >= n2) return;
    x3d += v * n3d;
    xm += (long long)v * nth * n2;
    long long b = map3d[q];
    int st = jstr[q];
    if (b < 0) return;
    if (st == 0 || nth == 1) {
        TM val = xm[q];
        x3d[b] = Sc<TX>::from(Sc<TM>::re(val), Sc<TM>::im(val));
        return;
    }
    double inv = 1.0 / (double)nth;
    for (int j = 0; j < nth; ++j) {
        double re = 0.0, im = 0.0;
        for (int m = 0; m < nth; ++m) {
            TM val = xm[m * n2 + q];
            double sn, cs_;
            sincospi(2.0 * (double)((j * m) % nth) / (double)nth, &sn, &cs_);
            double vr = Sc<TM>::re(val), vi = Sc<TM>::im(val);
            re += vr * cs_ - vi * sn;
            im += vr * sn + vi * cs_;
        }
        x3d[b + (long long)j * st] = Sc<TX>::from(re * inv, im * inv);
    }
}

int band_dft_gather(const BandPlan& bp, const void* x3d, bool x_cplx, void* xm, bool xm_cplx, int nth, long long n3d, int nvec, cudaStream_t st) {
    dim3 g((unsigned)cdiv(bp.n2, TPB), nvec);
    if (x_cplx && xm_cplx) k_dft_gather<cplx, cplx><<<g, TPB, 0, st>>>(bp.map3d, bp.jstr, bp.n2, nth, (const cplx*)x3d, n3d, (cplx*)xm);
    else if (!x_cplx && xm_cplx) k_dft_gather<double, cplx><<<g, TPB, 0, st>>>(bp.map3d, bp.jstr, bp.n2, nth, (const double*)x3d, n3d, (cplx*)xm);
    else if (!x_cplx && !xm_cplx) {
        CS_CHECK(nth == 1, "real mode vectors need nth == 1");
        k_dft_gather<double, double><<<g, TPB, 0, st>>>(bp.map3d, bp.jstr, bp.n2, nth, (const double*)x3d, n3d, (double*)xm);
    } else CS_CHECK(false, "complex vector into real mode storage");
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}
int band_dft_scatter(const BandPlan& bp, const void* xm, bool xm_cplx, void* x3d, bool x_cplx, int nth, long long n3d, int nvec, cudaStream_t st) {
    dim3 g((unsigned)cdiv(bp.n2, TPB), nvec);
    if (x_cplx && xm_cplx) k_dft_scatter<cplx, cplx><<<g, TPB, 0, st>>>(bp.map3d, bp.jstr, bp.n2, nth, (const cplx*)xm, (cplx*)x3d, n3d);
    else if (!x_cplx && xm_cplx) k_dft_scatter<double, cplx><<<g, TPB, 0, st>>>(bp.map3d, bp.jstr, bp.n2, nth, (const cplx*)xm, (double*)x3d, n3d);
    else if (!x_cplx && !xm_cplx) {
        CS_CHECK(nth == 1, "real mode vectors need nth == 1");
        k_dft_scatter<double, double><<<g, TPB, 0, st>>>(bp.map3d, bp.jstr, bp.n2, nth, (const double*)xm, (double*)x3d, n3d);
    } else CS_CHECK(false, "complex mode storage into real vector needs cplx modes");
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// ------------------------------------------------------------- probing ------
// Columns of the s-independent part K are obtained by applying the *same*
// stencil kernel that the Krylov loop uses to 2p+1 "coloured" probe vectors
// (ones at j = 0 for every band index q == color mod (2p+1)); the DFT of the
// response holds column q of every mode matrix at once.
template <typename TX>
__global__ void k_probe_fill(const long long* __restrict__ map3d, long long n2, int ncol, int color, TX* __restrict__ x3d) {
    long long q = color + (blockIdx.x * (long long)blockDim.x + threadIdx.x) * ncol;
    if (q >= n2) return;
    long long b = map3d[q];
    if (b >= 0) x3d[b] = Sc<TX>::one();
}
int band_probe_fill(const BandPlan& bp, int color, void* x3d, long long n3d, bool x_cplx, cudaStream_t st) {
    int ncol = 2 * bp.p + 1;
    CS_CUDA(cudaMemsetAsync(x3d, 0, (size_t)n3d * (x_cplx ? sizeof(cplx) : sizeof(double)), st));
    long long cnt = (bp.n2 + ncol - 1) / ncol;
    if (x_cplx) k_probe_fill<cplx><<<cdiv(cnt, TPB), TPB, 0, st>>>(bp.map3d, bp.n2, ncol, color, (cplx*)x3d);
    else k_probe_fill<double><<<cdiv(cnt, TPB), TPB, 0, st>>>(bp.map3d, bp.n2, ncol, color, (double*)x3d);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

template <typename TB>
__global__ void k_probe_store(const long long* __restrict__ map3d, const int* __restrict__ jstr, long long n2, int p, int pw, int ldab,
                              int color, const TB* __restrict__ ym, TB* __restrict__ K) {
    long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int m = blockIdx.y;
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
    K[(size_t)m * n2 * ldab + (size_t)(pw + r - q) + (size_t)q * ldab] = ym[(size_t)m * n2 + r];
}
int band_probe_store(const BandPlan& bp, int color, const void* ym, cudaStream_t st) {
    dim3 g((unsigned)cdiv(bp.n2, TPB), bp.nmodes);
    if (bp.cplx_band) k_probe_store<cplx><<<g, TPB, 0, st>>>(bp.map3d, bp.jstr, bp.n2, bp.p, bp.pw, bp.ldab, color, (const cplx*)ym, (cplx*)bp.Kband);
    else k_probe_store<double><<<g, TPB, 0, st>>>(bp.map3d, bp.jstr, bp.n2, bp.p, bp.pw, bp.ldab, color, (const double*)ym, (double*)bp.Kband);
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

template <typename TB>
__global__ void k_diag_weights(const long long* __restrict__ map3d, const int* __restrict__ jstr, const double* __restrict__ mdiag,
                               long long n2, int pw, int ldab, int nmodes, const TB* __restrict__ K,
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
    for (int m = 0; m < nm; ++m) {
        TB kv = K[(size_t)m * nel + (size_t)pw + (size_t)q * ldab];
        re += Sc<TB>::re(kv); im += Sc<TB>::im(kv);
    }
    re /= nm; im /= nm;
    re += s_d[2 * f] * mdiag[q];
    im += s_d[2 * f + 1] * mdiag[q];
    double ww = 1.0 / sqrt(re * re + im * im);
    double* wf = w + (size_t)f * n3d;
    if (st == 0) { wf[b] = ww; return; }
    for (int j = 0; j < nmodes; ++j) wf[b + (long long)j * st] = ww;
}
int band_diag_weights(const BandPlan& bp, const double* s_d, int nshift, long long n3d, double* w, cudaStream_t st) {
    dim3 g((unsigned)cdiv(bp.n2, TPB), nshift);
    if (bp.cplx_band) k_diag_weights<cplx><<<g, TPB, 0, st>>>(bp.map3d, bp.jstr, bp.mdiag, bp.n2, bp.pw, bp.ldab, bp.nmodes, (const cplx*)bp.Kband, s_d, n3d, w);
    else k_diag_weights<double><<<g, TPB, 0, st>>>(bp.map3d, bp.jstr, bp.mdiag, bp.n2, bp.pw, bp.ldab, bp.nmodes, (const double*)bp.Kband, s_d, n3d, w);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// ---------------------------------------------------------- blocked LU ------
// Step t factors block column [t, t+nbe): panel kernel (one CTA per system)
//   A11 = L11 U11,  L21 = A21 U11^-1 (R x nbe),  U12 = L11^-1 A12 (nbe x R)
// then the update kernel (32x32 tiles x systems)  A22 -= L21 U12  (R x R).
// NB is a compile-time constant so the per-thread row / column solves live in
// registers; a short last block is padded with an identity.
static const int NB = 16;
static const int NBMAX = NB;

// grid = (chunks, systems), 128 threads: every CTA factors the NB x NB diagonal block
// itself (one warp, ~1 us, redundant across chunks) and then solves 64 rows of A21 and
// 64 columns of A12.  A11 is NOT overwritten here (other chunks may still read it):
// chunk 0 parks the factored block in `scratch`, the update kernel copies it in place.
template <typename T>
__global__ void __launch_bounds__(128)
k_lu_panel(T* __restrict__ LU, size_t nel, long long n2, int p, int pw, int ld, long long t, T* __restrict__ scratch) {
    __shared__ T D[NB][NB + 1];                      // D[r][c]
    T* A = LU + (size_t)blockIdx.y * nel + pw;       // dense view base
    int nbe = (int)min((long long)NB, n2 - t);
    int R = (int)min((long long)p, n2 - (t + nbe));
    int tid = threadIdx.x;
    int idx = blockIdx.x * 64 + (tid & 63);          // row / column handled by this thread
    bool isrow = (tid < 64) && idx < R, iscol = (tid >= 64) && idx < R;
    // issue this thread's panel loads first: they overlap the A11 factorization
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
    for (int e = tid; e < NB * NB; e += blockDim.x) {
        int r = e % NB, c = e / NB;
        T val = (r == c) ? Sc<T>::one() : Sc<T>::zero();
        if (r < nbe && c < nbe) val = A[(t + r) + (t + c) * (long long)ld];
        D[r][c] = val;
    }
    __syncthreads();
    // unblocked LU of the NB x NB diagonal block by the first warp (lane r owns row r)
    if (tid < 32) {
        int r = tid & 15;
        T rowv[NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) rowv[c] = D[r][c];
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            T piv = Sc<T>::inv(shfl_row(rowv[c], c));       // pivot row c is final now
            T l = rowv[c] * piv;
            if (r > c) rowv[c] = l;
#pragma unroll
            for (int cc = c + 1; cc < NB; ++cc) {
                T u = shfl_row(rowv[cc], c);
                if (r > c) rowv[cc] -= l * u;
            }
        }
        if (tid < NB) {
#pragma unroll
            for (int c = 0; c < NB; ++c) D[r][c] = rowv[c];
        }
    }
    __syncthreads();
    if (blockIdx.x == 0) {
        T* S = scratch + (size_t)blockIdx.y * NB * NB;
        for (int e = tid; e < NB * NB; e += blockDim.x) S[e] = D[e % NB][e / NB];
    }
    if (isrow) {
        // row a of A21:  v U11 = row
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            T acc = v[c];
#pragma unroll
            for (int c2 = 0; c2 < c; ++c2) acc -= v[c2] * D[c2][c];
            v[c] = acc * Sc<T>::inv(D[c][c]);
        }
#pragma unroll
        for (int c = 0; c < NB; ++c) if (c < nbe) A[row + (t + c) * (long long)ld] = v[c];
    } else if (iscol) {
        // column b of A12:  L11 w = col (unit lower)
#pragma unroll
        for (int r = 0; r < NB; ++r) {
            T acc = v[r];
#pragma unroll
            for (int r2 = 0; r2 < r; ++r2) acc -= D[r][r2] * v[r2];
            v[r] = acc;
        }
        T* cp = A + col * (long long)ld + t;
#pragma unroll
        for (int r = 0; r < NB; ++r) if (r < nbe) cp[r] = v[r];
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_lu_update(T* __restrict__ LU, size_t nel, long long n2, int p, int pw, int ld, long long t, const T* __restrict__ scratch) {
    __shared__ T Ls[NB][32];
    __shared__ T Us[NB][33];
    T* A = LU + (size_t)blockIdx.z * nel + pw;
    int nbe = (int)min((long long)NB, n2 - t);
    int R = (int)min((long long)p, n2 - (t + nbe));
    int a0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
    int tid = threadIdx.x;
    if (blockIdx.x == 0 && blockIdx.y == 0) {
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
        acc00 += l0 * u0; acc01 += l0 * u1; acc10 += l1 * u0; acc11 += l1 * u1;
    }
    if (m0 && n0) c00[0] = old00 - acc00;
    if (m1 && n0) c00[16] = old10 - acc10;
    if (m0 && n1) c01[0] = old01 - acc01;
    if (m1 && n1) c01[16] = old11 - acc11;
}

// ---------------------------------------------------- two-level blocking -----
// The NB = 16 scheme above is a chain of 2 * n2 / 16 dependent launches whose length, not
// the flops, sets the factorisation time.  For panels that fit shared memory the outer block
// is NBB = 64 columns: k_panel64 factors a whole 64-column panel inside ONE CTA per 64-row
// chunk (four 16-column sub-steps with the panel resident in shared memory: no global
// round trips, no launches), k_update64 applies the rank-64 update.  Column chunks hold rows
// of A21 and the block A11; row chunks hold columns of A12 TRANSPOSED together with A11^T:
// running the same Doolittle code on the transposed view yields U' = D L11^T, so
// U12^T = (A12^T U'^-1) D -- both panel kinds share one code path.
static const int NBB = 64;
static const int GS = NBB + 1;      // shared-memory row pitch (elements)

template <typename T>
__global__ void __launch_bounds__(256)
k_panel64(T* __restrict__ LU, size_t nel, long long n2, int p, int pw, int ld, long long t0, T* __restrict__ scratch, int chunks) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* G = reinterpret_cast<T*>(smem_raw);            // [NBB][GS]  view of A11 (row chunk: transposed)
    T* Q = G + NBB * GS;                              // [NBB][GS]  64 panel rows x 64 panel columns
    T* A = LU + (size_t)blockIdx.y * nel + pw;        // dense view base
    const int tid = threadIdx.x;
    const bool rowpanel = (int)blockIdx.x >= chunks;
    const int ch = rowpanel ? blockIdx.x - chunks : blockIdx.x;
    const int nbb = (int)min((long long)NBB, n2 - t0);
    const int R = (int)min((long long)p, n2 - (t0 + nbb));
    const long long pbase = t0 + nbb + 64LL * ch;     // first global row (column) of this chunk
    const int nq = (int)max(0LL, min(64LL, (long long)R - 64LL * ch));   // valid panel rows in this chunk
    const T zero = Sc<T>::zero(), one = Sc<T>::one();
    // ---- load (band-masked: entries farther than pw from the diagonal are not stored = zero)
    for (int e = tid; e < NBB * NBB; e += 256) {
        int r = e % NBB, c = e / NBB;                  // consecutive threads -> consecutive global rows
        T g = (r == c) ? one : zero;
        if (r < nbb && c < nbb) g = (abs(r - c) <= pw) ? A[(t0 + r) + (t0 + c) * (long long)ld] : zero;
        if (!rowpanel) G[r * GS + c] = g; else G[c * GS + r] = g;
        T q = zero;
        if (r < nq && c < nbb) {
            long long gr = pbase + r, gc = t0 + c;
            if (gr - gc <= pw) q = rowpanel ? A[gc + gr * (long long)ld] : A[gr + gc * (long long)ld];
        }
        Q[r * GS + c] = q;
    }
    __syncthreads();
    for (int o = 0; o < NBB; o += NB) {
        // ---- a. unblocked LU of the 16 x 16 diagonal sub-block by warp 0 (lane r owns row r)
        if (tid < 32) {
            int r = tid & 15;
            T rowv[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) rowv[c] = G[(o + r) * GS + o + c];
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                T piv = Sc<T>::inv(shfl_row(rowv[c], c));
                T l = rowv[c] * piv;
                if (r > c) rowv[c] = l;
#pragma unroll
                for (int cc = c + 1; cc < NB; ++cc) {
                    T u = shfl_row(rowv[cc], c);
                    if (r > c) rowv[cc] -= l * u;
                }
            }
            if (tid < NB) {
#pragma unroll
                for (int c = 0; c < NB; ++c) G[(o + r) * GS + o + c] = rowv[c];
            }
        }
        __syncthreads();
        // ---- b. triangular solves against the sub-block
        //   threads   0.. 63 : panel rows      q U^-1
        //   threads  64..127 : rows of G below the sub-block   g U^-1   (L part)
        //   threads 128..191 : columns of G right of the sub-block   L^-1 g   (U part)
        {
            T v[NB];
            const int grp = tid >> 6, idx = tid & 63;
            if (grp == 0 || (grp == 1 && idx >= o + NB)) {
                T* rowp = (grp == 0 ? Q : G) + idx * GS + o;
#pragma unroll
                for (int c = 0; c < NB; ++c) v[c] = rowp[c];
#pragma unroll
                for (int c = 0; c < NB; ++c) {
                    T acc = v[c];
#pragma unroll
                    for (int c2 = 0; c2 < c; ++c2) acc -= v[c2] * G[(o + c2) * GS + o + c];
                    v[c] = acc * Sc<T>::inv(G[(o + c) * GS + o + c]);
                }
#pragma unroll
                for (int c = 0; c < NB; ++c) rowp[c] = v[c];
            } else if (grp == 2 && idx >= o + NB) {
#pragma unroll
                for (int r = 0; r < NB; ++r) v[r] = G[(o + r) * GS + idx];
#pragma unroll
                for (int r = 0; r < NB; ++r) {
                    T acc = v[r];
#pragma unroll
                    for (int r2 = 0; r2 < r; ++r2) acc -= G[(o + r) * GS + o + r2] * v[r2];
                    v[r] = acc;
                }
#pragma unroll
                for (int r = 0; r < NB; ++r) G[(o + r) * GS + idx] = v[r];
            }
        }
        __syncthreads();
        // ---- c. trailing update inside the panel: rows (64 panel rows + rows of G below) x later columns
        if (o + NB < NBB) {
            const int rr = tid & 127, cg = tid >> 7;     // row task, column parity
            const bool isq = rr < 64;
            const int gr = isq ? rr : (o + NB + (rr - 64));           // G row for rr >= 64
            if (isq || gr < NBB) {
                T* rowp = (isq ? Q : G) + (isq ? rr : gr) * GS;
                T l[NB];
#pragma unroll
                for (int k = 0; k < NB; ++k) l[k] = rowp[o + k];
                // four independent accumulator chains per thread (fp64 FMA latency, not rate, is the limit)
                for (int c = o + NB + cg; c < NBB; c += 8) {
                    T a0 = rowp[c], a1 = (c + 2 < NBB) ? rowp[c + 2] : zero, a2 = (c + 4 < NBB) ? rowp[c + 4] : zero, a3 = (c + 6 < NBB) ? rowp[c + 6] : zero;
#pragma unroll
                    for (int k = 0; k < NB; ++k) {
                        const T* gk = G + (o + k) * GS + c;
                        a0 -= l[k] * gk[0];
                        if (c + 2 < NBB) a1 -= l[k] * gk[2];
                        if (c + 4 < NBB) a2 -= l[k] * gk[4];
                        if (c + 6 < NBB) a3 -= l[k] * gk[6];
                    }
                    rowp[c] = a0;
                    if (c + 2 < NBB) rowp[c + 2] = a1;
                    if (c + 4 < NBB) rowp[c + 4] = a2;
                    if (c + 6 < NBB) rowp[c + 6] = a3;
                }
            }
        }
        __syncthreads();
    }
    // ---- store
    if (rowpanel) {
        // U12^T = (A12^T U'^-1) D, D = diag(U') ; transposed store
        for (int e = tid; e < NBB * NBB; e += 256) {
            int c = e % NBB, r = e / NBB;               // consecutive threads -> consecutive global rows (= panel col c)
            if (r < nq && c < nbb) {
                long long gr = pbase + r, gc = t0 + c;
                if (gr - gc <= pw) A[gc + gr * (long long)ld] = Q[r * GS + c] * G[c * GS + c];
            }
        }
    } else {
        for (int e = tid; e < NBB * NBB; e += 256) {
            int r = e % NBB, c = e / NBB;
            if (r < nq && c < nbb) {
                long long gr = pbase + r, gc = t0 + c;
                if (gr - gc <= pw) A[gr + gc * (long long)ld] = Q[r * GS + c];
            }
        }
        if (ch == 0) {
            // park the factored diagonal block (other CTAs may still be reading A11)
            T* S = scratch + (size_t)blockIdx.y * NBB * NBB;
            for (int e = tid; e < NBB * NBB; e += 256) S[e] = G[(e % NBB) * GS + e / NBB];
        }
    }
}

// A22 -= L21 (R x 64) U12 (64 x R) on 32 x 32 tiles; tile (0,0) also copies the parked diagonal block in place
template <typename T>
__global__ void __launch_bounds__(256)
k_update64(T* __restrict__ LU, size_t nel, long long n2, int p, int pw, int ld, long long t0, const T* __restrict__ scratch) {
    __shared__ T Ls[NB][32];
    __shared__ T Us[NB][33];
    T* A = LU + (size_t)blockIdx.z * nel + pw;
    const int nbb = (int)min((long long)NBB, n2 - t0);
    const int R = (int)min((long long)p, n2 - (t0 + nbb));
    const int a0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
    const int tid = threadIdx.x;
    if (blockIdx.x == 0 && blockIdx.y == 0) {
        const T* S = scratch + (size_t)blockIdx.z * NBB * NBB;
        for (int e = tid; e < NBB * NBB; e += 256) {
            int r = e % NBB, c = e / NBB;
            if (r < nbb && c < nbb && abs(r - c) <= pw) A[(t0 + r) + (t0 + c) * (long long)ld] = S[e];
        }
    }
    if (a0 >= R || b0 >= R) return;
    const long long r0 = t0 + nbb + a0, c0 = t0 + nbb + b0;
    const int ta = tid % 16, tb = tid / 16;
    const bool m0 = (a0 + ta < R), m1 = (a0 + ta + 16 < R), n0 = (b0 + tb < R), n1 = (b0 + tb + 16 < R);
    T* c00 = A + (r0 + ta) + (c0 + tb) * (long long)ld;
    T* c01 = A + (r0 + ta) + (c0 + tb + 16) * (long long)ld;
    T old00 = (m0 && n0) ? c00[0] : Sc<T>::zero(), old10 = (m1 && n0) ? c00[16] : Sc<T>::zero();
    T old01 = (m0 && n1) ? c01[0] : Sc<T>::zero(), old11 = (m1 && n1) ? c01[16] : Sc<T>::zero();
    T acc00 = Sc<T>::zero(), acc01 = Sc<T>::zero(), acc10 = Sc<T>::zero(), acc11 = Sc<T>::zero();
    for (int k0 = 0; k0 < nbb; k0 += NB) {
        for (int e = tid; e < NB * 32; e += 256) {
            int a = e % 32, c = e / 32;
            long long gr = r0 + a, gc = t0 + k0 + c;
            Ls[c][a] = (a0 + a < R && k0 + c < nbb && gr - gc <= pw) ? A[gr + gc * (long long)ld] : Sc<T>::zero();
        }
        for (int e = tid; e < NB * 32; e += 256) {
            int c = e % NB, bq = e / NB;
            long long gr = t0 + k0 + c, gc = c0 + bq;
            Us[c][bq] = (b0 + bq < R && k0 + c < nbb && gc - gr <= pw) ? A[gr + gc * (long long)ld] : Sc<T>::zero();
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            T l0 = Ls[c][ta], l1 = Ls[c][ta + 16];
            T u0 = Us[c][tb], u1 = Us[c][tb + 16];
            acc00 += l0 * u0; acc01 += l0 * u1; acc10 += l1 * u0; acc11 += l1 * u1;
        }
        __syncthreads();
    }
    if (m0 && n0) c00[0] = old00 - acc00;
    if (m1 && n0) c00[16] = old10 - acc10;
    if (m0 && n1) c01[0] = old01 - acc01;
    if (m1 && n1) c01[16] = old11 - acc11;
}

// After the factorisation every NB x NB diagonal block holds L11 \ U11.  The triangular solves
// with those blocks are the sequential heart of the substitution kernels, so they are replaced
// once, here (fully parallel: one warp per block), by the explicit inverses
//   strictly lower part <- inv(L11) (unit diagonal implicit),  upper part <- inv(U11),
// which turns every block step of k_band_solve into two small dependency-free mat-vecs.
template <typename T>
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
    T xl[NB], xu[NB];
    int j = lane & 15;
    // column j of inv(L), L unit lower
#pragma unroll
    for (int r = 0; r < NB; ++r) {
        T acc = (r == j) ? Sc<T>::one() : Sc<T>::zero();
#pragma unroll
        for (int k = 0; k < r; ++k) if (k >= j) acc -= D[r][k] * xl[k];
        xl[r] = (r >= j) ? acc : Sc<T>::zero();
    }
    // column j of inv(U), U upper with diagonal
#pragma unroll
    for (int r = NB - 1; r >= 0; --r) {
        T acc = (r == j) ? Sc<T>::one() : Sc<T>::zero();
#pragma unroll
        for (int k = r + 1; k < NB; ++k) if (k <= j) acc -= D[r][k] * xu[k];
        xu[r] = (r <= j) ? acc * Sc<T>::inv(D[r][r]) : Sc<T>::zero();
    }
    __syncwarp();
    if (lane < NB && j < nbe) {
#pragma unroll
        for (int r = 0; r < NB; ++r) {
            if (r < nbe) {
                if (r > j) A[(t + r) + (t + j) * (long long)ld] = xl[r];
                else A[(t + r) + (t + j) * (long long)ld] = xu[r];
            }
        }
    }
}

// The factorisation is ~2 * n2 / NB tiny dependent launches (9 784 for the C2 sweep).  Issued
// one by one they make the solve hostage to host-thread jitter, so the whole sequence is
// captured once into a CUDA graph per (buffer, shape) and replayed with a single launch.
struct FactorGraphKey {
    void* LU; void* scratch; long long n2; int p, nsys, cplx;
    bool operator==(const FactorGraphKey& o) const {
        return LU == o.LU && scratch == o.scratch && n2 == o.n2 && p == o.p && nsys == o.nsys && cplx == o.cplx;
    }
};
struct FactorGraph { FactorGraphKey key; cudaGraphExec_t exec; unsigned long long stamp; };
static std::vector<FactorGraph> g_graphs;
static unsigned long long g_stamp = 0;

// 16 = single-level blocking (default); 64 = two-level blocking (env CS_FACTOR_BLOCK=64).  The
// two-level path is bit-for-bit tested but measured slower on B200 for p = 109 (113 ms vs 92 ms
// for the C2 sweep): the 64-column panel kernel is a long chain of dependent fp64 operations in
// one CTA.  Kept selectable for wider bands / future tuning.
static int g_factor_block = -1;

template <typename T>
static int band_factor_enqueue(const BandPlan& bp, T* LU, int nsys, T* scratch, cudaStream_t st) {
    size_t nel = bp.band_elems();
    int ld = bp.ldab - 1;
    int tiles = (bp.p + 31) / 32;
    int chunks = (bp.p + 63) / 64;
    if (g_factor_block == 64) {
        size_t smb = 2 * (size_t)NBB * GS * sizeof(T);
        for (long long t = 0; t < bp.n2; t += NBB) {
            dim3 gp(2 * chunks, nsys);
            k_panel64<T><<<gp, 256, smb, st>>>(LU, nel, bp.n2, bp.p, bp.pw, ld, t, scratch, chunks);
            dim3 g(tiles, tiles, nsys);
            k_update64<T><<<g, 256, 0, st>>>(LU, nel, bp.n2, bp.p, bp.pw, ld, t, (const T*)scratch);
        }
        dim3 gi((unsigned)((bp.n2 + NB - 1) / NB), nsys);
        k_invert_diag<T><<<gi, 32, 0, st>>>(LU, nel, bp.n2, bp.pw, ld);
        return CS_OK;
    }
    for (long long t = 0; t < bp.n2; t += NB) {
        dim3 gp(chunks, nsys);
        k_lu_panel<T><<<gp, 128, 0, st>>>(LU, nel, bp.n2, bp.p, bp.pw, ld, t, scratch);
        dim3 g(tiles, tiles, nsys);
        k_lu_update<T><<<g, 256, 0, st>>>(LU, nel, bp.n2, bp.p, bp.pw, ld, t, (const T*)scratch);
    }
    dim3 gi((unsigned)((bp.n2 + NB - 1) / NB), nsys);
    k_invert_diag<T><<<gi, 32, 0, st>>>(LU, nel, bp.n2, bp.pw, ld);
    return CS_OK;
}

template <typename T>
static int band_factor_t(const BandPlan& bp, T* LU, int nsys, cudaStream_t st) {
    CS_CHECK(bp.nb == NB, "band plan block size must equal the compiled NB");
    if (g_factor_block < 0) {
        const char* e = getenv("CS_FACTOR_BLOCK");
        g_factor_block = (e && atoi(e) == 64) ? 64 : 16;
    }
    {
        static bool attr = false;                // per instantiation
        if (!attr) { CS_CUDA(cudaFuncSetAttribute(k_panel64<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * NBB * GS * sizeof(T)))); attr = true; }
    }
    static void* scratch = nullptr;
    static size_t scratch_bytes = 0;
    size_t need = (size_t)nsys * NBB * NBB * sizeof(T);
    if (need > scratch_bytes) {
        if (scratch) cudaFree(scratch);
        if (need < (1u << 20)) need = 1u << 20;
        CS_CUDA(cudaMalloc(&scratch, need));
        scratch_bytes = need;
    }
    FactorGraphKey key{(void*)LU, scratch, bp.n2, bp.p, nsys, (int)(sizeof(T) == sizeof(cplx)) + 2 * g_factor_block};
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
    static cudaStream_t aux[3] = {nullptr, nullptr, nullptr};
    static cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
    if (!ev_fork) {
        cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
        for (int g = 0; g < 3; ++g) { cudaStreamCreateWithFlags(&aux[g], cudaStreamNonBlocking); cudaEventCreateWithFlags(&ev_join[g], cudaEventDisableTiming); }
    }
    const int groups = (nsys >= 16) ? 4 : (nsys >= 6 ? 2 : 1);
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        bool okq = true;
        if (groups > 1) {
            okq = okq && cudaEventRecord(ev_fork, st) == cudaSuccess;
            for (int g = 1; g < groups; ++g) okq = okq && cudaStreamWaitEvent(aux[g - 1], ev_fork, 0) == cudaSuccess;
        }
        for (int g = 0; g < groups && okq; ++g) {
            int s0 = (int)((long long)nsys * g / groups), s1 = (int)((long long)nsys * (g + 1) / groups);
            cudaStream_t sg = (g == 0) ? st : aux[g - 1];
            okq = okq && band_factor_enqueue<T>(bp, LU + (size_t)s0 * bp.band_elems(), s1 - s0, (T*)scratch + (size_t)s0 * NBB * NBB, sg) == CS_OK;
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
        CS_TRY(band_factor_enqueue<T>(bp, LU, nsys, (T*)scratch, st));
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
int band_factor(const BandPlan& bp, void* LU, bool lu_cplx, int nsys_tot, cudaStream_t st) {
    if (lu_cplx) return band_factor_t<cplx>(bp, (cplx*)LU, nsys_tot, st);
    return band_factor_t<double>(bp, (double*)LU, nsys_tot, st);
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
template <typename TL, int TPR>
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

template <typename TL, typename TV, int TPR>
__global__ void __launch_bounds__(TPR == 1 ? 1024 : (TPR == 2 ? 768 : 512))
k_band_solve(const TL* __restrict__ LU, size_t nel, long long n2, int p, int pw, int ld,
                             int nmodes, const int* __restrict__ shift_of_vec, TV* __restrict__ xm, int wsz, int nst) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int v = blockIdx.y, m = blockIdx.x;
    int f = shift_of_vec ? shift_of_vec[v] : 0;
    const TL* A = LU + ((size_t)f * nmodes + m) * nel + pw;
    TV* b = xm + ((size_t)v * nmodes + m) * n2;
    const int ldd = NB + p;                                   // rows per staged block column
    const size_t ssz = (size_t)ldd * NB;                      // elements per stage
    TL* ring = reinterpret_cast<TL*>(smem_raw);               // [nst][NB][ldd]
    TV* win = reinterpret_cast<TV*>(ring + (size_t)nst * ssz);   // circular window, wsz = power of two >= p + 5 NB
    __shared__ TV yw[32][NB];                                 // per-warp copy of the block solution
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r16 = lane & 15, half = lane >> 4;              // block mat-vec: lane <-> (row, column parity)
    const int urow = tid / TPR, uq = tid % TPR;               // window update: thread <-> (row, column group)
    const int wm = wsz - 1;
    const long long nsteps = (n2 + NB - 1) / NB;
    const TV zero = Sc<TV>::zero();

    // ================= forward: L y = b (unit lower; diagonal blocks hold inv(L11)) =================
    auto fwd_prefetch = [&](int s) {
        if (s < nsteps) {
            long long t2 = (long long)s * NB;
            int nbe2 = (int)min((long long)NB, n2 - t2);
            int R2 = (int)min((long long)p, n2 - (t2 + nbe2));
            prefetch_cols<TL, TPR>(ring + (size_t)(s % nst) * ssz, ldd, A, ld, t2, nbe2 + R2, t2, nbe2);
        }
        __pipeline_commit();
    };
    for (int e = tid; e < 3 * NB + p && e < n2; e += blockDim.x) win[e & wm] = b[e];
    for (int q = 0; q < nst - 1; ++q) fwd_prefetch(q);
    {
        // entering rows are loaded one step before they are stored (the global-load latency
        // would otherwise stall warp 0 in front of every barrier)
        TV pend = zero;
        long long pe = 3 * NB + p + tid;                      // row loaded into pend (threads < NB)
        if (tid < NB && pe < n2) pend = b[pe];
        int slot = 0;
        const int ns = (int)nsteps;
        for (int s = 0; s < ns; ++s) {
            const long long t = (long long)s * NB;
            const int nbe = (int)min((long long)NB, n2 - t);
            const int R = (int)min((long long)p, n2 - (t + nbe));
            const TL* cur = ring + (size_t)slot * ssz;
            slot = (slot + 1 == nst) ? 0 : slot + 1;
            __pipeline_wait_prior(nst - 2);                   // stage s has landed
            __syncthreads();                                  // ... and is visible; window of step s-1 complete
            fwd_prefetch(s + nst - 1);                        // refills the slot step s-1 just released
            if (tid < NB) {
                if (pe < n2) win[pe & wm] = pend;             // rows t + 3 NB + p .. (first touched at step s + 2)
                pe += NB;
                if (pe < n2) pend = b[pe];
            }
            // every warp computes y1 = inv(L11) b1 for itself: lane (r, half) sums the columns of its parity
            TV yr = (half == 0) ? win[(t + r16) & wm] : zero;
#pragma unroll
            for (int i = 0; i < NB / 2; ++i) {
                int c = 2 * i + half;
                if (c < r16) yr += cur[r16 + c * ldd] * win[(t + c) & wm];
            }
            yr += shfl_xor_t<TV>(yr, 16);
            if (lane < NB) yw[warp][lane] = yr;
            __syncwarp();
            if (warp == 0 && lane < nbe) b[t + lane] = yr;
            // window update: row (nbe + urow) of the staged block, TPR threads per row
            TV acc = zero;
            if (urow < R) {
#pragma unroll
                for (int i = 0; i < NB / TPR; ++i) {
                    int c = uq + TPR * i;
                    if (c < nbe) acc += cur[(nbe + urow) + c * ldd] * yw[warp][c];
                }
            }
            if (TPR >= 2) acc += shfl_xor_t<TV>(acc, 1);
            if (TPR >= 4) acc += shfl_xor_t<TV>(acc, 2);
            if (uq == 0 && urow < R) win[(t + nbe + urow) & wm] -= acc;
        }
    }
    __pipeline_wait_prior(0);
    __syncthreads();

    // ================= backward: U x = y (diagonal blocks hold inv(U11)) =================
    const long long tlast = (nsteps - 1) * NB;
    auto bwd_prefetch = [&](int s) {
        if (s < nsteps) {
            long long t2 = tlast - (long long)s * NB;
            int nbe2 = (int)min((long long)NB, n2 - t2);
            int Ra2 = (int)min((long long)p, t2);
            prefetch_cols<TL, TPR>(ring + (size_t)(s % nst) * ssz, ldd, A, ld, t2 - Ra2, Ra2 + nbe2, t2, nbe2);
        }
        __pipeline_commit();
    };
    // window now slides downwards; (re)load the last p + 3 NB rows of y
    for (long long e = max(0LL, n2 - (4 * NB + p)) + tid; e < n2; e += blockDim.x) win[e & wm] = b[e];
    for (int q = 0; q < nst - 1; ++q) bwd_prefetch(q);
    {
        TV pend = zero;
        long long pe = n2 - (4 * NB + p) - NB + tid;          // next rows below the loaded window
        if (tid < NB && pe >= 0) pend = b[pe];
        int slot = 0;
        const int ns = (int)nsteps;
        for (int s = 0; s < ns; ++s) {
            const long long t = tlast - (long long)s * NB;
            const int nbe = (int)min((long long)NB, n2 - t);
            const int Ra = (int)min((long long)p, t);
            const TL* cur = ring + (size_t)slot * ssz;        // rows [t - Ra, t + nbe), local row = row - (t - Ra)
            slot = (slot + 1 == nst) ? 0 : slot + 1;
            __pipeline_wait_prior(nst - 2);
            __syncthreads();
            bwd_prefetch(s + nst - 1);
            if (tid < NB) {
                if (pe >= 0) win[pe & wm] = pend;
                pe -= NB;
                if (pe >= 0) pend = b[pe];
            }
            TV xr = zero;
#pragma unroll
            for (int i = 0; i < NB / 2; ++i) {
                int c = 2 * i + half;
                if (c >= r16 && c < nbe && r16 < nbe) xr += cur[(Ra + r16) + c * ldd] * win[(t + c) & wm];
            }
            xr += shfl_xor_t<TV>(xr, 16);
            if (lane < NB) yw[warp][lane] = xr;
            __syncwarp();
            if (warp == 0 && lane < nbe) b[t + lane] = xr;
            TV acc = zero;
            if (urow < Ra) {
#pragma unroll
                for (int i = 0; i < NB / TPR; ++i) {
                    int c = uq + TPR * i;
                    if (c < nbe) acc += cur[urow + c * ldd] * yw[warp][c];
                }
            }
            if (TPR >= 2) acc += shfl_xor_t<TV>(acc, 1);
            if (TPR >= 4) acc += shfl_xor_t<TV>(acc, 2);
            if (uq == 0 && urow < Ra) win[(t - Ra + urow) & wm] -= acc;
        }
    }
    __pipeline_wait_prior(0);
}

template <typename T, int TPR>
static int band_solve_launch(const BandPlan& bp, const T* LU, T* xm, int nvec, const int* sov, int th, size_t sm, int wsz, int nst, cudaStream_t st) {
    static bool attr = false;                    // one flag per instantiation
    if (!attr) { CS_CUDA(cudaFuncSetAttribute(k_band_solve<T, T, TPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 204 * 1024)); attr = true; }
    dim3 g(bp.nmodes, nvec);
    k_band_solve<T, T, TPR><<<g, th, sm, st>>>(LU, bp.band_elems(), bp.n2, bp.p, bp.pw, bp.ldab - 1, bp.nmodes, sov, xm, wsz, nst);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

template <typename T>
static int band_solve_t(const BandPlan& bp, const T* LU, T* xm, int nvec, const int* sov, cudaStream_t st) {
    int rowsn = bp.p + NB;
    int tpr = (4 * rowsn <= 512) ? 4 : (2 * rowsn <= 768 ? 2 : 1);
    int th = ((tpr * rowsn + 31) / 32) * 32;
    if (th < 64) th = 64;
    CS_CHECK(th <= 1024, "half bandwidth too large for the solve kernel (p + 16 <= 1024)");
    int wsz = 64;
    while (wsz < bp.p + 7 * NB) wsz <<= 1;
    size_t stage_b = (size_t)rowsn * NB * sizeof(T);
    int nst = (int)((200 * 1024 - (size_t)wsz * sizeof(T)) / stage_b);   // ring depth: as deep as shared memory allows
    if (nst > 6) nst = 6;
    CS_CHECK(nst >= 2, "half bandwidth too large for the shared-memory staged solve");
    size_t sm = (size_t)nst * stage_b + (size_t)wsz * sizeof(T);
    if (tpr == 4) return band_solve_launch<T, 4>(bp, LU, xm, nvec, sov, th, sm, wsz, nst, st);
    if (tpr == 2) return band_solve_launch<T, 2>(bp, LU, xm, nvec, sov, th, sm, wsz, nst, st);
    return band_solve_launch<T, 1>(bp, LU, xm, nvec, sov, th, sm, wsz, nst, st);
}

int band_solve(const BandPlan& bp, const void* LU, bool lu_cplx, void* xm, bool xm_cplx, int nvec, const int* shift_of_vec_d, cudaStream_t st) {
    CS_CHECK(bp.nb == NB, "band plan block size must equal the compiled NB");
    CS_CHECK(lu_cplx == xm_cplx, "LU and mode vectors must have the same scalar type");
    if (lu_cplx) return band_solve_t<cplx>(bp, (const cplx*)LU, (cplx*)xm, nvec, shift_of_vec_d, st);
    return band_solve_t<double>(bp, (const double*)LU, (double*)xm, nvec, shift_of_vec_d, st);
}

}  // namespace cs
