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

namespace cs {

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
static const int NBMAX = 32;

template <typename T>
__global__ void k_lu_panel(T* __restrict__ LU, size_t nel, long long n2, int p, int pw, int ld, int nb, long long t) {
    extern __shared__ unsigned char smem_raw[];
    T* D = reinterpret_cast<T*>(smem_raw);           // [nb][nb], D[r*nb + c]
    T* A = LU + (size_t)blockIdx.x * nel + pw;       // dense view base
    int nbe = (int)min((long long)nb, n2 - t);
    int R = (int)min((long long)p, n2 - (t + nbe));
    int tid = threadIdx.x;
    for (int e = tid; e < nbe * nbe; e += blockDim.x) {
        int r = e % nbe, c = e / nbe;
        D[r * nb + c] = A[(t + r) + (t + c) * (long long)ld];
    }
    __syncthreads();
    for (int c = 0; c < nbe; ++c) {
        T piv = Sc<T>::inv(D[c * nb + c]);
        __syncthreads();
        if (tid > c && tid < nbe) D[tid * nb + c] = D[tid * nb + c] * piv;
        __syncthreads();
        for (int e = tid; e < (nbe - c - 1) * (nbe - c - 1); e += blockDim.x) {
            int r = c + 1 + e % (nbe - c - 1), cc = c + 1 + e / (nbe - c - 1);
            D[r * nb + cc] -= D[r * nb + c] * D[c * nb + cc];
        }
        __syncthreads();
    }
    for (int e = tid; e < nbe * nbe; e += blockDim.x) {
        int r = e % nbe, c = e / nbe;
        A[(t + r) + (t + c) * (long long)ld] = D[r * nb + c];
    }
    T v[NBMAX];
    if (tid < R) {
        // row a of A21:  v U11 = row
        long long row = t + nbe + tid;
        for (int c = 0; c < nbe; ++c) v[c] = A[row + (t + c) * (long long)ld];
        for (int c = 0; c < nbe; ++c) {
            T acc = v[c];
            for (int c2 = 0; c2 < c; ++c2) acc -= v[c2] * D[c2 * nb + c];
            v[c] = acc * Sc<T>::inv(D[c * nb + c]);
        }
        for (int c = 0; c < nbe; ++c) A[row + (t + c) * (long long)ld] = v[c];
    } else if (tid >= R && tid < 2 * R) {
        // column b of A12:  L11 w = col (unit lower)
        long long col = t + nbe + (tid - R);
        T* cp = A + col * (long long)ld;
        for (int r = 0; r < nbe; ++r) v[r] = cp[t + r];
        for (int r = 0; r < nbe; ++r) {
            T acc = v[r];
            for (int r2 = 0; r2 < r; ++r2) acc -= D[r * nb + r2] * v[r2];
            v[r] = acc;
        }
        for (int r = 0; r < nbe; ++r) cp[t + r] = v[r];
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_lu_update(T* __restrict__ LU, size_t nel, long long n2, int p, int pw, int ld, int nb, long long t) {
    __shared__ T Ls[NBMAX][32];
    __shared__ T Us[NBMAX][33];
    T* A = LU + (size_t)blockIdx.z * nel + pw;
    int nbe = (int)min((long long)nb, n2 - t);
    int R = (int)min((long long)p, n2 - (t + nbe));
    int a0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
    if (a0 >= R || b0 >= R) return;
    int tid = threadIdx.x;
    long long r0 = t + nbe + a0, c0 = t + nbe + b0;
    for (int e = tid; e < nbe * 32; e += 256) {
        int a = e % 32, c = e / 32;
        Ls[c][a] = (a0 + a < R) ? A[(r0 + a) + (t + c) * (long long)ld] : Sc<T>::zero();
    }
    for (int e = tid; e < nbe * 32; e += 256) {
        int c = e % nbe, b = e / nbe;
        Us[c][b] = (b0 + b < R) ? A[(t + c) + (c0 + b) * (long long)ld] : Sc<T>::zero();
    }
    __syncthreads();
    // thread -> rows (ta, ta+16), cols (tb, tb+16)
    int ta = tid % 16, tb = tid / 16;
    T acc00 = Sc<T>::zero(), acc01 = Sc<T>::zero(), acc10 = Sc<T>::zero(), acc11 = Sc<T>::zero();
    for (int c = 0; c < nbe; ++c) {
        T l0 = Ls[c][ta], l1 = Ls[c][ta + 16];
        T u0 = Us[c][tb], u1 = Us[c][tb + 16];
        acc00 += l0 * u0; acc01 += l0 * u1; acc10 += l1 * u0; acc11 += l1 * u1;
    }
    int a, b;
    a = a0 + ta; b = b0 + tb;
    if (a < R && b < R) A[(r0 + ta) + (c0 + tb) * (long long)ld] -= acc00;
    b = b0 + tb + 16;
    if (a < R && b < R) A[(r0 + ta) + (c0 + tb + 16) * (long long)ld] -= acc01;
    a = a0 + ta + 16; b = b0 + tb;
    if (a < R && b < R) A[(r0 + ta + 16) + (c0 + tb) * (long long)ld] -= acc10;
    b = b0 + tb + 16;
    if (a < R && b < R) A[(r0 + ta + 16) + (c0 + tb + 16) * (long long)ld] -= acc11;
}

template <typename T>
static int band_factor_t(const BandPlan& bp, T* LU, int nsys, cudaStream_t st) {
    size_t nel = bp.band_elems();
    int ld = bp.ldab - 1;
    int pth = 2 * bp.p;
    if (pth < bp.nb * bp.nb) pth = bp.nb * bp.nb;
    pth = ((pth + 31) / 32) * 32;
    CS_CHECK(pth <= 1024, "half bandwidth too large for the panel kernel (2p <= 1024)");
    CS_CHECK(bp.nb <= NBMAX, "nb too large");
    size_t sm = (size_t)bp.nb * bp.nb * sizeof(T);
    int tiles = (bp.p + 31) / 32;
    for (long long t = 0; t < bp.n2; t += bp.nb) {
        k_lu_panel<T><<<nsys, pth, sm, st>>>(LU, nel, bp.n2, bp.p, bp.pw, ld, bp.nb, t);
        if (t + bp.nb < bp.n2) {
            dim3 g(tiles, tiles, nsys);
            k_lu_update<T><<<g, 256, 0, st>>>(LU, nel, bp.n2, bp.p, bp.pw, ld, bp.nb, t);
        }
    }
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}
int band_factor(const BandPlan& bp, void* LU, bool lu_cplx, int nsys_tot, cudaStream_t st) {
    if (lu_cplx) return band_factor_t<cplx>(bp, (cplx*)LU, nsys_tot, st);
    return band_factor_t<double>(bp, (double*)LU, nsys_tot, st);
}

// ---------------------------------------------------------- banded solve ----
// One CTA per (vector, mode): forward then backward substitution, both in the
// column-oriented (axpy) form so that every read of L / U is a contiguous
// column segment of the band.
template <typename T> __device__ __forceinline__ T shfl_t(T v, int src);
template <> __device__ __forceinline__ double shfl_t<double>(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <> __device__ __forceinline__ cplx shfl_t<cplx>(cplx v, int src) {
    return mk(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

template <typename TL, typename TV>
__global__ void k_band_solve(const TL* __restrict__ LU, size_t nel, long long n2, int p, int pw, int ld, int nb,
                             int nmodes, const int* __restrict__ shift_of_vec, TV* __restrict__ xm) {
    int v = blockIdx.y, m = blockIdx.x;
    int f = shift_of_vec ? shift_of_vec[v] : 0;
    const TL* A = LU + ((size_t)f * nmodes + m) * nel + pw;
    TV* b = xm + ((size_t)v * nmodes + m) * n2;
    __shared__ TV y1[NBMAX];
    int tid = threadIdx.x, lane = tid & 31;
    // ---- forward: L y = b (unit lower) ----
    for (long long t = 0; t < n2; t += nb) {
        int nbe = (int)min((long long)nb, n2 - t);
        int R = (int)min((long long)p, n2 - (t + nbe));
        if (tid < 32) {
            TV br = (lane < nbe) ? b[t + lane] : Sc<TV>::zero();
            for (int c = 0; c < nbe; ++c) {
                TV yc = shfl_t<TV>(br, c);
                if (lane > c && lane < nbe) br -= A[(t + lane) + (t + c) * (long long)ld] * yc;
            }
            if (lane < nbe) { y1[lane] = br; b[t + lane] = br; }
        }
        __syncthreads();
        if (tid < R) {
            long long row = t + nbe + tid;
            TV acc = b[row];
            for (int c = 0; c < nbe; ++c) acc -= A[row + (t + c) * (long long)ld] * y1[c];
            b[row] = acc;
        }
        __syncthreads();
    }
    // ---- backward: U x = y ----
    long long tlast = ((n2 - 1) / nb) * nb;
    for (long long t = tlast; t >= 0; t -= nb) {
        int nbe = (int)min((long long)nb, n2 - t);
        int Ra = (int)min((long long)p, t);
        if (tid < 32) {
            TV br = (lane < nbe) ? b[t + lane] : Sc<TV>::zero();
            for (int c = nbe - 1; c >= 0; --c) {
                TV xc = shfl_t<TV>(br, c);
                xc = Sc<TL>::inv(A[(t + c) + (t + c) * (long long)ld]) * xc;
                if (lane == c) br = xc;
                if (lane < c) br -= A[(t + lane) + (t + c) * (long long)ld] * xc;
            }
            if (lane < nbe) { y1[lane] = br; b[t + lane] = br; }
        }
        __syncthreads();
        if (tid < Ra) {
            long long row = t - Ra + tid;
            TV acc = b[row];
            for (int c = 0; c < nbe; ++c) acc -= A[row + (t + c) * (long long)ld] * y1[c];
            b[row] = acc;
        }
        __syncthreads();
    }
}

int band_solve(const BandPlan& bp, const void* LU, bool lu_cplx, void* xm, bool xm_cplx, int nvec, const int* shift_of_vec_d, cudaStream_t st) {
    int th = ((bp.p + 31) / 32) * 32;
    if (th < 32) th = 32;
    CS_CHECK(th <= 1024, "half bandwidth too large for the solve kernel");
    CS_CHECK(bp.nb <= 32, "nb must be <= 32 for the warp triangular solve");
    dim3 g(bp.nmodes, nvec);
    size_t nel = bp.band_elems();
    int ld = bp.ldab - 1;
    if (lu_cplx && xm_cplx) k_band_solve<cplx, cplx><<<g, th, 0, st>>>((const cplx*)LU, nel, bp.n2, bp.p, bp.pw, ld, bp.nb, bp.nmodes, shift_of_vec_d, (cplx*)xm);
    else if (!lu_cplx && !xm_cplx) k_band_solve<double, double><<<g, th, 0, st>>>((const double*)LU, nel, bp.n2, bp.p, bp.pw, ld, bp.nb, bp.nmodes, shift_of_vec_d, (double*)xm);
    else CS_CHECK(false, "LU and mode vectors must have the same scalar type");
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

}  // namespace cs
