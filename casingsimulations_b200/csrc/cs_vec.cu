// K7: Krylov vector phase -- batched dot products (deterministic two-stage
// reduction with warp shuffles) and fused axpby updates.  nsys independent
// systems share every launch (blockIdx.y = system).
#include "cs_internal.h"

namespace cs {

static const int TPB = 256;
static const int DOT_BLOCKS = 256;   // partial sums per system

template <typename T, bool CONJ>
__global__ void k_dot_partial(const T* __restrict__ a, const T* __restrict__ b, long long n, double* __restrict__ part,
                              const double* __restrict__ wt, const int* __restrict__ wsel, int winv, const double* __restrict__ mask) {
    int sys = blockIdx.y;
    a += sys * n;
    b += sys * n;
    if (wt) wt += (wsel ? wsel[sys] : 0) * n;
    double re = 0.0, im = 0.0;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        T av = a[t], bv = b[t];
        if (CONJ) av = Sc<T>::conj(av);
        T pr = av * bv;
        double ww = wt ? (winv ? 1.0 / wt[t] : wt[t]) : 1.0;
        if (mask) ww *= mask[t];                               // z-slab decomposition: owned entries only
        re += ww * Sc<T>::re(pr);
        im += ww * Sc<T>::im(pr);
    }
    __shared__ double sre[TPB / 32], sim[TPB / 32];
    for (int o = 16; o > 0; o >>= 1) {
        re += __shfl_down_sync(0xffffffffu, re, o);
        im += __shfl_down_sync(0xffffffffu, im, o);
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sre[w] = re; sim[w] = im; }
    __syncthreads();
    if (w == 0) {
        re = (l < TPB / 32) ? sre[l] : 0.0;
        im = (l < TPB / 32) ? sim[l] : 0.0;
        for (int o = 4; o > 0; o >>= 1) {
            re += __shfl_down_sync(0xffffffffu, re, o);
            im += __shfl_down_sync(0xffffffffu, im, o);
        }
        if (l == 0) {
            part[2 * (sys * (long long)gridDim.x + blockIdx.x)] = re;
            part[2 * (sys * (long long)gridDim.x + blockIdx.x) + 1] = im;
        }
    }
}

__global__ void k_dot_final(const double* __restrict__ part, int nblk, double* __restrict__ out) {
    int sys = blockIdx.x;
    double re = 0.0, im = 0.0;
    for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
        re += part[2 * (sys * (long long)nblk + b)];
        im += part[2 * (sys * (long long)nblk + b) + 1];
    }
    __shared__ double sre[32], sim[32];
    for (int o = 16; o > 0; o >>= 1) {
        re += __shfl_down_sync(0xffffffffu, re, o);
        im += __shfl_down_sync(0xffffffffu, im, o);
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sre[w] = re; sim[w] = im; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 0.0, i2 = 0.0;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) { r += sre[q]; i2 += sim[q]; }
        out[2 * sys] = r;
        out[2 * sys + 1] = i2;
    }
}

// scratch: at least 2*nsys*DOT_BLOCKS doubles, lives right after dots_d (caller provides 2*nsys*(1+DOT_BLOCKS))
int launch_dot(const void* a, const void* b, long long n, int nsys, bool cplx_vec, bool conj_a, double* dots_d, cudaStream_t st,
               const double* w, const int* wsel, int winv, const double* mask) {
    double* part = dots_d + 2 * (size_t)nsys;
    int nblk = (int)((n + TPB - 1) / TPB);
    if (nblk > DOT_BLOCKS) nblk = DOT_BLOCKS;
    if (nblk < 1) nblk = 1;
    dim3 g(nblk, nsys);
    if (cplx_vec) {
        if (conj_a) k_dot_partial<cplx, true><<<g, TPB, 0, st>>>((const cplx*)a, (const cplx*)b, n, part, w, wsel, winv, mask);
        else k_dot_partial<cplx, false><<<g, TPB, 0, st>>>((const cplx*)a, (const cplx*)b, n, part, w, wsel, winv, mask);
    } else {
        k_dot_partial<double, false><<<g, TPB, 0, st>>>((const double*)a, (const double*)b, n, part, w, wsel, winv, mask);
    }
    k_dot_final<<<nsys, 256, 0, st>>>(part, nblk, dots_d);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

template <typename T>
__global__ void k_axpby(T* __restrict__ y, const T* __restrict__ x, long long n, const double* __restrict__ a_d, const double* __restrict__ b_d) {
    int sys = blockIdx.y;
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n) return;
    T a = Sc<T>::from(a_d[2 * sys], a_d[2 * sys + 1]);
    T b = Sc<T>::from(b_d[2 * sys], b_d[2 * sys + 1]);
    long long o = sys * n + t;
    T yv = y[o];
    bool bz = (Sc<T>::re(b) == 0.0 && Sc<T>::im(b) == 0.0);
    y[o] = bz ? a * x[o] : a * x[o] + b * yv;
}

int launch_axpby(void* y, const void* x, long long n, int nsys, bool cplx_vec, const double* a_d, const double* b_d, cudaStream_t st) {
    dim3 g((unsigned)cdiv(n, TPB), nsys);
    if (cplx_vec) k_axpby<cplx><<<g, TPB, 0, st>>>((cplx*)y, (const cplx*)x, n, a_d, b_d);
    else k_axpby<double><<<g, TPB, 0, st>>>((double*)y, (const double*)x, n, a_d, b_d);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// mode: 0 w, 1 1/w, 2 sqrt(w), 3 1/sqrt(w); wsel (optional) picks the weight row per system
template <typename T>
__global__ void k_mul_real(T* __restrict__ z, const T* __restrict__ x, const double* __restrict__ w, long long n, int mode,
                           const int* __restrict__ wsel) {
    int sys = blockIdx.y;
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n) return;
    double ww = w[(wsel ? (long long)wsel[sys] * n : 0) + t];
    if (mode == 1) ww = 1.0 / ww;
    else if (mode == 2) ww = sqrt(ww);
    else if (mode == 3) ww = rsqrt(ww);
    z[sys * n + t] = ww * x[sys * n + t];
}
int launch_mul_real(void* z, const void* x, const double* w, long long n, int nsys, bool cplx_vec, int recip, cudaStream_t st,
                    const int* wsel) {
    dim3 g((unsigned)cdiv(n, TPB), nsys);
    if (cplx_vec) k_mul_real<cplx><<<g, TPB, 0, st>>>((cplx*)z, (const cplx*)x, w, n, recip, wsel);
    else k_mul_real<double><<<g, TPB, 0, st>>>((double*)z, (const double*)x, w, n, recip, wsel);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

__global__ void k_real_to_cplx(const double* __restrict__ x, cplx* __restrict__ z, long long n) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t < n) z[t] = mk(x[t], 0.0);
}
int launch_real_to_cplx(const double* x, cplx* z, long long n, cudaStream_t st) {
    k_real_to_cplx<<<cdiv(n, TPB), TPB, 0, st>>>(x, z, n);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

}  // namespace cs
