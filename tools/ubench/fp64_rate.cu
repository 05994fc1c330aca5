// Micro-benchmark: sustained DFMA issue rate and dependent-chain latency per SM on this GPU.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_rate fp64_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k_dfma(double* out, int iters, long long* cyc) {
    double a[ILP];
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3 + i;
    double b = 1.0000001, c = 1e-9;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b, c);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
void run(int threads, int iters) {
    double* out; long long* cyc;
    int nb = 148;
    cudaMalloc(&out, sizeof(double) * nb * threads);
    cudaMalloc(&cyc, sizeof(long long) * nb);
    k_dfma<ILP><<<nb, threads>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_dfma<ILP><<<nb, threads>>>(out, iters, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double fl = 2.0 * (double)nb * threads * iters * ILP;
    printf("threads/SM %4d ILP %2d: %.1f DFMA/clk/SM (cycles %lld)  %.2f TFLOP/s  cycles per dependent DFMA step %.1f\n", threads, ILP,
           (double)threads * iters * ILP / (double)h[0], h[0], fl / ms * 1e-9, (double)h[0] / iters);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1>(32, 20000);     // latency of a dependent DFMA chain
    run<8>(32, 20000);
    run<8>(128, 20000);    // one warp per SMSP
    run<8>(256, 20000);
    run<8>(512, 20000);
    run<8>(1024, 20000);
    run<16>(1024, 20000);
    return 0;
}
