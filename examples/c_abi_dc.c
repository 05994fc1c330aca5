/* Plain-C client of the C ABI (include/casing_b200.h): a DC resistivity solve on a small cylindrically
 * symmetric casing mesh, no Python and no torch anywhere.  Device buffers come from cudaMalloc.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/c_abi_dc.c \
 *       -L casingsimulations_b200 -lcasing_b200 -L /usr/local/cuda/lib64 -lcudart -lm \
 *       -Wl,-rpath,$PWD/casingsimulations_b200 -o examples/c_abi_dc
 *
 * Prints the Krylov statistics and checks  ||q - A phi|| / ||q||  with the library's own operator apply.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <cuda_runtime.h>
#include "casing_b200.h"

#define CK(call) do { if ((call) != 0) { fprintf(stderr, "%s failed: %s\n", #call, cs_last_error()); return 1; } } while (0)
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 1; } } while (0)

int main(void) {
    enum { NX = 40, NZ = 60 };
    double hx[NX], hz[NZ], hth[1] = {2.0 * M_PI};
    for (int i = 0; i < NX; ++i) hx[i] = (i < 12) ? 0.01 : 0.01 * pow(1.35, i - 11);
    for (int k = 0; k < NZ; ++k) { int d = k < 15 ? 15 - k : (k >= 45 ? k - 44 : 0); hz[k] = 0.5 * pow(1.4, d); }
    double zsum = 0.0;
    for (int k = 0; k < NZ; ++k) zsum += hz[k];

    cs_ctx* ctx = NULL;
    CK(cs_ctx_create(0, NULL, &ctx));
    CK(cs_mesh_set(ctx, NX, 1, NZ, hx, hth, hz, -0.6 * zsum, 1.0, 0.5));
    long long cnt[7];
    CK(cs_mesh_counts(ctx, cnt));
    const long long nC = cnt[0];

    /* CasingInHalfspace: sigma_back 1e-2, air 1e-6 above z = 0, steel casing 0.02 < r < 0.04, -8 < z < 0 */
    double par[18] = {1e-2, 1.0, 1e-6, 0.0, 1.0, 5.5e6, 1.0, 0.02, 0.04, -8.0, 0.0, 1.0, 50.0, 1.0, 0.0, 0.0, 0.0, 0.0};
    double *sigma_d, *mu_d, *q_d, *phi_d, *r_d;
    CU(cudaMalloc((void**)&sigma_d, nC * sizeof(double)));
    CU(cudaMalloc((void**)&mu_d, nC * sizeof(double)));
    CU(cudaMalloc((void**)&q_d, nC * sizeof(double)));
    CU(cudaMalloc((void**)&phi_d, nC * sizeof(double)));
    CU(cudaMalloc((void**)&r_d, nC * sizeof(double)));
    CK(cs_model_paint(ctx, par, sigma_d, mu_d));

    /* +1 A in one cell, -1 A in another (Problem3D_CC.getRHS) */
    double* q = (double*)calloc((size_t)nC, sizeof(double));
    q[3 + NX * 30] = 1.0;
    q[30 + NX * 30] = -1.0;
    CU(cudaMemcpy(q_d, q, nC * sizeof(double), cudaMemcpyHostToDevice));

    double info[CS_INFO_LEN];
    CK(cs_solve_dc(ctx, q_d, phi_d, 1, 1e-10, 50, info));
    printf("cells %lld  iterations %.0f  scaled relres %.2e  kernels launched %.0f  factor %.2f ms  solve %.2f ms\n",
           nC, info[CS_INFO_ITERS], info[CS_INFO_RELRES], info[CS_INFO_N_LAUNCH], info[CS_INFO_MS_FACTOR], info[CS_INFO_MS_SOLVE]);

    /* residual through the operator object */
    cs_op* op = NULL;
    CK(cs_op_get(ctx, CS_KIND_DC, &op));
    double shift[2] = {0.0, 0.0};
    CK(cs_op_apply(op, shift, phi_d, r_d, 1, 0));
    CK(cs_ctx_sync(ctx));
    double* r = (double*)malloc((size_t)nC * sizeof(double));
    CU(cudaMemcpy(r, r_d, nC * sizeof(double), cudaMemcpyDeviceToHost));
    double num = 0.0, den = 0.0;
    for (long long i = 0; i < nC; ++i) { num += (r[i] - q[i]) * (r[i] - q[i]); den += q[i] * q[i]; }
    printf("||A phi - q|| / ||q|| = %.3e\n", sqrt(num / den));
    int ok = sqrt(num / den) < 1e-6 && info[CS_INFO_CONVERGED] > 0.5;

    free(q); free(r);
    cudaFree(sigma_d); cudaFree(mu_d); cudaFree(q_d); cudaFree(phi_d); cudaFree(r_d);
    CK(cs_ctx_destroy(ctx));
    puts(ok ? "C ABI OK" : "C ABI FAILED");
    return ok ? 0 : 2;
}
