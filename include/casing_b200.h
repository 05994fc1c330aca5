/* casing_b200 -- C ABI of the B200-native cylindrical-mesh EM solver.
 *
 * Drop-in boundary for the hot path casingSimulations hands to SimPEG + a sparse
 * direct solver.  Every entry point replaces a call that the reference makes
 * into un-vendored third-party code; the reference call site is cited.
 *
 * Conventions: every function returns 0 on success; on failure a message is
 * available from cs_last_error().  Pointers suffixed _h are host pointers,
 * _d device pointers (owned by the caller, e.g. torch tensors).  Vectors use
 * the reference layout (casingSimulations/utils.py:41-47, 63-67): Fortran order,
 * i = r fastest, [Fx|Fy|Fz] / [Ex|Ey|Ez]; complex128 = interleaved doubles.
 * A cs_ctx is bound to one device and one stream; it is not thread safe,
 * independent contexts are.
 */
#ifndef CASING_B200_H
#define CASING_B200_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct cs_ctx cs_ctx;
typedef struct cs_op cs_op;

/* operator kinds (cs_op_create) */
#define CS_KIND_DC 0          /* D~ MfRhoI D~^T            Problem3D_CC, run.py:339-344            */
#define CS_KIND_DC_GROUNDED 1 /* same + 1 at [0,0]          SimPEG tdem getAdc, via run.py:287-296  */
#define CS_KIND_EDGE_H 2      /* C^T MfRho C + s MeMu       Problem3D_h, run.py:240-248             */
#define CS_KIND_EDGE_E 3      /* C^T MfMui C + s MeSigma    Problem3D_e                             */
#define CS_KIND_FACE_J 4      /* MfRho (C MeMuI C^T MfRho + s)   Problem3D_j, run.py:287-296        */
#define CS_KIND_FACE_B 5      /* MfMui (C MeSigmaI C^T MfMui + s) Problem3D_b                       */
#define CS_KIND_NODAL_E 6     /* G^T MeSigma G + 1 at [0,0]      nodal DC initial state of tdem Problem3D_e (3-D meshes) */

/* info_h layout (doubles) returned by the solve calls */
#define CS_INFO_ITERS 0      /* Krylov iterations (max over systems)          */
#define CS_INFO_RELRES 1     /* max over systems of ||b - A x|| / ||b|| in the Jacobi-scaled norm (r_i / sqrt|a_ii|) */
#define CS_INFO_MS_ASSEMBLE 2
#define CS_INFO_MS_FACTOR 3
#define CS_INFO_MS_SOLVE 4
#define CS_INFO_N_APPLY 5    /* stencil applies (per system)                  */
#define CS_INFO_N_LAUNCH 6   /* kernels launched by this call                 */
#define CS_INFO_BYTES 7      /* algorithmic bytes moved by the stencil applies */
#define CS_INFO_CONVERGED 8
#define CS_INFO_BACKWARD_ERR 9 /* set when relres stalls above rtol at the fp64 floor: ||r~||/(8||x~||+||b~||) <= 2e-15 is accepted */
#define CS_INFO_DD_STEPS 10    /* refinement steps taken with the double-double residual (edge-form operators)  */
#define CS_INFO_DD_CORR 11     /* size of the last such correction, ||D^1/2 dx|| / ||D^1/2 x||                    */
#define CS_INFO_LEN 16

const char* cs_last_error(void);
long long cs_launch_count(void);

/* lifetime; stream may be NULL (a private stream is created) */
int cs_ctx_create(int device, void* stream, cs_ctx** out);
int cs_ctx_destroy(cs_ctx* ctx);
int cs_ctx_sync(cs_ctx* ctx);
/* optional forward-accuracy polish of the edge-form solves (FDEM-h / -e): up to max_dd_steps steps of iterative
 * refinement x += M^-1 (b - A x) with the residual evaluated in double-double arithmetic (exact products, csrc/cs_dd.cu).
 * Default 0 (off): the fp64 data of a 3-D casing problem do not determine the plain-L2 solution beyond 1e-5 at 1 Hz
 * anyway; energy-norm fields and casing currents are at 1e-10 without it. */
int cs_ctx_set_refine(cs_ctx* ctx, int max_dd_steps);

/* measurement aid (no reference counterpart): DFMA throughput of the context's device in TFLOP/s, the denominator
 * of the factorisation's FP64 roofline in bench.py */
int cs_fp64_peak(cs_ctx* ctx, double* tflops_h);

/* ---- one large system over several GPUs: z-slab decomposition with NCCL halo exchange (BASELINE config 5) ----
 * No reference counterpart (run.py:50-53: `num_threads` is its only, unused, parallel knob).  One process per GPU.
 * Rank 0 draws an id (cs_comm_unique_id, 128 bytes) and hands it to the others out of band (torch.distributed, MPI,
 * a file); every rank then calls cs_comm_init BEFORE cs_mesh_set with the global cell levels [z_lo, z_hi) it owns
 * (contiguous, ascending with the rank, >= 2 levels each, covering 0 .. nz).  From then on cs_mesh_set takes the GLOBAL
 * mesh and the context works on its slab: the cells [max(z_lo-1, 0), z_hi) -- cs_comm_local_range -- in the usual vector
 * layout (cs_mesh_counts returns the LOCAL counts).  Vectors handed in (s_e) must be valid on that whole local range;
 * solutions come back valid on it.  Implemented for cs_solve_fdem(form 'h'): stencil applies exchange one packed halo
 * message per neighbour, all inner products of a Krylov phase travel in one all-reduce, and the theta-mode factors are
 * sharded by mode (each rank factors the modes it owns over the full z range; mode vectors are transposed between the
 * slab and the mode layout with grouped ncclSend / ncclRecv). */
int cs_comm_unique_id(void* id128_h);
int cs_comm_init(cs_ctx* ctx, const void* nccl_unique_id_h, int rank, int nranks, int z_lo, int z_hi);
int cs_comm_local_range(cs_ctx* ctx, int* k0_k1_h);
/* refresh the ghost levels of nvec local vectors (family: 0 cells, 1 edges, 2 faces) from the neighbouring slabs */
int cs_comm_halo(cs_ctx* ctx, int family, void* vec_d, int nvec, int is_complex);

/* discretize.CylMesh([hx, hy, hz], x0)            mesh.py:87-99, 518, 563 */
int cs_mesh_set(cs_ctx* ctx, int nx, int nth, int nz, const double* hx_h, const double* hth_h,
                const double* hz_h, double z0, double axis_weight_edge, double axis_weight_face);
/* counts: nC nFx nFy nFz nEx nEy nEz */
int cs_mesh_counts(cs_ctx* ctx, long long* out7_h);

/* CasingInHalfspace.sigma / .mur                   model.py:190-260, 507-580, 699-717
 * params_h: 18 doubles, see cs_kernels.cu k_paint.  Writes the context's model
 * and, when non-NULL, the caller's sigma_d / mu_d (nC doubles each).          */
int cs_model_paint(cs_ctx* ctx, const double* params_h, double* sigma_d, double* mu_d);
/* PhysicalProperties.model = [sigma; mu]           model.py:848-868 */
int cs_model_set(cs_ctx* ctx, const double* sigma_d, const double* mu_d);

/* mesh.getFaceInnerProduct / getEdgeInnerProduct diagonals  (SimPEG MfRho, MeMu, ...) */
int cs_face_inner_product(cs_ctx* ctx, const double* prop_d, int invProp, int invMat, double* out_d);
int cs_edge_inner_product(cs_ctx* ctx, const double* prop_d, int invProp, int invMat, double* out_d);

/* mesh.edgeCurl, its transpose, vol*faceDiv and its transpose (is_complex: 0/1) */
int cs_curl(cs_ctx* ctx, const void* xE_d, void* yF_d, int is_complex);
int cs_curlT(cs_ctx* ctx, const void* xF_d, void* yE_d, int is_complex);
int cs_div(cs_ctx* ctx, const void* xF_d, void* yC_d, int is_complex);
int cs_divT(cs_ctx* ctx, const void* xC_d, void* yF_d, int is_complex);

/* mesh.nodalGrad (3-D meshes; nodes per z level: the axis node, then nx*nth nodes) and its transpose: x_N -> y_E, x_E -> y_N */
int cs_grad(cs_ctx* ctx, const void* xN_d, void* yE_d, int is_complex);
int cs_gradT(cs_ctx* ctx, const void* xE_d, void* yN_d, int is_complex);

/* mesh.aveF2CCV / mesh.aveE2CCV applied to a vector: the viewer's re-gridding to cell centres (view.py:300-308).
 * Output blocks [x | y | z] of nC entries (symmetric mesh: [x | z] from faces, [y] from edges). */
int cs_ave_f2ccv(cs_ctx* ctx, const void* xF_d, void* yCCV_d, int is_complex);
int cs_ave_e2ccv(cs_ctx* ctx, const void* xE_d, void* yCCV_d, int is_complex);

/* operator objects: weights (K2) + band plan + probed theta-mode band matrices.
 * Cached per kind inside the context; invalidated by cs_model_* / cs_mesh_set. */
int cs_op_get(cs_ctx* ctx, int kind, cs_op** out);
/* weights only, no band matrices: stencil-apply measurements at sizes whose factors do not fit */
int cs_op_get_noplan(cs_ctx* ctx, int kind, cs_op** out);
/* reps launches of the apply kernel bracketed by CUDA events on the context's stream */
int cs_op_apply_timed(cs_op* op, const double* shifts_h, const void* x_d, void* y_d, int nvec, int is_complex,
                      int reps, double* ms_per_launch_h);
/* Solver(A, mesh=...) for an ASSEMBLED scipy.sparse system (CSR, host pointers):
 * pymatsolver.Pardiso(A) / SimPEG.SolverLU(A)        run.py:17-24, 246, 294, 343
 * family: 0 cell-, 1 edge-, 2 face-based unknowns in the reference layout        */
int cs_op_from_csr(cs_ctx* ctx, int family, long long n, long long nnz, const long long* indptr_h,
                   const int* indices_h, const void* vals_h, int is_complex, cs_op** out);
long long cs_op_size(cs_op* op);
/* y = A(s) x for nvec vectors, shifts_h: 2 doubles (re, im) per vector.  K3 / K4. */
int cs_op_apply(cs_op* op, const double* shifts_h, const void* x_d, void* y_d, int nvec, int is_complex);
/* Solver(A)  : factor A(s_f) for nshift shifts      run.py:246, 294, 343 */
int cs_op_factor(cs_op* op, const double* shifts_h, int nshift, int complex_shift);
/* Ainv * rhs : preconditioned Krylov to rtol; vector v uses shift shift_of_vec_h[v] */
int cs_op_solve(cs_op* op, const void* rhs_d, void* x_d, int nvec, const int* shift_of_vec_h,
                int is_complex, double rtol, int maxit, double* info_h);
/* Ainv.clean() */
int cs_op_clean(cs_op* op);

/* Problem3D_CC.fields: phi = A^-1 q, q (nC x nrhs) column-major    run.py:205, 339-352 */
int cs_solve_dc(cs_ctx* ctx, const double* q_d, double* phi_d, int nrhs, double rtol, int maxit, double* info_h);
/* Problem3D_{h,j}.fields: one solve per frequency (form 'h' or 'j'); s_e_d real face
 * vector (sources.py:119-123); u_d complex, nfreq columns of nE (h) or nF (j)   run.py:205, 237-257 */
int cs_solve_fdem(cs_ctx* ctx, int form, const double* freqs_h, int nfreq, const double* s_e_d,
                  void* u_d, double rtol, int maxit, double* info_h);
/* Problem3D_j.fields (backward Euler, step-off RawVec_Grounded): out_d is
 * (nF x (nsteps+1)) column-major                                     run.py:205, 284-301 */
int cs_tdem_run(cs_ctx* ctx, int form, const double* dts_h, int nsteps, const double* s_e_d,
                double* out_d, double rtol, int maxit, double* info_h);

/* the same for nsrc sources at once (SimPEG's (n x nSrc) right-hand sides): they share every factorisation.
 * s_e_d: [nsrc][nF]; out_d: [nsteps+1][nsrc][nF]                                  sources.py:948-973 (SourceList) */
int cs_tdem_run_multi(cs_ctx* ctx, int form, const double* dts_h, int nsteps, const double* s_e_d, int nsrc,
                      double* out_d, double rtol, int maxit, double* info_h);

/* Fields derived quantities (tests/test_cyl2D3D.py:212,266; run.py:211)
 * which: 'j','e','h','b' ; kind: the formulation the solution u came from */
int cs_fields_fdem(cs_ctx* ctx, int form, int which, double freq, const void* u_d, const double* s_e_d, void* out_d);
int cs_fields_dc(cs_ctx* ctx, int which, const double* phi_d, double* out_d, int nrhs);

/* physics.casing_currents / casing_charges            physics.py:10-83
 * ix_d: nz doubles (per z cell), iz_d: nz+1 doubles (per z node); j real or complex */
int cs_casing_currents(cs_ctx* ctx, const void* j_d, int is_complex, double casing_a, double casing_b,
                       double casing_z0, double casing_z1, void* ix_d, void* iz_d);
int cs_casing_charges(cs_ctx* ctx, const double* charge_d, double casing_b, double casing_z0,
                      double casing_z1, double* out_d);

#ifdef __cplusplus
}
#endif
#endif
