// Internal host-side interfaces between the translation units of libcasing_b200.
#pragma once
#include <algorithm>
#include "cs_common.cuh"
#include <vector>

namespace cs {

// -------- cs_kernels.cu : assembly + stencil applies ---------------------------
// K1: paint sigma / mu from CasingInHalfspace-style parameters (model.py:190-260, 507-580)
int launch_paint(const MeshD& m, const double* zn_d, const double* par_d, double* sigma, double* mu, cudaStream_t st);
// K2: diagonal inner products (discretize getFaceInnerProduct / getEdgeInnerProduct)
int launch_face_ip(const MeshD& m, const double* prop, int invProp, int invMat, double* out, cudaStream_t st);
int launch_edge_ip(const MeshD& m, const double* prop, int invProp, int invMat, double* out, cudaStream_t st);
// elementwise geometric rescalings: mode 0: area^2/in, 1: in/area^2, 2: in/area, 3: in*area
int launch_face_scale(const MeshD& m, const double* in, double* out, int mode, cudaStream_t st);
// mode 0: len^2*in, 1: len^2/in, 2: in/len^2
int launch_edge_scale(const MeshD& m, const double* in, double* out, int mode, cudaStream_t st);

// K3: y = Dt diag(g) Dt^T x (+ shift0 * x[0] on row 0), Dt = vol*faceDiv; nsys vectors of length nC
int launch_dc_apply(const MeshD& m, const double* g, double shift0, const void* x, void* y, int nsys, bool cplx_vec, cudaStream_t st);
// Dt x_F -> cells ; Dt^T x_C -> faces
int launch_div(const MeshD& m, const void* xF, void* yC, bool cplx_vec, cudaStream_t st);
int launch_divT(const MeshD& m, const void* xC, void* yF, bool cplx_vec, cudaStream_t st);
// y_F = edgeCurl x_E ; y_E = edgeCurl^T x_F
int launch_curl(const MeshD& m, const void* xE, void* yF, bool cplx_vec, cudaStream_t st);
int launch_curlT(const MeshD& m, const void* xF, void* yE, bool cplx_vec, cudaStream_t st);
// 3-D nodal operators (discretize mesh.nodalGrad; nodes per z level: the axis node, then nx*nth nodes)
int launch_node_grad(const MeshD& m, const void* xN, void* yE, bool cplx_vec, bool with_len, cudaStream_t st);
// w != null: y_N = G_topo^T diag(w) G_topo x_N (+ shift0 x[0]);  w == null: y_N = G_topo^T x_E
int launch_node_apply(const MeshD& m, const double* w, double shift0, const void* x, void* y, int nsys, bool cplx_vec, cudaStream_t st);
// mesh.aveF2CCV / aveE2CCV: face / edge vector -> cell-centred vector, blocks [x | y | z] (view.py:300-308)
int launch_ave_ccv(const MeshD& m, bool from_faces, const void* x, void* y, bool cplx_vec, cudaStream_t st);
// K4 edge form: y = C^T diag(alpha) C x + s[sys] * beta .* x, wF = alpha/area^2
int launch_edge_op(const MeshD& m, const double* wF, const double* beta, const double* s_d /*2 doubles per sys*/,
                   const void* x, void* y, int nsys, bool cplx_vec, cudaStream_t st, int jc0 = 0, int jcn = 0 /*theta columns wanted; 0 = all*/);
// K4 face form: y = alpha .* (C diag(gamma) C^T (alpha .* x) + s[sys] x), uF = alpha/area, gE = len^2 gamma
int launch_face_op(const MeshD& m, const double* uF, const double* gE, const double* s_d,
                   const void* x, void* y, int nsys, bool cplx_vec, void* ework /*nsys*nE workspace or nullptr*/, cudaStream_t st);

// cs_dd.cu: r = b - A(s) x of the edge-form operator with exact products and double-double sums (iterative refinement);
// s_h: 2 doubles per system on the HOST, work: 2 * nF doubles
int launch_edge_resid_dd(const MeshD& m, const double* wF, const double* beta, const double* s_h, const void* x, const void* b,
                         void* r, int nsys, bool cplx_vec, double* work, cudaStream_t st);

int launch_csr_spmv(long long n, const long long* indptr, const int* indices, const void* vals, bool a_cplx,
                    const void* x, void* y, int nsys, bool x_cplx, cudaStream_t st);

// -------- cs_vec.cu : BLAS-1 style kernels ------------------------------------------
// out[sys] = sum conj?(a) * b over n entries; results to dots_d[2*sys..]
// optional weights w[wsel[sys]*n + t] (real, >= 0) give the diagonally scaled norms used by the convergence tests
int launch_dot(const void* a, const void* b, long long n, int nsys, bool cplx_vec, bool conj_a, double* dots_d, cudaStream_t st,
               const double* w = nullptr, const int* wsel = nullptr, int winv = 0, const double* mask = nullptr);
// y = a*x + b*y with per-system complex scalars (host arrays of 2 doubles per system)
int launch_axpby(void* y, const void* x, long long n, int nsys, bool cplx_vec, const double* a_d, const double* b_d, cudaStream_t st);
// z = x .* w (w real, shared by all systems), optionally reciprocal
// mode: 0 w, 1 1/w, 2 sqrt(w), 3 1/sqrt(w); wsel (optional, device) picks the weight row per system
int launch_mul_real(void* z, const void* x, const double* w, long long n, int nsys, bool cplx_vec, int mode, cudaStream_t st,
                    const int* wsel = nullptr);
int launch_real_to_cplx(const double* x, cplx* z, long long n, cudaStream_t st);

// -------- cs_band.cu : theta-DFT + banded direct solves (preconditioner) -----------
struct BandPlan {
    // per-mode 2-D unknown ordering ("band ordering")
    long long n2 = 0;          // unknowns per mode system (incl. dummies)
    int p = 0;                 // half bandwidth of the true couplings
    int nb = 16;               // block size of the blocked LU
    int pw = 0;                // stored half bandwidth = p + max(nb, nbf)
    // block size of the FACTORISATION steps (step kernels only).  The scalar LU without pivoting does not depend on the
    // blocking, so wide bands are factored in 32-column steps (half as many dependent launch pairs) while the substitution
    // keeps its 16-column blocks; k_invert_diag prepares the diagonal blocks for the latter
    int nbf = 16;
    int ldab = 0;              // 2*pw + 1
    int nth = 1;               // azimuthal cells of the mesh
    // stored mode systems per shift.  With T_m = diag(1 | i exp(+-i pi m / nth)) on the theta-staggered component
    // K''_m = T_m^* K_m T_m is REAL symmetric and K''_{nth-m} = K''_m, so only modes 0 .. nth/2 are probed,
    // stored and factored (nth/2 + 1 systems; 1 for a cylindrically symmetric mesh)
    int nmodes = 1;
    bool cplx_band = false;    // storage scalar type of K band (complex only for assembled complex CSR matrices)
    long long* map3d = nullptr;  // [n2] index into the 3-D vector at j = 0, -1 = dummy
    int* jstr = nullptr;         // [n2] stride per theta step (0 = axis unknown)
    signed char* hs = nullptr;   // [n2] theta stagger of the unknown in half cells: 0, +1 (Ey) or -1 (Fy)
    double* mdiag = nullptr;     // [n2] diagonal "mass" part multiplied by the shift s
    void* Kband = nullptr;       // [nmodes][n2*ldab] band of the s-independent part
    // format of the factored diagonal blocks, set by band_factor: 0 = inv(L11) \ inv(U11) of the
    // scalar LU, 1 = full inverse of the block pivot (block LDU: L21 = A21 inv(D), U12 = A12)
    mutable int diag_full = 0;
    // one unknown per (r, z) node (cell potentials, azimuthal edges): the NB x NB pivot blocks are
    // well conditioned and may be inverted explicitly; curl-curl systems with several components
    // per node have nearly singular pivot blocks and need the scalar LU
    bool scalar_field = false;
    // theta column the band matrices are probed at: 0 for uniform azimuthal widths, otherwise the column whose
    // width is closest to the mean (the circulant preconditioner then models an "average" column)
    long long Lrow = 0;        // band rows per z level (n2 = Lrow * levels + an optional padding dummy)
    // real factors can be staged with TMA bulk copies in the substitution kernel: every column segment starts at an even
    // element and has an even length (p even, n2 even -- build_plan pads both on 3-D meshes)
    bool tma_real = false;
    int jprobe = 0;
    // the probing response is confined to theta columns jprobe-1 .. jprobe+1 (three-term DFT).  Not for face unknowns:
    // the Fy faces of the innermost cells share the single axis Ez edge, so C C^T couples them across ALL theta columns
    bool thin_probe = true;
    // the probed band matrices are symmetric (measured after probing, band_check_symmetric): the step LU then works on the
    // lower triangle of the trailing window only and mirrors the pivot rows, U12 = diag(U11) L21^T
    bool sym = false;
    size_t band_elems() const { return (size_t)n2 * ldab; }
    // ---- two-front ("twisted") elimination: every mode system is split at one z level (the separator, `tw_w` unknowns);
    // the levels below it are eliminated top-down, the levels above it bottom-up (the same kernels on the REVERSED
    // ordering), at the same time: half as many dependent block steps in the factorisation and in both substitution
    // sweeps, no extra flops.  Storage: two band matrices of order tw_n2h = tw_M + tw_w per system,
    //   T = [tw_padT identity dummies | unknowns 0 .. tw_nt - 1 | separator],      T index i <-> original i - tw_padT
    //   B = [tw_padB identity dummies | unknowns n2 - 1 .. tw_nt + tw_w, descending | separator descending], i <-> n2 - 1 - (i - tw_padB)
    // tw_M is a multiple of the block size, so both fronts stop at a block boundary; the separator block of T then
    // receives B's Schur complement and is factored last.
    bool twist = false;
    // set by band_form_and_factor: the factors in the LU buffer ARE twisted (a factorisation that broke down is redone
    // one-front, see cs_op_factor); read by band_solve
    mutable bool tw_active = false;
    long long tw_n2h = 0, tw_M = 0, tw_nt = 0, tw_w = 0, tw_padT = 0, tw_padB = 0;
    size_t band_elems_h() const { return (size_t)tw_n2h * ldab; }
    size_t lu_elems_per_system() const { return twist ? std::max(2 * band_elems_h(), band_elems()) : band_elems(); }   // (room for either format)
    // mode vectors per 3-D vector: complex factors take all nth modes (mode m uses factor min(m, nth - m));
    // real factors (real shift) take Re / Im of modes 0 .. nth/2 of a REAL 3-D vector as 2 * nmodes real vectors
    int nmv(bool xm_cplx) const { return fmap ? nmv_own : (nth == 1 ? 1 : (xm_cplx ? nth : 2 * nmodes)); }
    int fold(bool xm_cplx) const { return fmap ? 3 : (nth == 1 ? 0 : (xm_cplx ? 1 : 2)); }
    // mode-sharded factors (z-slab / multi-GPU solve of one system): this rank stores `nmodes` of the factors and solves the
    // `nmv_own` mode vectors that use them; fmap[mm] = index of the stored factor of local mode vector mm (device array)
    const int* fmap = nullptr;
    int nmv_own = 0;
};

// block size (16 or 32) for a plan with half bandwidth p
int band_choose_nb(int p);
int band_choose_nbf(int p, int nb);
void band_plan_twist(BandPlan& bp);       // decides BandPlan::twist and fills the tw_* geometry (after n2, p, nb, Lrow are final)
void band_release_stream(cudaStream_t st);   // frees the band scratch buffers keyed by this stream
int band_factor_launches(const BandPlan& bp, bool lu_cplx);
// theta-DFT into the rotated mode basis x''_m = T_m^* x_m; xm [nvec][nmv][n2]
int band_dft_gather(const BandPlan& bp, const void* x3d, bool x_cplx, void* xm, bool xm_cplx, long long n3d, int nvec, cudaStream_t st);
int band_dft_scatter(const BandPlan& bp, const void* xm, bool xm_cplx, void* x3d, bool x_cplx, long long n3d, int nvec, cudaStream_t st);
int band_probe_fill(const BandPlan& bp, int color, int nb, void* x3d /*[nb][n3d]*/, long long n3d, bool x_cplx, cudaStream_t st, int mode = 0);
// modes 0 .. nmodes-1 of the probing responses y3d [nb][n3d] (nonzero in theta columns jprobe-1 .. jprobe+1 only)
// -> ym complex [nb][nmodes][n2], then column `color + b` of every mode matrix
int band_probe_gather(const BandPlan& bp, const void* y3d, bool y_cplx, void* ym, long long n3d, int nb, cudaStream_t st);
int band_probe_store_thin(const BandPlan& bp, int color, int nb, const void* y3d, bool y_cplx, long long n3d, cudaStream_t st);
int band_probe_store(const BandPlan& bp, int color, int nb, const void* ym, const void* y3d, bool y_cplx, long long n3d, cudaStream_t st);
// w[f][3-D index] = 1 / |diag(A(s_f))| recovered from the mode matrices (mean over modes of K_m[q,q]) + s_f * mass
int band_check_symmetric(const BandPlan& bp, int stride, double* asym_rel /*host*/, cudaStream_t st);
int band_diag_weights(const BandPlan& bp, const double* s_d, int nshift, long long n3d, double* w, cudaStream_t st);
// LU[f][mode] = Kband[mode] + s[f] * diag(mdiag)   (f = 0..nshift-1)
int band_form_shifted(const BandPlan& bp, const double* s_d /*device, 2 per shift*/, int nshift, void* LU, bool lu_cplx, cudaStream_t st);
int band_form_shifted_mode(const BandPlan& bp, int mode, const double* s_d, void* LU, cudaStream_t st);
// twisted halves (BandPlan::twist) of nshift x bp.nmodes systems; mode_base = true theta mode of stored mode 0
int band_form_twisted(const BandPlan& bp, const double* s_d, int nshift, void* LU, bool lu_cplx, cudaStream_t st, int mode_base);
// in-place blocked LU without pivoting of nsys_tot = nshift*nmodes band matrices
int band_factor(const BandPlan& bp, void* LU, bool lu_cplx, int nsys_tot, cudaStream_t st);
int band_form_and_factor(const BandPlan& bp, const double* s_d /*device*/, int nshift, void* LU, bool lu_cplx, cudaStream_t st, bool allow_twist = true);
// all pivots of the factors finite? (a twisted factorisation of a nearly singular system can break down where the one-front one does not)
int band_factor_finite(const BandPlan& bp, const void* LU, bool lu_cplx, int nshift, bool* ok /*host*/, cudaStream_t st);
// solve: rhs/solution xm [nvec][nmv][n2]; vector v uses the factors of shift shift_of_vec[v]
int band_solve(const BandPlan& bp, const void* LU, bool lu_cplx, void* xm, bool xm_cplx, int nvec, const int* shift_of_vec_d, cudaStream_t st);

}  // namespace cs
