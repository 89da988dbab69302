/*
 * mimsem_gpu.h -- C ABI of the B200-native horizontal mixed-mimetic operator path.
 *
 * This is the drop-in boundary for the hot path of davelee2804/MiMSEM that BASELINE.json names:
 * Basis -> ElMats/Geom -> Assembly/Topo -> MatMult.  Plain pointers and sizes only; every
 * function returns 0 on success and a negative code on failure (mimsem_last_error() gives the
 * text); nothing throws across this boundary.  The reference has no error convention of its own
 * (PETSc return codes are discarded everywhere).
 *
 * Two groups:
 *   mimsem_basis_* / mimsem_topo_* / mimsem_mesh_*   host-only (no GPU needed)
 *   mimsem_gpu_*                                     device engine (sm_100a CUDA)
 *
 * Device field layout ("column layout"): a k-level field over n degrees of freedom is the array
 *   f[row(dof) * ld + k],   k = 0..nlev-1,   ld >= nlev,
 * i.e. DOF-major with the vertical level fastest; row() is the identity for 0- and 2-forms and the
 * engine's element-blocked edge order for 1-forms (mimsem_gpu_form_permutation).  The reference keeps one PETSc Vec per level
 * (`Vec velx[NK]`, eul/UMJS14.cpp:302-316); mimsem_gpu_levels_to_columns / _columns_to_levels
 * convert on the device.  A single-level apply (the MatShell adaptor) uses ld = 1.
 *
 * Element-local conventions are the reference's (eul/ElMats.cpp:38-44, 73-79, 105-111, 135-141):
 *   x-normal edges j = iy*(p+1)+ix, y-normal edges j = iy*p+ix, faces j = iy*p+ix,
 *   nodes j = iy*(p+1)+ix, quadrature points q = qy*(m+1)+qx.
 */
#ifndef MIMSEM_GPU_H
#define MIMSEM_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIMSEM_OK 0
#define MIMSEM_ERR_ARG (-1)
#define MIMSEM_ERR_CUDA (-2)
#define MIMSEM_ERR_STATE (-3)
#define MIMSEM_ERR_IO (-4)
#define MIMSEM_ERR_UNSUPPORTED (-5)

#define MIMSEM_MESH_SPHERE 0
#define MIMSEM_MESH_BOX 1

/* last error text of the calling thread ("" if none) */
const char* mimsem_last_error(void);

/* ------------------------------------------------------------------ Basis (host) */
/* GaussLobatto(n): x[n+1], w[n+1].                     replaces eul/Basis.cpp:22-98   */
int mimsem_basis_gll(int n, double* x, double* w);
/* LagrangeNode(p,quad m).ljxi [(m+1)*(p+1)] and LagrangeEdge.ejxi [(m+1)*p],
 * quadrature point first.                               replaces eul/Basis.cpp:105-151, 238-286 */
int mimsem_basis_tables(int p, int m, double* ljxi, double* ejxi);
/* ElMats tabulations at the quadrature points, row-major (quad point, dof):
 * which = 0 M1x_j_xy_i (U), 1 M1y_j_xy_i (V), 2 M2_j_xy_i (W), 3 M0_j_xy_i (P), 4 Wii diagonal.
 *                                                       replaces eul/ElMats.cpp:20-186 */
int mimsem_elmat(int which, int p, int m, double* A);

/* ------------------------------------------------------------------ Topo (host) */
/* Sizes of one reference rank's ghosted maps: out = {n0,n1x,n1y,n2,n0l,n1xl,n1yl,n2l}.
 * kind = MIMSEM_MESH_*; order = element order; nprocs = 6*n^2 (sphere) or n^2 (box).
 *                                                       replaces scr/Proc2.py:52-70 */
int mimsem_topo_patch_sizes(int kind, int order, int ne, int nprocs, int rank, int out[8]);
/* The maps themselves (what input/{nodes,edges_x,edges_y,faces}_RRRR.txt hold).
 *                                                       replaces scr/Proc2.py:73-230, scr/ProcBox.py:59-136 */
int mimsem_topo_patch(int kind, int order, int ne, int nprocs, int rank, int* loc0, int* loc1x, int* loc1y, int* loc2);
/* Write the complete input/ directory the reference reads at start-up (scr/Setup.py:42-78). */
int mimsem_topo_write_input(int kind, int p, int m, int ne, int nprocs, const char* dir);

/* ------------------------------------------------------------------ global mesh (host) */
typedef struct mimsem_mesh mimsem_mesh;
/* Canonical global mesh (one patch per cube face / one patch for the box); m = quadrature order;
 * signed_det != 0 keeps the sign of det J as src/Geom.cpp:251 does (eul, box: fabs). */
int mimsem_mesh_create(int kind, int p, int m, int ne, int signed_det, mimsem_mesh** out);
void mimsem_mesh_destroy(mimsem_mesh* mesh);
/* out = {p, m, ne, nel, N0, N1, N2, NQ} */
int mimsem_mesh_sizes(const mimsem_mesh* mesh, int64_t out[8]);
/* element -> global DOF tables: el0[nel][(p+1)^2], el1x[nel][p(p+1)], el1y[nel][(p+1)p],
 * el2[nel][p^2], elq[nel][(m+1)^2]; any pointer may be NULL.   (Topo::elInds*_g, eul/Topo.cpp:253-305) */
int mimsem_mesh_tables(const mimsem_mesh* mesh, int* el0, int* el1x, int* el1y, int* el2, int* elq);
/* J[nel][(m+1)^2][4] (J00 J01 J10 J11), det[nel][(m+1)^2]       (Geom::J, Geom::det, eul/Geom.cpp:245-326) */
int mimsem_mesh_geometry(const mimsem_mesh* mesh, double* J, double* det);
/* cartesian coordinates of the quadrature points, xyz[NQ][3]   (scr/Geom2.py:10-277) */
int mimsem_mesh_coords(const mimsem_mesh* mesh, double* xyz);

/* ------------------------------------------------------------------ device engine */
typedef struct mimsem_gpu_ctx mimsem_gpu_ctx;

int mimsem_gpu_create(int device, mimsem_gpu_ctx** out);
int mimsem_gpu_destroy(mimsem_gpu_ctx* ctx);
/* Tuning / test knobs of a context (never read from the environment on a launch path; mimsem_gpu_create consults
 * MIMSEM_M1_VARIANT, MIMSEM_K_VARIANT, MIMSEM_ELL_VEC, MIMSEM_PREFETCH, MIMSEM_M1_MINB, MIMSEM_HOST_CHUNK once):
 *   "m1_variant" 4 automatic (default: the ring kernel for plain M1 on >= 4000 elements under "pdl_independent", else the
 *   tile kernel) | 3 persistent warp-specialised ring kernel | 2 tile kernel | 1 line tasks | 0 thread per element-level;  "k_variant" 1 tile | 0 registers;  "m2_variant", "inc_variant" 1 tile / element kernels | 0;
 *   "pdl_independent" 1: the caller guarantees that consecutive launches on a stream do not depend on each other (several
 *   fields per time step): the M1 tile kernels are launched with programmatic stream serialization, so a launch starts
 *   as the CTAs of the previous one retire instead of after its last CTA (default 0: ordinary stream order);
 *   "ell_vec" 4 | 2 | 1 levels per thread of the incidence kernels;  "prefetch_ahead" L2 prefetch distance in tiles;
 *   "m1_min_blocks" register-budget variant of the M1 tile kernel;  "host_chunk" levels per stage of apply_host;
 *   "halo_max_levels" levels per ghost row the caller's halo inboxes hold (0 = unchecked);
 *   "n0_owned" 0-form operators compute node rows [0, n0_owned) only (-1 = all; see mimsem_gpu_set_element_keys). */
int mimsem_gpu_set_option(mimsem_gpu_ctx* ctx, const char* name, long long value);

/* Basis tables (host pointers): quadrature weights w[m+1], ljxi[(m+1)(p+1)], ejxi[(m+1)p].
 * The sum-factorised kernels require m == p, where ljxi is the identity. */
int mimsem_gpu_set_basis(mimsem_gpu_ctx* ctx, int p, int m, const double* h_w, const double* h_ljxi, const double* h_ejxi);

/* Subdomain topology (host pointers, local 0-based indices).  Elements [0, nel_owned) are
 * computed; elements [nel_owned, nel_total) are read-only halo elements that only contribute to
 * shared DOFs of owned elements.  x- and y-normal edge tables index the same 1-form array.
 * mode: 0 = owner-computes (outputs on DOFs touched by owned elements' west/south/interior side
 *           are complete; DOFs owned elsewhere are not written),
 *       1 = partial sums (the reference's ghosted-local convention: east/north DOFs without a
 *           local owner receive this subdomain's partial sum; finish with a reverse ADD scatter,
 *           eul/Assembly.cpp:2194-2195). */
int mimsem_gpu_set_topo(mimsem_gpu_ctx* ctx, int nel_total, int nel_owned, int n0, int n1, int n2, int nq, int mode,
                        const int* h_el0, const int* h_el1x, const int* h_el1y, const int* h_el2, const int* h_elq);

/* Canonical order of the elements of the NEXT set_topo (one key per element, e.g. global element ids): sums over the
 * elements around a node (M0, M0h, M0h_up, E01) then run in key order instead of local element order, which makes the
 * results of a partitioned mesh bitwise equal to the single-GPU ones.  nel = 0 clears it.  Together with the option
 * "n0_owned" (0-form operators compute rows [0, n0_owned) only: the caller numbers its owned nodes first and holds
 * every element around them) this is what a node-partitioned subdomain needs. */
int mimsem_gpu_set_element_keys(mimsem_gpu_ctx* ctx, int nel, const int* h_keys);

/* Declare which rows are ghosts (refreshed from other subdomains): caller indices >= n1_owned (1-forms) and
 * >= n2_owned (2-forms).  Builds the INTERIOR / BOUNDARY element subsets; out_counts = {n_interior, n_boundary}. */
int mimsem_gpu_set_ghosts(mimsem_gpu_ctx* ctx, int n1_owned, int n2_owned, int out_counts[2]);

/* Geometry of the nel_total elements: J[nel][(m+1)^2][4], det[nel][(m+1)^2] (host pointers). */
int mimsem_gpu_set_geom(mimsem_gpu_ctx* ctx, const double* h_J, const double* h_det);

/* Layer thickness thick[nk][nq] (host pointer, level-major as Geom::thick, eul/Geom.cpp:752-763).
 * The engine stores 1/thick in column layout. nk = 0 clears it (2-D shallow water, src/). */
int mimsem_gpu_set_thickness(mimsem_gpu_ctx* ctx, int nk, const double* h_thick);

/* out = {nel_total, nel_owned, n0, n1, n2, nq, nk, p, m} */
int mimsem_gpu_sizes(const mimsem_gpu_ctx* ctx, int64_t out[9]);

/* Layout conversion (device pointers) between the reference's per-level vectors in the caller's DOF
 * numbering, levels[k*n + dof], and the engine's column layout columns[perm_space[dof]*ld + k].
 * space = k of the k-form space (0 nodes, 1 edges, 2 faces); -1 = plain transpose without renumbering.
 * 1-forms are stored element-blocked with x- and y-normal edges separated ("data layout in
 * HBM"): every apply below expects and produces fields in this internal order. */
int mimsem_gpu_levels_to_columns(mimsem_gpu_ctx* ctx, int space, int64_t n, int nlev, int ld, const double* d_levels, double* d_columns, void* stream);
int mimsem_gpu_columns_to_levels(mimsem_gpu_ctx* ctx, int space, int64_t n, int nlev, int ld, const double* d_columns, double* d_levels, void* stream);
/* perm[dof] = row of `dof` (caller numbering) in the engine's column layout, for callers that fill device
 * fields themselves (host output, n_space ints). */
int mimsem_gpu_form_permutation(const mimsem_gpu_ctx* ctx, int space, int* perm);

/*
 * Operator applications.  All field pointers are DEVICE pointers in column layout with leading
 * dimension ld; column j of a field corresponds to vertical level lev0 + j (thickness level);
 * nlev columns are processed in one launch.  tpow = number of 1/thick factors per quadrature
 * point (0, 1 or 2), which is how the reference's vert_scale / const_vert flags and the `src`
 * (no thickness) variant map onto one kernel:
 *   Umat::assemble(lev,scale,vert_scale)        -> apply_M1 (tpow = vert_scale)          eul/Assembly.cpp:51-153
 *   Wmat::assemble(lev,scale,vert_scale)        -> apply_M2 (tpow = vert_scale)          eul/Assembly.cpp:311-373
 *   Pmat::assemble(lev,scale)                   -> apply_M0 (tpow = 1)                   eul/Assembly.cpp:2004-2043
 *   Pmat::assemble_h(lev,scale,h2)              -> apply_M0h (tpow = 2)                  eul/Assembly.cpp:2045-2098
 *   Uhmat::assemble(h2,lev,const_vert,scale)    -> apply_M1h (tpow = 1 + const_vert)     eul/Assembly.cpp:416-474
 *   Whmat::assemble(rho,lev,scale,vs_rho)       -> apply_M2h (tpow = 1 + vs_rho)         eul/Assembly.cpp:1243-1299
 *   WtQUmat::assemble(u1,lev,scale)             -> apply_K  (tpow = 2)                   eul/Assembly.cpp:933-986
 *   Ut_mat::assemble(lev,scale)                 -> apply_M1 (tpow = 1, MIMSEM_THICK_MEAN) eul/Assembly.cpp:1338-1388
 *   Ut_mat::assemble_h(lev,scale,rho)           -> apply_M1h (tpow = 0)                  eul/Assembly.cpp:1390-1440
 *   WtQdUdz_mat::assemble(u1,scale)             -> apply_K  (tpow = 0, 2*scale)          eul/Assembly.cpp:1581-1640
 *   src/ variants (no thickness)                -> the same with tpow = 0                src/Assembly.cpp:30-124 ...
 * followed, in each case, by the MatMult the reference performs on the assembled matrix.
 * flags: MIMSEM_FIXED_LEVEL uses thickness level lev0 for every column (box/: Umat and Wmat are
 * assembled once at level 0, box/Assembly.cpp:44-45).
 */
#define MIMSEM_FIXED_LEVEL 1
/* Element subsets for overlapping the ghost refresh with computation (after mimsem_gpu_set_ghosts):
 * INTERIOR = owned elements that read no ghost row, BOUNDARY = the others.  Neither flag: all owned elements. */
#define MIMSEM_SUBSET_INTERIOR 2
#define MIMSEM_SUBSET_BOUNDARY 4
/* apply_M1 only: each thickness factor is the MEAN THICKNESS of levels lev and lev+1 instead of the inverse thickness
 * of level lev -- Ut_mat::assemble(lev, scale) is apply_M1 with tpow = 1 and this flag (eul/Assembly.cpp:1338-1388). */
#define MIMSEM_THICK_MEAN 8

int mimsem_gpu_apply_M1(mimsem_gpu_ctx* ctx, int lev0, int nlev, int ld, double scale, int tpow, int flags,
                        const double* d_x, double* d_y, void* stream);
int mimsem_gpu_apply_M1h(mimsem_gpu_ctx* ctx, int lev0, int nlev, int ld, double scale, int tpow, int flags,
                         const double* d_h2, const double* d_x, double* d_y, void* stream);
int mimsem_gpu_apply_M2(mimsem_gpu_ctx* ctx, int lev0, int nlev, int ld, double scale, int tpow, int flags,
                        const double* d_x, double* d_y, void* stream);
int mimsem_gpu_apply_M2h(mimsem_gpu_ctx* ctx, int lev0, int nlev, int ld, double scale, int tpow, int flags,
                         const double* d_h2, const double* d_x, double* d_y, void* stream);
int mimsem_gpu_apply_M0(mimsem_gpu_ctx* ctx, int lev0, int nlev, int ld, double scale, int tpow, int flags,
                        const double* d_x, double* d_y, void* stream);
int mimsem_gpu_apply_M0h(mimsem_gpu_ctx* ctx, int lev0, int nlev, int ld, double scale, int tpow, int flags,
                         const double* d_h2, const double* d_x, double* d_y, void* stream);
int mimsem_gpu_apply_K(mimsem_gpu_ctx* ctx, int lev0, int nlev, int ld, double scale, int tpow, int flags,
                       const double* d_u1, const double* d_x, double* d_y, void* stream);

/*
 * Rotational term and the potential-vorticity-upwinded operators of the shallow-water solver (BASELINE config 2):
 *   RotMat::assemble(q0)                  -> apply_R      (src tpow = 0; eul/box (q0,lev,scale): tpow = 2)
 *                                            src/Assembly.cpp:1346-1395, eul/Assembly.cpp:1030-1083, box/Assembly.cpp:826-880
 *   RotMat_up::assemble(q0, ul, fac, dt)  -> apply_R_up   (tau = fac*dt)                    src/Assembly.cpp:1784-1853
 *   Phmat::assemble_up(ul, hl, fac, dt)   -> apply_M0h_up (tau = fac*dt)                    src/Assembly.cpp:499-567
 * d_q0: 0-form coefficient (node rows), d_u1: 1-form advecting velocity (engine edge rows), d_h2: 2-form depth.
 * The trial basis is evaluated at the departure points xi_q - tau J^-1 u_g(xi_q) (LagrangeNode::eval_q).
 */
int mimsem_gpu_apply_R(mimsem_gpu_ctx* ctx, int lev0, int nlev, int ld, double scale, int tpow, int flags,
                       const double* d_q0, const double* d_x, double* d_y, void* stream);
int mimsem_gpu_apply_R_up(mimsem_gpu_ctx* ctx, int lev0, int nlev, int ld, double scale, int tpow, int flags,
                          const double* d_q0, const double* d_u1, double tau, const double* d_x, double* d_y, void* stream);
int mimsem_gpu_apply_M0h_up(mimsem_gpu_ctx* ctx, int lev0, int nlev, int ld, double scale, int tpow, int flags,
                            const double* d_h2, const double* d_u1, double tau, const double* d_x, double* d_y, void* stream);

/*
 * Mass-matrix solves -- the step that follows almost every apply in the reference:
 *   KSPSolve(ksp1, b, x) with GMRES + element-block Jacobi on M1  -> solve_M1   eul/HorizSolve.cpp:77-84, 224, 310, 322
 *   KSPSolve(ksp0, b, x) on M0                                    -> solve_M0   eul/HorizSolve.cpp:87-96, 246, 490
 * M1 is symmetric positive definite: solve_M1 runs a diagonally preconditioned conjugate-gradient iteration with the
 * matrix-free M1 kernel as the operator, batched over the nlev levels (each level is its own system with its own step
 * lengths; a converged level stops moving).  Stops when every level has |b - M1 x| <= rtol |b| or after maxit
 * iterations; *iters = iterations performed, *relres = the worst level's relative residual (either may be NULL).
 * The arguments before d_b are those of the apply that defines the matrix.  Whole mesh on one GPU only.
 * M0 is diagonal when the quadrature order equals the element order, so solve_M0 is a pointwise division.
 * diag_M1 returns the diagonal of M1 (MatGetDiagonal for a Jacobi-preconditioned KSP on the MatShell).
 */
int mimsem_gpu_solve_M1(mimsem_gpu_ctx* ctx, int lev0, int nlev, int ld, double scale, int tpow, int flags,
                        const double* d_b, double* d_x, double rtol, int maxit, int* iters, double* relres, void* stream);
int mimsem_gpu_solve_M0(mimsem_gpu_ctx* ctx, int lev0, int nlev, int ld, double scale, int tpow, int flags,
                        const double* d_b, double* d_x, void* stream);
int mimsem_gpu_diag_M1(mimsem_gpu_ctx* ctx, int lev0, int nlev, int ld, double scale, int tpow, int flags,
                       double* d_diag, void* stream);
/* z = blockdiag(M1)^-1 r with one block per owned element -- the preconditioner the reference selects for ksp1:
 * PCBJACOBI with PCBJacobiSetTotalBlocks(pc, size * nElsX^2, NULL) (eul/HorizSolve.cpp:77-84).  With the reference's
 * element-blocked edge numbering PETSc's equal consecutive row blocks are exactly the 2 p^2 edges each element owns.
 * Every block is tabulated from closed forms, factorised (L D L^T) and solved in shared memory, one thread per
 * (element, level); the arguments before d_r are those of the apply_M1 that defines the matrix.  Owner-computes mode. */
int mimsem_gpu_pc_bjacobi_M1(mimsem_gpu_ctx* ctx, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* d_r,
                             double* d_z, void* stream);

/* Umat_ray::assemble(lev, scale, dt, exner_k, exner_s) + MatMult (eul/Assembly.cpp:1846-1979; eul/Euler_2.cpp:1218-1229,
 * 1276-1277, 1437-1448): the Rayleigh-friction mass matrix, point weight dt k_v(exner(q), exner_s(q)) / thick with
 * k_v = compute_k_v of the reference.  d_exner: the Exner-pressure 2-form in column layout (columns = levels lev0 ..),
 * d_exner_s: its level-0 values, ONE per face (n2 doubles).  Columns are levels lev0 + j as for apply_M1. */
int mimsem_gpu_apply_M1ray(mimsem_gpu_ctx* ctx, int lev0, int nlev, int ld, double scale, double dt, const double* d_exner,
                           const double* d_exner_s, const double* d_x, double* d_y, void* stream);

/*
 * Remaining coefficient operators of the vorticity / forcing terms:
 *   Pvec::assemble(lev, scale)            -> diag_M0 (d_h2 = NULL, tpow = 1): M0 is diagonal when m == p, so the lumped
 *                                            0-form mass "vector" IS its diagonal            eul/Assembly.cpp:602-628
 *   Phvec::assemble(hl, lev, scale)       -> diag_M0 (d_h2 = hl, tpow = 2)                   eul/Assembly.cpp:652-681
 *   WmatInv::assemble(lev, scale)         -> solve_M2 (d_h2 = NULL, tpow = 1): the 2-form mass matrix is block diagonal,
 *   WhmatInv::assemble(rho, lev, scale)      one p^2 x p^2 SPD block per element; the reference inverts each block
 *                                            (Gauss-Jordan) and MatMults, here every block is tabulated, Cholesky-factorised
 *                                            and solved in shared memory                     eul/Assembly.cpp:1658-1800
 *   UtQWmat::assemble(u1, scale)          -> apply_UtQW: 2-form -> 1-form; equals Uhmat(h2 := x2) applied to u1 without
 *                                            thickness factors (and WtQdUdz_mat^T)           eul/Assembly.cpp:1462-1538
 */
int mimsem_gpu_diag_M0(mimsem_gpu_ctx* ctx, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* d_h2,
                       double* d_diag, void* stream);
int mimsem_gpu_solve_M2(mimsem_gpu_ctx* ctx, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* d_h2,
                        const double* d_b, double* d_x, void* stream);
int mimsem_gpu_apply_UtQW(mimsem_gpu_ctx* ctx, int nlev, int ld, double scale, const double* d_u1, const double* d_x2, double* d_y1,
                          void* stream);

/*
 * L2Vecs::HorizToVert / VertToHoriz (eul/L2Vecs.cpp:55-101) for device-resident 2-form fields: between the engine's
 * column layout cols[face*ld + k] and the reference's per-element vertical vectors, all owned elements back to back,
 * vert[e*(nlev*p^2) + k*p^2 + i]  (== vz[e] of size nk*p^2, face = Topo::elInds2_l(e)[i]).  A relabelling: bit exact.
 */
int mimsem_gpu_columns_to_vertical(mimsem_gpu_ctx* ctx, int nlev, int ld, const double* d_cols, double* d_vert, void* stream);
int mimsem_gpu_vertical_to_columns(mimsem_gpu_ctx* ctx, int nlev, int ld, const double* d_vert, double* d_cols, void* stream);

/* Incidence operators (exact +-1 stencils), E10mat/E21mat of eul/Assembly.cpp:1102-1226:
 * which = 0 E10 (0-form -> 1-form), 1 E01 = -E10^T, 2 E21 (1-form -> 2-form), 3 E12 = -E21^T. */
#define MIMSEM_E10 0
#define MIMSEM_E01 1
#define MIMSEM_E21 2
#define MIMSEM_E12 3
int mimsem_gpu_apply_incidence(mimsem_gpu_ctx* ctx, int which, int nlev, int ld, const double* d_x, double* d_y, void* stream);
/* The stencils themselves as CSR over local indices (host output), for bit-exact comparison with
 * the reference's matrices.  Call with NULL arrays to get sizes: out_sizes = {nrows, ncols, nnz}. */
int mimsem_gpu_incidence_csr(const mimsem_gpu_ctx* ctx, int which, int64_t out_sizes[3], int64_t* indptr, int* indices, double* values);

/*
 * End-to-end convenience with HOST buffers in the reference's per-level layout
 * (levels[k*n + dof]): copies in, converts, applies, converts back and copies out on the
 * engine's own streams.  op: 0 M1, 1 M2, 2 M0, 3 M1h, 4 K, 5 M2h, 6 M0h, 10+which incidence, 14 UtQW (h_coeff = u1, h_x the
 * 2-form), 15 / 16 diagonal of M0 / M0(h) (Pvec / Phvec; h_x is read but ignored), 17 / 18 M2^-1 / M2(rho)^-1 (WmatInv / WhmatInv),
 * 19 diagonal of M1 (MatGetDiagonal of the Umat shell), 20 element-block Jacobi of M1 (h_x = r, h_y = z).
 * h_coeff may be NULL for operators without a coefficient field.
 */
int mimsem_gpu_apply_host(mimsem_gpu_ctx* ctx, int op, int lev0, int nlev, double scale, int tpow, int flags,
                          const double* h_coeff, const double* h_x, double* h_y);
/* The same with the operators of BASELINE config 2: op 7 R (h_coeff = q0), 8 R_up (h_coeff = q0, h_u1, tau),
 * 9 M0h_up (h_coeff = h2, h_u1, tau); and op 21 Umat_ray (h_coeff = Exner 2-form of the levels, h_u1 = its level-0 values,
 * n2 doubles, tau = dt).  Every other op ignores h_u1 and tau. */
int mimsem_gpu_apply_host_up(mimsem_gpu_ctx* ctx, int op, int lev0, int nlev, double scale, int tpow, int flags,
                             const double* h_coeff, const double* h_u1, double tau, const double* h_x, double* h_y);

/* Halo pack / unpack (device pointers; d_rows holds engine rows, i.e. values of mimsem_gpu_form_permutation):
 *   gather : packed[i*nlev + k] = field[rows[i]*ld + k]        (send side of the ghost refresh)
 *   scatter: field[rows[i]*ld + k] = packed[i*nlev + k]        (receive side)
 * Together with an NCCL send/recv of the packed buffers these replace Topo's
 * VecScatter(gtol_1 | gtol_0, INSERT_VALUES, SCATTER_FORWARD)  (eul/Topo.cpp:145-155, eul/Euler_2.cpp:1455-1456). */
int mimsem_gpu_gather_rows(mimsem_gpu_ctx* ctx, int64_t nrows, int nlev, int ld, const int* d_rows, const double* d_field,
                           double* d_packed, void* stream);
int mimsem_gpu_scatter_rows(mimsem_gpu_ctx* ctx, int64_t nrows, int nlev, int ld, const int* d_rows, const double* d_packed,
                            double* d_field, void* stream);

/*
 * Peer-to-peer ghost refresh over NVLink, without NCCL on the data path (the kernel is fused with its collective:
 * the push kernel stores into the peer's memory).  Buffers that peers write into are allocated with
 * mimsem_gpu_ipc_alloc (cudaMalloc + IPC handle, 64 bytes, to be exchanged by the caller's process-group plumbing)
 * and mapped on the peers with mimsem_gpu_ipc_open.  d_peers is a device array of `npeers` records
 *   { const int* rows; int nrows; int row0; double* inbox; long long inbox_parity_stride;
 *     unsigned long long* signal; const unsigned long long* wait; }                        (48 bytes each)
 * Inbox rows are packed with stride nlev; a peer's share starts at row `row0` of the space's inbox.
 * push: rows = owned rows to send, inbox/signal = the PEER's inbox of the space and flag for this rank, wait = the ack
 *       the peer writes into this rank's memory;
 * pull: rows = ghost rows to fill, inbox/wait = this rank's inbox region and flag for the peer, signal = the ack
 *       on the peer.  d_epoch is a device counter, one for the pushes and one for the pulls of a space, advanced by
 *       every call, so a captured CUDA graph can be replayed and several fields can be in flight;
 * d_err is set to 1 if a peer never answered (the kernels give up after ~2 s instead of hanging the GPU).
 * nbuf = inbox copies of the space (2..4, inbox_parity_stride doubles apart): data epoch e lives in copy e % nbuf and a push
 * of epoch e waits for the acknowledgement of epoch e - nbuf -- the SAME rule as mimsem_gpu_apply_M1_halo, so that the
 * two mechanisms can be mixed on one space without a host synchronisation in between.  Whoever allocates the inboxes
 * declares their capacity with mimsem_gpu_set_option(ctx, "halo_max_levels", n); calls with nlev > n are rejected.
 */
int mimsem_gpu_ipc_alloc(mimsem_gpu_ctx* ctx, int64_t bytes, void** d_ptr, unsigned char handle[64]);
int mimsem_gpu_ipc_open(mimsem_gpu_ctx* ctx, const unsigned char handle[64], void** d_ptr);
int mimsem_gpu_ipc_close(mimsem_gpu_ctx* ctx, void* d_ptr, int owned);
int mimsem_gpu_halo_push(mimsem_gpu_ctx* ctx, int npeers, const void* d_peers, int nlev, int ld, int nbuf, const double* d_field,
                         void* d_epoch, int* d_err, void* stream);
int mimsem_gpu_halo_pull(mimsem_gpu_ctx* ctx, int npeers, const void* d_peers, int nlev, int ld, int nbuf, double* d_field,
                         void* d_epoch, int* d_err, void* stream);

/*
 * M1 apply FUSED with the ghost refresh of its input (one launch per step and GPU; replaces the
 * VecScatter(gtol_1, INSERT_VALUES, SCATTER_FORWARD) + Umat MatMult + VecScatter(ADD_VALUES, SCATTER_REVERSE)
 * sequence of eul/Euler_2.cpp:1455-1456, eul/Assembly.cpp:2194-2195).  The first `push_ctas` CTAs of the grid store
 * this rank's boundary rows into the peers' inboxes over NVLink and raise their flags; interior element tiles run
 * meanwhile; boundary tiles (ordered last) wait for the flags and stage their ghost rows straight from the inbox
 * by TMA; the last CTA acknowledges the inbox and advances the device-side epoch (CUDA-graph replayable).
 * Requires mimsem_gpu_set_ghosts, ld == nlev (even, <= 64) and inbox rows in ghost order:
 * inbox row i holds caller row n1_owned + i (d_push[].row0 = first row of each peer's share).
 * d_push / d_pull as for mimsem_gpu_halo_push / _pull; d_inbox = this rank's inbox, `nbuf` copies (2..4, the same on
 * every rank) parity_stride doubles apart, data epoch e in copy e % nbuf (3 copies let a pipelined push proceed
 * without waiting for the consumer of the current epoch); d_epoch = the space's two epoch counters {push, pull} (both advanced, so this
 * call can be mixed with push / pull pairs).  Ghost rows of d_x itself are neither read nor written.
 *
 * Software pipelining over INDEPENDENT applies (several fields per time step, a ring of right-hand sides): mode 1
 * pushes the boundary rows of d_x_push -- the input of the NEXT call -- while this call's boundary tiles consume
 * what the previous call pushed, which takes the NVLink round trip of the push off the critical path of a step.  A
 * pipelined sequence starts with one mode-2 call (push d_x_push only; nothing is applied, the epoch is unchanged)
 * and ends with one mode-3 call (consume only, push nothing); every rank must use the same mode in the same call.
 * mode 0 (d_x_push ignored): push and consume in one call.
 *
 * Bursts (flag protocol, mode 0, option "pdl_independent" = 1): B <= 32 consecutive calls on INDEPENDENT fields, captured
 * in one CUDA graph, may overlap completely under programmatic dependent launch.  Announce each call with
 * mimsem_gpu_set_option(ctx, "halo_burst_len", B) and (ctx, "halo_burst_pos", i), i = 0 .. B-1 in launch order (reset
 * halo_burst_len to 0 afterwards): launch i works on epoch *d_epoch + 1 + i with its own counters, flags and
 * acknowledgements still rise in launch order, and the last launch advances d_epoch by B.  The graph must hold the whole
 * burst; replays of it (and anything else on the stream) are ordered after it as usual.
 */
int mimsem_gpu_apply_M1_halo(mimsem_gpu_ctx* ctx, int lev0, int nlev, int ld, double scale, int tpow, int flags,
                             const double* d_x, double* d_y, const double* d_x_push, int mode, int npush, const void* d_push,
                             int npull, const void* d_pull, const double* d_inbox, int64_t parity_stride, int nbuf,
                             int push_ctas, void* d_epoch, int* d_err, void* stream);

/*
 * The same with the IN-BAND hand-over of the ghost rows: a pushed value travels as one 16-byte cell
 * { lo32(value), epoch32, hi32(value), epoch32 } that validates itself, so the sender needs no system fence and no flag
 * after the copy and the receiver's boundary tiles poll the cells they read (ordinary loads instead of TMA for those
 * rows).  d_inbox_cells = this rank's inbox of 16-byte cells, [nbuf][ghost rows][nlev], cell_stride cells between
 * copies; d_push[].inbox / .inbox_parity_stride address the PEER's cell inbox in the same units.  d_push[].wait and
 * d_pull[].signal carry the acknowledgements as before; d_push[].signal and d_pull[].wait are unused.  The cell
 * inbox, its acknowledgement words and d_epoch must not be shared with mimsem_gpu_halo_push / _pull.
 */
int mimsem_gpu_apply_M1_halo_ll(mimsem_gpu_ctx* ctx, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* d_x,
                                double* d_y, const double* d_x_push, int mode, int npush, const void* d_push, int npull,
                                const void* d_pull, const void* d_inbox_cells, int64_t cell_stride, int nbuf, int push_ctas,
                                void* d_epoch, int* d_err, void* stream);

/*
 * Element-partitioned M1 solve (N GPUs of one box; KSPSolve(ksp1, ...) of eul/HorizSolve.cpp:224, 310, 322 on the
 * partitioned operator).  Same iteration as mimsem_gpu_solve_M1; the operator is the fused ghost-refresh + M1 launch
 * (descriptor fields as the arguments of mimsem_gpu_apply_M1_halo / _halo_ll, ll = 1 for the in-band cells), the vectors
 * live on this rank's OWNED rows (d_b, d_x: local fields; ghost rows of d_x are not written), and each of the two dot
 * products per iteration is completed over peer memory inside the finishing kernel: every rank stores its per-level
 * partial sums into the peers' reduction areas as self-validating 16-byte cells and adds all ranks' sums in rank order,
 * so every rank computes bit-identical step lengths and stops at the same iteration (the call is collective).
 * d_peer_areas: device array of `world` pointers to the ranks' reduction areas ([rank] = this rank's own, the others
 * peer-mapped), each 2 * world * 3 * 64 16-byte cells, zero-initialised; d_seq: device counter (starts at 0, advanced
 * by every reduction, in step on all ranks).
 */
typedef struct mimsem_halo_desc {
    int npush; const void* d_push; int npull; const void* d_pull;
    const void* d_inbox; int64_t stride; int nbuf; int push_ctas; void* d_epoch; int* d_err; int ll;
} mimsem_halo_desc;
typedef struct mimsem_reduce_desc {
    int world, rank; const void* d_peer_areas; void* d_seq; int* d_err;
} mimsem_reduce_desc;
int mimsem_gpu_solve_M1_dist(mimsem_gpu_ctx* ctx, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* d_b,
                             double* d_x, double rtol, int maxit, int* iters, double* relres, const mimsem_halo_desc* halo,
                             const mimsem_reduce_desc* reduce, void* stream);

/* Device memory for hosts that do not link the CUDA runtime themselves (the C++ host layer, mimsem_b200/host/DistEngine):
 * plain cudaMalloc (zero-filled) / cudaFree / cudaMemcpy on the context's device; kind 0 host -> device, 1 device -> host,
 * 2 device -> device (synchronous); _sync waits for the given stream (NULL: the whole device). */
int mimsem_gpu_dev_alloc(mimsem_gpu_ctx* ctx, int64_t bytes, void** d_ptr);
int mimsem_gpu_dev_free(mimsem_gpu_ctx* ctx, void* d_ptr);
int mimsem_gpu_dev_copy(mimsem_gpu_ctx* ctx, void* dst, const void* src, int64_t bytes, int kind);
int mimsem_gpu_dev_sync(mimsem_gpu_ctx* ctx, void* stream);
/* Streams and CUDA graphs for hosts that do not link the CUDA runtime themselves: a non-blocking stream of the context's
 * device; capture of everything subsequently launched on it (thread-local capture mode; warm the sequence up once before,
 * first calls allocate) into an executable graph -- the launch-bound inner loop of a time step, and the form in which a
 * burst of fused launches is issued (see "Bursts" above) --, its replay on a stream, and their release. */
int mimsem_gpu_stream_create(mimsem_gpu_ctx* ctx, void** stream);
int mimsem_gpu_stream_destroy(mimsem_gpu_ctx* ctx, void* stream);
int mimsem_gpu_graph_begin(mimsem_gpu_ctx* ctx, void* stream);
int mimsem_gpu_graph_end(mimsem_gpu_ctx* ctx, void* stream, void** graph_exec);
int mimsem_gpu_graph_launch(mimsem_gpu_ctx* ctx, void* graph_exec, void* stream);
int mimsem_gpu_graph_destroy(mimsem_gpu_ctx* ctx, void* graph_exec);
/* Page-locked host memory (cudaMallocHost / cudaFreeHost) for the buffers handed to mimsem_gpu_apply_host: from pageable
 * memory the copies of its pipeline stages cannot overlap each other or the kernels. */
int mimsem_gpu_host_alloc(mimsem_gpu_ctx* ctx, int64_t bytes, void** h_ptr);
int mimsem_gpu_host_free(mimsem_gpu_ctx* ctx, void* h_ptr);

/* number of kernels this library has launched since the context was created */
int64_t mimsem_gpu_launch_count(const mimsem_gpu_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* MIMSEM_GPU_H */
