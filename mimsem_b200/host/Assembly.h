// Host mirror of the reference's operator classes for the horizontal hot path (eul/Assembly.h).
//
// Same class names, constructor arguments, assemble(...) signatures and public `Mat` members as the
// reference, so HorizSolve / Euler / VertSolve compile against this header unchanged.  What differs is
// what happens behind them: assemble(...) only records (level, scale, flags, coefficient) -- no element
// loop, no MatSetValues, no MatAssembly -- and the public Mat is a PETSc MATSHELL whose MatMult
//   1. gathers the ghosted local input (Topo::gtol_*, INSERT_VALUES, SCATTER_FORWARD),
//   2. runs the matrix-free CUDA kernel of libmimsem_gpu.so on this rank's patch (mimsem_gpu_apply_host),
//   3. sums shared DOFs into the global output (Topo::gtol_*, ADD_VALUES, SCATTER_REVERSE),
// which is exactly the structure of the reference's own matrix-free twins (Uvec::assemble,
// eul/Assembly.cpp:2124-2196).  Compile with -DMIMSEM_HAVE_PETSC against real PETSc, or against
// petsc_compat.h where PETSc is unavailable.
//
// ONE header serves the three directories of the reference: the eul/ (3-D) signatures, the src/ (2-D shallow water)
// overloads without level / scale arguments (src/Assembly.h:11, 23, 34, 77, 153) and the box/ members (box/Assembly.h:9-25:
// Umat and Wmat carry a second matrix `Mo` and are built once, in the constructor, with the level-0 thickness -- the
// constructors do that when the Topo is a box).  The classes that exist only in src/ and carry BASELINE config 2's
// potential-vorticity upwinding (Phmat::assemble_up, RotMat_up) have their src/ signatures.
#ifndef MIMSEM_HOST_ASSEMBLY_H
#define MIMSEM_HOST_ASSEMBLY_H

#include "Basis.h"
#include "ElMats.h"
#include "Geom.h"
#include "Topo.h"
#include "petsc_compat.h"

#define SCALE 1.0e+8   /* eul/Assembly.cpp:20 */

struct MimsemShell;   // per-operator MatShell context (Assembly.cpp)

// 1-form mass matrix                                                   eul/Assembly.h:1-15
class Umat {
    public:
        Umat(Topo* _topo, Geom* _geom, LagrangeNode* _l, LagrangeEdge* _e);
        ~Umat();
        Topo* topo; Geom* geom; LagrangeNode* l; LagrangeEdge* e;
        Mat M;
        Mat MT;    // unused by the live reference code (assemble_up only), kept NULL
        Mat Mo;    // box/: the mass matrix without the vertical scaling (box/Assembly.h:13); NULL elsewhere
        void assemble(int lev, double scale, bool vert_scale);   // eul/Assembly.cpp:51-153
        void assemble();                                         // src/Assembly.cpp:30-124 (no layers, scale 1)
    private:
        MimsemShell* sh;
        MimsemShell* sho;
};

// 2-form mass matrix                                                   eul/Assembly.h:17-28
class Wmat {
    public:
        Wmat(Topo* _topo, Geom* _geom, LagrangeEdge* _e);
        ~Wmat();
        Topo* topo; Geom* geom; LagrangeEdge* e;
        Mat M;
        Mat Mo;    // box/Assembly.h:25
        void assemble(int lev, double scale, bool vert_scale);   // eul/Assembly.cpp:311-373
        void assemble();                                         // src/Assembly.cpp:260-309
    private:
        MimsemShell* sh;
        MimsemShell* sho;
};

// 0-form mass matrix                                                   eul/Assembly.h:333-347
class Pmat {
    public:
        Pmat(Topo* _topo, Geom* _geom, LagrangeNode* _node);
        ~Pmat();
        Topo* topo; Geom* geom; LagrangeNode* node;
        Mat M;
        void assemble(int lev, double scale);
        void assemble_h(int lev, double scale, Vec h2);
        void assemble();                                         // src/Assembly.cpp:324-372
    private:
        MimsemShell* sh;
};

// 1-form mass matrix weighted by a 2-form                              eul/Assembly.h:30-59
class Uhmat {
    public:
        Uhmat(Topo* _topo, Geom* _geom, LagrangeNode* _l, LagrangeEdge* _e);
        ~Uhmat();
        Topo* topo; Geom* geom; LagrangeNode* l; LagrangeEdge* e;
        Mat M;
        Mat MT;
        void assemble(Vec h2, int lev, bool const_vert, double scale);
        void assemble(Vec h2);                                   // src/Assembly.cpp:675-734
    private:
        MimsemShell* sh;
};

// 1-form mass matrix with the Rayleigh-friction point weight          eul/Assembly.h:325-335, eul/Assembly.cpp:1846-1979
class Umat_ray {
    public:
        Umat_ray(Topo* _topo, Geom* _geom, LagrangeNode* _l, LagrangeEdge* _e);
        ~Umat_ray();
        Topo* topo; Geom* geom; LagrangeNode* l; LagrangeEdge* e;
        Mat M;
        void assemble(int lev, double scale, double dt, Vec exner, Vec exner_s);
    private:
        MimsemShell* sh;
};

// 2-form mass matrix weighted by a 2-form                              eul/Assembly.h:186-199
class Whmat {
    public:
        Whmat(Topo* _topo, Geom* _geom, LagrangeEdge* _e);
        ~Whmat();
        Topo* topo; Geom* geom; LagrangeEdge* e;
        Mat M;
        void assemble(Vec rho, int lev, double scale, bool vert_scale_rho);
    private:
        MimsemShell* sh;
};

// kinetic-energy operator, 1-form -> 2-form                            eul/Assembly.h:121-146
class WtQUmat {
    public:
        WtQUmat(Topo* _topo, Geom* _geom, LagrangeNode* _l, LagrangeEdge* _e);
        ~WtQUmat();
        Topo* topo; Geom* geom; LagrangeNode* l; LagrangeEdge* e;
        Mat M;
        void assemble(Vec u1, int lev, double scale);   // u1: ghosted local 1-form (VecCreateSeq(topo->n1))
        void assemble(Vec u1);                          // src/Assembly.cpp:1172-1218
    private:
        MimsemShell* sh;
};

// rotational term: 1-form mass matrix weighted by a 0-form             eul/Assembly.h:148-170
class RotMat {
    public:
        RotMat(Topo* _topo, Geom* _geom, LagrangeNode* _l, LagrangeEdge* _e);
        ~RotMat();
        Topo* topo; Geom* geom; LagrangeNode* l; LagrangeEdge* e;
        Mat M;
        void assemble(Vec q0, int lev, double scale);   // q0: ghosted local 0-form; eul/Assembly.cpp:1030-1083
        void assemble(Vec q0);                          // src/Assembly.cpp:1346-1395 (no layers, scale 1)
    private:
        MimsemShell* sh;
};

// rotational term with the potential vorticity upwinded along the velocity      src/Assembly.h:227-250
class RotMat_up {
    public:
        RotMat_up(Topo* _topo, Geom* _geom, LagrangeNode* _l, LagrangeEdge* _e);
        ~RotMat_up();
        Topo* topo; Geom* geom; LagrangeNode* l; LagrangeEdge* e;
        Mat M;
        void assemble(Vec q0, Vec ul, double tau, double dt);   // src/Assembly.cpp:1784-1853; q0, ul ghosted local
    private:
        MimsemShell* sh;
};

// 0-form mass matrix weighted by a 2-form, optionally with the trial space upwinded      src/Assembly.h:37-49
class Phmat {
    public:
        Phmat(Topo* _topo, Geom* _geom, LagrangeNode* _node);
        ~Phmat();
        Topo* topo; Geom* geom; LagrangeNode* node;
        Mat M;
        void assemble(Vec h2);                                      // src/Assembly.cpp:396-442
        void assemble_up(Vec ul, Vec hl, double fac, double dt);    // src/Assembly.cpp:499-567
    private:
        MimsemShell* sh;
};

// 1-form mass matrix of the horizontal-vorticity terms                  eul/Assembly.h:201-229
class Ut_mat {
    public:
        Ut_mat(Topo* _topo, Geom* _geom, LagrangeNode* _l, LagrangeEdge* _e);
        ~Ut_mat();
        Topo* topo; Geom* geom; LagrangeNode* l; LagrangeEdge* e;
        Mat M;
        void assemble(int lev, double scale);              // weighted by the mean thickness of levels lev, lev+1 (eul/Assembly.cpp:1338-1388)
        void assemble_h(int lev, double scale, Vec rho);   // weighted by rho, no thickness factor (eul/Assembly.cpp:1390-1440)
    private:
        MimsemShell* sh;
};

// vertical-momentum vorticity term, 1-form -> 2-form                    eul/Assembly.h:257-280
class WtQdUdz_mat {
    public:
        WtQdUdz_mat(Topo* _topo, Geom* _geom, LagrangeNode* _l, LagrangeEdge* _e);
        ~WtQdUdz_mat();
        Topo* topo; Geom* geom; LagrangeNode* l; LagrangeEdge* e;
        Mat M;
        void assemble(Vec u1, double scale);               // u1: ghosted local 1-form (eul/Assembly.cpp:1581-1640)
    private:
        MimsemShell* sh;
};

// vertical-vorticity term of the horizontal momentum equation, 2-form -> 1-form (= WtQdUdz_mat^T)   eul/Assembly.h:231-255
class UtQWmat {
    public:
        UtQWmat(Topo* _topo, Geom* _geom, LagrangeNode* _l, LagrangeEdge* _e);
        ~UtQWmat();
        Topo* topo; Geom* geom; LagrangeNode* l; LagrangeEdge* e;
        Mat M;
        void assemble(Vec u1, double scale);               // u1: ghosted local 1-form (eul/Assembly.cpp:1490-1538)
    private:
        MimsemShell* sh;
};

// element-block inverse of the 2-form mass matrix                      eul/Assembly.h:282-303
class WmatInv {
    public:
        WmatInv(Topo* _topo, Geom* _geom, LagrangeEdge* _e);
        ~WmatInv();
        Topo* topo; Geom* geom; LagrangeEdge* e;
        Mat M;
        void assemble(int lev, double scale);              // eul/Assembly.cpp:1673-1722
    private:
        MimsemShell* sh;
};
class WhmatInv {
    public:
        WhmatInv(Topo* _topo, Geom* _geom, LagrangeEdge* _e);
        ~WhmatInv();
        Topo* topo; Geom* geom; LagrangeEdge* e;
        Mat M;
        void assemble(Vec rho, int lev, double scale);     // eul/Assembly.cpp:1744-1800
    private:
        MimsemShell* sh;
};

// lumped 0-form mass "vectors": the diagonal of Pmat / Pmat::assemble_h (exact when m == p)      eul/Assembly.h:61-87
class Pvec {
    public:
        Pvec(Topo* _topo, Geom* _geom, LagrangeNode* _l);
        ~Pvec();
        Topo* topo; Geom* geom; LagrangeNode* l;
        Vec vl;
        Vec vg;
        Vec vg1;                                           // box/Assembly.h:69: level 0 with scale 1, built once (box only)
        void assemble(int lev, double scale);              // eul/Assembly.cpp:602-628
};
class Phvec {
    public:
        Phvec(Topo* _topo, Geom* _geom, LagrangeNode* _l);
        ~Phvec();
        Topo* topo; Geom* geom; LagrangeNode* l;
        Vec vl;
        Vec vg;
        void assemble(Vec hl, int lev, double scale);      // eul/Assembly.cpp:652-681
};

// quadrature-point values -> 0-form (start-up only: the Coriolis vector, eul/HorizSolve.cpp:141-151); host loop
class PtQmat {
    public:
        PtQmat(Topo* _topo, Geom* _geom, LagrangeNode* _l);
        ~PtQmat();
        Topo* topo; Geom* geom; LagrangeNode* l;
        Mat M;
        void assemble();                                   // eul/Assembly.cpp:758-800: a no-op here, the shell is matrix-free
    private:
        Vec xl, yl;
};

// quadrature-point values -> 2-form: y2 = W^T Q x (start-up only: the initial density / pressure fields,
// eul/Euler_2.cpp:493, 535); host loop                                              eul/Assembly.h (WtQmat)
class WtQmat {
    public:
        WtQmat(Topo* _topo, Geom* _geom, LagrangeEdge* _e);
        ~WtQmat();
        Topo* topo; Geom* geom; LagrangeEdge* e;
        Mat M;
        void assemble();                                   // a no-op here, the shell is matrix-free
};

// quadrature-point vector values (two interleaved Cartesian-tangent components per point) -> 1-form:
// y1 = [U V]^T Q J^T u (start-up only: the initial velocity, eul/Euler_2.cpp:432); host loop    eul/Assembly.h (UtQmat)
class UtQmat {
    public:
        UtQmat(Topo* _topo, Geom* _geom, LagrangeNode* _l, LagrangeEdge* _e);
        ~UtQmat();
        Topo* topo; Geom* geom; LagrangeNode* l; LagrangeEdge* e;
        Mat M;
        void assemble();
};

// matrix-free twins: the action of the 1-form mass matrix (optionally weighted by a 2-form) on a ghosted local
// velocity, accumulated in the ghosted local vector vl; vg = reverse ADD scatter of vl      eul/Assembly.h:350-369
class Uvec {
    public:
        Uvec(Topo* _topo, Geom* _geom, LagrangeNode* _node, LagrangeEdge* _edge);
        ~Uvec();
        Topo* topo; Geom* geom; LagrangeNode* node; LagrangeEdge* edge;
        Vec vl;
        Vec vg;
        void assemble(int lev, double scale, bool vert_scale, Vec vel);                                 // eul/Assembly.cpp:2124-2196
        void assemble_hu(int lev, double scale, Vec vel, Vec rho, bool zero_and_scatter, double fac);   // eul/Assembly.cpp:2198-2279
        void assemble_hu(int lev, double scale, bool vert_scale, Vec vel, Vec rho);                     // box/Assembly.cpp:2862-2925
    private:
        void accumulate(int op, int lev, double scale, int tpow, Vec vel, Vec rho);
};

// 2-form twins                                                           eul/Assembly.h:371-384
class Wvec {
    public:
        Wvec(Topo* _topo, Geom* _geom, LagrangeEdge* _edge);
        ~Wvec();
        Topo* topo; Geom* geom; LagrangeEdge* edge;
        Vec vg;
        void assemble(int lev, double scale, bool vert_scale, Vec rho);    // vg = M2 rho              eul/Assembly.cpp:2441-2495
        // vg = WtQUmat(vel2) vel1 -- the kinetic-energy form the reference's routine means to evaluate (its own loop is
        // dead code in eul/ and mis-indexes in box/)                            eul/Assembly.cpp:2497-2545
        void assemble_K(int lev, double scale, Vec vel1, Vec vel2);
        void assemble_K(int lev, double scale, bool vert_scale, Vec vel1, Vec vel2);                    // box/Assembly.cpp:3047-3095
};

// edge-node incidence and its negative transpose                       eul/Assembly.h:171-177
class E10mat {
    public:
        E10mat(Topo* _topo);
        ~E10mat();
        Topo* topo;
        Mat E10;
        Mat E01;
    private:
        MimsemShell *sh10, *sh01;
};

// face-edge incidence and its negative transpose                       eul/Assembly.h:179-185
class E21mat {
    public:
        E21mat(Topo* _topo);
        ~E21mat();
        Topo* topo;
        Mat E21;
        Mat E12;
    private:
        MimsemShell *sh21, *sh12;
};

// Incidence operators need no geometry, but the device context is created per (Topo, Geom) patch; construct at
// least one geometric operator (or call this) before the first E10mat / E21mat MatMult of a Topo.
int MimsemAttachPatch(Topo* topo, Geom* geom, LagrangeNode* l, LagrangeEdge* e);
// last error text of the device library for the calling thread
const char* MimsemLastError(void);
// wall-clock seconds the shells of this process have spent inside the device library (host -> device copies, kernels,
// device -> host copies) since the last reset: separates the device half of a MatMult from PETSc's scatters around it
double MimsemDeviceSeconds(int reset);

// All levels of one operator in ONE device call: y[k] = Op(level lev0 + k) x[k], k = 0 .. nlev-1 -- the reference's
//     for (kk = 0; kk < nk; kk++) { M1->assemble(kk, SCALE, true); MatMult(M1->M, velx[kk], Mu[kk]); }   (eul/Euler_2.cpp:1427-1456)
// with the levels crossing PCIe and the kernels as one pipelined call (equal to the level-by-level result to rounding).  Scale,
// thickness power and flags are those of the shell's last assemble(); coeff[k] = the coefficient of level k in the
// convention of the class's assemble() (Uhmat: 2-form Vec, WtQUmat: ghosted local 1-form, ...; NULL when there is none).
// Nonzero: not a shell of this library / operator without a batched form (the upwinded ones, Umat_ray, the inverses).
PetscErrorCode MimsemMatMultLevels(Mat M, int lev0, int nlev, Vec* x, Vec* y, Vec* coeff);

// The reference's preconditioner of its 1-form solves -- PCBJACOBI with one block per element
// (PCBJacobiSetTotalBlocks(pc, size * nElsX * nElsX, NULL), eul/HorizSolve.cpp:77-84, 791-796) -- for a Umat shell:
// PETSc's PCBJACOBI cuts its blocks out of an ASSEMBLED matrix, which a MatShell does not have, so the blocks are
// tabulated, factorised and solved on the device (mimsem_gpu_pc_bjacobi_M1) and handed to the KSP as a PCSHELL:
//     KSPSetOperators(ksp1, M1->M, M1->M);  MimsemKSPSetElementBlockJacobi(ksp1, M1->M);   // instead of PCSetType(pc, PCBJACOBI)
// The block data follow M1->assemble(lev, scale, vert_scale) like the shell's MatMult.  (The in-tree compatibility layer
// routes a plain PCBJACOBI request on such a shell here by itself, so that the reference's callers run unchanged.)
// Also accepts the Pmat shell (KSPSolve(ksp0, ...), eul/HorizSolve.cpp:87-96): M0 is diagonal, its blocks are its diagonal.
// Returns nonzero when M is neither.
PetscErrorCode MimsemPCApplyBJacobi(PC pc, Vec r, Vec z);     /* PCShellSetApply callback; PCShellSetContext(pc, M) */
PetscErrorCode MimsemKSPSetElementBlockJacobi(KSP ksp, Mat M);
// host half of that preconditioner's set-up, for tests without a GPU: builds the element tables of its context (the patch
// plus copies of the west / south neighbours across the patch boundary, checked against the point coordinates);
// sizes = {elements, neighbour copies, n0, n1, n2, nq}.  0: fine, 1: configuration not covered, 2: inconsistent tables
int MimsemPCTablesCheck(Topo* topo, Geom* geom, int sizes[6]);

#endif
