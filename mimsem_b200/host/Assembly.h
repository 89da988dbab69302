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
// Signatures follow eul/ (the 3-D code).  src/ has the same classes without the level / scale arguments
// (src/Assembly.h:11, 23, 34, 77, 153): an eul-signature call with scale 1 on a Geom without layers is
// that operator; the classes that exist ONLY in src/ and carry BASELINE config 2's potential-vorticity
// upwinding (Phmat::assemble_up, RotMat_up) are mirrored below with their src/ signatures.  box/ builds
// Umat/Wmat once in the constructor (box/Assembly.h:9-25), i.e. assemble(0, SCALE, true) here.
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
        void assemble(int lev, double scale, bool vert_scale);
    private:
        MimsemShell* sh;
};

// 2-form mass matrix                                                   eul/Assembly.h:17-28
class Wmat {
    public:
        Wmat(Topo* _topo, Geom* _geom, LagrangeEdge* _e);
        ~Wmat();
        Topo* topo; Geom* geom; LagrangeEdge* e;
        Mat M;
        void assemble(int lev, double scale, bool vert_scale);
    private:
        MimsemShell* sh;
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

#endif
