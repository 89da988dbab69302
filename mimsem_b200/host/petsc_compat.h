// Minimal PETSc / MPI compatibility layer for building the host classes WITHOUT PETSc.
//
// When PETSc is available, compile with -DMIMSEM_HAVE_PETSC: this header then just includes <petsc.h>
// and the adaptor uses the real MATSHELL / Vec / VecScatter.  Otherwise the subset below provides the
// same entry points with the documented PETSc semantics for ONE process that plays R "ranks" in
// lockstep phases (the current rank is set with PetscCompatSetRank): the k-th VecCreateMPI call of
// every rank refers to the same global vector, exactly as a collective call does under MPI.
// This is product code (the host mirror needs *a* PETSc to compile against); it is independent of the
// oracle's shim under oracle/shim/, which exists to compile the reference's sources.
#ifndef MIMSEM_PETSC_COMPAT_H
#define MIMSEM_PETSC_COMPAT_H

#ifdef MIMSEM_HAVE_PETSC
#include <petsc.h>
#else

#include <math.h>
#include <stddef.h>
#include <stdlib.h>

typedef double PetscScalar;
typedef double PetscReal;
typedef int PetscInt;
typedef int PetscErrorCode;
typedef int MPI_Comm;
#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2
#define PETSC_NULL NULL
#define PETSC_DECIDE (-1)

typedef enum { NOT_SET_VALUES = 0, INSERT_VALUES = 1, ADD_VALUES = 2 } InsertMode;
typedef enum { SCATTER_FORWARD = 0, SCATTER_REVERSE = 1 } ScatterMode;
typedef enum { PETSC_COPY_VALUES = 0, PETSC_OWN_POINTER = 1, PETSC_USE_POINTER = 2 } PetscCopyMode;
typedef enum { MATOP_MULT = 3, MATOP_MULT_TRANSPOSE = 5, MATOP_GET_DIAGONAL = 17, MATOP_AXPY = 43, MATOP_DESTROY = 60,
               /* compat only: z = blockdiag(A)^-1 r for a PCBJACOBI request on a shell (PETSc cuts the blocks out of an
                  assembled matrix; a shell that can apply them itself offers this operation) */
               MATOP_COMPAT_PCBJACOBI = 1000 } MatOperation;
typedef enum { DIFFERENT_NONZERO_PATTERN = 0, SUBSET_NONZERO_PATTERN = 1, SAME_NONZERO_PATTERN = 2 } MatStructure;
typedef enum { MAT_FLUSH_ASSEMBLY = 1, MAT_FINAL_ASSEMBLY = 0 } MatAssemblyType;
typedef enum { NORM_1 = 0, NORM_2 = 1, NORM_FROBENIUS = 2, NORM_INFINITY = 3 } NormType;
#define PETSC_DEFAULT (-2)
typedef const char* KSPType;
typedef const char* PCType;
#define KSPGMRES "gmres"
#define KSPCG "cg"
#define PCBJACOBI "bjacobi"
#define PCJACOBI "jacobi"
#define PCSHELL "shell"
#define PCNONE "none"

typedef struct _mimsem_Vec* Vec;
typedef struct _mimsem_Mat* Mat;
typedef struct _mimsem_IS* IS;
typedef struct _mimsem_VecScatter* VecScatter;
typedef struct _mimsem_KSP* KSP;
typedef struct _mimsem_PC* PC;
typedef struct _mimsem_Viewer* PetscViewer;
typedef enum { FILE_MODE_READ = 0, FILE_MODE_WRITE = 1 } PetscFileMode;
#define PETSC_COMM_WORLD MPI_COMM_WORLD
#define PETSC_COMM_SELF MPI_COMM_SELF

/* which of the R in-process ranks is executing (compat only) */
void PetscCompatSetRank(int rank, int size);
/* forget every global vector (between independent tests) */
void PetscCompatReset(void);

int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);

PetscErrorCode ISCreateGeneral(MPI_Comm comm, PetscInt n, const PetscInt idx[], PetscCopyMode mode, IS* is);
PetscErrorCode ISCreateStride(MPI_Comm comm, PetscInt n, PetscInt first, PetscInt step, IS* is);
PetscErrorCode ISDestroy(IS* is);

PetscErrorCode VecCreateSeq(MPI_Comm comm, PetscInt n, Vec* v);
PetscErrorCode VecCreateMPI(MPI_Comm comm, PetscInt n, PetscInt N, Vec* v);
PetscErrorCode VecDestroy(Vec* v);
PetscErrorCode VecZeroEntries(Vec v);
PetscErrorCode VecGetArray(Vec v, PetscScalar** a);         /* the rank's owned slice */
PetscErrorCode VecRestoreArray(Vec v, PetscScalar** a);
PetscErrorCode VecGetLocalSize(Vec v, PetscInt* n);
PetscErrorCode VecGetSize(Vec v, PetscInt* N);
PetscErrorCode VecGetOwnershipRange(Vec v, PetscInt* lo, PetscInt* hi);

/* vector algebra on the calling rank's owned slice; VecDot / VecNorm reduce over the WHOLE vector (a collective) */
PetscErrorCode VecSet(Vec v, PetscScalar a);
PetscErrorCode VecCopy(Vec x, Vec y);
PetscErrorCode VecDuplicate(Vec x, Vec* y);
PetscErrorCode VecScale(Vec v, PetscScalar a);
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x);              /* y += a x */
PetscErrorCode VecAYPX(Vec y, PetscScalar a, Vec x);              /* y = x + a y */
PetscErrorCode VecDot(Vec x, Vec y, PetscScalar* val);
PetscErrorCode VecNorm(Vec x, NormType t, PetscReal* val);
PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y);             /* w = x .* y */
PetscErrorCode VecPointwiseDivide(Vec w, Vec x, Vec y);           /* w = x ./ y */

/* Viewers.  Binary = PETSc's on-disk Vec format (big-endian: int32 VEC_FILE_CLASSID 1211214, int32 N, N float64), the
 * format of the reference's restart files output/<field>_<lev>_<step>.vec (eul/UMJS14.cpp:238-267); ASCII = the layout
 * of VecView on a PetscViewerASCII (header lines, then one value per line).  Collective calls: with several in-process
 * ranks the LAST rank's VecView writes the whole vector, every rank's VecLoad reads its own slice. */
PetscErrorCode PetscViewerBinaryOpen(MPI_Comm comm, const char* name, PetscFileMode mode, PetscViewer* viewer);
PetscErrorCode PetscViewerASCIIOpen(MPI_Comm comm, const char* name, PetscViewer* viewer);
PetscErrorCode PetscViewerDestroy(PetscViewer* viewer);
PetscErrorCode VecView(Vec v, PetscViewer viewer);
PetscErrorCode VecLoad(Vec v, PetscViewer viewer);

PetscErrorCode VecScatterCreate(Vec x, IS ix, Vec y, IS iy, VecScatter* sc);
PetscErrorCode VecScatterBegin(VecScatter sc, Vec x, Vec y, InsertMode addv, ScatterMode mode);
PetscErrorCode VecScatterEnd(VecScatter sc, Vec x, Vec y, InsertMode addv, ScatterMode mode);
PetscErrorCode VecScatterDestroy(VecScatter* sc);

PetscErrorCode MatCreateShell(MPI_Comm comm, PetscInt m, PetscInt n, PetscInt M, PetscInt N, void* ctx, Mat* A);
PetscErrorCode MatShellSetOperation(Mat A, MatOperation op, void (*f)(void));
PetscErrorCode MatShellGetContext(Mat A, void* ctx);
PetscErrorCode MatMult(Mat A, Vec x, Vec y);
PetscErrorCode MatDestroy(Mat* A);
PetscErrorCode MatGetDiagonal(Mat A, Vec d);                      /* MATOP_GET_DIAGONAL of the shell */
PetscErrorCode MatAXPY(Mat Y, PetscScalar a, Mat X, MatStructure str);   /* Y += a X: MATOP_AXPY of the shell Y */
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType type);     /* shells have nothing to assemble */
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType type);

/* Krylov solver on a shell operator, ONE rank (with several in-process ranks KSPSolve is not available: a collective
 * cannot be played rank after rank).  KSPGMRES (PETSc's default type, what the reference runs): restarted GMRES(30), left
 * preconditioning, modified Gram-Schmidt, converged when the preconditioned residual norm falls below
 * max(rtol |B b|, abstol) -- PETSc's default test.  KSPCG: preconditioned conjugate gradients, same test on the true residual.
 * Preconditioner B: PCSHELL -> the callback; PCBJACOBI -> the shell's MATOP_COMPAT_PCBJACOBI if it offers one (the Umat
 * shell: the reference's element blocks); otherwise, and for PCJACOBI, the shell's MATOP_GET_DIAGONAL; PCNONE -> identity. */
PetscErrorCode KSPCreate(MPI_Comm comm, KSP* ksp);
PetscErrorCode KSPDestroy(KSP* ksp);
PetscErrorCode KSPSetOperators(KSP ksp, Mat A, Mat P);
PetscErrorCode KSPSetTolerances(KSP ksp, PetscReal rtol, PetscReal abstol, PetscReal dtol, PetscInt maxits);
PetscErrorCode KSPSetType(KSP ksp, KSPType type);
PetscErrorCode KSPSetOptionsPrefix(KSP ksp, const char* prefix);
PetscErrorCode KSPSetFromOptions(KSP ksp);
PetscErrorCode KSPGetPC(KSP ksp, PC* pc);
PetscErrorCode KSPSolve(KSP ksp, Vec b, Vec x);
PetscErrorCode KSPGetIterationNumber(KSP ksp, PetscInt* its);
PetscErrorCode KSPGetResidualNorm(KSP ksp, PetscReal* rnorm);
PetscErrorCode PCSetType(PC pc, PCType type);
PetscErrorCode PCBJacobiSetTotalBlocks(PC pc, PetscInt blocks, const PetscInt lens[]);
PetscErrorCode PCApply(PC pc, Vec x, Vec y);                      /* PCSHELL only */
PetscErrorCode PCShellSetApply(PC pc, PetscErrorCode (*apply)(PC, Vec, Vec));
PetscErrorCode PCShellSetContext(PC pc, void* ctx);
PetscErrorCode PCShellGetContext(PC pc, void* ctx);

#endif /* MIMSEM_HAVE_PETSC */
#endif
