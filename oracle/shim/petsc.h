/*
 * ORACLE-ONLY PETSc/MPI shim  (TEST INFRASTRUCTURE -- not part of the product path).
 *
 * Purpose: let the reference's own hot-path translation units
 *   /root/reference/{src,eul,box}/{Basis,LinAlg,ElMats,Topo,Geom,Assembly}.cpp
 * compile UNMODIFIED with plain g++ in a container that has no PETSc and no MPI
 * (oracle O1).  The shim emulates R MPI ranks inside one
 * process: the "current rank" is a thread_local integer set by the driver, so
 * every emulated rank may run on its own std::thread.
 *
 *   Mat   = triplet accumulator in GLOBAL indices that remembers INSERT vs ADD
 *           (what MatSetValues + MatAssemblyBegin/End mean for a MATMPIAIJ matrix,
 *            once the driver merges the triplets of all emulated ranks)
 *   Vec   = plain array; an "MPI" Vec holds only the rank's owned slice
 *   IS / VecScatter = the index list only (the driver performs cross-rank
 *           gathers / scatter-adds itself, through Topo::loc*)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may link this.
 */
#ifndef ORACLE_PETSC_SHIM_H
#define ORACLE_PETSC_SHIM_H

#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <vector>

typedef double PetscScalar;
typedef double PetscReal;
typedef int PetscInt;
typedef int PetscErrorCode;
typedef int PetscMPIInt;
typedef int MPI_Comm;
typedef bool PetscBool;

#define MPI_COMM_WORLD 91
#define MPI_COMM_SELF 92
#define PETSC_NULL NULL
#define PETSC_TRUE true
#define PETSC_FALSE false
#define MATMPIAIJ "mpiaij"
#define MATSEQAIJ "seqaij"

enum InsertMode { NOT_SET_VALUES = 0, INSERT_VALUES = 1, ADD_VALUES = 2 };
enum ScatterMode { SCATTER_FORWARD = 0, SCATTER_REVERSE = 1 };
enum MatAssemblyType { MAT_FLUSH_ASSEMBLY = 1, MAT_FINAL_ASSEMBLY = 0 };
enum MatReuse { MAT_INITIAL_MATRIX = 0, MAT_REUSE_MATRIX = 1, MAT_IGNORE_MATRIX = 2, MAT_INPLACE_MATRIX = 3 };
enum MatDuplicateOption { MAT_DO_NOT_COPY_VALUES = 0, MAT_COPY_VALUES = 1, MAT_SHARE_NONZERO_PATTERN = 2 };
enum MatStructure { DIFFERENT_NONZERO_PATTERN = 0, SUBSET_NONZERO_PATTERN = 1, SAME_NONZERO_PATTERN = 2 };
enum PetscCopyMode { PETSC_COPY_VALUES = 0, PETSC_OWN_POINTER = 1, PETSC_USE_POINTER = 2 };
enum PetscFileMode { FILE_MODE_READ = 0, FILE_MODE_WRITE = 1, FILE_MODE_APPEND = 2 };

struct ShimTriplet {
    int row, col;
    double val;
    int mode; /* InsertMode */
};

struct _p_Mat {
    int m, n, M, N;
    std::vector<ShimTriplet> t;
};
struct _p_Vec {
    int n;      /* local length                     */
    int N;      /* global length (== n for Seq)     */
    bool mpi;
    double* a;
};
struct _p_IS {
    std::vector<int> idx;
};
struct _p_VecScatter {
    std::vector<int> from, to;
};
struct _p_PetscViewer {
    int dummy;
};

typedef _p_Mat* Mat;
typedef _p_Vec* Vec;
typedef _p_IS* IS;
typedef _p_VecScatter* VecScatter;
typedef _p_PetscViewer* PetscViewer;

/* rank emulation (driver side) */
void ShimSetRank(int rank, int size);

int MPI_Comm_rank(MPI_Comm, int* rank);
int MPI_Comm_size(MPI_Comm, int* size);

int ISCreateGeneral(MPI_Comm, int n, const int* idx, PetscCopyMode, IS* is);
int ISCreateStride(MPI_Comm, int n, int first, int step, IS* is);
int ISDestroy(IS* is);

int VecCreateSeq(MPI_Comm, int n, Vec* v);
int VecCreateMPI(MPI_Comm, int n, int N, Vec* v);
int VecDestroy(Vec* v);
int VecZeroEntries(Vec v);
int VecGetArray(Vec v, PetscScalar** a);
int VecRestoreArray(Vec v, PetscScalar** a);
int VecSetValues(Vec v, int n, const int* ix, const PetscScalar* y, InsertMode mode);
int VecCopy(Vec x, Vec y);
int VecView(Vec v, PetscViewer viewer);
int VecAssemblyBegin(Vec v);
int VecAssemblyEnd(Vec v);

int VecScatterCreate(Vec x, IS ix, Vec y, IS iy, VecScatter* sc);
int VecScatterBegin(VecScatter sc, Vec x, Vec y, InsertMode, ScatterMode);
int VecScatterEnd(VecScatter sc, Vec x, Vec y, InsertMode, ScatterMode);
int VecScatterDestroy(VecScatter* sc);

int MatCreate(MPI_Comm, Mat* A);
int MatSetSizes(Mat A, int m, int n, int M, int N);
int MatSetType(Mat A, const char* type);
int MatMPIAIJSetPreallocation(Mat A, int dnz, const int* dnnz, int onz, const int* onnz);
int MatSeqAIJSetPreallocation(Mat A, int nz, const int* nnz);
int MatZeroEntries(Mat A);
int MatSetValues(Mat A, int m, const int* im, int n, const int* in, const PetscScalar* v, InsertMode mode);
int MatAssemblyBegin(Mat A, MatAssemblyType);
int MatAssemblyEnd(Mat A, MatAssemblyType);
int MatDestroy(Mat* A);
int MatTranspose(Mat A, MatReuse reuse, Mat* B);
int MatDuplicate(Mat A, MatDuplicateOption op, Mat* B);
int MatAXPY(Mat Y, PetscScalar a, Mat X, MatStructure);
int MatCopy(Mat A, Mat B, MatStructure);
int MatScale(Mat A, PetscScalar a);

int PetscViewerASCIIOpen(MPI_Comm, const char* name, PetscViewer* v);
int PetscViewerBinaryOpen(MPI_Comm, const char* name, PetscFileMode, PetscViewer* v);
int PetscViewerHDF5Open(MPI_Comm, const char* name, PetscFileMode, PetscViewer* v);
int PetscViewerDestroy(PetscViewer* v);

#endif
