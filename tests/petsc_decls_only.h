/* TEST FIXTURE -- declarations only, nothing here is implemented in this repository.
 *
 * The part of PETSc / MPI that the reference's callers use for their ASSEMBLED matrices (the vertical solver's SeqAIJ operators,
 * eul/VertOps.cpp, eul/VertSolve.cpp, eul/Euler_2.cpp:154-170, 255-294) and global reductions.  It lies outside the horizontal
 * operator path; with PETSc it is PETSc's.  tests/test_host_cpp.py appends it to petsc_compat.h so that the reference's
 * eul/Euler_2.cpp can be COMPILED against the host mirror and its use of the mirrored classes checked symbol by symbol.       */
#ifndef MIMSEM_TEST_PETSC_DECLS_ONLY_H
#define MIMSEM_TEST_PETSC_DECLS_ONLY_H
#include "petsc_compat.h"
#ifndef MIMSEM_HAVE_PETSC
typedef const char* MatType;
#define MATSEQAIJ "seqaij"
#define MATMPIAIJ "mpiaij"
#define PCLU "lu"
typedef enum { MAT_INITIAL_MATRIX = 0, MAT_REUSE_MATRIX = 1, MAT_IGNORE_MATRIX = 2, MAT_INPLACE_MATRIX = 3 } MatReuse;
typedef enum { MAT_DO_NOT_COPY_VALUES = 0, MAT_COPY_VALUES = 1, MAT_SHARE_NONZERO_PATTERN = 2 } MatDuplicateOption;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_DOUBLE 11
#define MPI_INT 12
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
PetscErrorCode PetscInitialize(int* argc, char*** args, const char* file, const char* help);
PetscErrorCode PetscFinalize(void);
PetscErrorCode MatCreate(MPI_Comm comm, Mat* A);
PetscErrorCode MatSetSizes(Mat A, PetscInt m, PetscInt n, PetscInt M, PetscInt N);
PetscErrorCode MatSetType(Mat A, MatType type);
PetscErrorCode MatSeqAIJSetPreallocation(Mat A, PetscInt nz, const PetscInt nnz[]);
PetscErrorCode MatMPIAIJSetPreallocation(Mat A, PetscInt dnz, const PetscInt dnnz[], PetscInt onz, const PetscInt onnz[]);
PetscErrorCode MatCreateSeqAIJ(MPI_Comm comm, PetscInt m, PetscInt n, PetscInt nz, const PetscInt nnz[], Mat* A);
PetscErrorCode MatZeroEntries(Mat A);
PetscErrorCode MatSetValues(Mat A, PetscInt m, const PetscInt im[], PetscInt n, const PetscInt in[], const PetscScalar v[], InsertMode mode);
PetscErrorCode MatMatMult(Mat A, Mat B, MatReuse scall, PetscReal fill, Mat* C);
PetscErrorCode MatTranspose(Mat A, MatReuse reuse, Mat* B);
PetscErrorCode MatScale(Mat A, PetscScalar a);
PetscErrorCode MatAYPX(Mat Y, PetscScalar a, Mat X, MatStructure str);
PetscErrorCode MatMultTranspose(Mat A, Vec x, Vec y);
PetscErrorCode MatDiagonalScale(Mat A, Vec l, Vec r);
PetscErrorCode MatDuplicate(Mat A, MatDuplicateOption op, Mat* B);
PetscErrorCode MatGetOwnershipRange(Mat A, PetscInt* lo, PetscInt* hi);
PetscErrorCode MatGetRow(Mat A, PetscInt row, PetscInt* ncols, const PetscInt** cols, const PetscScalar** vals);
PetscErrorCode MatRestoreRow(Mat A, PetscInt row, PetscInt* ncols, const PetscInt** cols, const PetscScalar** vals);
#endif
#endif
