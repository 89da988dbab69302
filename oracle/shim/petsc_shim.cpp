/*
 * ORACLE-ONLY PETSc/MPI shim, definitions (TEST INFRASTRUCTURE; see petsc.h in this directory).
 * Each function implements the documented PETSc semantics that the reference's hot path
 * relies on, restricted to what a rank-by-rank emulation needs.
 */
#include "petsc.h"

static thread_local int g_rank = 0;
static thread_local int g_size = 1;

void ShimSetRank(int rank, int size) {
    g_rank = rank;
    g_size = size;
}

int MPI_Comm_rank(MPI_Comm, int* rank) { *rank = g_rank; return 0; }
int MPI_Comm_size(MPI_Comm, int* size) { *size = g_size; return 0; }

/* ---------------------------------------------------------------- IS */
int ISCreateGeneral(MPI_Comm, int n, const int* idx, PetscCopyMode, IS* is) {
    *is = new _p_IS;
    (*is)->idx.assign(idx, idx + n);
    return 0;
}
int ISCreateStride(MPI_Comm, int n, int first, int step, IS* is) {
    *is = new _p_IS;
    (*is)->idx.resize(n);
    for (int i = 0; i < n; i++) (*is)->idx[i] = first + i * step;
    return 0;
}
int ISDestroy(IS* is) {
    delete *is;
    *is = NULL;
    return 0;
}

/* ---------------------------------------------------------------- Vec */
static Vec vec_new(int n, int N, bool mpi) {
    Vec v = new _p_Vec;
    v->n = n;
    v->N = N;
    v->mpi = mpi;
    v->a = (double*)calloc(n > 0 ? n : 1, sizeof(double));
    return v;
}
int VecCreateSeq(MPI_Comm, int n, Vec* v) { *v = vec_new(n, n, false); return 0; }
int VecCreateMPI(MPI_Comm, int n, int N, Vec* v) { *v = vec_new(n, N, true); return 0; }
int VecDestroy(Vec* v) {
    if (*v) {
        free((*v)->a);
        delete *v;
    }
    *v = NULL;
    return 0;
}
int VecZeroEntries(Vec v) { memset(v->a, 0, sizeof(double) * v->n); return 0; }
int VecGetArray(Vec v, PetscScalar** a) { *a = v->a; return 0; }
int VecRestoreArray(Vec, PetscScalar** a) { *a = NULL; return 0; }
int VecSetValues(Vec v, int n, const int* ix, const PetscScalar* y, InsertMode mode) {
    if (v->mpi) {
        fprintf(stderr, "oracle shim: VecSetValues on an MPI Vec is not emulated\n");
        abort();
    }
    for (int i = 0; i < n; i++) {
        if (ix[i] < 0) continue;
        if (mode == ADD_VALUES) v->a[ix[i]] += y[i];
        else v->a[ix[i]] = y[i];
    }
    return 0;
}
int VecCopy(Vec x, Vec y) { memcpy(y->a, x->a, sizeof(double) * x->n); return 0; }
int VecView(Vec, PetscViewer) { return 0; }
int VecAssemblyBegin(Vec) { return 0; }
int VecAssemblyEnd(Vec) { return 0; }

/* VecScatter: records the index lists.  Cross-rank data motion is the driver's job
 * (it owns every rank's arrays); Begin/End are therefore no-ops here. */
int VecScatterCreate(Vec, IS ix, Vec, IS iy, VecScatter* sc) {
    *sc = new _p_VecScatter;
    (*sc)->from = ix->idx;
    (*sc)->to = iy->idx;
    return 0;
}
int VecScatterBegin(VecScatter, Vec, Vec, InsertMode, ScatterMode) { return 0; }
int VecScatterEnd(VecScatter, Vec, Vec, InsertMode, ScatterMode) { return 0; }
int VecScatterDestroy(VecScatter* sc) {
    delete *sc;
    *sc = NULL;
    return 0;
}

/* ---------------------------------------------------------------- Mat */
int MatCreate(MPI_Comm, Mat* A) {
    *A = new _p_Mat;
    (*A)->m = (*A)->n = (*A)->M = (*A)->N = 0;
    return 0;
}
int MatSetSizes(Mat A, int m, int n, int M, int N) {
    A->m = m; A->n = n; A->M = M; A->N = N;
    return 0;
}
int MatSetType(Mat, const char*) { return 0; }
int MatMPIAIJSetPreallocation(Mat, int, const int*, int, const int*) { return 0; }
int MatSeqAIJSetPreallocation(Mat, int, const int*) { return 0; }
int MatZeroEntries(Mat A) { A->t.clear(); return 0; }
int MatSetValues(Mat A, int m, const int* im, int n, const int* in, const PetscScalar* v, InsertMode mode) {
    size_t base = A->t.size();
    A->t.resize(base + (size_t)m * n);
    ShimTriplet* t = A->t.data() + base;
    for (int i = 0; i < m; i++) {
        for (int j = 0; j < n; j++) {
            t->row = im[i];
            t->col = in[j];
            t->val = v[i * n + j]; /* row-major, the PETSc default */
            t->mode = (int)mode;
            t++;
        }
    }
    return 0;
}
int MatAssemblyBegin(Mat, MatAssemblyType) { return 0; }
int MatAssemblyEnd(Mat, MatAssemblyType) { return 0; }
int MatDestroy(Mat* A) {
    if (*A) delete *A;
    *A = NULL;
    return 0;
}
int MatTranspose(Mat A, MatReuse reuse, Mat* B) {
    if (reuse == MAT_INITIAL_MATRIX) MatCreate(0, B);
    (*B)->m = A->n; (*B)->n = A->m; (*B)->M = A->N; (*B)->N = A->M;
    (*B)->t = A->t;
    for (size_t i = 0; i < (*B)->t.size(); i++) {
        int r = (*B)->t[i].row;
        (*B)->t[i].row = (*B)->t[i].col;
        (*B)->t[i].col = r;
    }
    return 0;
}
int MatDuplicate(Mat A, MatDuplicateOption op, Mat* B) {
    MatCreate(0, B);
    **B = *A;
    if (op == MAT_DO_NOT_COPY_VALUES)
        for (size_t i = 0; i < (*B)->t.size(); i++) (*B)->t[i].val = 0.0;
    return 0;
}
int MatAXPY(Mat Y, PetscScalar a, Mat X, MatStructure) {
    size_t base = Y->t.size();
    Y->t.insert(Y->t.end(), X->t.begin(), X->t.end());
    for (size_t i = base; i < Y->t.size(); i++) {
        /* X's own INSERT duplicates all carry the same value in the reference (incidence stencils) */
        Y->t[i].val *= a;
        Y->t[i].mode = ADD_VALUES;
    }
    return 0;
}
int MatCopy(Mat A, Mat B, MatStructure) {
    B->t = A->t;
    return 0;
}
int MatScale(Mat A, PetscScalar a) {
    for (size_t i = 0; i < A->t.size(); i++) A->t[i].val *= a;
    return 0;
}

/* ---------------------------------------------------------------- viewers: no-ops */
int PetscViewerASCIIOpen(MPI_Comm, const char*, PetscViewer* v) { *v = NULL; return 0; }
int PetscViewerBinaryOpen(MPI_Comm, const char*, PetscFileMode, PetscViewer* v) { *v = NULL; return 0; }
int PetscViewerHDF5Open(MPI_Comm, const char*, PetscFileMode, PetscViewer* v) { *v = NULL; return 0; }
int PetscViewerDestroy(PetscViewer* v) { *v = NULL; return 0; }
