// See petsc_compat.h.  Compiled only when PETSc is absent.
#ifndef MIMSEM_HAVE_PETSC
#include "petsc_compat.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <utility>
#include <vector>

namespace {
int g_rank = 0, g_size = 1;

struct GlobalVec {            // one collective VecCreateMPI across all in-process ranks
    std::vector<double> a;    // the whole vector
    std::vector<int> nlocal;  // owned size per rank (-1: that rank has not created it yet)
    int refs = 0;
    // Reverse ADD scatters are collective: under MPI every rank zeroes / fills its slice BEFORE any remote
    // contribution lands.  With the ranks played one after the other the contributions are therefore deferred
    // until somebody reads the vector.
    std::vector<std::pair<int, double> > pending;
    void flush() {
        for (size_t i = 0; i < pending.size(); i++) a[pending[i].first] += pending[i].second;
        pending.clear();
    }
};
std::map<long, GlobalVec*> g_globals;       // creation sequence number -> storage
std::vector<long> g_seq;                    // per rank: number of VecCreateMPI calls so far
}  // namespace

struct _mimsem_Vec {
    bool mpi;
    int n, N, rank;
    std::vector<double> local;   // Seq storage
    GlobalVec* g;                // MPI storage (shared)
    long seq;
};
struct _mimsem_IS {
    std::vector<int> idx;
};
struct _mimsem_VecScatter {
    std::vector<int> from, to;   // from: indices into x (global numbering if x is MPI), to: into y
};
struct _mimsem_Mat {
    void* ctx;
    PetscErrorCode (*mult)(Mat, Vec, Vec);
    PetscErrorCode (*destroy)(Mat);
};

void PetscCompatSetRank(int rank, int size) {
    g_rank = rank;
    g_size = size;
    if ((int)g_seq.size() < size) g_seq.resize(size, 0);
}
void PetscCompatReset(void) {
    for (auto& kv : g_globals) delete kv.second;
    g_globals.clear();
    g_seq.assign(g_seq.size(), 0);
}
int MPI_Comm_rank(MPI_Comm, int* rank) { *rank = g_rank; return 0; }
int MPI_Comm_size(MPI_Comm, int* size) { *size = g_size; return 0; }

PetscErrorCode ISCreateGeneral(MPI_Comm, PetscInt n, const PetscInt idx[], PetscCopyMode, IS* is) {
    *is = new _mimsem_IS;
    (*is)->idx.assign(idx, idx + n);
    return 0;
}
PetscErrorCode ISCreateStride(MPI_Comm, PetscInt n, PetscInt first, PetscInt step, IS* is) {
    *is = new _mimsem_IS;
    (*is)->idx.resize(n);
    for (int i = 0; i < n; i++) (*is)->idx[i] = first + i * step;
    return 0;
}
PetscErrorCode ISDestroy(IS* is) { delete *is; *is = NULL; return 0; }

PetscErrorCode VecCreateSeq(MPI_Comm, PetscInt n, Vec* v) {
    Vec x = new _mimsem_Vec;
    x->mpi = false; x->n = n; x->N = n; x->rank = g_rank; x->g = NULL; x->seq = -1;
    x->local.assign(n, 0.0);
    *v = x;
    return 0;
}
PetscErrorCode VecCreateMPI(MPI_Comm, PetscInt n, PetscInt N, Vec* v) {
    if ((int)g_seq.size() < g_size) g_seq.resize(g_size, 0);
    const long seq = g_seq[g_rank]++;
    GlobalVec*& g = g_globals[seq];
    if (!g) {
        g = new GlobalVec;
        g->a.assign(N, 0.0);
        g->nlocal.assign(g_size, -1);
    }
    if ((int)g->a.size() != N) {
        std::fprintf(stderr, "petsc_compat: ranks disagree on the %ld-th VecCreateMPI (N = %d vs %zu)\n", seq, N, g->a.size());
        std::abort();
    }
    g->nlocal[g_rank] = n;
    g->refs++;
    Vec x = new _mimsem_Vec;
    x->mpi = true; x->n = n; x->N = N; x->rank = g_rank; x->g = g; x->seq = seq;
    *v = x;
    return 0;
}
PetscErrorCode VecDestroy(Vec* v) {
    if (*v) {
        if ((*v)->mpi && --(*v)->g->refs == 0) {
            g_globals.erase((*v)->seq);
            delete (*v)->g;
        }
        delete *v;
    }
    *v = NULL;
    return 0;
}
static int rstart_of(Vec v) {
    int lo = 0;
    for (int r = 0; r < v->rank; r++) {
        if (v->g->nlocal[r] < 0) {
            std::fprintf(stderr, "petsc_compat: rank %d has not created this MPI Vec yet (collective order)\n", r);
            std::abort();
        }
        lo += v->g->nlocal[r];
    }
    return lo;
}
PetscErrorCode VecZeroEntries(Vec v) {
    if (v->mpi) std::memset(v->g->a.data() + rstart_of(v), 0, sizeof(double) * v->n);
    else std::fill(v->local.begin(), v->local.end(), 0.0);
    return 0;
}
PetscErrorCode VecGetArray(Vec v, PetscScalar** a) {
    if (v->mpi) v->g->flush();
    *a = v->mpi ? v->g->a.data() + rstart_of(v) : v->local.data();
    return 0;
}
PetscErrorCode VecRestoreArray(Vec, PetscScalar** a) { *a = NULL; return 0; }
PetscErrorCode VecGetLocalSize(Vec v, PetscInt* n) { *n = v->n; return 0; }
PetscErrorCode VecGetSize(Vec v, PetscInt* N) { *N = v->N; return 0; }
PetscErrorCode VecGetOwnershipRange(Vec v, PetscInt* lo, PetscInt* hi) {
    const int s = v->mpi ? rstart_of(v) : 0;
    if (lo) *lo = s;
    if (hi) *hi = s + v->n;
    return 0;
}

PetscErrorCode VecScatterCreate(Vec, IS ix, Vec, IS iy, VecScatter* sc) {
    *sc = new _mimsem_VecScatter;
    (*sc)->from = ix->idx;
    (*sc)->to = iy->idx;
    return 0;
}
static double* base_of(Vec v) { return v->mpi ? v->g->a.data() : v->local.data(); }   // MPI: GLOBAL indexing
PetscErrorCode VecScatterBegin(VecScatter sc, Vec x, Vec y, InsertMode addv, ScatterMode mode) {
    // FORWARD: y[to[i]] (op)= x[from[i]];  REVERSE: roles of the index sets swap (PETSc manual, VecScatterBegin)
    const std::vector<int>& src = (mode == SCATTER_FORWARD) ? sc->from : sc->to;
    const std::vector<int>& dst = (mode == SCATTER_FORWARD) ? sc->to : sc->from;
    if (x->mpi) x->g->flush();
    const double* xs = base_of(x);
    double* yd = base_of(y);
    for (size_t i = 0; i < src.size(); i++) {
        if (addv == ADD_VALUES) {
            if (y->mpi) y->g->pending.push_back(std::make_pair(dst[i], xs[src[i]]));
            else yd[dst[i]] += xs[src[i]];
        } else {
            yd[dst[i]] = xs[src[i]];
        }
    }
    return 0;
}
PetscErrorCode VecScatterEnd(VecScatter, Vec, Vec, InsertMode, ScatterMode) { return 0; }
PetscErrorCode VecScatterDestroy(VecScatter* sc) { delete *sc; *sc = NULL; return 0; }

PetscErrorCode MatCreateShell(MPI_Comm, PetscInt, PetscInt, PetscInt, PetscInt, void* ctx, Mat* A) {
    *A = new _mimsem_Mat;
    (*A)->ctx = ctx;
    (*A)->mult = NULL;
    (*A)->destroy = NULL;
    return 0;
}
PetscErrorCode MatShellSetOperation(Mat A, MatOperation op, void (*f)(void)) {
    if (op == MATOP_MULT) A->mult = (PetscErrorCode(*)(Mat, Vec, Vec))f;
    else if (op == MATOP_DESTROY) A->destroy = (PetscErrorCode(*)(Mat))f;
    else return 56;   /* PETSC_ERR_SUP */
    return 0;
}
PetscErrorCode MatShellGetContext(Mat A, void* ctx) { *(void**)ctx = A->ctx; return 0; }
PetscErrorCode MatMult(Mat A, Vec x, Vec y) { return A->mult ? A->mult(A, x, y) : 56; }
PetscErrorCode MatDestroy(Mat* A) {
    if (*A) {
        if ((*A)->destroy) (*A)->destroy(*A);
        delete *A;
    }
    *A = NULL;
    return 0;
}
#endif
