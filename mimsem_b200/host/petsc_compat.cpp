// See petsc_compat.h.  Compiled only when PETSc is absent.
#ifndef MIMSEM_HAVE_PETSC
#include "petsc_compat.h"

#include <cmath>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

namespace {
int g_rank = 0, g_size = 1;

struct GlobalVec {            // one collective VecCreateMPI across all in-process ranks
    std::vector<double> a;    // the whole vector
    std::vector<int> nlocal;  // owned size per rank (-1: that rank has not created it yet)
    int refs = 0;
    int created = 0;          // ranks that have made this collective call so far (storage lives until all have come and gone)
    long zero_calls = 0;      // VecZeroEntries is collective: the first rank of a round clears the WHOLE vector
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
    PetscErrorCode (*getdiag)(Mat, Vec);
    PetscErrorCode (*pcbjacobi)(Mat, Vec, Vec);
    PetscErrorCode (*axpy)(Mat, PetscScalar, Mat, MatStructure);
};
struct _mimsem_PC {
    const char* type;
    PetscErrorCode (*apply)(PC, Vec, Vec);
    void* ctx;
    int blocks;
};
struct _mimsem_KSP {
    Mat A, P;
    double rtol, abstol, rnorm;
    int maxits, its;
    bool cg;
    _mimsem_PC pc;
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
    g->created++;
    Vec x = new _mimsem_Vec;
    x->mpi = true; x->n = n; x->N = N; x->rank = g_rank; x->g = g; x->seq = seq;
    *v = x;
    return 0;
}
PetscErrorCode VecDestroy(Vec* v) {
    if (*v) {
        // a temporary created and destroyed inside a collective routine (Geom::write*, ...) is visited by the in-process
        // ranks one after the other: its storage must outlive the ranks that are already done with it
        if ((*v)->mpi && --(*v)->g->refs == 0 && (*v)->g->created >= g_size) {
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
    if (v->mpi) {
        // Under MPI every rank clears its slice BEFORE any rank's contribution of the next phase arrives.  Played rank
        // after rank, a later rank's clear would wipe what earlier ranks have already added or inserted into its slice:
        // the first rank of each round therefore clears the whole vector and the other ranks' calls are no-ops.
        GlobalVec* g = v->g;
        if (g->zero_calls % g_size == 0) {
            g->pending.clear();
            std::fill(g->a.begin(), g->a.end(), 0.0);
        }
        g->zero_calls++;
    } else {
        std::fill(v->local.begin(), v->local.end(), 0.0);
    }
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
    (*A)->getdiag = NULL;
    (*A)->pcbjacobi = NULL;
    (*A)->axpy = NULL;
    return 0;
}
PetscErrorCode MatShellSetOperation(Mat A, MatOperation op, void (*f)(void)) {
    if (op == MATOP_MULT) A->mult = (PetscErrorCode(*)(Mat, Vec, Vec))f;
    else if (op == MATOP_DESTROY) A->destroy = (PetscErrorCode(*)(Mat))f;
    else if (op == MATOP_GET_DIAGONAL) A->getdiag = (PetscErrorCode(*)(Mat, Vec))f;
    else if (op == MATOP_COMPAT_PCBJACOBI) A->pcbjacobi = (PetscErrorCode(*)(Mat, Vec, Vec))f;
    else if (op == MATOP_AXPY) A->axpy = (PetscErrorCode(*)(Mat, PetscScalar, Mat, MatStructure))f;
    else return 56;   /* PETSC_ERR_SUP */
    return 0;
}
PetscErrorCode MatShellGetContext(Mat A, void* ctx) { *(void**)ctx = A->ctx; return 0; }
PetscErrorCode MatMult(Mat A, Vec x, Vec y) { return A->mult ? A->mult(A, x, y) : 56; }
PetscErrorCode MatAXPY(Mat Y, PetscScalar a, Mat X, MatStructure str) { return Y->axpy ? Y->axpy(Y, a, X, str) : 56; }
PetscErrorCode MatAssemblyBegin(Mat, MatAssemblyType) { return 0; }
PetscErrorCode MatAssemblyEnd(Mat, MatAssemblyType) { return 0; }
PetscErrorCode MatDestroy(Mat* A) {
    if (*A) {
        if ((*A)->destroy) (*A)->destroy(*A);
        delete *A;
    }
    *A = NULL;
    return 0;
}
PetscErrorCode MatGetDiagonal(Mat A, Vec d) { return A->getdiag ? A->getdiag(A, d) : 56; }

// ---------------------------------------------------------------------------------------------- vector algebra
static double* own_of(Vec v) { return v->mpi ? v->g->a.data() + rstart_of(v) : v->local.data(); }
static void settle(Vec v) { if (v->mpi) v->g->flush(); }
PetscErrorCode VecSet(Vec v, PetscScalar a) {
    settle(v);
    double* p = own_of(v);
    for (int i = 0; i < v->n; i++) p[i] = a;
    return 0;
}
PetscErrorCode VecCopy(Vec x, Vec y) {
    settle(x); settle(y);
    std::memcpy(own_of(y), own_of(x), sizeof(double) * x->n);
    return 0;
}
PetscErrorCode VecDuplicate(Vec x, Vec* y) { return x->mpi ? VecCreateMPI(MPI_COMM_WORLD, x->n, x->N, y) : VecCreateSeq(MPI_COMM_SELF, x->n, y); }
PetscErrorCode VecScale(Vec v, PetscScalar a) {
    settle(v);
    double* p = own_of(v);
    for (int i = 0; i < v->n; i++) p[i] *= a;
    return 0;
}
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x) {
    settle(x); settle(y);
    double* py = own_of(y);
    const double* px = own_of(x);
    for (int i = 0; i < y->n; i++) py[i] += a * px[i];
    return 0;
}
PetscErrorCode VecAYPX(Vec y, PetscScalar a, Vec x) {
    settle(x); settle(y);
    double* py = own_of(y);
    const double* px = own_of(x);
    for (int i = 0; i < y->n; i++) py[i] = px[i] + a * py[i];
    return 0;
}
PetscErrorCode VecPointwiseMult(Vec w, Vec x, Vec y) {
    settle(x); settle(y); settle(w);
    double* pw = own_of(w);
    const double *px = own_of(x), *py = own_of(y);
    for (int i = 0; i < w->n; i++) pw[i] = px[i] * py[i];
    return 0;
}
PetscErrorCode VecPointwiseDivide(Vec w, Vec x, Vec y) {
    settle(x); settle(y); settle(w);
    double* pw = own_of(w);
    const double *px = own_of(x), *py = own_of(y);
    for (int i = 0; i < w->n; i++) pw[i] = px[i] / py[i];
    return 0;
}
// reductions are collective: over the whole vector (every in-process rank has finished its phase when one of them asks)
PetscErrorCode VecDot(Vec x, Vec y, PetscScalar* val) {
    settle(x); settle(y);
    const double *px = x->mpi ? x->g->a.data() : x->local.data(), *py = y->mpi ? y->g->a.data() : y->local.data();
    double s = 0.0;
    for (int i = 0; i < x->N; i++) s += px[i] * py[i];
    *val = s;
    return 0;
}
PetscErrorCode VecNorm(Vec x, NormType t, PetscReal* val) {
    settle(x);
    const double* px = x->mpi ? x->g->a.data() : x->local.data();
    double s = 0.0;
    for (int i = 0; i < x->N; i++) {
        const double a = std::fabs(px[i]);
        if (t == NORM_1) s += a;
        else if (t == NORM_INFINITY) s = a > s ? a : s;
        else s += a * a;
    }
    *val = (t == NORM_1 || t == NORM_INFINITY) ? s : std::sqrt(s);
    return 0;
}

// ---------------------------------------------------------------------------------------------- viewers
struct _mimsem_Viewer {
    std::string name;
    bool binary, write;
};
PetscErrorCode PetscViewerBinaryOpen(MPI_Comm, const char* name, PetscFileMode mode, PetscViewer* viewer) {
    *viewer = new _mimsem_Viewer{name, true, mode == FILE_MODE_WRITE};
    return 0;
}
PetscErrorCode PetscViewerASCIIOpen(MPI_Comm, const char* name, PetscViewer* viewer) {
    *viewer = new _mimsem_Viewer{name, false, true};
    return 0;
}
PetscErrorCode PetscViewerDestroy(PetscViewer* viewer) { delete *viewer; *viewer = NULL; return 0; }
static void put_be32(FILE* f, unsigned v) {
    unsigned char b[4] = {(unsigned char)(v >> 24), (unsigned char)(v >> 16), (unsigned char)(v >> 8), (unsigned char)v};
    std::fwrite(b, 1, 4, f);
}
static void put_be64(FILE* f, double d) {
    unsigned long long v;
    std::memcpy(&v, &d, 8);
    unsigned char b[8];
    for (int i = 0; i < 8; i++) b[i] = (unsigned char)(v >> (56 - 8 * i));
    std::fwrite(b, 1, 8, f);
}
PetscErrorCode VecView(Vec v, PetscViewer viewer) {
    settle(v);
    if (v->mpi && g_rank != g_size - 1) return 0;   // collective: the last rank to arrive writes the whole vector
    const double* a = v->mpi ? v->g->a.data() : v->local.data();
    FILE* f = std::fopen(viewer->name.c_str(), viewer->binary ? "wb" : "w");
    if (!f) return 65;   /* PETSC_ERR_FILE_OPEN */
    if (viewer->binary) {
        put_be32(f, 1211214u);   /* VEC_FILE_CLASSID */
        put_be32(f, (unsigned)v->N);
        for (int i = 0; i < v->N; i++) put_be64(f, a[i]);
    } else {
        std::fprintf(f, "Vec Object: %d MPI processes\n  type: %s\n", g_size, v->mpi ? "mpi" : "seq");
        int lo = 0;
        for (int r = 0; r < (v->mpi ? g_size : 1); r++) {
            const int n = v->mpi ? v->g->nlocal[r] : v->n;
            if (v->mpi) std::fprintf(f, "Process [%d]\n", r);
            for (int i = 0; i < n; i++) std::fprintf(f, "%.16g\n", a[lo + i]);
            lo += n;
        }
    }
    std::fclose(f);
    return 0;
}
PetscErrorCode VecLoad(Vec v, PetscViewer viewer) {
    FILE* f = std::fopen(viewer->name.c_str(), "rb");
    if (!f) return 65;
    unsigned char h[8];
    if (std::fread(h, 1, 8, f) != 8) { std::fclose(f); return 66; }
    const unsigned cid = ((unsigned)h[0] << 24) | ((unsigned)h[1] << 16) | ((unsigned)h[2] << 8) | h[3];
    const unsigned N = ((unsigned)h[4] << 24) | ((unsigned)h[5] << 16) | ((unsigned)h[6] << 8) | h[7];
    if (cid != 1211214u || (int)N != v->N) { std::fclose(f); return 79; }   /* PETSC_ERR_FILE_UNEXPECTED */
    const int lo = v->mpi ? rstart_of(v) : 0;
    std::fseek(f, 8 + 8L * lo, SEEK_SET);
    double* dst = own_of(v);
    for (int i = 0; i < v->n; i++) {
        unsigned char b[8];
        if (std::fread(b, 1, 8, f) != 8) { std::fclose(f); return 66; }
        unsigned long long u = 0;
        for (int j = 0; j < 8; j++) u = (u << 8) | b[j];
        std::memcpy(&dst[i], &u, 8);
    }
    std::fclose(f);
    return 0;
}

// ---------------------------------------------------------------------------------------------- KSP (one rank)
PetscErrorCode KSPCreate(MPI_Comm, KSP* ksp) {
    KSP k = new _mimsem_KSP;
    k->A = k->P = NULL;
    k->rtol = 1.0e-5;   // PETSc's defaults
    k->abstol = 1.0e-50;
    k->maxits = 10000;
    k->its = 0;
    k->rnorm = 0.0;
    k->cg = false;      // KSPGMRES is PETSc's default type
    k->pc.type = PCJACOBI;
    k->pc.apply = NULL;
    k->pc.ctx = NULL;
    k->pc.blocks = 0;
    *ksp = k;
    return 0;
}
PetscErrorCode KSPDestroy(KSP* ksp) { delete *ksp; *ksp = NULL; return 0; }
PetscErrorCode KSPSetOperators(KSP ksp, Mat A, Mat P) { ksp->A = A; ksp->P = P; return 0; }
PetscErrorCode KSPSetTolerances(KSP ksp, PetscReal rtol, PetscReal abstol, PetscReal, PetscInt maxits) {
    if (rtol != PETSC_DEFAULT) ksp->rtol = rtol;
    if (abstol != PETSC_DEFAULT) ksp->abstol = abstol;
    if (maxits != PETSC_DEFAULT) ksp->maxits = maxits;
    return 0;
}
PetscErrorCode KSPSetType(KSP ksp, KSPType type) { ksp->cg = type && std::strcmp(type, KSPCG) == 0; return 0; }
PetscErrorCode KSPSetOptionsPrefix(KSP, const char*) { return 0; }
PetscErrorCode KSPSetFromOptions(KSP) { return 0; }
PetscErrorCode KSPGetPC(KSP ksp, PC* pc) { *pc = &ksp->pc; return 0; }
PetscErrorCode KSPGetIterationNumber(KSP ksp, PetscInt* its) { *its = ksp->its; return 0; }
PetscErrorCode KSPGetResidualNorm(KSP ksp, PetscReal* rnorm) { *rnorm = ksp->rnorm; return 0; }
PetscErrorCode PCSetType(PC pc, PCType type) { pc->type = type; return 0; }
PetscErrorCode PCBJacobiSetTotalBlocks(PC pc, PetscInt blocks, const PetscInt*) { pc->blocks = blocks; return 0; }
PetscErrorCode PCShellSetApply(PC pc, PetscErrorCode (*apply)(PC, Vec, Vec)) { pc->apply = apply; return 0; }
PetscErrorCode PCApply(PC pc, Vec x, Vec y) { return pc->apply ? pc->apply(pc, x, y) : 56; }
PetscErrorCode PCShellSetContext(PC pc, void* ctx) { pc->ctx = ctx; return 0; }
PetscErrorCode PCShellGetContext(PC pc, void* ctx) { *(void**)ctx = pc->ctx; return 0; }

namespace {

// the preconditioner of a solve: see the comment at KSPCreate in petsc_compat.h
struct Precond {
    KSP ksp;
    Vec d;
    int kind;   // 0 identity, 1 diagonal, 2 shell callback, 3 the operator's own element blocks
    Precond(KSP k, Vec like) : ksp(k), d(NULL), kind(0), like_(like) {
        const char* t = k->pc.type ? k->pc.type : PCJACOBI;
        Mat P = k->P ? k->P : k->A;
        if (k->pc.apply && std::strcmp(t, PCSHELL) == 0) kind = 2;
        else if (std::strcmp(t, PCNONE) == 0) kind = 0;
        else if (std::strcmp(t, PCBJACOBI) == 0 && P->pcbjacobi) kind = 3;
        else use_diagonal();
    }
    ~Precond() { if (d) VecDestroy(&d); }
    void use_diagonal() {
        Mat P = ksp->P ? ksp->P : ksp->A;
        kind = 0;
        if (!P->getdiag) return;
        VecDuplicate(like_, &d);
        if (MatGetDiagonal(P, d) == 0) kind = 1;
    }
    void operator()(Vec in, Vec out) {
        Mat P = ksp->P ? ksp->P : ksp->A;
        // the shell says with a nonzero code that it has no blocks for this operator: its first answer decides (the
        // first application opens the solve), the diagonal takes over
        if (kind == 3 && P->pcbjacobi(P, in, out) != 0) use_diagonal();
        if (kind == 3) return;
        if (kind == 2) ksp->pc.apply(&ksp->pc, in, out);
        else if (kind == 1) VecPointwiseDivide(out, in, d);
        else VecCopy(in, out);
    }
    Vec like_;
};

// preconditioned conjugate gradients (KSPCG)
void solve_cg(KSP ksp, Precond& B, Vec b, Vec x) {
    Vec r, z, p, q;
    VecDuplicate(b, &r); VecDuplicate(b, &z); VecDuplicate(b, &p); VecDuplicate(b, &q);
    double bn, rn, rz, rz_new, pq;
    VecNorm(b, NORM_2, &bn);
    MatMult(ksp->A, x, q);             // nonzero initial guess, as PETSc's KSPSolve with the caller's x
    VecCopy(b, r);
    VecAXPY(r, -1.0, q);
    B(r, z);
    VecCopy(z, p);
    VecDot(r, z, &rz);
    VecNorm(r, NORM_2, &rn);
    while (rn > ksp->rtol * bn && rn > ksp->abstol && ksp->its < ksp->maxits) {
        MatMult(ksp->A, p, q);
        VecDot(p, q, &pq);
        const double alpha = rz / pq;
        VecAXPY(x, alpha, p);
        VecAXPY(r, -alpha, q);
        B(r, z);
        VecDot(r, z, &rz_new);
        VecAYPX(p, rz_new / rz, z);
        rz = rz_new;
        VecNorm(r, NORM_2, &rn);
        ksp->its++;
    }
    ksp->rnorm = rn;
    VecDestroy(&r); VecDestroy(&z); VecDestroy(&p); VecDestroy(&q);
}

// restarted GMRES(30), left preconditioning: minimises |B (b - A x)| over x0 + K_m(B A, B r0)
void solve_gmres(KSP ksp, Precond& B, Vec b, Vec x) {
    const int m = 30;
    std::vector<Vec> V(m + 1);
    for (int i = 0; i <= m; i++) VecDuplicate(b, &V[i]);
    Vec w, t;
    VecDuplicate(b, &w); VecDuplicate(b, &t);
    std::vector<double> H((size_t)(m + 1) * m), cs(m), sn(m), g(m + 1), y(m);
    B(b, w);
    double bnorm;
    VecNorm(w, NORM_2, &bnorm);
    const double ttol = std::max(ksp->rtol * bnorm, ksp->abstol);
    bool done = false;
    while (!done) {
        MatMult(ksp->A, x, t);
        VecAYPX(t, -1.0, b);           // t = b - A x
        B(t, V[0]);
        double beta;
        VecNorm(V[0], NORM_2, &beta);
        ksp->rnorm = beta;
        if (beta <= ttol || ksp->its >= ksp->maxits) break;
        VecScale(V[0], 1.0 / beta);
        std::fill(g.begin(), g.end(), 0.0);
        g[0] = beta;
        int k = 0;
        for (; k < m && ksp->its < ksp->maxits; k++) {
            MatMult(ksp->A, V[k], t);
            B(t, w);
            for (int i = 0; i <= k; i++) {
                double h;
                VecDot(w, V[i], &h);
                H[(size_t)i * m + k] = h;
                VecAXPY(w, -h, V[i]);
            }
            double hn;
            VecNorm(w, NORM_2, &hn);
            H[(size_t)(k + 1) * m + k] = hn;
            for (int i = 0; i < k; i++) {   // the rotations so far, then a new one that removes the subdiagonal entry
                const double a = H[(size_t)i * m + k], c = H[(size_t)(i + 1) * m + k];
                H[(size_t)i * m + k] = cs[i] * a + sn[i] * c;
                H[(size_t)(i + 1) * m + k] = -sn[i] * a + cs[i] * c;
            }
            const double a = H[(size_t)k * m + k], rho = std::sqrt(a * a + hn * hn);
            cs[k] = rho > 0.0 ? a / rho : 1.0;
            sn[k] = rho > 0.0 ? hn / rho : 0.0;
            H[(size_t)k * m + k] = rho;
            H[(size_t)(k + 1) * m + k] = 0.0;
            g[k + 1] = -sn[k] * g[k];
            g[k] = cs[k] * g[k];
            ksp->its++;
            ksp->rnorm = std::fabs(g[k + 1]);
            if (hn > 0.0) {
                VecCopy(w, V[k + 1]);
                VecScale(V[k + 1], 1.0 / hn);
            }
            if (ksp->rnorm <= ttol || hn == 0.0) {
                done = true;
                k++;
                break;
            }
        }
        for (int i = k - 1; i >= 0; i--) {   // back substitution, x += V y
            double s = g[i];
            for (int j = i + 1; j < k; j++) s -= H[(size_t)i * m + j] * y[j];
            y[i] = s / H[(size_t)i * m + i];
        }
        for (int i = 0; i < k; i++) VecAXPY(x, y[i], V[i]);
        if (ksp->its >= ksp->maxits) break;
    }
    for (int i = 0; i <= m; i++) VecDestroy(&V[i]);
    VecDestroy(&w); VecDestroy(&t);
}

}  // namespace

PetscErrorCode KSPSolve(KSP ksp, Vec b, Vec x) {
    if (g_size != 1) {
        std::fprintf(stderr, "petsc_compat: KSPSolve needs real PETSc when more than one rank is played in-process\n");
        std::abort();
    }
    ksp->its = 0;
    Precond B(ksp, b);
    if (ksp->cg) solve_cg(ksp, B, b, x);
    else solve_gmres(ksp, B, b, x);
    return 0;
}
#endif
