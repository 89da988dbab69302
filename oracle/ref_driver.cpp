/*
 * ORACLE-ONLY driver around the UNMODIFIED reference sources (TEST INFRASTRUCTURE).
 *
 * Compiled three times (-DREF_EUL / -DREF_SRC / -DREF_BOX) together with
 *   /root/reference/<dir>/{Basis,LinAlg,ElMats,Topo,Geom,Assembly}.cpp   (where they lie; never copied)
 * and oracle/shim/petsc_shim.cpp into oracle/_ref/libref_<dir>.so by oracle/Makefile.
 *
 * What it does, per SURVEY.md section 8c (oracle O1): for every emulated MPI rank it constructs the
 * reference's own Topo / Geom / operator objects, calls the reference's own assemble(...), takes
 * the MatSetValues triplets the shim recorded (global indices), merges all ranks into one CSR
 * matrix (INSERT / ADD semantics honoured, summation in rank-then-insertion order) and applies it
 * with a plain CSR SpMV -- the algorithm PETSc's MatMult_SeqAIJ/MPIAIJ implements.
 *
 * Exposed as a C ABI for ctypes (oracle/refbind.py).  Nothing here is product code.
 */
#include <unistd.h>
#include <iostream>
#include <thread>
#include <vector>
#include <algorithm>
#include <chrono>
#include <mutex>

#include <petsc.h>
#include "LinAlg.h"
#include "Basis.h"
#include "Topo.h"
#include "Geom.h"
#include "ElMats.h"
#include "Assembly.h"

namespace {

double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct RankObjs {
    Topo* topo;
    Geom* geom;
    GaussLobatto* quad;
    LagrangeNode* node;
    LagrangeEdge* edge;
};

struct Ref {
    int nranks, nk;
    std::vector<RankObjs> r;
    /* merged CSR of the last assembled operator */
    long nrows, ncols;
    std::vector<long> indptr;
    std::vector<int> indices;
    std::vector<double> data;
};

template <class F>
void for_ranks(Ref* h, int nthreads, F f) {
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> th;
    std::mutex mu;
    int next = 0;
    for (int t = 0; t < std::min(nthreads, h->nranks); t++) {
        th.emplace_back([&]() {
            for (;;) {
                int rk;
                {
                    std::lock_guard<std::mutex> g(mu);
                    rk = next++;
                }
                if (rk >= h->nranks) break;
                ShimSetRank(rk, h->nranks);
                f(rk);
            }
        });
    }
    for (auto& t : th) t.join();
}

/* merge triplet lists (rank order, insertion order) into CSR */
void merge_csr(Ref* h, long nrows, long ncols, const std::vector<std::vector<ShimTriplet>*>& lists) {
    h->nrows = nrows;
    h->ncols = ncols;
    std::vector<long> cnt(nrows + 1, 0);
    for (auto* l : lists)
        for (const auto& t : *l) cnt[t.row + 1]++;
    for (long i = 0; i < nrows; i++) cnt[i + 1] += cnt[i];
    long total = cnt[nrows];
    std::vector<ShimTriplet> b(total);
    {
        std::vector<long> pos(cnt.begin(), cnt.end() - 1);
        for (auto* l : lists)
            for (const auto& t : *l) b[pos[t.row]++] = t;
    }
    h->indptr.assign(nrows + 1, 0);
    h->indices.clear();
    h->data.clear();
    h->indices.reserve(total / 2);
    h->data.reserve(total / 2);
    for (long i = 0; i < nrows; i++) {
        ShimTriplet* lo = b.data() + cnt[i];
        ShimTriplet* hi = b.data() + cnt[i + 1];
        std::stable_sort(lo, hi, [](const ShimTriplet& a, const ShimTriplet& c) { return a.col < c.col; });
        for (ShimTriplet* p = lo; p < hi;) {
            int col = p->col;
            double v = 0.0;
            for (; p < hi && p->col == col; p++) {
                if (p->mode == INSERT_VALUES) v = p->val;
                else v += p->val;
            }
            h->indices.push_back(col);
            h->data.push_back(v);
        }
        h->indptr[i + 1] = (long)h->indices.size();
    }
}

}  // namespace

extern "C" {

/* mesh_dir must contain input/*.txt as written by the reference's scr/Setup*.py */
void* ref_open(const char* mesh_dir, int nranks, int nk, int nthreads) {
    std::cout.setstate(std::ios_base::failbit); /* the reference prints per-file chatter */
    if (chdir(mesh_dir) != 0) return NULL;
    Ref* h = new Ref;
    h->nranks = nranks;
    h->nk = nk;
    h->r.resize(nranks);
    h->nrows = h->ncols = 0;
    for_ranks(h, nthreads, [&](int rk) {
        RankObjs& o = h->r[rk];
#if defined(REF_EUL)
        o.topo = new Topo(nk);
        o.geom = new Geom(o.topo, nk);
#elif defined(REF_BOX)
        o.topo = new Topo();
        o.geom = new Geom(o.topo, nk);
#else
        o.topo = new Topo();
        o.geom = new Geom(o.topo);
#endif
        /* same construction order as the reference's solver constructors
         * (eul/HorizSolve.cpp:40-42, src/SWEqn_Picard.cpp:52-54) */
        o.quad = new GaussLobatto(o.geom->quad->n);
        o.node = new LagrangeNode(o.topo->elOrd, o.quad);
        o.edge = new LagrangeEdge(o.topo->elOrd, o.node);
    });
    return h;
}

void ref_close(void* hv) {
    Ref* h = (Ref*)hv;
    for (int rk = 0; rk < h->nranks; rk++) {
        ShimSetRank(rk, h->nranks);
        RankObjs& o = h->r[rk];
        delete o.edge;
        delete o.node;
        delete o.quad;
        delete o.geom;
        delete o.topo;
    }
    delete h;
}

/* out[0..11] = elOrd, nElsX, quadOrd, n0, n1x, n1y, n2, n0l, n1l, n2l, nDofs0G, nDofs1G ; out[12] = nDofs2G ; out[13] = geom n0 */
void ref_info(void* hv, int rank, int* out) {
    Ref* h = (Ref*)hv;
    Topo* t = h->r[rank].topo;
    Geom* g = h->r[rank].geom;
    out[0] = t->elOrd; out[1] = t->nElsX; out[2] = g->quad->n;
    out[3] = t->n0; out[4] = t->n1x; out[5] = t->n1y; out[6] = t->n2;
    out[7] = t->n0l; out[8] = t->n1l; out[9] = t->n2l;
    out[10] = t->nDofs0G; out[11] = t->nDofs1G; out[12] = t->nDofs2G;
#if defined(REF_BOX)
    out[13] = t->n0;
#else
    out[13] = g->n0;
#endif
}

/* which: 0 loc0, 1 loc1x, 2 loc1y, 3 loc2, 4 loc1 (interleaved), 5 geom quad-point loc0 */
void ref_get_loc(void* hv, int rank, int which, int* out) {
    Ref* h = (Ref*)hv;
    Topo* t = h->r[rank].topo;
    const int* src = NULL;
    int n = 0;
    switch (which) {
        case 0: src = t->loc0; n = t->n0; break;
        case 1: src = t->loc1x; n = t->n1x; break;
        case 2: src = t->loc1y; n = t->n1y; break;
        case 3: src = t->loc2; n = t->n2; break;
        case 4: src = t->loc1; n = t->n1; break;
#if !defined(REF_BOX)
        case 5: src = h->r[rank].geom->loc0; n = h->r[rank].geom->n0; break;
#endif
    }
    for (int i = 0; i < n; i++) out[i] = src[i];
}

/* det[nel][mp12], J[nel][mp12][2][2] of one rank, as the reference computed them */
void ref_get_geom(void* hv, int rank, double* det, double* J) {
    Ref* h = (Ref*)hv;
    Topo* t = h->r[rank].topo;
    Geom* g = h->r[rank].geom;
    int nel = t->nElsX * t->nElsX;
    int mp12 = (g->quad->n + 1) * (g->quad->n + 1);
    for (int e = 0; e < nel; e++)
        for (int q = 0; q < mp12; q++) {
            det[e * mp12 + q] = g->det[e][q];
            for (int a = 0; a < 2; a++)
                for (int b = 0; b < 2; b++) J[((e * mp12 + q) * 2 + a) * 2 + b] = g->J[e][q][a][b];
        }
}

/* cartesian coordinates x[nl][3] of the rank's quadrature points (after the reference's own fix-ups) */
void ref_get_coords(void* hv, int rank, double* x) {
    Ref* h = (Ref*)hv;
    Geom* g = h->r[rank].geom;
    for (int i = 0; i < g->nl; i++)
        for (int c = 0; c < 3; c++) x[i * 3 + c] = g->x[i][c];
}

int ref_geom_nl(void* hv, int rank) { return ((Ref*)hv)->r[rank].geom->nl; }

/* basis tables of rank 0: gll x[m+1], w[m+1], ljxi[(m+1)*(n+1)], ejxi[(m+1)*n] */
void ref_get_basis(void* hv, double* x, double* w, double* ljxi, double* ejxi) {
    Ref* h = (Ref*)hv;
    RankObjs& o = h->r[0];
    int m = o.quad->n, n = o.node->n;
    for (int i = 0; i <= m; i++) {
        x[i] = o.quad->x[i];
        w[i] = o.quad->w[i];
        for (int j = 0; j <= n; j++) ljxi[i * (n + 1) + j] = o.node->ljxi[i][j];
        for (int j = 0; j < n; j++) ejxi[i * n + j] = o.edge->ejxi[i][j];
    }
}

#if !defined(REF_SRC)
/* thick[nk][n0q] in the rank's local quadrature-point (eul) / node (box) numbering */
void ref_set_thick(void* hv, int rank, const double* thick) {
    Ref* h = (Ref*)hv;
    Geom* g = h->r[rank].geom;
#if defined(REF_EUL)
    int n0 = g->n0;
#else
    int n0 = g->topo->n0;
#endif
    for (int k = 0; k < h->nk; k++)
        for (int i = 0; i < n0; i++) {
            g->thick[k][i] = thick[(long)k * n0 + i];
#if defined(REF_EUL)
            g->thickInv[k][i] = 1.0 / thick[(long)k * n0 + i];
#endif
        }
}
#endif

enum {
    OP_UMAT = 0, OP_WMAT = 1, OP_PMAT = 2, OP_UHMAT = 3, OP_WTQUMAT = 4,
    OP_E10 = 5, OP_E01 = 6, OP_E21 = 7, OP_E12 = 8, OP_PMAT_H = 9, OP_WHMAT = 10, OP_ROTMAT = 11,
    OP_PHMAT_UP = 12, OP_ROTMAT_UP = 13,
    /* eul/: operators of the horizontal-vorticity / vertical-momentum terms */
    OP_UT_MAT = 14, OP_UT_MAT_H = 15, OP_UTQWMAT = 16, OP_WTQDUDZ = 17,
    /* eul/: Rayleigh friction; c2 = Exner 2-form of the level, c1 = the LEVEL-0 Exner 2-form (a second global 2-form) */
    OP_UMAT_RAY = 18,
    /* eul/: quadrature-point projections of the initial conditions (eul/Euler_2.cpp:432, 493, 535; eul/Assembly.cpp WtQmat, UtQmat,
       PtQmat): columns = quadrature points of Geom (UtQmat: two interleaved components per point) */
    OP_WTQMAT = 19, OP_UTQMAT = 20, OP_PTQMAT = 21
};

/*
 * Run the reference's constructor + assemble for operator `op` on every emulated rank and merge.
 *   flag   : vert_scale (Umat/Wmat) / const_vert (Uhmat) / vert_scale_rho (Whmat)
 *   c2     : global 2-form coefficient vector (Uhmat, Pmat_h, Whmat) or NULL
 *   c1     : global 1-form coefficient vector (WtQUmat) or NULL
 *   c0     : global 0-form coefficient vector (RotMat) or NULL
 * secs[0] = wall time of the per-rank constructor+assemble phase, secs[1] = merge-to-CSR time.
 * Returns nnz, or -1 for an operator this variant does not have.
 */
long ref_assemble(void* hv, int op, int lev, double scale, int flag, const double* c2, const double* c1,
                  const double* c0, double tau, double dt, int nthreads, double* secs) {
    Ref* h = (Ref*)hv;
    std::vector<Mat> mats(h->nranks, (Mat)NULL);
    std::vector<std::vector<ShimTriplet> > keep(h->nranks);
    long nrows = 0, ncols = 0;
    bool bad = false;
    double t0 = now();
    for_ranks(h, nthreads, [&](int rk) {
        RankObjs& o = h->r[rk];
        Topo* topo = o.topo;
        Geom* geom = o.geom;
        Vec v2 = NULL, v1 = NULL, v0 = NULL;
        if (c2) {
            VecCreateMPI(MPI_COMM_WORLD, topo->n2l, topo->nDofs2G, &v2);
            /* faces: global = local + pi*n2 (reference Topo::elInds2_g) */
            for (int i = 0; i < topo->n2l; i++) v2->a[i] = c2[(long)rk * topo->n2 + i];
        }
        Vec v2b = NULL;
        if (c1 && op == OP_UMAT_RAY) {
            VecCreateMPI(MPI_COMM_WORLD, topo->n2l, topo->nDofs2G, &v2b);
            for (int i = 0; i < topo->n2l; i++) v2b->a[i] = c1[(long)rk * topo->n2 + i];
        } else if (c1) {
            VecCreateSeq(MPI_COMM_SELF, topo->n1, &v1);
            for (int i = 0; i < topo->n1; i++) v1->a[i] = c1[topo->loc1[i]];
        }
        if (c0) {
            VecCreateSeq(MPI_COMM_SELF, topo->n0, &v0);
            for (int i = 0; i < topo->n0; i++) v0->a[i] = c0[topo->loc0[i]];
        }
        Mat M = NULL;
        switch (op) {
#if defined(REF_EUL)
            case OP_UMAT: {
                Umat* A = new Umat(topo, geom, o.node, o.edge);
                A->assemble(lev, scale, flag != 0);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_WMAT: {
                Wmat* A = new Wmat(topo, geom, o.edge);
                A->assemble(lev, scale, flag != 0);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_PMAT: {
                Pmat* A = new Pmat(topo, geom, o.node);
                A->assemble(lev, scale);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_PMAT_H: {
                Pmat* A = new Pmat(topo, geom, o.node);
                A->assemble_h(lev, scale, v2);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_UHMAT: {
                Uhmat* A = new Uhmat(topo, geom, o.node, o.edge);
                A->assemble(v2, lev, flag != 0, scale);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_WTQUMAT: {
                WtQUmat* A = new WtQUmat(topo, geom, o.node, o.edge);
                A->assemble(v1, lev, scale);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_WHMAT: {
                Whmat* A = new Whmat(topo, geom, o.edge);
                A->assemble(v2, lev, scale, flag != 0);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_ROTMAT: {
                RotMat* A = new RotMat(topo, geom, o.node, o.edge);
                A->assemble(v0, lev, scale);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_UT_MAT: {
                Ut_mat* A = new Ut_mat(topo, geom, o.node, o.edge);
                A->assemble(lev, scale);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_UT_MAT_H: {
                Ut_mat* A = new Ut_mat(topo, geom, o.node, o.edge);
                A->assemble_h(lev, scale, v2);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_UTQWMAT: {
                UtQWmat* A = new UtQWmat(topo, geom, o.node, o.edge);
                A->assemble(v1, scale);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_WTQDUDZ: {
                WtQdUdz_mat* A = new WtQdUdz_mat(topo, geom, o.node, o.edge);
                A->assemble(v1, scale);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_UMAT_RAY: {
                Umat_ray* A = new Umat_ray(topo, geom, o.node, o.edge);
                A->assemble(lev, scale, dt, v2, v2b);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_WTQMAT: {
                WtQmat* A = new WtQmat(topo, geom, o.edge); /* ctor assembles */
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_UTQMAT: {
                UtQmat* A = new UtQmat(topo, geom, o.node, o.edge);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_PTQMAT: {
                PtQmat* A = new PtQmat(topo, geom, o.node);
                keep[rk] = A->M->t; delete A; break;
            }
#elif defined(REF_SRC)
            case OP_UMAT: {
                Umat* A = new Umat(topo, geom, o.node, o.edge); /* ctor assembles */
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_WMAT: {
                Wmat* A = new Wmat(topo, geom, o.edge);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_PMAT: {
                Pmat* A = new Pmat(topo, geom, o.node);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_UHMAT: {
                Uhmat* A = new Uhmat(topo, geom, o.node, o.edge);
                A->assemble(v2);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_WTQUMAT: {
                WtQUmat* A = new WtQUmat(topo, geom, o.node, o.edge);
                A->assemble(v1);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_WHMAT: {
                Whmat* A = new Whmat(topo, geom, o.edge);
                A->assemble(v2);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_ROTMAT: {
                RotMat* A = new RotMat(topo, geom, o.node, o.edge);
                A->assemble(v0);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_PHMAT_UP: {
                /* Phmat::assemble_up(ul, hl, fac, dt): ul ghosted local 1-form, hl owned 2-form array */
                Phmat* A = new Phmat(topo, geom, o.node);
                A->assemble_up(v1, v2, tau, dt);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_ROTMAT_UP: {
                RotMat_up* A = new RotMat_up(topo, geom, o.node, o.edge);
                A->assemble(v0, v1, tau, dt);
                keep[rk] = A->M->t; delete A; break;
            }
#else /* REF_BOX */
            case OP_UMAT: {
                /* box: assemble() is private; the ctor builds M (lev 0, vert-scaled) and Mo (unscaled) */
                Umat* A = new Umat(topo, geom, o.node, o.edge);
                keep[rk] = (flag ? A->M : A->Mo)->t; delete A; break;
            }
            case OP_WMAT: {
                Wmat* A = new Wmat(topo, geom, o.edge);
                keep[rk] = (flag ? A->M : A->Mo)->t; delete A; break;
            }
            case OP_UHMAT: {
                Uhmat* A = new Uhmat(topo, geom, o.node, o.edge);
                A->assemble(v2, lev, flag != 0, scale);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_WTQUMAT: {
                WtQUmat* A = new WtQUmat(topo, geom, o.node, o.edge);
                A->assemble(v1, lev, scale);
                keep[rk] = A->M->t; delete A; break;
            }
            case OP_ROTMAT: {
                RotMat* A = new RotMat(topo, geom, o.node, o.edge);
                A->assemble(v0, lev, scale);
                keep[rk] = A->M->t; delete A; break;
            }
#endif
            case OP_E10: case OP_E01: {
                E10mat* A = new E10mat(topo);
                keep[rk] = (op == OP_E10 ? A->E10 : A->E01)->t; delete A; break;
            }
            case OP_E21: case OP_E12: {
                E21mat* A = new E21mat(topo);
                keep[rk] = (op == OP_E21 ? A->E21 : A->E12)->t; delete A; break;
            }
            default: bad = true;
        }
        (void)M;
        if (v2) VecDestroy(&v2);
        if (v1) VecDestroy(&v1);
        if (v2b) VecDestroy(&v2b);
        if (v0) VecDestroy(&v0);
    });
    if (bad) return -1;
    double t1 = now();
    Topo* t = h->r[0].topo;
    switch (op) {
        case OP_UMAT: case OP_UHMAT: case OP_ROTMAT: case OP_ROTMAT_UP: case OP_UT_MAT: case OP_UT_MAT_H: case OP_UMAT_RAY: nrows = ncols = t->nDofs1G; break;
        case OP_WTQDUDZ: nrows = t->nDofs2G; ncols = t->nDofs1G; break;
        case OP_UTQWMAT: nrows = t->nDofs1G; ncols = t->nDofs2G; break;
        case OP_WMAT: case OP_WHMAT: nrows = ncols = t->nDofs2G; break;
        case OP_PMAT: case OP_PMAT_H: case OP_PHMAT_UP: nrows = ncols = t->nDofs0G; break;
        case OP_WTQUMAT: case OP_E21: nrows = t->nDofs2G; ncols = t->nDofs1G; break;
        case OP_E12: nrows = t->nDofs1G; ncols = t->nDofs2G; break;
        case OP_E10: nrows = t->nDofs1G; ncols = t->nDofs0G; break;
        case OP_E01: nrows = t->nDofs0G; ncols = t->nDofs1G; break;
#if defined(REF_EUL)
        case OP_WTQMAT: nrows = t->nDofs2G; ncols = h->r[0].geom->nDofs0G; break;
        case OP_UTQMAT: nrows = t->nDofs1G; ncols = 2L * h->r[0].geom->nDofs0G; break;
        case OP_PTQMAT: nrows = t->nDofs0G; ncols = h->r[0].geom->nDofs0G; break;
#endif
    }
    std::vector<std::vector<ShimTriplet>*> lists;
    for (int rk = 0; rk < h->nranks; rk++) lists.push_back(&keep[rk]);
    merge_csr(h, nrows, ncols, lists);
    double t2 = now();
    if (secs) {
        secs[0] = t1 - t0;
        secs[1] = t2 - t1;
    }
    return (long)h->indices.size();
}

void ref_csr_shape(void* hv, long* out) {
    Ref* h = (Ref*)hv;
    out[0] = h->nrows;
    out[1] = h->ncols;
    out[2] = (long)h->indices.size();
}

void ref_csr_get(void* hv, long* indptr, int* indices, double* data) {
    Ref* h = (Ref*)hv;
    std::copy(h->indptr.begin(), h->indptr.end(), indptr);
    std::copy(h->indices.begin(), h->indices.end(), indices);
    std::copy(h->data.begin(), h->data.end(), data);
}

/* y = A x with the merged CSR (row-parallel over nthreads); returns seconds */
double ref_spmv(void* hv, const double* x, double* y, int nthreads) {
    Ref* h = (Ref*)hv;
    if (nthreads < 1) nthreads = 1;
    double t0 = now();
    std::vector<std::thread> th;
    long chunk = (h->nrows + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        long lo = t * chunk, hi = std::min(h->nrows, lo + chunk);
        th.emplace_back([=]() {
            for (long i = lo; i < hi; i++) {
                double s = 0.0;
                for (long k = h->indptr[i]; k < h->indptr[i + 1]; k++) s += h->data[k] * x[h->indices[k]];
                y[i] = s;
            }
        });
    }
    for (auto& t : th) t.join();
    return now() - t0;
}

#if defined(REF_EUL)
/*
 * The reference's matrix-free twin Uvec::assemble (eul/Assembly.cpp:2124-2196) on every rank:
 * y(global) = sum over ranks of scatter-add of the rank's ghosted local result (what
 * VecScatter(gtol_1, ADD_VALUES, SCATTER_REVERSE) does).  Returns seconds.
 */
double ref_uvec_apply(void* hv, int lev, double scale, int vert_scale, const double* x, double* y, int nthreads) {
    Ref* h = (Ref*)hv;
    Topo* t0p = h->r[0].topo;
    std::vector<std::vector<double> > part(h->nranks);
    double t0 = now();
    for_ranks(h, nthreads, [&](int rk) {
        RankObjs& o = h->r[rk];
        Topo* topo = o.topo;
        Vec v1;
        VecCreateSeq(MPI_COMM_SELF, topo->n1, &v1);
        for (int i = 0; i < topo->n1; i++) v1->a[i] = x[topo->loc1[i]];
        Uvec* A = new Uvec(topo, o.geom, o.node, o.edge);
        A->assemble(lev, scale, vert_scale != 0, v1);
        part[rk].assign(A->vl->a, A->vl->a + topo->n1);
        delete A;
        VecDestroy(&v1);
    });
    for (long i = 0; i < t0p->nDofs1G; i++) y[i] = 0.0;
    for (int rk = 0; rk < h->nranks; rk++) {
        Topo* topo = h->r[rk].topo;
        for (int i = 0; i < topo->n1; i++) y[topo->loc1[i]] += part[rk][i];
    }
    return now() - t0;
}

/*
 * The reference's own matrix-free twin of the flux form, exactly as eul/HorizSolve.cpp:298-306 drives it: per rank
 * Uvec::assemble_hu(lev, scale, u_a, h_a, false, fac_a) for nterms (velocity, density, factor) triples accumulated in
 * the rank's ghosted local vector vl, then the reverse ADD scatter -- merged here over the emulated ranks.
 * u[t] are global 1-form vectors (gathered through loc1), h[t] global 2-form vectors (each rank reads its owned block).
 */
double ref_uvec_assemble_hu(void* hv, int lev, double scale, int nterms, const double* u, const double* h2, const double* fac,
                            double* y, int nthreads) {
    Ref* h = (Ref*)hv;
    Topo* t0p = h->r[0].topo;
    const long N1 = t0p->nDofs1G, N2 = t0p->nDofs2G;
    std::vector<std::vector<double> > part(h->nranks);
    double t0 = now();
    for_ranks(h, nthreads, [&](int rk) {
        RankObjs& o = h->r[rk];
        Topo* topo = o.topo;
        Uvec* A = new Uvec(topo, o.geom, o.node, o.edge);
        VecZeroEntries(A->vl);
        VecZeroEntries(A->vg);
        for (int t = 0; t < nterms; t++) {
            Vec v1, r2;
            VecCreateSeq(MPI_COMM_SELF, topo->n1, &v1);
            VecCreateSeq(MPI_COMM_SELF, topo->n2, &r2);
            for (int i = 0; i < topo->n1; i++) v1->a[i] = u[(size_t)t * N1 + topo->loc1[i]];
            for (int i = 0; i < topo->n2; i++) r2->a[i] = h2[(size_t)t * N2 + (size_t)topo->pi * topo->n2 + i];   // faces: global = local + pi n2
            A->assemble_hu(lev, scale, v1, r2, false, fac[t]);
            VecDestroy(&v1);
            VecDestroy(&r2);
        }
        part[rk].assign(A->vl->a, A->vl->a + topo->n1);
        delete A;
    });
    for (long i = 0; i < N1; i++) y[i] = 0.0;
    for (int rk = 0; rk < h->nranks; rk++) {
        Topo* topo = h->r[rk].topo;
        for (int i = 0; i < topo->n1; i++) y[topo->loc1[i]] += part[rk][i];
    }
    return now() - t0;
}

/*
 * Geom::interp0 / interp1_l / interp2_l / interp1_g / interp2_g (eul/Geom.cpp:328-417) of one rank at every quadrature
 * point of every element: which = 0..4, vec = ghosted local DOF vector of the matching space (2-forms: the owned array),
 * out[nel][q2] (which 1, 3: out[nel][q2][2]).
 */
void ref_geom_interp(void* hv, int rank, int which, const double* vec, double* out) {
    Ref* h = (Ref*)hv;
    RankObjs& o = h->r[rank];
    Geom* g = o.geom;
    Topo* t = o.topo;
    const int mp1 = g->quad->n + 1, q2 = mp1 * mp1;
    double* v = const_cast<double*>(vec);
    for (int ey = 0; ey < t->nElsX; ey++)
        for (int ex = 0; ex < t->nElsX; ex++) {
            const long el = (long)ey * t->nElsX + ex;
            for (int q = 0; q < q2; q++) {
                double val[2] = {0.0, 0.0};
                switch (which) {
                    case 0: g->interp0(ex, ey, q % mp1, q / mp1, v, val); break;
                    case 1: g->interp1_l(ex, ey, q % mp1, q / mp1, v, val); break;
                    case 2: g->interp2_l(ex, ey, q % mp1, q / mp1, v, val); break;
                    case 3: g->interp1_g(ex, ey, q % mp1, q / mp1, v, val); break;
                    default: g->interp2_g(ex, ey, q % mp1, q / mp1, v, val); break;
                }
                if (which == 1 || which == 3) {
                    out[(el * q2 + q) * 2 + 0] = val[0];
                    out[(el * q2 + q) * 2 + 1] = val[1];
                } else {
                    out[el * q2 + q] = val[0];
                }
            }
        }
}

/* Geom::initTopog (eul/Geom.cpp:743-764) with the level function of the baroclinic test case (eul/UMJS14.cpp:124-129:
 * stretched levels, mu = 15, ZTOP = 30 km) over a smooth hill  topog = 2000 (z/R)^2;  thick_out[nk][n0q]. */
static int g_topog_nk = 1;
static double ref_topog_fn(double* x) {
    const double r2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
    return 2000.0 * x[2] * x[2] / r2;
}
static double ref_level_fn(double* x, int ki) {
    (void)x;
    const double mu = 15.0, ztop = 30000.0, f = (double)ki / g_topog_nk;
    return ztop * (sqrt(mu * f * f + 1.0) - 1.0) / (sqrt(mu + 1.0) - 1.0);
}
void ref_init_topog(void* hv, int rank, double* thick_out) {
    Ref* h = (Ref*)hv;
    Geom* g = h->r[rank].geom;
    g_topog_nk = h->nk;
    g->initTopog(ref_topog_fn, ref_level_fn);
    for (int k = 0; k < h->nk; k++)
        for (int i = 0; i < g->n0; i++) thick_out[(long)k * g->n0 + i] = g->thick[k][i];
}
#endif

}  // extern "C"
