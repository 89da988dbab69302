#include "Assembly.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "../../include/mimsem_gpu.h"

// ------------------------------------------------------------------------------------------------
// one device context per Topo (patch), shared by every operator built on it

namespace {

struct Patch {
    mimsem_gpu_ctx* ctx = NULL;
    Geom* geom = NULL;
    unsigned long thick_version = (unsigned long)-1;
    int refs = 0;
};
std::map<Topo*, Patch> g_patches;

void die(const char* where) {
    // the reference has no error convention (PETSc codes are dropped everywhere); a device failure here
    // cannot be ignored, so it is fatal and loud
    std::fprintf(stderr, "mimsem host adaptor: %s failed: %s\n", where, mimsem_last_error());
    std::abort();
}

Patch* attach(Topo* topo, Geom* geom) {
    // the tables of Geom's own bases are the ones the reference's Geom::interp* use (eul/Geom.cpp:52-54)
    LagrangeNode* l = geom->node;
    LagrangeEdge* e = geom->edge;
    Patch& p = g_patches[topo];
    if (p.ctx) return &p;
    int dev = 0;
    if (const char* s = std::getenv("MIMSEM_DEVICE")) dev = std::atoi(s);
    if (mimsem_gpu_create(dev, &p.ctx)) die("mimsem_gpu_create");
    const int n = topo->elOrd, m = geom->quad->n, np1 = n + 1, mp1 = m + 1;
    std::vector<double> lj((size_t)mp1 * np1), ej((size_t)mp1 * n);
    for (int q = 0; q < mp1; q++) {
        for (int j = 0; j < np1; j++) lj[(size_t)q * np1 + j] = l->ljxi[q][j];
        for (int j = 0; j < n; j++) ej[(size_t)q * n + j] = e->ejxi[q][j];
    }
    if (mimsem_gpu_set_basis(p.ctx, n, m, geom->quad->w, lj.data(), ej.data())) die("mimsem_gpu_set_basis");
    const int nel = topo->nElsX * topo->nElsX;
    std::vector<int> e0((size_t)nel * np1 * np1), e1x((size_t)nel * n * np1), e1y((size_t)nel * n * np1), e2((size_t)nel * n * n),
        eq((size_t)nel * mp1 * mp1);
    for (int ey = 0; ey < topo->nElsX; ey++)
        for (int ex = 0; ex < topo->nElsX; ex++) {
            const size_t el = (size_t)ey * topo->nElsX + ex;
            std::memcpy(&e0[el * np1 * np1], topo->elInds0_l(ex, ey), sizeof(int) * np1 * np1);
            std::memcpy(&e1x[el * n * np1], topo->elInds1x_l(ex, ey), sizeof(int) * n * np1);
            std::memcpy(&e1y[el * n * np1], topo->elInds1y_l(ex, ey), sizeof(int) * n * np1);
            std::memcpy(&e2[el * n * n], topo->elInds2_l(ex, ey), sizeof(int) * n * n);
            std::memcpy(&eq[el * mp1 * mp1], geom->elInds0_l(ex, ey), sizeof(int) * mp1 * mp1);
        }
    // mode 1: this rank's ghosted-local convention -- east / north DOFs receive partial sums
    if (mimsem_gpu_set_topo(p.ctx, nel, nel, topo->n0, topo->n1, topo->n2, geom->n0, 1, e0.data(), e1x.data(), e1y.data(), e2.data(),
                            eq.data()))
        die("mimsem_gpu_set_topo");
    if (mimsem_gpu_set_geom(p.ctx, geom->flatJ(), geom->flatDet())) die("mimsem_gpu_set_geom");
    p.geom = geom;
    return &p;
}

void sync_thickness(Patch* p) {
    Geom* g = p->geom;
    if (p->thick_version == g->thick_version) return;
    std::vector<double> t((size_t)g->nk * g->n0);
    for (int k = 0; k < g->nk; k++)
        for (int i = 0; i < g->n0; i++) t[(size_t)k * g->n0 + i] = g->thick[k][i];
    if (mimsem_gpu_set_thickness(p->ctx, g->nk, t.data())) die("mimsem_gpu_set_thickness");
    p->thick_version = g->thick_version;
}

}  // namespace

int MimsemAttachPatch(Topo* topo, Geom* geom, LagrangeNode* l, LagrangeEdge* e) {
    (void)l;
    (void)e;
    attach(topo, geom);
    return 0;
}
const char* MimsemLastError(void) { return mimsem_last_error(); }

// ------------------------------------------------------------------------------------------------
// the MatShell

struct MimsemShell {
    Topo* topo = NULL;
    int op = 0;            // mimsem_gpu_apply_host operator id
    int sin = 0, sout = 0; // k-form degree of the input / output space
    int lev = 0, tpow = 0, flags = 0;
    double scale = 1.0;
    std::vector<double> coeff;   // coefficient field in the rank-local numbering, copied at assemble() time
    std::vector<double> u1;      // advecting velocity of the upwinded operators (ghosted local 1-form)
    double tau = 0.0;            // fac*dt
    Vec xl = NULL, yl = NULL;    // ghosted local work vectors
    Mat mat = NULL;
};

namespace {

int space_size_local(Topo* t, int s) { return s == 0 ? t->n0 : (s == 1 ? t->n1 : t->n2); }
int space_size_owned(Topo* t, int s) { return s == 0 ? t->n0l : (s == 1 ? t->n1l : t->n2l); }
int space_size_global(Topo* t, int s) { return s == 0 ? t->nDofs0G : (s == 1 ? t->nDofs1G : t->nDofs2G); }

PetscErrorCode shell_mult(Mat A, Vec x, Vec y) {
    MimsemShell* s;
    MatShellGetContext(A, &s);
    Topo* topo = s->topo;
    std::map<Topo*, Patch>::iterator it = g_patches.find(topo);
    if (it == g_patches.end() || !it->second.ctx) {
        std::fprintf(stderr, "mimsem host adaptor: no device patch for this Topo (construct a geometric operator or call MimsemAttachPatch first)\n");
        std::abort();
    }
    Patch* p = &it->second;
    if (s->tpow > 0) sync_thickness(p);
    PetscScalar *xa, *ya;
    // 1. ghosted local input
    if (s->sin == 2) {
        VecGetArray(x, &xa);   // faces have no ghosts: the owned array IS the local array (n2 == n2l)
    } else {
        VecScatter sc = s->sin == 0 ? topo->gtol_0 : topo->gtol_1;
        VecScatterBegin(sc, x, s->xl, INSERT_VALUES, SCATTER_FORWARD);
        VecScatterEnd(sc, x, s->xl, INSERT_VALUES, SCATTER_FORWARD);
        VecGetArray(s->xl, &xa);
    }
    if (s->sout == 2) VecGetArray(y, &ya);
    else VecGetArray(s->yl, &ya);
    // 2. the CUDA kernels (single level: one column)
    if (mimsem_gpu_apply_host_up(p->ctx, s->op, s->lev, 1, s->scale, s->tpow, s->flags, s->coeff.empty() ? NULL : s->coeff.data(),
                                 s->u1.empty() ? NULL : s->u1.data(), s->tau, xa, ya))
        die("mimsem_gpu_apply_host");
    if (s->sin == 2) VecRestoreArray(x, &xa);
    else VecRestoreArray(s->xl, &xa);
    // 3. sum the partial results of shared DOFs into the global vector
    if (s->sout == 2) {
        VecRestoreArray(y, &ya);
    } else {
        VecRestoreArray(s->yl, &ya);
        VecScatter sc = s->sout == 0 ? topo->gtol_0 : topo->gtol_1;
        VecZeroEntries(y);
        VecScatterBegin(sc, s->yl, y, ADD_VALUES, SCATTER_REVERSE);
        VecScatterEnd(sc, s->yl, y, ADD_VALUES, SCATTER_REVERSE);
    }
    return 0;
}

MimsemShell* make_shell(Topo* topo, int op, int sin, int sout, Mat* M) {
    MimsemShell* s = new MimsemShell;
    s->topo = topo;
    s->op = op;
    s->sin = sin;
    s->sout = sout;
    if (sin != 2) VecCreateSeq(MPI_COMM_SELF, space_size_local(topo, sin), &s->xl);
    if (sout != 2) VecCreateSeq(MPI_COMM_SELF, space_size_local(topo, sout), &s->yl);
    MatCreateShell(MPI_COMM_WORLD, space_size_owned(topo, sout), space_size_owned(topo, sin), space_size_global(topo, sout),
                   space_size_global(topo, sin), s, M);
    MatShellSetOperation(*M, MATOP_MULT, (void (*)(void))shell_mult);
    s->mat = *M;
    return s;
}

void free_shell(MimsemShell* s, Mat* M) {
    if (s->xl) VecDestroy(&s->xl);
    if (s->yl) VecDestroy(&s->yl);
    MatDestroy(M);
    delete s;
}

void copy_coeff(MimsemShell* s, Vec v, int n) {
    PetscScalar* a;
    VecGetArray(v, &a);
    s->coeff.assign(a, a + n);
    VecRestoreArray(v, &a);
}

enum { OP_M1 = 0, OP_M2 = 1, OP_M0 = 2, OP_M1H = 3, OP_K = 4, OP_M2H = 5, OP_M0H = 6, OP_R = 7, OP_R_UP = 8, OP_M0H_UP = 9, OP_INC = 10 };

}  // namespace

// ------------------------------------------------------------------------------------------------
// operators

Umat::Umat(Topo* _topo, Geom* _geom, LagrangeNode* _l, LagrangeEdge* _e) : topo(_topo), geom(_geom), l(_l), e(_e), MT(NULL) {
    attach(topo, geom);
    sh = make_shell(topo, OP_M1, 1, 1, &M);
    assemble(0, SCALE, false);   // the reference's constructor assembles level 0 without the vertical scaling (eul/Assembly.cpp:46)
}
Umat::~Umat() { free_shell(sh, &M); }
void Umat::assemble(int lev, double scale, bool vert_scale) {
    sh->lev = lev;
    sh->scale = scale;
    sh->tpow = vert_scale ? 1 : 0;
}

Wmat::Wmat(Topo* _topo, Geom* _geom, LagrangeEdge* _e) : topo(_topo), geom(_geom), e(_e) {
    attach(topo, geom);
    sh = make_shell(topo, OP_M2, 2, 2, &M);
    assemble(0, SCALE, false);   // eul/Assembly.cpp:306
}
Wmat::~Wmat() { free_shell(sh, &M); }
void Wmat::assemble(int lev, double scale, bool vert_scale) {
    sh->lev = lev;
    sh->scale = scale;
    sh->tpow = vert_scale ? 1 : 0;
}

Pmat::Pmat(Topo* _topo, Geom* _geom, LagrangeNode* _node) : topo(_topo), geom(_geom), node(_node) {
    attach(topo, geom);
    sh = make_shell(topo, OP_M0, 0, 0, &M);
    sh->scale = SCALE;
    sh->tpow = 1;
}
Pmat::~Pmat() { free_shell(sh, &M); }
void Pmat::assemble(int lev, double scale) {      // eul/Assembly.cpp:2004-2043: always one 1/thick factor
    sh->op = OP_M0;
    sh->coeff.clear();
    sh->lev = lev;
    sh->scale = scale;
    sh->tpow = 1;
}
void Pmat::assemble_h(int lev, double scale, Vec h2) {   // eul/Assembly.cpp:2045-2098: two 1/thick factors
    sh->op = OP_M0H;
    copy_coeff(sh, h2, topo->n2);
    sh->lev = lev;
    sh->scale = scale;
    sh->tpow = 2;
}

Uhmat::Uhmat(Topo* _topo, Geom* _geom, LagrangeNode* _l, LagrangeEdge* _e) : topo(_topo), geom(_geom), l(_l), e(_e), MT(NULL) {
    attach(topo, geom);
    sh = make_shell(topo, OP_M1H, 1, 1, &M);
}
Uhmat::~Uhmat() { free_shell(sh, &M); }
void Uhmat::assemble(Vec h2, int lev, bool const_vert, double scale) {   // eul/Assembly.cpp:416-474
    copy_coeff(sh, h2, topo->n2);
    sh->lev = lev;
    sh->scale = scale;
    sh->tpow = const_vert ? 2 : 1;
}

Whmat::Whmat(Topo* _topo, Geom* _geom, LagrangeEdge* _e) : topo(_topo), geom(_geom), e(_e) {
    attach(topo, geom);
    sh = make_shell(topo, OP_M2H, 2, 2, &M);
}
Whmat::~Whmat() { free_shell(sh, &M); }
void Whmat::assemble(Vec rho, int lev, double scale, bool vert_scale_rho) {   // eul/Assembly.cpp:1243-1299
    copy_coeff(sh, rho, topo->n2);
    sh->lev = lev;
    sh->scale = scale;
    sh->tpow = vert_scale_rho ? 2 : 1;
}

WtQUmat::WtQUmat(Topo* _topo, Geom* _geom, LagrangeNode* _l, LagrangeEdge* _e) : topo(_topo), geom(_geom), l(_l), e(_e) {
    attach(topo, geom);
    sh = make_shell(topo, OP_K, 1, 2, &M);
}
WtQUmat::~WtQUmat() { free_shell(sh, &M); }
void WtQUmat::assemble(Vec u1, int lev, double scale) {   // eul/Assembly.cpp:933-986; u1 is the ghosted local vector
    copy_coeff(sh, u1, topo->n1);
    sh->lev = lev;
    sh->scale = scale;
    sh->tpow = 2;
}

RotMat::RotMat(Topo* _topo, Geom* _geom, LagrangeNode* _l, LagrangeEdge* _e) : topo(_topo), geom(_geom), l(_l), e(_e) {
    attach(topo, geom);
    sh = make_shell(topo, OP_R, 1, 1, &M);
}
RotMat::~RotMat() { free_shell(sh, &M); }
void RotMat::assemble(Vec q0, int lev, double scale) {   // eul/Assembly.cpp:1030-1083: vort/thick and Qab/thick
    copy_coeff(sh, q0, topo->n0);
    sh->lev = lev;
    sh->scale = scale;
    sh->tpow = 2;
}
void RotMat::assemble(Vec q0) {   // src/Assembly.cpp:1346-1395
    copy_coeff(sh, q0, topo->n0);
    sh->lev = 0;
    sh->scale = 1.0;
    sh->tpow = 0;
}

RotMat_up::RotMat_up(Topo* _topo, Geom* _geom, LagrangeNode* _l, LagrangeEdge* _e) : topo(_topo), geom(_geom), l(_l), e(_e) {
    attach(topo, geom);
    sh = make_shell(topo, OP_R_UP, 1, 1, &M);
}
RotMat_up::~RotMat_up() { free_shell(sh, &M); }
void RotMat_up::assemble(Vec q0, Vec ul, double tau, double dt) {   // src/Assembly.cpp:1784-1853: tau_eff = tau*dt (:1818)
    copy_coeff(sh, q0, topo->n0);
    PetscScalar* a;
    VecGetArray(ul, &a);
    sh->u1.assign(a, a + topo->n1);
    VecRestoreArray(ul, &a);
    sh->tau = tau * dt;
    sh->lev = 0;
    sh->scale = 1.0;
    sh->tpow = 0;
}

Phmat::Phmat(Topo* _topo, Geom* _geom, LagrangeNode* _node) : topo(_topo), geom(_geom), node(_node) {
    attach(topo, geom);
    sh = make_shell(topo, OP_M0H, 0, 0, &M);
}
Phmat::~Phmat() { free_shell(sh, &M); }
void Phmat::assemble(Vec h2) {   // src/Assembly.cpp:396-442
    sh->op = OP_M0H;
    copy_coeff(sh, h2, topo->n2);
    sh->u1.clear();
    sh->lev = 0;
    sh->scale = 1.0;
    sh->tpow = 0;
}
void Phmat::assemble_up(Vec ul, Vec hl, double fac, double dt) {   // src/Assembly.cpp:499-567
    sh->op = OP_M0H_UP;
    copy_coeff(sh, hl, topo->n2);
    PetscScalar* a;
    VecGetArray(ul, &a);
    sh->u1.assign(a, a + topo->n1);
    VecRestoreArray(ul, &a);
    sh->tau = fac * dt;
    sh->lev = 0;
    sh->scale = 1.0;
    sh->tpow = 0;
}

Ut_mat::Ut_mat(Topo* _topo, Geom* _geom, LagrangeNode* _l, LagrangeEdge* _e) : topo(_topo), geom(_geom), l(_l), e(_e) {
    attach(topo, geom);
    sh = make_shell(topo, OP_M1, 1, 1, &M);
}
Ut_mat::~Ut_mat() { free_shell(sh, &M); }
void Ut_mat::assemble(int lev, double scale) {   // eul/Assembly.cpp:1338-1388: Umat times 0.5 (thick[lev] + thick[lev+1])
    sh->op = OP_M1;
    sh->coeff.clear();
    sh->lev = lev;
    sh->scale = scale;
    sh->tpow = 1;
    sh->flags = MIMSEM_THICK_MEAN;
}
void Ut_mat::assemble_h(int lev, double scale, Vec rho) {   // eul/Assembly.cpp:1390-1440: Uhmat without its 1/thick factors
    sh->op = OP_M1H;
    copy_coeff(sh, rho, topo->n2);
    sh->lev = lev;
    sh->scale = scale;
    sh->tpow = 0;
    sh->flags = 0;
}

WtQdUdz_mat::WtQdUdz_mat(Topo* _topo, Geom* _geom, LagrangeNode* _l, LagrangeEdge* _e) : topo(_topo), geom(_geom), l(_l), e(_e) {
    attach(topo, geom);
    sh = make_shell(topo, OP_K, 1, 2, &M);
}
WtQdUdz_mat::~WtQdUdz_mat() { free_shell(sh, &M); }
void WtQdUdz_mat::assemble(Vec u1, double scale) {   // eul/Assembly.cpp:1581-1640: WtQUmat without the 1/2 and the 1/thick^2
    copy_coeff(sh, u1, topo->n1);
    sh->lev = 0;
    sh->scale = 2.0 * scale;
    sh->tpow = 0;
}

E10mat::E10mat(Topo* _topo) : topo(_topo) {
    sh10 = make_shell(topo, OP_INC + MIMSEM_E10, 0, 1, &E10);
    sh01 = make_shell(topo, OP_INC + MIMSEM_E01, 1, 0, &E01);
}
E10mat::~E10mat() {
    free_shell(sh10, &E10);
    free_shell(sh01, &E01);
}

E21mat::E21mat(Topo* _topo) : topo(_topo) {
    sh21 = make_shell(topo, OP_INC + MIMSEM_E21, 1, 2, &E21);
    sh12 = make_shell(topo, OP_INC + MIMSEM_E12, 2, 1, &E12);
}
E21mat::~E21mat() {
    free_shell(sh21, &E21);
    free_shell(sh12, &E12);
}
