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

// MatGetDiagonal of the mass-matrix shells (Jacobi-type preconditioners of a KSP on the shell)
PetscErrorCode shell_getdiag(Mat A, Vec d) {
    MimsemShell* s;
    MatShellGetContext(A, &s);
    int dop;
    if (s->op == 0 /* M1 */ && !(s->flags & MIMSEM_THICK_MEAN)) dop = 19;
    else if (s->op == 2 /* M0 */) dop = 15;
    else if (s->op == 6 /* M0h */) dop = 16;
    else return 56;   // PETSC_ERR_SUP
    Topo* topo = s->topo;
    std::map<Topo*, Patch>::iterator it = g_patches.find(topo);
    if (it == g_patches.end() || !it->second.ctx) std::abort();
    Patch* p = &it->second;
    if (s->tpow > 0) sync_thickness(p);
    PetscScalar* ya;
    VecGetArray(s->yl, &ya);
    if (mimsem_gpu_apply_host_up(p->ctx, dop, s->lev, 1, s->scale, s->tpow, s->flags, s->coeff.empty() ? NULL : s->coeff.data(), NULL, 0.0, ya, ya))
        die("mimsem_gpu_apply_host (diagonal)");
    VecRestoreArray(s->yl, &ya);
    VecScatter sc = s->sout == 0 ? topo->gtol_0 : topo->gtol_1;
    VecZeroEntries(d);
    VecScatterBegin(sc, s->yl, d, ADD_VALUES, SCATTER_REVERSE);
    VecScatterEnd(sc, s->yl, d, ADD_VALUES, SCATTER_REVERSE);
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
    if (sin == sout && sin != 2) MatShellSetOperation(*M, MATOP_GET_DIAGONAL, (void (*)(void))shell_getdiag);
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

enum { OP_M1 = 0, OP_M2 = 1, OP_M0 = 2, OP_M1H = 3, OP_K = 4, OP_M2H = 5, OP_M0H = 6, OP_R = 7, OP_R_UP = 8, OP_M0H_UP = 9, OP_INC = 10,
       OP_UTQW = 14, OP_DIAG_M0 = 15, OP_DIAG_M0H = 16, OP_M2INV = 17, OP_M2HINV = 18 };

Patch* patch_of(Topo* topo) {
    std::map<Topo*, Patch>::iterator it = g_patches.find(topo);
    if (it == g_patches.end() || !it->second.ctx) {
        std::fprintf(stderr, "mimsem host adaptor: no device patch for this Topo\n");
        std::abort();
    }
    return &it->second;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// operators

Umat::Umat(Topo* _topo, Geom* _geom, LagrangeNode* _l, LagrangeEdge* _e) : topo(_topo), geom(_geom), l(_l), e(_e), MT(NULL), Mo(NULL), sho(NULL) {
    attach(topo, geom);
    sh = make_shell(topo, OP_M1, 1, 1, &M);
    if (topo->kind == MIMSEM_MESH_BOX) {
        // box/Assembly.cpp:44-45: M = assemble(0, SCALE, true), Mo = assemble(0, SCALE, false), both once and for all
        assemble(0, SCALE, true);
        sh->flags = MIMSEM_FIXED_LEVEL;
        sho = make_shell(topo, OP_M1, 1, 1, &Mo);
        sho->lev = 0;
        sho->scale = SCALE;
        sho->tpow = 0;
    } else if (geom->nk > 0) {
        assemble(0, SCALE, false);   // the eul/ constructor assembles level 0 without the vertical scaling (eul/Assembly.cpp:46)
    } else {
        assemble();                  // src/Assembly.cpp:27
    }
}
Umat::~Umat() {
    free_shell(sh, &M);
    if (sho) free_shell(sho, &Mo);
}
void Umat::assemble(int lev, double scale, bool vert_scale) {
    sh->lev = lev;
    sh->scale = scale;
    sh->tpow = vert_scale ? 1 : 0;
}
void Umat::assemble() {
    sh->lev = 0;
    sh->scale = 1.0;
    sh->tpow = 0;
}

Wmat::Wmat(Topo* _topo, Geom* _geom, LagrangeEdge* _e) : topo(_topo), geom(_geom), e(_e), Mo(NULL), sho(NULL) {
    attach(topo, geom);
    sh = make_shell(topo, OP_M2, 2, 2, &M);
    if (topo->kind == MIMSEM_MESH_BOX) {
        assemble(0, SCALE, true);    // box/Assembly.cpp:171-172
        sh->flags = MIMSEM_FIXED_LEVEL;
        sho = make_shell(topo, OP_M2, 2, 2, &Mo);
        sho->lev = 0;
        sho->scale = SCALE;
        sho->tpow = 0;
    } else if (geom->nk > 0) {
        assemble(0, SCALE, false);   // eul/Assembly.cpp:306
    } else {
        assemble();                  // src/Assembly.cpp:257
    }
}
Wmat::~Wmat() {
    free_shell(sh, &M);
    if (sho) free_shell(sho, &Mo);
}
void Wmat::assemble(int lev, double scale, bool vert_scale) {
    sh->lev = lev;
    sh->scale = scale;
    sh->tpow = vert_scale ? 1 : 0;
}
void Wmat::assemble() {
    sh->lev = 0;
    sh->scale = 1.0;
    sh->tpow = 0;
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
void Pmat::assemble() {   // src/Assembly.cpp:324-372
    sh->op = OP_M0;
    sh->coeff.clear();
    sh->lev = 0;
    sh->scale = 1.0;
    sh->tpow = 0;
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

void Uhmat::assemble(Vec h2) {   // src/Assembly.cpp:675-734
    copy_coeff(sh, h2, topo->n2);
    sh->lev = 0;
    sh->scale = 1.0;
    sh->tpow = 0;
}

Umat_ray::Umat_ray(Topo* _topo, Geom* _geom, LagrangeNode* _l, LagrangeEdge* _e) : topo(_topo), geom(_geom), l(_l), e(_e) {
    attach(topo, geom);
    sh = make_shell(topo, 21 /* Umat_ray */, 1, 1, &M);
}
Umat_ray::~Umat_ray() { free_shell(sh, &M); }
void Umat_ray::assemble(int lev, double scale, double dt, Vec exner, Vec exner_s) {   // eul/Assembly.cpp:1875-1979
    copy_coeff(sh, exner, topo->n2);
    // the level-0 Exner field rides in the shell's second coefficient slot, dt in its time-scale slot
    PetscScalar* a;
    VecGetArray(exner_s, &a);
    sh->u1.assign(a, a + topo->n2);
    VecRestoreArray(exner_s, &a);
    sh->tau = dt;
    sh->lev = lev;
    sh->scale = scale;
    sh->tpow = 1;
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

void WtQUmat::assemble(Vec u1) {   // src/Assembly.cpp:1172-1218
    copy_coeff(sh, u1, topo->n1);
    sh->lev = 0;
    sh->scale = 1.0;
    sh->tpow = 0;
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

UtQWmat::UtQWmat(Topo* _topo, Geom* _geom, LagrangeNode* _l, LagrangeEdge* _e) : topo(_topo), geom(_geom), l(_l), e(_e) {
    attach(topo, geom);
    sh = make_shell(topo, OP_UTQW, 2, 1, &M);
}
UtQWmat::~UtQWmat() { free_shell(sh, &M); }
void UtQWmat::assemble(Vec u1, double scale) {   // eul/Assembly.cpp:1490-1538
    copy_coeff(sh, u1, topo->n1);
    sh->lev = 0;
    sh->scale = scale;
    sh->tpow = 0;
}

WmatInv::WmatInv(Topo* _topo, Geom* _geom, LagrangeEdge* _e) : topo(_topo), geom(_geom), e(_e) {
    attach(topo, geom);
    sh = make_shell(topo, OP_M2INV, 2, 2, &M);
}
WmatInv::~WmatInv() { free_shell(sh, &M); }
void WmatInv::assemble(int lev, double scale) {   // eul/Assembly.cpp:1673-1722: always one 1/thick factor
    sh->lev = lev;
    sh->scale = scale;
    sh->tpow = 1;
}

WhmatInv::WhmatInv(Topo* _topo, Geom* _geom, LagrangeEdge* _e) : topo(_topo), geom(_geom), e(_e) {
    attach(topo, geom);
    sh = make_shell(topo, OP_M2HINV, 2, 2, &M);
}
WhmatInv::~WhmatInv() { free_shell(sh, &M); }
void WhmatInv::assemble(Vec rho, int lev, double scale) {   // eul/Assembly.cpp:1744-1800
    copy_coeff(sh, rho, topo->n2);
    sh->lev = lev;
    sh->scale = scale;
    sh->tpow = 2;
}

// ------------------------------------------------------------------------------------------------
// vectors

namespace {
// one device call on this rank's patch: ghosted local in (may be NULL), ghosted local out
void patch_apply(Topo* topo, int op, int lev, double scale, int tpow, const double* coeff, const double* x, double* y) {
    Patch* p = patch_of(topo);
    if (tpow > 0) sync_thickness(p);
    if (mimsem_gpu_apply_host_up(p->ctx, op, lev, 1, scale, tpow, 0, coeff, NULL, 0.0, x ? x : y, y)) die("mimsem_gpu_apply_host");
}
void reverse_add(Topo* topo, int space, Vec vl, Vec vg) {
    VecScatter sc = space == 0 ? topo->gtol_0 : topo->gtol_1;
    VecScatterBegin(sc, vl, vg, ADD_VALUES, SCATTER_REVERSE);
    VecScatterEnd(sc, vl, vg, ADD_VALUES, SCATTER_REVERSE);
}
}  // namespace

Pvec::Pvec(Topo* _topo, Geom* _geom, LagrangeNode* _l) : topo(_topo), geom(_geom), l(_l) {
    attach(topo, geom);
    VecCreateSeq(MPI_COMM_SELF, topo->n0, &vl);
    VecCreateMPI(MPI_COMM_WORLD, topo->n0l, topo->nDofs0G, &vg);
}
Pvec::~Pvec() {
    VecDestroy(&vl);
    VecDestroy(&vg);
}
void Pvec::assemble(int lev, double scale) {
    PetscScalar* a;
    VecZeroEntries(vg);
    VecGetArray(vl, &a);
    patch_apply(topo, OP_DIAG_M0, lev, scale, 1, NULL, NULL, a);
    VecRestoreArray(vl, &a);
    reverse_add(topo, 0, vl, vg);
    // and back to the local vector (eul/Assembly.cpp:625-627)
    VecScatterBegin(topo->gtol_0, vg, vl, INSERT_VALUES, SCATTER_FORWARD);
    VecScatterEnd(topo->gtol_0, vg, vl, INSERT_VALUES, SCATTER_FORWARD);
}

Phvec::Phvec(Topo* _topo, Geom* _geom, LagrangeNode* _l) : topo(_topo), geom(_geom), l(_l) {
    attach(topo, geom);
    VecCreateSeq(MPI_COMM_SELF, topo->n0, &vl);
    VecCreateMPI(MPI_COMM_WORLD, topo->n0l, topo->nDofs0G, &vg);
}
Phvec::~Phvec() {
    VecDestroy(&vl);
    VecDestroy(&vg);
}
void Phvec::assemble(Vec hl, int lev, double scale) {
    PetscScalar *a, *h;
    VecZeroEntries(vg);
    VecGetArray(vl, &a);
    VecGetArray(hl, &h);
    patch_apply(topo, OP_DIAG_M0H, lev, scale, 2, h, NULL, a);
    VecRestoreArray(hl, &h);
    VecRestoreArray(vl, &a);
    reverse_add(topo, 0, vl, vg);
}

Uvec::Uvec(Topo* _topo, Geom* _geom, LagrangeNode* _node, LagrangeEdge* _edge) : topo(_topo), geom(_geom), node(_node), edge(_edge) {
    attach(topo, geom);
    VecCreateSeq(MPI_COMM_SELF, topo->n1, &vl);
    VecCreateMPI(MPI_COMM_WORLD, topo->n1l, topo->nDofs1G, &vg);
}
Uvec::~Uvec() {
    VecDestroy(&vl);
    VecDestroy(&vg);
}
// vl += (this patch's partial sums of) M1 vel  resp.  M1(rho) vel
void Uvec::accumulate(int op, int lev, double scale, int tpow, Vec vel, Vec rho) {
    std::vector<double> y(topo->n1);
    PetscScalar *v, *u, *r = NULL;
    VecGetArray(vel, &u);
    if (rho) VecGetArray(rho, &r);
    patch_apply(topo, op, lev, scale, tpow, r, u, y.data());
    if (rho) VecRestoreArray(rho, &r);
    VecRestoreArray(vel, &u);
    VecGetArray(vl, &v);
    for (int i = 0; i < topo->n1; i++) v[i] += y[i];
    VecRestoreArray(vl, &v);
}
void Uvec::assemble(int lev, double scale, bool vert_scale, Vec vel) {
    (void)vert_scale;   // the reference multiplies by 1/thick unconditionally (eul/Assembly.cpp:2153-2156)
    VecZeroEntries(vl);
    VecZeroEntries(vg);
    accumulate(OP_M1, lev, scale, 1, vel, NULL);
    reverse_add(topo, 1, vl, vg);
}
void Uvec::assemble_hu(int lev, double scale, Vec vel, Vec rho, bool zero_and_scatter, double fac) {
    if (zero_and_scatter) {
        VecZeroEntries(vl);
        VecZeroEntries(vg);
    }
    accumulate(OP_M1H, lev, scale * fac, 2, vel, rho);   // fac multiplies the interpolated rho (eul/Assembly.cpp:2237)
    if (zero_and_scatter) reverse_add(topo, 1, vl, vg);
}

void Uvec::assemble_hu(int lev, double scale, bool vert_scale, Vec vel, Vec rho) {   // box/: always zeroes and scatters
    VecZeroEntries(vl);
    VecZeroEntries(vg);
    accumulate(OP_M1H, lev, scale, vert_scale ? 2 : 1, vel, rho);
    reverse_add(topo, 1, vl, vg);
}

Wvec::Wvec(Topo* _topo, Geom* _geom, LagrangeEdge* _edge) : topo(_topo), geom(_geom), edge(_edge) {
    attach(topo, geom);
    VecCreateMPI(MPI_COMM_WORLD, topo->n2l, topo->nDofs2G, &vg);
}
Wvec::~Wvec() { VecDestroy(&vg); }
void Wvec::assemble(int lev, double scale, bool vert_scale, Vec rho) {
    PetscScalar *y, *x;
    VecGetArray(vg, &y);
    VecGetArray(rho, &x);
    patch_apply(topo, OP_M2, lev, scale, vert_scale ? 1 : 0, NULL, x, y);
    VecRestoreArray(rho, &x);
    VecRestoreArray(vg, &y);
}
void Wvec::assemble_K(int lev, double scale, Vec vel1, Vec vel2) {
    PetscScalar *y, *x, *u;
    VecGetArray(vg, &y);
    VecGetArray(vel1, &x);
    VecGetArray(vel2, &u);
    patch_apply(topo, OP_K, lev, scale, 2, u, x, y);
    VecRestoreArray(vel2, &u);
    VecRestoreArray(vel1, &x);
    VecRestoreArray(vg, &y);
}

void Wvec::assemble_K(int lev, double scale, bool vert_scale, Vec vel1, Vec vel2) {
    (void)vert_scale;
    assemble_K(lev, scale, vel1, vel2);
}

// quadrature-point values -> 0-form: y_n = sum over the (element, point) pairs at node n of w_q det x_q.  With m == p the
// nodal table is the identity, so P^T Q is a weighted injection; start-up only, evaluated on the host.
namespace {
PetscErrorCode ptq_mult(Mat A, Vec x, Vec y) {
    PtQmat* self;
    MatShellGetContext(A, &self);
    Topo* topo = self->topo;
    Geom* geom = self->geom;
    const int mp1 = geom->quad->n + 1, mp12 = mp1 * mp1;
    std::vector<double> xl(geom->n0), yl(topo->n0, 0.0);
    PetscScalar* xa;
    {
        // ghosted local copy of the quadrature-point vector (Geom::gtol_0)
        Vec tmp;
        VecCreateSeq(MPI_COMM_SELF, geom->n0, &tmp);
        VecScatterBegin(geom->gtol_0, x, tmp, INSERT_VALUES, SCATTER_FORWARD);
        VecScatterEnd(geom->gtol_0, x, tmp, INSERT_VALUES, SCATTER_FORWARD);
        VecGetArray(tmp, &xa);
        for (int i = 0; i < geom->n0; i++) xl[i] = xa[i];
        VecRestoreArray(tmp, &xa);
        VecDestroy(&tmp);
    }
    for (int ey = 0; ey < topo->nElsX; ey++)
        for (int ex = 0; ex < topo->nElsX; ex++) {
            const int ei = ey * topo->nElsX + ex;
            const int* iq = geom->elInds0_l(ex, ey);
            std::vector<int> q(iq, iq + mp12);
            const int* in = topo->elInds0_l(ex, ey);
            for (int ii = 0; ii < mp12; ii++)
                yl[in[ii]] += geom->quad->w[ii % mp1] * geom->quad->w[ii / mp1] * geom->det[ei][ii] * xl[q[ii]];
        }
    Vec yv;
    VecCreateSeq(MPI_COMM_SELF, topo->n0, &yv);
    PetscScalar* ya;
    VecGetArray(yv, &ya);
    for (int i = 0; i < topo->n0; i++) ya[i] = yl[i];
    VecRestoreArray(yv, &ya);
    VecZeroEntries(y);
    VecScatterBegin(topo->gtol_0, yv, y, ADD_VALUES, SCATTER_REVERSE);
    VecScatterEnd(topo->gtol_0, yv, y, ADD_VALUES, SCATTER_REVERSE);
    VecDestroy(&yv);
    return 0;
}
}  // namespace

PtQmat::PtQmat(Topo* _topo, Geom* _geom, LagrangeNode* _l) : topo(_topo), geom(_geom), l(_l), xl(NULL), yl(NULL) {
    MatCreateShell(MPI_COMM_WORLD, topo->n0l, geom->n0l, topo->nDofs0G, geom->nDofs0G, this, &M);
    MatShellSetOperation(M, MATOP_MULT, (void (*)(void))ptq_mult);
}
PtQmat::~PtQmat() { MatDestroy(&M); }
void PtQmat::assemble() {}

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
