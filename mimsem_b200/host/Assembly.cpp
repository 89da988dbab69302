#include "Assembly.h"

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "../../include/mimsem_gpu.h"
#include "Partition.h"

// ------------------------------------------------------------------------------------------------
// one device context per Topo (patch), shared by every operator built on it

namespace {

struct Patch {
    mimsem_gpu_ctx* ctx = NULL;
    Geom* geom = NULL;
    unsigned long thick_version = (unsigned long)-1;
    int refs = 0;
    // second context of the element-block Jacobi preconditioner (pc_patch): the patch's elements in owner-computes
    // mode plus one copy of the west / south neighbour of every element on the patch's west / south boundary
    mimsem_gpu_ctx* pc_ctx = NULL;
    bool pc_failed = false;
    int pc_n1 = 0, pc_nq = 0;
    unsigned long pc_thick_version = (unsigned long)-1;
    std::vector<double> pc_x, pc_y;
    // page-locked staging of MimsemMatMultLevels: input, output, coefficient ([nlev][local size] each)
    double* stage[3] = {NULL, NULL, NULL};
    size_t stage_cap[3] = {0, 0, 0};
};
std::map<Topo*, Patch> g_patches;

// wall time spent inside the device library's host-buffer calls (copies in, kernels, copies out), for MimsemDeviceSeconds
double g_device_seconds = 0.0;
struct DeviceTimer {
    std::chrono::steady_clock::time_point t0;
    DeviceTimer() : t0(std::chrono::steady_clock::now()) {}
    ~DeviceTimer() { g_device_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

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


// canonical global meshes (closed form, csrc/mesh.cpp), shared by the patches of a process
const mimsem_host::GlobalMesh* canonical_mesh(int kind, int p, int ne) {
    static std::map<std::vector<int>, mimsem_host::GlobalMesh*> cache;
    const std::vector<int> key = {kind, p, ne};
    std::map<std::vector<int>, mimsem_host::GlobalMesh*>::iterator it = cache.find(key);
    if (it != cache.end()) return it->second;
    mimsem_host::GlobalMesh* m = new mimsem_host::GlobalMesh;
    if (m->create(kind, p, ne)) {
        delete m;
        m = NULL;
    }
    cache[key] = m;
    return m;
}

// The reference preconditions its 1-form solves with PCBJACOBI, one block per element (PCBJacobiSetTotalBlocks(pc,
// size * nElsX^2, NULL), eul/HorizSolve.cpp:77-84): block e = the 2 p^2 edges element e owns.  Its west / south edges also
// carry the far-line terms of the element on the other side, which for the elements along the patch's west / south boundary
// lives on another rank (or, on a one-patch periodic box, at the other end of the same patch, under different local
// rows).  The block kernel (csrc/m1_bjacobi.cuh) runs in owner-computes mode on an element list with neighbours, so the
// preconditioner gets a context of its own: the patch's elements with the patch's local rows, plus one read-only copy of
// each such neighbour whose shared line is mapped onto the patch's rows and whose geometry comes from the closed-form
// canonical mesh (an element's Jacobians do not depend on the partition; placement of a rank's patch: scr/Proc2.py:52-61).
// Returns NULL when the configuration is not covered (the caller falls back to the diagonal).
struct PcTables {
    int nel = 0, nhalo = 0, n0 = 0, n1 = 0, n2 = 0, nq = 0;
    std::vector<int> e0, e1x, e1y, e2, eq;
    std::vector<double> J, det;
};

// host part of pc_patch: element tables and geometry of the preconditioner's context.  false: configuration not covered.
bool pc_tables(Topo* topo, Geom* geom, PcTables& T) {
    const int p = topo->elOrd, m = geom->quad->n, np1 = p + 1, mp1 = m + 1, nelx = topo->nElsX, nel = nelx * nelx;
    if (m != p) return false;
    int size = 1;
    MPI_Comm_size(MPI_COMM_WORLD, &size);
    const int per_face = topo->kind == MIMSEM_MESH_SPHERE ? size / 6 : size;
    int npx = 1;
    while ((npx + 1) * (npx + 1) <= per_face) npx++;
    if (npx * npx != per_face || (topo->kind == MIMSEM_MESH_SPHERE && size % 6)) return false;
    const int ne = nelx * npx, rank = topo->pi;
    const mimsem_host::GlobalMesh* G = canonical_mesh(topo->kind, p, ne);
    if (!G) return false;
    const int face = topo->kind == MIMSEM_MESH_SPHERE ? rank / per_face : 0, pj = rank % per_face, px = pj % npx, py = pj / npx;
    const double* Jown = geom->flatJ();
    const double* Down = geom->flatDet();
    for (size_t i = 0; i < (size_t)nel * mp1 * mp1; i++)
        if (!(Down[i] > 0.0)) return false;   // signed determinants (src/Geom.cpp:251): not covered
    const int w0 = np1 * np1, w1 = p * np1, w2 = p * p, wq = mp1 * mp1;
    std::vector<int>&e0 = T.e0, &e1x = T.e1x, &e1y = T.e1y, &e2 = T.e2, &eq = T.eq;
    std::vector<double>&J = T.J, &det = T.det;
    auto canon = [&](int ex, int ey) { return (int64_t)face * ne * ne + (int64_t)(py * nelx + ey) * ne + px * nelx + ex; };
    // a canonical quadrature point and a point of this patch are the same point of the mesh (the placement of the patch
    // is checked against the coordinates, not trusted; the neighbour copies are tied to it through shared canonical ids)
    double diam = 0.0;
    for (int i = 0; i < 3; i++) diam = std::max(diam, std::fabs(G->xyz[i]) + std::fabs(G->xyz[(size_t)(G->NQ - 1) * 3 + i]));
    double period = 0.0;
    if (topo->kind == MIMSEM_MESH_BOX) {
        // side length of the periodic box from the point set itself: the gap that closes the period equals the first gap
        double lo = G->xyz[0], hi = G->xyz[0], second = 1.0e300;
        for (int64_t q = 0; q < G->NQ; q++) {
            lo = std::min(lo, G->xyz[(size_t)q * 3]);
            hi = std::max(hi, G->xyz[(size_t)q * 3]);
        }
        for (int64_t q = 0; q < G->NQ; q++)
            if (G->xyz[(size_t)q * 3] > lo + 1.0e-9 * (hi - lo)) second = std::min(second, G->xyz[(size_t)q * 3]);
        period = hi - lo + (second - lo);
    }
    auto same_point = [&](int64_t canon_q, int local_q) {
        double d = 0.0;
        for (int i = 0; i < 3; i++) d = std::max(d, std::fabs(G->xyz[(size_t)canon_q * 3 + i] - geom->x[local_q][i]));
        if (d <= 1.0e-6 * diam) return true;
        // doubly periodic box: the same point up to a period (an east / north ghost column of the patch)
        if (topo->kind != MIMSEM_MESH_BOX) return false;
        for (int i = 0; i < 2; i++) {
            const double dx = std::fabs(G->xyz[(size_t)canon_q * 3 + i] - geom->x[local_q][i]);
            if (dx > 1.0e-6 * diam && std::fabs(dx - period) > 1.0e-6 * diam) return false;
        }
        return true;
    };
    for (int ey = 0; ey < nelx; ey++)
        for (int ex = 0; ex < nelx; ex++) {
            const size_t el = (size_t)ey * nelx + ex;
            e0.insert(e0.end(), topo->elInds0_l(ex, ey), topo->elInds0_l(ex, ey) + w0);
            e1x.insert(e1x.end(), topo->elInds1x_l(ex, ey), topo->elInds1x_l(ex, ey) + w1);
            e1y.insert(e1y.end(), topo->elInds1y_l(ex, ey), topo->elInds1y_l(ex, ey) + w1);
            e2.insert(e2.end(), topo->elInds2_l(ex, ey), topo->elInds2_l(ex, ey) + w2);
            eq.insert(eq.end(), geom->elInds0_l(ex, ey), geom->elInds0_l(ex, ey) + wq);
            J.insert(J.end(), Jown + el * wq * 4, Jown + (el + 1) * wq * 4);
            det.insert(det.end(), Down + el * wq, Down + (el + 1) * wq);
            // (element corners: the points inside an element are re-projected per element by Geom, eul/Geom.cpp:682-724)
            const int64_t c = canon(ex, ey);
            const int corner[4] = {0, m, m * mp1, wq - 1};
            for (int j = 0; j < 4; j++)
                if (!same_point(G->elq[(size_t)c * wq + corner[j]], eq[el * wq + corner[j]])) return false;
        }
    int n0 = topo->n0, n1 = topo->n1, n2 = topo->n2, nq = geom->n0, nhalo = 0;
    for (int ey = 0; ey < nelx; ey++)
        for (int ex = 0; ex < nelx; ex++)
            for (int side = 0; side < 2; side++) {
                if ((side == 0 ? ex : ey) != 0) continue;   // the neighbour is one of the patch's own elements, same rows
                const int64_t c = canon(ex, ey), n = G->ws_nbr[(size_t)c * 2 + side];
                if (n < 0) continue;
                // canonical id -> patch-local row, for what the element has on the shared line
                std::map<int, int> edge_row, node_row, quad_row;
                const size_t el = (size_t)ey * nelx + ex;
                for (int i = 0; i < p; i++) {
                    const int j = side == 0 ? i * np1 : i;
                    if (side == 0) edge_row[G->el1x[(size_t)c * w1 + j]] = e1x[el * w1 + j];
                    else edge_row[G->el1y[(size_t)c * w1 + j]] = e1y[el * w1 + j];
                }
                for (int i = 0; i <= p; i++) {
                    const int j = side == 0 ? i * np1 : i;
                    node_row[G->el0[(size_t)c * w0 + j]] = e0[el * w0 + j];
                    quad_row[G->elq[(size_t)c * wq + j]] = eq[el * wq + j];
                }
                int shared = 0;
                auto row_of = [&](const std::map<int, int>& known, int id, int* fresh) {
                    std::map<int, int>::const_iterator it = known.find(id);
                    if (it == known.end()) return (*fresh)++;
                    shared++;
                    return it->second;
                };
                for (int j = 0; j < w0; j++) e0.push_back(row_of(node_row, G->el0[(size_t)n * w0 + j], &n0));
                shared = 0;
                for (int j = 0; j < w1; j++) e1x.push_back(row_of(edge_row, G->el1x[(size_t)n * w1 + j], &n1));
                for (int j = 0; j < w1; j++) e1y.push_back(row_of(edge_row, G->el1y[(size_t)n * w1 + j], &n1));
                if (shared != p) return false;
                shared = 0;
                for (int j = 0; j < wq; j++) eq.push_back(row_of(quad_row, G->elq[(size_t)n * wq + j], &nq));
                if (shared != p + 1) return false;
                for (int j = 0; j < w2; j++) e2.push_back(n2++);
                J.insert(J.end(), G->J.begin() + (size_t)n * wq * 4, G->J.begin() + (size_t)(n + 1) * wq * 4);
                det.insert(det.end(), G->det.begin() + (size_t)n * wq, G->det.begin() + (size_t)(n + 1) * wq);
                nhalo++;
            }
    T.nel = nel; T.nhalo = nhalo; T.n0 = n0; T.n1 = n1; T.n2 = n2; T.nq = nq;
    return true;
}

mimsem_gpu_ctx* pc_patch(Patch* P, Topo* topo) {
    if (P->pc_ctx || P->pc_failed) return P->pc_ctx;
    P->pc_failed = true;   // until everything below went through
    Geom* geom = P->geom;
    PcTables T;
    if (!pc_tables(topo, geom, T)) return NULL;
    const int p = topo->elOrd, m = geom->quad->n, np1 = p + 1, mp1 = m + 1;
    mimsem_gpu_ctx* ctx = NULL;
    int dev = 0;
    if (const char* s = std::getenv("MIMSEM_DEVICE")) dev = std::atoi(s);
    if (mimsem_gpu_create(dev, &ctx)) die("mimsem_gpu_create (preconditioner)");
    std::vector<double> lj((size_t)mp1 * np1), ej((size_t)mp1 * p);
    for (int q = 0; q < mp1; q++) {
        for (int j = 0; j < np1; j++) lj[(size_t)q * np1 + j] = geom->node->ljxi[q][j];
        for (int j = 0; j < p; j++) ej[(size_t)q * p + j] = geom->edge->ejxi[q][j];
    }
    if (mimsem_gpu_set_basis(ctx, p, m, geom->quad->w, lj.data(), ej.data())) die("mimsem_gpu_set_basis (preconditioner)");
    if (mimsem_gpu_set_topo(ctx, T.nel + T.nhalo, T.nel, T.n0, T.n1, T.n2, T.nq, 0, T.e0.data(), T.e1x.data(), T.e1y.data(), T.e2.data(),
                            T.eq.data()))
        die("mimsem_gpu_set_topo (preconditioner)");
    if (mimsem_gpu_set_geom(ctx, T.J.data(), T.det.data())) die("mimsem_gpu_set_geom (preconditioner)");
    P->pc_ctx = ctx;
    P->pc_n1 = T.n1;
    P->pc_nq = T.nq;
    P->pc_x.assign(T.n1, 0.0);
    P->pc_y.assign(T.n1, 0.0);
    P->pc_failed = false;
    return ctx;
}

void sync_pc_thickness(Patch* P) {
    Geom* g = P->geom;
    if (P->pc_thick_version == g->thick_version) return;
    // the neighbour copies' far lines are the patch's own quadrature points; their other points are never read
    std::vector<double> t((size_t)g->nk * P->pc_nq, 1.0);
    for (int k = 0; k < g->nk; k++)
        for (int i = 0; i < g->n0; i++) t[(size_t)k * P->pc_nq + i] = g->thick[k][i];
    if (mimsem_gpu_set_thickness(P->pc_ctx, g->nk, t.data())) die("mimsem_gpu_set_thickness (preconditioner)");
    P->pc_thick_version = g->thick_version;
}

}  // namespace

int MimsemAttachPatch(Topo* topo, Geom* geom, LagrangeNode* l, LagrangeEdge* e) {
    (void)l;
    (void)e;
    attach(topo, geom);
    return 0;
}
const char* MimsemLastError(void) { return mimsem_last_error(); }
double MimsemDeviceSeconds(int reset) {
    const double t = g_device_seconds;
    if (reset) g_device_seconds = 0.0;
    return t;
}
int MimsemPCTablesCheck(Topo* topo, Geom* geom, int sizes[6]) {
    PcTables T;
    if (!pc_tables(topo, geom, T)) return 1;
    const int v[6] = {T.nel, T.nhalo, T.n0, T.n1, T.n2, T.nq};
    for (int i = 0; i < 6; i++) sizes[i] = v[i];
    // what mimsem_gpu_set_topo will insist on: an edge row belongs to at most two elements of the context
    std::vector<int> uses(T.n1, 0);
    for (size_t i = 0; i < T.e1x.size(); i++) uses[T.e1x[i]]++;
    for (size_t i = 0; i < T.e1y.size(); i++) uses[T.e1y[i]]++;
    for (int i = 0; i < T.n1; i++)
        if (uses[i] > 2) return 2;
    return 0;
}

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
    // MatAXPY(this, a, other): the terms added since the last assemble() (the reference adds the Rayleigh friction matrix to
    // the assembled M1 this way, eul/Euler_2.cpp:1229, 1448); applied with the operator itself in MatMult
    std::vector<std::pair<double, MimsemShell*> > plus;
    std::vector<double> plus_tmp;
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
    {
        DeviceTimer timer;
        if (mimsem_gpu_apply_host_up(p->ctx, s->op, s->lev, 1, s->scale, s->tpow, s->flags, s->coeff.empty() ? NULL : s->coeff.data(),
                                     s->u1.empty() ? NULL : s->u1.data(), s->tau, xa, ya))
            die("mimsem_gpu_apply_host");
        // terms added with MatAXPY: same ghosted input, summed into the local result before the one shared-DOF sum
        const int nloc = space_size_local(topo, s->sout);
        for (size_t t = 0; t < s->plus.size(); t++) {
            MimsemShell* o = s->plus[t].second;
            if (o->tpow > 0) sync_thickness(p);
            s->plus_tmp.resize(nloc);
            if (mimsem_gpu_apply_host_up(p->ctx, o->op, o->lev, 1, o->scale, o->tpow, o->flags, o->coeff.empty() ? NULL : o->coeff.data(),
                                         o->u1.empty() ? NULL : o->u1.data(), o->tau, xa, s->plus_tmp.data()))
                die("mimsem_gpu_apply_host (MatAXPY term)");
            for (int i = 0; i < nloc; i++) ya[i] += s->plus[t].first * s->plus_tmp[i];
        }
    }
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

// z = blockdiag(M1)^-1 r with the reference's element blocks (PCBJACOBI, eul/HorizSolve.cpp:77-84) for the Umat shell:
// ghosted local r -> block solves on the device -> owned rows of z.  Nonzero (PETSC_ERR_SUP) when A is not such a shell.
PetscErrorCode shell_pcbjacobi(Mat A, Vec r, Vec z) {
    MimsemShell* s;
    MatShellGetContext(A, &s);
    if (s && (s->op == 2 /* M0 */ || s->op == 6 /* M0(h) */)) {
        // diagonal when the quadrature order equals the element order (eul/Assembly.cpp:602-628: the reference keeps it as a
        // vector for that reason): block Jacobi with any blocks is the exact inverse, KSPSolve(ksp0, ...) of
        // eul/HorizSolve.cpp:87-96 converges in one iteration
        Vec d;
        VecDuplicate(r, &d);
        PetscErrorCode rc = shell_getdiag(A, d);
        if (!rc) VecPointwiseDivide(z, r, d);
        VecDestroy(&d);
        return rc;
    }
    if (!s || s->op != 0 /* M1 */ || (s->flags & MIMSEM_THICK_MEAN)) return 56;
    Topo* topo = s->topo;
    std::map<Topo*, Patch>::iterator it = g_patches.find(topo);
    if (it == g_patches.end() || !it->second.ctx) return 56;
    Patch* p = &it->second;
    mimsem_gpu_ctx* pc = pc_patch(p, topo);
    if (!pc) return 56;
    if (s->tpow > 0) sync_pc_thickness(p);
    VecScatterBegin(topo->gtol_1, r, s->xl, INSERT_VALUES, SCATTER_FORWARD);
    VecScatterEnd(topo->gtol_1, r, s->xl, INSERT_VALUES, SCATTER_FORWARD);
    PetscScalar *xa, *ya;
    VecGetArray(s->xl, &xa);
    std::memcpy(p->pc_x.data(), xa, sizeof(double) * topo->n1);   // rows beyond n1 belong to the neighbour copies: never read
    VecRestoreArray(s->xl, &xa);
    if (mimsem_gpu_apply_host_up(pc, 20, s->lev, 1, s->scale, s->tpow, s->flags, NULL, NULL, 0.0, p->pc_x.data(), p->pc_y.data()))
        die("mimsem_gpu_apply_host (element-block Jacobi)");
    // the blocks cover exactly the edges the patch's elements own; its east / north ghost rows stay zero
    VecGetArray(s->yl, &ya);
    std::memcpy(ya, p->pc_y.data(), sizeof(double) * topo->n1);
    VecRestoreArray(s->yl, &ya);
    VecZeroEntries(z);
    VecScatterBegin(topo->gtol_1, s->yl, z, ADD_VALUES, SCATTER_REVERSE);
    VecScatterEnd(topo->gtol_1, s->yl, z, ADD_VALUES, SCATTER_REVERSE);
    return 0;
}

// MatAXPY(Y, a, X, ...) on two shells of this library over the same spaces: Y's MatMult becomes (Op_Y + a Op_X) until Y's
// next assemble() -- the reference's assemble() starts with MatZeroEntries, which drops what was added
PetscErrorCode shell_axpy(Mat Y, PetscScalar a, Mat X, MatStructure) {
    MimsemShell *y = NULL, *x = NULL;
    MatShellGetContext(Y, &y);
    MatShellGetContext(X, &x);
    if (!y || !x || y == x || y->topo != x->topo || y->sin != x->sin || y->sout != x->sout) return 56;
    y->plus.push_back(std::make_pair((double)a, x));
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
    MatShellSetOperation(*M, MATOP_AXPY, (void (*)(void))shell_axpy);
#ifndef MIMSEM_HAVE_PETSC
    // the compatibility layer's KSP asks the operator itself for the blocks of a PCBJACOBI request (real PETSc: MimsemPCApplyBJacobi)
    if (op == 0 || op == 2 || op == 6) MatShellSetOperation(*M, MATOP_COMPAT_PCBJACOBI, (void (*)(void))shell_pcbjacobi);
#endif
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

// y[k] = Op(level lev0 + k) x[k] for nlev consecutive levels in ONE device call: what the reference writes as
//   for (kk = 0; kk < nk; kk++) { M1->assemble(kk, SCALE, true); MatMult(M1->M, velx[kk], Mu[kk]); }      (eul/Euler_2.cpp:1427-1456)
// The ghost refresh and the shared-DOF sum stay PETSc's, level by level; between them all levels cross PCIe and the
// kernels as one pipelined mimsem_gpu_apply_host call (the shell's own MatMult pays a launch, two copies and a
// synchronisation per level).  Scale, thickness power, flags (and the operator) are the shell's current ones; coeff[k]:
// the coefficient of level k in the convention of the class's assemble() (NULL for M1 / M2 / M0 and the incidence shells).
PetscErrorCode MimsemMatMultLevels(Mat A, int lev0, int nlev, Vec* x, Vec* y, Vec* coeff) {
    MimsemShell* s = NULL;
    MatShellGetContext(A, &s);
    if (!s || nlev < 1 || !x || !y) return 62;   // PETSC_ERR_ARG_WRONG
    if (!s->u1.empty() || !s->plus.empty() || s->op == 21 /* Umat_ray */ || s->op == OP_M2INV || s->op == OP_M2HINV) return 56;
    const bool need_c = !s->coeff.empty();
    if (need_c && !coeff) return 62;
    Topo* topo = s->topo;
    Patch* p = patch_of(topo);
    if (s->tpow > 0) sync_thickness(p);
    const size_t nin = space_size_local(topo, s->sin), nout = space_size_local(topo, s->sout), nc = s->coeff.size();
    const size_t want[3] = {nin * nlev, nout * nlev, nc * nlev};
    for (int b = 0; b < 3; b++)
        if (want[b] > p->stage_cap[b]) {
            if (p->stage[b]) mimsem_gpu_host_free(p->ctx, p->stage[b]);
            if (mimsem_gpu_host_alloc(p->ctx, (int64_t)want[b] * 8, (void**)&p->stage[b])) die("mimsem_gpu_host_alloc");
            p->stage_cap[b] = want[b];
        }
    PetscScalar* a;
    for (int k = 0; k < nlev; k++) {
        if (s->sin == 2) {
            VecGetArray(x[k], &a);
            std::memcpy(p->stage[0] + k * nin, a, 8 * nin);
            VecRestoreArray(x[k], &a);
        } else {
            VecScatter sc = s->sin == 0 ? topo->gtol_0 : topo->gtol_1;
            VecScatterBegin(sc, x[k], s->xl, INSERT_VALUES, SCATTER_FORWARD);
            VecScatterEnd(sc, x[k], s->xl, INSERT_VALUES, SCATTER_FORWARD);
            VecGetArray(s->xl, &a);
            std::memcpy(p->stage[0] + k * nin, a, 8 * nin);
            VecRestoreArray(s->xl, &a);
        }
        if (need_c) {
            VecGetArray(coeff[k], &a);
            std::memcpy(p->stage[2] + k * nc, a, 8 * nc);
            VecRestoreArray(coeff[k], &a);
        }
    }
    {
        DeviceTimer timer;
        if (mimsem_gpu_apply_host_up(p->ctx, s->op, lev0, nlev, s->scale, s->tpow, s->flags, need_c ? p->stage[2] : NULL, NULL, 0.0,
                                     p->stage[0], p->stage[1]))
            die("mimsem_gpu_apply_host (all levels)");
    }
    for (int k = 0; k < nlev; k++) {
        if (s->sout == 2) {
            VecGetArray(y[k], &a);
            std::memcpy(a, p->stage[1] + k * nout, 8 * nout);
            VecRestoreArray(y[k], &a);
        } else {
            VecGetArray(s->yl, &a);
            std::memcpy(a, p->stage[1] + k * nout, 8 * nout);
            VecRestoreArray(s->yl, &a);
            VecScatter sc = s->sout == 0 ? topo->gtol_0 : topo->gtol_1;
            VecZeroEntries(y[k]);
            VecScatterBegin(sc, s->yl, y[k], ADD_VALUES, SCATTER_REVERSE);
            VecScatterEnd(sc, s->yl, y[k], ADD_VALUES, SCATTER_REVERSE);
        }
    }
    return 0;
}

PetscErrorCode MimsemPCApplyBJacobi(PC pc, Vec r, Vec z) {
    Mat M = NULL;
    PCShellGetContext(pc, &M);
    return M ? shell_pcbjacobi(M, r, z) : 56;
}
PetscErrorCode MimsemKSPSetElementBlockJacobi(KSP ksp, Mat M) {
    MimsemShell* s = NULL;
    MatShellGetContext(M, &s);
    if (!s || (s->op != 0 /* M1 */ && s->op != 2 /* M0 */ && s->op != 6 /* M0(h) */)) return 56;
    PC pc;
    KSPGetPC(ksp, &pc);
    PCSetType(pc, PCSHELL);
    PCShellSetContext(pc, M);
    PCShellSetApply(pc, MimsemPCApplyBJacobi);
    return 0;
}

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
    sh->plus.clear();
    sh->lev = lev;
    sh->scale = scale;
    sh->tpow = vert_scale ? 1 : 0;
}
void Umat::assemble() {
    sh->plus.clear();
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
    vg1 = NULL;
    if (topo->kind == MIMSEM_MESH_BOX) {
        // box/Assembly.cpp:357-372: the constructor assembles level 0 twice, vg with SCALE and vg1 with scale 1; assemble is private there
        VecCreateMPI(MPI_COMM_WORLD, topo->n0l, topo->nDofs0G, &vg1);
        Vec keep = vg;
        vg = vg1;
        assemble(0, 1.0);
        vg = keep;
        assemble(0, SCALE);
    }
}
Pvec::~Pvec() {
    VecDestroy(&vl);
    VecDestroy(&vg);
    if (vg1) VecDestroy(&vg1);
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

namespace {

// ghosted local copy (Geom's quadrature-point numbering) of component c of a global quadrature-point vector with
// ncomp interleaved components per point
void quad_values_local(Geom* geom, Vec x, int ncomp, int c, std::vector<double>& out) {
    // global index of component c of local point i: ncomp * loc0[i] + c (eul/Assembly.cpp, UtQmat::assemble: inds_0x = 2 inds_0 + 0)
    std::vector<PetscInt> idx(geom->n0);
    for (int i = 0; i < geom->n0; i++) idx[i] = ncomp * geom->loc0[i] + c;
    IS isg, isl;
    Vec tmp;
    VecScatter sc;
    ISCreateGeneral(MPI_COMM_WORLD, geom->n0, idx.data(), PETSC_COPY_VALUES, &isg);
    ISCreateStride(MPI_COMM_SELF, geom->n0, 0, 1, &isl);
    VecCreateSeq(MPI_COMM_SELF, geom->n0, &tmp);
    VecScatterCreate(x, isg, tmp, isl, &sc);
    VecScatterBegin(sc, x, tmp, INSERT_VALUES, SCATTER_FORWARD);
    VecScatterEnd(sc, x, tmp, INSERT_VALUES, SCATTER_FORWARD);
    PetscScalar* xa;
    VecGetArray(tmp, &xa);
    out.assign(xa, xa + geom->n0);
    VecRestoreArray(tmp, &xa);
    VecScatterDestroy(&sc);
    ISDestroy(&isg);
    ISDestroy(&isl);
    VecDestroy(&tmp);
}

// eul/Assembly.cpp, WtQmat::assemble: M = sum over elements of W^T diag(w_q); faces are element-interior (no ghosts)
PetscErrorCode wtq_mult(Mat A, Vec x, Vec y) {
    WtQmat* self;
    MatShellGetContext(A, &self);
    Topo* topo = self->topo;
    Geom* geom = self->geom;
    const int mp1 = geom->quad->n + 1, mp12 = mp1 * mp1;
    std::vector<double> xl;
    quad_values_local(geom, x, 1, 0, xl);
    M2_j_xy_i W(self->e);
    PetscScalar* ya;
    VecZeroEntries(y);
    VecGetArray(y, &ya);
    for (int ey = 0; ey < topo->nElsX; ey++)
        for (int ex = 0; ex < topo->nElsX; ex++) {
            const int* iq = geom->elInds0_l(ex, ey);
            const int* i2 = topo->elInds2_l(ex, ey);
            for (int j = 0; j < W.nDofsJ; j++) {
                double s = 0.0;
                for (int q = 0; q < mp12; q++) s += W.A[q * W.nDofsJ + j] * geom->quad->w[q % mp1] * geom->quad->w[q / mp1] * xl[iq[q]];
                ya[i2[j]] += s;
            }
        }
    VecRestoreArray(y, &ya);
    return 0;
}

// eul/Assembly.cpp, UtQmat::assemble: x-edges take U^T diag(w_q) (J00 u_x + J10 u_y), y-edges V^T diag(w_q) (J01 u_x + J11 u_y)
PetscErrorCode utq_mult(Mat A, Vec x, Vec y) {
    UtQmat* self;
    MatShellGetContext(A, &self);
    Topo* topo = self->topo;
    Geom* geom = self->geom;
    const int mp1 = geom->quad->n + 1, mp12 = mp1 * mp1;
    std::vector<double> ux, uy;
    quad_values_local(geom, x, 2, 0, ux);
    quad_values_local(geom, x, 2, 1, uy);
    M1x_j_xy_i U(self->l, self->e);
    M1y_j_xy_i V(self->l, self->e);
    Vec yv;
    VecCreateSeq(MPI_COMM_SELF, topo->n1, &yv);
    VecZeroEntries(yv);
    PetscScalar* ya;
    VecGetArray(yv, &ya);
    std::vector<double> fa(mp12), fb(mp12);
    for (int ey = 0; ey < topo->nElsX; ey++)
        for (int ex = 0; ex < topo->nElsX; ex++) {
            const int ei = ey * topo->nElsX + ex;
            const int* iq = geom->elInds0_l(ex, ey);
            const int* ix = topo->elInds1x_l(ex, ey);
            const int* iy = topo->elInds1y_l(ex, ey);
            for (int q = 0; q < mp12; q++) {
                const double w = geom->quad->w[q % mp1] * geom->quad->w[q / mp1];
                double** J = geom->J[ei][q];
                fa[q] = w * (J[0][0] * ux[iq[q]] + J[1][0] * uy[iq[q]]);
                fb[q] = w * (J[0][1] * ux[iq[q]] + J[1][1] * uy[iq[q]]);
            }
            for (int j = 0; j < U.nDofsJ; j++) {
                double sx = 0.0, sy = 0.0;
                for (int q = 0; q < mp12; q++) {
                    sx += U.A[q * U.nDofsJ + j] * fa[q];
                    sy += V.A[q * V.nDofsJ + j] * fb[q];
                }
                ya[ix[j]] += sx;
                ya[iy[j]] += sy;
            }
        }
    VecRestoreArray(yv, &ya);
    VecZeroEntries(y);
    VecScatterBegin(topo->gtol_1, yv, y, ADD_VALUES, SCATTER_REVERSE);
    VecScatterEnd(topo->gtol_1, yv, y, ADD_VALUES, SCATTER_REVERSE);
    VecDestroy(&yv);
    return 0;
}

}  // namespace

WtQmat::WtQmat(Topo* _topo, Geom* _geom, LagrangeEdge* _e) : topo(_topo), geom(_geom), e(_e) {
    MatCreateShell(MPI_COMM_WORLD, topo->n2l, geom->n0l, topo->nDofs2G, geom->nDofs0G, this, &M);
    MatShellSetOperation(M, MATOP_MULT, (void (*)(void))wtq_mult);
}
WtQmat::~WtQmat() { MatDestroy(&M); }
void WtQmat::assemble() {}

UtQmat::UtQmat(Topo* _topo, Geom* _geom, LagrangeNode* _l, LagrangeEdge* _e) : topo(_topo), geom(_geom), l(_l), e(_e) {
    MatCreateShell(MPI_COMM_WORLD, topo->n1l, 2 * geom->n0l, topo->nDofs1G, 2 * geom->nDofs0G, this, &M);
    MatShellSetOperation(M, MATOP_MULT, (void (*)(void))utq_mult);
}
UtQmat::~UtQmat() { MatDestroy(&M); }
void UtQmat::assemble() {}

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
