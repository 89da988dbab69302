// Host-only part of the C ABI (include/mimsem_gpu.h): basis, element tabulations, topology, mesh.
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mimsem_gpu.h"
#include "basis.hpp"
#include "errors.hpp"
#include "mesh.hpp"

namespace mimsem {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* get_error() { return g_err.c_str(); }
}  // namespace mimsem

using namespace mimsem;

struct mimsem_mesh {
    GlobalMesh g;
};

namespace {
int fail(int code, const std::string& msg) {
    set_error(msg);
    return code;
}
bool kind_ok(int kind) { return kind == MIMSEM_MESH_SPHERE || kind == MIMSEM_MESH_BOX; }
}  // namespace

extern "C" {

const char* mimsem_last_error(void) { return get_error(); }

int mimsem_basis_gll(int n, double* x, double* w) {
    std::vector<double> vx, vw;
    if (!x || !w) return fail(MIMSEM_ERR_ARG, "null argument");
    if (!gll_rule(n, vx, vw)) return fail(MIMSEM_ERR_ARG, "invalid gauss-lobatto quadrature order");
    std::memcpy(x, vx.data(), vx.size() * sizeof(double));
    std::memcpy(w, vw.data(), vw.size() * sizeof(double));
    return MIMSEM_OK;
}

int mimsem_basis_tables(int p, int m, double* ljxi, double* ejxi) {
    BasisTables b;
    if (!b.build(p, m)) return fail(MIMSEM_ERR_ARG, "invalid basis / quadrature order");
    if (ljxi) std::memcpy(ljxi, b.ljxi.data(), b.ljxi.size() * sizeof(double));
    if (ejxi) std::memcpy(ejxi, b.ejxi.data(), b.ejxi.size() * sizeof(double));
    return MIMSEM_OK;
}

int mimsem_elmat(int which, int p, int m, double* A) {
    BasisTables b;
    if (!A) return fail(MIMSEM_ERR_ARG, "null argument");
    if (!b.build(p, m)) return fail(MIMSEM_ERR_ARG, "invalid basis / quadrature order");
    const int mp1 = m + 1, np1 = p + 1, q2 = mp1 * mp1;
    auto L = [&](int q, int j) { return b.ljxi[(size_t)q * np1 + j]; };
    auto E = [&](int q, int i) { return b.ejxi[(size_t)q * p + i]; };
    switch (which) {
        case 0:   // U[q][j] = l_jx(qx) e_jy(qy), j = jy*(p+1)+jx       (eul/ElMats.cpp:38-44)
            for (int q = 0; q < q2; q++)
                for (int j = 0; j < p * np1; j++) A[(size_t)q * p * np1 + j] = L(q % mp1, j % np1) * E(q / mp1, j / np1);
            break;
        case 1:   // V[q][j] = e_jx(qx) l_jy(qy), j = jy*p+jx           (eul/ElMats.cpp:73-79)
            for (int q = 0; q < q2; q++)
                for (int j = 0; j < p * np1; j++) A[(size_t)q * p * np1 + j] = E(q % mp1, j % p) * L(q / mp1, j / p);
            break;
        case 2:   // W[q][j] = e_jx(qx) e_jy(qy)                        (eul/ElMats.cpp:105-111)
            for (int q = 0; q < q2; q++)
                for (int j = 0; j < p * p; j++) A[(size_t)q * p * p + j] = E(q % mp1, j % p) * E(q / mp1, j / p);
            break;
        case 3:   // P[q][j] = l_jx(qx) l_jy(qy)                        (eul/ElMats.cpp:135-141)
            for (int q = 0; q < q2; q++)
                for (int j = 0; j < np1 * np1; j++) A[(size_t)q * np1 * np1 + j] = L(q % mp1, j % np1) * L(q / mp1, j / np1);
            break;
        case 4:   // Wii diagonal                                        (eul/ElMats.cpp:177)
            for (int q = 0; q < q2; q++) A[q] = b.qw[q % mp1] * b.qw[q / mp1];
            break;
        default: return fail(MIMSEM_ERR_ARG, "unknown element matrix id");
    }
    return MIMSEM_OK;
}

int mimsem_topo_patch_sizes(int kind, int order, int ne, int nprocs, int rank, int out[8]) {
    if (!kind_ok(kind) || !out) return fail(MIMSEM_ERR_ARG, "bad argument");
    PatchTopo t;
    std::string err;
    if (!patch_topology((MeshKind)kind, order, ne, nprocs, rank, t, &err)) return fail(MIMSEM_ERR_ARG, err);
    const int v[8] = {t.n0, t.n1x, t.n1y, t.n2, t.n0l, t.n1xl, t.n1yl, t.n2l};
    std::memcpy(out, v, sizeof v);
    return MIMSEM_OK;
}

int mimsem_topo_patch(int kind, int order, int ne, int nprocs, int rank, int* loc0, int* loc1x, int* loc1y, int* loc2) {
    if (!kind_ok(kind)) return fail(MIMSEM_ERR_ARG, "bad mesh kind");
    PatchTopo t;
    std::string err;
    if (!patch_topology((MeshKind)kind, order, ne, nprocs, rank, t, &err)) return fail(MIMSEM_ERR_ARG, err);
    if (loc0) std::memcpy(loc0, t.loc0.data(), t.loc0.size() * sizeof(int));
    if (loc1x) std::memcpy(loc1x, t.loc1x.data(), t.loc1x.size() * sizeof(int));
    if (loc1y) std::memcpy(loc1y, t.loc1y.data(), t.loc1y.size() * sizeof(int));
    if (loc2) std::memcpy(loc2, t.loc2.data(), t.loc2.size() * sizeof(int));
    return MIMSEM_OK;
}

int mimsem_topo_write_input(int kind, int p, int m, int ne, int nprocs, const char* dir) {
    if (!kind_ok(kind) || !dir) return fail(MIMSEM_ERR_ARG, "bad argument");
    std::string err;
    if (!write_input_files((MeshKind)kind, p, m, ne, nprocs, dir, &err)) return fail(MIMSEM_ERR_IO, err);
    return MIMSEM_OK;
}

int mimsem_mesh_create(int kind, int p, int m, int ne, int signed_det, mimsem_mesh** out) {
    if (!kind_ok(kind) || !out) return fail(MIMSEM_ERR_ARG, "bad argument");
    *out = nullptr;
    mimsem_mesh* h = new mimsem_mesh;
    std::string err;
    if (!build_global_mesh((MeshKind)kind, p, m, ne, signed_det != 0, h->g, &err)) {
        delete h;
        return fail(MIMSEM_ERR_ARG, err);
    }
    *out = h;
    return MIMSEM_OK;
}

void mimsem_mesh_destroy(mimsem_mesh* mesh) { delete mesh; }

int mimsem_mesh_sizes(const mimsem_mesh* mesh, int64_t out[8]) {
    if (!mesh || !out) return fail(MIMSEM_ERR_ARG, "null argument");
    const GlobalMesh& g = mesh->g;
    const int64_t v[8] = {g.p, g.m, g.ne, g.nel, g.N0, g.N1, g.N2, g.NQ};
    std::memcpy(out, v, sizeof v);
    return MIMSEM_OK;
}

int mimsem_mesh_tables(const mimsem_mesh* mesh, int* el0, int* el1x, int* el1y, int* el2, int* elq) {
    if (!mesh) return fail(MIMSEM_ERR_ARG, "null argument");
    const GlobalMesh& g = mesh->g;
    if (el0) std::memcpy(el0, g.el0.data(), g.el0.size() * sizeof(int));
    if (el1x) std::memcpy(el1x, g.el1x.data(), g.el1x.size() * sizeof(int));
    if (el1y) std::memcpy(el1y, g.el1y.data(), g.el1y.size() * sizeof(int));
    if (el2) std::memcpy(el2, g.el2.data(), g.el2.size() * sizeof(int));
    if (elq) std::memcpy(elq, g.elq.data(), g.elq.size() * sizeof(int));
    return MIMSEM_OK;
}

int mimsem_mesh_geometry(const mimsem_mesh* mesh, double* J, double* det) {
    if (!mesh) return fail(MIMSEM_ERR_ARG, "null argument");
    const GlobalMesh& g = mesh->g;
    if (J) std::memcpy(J, g.J.data(), g.J.size() * sizeof(double));
    if (det) std::memcpy(det, g.det.data(), g.det.size() * sizeof(double));
    return MIMSEM_OK;
}

int mimsem_mesh_coords(const mimsem_mesh* mesh, double* xyz) {
    if (!mesh || !xyz) return fail(MIMSEM_ERR_ARG, "null argument");
    std::memcpy(xyz, mesh->g.xyz.data(), mesh->g.xyz.size() * sizeof(double));
    return MIMSEM_OK;
}

}  // extern "C"
