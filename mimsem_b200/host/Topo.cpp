#include "Topo.h"

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/mimsem_gpu.h"
#include "../csrc/mesh.hpp"

using mimsem::PatchTopo;

namespace {
int* dup(const std::vector<int>& v) {
    int* p = new int[v.size() ? v.size() : 1];
    for (size_t i = 0; i < v.size(); i++) p[i] = v[i];
    return p;
}
void adopt(Topo* t, const PatchTopo& p) {
    t->elOrd = p.p;
    t->nElsX = p.nelx;
    t->nDofsX = p.nx;
    t->n0 = p.n0; t->n1x = p.n1x; t->n1y = p.n1y; t->n2 = p.n2;
    t->n1 = p.n1x + p.n1y;
    t->n0l = p.n0l; t->n1xl = p.n1xl; t->n1yl = p.n1yl; t->n2l = p.n2l;
    t->n1l = p.n1xl + p.n1yl;
    t->loc0 = dup(p.loc0);
    t->loc1x = dup(p.loc1x);
    t->loc1y = dup(p.loc1y);
    t->loc2 = dup(p.loc2);
    // x- and y-normal edges interleaved into one local 1-form numbering (eul/Topo.cpp:82-86)
    t->loc1 = new int[t->n1];
    for (int i = 0; i < p.n1x; i++) {
        t->loc1[2 * i] = p.loc1x[i];
        t->loc1[2 * i + 1] = p.loc1y[i];
    }
}
}  // namespace

void Topo::finish(int nprocs) {
    const int np1 = elOrd + 1;
    inds0_l = new int[np1 * np1];  inds0_g = new int[np1 * np1];
    inds1x_l = new int[elOrd * np1]; inds1x_g = new int[elOrd * np1];
    inds1y_l = new int[elOrd * np1]; inds1y_g = new int[elOrd * np1];
    inds2_l = new int[elOrd * elOrd]; inds2_g = new int[elOrd * elOrd];
    // eul/Topo.cpp:113-115 ; box/Topo.cpp:112 has no hanging nodes
    nDofs2G = nprocs * nDofsX * nDofsX;
    nDofs1G = 2 * nDofs2G;
    nDofs0G = nDofs2G + (kind == MIMSEM_MESH_SPHERE ? 2 : 0);
    ISCreateGeneral(MPI_COMM_WORLD, n0, loc0, PETSC_COPY_VALUES, &is_g_0);
    ISCreateGeneral(MPI_COMM_WORLD, n1, loc1, PETSC_COPY_VALUES, &is_g_1);
    ISCreateStride(MPI_COMM_SELF, n0, 0, 1, &is_l_0);
    ISCreateStride(MPI_COMM_SELF, n1, 0, 1, &is_l_1);
    Vec vl, vg;
    VecCreateSeq(MPI_COMM_SELF, n0, &vl);
    VecCreateMPI(MPI_COMM_WORLD, n0l, nDofs0G, &vg);
    VecScatterCreate(vg, is_g_0, vl, is_l_0, &gtol_0);
    VecDestroy(&vl);
    VecDestroy(&vg);
    VecCreateSeq(MPI_COMM_SELF, n1, &vl);
    VecCreateMPI(MPI_COMM_WORLD, n1l, nDofs1G, &vg);
    VecScatterCreate(vg, is_g_1, vl, is_l_1, &gtol_1);
    VecDestroy(&vl);
    VecDestroy(&vg);
}

static void load_from_input(Topo* t, int kind) {
    int nprocs;
    MPI_Comm_rank(MPI_COMM_WORLD, &t->pi);
    MPI_Comm_size(MPI_COMM_WORLD, &nprocs);
    PatchTopo p;
    std::string err;
    if (!mimsem::load_patch_files("input", nprocs, t->pi, (mimsem::MeshKind)kind, p, &err)) {
        std::fprintf(stderr, "Topo: %s\n", err.c_str());
        std::abort();
    }
    adopt(t, p);
}

Topo::Topo(int _nk) : nk(_nk), kind(MIMSEM_MESH_SPHERE) {
    int nprocs;
    MPI_Comm_size(MPI_COMM_WORLD, &nprocs);
    load_from_input(this, kind);
    finish(nprocs);
}

Topo::Topo() : nk(1), kind(MIMSEM_MESH_SPHERE) {
    int nprocs;
    MPI_Comm_size(MPI_COMM_WORLD, &nprocs);
    // box/ has a perfect-square rank count and no quads_* files; src/ has 6 n^2 ranks (scr/Setup_Box.py vs scr/Setup.py)
    FILE* f = std::fopen("input/grid_res_quad.txt", "r");
    if (f) std::fclose(f);
    else kind = MIMSEM_MESH_BOX;
    load_from_input(this, kind);
    finish(nprocs);
}

Topo::Topo(int _kind, int p, int ne, int _nk) : nk(_nk), kind(_kind) {
    int nprocs;
    MPI_Comm_rank(MPI_COMM_WORLD, &pi);
    MPI_Comm_size(MPI_COMM_WORLD, &nprocs);
    PatchTopo t;
    std::string err;
    if (!mimsem::patch_topology((mimsem::MeshKind)kind, p, ne, nprocs, pi, t, &err)) {
        std::fprintf(stderr, "Topo: %s\n", err.c_str());
        std::abort();
    }
    adopt(this, t);
    finish(nprocs);
}

Topo::~Topo() {
    delete[] loc0; delete[] loc1; delete[] loc1x; delete[] loc1y; delete[] loc2;
    delete[] inds0_l; delete[] inds1x_l; delete[] inds1y_l; delete[] inds2_l;
    delete[] inds0_g; delete[] inds1x_g; delete[] inds1y_g; delete[] inds2_g;
    ISDestroy(&is_g_0); ISDestroy(&is_g_1); ISDestroy(&is_l_0); ISDestroy(&is_l_1);
    VecScatterDestroy(&gtol_0);
    VecScatterDestroy(&gtol_1);
}

// element -> local indices (eul/Topo.cpp:200-251).  As in the reference these return pointers to member scratch.
int* Topo::elInds0_l(int ex, int ey) {
    int k = 0;
    for (int iy = 0; iy <= elOrd; iy++)
        for (int ix = 0; ix <= elOrd; ix++) inds0_l[k++] = (ey * elOrd + iy) * (nDofsX + 1) + ex * elOrd + ix;
    return inds0_l;
}
int* Topo::elInds1x_l(int ex, int ey) {
    int k = 0;
    for (int iy = 0; iy < elOrd; iy++)
        for (int ix = 0; ix <= elOrd; ix++) inds1x_l[k++] = 2 * ((ey * elOrd + iy) * (nDofsX + 1) + ex * elOrd + ix);
    return inds1x_l;
}
int* Topo::elInds1y_l(int ex, int ey) {
    int k = 0;
    for (int iy = 0; iy <= elOrd; iy++)
        for (int ix = 0; ix < elOrd; ix++) inds1y_l[k++] = 2 * ((ey * elOrd + iy) * nDofsX + ex * elOrd + ix) + 1;
    return inds1y_l;
}
int* Topo::elInds2_l(int ex, int ey) {
    const int n2e = elOrd * elOrd;
    for (int k = 0; k < n2e; k++) inds2_l[k] = (ey * nElsX + ex) * n2e + k;
    return inds2_l;
}
// element -> global indices (eul/Topo.cpp:253-305)
int* Topo::elInds0_g(int ex, int ey) {
    int k = 0;
    for (int iy = 0; iy <= elOrd; iy++)
        for (int ix = 0; ix <= elOrd; ix++) inds0_g[k++] = loc0[(ey * elOrd + iy) * (nDofsX + 1) + ex * elOrd + ix];
    return inds0_g;
}
int* Topo::elInds1x_g(int ex, int ey) {
    int k = 0;
    for (int iy = 0; iy < elOrd; iy++)
        for (int ix = 0; ix <= elOrd; ix++) inds1x_g[k++] = loc1x[(ey * elOrd + iy) * (nDofsX + 1) + ex * elOrd + ix];
    return inds1x_g;
}
int* Topo::elInds1y_g(int ex, int ey) {
    int k = 0;
    for (int iy = 0; iy <= elOrd; iy++)
        for (int ix = 0; ix < elOrd; ix++) inds1y_g[k++] = loc1y[(ey * elOrd + iy) * nDofsX + ex * elOrd + ix];
    return inds1y_g;
}
int* Topo::elInds2_g(int ex, int ey) {
    elInds2_l(ex, ey);
    for (int k = 0; k < elOrd * elOrd; k++) inds2_g[k] = inds2_l[k] + pi * n2;
    return inds2_g;
}
