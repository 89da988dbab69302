#include "Geom.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/mimsem_gpu.h"
#include "../csrc/mesh.hpp"

Geom::Geom(Topo* _topo, int _nk) : pi(_topo->pi), nk(_nk), topo(_topo), thick_version(0) { build(false); }
Geom::Geom(Topo* _topo) : pi(_topo->pi), nk(1), topo(_topo), thick_version(0) { build(true); }

void Geom::build(bool signed_det) {
    int nprocs;
    MPI_Comm_size(MPI_COMM_WORLD, &nprocs);
    const int kind = topo->kind;
    // quadrature order: input/grid_res_quad.txt when present (eul/Geom.cpp:39-47), else the element order (box/Geom.cpp:33)
    int quad_ord = topo->elOrd;
    {
        std::ifstream f("input/grid_res_quad.txt");
        std::string line;
        if (f && std::getline(f, line)) quad_ord = std::atoi(line.c_str());
    }
    quad = new GaussLobatto(quad_ord);
    node = new LagrangeNode(topo->elOrd, quad);
    edge = new LagrangeEdge(topo->elOrd, node);
    const int m = quad_ord, nelx = topo->nElsX;
    nDofsX = m * nelx;
    n0 = (nDofsX + 1) * (nDofsX + 1);
    nl = n0;
    const int npx = (kind == MIMSEM_MESH_SPHERE) ? (int)std::lround(std::sqrt(nprocs / 6.0)) : (int)std::lround(std::sqrt((double)nprocs));
    const int ne_side = nelx * npx;

    // quadrature-point maps and coordinates: from input/ when the files exist, else generated
    mimsem::PatchTopo q;
    std::string err;
    std::vector<double> xl((size_t)n0 * 3);
    char fn[64];
    std::snprintf(fn, sizeof fn, "input/geom_%04d.txt", pi);
    std::ifstream gf(fn);
    if (!mimsem::patch_topology((mimsem::MeshKind)kind, m, ne_side, nprocs, pi, q, &err)) {
        std::fprintf(stderr, "Geom: %s\n", err.c_str());
        std::abort();
    }
    if (gf) {
        std::string line;
        int i = 0;
        while (i < n0 && std::getline(gf, line)) {
            std::stringstream ss(line);
            ss >> xl[(size_t)i * 3] >> xl[(size_t)i * 3 + 1] >> xl[(size_t)i * 3 + 2];
            i++;
        }
        if (i != n0) std::fprintf(stderr, "ERROR! geometry file reading: %d\n", i);
    } else {
        std::vector<double> xyz;
        if (kind == MIMSEM_MESH_SPHERE) mimsem::sphere_node_coords(m, ne_side, 6371220.0, xyz);
        else mimsem::box_node_coords(m, ne_side, 1000.0, xyz);
        for (int i = 0; i < n0; i++)
            for (int a = 0; a < 3; a++) xl[(size_t)i * 3 + a] = xyz[(size_t)q.loc0[i] * 3 + a];
    }
    n0l = q.n0l;
    nDofs0G = (int)q.N0;
    loc0 = new int[n0];
    for (int i = 0; i < n0; i++) loc0[i] = q.loc0[i];
    {
        Vec vl, vg;
        ISCreateGeneral(MPI_COMM_WORLD, n0, loc0, PETSC_COPY_VALUES, &is_g_0);
        ISCreateStride(MPI_COMM_SELF, n0, 0, 1, &is_l_0);
        VecCreateSeq(MPI_COMM_SELF, n0, &vl);
        VecCreateMPI(MPI_COMM_WORLD, n0l, nDofs0G, &vg);
        VecScatterCreate(vg, is_g_0, vl, is_l_0, &gtol_0);
        VecDestroy(&vl);
        VecDestroy(&vg);
    }
    inds0_l = new int[(m + 1) * (m + 1)];
    inds0_g = new int[(m + 1) * (m + 1)];

    std::vector<double> Jv, dv, ll;
    mimsem::patch_geometry((mimsem::MeshKind)kind, m, nelx, ne_side, 6371220.0, 1000.0, signed_det, xl, Jv, dv, &ll);
    const int nel = nelx * nelx, mp12 = (m + 1) * (m + 1);
    Jflat = new double[(size_t)nel * mp12 * 4];
    detflat = new double[(size_t)nel * mp12];
    for (size_t i = 0; i < Jv.size(); i++) Jflat[i] = Jv[i];
    for (size_t i = 0; i < dv.size(); i++) detflat[i] = dv[i];
    x = new double*[nl];
    s = new double*[nl];
    for (int i = 0; i < nl; i++) {
        x[i] = new double[3];
        s[i] = new double[2];
        for (int a = 0; a < 3; a++) x[i][a] = xl[(size_t)i * 3 + a];
        s[i][0] = ll.empty() ? 0.0 : ll[(size_t)i * 2];
        s[i][1] = ll.empty() ? 0.0 : ll[(size_t)i * 2 + 1];
    }
    det = new double*[nel];
    J = new double***[nel];
    for (int e = 0; e < nel; e++) {
        det[e] = detflat + (size_t)e * mp12;
        J[e] = new double**[mp12];
        for (int k = 0; k < mp12; k++) {
            J[e][k] = new double*[2];
            J[e][k][0] = Jflat + ((size_t)e * mp12 + k) * 4;
            J[e][k][1] = Jflat + ((size_t)e * mp12 + k) * 4 + 2;
        }
    }
    topog = new double[n0];
    levs = new double*[nk + 1];
    thick = new double*[nk];
    thickInv = new double*[nk];
    for (int k = 0; k <= nk; k++) levs[k] = new double[n0];
    for (int k = 0; k < nk; k++) {
        thick[k] = new double[n0];
        thickInv[k] = new double[n0];
        for (int i = 0; i < n0; i++) thick[k][i] = thickInv[k][i] = 1.0;
    }
}

Geom::~Geom() {
    const int nel = topo->nElsX * topo->nElsX, mp12 = (quad->n + 1) * (quad->n + 1);
    for (int e = 0; e < nel; e++) {
        for (int k = 0; k < mp12; k++) delete[] J[e][k];
        delete[] J[e];
    }
    delete[] J;
    delete[] det;
    delete[] Jflat;
    delete[] detflat;
    for (int i = 0; i < nl; i++) {
        delete[] x[i];
        delete[] s[i];
    }
    delete[] x;
    delete[] s;
    delete[] loc0;
    ISDestroy(&is_l_0);
    ISDestroy(&is_g_0);
    VecScatterDestroy(&gtol_0);
    delete[] inds0_l;
    delete[] inds0_g;
    delete[] topog;
    for (int k = 0; k <= nk; k++) delete[] levs[k];
    for (int k = 0; k < nk; k++) {
        delete[] thick[k];
        delete[] thickInv[k];
    }
    delete[] levs;
    delete[] thick;
    delete[] thickInv;
    delete edge;
    delete node;
    delete quad;
}

// eul/Geom.cpp:743-764
void Geom::initTopog(TopogFunc* ft, LevelFunc* fl) {
    const double max_height = fl ? fl(x[0], nk) : 1.0;
    for (int i = 0; i < n0; i++) topog[i] = ft(x[i]);
    for (int k = 0; k <= nk; k++)
        for (int i = 0; i < n0; i++) levs[k][i] = (max_height - topog[i]) * fl(x[i], k) / max_height + topog[i];
    for (int k = 0; k < nk; k++)
        for (int i = 0; i < n0; i++) {
            thick[k][i] = levs[k + 1][i] - levs[k][i];
            thickInv[k][i] = 1.0 / thick[k][i];
        }
    thick_version++;
}

int* Geom::elInds0_l(int ex, int ey) {
    const int m = quad->n;
    int k = 0;
    for (int iy = 0; iy <= m; iy++)
        for (int ix = 0; ix <= m; ix++) inds0_l[k++] = (ey * m + iy) * (nDofsX + 1) + ex * m + ix;
    return inds0_l;
}
int* Geom::elInds0_g(int ex, int ey) {
    elInds0_l(ex, ey);
    for (int k = 0; k < (quad->n + 1) * (quad->n + 1); k++) inds0_g[k] = loc0[inds0_l[k]];
    return inds0_g;
}

// DOFs -> quadrature point (px, py) of element (ex, ey): eul/Geom.cpp:328-417
void Geom::interp0(int ex, int ey, int px, int py, double* vec, double* val) {
    const int np1 = node->n + 1;
    int* i0 = topo->elInds0_l(ex, ey);
    double v = 0.0;
    for (int j = 0; j < np1 * np1; j++) v += vec[i0[j]] * node->ljxi[px][j % np1] * node->ljxi[py][j / np1];
    val[0] = v;
}
void Geom::interp1_l(int ex, int ey, int px, int py, double* vec, double* val) {
    const int n = topo->elOrd, np1 = n + 1;
    int* ix = topo->elInds1x_l(ex, ey);
    int* iy = topo->elInds1y_l(ex, ey);
    double a = 0.0, b = 0.0;
    for (int j = 0; j < n * np1; j++) {
        a += vec[ix[j]] * node->ljxi[px][j % np1] * edge->ejxi[py][j / np1];
        b += vec[iy[j]] * edge->ejxi[px][j % n] * node->ljxi[py][j / n];
    }
    val[0] = a;
    val[1] = b;
}
void Geom::interp2_l(int ex, int ey, int px, int py, double* vec, double* val) {
    const int n = topo->elOrd;
    int* i2 = topo->elInds2_l(ex, ey);
    double v = 0.0;
    for (int j = 0; j < n * n; j++) v += vec[i2[j]] * edge->ejxi[px][j % n] * edge->ejxi[py][j / n];
    val[0] = v;
}
void Geom::interp1_g(int ex, int ey, int px, int py, double* vec, double* val) {
    const int el = ey * topo->nElsX + ex, q = py * (quad->n + 1) + px;
    double l[2];
    interp1_l(ex, ey, px, py, vec, l);
    double** jac = J[el][q];
    val[0] = (jac[0][0] * l[0] + jac[0][1] * l[1]) / det[el][q];
    val[1] = (jac[1][0] * l[0] + jac[1][1] * l[1]) / det[el][q];
}
void Geom::interp2_g(int ex, int ey, int px, int py, double* vec, double* val) {
    const int el = ey * topo->nElsX + ex, q = py * (quad->n + 1) + px;
    double l;
    interp2_l(ex, ey, px, py, vec, &l);
    val[0] = l / det[el][q];
}

// ------------------------------------------------------------------------------------------------
// field writers (eul/Geom.cpp:419-631, the non-HDF5 branch)

namespace {
void view_to(Vec v, const char* name, bool binary) {
    PetscViewer viewer;
    if (binary) PetscViewerBinaryOpen(MPI_COMM_WORLD, name, FILE_MODE_WRITE, &viewer);
    else PetscViewerASCIIOpen(MPI_COMM_WORLD, name, &viewer);
    VecView(v, viewer);
    PetscViewerDestroy(&viewer);
}
}  // namespace

// lev < 0: the src/ forms (no levels: no thickness factor, file names without the level; src/Geom.cpp:326-520)
void Geom::write0(Vec q, char* fieldname, int tstep) { write0(q, fieldname, tstep, -1); }
void Geom::write1(Vec u, char* fieldname, int tstep) { write1(u, fieldname, tstep, -1); }
void Geom::write2(Vec h, char* fieldname, int tstep) { write2(h, fieldname, tstep, -1, false); }

namespace {
void field_file(char* out, size_t n, const char* fieldname, const char* comp, int lev, int tstep, const char* ext) {
    if (lev >= 0) std::snprintf(out, n, "output/%s%s_%.3u_%.4u.%s", fieldname, comp, lev, tstep, ext);
    else std::snprintf(out, n, "output/%s%s_%.4u.%s", fieldname, comp, tstep, ext);
}
}  // namespace

void Geom::write0(Vec q, char* fieldname, int tstep, int lev) {
    const int mp1 = quad->n + 1, mp12 = mp1 * mp1;
    char filename[200];
    Vec ql, qxl, qxg;
    PetscScalar *qArray, *qxArray;
    VecCreateSeq(MPI_COMM_SELF, topo->n0, &ql);
    VecScatterBegin(topo->gtol_0, q, ql, INSERT_VALUES, SCATTER_FORWARD);
    VecScatterEnd(topo->gtol_0, q, ql, INSERT_VALUES, SCATTER_FORWARD);
    VecCreateSeq(MPI_COMM_SELF, n0, &qxl);
    VecCreateMPI(MPI_COMM_WORLD, n0l, nDofs0G, &qxg);
    VecGetArray(ql, &qArray);
    VecGetArray(qxl, &qxArray);
    for (int ey = 0; ey < topo->nElsX; ey++)
        for (int ex = 0; ex < topo->nElsX; ex++) {
            int* inds0 = elInds0_l(ex, ey);
            for (int ii = 0; ii < mp12; ii++) {
                double val;
                interp0(ex, ey, ii % mp1, ii / mp1, qArray, &val);
                qxArray[inds0[ii]] = lev >= 0 ? val / thick[lev][inds0[ii]] : val;   // piecewise constant in the vertical
            }
        }
    VecRestoreArray(ql, &qArray);
    VecRestoreArray(qxl, &qxArray);
    VecScatterBegin(gtol_0, qxl, qxg, INSERT_VALUES, SCATTER_REVERSE);
    VecScatterEnd(gtol_0, qxl, qxg, INSERT_VALUES, SCATTER_REVERSE);
    field_file(filename, sizeof filename, fieldname, "", lev, tstep, "dat");
    view_to(qxg, filename, false);
    VecDestroy(&ql);
    VecDestroy(&qxl);
    VecDestroy(&qxg);
}

void Geom::write1(Vec u, char* fieldname, int tstep, int lev) {
    const int mp1 = quad->n + 1, mp12 = mp1 * mp1;
    char filename[200];
    Vec ul, uxl, vxl, uxg;
    PetscScalar *uArray, *uxArray, *vxArray;
    VecCreateSeq(MPI_COMM_SELF, topo->n1, &ul);
    VecScatterBegin(topo->gtol_1, u, ul, INSERT_VALUES, SCATTER_FORWARD);
    VecScatterEnd(topo->gtol_1, u, ul, INSERT_VALUES, SCATTER_FORWARD);
    VecCreateSeq(MPI_COMM_SELF, n0, &uxl);
    VecCreateSeq(MPI_COMM_SELF, n0, &vxl);
    VecCreateMPI(MPI_COMM_WORLD, n0l, nDofs0G, &uxg);
    VecGetArray(ul, &uArray);
    VecGetArray(uxl, &uxArray);
    VecGetArray(vxl, &vxArray);
    for (int ey = 0; ey < topo->nElsX; ey++)
        for (int ex = 0; ex < topo->nElsX; ex++) {
            int* inds0 = elInds0_l(ex, ey);
            for (int ii = 0; ii < mp12; ii++) {
                double val[2];
                interp1_g(ex, ey, ii % mp1, ii / mp1, uArray, val);
                uxArray[inds0[ii]] = lev >= 0 ? val[0] / thick[lev][inds0[ii]] : val[0];
                vxArray[inds0[ii]] = lev >= 0 ? val[1] / thick[lev][inds0[ii]] : val[1];
            }
        }
    VecRestoreArray(uxl, &uxArray);
    VecRestoreArray(vxl, &vxArray);
    VecRestoreArray(ul, &uArray);
    VecZeroEntries(uxg);
    VecScatterBegin(gtol_0, uxl, uxg, INSERT_VALUES, SCATTER_REVERSE);
    VecScatterEnd(gtol_0, uxl, uxg, INSERT_VALUES, SCATTER_REVERSE);
    field_file(filename, sizeof filename, fieldname, "_x", lev, tstep, "dat");
    view_to(uxg, filename, false);
    VecZeroEntries(uxg);
    VecScatterBegin(gtol_0, vxl, uxg, INSERT_VALUES, SCATTER_REVERSE);
    VecScatterEnd(gtol_0, vxl, uxg, INSERT_VALUES, SCATTER_REVERSE);
    field_file(filename, sizeof filename, fieldname, "_y", lev, tstep, "dat");
    view_to(uxg, filename, false);
    VecDestroy(&ul);
    VecDestroy(&uxl);
    VecDestroy(&vxl);
    VecDestroy(&uxg);
    field_file(filename, sizeof filename, fieldname, "", lev, tstep, "vec");
    view_to(u, filename, true);   // also the vector itself
}

void Geom::write2(Vec h, char* fieldname, int tstep, int lev, bool vert_scale) {
    const int mp1 = quad->n + 1, mp12 = mp1 * mp1;
    char filename[200];
    Vec hxl, hxg;
    PetscScalar *hxArray, *hArray;
    VecCreateSeq(MPI_COMM_SELF, n0, &hxl);
    VecCreateMPI(MPI_COMM_WORLD, n0l, nDofs0G, &hxg);
    VecZeroEntries(hxg);
    VecGetArray(h, &hArray);
    VecGetArray(hxl, &hxArray);
    for (int ey = 0; ey < topo->nElsX; ey++)
        for (int ex = 0; ex < topo->nElsX; ex++) {
            int* inds0 = elInds0_l(ex, ey);
            for (int ii = 0; ii < mp12; ii++) {
                double val;
                interp2_g(ex, ey, ii % mp1, ii / mp1, hArray, &val);
                if (vert_scale && lev >= 0) val /= thick[lev][inds0[ii]];
                hxArray[inds0[ii]] = val;
            }
        }
    VecRestoreArray(h, &hArray);
    VecRestoreArray(hxl, &hxArray);
    VecScatterBegin(gtol_0, hxl, hxg, INSERT_VALUES, SCATTER_REVERSE);
    VecScatterEnd(gtol_0, hxl, hxg, INSERT_VALUES, SCATTER_REVERSE);
    field_file(filename, sizeof filename, fieldname, "", lev, tstep, "dat");
    view_to(hxg, filename, false);
    VecDestroy(&hxg);
    VecDestroy(&hxl);
    field_file(filename, sizeof filename, fieldname, "", lev, tstep, "vec");
    view_to(h, filename, true);
}

void Geom::writeVertToHoriz(Vec* vecs, char* fieldname, int tstep, int nv) {
    const int n2e = topo->elOrd * topo->elOrd;
    std::vector<Vec> hvecs(nv);
    for (int kk = 0; kk < nv; kk++) {
        VecCreateMPI(MPI_COMM_WORLD, topo->n2l, topo->nDofs2G, &hvecs[kk]);
        VecZeroEntries(hvecs[kk]);
    }
    for (int ey = 0; ey < topo->nElsX; ey++)
        for (int ex = 0; ex < topo->nElsX; ex++) {
            const int* inds2 = topo->elInds2_l(ex, ey);
            PetscScalar *vArray, *hArray;
            VecGetArray(vecs[ey * topo->nElsX + ex], &vArray);
            for (int kk = 0; kk < nv; kk++) {
                VecGetArray(hvecs[kk], &hArray);
                for (int ii = 0; ii < n2e; ii++) hArray[inds2[ii]] += vArray[kk * n2e + ii];
                VecRestoreArray(hvecs[kk], &hArray);
            }
            VecRestoreArray(vecs[ey * topo->nElsX + ex], &vArray);
        }
    for (int kk = 0; kk < nv; kk++) {
        write2(hvecs[kk], fieldname, tstep, kk, false);
        VecDestroy(&hvecs[kk]);
    }
}
