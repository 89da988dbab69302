#include "ElMats.h"

namespace {
double** row_view(double* flat, int nrows, int ncols) {
    double** rows = new double*[nrows];
    for (int i = 0; i < nrows; i++) rows[i] = flat + (long)i * ncols;
    return rows;
}
}  // namespace

// index conventions: eul/ElMats.cpp:38-44 (U), 73-79 (V), 105-111 (W), 135-141 (P), 177 (Wii)
M1x_j_xy_i::M1x_j_xy_i(LagrangeNode* _node, LagrangeEdge* _edge) : node(_node), edge(_edge) {
    const int n = node->n, np1 = n + 1, mp1 = node->q->n + 1;
    nDofsI = mp1 * mp1;
    nDofsJ = n * np1;
    A = new double[nDofsI * nDofsJ];
    for (int q = 0; q < nDofsI; q++)
        for (int j = 0; j < nDofsJ; j++) A[q * nDofsJ + j] = node->ljxi[q % mp1][j % np1] * edge->ejxi[q / mp1][j / np1];
    Arows = row_view(A, nDofsI, nDofsJ);
}
M1x_j_xy_i::~M1x_j_xy_i() { delete[] Arows; delete[] A; }

M1y_j_xy_i::M1y_j_xy_i(LagrangeNode* _node, LagrangeEdge* _edge) : node(_node), edge(_edge) {
    const int n = node->n, np1 = n + 1, mp1 = node->q->n + 1;
    nDofsI = mp1 * mp1;
    nDofsJ = n * np1;
    A = new double[nDofsI * nDofsJ];
    for (int q = 0; q < nDofsI; q++)
        for (int j = 0; j < nDofsJ; j++) A[q * nDofsJ + j] = edge->ejxi[q % mp1][j % n] * node->ljxi[q / mp1][j / n];
    Arows = row_view(A, nDofsI, nDofsJ);
}
M1y_j_xy_i::~M1y_j_xy_i() { delete[] Arows; delete[] A; }

M2_j_xy_i::M2_j_xy_i(LagrangeEdge* _edge) : edge(_edge) {
    const int n = edge->n, mp1 = edge->l->q->n + 1;
    nDofsI = mp1 * mp1;
    nDofsJ = n * n;
    A = new double[nDofsI * nDofsJ];
    for (int q = 0; q < nDofsI; q++)
        for (int j = 0; j < nDofsJ; j++) A[q * nDofsJ + j] = edge->ejxi[q % mp1][j % n] * edge->ejxi[q / mp1][j / n];
    Arows = row_view(A, nDofsI, nDofsJ);
}
M2_j_xy_i::~M2_j_xy_i() { delete[] Arows; delete[] A; }

M0_j_xy_i::M0_j_xy_i(LagrangeNode* _node) : node(_node) {
    const int np1 = node->n + 1, mp1 = node->q->n + 1;
    nDofsI = mp1 * mp1;
    nDofsJ = np1 * np1;
    A = new double[nDofsI * nDofsJ];
    for (int q = 0; q < nDofsI; q++)
        for (int j = 0; j < nDofsJ; j++) A[q * nDofsJ + j] = node->ljxi[q % mp1][j % np1] * node->ljxi[q / mp1][j / np1];
    Arows = row_view(A, nDofsI, nDofsJ);
}
M0_j_xy_i::~M0_j_xy_i() { delete[] Arows; delete[] A; }

Wii::Wii(GaussLobatto* _quad, Geom* _geom) : quad(_quad), geom(_geom) {
    const int mp1 = quad->n + 1;
    nDofsI = nDofsJ = mp1 * mp1;
    A = new double[nDofsI];
    Afull = new double[nDofsI * nDofsJ]();
    for (int q = 0; q < nDofsI; q++) A[q] = Afull[q * nDofsJ + q] = quad->w[q % mp1] * quad->w[q / mp1];
    Arows = row_view(Afull, nDofsI, nDofsJ);
}
Wii::~Wii() {
    delete[] Arows;
    delete[] Afull;
    delete[] A;
}
