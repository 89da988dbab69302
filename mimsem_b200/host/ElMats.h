// Host mirror of the reference's ElMats (eul/ElMats.h:1-55): reference-element tabulations as dense
// row-major (quadrature point, dof) matrices.  The device kernels never form them (they contract the
// 1-D tables directly); they are provided because callers outside the hot path (VertOps) use them.
#ifndef MIMSEM_HOST_ELMATS_H
#define MIMSEM_HOST_ELMATS_H

#include "Basis.h"

class Geom;

// The reference's eul/ElMats.h stores a tabulation as one flat row-major array (A[q * nDofsJ + j]; Wii: its diagonal),
// src/ElMats.h and box/ElMats.h as rows (A[q][j]; Wii: a full matrix of which only A[q][q] is set).  Both views of the same
// numbers are kept in every object, in the same two slots, so that one library serves callers of either directory: compile
// src/ and box/ callers with -DMIMSEM_ELMATS_ROWS and `A` names the row view.
#ifdef MIMSEM_ELMATS_ROWS
#define MIMSEM_ELMAT_VIEWS double* Aflat; double** A;
#else
#define MIMSEM_ELMAT_VIEWS double* A; double** Arows;
#endif

class M1x_j_xy_i {
    public:
        M1x_j_xy_i(LagrangeNode* _node, LagrangeEdge* _edge);
        ~M1x_j_xy_i();
        int nDofsI, nDofsJ;
        MIMSEM_ELMAT_VIEWS
        LagrangeNode* node;
        LagrangeEdge* edge;
};
class M1y_j_xy_i {
    public:
        M1y_j_xy_i(LagrangeNode* _node, LagrangeEdge* _edge);
        ~M1y_j_xy_i();
        int nDofsI, nDofsJ;
        MIMSEM_ELMAT_VIEWS
        LagrangeNode* node;
        LagrangeEdge* edge;
};
class M2_j_xy_i {
    public:
        M2_j_xy_i(LagrangeEdge* _edge);
        ~M2_j_xy_i();
        int nDofsI, nDofsJ;
        MIMSEM_ELMAT_VIEWS
        LagrangeEdge* edge;
};
class M0_j_xy_i {
    public:
        M0_j_xy_i(LagrangeNode* _node);
        ~M0_j_xy_i();
        int nDofsI, nDofsJ;
        MIMSEM_ELMAT_VIEWS
        LagrangeNode* node;
};
class Wii {
    public:
        Wii(GaussLobatto* _quad, Geom* _geom);
        ~Wii();
        int nDofsI, nDofsJ;
        MIMSEM_ELMAT_VIEWS   // flat: the diagonal (eul/ElMats.cpp:177); rows: the full matrix, zero off the diagonal (box/ElMats.cpp)
        GaussLobatto* quad;
        Geom* geom;
    private:
        double* Afull;
};

#endif
