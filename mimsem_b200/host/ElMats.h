// Host mirror of the reference's ElMats (eul/ElMats.h:1-55): reference-element tabulations as dense
// row-major (quadrature point, dof) matrices.  The device kernels never form them (they contract the
// 1-D tables directly); they are provided because callers outside the hot path (VertOps) use them.
#ifndef MIMSEM_HOST_ELMATS_H
#define MIMSEM_HOST_ELMATS_H

#include "Basis.h"

class Geom;

class M1x_j_xy_i {
    public:
        M1x_j_xy_i(LagrangeNode* _node, LagrangeEdge* _edge);
        ~M1x_j_xy_i();
        int nDofsI, nDofsJ;
        double* A;
        LagrangeNode* node;
        LagrangeEdge* edge;
};
class M1y_j_xy_i {
    public:
        M1y_j_xy_i(LagrangeNode* _node, LagrangeEdge* _edge);
        ~M1y_j_xy_i();
        int nDofsI, nDofsJ;
        double* A;
        LagrangeNode* node;
        LagrangeEdge* edge;
};
class M2_j_xy_i {
    public:
        M2_j_xy_i(LagrangeEdge* _edge);
        ~M2_j_xy_i();
        int nDofsI, nDofsJ;
        double* A;
        LagrangeEdge* edge;
};
class M0_j_xy_i {
    public:
        M0_j_xy_i(LagrangeNode* _node);
        ~M0_j_xy_i();
        int nDofsI, nDofsJ;
        double* A;
        LagrangeNode* node;
};
class Wii {
    public:
        Wii(GaussLobatto* _quad, Geom* _geom);
        ~Wii();
        int nDofsI, nDofsJ;
        double* A;     // the diagonal, flat (eul/ElMats.cpp:177)
        GaussLobatto* quad;
        Geom* geom;
};

#endif
