// Host mirror of the reference's Basis.h (eul/Basis.h:1-36): same class and member names, so the
// reference's callers (HorizSolve, Euler, SWEqn, VertOps) compile against it unchanged.
// Tabulations are produced by mimsem::BasisTables (csrc/basis.hpp).
#ifndef MIMSEM_HOST_BASIS_H
#define MIMSEM_HOST_BASIS_H

class GaussLobatto {
    public:
        GaussLobatto(int _n);
        ~GaussLobatto();
        int n;
        double* x;
        double* w;
};

class LagrangeNode {
    public:
        LagrangeNode(int _n, GaussLobatto* _q);
        ~LagrangeNode();
        int n;
        double* a;         // 1 / prod_{j != i} (x_i - x_j) over the quadrature points (reference quirk, eul/Basis.cpp:121-127)
        double* x;         // nodal GLL points of order n
        double** ljxi;     // [q->n+1][n+1]
        double** ljxi_t;   // [n+1][q->n+1]
        GaussLobatto* q;
        double eval(double x, int i);
        double eval_q(double x, int i);
        double evalDeriv(double x, int i);
};

class LagrangeEdge {
    public:
        LagrangeEdge(int _n, LagrangeNode* _l);
        ~LagrangeEdge();
        int n;
        double** ejxi;     // [l->q->n+1][n]
        double** ejxi_t;   // [n][l->q->n+1]
        LagrangeNode* l;
        double eval(double x, int i);
};

#endif
