// Basis tabulations for mixed mimetic spectral elements (host side).
//
// Replaces the reference's Basis.{h,cpp} (GaussLobatto, LagrangeNode, LagrangeEdge):
//   GLL points/weights            <- eul/Basis.cpp:22-98
//   nodal Lagrange basis l_j(x)   <- eul/Basis.cpp:183-190 (eval_q), :192-213 (evalDeriv)
//   edge (histopolation) basis    <- eul/Basis.cpp:277-286   e_i(x) = - sum_{j<=i} l_j'(x)
// The tables keep the reference's orientation: ljxi[q][j], ejxi[q][i] (quadrature point first).
#pragma once
#include <vector>

namespace mimsem {

// Gauss-Lobatto-Legendre rule with n+1 points on [-1,1] (exact to degree 2n-1), n = 1..7.
// Returns false for an unsupported order.
bool gll_rule(int n, std::vector<double>& x, std::vector<double>& w);

struct BasisTables {
    int p = 0;                 // polynomial order of the nodal basis (p+1 nodes, p edge functions)
    int m = 0;                 // quadrature order (m+1 points)
    std::vector<double> qx, qw;   // quadrature points / weights, m+1
    std::vector<double> nx;       // nodal GLL points of order p, p+1
    std::vector<double> ljxi;     // (m+1) x (p+1): l_j(qx_q)
    std::vector<double> ejxi;     // (m+1) x p    : e_i(qx_q)
    bool build(int p_, int m_);
    // evaluate at an arbitrary abscissa (used by upwinded operators, src/Assembly.cpp:537-541)
    double node_eval(double x, int j) const;
    double node_deriv(double x, int j) const;
    double edge_eval(double x, int i) const;
};

}  // namespace mimsem
