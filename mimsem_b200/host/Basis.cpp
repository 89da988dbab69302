#include "Basis.h"

#include <cstdio>
#include <vector>

#include "../csrc/basis.hpp"

namespace {
double** alloc2(int r, int c) {
    double** p = new double*[r];
    for (int i = 0; i < r; i++) p[i] = new double[c];
    return p;
}
void free2(double** p, int r) {
    for (int i = 0; i < r; i++) delete[] p[i];
    delete[] p;
}
}  // namespace

GaussLobatto::GaussLobatto(int _n) : n(_n), x(new double[_n + 1]), w(new double[_n + 1]) {
    std::vector<double> vx, vw;
    if (!mimsem::gll_rule(n, vx, vw)) {
        std::fprintf(stderr, "invalid gauss-lobatto quadrature order: %d\n", n);   // the reference prints and continues
        for (int i = 0; i <= n; i++) x[i] = w[i] = 0.0;
        return;
    }
    for (int i = 0; i <= n; i++) {
        x[i] = vx[i];
        w[i] = vw[i];
    }
}
GaussLobatto::~GaussLobatto() {
    delete[] x;
    delete[] w;
}

LagrangeNode::LagrangeNode(int _n, GaussLobatto* _q) : n(_n), q(_q) {
    GaussLobatto own(n);
    x = new double[n + 1];
    a = new double[n + 1];
    for (int i = 0; i <= n; i++) x[i] = own.x[i];
    for (int i = 0; i <= n; i++) {
        double prod = 1.0;
        for (int j = 0; j <= n; j++)
            if (j != i) prod *= 1.0 / (q->x[i] - q->x[j]);
        a[i] = prod;
    }
    ljxi = alloc2(q->n + 1, n + 1);
    ljxi_t = alloc2(n + 1, q->n + 1);
    for (int iq = 0; iq <= q->n; iq++)
        for (int j = 0; j <= n; j++) ljxi[iq][j] = ljxi_t[j][iq] = eval_q(q->x[iq], j);
}
LagrangeNode::~LagrangeNode() {
    free2(ljxi, q->n + 1);
    free2(ljxi_t, n + 1);
    delete[] x;
    delete[] a;
}
double LagrangeNode::eval(double _x, int i) {
    double prod = 1.0;
    for (int j = 0; j <= n; j++)
        if (j != i) prod *= _x - q->x[j];
    return a[i] * prod;
}
double LagrangeNode::eval_q(double _x, int i) {
    double y = 1.0;
    for (int j = 0; j <= n; j++)
        if (j != i) y *= (_x - x[j]) / (x[i] - x[j]);
    return y;
}
double LagrangeNode::evalDeriv(double _x, int i) {
    double sum = 0.0;
    for (int j = 0; j <= n; j++) {
        if (j == i) continue;
        double prod = 1.0;
        for (int k = 0; k <= n; k++)
            if (k != i && k != j) prod *= (_x - x[k]) / (x[i] - x[k]);
        sum += prod / (x[i] - x[j]);
    }
    return sum;
}

LagrangeEdge::LagrangeEdge(int _n, LagrangeNode* _l) : n(_n), l(_l) {
    const int mq = l->q->n;
    ejxi = alloc2(mq + 1, n);
    ejxi_t = alloc2(n, mq + 1);
    for (int iq = 0; iq <= mq; iq++)
        for (int j = 0; j < n; j++) ejxi[iq][j] = ejxi_t[j][iq] = eval(l->q->x[iq], j);
}
LagrangeEdge::~LagrangeEdge() {
    free2(ejxi, l->q->n + 1);
    free2(ejxi_t, n);
}
double LagrangeEdge::eval(double x, int i) {
    double c = 0.0;
    for (int j = 0; j <= i; j++) c -= l->evalDeriv(x, j);
    return c;
}
