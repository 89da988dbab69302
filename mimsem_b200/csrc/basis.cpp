// Host-side basis tabulations; see basis.hpp for the reference lines each piece restates.
#include "basis.hpp"

#include <cmath>

namespace mimsem {

bool gll_rule(int n, std::vector<double>& x, std::vector<double>& w) {
    if (n < 1 || n > 7) return false;
    x.assign(n + 1, 0.0);
    w.assign(n + 1, 0.0);
    // interior abscissae (positive half) and the weights of [endpoint, interior..., centre]
    // are the closed forms the reference uses (eul/Basis.cpp:31-55); order 7 is tabulated to
    // 15 digits there (eul/Basis.cpp:82-86) and must be reproduced digit for digit.
    std::vector<double> pos;   // descending positive abscissae, excluding 1 and 0
    std::vector<double> wts;   // weights for x = 1, then pos[...], then (even n) x = 0
    switch (n) {
        case 1: wts = {1.0}; break;
        case 2: wts = {1.0 / 3.0, 4.0 / 3.0}; break;
        case 3: pos = {std::sqrt(0.2)}; wts = {1.0 / 6.0, 5.0 / 6.0}; break;
        case 4: pos = {std::sqrt(3.0 / 7.0)}; wts = {0.1, 49.0 / 90.0, 64.0 / 90.0}; break;
        case 5: {
            const double a = 2.0 * std::sqrt(7.0) / 21.0;
            pos = {std::sqrt(1.0 / 3.0 + a), std::sqrt(1.0 / 3.0 - a)};
            wts = {1.0 / 15.0, (14.0 - std::sqrt(7.0)) / 30.0, (14.0 + std::sqrt(7.0)) / 30.0};
            break;
        }
        case 6: {
            const double a = 2.0 * std::sqrt(5.0 / 3.0) / 11.0;
            pos = {std::sqrt(5.0 / 11.0 + a), std::sqrt(5.0 / 11.0 - a)};
            wts = {1.0 / 21.0, (124.0 - 7.0 * std::sqrt(15.0)) / 350.0, (124.0 + 7.0 * std::sqrt(15.0)) / 350.0,
                   256.0 / 525.0};
            break;
        }
        case 7:
            pos = {0.871740148509607, 0.591700181433142, 0.209299217902479};
            wts = {0.035714285714286, 0.210704227143506, 0.341122692483504, 0.412458794658704};
            break;
    }
    x[0] = -1.0;
    x[n] = +1.0;
    w[0] = w[n] = wts[0];
    for (size_t i = 0; i < pos.size(); i++) {
        x[1 + i] = -pos[i];
        x[n - 1 - i] = +pos[i];
        w[1 + i] = w[n - 1 - i] = wts[1 + i];
    }
    if (n % 2 == 0) {
        x[n / 2] = 0.0;
        w[n / 2] = wts.back();
    }
    return true;
}

double BasisTables::node_eval(double x, int j) const {
    double y = 1.0;
    for (int k = 0; k <= p; k++)
        if (k != j) y *= (x - nx[k]) / (nx[j] - nx[k]);
    return y;
}

double BasisTables::node_deriv(double x, int j) const {
    // d/dx prod_{k!=j} (x-x_k)/(x_j-x_k) = sum_{i!=j} [ prod_{k!=i,j} (x-x_k)/(x_j-x_k) ] / (x_j-x_i)
    double sum = 0.0;
    for (int i = 0; i <= p; i++) {
        if (i == j) continue;
        double prod = 1.0;
        for (int k = 0; k <= p; k++)
            if (k != j && k != i) prod *= (x - nx[k]) / (nx[j] - nx[k]);
        sum += prod / (nx[j] - nx[i]);
    }
    return sum;
}

double BasisTables::edge_eval(double x, int i) const {
    double c = 0.0;
    for (int j = 0; j <= i; j++) c -= node_deriv(x, j);
    return c;
}

bool BasisTables::build(int p_, int m_) {
    p = p_;
    m = m_;
    std::vector<double> nw;
    if (!gll_rule(m, qx, qw) || !gll_rule(p, nx, nw)) return false;
    ljxi.assign((size_t)(m + 1) * (p + 1), 0.0);
    ejxi.assign((size_t)(m + 1) * p, 0.0);
    for (int q = 0; q <= m; q++) {
        for (int j = 0; j <= p; j++) ljxi[(size_t)q * (p + 1) + j] = node_eval(qx[q], j);
        for (int i = 0; i < p; i++) ejxi[(size_t)q * p + i] = edge_eval(qx[q], i);
    }
    return true;
}

}  // namespace mimsem
