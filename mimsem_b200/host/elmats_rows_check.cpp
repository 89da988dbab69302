// CPU check of the row view of the element tabulations (src/ElMats.h, box/ElMats.h: double** A): this translation unit is
// compiled with -DMIMSEM_ELMATS_ROWS, the library without, and both must see the same numbers in the same slots.
#include <cstdio>

#include "Assembly.h"

template <class T>
static int check(const T& m, const char* name) {
    int bad = 0;
    for (int q = 0; q < m.nDofsI; q++)
        for (int j = 0; j < m.nDofsJ; j++)
            if (m.A[q][j] != m.Aflat[q * m.nDofsJ + j]) bad++;
    std::printf("%-12s %d x %d  %s\n", name, m.nDofsI, m.nDofsJ, bad ? "FAIL" : "ok");
    return bad;
}

int main() {
    int bad = 0;
    for (int p = 2; p <= 4; p++) {
        GaussLobatto q(p);
        LagrangeNode l(p, &q);
        LagrangeEdge e(p, &l);
        M1x_j_xy_i U(&l, &e);
        M1y_j_xy_i V(&l, &e);
        M2_j_xy_i W(&e);
        M0_j_xy_i P(&l);
        Wii Q(&q, NULL);
        bad += check(U, "M1x_j_xy_i") + check(V, "M1y_j_xy_i") + check(W, "M2_j_xy_i") + check(P, "M0_j_xy_i");
        for (int i = 0; i < Q.nDofsI; i++)
            for (int j = 0; j < Q.nDofsJ; j++)
                if (Q.A[i][j] != (i == j ? Q.Aflat[i] : 0.0)) bad++;
        // the first row of U is l_j(xi_0) e_j(xi_0): 1 at the first x-edge of the west line only for the nodal part
        if (U.A[0][0] != l.ljxi[0][0] * e.ejxi[0][0]) bad++;
    }
    std::printf("elmats_rows_check %s\n", bad ? "FAIL" : "ok");
    return bad ? 1 : 0;
}
