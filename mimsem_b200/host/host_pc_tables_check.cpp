// CPU check of the host half of the element-block Jacobi preconditioner of the Umat shell (Assembly.cpp: pc_tables): for
// every rank of an emulated `mpirun -np nprocs` the patch's elements plus one copy of the west / south neighbour of every
// element on the patch's west / south boundary; patch placement and neighbour tables are verified against the point
// coordinates inside, the edge-use counts here.  No GPU.
//   host_pc_tables_check <kind 0|1> <p> <ne> <nprocs> <nk>
#include <cstdio>
#include <cstdlib>

#include "Assembly.h"

int main(int argc, char** argv) {
    if (argc < 6) return 2;
    const int kind = std::atoi(argv[1]), p = std::atoi(argv[2]), ne = std::atoi(argv[3]), np = std::atoi(argv[4]), nk = std::atoi(argv[5]);
    int failures = 0;
    for (int r = 0; r < np; r++) {
        PetscCompatSetRank(r, np);
        Topo* topo = new Topo(kind, p, ne, nk);
        Geom* geom = new Geom(topo, nk);
        int sz[6] = {0, 0, 0, 0, 0, 0};
        const int rc = MimsemPCTablesCheck(topo, geom, sz);
        // a sphere patch has a neighbour across every west and south element side; so has a periodic box patch
        const int expect = 2 * topo->nElsX;
        const bool ok = rc == 0 && sz[0] == topo->nElsX * topo->nElsX && sz[1] == expect &&
                        sz[3] == topo->n1 + sz[1] * (2 * p * (p + 1) - p) && sz[5] == geom->n0 + sz[1] * ((p + 1) * (p + 1) - (p + 1));
        std::printf("rank %2d: rc %d, %d elements + %d neighbour copies, n1 %d -> %d, nq %d -> %d  %s\n", r, rc, sz[0], sz[1], topo->n1, sz[3],
                    geom->n0, sz[5], ok ? "ok" : "FAIL");
        if (!ok) failures++;
        delete geom;
        delete topo;
    }
    std::printf("host_pc_tables_check %s\n", failures ? "FAIL" : "ok");
    return failures ? 1 : 0;
}
