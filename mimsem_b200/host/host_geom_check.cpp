// Checks of the host Geom mirror that need no GPU: Geom::interp0 / interp1_l / interp2_l / interp1_g / interp2_g
// (eul/Geom.cpp:328-417), Geom::initTopog (:743-764), the field writers write0 / write1 / write2 (:419-631) and the
// PETSc binary Vec format through VecView / VecLoad.
//
//   host_geom_check <p> <ne> <nprocs> <nk> <rank> <in.bin> <out.bin> <scratch dir>
// in.bin : doubles v0[n0] v1[n1] v2[n2] (ghosted local vectors of `rank`; 2-forms: the owned array)
// out.bin: doubles interp0[nel][q2] interp1_l[nel][q2][2] interp2_l[nel][q2] interp1_g[nel][q2][2] interp2_g[nel][q2] thick[nk][n0q]
// The writers are exercised on every rank (collective); their files land in <scratch dir>/output/.
#include <sys/stat.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "Geom.h"

static int g_nk = 1;
static double topog_fn(double* x) {
    const double r2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
    return 2000.0 * x[2] * x[2] / r2;
}
static double level_fn(double* x, int ki) {   // eul/UMJS14.cpp:124-129
    (void)x;
    const double mu = 15.0, ztop = 30000.0, f = (double)ki / g_nk;
    return ztop * (std::sqrt(mu * f * f + 1.0) - 1.0) / (std::sqrt(mu + 1.0) - 1.0);
}

int main(int argc, char** argv) {
    if (argc < 9) return 2;
    const int p = std::atoi(argv[1]), ne = std::atoi(argv[2]), np = std::atoi(argv[3]), nk = std::atoi(argv[4]), rank = std::atoi(argv[5]);
    g_nk = nk;
    std::vector<Topo*> topo(np);
    std::vector<Geom*> geom(np);
    for (int r = 0; r < np; r++) {
        PetscCompatSetRank(r, np);
        topo[r] = new Topo(0 /* MIMSEM_MESH_SPHERE */, p, ne, nk);
        geom[r] = new Geom(topo[r], nk);
        geom[r]->initTopog(topog_fn, level_fn);
    }
    Topo* t = topo[rank];
    Geom* g = geom[rank];
    FILE* f = std::fopen(argv[6], "rb");
    if (!f) return 2;
    std::vector<double> v0(t->n0), v1(t->n1), v2(t->n2);
    if (std::fread(v0.data(), 8, v0.size(), f) != v0.size() || std::fread(v1.data(), 8, v1.size(), f) != v1.size() ||
        std::fread(v2.data(), 8, v2.size(), f) != v2.size())
        return 2;
    std::fclose(f);
    FILE* out = std::fopen(argv[7], "wb");
    if (!out) return 2;
    const int mp1 = g->quad->n + 1, q2 = mp1 * mp1, nel = t->nElsX * t->nElsX;
    PetscCompatSetRank(rank, np);
    for (int which = 0; which < 5; which++)
        for (int el = 0; el < nel; el++)
            for (int q = 0; q < q2; q++) {
                const int ex = el % t->nElsX, ey = el / t->nElsX;
                double val[2] = {0.0, 0.0};
                switch (which) {
                    case 0: g->interp0(ex, ey, q % mp1, q / mp1, v0.data(), val); break;
                    case 1: g->interp1_l(ex, ey, q % mp1, q / mp1, v1.data(), val); break;
                    case 2: g->interp2_l(ex, ey, q % mp1, q / mp1, v2.data(), val); break;
                    case 3: g->interp1_g(ex, ey, q % mp1, q / mp1, v1.data(), val); break;
                    default: g->interp2_g(ex, ey, q % mp1, q / mp1, v2.data(), val); break;
                }
                std::fwrite(val, 8, (which == 1 || which == 3) ? 2 : 1, out);
            }
    for (int k = 0; k < nk; k++) std::fwrite(g->thick[k], 8, g->n0, out);
    std::fclose(out);
    // writers + binary round trip: a global 2-form h with h[i] = i + 0.25, a global 1-form, a global 0-form
    if (chdir(argv[8]) != 0) return 2;
    mkdir("output", 0755);
    std::vector<Vec> h2(np), u1(np), q0(np), back(np);
    for (int r = 0; r < np; r++) {
        PetscCompatSetRank(r, np);
        VecCreateMPI(MPI_COMM_WORLD, topo[r]->n2l, topo[r]->nDofs2G, &h2[r]);
        VecCreateMPI(MPI_COMM_WORLD, topo[r]->n1l, topo[r]->nDofs1G, &u1[r]);
        VecCreateMPI(MPI_COMM_WORLD, topo[r]->n0l, topo[r]->nDofs0G, &q0[r]);
        VecCreateMPI(MPI_COMM_WORLD, topo[r]->n2l, topo[r]->nDofs2G, &back[r]);
    }
    for (int r = 0; r < np; r++) {
        PetscCompatSetRank(r, np);
        Vec vs[3] = {h2[r], u1[r], q0[r]};
        for (int s = 0; s < 3; s++) {
            PetscScalar* a;
            PetscInt lo, hi;
            VecGetOwnershipRange(vs[s], &lo, &hi);
            VecGetArray(vs[s], &a);
            for (int i = lo; i < hi; i++) a[i - lo] = (s == 0 ? 1.0e4 : 1.0) * (1.0 + 0.001 * (i % 97)) + 0.25;
            VecRestoreArray(vs[s], &a);
        }
    }
    char fh[] = "rho", fu[] = "velocity", fq[] = "vorticity";
    for (int r = 0; r < np; r++) { PetscCompatSetRank(r, np); geom[r]->write2(h2[r], fh, 7, 1, true); }
    for (int r = 0; r < np; r++) { PetscCompatSetRank(r, np); geom[r]->write1(u1[r], fu, 7, 1); }
    for (int r = 0; r < np; r++) { PetscCompatSetRank(r, np); geom[r]->write0(q0[r], fq, 7, 1); }
    double worst = 0.0;
    for (int r = 0; r < np; r++) {
        PetscCompatSetRank(r, np);
        PetscViewer viewer;
        PetscViewerBinaryOpen(PETSC_COMM_WORLD, "output/rho_001_0007.vec", FILE_MODE_READ, &viewer);
        if (VecLoad(back[r], viewer)) return 3;
        PetscViewerDestroy(&viewer);
        PetscScalar *a, *b;
        VecGetArray(back[r], &a);
        VecGetArray(h2[r], &b);
        for (int i = 0; i < topo[r]->n2l; i++) worst = std::fmax(worst, std::fabs(a[i] - b[i]));
        VecRestoreArray(back[r], &a);
        VecRestoreArray(h2[r], &b);
    }
    std::printf("host_geom_check ok: binary round trip max |diff| = %g\n", worst);
    return worst == 0.0 ? 0 : 4;
}
