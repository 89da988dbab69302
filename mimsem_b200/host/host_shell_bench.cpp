// The MatShell path measured: Umat through the C++ mirror on the six patches of an emulated `mpirun -np 6` (one GPU, the
// patches one after the other), host Vecs in, host Vecs out --
//   (a) the reference's loop   for kk: M1->assemble(kk, SCALE, true); MatMult(M1->M, x[kk], y[kk])     (eul/Euler_2.cpp:1427-1456)
//   (b) the same levels in one device call per patch: MimsemMatMultLevels(M1->M, 0, nk, x, y, NULL)
// and prints one JSON line: wall time of each, and the part of it spent inside the device library (host -> device copies,
// kernels, device -> host copies: MimsemDeviceSeconds) -- the rest is the ghost refresh and the shared-DOF sum, here the
// VecScatter of the in-tree compatibility layer (six ranks played by one thread), with PETSc its own.
//   host_shell_bench <p> <ne> <nk> [reps]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "Assembly.h"

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const int p = std::atoi(argv[1]), ne = std::atoi(argv[2]), nk = std::atoi(argv[3]), reps = argc > 4 ? std::atoi(argv[4]) : 3, np = 6;
    struct Rank { Topo* topo; Geom* geom; GaussLobatto* quad; LagrangeNode* node; LagrangeEdge* edge; Umat* M1; std::vector<Vec> x, y; Vec xl, yl; };
    std::vector<Rank> R(np);
    for (int r = 0; r < np; r++) {
        PetscCompatSetRank(r, np);
        Rank& k = R[r];
        k.topo = new Topo(0, p, ne, nk);
        k.geom = new Geom(k.topo, nk);
        k.quad = new GaussLobatto(k.geom->quad->n);
        k.node = new LagrangeNode(k.topo->elOrd, k.quad);
        k.edge = new LagrangeEdge(k.topo->elOrd, k.node);
        for (int lev = 0; lev < nk; lev++)
            for (int i = 0; i < k.geom->n0; i++) {
                k.geom->thick[lev][i] = (200.0 + 30.0 * lev) * (1.0 + 0.1 * std::sin(1.0e-6 * k.geom->x[i][2]));
                k.geom->thickInv[lev][i] = 1.0 / k.geom->thick[lev][i];
            }
        k.geom->thick_version++;
        k.M1 = new Umat(k.topo, k.geom, k.node, k.edge);
        VecCreateSeq(MPI_COMM_SELF, k.topo->n1, &k.xl);
        VecCreateSeq(MPI_COMM_SELF, k.topo->n1, &k.yl);
    }
    for (int lev = 0; lev < nk; lev++)
        for (int r = 0; r < np; r++) {
            PetscCompatSetRank(r, np);
            Vec v;
            VecCreateMPI(MPI_COMM_WORLD, R[r].topo->n1l, R[r].topo->nDofs1G, &v);
            PetscScalar* a;
            VecGetArray(v, &a);
            for (int i = 0; i < R[r].topo->n1l; i++) a[i] = std::sin(0.001 * i + lev + r);
            VecRestoreArray(v, &a);
            R[r].x.push_back(v);
            VecCreateMPI(MPI_COMM_WORLD, R[r].topo->n1l, R[r].topo->nDofs1G, &v);
            R[r].y.push_back(v);
        }
    const double dofs = (double)R[0].topo->nDofs1G * nk;
    double t[2] = {1e30, 1e30}, dev[2] = {0.0, 0.0};
    std::vector<double> keep;
    double diff = 0.0, norm = 0.0;
    for (int rep = 0; rep < reps + 1; rep++) {   // first round: warm-up (contexts, staging buffers)
        MimsemDeviceSeconds(1);
        double t0 = now();
        for (int lev = 0; lev < nk; lev++)
            for (int r = 0; r < np; r++) {
                PetscCompatSetRank(r, np);
                R[r].M1->assemble(lev, SCALE, true);
                MatMult(R[r].M1->M, R[r].x[lev], R[r].y[lev]);
            }
        if (rep && now() - t0 < t[0]) { t[0] = now() - t0; dev[0] = MimsemDeviceSeconds(0); }
        if (rep == 1) {
            PetscCompatSetRank(0, np);
            PetscScalar* a;
            VecGetArray(R[0].y[nk - 1], &a);
            keep.assign(a, a + R[0].topo->n1l);
            VecRestoreArray(R[0].y[nk - 1], &a);
        }
        MimsemDeviceSeconds(1);
        t0 = now();
        for (int r = 0; r < np; r++) {
            PetscCompatSetRank(r, np);
            R[r].M1->assemble(0, SCALE, true);
            if (MimsemMatMultLevels(R[r].M1->M, 0, nk, R[r].x.data(), R[r].y.data(), NULL)) return 1;
        }
        if (rep && now() - t0 < t[1]) { t[1] = now() - t0; dev[1] = MimsemDeviceSeconds(0); }
        if (rep == 1) {
            PetscCompatSetRank(0, np);
            PetscScalar* a;
            VecGetArray(R[0].y[nk - 1], &a);
            for (int i = 0; i < R[0].topo->n1l; i++) {
                diff += (a[i] - keep[i]) * (a[i] - keep[i]);
                norm += keep[i] * keep[i];
            }
            VecRestoreArray(R[0].y[nk - 1], &a);
        }
    }
    std::printf("{\"workload\": \"Umat through the MatShell, sphere p=%d ne=%d nk=%d, 6 emulated ranks on one GPU\", \"dof_levels\": %.0f, "
                "\"per_level_loop_ms\": %.3f, \"per_level_loop_device_ms\": %.3f, \"all_levels_call_ms\": %.3f, \"all_levels_call_device_ms\": %.3f, "
                "\"per_level_loop_device_gdofs\": %.3f, \"all_levels_call_device_gdofs\": %.3f, \"rel_diff_batched_vs_per_level\": %.2e}\n",
                p, ne, nk, dofs, t[0] * 1e3, dev[0] * 1e3, t[1] * 1e3, dev[1] * 1e3, dofs / dev[0] / 1e9, dofs / dev[1] / 1e9,
                std::sqrt(diff / std::max(norm, 1e-300)));
    return 0;
}
