// Parity driver for the matrix-free twins and the remaining coefficient operators through the C++ host mirror, used as
// eul/HorizSolve.cpp uses them: Uvec::assemble / assemble_hu (diagnose_fluxes, :298-306), UtQWmat (Rh, :575-604), Pvec,
// Phvec, WmatInv, WhmatInv on the six ranks of an emulated `mpirun -np 6`; then, on the doubly periodic box (one rank: a
// Krylov solve is a collective and cannot be played rank after rank), KSPSolve(ksp1, ...) on Umat::M as :77-84, 224 do.
//
//   host_apply_twins <p> <ne> <nk> <in.bin> <out.bin>
// in.bin : doubles  thick[nk][N0] x1[nk][N1] x1b[nk][N1] x2[nk][N2] h2[nk][N2] h2b[nk][N2] u1[nk][N1] ex2[nk][N2]  (global numbering)
// out.bin: doubles  per level: Uvec::assemble, Uvec::assemble_hu (4 terms), UtQWmat, Pvec, Phvec, WmatInv, WhmatInv (rho = h2b),
//                   Umat_ray (exner = ex2[lev], exner_s = ex2[0], dt = 300, as eul/Euler_2.cpp:1218-1229 calls it),
//                   Umat + Umat_ray through MatAXPY(M1->M, 1.0, M1ray->M, DIFFERENT_NONZERO_PATTERN) (eul/Euler_2.cpp:1229), and the
//                   plain Umat again after the next assemble() (which drops the added term);
//                   then { its, |x - x_true| / |x_true|, its with PCJACOBI } of the box solve, { its, error } of KSPSolve on the Pmat shell, and the box Pvec's vg, vg1 [N0 box each]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "Assembly.h"

static std::vector<double> read_all(const char* fn) {
    FILE* f = std::fopen(fn, "rb");
    if (!f) { std::perror(fn); std::exit(2); }
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f) / 8;
    std::fseek(f, 0, SEEK_SET);
    std::vector<double> v(n);
    if (std::fread(v.data(), 8, n, f) != (size_t)n) std::exit(2);
    std::fclose(f);
    return v;
}

int main(int argc, char** argv) {
    if (argc < 6) return 2;
    const int p = std::atoi(argv[1]), ne = std::atoi(argv[2]), nk = std::atoi(argv[3]);
    std::vector<double> in = read_all(argv[4]);
    const int np = 6;
    struct Rank { Topo* topo; Geom* geom; GaussLobatto* quad; LagrangeNode* node; LagrangeEdge* edge;
                  Uvec* m1; UtQWmat* Rh; Pvec* m0; Phvec* m0h; WmatInv* Wi; WhmatInv* Whi; Umat_ray* ray; Umat* M1; };
    std::vector<Rank> R(np);
    for (int r = 0; r < np; r++) {
        PetscCompatSetRank(r, np);
        Rank& k = R[r];
        k.topo = new Topo(0, p, ne, nk);
        k.geom = new Geom(k.topo, nk);
        k.quad = new GaussLobatto(k.geom->quad->n);
        k.node = new LagrangeNode(k.topo->elOrd, k.quad);
        k.edge = new LagrangeEdge(k.topo->elOrd, k.node);
    }
    const long N0 = R[0].topo->nDofs0G, N1 = R[0].topo->nDofs1G, N2 = R[0].topo->nDofs2G;
    if ((long)in.size() != (long)nk * (N0 + 3 * N1 + 4 * N2)) { std::fprintf(stderr, "bad input size\n"); return 2; }
    const double* thick = in.data();
    const double* x1 = thick + (long)nk * N0;
    const double* x1b = x1 + (long)nk * N1;
    const double* x2 = x1b + (long)nk * N1;
    const double* h2 = x2 + (long)nk * N2;
    const double* h2b = h2 + (long)nk * N2;
    const double* u1 = h2b + (long)nk * N2;
    const double* ex2 = u1 + (long)nk * N1;
    for (int r = 0; r < np; r++) {
        PetscCompatSetRank(r, np);
        Rank& k = R[r];
        for (int lev = 0; lev < nk; lev++)
            for (int i = 0; i < k.geom->n0; i++) {
                k.geom->thick[lev][i] = thick[(long)lev * N0 + k.geom->loc0[i]];
                k.geom->thickInv[lev][i] = 1.0 / k.geom->thick[lev][i];
            }
        k.geom->thick_version++;
        k.m1 = new Uvec(k.topo, k.geom, k.node, k.edge);
        k.Rh = new UtQWmat(k.topo, k.geom, k.node, k.edge);
        k.m0 = new Pvec(k.topo, k.geom, k.node);
        k.m0h = new Phvec(k.topo, k.geom, k.node);
        k.Wi = new WmatInv(k.topo, k.geom, k.edge);
        k.Whi = new WhmatInv(k.topo, k.geom, k.edge);
        k.ray = new Umat_ray(k.topo, k.geom, k.node, k.edge);
        k.M1 = new Umat(k.topo, k.geom, k.node, k.edge);
    }
    FILE* out = std::fopen(argv[5], "wb");
    if (!out) { std::perror(argv[5]); return 2; }
    std::vector<Vec> g1(np), g1b(np), g2(np), gh(np), ghb(np), gu(np), l1(np), l1b(np), lu(np), w1(np), w2(np), gex(np), gex0(np);
    for (int r = 0; r < np; r++) {
        PetscCompatSetRank(r, np);
        Topo* t = R[r].topo;
        VecCreateMPI(MPI_COMM_WORLD, t->n1l, t->nDofs1G, &g1[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n1l, t->nDofs1G, &g1b[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n2l, t->nDofs2G, &g2[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n2l, t->nDofs2G, &gh[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n2l, t->nDofs2G, &ghb[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n2l, t->nDofs2G, &gex[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n2l, t->nDofs2G, &gex0[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n1l, t->nDofs1G, &gu[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n1l, t->nDofs1G, &w1[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n2l, t->nDofs2G, &w2[r]);
        VecCreateSeq(MPI_COMM_SELF, t->n1, &l1[r]);
        VecCreateSeq(MPI_COMM_SELF, t->n1, &l1b[r]);
        VecCreateSeq(MPI_COMM_SELF, t->n1, &lu[r]);
    }
    auto fill = [&](std::vector<Vec>& v, const double* src) {
        for (int r = 0; r < np; r++) {
            PetscCompatSetRank(r, np);
            PetscScalar* a;
            PetscInt lo, hi;
            VecGetOwnershipRange(v[r], &lo, &hi);
            VecGetArray(v[r], &a);
            for (int i = lo; i < hi; i++) a[i - lo] = src[i];
            VecRestoreArray(v[r], &a);
        }
    };
    auto ghost = [&](std::vector<Vec>& g, std::vector<Vec>& l) {
        for (int r = 0; r < np; r++) {
            PetscCompatSetRank(r, np);
            VecScatterBegin(R[r].topo->gtol_1, g[r], l[r], INSERT_VALUES, SCATTER_FORWARD);
            VecScatterEnd(R[r].topo->gtol_1, g[r], l[r], INSERT_VALUES, SCATTER_FORWARD);
        }
    };
    auto dump = [&](int which) {   // 0: Uvec::vg, 1: w1, 2: w2, 3: Pvec::vg, 4: Phvec::vg
        for (int r = 0; r < np; r++) {
            PetscCompatSetRank(r, np);
            Vec v = which == 0 ? R[r].m1->vg : (which == 1 ? w1[r] : (which == 2 ? w2[r] : (which == 3 ? R[r].m0->vg : R[r].m0h->vg)));
            PetscScalar* a;
            PetscInt n;
            VecGetLocalSize(v, &n);
            VecGetArray(v, &a);
            std::fwrite(a, 8, n, out);
            VecRestoreArray(v, &a);
        }
    };
#define ALL_RANKS(stmt) for (int r = 0; r < np; r++) { PetscCompatSetRank(r, np); Rank& k = R[r]; stmt; }
    for (int lev = 0; lev < nk; lev++) {
        fill(g1, x1 + (long)lev * N1);
        fill(g1b, x1b + (long)lev * N1);
        fill(g2, x2 + (long)lev * N2);
        fill(gh, h2 + (long)lev * N2);
        fill(ghb, h2b + (long)lev * N2);
        fill(gu, u1 + (long)lev * N1);
        fill(gex, ex2 + (long)lev * N2);
        fill(gex0, ex2);
        ghost(g1, l1);
        ghost(g1b, l1b);
        ghost(gu, lu);
        ALL_RANKS(k.m1->assemble(lev, SCALE, true, l1[r]))                                                  dump(0);
        // eul/HorizSolve.cpp:298-305
        ALL_RANKS(VecZeroEntries(k.m1->vl); VecZeroEntries(k.m1->vg);
                  k.m1->assemble_hu(lev, SCALE, l1[r], gh[r], false, 1.0 / 3.0);
                  k.m1->assemble_hu(lev, SCALE, l1[r], ghb[r], false, 1.0 / 6.0);
                  k.m1->assemble_hu(lev, SCALE, l1b[r], gh[r], false, 1.0 / 6.0);
                  k.m1->assemble_hu(lev, SCALE, l1b[r], ghb[r], false, 1.0 / 3.0))
        ALL_RANKS(VecScatterBegin(k.topo->gtol_1, k.m1->vl, k.m1->vg, ADD_VALUES, SCATTER_REVERSE);
                  VecScatterEnd(k.topo->gtol_1, k.m1->vl, k.m1->vg, ADD_VALUES, SCATTER_REVERSE))          dump(0);
        ALL_RANKS(k.Rh->assemble(lu[r], SCALE); MatMult(k.Rh->M, g2[r], w1[r]))                             dump(1);
        ALL_RANKS(k.m0->assemble(lev, SCALE))                                                               dump(3);
        ALL_RANKS(k.m0h->assemble(gh[r], lev, SCALE))                                                       dump(4);
        ALL_RANKS(k.Wi->assemble(lev, SCALE); MatMult(k.Wi->M, g2[r], w2[r]))                               dump(2);
        ALL_RANKS(k.Whi->assemble(ghb[r], lev, SCALE); MatMult(k.Whi->M, g2[r], w2[r]))                     dump(2);
        ALL_RANKS(k.ray->assemble(lev, SCALE, 300.0, gex[r], gex0[r]))
        ALL_RANKS(MatMult(k.ray->M, g1[r], w1[r]))                                                          dump(1);
        // eul/Euler_2.cpp:1226-1231
        ALL_RANKS(k.M1->assemble(lev, SCALE, true); k.ray->assemble(lev, SCALE, 300.0, gex[r], gex0[r]);
                  if (MatAXPY(k.M1->M, 1.0, k.ray->M, DIFFERENT_NONZERO_PATTERN)) return 1;
                  MatAssemblyBegin(k.M1->M, MAT_FINAL_ASSEMBLY); MatAssemblyEnd(k.M1->M, MAT_FINAL_ASSEMBLY))
        ALL_RANKS(MatMult(k.M1->M, g1[r], w1[r]))                                                           dump(1);
        ALL_RANKS(k.M1->assemble(lev, SCALE, true); MatMult(k.M1->M, g1[r], w1[r]))                         dump(1);
    }
    // ---- KSPSolve on the box: x -> b = M1 x -> KSPSolve(M1, b) recovers x (GMRES + block Jacobi requested, as the reference does)
    {
        PetscCompatReset();
        PetscCompatSetRank(0, 1);
        Topo* bt = new Topo(1 /* MIMSEM_MESH_BOX */, 3, 4, 2);
        Geom* bg = new Geom(bt, 2);
        for (int lev = 0; lev < 2; lev++)
            for (int i = 0; i < bg->n0; i++) {
                bg->thick[lev][i] = 750.0 * (1.0 + 0.05 * ((bg->loc0[i] * 7 + lev) % 5));   // a function of the GLOBAL point
                bg->thickInv[lev][i] = 1.0 / bg->thick[lev][i];
            }
        bg->thick_version++;
        GaussLobatto* q = new GaussLobatto(bg->quad->n);
        LagrangeNode* n = new LagrangeNode(bt->elOrd, q);
        LagrangeEdge* e = new LagrangeEdge(bt->elOrd, n);
        Umat* M1 = new Umat(bt, bg, n, e);
        KSP ksp1;
        PC pc;
        KSPCreate(MPI_COMM_WORLD, &ksp1);
        KSPSetOperators(ksp1, M1->M, M1->M);
        KSPSetTolerances(ksp1, 1.0e-14, 1.0e-50, PETSC_DEFAULT, 1000);
        KSPSetType(ksp1, KSPGMRES);
        KSPGetPC(ksp1, &pc);
        PCSetType(pc, PCBJACOBI);
        PCBJacobiSetTotalBlocks(pc, bt->nElsX * bt->nElsX, NULL);
        KSPSetOptionsPrefix(ksp1, "ksp1_");
        KSPSetFromOptions(ksp1);
        Vec x, b, s;
        VecCreateMPI(MPI_COMM_WORLD, bt->n1l, bt->nDofs1G, &x);
        VecCreateMPI(MPI_COMM_WORLD, bt->n1l, bt->nDofs1G, &b);
        VecCreateMPI(MPI_COMM_WORLD, bt->n1l, bt->nDofs1G, &s);
        PetscScalar* a;
        VecGetArray(x, &a);
        for (int i = 0; i < bt->n1l; i++) a[i] = std::sin(0.37 * i) + 0.25 * std::cos(1.3 * i);
        VecRestoreArray(x, &a);
        MatMult(M1->M, x, b);
        VecZeroEntries(s);
        KSPSolve(ksp1, b, s);
        PetscInt its;
        KSPGetIterationNumber(ksp1, &its);
        double en, xn;
        VecAXPY(s, -1.0, x);
        VecNorm(s, NORM_2, &en);
        VecNorm(x, NORM_2, &xn);
        // the same solve with the shell's diagonal instead of its element blocks
        PCSetType(pc, PCJACOBI);
        VecZeroEntries(s);
        KSPSolve(ksp1, b, s);
        PetscInt its_diag;
        KSPGetIterationNumber(ksp1, &its_diag);
        const double res[3] = {(double)its, en / xn, (double)its_diag};
        std::fwrite(res, 8, 3, out);
        KSPDestroy(&ksp1);
        {   // KSPSolve(ksp0, ...) on the Pmat shell with the block-Jacobi request of eul/HorizSolve.cpp:87-96: M0 is diagonal
            Pmat* M0 = new Pmat(bt, bg, n);
            M0->assemble(0, SCALE);
            KSP ksp0;
            PC pc0;
            KSPCreate(MPI_COMM_WORLD, &ksp0);
            KSPSetOperators(ksp0, M0->M, M0->M);
            KSPSetTolerances(ksp0, 1.0e-14, 1.0e-50, PETSC_DEFAULT, 1000);
            KSPSetType(ksp0, KSPGMRES);
            KSPGetPC(ksp0, &pc0);
            PCSetType(pc0, PCBJACOBI);
            PCBJacobiSetTotalBlocks(pc0, bt->nElsX * bt->nElsX, NULL);
            Vec x0, b0, s0;
            VecCreateMPI(MPI_COMM_WORLD, bt->n0l, bt->nDofs0G, &x0);
            VecCreateMPI(MPI_COMM_WORLD, bt->n0l, bt->nDofs0G, &b0);
            VecCreateMPI(MPI_COMM_WORLD, bt->n0l, bt->nDofs0G, &s0);
            PetscScalar* pa;
            VecGetArray(x0, &pa);
            for (int i = 0; i < bt->n0l; i++) pa[i] = std::cos(0.21 * i) + 0.5;
            VecRestoreArray(x0, &pa);
            MatMult(M0->M, x0, b0);
            VecZeroEntries(s0);
            KSPSolve(ksp0, b0, s0);
            PetscInt its0;
            KSPGetIterationNumber(ksp0, &its0);
            double e0n, x0n;
            VecAXPY(s0, -1.0, x0);
            VecNorm(s0, NORM_2, &e0n);
            VecNorm(x0, NORM_2, &x0n);
            const double res0[2] = {(double)its0, e0n / x0n};
            std::fwrite(res0, 8, 2, out);
            KSPDestroy(&ksp0);
            VecDestroy(&x0); VecDestroy(&b0); VecDestroy(&s0);
            delete M0;
        }
        {   // box/Assembly.cpp:357-372: the Pvec constructor assembles vg (SCALE) and vg1 (scale 1) at level 0
            Pvec* m0 = new Pvec(bt, bg, n);
            PetscScalar* pa;
            VecGetArray(m0->vg, &pa);
            std::fwrite(pa, 8, bt->n0l, out);
            VecRestoreArray(m0->vg, &pa);
            VecGetArray(m0->vg1, &pa);
            std::fwrite(pa, 8, bt->n0l, out);
            VecRestoreArray(m0->vg1, &pa);
            delete m0;
        }
        VecDestroy(&x); VecDestroy(&b); VecDestroy(&s);
        delete M1; delete e; delete n; delete q; delete bg; delete bt;
    }
    std::fclose(out);
    std::printf("host_apply_twins ok: %d ranks, %d levels\n", np, nk);
    return 0;
}
