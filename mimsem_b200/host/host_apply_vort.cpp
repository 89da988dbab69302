// Parity driver for the horizontal-vorticity / vertical-momentum operators of eul/ through the C++ host mirror
// (Ut_mat, WtQdUdz_mat: eul/Euler_2.cpp:79-81, eul/HorizSolve.cpp:67), every rank of an emulated `mpirun -np nprocs`.
//
//   host_apply_vort <input-dir or -> <p> <ne> <nprocs> <nk> <in.bin> <out.bin>
// in.bin : doubles  thick[nk][N0] x1[nk][N1] h2[nk][N2] u1[nk][N1]      (global numbering)
// out.bin: doubles  per level: Ut_mat::assemble (levels < nk-1 only), Ut_mat::assemble_h, WtQdUdz_mat results
#include <unistd.h>

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

struct Rank {
    Topo* topo; Geom* geom; GaussLobatto* quad; LagrangeNode* node; LagrangeEdge* edge;
    Ut_mat* M1t; WtQdUdz_mat* Rz;
};

int main(int argc, char** argv) {
    if (argc < 8) return 2;
    const char* dir = argv[1];
    const int p = std::atoi(argv[2]), ne = std::atoi(argv[3]), np = std::atoi(argv[4]), nk = std::atoi(argv[5]);
    std::vector<double> in = read_all(argv[6]);
    const bool from_files = dir[0] != '-';
    if (from_files && chdir(dir) != 0) { std::perror(dir); return 2; }
    std::vector<Rank> R(np);
    for (int r = 0; r < np; r++) {
        PetscCompatSetRank(r, np);
        Rank& k = R[r];
        k.topo = from_files ? new Topo(nk) : new Topo(0 /* MIMSEM_MESH_SPHERE */, p, ne, nk);
        k.geom = new Geom(k.topo, nk);
        k.quad = new GaussLobatto(k.geom->quad->n);
        k.node = new LagrangeNode(k.topo->elOrd, k.quad);
        k.edge = new LagrangeEdge(k.topo->elOrd, k.node);
    }
    const long N0 = R[0].topo->nDofs0G, N1 = R[0].topo->nDofs1G, N2 = R[0].topo->nDofs2G;
    if ((long)in.size() != (long)nk * (N0 + 2 * N1 + N2)) { std::fprintf(stderr, "bad input size\n"); return 2; }
    const double* thick = in.data();
    const double* x1 = thick + (long)nk * N0;
    const double* h2 = x1 + (long)nk * N1;
    const double* u1 = h2 + (long)nk * N2;
    for (int r = 0; r < np; r++) {
        PetscCompatSetRank(r, np);
        Rank& k = R[r];
        for (int lev = 0; lev < nk; lev++)
            for (int i = 0; i < k.geom->n0; i++) {
                k.geom->thick[lev][i] = thick[(long)lev * N0 + k.geom->loc0[i]];
                k.geom->thickInv[lev][i] = 1.0 / k.geom->thick[lev][i];
            }
        k.geom->thick_version++;
        k.M1t = new Ut_mat(k.topo, k.geom, k.node, k.edge);
        k.Rz = new WtQdUdz_mat(k.topo, k.geom, k.node, k.edge);
    }
    FILE* out = std::fopen(argv[7], "wb");
    if (!out) { std::perror(argv[7]); return 2; }
    std::vector<Vec> v1(np), w1(np), w2(np), hv(np), uv(np), ul(np);
    for (int r = 0; r < np; r++) {
        PetscCompatSetRank(r, np);
        Topo* t = R[r].topo;
        VecCreateMPI(MPI_COMM_WORLD, t->n1l, t->nDofs1G, &v1[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n1l, t->nDofs1G, &w1[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n2l, t->nDofs2G, &w2[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n2l, t->nDofs2G, &hv[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n1l, t->nDofs1G, &uv[r]);
        VecCreateSeq(MPI_COMM_SELF, t->n1, &ul[r]);
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
    auto dump = [&](std::vector<Vec>& v) {
        for (int r = 0; r < np; r++) {
            PetscCompatSetRank(r, np);
            PetscScalar* a;
            PetscInt n;
            VecGetLocalSize(v[r], &n);
            VecGetArray(v[r], &a);
            std::fwrite(a, 8, n, out);
            VecRestoreArray(v[r], &a);
        }
    };
    for (int lev = 0; lev < nk; lev++) {
        fill(v1, x1 + (long)lev * N1);
        fill(hv, h2 + (long)lev * N2);
        fill(uv, u1 + (long)lev * N1);
        for (int r = 0; r < np; r++) {
            PetscCompatSetRank(r, np);
            VecScatterBegin(R[r].topo->gtol_1, uv[r], ul[r], INSERT_VALUES, SCATTER_FORWARD);
            VecScatterEnd(R[r].topo->gtol_1, uv[r], ul[r], INSERT_VALUES, SCATTER_FORWARD);
        }
#define ALL_RANKS(stmt) for (int r = 0; r < np; r++) { PetscCompatSetRank(r, np); Rank& k = R[r]; stmt; }
        if (lev < nk - 1) {
            ALL_RANKS(k.M1t->assemble(lev, SCALE); MatMult(k.M1t->M, v1[r], w1[r]))           dump(w1);
        }
        ALL_RANKS(k.M1t->assemble_h(lev, SCALE, hv[r]); MatMult(k.M1t->M, v1[r], w1[r]))      dump(w1);
        ALL_RANKS(k.Rz->assemble(ul[r], SCALE); MatMult(k.Rz->M, v1[r], w2[r]))               dump(w2);
    }
    std::fclose(out);
    for (int r = 0; r < np; r++) {
        PetscCompatSetRank(r, np);
        Rank& k = R[r];
        VecDestroy(&v1[r]); VecDestroy(&w1[r]); VecDestroy(&w2[r]); VecDestroy(&hv[r]); VecDestroy(&uv[r]); VecDestroy(&ul[r]);
        delete k.Rz; delete k.M1t;
        delete k.edge; delete k.node; delete k.quad; delete k.geom; delete k.topo;
    }
    std::printf("host_apply_vort ok: %d ranks, %d levels\n", np, nk);
    return 0;
}
