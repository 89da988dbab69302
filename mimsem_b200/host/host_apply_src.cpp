// Parity driver for the shallow-water (src/) operators of BASELINE config 2 through the C++ host mirror, used
// as src/SWEqn_Picard.cpp uses them (:331 M0h->assemble_up, :449/:468 R->assemble, :567-581 R_up->assemble),
// for every rank of an emulated `mpirun -np nprocs`.
//
//   host_apply_src <input-dir or -> <p> <ne> <nprocs> <fac> <dt> <in.bin> <out.bin>
// in.bin : doubles  x1[N1] x0[N0] h2[N2] q0[N0] u1[N1]     (global numbering)
// out.bin: doubles  RotMat, RotMat_up, Phmat::assemble, Phmat::assemble_up results
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
    RotMat* R; RotMat_up* R_up; Phmat* M0h;
};

int main(int argc, char** argv) {
    if (argc < 9) return 2;
    const char* dir = argv[1];
    const int p = std::atoi(argv[2]), ne = std::atoi(argv[3]), np = std::atoi(argv[4]);
    const double fac = std::atof(argv[5]), dt = std::atof(argv[6]);
    std::vector<double> in = read_all(argv[7]);
    const bool from_files = dir[0] != '-';
    if (from_files && chdir(dir) != 0) { std::perror(dir); return 2; }
    std::vector<Rank> R(np);
    for (int r = 0; r < np; r++) {
        PetscCompatSetRank(r, np);
        Rank& k = R[r];
        k.topo = from_files ? new Topo() : new Topo(0 /* MIMSEM_MESH_SPHERE */, p, ne, 1);
        k.geom = new Geom(k.topo);   // src/: signed determinant, no layers
        k.quad = new GaussLobatto(k.geom->quad->n);
        k.node = new LagrangeNode(k.topo->elOrd, k.quad);
        k.edge = new LagrangeEdge(k.topo->elOrd, k.node);
        k.R = new RotMat(k.topo, k.geom, k.node, k.edge);
        k.R_up = new RotMat_up(k.topo, k.geom, k.node, k.edge);
        k.M0h = new Phmat(k.topo, k.geom, k.node);
    }
    const long N0 = R[0].topo->nDofs0G, N1 = R[0].topo->nDofs1G, N2 = R[0].topo->nDofs2G;
    if ((long)in.size() != 2 * N1 + 2 * N0 + N2) { std::fprintf(stderr, "bad input size\n"); return 2; }
    const double* x1 = in.data();
    const double* x0 = x1 + N1;
    const double* h2 = x0 + N0;
    const double* q0 = h2 + N2;
    const double* u1 = q0 + N0;
    FILE* out = std::fopen(argv[8], "wb");
    if (!out) { std::perror(argv[8]); return 2; }
    std::vector<Vec> v0(np), v1(np), w0(np), w1(np), hv(np), qv(np), uv(np), ql(np), ul(np);
    for (int r = 0; r < np; r++) {
        PetscCompatSetRank(r, np);
        Topo* t = R[r].topo;
        VecCreateMPI(MPI_COMM_WORLD, t->n0l, t->nDofs0G, &v0[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n1l, t->nDofs1G, &v1[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n0l, t->nDofs0G, &w0[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n1l, t->nDofs1G, &w1[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n2l, t->nDofs2G, &hv[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n0l, t->nDofs0G, &qv[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n1l, t->nDofs1G, &uv[r]);
        VecCreateSeq(MPI_COMM_SELF, t->n0, &ql[r]);
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
    fill(v1, x1); fill(v0, x0); fill(hv, h2); fill(qv, q0); fill(uv, u1);
    for (int r = 0; r < np; r++) {   // ghosted local coefficient vectors, as src/SWEqn_Picard.cpp:323-329
        PetscCompatSetRank(r, np);
        VecScatterBegin(R[r].topo->gtol_1, uv[r], ul[r], INSERT_VALUES, SCATTER_FORWARD);
        VecScatterEnd(R[r].topo->gtol_1, uv[r], ul[r], INSERT_VALUES, SCATTER_FORWARD);
        VecScatterBegin(R[r].topo->gtol_0, qv[r], ql[r], INSERT_VALUES, SCATTER_FORWARD);
        VecScatterEnd(R[r].topo->gtol_0, qv[r], ql[r], INSERT_VALUES, SCATTER_FORWARD);
    }
#define ALL_RANKS(stmt) for (int r = 0; r < np; r++) { PetscCompatSetRank(r, np); Rank& k = R[r]; stmt; }
    ALL_RANKS(k.R->assemble(ql[r]); MatMult(k.R->M, v1[r], w1[r]))                          dump(w1);
    ALL_RANKS(k.R_up->assemble(ql[r], ul[r], fac, dt); MatMult(k.R_up->M, v1[r], w1[r]))    dump(w1);
    ALL_RANKS(k.M0h->assemble(hv[r]); MatMult(k.M0h->M, v0[r], w0[r]))                      dump(w0);
    ALL_RANKS(k.M0h->assemble_up(ul[r], hv[r], fac, dt); MatMult(k.M0h->M, v0[r], w0[r]))   dump(w0);
    std::fclose(out);
    for (int r = 0; r < np; r++) {
        PetscCompatSetRank(r, np);
        Rank& k = R[r];
        VecDestroy(&v0[r]); VecDestroy(&v1[r]); VecDestroy(&w0[r]); VecDestroy(&w1[r]); VecDestroy(&hv[r]); VecDestroy(&qv[r]);
        VecDestroy(&uv[r]); VecDestroy(&ql[r]); VecDestroy(&ul[r]);
        delete k.M0h; delete k.R_up; delete k.R;
        delete k.edge; delete k.node; delete k.quad; delete k.geom; delete k.topo;
    }
    std::printf("host_apply_src ok: %d ranks\n", np);
    return 0;
}
