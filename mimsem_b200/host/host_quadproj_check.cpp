// CPU parity driver for the quadrature-point projections of the C++ host mirror -- WtQmat, UtQmat, PtQmat, used as the
// reference initialises its fields (eul/Euler_2.cpp:420-440, 485-500: MatMult(WQ->M, bg, WQb) etc.) -- on the six ranks of an
// emulated `mpirun -np 6`, and for Geom::writeVertToHoriz (eul/Geom.cpp:633-679) against write2 level by level.  No GPU.
//   host_quadproj_check <p> <ne> <in.bin> <out.bin> <scratch dir>
// in.bin : doubles xq[nq] uq[2 nq]      (global quadrature-point numbering; uq: two interleaved components per point)
// out.bin: doubles WtQmat xq [N2], UtQmat uq [N1], PtQmat xq [N0], then 1.0 / 0.0: writeVertToHoriz wrote what write2 writes,
//                   then 1.0 / 0.0: the level-less src/ writers wrote what the eul/ writers write at unit thickness
#include <sys/stat.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "Assembly.h"

static std::vector<double> read_all(const char* fn) {
    FILE* f = std::fopen(fn, "rb");
    if (!f) { std::perror(fn); std::exit(2); }
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f) / 8;
    std::fseek(f, 0, SEEK_SET);
    std::vector<double> v(n);
    if (n && std::fread(v.data(), 8, n, f) != (size_t)n) std::exit(2);
    std::fclose(f);
    return v;
}
static std::vector<char> read_bytes(const std::string& fn) {
    FILE* f = std::fopen(fn.c_str(), "rb");
    std::vector<char> v;
    if (!f) return v;
    int c;
    while ((c = std::fgetc(f)) != EOF) v.push_back((char)c);
    std::fclose(f);
    return v;
}

int main(int argc, char** argv) {
    if (argc < 6) return 2;
    const int p = std::atoi(argv[1]), ne = std::atoi(argv[2]), np = 6, nk = 3;
    std::vector<double> in = read_all(argv[3]);
    if (chdir(argv[5]) != 0) { std::perror(argv[5]); return 2; }
    mkdir("output", 0777);
    struct Rank { Topo* topo; Geom* geom; GaussLobatto* quad; LagrangeNode* node; LagrangeEdge* edge; WtQmat* WQ; UtQmat* UQ; PtQmat* PQ; };
    std::vector<Rank> R(np);
    for (int r = 0; r < np; r++) {
        PetscCompatSetRank(r, np);
        Rank& k = R[r];
        k.topo = new Topo(0, p, ne, nk);
        k.geom = new Geom(k.topo, nk);
        k.quad = new GaussLobatto(k.geom->quad->n);
        k.node = new LagrangeNode(k.topo->elOrd, k.quad);
        k.edge = new LagrangeEdge(k.topo->elOrd, k.node);
        k.WQ = new WtQmat(k.topo, k.geom, k.edge);
        k.UQ = new UtQmat(k.topo, k.geom, k.node, k.edge);
        k.PQ = new PtQmat(k.topo, k.geom, k.node);
    }
    const long NQ = R[0].geom->nDofs0G;
    if ((long)in.size() != 3 * NQ) { std::fprintf(stderr, "bad input size\n"); return 2; }
    std::vector<Vec> xq(np), uq(np), y2(np), y1(np), y0(np);
    for (int r = 0; r < np; r++) {
        PetscCompatSetRank(r, np);
        Topo* t = R[r].topo;
        Geom* g = R[r].geom;
        VecCreateMPI(MPI_COMM_WORLD, g->n0l, g->nDofs0G, &xq[r]);
        VecCreateMPI(MPI_COMM_WORLD, 2 * g->n0l, 2 * g->nDofs0G, &uq[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n2l, t->nDofs2G, &y2[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n1l, t->nDofs1G, &y1[r]);
        VecCreateMPI(MPI_COMM_WORLD, t->n0l, t->nDofs0G, &y0[r]);
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
    FILE* out = std::fopen(argv[4], "wb");
    if (!out) { std::perror(argv[4]); return 2; }
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
    fill(xq, in.data());
    fill(uq, in.data() + NQ);
#define ALL_RANKS(stmt) for (int r = 0; r < np; r++) { PetscCompatSetRank(r, np); Rank& k = R[r]; stmt; }
    ALL_RANKS(MatMult(k.WQ->M, xq[r], y2[r]))   dump(y2);
    ALL_RANKS(MatMult(k.UQ->M, uq[r], y1[r]))   dump(y1);
    ALL_RANKS(MatMult(k.PQ->M, xq[r], y0[r]))   dump(y0);

    // writeVertToHoriz: vertical vectors vz[element][level p^2 + i] built from horizontal 2-forms (what L2Vecs::HorizToVert
    // does, eul/L2Vecs.cpp:55-77) must come out as write2 of those 2-forms, file for file
    double same = 1.0;
    {
        const int n2e = p * p;
        std::vector<std::vector<Vec> > H(np), VZ(np);
        for (int r = 0; r < np; r++) {
            PetscCompatSetRank(r, np);
            Topo* t = R[r].topo;
            for (int lev = 0; lev < nk; lev++) {
                Vec v;
                VecCreateMPI(MPI_COMM_WORLD, t->n2l, t->nDofs2G, &v);
                PetscScalar* a;
                VecGetArray(v, &a);
                for (int i = 0; i < t->n2l; i++) a[i] = std::sin(0.01 * (i + 1) * (lev + 1) + r);
                VecRestoreArray(v, &a);
                H[r].push_back(v);
            }
            for (int ei = 0; ei < t->nElsX * t->nElsX; ei++) {
                Vec v;
                VecCreateSeq(MPI_COMM_SELF, nk * n2e, &v);
                PetscScalar *a, *h;
                VecGetArray(v, &a);
                const int* inds2 = t->elInds2_l(ei % t->nElsX, ei / t->nElsX);
                for (int lev = 0; lev < nk; lev++) {
                    VecGetArray(H[r][lev], &h);
                    for (int i = 0; i < n2e; i++) a[lev * n2e + i] = h[inds2[i]];
                    VecRestoreArray(H[r][lev], &h);
                }
                VecRestoreArray(v, &a);
                VZ[r].push_back(v);
            }
        }
        char fa[8] = "vth", fb[8] = "hor";
        for (int lev = 0; lev < nk; lev++) ALL_RANKS(k.geom->write2(H[r][lev], fb, 7, lev, false))
        ALL_RANKS(k.geom->writeVertToHoriz(VZ[r].data(), fa, 7, nk))
        for (int lev = 0; lev < nk; lev++)
            for (int ext = 0; ext < 2; ext++) {
                char a[200], b[200];
                std::snprintf(a, sizeof a, "output/vth_%.3u_%.4u.%s", lev, 7, ext ? "vec" : "dat");
                std::snprintf(b, sizeof b, "output/hor_%.3u_%.4u.%s", lev, 7, ext ? "vec" : "dat");
                const std::vector<char> A = read_bytes(a), B = read_bytes(b);
                if (A.empty() || A != B) same = 0.0;
            }
    }
    std::fwrite(&same, 8, 1, out);
    // the level-less writers of src/ (src/Geom.cpp:326-520): with unit thickness they write what the eul/ forms write for a level
    double same_src = 1.0;
    {
        std::vector<Vec> q0(np), u1(np), h2(np);
        for (int r = 0; r < np; r++) {
            PetscCompatSetRank(r, np);
            Topo* t = R[r].topo;
            for (int i = 0; i < R[r].geom->n0; i++) R[r].geom->thick[0][i] = 1.0;
            VecCreateMPI(MPI_COMM_WORLD, t->n0l, t->nDofs0G, &q0[r]);
            VecCreateMPI(MPI_COMM_WORLD, t->n1l, t->nDofs1G, &u1[r]);
            VecCreateMPI(MPI_COMM_WORLD, t->n2l, t->nDofs2G, &h2[r]);
            PetscScalar* a;
            VecGetArray(q0[r], &a); for (int i = 0; i < t->n0l; i++) a[i] = std::cos(0.02 * i + r); VecRestoreArray(q0[r], &a);
            VecGetArray(u1[r], &a); for (int i = 0; i < t->n1l; i++) a[i] = std::sin(0.03 * i - r); VecRestoreArray(u1[r], &a);
            VecGetArray(h2[r], &a); for (int i = 0; i < t->n2l; i++) a[i] = 1.0 + 0.5 * std::sin(0.05 * i + r); VecRestoreArray(h2[r], &a);
        }
        char s0[8] = "sq", s1[8] = "su", s2[8] = "sh", e0[8] = "eq", e1[8] = "eu", e2[8] = "eh";
        ALL_RANKS(k.geom->write0(q0[r], s0, 3))  ALL_RANKS(k.geom->write0(q0[r], e0, 3, 0))
        ALL_RANKS(k.geom->write1(u1[r], s1, 3))  ALL_RANKS(k.geom->write1(u1[r], e1, 3, 0))
        ALL_RANKS(k.geom->write2(h2[r], s2, 3))  ALL_RANKS(k.geom->write2(h2[r], e2, 3, 0, true))
        const char* pairs[7][2] = {{"sq_0003.dat", "eq_000_0003.dat"}, {"su_x_0003.dat", "eu_x_000_0003.dat"}, {"su_y_0003.dat", "eu_y_000_0003.dat"},
                                   {"su_0003.vec", "eu_000_0003.vec"}, {"sh_0003.dat", "eh_000_0003.dat"}, {"sh_0003.vec", "eh_000_0003.vec"},
                                   {"sh_0003.dat", "sh_0003.dat"}};
        for (int i = 0; i < 7; i++) {
            const std::vector<char> A = read_bytes(std::string("output/") + pairs[i][0]), B = read_bytes(std::string("output/") + pairs[i][1]);
            if (A.empty() || A != B) same_src = 0.0;
        }
    }
    std::fwrite(&same_src, 8, 1, out);
    std::fclose(out);
    std::printf("host_quadproj_check ok\n");
    return 0;
}
