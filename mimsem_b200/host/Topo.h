// Host mirror of the reference's Topo (eul/Topo.h:1-52; the src/ and box/ variants differ only in the
// constructor signature and in nDofs0G).  Reads the same input/*.txt files relative to the working
// directory, or -- without any files -- generates the maps in closed form (csrc/mesh.hpp).
#ifndef MIMSEM_HOST_TOPO_H
#define MIMSEM_HOST_TOPO_H

#include "petsc_compat.h"

class Topo {
    public:
        Topo(int _nk);                                   // eul/: reads input/, rank from MPI_Comm_rank
        Topo();                                          // src/, box/
        Topo(int kind, int p, int ne, int _nk);          // file-free: kind = MIMSEM_MESH_*, ne = elements per face side
        ~Topo();
        int pi;
        int n0, n1, n1x, n1y, n2;
        int n0l, n1l, n1xl, n1yl, n2l;
        int elOrd, nElsX, nDofsX;
        int nDofs0G, nDofs1G, nDofs2G;
        int *loc0, *loc1, *loc1x, *loc1y, *loc2;
        int *inds0_l, *inds1x_l, *inds1y_l, *inds2_l;
        int *inds0_g, *inds1x_g, *inds1y_g, *inds2_g;
        IS is_l_0, is_g_0, is_l_1, is_g_1;
        VecScatter gtol_0, gtol_1;
        int* elInds0_l(int ex, int ey);
        int* elInds1x_l(int ex, int ey);
        int* elInds1y_l(int ex, int ey);
        int* elInds2_l(int ex, int ey);
        int* elInds0_g(int ex, int ey);
        int* elInds1x_g(int ex, int ey);
        int* elInds1y_g(int ex, int ey);
        int* elInds2_g(int ex, int ey);
        int nk;
        int kind;        // MIMSEM_MESH_SPHERE / MIMSEM_MESH_BOX (extension: the reference hard-codes it per directory)
    private:
        void finish(int nprocs);
};

#endif
