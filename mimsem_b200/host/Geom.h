// Host mirror of the reference's Geom (eul/Geom.h:1-56): coordinates, Jacobians, determinants, layer
// thicknesses and the DOF -> quadrature-point interpolations, with the reference's member names.
// The field writers write0/1/2 produce the ASCII / PETSc-binary files of a build without HDF5.
#ifndef MIMSEM_HOST_GEOM_H
#define MIMSEM_HOST_GEOM_H

#include "Basis.h"
#include "Topo.h"

typedef double(TopogFunc)(double* xi);
typedef double(LevelFunc)(double* xi, int ki);

class Geom {
    public:
        Geom(Topo* _topo, int _nk);          // eul/, box/
        Geom(Topo* _topo);                   // src/  (signed determinant, no levels)
        ~Geom();
        int pi;
        int nl;            // number of local quadrature points
        int nk;
        int nDofsX;        // quadrature points per patch side - 1
        int nDofs0G;
        int n0, n0l;
        int* loc0;         // global ids of the local quadrature points (quads_RRRR.txt)
        int* inds0_l;
        int* inds0_g;
        double** x;        // [nl][3]
        double** s;        // [nl][2]  (lon, lat)
        double** det;      // [nel][mp12]
        double**** J;      // [nel][mp12][2][2]
        double* topog;
        double** levs;     // [nk+1][n0]
        double** thick;    // [nk][n0]
        double** thickInv; // [nk][n0]
        Topo* topo;
        IS is_l_0, is_g_0;
        VecScatter gtol_0;   // quadrature-point vectors: global (MPI) <-> ghosted local (eul/Geom.cpp:107-113)
        GaussLobatto* quad;
        LagrangeNode* node;
        LagrangeEdge* edge;
        void interp0(int ex, int ey, int px, int py, double* vec, double* val);
        void interp1_l(int ex, int ey, int px, int py, double* vec, double* val);
        void interp2_l(int ex, int ey, int px, int py, double* vec, double* val);
        void interp1_g(int ex, int ey, int px, int py, double* vec, double* val);
        void interp2_g(int ex, int ey, int px, int py, double* vec, double* val);
        void initTopog(TopogFunc* ft, LevelFunc* fl);
        // fields interpolated to the quadrature points -> output/<field>_<lev>_<step>.dat (ASCII VecView) and, for 1- and
        // 2-forms, the vector itself -> .vec (PETSc binary, the restart format)        eul/Geom.cpp:419-631
        void write0(Vec q, char* fieldname, int tstep, int lev);
        void write1(Vec u, char* fieldname, int tstep, int lev);
        void write2(Vec h, char* fieldname, int tstep, int lev, bool vert_scale);
        // src/Geom.h: no levels -- output/<field>_<step>.dat (the ASCII branch; the reference's optional HDF5 branch is not provided)
        void write0(Vec q, char* fieldname, int tstep);
        void write1(Vec u, char* fieldname, int tstep);
        void write2(Vec h, char* fieldname, int tstep);
        // per-element vertical vectors (L2Vecs::vz: vecs[element][level p^2 + i]) relabelled to one 2-form per level and
        // written with write2, levels 0 .. nv-1                                        eul/Geom.cpp:633-679
        void writeVertToHoriz(Vec* vecs, char* fieldname, int tstep, int nv);
        int* elInds0_l(int ex, int ey);
        int* elInds0_g(int ex, int ey);
        // flat copies for the device engine
        const double* flatJ() const { return Jflat; }
        const double* flatDet() const { return detflat; }
        unsigned long thick_version;   // bumped by initTopog so that operators re-upload the table
    private:
        void build(bool signed_det);
        double* Jflat;
        double* detflat;
};

#endif
