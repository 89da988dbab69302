// Topology and geometry of the cubed sphere / doubly periodic box (host side).
//
// Replaces, for the horizontal operator path:
//   scr/Proc2.py, scr/ProcBox.py      (offline global numbering + seam stitching)
//   scr/Geom2.py, scr/GeomBox.py      (node coordinates)
//   {src,eul,box}/Topo.cpp            (per-rank index maps read from input/*.txt)
//   {src,eul,box}/Geom.cpp            (coordinate fix-up, Jacobians, determinants)
// Everything here is closed-form: no input files are needed, but the reference's file formats
// can be written (write_input_files) and read (load_patch_files) for compatibility.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace mimsem {

enum MeshKind { MESH_SPHERE = 0, MESH_BOX = 1 };

// What one reference MPI rank's Topo holds (eul/Topo.cpp:15-156): ghosted local -> global maps.
struct PatchTopo {
    int p = 0;       // element order
    int nelx = 0;    // elements per patch side
    int nx = 0;      // p*nelx degrees of freedom per patch side
    int n0 = 0, n1x = 0, n1y = 0, n2 = 0;          // local sizes incl. east/north ghosts
    int n0l = 0, n1xl = 0, n1yl = 0, n2l = 0;      // owned sizes (local_sizes_*.txt)
    int64_t N0 = 0, N1 = 0, N2 = 0;                // global sizes
    std::vector<int> loc0, loc1x, loc1y, loc2;
};

// Closed-form restatement of scr/Proc2.py:52-230,404-601 (sphere, nprocs = 6*npx^2) and
// scr/ProcBox.py:44-136,186-243 (box, nprocs = npx^2).  `order` is the polynomial order of the
// grid being numbered (the element order for Topo, the quadrature order for Geom's quads_*).
bool patch_topology(MeshKind kind, int order, int ne, int nprocs, int rank, PatchTopo& out, std::string* err);

// Read one rank's maps from input/{nodes,edges_x,edges_y,faces,local_sizes}_RRRR.txt + grid_res.txt.
bool load_patch_files(const std::string& dir, int nprocs, int rank, MeshKind kind, PatchTopo& out, std::string* err);

// Global (all faces) mesh in the canonical numbering: nprocs = 6 on the sphere, 1 on the box.
struct GlobalMesh {
    MeshKind kind = MESH_SPHERE;
    int p = 0, m = 0, ne = 0, nfaces = 0;
    int64_t nel = 0, N0 = 0, N1 = 0, N2 = 0, NQ = 0;
    double radius = 6371220.0;   // eul/Geom.cpp:20 RAD_SPHERE
    double lx = 1000.0;          // box/Geom.cpp:20 _LX
    // element -> global DOF tables, element e = face*ne^2 + ey*ne + ex
    std::vector<int> el0;    // [nel][(p+1)^2]  nodes,   j = iy*(p+1)+ix
    std::vector<int> el1x;   // [nel][p*(p+1)]  x-normal edges, j = iy*(p+1)+ix  (ids in the 1-form numbering)
    std::vector<int> el1y;   // [nel][(p+1)*p]  y-normal edges, j = iy*p+ix
    std::vector<int> el2;    // [nel][p^2]      faces,   j = iy*p+ix
    std::vector<int> elq;    // [nel][(m+1)^2]  quadrature points
    std::vector<double> xyz; // [NQ][3] cartesian coordinates of the quadrature points (as generated, before
                             //         the per-element re-projection of eul/Geom.cpp:682-724)
    std::vector<double> J;   // [nel][(m+1)^2][4]   J00 J01 J10 J11
    std::vector<double> det; // [nel][(m+1)^2]
    bool signed_det = false; // src/Geom.cpp:251 keeps the sign, eul/box take fabs (eul/Geom.cpp:325)
};

bool build_global_mesh(MeshKind kind, int p, int m, int ne, bool signed_det, GlobalMesh& out, std::string* err);

// Geometry of one reference rank's patch from the coordinates of its quadrature points (what Geom reads from
// geom_RRRR.txt, local row-major (m*nelx+1)^2 grid): the per-element re-projection of eul/Geom.cpp:682-724 (xl is
// updated in place, like Geom::x) followed by the Jacobians of eul/Geom.cpp:245-326 (box: box/Geom.cpp:132-143,
// ne_side = elements per box side).  J[nel][(m+1)^2][4], det[nel][(m+1)^2], optional (lon, lat) per point.
void patch_geometry(MeshKind kind, int m, int nelx, int ne_side, double radius, double lx, bool signed_det,
                    std::vector<double>& xl, std::vector<double>& J, std::vector<double>& det, std::vector<double>* lonlat_out);

// Node coordinates in global numbering, restating scr/Geom2.py:10-277 / scr/GeomBox.py:9-74.
void sphere_node_coords(int order, int ne, double radius, std::vector<double>& xyz);
void box_node_coords(int order, int ne, double lx, std::vector<double>& xyz);

// Write the reference's input/*.txt set (scr/Setup.py:42-78 formats) for `nprocs` ranks.
bool write_input_files(MeshKind kind, int p, int m, int ne, int nprocs, const std::string& dir, std::string* err);

}  // namespace mimsem
