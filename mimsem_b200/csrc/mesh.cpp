// Closed-form cubed-sphere / periodic-box topology and geometry.  See mesh.hpp.
#include "mesh.hpp"

#include <cmath>
#include <cstdio>
#include <fstream>
#include <sstream>

#include "basis.hpp"

namespace mimsem {

namespace {

int isqrt_exact(int v) {
    int r = (int)std::lround(std::sqrt((double)v));
    return (r * r == v) ? r : -1;
}

// Global numbering of one cube (or the periodic square) as functions of face-global grid
// coordinates.  Owned entities follow scr/Proc2.py:73-130:
//   nodes  : face-row-major, independent of the processor count
//   edges  : element-blocked inside each patch, x/y interleaved (2k, 2k+1)
//   faces  : element-blocked inside each patch
// Entities on a face's east column / north row belong to the neighbouring face; the seam rules
// restate scr/Proc2.py:133-230 with the face adjacency and orientation tables of :419-479:
//   even face: east  -> face+2, rotated  (its south row, reversed; y-edges become x-edges)
//              north -> face+1, aligned
//   odd face : east  -> face+1, aligned
//              north -> face+2, rotated  (its west column, reversed; x-edges become y-edges)
// and the two valence-3 corners that are no face's south-west corner get the last two node ids.
struct Numbering {
    MeshKind kind;
    int p, npx, nelx, nx, nxF;
    int64_t n_owned_face;   // nxF^2

    int rank_of(int f, int px, int py) const { return (kind == MESH_SPHERE ? f * npx * npx : 0) + py * npx + px; }

    int64_t blocked(int f, int gx, int gy) const {
        int px = gx / nx, ix = gx % nx, py = gy / nx, iy = gy % nx;
        int64_t blk = (int64_t)((iy / p) * nelx + ix / p) * p * p + (iy % p) * p + ix % p;
        return (int64_t)rank_of(f, px, py) * nx * nx + blk;
    }

    int64_t node(int f, int gx, int gy) const {
        if (kind == MESH_BOX) return (int64_t)(gy % nxF) * nxF + (gx % nxF);
        if (gx < nxF && gy < nxF) return (int64_t)f * n_owned_face + (int64_t)gy * nxF + gx;
        const bool even = (f % 2 == 0);
        if (gx == nxF) {
            if (even) return gy == 0 ? 6 * n_owned_face : node((f + 2) % 6, nxF - gy, 0);
            return gy < nxF ? node((f + 1) % 6, 0, gy) : node((f + 2) % 6, 0, 0);
        }
        // gy == nxF, gx < nxF
        if (even) return node((f + 1) % 6, gx, 0);
        return gx == 0 ? 6 * n_owned_face + 1 : node((f + 2) % 6, 0, nxF - gx);
    }

    int64_t yedge(int f, int gx, int gy) const;
    int64_t xedge(int f, int gx, int gy) const {   // gx in [0,nxF], gy in [0,nxF)
        if (gx < nxF) return 2 * blocked(f, gx, gy);
        if (kind == MESH_BOX) return xedge(f, 0, gy);
        if (f % 2 == 0) return yedge((f + 2) % 6, nxF - 1 - gy, 0);
        return xedge((f + 1) % 6, 0, gy);
    }
    int64_t face(int f, int gx, int gy) const { return blocked(f, gx, gy); }
};

int64_t Numbering::yedge(int f, int gx, int gy) const {   // gx in [0,nxF), gy in [0,nxF]
    if (gy < nxF) return 2 * blocked(f, gx, gy) + 1;
    if (kind == MESH_BOX) return yedge(f, gx, 0);
    if (f % 2 == 0) return yedge((f + 1) % 6, gx, 0);
    return xedge((f + 2) % 6, 0, nxF - 1 - gx);
}

bool make_numbering(MeshKind kind, int order, int ne, int nprocs, Numbering& nb, std::string* err) {
    int per_face = (kind == MESH_SPHERE) ? nprocs / 6 : nprocs;
    if (kind == MESH_SPHERE && nprocs % 6 != 0) {
        if (err) *err = "cubed sphere needs 6*n^2 patches";
        return false;
    }
    int npx = isqrt_exact(per_face);
    if (npx < 1 || ne % npx != 0 || order < 1) {
        if (err) *err = "patch count must be a perfect square per face that divides the elements per side";
        return false;
    }
    nb.kind = kind;
    nb.p = order;
    nb.npx = npx;
    nb.nelx = ne / npx;
    nb.nx = order * nb.nelx;
    nb.nxF = order * ne;
    nb.n_owned_face = (int64_t)nb.nxF * nb.nxF;
    return true;
}

}  // namespace

bool patch_topology(MeshKind kind, int order, int ne, int nprocs, int rank, PatchTopo& out, std::string* err) {
    Numbering nb;
    if (!make_numbering(kind, order, ne, nprocs, nb, err)) return false;
    if (rank < 0 || rank >= nprocs) {
        if (err) *err = "rank out of range";
        return false;
    }
    const int npx = nb.npx, nx = nb.nx;
    const int f = (kind == MESH_SPHERE) ? rank / (npx * npx) : 0;
    const int pj = rank % (npx * npx);
    const int px = pj % npx, py = pj / npx;
    const int nfaces = (kind == MESH_SPHERE) ? 6 : 1;

    out.p = order;
    out.nelx = nb.nelx;
    out.nx = nx;
    out.n0 = (nx + 1) * (nx + 1);
    out.n1x = (nx + 1) * nx;
    out.n1y = nx * (nx + 1);
    out.n2 = nx * nx;
    out.n0l = out.n1xl = out.n1yl = out.n2l = nx * nx;
    if (kind == MESH_SPHERE) {
        // the two hanging nodes are owned by the patches that first meet them (scr/Proc2.py:57-61)
        if (f == 0 && px == npx - 1 && py == 0) out.n0l += 1;
        if (f == 1 && px == 0 && py == npx - 1) out.n0l += 1;
    }
    out.N2 = (int64_t)nfaces * nb.n_owned_face;
    out.N1 = 2 * out.N2;
    out.N0 = out.N2 + (kind == MESH_SPHERE ? 2 : 0);

    out.loc0.resize(out.n0);
    out.loc1x.resize(out.n1x);
    out.loc1y.resize(out.n1y);
    out.loc2.resize(out.n2);
    const int ox = px * nx, oy = py * nx;
    for (int iy = 0; iy <= nx; iy++)
        for (int ix = 0; ix <= nx; ix++) out.loc0[iy * (nx + 1) + ix] = (int)nb.node(f, ox + ix, oy + iy);
    for (int iy = 0; iy < nx; iy++)
        for (int ix = 0; ix <= nx; ix++) out.loc1x[iy * (nx + 1) + ix] = (int)nb.xedge(f, ox + ix, oy + iy);
    for (int iy = 0; iy <= nx; iy++)
        for (int ix = 0; ix < nx; ix++) out.loc1y[iy * nx + ix] = (int)nb.yedge(f, ox + ix, oy + iy);
    for (int iy = 0; iy < nx; iy++)
        for (int ix = 0; ix < nx; ix++) out.loc2[iy * nx + ix] = (int)nb.face(f, ox + ix, oy + iy);
    return true;
}

// ---------------------------------------------------------------------------------------------
// input/*.txt compatibility (formats: scr/Setup.py:42-78; readers: eul/Topo.cpp:27-140)

namespace {

bool read_ints(const std::string& path, std::vector<int>& v) {
    std::ifstream f(path.c_str());
    if (!f) return false;
    v.clear();
    std::string line;
    while (std::getline(f, line)) {
        if (line.empty()) continue;
        v.push_back(std::atoi(line.c_str()));
    }
    return true;
}

std::string rank_file(const std::string& dir, const char* stem, int rank) {
    char buf[64];
    std::snprintf(buf, sizeof buf, "/%s_%04d.txt", stem, rank);
    return dir + buf;
}

bool write_ints(const std::string& path, const int* v, size_t n) {
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) return false;
    for (size_t i = 0; i < n; i++) std::fprintf(f, "%u\n", (unsigned)v[i]);
    std::fclose(f);
    return true;
}

}  // namespace

bool load_patch_files(const std::string& dir, int nprocs, int rank, MeshKind kind, PatchTopo& out, std::string* err) {
    std::vector<int> res, sizes;
    if (!read_ints(dir + "/grid_res.txt", res) || res.size() < 2) {
        if (err) *err = "cannot read " + dir + "/grid_res.txt";
        return false;
    }
    out.p = res[0];
    out.nelx = res[1];
    out.nx = out.p * out.nelx;
    if (!read_ints(rank_file(dir, "nodes", rank), out.loc0) || !read_ints(rank_file(dir, "edges_x", rank), out.loc1x) ||
        !read_ints(rank_file(dir, "edges_y", rank), out.loc1y) || !read_ints(rank_file(dir, "faces", rank), out.loc2) ||
        !read_ints(rank_file(dir, "local_sizes", rank), sizes) || sizes.size() < 4) {
        if (err) *err = "cannot read the per-rank topology files in " + dir;
        return false;
    }
    out.n0 = (int)out.loc0.size();
    out.n1x = (int)out.loc1x.size();
    out.n1y = (int)out.loc1y.size();
    out.n2 = (int)out.loc2.size();
    out.n0l = sizes[0];
    out.n1xl = sizes[1];
    out.n1yl = sizes[2];
    out.n2l = sizes[3];
    // global sizes as eul/Topo.cpp:113-115 (box/Topo.cpp:112 has no hanging nodes)
    out.N2 = (int64_t)nprocs * out.nx * out.nx;
    out.N1 = 2 * out.N2;
    out.N0 = out.N2 + (kind == MESH_SPHERE ? 2 : 0);
    return true;
}

// ---------------------------------------------------------------------------------------------
// coordinates

void sphere_node_coords(int order, int ne, double radius, std::vector<double>& xyz) {
    const int nx = order * ne;
    const double quarter_pi = 0.25 * M_PI;
    std::vector<double> gx, gw;
    gll_rule(order, gx, gw);
    // GLL-spaced equi-angular abscissae across one face (scr/Geom2.py:23-40)
    std::vector<double> X(nx + 1);
    const double dx = 0.5 * M_PI / ne;
    for (int el = 0; el < ne; el++)
        for (int j = 0; j < order; j++) X[el * order + j] = dx * 0.5 * (gx[j] + 1.0) + el * dx - quarter_pi;
    X[nx] = +quarter_pi;

    const int64_t nf = (int64_t)nx * nx;
    xyz.assign((size_t)(6 * nf + 2) * 3, 0.0);
    auto on_face0 = [&](double ax, double ay, double th, double* c) {
        // gnomonic point of face 0 at angles (ax, ay), longitude th  (scr/Geom2.py:53-63)
        double tx = std::tan(ax), ty = std::tan(ay);
        double phi = std::asin(ty / std::sqrt(1.0 + tx * tx + ty * ty));
        c[0] = std::cos(phi) * std::cos(th);
        c[1] = std::cos(phi) * std::sin(th);
        c[2] = std::sin(phi);
    };
    // successive quarter turns carry face 0 onto faces 1..5 (scr/Geom2.py:80-186)
    auto turn = [](int step, const double* a, double* b) {
        switch (step % 3) {
            case 1: b[0] = -a[2]; b[1] = a[1]; b[2] = a[0]; break;    // faces 0->1, 3->4
            case 2: b[0] = a[0]; b[1] = a[2]; b[2] = -a[1]; break;    // faces 1->2, 4->5
            default: b[0] = -a[1]; b[1] = a[0]; b[2] = a[2]; break;   // faces 2->3
        }
    };
    auto expand = [&](double* c) {
        // scr/Geom2.py:262-270: back to (lon, lat), then out to the sphere radius
        double th = std::atan2(c[1], c[0]);
        double ph = std::asin(c[2]);
        c[0] = radius * std::cos(ph) * std::cos(th);
        c[1] = radius * std::cos(ph) * std::sin(th);
        c[2] = radius * std::sin(ph);
    };
    for (int iy = 0; iy < nx; iy++) {
        for (int ix = 0; ix < nx; ix++) {
            double c[3], d[3];
            on_face0(X[ix], X[iy], X[ix], c);
            int64_t k = (int64_t)iy * nx + ix;
            for (int f = 0; f < 6; f++) {
                double e[3] = {c[0], c[1], c[2]};
                expand(e);
                for (int a = 0; a < 3; a++) xyz[(size_t)(f * nf + k) * 3 + a] = e[a];
                turn(f + 1, c, d);
                c[0] = d[0]; c[1] = d[1]; c[2] = d[2];
            }
        }
    }
    // hanging nodes (scr/Geom2.py:65-72, 96-104)
    double h0[3], h1[3], t[3];
    on_face0(X[0], X[0], X[nx], h0);
    on_face0(X[nx], X[nx], X[0], t);
    turn(1, t, h1);
    expand(h0);
    expand(h1);
    for (int a = 0; a < 3; a++) {
        xyz[(size_t)(6 * nf) * 3 + a] = h0[a];
        xyz[(size_t)(6 * nf + 1) * 3 + a] = h1[a];
    }
}

void box_node_coords(int order, int ne, double lx, std::vector<double>& xyz) {
    const int nx = order * ne;
    std::vector<double> gx, gw;
    gll_rule(order, gx, gw);
    const double dx = lx / ne;
    xyz.assign((size_t)nx * nx * 3, 0.0);
    for (int iy = 0; iy < nx; iy++)
        for (int ix = 0; ix < nx; ix++) {
            size_t k = (size_t)iy * nx + ix;
            xyz[k * 3 + 0] = (ix / order) * dx + 0.5 * dx * (1.0 + gx[ix % order]);
            xyz[k * 3 + 1] = (iy / order) * dx + 0.5 * dx * (1.0 + gx[iy % order]);
            xyz[k * 3 + 2] = 0.0;
        }
}

// ---------------------------------------------------------------------------------------------
// Jacobians

namespace {

// Sphere: J = (R / 4|r~|) A(lon) B(lon,lat) C(corners) D(xi)   (Guba et al. 2014; eul/Geom.cpp:245-319).
// `c` holds the element's four corner positions (SW, SE, NE, NW), `lon/lat` the spherical
// coordinates stored for this quadrature point, (x1,x2) its reference abscissae.
void sphere_jacobian(const double* const c[4], double lon, double lat, double x1, double x2, double R, double* J) {
    const double wgt[4] = {(1.0 - x1) * (1.0 - x2), (1.0 + x1) * (1.0 - x2), (1.0 + x1) * (1.0 + x2), (1.0 - x1) * (1.0 + x2)};
    double r[3];
    for (int a = 0; a < 3; a++) r[a] = 0.25 * (wgt[0] * c[0][a] + wgt[1] * c[1][a] + wgt[2] * c[2][a] + wgt[3] * c[3][a]);
    const double rinv = 1.0 / std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    const double sl = std::sin(lon), cl = std::cos(lon), st = std::sin(lat), ct = std::cos(lat);
    const double A[2][3] = {{-sl, +cl, 0.0}, {0.0, 0.0, 1.0}};
    const double B[3][3] = {{+sl * sl * ct * ct + st * st, -0.5 * std::sin(2.0 * lon) * ct * ct, -0.5 * cl * std::sin(2.0 * lat)},
                            {-0.5 * std::sin(2.0 * lon) * ct * ct, +cl * cl * ct * ct + st * st, -0.5 * sl * std::sin(2.0 * lat)},
                            {-cl * st, -sl * st, +ct}};
    const double D[4][2] = {{-1.0 + x2, -1.0 + x1}, {+1.0 - x2, -1.0 - x1}, {+1.0 + x2, +1.0 + x1}, {-1.0 - x2, +1.0 - x1}};
    double AB[2][3], ABC[2][4];
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += A[i][k] * B[k][j];
            AB[i][j] = s;
        }
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 4; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += AB[i][k] * c[j][k];
            ABC[i][j] = s;
        }
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) {
            double s = 0.0;
            for (int k = 0; k < 4; k++) s += ABC[i][k] * D[k][j];
            J[i * 2 + j] = s * (0.25 * R * rinv);
        }
}

}  // namespace

void patch_geometry(MeshKind kind, int m, int nelx, int ne_side, double radius, double lx, bool signed_det,
                    std::vector<double>& xl, std::vector<double>& J, std::vector<double>& det, std::vector<double>* lonlat_out) {
    const int mp1 = m + 1, nqe = mp1 * mp1;
    const size_t nel = (size_t)nelx * nelx;
    J.assign(nel * nqe * 4, 0.0);
    det.assign(nel * nqe, 0.0);
    if (kind == MESH_BOX) {
        // box/Geom.cpp:132-143: constant diagonal Jacobian, half an element width
        const double h = 0.5 * lx / ne_side;
        for (size_t e = 0; e < nel; e++)
            for (int k = 0; k < nqe; k++) {
                double* Jq = &J[(e * nqe + k) * 4];
                Jq[0] = h; Jq[1] = 0.0; Jq[2] = 0.0; Jq[3] = h;
                det[e * nqe + k] = std::fabs(h * h);
            }
        return;
    }
    std::vector<double> qx, qw;
    gll_rule(m, qx, qw);
    // Sphere.  The reference keeps, per rank, a private copy of its quadrature points' coordinates,
    // re-projects every non-corner point of every element from the element's corners
    // (eul/Geom.cpp:682-724; element loop order ey, ex; shared points keep the last writer's value)
    // and only then evaluates the Jacobians (eul/Geom.cpp:726-741).
    const int nxq = m * nelx, nq1 = nxq + 1;
    std::vector<double> lon((size_t)nq1 * nq1), lat((size_t)nq1 * nq1);
    for (int i = 0; i < nq1 * nq1; i++) {
        lon[i] = std::atan2(xl[(size_t)i * 3 + 1], xl[(size_t)i * 3 + 0]);
        lat[i] = std::asin(xl[(size_t)i * 3 + 2] / radius);
    }
    auto corner_ids = [&](int ex, int ey, int* id) {
        id[0] = (ey * m) * nq1 + ex * m;
        id[1] = (ey * m) * nq1 + ex * m + m;
        id[2] = (ey * m + m) * nq1 + ex * m + m;
        id[3] = (ey * m + m) * nq1 + ex * m;
    };
    for (int ey = 0; ey < nelx; ey++)
        for (int ex = 0; ex < nelx; ex++) {
            int id[4];
            corner_ids(ex, ey, id);
            const double* c[4] = {&xl[(size_t)id[0] * 3], &xl[(size_t)id[1] * 3], &xl[(size_t)id[2] * 3], &xl[(size_t)id[3] * 3]};
            for (int qy = 0; qy <= m; qy++)
                for (int qxi = 0; qxi <= m; qxi++) {
                    if ((qxi == 0 || qxi == m) && (qy == 0 || qy == m)) continue;
                    const double x1 = qx[qxi], x2 = qx[qy];
                    const double wgt[4] = {(1.0 - x1) * (1.0 - x2), (1.0 + x1) * (1.0 - x2), (1.0 + x1) * (1.0 + x2),
                                           (1.0 - x1) * (1.0 + x2)};
                    double r[3];
                    for (int a = 0; a < 3; a++)
                        r[a] = 0.25 * (wgt[0] * c[0][a] + wgt[1] * c[1][a] + wgt[2] * c[2][a] + wgt[3] * c[3][a]);
                    const double mag = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
                    const int k = (ey * m + qy) * nq1 + ex * m + qxi;
                    for (int a = 0; a < 3; a++) xl[(size_t)k * 3 + a] = radius * r[a] / mag;
                    lon[k] = std::atan2(xl[(size_t)k * 3 + 1], xl[(size_t)k * 3 + 0]);
                    lat[k] = std::asin(xl[(size_t)k * 3 + 2] / radius);
                }
        }
    for (int ey = 0; ey < nelx; ey++)
        for (int ex = 0; ex < nelx; ex++) {
            const size_t e = (size_t)ey * nelx + ex;
            int id[4];
            corner_ids(ex, ey, id);
            const double* c[4] = {&xl[(size_t)id[0] * 3], &xl[(size_t)id[1] * 3], &xl[(size_t)id[2] * 3], &xl[(size_t)id[3] * 3]};
            for (int qy = 0; qy <= m; qy++)
                for (int qxi = 0; qxi <= m; qxi++) {
                    const int k = (ey * m + qy) * nq1 + ex * m + qxi;
                    double* Jq = &J[(e * nqe + qy * mp1 + qxi) * 4];
                    sphere_jacobian(c, lon[k], lat[k], qx[qxi], qx[qy], radius, Jq);
                    const double d = Jq[0] * Jq[3] - Jq[1] * Jq[2];
                    det[e * nqe + qy * mp1 + qxi] = signed_det ? d : std::fabs(d);
                }
        }
    if (lonlat_out) {
        lonlat_out->resize((size_t)nq1 * nq1 * 2);
        for (int i = 0; i < nq1 * nq1; i++) {
            (*lonlat_out)[(size_t)i * 2] = lon[i];
            (*lonlat_out)[(size_t)i * 2 + 1] = lat[i];
        }
    }
}

bool build_global_mesh(MeshKind kind, int p, int m, int ne, bool signed_det, GlobalMesh& g, std::string* err) {
    const int nfaces = (kind == MESH_SPHERE) ? 6 : 1;
    const int nprocs = nfaces;   // canonical numbering: one patch per face
    g.kind = kind;
    g.p = p;
    g.m = m;
    g.ne = ne;
    g.nfaces = nfaces;
    g.signed_det = signed_det;
    g.nel = (int64_t)nfaces * ne * ne;
    const int np1 = p + 1, mp1 = m + 1;
    const int n0e = np1 * np1, n1e = p * np1, n2e = p * p, nqe = mp1 * mp1;
    g.el0.resize((size_t)g.nel * n0e);
    g.el1x.resize((size_t)g.nel * n1e);
    g.el1y.resize((size_t)g.nel * n1e);
    g.el2.resize((size_t)g.nel * n2e);
    g.elq.resize((size_t)g.nel * nqe);
    g.J.resize((size_t)g.nel * nqe * 4);
    g.det.resize((size_t)g.nel * nqe);

    std::vector<double> qx, qw;
    if (!gll_rule(m, qx, qw) || p < 1 || p > 7) {
        if (err) *err = "unsupported element / quadrature order";
        return false;
    }
    if (kind == MESH_SPHERE) sphere_node_coords(m, ne, g.radius, g.xyz);
    else box_node_coords(m, ne, g.lx, g.xyz);

    for (int f = 0; f < nfaces; f++) {
        PatchTopo t, q;
        if (!patch_topology(kind, p, ne, nprocs, f, t, err)) return false;
        if (!patch_topology(kind, m, ne, nprocs, f, q, err)) return false;
        if (f == 0) {
            g.N0 = t.N0;
            g.N1 = t.N1;
            g.N2 = t.N2;
            g.NQ = q.N0;
        }
        const int nx = t.nx, nxq = q.nx;
        // element tables: the reference's Topo::elInds*_g (eul/Topo.cpp:253-305) for every element
        for (int ey = 0; ey < ne; ey++)
            for (int ex = 0; ex < ne; ex++) {
                const size_t e = (size_t)f * ne * ne + (size_t)ey * ne + ex;
                for (int iy = 0; iy <= p; iy++)
                    for (int ix = 0; ix <= p; ix++)
                        g.el0[e * n0e + iy * np1 + ix] = t.loc0[(ey * p + iy) * (nx + 1) + ex * p + ix];
                for (int iy = 0; iy < p; iy++)
                    for (int ix = 0; ix <= p; ix++)
                        g.el1x[e * n1e + iy * np1 + ix] = t.loc1x[(ey * p + iy) * (nx + 1) + ex * p + ix];
                for (int iy = 0; iy <= p; iy++)
                    for (int ix = 0; ix < p; ix++)
                        g.el1y[e * n1e + iy * p + ix] = t.loc1y[(ey * p + iy) * nx + ex * p + ix];
                for (int iy = 0; iy < p; iy++)
                    for (int ix = 0; ix < p; ix++)
                        g.el2[e * n2e + iy * p + ix] = t.loc2[(ey * p + iy) * nx + ex * p + ix];
                for (int iy = 0; iy <= m; iy++)
                    for (int ix = 0; ix <= m; ix++)
                        g.elq[e * nqe + iy * mp1 + ix] = q.loc0[(ey * m + iy) * (nxq + 1) + ex * m + ix];
            }

        // per-patch geometry exactly as one reference rank computes it (canonical numbering: patch == face)
        std::vector<double> xl((size_t)(nxq + 1) * (nxq + 1) * 3), Jp, dp;
        for (int i = 0; i < (nxq + 1) * (nxq + 1); i++)
            for (int a = 0; a < 3; a++) xl[(size_t)i * 3 + a] = g.xyz[(size_t)q.loc0[i] * 3 + a];
        patch_geometry(kind, m, ne, ne, g.radius, g.lx, signed_det, xl, Jp, dp, NULL);
        const size_t e0 = (size_t)f * ne * ne;
        std::copy(Jp.begin(), Jp.end(), g.J.begin() + e0 * nqe * 4);
        std::copy(dp.begin(), dp.end(), g.det.begin() + e0 * nqe);
    }
    return true;
}

bool write_input_files(MeshKind kind, int p, int m, int ne, int nprocs, const std::string& dir, std::string* err) {
    Numbering nb;
    if (!make_numbering(kind, p, ne, nprocs, nb, err)) return false;
    std::vector<double> xyz;
    if (kind == MESH_SPHERE) sphere_node_coords(m, ne, 6371220.0, xyz);
    else box_node_coords(m, ne, 1000.0, xyz);
    for (int r = 0; r < nprocs; r++) {
        PatchTopo t, q;
        if (!patch_topology(kind, p, ne, nprocs, r, t, err)) return false;
        if (!patch_topology(kind, m, ne, nprocs, r, q, err)) return false;
        int sizes[4] = {t.n0l, t.n1xl, t.n1yl, t.n2l};
        bool ok = write_ints(rank_file(dir, "nodes", r), t.loc0.data(), t.loc0.size()) &&
                  write_ints(rank_file(dir, "edges_x", r), t.loc1x.data(), t.loc1x.size()) &&
                  write_ints(rank_file(dir, "edges_y", r), t.loc1y.data(), t.loc1y.size()) &&
                  write_ints(rank_file(dir, "faces", r), t.loc2.data(), t.loc2.size()) &&
                  write_ints(rank_file(dir, "local_sizes", r), sizes, 4);
        if (kind == MESH_SPHERE) {
            ok = ok && write_ints(rank_file(dir, "quads", r), q.loc0.data(), q.loc0.size()) &&
                 write_ints(rank_file(dir, "local_sizes_quad", r), &q.n0l, 1);
        }
        FILE* f = ok ? std::fopen(rank_file(dir, "geom", r).c_str(), "w") : NULL;
        if (!f) {
            if (err) *err = "cannot write input files under " + dir;
            return false;
        }
        for (size_t i = 0; i < q.loc0.size(); i++) {
            const double* c = &xyz[(size_t)q.loc0[i] * 3];
            std::fprintf(f, "%.18e %.18e %.18e\n", c[0], c[1], c[2]);
        }
        std::fclose(f);
    }
    FILE* f = std::fopen((dir + "/grid_res.txt").c_str(), "w");
    if (!f) return false;
    std::fprintf(f, "%d\n%d", p, nb.nelx);
    std::fclose(f);
    if (kind == MESH_SPHERE) {
        f = std::fopen((dir + "/grid_res_quad.txt").c_str(), "w");
        if (!f) return false;
        std::fprintf(f, "%d\n%d", m, nb.nelx);
        std::fclose(f);
    }
    return true;
}

}  // namespace mimsem
