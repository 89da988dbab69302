// Host side of the device engine + the mimsem_gpu_* C ABI (include/mimsem_gpu.h).
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mimsem_gpu.h"
#include "errors.hpp"
#include "kernels.cuh"
#include "launch.hpp"
#include "m2_tile.cuh"

using namespace mimsem;

namespace {

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    cudaError_t upload(const std::vector<T>& h) {
        cudaError_t e = resize(h.size());
        if (e != cudaSuccess || h.empty()) return e;
        return cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    }
    cudaError_t resize(size_t m) {
        if (m <= n && p) return cudaSuccess;
        release();
        n = m;
        return cudaMalloc((void**)&p, std::max<size_t>(m, 1) * sizeof(T));
    }
};

struct HostCsr {
    int64_t nrows = 0, ncols = 0;
    std::vector<int64_t> indptr;
    std::vector<int> indices;
    std::vector<double> values;
};

struct DevEll {
    int64_t nrows = 0;
    int width = 0;
    DevBuf<int> col;
    DevBuf<signed char> sgn;
    DevBuf<int> rows;
    int64_t nrows_active = 0;   // rows to compute (owned); == nrows when rows list is empty
    bool use_rows = false;
};

}  // namespace

struct mimsem_gpu_ctx {
    int device = 0;
    int p = 0, m = 0;
    bool have_basis = false, have_topo = false, have_geom = false;
    std::vector<double> w, ejxi;   // w[m+1], ejxi[(m+1)*p]

    int nel_total = 0, nel_owned = 0, n0 = 0, n1 = 0, n2 = 0, nq = 0, mode = 0;
    std::vector<int> h_el0, h_el1x, h_el1y, h_el2, h_elq;   // 1-form tables hold INTERNAL indices (perm1 applied)
    std::vector<int> h_perm1;                                // external (caller) 1-form index -> internal
    DevBuf<int> d_perm1;
    std::vector<int> h_nbr;
    std::vector<unsigned char> h_eflags;
    std::vector<int> h_adj_ptr, h_adj_eq, h_node_q;

    DevBuf<int> d_el0, d_el1x, d_el1y, d_el2, d_elq, d_nbr, d_adj_ptr, d_adj_eq, d_node_q;
    DevBuf<int> d_el1xT, d_elqT, d_far;      // line-task tables
    DevBuf<double> d_Gc, d_Gr, d_Gch, d_Grh;
    int n_far = 0;
    // element subsets (multi-GPU overlap): interior = owned elements that read no ghost row
    DevBuf<int> d_elist_int, d_elist_bnd;
    int n_int = 0, n_bnd = 0;
    std::vector<int> h_el1x_ext, h_el1y_ext;   // caller numbering (before perm1), kept for set_ghosts
    // tuning / test knobs: read from the environment ONCE at mimsem_gpu_create, or set with mimsem_gpu_set_option
    int m1_variant = 4;                      // 4: automatic (default: the persistent ring kernel for plain M1 on >= 4000 elements when
                                             //    launches may overlap (pdl_independent), else the TMA tile kernel), 3: ring kernel,
                                             //    2: tile kernel, 1: line tasks, 0: one thread per element-level
    int k_variant = 1;                       // 1: TMA tile kernel (default), 0: one thread per element-level
    int ell_vec = 4;                         // widest level group of the incidence kernels (4, 2, 1)
    int prefetch_ahead = 444;                // L2 prefetch distance of the tile kernels in tiles (0: off)
    int m1_min_blocks = 0;                   // register-budget variant of the M1 tile kernel (0: default)
    int halo_burst_pos = 0, halo_burst_len = 0;   // the next fused M1 launches are launch pos of a burst of len (see HaloFused)
    int pdl = 0;                             // 1: the caller guarantees that consecutive launches on a stream are independent
                                             //    (tile kernels are launched with programmatic stream serialization)
    int host_chunk = 8;                      // levels per pipeline stage of mimsem_gpu_apply_host (measured: scripts/tune_e2e.py)
    int halo_max_levels = 0;                 // levels per ghost row the caller's halo inboxes were allocated for (0: unknown)
    int n0_owned = -1;                       // subdomains: 0-form operators compute rows [0, n0_owned) only (-1: all rows)
    std::vector<int> h_elem_key;             // canonical (partition-independent) order of the elements, e.g. global ids
#ifdef MIMSEM_DIAG
    int diag_debug = 0;
    long long* diag_times = nullptr;
#endif
    // TMA tile plan (owner-computes mode)
    bool tma_ok = false, tma_h_ok = false;
    DevBuf<TileHdr> d_recs, d_recs_h;   // fixed-stride tile records (see TileHdr)
    int rec_stride = 0, rec_stride_h = 0, rec_list = 0, rec_list_h = 0;
    // fused ghost refresh (set_ghosts): records whose ghost rows are staged from the halo inbox; tiles ordered
    // interior first, boundary last
    DevBuf<TileHdr> d_recs_halo;
    int rec_stride_halo = 0, rec_list_halo = 0;
    bool halo_plan_ok = false;
    DevBuf<int> d_elist_all;
    bool elist_all_identity = false;   // the caller already stores interior elements first (no indirection needed)
    DevBuf<unsigned> d_fused_counters;
    DevBuf<TileHdr> d_recs_k;           // K (WtQUmat) tile records
    int rec_stride_k = 0;
    bool k_plan_ok = false;
    DevBuf<TileHdr> d_recs_m2, d_recs_m2h;   // M2 / M2(rho) tile records
    int rec_stride_m2 = 0, rec_stride_m2h = 0;
    bool m2_plan_ok = false;
    int m2_variant = 1;                      // 1: TMA tile kernel (default), 0: one thread per element-level
    DevBuf<int> d_inc_recs;                  // per owned element: e1, f0, xe[P], yn[P], wf[P], sf[P] (k_inc_tile)
    bool inc_plan_ok = false;
    int inc_variant = 1;                     // 1: element kernel for E21 / E12 (default), 0: ELL stencils
    DevBuf<double> d_geo, d_geo_h, d_geo_k, d_geo_m2, d_geo_m2h;
    DevBuf<unsigned char> d_eflags;
    DevBuf<double> d_G1, d_G1h, d_W2, d_W2h, d_D0, d_wq, d_tinv;
    DevBuf<double> d_tmean;   // [nq][nkT]: mean thickness of levels k and k+1 (Ut_mat::assemble), selected by MIMSEM_THICK_MEAN
    DevBuf<double> cg_r, cg_p, cg_q, cg_dinv, cg_partial, cg_scal;   // work space of mimsem_gpu_solve_M1
    DevBuf<double> d_J4, d_det, d_Wr;   // raw Jacobians for the upwinded operators, signed quadrature weight of R(q)
    std::vector<double> xn;             // GLL nodes of order p
    int nkT = 0;

    HostCsr csr[4];
    DevEll ell[4];

    DevBuf<unsigned> d_halo_counters;   // [2][64] finished-CTA counters of the p2p halo kernels (push, pull)
    // staging for the host-buffer entry point
    static constexpr int HOST_SLOTS = 4;     // chunks in flight in mimsem_gpu_apply_host (one stream and one buffer set each)
    DevBuf<double> s_lev, s_x, s_y, s_c, s_lev2[HOST_SLOTS], s_out2[HOST_SLOTS], s_x2[HOST_SLOTS], s_y2[HOST_SLOTS], s_c2[HOST_SLOTS],
        s_u2[HOST_SLOTS];
    DevBuf<double> s_ray;                    // level-0 Exner values of Umat_ray (apply_host)
    cudaStream_t stream = nullptr, host_stream[HOST_SLOTS - 1] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_host[3] = {nullptr, nullptr, nullptr};

    int64_t launches = 0;
};

namespace {

#define CUDA_OK(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                              \
            return MIMSEM_ERR_CUDA;                                                                     \
        }                                                                                               \
    } while (0)

int fail(int code, const std::string& msg) {
    set_error(msg);
    return code;
}

int bind_device(const mimsem_gpu_ctx* ctx) {
    CUDA_OK(cudaSetDevice(ctx->device));
    return MIMSEM_OK;
}

// ------------------------------------------------------------------------------------------
// topology preprocessing (host)

struct EdgeUse {
    int elem;
    int type;   // 0 x-normal, 1 y-normal
    int j;      // element-local index
};

// For each owned element find the element across its west / south side and which far side
// (east column of x-edges or north row of y-edges, possibly reversed -- cubed-sphere seams,
// scr/Proc2.py:165-172, 210-227) of that element the shared edges are.
int build_neighbours(mimsem_gpu_ctx* c) {
    const int P = c->p, NP1 = P + 1, N1E = P * NP1;
    std::vector<std::array<EdgeUse, 2>> use(c->n1);
    std::vector<unsigned char> cnt(c->n1, 0);
    auto add = [&](int dof, EdgeUse u) -> bool {
        if (dof < 0 || dof >= c->n1) return false;
        if (cnt[dof] >= 2) return false;
        use[dof][cnt[dof]++] = u;
        return true;
    };
    for (int e = 0; e < c->nel_total; e++)
        for (int j = 0; j < N1E; j++) {
            if (!add(c->h_el1x[(size_t)e * N1E + j], EdgeUse{e, 0, j}) ||
                !add(c->h_el1y[(size_t)e * N1E + j], EdgeUse{e, 1, j}))
                return fail(MIMSEM_ERR_ARG, "set_topo: an edge index is out of range or shared by more than two elements");
        }
    c->h_nbr.assign((size_t)c->nel_owned * 2, -1);
    c->h_eflags.assign(c->nel_owned, 0);
    for (int e = 0; e < c->nel_owned; e++) {
        for (int s = 0; s < 2; s++) {   // 0: west side (x-edges ix=0), 1: south side (y-edges iy=0)
            int nb = -1, side = -1, rev = -1;
            bool none = false;
            for (int i = 0; i < P; i++) {
                const int type = s, j = (s == 0) ? i * NP1 : i;
                const int dof = (s == 0) ? c->h_el1x[(size_t)e * N1E + j] : c->h_el1y[(size_t)e * N1E + j];
                const EdgeUse* other = nullptr;
                for (int u = 0; u < cnt[dof]; u++) {
                    const EdgeUse& q = use[dof][u];
                    if (!(q.elem == e && q.type == type && q.j == j)) other = &q;
                }
                if (!other) {
                    none = true;
                    continue;
                }
                int oside, oi;
                if (other->type == 0) {
                    if (other->j % NP1 != P) return fail(MIMSEM_ERR_ARG, "set_topo: west/south edge shared with a non-east side");
                    oside = 0;
                    oi = other->j / NP1;
                } else {
                    if (other->j / P != P) return fail(MIMSEM_ERR_ARG, "set_topo: west/south edge shared with a non-north side");
                    oside = 1;
                    oi = other->j % P;
                }
                int r = (oi == i) ? 0 : (oi == P - 1 - i ? 1 : -2);
                if (P > 1 && oi == i && oi == P - 1 - i) r = -1;   // middle edge: orientation undetermined
                if (r == -2) return fail(MIMSEM_ERR_ARG, "set_topo: inconsistent edge order across an element side");
                if (nb == -1) {
                    nb = other->elem;
                    side = oside;
                } else if (nb != other->elem || side != oside) {
                    return fail(MIMSEM_ERR_ARG, "set_topo: an element side touches more than one neighbour side");
                }
                if (r >= 0) {
                    if (rev >= 0 && rev != r) return fail(MIMSEM_ERR_ARG, "set_topo: inconsistent orientation across an element side");
                    rev = r;
                }
            }
            if (nb >= 0 && none) return fail(MIMSEM_ERR_ARG, "set_topo: element side only partially shared");
            if (nb >= 0) {
                if (nb >= (1 << 29)) return fail(MIMSEM_ERR_UNSUPPORTED, "set_topo: too many elements");
                c->h_nbr[(size_t)e * 2 + s] = nb | (side << 29) | ((rev > 0 ? 1 : 0) << 30);
            }
        }
        if (c->mode == 1) {
            // partial-sum mode: east / north edges nobody else in this subdomain computes
            const int de = c->h_el1x[(size_t)e * N1E + P];
            const int dn = c->h_el1y[(size_t)e * N1E + P * P];
            if (cnt[de] == 1) c->h_eflags[e] |= 1;
            if (cnt[dn] == 1) c->h_eflags[e] |= 2;
        }
    }
    return MIMSEM_OK;
}

void csr_from_triplets(int64_t nrows, int64_t ncols, std::vector<std::array<int64_t, 3>>& t, HostCsr& out) {
    // t = (row, col, sign); INSERT semantics: duplicates (same row, col) carry the same value
    std::sort(t.begin(), t.end());
    t.erase(std::unique(t.begin(), t.end(), [](const std::array<int64_t, 3>& a, const std::array<int64_t, 3>& b) {
                return a[0] == b[0] && a[1] == b[1];
            }),
            t.end());
    out.nrows = nrows;
    out.ncols = ncols;
    out.indptr.assign(nrows + 1, 0);
    out.indices.resize(t.size());
    out.values.resize(t.size());
    for (size_t i = 0; i < t.size(); i++) {
        out.indptr[t[i][0] + 1]++;
        out.indices[i] = (int)t[i][1];
        out.values[i] = (double)t[i][2];
    }
    for (int64_t r = 0; r < nrows; r++) out.indptr[r + 1] += out.indptr[r];
}

// Device stencil from (row, col, sign, key) entries.  The order in which a row's entries are ADDED is the
// ascending `key`, chosen so that it does not depend on the local numbering (and hence not on the partition):
// results are bitwise identical on 1 and on N GPUs.
struct StencilEnt {
    int64_t row, col;
    int sign;
    long long key;
};
int upload_ell(int64_t nrows, std::vector<StencilEnt> ents, DevEll& d, const std::vector<int>* rows,
               const std::vector<int>* rowperm, const std::vector<int>* colperm) {
    std::stable_sort(ents.begin(), ents.end(), [](const StencilEnt& a, const StencilEnt& b) {
        return a.row != b.row ? a.row < b.row : a.key < b.key;
    });
    std::vector<int> cnt(nrows, 0);
    int width = 1;
    for (auto& e : ents) width = std::max(width, ++cnt[e.row]);
    std::vector<int> col((size_t)nrows * width, -1);
    std::vector<signed char> sgn((size_t)nrows * width, 0);
    std::fill(cnt.begin(), cnt.end(), 0);
    for (auto& e : ents) {
        const size_t ri = rowperm ? (size_t)(*rowperm)[e.row] : (size_t)e.row;
        const int j = cnt[e.row]++;
        col[ri * width + j] = colperm ? (*colperm)[e.col] : (int)e.col;
        sgn[ri * width + j] = e.sign > 0 ? 1 : -1;
    }
    d.nrows = nrows;
    d.width = width;
    CUDA_OK(d.col.upload(col));
    CUDA_OK(d.sgn.upload(sgn));
    if (rows) {
        d.use_rows = true;
        d.nrows_active = (int64_t)rows->size();
        std::vector<int> rr(*rows);
        if (rowperm)
            for (auto& v : rr) v = (*rowperm)[v];
        std::sort(rr.begin(), rr.end());
        CUDA_OK(d.rows.upload(rr));
    } else {
        d.use_rows = false;
        d.nrows_active = nrows;
    }
    return MIMSEM_OK;
}

// Incidence stencils, restating E10mat / E21mat (eul/Assembly.cpp:1102-1162, 1170-1220) over the
// local element tables; E01 = -E10^T and E12 = -E21^T (ibid. :1156-1161, :1214-1219).
int build_incidence(mimsem_gpu_ctx* c) {
    const int P = c->p, NP1 = P + 1, N1E = P * NP1, N2E = P * P, N0E = NP1 * NP1;
    std::vector<std::array<int64_t, 3>> t10, t21, t01, t12;
    std::vector<int> rows10, rows21;
    for (int e = 0; e < c->nel_total; e++) {
        const int* ix = &c->h_el1x[(size_t)e * N1E];
        const int* iy = &c->h_el1y[(size_t)e * N1E];
        const int* i0 = &c->h_el0[(size_t)e * N0E];
        const int* i2 = &c->h_el2[(size_t)e * N2E];
        for (int b = 0; b < P; b++)
            for (int a = 0; a < P; a++) {
                // edges local to this element (west/south/interior); a = ix, b = iy
                const int ll = b * NP1 + a;
                const int rx = ix[b * NP1 + a], ry = iy[b * P + a];
                t10.push_back({rx, i0[ll], +1});
                t10.push_back({rx, i0[ll + NP1], -1});
                t10.push_back({ry, i0[ll], -1});
                t10.push_back({ry, i0[ll + 1], +1});
                // face (a, b)
                const int rf = i2[b * P + a];
                t21.push_back({rf, ix[b * NP1 + a], -1});
                t21.push_back({rf, ix[b * NP1 + a + 1], +1});
                t21.push_back({rf, iy[b * P + a], -1});
                t21.push_back({rf, iy[(b + 1) * P + a], +1});
                if (e < c->nel_owned) {
                    rows10.push_back(rx);
                    rows10.push_back(ry);
                    rows21.push_back(rf);
                }
            }
    }
    for (auto& v : t10) t01.push_back({v[1], v[0], -v[2]});
    for (auto& v : t21) t12.push_back({v[1], v[0], -v[2]});
    // device stencils: E10 / E21 add their entries in the reference's MatSetValues order (eul/Assembly.cpp:1133-1148,
    // 1196-1205); E12 rows add the +1 face (the edge's owner side) before the -1 face; E01 rows by column.
    std::vector<StencilEnt> s10, s21, s01, s12;
    for (size_t i = 0; i < t10.size(); i++) s10.push_back({t10[i][0], t10[i][1], (int)t10[i][2], (int)(i & 1)});
    for (size_t i = 0; i < t21.size(); i++) s21.push_back({t21[i][0], t21[i][1], (int)t21[i][2], (int)(i & 3)});
    for (auto& v : t12) s12.push_back({v[0], v[1], (int)v[2], v[2] > 0 ? 0 : 1});
    // E01 rows add their (up to four) edges in the order (element that lists the edge, position in that element), with the
    // elements in the caller's canonical order if one was given: independent of the local numbering
    for (size_t i = 0; i < t01.size(); i++) {
        const int e = (int)(i / (4 * (size_t)P * P)), slot = (int)(i % (4 * (size_t)P * P));
        const long long ek = c->h_elem_key.empty() ? e : c->h_elem_key[e];
        s01.push_back({t01[i][0], t01[i][1], (int)t01[i][2], ek * (4 * P * P) + slot});
    }
    csr_from_triplets(c->n1, c->n0, t10, c->csr[MIMSEM_E10]);
    csr_from_triplets(c->n0, c->n1, t01, c->csr[MIMSEM_E01]);
    csr_from_triplets(c->n2, c->n1, t21, c->csr[MIMSEM_E21]);
    csr_from_triplets(c->n1, c->n2, t12, c->csr[MIMSEM_E12]);
    std::sort(rows10.begin(), rows10.end());
    std::sort(rows21.begin(), rows21.end());
    const bool all = (c->nel_owned == c->nel_total);
    int rc;
    const std::vector<int>* pm = &c->h_perm1;
    if ((rc = upload_ell(c->n1, s10, c->ell[MIMSEM_E10], all ? nullptr : &rows10, pm, nullptr))) return rc;
    if ((rc = upload_ell(c->n2, s21, c->ell[MIMSEM_E21], all ? nullptr : &rows21, nullptr, pm))) return rc;
    if ((rc = upload_ell(c->n0, s01, c->ell[MIMSEM_E01], nullptr, nullptr, pm))) return rc;
    if ((rc = upload_ell(c->n1, s12, c->ell[MIMSEM_E12], nullptr, pm, nullptr))) return rc;
    return MIMSEM_OK;
}

// Records of the element kernel for E21 / E12 (k_inc_tile): needs the engine's contiguous edge and face blocks
int build_inc_plan(mimsem_gpu_ctx* c) {
    const int P = c->p, NP1 = P + 1, N1E = P * NP1, N2E = P * P;
    c->inc_plan_ok = false;
    // the face on the far side of every edge that is some element's EAST or NORTH edge (the -1 entry of its E12 row)
    std::vector<int> far_face(c->n1, -1);
    for (int e = 0; e < c->nel_total; e++) {
        const int* ix = &c->h_el1x[(size_t)e * N1E];
        const int* iy = &c->h_el1y[(size_t)e * N1E];
        const int* i2 = &c->h_el2[(size_t)e * N2E];
        for (int b = 0; b < P; b++) {
            far_face[ix[b * NP1 + P]] = i2[b * P + P - 1];          // east column  <- face (P-1, b)
            far_face[iy[P * P + b]] = i2[(P - 1) * P + b];          // north row    <- face (b, P-1)
        }
    }
    const int W = 2 + 4 * P;
    std::vector<int> rec((size_t)c->nel_owned * W);
    for (int e = 0; e < c->nel_owned; e++) {
        const int* ix = &c->h_el1x[(size_t)e * N1E];
        const int* iy = &c->h_el1y[(size_t)e * N1E];
        const int* i2 = &c->h_el2[(size_t)e * N2E];
        int* r = &rec[(size_t)e * W];
        r[0] = ix[0];
        r[1] = i2[0];
        for (int j = 0; j < N2E; j++)
            if (i2[j] != i2[0] + j) return MIMSEM_OK;                                   // faces not contiguous: keep the ELL kernels
        for (int a = 0; a < P; a++)
            for (int b = 0; b < P; b++)
                if (ix[b * NP1 + a] != r[0] + a * P + b || iy[b * P + a] != r[0] + P * P + b * P + a) return MIMSEM_OK;
        for (int b = 0; b < P; b++) {
            r[2 + b] = ix[b * NP1 + P];
            r[2 + P + b] = iy[P * P + b];
            r[2 + 2 * P + b] = far_face[ix[b * NP1]];     // across my west edge xx(0, b)
            r[2 + 3 * P + b] = far_face[iy[b]];           // across my south edge xy(b, 0)
        }
    }
    CUDA_OK(c->d_inc_recs.upload(rec));
    c->inc_plan_ok = true;
    return MIMSEM_OK;
}

int build_node_adjacency(mimsem_gpu_ctx* c) {
    const int NP1 = c->p + 1, Q2 = NP1 * NP1;
    c->h_adj_ptr.assign(c->n0 + 1, 0);
    c->h_node_q.assign(c->n0, 0);
    for (int e = 0; e < c->nel_total; e++)
        for (int q = 0; q < Q2; q++) {
            const int n = c->h_el0[(size_t)e * Q2 + q];
            if (n < 0 || n >= c->n0) return fail(MIMSEM_ERR_ARG, "set_topo: node index out of range");
            c->h_adj_ptr[n + 1]++;
            c->h_node_q[n] = c->h_elq[(size_t)e * Q2 + q];
        }
    for (int n = 0; n < c->n0; n++) c->h_adj_ptr[n + 1] += c->h_adj_ptr[n];
    c->h_adj_eq.assign(c->h_adj_ptr[c->n0], 0);
    std::vector<int> pos(c->h_adj_ptr.begin(), c->h_adj_ptr.end() - 1);
    for (int e = 0; e < c->nel_total; e++)
        for (int q = 0; q < Q2; q++) c->h_adj_eq[pos[c->h_el0[(size_t)e * Q2 + q]]++] = e * Q2 + q;
    // canonical order of every node's (element, point) pairs: by the caller's element keys (global element ids on a
    // partitioned mesh), so that sums over the pairs do not depend on the local element order -- N-GPU results of the
    // 0-form operators are then bitwise equal to the single-GPU ones
    if (!c->h_elem_key.empty()) {
        if ((int)c->h_elem_key.size() != c->nel_total) return fail(MIMSEM_ERR_ARG, "set_element_keys: one key per element of set_topo");
        const std::vector<int>& key = c->h_elem_key;
        for (int n = 0; n < c->n0; n++)
            std::sort(c->h_adj_eq.begin() + c->h_adj_ptr[n], c->h_adj_eq.begin() + c->h_adj_ptr[n + 1], [&](int a, int b) {
                const int ea = a / Q2, eb = b / Q2;
                return key[ea] != key[eb] ? key[ea] < key[eb] : a < b;
            });
    }
    return MIMSEM_OK;
}


// ------------------------------------------------------------------------------------------
// TMA tile plan: per owned element, the list of bulk copies that fills the tile's slots
// (see M1Slots in engine.cuh) and the list of bulk stores of its owned edges.
// Tile records of the M1 / M1(h) kernel.  ghost_from >= 0 builds only the fused-halo variant of the plain-M1 records:
// x rows whose CALLER index is >= ghost_from are read from inbox row (caller index - ghost_from) instead (copy kind 4
// for staged slots, negative entries of the explicit far-row list for the far lines).
template <int P>
int build_tma_plan_p(mimsem_gpu_ctx* c, int ghost_from = -1) {
    using S = M1Slots<P>;
    std::vector<int> inv1;
    if (ghost_from >= 0) {
        inv1.assign(c->n1, 0);
        for (int i = 0; i < c->n1; i++) inv1[c->h_perm1[i]] = i;
    }
    constexpr int NP1 = P + 1, N1E = P * NP1, N2E = P * P, Q2 = NP1 * NP1, NF = NP1 * P;
    struct Rec {
        TileHdr h;
        TileFar f;
        TileFarH fh;
        std::vector<CopyEnt> cp, cp_h;
        int nslots_h;
        int list[2][NF];
    };
    std::vector<Rec> recs(c->nel_owned);
    auto emit_runs = [](std::vector<std::pair<int, int>>& dof_slot, int kind, std::vector<CopyEnt>& out) {
        // merge (dof, slot) pairs that advance together into runs
        size_t i = 0;
        int filled = 0;
        while (i < dof_slot.size()) {
            size_t j = i + 1;
            while (j < dof_slot.size() && dof_slot[j].first == dof_slot[j - 1].first + 1 && dof_slot[j].second == dof_slot[j - 1].second + 1) j++;
            out.push_back(CopyEnt{kind, dof_slot[i].first, dof_slot[i].second, (int)(j - i)});
            filled += (int)(j - i);
            i = j;
        }
        return filled;
    };
    bool all_contig = true, h_contig = true, any_list = false;
    for (int e = 0; e < c->nel_owned; e++) {
        Rec& r = recs[e];
        const int* ex = &c->h_el1x[(size_t)e * N1E];
        const int* ey = &c->h_el1y[(size_t)e * N1E];
        std::vector<std::pair<int, int>> xs, ts, hs;
        // own block, in slot order
        for (int ix = 0; ix < P; ix++)
            for (int iy = 0; iy < P; iy++) xs.push_back({ex[iy * NP1 + ix], S::OX + ix * P + iy});
        for (int iy = 0; iy < P; iy++)
            for (int ix = 0; ix < P; ix++) xs.push_back({ey[iy * P + ix], S::OY + iy * P + ix});
        // the kernel stores the owned block as one run of rows starting at st_dof
        for (size_t i = 1; i < xs.size(); i++)
            if (xs[i].first != xs[0].first + (int)i) all_contig = false;
        r.h.st_dof = xs[0].first;
        for (int iy = 0; iy < P; iy++) xs.push_back({ex[iy * NP1 + P], S::XE + iy});
        for (int ix = 0; ix < P; ix++) xs.push_back({ey[P * P + ix], S::YN + ix});
        int flags = 0;
        r.f = TileFar{0, 0, 0, 0};
        r.fh = TileFarH{0, 0, 0, 0};
        for (int s = 0; s < 2; s++) {
            for (int i = 0; i < NF; i++) r.list[s][i] = 0;
            const int nb = c->h_nbr[(size_t)e * 2 + s];
            if (nb < 0) continue;
            const int n = nb & 0x1fffffff, side = (nb >> 29) & 1, rev = (nb >> 30) & 1;
            flags |= (s == 0 ? TF_HAS_W : TF_HAS_S) | (rev ? (s == 0 ? TF_REV_W : TF_REV_S) : 0) | (side ? (s == 0 ? TF_ROW_W : TF_ROW_S) : 0);
            // far line = east column (side 0): other family = y-edges xy(ix=t, qy=q)
            // far line = north row  (side 1): other family = x-edges xx(qx=q, iy=t)
            int rows[NF];
            bool ghost = false;
            for (int q = 0; q <= P; q++)
                for (int t = 0; t < P; t++) {
                    const int dof = side == 0 ? c->h_el1y[(size_t)n * N1E + q * P + t] : c->h_el1x[(size_t)n * N1E + t * NP1 + q];
                    rows[q * P + t] = dof;
                    if (ghost_from >= 0 && inv1[dof] >= ghost_from) ghost = true;
                }
            bool run16 = true, asc4 = true, desc4 = true;
            for (int i = 1; i < P * P; i++)
                if (rows[i] != rows[0] + i) run16 = false;
            for (int t = 1; t < P; t++) {
                if (rows[P * P + t] != rows[P * P] + t) asc4 = false;
                if (rows[P * P + t] != rows[P * P] - t) desc4 = false;
            }
            if (!ghost && run16 && (asc4 || desc4)) {
                (s == 0 ? r.f.w16 : r.f.s16) = rows[0];
                (s == 0 ? r.f.w4 : r.f.s4) = rows[P * P];
                if (!asc4) flags |= (s == 0 ? TF_W4_DESC : TF_S4_DESC);
            } else {
                flags |= (s == 0 ? TF_LIST_W : TF_LIST_S);
                any_list = true;
                for (int i = 0; i < NF; i++) {
                    const int ext = ghost_from >= 0 ? inv1[rows[i]] : -1;
                    r.list[s][i] = (ghost_from >= 0 && ext >= ghost_from) ? -(ext - ghost_from) - 1 : rows[i];
                }
            }
            // neighbour's 2-form block (M1h)
            const int* e2 = &c->h_el2[(size_t)n * N2E];
            for (int j = 1; j < N2E; j++)
                if (e2[j] != e2[0] + j) h_contig = false;
            (s == 0 ? r.fh.hw : r.fh.hs) = e2[0];
        }
        for (int q = 0; q < Q2; q++) ts.push_back({c->h_elq[(size_t)e * Q2 + q], S::T + q});
        for (int j = 0; j < N2E; j++) hs.push_back({c->h_el2[(size_t)e * N2E + j], S::H + j});
        r.h.flags = flags;
        // plain M1
        r.cp.push_back(CopyEnt{3, e, 0, 1});
        int nx, ng = 0;
        if (ghost_from >= 0) {
            std::vector<std::pair<int, int>> xo, xg;
            for (auto& ds : xs) {
                const int ext = inv1[ds.first];
                if (ext >= ghost_from) xg.push_back({ext - ghost_from, ds.second});
                else xo.push_back(ds);
            }
            nx = emit_runs(xo, 0, r.cp);
            r.fh.first4 = (int)r.cp.size();
            ng = emit_runs(xg, 4, r.cp);
            r.fh.n4 = (int)r.cp.size() - r.fh.first4;
        } else {
            nx = emit_runs(xs, 0, r.cp);
        }
        const int nt = emit_runs(ts, 2, r.cp);
        r.h.cp_count = (int)r.cp.size();
        r.h.nslots = nx | (ng << 12) | (nt << 20);
        // M1(h)
        r.cp_h.push_back(CopyEnt{3, e, 0, 1});
        emit_runs(xs, 0, r.cp_h);
        const int nh = emit_runs(hs, 1, r.cp_h);
        emit_runs(ts, 2, r.cp_h);
        r.nslots_h = (nx + nh) | (nt << 20);
    }
    // pack into fixed-stride records: header words, copy entries, explicit far-row lists (only if some tile needs one)
    static_assert(sizeof(TileHdr) == 16 && sizeof(CopyEnt) == 16 && sizeof(TileFar) == 16 && sizeof(TileFarH) == 16, "16-byte records");
    constexpr int LIST_WORDS = (2 * NF * 4 + 15) / 16;
    auto pack = [&](bool with_h, int& nents, int& stride, int& list_off, DevBuf<TileHdr>& out) -> cudaError_t {
        nents = 1;
        for (auto& r : recs) nents = std::max(nents, (int)(with_h ? r.cp_h.size() : r.cp.size()));
        list_off = any_list ? kRecHdr + nents : 0;
        stride = kRecHdr + nents + (any_list ? LIST_WORDS : 0);
        std::vector<TileHdr> rec((size_t)recs.size() * stride);
        std::memset(rec.data(), 0, rec.size() * sizeof(TileHdr));
        for (size_t e = 0; e < recs.size(); e++) {
            const Rec& r = recs[e];
            TileHdr* w = &rec[e * stride];
            w[0] = r.h;
            const std::vector<CopyEnt>& cp = with_h ? r.cp_h : r.cp;
            if (with_h) {
                w[0].cp_count = (int)cp.size();
                w[0].nslots = r.nslots_h;
            }
            std::memcpy(&w[1], &r.f, 16);
            std::memcpy(&w[2], &r.fh, 16);
            std::memcpy(&w[kRecHdr], cp.data(), cp.size() * sizeof(CopyEnt));
            if (any_list) std::memcpy(&w[list_off], &r.list[0][0], 2 * NF * sizeof(int));
        }
        return out.upload(rec);
    };
    int nents = 0;
    if (ghost_from >= 0) {
        CUDA_OK(pack(false, nents, c->rec_stride_halo, c->rec_list_halo, c->d_recs_halo));
        c->halo_plan_ok = all_contig;
        return MIMSEM_OK;
    }
    CUDA_OK(pack(false, nents, c->rec_stride, c->rec_list, c->d_recs));
    CUDA_OK(pack(true, nents, c->rec_stride_h, c->rec_list_h, c->d_recs_h));
    c->tma_ok = all_contig;
    c->tma_h_ok = all_contig && h_contig;
    return MIMSEM_OK;
}

// Tile records of the K kernel (KSlots): x and u1 edges of the element, thickness; output = the element's faces.
template <int P>
int build_k_plan_p(mimsem_gpu_ctx* c) {
    using S = KSlots<P>;
    constexpr int NP1 = P + 1, N1E = P * NP1, N2E = P * P, Q2 = NP1 * NP1;
    std::vector<TileHdr> rec;
    std::vector<std::vector<CopyEnt>> ents(c->nel_owned);
    std::vector<TileHdr> hdr(c->nel_owned);
    bool ok = true;
    int nents = 1;
    for (int e = 0; e < c->nel_owned; e++) {
        const int* ex = &c->h_el1x[(size_t)e * N1E];
        const int* ey = &c->h_el1y[(size_t)e * N1E];
        std::vector<std::pair<int, int>> xs, ts;
        for (int ix = 0; ix < P; ix++)
            for (int iy = 0; iy < P; iy++) xs.push_back({ex[iy * NP1 + ix], S::OX + ix * P + iy});
        for (int iy = 0; iy < P; iy++)
            for (int ix = 0; ix < P; ix++) xs.push_back({ey[iy * P + ix], S::OY + iy * P + ix});
        for (int iy = 0; iy < P; iy++) xs.push_back({ex[iy * NP1 + P], S::XE + iy});
        for (int ix = 0; ix < P; ix++) xs.push_back({ey[P * P + ix], S::YN + ix});
        for (int q = 0; q < Q2; q++) ts.push_back({c->h_elq[(size_t)e * Q2 + q], S::T + q});
        auto runs = [&](const std::vector<std::pair<int, int>>& ds, int kind, int slot_off) {
            size_t i = 0;
            while (i < ds.size()) {
                size_t j = i + 1;
                while (j < ds.size() && ds[j].first == ds[j - 1].first + 1 && ds[j].second == ds[j - 1].second + 1) j++;
                ents[e].push_back(CopyEnt{kind, ds[i].first, ds[i].second + slot_off, (int)(j - i)});
                i = j;
            }
        };
        ents[e].push_back(CopyEnt{3, e, 0, 1});
        runs(xs, 0, 0);
        runs(xs, 1, S::U0);
        runs(ts, 2, 0);
        const int* e2 = &c->h_el2[(size_t)e * N2E];
        for (int j = 0; j < N2E; j++)
            if (e2[j] != e2[0] + j) ok = false;
        TileHdr h;
        h.st_dof = e2[0];
        h.cp_count = (int)ents[e].size();
        h.flags = 0;
        h.nslots = (int)(2 * xs.size()) | ((int)ts.size() << 20);
        hdr[e] = h;
        nents = std::max(nents, h.cp_count);
    }
    rec.assign((size_t)c->nel_owned * (1 + nents), TileHdr{0, 0, 0, 0});
    for (int e = 0; e < c->nel_owned; e++) {
        rec[(size_t)e * (1 + nents)] = hdr[e];
        std::memcpy(&rec[(size_t)e * (1 + nents) + 1], ents[e].data(), ents[e].size() * sizeof(CopyEnt));
    }
    CUDA_OK(c->d_recs_k.upload(rec));
    c->rec_stride_k = 1 + nents;
    c->k_plan_ok = ok;
    return MIMSEM_OK;
}

// Tile records of the M2 / M2(rho) kernel (M2Slots): the element's faces, thickness, optionally the coefficient's faces
template <int P>
int build_m2_plan_p(mimsem_gpu_ctx* c) {
    using S = M2Slots<P>;
    constexpr int NP1 = P + 1, N2E = P * P, Q2 = NP1 * NP1;
    bool ok = true;
    for (int with_h = 0; with_h < 2; with_h++) {
        std::vector<std::vector<CopyEnt>> ents(c->nel_owned);
        std::vector<TileHdr> hdr(c->nel_owned);
        int nents = 1;
        for (int e = 0; e < c->nel_owned; e++) {
            const int* e2 = &c->h_el2[(size_t)e * N2E];
            std::vector<std::pair<int, int>> xs, ts;
            for (int j = 0; j < N2E; j++) xs.push_back({e2[j], S::X + j});
            for (int q = 0; q < Q2; q++) ts.push_back({c->h_elq[(size_t)e * Q2 + q], S::T + q});
            auto runs = [&](const std::vector<std::pair<int, int>>& ds, int kind, int slot_off) {
                size_t i = 0;
                while (i < ds.size()) {
                    size_t j = i + 1;
                    while (j < ds.size() && ds[j].first == ds[j - 1].first + 1 && ds[j].second == ds[j - 1].second + 1) j++;
                    ents[e].push_back(CopyEnt{kind, ds[i].first, ds[i].second + slot_off, (int)(j - i)});
                    i = j;
                }
            };
            ents[e].push_back(CopyEnt{3, e, 0, 1});
            runs(xs, 0, 0);
            if (with_h) runs(xs, 1, S::H - S::X);
            runs(ts, 2, 0);
            for (int j = 0; j < N2E; j++)
                if (e2[j] != e2[0] + j) ok = false;
            TileHdr h;
            h.st_dof = e2[0];
            h.cp_count = (int)ents[e].size();
            h.flags = 0;
            h.nslots = (int)((with_h ? 2 : 1) * xs.size()) | ((int)ts.size() << 20);
            hdr[e] = h;
            nents = std::max(nents, h.cp_count);
        }
        std::vector<TileHdr> rec((size_t)c->nel_owned * (1 + nents), TileHdr{0, 0, 0, 0});
        for (int e = 0; e < c->nel_owned; e++) {
            rec[(size_t)e * (1 + nents)] = hdr[e];
            std::memcpy(&rec[(size_t)e * (1 + nents) + 1], ents[e].data(), ents[e].size() * sizeof(CopyEnt));
        }
        CUDA_OK((with_h ? c->d_recs_m2h : c->d_recs_m2).upload(rec));
        (with_h ? c->rec_stride_m2h : c->rec_stride_m2) = 1 + nents;
    }
    c->m2_plan_ok = ok;
    return MIMSEM_OK;
}

// geometry records of the M2 tile kernel: the point weights w/det (M2) resp. w/det^2 (M2h), padded to an even count
int build_m2_geo(mimsem_gpu_ctx* c, const std::vector<double>& W, DevBuf<double>& out) {
    const int Q2 = (c->p + 1) * (c->p + 1), GM = (Q2 + 1) / 2 * 2;
    std::vector<double> geo((size_t)c->nel_owned * GM, 0.0);
    for (int e = 0; e < c->nel_owned; e++)
        for (int q = 0; q < Q2; q++) geo[(size_t)e * GM + q] = W[(size_t)e * Q2 + q];
    CUDA_OK(out.upload(geo));
    return MIMSEM_OK;
}

// geometry records of the M1 tile kernel: gl[part][line][q] = (g_own, g_oth) of the element's own lines, then the
// (g_own, g_oth) pairs of the west and south far lines (M1Slots)
template <int P>
int build_tma_geo_p(mimsem_gpu_ctx* c, const std::vector<double>& G, DevBuf<double>& out) {
    using S = M1Slots<P>;
    constexpr int NP1 = P + 1, Q2 = NP1 * NP1;
    std::vector<double> geo((size_t)c->nel_owned * S::GEO, 0.0);
    for (int e = 0; e < c->nel_owned; e++) {
        double* g = &geo[(size_t)e * S::GEO];
        const double* Ge = &G[(size_t)e * Q2 * 3];
        for (int ln = 0; ln < P; ln++)
            for (int q = 0; q <= P; q++) {
                // part 0: x-line ln = GLL column qx = ln, points qy = q: f0 = c (Gaa ul0 + Gab ul1)
                const int q0 = q * NP1 + ln;
                g[S::GL + (ln * NP1 + q) * 2 + 0] = Ge[q0 * 3 + 0];
                g[S::GL + (ln * NP1 + q) * 2 + 1] = Ge[q0 * 3 + 1];
                // part 1: y-line ln = GLL row qy = ln, points qx = q: f1 = c (Gbb ul1 + Gab ul0)
                const int q1 = ln * NP1 + q;
                g[S::GL + P * NP1 * 2 + (ln * NP1 + q) * 2 + 0] = Ge[q1 * 3 + 2];
                g[S::GL + P * NP1 * 2 + (ln * NP1 + q) * 2 + 1] = Ge[q1 * 3 + 1];
            }
        for (int s = 0; s < 2; s++) {
            const int nb = c->h_nbr[(size_t)e * 2 + s];
            if (nb < 0) continue;
            const int n = nb & 0x1fffffff, side = (nb >> 29) & 1;
            double* gf = g + (s == 0 ? S::GW : S::GS);
            for (int q = 0; q <= P; q++) {
                const size_t pt = (size_t)n * Q2 + (side == 0 ? q * NP1 + P : P * NP1 + q);
                // east column: f = c (Gaa u_own + Gab u_oth) ; north row: f = c (Gbb u_own + Gab u_oth)
                gf[q * 2 + 0] = side == 0 ? G[pt * 3 + 0] : G[pt * 3 + 2];
                gf[q * 2 + 1] = G[pt * 3 + 1];
            }
        }
    }
    CUDA_OK(out.upload(geo));
    return MIMSEM_OK;
}

// geometry records of the K tile kernel: G[q][3] of the element, padded to an even number of doubles
int build_k_geo(mimsem_gpu_ctx* c, const std::vector<double>& G, DevBuf<double>& out) {
    const int Q2 = (c->p + 1) * (c->p + 1), GK = (Q2 * 3 + 1) / 2 * 2;
    std::vector<double> geo((size_t)c->nel_owned * GK, 0.0);
    for (int e = 0; e < c->nel_owned; e++)
        for (int i = 0; i < Q2 * 3; i++) geo[(size_t)e * GK + i] = G[(size_t)e * Q2 * 3 + i];
    CUDA_OK(out.upload(geo));
    return MIMSEM_OK;
}

template <class Args>
void copy_basis(const mimsem_gpu_ctx* c, Args& a) {
    std::memset(a.E, 0, sizeof(a.E));
    std::memcpy(a.E, c->ejxi.data(), c->ejxi.size() * sizeof(double));
}

int check_ready(const mimsem_gpu_ctx* c, bool need_thick, int lev0, int nlev, int ld, int flags) {
    if (!c) return fail(MIMSEM_ERR_ARG, "null context");
    if ((flags & (MIMSEM_SUBSET_INTERIOR | MIMSEM_SUBSET_BOUNDARY)) && c->n_int + c->n_bnd != c->nel_owned)
        return fail(MIMSEM_ERR_STATE, "element subsets need mimsem_gpu_set_ghosts");
    if ((flags & MIMSEM_SUBSET_INTERIOR) && (flags & MIMSEM_SUBSET_BOUNDARY)) return fail(MIMSEM_ERR_ARG, "choose one subset");
    if (!c->have_basis || !c->have_topo || !c->have_geom) return fail(MIMSEM_ERR_STATE, "set_basis, set_topo and set_geom must precede an apply");
    if (c->m != c->p) return fail(MIMSEM_ERR_UNSUPPORTED, "the sum-factorised kernels require quadrature order == element order");
    if (nlev < 1 || ld < nlev || lev0 < 0) return fail(MIMSEM_ERR_ARG, "bad level range / leading dimension");
    if (need_thick) {
        const int last = (flags & MIMSEM_FIXED_LEVEL) ? lev0 : lev0 + nlev - 1;
        if (c->nkT == 0) return fail(MIMSEM_ERR_STATE, "tpow > 0 needs set_thickness");
        if (last >= c->nkT) return fail(MIMSEM_ERR_ARG, "level range exceeds the thickness table");
    }
    return MIMSEM_OK;
}

void fill_common(const mimsem_gpu_ctx* c, KArgs& a, int lev0, int nlev, int ld, double scale, int tpow, int flags) {
    a.nel = c->nel_owned;
    a.elist = nullptr;
    if (flags & (MIMSEM_SUBSET_INTERIOR | MIMSEM_SUBSET_BOUNDARY)) {
        const bool in = flags & MIMSEM_SUBSET_INTERIOR;
        a.nel = in ? c->n_int : c->n_bnd;
        a.elist = in ? c->d_elist_int.p : c->d_elist_bnd.p;
    }
    a.nlev = nlev;
    a.ld = ld;
    a.lev0 = lev0;
    a.lev_stride = (flags & MIMSEM_FIXED_LEVEL) ? 0 : 1;
    a.nkT = c->nkT;
    a.tpow = tpow;
    a.scale = scale;
    a.el1x = c->d_el1x.p;
    a.el1y = c->d_el1y.p;
    a.el2 = c->d_el2.p;
    a.elq = c->d_elq.p;
    a.nbr = c->d_nbr.p;
    a.eflags = c->d_eflags.p;
    a.tinv = (flags & MIMSEM_THICK_MEAN) ? c->d_tmean.p : c->d_tinv.p;
    a.c = nullptr;
    a.c2 = nullptr;
    a.ray_dt = 0.0;
    a.det = c->d_det.p;
    a.el1xT = c->d_el1xT.p;
    a.elqT = c->d_elqT.p;
    a.Gc = a.Gr = nullptr;
    const FastDiv fd = make_fastdiv((unsigned)nlev);
    a.div_m = fd.m;
    a.div_s = fd.s;
    copy_basis(c, a);
}

template <class F>
int dispatch_p(int p, F f) {
    switch (p) {
        case 2: return f(std::integral_constant<int, 2>());
        case 3: return f(std::integral_constant<int, 3>());
        case 4: return f(std::integral_constant<int, 4>());
        case 5: return f(std::integral_constant<int, 5>());
        default: return fail(MIMSEM_ERR_UNSUPPORTED, "element order must be 2..5");
    }
}

int finish_launch(mimsem_gpu_ctx* c, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(MIMSEM_ERR_CUDA, std::string(what) + " launch: " + cudaGetErrorString(e));
    c->launches++;
    return MIMSEM_OK;
}

unsigned grid_for(int64_t threads, int block) { return (unsigned)((threads + block - 1) / block); }

int apply_m1(mimsem_gpu_ctx* c, bool with_h, int lev0, int nlev, int ld, double scale, int tpow, int flags,
             const double* h2, const double* x, double* y, cudaStream_t st, const HaloFused* hf = nullptr) {
    int rc = check_ready(c, tpow > 0, lev0, nlev, ld, flags);
    if (rc) return rc;
    if ((rc = bind_device(c))) return rc;
    KArgs a;
    fill_common(c, a, lev0, nlev, ld, scale, tpow, flags);
    a.G = with_h ? c->d_G1h.p : c->d_G1.p;
    a.c = h2;
    a.x = x;
    a.y = y;
    const int64_t threads = (int64_t)a.nel * nlev;
    if (threads >= (1ll << 31)) return fail(MIMSEM_ERR_UNSUPPORTED, "more than 2^31 element-levels in one launch");
    if (threads == 0) return MIMSEM_OK;
    // TMA tile kernel: needs 16-byte aligned, even-length level runs and one thread per level
    const bool aligned = ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && (!with_h || (uintptr_t)h2 % 16 == 0);
    const bool thick_ok = tpow == 0 || (!(flags & MIMSEM_FIXED_LEVEL) && c->nkT % 2 == 0 && lev0 % 2 == 0);
    const bool tma_path = c->m1_variant >= 2 && (with_h ? c->tma_h_ok : c->tma_ok) && aligned && thick_ok && nlev % 2 == 0 && ld % 2 == 0 && nlev <= 64;
    if (hf) {
        if (with_h || !tma_path || !c->halo_plan_ok || a.elist || ld != nlev)
            return fail(MIMSEM_ERR_UNSUPPORTED, "fused ghost refresh needs the TMA tile path (plain M1, all owned elements, even nlev == ld <= 64, set_ghosts)");
        if (hf->npush > kMaxPushPeers || hf->npull > 32) return fail(MIMSEM_ERR_ARG, "too many halo peers");
    }
    if (tma_path) {
        TArgs t;
        std::memset(&t, 0, sizeof(t));
        t.nlev = nlev; t.ld = ld; t.lev0 = lev0; t.nkT = c->nkT; t.tpow = tpow;
        t.contig_x = (ld == nlev); t.contig_t = (c->nkT == nlev);
        t.scale = scale;
        t.prefetch_ahead = c->prefetch_ahead;
        t.prefetch_own_slots = 2 * c->p * c->p;
        t.pdl = c->pdl;
        t.elist = a.elist;
        t.recs = with_h ? c->d_recs_h.p : c->d_recs.p;
        t.rec_stride = with_h ? c->rec_stride_h : c->rec_stride;
        t.rec_list = with_h ? c->rec_list_h : c->rec_list;
        t.rec_hdr = kRecHdr;
        t.geo = with_h ? c->d_geo_h.p : c->d_geo.p;
        t.x = x; t.c = h2; t.tinv = a.tinv; t.y = y;
        copy_basis(c, t);
        for (size_t i = 0; i < sizeof(t.E) / sizeof(double); i++) t.Es[i] = scale * t.E[i];
        M1TileLaunch l{c->p, with_h, hf ? (hf->ll ? 2 : 1) : 0, a.nel, 0, false, c->m1_min_blocks};
        if (hf) {
            constexpr int NCTR = (kMaxBurst + 1) * kCtrStride;   // counter slots of a burst, then the sequence words
            if (!c->d_fused_counters.p) {
                CUDA_OK(c->d_fused_counters.resize(NCTR));
                CUDA_OK(cudaMemset(c->d_fused_counters.p, 0, NCTR * sizeof(unsigned)));
            }
            t.halo_on = 1;
            t.halo = *hf;
            t.halo.seq = c->d_fused_counters.p + kMaxBurst * kCtrStride;
            if (c->halo_burst_len > 1 && c->pdl && !hf->ll && !hf->push_only) {
                if (c->halo_burst_pos < 0 || c->halo_burst_pos >= c->halo_burst_len || c->halo_burst_len > kMaxBurst)
                    return fail(MIMSEM_ERR_ARG, "fused ghost refresh: burst position / length out of range");
                t.halo.burst_pos = c->halo_burst_pos;
                t.halo.burst_len = c->halo_burst_len;
            }
            t.halo.n_int = hf->push_only ? 0 : c->n_int;
            t.halo.counters = c->d_fused_counters.p;
            t.elist = c->elist_all_identity ? nullptr : c->d_elist_all.p;
            t.recs = c->d_recs_halo.p;
            t.rec_stride = c->rec_stride_halo;
            t.rec_list = c->rec_list_halo;
            l.push_ctas = t.halo.push_ctas = hf->npush > 0 ? std::max(1, std::min(296, hf->push_ctas)) : 0;
            l.push_only = hf->push_only != 0;
        }
#ifdef MIMSEM_DIAG
        t.debug = c->diag_debug;
        t.dbg_times = c->diag_times;
#endif
        std::string err;
        // the ring kernel wants many tiles per SM and neighbouring launches to hide its ramp and tail behind
        const bool ring = !hf && (c->m1_variant == 3 || (c->m1_variant == 4 && !with_h && c->pdl && a.nel >= 4000 && !a.elist));
        int rc3 = ring ? launch_m1_pipe(l, t, st, &err) : 1;
        if (rc3 == 1) rc3 = launch_m1_tile(l, t, st, &err);
        if (rc3 < 0) return fail(MIMSEM_ERR_CUDA, err);
        if (rc3 == 0) {
            if (l.push_only && l.push_ctas == 0) return MIMSEM_OK;
            return finish_launch(c, l.push_only ? "halo push (prologue)" : "apply_M1 (tile)");
        }
        if (hf) return fail(MIMSEM_ERR_UNSUPPORTED, "fused ghost refresh: the tile does not fit in shared memory");
    }
    if (c->m1_variant == 0) {
        launch_m1_regs(c->p, with_h, a, grid_for(threads, 128), st);
        return finish_launch(c, "apply_M1");
    }
    a.Gc = with_h ? c->d_Gch.p : c->d_Gc.p;
    a.Gr = with_h ? c->d_Grh.p : c->d_Gr.p;
    launch_m1_lines(c->p, with_h, false, a, dim3(grid_for(threads, 128), 2 * c->p), st);
    int rc2 = finish_launch(c, "apply_M1");
    if (rc2 || c->n_far == 0) return rc2;
    // partial-sum mode: far lines nobody else in the subdomain computes
    KArgs b = a;
    b.nel = c->n_far;
    b.nbr = c->d_far.p;
    launch_m1_lines(c->p, with_h, true, b, dim3(grid_for((int64_t)b.nel * nlev, 128)), st);
    return finish_launch(c, "apply_M1 far lines");
}

// y = M1_ray(exner, exner_s) x: Umat_ray::assemble(lev, scale, dt, exner_k, exner_s) + MatMult (eul/Assembly.cpp:1858-1979,
// called from eul/Euler_2.cpp:1218-1229, 1276-1277, 1437-1448).  A mass matrix whose point weight is
// dt k_v(exner(q), exner_s(q)) / thick: the M1(h) register kernel with the coefficient turned into the friction weight.
int apply_m1_ray(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, double dt, const double* exner, const double* exner_s,
                 const double* x, double* y, cudaStream_t st) {
    int rc = check_ready(c, true, lev0, nlev, ld, 0);
    if (rc) return rc;
    if ((rc = bind_device(c))) return rc;
    if (!exner || !exner_s || !x || !y) return fail(MIMSEM_ERR_ARG, "null field");
    if (!c->d_det.p) return fail(MIMSEM_ERR_STATE, "set_geom first");
    KArgs a;
    fill_common(c, a, lev0, nlev, ld, scale, 1, 0);
    a.G = c->d_G1.p;   // the weight carries its own 1/det
    a.c = exner;
    a.c2 = exner_s;
    a.ray_dt = dt;
    a.x = x;
    a.y = y;
    const int64_t threads = (int64_t)a.nel * nlev;
    if (threads >= (1ll << 31)) return fail(MIMSEM_ERR_UNSUPPORTED, "more than 2^31 element-levels in one launch");
    if (threads == 0 || dt == 0.0) {
        if (threads && dt == 0.0) return fail(MIMSEM_ERR_ARG, "Umat_ray with dt = 0 is the zero operator");
        return MIMSEM_OK;
    }
    launch_m1_regs(c->p, true, a, grid_for(threads, 128), st);
    return finish_launch(c, "apply_M1ray");
}

int apply_m2(mimsem_gpu_ctx* c, bool with_h, int lev0, int nlev, int ld, double scale, int tpow, int flags,
             const double* h2, const double* x, double* y, cudaStream_t st) {
    int rc = check_ready(c, tpow > 0, lev0, nlev, ld, flags);
    if (rc) return rc;
    if ((rc = bind_device(c))) return rc;
    KArgs a;
    fill_common(c, a, lev0, nlev, ld, scale, tpow, flags);
    a.G = with_h ? c->d_W2h.p : c->d_W2.p;
    a.c = h2;
    a.x = x;
    a.y = y;
    const int64_t threads = (int64_t)a.nel * nlev;
    if (threads == 0) return MIMSEM_OK;
    // TMA tile kernel (same eligibility as M1's)
    const bool aligned = ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && (!with_h || (uintptr_t)h2 % 16 == 0);
    const bool thick_ok = tpow == 0 || (!(flags & MIMSEM_FIXED_LEVEL) && c->nkT % 2 == 0 && lev0 % 2 == 0);
    if (c->m2_variant != 0 && c->m2_plan_ok && aligned && thick_ok && nlev % 2 == 0 && ld % 2 == 0 && nlev <= 64) {
        TArgs t;
        std::memset(&t, 0, sizeof(t));
        t.nlev = nlev; t.ld = ld; t.lev0 = lev0; t.nkT = c->nkT; t.tpow = tpow;
        t.contig_x = (ld == nlev); t.contig_t = (c->nkT == nlev);
        t.scale = scale;
        t.prefetch_ahead = c->prefetch_ahead * 2;   // 64-thread CTAs: about twice as many tiles are resident
        t.pdl = c->pdl;
        t.prefetch_own_slots = c->p * c->p;
        t.elist = a.elist;
        t.recs = with_h ? c->d_recs_m2h.p : c->d_recs_m2.p;
        t.rec_stride = with_h ? c->rec_stride_m2h : c->rec_stride_m2;
        t.rec_hdr = 1;
        t.geo = with_h ? c->d_geo_m2h.p : c->d_geo_m2.p;
        t.x = x; t.c = h2; t.tinv = a.tinv; t.y = y;
        copy_basis(c, t);
        for (size_t i = 0; i < sizeof(t.E) / sizeof(double); i++) t.Es[i] = scale * t.E[i];
        std::string err;
        const int rc3 = launch_m2_tile(c->p, with_h, t, a.nel, st, &err);
        if (rc3 < 0) return fail(MIMSEM_ERR_CUDA, err);
        if (rc3 == 0) return finish_launch(c, "apply_M2 (tile)");
    }
    launch_m2(c->p, with_h, a, grid_for(threads, 128), st);
    return finish_launch(c, "apply_M2");
}

int apply_k(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* u1,
            const double* x, double* y, cudaStream_t st) {
    int rc = check_ready(c, tpow > 0, lev0, nlev, ld, flags);
    if (rc) return rc;
    if ((rc = bind_device(c))) return rc;
    KArgs a;
    fill_common(c, a, lev0, nlev, ld, scale, tpow, flags);
    a.G = c->d_G1h.p;
    a.c = u1;
    a.x = x;
    a.y = y;
    const int64_t threads = (int64_t)a.nel * nlev;
    if (threads == 0) return MIMSEM_OK;
    // TMA tile kernel (same eligibility as M1's)
    const bool aligned = ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)u1 % 16 == 0);
    const bool thick_ok = tpow == 0 || (!(flags & MIMSEM_FIXED_LEVEL) && c->nkT % 2 == 0 && lev0 % 2 == 0);
    if (c->k_variant != 0 && c->k_plan_ok && c->tma_ok && aligned && thick_ok && nlev % 2 == 0 && ld % 2 == 0 && nlev <= 64) {
        TArgs t;
        std::memset(&t, 0, sizeof(t));
        t.nlev = nlev; t.ld = ld; t.lev0 = lev0; t.nkT = c->nkT; t.tpow = tpow;
        t.contig_x = (ld == nlev); t.contig_t = (c->nkT == nlev);
        t.scale = scale;
        t.prefetch_ahead = c->prefetch_ahead;
        t.pdl = c->pdl;
        t.prefetch_own_slots = 2 * c->p * c->p + 2 * c->p;
        t.elist = a.elist;
        t.recs = c->d_recs_k.p;
        t.rec_stride = c->rec_stride_k;
        t.rec_hdr = 1;
        t.geo = c->d_geo_k.p;
        t.x = x; t.c = u1; t.tinv = a.tinv; t.y = y;
        copy_basis(c, t);
        std::string err;
        const int rc3 = launch_k_tile(c->p, t, a.nel, st, &err);
        if (rc3 < 0) return fail(MIMSEM_ERR_CUDA, err);
        if (rc3 == 0) return finish_launch(c, "apply_K (tile)");
    }
    launch_k_regs(c->p, a, grid_for(threads, 128), st);
    return finish_launch(c, "apply_K");
}

int apply_m0(mimsem_gpu_ctx* c, bool with_h, int lev0, int nlev, int ld, double scale, int tpow, int flags,
             const double* h2, const double* x, double* y, cudaStream_t st) {
    int rc = check_ready(c, tpow > 0, lev0, nlev, ld, flags);
    if (rc) return rc;
    if ((rc = bind_device(c))) return rc;
    NodeArgs a;
    a.n0 = c->n0_owned >= 0 ? c->n0_owned : c->n0;   // subdomain: owned nodes first, their (element, point) pairs all local
    a.nlev = nlev;
    a.ld = ld;
    a.lev0 = lev0;
    a.lev_stride = (flags & MIMSEM_FIXED_LEVEL) ? 0 : 1;
    a.nkT = c->nkT;
    a.tpow = tpow;
    a.scale = scale;
    a.adj_ptr = c->d_adj_ptr.p;
    a.adj_eq = c->d_adj_eq.p;
    a.node_q = c->d_node_q.p;
    a.el2 = c->d_el2.p;
    a.D0 = c->d_D0.p;
    a.wq = c->d_wq.p;
    a.tinv = c->d_tinv.p;
    a.c = h2;
    a.x = x;
    a.y = y;
    copy_basis(c, a);
    if (a.n0 == 0) return MIMSEM_OK;
    {
        const FastDiv fd = make_fastdiv((unsigned)nlev);
        a.div_m = fd.m;
        a.div_s = fd.s;
    }
    const bool vec2 = x && !with_h && nlev % 2 == 0 && ld % 2 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0 &&
                      (tpow == 0 || (!(flags & MIMSEM_FIXED_LEVEL) && c->nkT % 2 == 0 && lev0 % 2 == 0));
    if (vec2) {
        a.nlev = nlev / 2;
        const FastDiv fd = make_fastdiv((unsigned)a.nlev);
        a.div_m = fd.m;
        a.div_s = fd.s;
        k_apply_m0_vec2<<<grid_for((int64_t)a.n0 * a.nlev, 256), 256, 0, st>>>(a);
        return finish_launch(c, "apply_M0");
    }
    const int64_t threads = (int64_t)a.n0 * nlev;
    launch_m0(c->p, with_h, a, grid_for(threads, 128), st);
    return finish_launch(c, "apply_M0");
}

// RotMat::assemble(q0) / RotMat_up::assemble(q0, ul, fac, dt) + MatMult
int apply_rot(mimsem_gpu_ctx* c, bool up, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* q0,
              const double* u1, double tau, const double* x, double* y, cudaStream_t st) {
    int rc = check_ready(c, tpow > 0, lev0, nlev, ld, flags);
    if (rc) return rc;
    if ((rc = bind_device(c))) return rc;
    if (!q0 || !x || !y || (up && !u1)) return fail(MIMSEM_ERR_ARG, "null field");
    KArgs a;
    fill_common(c, a, lev0, nlev, ld, scale, tpow, flags);
    a.G = nullptr;
    a.x = x;
    a.y = y;
    a.el0 = c->d_el0.p;
    a.q0 = q0;
    a.u1 = u1;
    a.J4 = c->d_J4.p;
    a.det = c->d_det.p;
    a.Wr = c->d_Wr.p;
    a.tau = tau;
    for (int i = 0; i <= kMaxP; i++) a.xn[i] = i <= c->p ? c->xn[i] : 0.0;
    for (int i = 0; i <= kMaxP; i++) {
        double d = 1.0;
        for (int j = 0; j <= c->p; j++)
            if (j != i && i <= c->p) d *= c->xn[i] - c->xn[j];
        a.wb[i] = i <= c->p ? 1.0 / d : 0.0;
    }
    const int64_t threads = (int64_t)a.nel * nlev;
    if (threads >= (1ll << 31)) return fail(MIMSEM_ERR_UNSUPPORTED, "more than 2^31 element-levels in one launch");
    if (threads == 0) return MIMSEM_OK;
    launch_rot(c->p, up, a, grid_for(threads, 128), st);
    return finish_launch(c, up ? "apply_R_up" : "apply_R");
}

// Phmat::assemble_up(ul, hl, fac, dt) + MatMult
int apply_m0h_up(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* h2,
                 const double* u1, double tau, const double* x, double* y, cudaStream_t st) {
    int rc = check_ready(c, tpow > 0, lev0, nlev, ld, flags);
    if (rc) return rc;
    if ((rc = bind_device(c))) return rc;
    if (!h2 || !u1 || !x || !y) return fail(MIMSEM_ERR_ARG, "null field");
    if (c->nel_owned != c->nel_total && c->n0_owned < 0)
        return fail(MIMSEM_ERR_UNSUPPORTED, "0-form operators on a subdomain need the option n0_owned (owned nodes first, all their elements local)");
    NodeArgs a;
    a.n0 = c->n0_owned >= 0 ? c->n0_owned : c->n0;
    a.nlev = nlev;
    a.ld = ld;
    a.lev0 = lev0;
    a.lev_stride = (flags & MIMSEM_FIXED_LEVEL) ? 0 : 1;
    a.nkT = c->nkT;
    a.tpow = tpow;
    a.scale = scale;
    a.adj_ptr = c->d_adj_ptr.p;
    a.adj_eq = c->d_adj_eq.p;
    a.node_q = c->d_node_q.p;
    a.el2 = c->d_el2.p;
    a.D0 = c->d_D0.p;
    a.wq = c->d_wq.p;
    a.tinv = c->d_tinv.p;
    a.c = h2;
    a.x = x;
    a.y = y;
    a.el0 = c->d_el0.p;
    a.el1x = c->d_el1x.p;
    a.el1y = c->d_el1y.p;
    a.u1 = u1;
    a.J4 = c->d_J4.p;
    a.det = c->d_det.p;
    a.tau = tau;
    for (int i = 0; i <= kMaxP; i++) a.xn[i] = i <= c->p ? c->xn[i] : 0.0;
    for (int i = 0; i <= kMaxP; i++) {
        double d = 1.0;
        for (int j = 0; j <= c->p; j++)
            if (j != i && i <= c->p) d *= c->xn[i] - c->xn[j];
        a.wb[i] = i <= c->p ? 1.0 / d : 0.0;
    }
    copy_basis(c, a);
    const FastDiv fd = make_fastdiv((unsigned)nlev);
    a.div_m = fd.m;
    a.div_s = fd.s;
    const int64_t threads = (int64_t)a.n0 * nlev;
    launch_m0h_up(c->p, a, grid_for(threads, 128), st);
    return finish_launch(c, "apply_M0h_up");
}

// diag(M1) (invert: its reciprocal) for the same arguments as apply_m1
int diag_m1(mimsem_gpu_ctx* c, bool invert, int lev0, int nlev, int ld, double scale, int tpow, int flags, double* d, cudaStream_t st) {
    int rc = check_ready(c, tpow > 0, lev0, nlev, ld, flags);
    if (rc) return rc;
    if ((rc = bind_device(c))) return rc;
    if (!d) return fail(MIMSEM_ERR_ARG, "null output");
    KArgs a;
    fill_common(c, a, lev0, nlev, ld, scale, tpow, flags);
    a.G = c->d_G1.p;
    a.x = nullptr;
    a.y = d;
    const int64_t threads = (int64_t)a.nel * nlev;
    if (threads == 0) return MIMSEM_OK;
    launch_diag_m1(c->p, invert, a, grid_for(threads, 128), st);
    return finish_launch(c, "diag_M1");
}

// z = blockdiag(M1)^-1 r, one block per owned element (PCBJACOBI of eul/HorizSolve.cpp:77-84)
int bjacobi_m1(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* r, double* z, cudaStream_t st) {
    int rc = check_ready(c, tpow > 0, lev0, nlev, ld, flags);
    if (rc) return rc;
    if ((rc = bind_device(c))) return rc;
    if (!r || !z) return fail(MIMSEM_ERR_ARG, "null field");
    if (c->mode != 0) return fail(MIMSEM_ERR_UNSUPPORTED, "element-block Jacobi needs owner-computes mode");
    KArgs a;
    fill_common(c, a, lev0, nlev, ld, scale, tpow, flags);
    a.G = c->d_G1.p;
    a.x = r;
    a.y = z;
    if ((int64_t)a.nel * nlev == 0) return MIMSEM_OK;
    std::string err;
    const int rl = launch_bjacobi_m1(c->p, a, st, &err);
    if (rl == -2) return fail(MIMSEM_ERR_UNSUPPORTED, err);
    if (rl < 0) return fail(MIMSEM_ERR_CUDA, err);
    return finish_launch(c, "pc_bjacobi_M1");
}

// HaloFused of one fused launch in mode 0 (push and consume the rows of `x` itself) from the C-ABI descriptor
HaloFused fused_from_desc(const mimsem_halo_desc* d, const double* x) {
    HaloFused hf;
    std::memset(&hf, 0, sizeof(hf));
    hf.npush = d->npush;
    hf.npull = d->npull;
    hf.push_ctas = d->push_ctas;
    hf.push = (const HaloPeer*)d->d_push;
    hf.pull = (const HaloPeer*)d->d_pull;
    hf.inbox = (const double*)d->d_inbox;
    hf.parity_stride = d->stride;
    hf.nbuf = d->nbuf;
    hf.ll = d->ll;
    hf.epoch = (unsigned long long*)d->d_epoch;
    hf.err = d->d_err;
    hf.x_push = x;
    return hf;
}

// x = M1^-1 b by Jacobi-preconditioned CG, all levels at once (per-level step lengths).  hd / rd: element-partitioned
// solve -- the operator is the fused ghost-refresh + M1 launch, vectors live on the OWNED rows (the first
// nel_owned 2 p^2 rows of the engine's edge order) and the dot products are completed over peer memory in k_cg_finish.
int solve_m1(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* b, double* x,
             double rtol, int maxit, int* iters, double* relres, cudaStream_t st, const mimsem_halo_desc* hd = nullptr,
             const mimsem_reduce_desc* rd = nullptr) {
    int rc = check_ready(c, tpow > 0, lev0, nlev, ld, flags);
    if (rc) return rc;
    if ((rc = bind_device(c))) return rc;
    if (!b || !x) return fail(MIMSEM_ERR_ARG, "null field");
    const bool dist = hd != nullptr || rd != nullptr;
    if (dist && (!hd || !rd || rd->world < 2 || rd->rank < 0 || rd->rank >= rd->world || !rd->d_peer_areas || !rd->d_seq || !rd->d_err))
        return fail(MIMSEM_ERR_ARG, "partitioned solve_M1 needs both a halo and a reduction descriptor");
    if (c->mode != 0 || (!dist && c->nel_owned != c->nel_total))
        return fail(MIMSEM_ERR_UNSUPPORTED, "solve_M1 on a subdomain needs mimsem_gpu_solve_M1_dist (ghost refresh + reduction descriptors)");
    if (nlev > 64) return fail(MIMSEM_ERR_UNSUPPORTED, "solve_M1: at most 64 levels per call");
    if (flags & (MIMSEM_SUBSET_INTERIOR | MIMSEM_SUBSET_BOUNDARY)) return fail(MIMSEM_ERR_ARG, "solve_M1 takes no element subset");
    if (!(rtol > 0.0) || maxit < 1) return fail(MIMSEM_ERR_ARG, "bad tolerance / iteration limit");
    const size_t nfield = (size_t)c->n1 * ld;
    CgArgs g;
    std::memset(&g, 0, sizeof(g));
    g.nrows = dist ? (int64_t)c->nel_owned * 2 * c->p * c->p : c->n1;
    g.nlev = nlev;
    g.ld = ld;
    g.nblocks = (int)((g.nrows + CG_ROWS - 1) / CG_ROWS);
    g.world = dist ? rd->world : 1;
    g.rank = dist ? rd->rank : 0;
    g.peer_area = dist ? (uint4* const*)rd->d_peer_areas : nullptr;
    g.seq = dist ? (unsigned long long*)rd->d_seq : nullptr;
    g.err = dist ? rd->d_err : nullptr;
    CUDA_OK(c->cg_r.resize(nfield));
    CUDA_OK(c->cg_p.resize(nfield));
    CUDA_OK(c->cg_q.resize(nfield));
    CUDA_OK(c->cg_dinv.resize(nfield));
    CUDA_OK(c->cg_partial.resize((size_t)3 * g.nblocks * 64));
    CUDA_OK(c->cg_scal.resize(8 * 64));
    CUDA_OK(cudaMemsetAsync(c->cg_scal.p, 0, 8 * 64 * sizeof(double), st));
    g.x = x; g.r = c->cg_r.p; g.p = c->cg_p.p; g.q = c->cg_q.p; g.dinv = c->cg_dinv.p; g.b = b;
    g.partial = c->cg_partial.p;
    g.scal = c->cg_scal.p;
    g.tol2 = rtol * rtol;
    if ((rc = diag_m1(c, true, lev0, nlev, ld, scale, tpow, flags, c->cg_dinv.p, st))) return rc;
    k_cg_step<0><<<g.nblocks, 256, 0, st>>>(g);
    k_cg_finish<0><<<1, 64 * CG_FIN_G, 0, st>>>(g);
    c->launches += 2;
    std::vector<double> h(8 * 64);
    int it = 0;
    double worst = 0.0;
    const int check_every = 4;
    bool done = false;
    while (!done) {
        for (int j = 0; j < check_every && it < maxit; j++, it++) {
            if (dist) {
                const HaloFused hf = fused_from_desc(hd, g.p);
                if ((rc = apply_m1(c, false, lev0, nlev, ld, scale, tpow, flags, nullptr, g.p, c->cg_q.p, st, &hf))) return rc;
            } else if ((rc = apply_m1(c, false, lev0, nlev, ld, scale, tpow, flags, nullptr, g.p, c->cg_q.p, st))) {
                return rc;
            }
            k_cg_step<1><<<g.nblocks, 256, 0, st>>>(g);
            k_cg_finish<1><<<1, 64 * CG_FIN_G, 0, st>>>(g);
            k_cg_step<2><<<g.nblocks, 256, 0, st>>>(g);
            k_cg_finish<2><<<1, 64 * CG_FIN_G, 0, st>>>(g);
            k_cg_step<3><<<g.nblocks, 256, 0, st>>>(g);
            c->launches += 5;
        }
        CUDA_OK(cudaMemcpyAsync(h.data(), c->cg_scal.p, 8 * 64 * sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        done = true;
        worst = 0.0;
        for (int k = 0; k < nlev; k++) {
            const double bb = h[3 * 64 + k], rr = h[2 * 64 + k];
            if (bb > 0.0) worst = std::max(worst, std::sqrt(rr / bb));
            if (h[7 * 64 + k] == 0.0) done = false;
        }
        if (it >= maxit) done = true;
    }
    CUDA_OK(cudaGetLastError());
    if (iters) *iters = it;
    if (relres) *relres = worst;
    return MIMSEM_OK;
}

int solve_m0(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* b, double* x, cudaStream_t st) {
    int rc = check_ready(c, tpow > 0, lev0, nlev, ld, flags);
    if (rc) return rc;
    if ((rc = bind_device(c))) return rc;
    if (!b || !x) return fail(MIMSEM_ERR_ARG, "null field");
    NodeArgs a;
    std::memset(&a, 0, sizeof(a));
    a.n0 = c->n0_owned >= 0 ? c->n0_owned : c->n0;
    a.nlev = nlev;
    a.ld = ld;
    a.lev0 = lev0;
    a.lev_stride = (flags & MIMSEM_FIXED_LEVEL) ? 0 : 1;
    a.nkT = c->nkT;
    a.tpow = tpow;
    a.scale = scale;
    a.node_q = c->d_node_q.p;
    a.D0 = c->d_D0.p;
    a.tinv = c->d_tinv.p;
    a.x = b;
    a.y = x;
    const FastDiv fd = make_fastdiv((unsigned)nlev);
    a.div_m = fd.m;
    a.div_s = fd.s;
    k_solve_m0<<<grid_for((int64_t)a.n0 * nlev, 256), 256, 0, st>>>(a);
    return finish_launch(c, "solve_M0");
}

int apply_inc(mimsem_gpu_ctx* c, int which, int nlev, int ld, const double* x, double* y, cudaStream_t st) {
    if (!c || !c->have_topo) return fail(MIMSEM_ERR_STATE, "set_topo must precede apply_incidence");
    if (which < 0 || which > 3) return fail(MIMSEM_ERR_ARG, "incidence operator id must be 0..3");
    if (nlev < 1 || ld < nlev) return fail(MIMSEM_ERR_ARG, "bad level count / leading dimension");
    int rc = bind_device(c);
    if (rc) return rc;
    const DevEll& d = c->ell[which];
    EllArgs a;
    a.nrows = d.nrows_active;
    if (which == MIMSEM_E01 && c->n0_owned >= 0) a.nrows = c->n0_owned;   // node rows are identity-ordered: owned nodes first
    a.width = d.width;
    a.nlev = nlev;
    a.ld = ld;
    a.rows = d.use_rows ? d.rows.p : nullptr;
    a.col = d.col.p;
    a.sgn = d.sgn.p;
    a.x = x;
    a.y = y;
    {
        const FastDiv fd = make_fastdiv((unsigned)nlev);
        a.div_m = fd.m;
        a.div_s = fd.s;
    }
    if (a.nrows == 0) return MIMSEM_OK;
    if ((which == MIMSEM_E21 || which == MIMSEM_E12) && c->inc_variant != 0 && c->inc_plan_ok && nlev % 2 == 0 && ld % 2 == 0 && nlev <= 64 &&
        (uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0) {
        IncArgs ia;
        ia.nel = c->nel_owned;
        ia.nl2 = nlev / 2;
        ia.ld = ld;
        ia.recs = c->d_inc_recs.p;
        ia.x = x;
        ia.y = y;
        launch_inc_tile(c->p, which == MIMSEM_E21, ia, (cudaStream_t)st);
        return finish_launch(c, "apply_incidence (element kernel)");
    }
    if (a.width > 4) return fail(MIMSEM_ERR_UNSUPPORTED, "incidence stencil wider than 4");
    const int vmax = c->ell_vec;
    if (vmax >= 4 && nlev % 4 == 0 && ld % 4 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0) {
        a.nlev = nlev / 4;
        const FastDiv fd4 = make_fastdiv((unsigned)a.nlev);
        a.div_m = fd4.m;
        a.div_s = fd4.s;
        k_apply_ell<4><<<grid_for(a.nrows * a.nlev, 256), 256, 0, st>>>(a);
        return finish_launch(c, "apply_incidence");
    }
    if (vmax >= 2 && nlev % 2 == 0 && ld % 2 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0) {
        a.nlev = nlev / 2;
        const FastDiv fd2 = make_fastdiv((unsigned)a.nlev);
        a.div_m = fd2.m;
        a.div_s = fd2.s;
        k_apply_ell<2><<<grid_for(a.nrows * a.nlev, 256), 256, 0, st>>>(a);
        return finish_launch(c, "apply_incidence");
    }
    k_apply_ell<1><<<grid_for(a.nrows * nlev, 256), 256, 0, st>>>(a);
    return finish_launch(c, "apply_incidence");
}

// x = M2^-1 b (WITH_H: M2(rho)^-1 b), element-local dense solves
int solve_m2(mimsem_gpu_ctx* c, bool with_h, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* h2,
             const double* b, double* x, cudaStream_t st) {
    int rc = check_ready(c, tpow > 0, lev0, nlev, ld, flags);
    if (rc) return rc;
    if ((rc = bind_device(c))) return rc;
    if (!b || !x || (with_h && !h2)) return fail(MIMSEM_ERR_ARG, "null field");
    KArgs a;
    fill_common(c, a, lev0, nlev, ld, scale, tpow, flags);
    a.G = with_h ? c->d_W2h.p : c->d_W2.p;
    a.c = h2;
    a.x = b;
    a.y = x;
    if ((int64_t)a.nel * nlev == 0) return MIMSEM_OK;
    std::string err;
    if (launch_solve_m2(c->p, with_h, a, st, &err) < 0) return fail(MIMSEM_ERR_CUDA, err);
    return finish_launch(c, "solve_M2");
}

// L2Vecs::HorizToVert / VertToHoriz on device-resident 2-form fields
int l2vecs(mimsem_gpu_ctx* c, bool to_vert, int nlev, int ld, const double* in, double* out, cudaStream_t st) {
    if (!c || !c->have_topo) return fail(MIMSEM_ERR_STATE, "set_topo first");
    if (!in || !out) return fail(MIMSEM_ERR_ARG, "null field");
    if (nlev < 1 || ld < nlev) return fail(MIMSEM_ERR_ARG, "bad level count / leading dimension");
    int rc = bind_device(c);
    if (rc) return rc;
    const int p2 = c->p * c->p;
    const size_t smem = (size_t)nlev * (p2 + 1) * sizeof(double);
    if (smem > 48 * 1024) return fail(MIMSEM_ERR_UNSUPPORTED, "L2Vecs relabelling: more than 48 KB per element column");
    if (to_vert) k_l2vecs<true><<<c->nel_owned, 256, smem, st>>>(p2, nlev, ld, c->d_el2.p, in, out);
    else k_l2vecs<false><<<c->nel_owned, 256, smem, st>>>(p2, nlev, ld, c->d_el2.p, in, out);
    return finish_launch(c, "L2Vecs relabel");
}

int transpose(mimsem_gpu_ctx* c, bool to_columns, int space, int64_t n, int nlev, int ld, const double* in, double* out,
              cudaStream_t st) {
    if (!c) return fail(MIMSEM_ERR_ARG, "null context");
    if (n < 1 || nlev < 1 || ld < nlev) return fail(MIMSEM_ERR_ARG, "bad transpose shape");
    const int* perm = nullptr;
    if (space == 1) {
        if (!c->have_topo || n != c->n1) return fail(MIMSEM_ERR_ARG, "1-form layout conversion needs n == n1 after set_topo");
        perm = c->d_perm1.p;
    } else if (space == 0 || space == 2) {
        if (c->have_topo && n != (space == 0 ? c->n0 : c->n2)) return fail(MIMSEM_ERR_ARG, "field length does not match the space");
    } else if (space != -1) {
        return fail(MIMSEM_ERR_ARG, "space must be 0, 1, 2 (k-forms) or -1 (no renumbering)");
    }
    int rc = bind_device(c);
    if (rc) return rc;
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((nlev + 31) / 32));
    if (to_columns) k_transpose<true><<<grid, 256, 0, st>>>(n, nlev, ld, perm, in, out);
    else k_transpose<false><<<grid, 256, 0, st>>>(n, nlev, ld, perm, in, out);
    return finish_launch(c, "transpose");
}

}  // namespace

// ==========================================================================================
// C ABI

// the inbox of a space holds halo_max_levels levels per ghost row (declared by whoever allocated it)
static int check_halo_levels(const mimsem_gpu_ctx* c, int nlev, int ld, int nbuf) {
    if (nlev < 1 || ld < nlev) return fail(MIMSEM_ERR_ARG, "halo exchange: bad level count / leading dimension");
    if (nbuf < 2 || nbuf > 4) return fail(MIMSEM_ERR_ARG, "the inbox has 2..4 copies");
    if (c->halo_max_levels > 0 && nlev > c->halo_max_levels)
        return fail(MIMSEM_ERR_ARG, "halo exchange: more levels than the inbox was allocated for (option halo_max_levels)");
    return MIMSEM_OK;
}

extern "C" {

int mimsem_gpu_create(int device, mimsem_gpu_ctx** out) {
    if (!out) return fail(MIMSEM_ERR_ARG, "null output pointer");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(MIMSEM_ERR_CUDA, std::string("no CUDA device available (this library has no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(MIMSEM_ERR_ARG, "device ordinal out of range");
    CUDA_OK(cudaSetDevice(device));
    mimsem_gpu_ctx* c = new mimsem_gpu_ctx;
    c->device = device;
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete c;
        return fail(MIMSEM_ERR_CUDA, cudaGetErrorString(e));
    }
    // test / tuning knobs: the environment is consulted here, once -- never on a launch path
    static const char* const names[] = {"m1_variant", "k_variant", "ell_vec", "prefetch_ahead", "m1_min_blocks", "host_chunk", "m2_variant", "inc_variant"};
    static const char* const envs[] = {"MIMSEM_M1_VARIANT", "MIMSEM_K_VARIANT", "MIMSEM_ELL_VEC", "MIMSEM_PREFETCH", "MIMSEM_M1_MINB", "MIMSEM_HOST_CHUNK", "MIMSEM_M2_VARIANT", "MIMSEM_INC_VARIANT"};
    for (int i = 0; i < 8; i++)
        if (const char* v = getenv(envs[i])) mimsem_gpu_set_option(c, names[i], atoll(v));
    *out = c;
    return MIMSEM_OK;
}

int mimsem_gpu_set_option(mimsem_gpu_ctx* c, const char* name, long long value) {
    if (!c || !name) return fail(MIMSEM_ERR_ARG, "null argument");
    const std::string n(name);
    const int v = (int)value;
    if (n == "m1_variant" && v >= 0 && v <= 4) c->m1_variant = v;
    else if (n == "k_variant" && v >= 0 && v <= 1) c->k_variant = v;
    else if (n == "m2_variant" && v >= 0 && v <= 1) c->m2_variant = v;
    else if (n == "inc_variant" && v >= 0 && v <= 1) c->inc_variant = v;
    else if (n == "ell_vec" && (v == 1 || v == 2 || v == 4)) c->ell_vec = v;
    else if (n == "prefetch_ahead" && v >= 0) c->prefetch_ahead = v;
    else if (n == "m1_min_blocks" && v >= 0 && v <= 8) c->m1_min_blocks = v;
    else if (n == "pdl_independent" && v >= 0 && v <= 1) c->pdl = v;
    else if (n == "halo_burst_len" && v >= 0 && v <= kMaxBurst) c->halo_burst_len = v;
    else if (n == "halo_burst_pos" && v >= 0 && v < kMaxBurst) c->halo_burst_pos = v;
    else if (n == "host_chunk" && v >= 1) c->host_chunk = v;
    else if (n == "halo_max_levels" && v >= 0) c->halo_max_levels = v;
    else if (n == "n0_owned" && v >= -1) c->n0_owned = v;
#ifdef MIMSEM_DIAG
    else if (n == "diag_debug") c->diag_debug = v;
    else if (n == "diag_times") c->diag_times = (long long*)value;
#endif
    else return fail(MIMSEM_ERR_ARG, "unknown option or value out of range: " + n);
    return MIMSEM_OK;
}

int mimsem_gpu_destroy(mimsem_gpu_ctx* ctx) {
    if (!ctx) return MIMSEM_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    for (int b = 0; b < mimsem_gpu_ctx::HOST_SLOTS - 1; b++)
        if (ctx->host_stream[b]) cudaStreamDestroy(ctx->host_stream[b]);
    for (int b = 0; b < 3; b++)
        if (ctx->ev_host[b]) cudaEventDestroy(ctx->ev_host[b]);
    delete ctx;
    return MIMSEM_OK;
}

int mimsem_gpu_set_basis(mimsem_gpu_ctx* c, int p, int m, const double* h_w, const double* h_ljxi, const double* h_ejxi) {
    if (!c || !h_w || !h_ljxi || !h_ejxi) return fail(MIMSEM_ERR_ARG, "null argument");
    if (p < 1 || p > kMaxP || m < 1) return fail(MIMSEM_ERR_ARG, "unsupported order");
    if (m != p) return fail(MIMSEM_ERR_UNSUPPORTED, "quadrature order must equal the element order (every BASELINE configuration)");
    // with m == p the nodal table must be the identity (up to signed zeros / a few ulps in box/)
    for (int q = 0; q <= m; q++)
        for (int j = 0; j <= p; j++)
            if (std::fabs(h_ljxi[q * (p + 1) + j] - (q == j ? 1.0 : 0.0)) > 1e-12)
                return fail(MIMSEM_ERR_ARG, "ljxi is not the identity although m == p");
    c->p = p;
    c->m = m;
    c->w.assign(h_w, h_w + m + 1);
    c->ejxi.assign(h_ejxi, h_ejxi + (size_t)(m + 1) * p);
    {
        c->xn.assign(p + 1, 0.0);
        std::vector<double> wtmp(p + 1);
        if (mimsem_basis_gll(p, c->xn.data(), wtmp.data()) != MIMSEM_OK) return MIMSEM_ERR_ARG;
    }
    c->have_basis = true;
    c->have_topo = c->have_geom = false;
    return MIMSEM_OK;
}

int mimsem_gpu_set_topo(mimsem_gpu_ctx* c, int nel_total, int nel_owned, int n0, int n1, int n2, int nq, int mode,
                        const int* h_el0, const int* h_el1x, const int* h_el1y, const int* h_el2, const int* h_elq) {
    if (!c || !h_el0 || !h_el1x || !h_el1y || !h_el2 || !h_elq) return fail(MIMSEM_ERR_ARG, "null argument");
    if (!c->have_basis) return fail(MIMSEM_ERR_STATE, "set_basis first");
    if (nel_owned < 1 || nel_total < nel_owned || n0 < 1 || n1 < 1 || n2 < 1 || nq < 1 || mode < 0 || mode > 1)
        return fail(MIMSEM_ERR_ARG, "bad sizes");
    int rc = bind_device(c);
    if (rc) return rc;
    const int P = c->p, NP1 = P + 1;
    c->nel_total = nel_total;
    c->nel_owned = nel_owned;
    c->n0 = n0;
    c->n1 = n1;
    c->n2 = n2;
    c->nq = nq;
    c->mode = mode;
    c->h_el0.assign(h_el0, h_el0 + (size_t)nel_total * NP1 * NP1);
    c->h_el1x.assign(h_el1x, h_el1x + (size_t)nel_total * P * NP1);
    c->h_el1y.assign(h_el1y, h_el1y + (size_t)nel_total * P * NP1);
    c->h_el2.assign(h_el2, h_el2 + (size_t)nel_total * P * P);
    c->h_elq.assign(h_elq, h_elq + (size_t)nel_total * NP1 * NP1);
    for (int v : c->h_el2)
        if (v < 0 || v >= n2) return fail(MIMSEM_ERR_ARG, "set_topo: face index out of range");
    for (int v : c->h_elq)
        if (v < 0 || v >= nq) return fail(MIMSEM_ERR_ARG, "set_topo: quadrature-point index out of range");
    for (int v : c->h_el1x)
        if (v < 0 || v >= n1) return fail(MIMSEM_ERR_ARG, "set_topo: edge index out of range");
    for (int v : c->h_el1y)
        if (v < 0 || v >= n1) return fail(MIMSEM_ERR_ARG, "set_topo: edge index out of range");
    {
        // Internal 1-form numbering: element-blocked in local element order, each block = the element's
        // west/interior x-edges column-major, then its south/interior y-edges row-major.  With it an
        // element's own edges, a neighbour's west column, south row or whole edge family are each ONE
        // contiguous run of memory, i.e. one TMA bulk copy.  (The reference's own numbering is
        // element-blocked too, but interleaves x and y edges, scr/Proc2.py:105-123.)
        const int N1E = P * NP1;
        c->h_perm1.assign(n1, -1);
        int next = 0;
        for (int e = 0; e < nel_total; e++) {
            for (int ix = 0; ix < P; ix++)
                for (int iy = 0; iy < P; iy++) {
                    int& d = c->h_perm1[c->h_el1x[(size_t)e * N1E + iy * NP1 + ix]];
                    if (d < 0) d = next++;
                }
            for (int iy = 0; iy < P; iy++)
                for (int ix = 0; ix < P; ix++) {
                    int& d = c->h_perm1[c->h_el1y[(size_t)e * N1E + iy * P + ix]];
                    if (d < 0) d = next++;
                }
        }
        for (int i = 0; i < n1; i++)
            if (c->h_perm1[i] < 0) c->h_perm1[i] = next++;
        CUDA_OK(c->d_perm1.upload(c->h_perm1));
    }
    if ((rc = build_incidence(c))) return rc;   // external numbering (exported CSR), device stencils permuted
    c->h_el1x_ext = c->h_el1x;
    c->h_el1y_ext = c->h_el1y;
    c->n_int = c->n_bnd = 0;
    for (auto& v : c->h_el1x) v = c->h_perm1[v];
    for (auto& v : c->h_el1y) v = c->h_perm1[v];
    if ((rc = build_neighbours(c))) return rc;
    if ((rc = build_node_adjacency(c))) return rc;
    if ((rc = build_inc_plan(c))) return rc;
    {
        // column-major copies so that every GLL line of an element is contiguous (line-task kernels)
        const int N1E = P * NP1, Q2 = NP1 * NP1;
        std::vector<int> xT((size_t)nel_total * N1E), qT((size_t)nel_total * Q2), far;
        for (int e = 0; e < nel_total; e++) {
            for (int iy = 0; iy < P; iy++)
                for (int ix = 0; ix <= P; ix++) xT[(size_t)e * N1E + ix * P + iy] = c->h_el1x[(size_t)e * N1E + iy * NP1 + ix];
            for (int qy = 0; qy <= P; qy++)
                for (int qx = 0; qx <= P; qx++) qT[(size_t)e * Q2 + qx * NP1 + qy] = c->h_elq[(size_t)e * Q2 + qy * NP1 + qx];
        }
        for (int e = 0; e < nel_owned; e++) {
            if (c->h_eflags[e] & 1) far.push_back(e << 1);
            if (c->h_eflags[e] & 2) far.push_back((e << 1) | 1);
        }
        c->n_far = (int)far.size();
        CUDA_OK(c->d_el1xT.upload(xT));
        CUDA_OK(c->d_elqT.upload(qT));
        CUDA_OK(c->d_far.upload(far));
    }
    c->tma_ok = false;
    if (mode == 0) {
        switch (P) {
            case 2: rc = build_tma_plan_p<2>(c); break;
            case 3: rc = build_tma_plan_p<3>(c); break;
            case 4: rc = build_tma_plan_p<4>(c); break;
            case 5: rc = build_tma_plan_p<5>(c); break;
            default: rc = MIMSEM_OK;
        }
        if (rc) return rc;
    }
    c->m2_plan_ok = false;
    switch (P) {
        case 2: rc = build_m2_plan_p<2>(c); break;
        case 3: rc = build_m2_plan_p<3>(c); break;
        case 4: rc = build_m2_plan_p<4>(c); break;
        case 5: rc = build_m2_plan_p<5>(c); break;
        default: rc = MIMSEM_OK;
    }
    if (rc) return rc;
    c->k_plan_ok = false;
    switch (P) {
        case 2: rc = build_k_plan_p<2>(c); break;
        case 3: rc = build_k_plan_p<3>(c); break;
        case 4: rc = build_k_plan_p<4>(c); break;
        case 5: rc = build_k_plan_p<5>(c); break;
        default: rc = MIMSEM_OK;
    }
    if (rc) return rc;
    CUDA_OK(c->d_el0.upload(c->h_el0));
    CUDA_OK(c->d_el1x.upload(c->h_el1x));
    CUDA_OK(c->d_el1y.upload(c->h_el1y));
    CUDA_OK(c->d_el2.upload(c->h_el2));
    CUDA_OK(c->d_elq.upload(c->h_elq));
    CUDA_OK(c->d_nbr.upload(c->h_nbr));
    CUDA_OK(c->d_eflags.upload(c->h_eflags));
    CUDA_OK(c->d_adj_ptr.upload(c->h_adj_ptr));
    CUDA_OK(c->d_adj_eq.upload(c->h_adj_eq));
    CUDA_OK(c->d_node_q.upload(c->h_node_q));
    c->have_topo = true;
    c->have_geom = false;
    return MIMSEM_OK;
}

int mimsem_gpu_set_element_keys(mimsem_gpu_ctx* c, int nel, const int* h_keys) {
    if (!c || (nel > 0 && !h_keys) || nel < 0) return fail(MIMSEM_ERR_ARG, "bad argument");
    c->h_elem_key.assign(h_keys, h_keys + nel);
    c->have_topo = c->have_geom = false;   // takes effect at the next set_topo / set_geom
    return MIMSEM_OK;
}

int mimsem_gpu_set_ghosts(mimsem_gpu_ctx* c, int n1_owned, int n2_owned, int out_counts[2]) {
    if (!c || !c->have_topo) return fail(MIMSEM_ERR_STATE, "set_topo first");
    if (n1_owned < 0 || n1_owned > c->n1 || n2_owned < 0 || n2_owned > c->n2) return fail(MIMSEM_ERR_ARG, "bad owned counts");
    int rc = bind_device(c);
    if (rc) return rc;
    const int P = c->p, N1E = P * (P + 1), N2E = P * P;
    auto reads_ghost = [&](int e) {
        for (int j = 0; j < N1E; j++)
            if (c->h_el1x_ext[(size_t)e * N1E + j] >= n1_owned || c->h_el1y_ext[(size_t)e * N1E + j] >= n1_owned) return true;
        for (int j = 0; j < N2E; j++)
            if (c->h_el2[(size_t)e * N2E + j] >= n2_owned) return true;
        return false;
    };
    std::vector<int> in, bd;
    for (int e = 0; e < c->nel_owned; e++) {
        bool g = reads_ghost(e);
        for (int s = 0; s < 2 && !g; s++) {
            const int nb = c->h_nbr[(size_t)e * 2 + s];
            if (nb >= 0) g = reads_ghost(nb & 0x1fffffff);
        }
        (g ? bd : in).push_back(e);
    }
    c->n_int = (int)in.size();
    c->n_bnd = (int)bd.size();
    CUDA_OK(c->d_elist_int.upload(in));
    CUDA_OK(c->d_elist_bnd.upload(bd));
    {
        std::vector<int> all(in);
        all.insert(all.end(), bd.begin(), bd.end());
        CUDA_OK(c->d_elist_all.upload(all));
        c->elist_all_identity = true;
        for (size_t i = 0; i < all.size(); i++)
            if (all[i] != (int)i) c->elist_all_identity = false;
    }
    c->halo_plan_ok = false;
    if (c->mode == 0 && c->tma_ok) {
        switch (P) {
            case 2: rc = build_tma_plan_p<2>(c, n1_owned); break;
            case 3: rc = build_tma_plan_p<3>(c, n1_owned); break;
            case 4: rc = build_tma_plan_p<4>(c, n1_owned); break;
            case 5: rc = build_tma_plan_p<5>(c, n1_owned); break;
            default: rc = MIMSEM_OK;
        }
        if (rc) return rc;
    }
    if (out_counts) {
        out_counts[0] = c->n_int;
        out_counts[1] = c->n_bnd;
    }
    return MIMSEM_OK;
}

int mimsem_gpu_set_geom(mimsem_gpu_ctx* c, const double* h_J, const double* h_det) {
    if (!c || !h_J || !h_det) return fail(MIMSEM_ERR_ARG, "null argument");
    if (!c->have_topo) return fail(MIMSEM_ERR_STATE, "set_topo first");
    int rc = bind_device(c);
    if (rc) return rc;
    const int NP1 = c->p + 1, Q2 = NP1 * NP1;
    const size_t npts = (size_t)c->nel_total * Q2;
    std::vector<double> G1(npts * 3), G1h(npts * 3), W2(npts), W2h(npts), D0(c->n0, 0.0), wq(Q2);
    for (int q = 0; q < Q2; q++) wq[q] = c->w[q % NP1] * c->w[q / NP1];   // Wii, eul/ElMats.cpp:177
    for (size_t i = 0; i < npts; i++) {
        const double* J = h_J + i * 4;
        const double det = h_det[i];
        if (!(det != 0.0)) return fail(MIMSEM_ERR_ARG, "set_geom: zero Jacobian determinant");
        const double w = wq[i % Q2];
        // metric of the H(div) Piola map, eul/Assembly.cpp:103-105
        const double gaa = J[0] * J[0] + J[2] * J[2];
        const double gab = J[0] * J[1] + J[2] * J[3];
        const double gbb = J[1] * J[1] + J[3] * J[3];
        const double wd = w / det, wdd = w / (det * det);
        G1[i * 3 + 0] = gaa * wd;  G1[i * 3 + 1] = gab * wd;  G1[i * 3 + 2] = gbb * wd;
        G1h[i * 3 + 0] = gaa * wdd; G1h[i * 3 + 1] = gab * wdd; G1h[i * 3 + 2] = gbb * wdd;
        W2[i] = wd;
        W2h[i] = wdd;
    }
    // sum over the (element, point) pairs at each node in adjacency order: element order as MatSetValues(ADD_VALUES)
    // accumulates, or the caller's canonical order (set_element_keys)
    for (int n = 0; n < c->n0; n++)
        for (int j = c->h_adj_ptr[n]; j < c->h_adj_ptr[n + 1]; j++) {
            const size_t i = (size_t)c->h_adj_eq[j];
            D0[n] += wq[i % Q2] * h_det[i];
        }
    {
        // line-ordered copies of the metric: Gc[e][qx][qy] = (g0,g1), Gr[e][qy][qx] = (g1,g2)
        std::vector<double> Gc(npts * 2), Gr(npts * 2), Gch(npts * 2), Grh(npts * 2);
        for (size_t e = 0; e < (size_t)c->nel_total; e++)
            for (int qy = 0; qy < NP1; qy++)
                for (int qx = 0; qx < NP1; qx++) {
                    const size_t i = e * Q2 + qy * NP1 + qx, ic = e * Q2 + qx * NP1 + qy;
                    Gc[ic * 2 + 0] = G1[i * 3 + 0];   Gc[ic * 2 + 1] = G1[i * 3 + 1];
                    Gr[i * 2 + 0] = G1[i * 3 + 1];    Gr[i * 2 + 1] = G1[i * 3 + 2];
                    Gch[ic * 2 + 0] = G1h[i * 3 + 0]; Gch[ic * 2 + 1] = G1h[i * 3 + 1];
                    Grh[i * 2 + 0] = G1h[i * 3 + 1];  Grh[i * 2 + 1] = G1h[i * 3 + 2];
                }
        CUDA_OK(c->d_Gc.upload(Gc));
        CUDA_OK(c->d_Gr.upload(Gr));
        CUDA_OK(c->d_Gch.upload(Gch));
        CUDA_OK(c->d_Grh.upload(Grh));
    }
    if (c->tma_ok) {
        switch (c->p) {
            case 2: rc = build_tma_geo_p<2>(c, G1, c->d_geo); if (!rc) rc = build_tma_geo_p<2>(c, G1h, c->d_geo_h); break;
            case 3: rc = build_tma_geo_p<3>(c, G1, c->d_geo); if (!rc) rc = build_tma_geo_p<3>(c, G1h, c->d_geo_h); break;
            case 4: rc = build_tma_geo_p<4>(c, G1, c->d_geo); if (!rc) rc = build_tma_geo_p<4>(c, G1h, c->d_geo_h); break;
            case 5: rc = build_tma_geo_p<5>(c, G1, c->d_geo); if (!rc) rc = build_tma_geo_p<5>(c, G1h, c->d_geo_h); break;
        }
        if (rc) return rc;
        if ((rc = build_k_geo(c, G1h, c->d_geo_k))) return rc;
    }
    if ((rc = build_m2_geo(c, W2, c->d_geo_m2))) return rc;
    if ((rc = build_m2_geo(c, W2h, c->d_geo_m2h))) return rc;
    CUDA_OK(c->d_G1.upload(G1));
    CUDA_OK(c->d_G1h.upload(G1h));
    CUDA_OK(c->d_W2.upload(W2));
    CUDA_OK(c->d_W2h.upload(W2h));
    CUDA_OK(c->d_D0.upload(D0));
    CUDA_OK(c->d_wq.upload(wq));
    {
        std::vector<double> J4(h_J, h_J + npts * 4), dt(h_det, h_det + npts), Wr(npts);
        // RotMat: (-/+)(J00 J11 - J01 J10) w / det   (src/Assembly.cpp:1369-1370)
        for (size_t i = 0; i < npts; i++) Wr[i] = (h_J[i * 4 + 0] * h_J[i * 4 + 3] - h_J[i * 4 + 1] * h_J[i * 4 + 2]) * wq[i % Q2] / h_det[i];
        CUDA_OK(c->d_J4.upload(J4));
        CUDA_OK(c->d_det.upload(dt));
        CUDA_OK(c->d_Wr.upload(Wr));
    }
    c->have_geom = true;
    return MIMSEM_OK;
}

int mimsem_gpu_set_thickness(mimsem_gpu_ctx* c, int nk, const double* h_thick) {
    if (!c) return fail(MIMSEM_ERR_ARG, "null context");
    if (!c->have_topo) return fail(MIMSEM_ERR_STATE, "set_topo first");
    int rc = bind_device(c);
    if (rc) return rc;
    if (nk == 0) {
        c->nkT = 0;
        return MIMSEM_OK;
    }
    if (nk < 0 || !h_thick) return fail(MIMSEM_ERR_ARG, "bad thickness table");
    std::vector<double> tinv((size_t)c->nq * nk);
    for (int k = 0; k < nk; k++)
        for (int q = 0; q < c->nq; q++) tinv[(size_t)q * nk + k] = 1.0 / h_thick[(size_t)k * c->nq + q];   // eul/Geom.cpp:761
    CUDA_OK(c->d_tinv.upload(tinv));
    {
        // 0.5 (thick[k] + thick[k+1]) per point (eul/Assembly.cpp:1362-1364); the last level has no upper neighbour and
        // keeps its own thickness so that the table has the same shape as the inverse-thickness one
        std::vector<double> tm((size_t)c->nq * nk);
        for (int k = 0; k < nk; k++)
            for (int q = 0; q < c->nq; q++) {
                const double t0 = h_thick[(size_t)k * c->nq + q];
                const double t1 = k + 1 < nk ? h_thick[(size_t)(k + 1) * c->nq + q] : t0;
                tm[(size_t)q * nk + k] = 0.5 * (t0 + t1);
            }
        CUDA_OK(c->d_tmean.upload(tm));
    }
    c->nkT = nk;
    return MIMSEM_OK;
}

int mimsem_gpu_sizes(const mimsem_gpu_ctx* c, int64_t out[9]) {
    if (!c || !out) return fail(MIMSEM_ERR_ARG, "null argument");
    out[0] = c->nel_total; out[1] = c->nel_owned; out[2] = c->n0; out[3] = c->n1; out[4] = c->n2;
    out[5] = c->nq; out[6] = c->nkT; out[7] = c->p; out[8] = c->m;
    return MIMSEM_OK;
}

int mimsem_gpu_levels_to_columns(mimsem_gpu_ctx* c, int space, int64_t n, int nlev, int ld, const double* in, double* out,
                                 void* stream) {
    return transpose(c, true, space, n, nlev, ld, in, out, (cudaStream_t)stream);
}
int mimsem_gpu_columns_to_levels(mimsem_gpu_ctx* c, int space, int64_t n, int nlev, int ld, const double* in, double* out,
                                 void* stream) {
    return transpose(c, false, space, n, nlev, ld, in, out, (cudaStream_t)stream);
}
int mimsem_gpu_form_permutation(const mimsem_gpu_ctx* c, int space, int* perm) {
    if (!c || !perm || !c->have_topo) return fail(MIMSEM_ERR_STATE, "set_topo first");
    if (space == 1) {
        std::copy(c->h_perm1.begin(), c->h_perm1.end(), perm);
    } else if (space == 0 || space == 2) {
        const int n = space == 0 ? c->n0 : c->n2;
        for (int i = 0; i < n; i++) perm[i] = i;
    } else {
        return fail(MIMSEM_ERR_ARG, "space must be 0, 1 or 2");
    }
    return MIMSEM_OK;
}

int mimsem_gpu_apply_M1(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* x,
                        double* y, void* st) {
    return apply_m1(c, false, lev0, nlev, ld, scale, tpow, flags, nullptr, x, y, (cudaStream_t)st);
}
static int apply_m1_halo_impl(int ll, mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* x,
                             double* y, const double* x_push, int mode, int npush, const void* d_push, int npull,
                             const void* d_pull, const double* d_inbox, int64_t parity_stride, int nbuf, int push_ctas,
                             void* d_epoch, int* d_err, void* st) {
    if (!c) return fail(MIMSEM_ERR_ARG, "null context");
    if (int rcl = check_halo_levels(c, nlev, ld, nbuf)) return rcl;
    // bulk copies and 16-byte stores address the inbox copies: every copy must start on a 16-byte boundary
    if ((!ll && parity_stride % 2 != 0) || (uintptr_t)d_inbox % 16 != 0) return fail(MIMSEM_ERR_ARG, "fused ghost refresh: inbox copies must be 16-byte aligned (even stride)");
    if (mode < 0 || mode > 3 || ((mode == 1 || mode == 2) && !x_push)) return fail(MIMSEM_ERR_ARG, "bad pipelining mode / missing field to push");
    if (mode == 3) npush = 0;   // last call of a pipelined sequence: consume what the previous call pushed, push nothing
    if (!c || !d_epoch || !d_err || (npush > 0 && !d_push) || (npull > 0 && (!d_pull || !d_inbox)))
        return fail(MIMSEM_ERR_ARG, "null argument");
    HaloFused hf;
    std::memset(&hf, 0, sizeof(hf));
    hf.npush = npush;
    hf.npull = npull;
    hf.push_ctas = push_ctas;
    hf.push = (const HaloPeer*)d_push;
    hf.pull = (const HaloPeer*)d_pull;
    hf.inbox = d_inbox;
    hf.parity_stride = parity_stride;
    hf.nbuf = nbuf;
    hf.ll = ll;
    hf.epoch = (unsigned long long*)d_epoch;
    hf.err = d_err;
    hf.x_push = mode == 0 ? x : x_push;
    hf.lead = mode == 1 ? 1 : 0;
    hf.push_only = mode == 2 ? 1 : 0;
    return apply_m1(c, false, lev0, nlev, ld, scale, tpow, flags, nullptr, x, y, (cudaStream_t)st, &hf);
}
int mimsem_gpu_apply_M1_halo(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* x,
                             double* y, const double* x_push, int mode, int npush, const void* d_push, int npull,
                             const void* d_pull, const double* d_inbox, int64_t parity_stride, int nbuf, int push_ctas,
                             void* d_epoch, int* d_err, void* st) {
    return apply_m1_halo_impl(0, c, lev0, nlev, ld, scale, tpow, flags, x, y, x_push, mode, npush, d_push, npull, d_pull, d_inbox,
                              parity_stride, nbuf, push_ctas, d_epoch, d_err, st);
}
int mimsem_gpu_apply_M1_halo_ll(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* x,
                                double* y, const double* x_push, int mode, int npush, const void* d_push, int npull,
                                const void* d_pull, const void* d_inbox_cells, int64_t cell_stride, int nbuf, int push_ctas,
                                void* d_epoch, int* d_err, void* st) {
    return apply_m1_halo_impl(1, c, lev0, nlev, ld, scale, tpow, flags, x, y, x_push, mode, npush, d_push, npull, d_pull,
                              (const double*)d_inbox_cells, cell_stride, nbuf, push_ctas, d_epoch, d_err, st);
}
int mimsem_gpu_apply_M1h(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* h2,
                         const double* x, double* y, void* st) {
    if (!h2) return fail(MIMSEM_ERR_ARG, "null coefficient field");
    return apply_m1(c, true, lev0, nlev, ld, scale, tpow, flags, h2, x, y, (cudaStream_t)st);
}
int mimsem_gpu_apply_M1ray(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, double dt, const double* exner,
                           const double* exner_s, const double* x, double* y, void* st) {
    return apply_m1_ray(c, lev0, nlev, ld, scale, dt, exner, exner_s, x, y, (cudaStream_t)st);
}
int mimsem_gpu_apply_M2(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* x,
                        double* y, void* st) {
    return apply_m2(c, false, lev0, nlev, ld, scale, tpow, flags, nullptr, x, y, (cudaStream_t)st);
}
int mimsem_gpu_apply_M2h(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* h2,
                         const double* x, double* y, void* st) {
    if (!h2) return fail(MIMSEM_ERR_ARG, "null coefficient field");
    return apply_m2(c, true, lev0, nlev, ld, scale, tpow, flags, h2, x, y, (cudaStream_t)st);
}
int mimsem_gpu_apply_M0(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* x,
                        double* y, void* st) {
    return apply_m0(c, false, lev0, nlev, ld, scale, tpow, flags, nullptr, x, y, (cudaStream_t)st);
}
int mimsem_gpu_apply_M0h(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* h2,
                         const double* x, double* y, void* st) {
    if (!h2) return fail(MIMSEM_ERR_ARG, "null coefficient field");
    return apply_m0(c, true, lev0, nlev, ld, scale, tpow, flags, h2, x, y, (cudaStream_t)st);
}
int mimsem_gpu_apply_K(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* u1,
                       const double* x, double* y, void* st) {
    if (!u1) return fail(MIMSEM_ERR_ARG, "null coefficient field");
    return apply_k(c, lev0, nlev, ld, scale, tpow, flags, u1, x, y, (cudaStream_t)st);
}
int mimsem_gpu_apply_R(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* q0,
                       const double* x, double* y, void* st) {
    return apply_rot(c, false, lev0, nlev, ld, scale, tpow, flags, q0, nullptr, 0.0, x, y, (cudaStream_t)st);
}
int mimsem_gpu_apply_R_up(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* q0,
                          const double* u1, double tau, const double* x, double* y, void* st) {
    return apply_rot(c, true, lev0, nlev, ld, scale, tpow, flags, q0, u1, tau, x, y, (cudaStream_t)st);
}
int mimsem_gpu_apply_M0h_up(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* h2,
                            const double* u1, double tau, const double* x, double* y, void* st) {
    return apply_m0h_up(c, lev0, nlev, ld, scale, tpow, flags, h2, u1, tau, x, y, (cudaStream_t)st);
}
int mimsem_gpu_solve_M1(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* b, double* x,
                        double rtol, int maxit, int* iters, double* relres, void* st) {
    return solve_m1(c, lev0, nlev, ld, scale, tpow, flags, b, x, rtol, maxit, iters, relres, (cudaStream_t)st);
}
int mimsem_gpu_pc_bjacobi_M1(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* r, double* z,
                             void* st) {
    return bjacobi_m1(c, lev0, nlev, ld, scale, tpow, flags, r, z, (cudaStream_t)st);
}
int mimsem_gpu_solve_M1_dist(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* b, double* x,
                             double rtol, int maxit, int* iters, double* relres, const mimsem_halo_desc* halo,
                             const mimsem_reduce_desc* reduce, void* st) {
    if (!halo || !reduce) return fail(MIMSEM_ERR_ARG, "null descriptor");
    if (int rcl = check_halo_levels(c, nlev, ld, halo->nbuf)) return rcl;
    return solve_m1(c, lev0, nlev, ld, scale, tpow, flags, b, x, rtol, maxit, iters, relres, (cudaStream_t)st, halo, reduce);
}
int mimsem_gpu_solve_M0(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* b, double* x,
                        void* st) {
    return solve_m0(c, lev0, nlev, ld, scale, tpow, flags, b, x, (cudaStream_t)st);
}
int mimsem_gpu_diag_M1(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, double* d, void* st) {
    return diag_m1(c, false, lev0, nlev, ld, scale, tpow, flags, d, (cudaStream_t)st);
}
int mimsem_gpu_diag_M0(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* h2, double* d,
                       void* st) {
    if (!d) return fail(MIMSEM_ERR_ARG, "null output");
    return apply_m0(c, h2 != nullptr, lev0, nlev, ld, scale, tpow, flags, h2, nullptr, d, (cudaStream_t)st);
}
int mimsem_gpu_solve_M2(mimsem_gpu_ctx* c, int lev0, int nlev, int ld, double scale, int tpow, int flags, const double* h2,
                        const double* b, double* x, void* st) {
    return solve_m2(c, h2 != nullptr, lev0, nlev, ld, scale, tpow, flags, h2, b, x, (cudaStream_t)st);
}
int mimsem_gpu_apply_UtQW(mimsem_gpu_ctx* c, int nlev, int ld, double scale, const double* u1, const double* x2, double* y1, void* st) {
    // UtQWmat(u1) x2 == Uhmat(h2 := x2) u1 without thickness factors: the form is bilinear in (2-form, 1-form)
    if (!u1 || !x2) return fail(MIMSEM_ERR_ARG, "null field");
    return apply_m1(c, true, 0, nlev, ld, scale, 0, 0, x2, u1, y1, (cudaStream_t)st);
}
int mimsem_gpu_columns_to_vertical(mimsem_gpu_ctx* c, int nlev, int ld, const double* cols, double* vert, void* st) {
    return l2vecs(c, true, nlev, ld, cols, vert, (cudaStream_t)st);
}
int mimsem_gpu_vertical_to_columns(mimsem_gpu_ctx* c, int nlev, int ld, const double* vert, double* cols, void* st) {
    return l2vecs(c, false, nlev, ld, vert, cols, (cudaStream_t)st);
}
int mimsem_gpu_apply_incidence(mimsem_gpu_ctx* c, int which, int nlev, int ld, const double* x, double* y, void* st) {
    return apply_inc(c, which, nlev, ld, x, y, (cudaStream_t)st);
}

int mimsem_gpu_incidence_csr(const mimsem_gpu_ctx* c, int which, int64_t out_sizes[3], int64_t* indptr, int* indices, double* values) {
    if (!c || !c->have_topo) return fail(MIMSEM_ERR_STATE, "set_topo first");
    if (which < 0 || which > 3) return fail(MIMSEM_ERR_ARG, "incidence operator id must be 0..3");
    const HostCsr& m = c->csr[which];
    if (out_sizes) {
        out_sizes[0] = m.nrows;
        out_sizes[1] = m.ncols;
        out_sizes[2] = (int64_t)m.indices.size();
    }
    if (indptr) std::copy(m.indptr.begin(), m.indptr.end(), indptr);
    if (indices) std::copy(m.indices.begin(), m.indices.end(), indices);
    if (values) std::copy(m.values.begin(), m.values.end(), values);
    return MIMSEM_OK;
}

int mimsem_gpu_apply_host(mimsem_gpu_ctx* c, int op, int lev0, int nlev, double scale, int tpow, int flags,
                          const double* h_coeff, const double* h_x, double* h_y) {
    return mimsem_gpu_apply_host_up(c, op, lev0, nlev, scale, tpow, flags, h_coeff, nullptr, 0.0, h_x, h_y);
}

int mimsem_gpu_apply_host_up(mimsem_gpu_ctx* c, int op, int lev0, int nlev, double scale, int tpow, int flags,
                             const double* h_coeff, const double* h_u1, double tau, const double* h_x, double* h_y) {
    if (!c || !h_x || !h_y) return fail(MIMSEM_ERR_ARG, "null argument");
    if (!c->have_topo) return fail(MIMSEM_ERR_STATE, "set_topo first");
    int rc = bind_device(c);
    if (rc) return rc;
    // input / output / coefficient spaces
    int sin, sout, scoef = -1;   // k of the k-form spaces
    switch (op) {
        case 0: sin = sout = 1; break;
        case 3: sin = sout = 1; scoef = 2; break;
        case 1: sin = sout = 2; break;
        case 5: sin = sout = 2; scoef = 2; break;
        case 2: sin = sout = 0; break;
        case 6: sin = sout = 0; scoef = 2; break;
        case 4: sin = 1; sout = 2; scoef = 1; break;
        case 7: case 8: sin = sout = 1; scoef = 0; break;   // R(q0), R_up(q0, u1)
        case 9: sin = sout = 0; scoef = 2; break;           // M0h_up(h2, u1)
        case 10 + MIMSEM_E10: sin = 0; sout = 1; break;
        case 10 + MIMSEM_E01: sin = 1; sout = 0; break;
        case 10 + MIMSEM_E21: sin = 1; sout = 2; break;
        case 10 + MIMSEM_E12: sin = 2; sout = 1; break;
        case 14: sin = 2; sout = 1; scoef = 1; break;       // UtQW(u1) x2
        case 15: sin = sout = 0; break;                     // diag M0 (Pvec): h_x is ignored
        case 16: sin = sout = 0; scoef = 2; break;          // diag M0(h) (Phvec): h_x is ignored
        case 17: sin = sout = 2; break;                     // M2^-1 (WmatInv)
        case 18: sin = sout = 2; scoef = 2; break;          // M2(rho)^-1 (WhmatInv)
        case 19: sin = sout = 1; break;                     // diag M1 (MatGetDiagonal of the Umat shell): h_x is ignored
        case 20: sin = sout = 1; break;                     // element-block Jacobi of M1 (PCBJACOBI of the Umat shell)
        case 21: sin = sout = 1; scoef = 2; break;          // Umat_ray: h_coeff = Exner 2-form of the levels, h_u1 = its level-0 values (n2), tau = dt
        default: return fail(MIMSEM_ERR_ARG, "unknown operator id");
    }
    const int64_t nsp[3] = {c->n0, c->n1, c->n2};
    const int64_t nin = nsp[sin], nout = nsp[sout], ncoef = scoef >= 0 ? nsp[scoef] : 0;
    if (ncoef && !h_coeff) return fail(MIMSEM_ERR_ARG, "this operator needs a coefficient field");
    const bool need_u = (op == 8 || op == 9);
    if (need_u && !h_u1) return fail(MIMSEM_ERR_ARG, "this operator needs the advecting velocity");
    if (op == 21) {
        if (!h_u1) return fail(MIMSEM_ERR_ARG, "Umat_ray needs the level-0 Exner field");
        CUDA_OK(c->s_ray.resize((size_t)c->n2));
        CUDA_OK(cudaMemcpy(c->s_ray.p, h_u1, (size_t)c->n2 * sizeof(double), cudaMemcpyHostToDevice));
    }
    // Pipeline over chunks of levels (levels are independent).  A chunk goes host -> device, through the operator and back
    // on ONE stream; consecutive chunks use HOST_SLOTS streams (and buffer sets) in turn, so the upload of chunk i + 2 is
    // not held back by the download of chunk i that shares a stream with chunk i + HOST_SLOTS: the host -> device copy
    // engine runs back to back from the first chunk to the last and the device -> host engine trails it by one chunk.
    // (With two slots every stream alternated upload / download and each PCIe direction idled half of the time.)
    const int ld = nlev;
    constexpr int NS = mimsem_gpu_ctx::HOST_SLOTS;
    int CH = std::max(1, c->host_chunk);
    if (CH % 2) CH++;
    CH = std::min(CH, nlev);
    const int nchunk = (nlev + CH - 1) / CH;
    const int nslot = std::min(NS, nchunk);
    const size_t big = (size_t)std::max(std::max(std::max(nin, nout), ncoef), need_u ? (int64_t)c->n1 : (int64_t)0) * CH;
    for (int b = 0; b < nslot; b++) {
        CUDA_OK(c->s_lev2[b].resize(big));
        CUDA_OK(c->s_out2[b].resize((size_t)nout * CH));
        CUDA_OK(c->s_x2[b].resize((size_t)nin * CH));
        CUDA_OK(c->s_y2[b].resize((size_t)nout * CH));
        if (ncoef) CUDA_OK(c->s_c2[b].resize((size_t)ncoef * CH));
        if (need_u) CUDA_OK(c->s_u2[b].resize((size_t)c->n1 * CH));
    }
    cudaStream_t sts[NS];
    sts[0] = c->stream;
    for (int b = 1; b < nslot; b++) {
        if (!c->host_stream[b - 1]) CUDA_OK(cudaStreamCreateWithFlags(&c->host_stream[b - 1], cudaStreamNonBlocking));
        sts[b] = c->host_stream[b - 1];
    }
    for (int ck = 0; ck < nchunk; ck++) {
        const int b = ck % nslot;
        cudaStream_t st = sts[b];
        const int k0 = ck * CH, nl = std::min(CH, nlev - k0);
        // each chunk lives in its own column-layout buffers with leading dimension nl (contiguous DOF runs)
        if (ncoef) {
            CUDA_OK(cudaMemcpyAsync(c->s_lev2[b].p, h_coeff + (size_t)k0 * ncoef, (size_t)ncoef * nl * sizeof(double), cudaMemcpyHostToDevice, st));
            if ((rc = transpose(c, true, scoef, ncoef, nl, nl, c->s_lev2[b].p, c->s_c2[b].p, st))) return rc;
        }
        if (need_u) {
            CUDA_OK(cudaMemcpyAsync(c->s_lev2[b].p, h_u1 + (size_t)k0 * c->n1, (size_t)c->n1 * nl * sizeof(double), cudaMemcpyHostToDevice, st));
            if ((rc = transpose(c, true, 1, c->n1, nl, nl, c->s_lev2[b].p, c->s_u2[b].p, st))) return rc;
        }
        CUDA_OK(cudaMemcpyAsync(c->s_lev2[b].p, h_x + (size_t)k0 * nin, (size_t)nin * nl * sizeof(double), cudaMemcpyHostToDevice, st));
        if ((rc = transpose(c, true, sin, nin, nl, nl, c->s_lev2[b].p, c->s_x2[b].p, st))) return rc;
        // rows the operator does not write (halo rows in owner-computes mode) stay zero
        CUDA_OK(cudaMemsetAsync(c->s_y2[b].p, 0, (size_t)nout * nl * sizeof(double), st));
        const double* xc = c->s_x2[b].p;
        const double* cc = ncoef ? c->s_c2[b].p : nullptr;
        double* yc = c->s_y2[b].p;
        // MIMSEM_FIXED_LEVEL: every column of every chunk uses thickness level lev0 (box/Assembly.cpp:44-45)
        const int levk = (flags & MIMSEM_FIXED_LEVEL) ? lev0 : lev0 + k0;
        switch (op) {
            case 0: rc = apply_m1(c, false, levk, nl, nl, scale, tpow, flags, nullptr, xc, yc, st); break;
            case 3: rc = apply_m1(c, true, levk, nl, nl, scale, tpow, flags, cc, xc, yc, st); break;
            case 1: rc = apply_m2(c, false, levk, nl, nl, scale, tpow, flags, nullptr, xc, yc, st); break;
            case 5: rc = apply_m2(c, true, levk, nl, nl, scale, tpow, flags, cc, xc, yc, st); break;
            case 2: rc = apply_m0(c, false, levk, nl, nl, scale, tpow, flags, nullptr, xc, yc, st); break;
            case 6: rc = apply_m0(c, true, levk, nl, nl, scale, tpow, flags, cc, xc, yc, st); break;
            case 4: rc = apply_k(c, levk, nl, nl, scale, tpow, flags, cc, xc, yc, st); break;
            case 7: rc = apply_rot(c, false, levk, nl, nl, scale, tpow, flags, cc, nullptr, 0.0, xc, yc, st); break;
            case 8: rc = apply_rot(c, true, levk, nl, nl, scale, tpow, flags, cc, c->s_u2[b].p, tau, xc, yc, st); break;
            case 9: rc = apply_m0h_up(c, levk, nl, nl, scale, tpow, flags, cc, c->s_u2[b].p, tau, xc, yc, st); break;
            case 14: rc = apply_m1(c, true, 0, nl, nl, scale, 0, 0, xc, cc, yc, st); break;   // Uhmat(h2 := x2) u1, no thickness
            case 15: rc = apply_m0(c, false, levk, nl, nl, scale, tpow, flags, nullptr, nullptr, yc, st); break;
            case 16: rc = apply_m0(c, true, levk, nl, nl, scale, tpow, flags, cc, nullptr, yc, st); break;
            case 17: rc = solve_m2(c, false, levk, nl, nl, scale, tpow, flags, nullptr, xc, yc, st); break;
            case 18: rc = solve_m2(c, true, levk, nl, nl, scale, tpow, flags, cc, xc, yc, st); break;
            case 19: rc = diag_m1(c, false, levk, nl, nl, scale, tpow, flags, yc, st); break;
            case 20: rc = bjacobi_m1(c, levk, nl, nl, scale, tpow, flags, xc, yc, st); break;
            case 21: rc = apply_m1_ray(c, levk, nl, nl, scale, tau, cc, c->s_ray.p, xc, yc, st); break;
            default: rc = apply_inc(c, op - 10, nl, nl, xc, yc, st); break;
        }
        if (rc) return rc;
        if ((rc = transpose(c, false, sout, nout, nl, nl, yc, c->s_out2[b].p, st))) return rc;
        CUDA_OK(cudaMemcpyAsync(h_y + (size_t)k0 * nout, c->s_out2[b].p, (size_t)nout * nl * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    for (int b = nslot - 1; b >= 0; b--) CUDA_OK(cudaStreamSynchronize(sts[b]));
    return MIMSEM_OK;
}

int mimsem_gpu_gather_rows(mimsem_gpu_ctx* c, int64_t nrows, int nlev, int ld, const int* d_rows, const double* d_field,
                           double* d_packed, void* stream) {
    if (!c || !d_rows || !d_field || !d_packed) return fail(MIMSEM_ERR_ARG, "null argument");
    if (nrows == 0) return MIMSEM_OK;
    if (nrows < 0 || nlev < 1 || ld < nlev || nrows * nlev >= (1ll << 32)) return fail(MIMSEM_ERR_ARG, "bad pack shape");
    int rc = bind_device(c);
    if (rc) return rc;
    const FastDiv fd = make_fastdiv((unsigned)nlev);
    k_rows<true><<<grid_for(nrows * nlev, 256), 256, 0, (cudaStream_t)stream>>>(nrows, nlev, ld, fd.m, fd.s, d_rows, d_field, d_packed);
    return finish_launch(c, "gather_rows");
}

int mimsem_gpu_scatter_rows(mimsem_gpu_ctx* c, int64_t nrows, int nlev, int ld, const int* d_rows, const double* d_packed,
                            double* d_field, void* stream) {
    if (!c || !d_rows || !d_field || !d_packed) return fail(MIMSEM_ERR_ARG, "null argument");
    if (nrows == 0) return MIMSEM_OK;
    if (nrows < 0 || nlev < 1 || ld < nlev || nrows * nlev >= (1ll << 32)) return fail(MIMSEM_ERR_ARG, "bad unpack shape");
    int rc = bind_device(c);
    if (rc) return rc;
    const FastDiv fd = make_fastdiv((unsigned)nlev);
    k_rows<false><<<grid_for(nrows * nlev, 256), 256, 0, (cudaStream_t)stream>>>(nrows, nlev, ld, fd.m, fd.s, d_rows, d_packed, d_field);
    return finish_launch(c, "scatter_rows");
}

int mimsem_gpu_dev_alloc(mimsem_gpu_ctx* c, int64_t bytes, void** d_ptr) {
    if (!c || !d_ptr || bytes < 1) return fail(MIMSEM_ERR_ARG, "bad argument");
    int rc = bind_device(c);
    if (rc) return rc;
    CUDA_OK(cudaMalloc(d_ptr, (size_t)bytes));
    CUDA_OK(cudaMemset(*d_ptr, 0, (size_t)bytes));
    return MIMSEM_OK;
}
int mimsem_gpu_dev_free(mimsem_gpu_ctx* c, void* d_ptr) {
    if (!c || !d_ptr) return MIMSEM_OK;
    int rc = bind_device(c);
    if (rc) return rc;
    CUDA_OK(cudaFree(d_ptr));
    return MIMSEM_OK;
}
int mimsem_gpu_stream_create(mimsem_gpu_ctx* c, void** stream) {
    if (!c || !stream) return fail(MIMSEM_ERR_ARG, "null argument");
    int rc = bind_device(c);
    if (rc) return rc;
    cudaStream_t st;
    CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    *stream = (void*)st;
    return MIMSEM_OK;
}
int mimsem_gpu_stream_destroy(mimsem_gpu_ctx* c, void* stream) {
    if (!c || !stream) return MIMSEM_OK;
    int rc = bind_device(c);
    if (rc) return rc;
    CUDA_OK(cudaStreamDestroy((cudaStream_t)stream));
    return MIMSEM_OK;
}
int mimsem_gpu_graph_begin(mimsem_gpu_ctx* c, void* stream) {
    if (!c || !stream) return fail(MIMSEM_ERR_ARG, "capture needs a stream of its own (mimsem_gpu_stream_create)");
    int rc = bind_device(c);
    if (rc) return rc;
    CUDA_OK(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal));
    return MIMSEM_OK;
}
int mimsem_gpu_graph_end(mimsem_gpu_ctx* c, void* stream, void** graph_exec) {
    if (!c || !stream || !graph_exec) return fail(MIMSEM_ERR_ARG, "null argument");
    int rc = bind_device(c);
    if (rc) return rc;
    cudaGraph_t g = nullptr;
    CUDA_OK(cudaStreamEndCapture((cudaStream_t)stream, &g));
    cudaGraphExec_t ex = nullptr;
    cudaError_t e = cudaGraphInstantiate(&ex, g, 0);
    cudaGraphDestroy(g);
    CUDA_OK(e);
    *graph_exec = (void*)ex;
    return MIMSEM_OK;
}
int mimsem_gpu_graph_launch(mimsem_gpu_ctx* c, void* graph_exec, void* stream) {
    if (!c || !graph_exec) return fail(MIMSEM_ERR_ARG, "null argument");
    int rc = bind_device(c);
    if (rc) return rc;
    CUDA_OK(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream));
    return MIMSEM_OK;
}
int mimsem_gpu_graph_destroy(mimsem_gpu_ctx* c, void* graph_exec) {
    if (!c || !graph_exec) return MIMSEM_OK;
    int rc = bind_device(c);
    if (rc) return rc;
    CUDA_OK(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
    return MIMSEM_OK;
}
int mimsem_gpu_host_alloc(mimsem_gpu_ctx* c, int64_t bytes, void** h_ptr) {
    if (!c || !h_ptr || bytes < 1) return fail(MIMSEM_ERR_ARG, "bad argument");
    int rc = bind_device(c);
    if (rc) return rc;
    CUDA_OK(cudaMallocHost(h_ptr, (size_t)bytes));
    return MIMSEM_OK;
}
int mimsem_gpu_host_free(mimsem_gpu_ctx* c, void* h_ptr) {
    if (!c || !h_ptr) return MIMSEM_OK;
    int rc = bind_device(c);
    if (rc) return rc;
    CUDA_OK(cudaFreeHost(h_ptr));
    return MIMSEM_OK;
}
int mimsem_gpu_dev_copy(mimsem_gpu_ctx* c, void* dst, const void* src, int64_t bytes, int kind) {
    if (!c || !dst || !src || bytes < 0 || kind < 0 || kind > 2) return fail(MIMSEM_ERR_ARG, "bad argument");
    int rc = bind_device(c);
    if (rc) return rc;
    const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : (kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
    if (bytes > 0) CUDA_OK(cudaMemcpy(dst, src, (size_t)bytes, k));
    return MIMSEM_OK;
}
int mimsem_gpu_dev_sync(mimsem_gpu_ctx* c, void* st) {
    if (!c) return fail(MIMSEM_ERR_ARG, "null context");
    int rc = bind_device(c);
    if (rc) return rc;
    if (st) CUDA_OK(cudaStreamSynchronize((cudaStream_t)st));
    else CUDA_OK(cudaDeviceSynchronize());
    return MIMSEM_OK;
}

int mimsem_gpu_ipc_alloc(mimsem_gpu_ctx* c, int64_t bytes, void** d_ptr, unsigned char handle[64]) {
    if (!c || !d_ptr || !handle || bytes < 1) return fail(MIMSEM_ERR_ARG, "bad argument");
    int rc = bind_device(c);
    if (rc) return rc;
    CUDA_OK(cudaMalloc(d_ptr, (size_t)bytes));
    CUDA_OK(cudaMemset(*d_ptr, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    CUDA_OK(cudaIpcGetMemHandle(&h, *d_ptr));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(handle, &h, 64);
    return MIMSEM_OK;
}

int mimsem_gpu_ipc_open(mimsem_gpu_ctx* c, const unsigned char handle[64], void** d_ptr) {
    if (!c || !d_ptr || !handle) return fail(MIMSEM_ERR_ARG, "bad argument");
    int rc = bind_device(c);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    CUDA_OK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return MIMSEM_OK;
}

int mimsem_gpu_ipc_close(mimsem_gpu_ctx* c, void* d_ptr, int owned) {
    if (!c || !d_ptr) return MIMSEM_OK;
    cudaSetDevice(c->device);
    if (owned) cudaFree(d_ptr);
    else cudaIpcCloseMemHandle(d_ptr);
    return MIMSEM_OK;
}

int mimsem_gpu_halo_push(mimsem_gpu_ctx* c, int npeers, const void* d_peers, int nlev, int ld, int nbuf, const double* d_field,
                         void* d_epoch, int* d_err, void* stream) {
    if (!c || !d_epoch || (npeers > 0 && (!d_peers || !d_field || !d_err))) return fail(MIMSEM_ERR_ARG, "null argument");
    int rc = check_halo_levels(c, nlev, ld, nbuf);
    if (rc) return rc;
    if ((rc = bind_device(c))) return rc;
    if (npeers > 64) return fail(MIMSEM_ERR_ARG, "at most 64 halo peers");
    if (!c->d_halo_counters.p) {
        CUDA_OK(c->d_halo_counters.resize(256));
        CUDA_OK(cudaMemset(c->d_halo_counters.p, 0, 256 * sizeof(unsigned)));
    }
    if (npeers > 0) {
        k_halo<true><<<dim3(npeers, HALO_NB), 256, 0, (cudaStream_t)stream>>>((const HaloPeer*)d_peers, nlev, ld, nbuf,
                                                                             const_cast<double*>(d_field),
                                                                             (const unsigned long long*)d_epoch,
                                                                             c->d_halo_counters.p + (((uintptr_t)d_epoch >> 3) & 1) * 128, d_err);
        if ((rc = finish_launch(c, "halo_push"))) return rc;
    }
    k_epoch_inc<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)d_epoch);
    return finish_launch(c, "halo epoch");
}

int mimsem_gpu_halo_pull(mimsem_gpu_ctx* c, int npeers, const void* d_peers, int nlev, int ld, int nbuf, double* d_field, void* d_epoch,
                         int* d_err, void* stream) {
    if (!c || !d_epoch || (npeers > 0 && (!d_peers || !d_field || !d_err))) return fail(MIMSEM_ERR_ARG, "null argument");
    int rc = check_halo_levels(c, nlev, ld, nbuf);
    if (rc) return rc;
    if ((rc = bind_device(c))) return rc;
    if (npeers > 64) return fail(MIMSEM_ERR_ARG, "at most 64 halo peers");
    if (!c->d_halo_counters.p) {
        CUDA_OK(c->d_halo_counters.resize(256));
        CUDA_OK(cudaMemset(c->d_halo_counters.p, 0, 256 * sizeof(unsigned)));
    }
    if (npeers > 0) {
        k_halo<false><<<dim3(npeers, HALO_NB), 256, 0, (cudaStream_t)stream>>>((const HaloPeer*)d_peers, nlev, ld, nbuf, d_field,
                                                                              (const unsigned long long*)d_epoch,
                                                                              c->d_halo_counters.p + 64 + (((uintptr_t)d_epoch >> 3) & 1) * 128, d_err);
        if ((rc = finish_launch(c, "halo_pull"))) return rc;
    }
    k_epoch_inc<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)d_epoch);
    return finish_launch(c, "halo epoch");
}

int64_t mimsem_gpu_launch_count(const mimsem_gpu_ctx* c) { return c ? c->launches : 0; }

}  // extern "C"
