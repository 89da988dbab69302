// Element-block partition (see Partition.h).  Construction identical to mimsem_b200/parallel.py.
#include "Partition.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>

#include "mimsem_gpu.h"

namespace mimsem_host {

namespace {

typedef std::vector<int64_t> IVec;

IVec unique_sorted(IVec v) {
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
    return v;
}
IVec set_diff(const IVec& a, const IVec& b) {   // both sorted unique
    IVec out;
    std::set_difference(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(out));
    return out;
}
IVec set_union(const IVec& a, const IVec& b) {
    IVec out;
    std::set_union(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(out));
    return out;
}
IVec set_inter(const IVec& a, const IVec& b) {
    IVec out;
    std::set_intersection(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(out));
    return out;
}
void append_row(IVec& out, const std::vector<int>& table, int64_t e, int width) {
    for (int i = 0; i < width; i++) out.push_back(table[(size_t)e * width + i]);
}

// local numbering of a set of global ids: owned (id / block in [e0, e1)) first, then the ghosts -- those some kernel
// reads (`needed`) before the others; block == 0: everything is "owned" (quadrature points)
void local_numbering(const IVec& ids_in, int block, int64_t e0, int64_t e1, const IVec* needed, IVec* out, int* n_own, int* n_need) {
    const IVec ids = unique_sorted(ids_in);
    IVec own, ghost;
    for (size_t i = 0; i < ids.size(); i++) {
        const bool mine = block == 0 || (ids[i] / block >= e0 && ids[i] / block < e1);
        (mine ? own : ghost).push_back(ids[i]);
    }
    *n_own = (int)own.size();
    *out = own;
    if (needed) {
        const IVec need = set_inter(ghost, *needed);
        const IVec rest = set_diff(ghost, need);
        out->insert(out->end(), need.begin(), need.end());
        out->insert(out->end(), rest.begin(), rest.end());
        *n_need = (int)need.size();
    } else {
        out->insert(out->end(), ghost.begin(), ghost.end());
        *n_need = (int)ghost.size();
    }
}

// table of global ids -> local rows (every id must occur in gids)
std::vector<int> to_local(const IVec& gids, const std::vector<int>& table, const IVec& elements, int width) {
    std::vector<int> order(gids.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return gids[a] < gids[b]; });
    IVec sorted(gids.size());
    for (size_t i = 0; i < gids.size(); i++) sorted[i] = gids[order[i]];
    std::vector<int> out(elements.size() * (size_t)width);
    for (size_t l = 0; l < elements.size(); l++)
        for (int i = 0; i < width; i++) {
            const int64_t g = table[(size_t)elements[l] * width + i];
            const size_t pos = std::lower_bound(sorted.begin(), sorted.end(), g) - sorted.begin();
            if (pos >= sorted.size() || sorted[pos] != g) {
                std::fprintf(stderr, "mimsem_host::Partition: id %lld missing from the local numbering\n", (long long)g);
                std::abort();
            }
            out[l * width + i] = order[pos];
        }
    return out;
}

}  // namespace

void element_range(int64_t nel, int rank, int world, int64_t* e0, int64_t* e1) {
    *e0 = ((int64_t)rank * nel) / world;
    *e1 = ((int64_t)(rank + 1) * nel) / world;
}

int owner_rank_of_element(int64_t e, int64_t nel, int world) {
    int64_t r = (e * world) / nel;
    const int64_t lo = (r * nel) / world, hi = ((r + 1) * nel) / world;
    if (e < lo) r--;
    else if (e >= hi) r++;
    return (int)r;
}

int GlobalMesh::create(int kind_, int p_, int ne_, bool signed_det, bool with_geometry) {
    mimsem_mesh* M = NULL;
    int rc = mimsem_mesh_create(kind_, p_, p_, ne_, signed_det ? 1 : 0, &M);
    if (rc) return rc;
    int64_t sz[8];
    mimsem_mesh_sizes(M, sz);
    kind = kind_;
    p = (int)sz[0]; m = (int)sz[1]; ne = (int)sz[2]; nel = (int)sz[3];
    N0 = sz[4]; N1 = sz[5]; N2 = sz[6]; NQ = sz[7];
    el0.resize((size_t)nel * n0e()); el1x.resize((size_t)nel * n1e()); el1y.resize((size_t)nel * n1e());
    el2.resize((size_t)nel * n2e()); elq.resize((size_t)nel * nqe());
    rc = mimsem_mesh_tables(M, el0.data(), el1x.data(), el1y.data(), el2.data(), elq.data());
    if (!rc && with_geometry) {
        J.resize((size_t)nel * nqe() * 4);
        det.resize((size_t)nel * nqe());
        rc = mimsem_mesh_geometry(M, J.data(), det.data());
        xyz.resize((size_t)NQ * 3);
        if (!rc) rc = mimsem_mesh_coords(M, xyz.data());
    }
    mimsem_mesh_destroy(M);
    if (rc) return rc;
    // west / south neighbours: the other element that uses my first x-normal / y-normal edge
    const int w1 = n1e();
    std::vector<std::pair<int64_t, int64_t> > uses;   // (edge, element), el1x uses first, then el1y (stable sort keeps that)
    uses.reserve((size_t)nel * w1 * 2);
    for (int64_t e = 0; e < nel; e++)
        for (int i = 0; i < w1; i++) uses.push_back(std::make_pair((int64_t)el1x[(size_t)e * w1 + i], e));
    for (int64_t e = 0; e < nel; e++)
        for (int i = 0; i < w1; i++) uses.push_back(std::make_pair((int64_t)el1y[(size_t)e * w1 + i], e));
    std::stable_sort(uses.begin(), uses.end(), [](const std::pair<int64_t, int64_t>& a, const std::pair<int64_t, int64_t>& b) { return a.first < b.first; });
    std::vector<int64_t> first(N1 + 1, 0);
    {
        size_t i = 0;
        for (int64_t d = 0; d <= N1; d++) {
            while (i < uses.size() && uses[i].first < d) i++;
            first[d] = (int64_t)i;
        }
    }
    ws_nbr.assign((size_t)nel * 2, -1);
    for (int64_t e = 0; e < nel; e++)
        for (int side = 0; side < 2; side++) {
            const int64_t d = side == 0 ? el1x[(size_t)e * w1] : el1y[(size_t)e * w1];
            const int64_t f = first[d], l = first[d + 1];
            if (l - f < 2) continue;
            const int64_t a = uses[f].second, b = uses[l - 1].second;
            ws_nbr[(size_t)e * 2 + side] = a == e ? b : a;   // a one-element-wide periodic mesh: its own neighbour
        }
    // nodes -> elements
    const int w0 = n0e();
    node_min_el.assign(N0, nel);
    std::vector<int64_t> cnt(N0 + 1, 0);
    for (int64_t e = 0; e < nel; e++)
        for (int i = 0; i < w0; i++) {
            const int64_t n = el0[(size_t)e * w0 + i];
            node_min_el[n] = std::min(node_min_el[n], e);
            cnt[n + 1]++;
        }
    node_ptr.assign(N0 + 1, 0);
    for (int64_t n = 0; n < N0; n++) node_ptr[n + 1] = node_ptr[n] + cnt[n + 1];
    node_els.resize(node_ptr[N0]);
    std::vector<int64_t> fill(node_ptr.begin(), node_ptr.end() - 1);
    for (int64_t e = 0; e < nel; e++)
        for (int i = 0; i < w0; i++) node_els[fill[el0[(size_t)e * w0 + i]]++] = e;
    return 0;
}

Partition::Partition(const GlobalMesh& mesh, int rank_, int world_) {
    p = mesh.p; rank = rank_; world = world_;
    const int64_t nel = mesh.nel;
    element_range(nel, rank, world, &e0, &e1);
    const int w1 = mesh.n1e(), w2 = mesh.n2e(), w0 = mesh.n0e(), wq = mesh.nqe();
    const int b1 = 2 * p * p, b2 = p * p;
    IVec owned;
    for (int64_t e = e0; e < e1; e++) owned.push_back(e);
    // west / south halo elements
    IVec nb;
    for (int64_t e = e0; e < e1; e++)
        for (int s = 0; s < 2; s++)
            if (mesh.ws_nbr[(size_t)e * 2 + s] >= 0) nb.push_back(mesh.ws_nbr[(size_t)e * 2 + s]);
    halo_ws = set_diff(unique_sorted(nb), owned);
    // every element around an owned node (node-sum operators)
    IVec around;
    for (int64_t n = 0; n < mesh.N0; n++)
        if (mesh.node_min_el[n] >= e0 && mesh.node_min_el[n] < e1)
            for (int64_t i = mesh.node_ptr[n]; i < mesh.node_ptr[n + 1]; i++) around.push_back(mesh.node_els[i]);
    const IVec halo_node = set_diff(unique_sorted(around), owned);
    const IVec halo = set_union(halo_ws, halo_node);
    // interior (reads no row owned elsewhere, directly or through a west / south neighbour) first, boundary last
    auto foreign = [&](int64_t e) {
        for (int i = 0; i < w1; i++) {
            const int64_t dx = mesh.el1x[(size_t)e * w1 + i] / b1, dy = mesh.el1y[(size_t)e * w1 + i] / b1;
            if (dx < e0 || dx >= e1 || dy < e0 || dy >= e1) return true;
        }
        for (int i = 0; i < w2; i++) {
            const int64_t d = mesh.el2[(size_t)e * w2 + i] / b2;
            if (d < e0 || d >= e1) return true;
        }
        return false;
    };
    IVec inner, bnd;
    for (int64_t e = e0; e < e1; e++) {
        bool b = foreign(e);
        for (int s = 0; s < 2 && !b; s++) {
            const int64_t n = mesh.ws_nbr[(size_t)e * 2 + s];
            if (n >= 0 && foreign(n)) b = true;
        }
        (b ? bnd : inner).push_back(e);
    }
    n_interior = (int)inner.size();
    elements = inner;
    elements.insert(elements.end(), bnd.begin(), bnd.end());
    nel_owned = (int)elements.size();
    elements.insert(elements.end(), halo.begin(), halo.end());
    nel_total = (int)elements.size();
    const IVec& L = elements;
    // 1-form rows the kernels read: every edge of an owned element and, of a west / south halo element, the edge family
    // ACROSS its far line
    IVec need;
    for (int64_t e = e0; e < e1; e++) {
        append_row(need, mesh.el1x, e, w1);
        append_row(need, mesh.el1y, e, w1);
    }
    for (int side = 0; side < 2; side++)
        for (int64_t e = e0; e < e1; e++) {
            const int64_t n = mesh.ws_nbr[(size_t)e * 2 + side];
            if (n < 0) continue;
            const int64_t sh = side == 0 ? mesh.el1x[(size_t)e * w1] : mesh.el1y[(size_t)e * w1];
            bool far_is_east = false;
            for (int i = 0; i < w1; i++)
                if (mesh.el1x[(size_t)n * w1 + i] == sh) far_is_east = true;
            append_row(need, far_is_east ? mesh.el1y : mesh.el1x, n, w1);
        }
    const IVec needed1 = unique_sorted(need);
    IVec ids;
    for (size_t l = 0; l < L.size(); l++) append_row(ids, mesh.el1x, L[l], w1);
    for (size_t l = 0; l < L.size(); l++) append_row(ids, mesh.el1y, L[l], w1);
    int n_need = 0;
    local_numbering(ids, b1, e0, e1, &needed1, &g1, &n1_owned, &n_need);
    n1_halo = n1_owned + n_need;
    // 2-forms the element kernels read: the faces of the owned and the west / south halo elements
    IVec need2;
    for (int l = 0; l < nel_owned; l++) append_row(need2, mesh.el2, L[l], w2);
    for (size_t i = 0; i < halo_ws.size(); i++) append_row(need2, mesh.el2, halo_ws[i], w2);
    const IVec needed2 = unique_sorted(need2);
    ids.clear();
    for (size_t l = 0; l < L.size(); l++) append_row(ids, mesh.el2, L[l], w2);
    local_numbering(ids, b2, e0, e1, &needed2, &g2, &n2_owned, &n_need);
    n2_halo = n2_owned + n_need;
    // nodes: owned first (ascending), then the ghosts grouped by owner rank
    ids.clear();
    for (size_t l = 0; l < L.size(); l++) append_row(ids, mesh.el0, L[l], w0);
    const IVec ids0 = unique_sorted(ids);
    IVec own0, gh0;
    for (size_t i = 0; i < ids0.size(); i++) {
        const int64_t me = mesh.node_min_el[ids0[i]];
        (me >= e0 && me < e1 ? own0 : gh0).push_back(ids0[i]);
    }
    auto node_owner = [&](int64_t n) { return owner_rank_of_element(mesh.node_min_el[n], nel, world); };
    std::stable_sort(gh0.begin(), gh0.end(), [&](int64_t a, int64_t b) {
        const int oa = node_owner(a), ob = node_owner(b);
        return oa != ob ? oa < ob : a < b;
    });
    g0 = own0;
    g0.insert(g0.end(), gh0.begin(), gh0.end());
    n0_owned = (int)own0.size();
    ids.clear();
    for (size_t l = 0; l < L.size(); l++) append_row(ids, mesh.elq, L[l], wq);
    int nq_own = 0;
    local_numbering(ids, 0, e0, e1, NULL, &gq, &nq_own, &n_need);
    n0 = (int)g0.size(); n1 = (int)g1.size(); n2 = (int)g2.size(); nq = (int)gq.size();
    el1x = to_local(g1, mesh.el1x, L, w1);
    el1y = to_local(g1, mesh.el1y, L, w1);
    el2 = to_local(g2, mesh.el2, L, w2);
    el0 = to_local(g0, mesh.el0, L, w0);
    elq = to_local(gq, mesh.elq, L, wq);
    // ghosts grouped by owner rank, in ghost order
    auto group = [&](const IVec& g, int from, int to, int block, std::map<int, GhostGroup>& out) {
        for (int i = from; i < to; i++) {
            const int q = block ? owner_rank_of_element(g[i] / block, nel, world) : node_owner(g[i]);
            GhostGroup& gg = out[q];
            gg.local.push_back(i);
            gg.glob.push_back(g[i]);
        }
    };
    group(g0, n0_owned, n0, 0, recv[0]);
    group(g1, n1_owned, n1_halo, b1, recv[1]);
    group(g2, n2_owned, n2_halo, b2, recv[2]);
    group(g1, n1_owned, n1, b1, recv_ext[1]);
    group(g2, n2_owned, n2, b2, recv_ext[2]);
}

void Partition::build_send_lists(const GlobalMesh& mesh) {
    for (int q = 0; q < world; q++) {
        if (q == rank) continue;
        const Partition other(mesh, q, world);
        for (int space = 0; space < 3; space++) {
            const IVec& g = gids(space);
            const int nown = n_owned(space);   // owned ids are ascending: local id of an owned DOF = its position
            for (int ext = 0; ext < 2; ext++) {
                if (ext && space == 0) continue;
                const std::map<int, GhostGroup>& table = ext ? other.recv_ext[space] : other.recv[space];
                std::map<int, GhostGroup>::const_iterator it = table.find(rank);
                if (it == table.end()) continue;
                std::vector<int>& dst = (ext ? send_ext : send)[space][q];
                for (size_t i = 0; i < it->second.glob.size(); i++) {
                    const int64_t gid = it->second.glob[i];
                    const size_t pos = std::lower_bound(g.begin(), g.begin() + nown, gid) - g.begin();
                    if (pos >= (size_t)nown || g[pos] != gid) {
                        std::fprintf(stderr, "mimsem_host::Partition: rank %d asks rank %d for id %lld, which it does not own\n", q, rank, (long long)gid);
                        std::abort();
                    }
                    dst.push_back((int)pos);
                }
            }
        }
    }
}

}  // namespace mimsem_host

// ------------------------------------------------------------------------------------------------------------------
struct mimsem_host_partition {
    mimsem_host::GlobalMesh mesh;
    mimsem_host::Partition* part;
};

extern "C" {

int mimsem_host_partition_create(int kind, int p, int ne, int rank, int world, mimsem_host_partition** out) {
    if (!out || world < 1 || rank < 0 || rank >= world) return MIMSEM_ERR_ARG;
    mimsem_host_partition* h = new mimsem_host_partition;
    const int rc = h->mesh.create(kind, p, ne, false, false);
    if (rc) {
        delete h;
        return rc;
    }
    h->part = new mimsem_host::Partition(h->mesh, rank, world);
    h->part->build_send_lists(h->mesh);
    *out = h;
    return 0;
}

void mimsem_host_partition_destroy(mimsem_host_partition* h) {
    if (!h) return;
    delete h->part;
    delete h;
}

int mimsem_host_partition_sizes(const mimsem_host_partition* h, int64_t s[12]) {
    if (!h || !s) return MIMSEM_ERR_ARG;
    const mimsem_host::Partition& P = *h->part;
    const int64_t v[12] = {P.nel_owned, P.nel_total, P.n_interior, P.n0, P.n1, P.n2, P.nq, P.n0_owned, P.n1_owned, P.n2_owned, P.n1_halo, P.n2_halo};
    std::memcpy(s, v, sizeof(v));
    return 0;
}

int64_t mimsem_host_partition_array(const mimsem_host_partition* h, const char* name, int64_t* out) {
    if (!h || !name) return -1;
    const mimsem_host::Partition& P = *h->part;
    const std::string n(name);
    std::vector<int64_t> tmp;
    const std::vector<int64_t>* src64 = NULL;
    const std::vector<int>* src32 = NULL;
    if (n == "elements") src64 = &P.elements;
    else if (n == "g0") src64 = &P.g0;
    else if (n == "g1") src64 = &P.g1;
    else if (n == "g2") src64 = &P.g2;
    else if (n == "gq") src64 = &P.gq;
    else if (n == "el0") src32 = &P.el0;
    else if (n == "el1x") src32 = &P.el1x;
    else if (n == "el1y") src32 = &P.el1y;
    else if (n == "el2") src32 = &P.el2;
    else if (n == "elq") src32 = &P.elq;
    else {
        int space = -1, peer = -1;
        if (std::sscanf(name, "recv%d_%d", &space, &peer) == 2 && space >= 0 && space < 3) {
            std::map<int, mimsem_host::GhostGroup>::const_iterator it = P.recv[space].find(peer);
            if (it == P.recv[space].end()) return 0;
            src32 = &it->second.local;
        } else if (std::sscanf(name, "send%d_%d", &space, &peer) == 2 && space >= 0 && space < 3) {
            std::map<int, std::vector<int> >::const_iterator it = P.send[space].find(peer);
            if (it == P.send[space].end()) return 0;
            src32 = &it->second;
        } else {
            return -1;
        }
    }
    if (src64) {
        if (out) std::memcpy(out, src64->data(), src64->size() * sizeof(int64_t));
        return (int64_t)src64->size();
    }
    if (out)
        for (size_t i = 0; i < src32->size(); i++) out[i] = (*src32)[i];
    return (int64_t)src32->size();
}
}
