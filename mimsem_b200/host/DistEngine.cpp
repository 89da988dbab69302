// One rank of the element-partitioned engine (see DistEngine.h).  Set-up follows mimsem_b200/parallel.py
// (DistributedEngine.__init__, _setup_p2p) step by step; the data path is the C ABI of libmimsem_gpu.
#include "DistEngine.h"

#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace mimsem_host {

// ------------------------------------------------------------------------------------------------ file rendezvous
FileComm::FileComm(const std::string& dir_, int rank_, int world_) : dir(dir_), seq(0) {
    rank = rank_;
    world = world_;
    mkdir(dir.c_str(), 0777);
}

void FileComm::allgather(const void* send, int64_t bytes, void* recv) {
    char name[512], tmp[520];
    std::snprintf(name, sizeof(name), "%s/ag%ld.%d", dir.c_str(), seq, rank);
    std::snprintf(tmp, sizeof(tmp), "%s.tmp", name);
    FILE* f = std::fopen(tmp, "wb");
    if (!f || std::fwrite(send, 1, (size_t)bytes, f) != (size_t)bytes) throw std::runtime_error("FileComm: cannot write " + std::string(tmp));
    std::fclose(f);
    if (std::rename(tmp, name)) throw std::runtime_error("FileComm: rename failed");
    for (int q = 0; q < world; q++) {
        std::snprintf(name, sizeof(name), "%s/ag%ld.%d", dir.c_str(), seq, q);
        for (long spin = 0;; spin++) {
            struct stat st;
            if (stat(name, &st) == 0 && st.st_size == bytes) break;
            if (spin > 600000) throw std::runtime_error("FileComm: rank " + std::to_string(q) + " never arrived");
            usleep(100);
        }
        f = std::fopen(name, "rb");
        if (!f || std::fread((char*)recv + (size_t)q * bytes, 1, (size_t)bytes, f) != (size_t)bytes) throw std::runtime_error("FileComm: short read");
        std::fclose(f);
    }
    // everybody has passed round seq - 1 once it has written round seq: my file of that round can go
    if (seq > 0) {
        std::snprintf(name, sizeof(name), "%s/ag%ld.%d", dir.c_str(), seq - 1, rank);
        std::remove(name);
    }
    seq++;
}

void FileComm::barrier() {
    char c = 0;
    std::vector<char> all(world);
    allgather(&c, 1, all.data());
}

// ------------------------------------------------------------------------------------------------ the engine
namespace {

const int MAXW = 64;            // ranks of one box
struct Wire {                   // what a rank tells the others about its IPC buffer
    unsigned char handle[64];
    // plans 0, 1, 2 = the k-form spaces (rows the element kernels read); 3, 4 = ALL ghost rows of the 1- and 2-forms
    int64_t region_off[5], region_rows[5], region_stride[5];
    int64_t red_off;
    int layout_slot[5][MAXW], layout_row0[5][MAXW], layout_n[5][MAXW];   // my inbox share of peer q (slot -1: none)
    int send_slot[5][MAXW];                                              // position of q among the peers I send to
};
struct PeerDesc {               // HaloPeer of csrc/engine.cuh (48 bytes)
    const int* rows;
    int nrows, row0;
    double* inbox;
    long long stride;
    unsigned long long* signal;
    const unsigned long long* wait;
};

}  // namespace

void DistEngine::check(int rc, const char* what) {
    if (rc) throw std::runtime_error(std::string(what) + ": " + mimsem_last_error());
}

DistEngine::DistEngine(const GlobalMesh& mesh, const double* thick, int nk, Comm* comm, int device, int max_levels)
    : mesh_(mesh), comm_(comm), part_(NULL), ctx_(NULL), device_(device), base_(NULL), d_err_(NULL), stream_(NULL), d_red_areas_(NULL), d_red_seq_(NULL) {
    static_assert(sizeof(PeerDesc) == 48, "descriptor layout of csrc/engine.cuh");
    std::memset(plan_, 0, sizeof(plan_));
    const int rank = comm->rank, world = comm->world;
    if (world > MAXW) throw std::runtime_error("DistEngine: at most 64 ranks");
    part_ = new Partition(mesh, rank, world);
    part_->build_send_lists(mesh);
    const Partition& P = *part_;
    check(mimsem_gpu_create(device, &ctx_), "mimsem_gpu_create");
    const int p = mesh.p;
    {
        std::vector<double> x(p + 1), w(p + 1), lj((size_t)(p + 1) * (p + 1)), ej((size_t)(p + 1) * p);
        check(mimsem_basis_gll(p, x.data(), w.data()), "mimsem_basis_gll");
        check(mimsem_basis_tables(p, p, lj.data(), ej.data()), "mimsem_basis_tables");
        check(mimsem_gpu_set_basis(ctx_, p, p, w.data(), lj.data(), ej.data()), "mimsem_gpu_set_basis");
    }
    // sums over the elements around a node run in GLOBAL element order: bitwise equal to the single-GPU result
    std::vector<int> keys(P.elements.begin(), P.elements.end());
    check(mimsem_gpu_set_element_keys(ctx_, (int)keys.size(), keys.data()), "mimsem_gpu_set_element_keys");
    check(mimsem_gpu_set_topo(ctx_, P.nel_total, P.nel_owned, P.n0, P.n1, P.n2, P.nq, 0, P.el0.data(), P.el1x.data(), P.el1y.data(),
                              P.el2.data(), P.elq.data()),
          "mimsem_gpu_set_topo");
    check(mimsem_gpu_set_option(ctx_, "n0_owned", P.n0_owned), "n0_owned");
    int counts[2];
    check(mimsem_gpu_set_ghosts(ctx_, P.n1_owned, P.n2_owned, counts), "mimsem_gpu_set_ghosts");
    if (counts[0] != P.n_interior) throw std::runtime_error("DistEngine: partition and engine disagree on the interior / boundary split");
    {
        const int wq = mesh.nqe();
        std::vector<double> J((size_t)P.nel_total * wq * 4), det((size_t)P.nel_total * wq);
        for (int l = 0; l < P.nel_total; l++) {
            std::memcpy(&J[(size_t)l * wq * 4], &mesh.J[(size_t)P.elements[l] * wq * 4], sizeof(double) * wq * 4);
            std::memcpy(&det[(size_t)l * wq], &mesh.det[(size_t)P.elements[l] * wq], sizeof(double) * wq);
        }
        check(mimsem_gpu_set_geom(ctx_, J.data(), det.data()), "mimsem_gpu_set_geom");
    }
    if (thick) {
        std::vector<double> t((size_t)nk * P.nq);
        for (int k = 0; k < nk; k++)
            for (int i = 0; i < P.nq; i++) t[(size_t)k * P.nq + i] = thick[(size_t)k * mesh.NQ + P.gq[i]];
        check(mimsem_gpu_set_thickness(ctx_, nk, t.data()), "mimsem_gpu_set_thickness");
    }
    for (int s = 0; s < 3; s++) {
        perm_[s].resize(n_rows(s));
        check(mimsem_gpu_form_permutation(ctx_, s, perm_[s].data()), "mimsem_gpu_form_permutation");
    }
    nk_max_ = max_levels > 0 ? max_levels : (thick ? nk : 1);
    check(mimsem_gpu_set_option(ctx_, "halo_max_levels", nk_max_), "halo_max_levels");
    if (world > 1) setup_p2p();
}

void DistEngine::setup_p2p() {
    const Partition& P = *part_;
    const int rank = comm_->rank, world = comm_->world, nsp = 5, nk = nk_max_;
    auto space_of = [](int s) { return s < 3 ? s : s - 2; };
    auto recv_of = [&](int s) -> const std::map<int, GhostGroup>& { return s < 3 ? P.recv[s] : P.recv_ext[s - 2]; };
    auto send_of = [&](int s) -> const std::map<int, std::vector<int> >& { return s < 3 ? P.send[s] : P.send_ext[s - 2]; };
    Wire mine;
    std::memset(&mine, 0, sizeof(mine));
    for (int s = 0; s < nsp; s++)
        for (int q = 0; q < MAXW; q++) mine.layout_slot[s][q] = mine.send_slot[s][q] = -1;
    std::vector<int> recv_peers[5], send_peers[5];
    const int64_t hdr_bytes = (int64_t)(2 * nsp + 1) * MAXP * 8;   // flags[nsp][MAXP], acks[nsp][MAXP] (+ one spare row)
    int64_t off = hdr_bytes;
    for (int s = 0; s < nsp; s++) {
        for (std::map<int, GhostGroup>::const_iterator it = recv_of(s).begin(); it != recv_of(s).end(); ++it) recv_peers[s].push_back(it->first);
        for (std::map<int, std::vector<int> >::const_iterator it = send_of(s).begin(); it != send_of(s).end(); ++it) send_peers[s].push_back(it->first);
        if ((int)recv_peers[s].size() > MAXP || (int)send_peers[s].size() > MAXP) throw std::runtime_error("DistEngine: too many halo peers");
        // inbox of a space: [NBUF copies][ghost rows of the space, in ghost order][nk]; a peer's share is one run of rows
        int row = 0;
        for (size_t slot = 0; slot < recv_peers[s].size(); slot++) {
            const int q = recv_peers[s][slot];
            const std::vector<int>& loc = recv_of(s).find(q)->second.local;
            // (the fused M1 launch reads inbox row i as ghost row n_owned + i; the extended plans only go through the pull kernel)
            for (size_t i = 0; s < 3 && i < loc.size(); i++)
                if (loc[i] != P.n_owned(s) + row + (int)i) throw std::runtime_error("DistEngine: ghost rows of a peer must be one run");
            mine.layout_slot[s][q] = (int)slot;
            mine.layout_row0[s][q] = row;
            mine.layout_n[s][q] = (int)loc.size();
            row += (int)loc.size();
        }
        for (size_t i = 0; i < send_peers[s].size(); i++) mine.send_slot[s][send_peers[s][i]] = (int)i;
        const int64_t stride = (int64_t)row * nk + (((int64_t)row * nk) & 1);   // even: every copy starts 16-byte aligned
        mine.region_off[s] = off;
        mine.region_rows[s] = row;
        mine.region_stride[s] = stride;
        off += (int64_t)NBUF * stride * 8;
    }
    off = (off + 15) / 16 * 16;
    mine.red_off = off;   // reduction area of the partitioned CG: [2 parities][world][3 sums][64 levels] 16-byte cells
    off += (int64_t)2 * world * 3 * 64 * 16;
    void* base = NULL;
    check(mimsem_gpu_ipc_alloc(ctx_, std::max(off, hdr_bytes + 16), &base, mine.handle), "mimsem_gpu_ipc_alloc");
    base_ = (char*)base;
    std::vector<Wire> all(world);
    comm_->allgather(&mine, sizeof(Wire), all.data());
    peer_base_.assign(world, (char*)NULL);
    for (int q = 0; q < world; q++) {
        if (q == rank) continue;
        void* ptr = NULL;   // every peer is mapped: the CG reduction is all-to-all even where no ghost row is shared
        check(mimsem_gpu_ipc_open(ctx_, all[q].handle, &ptr), "mimsem_gpu_ipc_open");
        peer_base_[q] = (char*)ptr;
    }
    auto upload_rows = [&](const std::vector<int>& local, int space) {
        std::vector<int> rows(local.size());
        for (size_t i = 0; i < local.size(); i++) rows[i] = perm_[space][local[i]];
        void* d = NULL;
        check(mimsem_gpu_dev_alloc(ctx_, std::max<int64_t>(4, (int64_t)rows.size() * 4), &d), "dev_alloc");
        check(mimsem_gpu_dev_copy(ctx_, d, rows.data(), (int64_t)rows.size() * 4, 0), "dev_copy");
        keep_.push_back(d);
        return (const int*)d;
    };
    for (int s = 0; s < nsp; s++) {
        std::vector<PeerDesc> push(send_peers[s].size()), pull(recv_peers[s].size());
        int push_rows = 0;
        for (size_t i = 0; i < push.size(); i++) {
            const int q = send_peers[s][i];
            const std::vector<int>& loc = send_of(s).find(q)->second;
            const Wire& W = all[q];
            if (W.layout_slot[s][rank] < 0 || W.layout_n[s][rank] != (int)loc.size()) throw std::runtime_error("DistEngine: send / receive lists disagree");
            push[i].rows = upload_rows(loc, space_of(s));
            push[i].nrows = (int)loc.size();
            push[i].row0 = W.layout_row0[s][rank];
            push[i].inbox = (double*)(peer_base_[q] + W.region_off[s]);
            push[i].stride = W.region_stride[s];
            push[i].signal = (unsigned long long*)(peer_base_[q] + ((int64_t)s * MAXP + W.layout_slot[s][rank]) * 8);   // flag on q
            push[i].wait = (const unsigned long long*)(base_ + ((int64_t)(nsp + s) * MAXP + (int64_t)i) * 8);           // ack from q, in my memory
            push_rows += push[i].nrows;
        }
        for (size_t i = 0; i < pull.size(); i++) {
            const int q = recv_peers[s][i];
            const Wire& W = all[q];
            if (W.send_slot[s][rank] < 0) throw std::runtime_error("DistEngine: peer does not send what this rank receives");
            pull[i].rows = upload_rows(recv_of(s).find(q)->second.local, space_of(s));
            pull[i].nrows = mine.layout_n[s][q];
            pull[i].row0 = mine.layout_row0[s][q];
            pull[i].inbox = (double*)(base_ + mine.region_off[s]);
            pull[i].stride = mine.region_stride[s];
            pull[i].signal = (unsigned long long*)(peer_base_[q] + ((int64_t)(nsp + s) * MAXP + W.send_slot[s][rank]) * 8);   // ack on q
            pull[i].wait = (const unsigned long long*)(base_ + ((int64_t)s * MAXP + mine.layout_slot[s][q]) * 8);             // flag from q
        }
        Plan& pl = plan_[s];
        pl.npush = (int)push.size();
        pl.npull = (int)pull.size();
        check(mimsem_gpu_dev_alloc(ctx_, std::max<int64_t>(48, (int64_t)push.size() * 48), &pl.d_push), "dev_alloc");
        check(mimsem_gpu_dev_alloc(ctx_, std::max<int64_t>(48, (int64_t)pull.size() * 48), &pl.d_pull), "dev_alloc");
        check(mimsem_gpu_dev_copy(ctx_, pl.d_push, push.data(), (int64_t)push.size() * 48, 0), "dev_copy");
        check(mimsem_gpu_dev_copy(ctx_, pl.d_pull, pull.data(), (int64_t)pull.size() * 48, 0), "dev_copy");
        void* ep = NULL;
        check(mimsem_gpu_dev_alloc(ctx_, 16, &ep), "dev_alloc");
        pl.d_epochs = (unsigned long long*)ep;
        pl.inbox = base_ + mine.region_off[s];
        pl.stride = mine.region_stride[s];
        pl.push_rows = push_rows;
    }
    std::vector<unsigned long long> areas(world);
    for (int q = 0; q < world; q++) areas[q] = (unsigned long long)(uintptr_t)((q == rank ? base_ : peer_base_[q]) + all[q].red_off);
    check(mimsem_gpu_dev_alloc(ctx_, (int64_t)world * 8, &d_red_areas_), "dev_alloc");
    check(mimsem_gpu_dev_copy(ctx_, d_red_areas_, areas.data(), (int64_t)world * 8, 0), "dev_copy");
    check(mimsem_gpu_dev_alloc(ctx_, 8, &d_red_seq_), "dev_alloc");
    void* e = NULL;
    check(mimsem_gpu_dev_alloc(ctx_, 4, &e), "dev_alloc");
    d_err_ = (int*)e;
    sync();
    comm_->barrier();   // every inbox is allocated and zeroed before anybody pushes
}

DistEngine::~DistEngine() {
    if (ctx_) {
        mimsem_gpu_dev_sync(ctx_, NULL);
        if (stream_) mimsem_gpu_stream_destroy(ctx_, stream_);
        if (comm_ && comm_->world > 1) comm_->barrier();   // nobody unmaps a buffer a peer may still write
        for (size_t i = 0; i < keep_.size(); i++) mimsem_gpu_dev_free(ctx_, keep_[i]);
        for (int s = 0; s < 5; s++) {
            mimsem_gpu_dev_free(ctx_, plan_[s].d_push);
            mimsem_gpu_dev_free(ctx_, plan_[s].d_pull);
            mimsem_gpu_dev_free(ctx_, plan_[s].d_epochs);
        }
        mimsem_gpu_dev_free(ctx_, d_red_areas_);
        mimsem_gpu_dev_free(ctx_, d_red_seq_);
        mimsem_gpu_dev_free(ctx_, d_err_);
        for (size_t q = 0; q < peer_base_.size(); q++)
            if (peer_base_[q]) mimsem_gpu_ipc_close(ctx_, peer_base_[q], 0);
        if (base_) mimsem_gpu_ipc_close(ctx_, base_, 1);
        mimsem_gpu_destroy(ctx_);
    }
    delete part_;
}

void DistEngine::sync() { check(mimsem_gpu_dev_sync(ctx_, NULL), "sync"); }

bool DistEngine::halo_error() {
    if (!d_err_) return false;
    int e = 0;
    check(mimsem_gpu_dev_copy(ctx_, &e, d_err_, 4, 1), "dev_copy");
    return e != 0;
}

double* DistEngine::alloc_field(int space, int nlev) {
    void* d = NULL;
    check(mimsem_gpu_dev_alloc(ctx_, std::max<int64_t>(8, (int64_t)n_rows(space) * nlev * 8), &d), "dev_alloc");
    return (double*)d;
}
void DistEngine::free_field(double* d) { mimsem_gpu_dev_free(ctx_, d); }

void DistEngine::scatter_from_global(const double* levels, int space, int nlev, double* d_field) {
    const std::vector<int64_t>& g = part_->gids(space);
    const int64_t N = space == 0 ? mesh_.N0 : (space == 1 ? mesh_.N1 : mesh_.N2);
    std::vector<double> cols((size_t)g.size() * nlev);
    for (size_t i = 0; i < g.size(); i++)
        for (int k = 0; k < nlev; k++) cols[(size_t)perm_[space][i] * nlev + k] = levels[(size_t)k * N + g[i]];
    check(mimsem_gpu_dev_copy(ctx_, d_field, cols.data(), (int64_t)cols.size() * 8, 0), "dev_copy");
}

void DistEngine::owned_to_global(const double* d_field, int space, int nlev, double* levels) {
    const std::vector<int64_t>& g = part_->gids(space);
    const int64_t N = space == 0 ? mesh_.N0 : (space == 1 ? mesh_.N1 : mesh_.N2);
    std::vector<double> cols((size_t)g.size() * nlev);
    check(mimsem_gpu_dev_copy(ctx_, cols.data(), d_field, (int64_t)cols.size() * 8, 1), "dev_copy");
    for (int i = 0; i < part_->n_owned(space); i++)
        for (int k = 0; k < nlev; k++) levels[(size_t)k * N + g[i]] = cols[(size_t)perm_[space][i] * nlev + k];
}

void DistEngine::exchange(double* d_field, int space, int nlev, bool ext) {
    if (comm_->world == 1) return;
    if (nlev > nk_max_) throw std::runtime_error("DistEngine: more levels than the halo inboxes hold");
    if (ext && space == 0) ext = false;   // every ghost node is already in the plain plan
    const Plan& pl = plan_[ext ? space + 2 : space];
    check(mimsem_gpu_halo_push(ctx_, pl.npush, pl.d_push, nlev, nlev, NBUF, d_field, pl.d_epochs, d_err_, NULL), "mimsem_gpu_halo_push");
    check(mimsem_gpu_halo_pull(ctx_, pl.npull, pl.d_pull, nlev, nlev, NBUF, d_field, pl.d_epochs + 1, d_err_, NULL), "mimsem_gpu_halo_pull");
}

void DistEngine::apply_M1(const double* d_x, double* d_y, int nlev, double scale, int tpow, int lev0, int flags, void* stream) {
    if (comm_->world == 1) {
        check(mimsem_gpu_apply_M1(ctx_, lev0, nlev, nlev, scale, tpow, flags, d_x, d_y, stream), "mimsem_gpu_apply_M1");
        return;
    }
    const Plan& pl = plan_[1];
    if (nlev % 2 || nlev > 64 || nlev > nk_max_) {   // the fused launch needs even nlev <= 64: refresh, then apply
        if (stream) throw std::runtime_error("DistEngine::apply_M1: only the fused launch runs on a caller's stream");
        exchange(const_cast<double*>(d_x), 1, nlev);
        check(mimsem_gpu_apply_M1(ctx_, lev0, nlev, nlev, scale, tpow, flags, d_x, d_y, NULL), "mimsem_gpu_apply_M1");
        return;
    }
    const int push_ctas = std::max(1, std::min(148, pl.push_rows / 16));
    check(mimsem_gpu_apply_M1_halo(ctx_, lev0, nlev, nlev, scale, tpow, flags, d_x, d_y, d_x, 0, pl.npush, pl.d_push, pl.npull, pl.d_pull,
                                   (const double*)pl.inbox, pl.stride, NBUF, push_ctas, pl.d_epochs, d_err_, stream),
          "mimsem_gpu_apply_M1_halo");
}

struct DistEngine::Burst {
    void* graph;
};

DistEngine::Burst* DistEngine::capture_burst_M1(const std::vector<const double*>& xs, const std::vector<double*>& ys, int nsteps, int nlev,
                                                 double scale, int tpow, int lev0) {
    const int n = (int)xs.size();
    if (n < 1 || ys.size() != xs.size() || nsteps < 1 || nsteps > 32) throw std::runtime_error("DistEngine::capture_burst_M1: bad arguments");
    const bool fused = comm_->world > 1 && !(nlev % 2 || nlev > 64 || nlev > nk_max_);
    if (comm_->world > 1 && !fused) throw std::runtime_error("DistEngine::capture_burst_M1: needs the fused launch (even nlev <= 64)");
    if (!stream_) check(mimsem_gpu_stream_create(ctx_, &stream_), "mimsem_gpu_stream_create");
    check(mimsem_gpu_set_option(ctx_, "pdl_independent", 1), "pdl_independent");
    // warm-up outside the capture (first calls allocate counters; every pair is touched), in ordinary epochs
    sync();
    for (int i = 0; i < 2 * n; i++) apply_M1(xs[i % n], ys[i % n], nlev, scale, tpow, lev0, 0, stream_);
    check(mimsem_gpu_dev_sync(ctx_, stream_), "mimsem_gpu_dev_sync");
    if (comm_->world > 1) comm_->barrier();
    Burst* b = new Burst;
    b->graph = NULL;
    check(mimsem_gpu_graph_begin(ctx_, stream_), "mimsem_gpu_graph_begin");
    int rc = 0;
    try {
        for (int i = 0; i < nsteps; i++) {
            if (fused && nsteps > 1) {
                // launch i of the burst works on epoch *epoch + 1 + i with its own counter slot (include/mimsem_gpu.h, "Bursts")
                check(mimsem_gpu_set_option(ctx_, "halo_burst_len", nsteps), "halo_burst_len");
                check(mimsem_gpu_set_option(ctx_, "halo_burst_pos", i), "halo_burst_pos");
            }
            apply_M1(xs[i % n], ys[i % n], nlev, scale, tpow, lev0, 0, stream_);
        }
    } catch (...) {
        rc = 1;
    }
    mimsem_gpu_set_option(ctx_, "halo_burst_len", 0);
    mimsem_gpu_set_option(ctx_, "halo_burst_pos", 0);
    // the programmatic edges are part of the graph now; launches outside it are dependent unless the caller says otherwise
    mimsem_gpu_set_option(ctx_, "pdl_independent", 0);
    const int rc_end = mimsem_gpu_graph_end(ctx_, stream_, &b->graph);
    if (rc || rc_end) {
        delete b;
        throw std::runtime_error(std::string("DistEngine::capture_burst_M1: capture failed: ") + mimsem_last_error());
    }
    return b;
}

void DistEngine::replay(Burst* b) { check(mimsem_gpu_graph_launch(ctx_, b->graph, stream_), "mimsem_gpu_graph_launch"); }

void DistEngine::free_burst(Burst* b) {
    if (!b) return;
    mimsem_gpu_dev_sync(ctx_, NULL);
    mimsem_gpu_graph_destroy(ctx_, b->graph);
    delete b;
}

void DistEngine::apply(const std::string& op, double* x, double* c, double* y, int nlev, double scale, int tpow, int lev0, double* u1,
                       double tau) {
    const int ld = nlev;
    if (op == "M1") {
        apply_M1(x, y, nlev, scale, tpow, lev0);
    } else if (op == "M1h") {
        exchange(x, 1, nlev);
        exchange(c, 2, nlev);
        check(mimsem_gpu_apply_M1h(ctx_, lev0, nlev, ld, scale, tpow, 0, c, x, y, NULL), "mimsem_gpu_apply_M1h");
    } else if (op == "K") {
        exchange(x, 1, nlev);
        exchange(c, 1, nlev);
        check(mimsem_gpu_apply_K(ctx_, lev0, nlev, ld, scale, tpow, 0, c, x, y, NULL), "mimsem_gpu_apply_K");
    } else if (op == "M2") {
        check(mimsem_gpu_apply_M2(ctx_, lev0, nlev, ld, scale, tpow, 0, x, y, NULL), "mimsem_gpu_apply_M2");
    } else if (op == "M2h") {
        check(mimsem_gpu_apply_M2h(ctx_, lev0, nlev, ld, scale, tpow, 0, c, x, y, NULL), "mimsem_gpu_apply_M2h");
    } else if (op == "M0") {
        check(mimsem_gpu_apply_M0(ctx_, lev0, nlev, ld, scale, tpow, 0, x, y, NULL), "mimsem_gpu_apply_M0");
    } else if (op == "E21" || op == "E12" || op == "E10") {
        exchange(x, op == "E21" ? 1 : (op == "E12" ? 2 : 0), nlev);
        const int which = op == "E21" ? MIMSEM_E21 : (op == "E12" ? MIMSEM_E12 : MIMSEM_E10);
        check(mimsem_gpu_apply_incidence(ctx_, which, nlev, ld, x, y, NULL), "mimsem_gpu_apply_incidence");
    } else if (op == "R") {
        exchange(x, 1, nlev);
        exchange(c, 0, nlev);
        check(mimsem_gpu_apply_R(ctx_, lev0, nlev, ld, scale, tpow, 0, c, x, y, NULL), "mimsem_gpu_apply_R");
    } else if (op == "R_up") {
        exchange(x, 1, nlev);
        exchange(c, 0, nlev);
        exchange(u1, 1, nlev);
        check(mimsem_gpu_apply_R_up(ctx_, lev0, nlev, ld, scale, tpow, 0, c, u1, tau, x, y, NULL), "mimsem_gpu_apply_R_up");
    } else if (op == "UtQW") {
        exchange(x, 2, nlev);
        exchange(c, 1, nlev);
        check(mimsem_gpu_apply_UtQW(ctx_, nlev, ld, scale, c, x, y, NULL), "mimsem_gpu_apply_UtQW");
    } else if (op == "M0h") {   // sums over the elements around a node: every ghost face of the coefficient
        exchange(c, 2, nlev, true);
        check(mimsem_gpu_apply_M0h(ctx_, lev0, nlev, ld, scale, tpow, 0, c, x, y, NULL), "mimsem_gpu_apply_M0h");
    } else if (op == "E01") {
        exchange(x, 1, nlev, true);
        check(mimsem_gpu_apply_incidence(ctx_, MIMSEM_E01, nlev, ld, x, y, NULL), "mimsem_gpu_apply_incidence");
    } else if (op == "M0h_up") {
        exchange(x, 0, nlev);
        exchange(c, 2, nlev, true);
        exchange(u1, 1, nlev, true);
        check(mimsem_gpu_apply_M0h_up(ctx_, lev0, nlev, ld, scale, tpow, 0, c, u1, tau, x, y, NULL), "mimsem_gpu_apply_M0h_up");
    } else {
        throw std::runtime_error("DistEngine::apply: operator " + op + " is not available in the C++ host layer");
    }
}

int DistEngine::solve_M1(const double* b, double* x, int nlev, double scale, int tpow, double rtol, int maxit, double* relres) {
    int iters = 0;
    double rr = 0.0;
    if (comm_->world == 1) {
        check(mimsem_gpu_solve_M1(ctx_, 0, nlev, nlev, scale, tpow, 0, b, x, rtol, maxit, &iters, &rr, NULL), "mimsem_gpu_solve_M1");
    } else {
        const Plan& pl = plan_[1];
        mimsem_halo_desc hd;
        std::memset(&hd, 0, sizeof(hd));
        hd.npush = pl.npush; hd.d_push = pl.d_push; hd.npull = pl.npull; hd.d_pull = pl.d_pull;
        hd.d_inbox = pl.inbox; hd.stride = pl.stride; hd.nbuf = NBUF;
        hd.push_ctas = std::max(1, std::min(148, pl.push_rows / 16));
        hd.d_epoch = pl.d_epochs; hd.d_err = d_err_; hd.ll = 0;
        mimsem_reduce_desc rd;
        rd.world = comm_->world; rd.rank = comm_->rank; rd.d_peer_areas = d_red_areas_; rd.d_seq = d_red_seq_; rd.d_err = d_err_;
        check(mimsem_gpu_solve_M1_dist(ctx_, 0, nlev, nlev, scale, tpow, 0, b, x, rtol, maxit, &iters, &rr, &hd, &rd, NULL), "mimsem_gpu_solve_M1_dist");
    }
    if (relres) *relres = rr;
    return iters;
}

}  // namespace mimsem_host
