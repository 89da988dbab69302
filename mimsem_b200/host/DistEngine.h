// One rank of the element-partitioned engine on the GPUs of one box (C++ host layer; what mimsem_b200/parallel.py's
// DistributedEngine is for the Python harness): builds the rank's subdomain (Partition.h), configures a
// mimsem_gpu context for it, allocates the ghost inboxes, exchanges the IPC handles and layouts through a caller-supplied
// allgather (MPI_Allgather in the reference's world; a file rendezvous for plain processes of one box is included), maps
// the peers' inboxes and builds the push / pull descriptor arrays.  After that every call is one or two library
// launches: there is no NCCL and no MPI on the data path.
//
//   replaces, per operator apply of the reference:
//     VecScatterBegin/End(gtol_k, x, xl, INSERT_VALUES, SCATTER_FORWARD)           eul/Euler_2.cpp:1455-1456
//     Op->assemble(...); MatMult(Op->M, ...)                                        eul/Assembly.cpp
//     VecScatterBegin/End(gtol_k, yl, y, ADD_VALUES, SCATTER_REVERSE)               eul/Assembly.cpp:2194-2195
#pragma once
#include <string>
#include <vector>

#include "Partition.h"
#include "mimsem_gpu.h"

namespace mimsem_host {

// control-plane transport: every rank contributes `bytes` bytes, receives world * bytes (rank order); collective
struct Comm {
    int rank, world;
    virtual void allgather(const void* send, int64_t bytes, void* recv) = 0;
    virtual void barrier() = 0;
    virtual ~Comm() {}
};

// processes of one box without MPI: rank r writes <dir>/<seq>.<r> and polls for the others' files
struct FileComm : Comm {
    std::string dir;
    long seq;
    FileComm(const std::string& dir, int rank, int world);
    void allgather(const void* send, int64_t bytes, void* recv);
    void barrier();
};

class DistEngine {
public:
    // thick: [nk][NQ] of the GLOBAL mesh (level-major as Geom::thick) or NULL; max_levels: levels per ghost row the
    // inboxes hold (0: nk, or 1 without thickness)
    DistEngine(const GlobalMesh& mesh, const double* thick, int nk, Comm* comm, int device, int max_levels = 0);
    ~DistEngine();

    mimsem_gpu_ctx* ctx() const { return ctx_; }
    const Partition& part() const { return *part_; }
    int n_rows(int space) const { return space == 0 ? part_->n0 : (space == 1 ? part_->n1 : part_->n2); }

    // local device field of a space in the engine's column layout, [n_rows(space)][nlev] doubles, zero-filled
    double* alloc_field(int space, int nlev);
    void free_field(double* d);
    // global per-level host field levels[k * N_space + gid] -> local device field (owned AND ghost rows), and the owned
    // rows of a local device field back into a global per-level host array (other entries untouched)
    void scatter_from_global(const double* levels_global, int space, int nlev, double* d_field);
    void owned_to_global(const double* d_field, int space, int nlev, double* levels_global);

    // ghost refresh of a local field (push kernel into the peers' inboxes + pull kernel), collective.  ext: ALL ghost rows of
    // the space (1- and 2-forms), what the operators that sum over the elements around a node read (M0h, E01, M0h_up)
    void exchange(double* d_field, int space, int nlev, bool ext = false);
    // y = M1 x with the ghost refresh of x fused into the launch (mimsem_gpu_apply_M1_halo), collective
    void apply_M1(const double* d_x, double* d_y, int nlev, double scale, int tpow, int lev0 = 0, int flags = 0, void* stream = NULL);
    // A burst: nsteps consecutive M1 applies, step i on field pair i % n (n = xs.size(); the pairs must be independent
    // fields), captured in ONE CUDA graph.  The launches carry programmatic dependencies and work on consecutive epochs of
    // the ghost hand-over, so that a step starts while the previous one drains and the boundary rows of step i + 1 travel
    // while step i computes (engine option "pdl_independent" is on during the capture only: the edges live in the graph).
    // Collective: every rank captures and replays the same bursts in the same order.  replay() only enqueues, on the engine's
    // own non-blocking stream: call sync() before the fields are used by any other call of this class (they run on the
    // default stream) and before the host reads them.
    struct Burst;
    Burst* capture_burst_M1(const std::vector<const double*>& xs, const std::vector<double*>& ys, int nsteps, int nlev, double scale,
                            int tpow, int lev0 = 0);
    void replay(Burst* burst);
    void free_burst(Burst* burst);
    // the other operators of the path: ghost refresh of the inputs that need one, then the local apply.
    // op: "M1", "M1h", "M2", "M2h", "K", "UtQW", "E21", "E12", "M0", "M0h", "E10", "E01", "R", "R_up", "M0h_up" -- all fifteen
    // operators of the path; d_u1 / tau: advecting velocity and time scale of the upwinded ones
    void apply(const std::string& op, double* d_x, double* d_coeff, double* d_y, int nlev, double scale, int tpow, int lev0 = 0,
               double* d_u1 = NULL, double tau = 0.0);
    // x = M1^-1 b on the partitioned mesh (mimsem_gpu_solve_M1_dist); returns the iteration count
    int solve_M1(const double* d_b, double* d_x, int nlev, double scale, int tpow, double rtol, int maxit, double* relres);
    bool halo_error();
    void sync();

private:
    struct Plan {
        int npush, npull;
        void* d_push;
        void* d_pull;
        unsigned long long* d_epochs;   // [push counter, pull counter]
        char* inbox;                    // my inbox of the space
        long long stride;               // doubles between inbox copies
        int push_rows;
    };
    void setup_p2p();
    void check(int rc, const char* what);
    const GlobalMesh& mesh_;
    Comm* comm_;
    Partition* part_;
    mimsem_gpu_ctx* ctx_;
    int device_, nk_max_;
    std::vector<int> perm_[3];          // caller row -> engine row
    Plan plan_[5];                      // spaces 0, 1, 2, then the extended plans of spaces 1 and 2
    char* base_;                        // my IPC buffer
    std::vector<char*> peer_base_;
    std::vector<void*> keep_;           // device row lists
    int* d_err_;
    void* stream_;                      // stream of the captured bursts
    void* d_red_areas_;
    void* d_red_seq_;
    static const int MAXP = 16, NBUF = 3;
};

}  // namespace mimsem_host
