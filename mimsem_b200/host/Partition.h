// Element-block partition of the canonical global mesh over the GPUs of one box (C++ host layer of the multi-GPU path).
//
// Rank g of G owns the contiguous element range [g Nel / G, (g+1) Nel / G) of the canonical numbering
// (e = face ne^2 + ey ne + ex) -- free of the reference's 6 n^2 rank constraint (README.md:32, scr/Setup.py:25-29) --
// and with it the edges and faces whose global ids fall into those elements' blocks (scr/Proc2.py:105-123); a node
// belongs to the rank of the lowest-numbered element around it.  Operators run owner-computes: besides its owned
// elements a rank holds the west / south neighbours of its owned elements (and every element around an owned node) as
// read-only halo elements, so outputs on owned DOFs are complete and the reference's two scatters
// (eul/Assembly.cpp:2194-2195, eul/Euler_2.cpp:1455-1456) collapse into ONE ghost refresh of the input field.
//
// Same construction, array for array, as mimsem_b200/parallel.py (class Partition, send_lists); the CPU test compares
// the two.  No CUDA, no PETSc: plain index arithmetic.
#pragma once
#include <cstdint>
#include <map>
#include <vector>

struct mimsem_mesh;

namespace mimsem_host {

// canonical global mesh tables (mimsem_mesh_*, csrc/mesh.cpp)
struct GlobalMesh {
    int kind = 0, p = 0, m = 0, ne = 0, nel = 0;
    int64_t N0 = 0, N1 = 0, N2 = 0, NQ = 0;
    std::vector<int> el0, el1x, el1y, el2, elq;   // [nel][(p+1)^2], [nel][p(p+1)], [nel][(p+1)p], [nel][p^2], [nel][(m+1)^2]
    std::vector<double> J, det;                   // [nel][(m+1)^2][4], [nel][(m+1)^2]
    std::vector<double> xyz;                      // [NQ][3]
    std::vector<int64_t> ws_nbr;                  // [nel][2]: element across the west / south side (-1: none)
    std::vector<int64_t> node_min_el;             // [N0]: lowest-numbered element around a node (its owner element)
    std::vector<int64_t> node_ptr, node_els;      // CSR node -> elements around it
    int n0e() const { return (p + 1) * (p + 1); }
    int n1e() const { return p * (p + 1); }
    int n2e() const { return p * p; }
    int nqe() const { return (m + 1) * (m + 1); }
    // kind = MIMSEM_MESH_SPHERE | MIMSEM_MESH_BOX; returns 0 or a mimsem error code
    int create(int kind, int p, int ne, bool signed_det = false, bool with_geometry = true);
};

void element_range(int64_t nel, int rank, int world, int64_t* e0, int64_t* e1);
int owner_rank_of_element(int64_t e, int64_t nel, int world);

struct GhostGroup {
    std::vector<int> local;       // caller (local) row ids, ascending
    std::vector<int64_t> glob;    // their global ids
};

struct Partition {
    int p = 0, rank = 0, world = 1;
    int64_t e0 = 0, e1 = 0;
    int n_interior = 0, nel_owned = 0, nel_total = 0;
    std::vector<int64_t> elements;               // local element -> global element (owned: interior first, boundary last; then halo)
    std::vector<int64_t> halo_ws;                // west / south halo elements
    std::vector<int64_t> g0, g1, g2, gq;         // local row -> global id (owned first)
    int n0 = 0, n1 = 0, n2 = 0, nq = 0;
    int n0_owned = 0, n1_owned = 0, n2_owned = 0;
    int n1_halo = 0, n2_halo = 0;                // rows [n_owned, n_halo) are refreshed from their owners
    std::vector<int> el0, el1x, el1y, el2, elq;  // element -> local row tables
    // ghost rows the element kernels read, grouped by owner rank (space 0, 1, 2); ext: ALL ghost rows (space 1, 2)
    std::map<int, GhostGroup> recv[3], recv_ext[3];
    // for every peer q: my owned local ids that q holds as ghosts, in q's receive order
    std::map<int, std::vector<int> > send[3], send_ext[3];

    Partition(const GlobalMesh& mesh, int rank, int world);
    // fills send / send_ext (builds every peer's partition: every rank can, so no index exchange is needed at set-up)
    void build_send_lists(const GlobalMesh& mesh);
    int n_owned(int space) const { return space == 0 ? n0_owned : (space == 1 ? n1_owned : n2_owned); }
    const std::vector<int64_t>& gids(int space) const { return space == 0 ? g0 : (space == 1 ? g1 : g2); }
};

}  // namespace mimsem_host

// C ABI for tests and foreign hosts: arrays of a rank's partition by name
// ("elements", "g0", "g1", "g2", "gq", "el0", "el1x", "el1y", "el2", "elq", "recv<space>_<peer>", "send<space>_<peer>";
//  sizes[12] = {nel_owned, nel_total, n_interior, n0, n1, n2, nq, n0_owned, n1_owned, n2_owned, n1_halo, n2_halo}).
extern "C" {
typedef struct mimsem_host_partition mimsem_host_partition;
int mimsem_host_partition_create(int kind, int p, int ne, int rank, int world, mimsem_host_partition** out);
void mimsem_host_partition_destroy(mimsem_host_partition* part);
int mimsem_host_partition_sizes(const mimsem_host_partition* part, int64_t sizes[12]);
/* returns the number of entries (as int64) of the named array, -1 if unknown; copies them to out if out != NULL */
int64_t mimsem_host_partition_array(const mimsem_host_partition* part, const char* name, int64_t* out);
}
