// Device-side data structures shared by the kernels and the host engine (sm_100a).
#pragma once
#include <cstdint>

namespace mimsem {

constexpr int kMaxP = 6;

// Kernel arguments, passed by value: they live in the constant bank, so the (p+1) x p edge table
// E and the quadrature weights are broadcast operands of the FP64 FMAs (no shared-memory or
// global traffic for the basis).  The reference re-tabulates these inside every assemble call
// (eul/Assembly.cpp:69-71).
struct KArgs {
    // work
    int nel;            // elements to compute (owned)
    int nlev;           // columns per field
    int ld;             // leading dimension of the fields
    int lev0;           // thickness level of column 0
    int lev_stride;     // 1, or 0 when every column uses thickness level lev0
    int nkT;            // leading dimension of the thickness table
    int tpow;           // number of 1/thick factors per quadrature point (0..2)
    double scale;
    // topology (device pointers, local indices)
    const int* el1x;
    const int* el1y;
    const int* el2;
    const int* elq;
    const int* nbr;               // [nel][2] west / south neighbour: elem | side<<29 | rev<<30, or -1
    const unsigned char* eflags;  // [nel] bit0: write east side, bit1: write north side (partial-sum mode)
    const int* elist;             // optional element subset (interior / boundary); nullptr: elements 0..nel-1
    // line-task tables (transposed copies so that one GLL line is contiguous)
    const int* el1xT;             // [nel][ix][iy]  x-normal edges, column-major
    const int* elqT;              // [nel][qx][qy]  quadrature points, column-major
    const double* Gc;             // [nel][qx][qy][2] = (g0, g1) of G, column-major   (x-line tasks)
    const double* Gr;             // [nel][qy][qx][2] = (g1, g2) of G, row-major      (y-line tasks)
    unsigned div_m, div_s;        // magic number / shift for idx / nlev
    // geometry (device pointers)
    const double* G;              // [nel_total][q2][3] or [nel_total][q2]
    const double* tinv;           // [nq][nkT] inverse layer thickness, level fastest
    // fields
    const double* c;              // coefficient field (h2 or u1) or nullptr
    const double* x;
    double* y;
    // basis
    double E[(kMaxP + 1) * kMaxP];   // E[q*p + i] = e_i(x_q)
    // rotational / upwinded operators (RotMat, RotMat_up, Phmat::assemble_up; src/Assembly.cpp:1346-1395, 1784-1853, 499-567)
    const int* el0;               // [nel_total][(p+1)^2] element -> node
    const double* q0;             // 0-form coefficient (potential vorticity), column layout
    const double* u1;             // 1-form advecting velocity (upwinded variants), column layout
    const double* J4;             // [nel_total][q2][4] Jacobian J00 J01 J10 J11
    const double* det;            // [nel_total][q2]
    const double* Wr;             // [nel_total][q2] w_q * (J00 J11 - J01 J10) / det   (= +-w_q)
    double tau;                   // upwinding time scale fac*dt
    double xn[kMaxP + 1];         // GLL nodes
};

// Division of a 32-bit index by a launch-constant divisor (Granlund-Montgomery):
// q = (umulhi(m, n) + n) >> s  with  s = ceil(log2 d),  m = floor(2^32 (2^s - d) / d) + 1.
struct FastDiv {
    unsigned m = 0, s = 0;
};
inline FastDiv make_fastdiv(unsigned d) {
    FastDiv f;
    unsigned s = 0;
    while ((1ull << s) < d) s++;
    f.s = s;
    f.m = (unsigned)(((1ull << 32) * ((1ull << s) - d)) / d + 1);
    return f;
}
#ifdef __CUDACC__
__device__ __forceinline__ unsigned fastdiv(unsigned n, unsigned m, unsigned s) {
    return (unsigned)(((unsigned long long)__umulhi(m, n) + n) >> s);
}
#endif

// Peer-to-peer halo exchange (no NCCL on the data path): the sender's kernel stores its boundary rows
// straight into the receiver's inbox over NVLink and then raises a flag there; the receiver's kernel waits on the
// flag, scatters the inbox into its ghost rows and acknowledges.  Epoch counters live in device memory so the
// whole step (push, interior kernel, pull, boundary kernel) replays as one CUDA graph.
struct HaloPeer {
    const int* rows;          // local rows to send (push) / ghost rows to fill (pull)
    int nrows;
    int row0;                 // first inbox row of this peer's share (inbox rows are packed with stride nlev)
    double* inbox;            // push: the PEER's inbox of this space (peer address); pull: my inbox of this space
    long long inbox_parity_stride;   // doubles between the two parity buffers
    unsigned long long* signal;      // push: flag on the peer (I write) ; pull: ack on the peer (I write)
    const unsigned long long* wait;  // push: ack from the peer (peer writes, local memory) ; pull: flag from the peer
};

// Padded ELL incidence stencil: row r has entries (col[r*width+j], sgn[r*width+j]), col = -1 pads.
struct EllArgs {
    int64_t nrows;
    int width;
    int nlev, ld;
    unsigned div_m, div_s;
    const int* rows;     // optional row list (nullptr: rows 0..nrows-1)
    const int* col;
    const signed char* sgn;
    const double* x;
    double* y;
};

// node -> (element, quadrature point) adjacency for the 0-form operators
struct NodeArgs {
    int n0;
    int nlev, ld, lev0, lev_stride, nkT, tpow;
    unsigned div_m, div_s;
    double scale;
    const int* adj_ptr;   // [n0+1]
    const int* adj_eq;    // elem*q2 + q
    const int* node_q;    // [n0] quadrature-point index of the node (thickness lookup)
    const int* el2;
    const double* D0;     // [n0] sum_e w_q det
    const double* wq;     // [q2] w_qx * w_qy
    const double* tinv;
    const double* c;      // h2 or nullptr
    const double* x;
    double* y;
    double E[(kMaxP + 1) * kMaxP];
    // Phmat::assemble_up
    const int* el0;
    const int* el1x;
    const int* el1y;
    const double* u1;
    const double* J4;
    const double* det;
    double tau;
    double xn[kMaxP + 1];
};

}  // namespace mimsem

namespace mimsem {

// ---------------------------------------------------------------------------------------------
// TMA tile kernels: one CTA per element, all levels; every operand of the element-level
// contraction is staged in shared memory by 1-D bulk async copies (cp.async.bulk, the TMA unit),
// issued from a per-element copy list that the host builds once at set_topo.  All topology
// irregularity (cubed-sphere seams, reversed sides, ghost elements) lives in that list; the
// kernel itself addresses shared memory with compile-time slot numbers.
//
// Slot s of the tile holds nlev consecutive doubles (one value per level).  Slot map for M1:
//   [0, 2P^2)            own edges in the engine's internal block order (== memory order, so the block is ONE
//                        bulk copy): x-edge (ix<P,iy) -> OX + ix P + iy (column-major: the west column is
//                        contiguous), y-edge (ix,iy<P) -> OY + iy P + ix (row-major: the south row is contiguous)
//   XE  + iy             east column of x-edges xx(P,iy)            (owned by the east neighbour)
//   YN  + ix             north row of y-edges   xy(ix,P)            (owned by the north neighbour)
//   WOTH + q P + t       the west neighbour's other-family edges along its far line
//   SOTH + q P + t       the south neighbour's other-family edges along its far line
//   T + qy (P+1) + qx    inverse layer thickness at the element's quadrature points
//   H, HW, HS (+P^2 each) 2-form coefficient of the element / west / south neighbour (M1h only)
template <int P>
struct M1Slots {
    static constexpr int OX = 0;
    static constexpr int OY = P * P;
    static constexpr int XE = 2 * P * P;
    static constexpr int YN = XE + P;
    static constexpr int WOTH = YN + P;
    static constexpr int SOTH = WOTH + (P + 1) * P;
    static constexpr int T = SOTH + (P + 1) * P;
    static constexpr int NS = T + (P + 1) * (P + 1);
    static constexpr int H = NS;
    static constexpr int HW = H + P * P;
    static constexpr int HS = HW + P * P;
    static constexpr int NS_H = HS + P * P;
    // geometry record (doubles): G[q][3], then (c_own, c_oth)[q] for the west and the south far line
    static constexpr int GW = (P + 1) * (P + 1) * 3;
    static constexpr int GS = GW + 2 * (P + 1);
    static constexpr int GEO = ((GS + 2 * (P + 1)) + 1) / 2 * 2;
};

// Slot map of the K (WtQUmat) tile: the element's x edges as in M1Slots, the same block again for the velocity
// coefficient u1, then the inverse thickness; element-local (no neighbour data).  The geometry record is M1(h)'s.
template <int P>
struct KSlots {
    static constexpr int OX = 0;
    static constexpr int OY = P * P;
    static constexpr int XE = 2 * P * P;
    static constexpr int YN = XE + P;
    static constexpr int U0 = YN + P;            // u1 block: same layout, offset U0
    static constexpr int T = 2 * U0;
    static constexpr int NS = T + (P + 1) * (P + 1);
};

struct CopyEnt {      // 16 bytes
    int kind;         // 0 x field, 1 coefficient field, 2 inverse thickness, 3 geometry record, 4 x field from the halo inbox
    int src;          // first DOF / quadrature point / element / inbox row
    int slot;         // first destination slot (kind 3: ignored)
    int count;        // consecutive DOFs -> consecutive slots
};

// Per-element tile record in global memory: one header followed by rec_ents copy entries (fixed stride), so
// that lane l fetches its entry with ONE load whose address depends only on the element number.
struct TileHdr {      // 16 bytes
    int st_dof;       // first output row of the element's owned block (-1: use the general store list)
    int cp_count;
    int flags;        // bit0 has west nbr, bit1 west reversed, bit2 has south nbr, bit3 south reversed
    int nslots;       // slots filled by the list (for the mbarrier transaction count)
};

struct StoreEnt {     // 16 bytes
    int slot, dof, count, pad;
};

// Ghost refresh fused into the tile kernel (multi-GPU): the first push_ctas CTAs of the grid store this rank's
// boundary rows into the peers' inboxes over NVLink and raise the peers' flags; tiles [0, n_int) read no ghost row;
// tiles >= n_int wait for the flags of this epoch and then stage their ghost rows straight from the inbox (copy-list
// kind 4) -- no pull kernel, no ghost rows in x.  The last CTA to finish acknowledges the inbox to the peers and
// advances the device-side epoch, so the launch replays inside a CUDA graph.
struct HaloFused {
    int npush, npull;
    int push_ctas;
    int n_int;                       // tiles below this index never touch the inbox
    const HaloPeer* push;
    const HaloPeer* pull;
    const double* inbox;             // my inbox of the x field's space, parity 0
    long long parity_stride;         // doubles between consecutive inbox copies
    unsigned long long* epoch;
    unsigned* counters;              // [1 + npush]: finished CTAs of the launch, finished push CTAs per peer
    int* err;
    // software pipelining over independent applies: with lead = 1 the push CTAs send the boundary rows of the NEXT
    // call's input (x_push, data epoch e+1) while this call's boundary tiles consume what the previous call pushed
    const double* x_push;            // field whose boundary rows are pushed (lead 0: the input itself)
    int lead;                        // 0 or 1
    int nbuf;                        // inbox copies (2, or 3 so that a pipelined push never waits for the current consumer)
    int push_only;                   // prologue of a pipelined sequence: push data epoch e+1... nothing else, epoch unchanged
};

struct TArgs {
    int halo_on;
    HaloFused halo;
    int nlev, ld, lev0, nkT, tpow;
    int contig_x;     // ld == nlev: a run of DOFs is one contiguous copy
    int contig_t;     // nkT == nlev
    int geo_doubles;
    long long* dbg_times;     // optional [grid][6] phase timestamps (globaltimer ns) for latency breakdowns; nullptr: off
    int ntiles;               // tiles (elements) of this launch; the grid is persistent
    int prefetch_ahead;       // L2-prefetch the tile this many CTAs ahead (0: off)
    int prefetch_own_slots;   // x-field slots below this number are the element's own block
    int debug_slot_lo, debug_slot_hi;   // bit1 of debug: skip x-field copies into slots [lo, hi) (traffic experiment)
    int debug;        // bit0: skip the arithmetic (data-movement-only timing experiment, MIMSEM_DEBUG=1)
    double scale;
    const int* elist;         // optional element subset; nullptr: element = blockIdx.x
    const TileHdr* recs;      // [nel][1 + rec_ents] 16-byte words: header, then the copy entries
    int rec_ents;
    const int* st_ptr;        // [nel+1]
    const StoreEnt* stores;
    const double* geo;
    const double* x;
    const double* c;
    const double* tinv;
    double* y;
    double E[(kMaxP + 1) * kMaxP];
};

}  // namespace mimsem
