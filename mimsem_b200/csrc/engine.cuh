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
    // Umat_ray (eul/Assembly.cpp:1846-1979): Rayleigh friction; c = Exner pressure 2-form of the column's level, c2 = the same
    // field's LEVEL-0 column (one value per face), ray_dt = dt (0: off); needs det and the level-0 thickness
    const double* c2;
    double ray_dt;
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
    double wb[kMaxP + 1];         // barycentric weights 1 / prod_{j != i} (xn[i] - xn[j])
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
    double wb[kMaxP + 1];
};

}  // namespace mimsem

namespace mimsem {

// ---------------------------------------------------------------------------------------------
// TMA tile kernels: one CTA per element, all levels.  The operands every thread of the tile reads many times
// (the element's own edges, its east column / north row, the inverse thickness, the 2-form coefficient) are
// staged in shared memory by 1-D bulk async copies (cp.async.bulk, the TMA unit) issued from a per-element
// copy list that the host builds once at set_topo; the operands that are read ONCE per tile -- the west / south
// neighbour's edge family behind its far line (and its 2-form coefficient for M1(h)) -- are loaded straight into
// registers (lanes = levels: coalesced) while the bulk copies are in flight and reduced to P+1 (resp. P) values per
// thread.  Keeping them out of shared memory cuts the tile from 105 to 65 slots (M1, p = 4: 31 KB instead of 50 KB),
// which is what bounds the number of resident tiles per SM.  All topology irregularity (cubed-sphere seams,
// reversed sides, ghost elements) lives in the per-element record; the kernel itself addresses shared memory with
// compile-time slot numbers.
//
// Slot s of the tile holds nlev consecutive doubles (one value per level).  Slot map for M1:
//   [0, 2P^2)            own edges in the engine's internal block order (== memory order, so the block is ONE
//                        bulk copy): x-edge (ix<P,iy) -> OX + ix P + iy (column-major: the west column is
//                        contiguous), y-edge (ix,iy<P) -> OY + iy P + ix (row-major: the south row is contiguous)
//   XE  + iy             east column of x-edges xx(P,iy)            (owned by the east neighbour)
//   YN  + ix             north row of y-edges   xy(ix,P)            (owned by the north neighbour)
//   T + qy (P+1) + qx    inverse layer thickness at the element's quadrature points
//   H  (+P^2)            2-form coefficient of the element (M1h only)
template <int P>
struct M1Slots {
    static constexpr int OX = 0;
    static constexpr int OY = P * P;
    static constexpr int XE = 2 * P * P;
    static constexpr int YN = XE + P;
    static constexpr int T = YN + P;
    static constexpr int NS = T + (P + 1) * (P + 1);
    static constexpr int H = NS;
    static constexpr int NS_H = H + P * P;
    // geometry record (doubles): per part (x-lines, y-lines) gl[part][line][q] = (g_own, g_oth), then (g_own, g_oth)[q] of
    // the west and the south far line.  (The K tile kernel shares the record type but reads G[q][3] at offset 0: its
    // records are built separately, see build_tile_geo.)
    static constexpr int GL = 0;
    static constexpr int GW = 2 * P * (P + 1) * 2;
    static constexpr int GS = GW + 2 * (P + 1);
    static constexpr int GEO = ((GS + 2 * (P + 1)) + 1) / 2 * 2;
    // K tile: plain G[q][3]
    static constexpr int GEO_K = (((P + 1) * (P + 1) * 3) + 1) / 2 * 2;
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

// Per-element tile record in global memory: kRecHdr 16-byte header words, then rec_ents copy entries, then (only in
// plans that need it) 2 x (P+1)P explicit far-line rows; fixed stride, so that lane l fetches its entry with ONE load
// whose address depends only on the element number.
struct TileHdr {      // word 0
    int st_dof;       // first output row of the element's owned block
    int cp_count;
    int flags;        // TF_* bits
    int nslots;       // slots filled by the list (for the mbarrier transaction count):
                      //   x / coefficient rows | ghost rows staged from the halo inbox << 12 | thickness << 20
};
struct TileFar {      // word 1: first rows of the far-line runs of the x field.  OTH(q,t), q <= P, t < P:
    int w16, w4;      //   west neighbour:  q < P -> row w16 + q P + t ; q == P -> row w4 +- t  (TF_W4_DESC)
    int s16, s4;      //   south neighbour: likewise
};
struct TileFarH {     // word 2
    int hw, hs;       // first row of the west / south neighbour's 2-form block (M1h)
    int n4, first4;   // copy entries [first4, first4 + n4) are the kind-4 ones (ghost rows staged from the halo inbox)
};
constexpr int kRecHdr = 3;
enum : int {
    TF_HAS_W = 1, TF_REV_W = 2, TF_HAS_S = 4, TF_REV_S = 8, TF_ROW_W = 16, TF_ROW_S = 32,   // ROW: the far line is a north row
    TF_W4_DESC = 64, TF_S4_DESC = 128,      // the 4-run is stored in descending row order
    TF_LIST_W = 256, TF_LIST_S = 512        // explicit row list instead of runs (entries < 0: inbox row -(r)-1)
};

// Ghost refresh fused into the tile kernel (multi-GPU): the first push_ctas CTAs of the grid store this rank's
// boundary rows into the peers' inboxes over NVLink and raise the peers' flags; tiles [0, n_int) read no ghost row;
// tiles >= n_int wait for the flags of this epoch and then stage their ghost rows straight from the inbox (copy-list
// kind 4) -- no pull kernel, no ghost rows in x.  The last CTA to finish acknowledges the inbox to the peers and
// advances the device-side epoch, so the launch replays inside a CUDA graph.
constexpr int kMaxPushPeers = 16;   // push descriptors staged in shared memory by the push role
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
    int ll;                          // 1: in-band protocol -- inbox cells are 16 bytes {lo, epoch, hi, epoch}, strides count cells
    const double* x_push;            // field whose boundary rows are pushed (lead 0: the input itself)
    int lead;                        // 0 or 1
    int nbuf;                        // inbox copies (2, or 3 so that a pipelined push never waits for the current consumer)
    int push_only;                   // prologue of a pipelined sequence: push data epoch e+1... nothing else, epoch unchanged
    // burst of launches that overlap under programmatic dependent launch (flag protocol, mode 0): launch `burst_pos` of
    // `burst_len` works on epoch  *epoch + 1 + burst_pos  (every launch of the burst reads the epoch word before the last
    // one advances it by burst_len), keeps its counters in slot burst_pos of `counters`, and hands its flags and
    // acknowledgements over in launch order through the local sequence words seq[0] (acks) and seq[1 + peer] (flags)
    int burst_pos, burst_len;
    unsigned* seq;
};
constexpr int kMaxBurst = 32;                    // launches of one burst (counter slots)
constexpr int kCtrStride = 1 + kMaxPushPeers;    // words per counter slot

struct TArgs {
    int halo_on;
    HaloFused halo;
    int nlev, ld, lev0, nkT, tpow;
    int contig_x;     // ld == nlev: a run of DOFs is one contiguous copy
    int contig_t;     // nkT == nlev
    int geo_doubles;
    int ntiles;               // tiles (elements) of this launch
    int prefetch_ahead;       // L2-prefetch the tile this many CTAs ahead (0: off)
    int prefetch_own_slots;   // x-field slots below this number are the element's own block
    int pdl;                  // programmatic dependent launch: let the next launch on the stream start as CTAs of this one retire
    double scale;
    const int* elist;         // optional element subset; nullptr: element = blockIdx.x
    const TileHdr* recs;      // [nel][rec_stride] 16-byte words: kRecHdr header words (K tile: 1), copy entries, far-row list
    int rec_stride;           // words per record
    int rec_hdr;              // header words before the copy entries
    int rec_list;             // word offset of the explicit far-row list (0: the plan has none)
    const double* geo;
    const double* x;
    const double* c;
    const double* tinv;
    double* y;
    double E[(kMaxP + 1) * kMaxP];
    double Es[(kMaxP + 1) * kMaxP];   // scale * E: the operator's scale factor rides on the last contraction (M1 tile kernel)
#ifdef MIMSEM_DIAG
    long long* dbg_times;     // [grid][6] phase timestamps (globaltimer ns), diagnostics build only
    int debug;                // bit0 skip arithmetic, bit4 skip copies (phase-isolation timing experiments)
#endif
};

}  // namespace mimsem
