// y = M1 x and y = M1(h) x as a PERSISTENT, warp-specialised tile kernel (Umat / Uhmat / Ut_mat, eul/Assembly.cpp:51-153,
// 416-474, 1338-1440).  Same tile record, copy list, slot map and line contraction as k_apply_m1_tile (m1_tile.cuh);
// what changes is who waits for memory.  Measured on this kernel's first version, which looked one tile ahead
// (profiles/r02_summary.md): the contraction of a tile takes 0.75 us on warps that never touch global memory, while
// record fetch -> bulk copies and record fetch -> far-line rows are two dependent chains of 2-3 us.  So:
//
//   * ONE CTA per SM, a ring of NB tile buffers (all the shared memory there is); the CTA owns a contiguous chunk of
//     the tile list, split in two halves; "job" j of the CTA is tile (j >> 1) of half (j & 1) and lives in buffer j % NB;
//   * warps 0-7 = G = 2 contraction groups (two warp-pairs x 64 level lanes each, as in the tile kernel); group g takes
//     jobs g, g + G, ..., i.e. it walks through ITS half tile by tile: waits for full[j % NB], contracts, stores straight
//     from registers, releases empty[j % NB].  Consecutive tiles of a half are west-east neighbours on the mesh, so the
//     far-line operands of the west neighbour are what the same threads interpolated one tile earlier from registers
//     (tile_lines<CARRY>): they stay in registers and neither the stagers nor the memory system see that side again;
//   * warps 8.. = NB stagers, one WARP per ring buffer; stager w takes jobs w, w + NB, ...: with the tile record already in
//     registers (fetched during its previous job) it loads the far-line rows of the west / south neighbours straight
//     from global memory (16-byte loads, lane = level pair) and interpolates them onto the far line, waits for
//     empty[], walks the copy list (bulk copies -> full[]), leaves the far-line operands in 2(P+1) (+ 2P) extra slots
//     of the buffer and arrives on full[].  All stagers are in flight on different tiles, so their latency chains
//     (measured: 1-1.2 us per dependent L2 round trip under load) overlap each other and the contraction.
//
// No CTA-wide barrier inside the loop: buffers are handed over through mbarriers only.
#pragma once
#include "m1_tile.cuh"

namespace mimsem {

template <int P, bool WITH_H>
struct M1Pipe {
    using S = M1Slots<P>;
    static constexpr int G = 2;                     // contraction groups (128 threads each)
    static constexpr int MAXB = 8;                  // ring size limit (barrier storage)
    static constexpr int MAX_THREADS = G * 128 + MAXB * 32;   // one stager warp per ring buffer
    static __host__ __device__ constexpr int threads(int nb) { return G * 128 + nb * 32; }
    static constexpr int NSLOT = WITH_H ? S::NS_H : S::NS;
    // far-line operands of the west (side 0) / south (side 1) neighbour, reduced to the far line by the stagers:
    // ubf[side][q] -> slot FARU + side (P+1) + q ; hs[side][j] (M1(h) only) -> slot FARH + side P + j
    static constexpr int FARU = NSLOT;
    static constexpr int FARH = FARU + 2 * (P + 1);
    static constexpr int SLOTS = FARH + (WITH_H ? 2 * P : 0);
    static constexpr int HDR_BYTES = 2 * MAXB * 8 + MAXB * 32;   // full[], empty[], tile headers (PipeHdr)
    static __host__ __device__ constexpr size_t buf_doubles(int nl) { return (size_t)S::GEO + (size_t)SLOTS * nl; }
    static __host__ __device__ constexpr int ring(int nl, size_t smem_max) {
        const size_t n = (smem_max - HDR_BYTES) / (buf_doubles(nl) * sizeof(double));
        // a multiple of G, so that a buffer is always consumed by the same contraction group: every waiter then follows the
        // phases of its barriers in order (a parity wait cannot tell "two phases behind" from "done")
        return n > (size_t)MAXB ? MAXB : (int)n / G * G;
    }
    static __host__ __device__ constexpr size_t smem_bytes(int nl, int nb) { return HDR_BYTES + (size_t)nb * buf_doubles(nl) * sizeof(double); }
};

struct PipeHdr {      // 32 bytes per ring buffer, written by the stager
    TileHdr hd;
    int carry_w;      // the west far-line operands are the ones the contraction group carried over from its previous tile
    int pad[3];
};

// job j of a CTA whose chunk is [start, start + n): half g = j & 1 covers [start, start + n0) (g = 0, n0 = ceil(n / 2)) or
// [start + n0, start + n); tile = half start + (j >> 1)
__device__ __forceinline__ int pipe_tile(int start, int n0, int j) { return start + ((j & 1) ? n0 : 0) + (j >> 1); }

__device__ __forceinline__ void named_barrier_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Far-line operands of one side for a whole tile by ONE warp: lane l holds levels 2l, 2l+1 (16-byte loads; rows are
// 16-byte aligned and of even length on this path).  Arithmetic and summation order are those of far_fetch (m1_tile.cuh),
// so the result is bit for bit the same.  The loads bypass L1 (ld.global.cg): with the ring taking all but ~27 KB of the
// unified array, rows streaming through L1 evicted whatever else lived there.
template <int P, bool WITH_H, class PP>
__device__ __forceinline__ void far_fetch_warp(const TArgs& a, const TileHdr* rec, int flags, const TileFar& fr, const TileFarH& fh, int side,
                                               int lane, int nl, double2 (&ubf)[P + 1], double2 (&hs)[P]) {
#pragma unroll
    for (int q = 0; q <= P; q++) ubf[q] = make_double2(0.0, 0.0);
#pragma unroll
    for (int j = 0; j < P; j++) hs[j] = make_double2(0.0, 0.0);
    if (!(flags & (side ? TF_HAS_S : TF_HAS_W)) || 2 * lane >= nl) return;
    constexpr int NF = (P + 1) * P;
    double2 oth[P + 1][P];
    const double2* xk = reinterpret_cast<const double2*>(a.x) + lane;
    const size_t ld2 = (size_t)a.ld >> 1;
    if (!(flags & (side ? TF_LIST_S : TF_LIST_W))) {
        const double2* b16 = xk + (size_t)(side ? fr.s16 : fr.w16) * ld2;
        const double2* b4 = xk + (size_t)(side ? fr.s4 : fr.w4) * ld2;
        const long long st4 = (flags & (side ? TF_S4_DESC : TF_W4_DESC)) ? -(long long)ld2 : (long long)ld2;
#pragma unroll
        for (int q = 0; q < P; q++)
#pragma unroll
            for (int t = 0; t < P; t++) oth[q][t] = __ldcg(b16 + (size_t)(q * P + t) * ld2);
#pragma unroll
        for (int t = 0; t < P; t++) oth[P][t] = __ldcg(b4 + t * st4);
    } else {
        const int* list = reinterpret_cast<const int*>(rec + a.rec_list) + side * NF;
#pragma unroll
        for (int q = 0; q <= P; q++)
#pragma unroll
            for (int t = 0; t < P; t++) oth[q][t] = __ldcg(xk + (size_t)list[q * P + t] * ld2);
    }
    double2 hv[WITH_H ? P : 1][WITH_H ? P : 1];
    if (WITH_H) {
        const double2* hb = reinterpret_cast<const double2*>(a.c) + (size_t)(side ? fh.hs : fh.hw) * ld2 + lane;
#pragma unroll
        for (int iy = 0; iy < P; iy++)
#pragma unroll
            for (int ix = 0; ix < P; ix++) hv[WITH_H ? iy : 0][WITH_H ? ix : 0] = __ldcg(hb + (size_t)(iy * P + ix) * ld2);
    }
#pragma unroll
    for (int q = 0; q <= P; q++) {
        double sx = 0.0, sy = 0.0;
#pragma unroll
        for (int t = 0; t < P; t++) {
            sx += a.E[P * P + t] * oth[q][t].x;
            sy += a.E[P * P + t] * oth[q][t].y;
        }
        ubf[q] = make_double2(sx, sy);
    }
    if (WITH_H) {
        const bool far_is_row = flags & (side ? TF_ROW_S : TF_ROW_W);
#pragma unroll
        for (int j = 0; j < P; j++) {
            double sx = 0.0, sy = 0.0;
#pragma unroll
            for (int t = 0; t < P; t++) {
                const double2 v = far_is_row ? hv[WITH_H ? t : 0][WITH_H ? j : 0] : hv[WITH_H ? j : 0][WITH_H ? t : 0];
                sx += a.E[P * P + t] * v.x;
                sy += a.E[P * P + t] * v.y;
            }
            hs[j] = make_double2(sx, sy);
        }
    }
}

#ifdef MIMSEM_DIAG
// per CTA and job: stager start / copies issued / far operands written, contraction start / copies landed / done (globaltimer ns)
#define PIPE_T(j, i, cond) do { if (a.dbg_times && (cond)) a.dbg_times[((size_t)blockIdx.x * 128 + ((j) & 127)) * 8 + (i)] = gtime_ns(); } while (0)
#else
#define PIPE_T(j, i, cond) do { } while (0)
#endif

template <int P, bool WITH_H, int NL, int TPOW>
__global__ void __launch_bounds__(M1Pipe<P, WITH_H>::MAX_THREADS, 1) k_apply_m1_pipe(const __grid_constant__ TArgs a, const int NB) {
    using S = M1Slots<P>;
    using PP = M1Pipe<P, WITH_H>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);                       // [MAXB]
    uint64_t* empty = full + PP::MAXB;                                            // [MAXB]
    PipeHdr* hdrs = reinterpret_cast<PipeHdr*>(smem_raw + 2 * PP::MAXB * 8);      // [MAXB]
    double* buf0 = reinterpret_cast<double*>(smem_raw + PP::HDR_BYTES);
    const int nl = NL ? NL : a.nlev;
    const size_t bufd = PP::buf_doubles(nl);
    const int k = threadIdx.x & 63;
    const bool active = k < nl;   // level lanes beyond nlev only take part in the barriers
    // this CTA's chunk of the tile list
    const int per = a.ntiles / (int)gridDim.x, rem = a.ntiles % (int)gridDim.x;
    const int njobs = per + ((int)blockIdx.x < rem ? 1 : 0);
    const int start = (int)blockIdx.x * per + min((int)blockIdx.x, rem);
    const int n0 = (njobs + 1) >> 1;
    if (njobs == 0) return;
    if (threadIdx.x == 0) {
        for (int b = 0; b < NB; b++) {
            mbar_init(&full[b], 1 + 32);   // expect_tx arrival of the copy-list walk + the 32 lanes of the stager
            mbar_init(&empty[b], 4);       // one lane of each warp of the contraction group
        }
        fence_async_smem();
    }
    if (a.pdl) pdl_launch_dependents();
    __syncthreads();

    if (threadIdx.x >= PP::G * 128) {
        // ---- stager warp s ------------------------------------------------------------------------------------------
        const int s = ((int)threadIdx.x - PP::G * 128) >> 5;
        const int lane = threadIdx.x & 31;
        const int b = s;
        int n = 0;   // use count of buffer b
        if (s >= njobs) return;
        // record words of the current job (header, far rows, this lane's first copy entry): always one job ahead in registers
        const TileHdr* rec;
        TileHdr hd;
        TileFar fr;
        TileFarH fh;
        CopyEnt ent;
        int prev_st;   // first owned row of the tile before this one in its half (-1: none)
        auto fetch_record = [&](int j, const TileHdr*& r, TileHdr& h, TileFar& f, TileFarH& f2, CopyEnt& en, int& pst) {
            const int ti = pipe_tile(start, n0, j);
            r = tile_record(a, a.elist ? a.elist[ti] : ti);
            h = r[0];
            f = *reinterpret_cast<const TileFar*>(r + 1);
            f2 = *reinterpret_cast<const TileFarH*>(r + 2);
            en = tile_first_entry(a, r);
            pst = -1;
            if ((j >> 1) > 0) pst = tile_record(a, a.elist ? a.elist[ti - 1] : ti - 1)->st_dof;
        };
        fetch_record(s, rec, hd, fr, fh, ent, prev_st);
        // stager s owns buffer s (NS == NB): its waits on empty[s] follow the phases of that one barrier in order
        for (int j = s; j < njobs; j += NB) {
            PIPE_T(j, 0, lane == 0);
            const TileHdr* rec_n = rec;
            TileHdr hd_n = hd;
            TileFar fr_n = fr;
            TileFarH fh_n = fh;
            CopyEnt ent_n = ent;
            int prev_st_n = -1;
            if (j + NB < njobs) fetch_record(j + NB, rec_n, hd_n, fr_n, fh_n, ent_n, prev_st_n);
            // plain west neighbour (same orientation, rows in runs) that is the previous tile of this half: carried
            constexpr int WBITS = TF_HAS_W | TF_REV_W | TF_ROW_W | TF_LIST_W | TF_W4_DESC;
            const bool carry_w = prev_st >= 0 && (hd.flags & WBITS) == TF_HAS_W && fr.w16 - P * P == prev_st &&
                                 (!WITH_H || fh.hw >= 0);
            double* geo = buf0 + (size_t)b * bufd;
            double* tile = geo + S::GEO;
            // far-line operands: global -> registers before the buffer is free (they do not need it yet)
            double2 ubf[2][P + 1], hs[2][P];
#ifdef MIMSEM_DIAG
            if (a.debug & 32) {
#pragma unroll
                for (int side = 0; side < 2; side++) {
#pragma unroll
                    for (int q = 0; q <= P; q++) ubf[side][q] = make_double2(0.0, 0.0);
#pragma unroll
                    for (int jj = 0; jj < P; jj++) hs[side][jj] = make_double2(0.0, 0.0);
                }
            } else
#endif
            {
                if (!carry_w) far_fetch_warp<P, WITH_H, PP>(a, rec, hd.flags, fr, fh, 0, lane, nl, ubf[0], hs[0]);
                far_fetch_warp<P, WITH_H, PP>(a, rec, hd.flags, fr, fh, 1, lane, nl, ubf[1], hs[1]);
            }
            if (n > 0) mbar_wait(&empty[b], (unsigned)(n - 1) & 1u);
            tile_load_pre(a, rec, hd, ent, nullptr, &full[b], geo, tile);
            PIPE_T(j, 1, lane == 0);
            if (lane == 0) {
                hdrs[b].hd = hd;
                hdrs[b].carry_w = carry_w ? 1 : 0;
            }
            if (2 * lane < nl) {
#pragma unroll
                for (int side = 0; side < 2; side++) {
                    if (side == 0 && carry_w) continue;
#pragma unroll
                    for (int q = 0; q <= P; q++) reinterpret_cast<double2*>(tile + (size_t)(PP::FARU + side * (P + 1) + q) * nl)[lane] = ubf[side][q];
                    if (WITH_H) {
#pragma unroll
                        for (int jj = 0; jj < P; jj++) reinterpret_cast<double2*>(tile + (size_t)(PP::FARH + side * P + jj) * nl)[lane] = hs[side][jj];
                    }
                }
            }
            mbar_arrive(&full[b]);
            PIPE_T(j, 2, lane == 0);
            rec = rec_n;
            hd = hd_n;
            fr = fr_n;
            fh = fh_n;
            ent = ent_n;
            prev_st = prev_st_n;
            n++;
        }
    } else {
        // ---- contraction group g ------------------------------------------------------------------------------------
        const int g = threadIdx.x >> 7;
        const int part = (threadIdx.x >> 6) & 1;
        int b = g % NB, n = g / NB;
        // (loop bound straight from the parameter bank: a register held across the contraction is one too many)
        double cw[P + 1], chs[P];   // west far-line operands of the NEXT tile of this half (part 0 threads)
#pragma unroll
        for (int q = 0; q <= P; q++) cw[q] = 0.0;
#pragma unroll
        for (int jj = 0; jj < P; jj++) chs[jj] = 0.0;
        for (int j = g; j < njobs; j += PP::G) {
            double* geo = buf0 + (size_t)b * bufd;
            double* tile = geo + S::GEO;
            double* col = tile + k;
            PIPE_T(j, 3, (threadIdx.x & 127) == 0);
            mbar_wait(&full[b], (unsigned)n & 1u);
            PIPE_T(j, 4, (threadIdx.x & 127) == 0);
            const TileHdr hd = hdrs[b].hd;
            const bool carried = part == 0 && hdrs[b].carry_w;
            double cfar[P];
#ifdef MIMSEM_DIAG
            // timing experiments (results are garbage): 1 = no contraction, no stores; 64 = stores only
            if (a.debug & 64) {
                double* yy = a.y + (size_t)(hd.st_dof + (part ? S::OY : S::OX)) * a.ld + k;
                if (active)
                    for (int r = 0; r < P * P; r++) yy[(size_t)r * a.ld] = col[(size_t)r * nl];
            }
            if (!(a.debug & 65)) {
#endif
            if (active) {
                double ubf[P + 1], hs[P];
                if (carried) {
#pragma unroll
                    for (int q = 0; q <= P; q++) ubf[q] = cw[q];
#pragma unroll
                    for (int jj = 0; jj < P; jj++) hs[jj] = chs[jj];
                } else {
#pragma unroll
                    for (int q = 0; q <= P; q++) ubf[q] = col[(size_t)(PP::FARU + part * (P + 1) + q) * nl];
#pragma unroll
                    for (int jj = 0; jj < P; jj++) hs[jj] = WITH_H ? col[(size_t)(PP::FARH + part * P + jj) * nl] : 0.0;
                }
                far_line<P, WITH_H, NL, TPOW>(a, col, geo, hd.flags, part, ubf, hs, cfar);
            }
            if (WITH_H) {
                named_barrier_sync(1 + g, 128);   // the far lines have read the raw thickness of the west column / south row
                if (active) h_prepass<P, NL, TPOW, true>(a, col, part, chs);
                named_barrier_sync(1 + g, 128);
            }
            if (active) tile_lines<P, WITH_H, NL, TPOW, true>(a, col, geo, part, cfar, a.y + (size_t)(hd.st_dof + (part ? S::OY : S::OX)) * a.ld + k, cw);
#ifdef MIMSEM_DIAG
            }
#endif
            if (WITH_H) fence_async_smem();   // the thickness slots were rewritten through the generic proxy; a bulk copy overwrites them next
            PIPE_T(j, 5, (threadIdx.x & 127) == 0);
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[b]);
            b += PP::G;
            while (b >= NB) {
                b -= NB;
                n++;
            }
        }
    }
}

}  // namespace mimsem
