// TMA-staged tile kernels (sm_100a): see the comment above M1Slots in engine.cuh.
#pragma once
#include <cstdint>

#include "engine.cuh"
#include "kernels.cuh"

namespace mimsem {

__device__ __forceinline__ long long gtime_ns() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define DBG_T(i) do { if (a.dbg_times && threadIdx.x == 64) a.dbg_times[(size_t)tile_i * 6 + (i)] = gtime_ns(); } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared 1-D bulk copy (TMA), completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global 1-D bulk copy
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
// commit the bulk stores and wait until their shared-memory source has been read (the CTA may then exit;
// global visibility is guaranteed at kernel completion)
__device__ __forceinline__ void bulk_commit_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// ask the TMA unit to pull a range into L2 (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Push role of the fused ghost refresh: CTA b of push_ctas copies its share of every peer's rows into that peer's
// inbox (16-byte accesses, lanes along the levels) and the last CTA to finish a peer raises that peer's flag.
#define DBG_PUSH(i) do { if (a.dbg_times && threadIdx.x == 0) a.dbg_times[(size_t)(a.ntiles + blockIdx.x) * 6 + (i)] = gtime_ns(); } while (0)
// Push role of the fused ghost refresh.  The rows of all peers form one flat list that is split evenly over the push
// CTAs (a CTA's share may straddle peers), sized so that a CTA moves its share in a single pass: every thread has up to
// four independent 16-byte loads in flight, then stores them into the peers' inboxes over NVLink.  The copy is
// latency-bound (index load -> HBM load -> remote store -> system fence), hence many small CTAs rather than few big ones.
constexpr int kMaxPushPeers = 16;
__device__ __noinline__ void halo_push_role(const TArgs& a, unsigned long long epoch) {
    const HaloFused& h = a.halo;
    __shared__ HaloPeer peers[kMaxPushPeers];
    __shared__ int pre[kMaxPushPeers + 1];
    DBG_PUSH(0);
    if (threadIdx.x < h.npush) {
        peers[threadIdx.x] = h.push[threadIdx.x];
        // the peer must have consumed the inbox copy of epoch - nbuf (the one this push overwrites)
        if (epoch > (unsigned long long)h.nbuf) spin_until(h.push[threadIdx.x].wait, epoch - h.nbuf, h.err);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int p = 0; p < h.npush; p++) {
            pre[p] = acc;
            acc += peers[p].nrows;
        }
        pre[h.npush] = acc;
    }
    __syncthreads();
    DBG_PUSH(1);
    const int nl2 = a.nlev >> 1;
    const int R = pre[h.npush];
    const int f0 = (int)(((long long)R * blockIdx.x) / h.push_ctas);
    const int f1 = (int)(((long long)R * (blockIdx.x + 1)) / h.push_ctas);
    const int total = (a.debug & 4) ? 0 : (f1 - f0) * nl2;   // debug bit 2: signal without copying (timing experiment)
    constexpr int U = 4;
    for (int base = threadIdx.x; base < total; base += U * blockDim.x) {
        double2 v[U];
        double2* dst[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int i = base + u * blockDim.x;
            dst[u] = nullptr;
            if (i < total) {
                const int f = f0 + i / nl2, k2 = i - (f - f0) * nl2;
                int p = 0;
                while (f >= pre[p + 1]) p++;
                const int r = f - pre[p];
                const HaloPeer& pp = peers[p];
                v[u] = __ldg(reinterpret_cast<const double2*>(h.x_push + (size_t)pp.rows[r] * a.ld) + k2);
                dst[u] = reinterpret_cast<double2*>(pp.inbox + (epoch % h.nbuf) * pp.inbox_parity_stride + (size_t)(pp.row0 + r) * a.nlev) + k2;
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            if (dst[u]) *dst[u] = v[u];
    }
    DBG_PUSH(2);
    // the CTA barrier orders every thread's stores before the npush threads below; their system-scope fences are
    // cumulative, so only they (not all 128 threads) pay the NVLink round trip
    __syncthreads();
    if (threadIdx.x < h.npush) {
        const int p = threadIdx.x;
        __threadfence_system();
        const unsigned done = atomicAdd(&h.counters[1 + p], 1u);
        if (done == (unsigned)h.push_ctas - 1) {
            h.counters[1 + p] = 0;
            __threadfence_system();
            st_release_sys(peers[p].signal, epoch);
        }
    }
    DBG_PUSH(3);
    DBG_PUSH(4);
}

// Every CTA of a fused launch ends here; the last one acknowledges the inbox of this epoch to the peers (all bulk
// loads from it have completed) and advances the epoch for the next launch / graph replay.
__device__ __noinline__ void halo_cta_done(const TArgs& a, unsigned long long epoch) {
    // only the push CTAs and the boundary tiles take part (interior tiles neither read the epoch nor touch the inbox)
    __shared__ int last;
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(&a.halo.counters[0], 1u);
        last = (done == (unsigned)(a.halo.push_ctas + a.ntiles - a.halo.n_int) - 1) ? 1 : 0;
        if (last) a.halo.counters[0] = 0;
    }
    __syncthreads();
    if (!last || a.halo.push_only) return;
    // one thread per peer: the release stores cross NVLink concurrently (a serial loop costs a round trip per peer)
    if ((int)threadIdx.x < a.halo.npull) st_release_sys(a.halo.pull[threadIdx.x].signal, epoch);
    if (threadIdx.x == 0) {
        a.halo.epoch[0] = epoch;   // [push counter, pull counter] of mimsem_gpu_halo_push / _pull: kept in step
        a.halo.epoch[1] = epoch;
        __threadfence();
    }
}

// Boundary tile, warp 0: wait until every peer's rows of this epoch have landed (lane i watches peer i) and return the
// inbox copy to stage from.  Out of line: cold code stays out of the tile loop.
__device__ __noinline__ const double* halo_wait_peers(const TArgs& a, unsigned long long epoch, int lane) {
    if (!(a.debug & 8)) {   // debug bit 3: do not wait for the peers (timing experiment)
#pragma unroll 1
        for (int i = lane; i < a.halo.npull; i += 32) spin_until(a.halo.pull[i].wait, epoch, a.halo.err);
    }
    __syncwarp();
    fence_async_all();   // the peers' generic-proxy stores are read by the async proxy (TMA) next
    return a.halo.inbox + (epoch % a.halo.nbuf) * a.halo.parity_stride;
}

// Stage one tile: warp 0 walks the element's copy list (one entry per lane and round)
template <bool HALO>
__device__ __forceinline__ void tile_load(const TArgs& a, int e, int tile_i, unsigned long long epoch, uint64_t* bar, double* geo,
                                          double* tile, int lane_in = -1) {
    const TileHdr* rec = a.recs + (size_t)e * (1 + a.rec_ents);
    const CopyEnt* ents = reinterpret_cast<const CopyEnt*>(rec + 1);
    const int lane = lane_in >= 0 ? lane_in : (int)threadIdx.x;
    // header and this lane's first entry are fetched together (no dependent load on the critical path)
    const TileHdr h = rec[0];
    CopyEnt first = ents[lane < a.rec_ents ? lane : 0];
    const unsigned slot_bytes = (unsigned)a.nlev * 8u;
    const bool with_t = a.tpow > 0;
    int skipped = 0;
    const double* inbox = nullptr;
    if (HALO && tile_i >= a.halo.n_int) {
        // boundary tile: the peers' rows of this epoch must have landed in my inbox (lane i watches peer i)
        inbox = halo_wait_peers(a, epoch, lane);
    }
    for (int ci = lane; ci < h.cp_count; ci += 32) {
        const CopyEnt c = (ci == lane) ? first : ents[ci];
        if (c.kind == 2 && !with_t) continue;
        if ((a.debug & 2) && c.kind == 0 && c.slot >= a.debug_slot_lo && c.slot < a.debug_slot_hi) {   // traffic experiment only
            skipped += c.count;
            continue;
        }
        if (c.kind == 3) {
            bulk_g2s(geo, a.geo + (size_t)c.src * a.geo_doubles, (unsigned)a.geo_doubles * 8u, bar);
        } else if (HALO && c.kind == 4) {
            // ghost rows of x, straight from the inbox (rows packed with stride nlev)
            bulk_g2s(tile + (size_t)c.slot * a.nlev, inbox + (size_t)c.src * a.nlev, slot_bytes * c.count, bar);
        } else if (c.kind == 2) {
            const double* src = a.tinv + (size_t)c.src * a.nkT + a.lev0;
            double* dst = tile + (size_t)c.slot * a.nlev;
            if (a.contig_t) bulk_g2s(dst, src, slot_bytes * c.count, bar);
            else
                for (int j = 0; j < c.count; j++) bulk_g2s(dst + (size_t)j * a.nlev, src + (size_t)j * a.nkT, slot_bytes, bar);
        } else {
            const double* src = (c.kind == 0 ? a.x : a.c) + (size_t)c.src * a.ld;
            double* dst = tile + (size_t)c.slot * a.nlev;
            if (a.contig_x) bulk_g2s(dst, src, slot_bytes * c.count, bar);
            else
                for (int j = 0; j < c.count; j++) bulk_g2s(dst + (size_t)j * a.nlev, src + (size_t)j * a.ld, slot_bytes, bar);
        }
    }
    // nslots = slots filled by x / coefficient entries (low 16 bits) and by thickness entries (high 16 bits)
    unsigned nslots = (unsigned)(h.nslots & 0xffff) + (with_t ? (unsigned)(h.nslots >> 16) : 0u);
    if (a.debug & 2) {
        for (int o = 16; o > 0; o >>= 1) skipped += __shfl_xor_sync(0xffffffffu, skipped, o);
        nslots -= (unsigned)skipped;
    }
    if (lane == 0) mbar_arrive_expect_tx(bar, nslots * slot_bytes + (unsigned)a.geo_doubles * 8u);
}

// Contribution of the west (SIDE 0) / south (SIDE 1) neighbour's far GLL line to my P shared edges.
// REV: the neighbour numbers the shared edges in the opposite direction (rotated cubed-sphere seam).
template <int P, bool WITH_H, int SIDE, bool REV, int NL>
__device__ __forceinline__ void tile_far_line(const TArgs& a, const double* col, const double* geo, bool far_is_row,
                                              double (&cfar)[P]) {
    using S = M1Slots<P>;
    constexpr int NP1 = P + 1;
    const int nl = NL ? NL : a.nlev;
#define SLOT(s) col[(size_t)(s) * nl]
    constexpr int OTH = SIDE == 0 ? S::WOTH : S::SOTH;
    constexpr int OWN = SIDE == 0 ? S::OX : S::OY;   // west column xx(0,iy) -> OX+iy ; south row xy(ix,0) -> OY+ix
    const double* gf = geo + (SIDE == 0 ? S::GW : S::GS);
    double own[P];   // the shared edges in the neighbour's order
#pragma unroll
    for (int j = 0; j < P; j++) own[j] = SLOT(OWN + (REV ? P - 1 - j : j));
    double hs[P];
    if (WITH_H) {
        // neighbour's h contracted across its far line (east column: over ix; north row: over iy)
        constexpr int HN = SIDE == 0 ? S::HW : S::HS;
#pragma unroll
        for (int j = 0; j < P; j++) hs[j] = 0.0;
#pragma unroll
        for (int iy = 0; iy < P; iy++)
#pragma unroll
            for (int ix = 0; ix < P; ix++) {
                const double hv = SLOT(HN + iy * P + ix);
                if (!far_is_row) hs[iy] += a.E[P * P + ix] * hv;
                else hs[ix] += a.E[P * P + iy] * hv;
            }
    }
    double f[P + 1];
#pragma unroll
    for (int q = 0; q <= P; q++) {
        double ua = 0.0, ub = 0.0;
#pragma unroll
        for (int j = 0; j < P; j++) ua += a.E[q * P + j] * own[j];
#pragma unroll
        for (int t = 0; t < P; t++) ub += a.E[P * P + t] * SLOT(OTH + q * P + t);
        // the far line's quadrature points are my own west column / south row points
        constexpr int dummy = 0;
        (void)dummy;
        const int qm = REV ? P - q : q;
        double c = a.scale;
        if (a.tpow > 0) {
            const double t = SLOT(S::T + (SIDE == 0 ? qm * NP1 : qm));
            c *= t;
            if (a.tpow > 1) c *= t;
        }
        if (WITH_H) {
            double hl = 0.0;
#pragma unroll
            for (int j = 0; j < P; j++) hl += a.E[q * P + j] * hs[j];
            c *= hl;
        }
        f[q] = c * (gf[q * 2 + 0] * ua + gf[q * 2 + 1] * ub);
    }
#pragma unroll
    for (int j = 0; j < P; j++) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q <= P; q++) s += a.E[q * P + j] * f[q];
        cfar[REV ? P - 1 - j : j] = s;   // neighbour's edge j is my edge (REV ? P-1-j : j)
    }
#undef SLOT
}

// L2 prefetch of the DRAM-unique part of a LATER tile (its own edge block and its thickness rows): by the time that
// tile's CTA starts, its bulk loads hit L2, which takes the HBM latency out of the per-tile critical path.
__device__ __forceinline__ void tile_prefetch(const TArgs& a, int e) {
    const TileHdr* rec = a.recs + (size_t)e * (1 + a.rec_ents);
    const CopyEnt* ents = reinterpret_cast<const CopyEnt*>(rec + 1);
    const TileHdr h = rec[0];
    const int lane = threadIdx.x & 31;
    const unsigned slot_bytes = (unsigned)a.nlev * 8u;
    const int own_slots = a.prefetch_own_slots;
    for (int ci = lane; ci < h.cp_count; ci += 32) {
        const CopyEnt c = ents[ci];
        if (c.kind == 3) {
            bulk_prefetch_l2(a.geo + (size_t)c.src * a.geo_doubles, (unsigned)a.geo_doubles * 8u);
        } else if (c.kind == 2 && a.tpow > 0 && a.contig_t) {
            bulk_prefetch_l2(a.tinv + (size_t)c.src * a.nkT + a.lev0, slot_bytes * c.count);
        } else if ((c.kind == 0 && c.slot < own_slots && a.contig_x) || (c.kind == 1 && a.contig_x)) {
            bulk_prefetch_l2((c.kind == 0 ? a.x : a.c) + (size_t)c.src * a.ld, slot_bytes * c.count);
        }
    }
}

// One warp-pair's share of an element tile: DIR 0 = x-normal edges of GLL columns [LO,HI), DIR 1 = y-normal edges of
// GLL rows [LO,HI); HALF 0 additionally gathers the west (DIR 0) / south (DIR 1) neighbour's far line.
template <int P, bool WITH_H, int NL, int DIR, int HALF>
__device__ __forceinline__ void tile_compute(const TArgs& a, const double* col, const double* geo, int flags, int st_dof, int k) {
    using S = M1Slots<P>;
    constexpr int NP1 = P + 1;
    // HALF 2 = all lines of the direction (two warp-pairs per tile: the configuration that measured fastest)
    constexpr int LO = HALF == 1 ? P / 2 : 0;
    constexpr int HI = HALF == 0 ? P / 2 : P;
    constexpr int NLN = HI - LO > 0 ? HI - LO : 1;
    const int nl = NL ? NL : a.nlev;
#define SLOT(s) col[(size_t)(s) * nl]
    auto tf = [&](int q) {
        double f = a.scale;
        if (a.tpow > 0) {
            const double t = SLOT(S::T + q);
            f *= t;
            if (a.tpow > 1) f *= t;
        }
        return f;
    };
    double cfar[P];
#pragma unroll
    for (int j = 0; j < P; j++) cfar[j] = 0.0;
    if (HALF != 1) {
        const int has = DIR == 0 ? 1 : 4, rv = DIR == 0 ? 2 : 8, row = DIR == 0 ? 16 : 32;
        if (flags & has) {
            if (flags & rv) tile_far_line<P, WITH_H, DIR, true, NL>(a, col, geo, flags & row, cfar);
            else tile_far_line<P, WITH_H, DIR, false, NL>(a, col, geo, flags & row, cfar);
        }
    }
    // the other family's edges oth(q,t) = xy(ix=t, qy=q) (DIR 0) / xx(qx=q, iy=t) (DIR 1) are read from shared memory at
    // their point of use (keeping all (P+1)P of them in registers would halve the number of resident CTAs)
#define OTH_SMEM(q, t) ((DIR == 0) ? ((q) < P ? SLOT(S::OY + (q) * P + (t)) : SLOT(S::YN + (t))) : ((q) < P ? SLOT(S::OX + (q) * P + (t)) : SLOT(S::XE + (t))))
    // with two warp-pairs per tile the (P+1)P values are used P times each: keep them in registers
    constexpr bool OTH_REGS = (HALF == 2);
    double othr[OTH_REGS ? P + 1 : 1][OTH_REGS ? P : 1];
    if (OTH_REGS) {
#pragma unroll
        for (int q = 0; q <= P; q++)
#pragma unroll
            for (int t = 0; t < P; t++) othr[OTH_REGS ? q : 0][OTH_REGS ? t : 0] = OTH_SMEM(q, t);
    }
#define OTH(q, t) (OTH_REGS ? othr[OTH_REGS ? (q) : 0][OTH_REGS ? (t) : 0] : OTH_SMEM(q, t))
    double hc[NLN][P];   // h contracted across the line direction: hc[line][j]
    if (WITH_H) {
#pragma unroll
        for (int j = 0; j < P; j++) {
            double hv[P];
#pragma unroll
            for (int t = 0; t < P; t++) hv[t] = (DIR == 0) ? SLOT(S::H + j * P + t) : SLOT(S::H + t * P + j);   // h(ix=t,iy=j) / h(ix=j,iy=t)
#pragma unroll
            for (int ln = LO; ln < HI; ln++) {
                double s = 0.0;
#pragma unroll
                for (int t = 0; t < P; t++) s += a.E[ln * P + t] * hv[t];
                hc[ln - LO][j] = s;
            }
        }
    }
    double out[NLN][P];
#pragma unroll
    for (int ln = LO; ln < HI; ln++) {
        // edges ON the line: xx(ln, iy) (DIR 0) / xy(ix, ln) (DIR 1)
        double own[P];
#pragma unroll
        for (int j = 0; j < P; j++) own[j] = SLOT((DIR == 0 ? S::OX : S::OY) + ln * P + j);
        double f[P + 1];
#pragma unroll
        for (int q = 0; q <= P; q++) {
            double ua = 0.0, ub = 0.0;   // along-line interpolation of own, across-line interpolation of oth
#pragma unroll
            for (int j = 0; j < P; j++) ua += a.E[q * P + j] * own[j];
#pragma unroll
            for (int t = 0; t < P; t++) ub += a.E[ln * P + t] * OTH(q, t);
            const int qq = (DIR == 0) ? q * NP1 + ln : ln * NP1 + q;
            double c = tf(qq);
            if (WITH_H) {
                double hl = 0.0;
#pragma unroll
                for (int j = 0; j < P; j++) hl += a.E[q * P + j] * hc[ln - LO][j];
                c *= hl;
            }
            // DIR 0: f0 = c (Gaa ul0 + Gab ul1), ul0 = ua ; DIR 1: f1 = c (Gab ul0 + Gbb ul1), ul1 = ua
            f[q] = (DIR == 0) ? c * (geo[qq * 3 + 0] * ua + geo[qq * 3 + 1] * ub) : c * (geo[qq * 3 + 1] * ub + geo[qq * 3 + 2] * ua);
        }
#pragma unroll
        for (int j = 0; j < P; j++) {
            double s = (ln == 0) ? cfar[j] : 0.0;
#pragma unroll
            for (int q = 0; q <= P; q++) s += a.E[q * P + j] * f[q];
            out[ln - LO][j] = s;
        }
    }
    // results straight from registers to global memory (lanes = levels: coalesced 8-byte stores)
    if (st_dof >= 0) {
        double* __restrict__ y = a.y + (size_t)(st_dof + (DIR == 0 ? S::OX : S::OY) + LO * P) * a.ld + k;
#pragma unroll
        for (int i = 0; i < HI - LO; i++)
#pragma unroll
            for (int j = 0; j < P; j++) y[(size_t)(i * P + j) * a.ld] = out[i][j];
    }
#undef OTH
#undef OTH_SMEM
#undef SLOT
}

// y = M1 x (WITH_H: M1(h) x).  One CTA per element; 128 threads = 2 warp-pairs (x-normal / y-normal edges) x 64 level
// lanes.  (A 4-warp-pair split -- direction x half of the lines -- was measured slower: 165 vs 121 us on C5.)
// NL = compile-time number of levels (0: runtime) so that shared-memory operands use immediate offsets.
// HALO: ghost refresh fused into the launch (see HaloFused); a separate instantiation so that the single-GPU kernel
// carries none of its code.
template <int P, bool WITH_H, int NL, bool HALO>
__global__ void __launch_bounds__(128) k_apply_m1_tma(const __grid_constant__ TArgs a) {
    using S = M1Slots<P>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    double* geo = reinterpret_cast<double*>(smem_raw + 16);
    double* tile = geo + S::GEO;
    const int part = threadIdx.x >> 6;
    const int k = threadIdx.x & 63;
    const int nl = NL ? NL : a.nlev;
    unsigned long long epoch = 0;
    int first_tile = blockIdx.x, tile_stride = gridDim.x;
    if (HALO) {
        if ((int)blockIdx.x < a.halo.push_ctas || (int)blockIdx.x - a.halo.push_ctas >= a.halo.n_int) epoch = *a.halo.epoch + 1;
        if ((int)blockIdx.x < a.halo.push_ctas) {
            halo_push_role(a, epoch + (unsigned long long)a.halo.lead);   // data epoch of the pushed field
            halo_cta_done(a, epoch);
            return;
        }
        first_tile -= a.halo.push_ctas;
        tile_stride -= a.halo.push_ctas;
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_async_smem();
    }
    __syncthreads();
    // optionally persistent: tiles first_tile, first_tile + tile_stride, ...
    const int tile_end = a.ntiles;
    unsigned phase = 0;
    for (int tile_i = first_tile; tile_i < tile_end; tile_i += tile_stride, phase ^= 1) {
        const int e = a.elist ? a.elist[tile_i] : tile_i;
        DBG_T(0);
        if (a.debug & 16) {
            // timing experiment: contraction + stores on whatever the shared memory holds (no copies, no wait)
        } else if (threadIdx.x < 32) tile_load<HALO>(a, e, tile_i, epoch, bar, geo, tile);
        else if (threadIdx.x < 64 && a.prefetch_ahead > 0 && tile_i + a.prefetch_ahead < a.ntiles) {
            const int bn = tile_i + a.prefetch_ahead;
            tile_prefetch(a, a.elist ? a.elist[bn] : bn);
        }
        const TileHdr hd = a.recs[(size_t)e * (1 + a.rec_ents)];
        DBG_T(1);
        if (!(a.debug & 16)) mbar_wait(bar, phase);
        DBG_T(2);
        if (k < nl && !(a.debug & 1)) {
            const double* col = tile + k;
            if (hd.st_dof < 0) {
                // owned rows not contiguous (generic numbering): not produced by this library's own plans
                if (threadIdx.x == 0 && blockIdx.x == 0) printf("mimsem: non-contiguous owned block is not supported by the TMA kernel\n");
            } else if (part == 0) {
                tile_compute<P, WITH_H, NL, 0, 2>(a, col, geo, hd.flags, hd.st_dof, k);
            } else {
                tile_compute<P, WITH_H, NL, 1, 2>(a, col, geo, hd.flags, hd.st_dof, k);
            }
        }
        DBG_T(3);
        DBG_T(4);
        if (tile_i + tile_stride < tile_end) __syncthreads();   // the next tile's bulk loads overwrite the buffer
    }
    if (HALO && first_tile >= a.halo.n_int) halo_cta_done(a, epoch);
}

// ---------------------------------------------------------------------------------------------
// y = K(u1) x (WtQUmat, eul/Assembly.cpp:933-986): 1-form -> 2-form, element-local.  One CTA per element, 128 threads =
// 2 parts x 64 level lanes; part 0 takes the quadrature columns qx < (P+1)/2, part 1 the rest; the two partial
// P x P results are exchanged through the (by then dead) thickness slots and each part stores half of the rows.
template <int P, int NL, int PART>
__device__ __forceinline__ void k_tile_compute(const TArgs& a, const double* col, const double* geo, double (&out)[P][P]) {
    using S = KSlots<P>;
    constexpr int NP1 = P + 1;
    constexpr int Q0 = PART == 0 ? 0 : NP1 / 2, Q1 = PART == 0 ? NP1 / 2 : NP1;
    const int nl = NL ? NL : a.nlev;
#define SLOT(s) col[(size_t)(s) * nl]
#pragma unroll
    for (int iy = 0; iy < P; iy++)
#pragma unroll
        for (int ix = 0; ix < P; ix++) out[iy][ix] = 0.0;
#pragma unroll
    for (int qx = Q0; qx < Q1; qx++) {
        double xc[P], uc[P];
#pragma unroll
        for (int iy = 0; iy < P; iy++) {
            const int s = qx < P ? S::OX + qx * P + iy : S::XE + iy;
            xc[iy] = SLOT(s);
            uc[iy] = SLOT(S::U0 + s);
        }
        double g[NP1];
#pragma unroll
        for (int qy = 0; qy <= P; qy++) {
            double x0 = 0.0, x1 = 0.0, a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int iy = 0; iy < P; iy++) {
                x0 += a.E[qy * P + iy] * xc[iy];
                a0 += a.E[qy * P + iy] * uc[iy];
            }
#pragma unroll
            for (int ix = 0; ix < P; ix++) {
                const int s = qy < P ? S::OY + qy * P + ix : S::YN + ix;
                x1 += a.E[qx * P + ix] * SLOT(s);
                a1 += a.E[qx * P + ix] * SLOT(S::U0 + s);
            }
            const int q = qy * NP1 + qx;
            double f = a.scale;
            if (a.tpow > 0) {
                const double t = SLOT(S::T + q);
                f *= t;
                if (a.tpow > 1) f *= t;
            }
            const double c = 0.5 * f;
            const double ka = geo[q * 3 + 0] * a0 + geo[q * 3 + 1] * a1;
            const double kb = geo[q * 3 + 1] * a0 + geo[q * 3 + 2] * a1;
            g[qy] = c * (ka * x0 + kb * x1);
        }
#pragma unroll
        for (int iy = 0; iy < P; iy++) {
            double b = 0.0;
#pragma unroll
            for (int qy = 0; qy <= P; qy++) b += a.E[qy * P + iy] * g[qy];
#pragma unroll
            for (int ix = 0; ix < P; ix++) out[iy][ix] += a.E[qx * P + ix] * b;
        }
    }
#undef SLOT
}

template <int P, int NL>
__global__ void __launch_bounds__(128, (P <= 4 ? 4 : 2)) k_apply_k_tma(const __grid_constant__ TArgs a) {
    using S = KSlots<P>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    double* geo = reinterpret_cast<double*>(smem_raw + 16);
    double* tile = geo + M1Slots<P>::GEO;
    const int part = threadIdx.x >> 6;
    const int k = threadIdx.x & 63;
    const int nl = NL ? NL : a.nlev;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_async_smem();
    }
    __syncthreads();
    unsigned phase = 0;
    for (int tile_i = blockIdx.x; tile_i < a.ntiles; tile_i += gridDim.x, phase ^= 1) {
        const int e = a.elist ? a.elist[tile_i] : tile_i;
        if (threadIdx.x < 32) tile_load<false>(a, e, tile_i, 0ull, bar, geo, tile);
        else if (threadIdx.x < 64 && a.prefetch_ahead > 0 && tile_i + a.prefetch_ahead < a.ntiles) {
            const int bn = tile_i + a.prefetch_ahead;
            tile_prefetch(a, a.elist ? a.elist[bn] : bn);
        }
        const TileHdr hd = a.recs[(size_t)e * (1 + a.rec_ents)];
        mbar_wait(bar, phase);
        const bool active = k < nl;
        double* col = tile + k;
        double out[P][P];
        if (active) {
            if (part == 0) k_tile_compute<P, NL, 0>(a, col, geo, out);
            else k_tile_compute<P, NL, 1>(a, col, geo, out);
        }
        __syncthreads();   // every read of the tile is done: the thickness slots become the exchange buffer
        constexpr int H0 = P / 2;   // part 0 finishes rows [0, H0), part 1 rows [H0, P)
        if (active) {
            // hand the rows the OTHER part finishes over to it (slot = row-major face index)
#pragma unroll
            for (int iy = 0; iy < P; iy++)
#pragma unroll
                for (int ix = 0; ix < P; ix++)
                    if ((iy < H0) != (part == 0)) col[(size_t)(S::T + iy * P + ix) * nl] = out[iy][ix];
        }
        __syncthreads();
        if (active) {
            double* __restrict__ y = a.y + (size_t)hd.st_dof * a.ld + k;
#pragma unroll
            for (int iy = 0; iy < P; iy++)
#pragma unroll
                for (int ix = 0; ix < P; ix++)
                    if ((iy < H0) == (part == 0)) {
                        // part 0's share (low qx) is always the first addend: the result does not depend on the part
                        const double o = col[(size_t)(S::T + iy * P + ix) * nl];
                        y[(size_t)(iy * P + ix) * a.ld] = part == 0 ? out[iy][ix] + o : o + out[iy][ix];
                    }
        }
        if (gridDim.x < (unsigned)a.ntiles) __syncthreads();
    }
}

}  // namespace mimsem
