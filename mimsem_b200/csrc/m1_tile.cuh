// y = M1 x and y = M1(h) x as a TMA-staged tile kernel (Umat / Uhmat / Ut_mat, eul/Assembly.cpp:51-153, 416-474,
// 1338-1440), optionally fused with the peer-to-peer ghost refresh of its input (HaloFused, engine.cuh).
#pragma once
#include "tile_common.cuh"

namespace mimsem {

// ---------------------------------------------------------------------------------------------------------------
// Fused ghost refresh, push role.  The rows of all peers form one flat list that is split evenly over the push CTAs (a
// CTA's share may straddle peers), sized so that a CTA moves its share in a single pass: every thread has up to four
// independent 16-byte loads in flight, then stores them into the peers' inboxes over NVLink.  The copy is
// latency-bound (index load -> HBM load -> remote store -> system fence), hence many small CTAs rather than few big ones.
static __device__ __noinline__ void halo_push_role(const TArgs& a, unsigned long long epoch) {
    const HaloFused& h = a.halo;
    __shared__ HaloPeer peers[kMaxPushPeers];
    __shared__ int pre[kMaxPushPeers + 1];
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
    const int nl2 = a.nlev >> 1;
    const int R = pre[h.npush];
    const int f0 = (int)(((long long)R * blockIdx.x) / h.push_ctas);
    const int f1 = (int)(((long long)R * (blockIdx.x + 1)) / h.push_ctas);
    const int total = (f1 - f0) * nl2;
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
    // the CTA barrier orders every thread's stores before the npush threads below; their system-scope fences are
    // cumulative, so only they (not all 128 threads) pay the NVLink round trip
    __syncthreads();
    if (threadIdx.x < h.npush) {
        const int p = threadIdx.x;
        __threadfence_system();
        unsigned* ctr = h.counters + h.burst_pos * kCtrStride;
        const unsigned done = atomicAdd(&ctr[1 + p], 1u);
        if (done == (unsigned)h.push_ctas - 1) {
            ctr[1 + p] = 0;
            if (h.burst_len > 1) {
                // launches of a burst overlap: flags must still rise in launch order (a receiver that sees flag e takes
                // every epoch <= e for delivered)
                volatile unsigned* sq = h.seq + 1 + p;
                while (*sq != (unsigned)h.burst_pos) __nanosleep(20);
            }
            __threadfence_system();
            st_release_sys(peers[p].signal, epoch);
            if (h.burst_len > 1) {
                __threadfence();
                h.seq[1 + p] = h.burst_pos == h.burst_len - 1 ? 0u : (unsigned)h.burst_pos + 1u;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// In-band ("LL") variant of the fused ghost refresh.  A pushed value travels as ONE 16-byte cell
//   { lo32(value), epoch32, hi32(value), epoch32 }
// so that the data validates itself: the receiver polls the cell until both flag words carry the epoch it expects
// (each 8-byte half is written atomically, so a torn cell is never mistaken for a complete one).  Nothing else is
// needed to hand the rows over -- no system fence after the copy, no per-peer flag, no counter of finished push
// CTAs -- which takes the push off the critical path of a step: the rows are visible one NVLink traversal after
// the stores were issued instead of 15-20 us into the launch.  The inbox is never reset: cells of the copy that is
// overwritten (data epoch e - nbuf) carry a different epoch.  Cost: twice the NVLink bytes (still < 2 MB per rank
// and step on C5 at 8 GPUs) and no TMA on the receiving side (the consumer unpacks cells with ordinary loads).
__device__ __forceinline__ void ll_store(uint4* cell, double v, unsigned ep) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(cell), "r"((unsigned)b), "r"(ep), "r"((unsigned)(b >> 32)), "r"(ep)
                 : "memory");
}
// poll a cell until it carries epoch `ep`; gives up after ~2 s (a dead peer must not hang the GPU) and records the failure
__device__ __forceinline__ double ll_load(const uint4* cell, unsigned ep, int* err) {
    unsigned x, f0, y, f1;
    long long t0 = 0;
    for (unsigned spin = 0;; spin++) {
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(f0), "=r"(y), "=r"(f1) : "l"(cell) : "memory");
        if (f0 == ep && f1 == ep) break;
        if ((spin & 1023u) == 1023u) {
            const long long t = clock64();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 4000000000ll) {
                atomicExch(err, 1);
                break;
            }
        }
    }
    return __longlong_as_double((long long)(((unsigned long long)y << 32) | (unsigned long long)x));
}

// N cells at once: all loads are issued before the first one is examined (memory-level parallelism; a cell-by-cell poll
// serialises N round trips to L2 / HBM), then only the cells that had not landed yet are polled again.
template <int N>
__device__ __forceinline__ void ll_load_many(const uint4* const (&cell)[N], unsigned ep, int* err, double (&out)[N]) {
    unsigned x[N], f0[N], y[N], f1[N];
#pragma unroll
    for (int i = 0; i < N; i++)
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x[i]), "=r"(f0[i]), "=r"(y[i]), "=r"(f1[i]) : "l"(cell[i]) : "memory");
#pragma unroll
    for (int i = 0; i < N; i++) {
        if (f0[i] == ep && f1[i] == ep) out[i] = __longlong_as_double((long long)(((unsigned long long)y[i] << 32) | (unsigned long long)x[i]));
        else out[i] = ll_load(cell[i], ep, err);
    }
}

static __device__ __noinline__ void halo_push_role_ll(const TArgs& a, unsigned long long epoch) {
    const HaloFused& h = a.halo;
    __shared__ HaloPeer peers[kMaxPushPeers];
    __shared__ int pre[kMaxPushPeers + 1];
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
    // one VALUE per lane and store: consecutive lanes write consecutive 16-byte cells, i.e. every warp store is 512
    // contiguous bytes = four full 128-byte lines on the wire (two cells per lane left 16-byte holes in every sector of
    // each store and cost more NVLink time than the whole tile work of a step)
    const int nl = a.nlev;
    const int R = pre[h.npush];
    const int f0 = (int)(((long long)R * blockIdx.x) / h.push_ctas);
    const int f1 = (int)(((long long)R * (blockIdx.x + 1)) / h.push_ctas);
    const int total = (f1 - f0) * nl;
    const unsigned ep = (unsigned)epoch;
    constexpr int U = 8;
    for (int base = threadIdx.x; base < total; base += U * blockDim.x) {
        double v[U];
        uint4* dst[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int i = base + u * blockDim.x;
            dst[u] = nullptr;
            if (i < total) {
                const int f = f0 + i / nl, k = i - (f - f0) * nl;
                int p = 0;
                while (f >= pre[p + 1]) p++;
                const int r = f - pre[p];
                const HaloPeer& pp = peers[p];
                v[u] = __ldg(h.x_push + (size_t)pp.rows[r] * a.ld + k);
                // LL inbox of the peer: 16-byte cells, [copy][row][level]
                dst[u] = reinterpret_cast<uint4*>(pp.inbox) + (epoch % h.nbuf) * pp.inbox_parity_stride + (size_t)(pp.row0 + r) * a.nlev + k;
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            if (dst[u]) ll_store(dst[u], v[u], ep);
    }
}

// Every push CTA and every boundary tile of a fused launch ends here; the last one acknowledges the inbox of this epoch
// to the peers (all reads from it have completed) and advances the epoch for the next launch / graph replay.
static __device__ __noinline__ void halo_cta_done(const TArgs& a, unsigned long long epoch) {
    __shared__ int last;
    __syncthreads();
    const bool burst = a.halo.burst_len > 1;
    if (threadIdx.x == 0) {
        unsigned* ctr = a.halo.counters + a.halo.burst_pos * kCtrStride;
        const unsigned done = atomicAdd(&ctr[0], 1u);
        last = (done == (unsigned)(a.halo.push_ctas + a.ntiles - a.halo.n_int) - 1) ? 1 : 0;
        if (last) {
            ctr[0] = 0;
            if (burst) {
                // acknowledgements (and the epoch word) in launch order
                volatile unsigned* sq = a.halo.seq;
                while (*sq != (unsigned)a.halo.burst_pos) __nanosleep(20);
            }
        }
    }
    __syncthreads();
    if (!last || a.halo.push_only) return;
    // one thread per peer: the release stores cross NVLink concurrently (a serial loop costs a round trip per peer)
    if ((int)threadIdx.x < a.halo.npull) st_release_sys(a.halo.pull[threadIdx.x].signal, epoch);
    if (burst) __syncthreads();
    if (threadIdx.x == 0) {
        if (!burst || a.halo.burst_pos == a.halo.burst_len - 1) {
            a.halo.epoch[0] = epoch;   // [push counter, pull counter] of mimsem_gpu_halo_push / _pull: kept in step
            a.halo.epoch[1] = epoch;
        }
        __threadfence();
        if (burst) a.halo.seq[0] = a.halo.burst_pos == a.halo.burst_len - 1 ? 0u : (unsigned)a.halo.burst_pos + 1u;
    }
}

// Boundary tile, warp 0: wait until every peer's rows of this epoch have landed (lane i watches peer i).
// Out of line: cold code stays out of the tile body.
static __device__ __noinline__ void halo_wait_peers(const TArgs& a, unsigned long long epoch, int lane) {
#pragma unroll 1
    for (int i = lane; i < a.halo.npull; i += 32) spin_until(a.halo.pull[i].wait, epoch, a.halo.err);
    __syncwarp();
    fence_async_all();   // the peers' generic-proxy stores are read by the async proxy (TMA) next
}

// ---------------------------------------------------------------------------------------------------------------
// Far-line operands, straight from global memory into registers (side 0: the west neighbour behind my west column,
// side 1: the south neighbour behind my south row).  Of the neighbour's other edge family oth(q,t) only its
// interpolation onto the far line is needed,  ubf[q] = sum_t E[P][t] oth(q,t)  (and of its 2-form coefficient only
// hs[j] = sum_t E[P][t] h(t across, j along)), so (P+1)P + P^2 loads collapse into 2P+1 registers that stay live
// across the wait for the bulk copies.  Lanes are levels: every load is a coalesced run of the row.
template <int P, bool WITH_H, int HALO>
__device__ __forceinline__ void far_fetch(const TArgs& a, const TileHdr* rec, int flags, int side, const double* inbox, unsigned ep, int k,
                                          double (&ubf)[P + 1], double (&hs)[P]) {
    const int has = side ? TF_HAS_S : TF_HAS_W;
#pragma unroll
    for (int q = 0; q <= P; q++) ubf[q] = 0.0;
#pragma unroll
    for (int j = 0; j < P; j++) hs[j] = 0.0;
    if (!(flags & has)) return;
    // rows of oth(q,t): two runs (the neighbour's own block, then the far row / column owned by ITS neighbour), or an
    // explicit list (ghost rows of a fused launch live in the inbox: entries < 0; generic numberings)
    constexpr int NF = (P + 1) * P;
    double oth[P + 1][P];
    const double* xk = a.x + k;
    if (!(flags & (side ? TF_LIST_S : TF_LIST_W))) {
        const TileFar fr = *reinterpret_cast<const TileFar*>(rec + 1);
        const double* b16 = xk + (size_t)(side ? fr.s16 : fr.w16) * a.ld;
        const double* b4 = xk + (size_t)(side ? fr.s4 : fr.w4) * a.ld;
        const long long st4 = (flags & (side ? TF_S4_DESC : TF_W4_DESC)) ? -(long long)a.ld : (long long)a.ld;
#pragma unroll
        for (int q = 0; q < P; q++)
#pragma unroll
            for (int t = 0; t < P; t++) oth[q][t] = b16[(size_t)(q * P + t) * a.ld];
#pragma unroll
        for (int t = 0; t < P; t++) oth[P][t] = b4[t * st4];
    } else {
        int rows[NF];
        const int* list = reinterpret_cast<const int*>(rec + a.rec_list) + side * NF;
#pragma unroll
        for (int i = 0; i < NF; i++) rows[i] = list[i];
        if (HALO == 2 && rows[0] < 0 && rows[NF - 1] < 0) {
            // every row of this side lives in the cell inbox (the common case of a ghost far line): two batches of loads
            constexpr int H0 = NF / 2, H1 = NF - H0;
            const uint4* cb = reinterpret_cast<const uint4*>(inbox) + k;
            double* of = &oth[0][0];
            {
                const uint4* cell[H0];
                double v[H0];
#pragma unroll
                for (int i = 0; i < H0; i++) cell[i] = rows[i] < 0 ? cb + (size_t)(-rows[i] - 1) * a.nlev : nullptr;
                bool all = true;
#pragma unroll
                for (int i = 0; i < H0; i++) all = all && cell[i];
                if (all) {
                    ll_load_many<H0>(cell, ep, a.halo.err, v);
#pragma unroll
                    for (int i = 0; i < H0; i++) of[i] = v[i];
                } else {
#pragma unroll
                    for (int i = 0; i < H0; i++) of[i] = cell[i] ? ll_load(cell[i], ep, a.halo.err) : xk[(size_t)rows[i] * a.ld];
                }
            }
            {
                const uint4* cell[H1];
                double v[H1];
#pragma unroll
                for (int i = 0; i < H1; i++) cell[i] = rows[H0 + i] < 0 ? cb + (size_t)(-rows[H0 + i] - 1) * a.nlev : nullptr;
                bool all = true;
#pragma unroll
                for (int i = 0; i < H1; i++) all = all && cell[i];
                if (all) {
                    ll_load_many<H1>(cell, ep, a.halo.err, v);
#pragma unroll
                    for (int i = 0; i < H1; i++) of[H0 + i] = v[i];
                } else {
#pragma unroll
                    for (int i = 0; i < H1; i++) of[H0 + i] = cell[i] ? ll_load(cell[i], ep, a.halo.err) : xk[(size_t)rows[H0 + i] * a.ld];
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q <= P; q++)
#pragma unroll
                for (int t = 0; t < P; t++) {
                    const int r = rows[q * P + t];
                    if (HALO == 1 && r < 0) oth[q][t] = __ldcg(inbox + (size_t)(-r - 1) * a.nlev + k);
                    else if (HALO == 2 && r < 0) oth[q][t] = ll_load(reinterpret_cast<const uint4*>(inbox) + (size_t)(-r - 1) * a.nlev + k, ep, a.halo.err);
                    else oth[q][t] = xk[(size_t)r * a.ld];
                }
        }
    }
    double hv[WITH_H ? P : 1][WITH_H ? P : 1];
    if (WITH_H) {
        const TileFarH fh = *reinterpret_cast<const TileFarH*>(rec + 2);
        const double* hb = a.c + (size_t)(side ? fh.hs : fh.hw) * a.ld + k;
#pragma unroll
        for (int iy = 0; iy < P; iy++)
#pragma unroll
            for (int ix = 0; ix < P; ix++) hv[WITH_H ? iy : 0][WITH_H ? ix : 0] = hb[(size_t)(iy * P + ix) * a.ld];
    }
#pragma unroll
    for (int q = 0; q <= P; q++) {
        double s = 0.0;
#pragma unroll
        for (int t = 0; t < P; t++) s += a.E[P * P + t] * oth[q][t];
        ubf[q] = s;
    }
    if (WITH_H) {
        // neighbour's h contracted across its far line (east column: over ix; north row: over iy)
        const bool far_is_row = flags & (side ? TF_ROW_S : TF_ROW_W);
#pragma unroll
        for (int j = 0; j < P; j++) {
            double s = 0.0;
#pragma unroll
            for (int t = 0; t < P; t++) s += a.E[P * P + t] * (far_is_row ? hv[WITH_H ? t : 0][WITH_H ? j : 0] : hv[WITH_H ? j : 0][WITH_H ? t : 0]);
            hs[j] = s;
        }
    }
}

// Contribution of the neighbour's far GLL line to my P shared edges (one routine for both sides and both orientations:
// side and `rev` only select shared-memory strides).  rev: the neighbour numbers the shared edges in the opposite
// direction (rotated cubed-sphere seam).
template <int P, bool WITH_H, int NL, int TPOW>
__device__ __forceinline__ void far_line(const TArgs& a, const double* col, const double* geo, int flags, int side,
                                         const double (&ubf)[P + 1], const double (&hs)[P], double (&cfar)[P]) {
    using S = M1Slots<P>;
    constexpr int NP1 = P + 1;
    const int nl = NL ? NL : a.nlev;
#pragma unroll
    for (int j = 0; j < P; j++) cfar[j] = 0.0;
    if (!(flags & (side ? TF_HAS_S : TF_HAS_W))) return;
    const bool rev = flags & (side ? TF_REV_S : TF_REV_W);
    // the shared edges in the neighbour's order: my west column xx(0,iy) = OX + iy, my south row xy(ix,0) = OY + ix
    const double* ownp = col + (size_t)((side ? S::OY : S::OX) + (rev ? P - 1 : 0)) * nl;
    const int ostep = rev ? -nl : nl;
    // the far line's quadrature points are my own west column (T + q (P+1)) / south row (T + q) points
    const int tq = side ? 1 : NP1;
    const double* tp = col + (size_t)(S::T + (rev ? P * tq : 0)) * nl;
    const int tstep = (rev ? -tq : tq) * nl;
    const double* gf = geo + (side ? S::GS : S::GW);
    double own[P];
#pragma unroll
    for (int j = 0; j < P; j++) own[j] = ownp[j * ostep];
    double f[P + 1];
#pragma unroll
    for (int q = 0; q <= P; q++) {
        double ua = 0.0;
#pragma unroll
        for (int j = 0; j < P; j++) ua += a.E[q * P + j] * own[j];
        // (the operator's scale factor rides on the final contraction: a.Es = scale * E)
        double g = gf[q * 2 + 0] * ua + gf[q * 2 + 1] * ubf[q];
        const int tpw = TPOW >= 0 ? TPOW : a.tpow;
        if (tpw > 0) {
            const double t = tp[q * tstep];
            g *= t;
            if (tpw > 1) g *= t;
        }
        if (WITH_H) {
            double hl = 0.0;
#pragma unroll
            for (int j = 0; j < P; j++) hl += a.E[q * P + j] * hs[j];
            g *= hl;
        }
        f[q] = g;
    }
    double s[P];
#pragma unroll
    for (int j = 0; j < P; j++) {
        double v = 0.0;
#pragma unroll
        for (int q = 0; q <= P; q++) v += a.Es[q * P + j] * f[q];
        s[j] = v;
    }
    // the neighbour's edge j is my edge (rev ? P-1-j : j)
#pragma unroll
    for (int j = 0; j < P; j++) cfar[j] = rev ? s[P - 1 - j] : s[j];
}

// M1(h): the point factor t^tpow * hl(q), hl = the element's 2-form coefficient interpolated to the quadrature point
// (interp2_g without its 1/det, which sits in the geometry record), is the same for both edge directions.  The two
// warp-pairs of a tile tabulate it ONCE -- part 0 the GLL columns qx < SPLIT, part 1 the others, sum-factorised --
// and leave it in the thickness slots, so that the line contraction below is exactly the plain M1 one (one multiply per
// point, no coefficient data in registers).  Runs between two CTA barriers, after the far lines have read the raw
// thickness of the west column / south row.
// CARRY (persistent kernel): also returns hsn[j] = sum_t E[P][t] h(ix = t, iy = j), the coefficient contracted across this
// element's east column -- what far_fetch computes for the WEST far line of the element's east neighbour.
template <int P, int NL, int TPOW, bool CARRY = false>
__device__ __forceinline__ void h_prepass(const TArgs& a, double* col, int part, double* hsn = nullptr) {
    using S = M1Slots<P>;
    constexpr int NP1 = P + 1, SPLIT = (NP1 + 1) / 2;
    const int nl = NL ? NL : a.nlev;
    const int tpw = TPOW >= 0 ? TPOW : a.tpow;
    double hv[P][P];
#pragma unroll
    for (int iy = 0; iy < P; iy++)
#pragma unroll
        for (int ix = 0; ix < P; ix++) hv[iy][ix] = col[(size_t)(S::H + iy * P + ix) * nl];
    if (CARRY) {
#pragma unroll
        for (int j = 0; j < P; j++) {
            double s = 0.0;
#pragma unroll
            for (int t = 0; t < P; t++) s += a.E[P * P + t] * hv[j][t];
            hsn[j] = s;
        }
    }
#pragma unroll
    for (int qx = 0; qx <= P; qx++) {
        if ((qx < SPLIT) != (part == 0)) continue;
        double ax[P];
#pragma unroll
        for (int iy = 0; iy < P; iy++) {
            double s = 0.0;
#pragma unroll
            for (int ix = 0; ix < P; ix++) s += a.E[qx * P + ix] * hv[iy][ix];
            ax[iy] = s;
        }
#pragma unroll
        for (int qy = 0; qy <= P; qy++) {
            double hl = 0.0;
#pragma unroll
            for (int iy = 0; iy < P; iy++) hl += a.E[qy * P + iy] * ax[iy];
            double* tq = col + (size_t)(S::T + qy * NP1 + qx) * nl;
            if (tpw > 0) {
                const double t = *tq;
                hl *= t;
                if (tpw > 1) hl *= t;
            }
            *tq = hl;
        }
    }
}

// One warp-pair's share of an element tile: part 0 = the x-normal edges of GLL columns 0..P-1, part 1 = the y-normal
// edges of GLL rows 0..P-1; line 0 additionally receives the neighbour's far-line contribution cfar.  ONE code path
// serves both directions -- `part` only selects base pointers and two shared-memory strides -- so that all four warps
// of a tile run the same instructions (the unrolled contraction of one direction is ~12 KB of SASS; two copies did
// not fit the 32 KB instruction cache and 14 % of the warp samples were waiting for instructions).  The geometry
// record is laid out per part for this, gl[part][line][q] = (g_own, g_oth):
//   part 0: f0 = c (Gaa ul0 + Gab ul1), ul0 = ua ; part 1: f1 = c (Gbb ul1 + Gab ul0), ul1 = ua.
// Each line is stored as soon as it is finished (lanes = levels: coalesced 8-byte stores straight from registers).
// CARRY (persistent kernel): also returns ubn[q] = sum_t E[P][t] oth(q,t), this element's other edge family interpolated onto
// its east column (part 0) / north row (part 1) -- what far_fetch computes for the far line of the east / north neighbour.
template <int P, bool WITH_H, int NL, int TPOW, bool CARRY = false>
__device__ __forceinline__ void tile_lines(const TArgs& a, const double* col, const double* geo, int part, const double (&cfar)[P],
                                           double* __restrict__ y, double* ubn = nullptr) {
    using S = M1Slots<P>;
    constexpr int NP1 = P + 1;
    const int nl = NL ? NL : a.nlev;
    const double* ownp = col + (size_t)(part ? S::OY : S::OX) * nl;      // edges ON line ln: ownp[(ln P + j) nl]
    const double* oth16 = col + (size_t)(part ? S::OX : S::OY) * nl;     // the other family: oth(q < P, t) = oth16[(q P + t) nl]
    const double* oth4 = col + (size_t)(part ? S::XE : S::YN) * nl;      //                   oth(P, t)     = oth4[t nl]
    const double* tp = col + (size_t)S::T * nl;                          // thickness at point (line ln, q): tp[q sq + ln sl]
    const int sq = (part ? 1 : NP1) * nl, sl = (part ? NP1 : 1) * nl;
    const double* gl = geo + S::GL + part * (P * NP1 * 2);
    // the other family's edges oth(q,t) = xy(ix=t, qy=q) (part 0) / xx(qx=q, iy=t) (part 1): used P times each, kept in registers
    // (reading them from shared memory at every use to fit 6-7 tiles per SM was measured slower: 0.72 / 0.66 of the HBM peak)
    double othr[P + 1][P];
#pragma unroll
    for (int q = 0; q <= P; q++)
#pragma unroll
        for (int t = 0; t < P; t++) othr[q][t] = q < P ? oth16[(size_t)(q * P + t) * nl] : oth4[(size_t)t * nl];
    if (CARRY) {
#pragma unroll
        for (int q = 0; q <= P; q++) {
            double s = 0.0;
#pragma unroll
            for (int t = 0; t < P; t++) s += a.E[P * P + t] * othr[q][t];
            ubn[q] = s;
        }
    }
#pragma unroll
    for (int ln = 0; ln < P; ln++) {
        double own[P];
#pragma unroll
        for (int j = 0; j < P; j++) own[j] = ownp[(size_t)(ln * P + j) * nl];
        double f[P + 1];
#pragma unroll
        for (int q = 0; q <= P; q++) {
            double ua = 0.0, ub = 0.0;   // along-line interpolation of own, across-line interpolation of oth
#pragma unroll
            for (int j = 0; j < P; j++) ua += a.E[q * P + j] * own[j];
#pragma unroll
            for (int t = 0; t < P; t++) ub += a.E[ln * P + t] * othr[q][t];
            double g = gl[(ln * NP1 + q) * 2 + 0] * ua + gl[(ln * NP1 + q) * 2 + 1] * ub;
            if (WITH_H) {
                g *= tp[q * sq + ln * sl];   // h_prepass left t^tpow * hl(q) in the thickness slot
            } else {
                const int tpw = TPOW >= 0 ? TPOW : a.tpow;
                if (tpw > 0) {
                    const double t = tp[q * sq + ln * sl];
                    g *= t;
                    if (tpw > 1) g *= t;
                }
            }
            f[q] = g;
        }
#pragma unroll
        for (int j = 0; j < P; j++) {
            double s = (ln == 0) ? cfar[j] : 0.0;
#pragma unroll
            for (int q = 0; q <= P; q++) s += a.Es[q * P + j] * f[q];
            y[(size_t)(ln * P + j) * a.ld] = s;
        }
    }
}

// y = M1 x (WITH_H: M1(h) x).  One CTA per element; 128 threads = 2 warp-pairs (x-normal / y-normal edges) x 64 level
// lanes.  NL = compile-time number of levels (0: runtime) so that shared-memory operands use immediate offsets.
// HALO: ghost refresh fused into the launch (see HaloFused; 1 = flag protocol, 2 = in-band LL cells); separate
// instantiations so that the single-GPU kernel
// carries none of its code.  MINB = resident CTAs per SM the register allocation is budgeted for.  TPOW = compile-time
// number of thickness factors (-1: runtime) for the BASELINE shapes.
template <int P, bool WITH_H, int NL, int HALO, int MINB, int TPOW>
__global__ void __launch_bounds__(128, MINB) k_apply_m1_tile(const __grid_constant__ TArgs a) {
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
        if (a.pdl) pdl_launch_dependents();
        if ((int)blockIdx.x < a.halo.push_ctas || (int)blockIdx.x - a.halo.push_ctas >= a.halo.n_int) {
            // push CTAs and boundary tiles share the epoch word and the counters with the previous launch: under programmatic
            // dependent launch they wait for it (interior tiles, no ghost rows, run on while it drains) -- unless the launch
            // is part of a burst, whose launches have their own epoch offsets and counter slots and overlap completely
            if (a.pdl && a.halo.burst_len <= 1) pdl_wait_primary();
            epoch = *reinterpret_cast<volatile unsigned long long*>(a.halo.epoch) + 1 + (unsigned long long)a.halo.burst_pos;
        }
        if ((int)blockIdx.x < a.halo.push_ctas) {
            // data epoch of the pushed field = epoch + lead
            if (HALO == 2) halo_push_role_ll(a, epoch + (unsigned long long)a.halo.lead);
            else halo_push_role(a, epoch + (unsigned long long)a.halo.lead);
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
    if (!HALO && a.pdl) pdl_launch_dependents();
    __syncthreads();
    unsigned phase = 0;
    for (int tile_i = first_tile; tile_i < a.ntiles; tile_i += tile_stride, phase ^= 1) {
        const int e = a.elist ? a.elist[tile_i] : tile_i;
        const TileHdr* rec = tile_record(a, e);
        const TileHdr hd = rec[0];
        DBG_T(0);
        const double* inbox = nullptr;
        if (HALO && tile_i >= a.halo.n_int) {
            if (HALO == 1) {
                // boundary tile: the peers' rows of this epoch must have landed in my inbox before anybody reads it
                if (threadIdx.x < 32) halo_wait_peers(a, epoch, threadIdx.x);
                __syncthreads();
                inbox = a.halo.inbox + (epoch % a.halo.nbuf) * a.halo.parity_stride;
            } else {
                // LL: cells validate themselves; parity_stride counts 16-byte cells (= 2 doubles)
                inbox = a.halo.inbox + (epoch % a.halo.nbuf) * a.halo.parity_stride * 2;
            }
        }
#ifdef MIMSEM_DIAG
        if (!(a.debug & 16))
#endif
        {
            if (threadIdx.x < 32) tile_load(a, e, inbox, bar, geo, tile);
            else if (threadIdx.x < 64 && a.prefetch_ahead > 0 && tile_i + a.prefetch_ahead < a.ntiles) {
                const int bn = tile_i + a.prefetch_ahead;
                tile_prefetch(a, a.elist ? a.elist[bn] : bn);
            }
        }
        double ubf[P + 1], hs[P];
        if (k < nl) far_fetch<P, WITH_H, HALO>(a, rec, hd.flags, part, inbox, (unsigned)epoch, k, ubf, hs);
        if (HALO == 2 && inbox) {
            // ghost rows that are STAGED (east column / north row owned elsewhere): unpack their cells into the tile
            // slots with ordinary loads (the copy-list walker skips kind-4 entries in this mode); rows alternate
            // between the two warp-pairs, lanes are levels
            const TileFarH fh = *reinterpret_cast<const TileFarH*>(rec + 2);
            const CopyEnt* ents = reinterpret_cast<const CopyEnt*>(rec + a.rec_hdr);
            if (k < nl)
                for (int ci = fh.first4; ci < fh.first4 + fh.n4; ci++) {
                    const CopyEnt c = ents[ci];
                    const uint4* cb = reinterpret_cast<const uint4*>(inbox) + (size_t)c.src * a.nlev + k;
                    int j = part;
                    for (; j + 2 < c.count; j += 4) {   // two rows of this warp-pair in flight
                        const uint4* cell[2] = {cb + (size_t)j * a.nlev, cb + (size_t)(j + 2) * a.nlev};
                        double v[2];
                        ll_load_many<2>(cell, (unsigned)epoch, a.halo.err, v);
                        tile[(size_t)(c.slot + j) * nl + k] = v[0];
                        tile[(size_t)(c.slot + j + 2) * nl + k] = v[1];
                    }
                    for (; j < c.count; j += 2) tile[(size_t)(c.slot + j) * nl + k] = ll_load(cb + (size_t)j * a.nlev, (unsigned)epoch, a.halo.err);
                }
            __syncthreads();   // generic-proxy writes of one warp-pair are read by the other
        }
        DBG_T(1);
#ifdef MIMSEM_DIAG
        if (!(a.debug & 16))
#endif
        mbar_wait(bar, phase);
        DBG_T(2);
#ifdef MIMSEM_DIAG
        if (!(a.debug & 1))
#endif
        {
            const bool active = k < nl;   // level lanes beyond nlev only take part in the barriers
            double* col = tile + k;
            double cfar[P];
            if (active) far_line<P, WITH_H, NL, TPOW>(a, col, geo, hd.flags, part, ubf, hs, cfar);
            if (WITH_H) {
                __syncthreads();   // the far lines have read the raw thickness of the west column / south row
                if (active) h_prepass<P, NL, TPOW>(a, col, part);
                __syncthreads();
            }
            if (active) tile_lines<P, WITH_H, NL, TPOW>(a, col, geo, part, cfar, a.y + (size_t)(hd.st_dof + (part ? S::OY : S::OX)) * a.ld + k);
        }
        DBG_T(3);
        if (tile_i + tile_stride < a.ntiles) __syncthreads();   // the next tile's bulk loads overwrite the buffer
    }
    if (HALO && first_tile >= a.halo.n_int) halo_cta_done(a, epoch);
}

}  // namespace mimsem
