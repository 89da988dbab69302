// TMA-staged tile kernels (sm_100a): see the comment above M1Slots in engine.cuh.
#pragma once
#include <cstdint>

#include "engine.cuh"

namespace mimsem {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void bulk_commit_wait_all() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Stage one tile: warp 0 walks the element's copy list (one entry per lane and round)
__device__ __forceinline__ void tile_load(const TArgs& a, int e, uint64_t* bar, double* geo, double* tile) {
    const TileHdr h = a.hdr[e];
    const int lane = threadIdx.x;
    const unsigned slot_bytes = (unsigned)a.nlev * 8u;
    const bool with_t = a.tpow > 0;
    for (int ci = lane; ci < h.cp_count; ci += 32) {
        const CopyEnt c = a.cps[h.cp_begin + ci];
        if (c.kind == 2 && !with_t) continue;
        if (c.kind == 3) {
            bulk_g2s(geo, a.geo + (size_t)c.src * a.geo_doubles, (unsigned)a.geo_doubles * 8u, bar);
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
    const unsigned nslots = (unsigned)(h.nslots & 0xffff) + (with_t ? (unsigned)(h.nslots >> 16) : 0u);
    if (lane == 0) mbar_arrive_expect_tx(bar, nslots * slot_bytes + (unsigned)a.geo_doubles * 8u);
}

// y = M1 x (WITH_H: M1(h) x), one CTA per element, thread k = level k.
template <int P, bool WITH_H>
__global__ void __launch_bounds__(64) k_apply_m1_tma(const __grid_constant__ TArgs a) {
    using S = M1Slots<P>;
    constexpr int NP1 = P + 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    double* geo = reinterpret_cast<double*>(smem_raw + 16);
    double* tile = geo + S::GEO;
    const int e = blockIdx.x;
    const int k = threadIdx.x;
    if (k == 0) mbar_init(bar, 1);
    __syncthreads();
    if (k < 32) tile_load(a, e, bar, geo, tile);
    const int flags = a.hdr[e].flags;
    mbar_wait(bar, 0);

    if (k < a.nlev) {
        const int nl = a.nlev;
        double* col = tile + k;
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
        // west / south neighbours' far lines -> contributions to my west x-edges / south y-edges
        double cw[P], cs[P];
#pragma unroll
        for (int j = 0; j < P; j++) cw[j] = cs[j] = 0.0;
#pragma unroll
        for (int side = 0; side < 2; side++) {
            const bool has = flags & (side == 0 ? 1 : 4);
            const bool rev = flags & (side == 0 ? 2 : 8);
            if (!has) continue;
            const int OTH = side == 0 ? S::WOTH : S::SOTH;
            const double* gf = geo + (side == 0 ? S::GW : S::GS);
            // the shared edges, in the neighbour's order
            double own[P];
#pragma unroll
            for (int j = 0; j < P; j++) {
                const int mine = rev ? P - 1 - j : j;   // my iy (west) / ix (south)
                // west: xx(0, iy) -> slot OX + iy ; south: xy(ix, 0) -> slot OY + ix
                own[j] = rev ? (side == 0 ? SLOT(S::OX + (P - 1 - j)) : SLOT(S::OY + (P - 1 - j)))
                             : (side == 0 ? SLOT(S::OX + j) : SLOT(S::OY + j));
                (void)mine;
            }
            double hs[P];
            if (WITH_H) {
                // neighbour's h contracted across its far line: which index is "across" depends on whether the far
                // line is its east column (contract ix) or its north row (contract iy): bit 4/5 of flags
                const bool far_is_row = flags & (side == 0 ? 16 : 32);
                const int HN = side == 0 ? S::HW : S::HS;
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
                const int qm = rev ? P - q : q;
                double c = tf(side == 0 ? qm * NP1 : qm);
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
                // neighbour's edge j is my edge (rev ? P-1-j : j)
                if (side == 0) {
                    if (rev) cw[P - 1 - j] = s;
                    else cw[j] = s;
                } else {
                    if (rev) cs[P - 1 - j] = s;
                    else cs[j] = s;
                }
            }
        }

        // own element
        double xy[P + 1][P];
#pragma unroll
        for (int iy = 0; iy <= P; iy++)
#pragma unroll
            for (int ix = 0; ix < P; ix++) xy[iy][ix] = (iy < P) ? SLOT(S::OY + iy * P + ix) : SLOT(S::YN + ix);
        double hx[P][P + 1];
        if (WITH_H) {
#pragma unroll
            for (int iy = 0; iy < P; iy++)
#pragma unroll
                for (int qx = 0; qx <= P; qx++) {
                    double s = 0.0;
#pragma unroll
                    for (int ix = 0; ix < P; ix++) s += a.E[qx * P + ix] * SLOT(S::H + iy * P + ix);
                    hx[iy][qx] = s;
                }
        }
        double yy[P][P];
#pragma unroll
        for (int iy = 0; iy < P; iy++)
#pragma unroll
            for (int ix = 0; ix < P; ix++) yy[iy][ix] = 0.0;
#pragma unroll
        for (int qx = 0; qx <= P; qx++) {
            double xc[P];
#pragma unroll
            for (int iy = 0; iy < P; iy++) xc[iy] = (qx < P) ? SLOT(S::OX + qx * P + iy) : SLOT(S::XE + iy);
            double f0[P + 1];
#pragma unroll
            for (int qy = 0; qy <= P; qy++) {
                double ul0 = 0.0, ul1 = 0.0;
#pragma unroll
                for (int iy = 0; iy < P; iy++) ul0 += a.E[qy * P + iy] * xc[iy];
#pragma unroll
                for (int ix = 0; ix < P; ix++) ul1 += a.E[qx * P + ix] * xy[qy][ix];
                const int q = qy * NP1 + qx;
                double c = tf(q);
                if (WITH_H) {
                    double hl = 0.0;
#pragma unroll
                    for (int iy = 0; iy < P; iy++) hl += a.E[qy * P + iy] * hx[iy][qx];
                    c *= hl;
                }
                const double g0 = geo[q * 3 + 0], g1 = geo[q * 3 + 1], g2 = geo[q * 3 + 2];
                f0[qy] = c * (g0 * ul0 + g1 * ul1);
                if (qy < P) {
                    const double f1 = c * (g1 * ul0 + g2 * ul1);
#pragma unroll
                    for (int ix = 0; ix < P; ix++) yy[qy][ix] += a.E[qx * P + ix] * f1;
                }
            }
            if (qx < P) {
#pragma unroll
                for (int iy = 0; iy < P; iy++) {
                    double s = (qx == 0) ? cw[iy] : 0.0;
#pragma unroll
                    for (int qy = 0; qy <= P; qy++) s += a.E[qy * P + iy] * f0[qy];
                    // the column's own x-edge slots were read above (xc) and by the west far line: overwrite in place
                    SLOT(S::OX + qx * P + iy) = s;
                }
            }
        }
        // results replace the inputs in the own-block slots (each thread only ever touches its own column)
#pragma unroll
        for (int iy = 0; iy < P; iy++)
#pragma unroll
            for (int ix = 0; ix < P; ix++) SLOT(S::OY + iy * P + ix) = yy[iy][ix] + (iy == 0 ? cs[ix] : 0.0);
#undef SLOT
    }
    fence_async_smem();
    __syncthreads();
    // bulk stores of the owned block
    if (k < 32) {
        const unsigned slot_bytes = (unsigned)a.nlev * 8u;
        for (int si = a.st_ptr[e] + k; si < a.st_ptr[e + 1]; si += 32) {
            const StoreEnt s = a.stores[si];
            double* dst = a.y + (size_t)s.dof * a.ld;
            const double* src = tile + (size_t)s.slot * a.nlev;
            if (a.contig_x) bulk_s2g(dst, src, slot_bytes * s.count);
            else
                for (int j = 0; j < s.count; j++) bulk_s2g(dst + (size_t)j * a.ld, src + (size_t)j * a.nlev, slot_bytes);
        }
        bulk_commit_wait_all();
    }
}

}  // namespace mimsem
