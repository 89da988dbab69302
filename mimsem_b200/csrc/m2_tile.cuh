// y = M2 x and y = M2(rho) x as a TMA-staged tile kernel (Wmat / Whmat, eul/Assembly.cpp:311-373, 1243-1299): element-local,
// one CTA of 64 level lanes per element; the element's faces, the inverse thickness at its quadrature points (and the
// coefficient's faces) are staged by bulk copies, each thread contracts one level, sum-factorised:
//   ax[iy][qx] = sum_ix E[qx][ix] x(ix,iy) ;  xl(q) = sum_iy E[qy][iy] ax[iy][qx] ;  g(q) = c(q) xl(q), c = s w/det t^tpow [rho_l/det]
//   y(ix,iy) = sum_qx E[qx][ix] sum_qy E[qy][iy] g(q)
#pragma once
#include "tile_common.cuh"

namespace mimsem {

template <int P>
struct M2Slots {
    static constexpr int X = 0;
    static constexpr int T = P * P;
    static constexpr int NS = T + (P + 1) * (P + 1);
    static constexpr int H = NS;
    static constexpr int NS_H = H + P * P;
    static constexpr int GEO = (((P + 1) * (P + 1)) + 1) / 2 * 2;   // w/det (M2) or w/det^2 (M2h) per quadrature point
};

template <int P, bool WITH_H, int NL>
__global__ void __launch_bounds__(64) k_apply_m2_tile(const __grid_constant__ TArgs a) {
    using S = M2Slots<P>;
    constexpr int NP1 = P + 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    double* geo = reinterpret_cast<double*>(smem_raw + 16);
    double* tile = geo + S::GEO;
    const int k = threadIdx.x;
    const int nl = NL ? NL : a.nlev;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_async_smem();
    }
    if (a.pdl) pdl_launch_dependents();
    __syncthreads();
    const int e = a.elist ? a.elist[blockIdx.x] : (int)blockIdx.x;
    if (threadIdx.x < 32) tile_load(a, e, nullptr, bar, geo, tile);
    else if (a.prefetch_ahead > 0 && (int)blockIdx.x + a.prefetch_ahead < a.ntiles) {
        const int bn = blockIdx.x + a.prefetch_ahead;
        tile_prefetch(a, a.elist ? a.elist[bn] : bn);
    }
    const TileHdr hd = *tile_record(a, e);
    mbar_wait(bar, 0);
    if (k >= nl) return;
    const double* col = tile + k;
#define SLOT(s) col[(size_t)(s) * nl]
    double ax[P][NP1], rx[WITH_H ? P : 1][WITH_H ? NP1 : 1];
#pragma unroll
    for (int iy = 0; iy < P; iy++) {
        double xv[P], rv[WITH_H ? P : 1];
#pragma unroll
        for (int ix = 0; ix < P; ix++) {
            xv[ix] = SLOT(S::X + iy * P + ix);
            if (WITH_H) rv[WITH_H ? ix : 0] = SLOT(S::H + iy * P + ix);
        }
#pragma unroll
        for (int qx = 0; qx <= P; qx++) {
            double s = 0.0, r = 0.0;
#pragma unroll
            for (int ix = 0; ix < P; ix++) {
                s += a.E[qx * P + ix] * xv[ix];
                if (WITH_H) r += a.E[qx * P + ix] * rv[WITH_H ? ix : 0];
            }
            ax[iy][qx] = s;
            if (WITH_H) rx[WITH_H ? iy : 0][WITH_H ? qx : 0] = r;
        }
    }
    double* __restrict__ y = a.y + (size_t)hd.st_dof * a.ld + k;
    double out[P][P];
#pragma unroll
    for (int iy = 0; iy < P; iy++)
#pragma unroll
        for (int ix = 0; ix < P; ix++) out[iy][ix] = 0.0;
#pragma unroll
    for (int qx = 0; qx <= P; qx++) {
        double g[NP1];
#pragma unroll
        for (int qy = 0; qy <= P; qy++) {
            double xl = 0.0, rl = 0.0;
#pragma unroll
            for (int iy = 0; iy < P; iy++) {
                xl += a.E[qy * P + iy] * ax[iy][qx];
                if (WITH_H) rl += a.E[qy * P + iy] * rx[WITH_H ? iy : 0][WITH_H ? qx : 0];
            }
            const int q = qy * NP1 + qx;
            double c = geo[q];
            if (a.tpow > 0) {
                const double t = SLOT(S::T + q);
                c *= t;
                if (a.tpow > 1) c *= t;
            }
            if (WITH_H) c *= rl;
            g[qy] = c * xl;
        }
#pragma unroll
        for (int iy = 0; iy < P; iy++) {
            double b = 0.0;
#pragma unroll
            for (int qy = 0; qy <= P; qy++) b += a.Es[qy * P + iy] * g[qy];   // Es = scale * E
#pragma unroll
            for (int ix = 0; ix < P; ix++) out[iy][ix] += a.E[qx * P + ix] * b;
        }
    }
#pragma unroll
    for (int iy = 0; iy < P; iy++)
#pragma unroll
        for (int ix = 0; ix < P; ix++) y[(size_t)(iy * P + ix) * a.ld] = out[iy][ix];
#undef SLOT
}

}  // namespace mimsem
