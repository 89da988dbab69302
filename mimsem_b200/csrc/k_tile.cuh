// y = K(u1) x as a TMA-staged tile kernel (WtQUmat / WtQdUdz_mat, eul/Assembly.cpp:933-986, 1581-1640).
#pragma once
#include "tile_common.cuh"

namespace mimsem {

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
    double* tile = geo + M1Slots<P>::GEO_K;
    const int part = threadIdx.x >> 6;
    const int k = threadIdx.x & 63;
    const int nl = NL ? NL : a.nlev;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_async_smem();
    }
    if (a.pdl) pdl_launch_dependents();
    __syncthreads();
    unsigned phase = 0;
    for (int tile_i = blockIdx.x; tile_i < a.ntiles; tile_i += gridDim.x, phase ^= 1) {
        const int e = a.elist ? a.elist[tile_i] : tile_i;
        if (threadIdx.x < 32) tile_load(a, e, nullptr, bar, geo, tile);
        else if (threadIdx.x < 64 && a.prefetch_ahead > 0 && tile_i + a.prefetch_ahead < a.ntiles) {
            const int bn = tile_i + a.prefetch_ahead;
            tile_prefetch(a, a.elist ? a.elist[bn] : bn);
        }
        const TileHdr hd = *tile_record(a, e);
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
