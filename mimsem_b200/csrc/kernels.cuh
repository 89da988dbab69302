// Matrix-free, sum-factorised FP64 element kernels for sm_100a.
//
// Work decomposition: one thread per (element, vertical level), level fastest, over fields in
// column layout f[dof*ld + k].  Consecutive lanes therefore touch consecutive addresses for
// every DOF of the element (perfectly coalesced 8-byte accesses, runs of nlev*8 bytes), the
// whole element contraction happens in registers (no shuffles, no shared memory), geometry and
// index tables are warp-uniform broadcast loads amortised over the nlev levels of the launch,
// and the edge basis table sits in the constant bank (kernel parameters).
//
// Shared degrees of freedom are summed by GATHER, not by atomics: an element owns its west,
// south and interior edges (the reference's global numbering has exactly this ownership,
// scr/Proc2.py:105-123); the contribution of the west / south neighbour to those edges only
// needs that neighbour's east column / north row of quadrature points, which the owner
// recomputes.  Results are therefore bit-reproducible and independent of the partition.
//
// Formula sheet (restated from the reference, m == p so the nodal table is the identity):
//   ul0(qx,qy) = sum_iy E[qy][iy] xx(qx,iy)      ul1(qx,qy) = sum_ix E[qx][ix] xy(ix,qy)
//   hl (qx,qy) = sum_iy E[qy][iy] sum_ix E[qx][ix] h(ix,iy)
//   M1 : f = c (Gaa ul0 + Gab ul1 , Gab ul0 + Gbb ul1),  c = s w/det t^tpow [hl/det]
//        y^x(ix,iy) = sum_qy E[qy][iy] f0(ix,qy) ;  y^y(ix,iy) = sum_qx E[qx][ix] f1(qx,iy)
//        (eul/Assembly.cpp:99-131 Umat, :432-467 Uhmat; matrix-free twins :2144-2188, :2221-2268)
//   M2 : y = W^T c W x,  c = s w/det t^tpow [rho_l/det]          (eul/Assembly.cpp:347-358, :1268-1285)
//   K  : y = W^T (ka ul0(x) + kb ul1(x)),  (ka,kb) = 1/2 s t^2 w/det^2 G ul(u1)   (eul/Assembly.cpp:951-979)
#pragma once
#include "engine.cuh"
#include "p2p_sync.cuh"

namespace mimsem {

template <int P>
struct ElDim {
    static constexpr int NP1 = P + 1;
    static constexpr int N1E = P * (P + 1);
    static constexpr int N2E = P * P;
    static constexpr int Q2 = (P + 1) * (P + 1);
};

__device__ __forceinline__ double ldro(const double* p) { return __ldg(p); }

// scale * (1/thick)^tpow at local quadrature point gq, column k
__device__ __forceinline__ double thick_factor(const KArgs& a, int gq, int k) {
    double f = a.scale;
    if (a.tpow > 0) {
        const double t = ldro(a.tinv + (size_t)gq * a.nkT + a.lev0 + k * a.lev_stride);
        f *= t;
        if (a.tpow > 1) f *= t;
    }
    return f;
}

// Umat_ray's point weight dt k_v(exner, exner_s) (compute_k_v, eul/Assembly.cpp:1846-1856) from the interpolated (without
// 1/det) Exner 2-forms hl (this level) and hls (level 0) at local quadrature point gq of element point (n, q)
__device__ __forceinline__ double ray_weight(const KArgs& a, double hl, double hls, int n_q2_plus_q, int gq, int k) {
    const double idet = 1.0 / ldro(a.det + n_q2_plus_q);
    const double t = ldro(a.tinv + (size_t)gq * a.nkT + a.lev0 + k * a.lev_stride), t0 = ldro(a.tinv + (size_t)gq * a.nkT);
    const double ex = hl * idet * t, exs = hls * idet * t0;
    const double CP = 1004.5, RD = 287.0;
    const double sigma = pow(ex / CP, CP / RD) / pow(exs / CP, CP / RD);
    const double sigma_b = 0.7, k_f = 1.1574074074074073e-05;
    if (sigma < sigma_b) return 0.0;
    return a.ray_dt * (k_f * (sigma - sigma_b) / (1.0 - sigma_b));
}

// Contribution of neighbour element n to the P edges of one of its far sides:
//   side 0: n's east column of x-normal edges  (quadrature points (P, qy))
//   side 1: n's north row of y-normal edges    (quadrature points (qx, P))
template <int P, bool WITH_H>
__device__ __forceinline__ void m1_far_side(const KArgs& a, int n, int side, bool rev, int k, double (&out)[P]) {
    using D = ElDim<P>;
    const size_t ld = a.ld;
    const double* __restrict__ x = a.x + k;
    const int* __restrict__ nx = a.el1x + (size_t)n * D::N1E;
    const int* __restrict__ ny = a.el1y + (size_t)n * D::N1E;
    const int* __restrict__ nq = a.elq + (size_t)n * D::Q2;
    const double* __restrict__ G = a.G + (size_t)n * D::Q2 * 3;
    double f[P + 1];
    double hs[P];   // h contracted along the side-normal direction at the far abscissa
    double hss[P];  // Umat_ray: the same for the level-0 Exner field
    const bool ray = WITH_H && a.ray_dt != 0.0;
    if (WITH_H) {
        const double* __restrict__ h = a.c + k;
        const int* __restrict__ n2 = a.el2 + (size_t)n * D::N2E;
#pragma unroll
        for (int i = 0; i < P; i++) hs[i] = hss[i] = 0.0;
#pragma unroll
        for (int iy = 0; iy < P; iy++)
#pragma unroll
            for (int ix = 0; ix < P; ix++) {
                const double hv = ldro(h + (size_t)n2[iy * P + ix] * ld);
                if (side == 0) hs[iy] += a.E[P * P + ix] * hv;   // E[P][ix]
                else hs[ix] += a.E[P * P + iy] * hv;             // E[P][iy]
                if (ray) {
                    const double hv0 = ldro(a.c2 + (size_t)n2[iy * P + ix]);
                    if (side == 0) hss[iy] += a.E[P * P + ix] * hv0;
                    else hss[ix] += a.E[P * P + iy] * hv0;
                }
            }
    }
    if (side == 0) {
        double xe[P];
#pragma unroll
        for (int iy = 0; iy < P; iy++) xe[iy] = ldro(x + (size_t)nx[iy * D::NP1 + P] * ld);
#pragma unroll
        for (int qy = 0; qy <= P; qy++) {
            double ul0 = 0.0, ul1 = 0.0;
#pragma unroll
            for (int iy = 0; iy < P; iy++) ul0 += a.E[qy * P + iy] * xe[iy];
#pragma unroll
            for (int ix = 0; ix < P; ix++) ul1 += a.E[P * P + ix] * ldro(x + (size_t)ny[qy * P + ix] * ld);
            const int q = qy * D::NP1 + P;
            double c = thick_factor(a, nq[q], k);
            if (WITH_H) {
                double hl = 0.0;
#pragma unroll
                for (int iy = 0; iy < P; iy++) hl += a.E[qy * P + iy] * hs[iy];
                if (ray) {
                    double hls = 0.0;
#pragma unroll
                    for (int iy = 0; iy < P; iy++) hls += a.E[qy * P + iy] * hss[iy];
                    hl = ray_weight(a, hl, hls, n * D::Q2 + q, nq[q], k);
                }
                c *= hl;
            }
            f[qy] = c * (G[q * 3 + 0] * ul0 + G[q * 3 + 1] * ul1);
        }
    } else {
        double ye[P];
#pragma unroll
        for (int ix = 0; ix < P; ix++) ye[ix] = ldro(x + (size_t)ny[P * P + ix] * ld);
#pragma unroll
        for (int qx = 0; qx <= P; qx++) {
            double ul0 = 0.0, ul1 = 0.0;
#pragma unroll
            for (int ix = 0; ix < P; ix++) ul1 += a.E[qx * P + ix] * ye[ix];
#pragma unroll
            for (int iy = 0; iy < P; iy++) ul0 += a.E[P * P + iy] * ldro(x + (size_t)nx[iy * D::NP1 + qx] * ld);
            const int q = P * D::NP1 + qx;
            double c = thick_factor(a, nq[q], k);
            if (WITH_H) {
                double hl = 0.0;
#pragma unroll
                for (int ix = 0; ix < P; ix++) hl += a.E[qx * P + ix] * hs[ix];
                if (ray) {
                    double hls = 0.0;
#pragma unroll
                    for (int ix = 0; ix < P; ix++) hls += a.E[qx * P + ix] * hss[ix];
                    hl = ray_weight(a, hl, hls, n * D::Q2 + q, nq[q], k);
                }
                c *= hl;
            }
            f[qx] = c * (G[q * 3 + 1] * ul0 + G[q * 3 + 2] * ul1);
        }
    }
    double o[P];
#pragma unroll
    for (int i = 0; i < P; i++) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q <= P; q++) s += a.E[q * P + i] * f[q];
        o[i] = s;
    }
#pragma unroll
    for (int i = 0; i < P; i++) out[i] = rev ? o[P - 1 - i] : o[i];
}

// y = M1 x   (WITH_H: M1(h) x with the 2-form coefficient a.c)
template <int P, bool WITH_H>
__global__ void __launch_bounds__(128) k_apply_m1(const __grid_constant__ KArgs a) {
    using D = ElDim<P>;
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)a.nel * (unsigned)a.nlev) return;
    int e = (int)fastdiv(idx, a.div_m, a.div_s);
    const int k = (int)(idx - (unsigned)e * (unsigned)a.nlev);
    if (a.elist) e = a.elist[e];
    const size_t ld = a.ld;
    const double* __restrict__ x = a.x + k;
    double* __restrict__ y = a.y + k;
    const int* __restrict__ ex = a.el1x + (size_t)e * D::N1E;
    const int* __restrict__ ey = a.el1y + (size_t)e * D::N1E;
    const int* __restrict__ eq = a.elq + (size_t)e * D::Q2;
    const double* __restrict__ G = a.G + (size_t)e * D::Q2 * 3;

    // own degrees of freedom
    double xy[P + 1][P];
#pragma unroll
    for (int iy = 0; iy <= P; iy++)
#pragma unroll
        for (int ix = 0; ix < P; ix++) xy[iy][ix] = ldro(x + (size_t)ey[iy * P + ix] * ld);
    double xx[P][P + 1];
#pragma unroll
    for (int iy = 0; iy < P; iy++)
#pragma unroll
        for (int ix = 0; ix <= P; ix++) xx[iy][ix] = ldro(x + (size_t)ex[iy * D::NP1 + ix] * ld);
    double hx[P][P + 1];   // h contracted in x: hx[iy][qx]
    double hxs[WITH_H ? P : 1][WITH_H ? P + 1 : 1];   // Umat_ray: the same for the level-0 Exner field
    const bool ray = WITH_H && a.ray_dt != 0.0;
    if (WITH_H) {
        const double* __restrict__ h = a.c + k;
        const int* __restrict__ e2 = a.el2 + (size_t)e * D::N2E;
#pragma unroll
        for (int iy = 0; iy < P; iy++) {
            double hv[P], hv0[P];
#pragma unroll
            for (int ix = 0; ix < P; ix++) {
                hv[ix] = ldro(h + (size_t)e2[iy * P + ix] * ld);
                hv0[ix] = ray ? ldro(a.c2 + (size_t)e2[iy * P + ix]) : 0.0;
            }
#pragma unroll
            for (int qx = 0; qx <= P; qx++) {
                double s = 0.0, s0 = 0.0;
#pragma unroll
                for (int ix = 0; ix < P; ix++) {
                    s += a.E[qx * P + ix] * hv[ix];
                    s0 += a.E[qx * P + ix] * hv0[ix];
                }
                hx[iy][qx] = s;
                hxs[WITH_H ? iy : 0][WITH_H ? qx : 0] = s0;
            }
        }
    }

    // neighbours' contributions to the west x-edges and south y-edges
    double cw[P], cs[P];
#pragma unroll
    for (int i = 0; i < P; i++) cw[i] = cs[i] = 0.0;
    const int nw = a.nbr[2 * e + 0], ns = a.nbr[2 * e + 1];
    if (nw >= 0) m1_far_side<P, WITH_H>(a, nw & 0x1fffffff, (nw >> 29) & 1, (nw >> 30) & 1, k, cw);
    if (ns >= 0) m1_far_side<P, WITH_H>(a, ns & 0x1fffffff, (ns >> 29) & 1, (ns >> 30) & 1, k, cs);

    const unsigned flags = a.eflags[e];
    double yy[P + 1][P];
#pragma unroll
    for (int iy = 0; iy <= P; iy++)
#pragma unroll
        for (int ix = 0; ix < P; ix++) yy[iy][ix] = 0.0;

#pragma unroll
    for (int qx = 0; qx <= P; qx++) {
        double f0[P + 1];
#pragma unroll
        for (int qy = 0; qy <= P; qy++) {
            double ul0 = 0.0, ul1 = 0.0;
#pragma unroll
            for (int iy = 0; iy < P; iy++) ul0 += a.E[qy * P + iy] * xx[iy][qx];
#pragma unroll
            for (int ix = 0; ix < P; ix++) ul1 += a.E[qx * P + ix] * xy[qy][ix];
            const int q = qy * D::NP1 + qx;
            double c = thick_factor(a, eq[q], k);
            if (WITH_H) {
                double hl = 0.0;
#pragma unroll
                for (int iy = 0; iy < P; iy++) hl += a.E[qy * P + iy] * hx[iy][qx];
                if (ray) {
                    double hls = 0.0;
#pragma unroll
                    for (int iy = 0; iy < P; iy++) hls += a.E[qy * P + iy] * hxs[WITH_H ? iy : 0][WITH_H ? qx : 0];
                    hl = ray_weight(a, hl, hls, e * D::Q2 + q, eq[q], k);
                }
                c *= hl;
            }
            const double g0 = G[q * 3 + 0], g1 = G[q * 3 + 1], g2 = G[q * 3 + 2];
            f0[qy] = c * (g0 * ul0 + g1 * ul1);
            const double f1 = c * (g1 * ul0 + g2 * ul1);
#pragma unroll
            for (int ix = 0; ix < P; ix++) yy[qy][ix] += a.E[qx * P + ix] * f1;
        }
        // x-normal edges of column ix = qx
        if (qx < P || (flags & 1u)) {
#pragma unroll
            for (int iy = 0; iy < P; iy++) {
                double s = (qx == 0) ? cw[iy] : 0.0;
#pragma unroll
                for (int qy = 0; qy <= P; qy++) s += a.E[qy * P + iy] * f0[qy];
                y[(size_t)ex[iy * D::NP1 + qx] * ld] = s;
            }
        }
    }
#pragma unroll
    for (int iy = 0; iy < P; iy++)
#pragma unroll
        for (int ix = 0; ix < P; ix++) y[(size_t)ey[iy * P + ix] * ld] = yy[iy][ix] + (iy == 0 ? cs[ix] : 0.0);
    if (flags & 2u) {
#pragma unroll
        for (int ix = 0; ix < P; ix++) y[(size_t)ey[P * P + ix] * ld] = yy[P][ix];
    }
}


// ---------------------------------------------------------------------------------------------
// M1 as independent GLL-line tasks (v1).
//
// The 2x2 block structure of M1 couples x- and y-edges only through the quadrature points, so the
// P x-edge outputs of one GLL column (ix fixed) need exactly the P+1 quadrature points of that
// column, and the P y-edge outputs of one GLL row (iy fixed) the P+1 points of that row:
//   column i:  out[iy] = sum_qy E[qy][iy] c (Gaa ul0 + Gab ul1)(i,qy)
//   row    i:  out[ix] = sum_qx E[qx][ix] c (Gab ul0 + Gbb ul1)(qx,i)
// One thread handles one (element, line, level): ~24 loads, ~80 FP64 FMAs, ~40 registers, so an SM
// keeps 40+ warps resident instead of 11 (the thread-per-element kernel was latency-bound by
// occupancy, profiles/r01_m1_v0_thread_per_element.md).  The west-most column / south-most row
// additionally gathers the neighbour's far line (i = P), exactly as before.
template <int P, int DIR, bool WITH_H>
__device__ __forceinline__ void m1_line(const KArgs& a, int n, int i, int k, double (&out)[P]) {
    using D = ElDim<P>;
    const size_t ld = a.ld;
    const double* __restrict__ x = a.x + k;
    // line-local DOFs (the P edges normal to the line direction that sit ON the line's abscissa)
    // and the (P+1) x P edges of the other family that the line's quadrature points interpolate
    const int* __restrict__ iown = (DIR == 0) ? a.el1xT + (size_t)n * D::N1E + i * P      // xx(i, iy), iy = 0..P-1
                                              : a.el1y + (size_t)n * D::N1E + i * P;      // xy(ix, i), ix = 0..P-1
    const int* __restrict__ ioth = (DIR == 0) ? a.el1y + (size_t)n * D::N1E               // xy(ix, qy): [qy][ix]
                                              : a.el1xT + (size_t)n * D::N1E;             // xx(qx, iy): [qx][iy]
    const int* __restrict__ iq = (DIR == 0) ? a.elqT + (size_t)n * D::Q2 + i * D::NP1 : a.elq + (size_t)n * D::Q2 + i * D::NP1;
    const double* __restrict__ G = ((DIR == 0) ? a.Gc : a.Gr) + ((size_t)n * D::Q2 + i * D::NP1) * 2;

    double own[P];
#pragma unroll
    for (int j = 0; j < P; j++) own[j] = ldro(x + (size_t)iown[j] * ld);
    double hs[P];
    if (WITH_H) {
        // h contracted across the line at abscissa i: hs[j] = sum_t E[i][t] h(t, j) (DIR 0) / h(j, t) (DIR 1)
        const double* __restrict__ h = a.c + k;
        const int* __restrict__ n2 = a.el2 + (size_t)n * D::N2E;
#pragma unroll
        for (int j = 0; j < P; j++) hs[j] = 0.0;
#pragma unroll
        for (int iy = 0; iy < P; iy++)
#pragma unroll
            for (int ix = 0; ix < P; ix++) {
                const double hv = ldro(h + (size_t)n2[iy * P + ix] * ld);
                if (DIR == 0) hs[iy] += a.E[i * P + ix] * hv;
                else hs[ix] += a.E[i * P + iy] * hv;
            }
    }
    double f[P + 1];
#pragma unroll
    for (int q = 0; q <= P; q++) {
        double ua = 0.0, ub = 0.0;   // along-line interpolation of `own`, across-line interpolation of `oth`
#pragma unroll
        for (int j = 0; j < P; j++) ua += a.E[q * P + j] * own[j];
#pragma unroll
        for (int t = 0; t < P; t++) ub += a.E[i * P + t] * ldro(x + (size_t)ioth[q * P + t] * ld);
        double c = thick_factor(a, iq[q], k);
        if (WITH_H) {
            double hl = 0.0;
#pragma unroll
            for (int j = 0; j < P; j++) hl += a.E[q * P + j] * hs[j];
            c *= hl;
        }
        // DIR 0: (g0,g1) = (Gaa,Gab), ua = ul0, ub = ul1 ; DIR 1: (g1,g2) = (Gab,Gbb), ub = ul0, ua = ul1
        f[q] = (DIR == 0) ? c * (G[q * 2 + 0] * ua + G[q * 2 + 1] * ub) : c * (G[q * 2 + 0] * ub + G[q * 2 + 1] * ua);
    }
#pragma unroll
    for (int j = 0; j < P; j++) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q <= P; q++) s += a.E[q * P + j] * f[q];
        out[j] = s;
    }
}

// runtime line index -> compile-time E row (the table sits in the constant bank; a runtime row
// index would turn every operand into an indexed constant load)
template <int P, int DIR, bool WITH_H>
__device__ __forceinline__ void m1_line_rt(const KArgs& a, int n, int i, int k, double (&out)[P]) {
    // i is warp-uniform; dispatch keeps the E[i][*] operands immediate
    switch (i) {
        case 0: m1_line<P, DIR, WITH_H>(a, n, 0, k, out); break;
        case 1: m1_line<P, DIR, WITH_H>(a, n, 1, k, out); break;
        case 2: if (P >= 2) m1_line<P, DIR, WITH_H>(a, n, P >= 2 ? 2 : 0, k, out); break;
        case 3: if (P >= 3) m1_line<P, DIR, WITH_H>(a, n, P >= 3 ? 3 : 0, k, out); break;
        case 4: if (P >= 4) m1_line<P, DIR, WITH_H>(a, n, P >= 4 ? 4 : 0, k, out); break;
        case 5: if (P >= 5) m1_line<P, DIR, WITH_H>(a, n, P >= 5 ? 5 : 0, k, out); break;
        default: break;
    }
}

// grid: x over (element, level) flattened with the level fastest, y over the 2P lines
// (y < P: x-edge columns, y >= P: y-edge rows).  FAR = true launches only the far lines (i = P) of the
// elements flagged in partial-sum mode (a.nbr then holds the flagged element list).
template <int P, bool WITH_H, bool FAR>
__global__ void __launch_bounds__(128) k_apply_m1_lines(const __grid_constant__ KArgs a) {
    using D = ElDim<P>;
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)a.nel * (unsigned)a.nlev) return;
    int e = (int)fastdiv(idx, a.div_m, a.div_s);
    const int k = (int)(idx - (unsigned)e * (unsigned)a.nlev);
    if (!FAR && a.elist) e = a.elist[e];
    const size_t ld = a.ld;
    double* __restrict__ y = a.y + k;
    double out[P];
    if (FAR) {
        // a.nbr = list of (element, dir) pairs whose far line nobody else in the subdomain computes
        const int code = a.nbr[e];
        e = code >> 1;
        if ((code & 1) == 0) {
            m1_line<P, 0, WITH_H>(a, e, P, k, out);
            const int* __restrict__ io = a.el1xT + (size_t)e * D::N1E + P * P;
#pragma unroll
            for (int j = 0; j < P; j++) y[(size_t)io[j] * ld] = out[j];
        } else {
            m1_line<P, 1, WITH_H>(a, e, P, k, out);
            const int* __restrict__ io = a.el1y + (size_t)e * D::N1E + P * P;
#pragma unroll
            for (int j = 0; j < P; j++) y[(size_t)io[j] * ld] = out[j];
        }
        return;
    }
    const int line = blockIdx.y;
    const int dir = line >= P;
    const int i = dir ? line - P : line;
    if (!dir) m1_line_rt<P, 0, WITH_H>(a, e, i, k, out);
    else m1_line_rt<P, 1, WITH_H>(a, e, i, k, out);
    if (i == 0) {
        const int nb = a.nbr[2 * e + dir];
        if (nb >= 0) {
            double far[P];
            const int n = nb & 0x1fffffff;
            if (((nb >> 29) & 1) == 0) m1_line<P, 0, WITH_H>(a, n, P, k, far);
            else m1_line<P, 1, WITH_H>(a, n, P, k, far);
            const bool rev = (nb >> 30) & 1;
#pragma unroll
            for (int j = 0; j < P; j++) out[j] += rev ? far[P - 1 - j] : far[j];
        }
    }
    const int* __restrict__ io = dir ? a.el1y + (size_t)e * D::N1E + i * P : a.el1xT + (size_t)e * D::N1E + i * P;
#pragma unroll
    for (int j = 0; j < P; j++) y[(size_t)io[j] * ld] = out[j];
}

// ---------------------------------------------------------------------------------------------
// Rotational term R(q) and its potential-vorticity-upwinded variant (BASELINE config 2).
//   RotMat::assemble(q0)             y^x = U^T [-c ul1(x)],  y^y = V^T [+c ul0(x)],  c = s t^tpow w sigma q_q
//   RotMat_up::assemble(q0,ul,fac,dt) the same with q_q replaced by the element's nodal interpolant of q0 at the
//                                    departure point xi_q - tau J^-1 u_g(xi_q), tau = fac dt
// (src/Assembly.cpp:1346-1395, 1784-1853; eul/Assembly.cpp:1030-1083 adds t^2 and scale; sigma = det J / |det|).
// One thread per (element, level), structured like k_apply_m1: shared edges by gather from the west / south neighbour.

// Lagrange polynomials through the GLL nodes at x (LagrangeNode::eval_q, src/Basis.cpp:183-190: prod_{j != i}
// (x - x_j) / (x_i - x_j)), with the denominators folded into the barycentric weights wb[i] on the host: the reference's 2 P (P+1)
// FP64 divisions per quadrature point were the whole cost of the upwinded kernels (a few ulps of difference).
template <int P>
__device__ __forceinline__ void lagrange_at(const double* xn, const double* wb, double x, double (&l)[P + 1]) {
    double d[P + 1];
#pragma unroll
    for (int j = 0; j <= P; j++) d[j] = x - xn[j];
#pragma unroll
    for (int i = 0; i <= P; i++) {
        double y = wb[i];
#pragma unroll
        for (int j = 0; j <= P; j++)
            if (j != i) y *= d[j];
        l[i] = y;
    }
}

// departure point of quadrature point (qx,qy) of element n given the local velocity components there
template <int P, class A>
__device__ __forceinline__ void departure_basis(const A& a, int n, int qx, int qy, double ul0, double ul1, double (&lx)[P + 1],
                                                double (&ly)[P + 1]) {
    constexpr int Q2 = (P + 1) * (P + 1);
    const int q = qy * (P + 1) + qx;
    const double* __restrict__ J = a.J4 + ((size_t)n * Q2 + q) * 4;
    const double det = a.det[(size_t)n * Q2 + q];
    // interp1_g (src/Geom.cpp:302-313), then J^-1 (src/Assembly.cpp:1815-1816)
    const double idet = 1.0 / det;
    const double ux0 = (J[0] * ul0 + J[1] * ul1) * idet;
    const double ux1 = (J[2] * ul0 + J[3] * ul1) * idet;
    const double v0 = (J[3] * ux0 - J[1] * ux1) * idet;
    const double v1 = (J[0] * ux1 - J[2] * ux0) * idet;
    lagrange_at<P>(a.xn, a.wb, a.xn[qx] - a.tau * v0, lx);
    lagrange_at<P>(a.xn, a.wb, a.xn[qy] - a.tau * v1, ly);
}

// potential vorticity seen by quadrature point (qx,qy) of element n, column k
template <int P, bool UP>
__device__ __forceinline__ double rot_vort(const KArgs& a, int n, int qx, int qy, int k) {
    using D = ElDim<P>;
    const int* __restrict__ n0 = a.el0 + (size_t)n * D::Q2;
    const size_t ld = a.ld;
    if (!UP) return ldro(a.q0 + (size_t)n0[qy * D::NP1 + qx] * ld + k);
    const double* __restrict__ u = a.u1 + k;
    const int* __restrict__ nx = a.el1x + (size_t)n * D::N1E;
    const int* __restrict__ ny = a.el1y + (size_t)n * D::N1E;
    double ul0 = 0.0, ul1 = 0.0;
#pragma unroll
    for (int iy = 0; iy < P; iy++) ul0 += a.E[qy * P + iy] * ldro(u + (size_t)nx[iy * D::NP1 + qx] * ld);
#pragma unroll
    for (int ix = 0; ix < P; ix++) ul1 += a.E[qx * P + ix] * ldro(u + (size_t)ny[qy * P + ix] * ld);
    double lx[P + 1], ly[P + 1];
    departure_basis<P>(a, n, qx, qy, ul0, ul1, lx, ly);
    double v = 0.0;
#pragma unroll
    for (int jy = 0; jy <= P; jy++)
#pragma unroll
        for (int jx = 0; jx <= P; jx++) v += ldro(a.q0 + (size_t)n0[jy * D::NP1 + jx] * ld + k) * lx[jx] * ly[jy];
    return v;
}

// neighbour n's contribution to the P edges of its far side (side 0: east column of x-edges, 1: north row of y-edges)
template <int P, bool UP>
__device__ __forceinline__ void rot_far_side(const KArgs& a, int n, int side, bool rev, int k, double (&out)[P]) {
    using D = ElDim<P>;
    const size_t ld = a.ld;
    const double* __restrict__ x = a.x + k;
    const int* __restrict__ nx = a.el1x + (size_t)n * D::N1E;
    const int* __restrict__ ny = a.el1y + (size_t)n * D::N1E;
    const int* __restrict__ nq = a.elq + (size_t)n * D::Q2;
    double f[P + 1];
#pragma unroll
    for (int t = 0; t <= P; t++) {
        const int qx = side == 0 ? P : t, qy = side == 0 ? t : P;
        const int q = qy * D::NP1 + qx;
        const double c = thick_factor(a, nq[q], k) * a.Wr[(size_t)n * D::Q2 + q] * rot_vort<P, UP>(a, n, qx, qy, k);
        if (side == 0) {
            double ul1 = 0.0;   // east column x-edges receive -c ul1
#pragma unroll
            for (int ix = 0; ix < P; ix++) ul1 += a.E[qx * P + ix] * ldro(x + (size_t)ny[qy * P + ix] * ld);
            f[t] = -c * ul1;
        } else {
            double ul0 = 0.0;   // north row y-edges receive +c ul0
#pragma unroll
            for (int iy = 0; iy < P; iy++) ul0 += a.E[qy * P + iy] * ldro(x + (size_t)nx[iy * D::NP1 + qx] * ld);
            f[t] = c * ul0;
        }
    }
    double o[P];
#pragma unroll
    for (int i = 0; i < P; i++) {
        double s = 0.0;
#pragma unroll
        for (int t = 0; t <= P; t++) s += a.E[t * P + i] * f[t];
        o[i] = s;
    }
#pragma unroll
    for (int i = 0; i < P; i++) out[i] = rev ? o[P - 1 - i] : o[i];
}

template <int P, bool UP>
__global__ void __launch_bounds__(128) k_apply_rot(const __grid_constant__ KArgs a) {
    using D = ElDim<P>;
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)a.nel * (unsigned)a.nlev) return;
    int e = (int)fastdiv(idx, a.div_m, a.div_s);
    const int k = (int)(idx - (unsigned)e * (unsigned)a.nlev);
    if (a.elist) e = a.elist[e];
    const size_t ld = a.ld;
    const double* __restrict__ x = a.x + k;
    double* __restrict__ y = a.y + k;
    const int* __restrict__ ex = a.el1x + (size_t)e * D::N1E;
    const int* __restrict__ ey = a.el1y + (size_t)e * D::N1E;
    const int* __restrict__ eq = a.elq + (size_t)e * D::Q2;
    double xy[P + 1][P];
#pragma unroll
    for (int iy = 0; iy <= P; iy++)
#pragma unroll
        for (int ix = 0; ix < P; ix++) xy[iy][ix] = ldro(x + (size_t)ey[iy * P + ix] * ld);
    double xx[P][P + 1];
#pragma unroll
    for (int iy = 0; iy < P; iy++)
#pragma unroll
        for (int ix = 0; ix <= P; ix++) xx[iy][ix] = ldro(x + (size_t)ex[iy * D::NP1 + ix] * ld);
    double cw[P], cs[P];
#pragma unroll
    for (int i = 0; i < P; i++) cw[i] = cs[i] = 0.0;
    const int nw = a.nbr[2 * e + 0], ns = a.nbr[2 * e + 1];
    if (nw >= 0) rot_far_side<P, UP>(a, nw & 0x1fffffff, (nw >> 29) & 1, (nw >> 30) & 1, k, cw);
    if (ns >= 0) rot_far_side<P, UP>(a, ns & 0x1fffffff, (ns >> 29) & 1, (ns >> 30) & 1, k, cs);
    const unsigned flags = a.eflags[e];
    double yy[P + 1][P];
#pragma unroll
    for (int iy = 0; iy <= P; iy++)
#pragma unroll
        for (int ix = 0; ix < P; ix++) yy[iy][ix] = 0.0;
#pragma unroll
    for (int qx = 0; qx <= P; qx++) {
        double f0[P + 1];
#pragma unroll
        for (int qy = 0; qy <= P; qy++) {
            double ul0 = 0.0, ul1 = 0.0;
#pragma unroll
            for (int iy = 0; iy < P; iy++) ul0 += a.E[qy * P + iy] * xx[iy][qx];
#pragma unroll
            for (int ix = 0; ix < P; ix++) ul1 += a.E[qx * P + ix] * xy[qy][ix];
            const int q = qy * D::NP1 + qx;
            const double c = thick_factor(a, eq[q], k) * a.Wr[(size_t)e * D::Q2 + q] * rot_vort<P, UP>(a, e, qx, qy, k);
            f0[qy] = -c * ul1;
            const double f1 = c * ul0;
#pragma unroll
            for (int ix = 0; ix < P; ix++) yy[qy][ix] += a.E[qx * P + ix] * f1;
        }
        if (qx < P || (flags & 1u)) {
#pragma unroll
            for (int iy = 0; iy < P; iy++) {
                double s = (qx == 0) ? cw[iy] : 0.0;
#pragma unroll
                for (int qy = 0; qy <= P; qy++) s += a.E[qy * P + iy] * f0[qy];
                y[(size_t)ex[iy * D::NP1 + qx] * ld] = s;
            }
        }
    }
#pragma unroll
    for (int iy = 0; iy < P; iy++)
#pragma unroll
        for (int ix = 0; ix < P; ix++) y[(size_t)ey[iy * P + ix] * ld] = yy[iy][ix] + (iy == 0 ? cs[ix] : 0.0);
    if (flags & 2u) {
#pragma unroll
        for (int ix = 0; ix < P; ix++) y[(size_t)ey[P * P + ix] * ld] = yy[P][ix];
    }
}

// Phmat::assemble_up(ul, hl, fac, dt) followed by MatMult (src/Assembly.cpp:499-567): with m == p the test basis is
// nodal, so row n collects, from every (element, quadrature point) at the node,
//   y_n = sum_{(e,q) at n} w_q hl^e_q(h) * [element e's nodal interpolant of x at xi_q - tau J^-1 u_g(xi_q)].
// One thread per (node, level).
template <int P>
__global__ void __launch_bounds__(128) k_apply_m0h_up(const __grid_constant__ NodeArgs a) {
    using D = ElDim<P>;
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)a.n0 * (unsigned)a.nlev) return;
    const int n = (int)fastdiv(idx, a.div_m, a.div_s);
    const int k = (int)(idx - (unsigned)n * (unsigned)a.nlev);
    const size_t ld = a.ld;
    double f = a.scale;
    if (a.tpow > 0) {
        const double t = ldro(a.tinv + (size_t)a.node_q[n] * a.nkT + a.lev0 + k * a.lev_stride);
        f *= t;
        if (a.tpow > 1) f *= t;
    }
    const double* __restrict__ h = a.c + k;
    const double* __restrict__ u = a.u1 + k;
    double acc = 0.0;
    for (int j = a.adj_ptr[n]; j < a.adj_ptr[n + 1]; j++) {
        const int eqv = a.adj_eq[j];
        const int e = eqv / D::Q2, q = eqv - e * D::Q2;
        const int qx = q % D::NP1, qy = q / D::NP1;
        const int* __restrict__ e2 = a.el2 + (size_t)e * D::N2E;
        const int* __restrict__ e0 = a.el0 + (size_t)e * D::Q2;
        const int* __restrict__ ex = a.el1x + (size_t)e * D::N1E;
        const int* __restrict__ ey = a.el1y + (size_t)e * D::N1E;
        double hl = 0.0;
#pragma unroll
        for (int iy = 0; iy < P; iy++) {
            double s = 0.0;
#pragma unroll
            for (int ix = 0; ix < P; ix++) s += a.E[qx * P + ix] * ldro(h + (size_t)e2[iy * P + ix] * ld);
            hl += a.E[qy * P + iy] * s;
        }
        double ul0 = 0.0, ul1 = 0.0;
        for (int iy = 0; iy < P; iy++) ul0 += a.E[qy * P + iy] * ldro(u + (size_t)ex[iy * D::NP1 + qx] * ld);
        for (int ix = 0; ix < P; ix++) ul1 += a.E[qx * P + ix] * ldro(u + (size_t)ey[qy * P + ix] * ld);
        double lx[P + 1], ly[P + 1];
        departure_basis<P>(a, e, qx, qy, ul0, ul1, lx, ly);
        double v = 0.0;
#pragma unroll
        for (int jy = 0; jy <= P; jy++)
#pragma unroll
            for (int jx = 0; jx <= P; jx++) v += ldro(a.x + (size_t)e0[jy * D::NP1 + jx] * ld + k) * lx[jx] * ly[jy];
        acc += a.wq[q] * hl * v;
    }
    a.y[(size_t)n * ld + k] = f * acc;
}

// ---------------------------------------------------------------------------------------------
// Mass-matrix solves (KSPSolve(ksp1, ...) after almost every apply, eul/HorizSolve.cpp:77-96, 224, 310).
// M1 is symmetric positive definite, so the reference's GMRES + element-block Jacobi is replaced by a
// diagonally preconditioned conjugate-gradient iteration, batched over the levels (every level is its own system,
// so alpha and beta are per-level vectors) with the matrix-free M1 kernel as the operator.

// diag(M1) for the element's owned edges: y^x(ix,iy) = sum_qy E[qy][iy]^2 c Gaa (ix,qy) (+ west neighbour's far line),
//                                         y^y(ix,iy) = sum_qx E[qx][ix]^2 c Gbb (qx,iy) (+ south neighbour's far line)
template <int P>
__device__ __forceinline__ void m1_diag_far(const KArgs& a, int n, int side, bool rev, int k, double (&out)[P]) {
    using D = ElDim<P>;
    const int* __restrict__ nq = a.elq + (size_t)n * D::Q2;
    const double* __restrict__ G = a.G + (size_t)n * D::Q2 * 3;
    double o[P];
#pragma unroll
    for (int i = 0; i < P; i++) o[i] = 0.0;
#pragma unroll
    for (int t = 0; t <= P; t++) {
        const int q = side == 0 ? t * D::NP1 + P : P * D::NP1 + t;
        const double c = thick_factor(a, nq[q], k) * (side == 0 ? G[q * 3 + 0] : G[q * 3 + 2]);
#pragma unroll
        for (int i = 0; i < P; i++) o[i] += a.E[t * P + i] * a.E[t * P + i] * c;
    }
#pragma unroll
    for (int i = 0; i < P; i++) out[i] = rev ? o[P - 1 - i] : o[i];
}

// INVERT: store 1 / diag (the preconditioner)
template <int P, bool INVERT>
__global__ void __launch_bounds__(128) k_diag_m1(const __grid_constant__ KArgs a) {
    using D = ElDim<P>;
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)a.nel * (unsigned)a.nlev) return;
    int e = (int)fastdiv(idx, a.div_m, a.div_s);
    const int k = (int)(idx - (unsigned)e * (unsigned)a.nlev);
    if (a.elist) e = a.elist[e];
    const size_t ld = a.ld;
    double* __restrict__ y = a.y + k;
    const int* __restrict__ ex = a.el1x + (size_t)e * D::N1E;
    const int* __restrict__ ey = a.el1y + (size_t)e * D::N1E;
    const int* __restrict__ eq = a.elq + (size_t)e * D::Q2;
    const double* __restrict__ G = a.G + (size_t)e * D::Q2 * 3;
    double cw[P], cs[P];
#pragma unroll
    for (int i = 0; i < P; i++) cw[i] = cs[i] = 0.0;
    const int nw = a.nbr[2 * e + 0], ns = a.nbr[2 * e + 1];
    if (nw >= 0) m1_diag_far<P>(a, nw & 0x1fffffff, (nw >> 29) & 1, (nw >> 30) & 1, k, cw);
    if (ns >= 0) m1_diag_far<P>(a, ns & 0x1fffffff, (ns >> 29) & 1, (ns >> 30) & 1, k, cs);
    const unsigned flags = a.eflags[e];
    double dx[P][P + 1], dy[P + 1][P];
#pragma unroll
    for (int iy = 0; iy < P; iy++)
#pragma unroll
        for (int ix = 0; ix <= P; ix++) dx[iy][ix] = (ix == 0) ? cw[iy] : 0.0;
#pragma unroll
    for (int iy = 0; iy <= P; iy++)
#pragma unroll
        for (int ix = 0; ix < P; ix++) dy[iy][ix] = (iy == 0) ? cs[ix] : 0.0;
#pragma unroll
    for (int qy = 0; qy <= P; qy++)
#pragma unroll
        for (int qx = 0; qx <= P; qx++) {
            const int q = qy * D::NP1 + qx;
            const double c = thick_factor(a, eq[q], k);
            const double caa = c * G[q * 3 + 0], cbb = c * G[q * 3 + 2];
#pragma unroll
            for (int iy = 0; iy < P; iy++) dx[iy][qx] += a.E[qy * P + iy] * a.E[qy * P + iy] * caa;
#pragma unroll
            for (int ix = 0; ix < P; ix++) dy[qy][ix] += a.E[qx * P + ix] * a.E[qx * P + ix] * cbb;
        }
#pragma unroll
    for (int iy = 0; iy < P; iy++)
#pragma unroll
        for (int ix = 0; ix <= P; ix++)
            if (ix < P || (flags & 1u)) y[(size_t)ex[iy * D::NP1 + ix] * ld] = INVERT ? 1.0 / dx[iy][ix] : dx[iy][ix];
#pragma unroll
    for (int iy = 0; iy <= P; iy++)
#pragma unroll
        for (int ix = 0; ix < P; ix++)
            if (iy < P || (flags & 2u)) y[(size_t)ey[iy * P + ix] * ld] = INVERT ? 1.0 / dy[iy][ix] : dy[iy][ix];
}

// Batched CG building blocks.  Fields are [rows][ld] with lanes along the levels; every reduction is per level and
// deterministic: a block sums 256 rows into partial[block][k], k_cg_finish adds the partials in block order.
constexpr int CG_ROWS = 256;
constexpr int CG_FIN_G = 16;   // lane groups of k_cg_finish (threads = 64 * CG_FIN_G)
struct CgArgs {
    int64_t nrows;
    int nlev, ld, nblocks;
    double* x; double* r; double* p; const double* q; const double* dinv; const double* b;
    double* partial;      // [3][nblocks][64]
    double* scal;         // [8][64]: 0 rz, 1 pq, 2 rr, 3 bb, 4 rz_new, 5 alpha, 6 beta, 7 frozen (converged levels)
    double tol2;
    // element-partitioned solve (N GPUs): every rank contributes its sums over OWNED rows; k_cg_finish exchanges them
    // over peer memory and adds them in rank order, so that every rank computes bit-identical step lengths
    int world, rank;
    uint4* const* peer_area;          // [world] reduction areas (mine at [rank]): [2 parities][world][3][64] 16-byte cells
    unsigned long long* seq;          // device counter of reductions performed so far (same on every rank)
    int* err;
};

// 16-byte self-validating cells {lo32, tag, hi32, tag} (the in-band protocol of the fused ghost refresh, m1_tile.cuh)
__device__ __forceinline__ void cg_cell_store(uint4* cell, double v, unsigned tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(cell), "r"((unsigned)b), "r"(tag), "r"((unsigned)(b >> 32)), "r"(tag)
                 : "memory");
}
__device__ __forceinline__ double cg_cell_load(const uint4* cell, unsigned tag, int* err) {
    unsigned x, f0, y, f1;
    long long t0 = 0;
    for (unsigned spin = 0;; spin++) {
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(f0), "=r"(y), "=r"(f1) : "l"(cell) : "memory");
        if (f0 == tag && f1 == tag) break;
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

__device__ __forceinline__ void cg_block_reduce(double v0, double v1, double v2, double* partial, int nblocks, int nsums) {
    __shared__ double red[3][4][64];
    const int k = threadIdx.x & 63, g = threadIdx.x >> 6;
    red[0][g][k] = v0; red[1][g][k] = v1; red[2][g][k] = v2;
    __syncthreads();
    if (g == 0) {
        for (int s = 0; s < nsums; s++)
            partial[((size_t)s * nblocks + blockIdx.x) * 64 + k] = ((red[s][0][k] + red[s][1][k]) + red[s][2][k]) + red[s][3][k];
    }
}

// MODE 0: r = b, x = 0, z = dinv r, p = z; sums rz, bb, rr.   MODE 1: sums pq.
// MODE 2: alpha = rz/pq; x += alpha p; r -= alpha q; sums rz_new (r . dinv r), rr.   MODE 3: beta = rz_new/rz; p = dinv r + beta p.
template <int MODE>
__global__ void __launch_bounds__(256) k_cg_step(const __grid_constant__ CgArgs a) {
    const int k = threadIdx.x & 63, g = threadIdx.x >> 6;
    const int64_t r0 = (int64_t)blockIdx.x * CG_ROWS;
    const int64_t r1 = r0 + CG_ROWS < a.nrows ? r0 + CG_ROWS : a.nrows;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    if (k < a.nlev) {
        double alpha = 0.0, beta = 0.0;
        if (MODE == 2) alpha = a.scal[5 * 64 + k];
        if (MODE == 3) beta = a.scal[6 * 64 + k];
        // U rows per pass, every load issued before the first store: the field pointers may alias as far as the compiler
        // knows, so a row-by-row loop kept ONE row of each field in flight per thread (45 % of the HBM rate)
        constexpr int U = 8;
        for (int64_t rb = r0 + g; rb < r1; rb += 4 * U) {
            size_t i[U];
            bool ok[U];
            double v0[U], v1[U], v2[U], v3[U], v4[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int64_t row = rb + 4 * u;
                ok[u] = row < r1;
                i[u] = (size_t)(ok[u] ? row : rb) * a.ld + k;
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (MODE == 0) { v0[u] = a.b[i[u]]; v1[u] = a.dinv[i[u]]; }
                else if (MODE == 1) { v0[u] = a.p[i[u]]; v1[u] = a.q[i[u]]; }
                else if (MODE == 2) { v0[u] = a.x[i[u]]; v1[u] = a.p[i[u]]; v2[u] = a.r[i[u]]; v3[u] = a.q[i[u]]; v4[u] = a.dinv[i[u]]; }
                else { v0[u] = a.dinv[i[u]]; v1[u] = a.r[i[u]]; v2[u] = a.p[i[u]]; }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (!ok[u]) continue;
                if (MODE == 0) {
                    const double bv = v0[u], z = v1[u] * bv;
                    a.x[i[u]] = 0.0; a.r[i[u]] = bv; a.p[i[u]] = z;
                    s0 += bv * z; s1 += bv * bv; s2 += bv * bv;
                } else if (MODE == 1) {
                    s0 += v0[u] * v1[u];
                } else if (MODE == 2) {
                    a.x[i[u]] = v0[u] + alpha * v1[u];
                    const double rv = v2[u] - alpha * v3[u];
                    a.r[i[u]] = rv;
                    s0 += rv * (v4[u] * rv); s1 += rv * rv;
                } else {
                    a.p[i[u]] = v0[u] * v1[u] + beta * v2[u];
                }
            }
        }
    }
    if (MODE != 3) cg_block_reduce(s0, s1, s2, a.partial, a.nblocks, MODE == 0 ? 3 : (MODE == 1 ? 1 : 2));
}

// one thread per level: add the block partials in order and update the per-level scalars
template <int MODE>
__global__ void k_cg_finish(const __grid_constant__ CgArgs a) {
    // CG_FIN_G groups of 64 level lanes: group g adds the block partials g, g + G, ... in order, then lane group 0 adds the G group
    // sums in order (fixed order: deterministic).  One group alone walked 1 700 partials per sum, 0.2 ms per call on C5.
    const int k = threadIdx.x & 63, grp = threadIdx.x >> 6;
    const bool active = k < a.nlev;
    constexpr int NS = MODE == 0 ? 3 : (MODE == 1 ? 1 : 2);
    __shared__ double part[3][CG_FIN_G][64];
    {
        double t[3] = {0.0, 0.0, 0.0};
        if (active)
            for (int bI = grp; bI < a.nblocks; bI += CG_FIN_G)
#pragma unroll
                for (int j = 0; j < NS; j++) t[j] += a.partial[((size_t)j * a.nblocks + bI) * 64 + k];
#pragma unroll
        for (int j = 0; j < NS; j++) part[j][grp][k] = t[j];
    }
    __syncthreads();
    double s[3] = {0.0, 0.0, 0.0};
    const int nsums = !active ? 0 : NS;
    if (grp == 0)
        for (int j = 0; j < nsums; j++)
            for (int g2 = 0; g2 < CG_FIN_G; g2++) s[j] += part[j][g2][k];
    if (a.world > 1 && grp > 0) {
        __syncthreads();   // (the barrier of the exchange below)
        return;
    }
    if (grp > 0) return;
    if (a.world > 1) {
        // all-gather of the per-rank sums over peer memory, then the same rank-ordered sum on every rank.  Two parities:
        // a rank can be at most one reduction ahead of a peer (it cannot finish the next one without that peer's sums).
        const unsigned long long n = *a.seq + 1;
        const unsigned tag = (unsigned)n;
        const size_t base = (size_t)(n & 1) * a.world * 3 * 64;
        for (int j = 0; j < nsums; j++)
            for (int r = 0; r < a.world; r++)
                if (r != a.rank) cg_cell_store(a.peer_area[r] + base + ((size_t)a.rank * 3 + j) * 64 + k, s[j], tag);
        for (int j = 0; j < nsums; j++) {
            double t = 0.0;
            for (int r = 0; r < a.world; r++)
                t += (r == a.rank) ? s[j] : cg_cell_load(a.peer_area[a.rank] + base + ((size_t)r * 3 + j) * 64 + k, tag, a.err);
            s[j] = t;
        }
        __syncthreads();
        if (k == 0) *a.seq = n;
    }
    double* S = a.scal;
    if (!active) return;
    if (MODE == 0) {
        S[0 * 64 + k] = s[0]; S[3 * 64 + k] = s[1]; S[2 * 64 + k] = s[2];
        S[7 * 64 + k] = (s[1] == 0.0) ? 1.0 : 0.0;                    // zero right-hand side: x = 0 is the solution
    } else if (MODE == 1) {
        S[1 * 64 + k] = s[0];
        const bool frozen = S[7 * 64 + k] != 0.0 || s[0] <= 0.0;
        S[5 * 64 + k] = frozen ? 0.0 : S[0 * 64 + k] / s[0];          // alpha
    } else {
        const bool frozen = S[7 * 64 + k] != 0.0;
        S[6 * 64 + k] = (frozen || S[0 * 64 + k] == 0.0) ? 0.0 : s[0] / S[0 * 64 + k];   // beta = rz_new / rz
        if (!frozen) {
            S[0 * 64 + k] = s[0];
            S[2 * 64 + k] = s[1];
            if (s[1] <= a.tol2 * S[3 * 64 + k]) S[7 * 64 + k] = 1.0;   // converged: this level stops moving
        }
    }
}

// y = M2 x   (WITH_H: M2(rho) x)
template <int P, bool WITH_H>
__global__ void __launch_bounds__(128) k_apply_m2(const __grid_constant__ KArgs a) {
    using D = ElDim<P>;
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)a.nel * (unsigned)a.nlev) return;
    int e = (int)fastdiv(idx, a.div_m, a.div_s);
    const int k = (int)(idx - (unsigned)e * (unsigned)a.nlev);
    if (a.elist) e = a.elist[e];
    const size_t ld = a.ld;
    const double* __restrict__ x = a.x + k;
    double* __restrict__ y = a.y + k;
    const int* __restrict__ e2 = a.el2 + (size_t)e * D::N2E;
    const int* __restrict__ eq = a.elq + (size_t)e * D::Q2;
    const double* __restrict__ G = a.G + (size_t)e * D::Q2;

    double ax[P][P + 1];   // x contracted in x-direction: ax[iy][qx]
    double rx[P][P + 1];
#pragma unroll
    for (int iy = 0; iy < P; iy++) {
        double hv[P], rv[P];
#pragma unroll
        for (int ix = 0; ix < P; ix++) {
            hv[ix] = ldro(x + (size_t)e2[iy * P + ix] * ld);
            if (WITH_H) rv[ix] = ldro(a.c + k + (size_t)e2[iy * P + ix] * ld);
        }
#pragma unroll
        for (int qx = 0; qx <= P; qx++) {
            double s = 0.0, r = 0.0;
#pragma unroll
            for (int ix = 0; ix < P; ix++) {
                s += a.E[qx * P + ix] * hv[ix];
                if (WITH_H) r += a.E[qx * P + ix] * rv[ix];
            }
            ax[iy][qx] = s;
            if (WITH_H) rx[iy][qx] = r;
        }
    }
    double out[P][P];
#pragma unroll
    for (int iy = 0; iy < P; iy++)
#pragma unroll
        for (int ix = 0; ix < P; ix++) out[iy][ix] = 0.0;
#pragma unroll
    for (int qx = 0; qx <= P; qx++) {
        double g[P + 1];
#pragma unroll
        for (int qy = 0; qy <= P; qy++) {
            double hl = 0.0, rl = 0.0;
#pragma unroll
            for (int iy = 0; iy < P; iy++) {
                hl += a.E[qy * P + iy] * ax[iy][qx];
                if (WITH_H) rl += a.E[qy * P + iy] * rx[iy][qx];
            }
            const int q = qy * D::NP1 + qx;
            double c = thick_factor(a, eq[q], k) * G[q];
            if (WITH_H) c *= rl;
            g[qy] = c * hl;
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
#pragma unroll
    for (int iy = 0; iy < P; iy++)
#pragma unroll
        for (int ix = 0; ix < P; ix++) y[(size_t)e2[iy * P + ix] * ld] = out[iy][ix];
}

// y = K(u1) x : 1-form -> 2-form, coefficient a.c = u1 (1-form, same array indexing as x)
template <int P>
__global__ void __launch_bounds__(128) k_apply_k(const __grid_constant__ KArgs a) {
    using D = ElDim<P>;
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)a.nel * (unsigned)a.nlev) return;
    int e = (int)fastdiv(idx, a.div_m, a.div_s);
    const int k = (int)(idx - (unsigned)e * (unsigned)a.nlev);
    if (a.elist) e = a.elist[e];
    const size_t ld = a.ld;
    const double* __restrict__ x = a.x + k;
    const double* __restrict__ u = a.c + k;
    double* __restrict__ y = a.y + k;
    const int* __restrict__ ex = a.el1x + (size_t)e * D::N1E;
    const int* __restrict__ ey = a.el1y + (size_t)e * D::N1E;
    const int* __restrict__ e2 = a.el2 + (size_t)e * D::N2E;
    const int* __restrict__ eq = a.elq + (size_t)e * D::Q2;
    const double* __restrict__ G = a.G + (size_t)e * D::Q2 * 3;

    double xy[P + 1][P], uy[P + 1][P];
#pragma unroll
    for (int iy = 0; iy <= P; iy++)
#pragma unroll
        for (int ix = 0; ix < P; ix++) {
            const size_t o = (size_t)ey[iy * P + ix] * ld;
            xy[iy][ix] = ldro(x + o);
            uy[iy][ix] = ldro(u + o);
        }
    double out[P][P];
#pragma unroll
    for (int iy = 0; iy < P; iy++)
#pragma unroll
        for (int ix = 0; ix < P; ix++) out[iy][ix] = 0.0;
#pragma unroll
    for (int qx = 0; qx <= P; qx++) {
        double xc[P], uc[P];
#pragma unroll
        for (int iy = 0; iy < P; iy++) {
            const size_t o = (size_t)ex[iy * D::NP1 + qx] * ld;
            xc[iy] = ldro(x + o);
            uc[iy] = ldro(u + o);
        }
        double g[P + 1];
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
                x1 += a.E[qx * P + ix] * xy[qy][ix];
                a1 += a.E[qx * P + ix] * uy[qy][ix];
            }
            const int q = qy * D::NP1 + qx;
            const double c = 0.5 * thick_factor(a, eq[q], k);
            const double ka = G[q * 3 + 0] * a0 + G[q * 3 + 1] * a1;
            const double kb = G[q * 3 + 1] * a0 + G[q * 3 + 2] * a1;
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
#pragma unroll
    for (int iy = 0; iy < P; iy++)
#pragma unroll
        for (int ix = 0; ix < P; ix++) y[(size_t)e2[iy * P + ix] * ld] = out[iy][ix];
}

// 0-form mass matrix.  With m == p the nodal tabulation is the identity, so M0 is diagonal:
//   y_n = s t^tpow x_n sum_{(e,q) at n} w_q det_{e,q}                       (eul/Assembly.cpp:2021-2036)
//   assemble_h: y_n = s t^2 x_n sum_{(e,q) at n} w_q hl^e_q(h)              (eul/Assembly.cpp:2067-2086)
template <int P, bool WITH_H>
__global__ void __launch_bounds__(128) k_apply_m0(const __grid_constant__ NodeArgs a) {
    using D = ElDim<P>;
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)a.n0 * (unsigned)a.nlev) return;
    const int n = (int)fastdiv(idx, a.div_m, a.div_s);
    const int k = (int)(idx - (unsigned)n * (unsigned)a.nlev);
    const size_t ld = a.ld;
    double f = a.scale;
    if (a.tpow > 0) {
        const double t = ldro(a.tinv + (size_t)a.node_q[n] * a.nkT + a.lev0 + k * a.lev_stride);
        f *= t;
        if (a.tpow > 1) f *= t;
    }
    double d;
    if (!WITH_H) {
        d = a.D0[n];
    } else {
        d = 0.0;
        const double* __restrict__ h = a.c + k;
        for (int j = a.adj_ptr[n]; j < a.adj_ptr[n + 1]; j++) {
            const int eqv = a.adj_eq[j];
            const int e = eqv / D::Q2, q = eqv - e * D::Q2;
            const int qx = q % D::NP1, qy = q / D::NP1;
            const int* __restrict__ e2 = a.el2 + (size_t)e * D::N2E;
            double hl = 0.0;
#pragma unroll
            for (int iy = 0; iy < P; iy++) {
                double s = 0.0;
#pragma unroll
                for (int ix = 0; ix < P; ix++) s += a.E[qx * P + ix] * ldro(h + (size_t)e2[iy * P + ix] * ld);
                hl += a.E[qy * P + iy] * s;
            }
            d += a.wq[q] * hl;
        }
    }
    // x == nullptr: the diagonal itself (Pvec::assemble, Phvec::assemble, eul/Assembly.cpp:602-628, 652-681)
    a.y[(size_t)n * ld + k] = f * d * (a.x ? ldro(a.x + (size_t)n * ld + k) : 1.0);
}

// x = M0^-1 b: M0 is diagonal when the quadrature order equals the element order (KSPSolve(ksp0, ...) of
// eul/HorizSolve.cpp:87-96, 246 becomes a pointwise division)
static __global__ void __launch_bounds__(256) k_solve_m0(const __grid_constant__ NodeArgs a) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)a.n0 * (unsigned)a.nlev) return;
    const int n = (int)fastdiv(idx, a.div_m, a.div_s);
    const int k = (int)(idx - (unsigned)n * (unsigned)a.nlev);
    double f = a.scale;
    if (a.tpow > 0) {
        const double t = ldro(a.tinv + (size_t)a.node_q[n] * a.nkT + a.lev0 + k * a.lev_stride);
        f *= t;
        if (a.tpow > 1) f *= t;
    }
    a.y[(size_t)n * a.ld + k] = ldro(a.x + (size_t)n * a.ld + k) / (f * a.D0[n]);
}

// M0 without a coefficient field, two levels per thread (16-byte accesses); a.nlev counts level pairs
static __global__ void __launch_bounds__(256) k_apply_m0_vec2(const __grid_constant__ NodeArgs a) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)a.n0 * (unsigned)a.nlev) return;
    const int n = (int)fastdiv(idx, a.div_m, a.div_s);
    const int k = (int)(idx - (unsigned)n * (unsigned)a.nlev) * 2;
    const double2 x = __ldg(reinterpret_cast<const double2*>(a.x + (size_t)n * a.ld + k));
    double2 f = make_double2(a.scale, a.scale);
    if (a.tpow > 0) {
        const double2 t = __ldg(reinterpret_cast<const double2*>(a.tinv + (size_t)a.node_q[n] * a.nkT + a.lev0 + k));
        f.x *= t.x; f.y *= t.y;
        if (a.tpow > 1) { f.x *= t.x; f.y *= t.y; }
    }
    const double d = a.D0[n];
    *reinterpret_cast<double2*>(a.y + (size_t)n * a.ld + k) = make_double2(f.x * d * x.x, f.y * d * x.y);
}

// y[r][k] = sum_j sgn[r][j] x[col[r][j]][k]   (entries in a partition-invariant order, see upload_ell)
// VEC = 2 / 4: two / four consecutive levels per thread through 16-byte accesses (nlev, ld multiples of VEC, 16-byte
// aligned fields); a.nlev then counts level GROUPS.  The kernels are bound by instruction issue per thread (index
// loads, address arithmetic), not by HBM, so fewer, fatter threads are faster.
template <int VEC>
__global__ void __launch_bounds__(256) k_apply_ell(const __grid_constant__ EllArgs a) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)a.nrows * (unsigned)a.nlev) return;
    const int64_t rr = fastdiv(idx, a.div_m, a.div_s);
    const int k = (int)(idx - (unsigned)rr * (unsigned)a.nlev) * VEC;
    const int64_t r = a.rows ? a.rows[rr] : rr;
    int cols[4];
    signed char sg[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        cols[j] = j < a.width ? a.col[r * a.width + j] : -1;
        sg[j] = j < a.width ? a.sgn[r * a.width + j] : 0;
    }
    if (VEC >= 2) {
        constexpr int NV = VEC >= 2 ? VEC / 2 : 1;
        double2 v[4][NV];
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (cols[j] >= 0) {
                const double2* src = reinterpret_cast<const double2*>(a.x + (size_t)cols[j] * a.ld + k);
#pragma unroll
                for (int u = 0; u < NV; u++) v[j][u] = __ldg(src + u);
            }
        double2 s[NV];
#pragma unroll
        for (int u = 0; u < NV; u++) s[u] = make_double2(0.0, 0.0);
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (cols[j] >= 0) {
#pragma unroll
                for (int u = 0; u < NV; u++) {
                    s[u].x += sg[j] > 0 ? v[j][u].x : -v[j][u].x;
                    s[u].y += sg[j] > 0 ? v[j][u].y : -v[j][u].y;
                }
            }
        double2* dst = reinterpret_cast<double2*>(a.y + (size_t)r * a.ld + k);
#pragma unroll
        for (int u = 0; u < NV; u++) dst[u] = s[u];
    } else {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (cols[j] >= 0) {
                const double v = ldro(a.x + (size_t)cols[j] * a.ld + k);
                s += sg[j] > 0 ? v : -v;
            }
        a.y[(size_t)r * a.ld + k] = s;
    }
}

// PUSH (HALO_NB CTAs per send peer): field rows -> the peer's inbox; the last CTA to finish raises the peer's flag
// to the new epoch.  PULL (HALO_NB CTAs per receive peer): wait for the peer's flag, inbox -> ghost rows; the last
// CTA acknowledges.  counters[peer] counts finished CTAs (reset by the last one).
constexpr int HALO_NB = 96;
template <bool PUSH>
__global__ void __launch_bounds__(256) k_halo(const HaloPeer* __restrict__ peers, int nlev, int ld, int nbuf, double* field,
                                               const unsigned long long* __restrict__ epoch, unsigned* counters, int* err) {
    const HaloPeer p = peers[blockIdx.x];
    const unsigned long long e = *epoch + 1;
    // ONE buffering rule for this kernel and the fused M1 launch (they share inbox, flags and epochs of a space):
    // data epoch e lives in inbox copy e % nbuf; a push of epoch e waits for the acknowledgement of epoch e - nbuf
    double* box = p.inbox + (e % (unsigned long long)nbuf) * p.inbox_parity_stride + (size_t)p.row0 * nlev;
    if (threadIdx.x == 0) {
        // PUSH: the receiver must have consumed the copy this push overwrites ; PULL: the data of epoch e must have landed
        if (PUSH) {
            if (e > (unsigned long long)nbuf) spin_until(p.wait, e - nbuf, err);
        } else {
            spin_until(p.wait, e, err);
        }
    }
    __syncthreads();
    const int total = p.nrows * nlev;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < total; i += gridDim.y * blockDim.x) {
        const int r = i / nlev, k = i - r * nlev;
        if (PUSH) box[i] = field[(size_t)p.rows[r] * ld + k];
        else field[(size_t)p.rows[r] * ld + k] = box[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned done = atomicAdd(&counters[blockIdx.x], 1u);
        if (done == gridDim.y - 1) {
            counters[blockIdx.x] = 0;
            __threadfence_system();
            st_release_sys(p.signal, e);
        }
    }
}
static __global__ void k_epoch_inc(unsigned long long* epoch) { *epoch += 1; }

// Halo pack / unpack: packed[i*nlev + k] <-> field[rows[i]*ld + k]   (rows = ghost or send lists)
template <bool GATHER>
__global__ void __launch_bounds__(256) k_rows(int64_t nrows, int nlev, int ld, unsigned div_m, unsigned div_s,
                                              const int* __restrict__ rows, const double* __restrict__ src,
                                              double* __restrict__ dst) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)nrows * (unsigned)nlev) return;
    const unsigned i = fastdiv(idx, div_m, div_s);
    const unsigned k = idx - i * (unsigned)nlev;
    if (GATHER) dst[idx] = src[(size_t)rows[i] * ld + k];
    else dst[(size_t)rows[i] * ld + k] = src[idx];
}

// E21 (div) and E12 = -E21^T (weak gradient) element by element, without index loads: one record of 2 + 4 P row numbers per
// owned element -- its first edge row and first face row (both blocks are contiguous in the engine's numbering), its
// east column / north row of edges (E21) and the faces across its west column / south row (E12; -1 where the mesh
// ends) -- and closed-form positions inside the blocks.  A thread owns one element and two adjacent levels (16-byte
// accesses, lanes along the levels).  The sums run in the order of the reference's MatSetValues calls (E21: -x(a,b)
// +x(a+1,b) -y(a,b) +y(a,b+1), eul/Assembly.cpp:1196-1205; E12: the edge's own face first), exactly as the ELL kernel
// does, so integer data stays exact and the two kernels agree bitwise.
struct IncArgs {
    int nel, nl2, ld;            // owned elements, level PAIRS per row, leading dimension
    const int* recs;             // [nel][2 + 4 P]: e1, f0, xe[P], yn[P], wf[P], sf[P]
    const double* x;
    double* y;
};

template <int P, bool DIV>
__global__ void __launch_bounds__(128) k_inc_tile(const __grid_constant__ IncArgs a) {
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (e >= a.nel || lane >= a.nl2) return;
    const int* __restrict__ r = a.recs + (size_t)e * (2 + 4 * P);
    const int e1 = r[0], f0 = r[1];
    const size_t ld = a.ld;
    const double2* __restrict__ x = reinterpret_cast<const double2*>(a.x) + lane;
    double2* __restrict__ y = reinterpret_cast<double2*>(a.y) + lane;
    auto ldx = [&](int row) { return __ldg(x + (size_t)row * (ld >> 1)); };
    const double2 zero = make_double2(0.0, 0.0);
    if (DIV) {
        // faces row by row: y_f(a,b) = ((0 - xx(a,b)) + xx(a+1,b)) - xy(a,b)) + xy(a,b+1)
        double2 ylo[P];
#pragma unroll
        for (int a_ = 0; a_ < P; a_++) ylo[a_] = ldx(e1 + P * P + a_);   // xy(a, 0)
#pragma unroll
        for (int b = 0; b < P; b++) {
            double2 xv[P + 1], yhi[P];
#pragma unroll
            for (int a_ = 0; a_ < P; a_++) xv[a_] = ldx(e1 + a_ * P + b);            // xx(a, b): column-major block
            xv[P] = ldx(r[2 + b]);                                                   // east column
#pragma unroll
            for (int a_ = 0; a_ < P; a_++) yhi[a_] = (b + 1 < P) ? ldx(e1 + P * P + (b + 1) * P + a_) : ldx(r[2 + P + a_]);   // xy(a, b+1)
#pragma unroll
            for (int a_ = 0; a_ < P; a_++) {
                double2 s;
                s.x = 0.0 - xv[a_].x; s.y = 0.0 - xv[a_].y;
                s.x += xv[a_ + 1].x;  s.y += xv[a_ + 1].y;
                s.x -= ylo[a_].x;     s.y -= ylo[a_].y;
                s.x += yhi[a_].x;     s.y += yhi[a_].y;
                y[(size_t)(f0 + b * P + a_) * (ld >> 1)] = s;
                ylo[a_] = yhi[a_];
            }
        }
    } else {
        // edges: y_x(a,b) = (0 + f(a,b)) - f(a-1,b) ; y_y(a,b) = (0 + f(a,b)) - f(a,b-1)
        double2 f[P][P];
#pragma unroll
        for (int b = 0; b < P; b++)
#pragma unroll
            for (int a_ = 0; a_ < P; a_++) f[b][a_] = ldx(f0 + b * P + a_);
#pragma unroll
        for (int b = 0; b < P; b++) {
            const int wr = r[2 + 2 * P + b];
            const double2 w = wr >= 0 ? ldx(wr) : zero;
#pragma unroll
            for (int a_ = 0; a_ < P; a_++) {
                const double2 o = a_ > 0 ? f[b][a_ - 1] : w;
                double2 s;
                s.x = (0.0 + f[b][a_].x) - o.x; s.y = (0.0 + f[b][a_].y) - o.y;
                if (a_ == 0 && wr < 0) s = f[b][a_];
                y[(size_t)(e1 + a_ * P + b) * (ld >> 1)] = s;
            }
        }
#pragma unroll
        for (int a_ = 0; a_ < P; a_++) {
            const int sr = r[2 + 3 * P + a_];
            const double2 sv = sr >= 0 ? ldx(sr) : zero;
#pragma unroll
            for (int b = 0; b < P; b++) {
                const double2 o = b > 0 ? f[b - 1][a_] : sv;
                double2 s;
                s.x = (0.0 + f[b][a_].x) - o.x; s.y = (0.0 + f[b][a_].y) - o.y;
                if (b == 0 && sr < 0) s = f[b][a_];
                y[(size_t)(e1 + P * P + b * P + a_) * (ld >> 1)] = s;
            }
        }
    }
}

// L2Vecs::HorizToVert / VertToHoriz (eul/L2Vecs.cpp:55-101): 2-form fields between the engine's column layout
// cols[el2[e][i]*ld + k] and the reference's per-element vertical vectors vert[e][k*p2 + i] (vz[ei] of size nk*p2, all
// elements back to back).  A pure relabelling -- bit exact -- through a padded shared-memory tile, one CTA per element,
// so that both sides are accessed in runs.
template <bool TO_VERT>
__global__ void __launch_bounds__(256) k_l2vecs(int p2, int nlev, int ld, const int* __restrict__ el2, const double* __restrict__ in,
                                                double* __restrict__ out) {
    extern __shared__ double l2tile[];   // [nlev][p2 + 1]
    const int e = blockIdx.x, n = p2 * nlev, pad = p2 + 1;
    const int* __restrict__ rows = el2 + (size_t)e * p2;
    if (TO_VERT) {
        for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
            const int i = idx / nlev, k = idx - i * nlev;
            l2tile[k * pad + i] = in[(size_t)rows[i] * ld + k];
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
            const int k = idx / p2, i = idx - k * p2;
            out[(size_t)e * n + idx] = l2tile[k * pad + i];
        }
    } else {
        for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
            const int k = idx / p2, i = idx - k * p2;
            l2tile[k * pad + i] = in[(size_t)e * n + idx];
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
            const int i = idx / nlev, k = idx - i * nlev;
            out[(size_t)rows[i] * ld + k] = l2tile[k * pad + i];
        }
    }
}

// levels[k*n + dof] <-> columns[perm[dof]*ld + k] through a padded shared-memory tile
// (perm = the engine's internal numbering of the space, nullptr = identity)
template <bool TO_COLUMNS>
__global__ void __launch_bounds__(256) k_transpose(int64_t n, int nlev, int ld, const int* __restrict__ perm,
                                                   const double* __restrict__ in, double* __restrict__ out) {
    __shared__ double tile[32][33];
    const int64_t d0 = (int64_t)blockIdx.x * 32;
    const int k0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    if (TO_COLUMNS) {
        for (int j = ty; j < 32; j += 8) {
            const int k = k0 + j;
            const int64_t d = d0 + tx;
            if (k < nlev && d < n) tile[j][tx] = in[(size_t)k * n + d];
        }
        __syncthreads();
        for (int j = ty; j < 32; j += 8) {
            const int64_t d = d0 + j;
            const int k = k0 + tx;
            if (k < nlev && d < n) out[(size_t)(perm ? perm[d] : d) * ld + k] = tile[tx][j];
        }
    } else {
        for (int j = ty; j < 32; j += 8) {
            const int64_t d = d0 + j;
            const int k = k0 + tx;
            if (k < nlev && d < n) tile[j][tx] = in[(size_t)(perm ? perm[d] : d) * ld + k];
        }
        __syncthreads();
        for (int j = ty; j < 32; j += 8) {
            const int k = k0 + j;
            const int64_t d = d0 + tx;
            if (k < nlev && d < n) out[(size_t)k * n + d] = tile[tx][j];
        }
    }
}

}  // namespace mimsem
