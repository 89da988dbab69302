// x = M2^-1 b, element by element: WmatInv::assemble(lev, scale) / WhmatInv::assemble(rho, lev, scale) + MatMult
// (eul/Assembly.cpp:1658-1722, 1730-1800).  The 2-form mass matrix is block diagonal -- one p^2 x p^2 symmetric positive
// definite block  B = W^T diag(c) W  per element and level -- and the reference inverts every block with Gauss-Jordan
// elimination (LinAlg Inv).  Here one thread owns one (element, level) pair: it tabulates the lower triangle of its block
// in shared memory (entry-major, lanes = levels: conflict free), factorises it in place (L D L^T) and solves; the block
// never exists in global memory.  The tabulation exploits the tensor structure of W:
//   B[(iy,ix),(jy,jx)] = sum_qy E[qy][iy] E[qy][jy] ( sum_qx E[qx][ix] E[qx][jx] c[qy][qx] ).
#pragma once
#include "kernels.cuh"

namespace mimsem {

template <int P>
struct M2SolveSmem {
    static constexpr int N = P * P;
    static constexpr int TRI = N * (N + 1) / 2;
    static constexpr int PAIRS = P * (P + 1) / 2;
    static constexpr int LANES = 32;
    // per lane: TRI block entries + (P+1) x PAIRS partial sums + N right-hand side / solution
    static constexpr int DOUBLES = (TRI + (P + 1) * PAIRS + N) * LANES;
};

template <int P, bool WITH_H>
__global__ void __launch_bounds__(32) k_solve_m2(const __grid_constant__ KArgs a) {
    using D = ElDim<P>;
    using S = M2SolveSmem<P>;
    constexpr int N = S::N, NP1 = P + 1, L = S::LANES, PAIRS = S::PAIRS;
    extern __shared__ double sm2[];
    double* B = sm2 + threadIdx.x;                     // B[t * L], t = i (i+1)/2 + j, i >= j
    double* T = sm2 + (size_t)S::TRI * L + threadIdx.x;   // T[(qy * PAIRS + pr) * L]
    double* R = T + (size_t)NP1 * PAIRS * L;             // R[i * L]
    const unsigned idx = blockIdx.x * L + threadIdx.x;
    if (idx >= (unsigned)a.nel * (unsigned)a.nlev) return;
    int e = (int)fastdiv(idx, a.div_m, a.div_s);
    const int k = (int)(idx - (unsigned)e * (unsigned)a.nlev);
    if (a.elist) e = a.elist[e];
    const size_t ld = a.ld;
    const int* __restrict__ e2 = a.el2 + (size_t)e * D::N2E;
    const int* __restrict__ eq = a.elq + (size_t)e * D::Q2;
    const double* __restrict__ G = a.G + (size_t)e * D::Q2;
    // point weights c[qy][qx] = s w/det t^tpow [rho_l/det]
    double c[NP1][NP1];
    double rx[WITH_H ? P : 1][WITH_H ? NP1 : 1];
    if (WITH_H) {
#pragma unroll
        for (int iy = 0; iy < P; iy++) {
            double rv[P];
#pragma unroll
            for (int ix = 0; ix < P; ix++) rv[ix] = ldro(a.c + k + (size_t)e2[iy * P + ix] * ld);
#pragma unroll
            for (int qx = 0; qx <= P; qx++) {
                double r = 0.0;
#pragma unroll
                for (int ix = 0; ix < P; ix++) r += a.E[qx * P + ix] * rv[ix];
                rx[WITH_H ? iy : 0][WITH_H ? qx : 0] = r;
            }
        }
    }
#pragma unroll
    for (int qy = 0; qy <= P; qy++)
#pragma unroll
        for (int qx = 0; qx <= P; qx++) {
            const int q = qy * NP1 + qx;
            double v = thick_factor(a, eq[q], k) * G[q];
            if (WITH_H) {
                double rl = 0.0;
#pragma unroll
                for (int iy = 0; iy < P; iy++) rl += a.E[qy * P + iy] * rx[WITH_H ? iy : 0][WITH_H ? qx : 0];
                v *= rl;
            }
            c[qy][qx] = v;
        }
    // T[qy][(ix >= jx)] = sum_qx E[qx][ix] E[qx][jx] c[qy][qx]
#pragma unroll
    for (int qy = 0; qy <= P; qy++) {
        int pr = 0;
#pragma unroll
        for (int ix = 0; ix < P; ix++)
#pragma unroll
            for (int jx = 0; jx <= ix; jx++, pr++) {
                double s = 0.0;
#pragma unroll
                for (int qx = 0; qx <= P; qx++) s += (a.E[qx * P + ix] * a.E[qx * P + jx]) * c[qy][qx];
                T[(size_t)(qy * PAIRS + pr) * L] = s;
            }
    }
    // lower triangle of the block, rows i = iy P + ix >= j = jy P + jx
    for (int i = 0; i < N; i++) {
        const int iy = i / P, ix = i - iy * P;
        for (int j = 0; j <= i; j++) {
            const int jy = j / P, jx = j - jy * P;
            const int hi = ix > jx ? ix : jx, lo = ix > jx ? jx : ix;
            const int pr = hi * (hi + 1) / 2 + lo;
            double s = 0.0;
#pragma unroll
            for (int qy = 0; qy <= P; qy++) s += (a.E[qy * P + iy] * a.E[qy * P + jy]) * T[(size_t)(qy * PAIRS + pr) * L];
            B[(size_t)(i * (i + 1) / 2 + j) * L] = s;
        }
    }
    // in-place B = L D L^T (unit lower L below the diagonal, 1/D on it); no square roots, so a block that is not positive
    // definite (a coefficient whose interpolant changes sign) still factorises as long as no pivot vanishes
    for (int i = 0; i < N; i++) {
        const int ri = i * (i + 1) / 2;
        // row i holds W_it = L_it D_t until the row is finished
        for (int j = 0; j < i; j++) {
            const int rj = j * (j + 1) / 2;
            double s = B[(size_t)(ri + j) * L];
            for (int t = 0; t < j; t++) s -= B[(size_t)(ri + t) * L] * B[(size_t)(rj + t) * L];
            B[(size_t)(ri + j) * L] = s;
        }
        double d = B[(size_t)(ri + i) * L];
        for (int t = 0; t < i; t++) {
            const double w = B[(size_t)(ri + t) * L];
            const double l = w * B[(size_t)(t * (t + 1) / 2 + t) * L];   // L_it = W_it / D_t
            d -= w * l;
            B[(size_t)(ri + t) * L] = l;
        }
        B[(size_t)(ri + i) * L] = 1.0 / d;
    }
    // L z = b, z /= D, L^T x = z
    for (int i = 0; i < N; i++) R[(size_t)i * L] = ldro(a.x + k + (size_t)e2[i] * ld);
    for (int i = 0; i < N; i++) {
        const int ri = i * (i + 1) / 2;
        double s = R[(size_t)i * L];
        for (int t = 0; t < i; t++) s -= B[(size_t)(ri + t) * L] * R[(size_t)t * L];
        R[(size_t)i * L] = s;
    }
    for (int i = 0; i < N; i++) R[(size_t)i * L] *= B[(size_t)(i * (i + 1) / 2 + i) * L];
    for (int i = N - 1; i >= 0; i--) {
        double s = R[(size_t)i * L];
        for (int t = i + 1; t < N; t++) s -= B[(size_t)(t * (t + 1) / 2 + i) * L] * R[(size_t)t * L];
        R[(size_t)i * L] = s;
    }
    for (int i = 0; i < N; i++) a.y[k + (size_t)e2[i] * ld] = R[(size_t)i * L];
}

}  // namespace mimsem
