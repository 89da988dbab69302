// z = blockdiag(M1)^-1 r, one block per element: the element-block Jacobi preconditioner the reference puts on its
// 1-form mass-matrix solves (KSPGMRES + PCBJACOBI with PCBJacobiSetTotalBlocks(pc, size * nElsX^2, NULL),
// eul/HorizSolve.cpp:77-84, 791-796).  PETSc cuts the assembled matrix into equal consecutive row blocks; with the
// reference's element-blocked edge numbering (scr/Proc2.py:105-123) block e holds exactly the 2 p^2 edges element e
// owns (its west / interior x-edges and south / interior y-edges) -- the same rows as the engine's edge block of e.
//
// The block is never assembled in global memory.  One thread owns one (element, level) pair: it tabulates the lower
// triangle in shared memory (entry-major, lanes = levels) from closed forms that follow from the tensor structure of
// the edge bases when the quadrature order equals the element order,
//   B_xx[(ix,iy),(jx,jy)] = delta(ix,jx) sum_qy E[qy][iy] E[qy][jy] c Gaa(ix,qy)      (+ the west neighbour's far line on ix = 0)
//   B_yy[(ix,iy),(jx,jy)] = delta(iy,jy) sum_qx E[qx][ix] E[qx][jx] c Gbb(qx,iy)      (+ the south neighbour's far line on iy = 0)
//   B_yx[(ix,iy),(jx,jy)] = E[jx][ix] E[iy][jy] c Gab(jx,iy)                          (x-edge column jx, y-edge row iy)
// with c = s t^tpow (the geometry is pre-multiplied by w/det), factorises it in place (L D L^T) and solves.
#pragma once
#include "kernels.cuh"

namespace mimsem {

template <int P>
struct BJacobiSmem {
    static constexpr int N = 2 * P * P;
    static constexpr int TRI = N * (N + 1) / 2;
    static constexpr int LANES = 32;
    static constexpr int DOUBLES = (TRI + N) * LANES;
};

template <int P>
__global__ void __launch_bounds__(32) k_bjacobi_m1(const __grid_constant__ KArgs a) {
    using D = ElDim<P>;
    using S = BJacobiSmem<P>;
    constexpr int N = S::N, NP1 = P + 1, L = S::LANES, PP = P * P;
    extern __shared__ double smb[];
    double* B = smb + threadIdx.x;                        // B[t * L], t = i (i+1)/2 + j, i >= j
    double* R = smb + (size_t)S::TRI * L + threadIdx.x;   // R[i * L]
    const unsigned idx = blockIdx.x * L + threadIdx.x;
    if (idx >= (unsigned)a.nel * (unsigned)a.nlev) return;
    int e = (int)fastdiv(idx, a.div_m, a.div_s);
    const int k = (int)(idx - (unsigned)e * (unsigned)a.nlev);
    if (a.elist) e = a.elist[e];
    const size_t ld = a.ld;
    const int* __restrict__ ex = a.el1x + (size_t)e * D::N1E;
    const int* __restrict__ ey = a.el1y + (size_t)e * D::N1E;
    const int* __restrict__ eq = a.elq + (size_t)e * D::Q2;
    const double* __restrict__ G = a.G + (size_t)e * D::Q2 * 3;
    auto Bij = [&](int i, int j) -> double& { return B[(size_t)(i * (i + 1) / 2 + j) * L]; };
    // block rows: x-edge (ix, iy) -> ix P + iy ; y-edge (ix, iy) -> P^2 + iy P + ix   (the engine's edge block order)
    for (int t = 0; t < S::TRI; t++) B[(size_t)t * L] = 0.0;
    // point factors c(q) G(q)
    double caa[NP1][NP1], cab[NP1][NP1], cbb[NP1][NP1];   // [qy][qx]
#pragma unroll
    for (int qy = 0; qy <= P; qy++)
#pragma unroll
        for (int qx = 0; qx <= P; qx++) {
            const int q = qy * NP1 + qx;
            const double c = thick_factor(a, eq[q], k);
            caa[qy][qx] = c * G[q * 3 + 0];
            cab[qy][qx] = c * G[q * 3 + 1];
            cbb[qy][qx] = c * G[q * 3 + 2];
        }
    // far lines of the west / south neighbour: only their own-edge self coupling lands inside the block
    double fw[NP1], fs[NP1];
#pragma unroll
    for (int q = 0; q <= P; q++) fw[q] = fs[q] = 0.0;
    bool revw = false, revs = false;
    {
        const int nw = a.nbr[2 * e + 0], ns = a.nbr[2 * e + 1];
        if (nw >= 0) {
            const int n = nw & 0x1fffffff, side = (nw >> 29) & 1;
            revw = (nw >> 30) & 1;
            const int* __restrict__ nq = a.elq + (size_t)n * D::Q2;
            const double* __restrict__ Gn = a.G + (size_t)n * D::Q2 * 3;
            for (int t = 0; t <= P; t++) {
                const int q = side == 0 ? t * NP1 + P : P * NP1 + t;
                fw[t] = thick_factor(a, nq[q], k) * (side == 0 ? Gn[q * 3 + 0] : Gn[q * 3 + 2]);
            }
        }
        if (ns >= 0) {
            const int n = ns & 0x1fffffff, side = (ns >> 29) & 1;
            revs = (ns >> 30) & 1;
            const int* __restrict__ nq = a.elq + (size_t)n * D::Q2;
            const double* __restrict__ Gn = a.G + (size_t)n * D::Q2 * 3;
            for (int t = 0; t <= P; t++) {
                const int q = side == 0 ? t * NP1 + P : P * NP1 + t;
                fs[t] = thick_factor(a, nq[q], k) * (side == 0 ? Gn[q * 3 + 0] : Gn[q * 3 + 2]);
            }
        }
    }
    // x-x: same GLL column ix
    for (int ix = 0; ix < P; ix++)
        for (int iy = 0; iy < P; iy++)
            for (int jy = 0; jy <= iy; jy++) {
                double s = 0.0;
#pragma unroll
                for (int qy = 0; qy <= P; qy++) s += a.E[qy * P + iy] * a.E[qy * P + jy] * caa[qy][ix];
                if (ix == 0) {
                    // the neighbour numbers the shared edges in its own direction: its edge j is my edge (rev ? P-1-j : j)
                    const int ni = revw ? P - 1 - iy : iy, nj = revw ? P - 1 - jy : jy;
#pragma unroll
                    for (int t = 0; t <= P; t++) s += a.E[t * P + ni] * a.E[t * P + nj] * fw[t];
                }
                Bij(ix * P + iy, ix * P + jy) = s;
            }
    // y-y: same GLL row iy
    for (int iy = 0; iy < P; iy++)
        for (int ix = 0; ix < P; ix++)
            for (int jx = 0; jx <= ix; jx++) {
                double s = 0.0;
#pragma unroll
                for (int qx = 0; qx <= P; qx++) s += a.E[qx * P + ix] * a.E[qx * P + jx] * cbb[iy][qx];
                if (iy == 0) {
                    const int ni = revs ? P - 1 - ix : ix, nj = revs ? P - 1 - jx : jx;
#pragma unroll
                    for (int t = 0; t <= P; t++) s += a.E[t * P + ni] * a.E[t * P + nj] * fs[t];
                }
                Bij(PP + iy * P + ix, PP + iy * P + jx) = s;
            }
    // y-x (rows of the y block, columns of the x block): one quadrature point each
    for (int iy = 0; iy < P; iy++)
        for (int ix = 0; ix < P; ix++)
            for (int jx = 0; jx < P; jx++)
                for (int jy = 0; jy < P; jy++) Bij(PP + iy * P + ix, jx * P + jy) = a.E[jx * P + ix] * a.E[iy * P + jy] * cab[iy][jx];
    // in-place B = L D L^T (unit lower L, 1/D on the diagonal)
    for (int i = 0; i < N; i++) {
        const int ri = i * (i + 1) / 2;
        for (int j = 0; j < i; j++) {
            const int rj = j * (j + 1) / 2;
            double s = B[(size_t)(ri + j) * L];
            for (int t = 0; t < j; t++) s -= B[(size_t)(ri + t) * L] * B[(size_t)(rj + t) * L];
            B[(size_t)(ri + j) * L] = s;
        }
        double d = B[(size_t)(ri + i) * L];
        for (int t = 0; t < i; t++) {
            const double w = B[(size_t)(ri + t) * L];
            const double l = w * B[(size_t)(t * (t + 1) / 2 + t) * L];
            d -= w * l;
            B[(size_t)(ri + t) * L] = l;
        }
        B[(size_t)(ri + i) * L] = 1.0 / d;
    }
    // rows of the block in the field: x-edge (ix,iy) = ex[iy (P+1) + ix], y-edge (ix,iy) = ey[iy P + ix]
    auto row_of = [&](int i) { return i < PP ? ex[(i % P) * NP1 + i / P] : ey[i - PP]; };
    for (int i = 0; i < N; i++) R[(size_t)i * L] = ldro(a.x + k + (size_t)row_of(i) * ld);
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
    for (int i = 0; i < N; i++) a.y[k + (size_t)row_of(i) * ld] = R[(size_t)i * L];
}

}  // namespace mimsem
