// CPU check of the compatibility layer's Krylov solver (petsc_compat.cpp: KSPSolve) on small dense shell operators:
// GMRES(30) with restarts on a nonsymmetric matrix, CG on a symmetric positive definite one, the four preconditioner
// routes (none, the shell's diagonal, a PCSHELL callback, the shell's own blocks for a PCBJACOBI request) and the
// fall-back of a PCBJACOBI request when the shell declines.  No GPU.
//   compat_ksp_check            prints "compat_ksp_check ok"
#include <cmath>
#include <cstdio>
#include <vector>

#include "petsc_compat.h"

#ifndef MIMSEM_HAVE_PETSC
namespace {

struct Dense {
    int n;
    std::vector<double> a;
    int block;        // block size of the block-diagonal preconditioner
    bool decline;     // the PCBJACOBI hook answers "not supported"
    int block_calls;
};

PetscErrorCode mult(Mat A, Vec x, Vec y) {
    Dense* d;
    MatShellGetContext(A, &d);
    PetscScalar *xa, *ya;
    VecGetArray(x, &xa);
    VecGetArray(y, &ya);
    for (int i = 0; i < d->n; i++) {
        double s = 0.0;
        for (int j = 0; j < d->n; j++) s += d->a[(size_t)i * d->n + j] * xa[j];
        ya[i] = s;
    }
    VecRestoreArray(x, &xa);
    VecRestoreArray(y, &ya);
    return 0;
}
PetscErrorCode getdiag(Mat A, Vec v) {
    Dense* d;
    MatShellGetContext(A, &d);
    PetscScalar* a;
    VecGetArray(v, &a);
    for (int i = 0; i < d->n; i++) a[i] = d->a[(size_t)i * d->n + i];
    VecRestoreArray(v, &a);
    return 0;
}
// z = blockdiag(A)^-1 r by Gaussian elimination with partial pivoting on every diagonal block
void block_solve(Dense* d, const double* r, double* z) {
    const int nb = d->block;
    for (int b0 = 0; b0 < d->n; b0 += nb) {
        const int m = std::min(nb, d->n - b0);
        std::vector<double> B((size_t)m * (m + 1));
        for (int i = 0; i < m; i++) {
            for (int j = 0; j < m; j++) B[(size_t)i * (m + 1) + j] = d->a[(size_t)(b0 + i) * d->n + b0 + j];
            B[(size_t)i * (m + 1) + m] = r[b0 + i];
        }
        for (int c = 0; c < m; c++) {
            int piv = c;
            for (int i = c + 1; i < m; i++)
                if (std::fabs(B[(size_t)i * (m + 1) + c]) > std::fabs(B[(size_t)piv * (m + 1) + c])) piv = i;
            for (int j = 0; j <= m; j++) std::swap(B[(size_t)c * (m + 1) + j], B[(size_t)piv * (m + 1) + j]);
            for (int i = c + 1; i < m; i++) {
                const double f = B[(size_t)i * (m + 1) + c] / B[(size_t)c * (m + 1) + c];
                for (int j = c; j <= m; j++) B[(size_t)i * (m + 1) + j] -= f * B[(size_t)c * (m + 1) + j];
            }
        }
        for (int i = m - 1; i >= 0; i--) {
            double s = B[(size_t)i * (m + 1) + m];
            for (int j = i + 1; j < m; j++) s -= B[(size_t)i * (m + 1) + j] * z[b0 + j];
            z[b0 + i] = s / B[(size_t)i * (m + 1) + i];
        }
    }
}
PetscErrorCode blocks(Mat A, Vec r, Vec z) {
    Dense* d;
    MatShellGetContext(A, &d);
    if (d->decline) return 56;
    d->block_calls++;
    PetscScalar *ra, *za;
    VecGetArray(r, &ra);
    VecGetArray(z, &za);
    block_solve(d, ra, za);
    VecRestoreArray(r, &ra);
    VecRestoreArray(z, &za);
    return 0;
}
PetscErrorCode pc_callback(PC pc, Vec r, Vec z) {
    Mat A;
    PCShellGetContext(pc, &A);
    return blocks(A, r, z);
}

struct Result { int its; double err; };

Result solve(Dense& d, bool cg, const char* pctype, bool callback, double rtol) {
    Mat A;
    MatCreateShell(MPI_COMM_WORLD, d.n, d.n, d.n, d.n, &d, &A);
    MatShellSetOperation(A, MATOP_MULT, (void (*)(void))mult);
    MatShellSetOperation(A, MATOP_GET_DIAGONAL, (void (*)(void))getdiag);
    MatShellSetOperation(A, MATOP_COMPAT_PCBJACOBI, (void (*)(void))blocks);
    Vec x, b, s;
    VecCreateMPI(MPI_COMM_WORLD, d.n, d.n, &x);
    VecDuplicate(x, &b);
    VecDuplicate(x, &s);
    PetscScalar* a;
    VecGetArray(x, &a);
    for (int i = 0; i < d.n; i++) a[i] = std::sin(0.7 * i) + 0.3 * std::cos(2.1 * i);
    VecRestoreArray(x, &a);
    MatMult(A, x, b);
    VecSet(s, 0.0);
    KSP ksp;
    PC pc;
    KSPCreate(MPI_COMM_WORLD, &ksp);
    KSPSetOperators(ksp, A, A);
    KSPSetTolerances(ksp, rtol, 1.0e-50, PETSC_DEFAULT, 2000);
    KSPSetType(ksp, cg ? KSPCG : KSPGMRES);
    KSPGetPC(ksp, &pc);
    PCSetType(pc, pctype);
    if (callback) {
        PCShellSetContext(pc, A);
        PCShellSetApply(pc, pc_callback);
    }
    KSPSolve(ksp, b, s);
    Result r;
    KSPGetIterationNumber(ksp, &r.its);
    double en, xn;
    VecAXPY(s, -1.0, x);
    VecNorm(s, NORM_2, &en);
    VecNorm(x, NORM_2, &xn);
    r.err = en / xn;
    KSPDestroy(&ksp);
    VecDestroy(&x); VecDestroy(&b); VecDestroy(&s);
    MatDestroy(&A);
    return r;
}

int failures = 0;
void expect(bool ok, const char* what, const Result& r) {
    std::printf("%-64s its %4d  err %.2e  %s\n", what, r.its, r.err, ok ? "ok" : "FAIL");
    if (!ok) failures++;
}

}  // namespace

int main() {
    PetscCompatSetRank(0, 1);
    const int n = 96;
    // symmetric positive definite: strong 8 x 8 diagonal blocks, weak coupling between them, diagonal spread over decades
    Dense S{n, std::vector<double>((size_t)n * n, 0.0), 8, false, 0};
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            const double scale = std::pow(10.0, 0.5 * ((i / 8) % 5)) * std::pow(10.0, 0.5 * ((j / 8) % 5));
            double v = (i / 8 == j / 8) ? 1.0 / (1.0 + std::abs(i - j)) : 0.02 / (1.0 + std::abs(i - j));
            if (i == j) v += 2.0;
            S.a[(size_t)i * n + j] = v * std::sqrt(scale);
        }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++) S.a[(size_t)i * n + j] = S.a[(size_t)j * n + i];
    // nonsymmetric: the same plus a skew part
    Dense N = S;
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            const double k = 0.3 * S.a[(size_t)i * n + j];
            N.a[(size_t)i * n + j] += k;
            N.a[(size_t)j * n + i] -= k;
        }
    const double rtol = 1.0e-12;
    Result none = solve(N, false, PCNONE, false, rtol);
    expect(none.err < 1e-9 && none.its > 30 && none.its <= 2000, "GMRES(30), no preconditioner, nonsymmetric (restarts)", none);
    Result jac = solve(N, false, PCJACOBI, false, rtol);
    expect(jac.err < 1e-9 && jac.its < none.its, "GMRES(30), diagonal of the shell", jac);
    Result bj = solve(N, false, PCBJACOBI, false, rtol);
    expect(bj.err < 1e-9 && bj.its < jac.its && N.block_calls > 0, "GMRES(30), PCBJACOBI -> the shell's own blocks", bj);
    Result sh = solve(N, false, PCSHELL, true, rtol);
    expect(sh.err < 1e-9 && sh.its == bj.its, "GMRES(30), PCSHELL callback with the same blocks", sh);
    N.decline = true;
    Result fb = solve(N, false, PCBJACOBI, false, rtol);
    expect(fb.err < 1e-9 && fb.its == jac.its, "GMRES(30), PCBJACOBI declined by the shell -> its diagonal", fb);
    Result cgj = solve(S, true, PCJACOBI, false, rtol);
    expect(cgj.err < 1e-9, "CG, diagonal of the shell, symmetric positive definite", cgj);
    Result cgb = solve(S, true, PCBJACOBI, false, rtol);
    expect(cgb.err < 1e-9 && cgb.its < cgj.its, "CG, PCBJACOBI -> the shell's own blocks", cgb);
    Result gs = solve(S, false, PCBJACOBI, false, rtol);
    expect(gs.err < 1e-9 && gs.its <= cgb.its + 2, "GMRES(30) on the symmetric matrix, same blocks", gs);
    // a block-diagonal matrix with its own blocks as the preconditioner: one iteration
    Dense D = S;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
            if (i / 8 != j / 8) D.a[(size_t)i * n + j] = 0.0;
    Result one = solve(D, false, PCBJACOBI, false, rtol);
    expect(one.err < 1e-12 && one.its == 1, "GMRES, exact inverse as the preconditioner", one);
    std::printf("compat_ksp_check %s\n", failures ? "FAIL" : "ok");
    return failures ? 1 : 0;
}
#else
int main() { std::printf("compat_ksp_check ok (real PETSc: nothing to check)\n"); return 0; }
#endif
