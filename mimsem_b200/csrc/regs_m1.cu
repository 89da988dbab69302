// Instantiations and launchers of the register (LDG) kernels of the 1-form mass matrix: thread per element-level,
// line tasks, diagonal.
#include "kernels.cuh"
#include "launch.hpp"
#include "m1_bjacobi.cuh"

namespace mimsem {

void launch_m1_regs(int p, bool with_h, const KArgs& a, unsigned grid, cudaStream_t st) {
    for_p(p, [&](auto Pc) {
        constexpr int P = decltype(Pc)::value;
        if (with_h) k_apply_m1<P, true><<<grid, 128, 0, st>>>(a);
        else k_apply_m1<P, false><<<grid, 128, 0, st>>>(a);
    });
}

void launch_m1_lines(int p, bool with_h, bool far, const KArgs& a, dim3 grid, cudaStream_t st) {
    for_p(p, [&](auto Pc) {
        constexpr int P = decltype(Pc)::value;
        if (far) {
            if (with_h) k_apply_m1_lines<P, true, true><<<grid, 128, 0, st>>>(a);
            else k_apply_m1_lines<P, false, true><<<grid, 128, 0, st>>>(a);
        } else {
            if (with_h) k_apply_m1_lines<P, true, false><<<grid, 128, 0, st>>>(a);
            else k_apply_m1_lines<P, false, false><<<grid, 128, 0, st>>>(a);
        }
    });
}

int launch_bjacobi_m1(int p, const KArgs& a, cudaStream_t st, std::string* err) {
    int rc = 0;
    for_p(p, [&](auto Pc) {
        constexpr int P = decltype(Pc)::value;
        using S = BJacobiSmem<P>;
        const size_t smem = (size_t)S::DOUBLES * sizeof(double);
        if (smem > 227 * 1024) {
            *err = "element-block Jacobi: the block does not fit in shared memory at this order";
            rc = -2;
            return;
        }
        cudaError_t ce = cudaFuncSetAttribute(k_bjacobi_m1<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) {
            *err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(ce);
            rc = -1;
            return;
        }
        const unsigned grid = (unsigned)(((int64_t)a.nel * a.nlev + S::LANES - 1) / S::LANES);
        k_bjacobi_m1<P><<<grid, S::LANES, smem, st>>>(a);
    });
    return rc;
}

void launch_diag_m1(int p, bool invert, const KArgs& a, unsigned grid, cudaStream_t st) {
    for_p(p, [&](auto Pc) {
        constexpr int P = decltype(Pc)::value;
        if (invert) k_diag_m1<P, true><<<grid, 128, 0, st>>>(a);
        else k_diag_m1<P, false><<<grid, 128, 0, st>>>(a);
    });
}

}  // namespace mimsem
