// Instantiations and launchers of the remaining register kernels: M2, K (fallback), rotational / upwinded operators,
// 0-form mass matrix.
#include "kernels.cuh"
#include "launch.hpp"
#include "m2_solve.cuh"

namespace mimsem {

void launch_m2(int p, bool with_h, const KArgs& a, unsigned grid, cudaStream_t st) {
    for_p(p, [&](auto Pc) {
        constexpr int P = decltype(Pc)::value;
        if (with_h) k_apply_m2<P, true><<<grid, 128, 0, st>>>(a);
        else k_apply_m2<P, false><<<grid, 128, 0, st>>>(a);
    });
}

int launch_solve_m2(int p, bool with_h, const KArgs& a, cudaStream_t st, std::string* err) {
    int rc = 0;
    for_p(p, [&](auto Pc) {
        constexpr int P = decltype(Pc)::value;
        using S = M2SolveSmem<P>;
        const size_t smem = (size_t)S::DOUBLES * sizeof(double);
        void (*kern)(const KArgs) = with_h ? k_solve_m2<P, true> : k_solve_m2<P, false>;
        cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) {
            *err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(ce);
            rc = -1;
            return;
        }
        const unsigned grid = (unsigned)(((int64_t)a.nel * a.nlev + S::LANES - 1) / S::LANES);
        kern<<<grid, S::LANES, smem, st>>>(a);
    });
    return rc;
}

void launch_inc_tile(int p, bool div, const IncArgs& a, cudaStream_t st) {
    const unsigned grid = (unsigned)((a.nel + 3) / 4);
    for_p(p, [&](auto Pc) {
        constexpr int P = decltype(Pc)::value;
        if (div) k_inc_tile<P, true><<<grid, 128, 0, st>>>(a);
        else k_inc_tile<P, false><<<grid, 128, 0, st>>>(a);
    });
}

void launch_k_regs(int p, const KArgs& a, unsigned grid, cudaStream_t st) {
    for_p(p, [&](auto Pc) { k_apply_k<decltype(Pc)::value><<<grid, 128, 0, st>>>(a); });
}

void launch_m0h_up(int p, const NodeArgs& a, unsigned grid, cudaStream_t st) {
    for_p(p, [&](auto Pc) { k_apply_m0h_up<decltype(Pc)::value><<<grid, 128, 0, st>>>(a); });
}

void launch_m0(int p, bool with_h, const NodeArgs& a, unsigned grid, cudaStream_t st) {
    for_p(p, [&](auto Pc) {
        constexpr int P = decltype(Pc)::value;
        if (with_h) k_apply_m0<P, true><<<grid, 128, 0, st>>>(a);
        else k_apply_m0<P, false><<<grid, 128, 0, st>>>(a);
    });
}

}  // namespace mimsem
