// Instantiations and launchers of the remaining register kernels: M2, K (fallback), rotational / upwinded operators,
// 0-form mass matrix.
#include "kernels.cuh"
#include "launch.hpp"

namespace mimsem {

void launch_m2(int p, bool with_h, const KArgs& a, unsigned grid, cudaStream_t st) {
    for_p(p, [&](auto Pc) {
        constexpr int P = decltype(Pc)::value;
        if (with_h) k_apply_m2<P, true><<<grid, 128, 0, st>>>(a);
        else k_apply_m2<P, false><<<grid, 128, 0, st>>>(a);
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
