// Instantiations and launcher of the rotational-term kernels (RotMat, RotMat_up).
#include "kernels.cuh"
#include "launch.hpp"

namespace mimsem {

void launch_rot(int p, bool up, const KArgs& a, unsigned grid, cudaStream_t st) {
    for_p(p, [&](auto Pc) {
        constexpr int P = decltype(Pc)::value;
        if (up) k_apply_rot<P, true><<<grid, 128, 0, st>>>(a);
        else k_apply_rot<P, false><<<grid, 128, 0, st>>>(a);
    });
}

}  // namespace mimsem
