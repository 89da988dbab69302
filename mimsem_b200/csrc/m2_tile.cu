// Instantiations and launcher of the M2 / M2(rho) tile kernel.
#include "launch.hpp"
#include "m2_tile.cuh"

namespace mimsem {

int launch_m2_tile(int p, bool with_h, TArgs& t, int nel, cudaStream_t st, std::string* err) {
    int rc = 1;
    for_p(p, [&](auto Pc) {
        constexpr int P = decltype(Pc)::value;
        using S = M2Slots<P>;
        t.geo_doubles = S::GEO;
        const size_t smem = 16 + ((size_t)S::GEO + (size_t)(with_h ? S::NS_H : S::NS) * t.nlev) * sizeof(double);
        if (smem > 227 * 1024) return;
        void (*kern)(const TArgs) = nullptr;
        auto pick = [&](auto H) {
            constexpr bool WH = decltype(H)::value;
            if ((P == 3 || P == 4) && t.nlev == 60) kern = k_apply_m2_tile<P, WH, 60>;
            else if (P == 3 && t.nlev == 30) kern = k_apply_m2_tile<P, WH, 30>;
            else if (P == 3 && t.nlev == 40) kern = k_apply_m2_tile<P, WH, 40>;
            else kern = k_apply_m2_tile<P, WH, 0>;
        };
        if (with_h) pick(std::true_type());
        else pick(std::false_type());
        cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) {
            *err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(ce);
            rc = -1;
            return;
        }
        rc = 0;
        t.ntiles = nel;
        if (nel > 0 && launch_maybe_pdl(kern, dim3(nel), dim3(64), smem, st, t.pdl != 0, t) != cudaSuccess) {
            *err = std::string("cudaLaunchKernelEx: ") + cudaGetErrorString(cudaGetLastError());
            rc = -1;
        }
    });
    return rc;
}

}  // namespace mimsem
