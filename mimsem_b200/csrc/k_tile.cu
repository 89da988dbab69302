// Instantiations and launcher of the K (WtQUmat) tile kernel.
#include "k_tile.cuh"
#include "launch.hpp"

namespace mimsem {

int launch_k_tile(int p, TArgs& t, int nel, cudaStream_t st, std::string* err) {
    int rc = 1;
    for_p(p, [&](auto Pc) {
        constexpr int P = decltype(Pc)::value;
        t.geo_doubles = M1Slots<P>::GEO_K;
        const size_t smem = 16 + ((size_t)M1Slots<P>::GEO_K + (size_t)KSlots<P>::NS * t.nlev) * sizeof(double);
        if (smem > 227 * 1024) return;
        void (*kern)(const TArgs) = nullptr;
        if ((P == 3 || P == 4) && t.nlev == 60) kern = k_apply_k_tma<P, 60>;
        else if (P == 3 && t.nlev == 30) kern = k_apply_k_tma<P, 30>;
        else if (P == 3 && t.nlev == 40) kern = k_apply_k_tma<P, 40>;
        else kern = k_apply_k_tma<P, 0>;
        cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) {
            *err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(ce);
            rc = -1;
            return;
        }
        rc = 0;
        t.ntiles = nel;
        if (nel > 0 && launch_maybe_pdl(kern, dim3(nel), dim3(128), smem, st, t.pdl != 0, t) != cudaSuccess) {
            *err = std::string("cudaLaunchKernelEx: ") + cudaGetErrorString(cudaGetLastError());
            rc = -1;
        }
    });
    return rc;
}

}  // namespace mimsem
