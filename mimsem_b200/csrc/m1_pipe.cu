// Instantiations and launcher of the persistent, warp-specialised M1 / M1(h) tile kernel.
#include <algorithm>

#include "launch.hpp"
#include "m1_pipe.cuh"

namespace mimsem {

namespace {

template <int P, bool WITH_H>
void (*pick_pipe(int nlev, int tpow))(const TArgs, const int) {
    constexpr int TP = WITH_H ? 2 : 1;
    if ((P == 3 || P == 4) && nlev == 60) return tpow == TP ? k_apply_m1_pipe<P, WITH_H, 60, TP> : k_apply_m1_pipe<P, WITH_H, 60, -1>;
    if (P == 3 && nlev == 30) return tpow == TP ? k_apply_m1_pipe<P, WITH_H, 30, TP> : k_apply_m1_pipe<P, WITH_H, 30, -1>;
    if (P == 3 && nlev == 40) return tpow == TP ? k_apply_m1_pipe<P, WITH_H, 40, TP> : k_apply_m1_pipe<P, WITH_H, 40, -1>;
    return k_apply_m1_pipe<P, WITH_H, 0, -1>;
}

constexpr size_t kSmemMax = 227 * 1024;

}  // namespace

int launch_m1_pipe(const M1TileLaunch& l, TArgs& t, cudaStream_t st, std::string* err) {
    int rc = 1;
    for_p(l.p, [&](auto Pc) {
        constexpr int P = decltype(Pc)::value;
        t.geo_doubles = M1Slots<P>::GEO;
        const int nb = l.with_h ? M1Pipe<P, true>::ring(t.nlev, kSmemMax) : M1Pipe<P, false>::ring(t.nlev, kSmemMax);
        if (nb < M1Pipe<P, false>::G + 1) return;   // no look-ahead left: the tile kernel does better
        const size_t smem = l.with_h ? M1Pipe<P, true>::smem_bytes(t.nlev, nb) : M1Pipe<P, false>::smem_bytes(t.nlev, nb);
        void (*kern)(const TArgs, const int) = l.with_h ? pick_pipe<P, true>(t.nlev, t.tpow) : pick_pipe<P, false>(t.nlev, t.tpow);
        cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int dev = 0, sms = 0;
        if (ce == cudaSuccess) ce = cudaGetDevice(&dev);
        if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (ce != cudaSuccess) {
            *err = std::string("launch_m1_pipe: ") + cudaGetErrorString(ce);
            rc = -1;
            return;
        }
        rc = 0;
        if (l.nel == 0) return;
        t.ntiles = l.nel;
        const int grid = std::min(l.nel, sms);
        ce = launch_maybe_pdl(kern, dim3(grid), dim3(M1Pipe<P, false>::threads(nb)), smem, st, t.pdl != 0, t, nb);
        if (ce != cudaSuccess) {
            *err = std::string("cudaLaunchKernelEx: ") + cudaGetErrorString(ce);
            rc = -1;
        }
    });
    return rc;
}

}  // namespace mimsem
