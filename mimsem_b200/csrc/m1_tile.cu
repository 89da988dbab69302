// Instantiations and launcher of the M1 / M1(h) tile kernel.
#include "launch.hpp"
#include "m1_tile.cuh"

namespace mimsem {

namespace {

template <int P, bool WITH_H, int HALO, int MINB>
void (*pick_nl(int nlev, int tpow))(const TArgs) {
    // compile-time level counts of the BASELINE configurations (C3: 30, C4: 40, C5: 60) together with the number of
    // thickness factors their Umat (1) / Uhmat (2) applies use; anything else: runtime
    constexpr int TP = WITH_H ? 2 : 1;
    if ((P == 3 || P == 4) && nlev == 60) return tpow == TP ? k_apply_m1_tile<P, WITH_H, 60, HALO, MINB, TP> : k_apply_m1_tile<P, WITH_H, 60, HALO, MINB, -1>;
    if (P == 3 && nlev == 30) return tpow == TP ? k_apply_m1_tile<P, WITH_H, 30, HALO, MINB, TP> : k_apply_m1_tile<P, WITH_H, 30, HALO, MINB, -1>;
    if (P == 3 && nlev == 40) return tpow == TP ? k_apply_m1_tile<P, WITH_H, 40, HALO, MINB, TP> : k_apply_m1_tile<P, WITH_H, 40, HALO, MINB, -1>;
    return k_apply_m1_tile<P, WITH_H, 0, HALO, MINB, -1>;
}

// register budget: CTAs per SM the shared-memory footprint allows at the BASELINE shapes
template <int P, bool WITH_H>
// (p = 4: 126 registers without spills at 4 tiles per SM; 5 tiles at 96 registers spill a few values and measure the same or
// worse -- profiles/r02_summary.md)
constexpr int default_minb() { return P <= 3 ? (WITH_H ? 5 : 6) : (P == 4 ? 4 : 3); }

}  // namespace

int launch_m1_tile(const M1TileLaunch& l, TArgs& t, cudaStream_t st, std::string* err) {
    int rc = 1;
    for_p(l.p, [&](auto Pc) {
        constexpr int P = decltype(Pc)::value;
        using S = M1Slots<P>;
        t.geo_doubles = S::GEO;
        const size_t smem = 16 + ((size_t)S::GEO + (size_t)(l.with_h ? S::NS_H : S::NS) * t.nlev) * sizeof(double);
        if (smem > 227 * 1024) return;
        void (*kern)(const TArgs) = nullptr;
        if (l.halo == 2) kern = pick_nl<P, false, 2, default_minb<P, false>()>(t.nlev, t.tpow);
        else if (l.halo) kern = pick_nl<P, false, 1, default_minb<P, false>()>(t.nlev, t.tpow);
        else if (l.with_h) {
            kern = pick_nl<P, true, 0, default_minb<P, true>()>(t.nlev, t.tpow);
            if constexpr (P == 4) {
                if (l.min_blocks == 5) kern = pick_nl<P, true, 0, 5>(t.nlev, t.tpow);
            }
        }
        else {
            kern = pick_nl<P, false, 0, default_minb<P, false>()>(t.nlev, t.tpow);
            if constexpr (P == 4) {   // register-budget variants kept for tuning runs (mimsem_gpu_set_option "m1_min_blocks")
                if (l.min_blocks == 5) kern = pick_nl<P, false, 0, 5>(t.nlev, t.tpow);
                if (l.min_blocks == 6) kern = pick_nl<P, false, 0, 6>(t.nlev, t.tpow);
            }
        }
        cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) {
            *err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(ce);
            rc = -1;
            return;
        }
        rc = 0;
        if (l.push_only) {
            if (l.push_ctas == 0) return;
            t.ntiles = 0;
            kern<<<l.push_ctas, 128, smem, st>>>(t);
            return;
        }
        if (l.nel == 0) return;
        t.ntiles = l.nel;
        if (t.pdl) {
            ce = launch_maybe_pdl(kern, dim3(l.nel + l.push_ctas), dim3(128), smem, st, true, t);
            if (ce != cudaSuccess) {
                *err = std::string("cudaLaunchKernelEx: ") + cudaGetErrorString(ce);
                rc = -1;
            }
            return;
        }
        kern<<<l.nel + l.push_ctas, 128, smem, st>>>(t);
    });
    return rc;
}

}  // namespace mimsem
