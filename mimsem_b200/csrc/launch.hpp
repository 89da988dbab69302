// Kernel launchers, one translation unit per kernel family so that the sm_100a build runs in parallel
// (`make -j`): every template instantiation of a family lives in the .cu file that defines its launcher.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "engine.cuh"
#include "kernels.cuh"

namespace mimsem {

// Options of the M1 tile launch that are not kernel arguments.
struct M1TileLaunch {
    int p;
    bool with_h;
    int halo;           // 0: none, 1: fused ghost refresh with flags, 2: with in-band (LL) cells
    int nel;            // tiles
    int push_ctas;      // fused ghost refresh: CTAs of the push role (0: none)
    bool push_only;     // prologue of a pipelined sequence
    int min_blocks;     // register budget variant: 0 = default, else resident CTAs per SM to budget for (tuning)
};
// return: 0 launched, 1 not applicable (tile does not fit in shared memory: use a register kernel), < 0 CUDA error (*err set)
int launch_m1_tile(const M1TileLaunch& l, TArgs& t, cudaStream_t st, std::string* err);
// persistent double-buffered variant (m1_pipe.cuh); same return convention, single GPU launches only
int launch_m1_pipe(const M1TileLaunch& l, TArgs& t, cudaStream_t st, std::string* err);
int launch_k_tile(int p, TArgs& t, int nel, cudaStream_t st, std::string* err);
int launch_m2_tile(int p, bool with_h, TArgs& t, int nel, cudaStream_t st, std::string* err);

void launch_m1_regs(int p, bool with_h, const KArgs& a, unsigned grid, cudaStream_t st);
void launch_m1_lines(int p, bool with_h, bool far, const KArgs& a, dim3 grid, cudaStream_t st);
int launch_bjacobi_m1(int p, const KArgs& a, cudaStream_t st, std::string* err);
void launch_diag_m1(int p, bool invert, const KArgs& a, unsigned grid, cudaStream_t st);
void launch_m2(int p, bool with_h, const KArgs& a, unsigned grid, cudaStream_t st);
int launch_solve_m2(int p, bool with_h, const KArgs& a, cudaStream_t st, std::string* err);
void launch_k_regs(int p, const KArgs& a, unsigned grid, cudaStream_t st);
void launch_rot(int p, bool up, const KArgs& a, unsigned grid, cudaStream_t st);
void launch_m0h_up(int p, const NodeArgs& a, unsigned grid, cudaStream_t st);
void launch_m0(int p, bool with_h, const NodeArgs& a, unsigned grid, cudaStream_t st);
void launch_inc_tile(int p, bool div, const IncArgs& a, cudaStream_t st);

// p = 2..5 -> integral_constant dispatch (callers have validated p)
template <class F>
inline void for_p(int p, F f) {
    switch (p) {
        case 2: f(std::integral_constant<int, 2>()); break;
        case 3: f(std::integral_constant<int, 3>()); break;
        case 4: f(std::integral_constant<int, 4>()); break;
        case 5: f(std::integral_constant<int, 5>()); break;
        default: break;
    }
}

}  // namespace mimsem
