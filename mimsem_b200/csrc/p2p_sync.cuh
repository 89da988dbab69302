// System-scope flag primitives of the peer-to-peer ghost refresh (flags live in peer-mapped device memory).
#pragma once

namespace mimsem {

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// spin until *p >= want; gives up after ~4e9 cycles (a dead peer must not hang the GPU) and records the failure
static __device__ __noinline__ void spin_until(const unsigned long long* p, unsigned long long want, int* err) {
    const long long t0 = clock64();
    while (ld_acquire_sys(p) < want) {
        if (clock64() - t0 > 4000000000ll) {
            atomicExch(err, 1);
            break;
        }
        __nanosleep(64);
    }
}

}  // namespace mimsem
