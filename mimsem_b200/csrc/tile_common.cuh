// Shared pieces of the TMA-staged tile kernels (sm_100a): PTX wrappers, the copy-list walker, the L2 prefetcher.
// See the comment above M1Slots in engine.cuh.
#pragma once
#include <cstdint>

#include "engine.cuh"
#include "p2p_sync.cuh"

namespace mimsem {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared 1-D bulk copy (TMA), completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// ask the TMA unit to pull a range into L2 (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
// programmatic dependent launch: the next kernel on the stream (launched with programmatic stream serialization) may
// begin once every CTA of this grid has passed this point or exited
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// ... and the dependent side: blocks until the previous kernel on the stream has completed and its memory is visible (a no-op
// in a launch without the attribute)
__device__ __forceinline__ void pdl_wait_primary() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// launch with (pdl) or without the programmatic-stream-serialization attribute
template <class... KArgsT, class... ArgsT>
inline cudaError_t launch_maybe_pdl(void (*kern)(KArgsT...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, ArgsT... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifdef MIMSEM_DIAG
__device__ __forceinline__ long long gtime_ns() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define DBG_T(i) do { if (a.dbg_times && threadIdx.x == 64) a.dbg_times[(size_t)tile_i * 6 + (i)] = gtime_ns(); } while (0)
#else
#define DBG_T(i) do { } while (0)
#endif

__device__ __forceinline__ const TileHdr* tile_record(const TArgs& a, int e) { return a.recs + (size_t)e * a.rec_stride; }

// Stage one tile: one warp walks the element's copy list (one entry per lane and round).  `inbox` = the halo inbox copy
// of this epoch (kind-4 entries), or nullptr.
// `h` = the record's header word and `first` = this lane's first copy entry (tile_first_entry), already in registers.
__device__ __forceinline__ CopyEnt tile_first_entry(const TArgs& a, const TileHdr* rec) {
    const int lane = (int)threadIdx.x & 31;
    const int nents = a.rec_stride - a.rec_hdr;
    return reinterpret_cast<const CopyEnt*>(rec + a.rec_hdr)[lane < nents ? lane : 0];
}
__device__ __forceinline__ void tile_load_pre(const TArgs& a, const TileHdr* rec, const TileHdr h, const CopyEnt first, const double* inbox,
                                              uint64_t* bar, double* geo, double* tile) {
    const CopyEnt* ents = reinterpret_cast<const CopyEnt*>(rec + a.rec_hdr);
    const int lane = (int)threadIdx.x & 31;
    const unsigned slot_bytes = (unsigned)a.nlev * 8u;
    const bool with_t = a.tpow > 0;
    for (int ci = lane; ci < h.cp_count; ci += 32) {
        const CopyEnt c = (ci == lane) ? first : ents[ci];
        if (c.kind == 2 && !with_t) continue;
        if (c.kind == 4 && a.halo.ll) continue;   // in-band protocol: unpacked by the tile's threads, not by TMA
        // one copy site for every kind (the walker is cold code on one warp; keep its footprint small): a run of
        // `count` rows is ONE bulk copy when the rows are packed (row stride == nlev), else one copy per row
        const double* src;
        double* dst = tile + (size_t)c.slot * a.nlev;
        size_t sstride = (size_t)a.nlev;
        bool packed = true;
        if (c.kind == 3) {
            src = a.geo + (size_t)c.src * a.geo_doubles;
            dst = geo;
        } else if (c.kind == 4) {
            src = inbox + (size_t)c.src * a.nlev;   // ghost rows of x, straight from the inbox (rows packed with stride nlev)
        } else if (c.kind == 2) {
            src = a.tinv + (size_t)c.src * a.nkT + a.lev0;
            sstride = (size_t)a.nkT;
            packed = a.contig_t;
        } else {
            src = (c.kind == 0 ? a.x : a.c) + (size_t)c.src * a.ld;
            sstride = (size_t)a.ld;
            packed = a.contig_x;
        }
        const int n = packed ? 1 : c.count;
        const unsigned bytes = c.kind == 3 ? (unsigned)a.geo_doubles * 8u : (packed ? slot_bytes * (unsigned)c.count : slot_bytes);
#pragma unroll 1
        for (int j = 0; j < n; j++) bulk_g2s(dst + (size_t)j * a.nlev, src + (size_t)j * sstride, bytes, bar);
    }
    // bytes the TMA unit will deliver (the LL protocol unpacks ghost slots with ordinary loads instead)
    const unsigned nslots = (unsigned)(h.nslots & 0xfff) + (a.halo.ll ? 0u : (unsigned)((h.nslots >> 12) & 0xff)) +
                            (with_t ? (unsigned)(h.nslots >> 20) : 0u);
    if (lane == 0) mbar_arrive_expect_tx(bar, nslots * slot_bytes + (unsigned)a.geo_doubles * 8u);
}
__device__ __forceinline__ void tile_load(const TArgs& a, int e, const double* inbox, uint64_t* bar, double* geo, double* tile) {
    const TileHdr* rec = tile_record(a, e);
    // header and this lane's first entry are fetched together (no dependent load on the critical path)
    const TileHdr h = rec[0];
    const CopyEnt first = tile_first_entry(a, rec);
    tile_load_pre(a, rec, h, first, inbox, bar, geo, tile);
}

// L2 prefetch of the DRAM-unique part of a LATER tile (its record, its own edge block, its coefficient and thickness
// rows): by the time that tile's CTA starts, its bulk loads hit L2, which takes the HBM latency out of the per-tile
// critical path.
__device__ __forceinline__ void tile_prefetch(const TArgs& a, int e) {
    const TileHdr* rec = tile_record(a, e);
    const CopyEnt* ents = reinterpret_cast<const CopyEnt*>(rec + a.rec_hdr);
    const TileHdr h = rec[0];
    const int lane = threadIdx.x & 31;
    const unsigned slot_bytes = (unsigned)a.nlev * 8u;
    const int own_slots = a.prefetch_own_slots;
    for (int ci = lane; ci < h.cp_count; ci += 32) {
        const CopyEnt c = ents[ci];
        const double* src = nullptr;
        unsigned bytes = slot_bytes * (unsigned)c.count;
        if (c.kind == 3) {
            src = a.geo + (size_t)c.src * a.geo_doubles;
            bytes = (unsigned)a.geo_doubles * 8u;
        } else if (c.kind == 2) {
            if (a.tpow > 0 && a.contig_t) src = a.tinv + (size_t)c.src * a.nkT + a.lev0;
        } else if (a.contig_x && ((c.kind == 0 && c.slot < own_slots) || c.kind == 1)) {
            src = (c.kind == 0 ? a.x : a.c) + (size_t)c.src * a.ld;
        }
        if (src) bulk_prefetch_l2(src, bytes);
    }
}

}  // namespace mimsem
