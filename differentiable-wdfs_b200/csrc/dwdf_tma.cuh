// dwdf_tma.cuh — thin inline-PTX wrappers for the sm_100a async-copy machinery the clipper kernels
// use: mbarrier, cp.async.bulk.tensor (TMA) 2-D tile loads/stores, proxy fences, bulk-group waits.
#pragma once
#if defined(__CUDACC_RTC__) // NVRTC: no host headers; the tensor map is an opaque 128-byte kernel parameter
struct alignas (64) CUtensorMap_st
{
    unsigned long long opaque[16];
};
typedef CUtensorMap_st CUtensorMap;
#else
#include <cuda.h>
#include <cstdint>
#endif

namespace dwdf
{

// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start before
// its predecessor in the stream has finished (its launch latency and prologue overlap the predecessor's tail); this waits until
// the predecessor has completed and its memory operations are visible. A no-op for a normally launched kernel.
__device__ __forceinline__ void grid_dependency_wait () { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32 (const void* p) { return (uint32_t) __cvta_generic_to_shared (p); }

__device__ __forceinline__ void mbar_init (uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// makes the mbarrier initialisation visible to the async proxy (TMA unit)
__device__ __forceinline__ void fence_mbar_init () { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// orders generic-proxy shared-memory accesses (ld/st.shared) against async-proxy ones (TMA)
__device__ __forceinline__ void fence_proxy_async () { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx (uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait (uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do
    {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (! done);
}

// global (tensor map, coords {c0 = time, c1 = sequence}) -> shared tile, completion on `bar`
__device__ __forceinline__ void tma_load_2d (uint32_t dst, const CUtensorMap* tm, int32_t c0, int32_t c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
// global tile -> L2 only (no shared-memory destination, no completion to wait for): run a few tiles ahead
// of the shared-memory ring so that the ring's own loads find their data in L2 instead of paying HBM latency
__device__ __forceinline__ void tma_prefetch_l2_2d (const CUtensorMap* tm, int32_t c0, int32_t c1)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tm), "r"(c0), "r"(c1) : "memory");
}
// shared tile -> global; out-of-bounds rows / columns are clipped by the TMA unit
__device__ __forceinline__ void tma_store_2d (const CUtensorMap* tm, int32_t c0, int32_t c1, uint32_t src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void tma_commit () { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still READ their shared-memory source
template <int N>
__device__ __forceinline__ void tma_wait_read ()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_wait_all ()
{
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc (const CUtensorMap* tm) { asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory"); }

__device__ __forceinline__ float4 lds128 (uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128 (uint32_t addr, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

} // namespace dwdf
