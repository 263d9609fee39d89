// Device helpers shared by the streaming kernels (hmi_stream.cu, hmi_stream_nl.cu).
#pragma once

#include "hmi_common.cuh"

namespace hmi {

template <typename T> struct Chunk16;
template <> struct Chunk16<float> { typedef float4 type; static constexpr int N = 4; };
template <> struct Chunk16<double> { typedef double2 type; static constexpr int N = 2; };

__device__ __forceinline__ unsigned smem_u32(const void *p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cp_async16(unsigned dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(unsigned dst_smem, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <typename T>
__device__ __forceinline__ T shfl_up1(T v) { return __shfl_up_sync(0xffffffffu, v, 1); }
template <typename T>
__device__ __forceinline__ T shfl_down1(T v) { return __shfl_down_sync(0xffffffffu, v, 1); }

// Sum of the four neighbours of centre row b (rows above/below + lane halos).  The association
// (up + dn) + (l + r) is invariant under mirroring either axis (IEEE addition commutes), so a ghost
// cell loaded from its mirror cell evolves *bit for bit* like that cell through the fused sweeps and
// the result does not depend on how the sweeps were cut into launches or slabs.
template <typename T, int V>
__device__ __forceinline__ void sum4_row(const T (&up)[V], const T (&b)[V], const T (&dn)[V], T bl, T br,
                                         T (&out)[V])
{
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const T l = (v > 0) ? b[v - 1] : bl;
        const T r = (v < V - 1) ? b[v + 1] : br;
        out[v] = (up[v] + dn[v]) + (l + r);
    }
}

template <int PH> struct Phase { static constexpr int value = PH; };

__device__ __forceinline__ void lds16(unsigned addr, double (&v)[2])
{
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v[0]), "=d"(v[1]) : "r"(addr));
}
__device__ __forceinline__ void lds16(unsigned addr, float (&v)[4])
{
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
}
__device__ __forceinline__ unsigned lds32(unsigned addr)
{
    unsigned r;
    asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(r) : "r"(addr));
    return r;
}

__device__ __forceinline__ double lds1(unsigned addr, double)
{
    double r;
    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ float lds1(unsigned addr, float)
{
    float r;
    asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ unsigned lds16u(unsigned addr)
{
    unsigned short r;
    asm volatile("ld.shared.u16 %0, [%1];\n" : "=h"(r) : "r"(addr));
    return (unsigned)r;
}
__device__ __forceinline__ unsigned lds8(unsigned addr)
{
    unsigned r;
    asm volatile("ld.shared.u8 %0, [%1];\n" : "=r"(r) : "r"(addr));
    return r;
}

// ---- TMA (bulk asynchronous copy) + mbarrier helpers ------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(unsigned dst_smem, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// one box of a 2-D tensor map (innermost coordinate first) into shared memory, completing on `bar`
__device__ __forceinline__ void tma_load_2d(unsigned dst_smem, const void *tmap, int x, int y, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                 ::"r"(dst_smem), "l"(tmap), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

}  // namespace hmi
