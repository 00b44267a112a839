// Thin wrappers over the sm_100a bulk-copy engine (1-D TMA: cp.async.bulk) and
// mbarrier, as used by the streaming IIR kernels (zfb_iirstream.cuh): every
// lane moves its own contiguous piece of a stream between global and shared
// memory without touching the LSU/L1 wavefront path (a per-lane LDG.128 of 32
// different cache lines costs 32 L1 wavefronts; a 256-byte bulk copy costs none).
//
// Under ZFB_EMULATE (tests/emu: CPU stand-in, test infrastructure only) the
// copies are synchronous memcpy's and the barriers are no-ops, so the kernels'
// indexing and arithmetic run unchanged in the GPU-less build container.
#pragma once
#include <stdint.h>

#include "zfb_platform.h"

namespace zfb {

#ifndef ZFB_EMULATE

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "ZFB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra ZFB_DONE_%=;\n"
        "bra ZFB_WAIT_%=;\n"
        "ZFB_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, completion counted in bytes on `bar` (all of src, dst, bytes: multiples of 16)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-groups
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N of this thread's bulk groups may still be READING shared memory
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// at most N of this thread's bulk groups may still be incomplete (writes not yet performed)
template <int N>
__device__ __forceinline__ void bulk_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy writes to shared memory -> visible to the bulk-copy engine
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#else  // ---------------------------------------------------------------- emulation

inline void mbar_init(uint64_t *, int) {}
inline void mbar_fence_init() {}
inline void mbar_arrive(uint64_t *) {}
inline void mbar_arrive_expect_tx(uint64_t *, uint32_t) {}
inline void mbar_wait(uint64_t *, uint32_t) {}
inline void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *) { memcpy(dst, src, bytes); }
inline void bulk_s2g(void *dst, const void *src, uint32_t bytes) { memcpy(dst, src, bytes); }
inline void bulk_commit() {}
template <int N> inline void bulk_wait_read() {}
template <int N> inline void bulk_wait() {}
inline void fence_proxy_async() {}

#endif

}  // namespace zfb
