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
#ifndef ZFB_EMULATE
#include <cuda.h>
#endif

namespace zfb {

#ifndef ZFB_EMULATE

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
// bytes to skip so that a shared-memory pointer becomes `align`-aligned (keeps the pointer's
// address space visible to the compiler: LDS/STS, not generic LD/ST)
__device__ __forceinline__ unsigned smem_align_pad(const void *p, unsigned align) {
    return (align - (smem_u32(p) & (align - 1))) & (align - 1);
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

// ---- tiled TMA over a 4-D tensor map (one request moves a whole [rows][bytes] box) ----
typedef CUtensorMap TensorMap;
#define ZFB_TMAP_PARAM const __grid_constant__ ::zfb::TensorMap
__device__ __forceinline__ void tma_load_4d(void *dst_smem, const TensorMap *map, int c0, int c1, int c2, int c3,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
            "r"(smem_u32(dst_smem)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
// the same with an L2 eviction-priority hint (createpolicy): pol != 0
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_4d_hint(void *dst_smem, const TensorMap *map, int c0, int c1, int c2,
                                                 int c3, uint64_t *bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;" ::
            "r"(smem_u32(dst_smem)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d_hint(const TensorMap *map, int c0, int c1, int c2, int c3,
                                                  const void *src_smem, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2, %3, %4}], [%5], %6;" ::"l"(map),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(src_smem)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const TensorMap *map, int c0, int c1, int c2, int c3,
                                             const void *src_smem) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(map),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(src_smem))
                 : "memory");
}

#else  // ---------------------------------------------------------------- emulation

inline unsigned smem_align_pad(const void *p, unsigned align) {
    return (unsigned)((align - ((uintptr_t)p & (align - 1))) & (align - 1));
}
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

// tensor map stand-in: the fields cuTensorMapEncodeTiled takes, interpreted by the two
// functions below exactly as the hardware does (out-of-bounds elements read as zero and are
// not written; SWIZZLE_128B / SWIZZLE_64B XOR the 16-byte chunk index with the row's bits)
struct TensorMap {
    unsigned char *base;
    unsigned long long dim[4], stride[4];      // elements, bytes (stride[0] = element size)
    unsigned int box[4];
    int swizzle;                               // bytes: 0, 64 or 128
};
#define ZFB_TMAP_PARAM const ::zfb::TensorMap
template <bool STORE>
inline void tma_emu_4d(unsigned char *smem, const TensorMap *m, int c0, int c1, int c2, int c3) {
    const size_t es = (size_t)m->stride[0];
    const size_t row_bytes = (size_t)m->box[0] * es;
    size_t row = 0;
    for (unsigned i3 = 0; i3 < m->box[3]; ++i3)
        for (unsigned i2 = 0; i2 < m->box[2]; ++i2)
            for (unsigned i1 = 0; i1 < m->box[1]; ++i1, ++row) {
                const long long g3 = (long long)c3 + i3, g2 = (long long)c2 + i2, g1 = (long long)c1 + i1;
                const bool row_ok = g3 >= 0 && g3 < (long long)m->dim[3] && g2 >= 0 && g2 < (long long)m->dim[2] &&
                                    g1 >= 0 && g1 < (long long)m->dim[1];
                for (unsigned i0 = 0; i0 < m->box[0]; ++i0) {
                    const long long g0 = (long long)c0 + i0;
                    const bool ok = row_ok && g0 >= 0 && g0 < (long long)m->dim[0];
                    size_t off = row * row_bytes + (size_t)i0 * es;          // dense box offset
                    if (m->swizzle) {
                        const size_t chunk = off >> 4, within = off & 15;
                        const size_t mask = m->swizzle == 128 ? ((off >> 7) & 7) : ((off >> 7) & 3);
                        off = ((chunk ^ mask) << 4) | within;
                    }
                    unsigned char *g = m->base + (size_t)g3 * m->stride[3] + (size_t)g2 * m->stride[2] +
                                       (size_t)g1 * m->stride[1] + (size_t)g0 * es;
                    if (STORE) { if (ok) memcpy(g, smem + off, es); }
                    else if (ok) memcpy(smem + off, g, es);
                    else memset(smem + off, 0, es);
                }
            }
}
inline void tma_load_4d(void *dst, const TensorMap *m, int c0, int c1, int c2, int c3, uint64_t *) {
    tma_emu_4d<false>((unsigned char *)dst, m, c0, c1, c2, c3);
}
inline void tma_store_4d(const TensorMap *m, int c0, int c1, int c2, int c3, const void *src) {
    tma_emu_4d<true>((unsigned char *)src, m, c0, c1, c2, c3);
}
inline uint64_t l2_policy_evict_last() { return 1; }
inline uint64_t l2_policy_evict_first() { return 2; }
inline void tma_load_4d_hint(void *dst, const TensorMap *m, int c0, int c1, int c2, int c3, uint64_t *b, uint64_t) {
    tma_load_4d(dst, m, c0, c1, c2, c3, b);
}
inline void tma_store_4d_hint(const TensorMap *m, int c0, int c1, int c2, int c3, const void *src, uint64_t) {
    tma_store_4d(m, c0, c1, c2, c3, src);
}

#endif

// host side: rank-4 map over 8-byte elements; dims in elements, strides in bytes (dim 0 is dense)
struct TensorMapSpec {
    void *base;
    unsigned long long dim[4];
    unsigned long long stride[4];          // stride[0] ignored (8)
    unsigned int box[4];
    int swizzle;                           // 0, 64, 128
};
// returns 0 on success (the sm_100a build resolves cuTensorMapEncodeTiled through the runtime)
int make_tensor_map(const TensorMapSpec &spec, TensorMap *out);

}  // namespace zfb
