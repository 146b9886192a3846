// sm_100a primitives used by the hand-written network tower (hz_tower.cu): mbarriers, the bulk
// copy engine (TMA, cp.async.bulk), tensor memory (TMEM) management, tcgen05.mma / commit / ld and
// the shared-memory / instruction descriptors.  Inline PTX only; the bit layouts follow the PTX ISA
// (the same fields CUTLASS's cute/arch/mma_sm100_desc.hpp names).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace hz {
namespace sm100 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier --------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must end as a trapped launch (cudaErrorLaunchFailure on the host),
// never as a hung GPU.  `fault` (may be null) receives a code identifying the waiter.
#ifndef HZ_MBAR_SPIN_LIMIT
#define HZ_MBAR_SPIN_LIMIT (1u << 24)
#endif
// (the time-out path is one out-of-line function: inlined at every wait it made the tower's issue loop several times
// larger than the instruction cache)
__device__ __noinline__ void mbar_timeout(unsigned int* fault, unsigned int code) {
    if (fault) {
        atomicExch(fault, code);
        __threadfence_system();
    }
    __trap();
}
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, unsigned int* fault, unsigned int code) {
    for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it) {
        if (it > HZ_MBAR_SPIN_LIMIT) mbar_timeout(fault, code);
    }
}
// inline: one try; the spin loop is shared code (a waiter that has to spin is not in a hurry)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, unsigned int* fault, unsigned int code) {
    if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity, fault, code);
}

// ---- proxies / fences ------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- bulk copy engine (TMA, non-tensor form): global -> shared, completion on an mbarrier -----
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// multicast form: the bytes land at the same CTA-relative offset in every CTA of `mask`, and each
// destination CTA's mbarrier (same offset) receives the complete_tx
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            dst_smem),
        "l"(src), "r"(bytes), "r"(bar), "h"(mask)
        : "memory");
}
// shared -> global (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---- tensor memory -----------------------------------------------------------------------------
// one full warp; writes the TMEM base address (lane 0, column c) to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- descriptors -------------------------------------------------------------------------------
// K-major operand tile in SWIZZLE_128B layout: rows of 128 bytes (64 bf16 of K), 8-row atoms of
// 1024 bytes (16-byte chunk index XOR row%8), atoms stacked along M/N at `sbo` bytes.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr, uint32_t sbo = 1024) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);        // start address        bits [0,14)
    d |= (uint64_t)1 << 16;                          // leading byte offset  bits [16,30) (unused for swizzled K-major)
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;     // stride byte offset   bits [32,46)
    d |= (uint64_t)1 << 46;                          // descriptor version 1 (sm_100)
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}
// kind::f16, A = B = bf16 (K-major both), D = f32, M x N
__host__ __device__ constexpr uint32_t idesc_bf16_f32(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- MMA / commit / TMEM load (one thread issues) ------------------------------------------------
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// The same MMA with a collector hint for the A operand: consecutive MMAs that share A (here: one
// weight tile applied to several board rows) read it from shared memory once.  FILL = read A and keep
// it in the collector (SASS .A_KEEP), USE = take it from the collector and keep it (.A_REUSE.A_KEEP),
// LASTUSE = take it from the collector and release it (.A_REUSE).
enum { COLL_DISCARD = 0, COLL_FILL = 1, COLL_USE = 2, COLL_LASTUSE = 3 };
template <int COLL>
__device__ __forceinline__ void umma_bf16_coll(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if (COLL == COLL_FILL)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b),
                     "r"(idesc), "r"(accumulate) : "memory");
    else if (COLL == COLL_USE)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.collector::a::use [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b),
                     "r"(idesc), "r"(accumulate) : "memory");
    else if (COLL == COLL_LASTUSE)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b),
                     "r"(idesc), "r"(accumulate) : "memory");
    else
        umma_bf16(tmem_d, desc_a, desc_b, idesc, accumulate);
}
// arrives (count 1) on the mbarrier once every tcgen05 operation issued so far by this thread is done
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp reads lane (taddr.lane + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- cluster helpers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

}  // namespace sm100
}  // namespace hz
