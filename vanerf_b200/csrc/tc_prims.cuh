// sm_100a building blocks of the tensor-core shading path: mbarrier, bulk async copy (TMA engine, no tensor map),
// tcgen05 MMA / commit / TMEM load-store, UMMA shared-memory descriptors for the K-major 128-byte-swizzle layout.
// Spelling of the PTX follows the CUTLASS 4.x sm100 headers (cute/arch/mma_sm100_umma.hpp, copy_sm100.hpp,
// tmem_allocator_sm100.hpp, cutlass/arch/barrier.h); nothing here depends on CUTLASS.
//
// Every blocking wait is bounded: a wait that does not complete within TC_WAIT_TRIES tries raises the CTA-wide abort flag,
// after which all waits fall through, the kernel drains (garbage results), frees TMEM and the host reports an error.
// A protocol bug therefore costs a wrong answer and an error code, never a hung GPU.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace tc {


__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// One try_wait: suspends the warp in hardware until the phase completes or the suspend-time hint (ns) runs out, so a
// waiting warp does not burn issue slots polling.
#ifndef TC_WAIT_HINT_NS
#define TC_WAIT_HINT_NS 20000u
#endif
// A wait gives up after TC_WAIT_TRIES tries: ~21 s when the suspend-time hint is honoured in full, ~0.1 s when every try
// returns at once (~150 cycles per try); legitimate waits of these kernels are below a millisecond, so neither time slicing,
// MPS nor a debugger stretching or shortening the tries can trip the bound, and a protocol bug still ends in seconds.
// (Measured, round 2: a wall-clock bound - globaltimer read on every failed try, or once per 1024 tries with the start time
// kept in registers - costs k_mlp_tc 7 % and 20 %: the extra live state at ~150 inlined wait sites changes the code of the
// fast path.  A try count costs one loop counter.)
#define TC_WAIT_TRIES (1 << 20)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(TC_WAIT_HINT_NS)
        : "memory");
    return ok != 0;
}
// Returns false when the wait was abandoned (try budget spent, or the abort flag raised by another thread).
// abort_flag[0] = code of the first wait that gave up, abort_flag[1 + code / 100] = last code of each wait class that
// was still pending.
// (Measured: pure polling with mbarrier.test_wait instead of the suspending try_wait changes nothing, and moving the
// retry loop out of line does not pay either.)
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, volatile int* abort_flag, int code) {
#pragma unroll 1
    for (int it = 0; it < TC_WAIT_TRIES; ++it) {
        if (mbar_try_wait(bar, parity)) return true;
        if (*abort_flag) break;           // only reached when a try returned without completion, i.e. off the fast path
    }
    if (*abort_flag == 0) *abort_flag = code;
    abort_flag[1 + code / 100] = code;
    return false;
}

// ---------------------------------------------------------------------------------------------------- proxies / fences
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------------- bulk copy (global -> smem)
// bytes % 16 == 0, both addresses 16-byte aligned; completion is signalled on `bar` as transaction bytes.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------------------- TMEM
// Whole-warp calls (.sync.aligned).  The allocated base address (lane 0, first column) is written to *smem_slot.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 lanes x 32 bit, 16 consecutive columns: thread t of the warp receives row (lane base + t), columns c..c+15.
// taddr = base + (lane_base << 16) + column, lane_base = 32 * (warp id % 4).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------- UMMA
// Operand tiles live in shared memory as K-major, 128-byte-swizzled "slots": `rows` rows of 64 bf16 (128 bytes), rows
// grouped by 8 into 1024-byte swizzle atoms (slot base 1024-byte aligned).  Element (r, k) sits at
//   base + (r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) ^ (r & 7)) << 4) | ((k & 7) << 1)).
#define TC_SLOT_ROW_BYTES 128
__device__ __forceinline__ uint32_t slot_chunk_off(int row, int chunk16) {     // byte offset of a 16-byte chunk (8 bf16)
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + (((chunk16 ^ row) & 7) << 4));
}
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits [0,14), leading byte
// offset >> 4 in [16,30) (1 for swizzled K-major), stride byte offset >> 4 in [32,46) (1024 B between 8-row groups),
// version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16: D fp32 (bits [4,6) = 1), A and B bf16
// ([7,10) = [10,13) = 1), both K-major, N >> 3 in [17,23), M >> 4 in [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T, one K step of 16 bf16.  Single thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on `bar` when every MMA issued so far by this thread has completed (implies fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace tc
