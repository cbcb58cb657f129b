// mbarrier / TMA (bulk async copy) PTX wrappers shared by the chain kernels.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace pmb {

// ---- mbarrier / TMA (PTX) -------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// Same wait with a suspend-time hint (ns): the loader / drainer threads of fk_rows_kernel wait for whole tiles;
// without the hint they were measured spinning (4e8 warp instructions per launch at 4M x 65) in the issue slots
// the row warps need.
__device__ __forceinline__ void mbar_wait_long(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity), "r"(200000u) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(tm), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}

// 1-D bulk copy global -> shared (contiguous range, multiple of 16 bytes), completion on an mbarrier
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void *gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(gsrc),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

// 1-D bulk copy shared -> global (TMA engine, no register traffic); completion tracked per thread by bulk groups
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void *gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
// same with an L2 eviction-priority hint for the written lines (experiment knob PMB_ST_HINT: the outputs are never
// read back by this kernel)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_store_hint(void *gdst, uint32_t ssrc, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(ssrc),
                 "r"(bytes), "l"(policy)
                 : "memory");
}
// Ask L2 to fetch a contiguous global range (multiple of 16 bytes) ahead of the TMA boxes that will read it: one
// sequential DRAM burst per tile instead of 128-byte row pieces.
__device__ __forceinline__ void bulk_prefetch_l2(const void *gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }


// 16-byte shared-memory reads through 32-bit shared addresses
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// read-only data (the joint table): free to be scheduled and merged by the compiler
__device__ __forceinline__ float4 lds128_ro(uint32_t addr) {
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}


constexpr int kChunk = 8;                         // joints per TMA box (128-byte rows, SWIZZLE_128B)
constexpr int kBoxBytes = 32 * kChunk * 16;       // one box: dense (swizzled) [32 frames][8 joints] float4
constexpr int kBoxStages = 2;                     // boxes in flight per warp (prefetch depth in chunks)

}  // namespace pmb
