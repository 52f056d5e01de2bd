// apm_common.cuh -- small device/host helpers shared by the kernels of libapm_b200.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace apm {

constexpr int kThreads = 256;   // threads per CTA of the count kernels
constexpr int kMaxWords = 8;    // bit-parallel kernel handles m <= 32 * kMaxWords
constexpr int kMaxMyersLen = 32 * kMaxWords;

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// Words per (group, code) entry of the Peq table: R patterns x NW words, padded so that the entry
// can be fetched with the widest shared-memory vector loads.
__host__ __device__ constexpr int entry_words(int rnw) {
    return rnw >= 4 ? ((rnw + 3) / 4) * 4 : (rnw >= 2 ? 2 : 1);
}

// Optional output of the match POSITIONS (apm_find_matches / apm_plan_set_hit_buffer): every kernel that counts a
// matching window also appends (pattern index << 40 | global window start) when a sink is attached.
struct HitSink {
    unsigned long long *buf = nullptr;    // device array of `cap` entries, or nullptr (positions not wanted)
    unsigned long long cap = 0;
    unsigned long long *count = nullptr;  // device counter: total hits, may exceed cap
    long long base = 0;                   // global index of local window start 0
};
constexpr int kHitPosBits = 40;

#ifdef __CUDACC__
__device__ __forceinline__ void hit_emit(const HitSink &h, int pattern, long long local_start) {
    const unsigned long long pos = atomicAdd(h.count, 1ull);
    if (pos < h.cap) h.buf[pos] = ((unsigned long long)pattern << kHitPosBits) | (unsigned long long)(h.base + local_start);
}
// all set bits of a 32-window hit mask whose bit 0 is local window start j0
__device__ __forceinline__ void hit_emit_mask(const HitSink &h, int pattern, long long j0, uint32_t mask) {
    while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1u;
        hit_emit(h, pattern, j0 + b);
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + TMA 1-D bulk copy (cp.async.bulk -> SASS UBLKCP) --------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                             uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(__cvta_generic_to_global(gmem_src)), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
#endif  // __CUDACC__

}  // namespace apm
