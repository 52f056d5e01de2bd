// apm_filter.cuh -- exact filter mode: seed scan + verification, bit-identical counts.
//
// Pigeonhole: cut a pattern of length m into k+1 pieces.  If levenshtein(P, W) <= k for a window W of the same
// length, at most k pieces contain an edit, so some piece is copied verbatim into W, shifted by the number of
// insertions minus deletions before it: |shift| <= k.  A SEED is the first s symbols of a piece (s = the shortest
// piece of the plan's filtered patterns, <= 16), so
//
//   (p, j) can match  ==>  some piece i of p has its seed at text position j + o_i + d, |d| <= k, inside W
//
// with o_i = floor(i * m / (k + 1)).  The scan kernel computes a rolling hash of every s-byte text substring (one
// pass over the text, independent of the number of patterns) and probes a 64 KB one-word Bloom digest of all seed
// hashes in shared memory; the ~0.1 % of the positions that pass are queued per warp and, 32 at a time, re-hashed,
// checked against the full seed bitmap (L2), looked up through a directory, byte-compared and turned into
// (pattern, window, piece, shift) candidates (staged per warp, one global atomic per flush).  The verify kernels
// evaluate the banded DP |row - col| <= k of src/utils.c:76-99 for each candidate -- exact for the decision D <= k;
// one thread per candidate for the first 32 rows, a whole warp for the few that survive -- and count a matching
// window only from its CANONICAL witness (the lexicographically smallest (piece, shift) whose seed occurs), so a
// window witnessed several times is counted once.  No sort, no host round trip.  If the candidate buffer
// overflows (low-complexity text) a flag makes the verify kernels no-ops and switches on the band kernel for the
// same patterns and windows instead: always exact.
//
// Work: O(text bytes) + O(candidates * m * (2k+1)) instead of O(text bytes * patterns * m * (2k+1)).
#pragma once
#include "apm_common.cuh"
#include "apm_dp.cuh"

namespace apm {

constexpr int kFilterMinSeed = 8;
constexpr int kFilterMaxSeed = 16;
constexpr int kFilterMaxK = 16;  // == kBandDpMaxK (apm_dp.cuh)
constexpr int kFilterThreads = 512;
constexpr int kFilterPosPerThread = 32;
constexpr uint32_t kFilterHashB = 0x9E3779B1u;
constexpr int kFilterSlabLog = 28;  // window starts per scan/verify round: candidates carry a 28-bit local start

// table index of a seed hash (hb bits)
// (the leading bits of the polynomial hash depend on every byte of the seed; no extra mixing step)
__host__ __device__ __forceinline__ uint32_t filter_index(uint32_t h, int hb) { return h >> (32 - hb); }
// polynomial hash of s bytes: sum c_i * B^(s-1-i)
__host__ __device__ inline uint32_t filter_hash(const uint8_t *p, int s) {
    uint32_t h = 0;
    for (int i = 0; i < s; ++i) h = h * kFilterHashB + p[i];
    return h;
}
__host__ __device__ __forceinline__ int filter_piece_offset(int i, int m, int k) { return (int)((long long)i * m / (k + 1)); }

// candidate: slot (24) | piece (6) | shift + k (6) | local window start (28)
__host__ __device__ __forceinline__ uint64_t filter_pack(uint32_t slot, int piece, int dk, uint32_t jl) {
    return ((uint64_t)slot << 40) | ((uint64_t)piece << 34) | ((uint64_t)dk << 28) | jl;
}

#ifdef __CUDACC__

struct FilterArgs {
    const uint8_t *buf;           // device text; buf[0] is global byte buf_offset
    long long buf_len, n_end;     // valid bytes; local index of the global end of text
    long long w0, w1;             // local window-start range of this round (w1 - w0 <= 2^28)
    int s, k, hb, mmax;
    uint32_t bs;                  // B^s
    const uint32_t *bitmap;       // 2^hb bits
    const uint32_t *digest;       // 2^19 bits: OR of the bitmap bits with the same leading index bits
    const uint32_t *coarse;       // [2^16 + 1] first entry of every leading-16-bit class of the table index
    const uint32_t *ent_hash;     // [nent] full 32-bit hash of every seed (cheap reject before the byte compare)
    const uint32_t *ent_idx;      // [nent] table index of every seed, sorted
    const uint32_t *ent_slot;     // [nent] filtered-pattern slot
    const uint8_t *ent_piece;     // [nent] piece index
    int nent;
    const int *fp_id;             // [nfp] index into counts
    const int *fp_m;              // [nfp]
    const long long *fp_off;      // [nfp] offset of the pattern in pat_bytes
    const uint8_t *pat_bytes;
    uint64_t *cand;               // candidate buffer
    unsigned long long cap;
    unsigned long long *ncand;    // zeroed before the scan
    unsigned int *overflow;       // zeroed before the scan; set when cand is full
    uint64_t *longq;              // candidates still alive after kFilterShortRows rows: finished by whole warps
    unsigned long long longcap;
    unsigned long long *nlong;    // zeroed before the scan
    unsigned long long *counts;
    HitSink sink;                 // optional match-position output
};

__device__ __forceinline__ uint4 filter_load16(const FilterArgs &a, long long pos) {
    if (pos >= 0 && pos + 16 <= a.buf_len) return *reinterpret_cast<const uint4 *>(a.buf + pos);
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    for (int i = 0; i < 16; ++i) {
        const long long q = pos + i;
        if (q >= 0 && q < a.buf_len) w[i >> 2] |= (uint32_t)a.buf[q] << (8 * (i & 3));
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// slow path of the scan: text position t has hash h / table index idx and its bit is set in the seed bitmap.
// The entries are sorted by table index; a coarse directory over the leading 16 index bits gives the (tiny)
// range to look at in one load.
// Candidates are staged in a small shared-memory list per WARP and flushed with one global atomic per flush
// instead of one per candidate (a single contended counter admits ~1 atomic per clock, which bounded the scan).
constexpr int kFilterStage = 96;
constexpr int kFilterQueue = 64;
struct FilterStage {
    uint64_t item[kFilterStage];
    unsigned int count;   // may exceed kFilterStage: the excess went straight to global memory
    unsigned int qcount;  // text positions that passed the digest and wait for the second-level probe
    uint32_t queue[kFilterQueue];  // ... as offsets from the first scanned position of the round
};

__device__ __forceinline__ void filter_emit(const FilterArgs &a, FilterStage *stg, uint64_t packed) {
    const unsigned int slot = atomicAdd(&stg->count, 1u);
    if (slot < (unsigned)kFilterStage) {
        stg->item[slot] = packed;
    } else {  // staging list full (low-complexity text): append directly
        const unsigned long long pos = atomicAdd(a.ncand, 1ull);
        if (pos < a.cap) a.cand[pos] = packed;
        else *a.overflow = 1u;
    }
}

// all lanes of the warp (converged): move the warp's staged candidates to the global buffer
__device__ __forceinline__ void filter_flush(const FilterArgs &a, FilterStage *stg) {
    __syncwarp();
    const unsigned int n = min(stg->count, (unsigned)kFilterStage);
    if (n == 0) return;  // warp-uniform
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(a.ncand, (unsigned long long)n);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    for (unsigned int i = lane; i < n; i += 32) {
        if (base + i < a.cap) a.cand[base + i] = stg->item[i];
        else *a.overflow = 1u;
    }
    __syncwarp();
    if (lane == 0) stg->count = 0u;
    __syncwarp();
}

__device__ __noinline__ void filter_hit(const FilterArgs &a, FilterStage *stg, long long t, uint32_t idx, uint32_t h) {
    const uint32_t c = idx >> (a.hb - 16);
    const int lo = (int)__ldg(a.coarse + c), hi = (int)__ldg(a.coarse + c + 1);
    for (int e = lo; e < hi; ++e) {
        if (__ldg(a.ent_idx + e) != idx || __ldg(a.ent_hash + e) != h) continue;
        const uint32_t slot = __ldg(a.ent_slot + e);
        const int piece = __ldg(a.ent_piece + e);
        const int m = __ldg(a.fp_m + slot);
        const int o = filter_piece_offset(piece, m, a.k);
        const uint8_t *seed = a.pat_bytes + __ldg(a.fp_off + slot) + o;
        bool same = true;
        for (int x = 0; x < a.s && same; ++x) same = seed[x] == a.buf[t + x];
        if (!same) continue;
        for (int d = -a.k; d <= a.k; ++d) {
            const int off = o + d;  // where the seed sits inside the window
            if (off < 0 || off + a.s > m) continue;
            const long long j = t - off;
            if (j < a.w0 || j >= a.w1 || j + m > a.n_end) continue;  // full windows of this round only
            filter_emit(a, stg, filter_pack(slot, piece, d + a.k, (uint32_t)(j - a.w0)));
        }
    }
}

// Scan: every text position that can hold a seed of a window of [w0, w1).  A thread owns 32 consecutive
// positions: its own 32 bytes come from two 128-bit loads, the next 16 from the neighbouring lane by
// shuffle; rolling hash in registers.  Two-level probe: a 64 KB digest of the seed set lives in shared memory
// (copied once per persistent CTA), so all but ~0.1 % of the positions are rejected by one LDS and only the
// rest recompute their hash and touch the full bitmap in L2.
constexpr int kFilterSmemLog = 19;
constexpr int kFilterSmemBytes = 1 << (kFilterSmemLog - 3);

// slow path of the scan, second level: position t passed the shared-memory digest
template <int S>
__device__ __forceinline__ void filter_probe(const FilterArgs &a, FilterStage *stg, long long t) {
    uint32_t c[S];
#pragma unroll
    for (int i = 0; i < S; ++i) c[i] = a.buf[t + i];
    uint32_t h = 0;
#pragma unroll
    for (int i = 0; i < S; ++i) h = h * kFilterHashB + c[i];
    const uint32_t idx = filter_index(h, a.hb);
    if ((__ldg(a.bitmap + (idx >> 5)) >> (idx & 31)) & 1u) filter_hit(a, stg, t, idx, h);
}

// Shared-memory digest: 2^14 words; a seed with hash h sets TWO bits, (h >> 13) & 31 and (h >> 8) & 31, of word
// h >> 18 (a one-word Bloom filter: one LDS per text position, false-positive rate ~ (bits set per word / 32)^2).
__host__ __device__ __forceinline__ uint32_t filter_digest_word(uint32_t h) { return h >> 18; }
__host__ __device__ __forceinline__ uint32_t filter_digest_mask(uint32_t h) {
    return (1u << ((h >> 13) & 31)) | (1u << ((h >> 8) & 31));
}

template <int S>
__global__ void __launch_bounds__(kFilterThreads, 2) filter_scan_kernel(const __grid_constant__ FilterArgs a) {
    extern __shared__ __align__(16) uint32_t s_digest[];
    FilterStage *stg = reinterpret_cast<FilterStage *>(reinterpret_cast<unsigned char *>(s_digest) + kFilterSmemBytes) +
                       (threadIdx.x >> 5);  // this warp's staging list
    const long long t_begin = a.w0;
    const long long t_end = min(a.w1 - 1 + a.mmax, a.n_end) - S + 1;  // exclusive
    if (t_end <= t_begin) return;
    for (int i = threadIdx.x; i < kFilterSmemBytes / 16; i += kFilterThreads)
        reinterpret_cast<uint4 *>(s_digest)[i] = __ldg(reinterpret_cast<const uint4 *>(a.digest) + i);
    if ((threadIdx.x & 31) == 0) { stg->count = 0u; stg->qcount = 0u; }
    __syncthreads();
    // align the tile grid to 16-byte addresses of the buffer
    const long long mis = (long long)(reinterpret_cast<uintptr_t>(a.buf + t_begin) & 15);
    const long long base0 = t_begin - mis;
    constexpr long long kTile = (long long)kFilterThreads * kFilterPosPerThread;
    const int lane = threadIdx.x & 31;
    const unsigned char *dig = reinterpret_cast<const unsigned char *>(s_digest);
    const long long stride = (long long)gridDim.x * kTile;
    long long base = base0 + (long long)blockIdx.x * kTile;
    // software prefetch: the next tile's 32 bytes per thread are requested a tile ahead
    uint4 pre0 = make_uint4(0u, 0u, 0u, 0u), pre1 = pre0;
    if (base < t_end) {
        pre0 = filter_load16(a, base + (long long)threadIdx.x * kFilterPosPerThread);
        pre1 = filter_load16(a, base + (long long)threadIdx.x * kFilterPosPerThread + 16);
    }
    for (; base < t_end; base += stride) {
        const long long p0 = base + (long long)threadIdx.x * kFilterPosPerThread;
        const uint4 own0 = pre0, own1 = pre1;
        if (base + stride < t_end) {
            pre0 = filter_load16(a, p0 + stride);
            pre1 = filter_load16(a, p0 + stride + 16);
        }
        uint4 nxt;  // the 16 bytes behind this thread's 32: the neighbouring lane's first 16
        nxt.x = __shfl_down_sync(0xFFFFFFFFu, own0.x, 1);
        nxt.y = __shfl_down_sync(0xFFFFFFFFu, own0.y, 1);
        nxt.z = __shfl_down_sync(0xFFFFFFFFu, own0.z, 1);
        nxt.w = __shfl_down_sync(0xFFFFFFFFu, own0.w, 1);
        if (lane == 31) nxt = filter_load16(a, p0 + 32);
        const uint32_t w[12] = {own0.x, own0.y, own0.z, own0.w, own1.x, own1.y, own1.z, own1.w, nxt.x, nxt.y, nxt.z, nxt.w};
        auto byte_at = [&](int i) -> uint32_t { return __byte_perm(w[i >> 2], 0u, 0x4440 + (i & 3)); };
        uint32_t h = 0;
#pragma unroll
        for (int i = 0; i < S; ++i) h = h * kFilterHashB + byte_at(i);
        // digest probe of the 32 positions (filter_digest_word / _mask); the answers are shifted into `maybe` from
        // the top, so after 32 steps position i sits at bit i
        uint32_t maybe = 0u;
#pragma unroll
        for (int i = 0; i < kFilterPosPerThread; ++i) {
            const uint32_t word = *reinterpret_cast<const uint32_t *>(dig + ((h >> 16) & 0xFFFCu));
            maybe = __funnelshift_r(maybe, __funnelshift_r(word, 0u, h >> 13) & __funnelshift_r(word, 0u, h >> 8), 1);
            if (i + 1 < kFilterPosPerThread) h = h * kFilterHashB - byte_at(i) * a.bs + byte_at(i + S);  // roll by one byte
        }
        // positions outside [t_begin, t_end) do not count
        const long long lo = t_begin - p0, hi = t_end - p0;
        if (lo > 0) maybe &= lo >= 32 ? 0u : ~((1u << (int)lo) - 1u);
        if (hi < 32) maybe &= hi <= 0 ? 0u : ((1u << (int)hi) - 1u);
        // ~0.1 % of the positions pass the digest.  They are queued per warp and probed 32 at a time with all
        // lanes busy (probing inline would stall the whole warp behind one lane's dependent loads).
        while (maybe) {
            const int i = __ffs(maybe) - 1;
            maybe &= maybe - 1u;
            const unsigned int slot = atomicAdd(&stg->qcount, 1u);
            if (slot < (unsigned)kFilterQueue) stg->queue[slot] = (uint32_t)(p0 + i - base0);
            else filter_probe<S>(a, stg, p0 + i);  // queue full: probe right away
        }
        __syncwarp();
        if (stg->qcount >= 32u) {  // warp-uniform
            const unsigned int n = min(stg->qcount, (unsigned)kFilterQueue);
            for (unsigned int q = lane; q < n; q += 32) filter_probe<S>(a, stg, base0 + stg->queue[q]);
            __syncwarp();
            if (lane == 0) stg->qcount = 0u;
            __syncwarp();
        }
        if (stg->count >= (unsigned)kFilterStage / 3) filter_flush(a, stg);
    }
    __syncwarp();
    {
        const unsigned int n = min(stg->qcount, (unsigned)kFilterQueue);
        for (unsigned int q = lane; q < n; q += 32) filter_probe<S>(a, stg, base0 + stg->queue[q]);
    }
    filter_flush(a, stg);
}

// plan construction: the seed bitmap (up to 16 MiB) is built on the device from the sorted entry indices
__global__ void __launch_bounds__(256) filter_bitmap_kernel(uint32_t *bitmap, const uint32_t *__restrict__ ent_idx, int nent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nent) atomicOr(bitmap + (ent_idx[i] >> 5), 1u << (ent_idx[i] & 31));
}

// canonical witness: no (piece', shift') < (piece, shift) whose seed occurs inside the window
__device__ __forceinline__ bool filter_is_canonical(const uint8_t *P, const uint8_t *W, int m, int k, int s, int piece, int dk) {
    for (int i2 = 0; i2 <= piece; ++i2) {
        const int o2 = filter_piece_offset(i2, m, k);
        const int dmax = i2 == piece ? dk - 1 : 2 * k;
        for (int d2 = 0; d2 <= dmax; ++d2) {
            const int off = o2 + d2 - k;
            if (off < 0 || off + s > m) continue;
            bool same = true;
            for (int x = 0; x < s && same; ++x) same = P[o2 + x] == W[off + x];
            if (same) return false;
        }
    }
    return true;
}

// Verify: banded DP of one candidate per thread (cells with |row - col| > k are "infinite"): D[m][m] <= k is
// decided exactly.  A matching window is counted only by its canonical witness.  Random seed hits die within a
// few rows; a candidate that is still alive after kFilterShortRows rows is almost certainly a real match and is
// handed to filter_verify_long_kernel, where a whole warp finishes it.
constexpr int kFilterShortRows = 32;
__global__ void __launch_bounds__(128) filter_verify_kernel(const __grid_constant__ FilterArgs a) {
    if (*a.overflow) return;  // the band kernel takes this round instead
    const unsigned long long n = min(*a.ncand, a.cap);
    const int k = a.k;
    for (unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; c < n;
         c += (unsigned long long)gridDim.x * blockDim.x) {
        const uint64_t e = a.cand[c];
        const uint32_t slot = (uint32_t)(e >> 40);
        const int piece = (int)((e >> 34) & 63), dk = (int)((e >> 28) & 63);
        const long long j = a.w0 + (long long)(e & 0xFFFFFFFu);
        const int m = __ldg(a.fp_m + slot);
        const uint8_t *P = a.pat_bytes + __ldg(a.fp_off + slot);
        const uint8_t *W = a.buf + j;
        int st = band_dp_status(P, W, m, k, kFilterShortRows);
        if (st == 2) {
            if (k <= 15) {
                const unsigned long long pos = atomicAdd(a.nlong, 1ull);
                if (pos < a.longcap) {
                    a.longq[pos] = e;
                    continue;
                }
            }
            st = band_dp_status(P, W, m, k, 0x7FFFFFFF);  // queue full (or k = 16): finish here
        }
        if (st != 1 || !filter_is_canonical(P, W, m, k, a.s, piece, dk)) continue;
        atomicAdd(&a.counts[__ldg(a.fp_id + slot)], 1ull);
        if (a.sink.buf) hit_emit(a.sink, __ldg(a.fp_id + slot), j);
    }
}

// One warp per long-lived candidate: warp-parallel banded DP, then the canonical-witness test spread over the lanes.
__global__ void __launch_bounds__(128) filter_verify_long_kernel(const __grid_constant__ FilterArgs a) {
    if (*a.overflow) return;
    const unsigned long long n = min(*a.nlong, a.longcap);
    const int k = a.k, lane = threadIdx.x & 31;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    for (unsigned long long c = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < n; c += nwarps) {
        const uint64_t e = a.longq[c];
        const uint32_t slot = (uint32_t)(e >> 40);
        const int piece = (int)((e >> 34) & 63), dk = (int)((e >> 28) & 63);
        const long long j = a.w0 + (long long)(e & 0xFFFFFFFu);
        const int m = __ldg(a.fp_m + slot);
        const uint8_t *P = a.pat_bytes + __ldg(a.fp_off + slot);
        const uint8_t *W = a.buf + j;
        if (!band_dp_within_k_warp(P, W, m, k)) continue;  // warp-uniform
        // smaller witnesses (piece', shift') in linear order q = piece' * (2k+1) + shift'; lane l tests q = l, l+32, ..
        const int nq = piece * (2 * k + 1) + dk;
        bool found = false;
        for (int q = lane; q < nq && !found; q += 32) {
            const int i2 = q / (2 * k + 1), d2 = q - i2 * (2 * k + 1);
            const int o2 = filter_piece_offset(i2, m, k), off = o2 + d2 - k;
            if (off < 0 || off + a.s > m) continue;
            bool same = true;
            for (int x = 0; x < a.s && same; ++x) same = P[o2 + x] == W[off + x];
            found = same;
        }
        if (__any_sync(0xFFFFFFFFu, found)) continue;
        if (lane == 0) {
            atomicAdd(&a.counts[__ldg(a.fp_id + slot)], 1ull);
            if (a.sink.buf) hit_emit(a.sink, __ldg(a.fp_id + slot), j);
        }
    }
}

#endif  // __CUDACC__

}  // namespace apm
