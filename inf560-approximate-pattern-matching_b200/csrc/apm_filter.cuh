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
// pass over the text, independent of the number of patterns), probes a bitmap of all seed hashes (L2 resident),
// and for the rare hits looks the seed up, compares the bytes and emits the (pattern, window, piece, shift)
// candidates.  The verify kernel evaluates the banded DP |row - col| <= k of src/utils.c:76-99 for each
// candidate -- exact for the decision D <= k -- and counts a matching window only from its CANONICAL witness (the
// lexicographically smallest (piece, shift) whose seed occurs), so a window witnessed several times is counted
// once.  No sort, no host round trip.  If the candidate buffer overflows (low-complexity text) a flag makes the
// verify kernel a no-op and switches on the band kernel for the same patterns and windows instead: always exact.
//
// Work: O(text bytes) + O(candidates * m * (2k+1)) instead of O(text bytes * patterns * m * (2k+1)).
#pragma once
#include "apm_common.cuh"

namespace apm {

constexpr int kFilterMinSeed = 8;
constexpr int kFilterMaxSeed = 16;
constexpr int kFilterMaxK = 16;
constexpr int kFilterThreads = 128;
constexpr int kFilterPosPerThread = 16;
constexpr uint32_t kFilterHashB = 0x9E3779B1u;
constexpr uint32_t kFilterMixM = 0x2C1B3C6Du;
constexpr int kFilterSlabLog = 27;  // window starts per scan/verify round: candidates carry a 28-bit local start

// table index of a seed hash (hb bits)
__host__ __device__ __forceinline__ uint32_t filter_index(uint32_t h, int hb) {
    h ^= h >> 15;
    h *= kFilterMixM;
    return h >> (32 - hb);
}
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
    unsigned long long *counts;
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

// slow path of the scan: text position t produced table index idx
__device__ __noinline__ void filter_hit(const FilterArgs &a, long long t, uint32_t idx) {
    int lo = 0, hi = a.nent;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a.ent_idx + mid) < idx) lo = mid + 1;
        else hi = mid;
    }
    for (int e = lo; e < a.nent && __ldg(a.ent_idx + e) == idx; ++e) {
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
            const unsigned long long pos = atomicAdd(a.ncand, 1ull);
            if (pos < a.cap) a.cand[pos] = filter_pack(slot, piece, d + a.k, (uint32_t)(j - a.w0));
            else *a.overflow = 1u;
        }
    }
}

// Scan: every text position that can hold a seed of a window of [w0, w1).  A thread owns 16 consecutive
// positions: its own 16 bytes come from one coalesced 128-bit load, the next 16 from the neighbouring lane by
// shuffle; rolling hash in registers; the 16 bitmap probes are issued together.
template <int S>
__global__ void __launch_bounds__(kFilterThreads) filter_scan_kernel(const FilterArgs a) {
    const long long t_begin = a.w0;
    const long long t_end = min(a.w1 - 1 + a.mmax, a.n_end) - S + 1;  // exclusive
    if (t_end <= t_begin) return;
    // align the tile grid to 16-byte addresses of the buffer
    const long long mis = (long long)(reinterpret_cast<uintptr_t>(a.buf + t_begin) & 15);
    const long long base0 = t_begin - mis;
    constexpr long long kTile = (long long)kFilterThreads * kFilterPosPerThread;
    const int lane = threadIdx.x & 31;
    for (long long base = base0 + (long long)blockIdx.x * kTile; base < t_end; base += (long long)gridDim.x * kTile) {
        const long long p0 = base + (long long)threadIdx.x * kFilterPosPerThread;
        const uint4 own = filter_load16(a, p0);
        uint4 nxt;
        nxt.x = __shfl_down_sync(0xFFFFFFFFu, own.x, 1);
        nxt.y = __shfl_down_sync(0xFFFFFFFFu, own.y, 1);
        nxt.z = __shfl_down_sync(0xFFFFFFFFu, own.z, 1);
        nxt.w = __shfl_down_sync(0xFFFFFFFFu, own.w, 1);
        if (lane == 31) nxt = filter_load16(a, p0 + 16);
        const uint32_t w[8] = {own.x, own.y, own.z, own.w, nxt.x, nxt.y, nxt.z, nxt.w};
        auto byte_at = [&](int i) -> uint32_t { return (w[i >> 2] >> (8 * (i & 3))) & 0xFFu; };
        uint32_t h = 0;
#pragma unroll
        for (int i = 0; i < S; ++i) h = h * kFilterHashB + byte_at(i);
        uint32_t idx[kFilterPosPerThread], word[kFilterPosPerThread];
#pragma unroll
        for (int i = 0; i < kFilterPosPerThread; ++i) {
            idx[i] = filter_index(h, a.hb);
            word[i] = __ldg(a.bitmap + (idx[i] >> 5));
            if (i + 1 < kFilterPosPerThread) h = h * kFilterHashB - byte_at(i) * a.bs + byte_at(i + S);  // roll by one byte
        }
#pragma unroll
        for (int i = 0; i < kFilterPosPerThread; ++i) {
            const long long t = p0 + i;
            if (((word[i] >> (idx[i] & 31)) & 1u) && t >= t_begin && t < t_end) filter_hit(a, t, idx[i]);
        }
    }
}

// Verify: banded DP of one candidate (cells with |row - col| > k are "infinite"): D[m][m] <= k is decided
// exactly.  A matching window is counted only by its canonical witness.
__global__ void __launch_bounds__(128) filter_verify_kernel(const FilterArgs a) {
    if (*a.overflow) return;  // the band kernel takes this round instead
    const unsigned long long n = min(*a.ncand, a.cap);
    constexpr int INF = 1 << 20;
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
        int band[2 * kFilterMaxK + 3];  // band[x] = D[r][r + x - k]
        for (int x = 0; x <= 2 * k; ++x) band[x] = x >= k ? x - k : INF;  // row 0: D[0][c] = c
        band[2 * k + 1] = INF;
        bool alive = true;
        for (int r = 1; r <= m && alive; ++r) {
            const uint32_t pc = P[r - 1];
            int left = INF, best = INF;
            for (int x = 0; x <= 2 * k; ++x) {
                const int col = r + x - k;
                int v;
                if (col < 0 || col > m) v = INF;
                else if (col == 0) v = r;
                else {
                    const int diag = band[x] + (pc == W[col - 1] ? 0 : 1);
                    v = min(diag, min(band[x + 1], left) + 1);
                }
                band[x] = v;  // band[x + 1] (row r-1) is still untouched when the next x reads it as "diag"
                left = v;
                best = min(best, v);
            }
            alive = best <= k;
        }
        if (!alive || band[k] > k) continue;
        // canonical witness: no (piece', shift') < (piece, shift) whose seed occurs inside the window
        bool canonical = true;
        for (int i2 = 0; i2 <= piece && canonical; ++i2) {
            const int o2 = filter_piece_offset(i2, m, k);
            const int dmax = i2 == piece ? dk - 1 : 2 * k;
            for (int d2 = 0; d2 <= dmax && canonical; ++d2) {
                const int off = o2 + d2 - k;
                if (off < 0 || off + a.s > m) continue;
                bool same = true;
                for (int x = 0; x < a.s && same; ++x) same = P[o2 + x] == W[off + x];
                if (same) canonical = false;
            }
        }
        if (canonical) atomicAdd(&a.counts[__ldg(a.fp_id + slot)], 1ull);
    }
}

#endif  // __CUDACC__

}  // namespace apm
