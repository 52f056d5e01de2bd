// apm_dp.cuh -- explicit-DP fallback kernels (BASELINE north_star item 2).
//
// Evaluates src/utils.c:76-99 cell by cell for (a) the truncated tail windows of every pattern
// (src/sequential.c:131-134: size = n_bytes - j < m, the pattern PREFIX of that size is compared),
// (b) every window of patterns longer than the bit-parallel kernel supports (m > 32*kMaxWords), and
// (c) everything when option kernel=dp is set (an in-GPU cross-check of the bit-parallel kernel).
//
// One thread per (pattern, window).  The DP column of a thread lives in a global scratch array laid
// out [row][thread] so that the row sweep of a warp is coalesced and stays in L1/L2; unlike the
// reference's GPU clone (patterns_over_ranks.cu:31) there is no device-heap malloc, and the count is
// a proper 64-bit atomic (the reference's `(*local_matches)++` at :68 is a data race).
#pragma once
#include "apm_common.cuh"

namespace apm {

#ifdef __CUDACC__

struct DpArgs {
    const uint8_t *buf;           // device text; buf[0] = global byte buf_offset
    long long buf_offset;         // global index of buf[0]
    long long n_total;            // global text length
    long long j_begin, j_end;     // global window-start range, already clamped to n_total - k
    const uint8_t *pat_bytes;     // all patterns, concatenated
    const long long *pat_off;     // [P]
    const int *pat_len;           // [P]
    const int *pat_list;          // patterns handled by this launch
    int npat;
    int k;
    int tail_width;               // tail mode: window slots per pattern (>= max m-1-k)
    uint32_t *scratch;            // [(mmax+1)][scratch_stride]
    long long scratch_stride;     // = total threads of the launch
    unsigned long long *counts;
    HitSink sink;                 // optional match-position output (window starts are global here: base = 0)
};

// Distance of pattern[0..size) vs text window [j, j+size): the DP of utils.c:76-99 with the column
// in scratch[row * stride].  Rows/cols are 1-based as in the reference.
__device__ __forceinline__ uint32_t dp_window(const uint8_t *__restrict__ pat, const uint8_t *__restrict__ win,
                                              int size, uint32_t *__restrict__ col, long long stride) {
    for (int r = 1; r <= size; ++r) col[(long long)r * stride] = (uint32_t)r;
    uint32_t last = (uint32_t)size;
    for (int c = 1; c <= size; ++c) {
        const uint32_t tc = win[c - 1];
        uint32_t diag = (uint32_t)(c - 1);  // D[0][c-1]
        uint32_t up = (uint32_t)c;          // D[0][c]
        for (int r = 1; r <= size; ++r) {
            const uint32_t left = col[(long long)r * stride];  // D[r][c-1]
            const uint32_t sub = diag + (pat[r - 1] == tc ? 0u : 1u);
            const uint32_t v = min(sub, min(left, up) + 1u);
            col[(long long)r * stride] = v;
            diag = left;
            up = v;
        }
        last = up;
    }
    return last;
}

// Banded version of the same DP for the DECISION distance <= k (k <= kBandDpMaxK): cells with |row - col| > k are
// treated as infinite, which leaves every value <= k exact (Ukkonen).  One row of 2k+1 cells in local memory, no
// global scratch, early exit when a whole row exceeds k.  Used by the filter's verification and, in the exact
// shortcut modes, for the truncated tail windows (size x size instead of m x m).
constexpr int kBandDpMaxK = 16;

// KB = compile-time band half-width >= k (a wider band is just as exact): the row lives in registers and the
// sweep over its 2 KB + 1 cells is fully unrolled, so a row costs one dependent min/add chain instead of a chain
// of local-memory round trips.
// Returns 0 = distance > k, 1 = distance <= k, 2 = still undecided after max_rows rows (size > max_rows).
template <int KB>
__device__ __forceinline__ int band_dp_status_t(const uint8_t *__restrict__ P, const uint8_t *__restrict__ W, int size, int k,
                                                int max_rows) {
    constexpr int INF = 1 << 20;
    int band[2 * KB + 2];  // band[x] = D[r][r + x - KB]
#pragma unroll
    for (int x = 0; x <= 2 * KB; ++x) band[x] = x >= KB ? x - KB : INF;  // row 0: D[0][c] = c
    band[2 * KB + 1] = INF;
    const int rows = min(size, max_rows);
    for (int r = 1; r <= rows; ++r) {
        const uint32_t pc = P[r - 1];
        int left = INF, best = INF;
#pragma unroll
        for (int x = 0; x <= 2 * KB; ++x) {
            const int col = r + x - KB;
            int v;
            if (col < 0 || col > size) v = INF;
            else if (col == 0) v = r;
            else {
                const int diag = band[x] + (pc == W[col - 1] ? 0 : 1);
                v = min(diag, min(band[x + 1], left) + 1);
            }
            band[x] = v;  // band[x + 1] (row r-1) is still untouched when the next x reads it as "diag"
            left = v;
            best = min(best, v);
        }
        if (best > k) return 0;
    }
    if (rows < size) return 2;
    return band[KB] <= k ? 1 : 0;
}

__device__ __forceinline__ int band_dp_status(const uint8_t *__restrict__ P, const uint8_t *__restrict__ W, int size, int k,
                                              int max_rows) {
    if (k <= 2) return band_dp_status_t<2>(P, W, size, k, max_rows);
    if (k <= 4) return band_dp_status_t<4>(P, W, size, k, max_rows);
    if (k <= 6) return band_dp_status_t<6>(P, W, size, k, max_rows);
    if (k <= 8) return band_dp_status_t<8>(P, W, size, k, max_rows);
    if (k <= 12) return band_dp_status_t<12>(P, W, size, k, max_rows);
    return band_dp_status_t<16>(P, W, size, k, max_rows);
}
__device__ __forceinline__ bool band_dp_within_k(const uint8_t *__restrict__ P, const uint8_t *__restrict__ W, int size, int k) {
    return band_dp_status(P, W, size, k, 0x7FFFFFFF) == 1;
}

// The same decision by a whole WARP (k <= 15): lane x holds band cell x of the current row.  The serial
// dependency along the row, v[x] = min(c[x], v[x-1] + 1), is a prefix minimum of c[y] - y (five shuffles), so a
// row costs ~25 instructions of latency instead of 2k+1 dependent min/add pairs: ~10x faster for long patterns
// whose candidates survive to the last row.  All lanes must call it with the same arguments.
__device__ __forceinline__ bool band_dp_within_k_warp(const uint8_t *__restrict__ P, const uint8_t *__restrict__ W, int size, int k) {
    constexpr int INF = 1 << 20;
    const int x = threadIdx.x & 31;
    const bool lane_used = x <= 2 * k;
    int old = (lane_used && x >= k) ? x - k : INF;  // row 0: D[0][c] = c
    for (int r = 1; r <= size; ++r) {
        const uint32_t pc = P[r - 1];
        const int col = r + x - k;
        int up = __shfl_down_sync(0xFFFFFFFFu, old, 1);
        if (x >= 2 * k) up = INF;
        int c = INF;
        if (lane_used && col >= 0 && col <= size) c = col == 0 ? r : min(old + (pc == W[col - 1] ? 0 : 1), up + 1);
        int t = c - x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xFFFFFFFFu, t, d);
            if (x >= d) t = min(t, y);
        }
        old = (lane_used && col >= 0 && col <= size) ? min(t + x, INF) : INF;
        if (__reduce_min_sync(0xFFFFFFFFu, old) > k) return false;
    }
    return __shfl_sync(0xFFFFFFFFu, old, k) <= k;
}

// Tail mode: thread (p, t) evaluates window j = max(0, n_total - m_p + 1) + t if it lies in
// [j_begin, j_end).  All such windows are truncated (size < m) -- or the whole text is shorter
// than the pattern.
__global__ void __launch_bounds__(128) dp_tail_kernel(const DpArgs a) {
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)a.npat * a.tail_width;
    if (gtid >= total) return;
    const int pi = (int)(gtid / a.tail_width);
    const int t = (int)(gtid % a.tail_width);
    const int p = a.pat_list[pi];
    const int m = a.pat_len[p];
    const long long first = a.n_total - m + 1 > 0 ? a.n_total - m + 1 : 0;
    const long long j = first + t;
    if (j < a.j_begin || j >= a.j_end) return;
    const long long left = a.n_total - j;
    const int size = left < m ? (int)left : m;
    const uint32_t d = dp_window(a.pat_bytes + a.pat_off[p], a.buf + (j - a.buf_offset), size,
                                 a.scratch + gtid, a.scratch_stride);
    if (d <= (uint32_t)a.k) {
        atomicAdd(&a.counts[p], 1ull);
        if (a.sink.buf) hit_emit(a.sink, p, j);
    }
}

// Tail mode with the banded DP (exact shortcut modes, k <= kBandDpMaxK): same mapping, no scratch.
__global__ void __launch_bounds__(128) dp_tail_band_kernel(const DpArgs a) {
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)a.npat * a.tail_width;
    if (gtid >= total) return;
    const int pi = (int)(gtid / a.tail_width);
    const int t = (int)(gtid % a.tail_width);
    const int p = a.pat_list[pi];
    const int m = a.pat_len[p];
    const long long first = a.n_total - m + 1 > 0 ? a.n_total - m + 1 : 0;
    const long long j = first + t;
    if (j < a.j_begin || j >= a.j_end) return;
    const long long left = a.n_total - j;
    const int size = left < m ? (int)left : m;
    if (size <= a.k || band_dp_within_k(a.pat_bytes + a.pat_off[p], a.buf + (j - a.buf_offset), size, a.k)) {
        atomicAdd(&a.counts[p], 1ull);
        if (a.sink.buf) hit_emit(a.sink, p, j);
    }
}

// All-windows mode: blockIdx.y = pattern slot, grid-stride over window starts [j_begin, j_end).
__global__ void __launch_bounds__(128) dp_all_kernel(const DpArgs a) {
    const int p = a.pat_list[blockIdx.y];
    const int m = a.pat_len[p];
    const uint8_t *pat = a.pat_bytes + a.pat_off[p];
    const long long tx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nx = (long long)gridDim.x * blockDim.x;
    uint32_t *col = a.scratch + ((long long)blockIdx.y * nx + tx);
    unsigned int hits = 0;
    for (long long j = a.j_begin + tx; j < a.j_end; j += nx) {
        const long long left = a.n_total - j;
        const int size = left < m ? (int)left : m;
        const uint32_t d = dp_window(pat, a.buf + (j - a.buf_offset), size, col, a.scratch_stride);
        hits += (d <= (uint32_t)a.k);
        if (a.sink.buf && d <= (uint32_t)a.k) hit_emit(a.sink, p, j);
    }
    // warp-aggregate (all lanes of a warp share the pattern)
    for (int o = 16; o > 0; o >>= 1) hits += __shfl_down_sync(0xFFFFFFFFu, hits, o);
    if ((threadIdx.x & 31) == 0 && hits) atomicAdd(&a.counts[p], (unsigned long long)hits);
}

#endif  // __CUDACC__

}  // namespace apm
