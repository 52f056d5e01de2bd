// apm_tail.cuh -- bit-parallel evaluation of the TRUNCATED TAIL WINDOWS (src/sequential.c:131-136).
//
// For window starts j > n_bytes - m the reference compares the pattern PREFIX of length s = n_bytes - j with the
// s remaining text bytes: levenshtein(P[0..s), T[j..j+s), s) (src/utils.c:76-99).  There are m - 1 - k such
// windows per pattern (s = m-1 .. k+1).  The explicit-DP kernel (apm_dp.cuh: one thread per window, s^2 serial
// cells through a global scratch column) made them the critical path of small calls (4.6 of 5.7 ms in the
// round-1 smoke run).  Here the same cells are evaluated with Hyyro's global variant of Myers' bit-vector
// recurrence (apm_myers.cuh: myers_step): rows -> bits, one column step per text byte, so a window costs
// s * ceil(m/32) word steps instead of s^2 cells.  The Peq table of the FULL pattern is used -- rows below s never
// influence rows <= s -- and the distance is read at row s: D[s][s] = s + popc(Pv & rows<=s) - popc(Mv & rows<=s).
//
//   tail_myers_kernel       m <= 256: one THREAD per window, ceil(m/32) <= 8 words in registers
//   tail_myers_warp_kernel  m <= 1024: one WARP per window, lane = word; the carry of the column step's addition
//                           crosses lanes by a carry-lookahead on two warp ballots, the 1-bit shifts by one shuffle
//
// One CTA serves windows of ONE pattern and first builds that pattern's Peq in shared memory (compact alphabet of
// the plan: ncodes symbols).  Every cell of the truncated DP is accounted for, so the result is exact in every mode.
#pragma once
#include "apm_common.cuh"
#include "apm_dp.cuh"
#include "apm_myers.cuh"

namespace apm {

#ifdef __CUDACC__

constexpr int kTailThreads = 128;      // thread-per-window kernel: windows per CTA
constexpr int kTailWarpThreads = 256;  // warp-per-window kernel: 8 windows per CTA
constexpr int kTailThreadMaxWords = 8;

struct TailArgs {
    DpArgs d;                 // text, patterns, window range, counts, sink (scratch unused)
    const uint8_t *code_of;   // [256] byte -> compact alphabet code of the plan
    int ncodes;
    int chunks;               // CTAs per pattern
};

// Peq of pattern p in shared memory: peq[code * NW + w], bit x of word w <-> pattern row 32 w + x + 1
__device__ __forceinline__ void tail_build_peq(uint32_t *s_peq, const uint8_t *pat, int m, int NW, int ncodes,
                                               const uint8_t *code_of) {
    for (int i = threadIdx.x; i < ncodes * NW; i += blockDim.x) s_peq[i] = 0u;
    __syncthreads();
    for (int x = threadIdx.x; x < m; x += blockDim.x)
        atomicOr(&s_peq[(int)__ldg(code_of + pat[x]) * NW + (x >> 5)], 1u << (x & 31));
    __syncthreads();
}

template <int NW>
__device__ __forceinline__ int tail_thread_distance(const uint32_t *s_peq, const uint8_t *code_of,
                                                    const uint8_t *__restrict__ win, int size) {
    uint32_t Pv[NW], Mv[NW], Eq[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) { Pv[w] = 0xFFFFFFFFu; Mv[w] = 0u; }  // D[r][0] = r
    for (int c = 0; c < size; ++c) {
        const uint32_t *e = s_peq + (int)__ldg(code_of + win[c]) * NW;
#pragma unroll
        for (int w = 0; w < NW; ++w) Eq[w] = e[w];
        myers_step<NW>(Pv, Mv, Eq);
    }
    int d = size;  // D[0][size] = size, then the vertical deltas of rows 1..size
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const int rows = size - 32 * w;
        const uint32_t mask = rows >= 32 ? 0xFFFFFFFFu : (rows <= 0 ? 0u : ((1u << rows) - 1u));
        d += __popc(Pv[w] & mask) - __popc(Mv[w] & mask);
    }
    return d;
}

// grid.x = npat * chunks; CTA (pattern slot pi, chunk c) takes tail windows t = c * kTailThreads + tid
__global__ void __launch_bounds__(kTailThreads) tail_myers_kernel(const TailArgs a) {
    __shared__ uint32_t s_peq[256 * kTailThreadMaxWords];
    const int pi = blockIdx.x / a.chunks, chunk = blockIdx.x - pi * a.chunks;
    const int p = a.d.pat_list[pi];
    const int m = a.d.pat_len[p];
    const int NW = (m + 31) >> 5;
    const long long first = a.d.n_total - m + 1 > 0 ? a.d.n_total - m + 1 : 0;
    if (first + (long long)chunk * kTailThreads >= a.d.j_end) return;  // CTA-uniform: no tail window here
    const uint8_t *pat = a.d.pat_bytes + a.d.pat_off[p];
    tail_build_peq(s_peq, pat, m, NW, a.ncodes, a.code_of);
    const int t = chunk * kTailThreads + threadIdx.x;
    const long long j = first + t;
    if (j < a.d.j_begin || j >= a.d.j_end) return;
    const long long left = a.d.n_total - j;
    const int size = left < m ? (int)left : m;
    const uint8_t *win = a.d.buf + (j - a.d.buf_offset);
    int d;
    switch (NW) {
        case 1: d = tail_thread_distance<1>(s_peq, a.code_of, win, size); break;
        case 2: d = tail_thread_distance<2>(s_peq, a.code_of, win, size); break;
        case 3: d = tail_thread_distance<3>(s_peq, a.code_of, win, size); break;
        case 4: d = tail_thread_distance<4>(s_peq, a.code_of, win, size); break;
        case 5: d = tail_thread_distance<5>(s_peq, a.code_of, win, size); break;
        case 6: d = tail_thread_distance<6>(s_peq, a.code_of, win, size); break;
        case 7: d = tail_thread_distance<7>(s_peq, a.code_of, win, size); break;
        default: d = tail_thread_distance<8>(s_peq, a.code_of, win, size); break;
    }
    if (d <= a.d.k) {
        atomicAdd(&a.d.counts[p], 1ull);
        if (a.d.sink.buf) hit_emit(a.d.sink, p, j);
    }
}

// grid.x = npat * chunks; CTA (pattern slot pi, chunk c) takes tail windows t = c * 8 + warp.  Lane l holds rows
// 32 l + 1 .. 32 l + 32 of the column; dynamic shared memory: ncodes * 32 words.
__global__ void __launch_bounds__(kTailWarpThreads) tail_myers_warp_kernel(const TailArgs a) {
    extern __shared__ __align__(16) uint32_t s_peq_dyn[];
    constexpr int kWarps = kTailWarpThreads / 32;
    const int pi = blockIdx.x / a.chunks, chunk = blockIdx.x - pi * a.chunks;
    const int p = a.d.pat_list[pi];
    const int m = a.d.pat_len[p];
    const long long first = a.d.n_total - m + 1 > 0 ? a.d.n_total - m + 1 : 0;
    if (first + (long long)chunk * kWarps >= a.d.j_end) return;  // CTA-uniform
    const uint8_t *pat = a.d.pat_bytes + a.d.pat_off[p];
    tail_build_peq(s_peq_dyn, pat, m, 32, a.ncodes, a.code_of);
    const int lane = threadIdx.x & 31;
    const int t = chunk * kWarps + (threadIdx.x >> 5);
    const long long j = first + t;
    if (j < a.d.j_begin || j >= a.d.j_end) return;  // warp-uniform
    const long long left = a.d.n_total - j;
    const int size = left < m ? (int)left : m;
    const uint8_t *win = a.d.buf + (j - a.d.buf_offset);
    uint32_t Pv = 0xFFFFFFFFu, Mv = 0u;
    for (int c = 0; c < size; ++c) {
        const uint32_t Eq = s_peq_dyn[(int)__ldg(a.code_of + win[c]) * 32 + lane];
        // t + Pv over the 1024-bit column: per-lane sum, then the carries INTO every lane from a carry-lookahead
        // on the generate / propagate ballots (the carries of U + V with U = G | P, V = G are exactly that chain)
        const uint32_t tt = Eq & Pv;
        const uint32_t s0 = tt + Pv;
        const uint32_t G = __ballot_sync(0xFFFFFFFFu, s0 < tt);
        const uint32_t Pr = __ballot_sync(0xFFFFFFFFu, s0 == 0xFFFFFFFFu);
        const uint32_t U = G | Pr;
        const uint32_t cin = (((U + G) ^ U ^ G) >> lane) & 1u;
        const uint32_t s = s0 + cin;
        const uint32_t Xv = Eq | Mv;
        const uint32_t Xh = (s ^ Pv) | Eq;
        const uint32_t Ph = Mv | ~(Xh | Pv);
        const uint32_t Mh = Pv & Xh;
        // 1-bit shifts across lanes: both top bits travel in one shuffle; row 0 shifts in (+1, 0): D[0][c] = c
        uint32_t tops = __shfl_up_sync(0xFFFFFFFFu, (Ph >> 31) | ((Mh >> 31) << 1), 1);
        if (lane == 0) tops = 1u;
        const uint32_t Phs = (Ph << 1) | (tops & 1u);
        const uint32_t Mhs = (Mh << 1) | (tops >> 1);
        Pv = Mhs | ~(Xv | Phs);
        Mv = Phs & Xv;
    }
    const int rows = size - 32 * lane;
    const uint32_t mask = rows >= 32 ? 0xFFFFFFFFu : (rows <= 0 ? 0u : ((1u << rows) - 1u));
    const int d = size + __reduce_add_sync(0xFFFFFFFFu, __popc(Pv & mask) - __popc(Mv & mask));
    if (lane == 0 && d <= a.d.k) {
        atomicAdd(&a.d.counts[p], 1ull);
        if (a.d.sink.buf) hit_emit(a.d.sink, p, j);
    }
}

#endif  // __CUDACC__

}  // namespace apm
