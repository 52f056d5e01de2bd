// apm_sliced.cuh -- window-sliced bit-parallel kernel: the same DP as src/utils.c:76-99, every one of the
// m x m cells evaluated, but with the 32 bits of a register holding the SAME cell of 32 CONSECUTIVE WINDOWS
// (bit b <-> window start j0 + b) instead of 32 rows of one window.
//
// Why: in the row-parallel (Hyyro/Myers) formulation of apm_myers.cuh a column step costs 7 boolean
// instructions + 1 carry add + 2 shifts per 32 cells, and on sm_100 it is bound by the ALU pipe (LOP3/SHF
// /IADD3 with carry: 64 lanes/clk/SM); the cross-word carries of m > 32 add more ALU work.  Slicing across
// windows makes every cell's recurrence purely boolean -- no adder, no shifts, no carries between words:
//
//   a = D[i][j-1]-D[i-1][j-1] (vertical delta, left neighbour)    b = D[i-1][j]-D[i-1][j-1] (horizontal, above)
//   d0 = eq | a- | b-                      diagonal step is 0
//   v+ = b- | ~(d0 | b+)   v- = b+ & d0    vertical delta of this cell   (goes right)
//   h+ = a- | ~(d0 | a+)   h- = a+ & d0    horizontal delta of this cell (goes down)
//
// = 5 LOP3 per cell per 32 windows (0.156 instructions per DP cell, vs >= 0.27 ALU-pipe instructions per cell
// for the row-parallel step), identical for every pattern length.  These are Myers' delta equations applied
// per cell; the boundary D[0][j] = j, D[i][0] = i (utils.c:84-88) is "+1" in both directions.
//
// Mapping
//   thread t of a CTA  <-> the 32 windows starting at tile_start + 32 t
//   match bits         <-> eq(i, j) for those 32 windows is bits [32t + j-1, +32) of the occurrence bit-vector
//                          of symbol P[i-1] in the text tile.  All 32-bit windows of those bit-vectors are
//                          precomputed once per tile into shared memory (U table, one plane per symbol),
//                          laid out so that the 32 lanes of a warp read conflict-free with LDS.64 and an
//                          immediate offset per column.
//   state              <-> the horizontal deltas of the previous row for all MC columns live in 2*MC
//                          registers; the sweep is row-major (pattern symbol outer, text column inner), so
//                          the plane of the row's symbol is selected once per row.
//   distance           <-> D[m][m] = m + sum_j (h+ - h-) of the last row: a carry-save adder tree over the
//                          bit-planes (Harley-Seal), then a bit-sliced compare with the threshold.
//   text tile          <-> same TMA 1-D bulk staging + re-coding as the row-parallel kernel (apm_tile.cuh).
#pragma once
#include "apm_common.cuh"
#include "apm_tile.cuh"

namespace apm {

constexpr int kSlicedThreads = 128;
constexpr int kSlicedTile = 32 * kSlicedThreads;  // window starts per tile
constexpr int kURowBytes = 136;                   // 34 words per U row: lanes hit distinct banks with LDS.64
constexpr int kSlicedMaxPlanes = 8;               // symbols with their own occurrence plane
constexpr uint8_t kNoPlane = 0xFF;

#ifdef __CUDACC__

struct SlicedArgs {
    const uint8_t *buf;          // device text; buf[0] is global byte buf_offset
    long long buf_len, n_end;    // valid bytes; local index of the global end of text
    long long w0, w1;            // local window-start range
    const uint8_t *pat_codes;    // [npat][MC] plane index of every pattern symbol (padded)
    const int *pat_m;            // [npat]
    const int *pat_id;           // [npat] index into counts
    const uint8_t *plane_of;     // [256] byte -> plane index or kNoPlane
    unsigned long long *counts;
    int npat, pats_per_chunk, nplanes, k;
};

// geometry shared by host and device
__host__ __device__ constexpr int sliced_rowsU(int MC) { return kSlicedTile / 32 + MC / 32 + 1; }  // rows per U plane
__host__ __device__ constexpr int sliced_nB(int MC) { return sliced_rowsU(MC) + 1; }  // words per bit-vector
__host__ __device__ constexpr int sliced_span(int MC) { return sliced_nB(MC) * 32; }  // text positions per tile

// Shared-memory layout: [barrier | byte->plane map | B bit-vectors | pattern chunk | U table].  The raw
// text tile is staged by TMA INSIDE the (not yet built) U region, so it costs no extra shared memory and
// three CTAs fit an SM for the 4-symbol DNA alphabet.
template <int MC>
__host__ __device__ inline size_t sliced_smem_fixed(int nplanes) {
    size_t b = 64 + 256 + (size_t)nplanes * sliced_nB(MC) * 4;
    return (b + 15) / 16 * 16;
}
template <int MC>
__host__ __device__ inline size_t sliced_smem_chunk(int pats_per_chunk) {
    return ((size_t)pats_per_chunk * (MC + 3 * sizeof(int)) + 16 + 15) / 16 * 16;
}
template <int MC>
__host__ __device__ inline size_t sliced_smem_bytes(int pats_per_chunk, int nplanes) {
    return sliced_smem_fixed<MC>(nplanes) + sliced_smem_chunk<MC>(pats_per_chunk) +
           (size_t)nplanes * sliced_rowsU(MC) * kURowBytes;
}

// one LOP3: any boolean function of three words, LUT evaluated on a = 0xF0, b = 0xCC, c = 0xAA.
// Written as PTX so that the five-instruction decomposition of a cell is exactly what gets executed
// (left to itself the compiler re-associates the expressions into six LOP3s per cell).
template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
}
constexpr int kLutOr3 = 0xFE;        // a | b | c
constexpr int kLutAndOr = 0xE0;      // a & (b | c)
constexpr int kLutOrNor = 0xF1;      // a | ~(b | c)

// full adder on bit-planes: a <- a ^ b ^ c, returns majority(a, b, c)   (2 LOP3)
__device__ __forceinline__ uint32_t plane_fa(uint32_t &a, uint32_t b, uint32_t c) {
    const uint32_t carry = (a & b) | (c & (a | b));
    a = a ^ b ^ c;
    return carry;
}

// Harley-Seal style accumulation: acc[l] holds the weight-2^l plane of the running sum, pend[l] a pending
// carry of the same weight.  IDX is the (compile-time) index of the pair being pushed.
template <int L, int IDX, int NL>
struct PushCarry {
    static __device__ __forceinline__ void run(uint32_t (&acc)[NL], uint32_t (&pend)[NL], uint32_t c) {
        if constexpr (IDX & 1) {
            const uint32_t c2 = plane_fa(acc[L], pend[L], c);
            PushCarry<L + 1, (IDX >> 1), NL>::run(acc, pend, c2);
        } else {
            pend[L] = c;
        }
    }
};

template <int MC, int I, int NL>
struct SumPlanes {  // adds the planes of columns 2I and 2I+1 (h+ and ~h- of each), recursion over I
    static __device__ __forceinline__ void run(const uint32_t (&hp)[MC], const uint32_t (&hm)[MC], uint32_t (&acc)[NL],
                                               uint32_t (&pend)[NL]) {
        if constexpr (I < MC) {
            // pair index 2I: (h+[I], ~h-[I]) ; the second pair of this step comes from column I+1
            const uint32_t c = plane_fa(acc[0], hp[I], ~hm[I]);
            PushCarry<1, I, NL>::run(acc, pend, c);
            SumPlanes<MC, I + 1, NL>::run(hp, hm, acc, pend);
        }
    }
};

// One DP cell for 32 windows.  (ap, am) = vertical delta coming from the left neighbour, (bp, bm) =
// horizontal delta coming from the row above; both are replaced by the deltas of this cell.
__device__ __forceinline__ void sliced_cell(uint32_t q, uint32_t &ap, uint32_t &am, uint32_t &bp, uint32_t &bm) {
    // a+ & a- = 0 and b+ & b- = 0, so b+ & d0 = b+ & (eq | a-): the chain a- -> a-' that links a cell to
    // its right neighbour is a single LOP3 deep
    const uint32_t d0 = lop3<kLutOr3>(q, am, bm);
    const uint32_t vm = lop3<kLutAndOr>(bp, q, am);
    const uint32_t vp = lop3<kLutOrNor>(bm, d0, bp);
    const uint32_t hp2 = lop3<kLutOrNor>(am, d0, ap);
    const uint32_t hm2 = lop3<kLutAndOr>(ap, q, bm);
    bp = hp2;
    bm = hm2;
    ap = vp;
    am = vm;
}

// Row-major sweep of the m x m matrix for one pattern.  urow = this thread's row origin inside plane 0 of
// the U table, pc = plane index of every pattern symbol (one readable byte past the end).
// FULL (m == MC): branch-free, software pipelined -- the match words of the next column group (same row,
// or the first group of the next row) are requested before the current group is computed.
template <int MC, bool FULL>
__device__ __forceinline__ void sliced_sweep(const unsigned char *__restrict__ urow, const uint8_t *__restrict__ pc,
                                             int m, uint32_t plane_bytes, uint32_t (&hp)[MC], uint32_t (&hm)[MC]) {
    if constexpr (FULL) {
        constexpr int G = 8;  // columns per pipeline group: 4 LDS.64 in flight
        constexpr int NG = MC / G;
        const unsigned char *e = urow + (uint32_t)pc[0] * plane_bytes;
        uint2 buf[G / 2];
#pragma unroll
        for (int q = 0; q < G / 2; ++q) buf[q] = *reinterpret_cast<const uint2 *>(e + q * 8);
#pragma unroll 1
        for (int i = 0; i < MC; ++i) {
            const unsigned char *e_next = urow + (uint32_t)pc[i + 1] * plane_bytes;
            uint32_t ap = 0xFFFFFFFFu, am = 0u;  // D[i][0] - D[i-1][0] = +1
#pragma unroll
            for (int gi = 0; gi < NG; ++gi) {
                uint2 cur[G / 2];
#pragma unroll
                for (int q = 0; q < G / 2; ++q) cur[q] = buf[q];
                const int c1 = (gi + 1) * G;
#pragma unroll
                for (int q = 0; q < G / 2; ++q) {
                    if (gi + 1 < NG)
                        buf[q] = *reinterpret_cast<const uint2 *>(e + ((c1 + 2 * q) >> 5) * kURowBytes + ((c1 + 2 * q) & 31) * 4);
                    else
                        buf[q] = *reinterpret_cast<const uint2 *>(e_next + q * 8);
                }
#pragma unroll
                for (int cc = 0; cc < G; ++cc)
                    sliced_cell((cc & 1) ? cur[cc >> 1].y : cur[cc >> 1].x, ap, am, hp[gi * G + cc], hm[gi * G + cc]);
            }
            e = e_next;
        }
    } else {
        constexpr int CH = 8;  // columns per uniform early-exit check
#pragma unroll 1
        for (int i = 0; i < m; ++i) {
            const unsigned char *e = urow + (uint32_t)pc[i] * plane_bytes;
            uint32_t ap = 0xFFFFFFFFu, am = 0u;
#pragma unroll
            for (int c = 0; c < MC; c += 2) {
                if ((c % CH) == 0 && c >= m) break;  // uniform: m is the same for the whole CTA
                const uint2 eq = *reinterpret_cast<const uint2 *>(e + (c >> 5) * kURowBytes + (c & 31) * 4);
                sliced_cell(eq.x, ap, am, hp[c], hm[c]);
                sliced_cell(eq.y, ap, am, hp[c + 1], hm[c + 1]);
            }
        }
#pragma unroll
        for (int j = 0; j < MC; ++j)  // columns >= m: back to the neutral boundary value (they add a constant)
            if (j >= m) { hp[j] = 0xFFFFFFFFu; hm[j] = 0u; }
    }
}

template <int MC>
__global__ void __launch_bounds__(kSlicedThreads, MC == 64 ? 3 : 4) sliced_count_kernel(const SlicedArgs a) {
    static_assert(MC == 32 || MC == 64, "MC must be 32 or 64");
    constexpr int LOG = MC == 32 ? 6 : 7;  // 2*MC planes are summed: value range [0, 2*MC]
    constexpr int NL = LOG + 1;
    extern __shared__ __align__(128) unsigned char smem[];

    const int tid = threadIdx.x;
    constexpr int nB = sliced_nB(MC);        // words per occurrence bit-vector
    constexpr int rowsU = sliced_rowsU(MC);  // rows per U plane
    const size_t off_B = 64 + 256;
    const size_t off_pc = sliced_smem_fixed<MC>(a.nplanes);
    const size_t off_pm = off_pc + (((size_t)a.pats_per_chunk * MC + 16 + 3) & ~size_t(3));
    const size_t off_pid = off_pm + sizeof(int) * a.pats_per_chunk;
    const size_t off_cnt = off_pid + sizeof(int) * a.pats_per_chunk;
    const size_t off_U = off_pc + sliced_smem_chunk<MC>(a.pats_per_chunk);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    uint8_t *s_map = smem + 64;
    uint32_t *s_B = reinterpret_cast<uint32_t *>(smem + off_B);
    uint8_t *s_pc = smem + off_pc;
    int *s_pm = reinterpret_cast<int *>(smem + off_pm);
    int *s_pid = reinterpret_cast<int *>(smem + off_pid);
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(smem + off_cnt);
    uint8_t *s_raw = smem + off_U;  // raw text tile, overwritten by the U table once the bit-vectors exist
    const uint32_t plane_bytes = rowsU * kURowBytes;

    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    for (int i = tid; i < 256; i += kSlicedThreads) s_map[i] = a.plane_of[i];
    __syncthreads();

    const long long nwin = a.w1 - a.w0;
    const long long ntiles = (nwin + kSlicedTile - 1) / kSlicedTile;
    constexpr int span = sliced_span(MC);
    uint32_t phase = 0;

    // Work items = (pattern chunk, text tile), chunk-major, dealt round-robin to the persistent CTAs.
    const int nchunks = (a.npat + a.pats_per_chunk - 1) / a.pats_per_chunk;
    const long long nitems = ntiles * nchunks;
    int cur_chunk = -1, pcount = 0;
    for (long long it = blockIdx.x; it < nitems; it += gridDim.x) {
        const int chunk = (int)(it / ntiles);
        const long long t = it % ntiles;
        // the tile covers text positions [ts, ts + span); te clips it to the buffer
        const TileGeom g = tile_geometry(a.buf, a.buf_len, a.w0, t, kSlicedTile, span - kSlicedTile);
        if (tid == 0) tile_issue(g, a.buf, s_raw, bar);  // TMA bulk copy; overlaps the chunk reload below
        if (chunk != cur_chunk) {  // (re)load the chunk's patterns; thread i owns slot i of both loops
            const int p0 = chunk * a.pats_per_chunk;
            for (int i = tid; i < pcount; i += kSlicedThreads) {
                const uint32_t c = s_cnt[i];
                if (c) atomicAdd(&a.counts[s_pid[i]], (unsigned long long)c);
            }
            pcount = min(a.pats_per_chunk, a.npat - p0);
            for (int i = tid; i < pcount; i += kSlicedThreads) {
                s_pm[i] = a.pat_m[p0 + i];
                s_pid[i] = a.pat_id[p0 + i];
                s_cnt[i] = 0;
            }
            for (int i = tid; i < pcount * MC; i += kSlicedThreads) s_pc[i] = a.pat_codes[(size_t)p0 * MC + i];
            if (tid < 16) s_pc[pcount * MC + tid] = 0;  // the row loop reads one code past the last row
            cur_chunk = chunk;
        }
        if (g.tb > g.ta) {
            mbar_wait(bar, phase);
            phase ^= 1u;
        }
        // ---- occurrence bit-vectors: bit x of B[p] <-> text position ts + x holds the symbol of plane p.
        //      Bytes outside the TMA box (<= 15 on either side) are read from global memory.
        {
            const int lane = tid & 31;
            for (int w = tid >> 5; w < nB; w += kSlicedThreads / 32) {
                const long long i = g.ts + 32 * w + lane;
                uint32_t c = kNoPlane;
                if (i < g.te) c = s_map[(i >= g.ta && i < g.tb) ? s_raw[i - g.a0] : a.buf[i]];
                for (int p = 0; p < a.nplanes; ++p) {
                    const uint32_t bits = __ballot_sync(0xFFFFFFFFu, c == (uint32_t)p);
                    if (lane == 0) s_B[p * nB + w] = bits;
                }
            }
        }
        __syncthreads();
        // ---- U table: U[p][w][s] = bits [32 w + s, +32) of B[p]
        for (int idx = tid; idx < a.nplanes * rowsU; idx += kSlicedThreads) {
            const int p = idx / rowsU, w = idx - p * rowsU;
            const uint32_t lo = s_B[p * nB + w], hi = s_B[p * nB + w + 1];
            uint32_t *dst = reinterpret_cast<uint32_t *>(smem + off_U + (size_t)p * plane_bytes + (size_t)w * kURowBytes);
#pragma unroll
            for (int s = 0; s < 32; ++s) dst[s] = __funnelshift_r(lo, hi, s);
        }
        __syncthreads();

        // ---- hot loop: patterns x rows x columns, 5 LOP3 per cell
        const long long tile_end = min(g.ts + (long long)kSlicedTile, a.w1);
        const long long jbase = g.ts + 32ll * tid;
        const unsigned char *urow = smem + off_U + (size_t)tid * kURowBytes;
        for (int pi = 0; pi < pcount; ++pi) {
            const int m = s_pm[pi];
            const long long lim = min(tile_end, a.n_end - m + 1);  // full windows only
            const long long nvalid = lim - jbase;
            const uint32_t validmask = nvalid >= 32 ? 0xFFFFFFFFu : (nvalid <= 0 ? 0u : ((1u << (int)nvalid) - 1u));
            uint32_t hits = 0;
            if (validmask != 0u) {
                uint32_t hp[MC], hm[MC];
#pragma unroll
                for (int j = 0; j < MC; ++j) { hp[j] = 0xFFFFFFFFu; hm[j] = 0u; }  // D[0][j] - D[0][j-1] = +1
                const uint8_t *pc = s_pc + pi * MC;
                if (m == MC) sliced_sweep<MC, true>(urow, pc, m, plane_bytes, hp, hm);
                else sliced_sweep<MC, false>(urow, pc, m, plane_bytes, hp, hm);
                // V = sum_j (h+[j] + ~h-[j]) = D[m][m] + 2 (MC - m), bit-sliced in acc[0..LOG-1] + pend[LOG]
                uint32_t acc[NL], pend[NL];
#pragma unroll
                for (int l = 0; l < NL; ++l) { acc[l] = 0u; pend[l] = 0u; }
                SumPlanes<MC, 0, NL>::run(hp, hm, acc, pend);
                acc[LOG] = pend[LOG];
                // windows with V <= T
                const int T = a.k + 2 * (MC - m);
                uint32_t le;
                if (T >= 2 * MC) le = 0xFFFFFFFFu;
                else {
                    uint32_t lt = 0u, eqm = 0xFFFFFFFFu;
#pragma unroll
                    for (int l = LOG; l >= 0; --l) {
                        const uint32_t tb = ((T >> l) & 1) ? 0xFFFFFFFFu : 0u;
                        lt |= eqm & ~acc[l] & tb;
                        eqm &= ~(acc[l] ^ tb);
                    }
                    le = lt | eqm;
                }
                hits = __popc(le & validmask);
            }
            hits = __reduce_add_sync(0xFFFFFFFFu, hits);
            if ((tid & 31) == 0 && hits) atomicAdd(&s_cnt[pi], hits);
        }
        __syncthreads();  // U (and the raw staging area inside it) free for the next item
    }
    for (int i = tid; i < pcount; i += kSlicedThreads) {
        const uint32_t c = s_cnt[i];
        if (c) atomicAdd(&a.counts[s_pid[i]], (unsigned long long)c);
    }
}

#endif  // __CUDACC__

}  // namespace apm
