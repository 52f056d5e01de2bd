// apm_band.cuh -- exact band mode: the window-sliced DP restricted to the
// Ukkonen band |i - j| <= K, K >= approx_factor.
//
// D[m][m] <= k can only be reached by paths that stay within k diagonals of the main diagonal, so cells with
// |i - j| > K >= k never matter for the threshold decision: treating them as "infinitely expensive" leaves
// every value <= k exact and every other value > k.  In the delta encoding an out-of-band neighbour is the
// constant delta +1 (then its candidate is never strictly better than the diagonal one).  The band has
// 2K+1 cells per row instead of m: (2K+1)*5 + (K+1) LOP3 per row and 32 windows -- m=64, k=4: 100 LOP3 per
// (pattern, window) instead of 640; m=200, k=10: 725 instead of 6250.  Results are bit-identical to the
// full evaluation (tests/test_gpu_parity.py compares band, direct, DP kernel and the oracle).
//
// Layout tricks
//   * band index d = j - i + K: the cell of row i+1 at index d sits under the cell of row i at index d+1,
//     so the row-to-row state (horizontal deltas) is updated in place, hb[d] <- f(hb[d+1]), no shifting.
//   * left of the matrix (j <= 0) the DP boundary D[i][0] = i is continued as D[i][j] = i - j: those cells
//     have horizontal delta -1, vertical delta +1 and reproduce themselves under the recurrence whatever
//     eq says (d0 is forced by b-), so the first K rows need no special case; cells right of the matrix
//     (j > m) never feed a valid cell.  Every row is the same branch-free sequence of 2K+1 cells.
//   * the distance is accumulated along the main diagonal: D[m][m] = #{ i : D[i][i] > D[i-1][i-1] } =
//     number of rows whose diagonal step is not free (~d0 at d = K), kept as a saturating thermometer
//     T[q] = (count > q), one LOP3 per plane and row.
//   * the U rows of this kernel hold 32 + 2K windows (the last 2K repeat the start of the next row), so
//     the 2K+1 match words of a row are one row pointer + immediate offsets.
// State per thread: 2(2K+1) + K+1 registers, for any pattern length up to kSlicedMaxLen.
#pragma once
#include "apm_sliced.cuh"

namespace apm {

constexpr int kBandMaxK = 16;

#ifdef __CUDACC__

// address of the match word of text offset c0 inside a U plane row (c0 may be negative: previous row)
__device__ __forceinline__ const unsigned char *band_row_ptr(const unsigned char *plane_row, int c0, int row_bytes) {
    return plane_row + (c0 >> 5) * row_bytes + (c0 & 31) * 4;
}

template <int K, int CELL>
__global__ void __launch_bounds__(kSlicedThreads, 3) band_count_kernel(const SlicedArgs a) {
    constexpr int BW = 2 * K + 1;
    if (a.run_if && *a.run_if == 0u) return;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const size_t off_U = sliced_smem_fixed(a.nplanes, a.rowsU);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    uint8_t *s_map = smem + 64;
    const uint32_t plane_bytes = (uint32_t)a.rowsU * a.row_bytes;
    const int row_bytes = a.row_bytes;

    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    for (int i = tid; i < 256; i += kSlicedThreads) s_map[i] = a.plane_of[i];
    __syncthreads();

    const long long nwin = a.w1 - a.w0;
    const long long ntiles = (nwin + kSlicedTile - 1) / kSlicedTile;
    uint32_t phase = 0;
    const uint32_t neg1 = a.c_neg1;
    const long long nitems = ntiles * a.nsplits;
    const int per_split = (a.npat + a.nsplits - 1) / a.nsplits;
    __shared__ long long s_item;
    for (;;) {
        if (tid == 0) s_item = (long long)atomicAdd(a.work_counter, 1ull);
        __syncthreads();
        const long long it = s_item;
        if (it >= nitems) break;
        const long long t = it / a.nsplits;
        const int split = (int)(it % a.nsplits);
        const int p_begin = split * per_split, p_end = min(a.npat, p_begin + per_split);
        const TileGeom g = sliced_stage_tile(a, t, smem, bar, phase);

        const long long tile_end = min(g.ts + (long long)kSlicedTile, a.w1);
        const long long jbase = g.ts + 32ll * tid;
        // U row 0 holds the 32 positions before the tile (a.lead = 32): the first K rows of the DP read their
        // match words through the tail of the PREVIOUS U row, which must exist for thread 0 as well
        const unsigned char *urow = smem + off_U + (size_t)(tid + 1) * row_bytes;
        for (int pi = p_begin; pi < p_end; ++pi) {
            const int m = __ldg(a.pat_m + pi);
            const long long lim = min(tile_end, a.n_end - m + 1);  // full windows only
            const long long nvalid = lim - jbase;
            const uint32_t validmask = nvalid >= 32 ? 0xFFFFFFFFu : (nvalid <= 0 ? 0u : ((1u << (int)nvalid) - 1u));
            uint32_t hits = 0;
            if (validmask != 0u) {
                const uint8_t *pc = a.pat_codes + (size_t)pi * a.mcp;
                // row 0: horizontal delta +1 inside the matrix (j >= 1 <=> d > K), -1 in its left continuation
                uint32_t hbp[BW], hbm[BW];
#pragma unroll
                for (int d = 0; d < BW; ++d) {
                    hbp[d] = d > K ? 0xFFFFFFFFu : 0u;
                    hbm[d] = (d > K && CELL != 2) ? 0u : 0xFFFFFFFFu;  // CELL 2: second plane = [h != 0]
                }
                // T[q] = (K - k + number of non-free diagonal steps so far) > q: starting the count at K - k makes
                // the final test "count > k" the compile-time plane T[K] whatever the runtime k <= K is
                uint32_t T[K + 1];
#pragma unroll
                for (int q = 0; q <= K; ++q) T[q] = q < K - a.k ? 0xFFFFFFFFu : 0u;
                uint32_t code_next = __ldg(pc + 1);
                // match words of band index 0..2K of row i: text offset c0 = i - K (negative in the first K
                // rows: left continuation, read through the tail of the previous U row); U rows hold 32 + 2K
                // windows, so the 2K+1 words never change rows: one row pointer + immediate offsets
                uint32_t qn[BW];  // software pipeline: the words of the NEXT row are requested a row ahead
                {
                    const unsigned char *rp = band_row_ptr(urow + (uint32_t)__ldg(pc) * plane_bytes, -K, row_bytes);
#pragma unroll
                    for (int d = 0; d < BW; ++d) qn[d] = *reinterpret_cast<const uint32_t *>(rp + 4 * d);
                }
#pragma unroll 1
                for (int i = 0; i < m; ++i) {  // row i+1 of the DP; band index d <-> text offset c = i + d - K
                    uint32_t qc[BW];
#pragma unroll
                    for (int d = 0; d < BW; ++d) qc[d] = qn[d];
                    {
                        const unsigned char *rp = band_row_ptr(urow + code_next * plane_bytes, i + 1 - K, row_bytes);
#pragma unroll
                        for (int d = 0; d < BW; ++d) qn[d] = *reinterpret_cast<const uint32_t *>(rp + 4 * d);
                    }
                    code_next = __ldg(pc + i + 2);
                    uint32_t ap = 0xFFFFFFFFu, am = 0u;  // left of the band: +1 (boundary or out-of-band)
                    uint32_t z = 0u;
#pragma unroll
                    for (int d = 0; d < BW; ++d) {
                        const uint32_t q = qc[d];
                        const uint32_t bp = d + 1 < BW ? hbp[d + 1] : 0xFFFFFFFFu;  // above the band: +1
                        const uint32_t bm = d + 1 < BW ? hbm[d + 1] : cell_plus_second_plane<CELL>();
                        uint32_t nbp = bp, nbm = bm;
                        if constexpr (CELL == 0) {  // five LOP3; ~d0 of the main diagonal is the cost of the step
                            const uint32_t d0 = lop3<kLutOr3>(q, am, bm);
                            const uint32_t vm = lop3<kLutAndOr>(bp, q, am);
                            const uint32_t vp = lop3<kLutOrNor>(bm, d0, bp);
                            nbp = lop3<kLutOrNor>(am, d0, ap);
                            nbm = lop3<kLutAndOr>(ap, q, bm);
                            ap = vp;
                            am = vm;
                            if (d == K) z = ~d0;
                        } else {  // FMA-pipe variants of the cell (apm_sliced.cuh); x = ~d0 needs its own LOP3
                            if (d == K) z = lop3<kLutNor3>(q, am, cell_minus_plane<CELL>(bp, bm));
                            sliced_cell<CELL>(q, ap, am, nbp, nbm, neg1);
                        }
                        hbp[d] = nbp;
                        hbm[d] = nbm;
                    }
#pragma unroll
                    for (int q = K; q >= 1; --q) T[q] |= T[q - 1] & z;
                    T[0] |= z;
                }
                // D[m][m] <= k  <=>  not (count > k)
                const uint32_t hitmask = ~T[K] & validmask;
                hits = __popc(hitmask);
                if (a.sink.buf && hitmask) hit_emit_mask(a.sink, __ldg(a.pat_id + pi), jbase, hitmask);
            }
            hits = __reduce_add_sync(0xFFFFFFFFu, hits);
            if ((tid & 31) == 0 && hits) atomicAdd(&a.counts[__ldg(a.pat_id + pi)], (unsigned long long)hits);
        }
        __syncthreads();  // U free for the next item
    }
}

#endif  // __CUDACC__

}  // namespace apm
