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
constexpr int kSlicedMaxLen = 1024;               // longest pattern handled (column blocks of 64)
constexpr int kSlicedThreeCtaLen = 224;            // longest pattern whose U table still lets three CTAs share an SM (4 planes)
constexpr int kSlicedTotPlanes = 12;              // bit-planes of the running distance sum (values <= 2*1024+k)
constexpr uint8_t kNoPlane = 0xFF;

// rows of one U plane for patterns up to mmax symbols: one per thread + the rows the halo reaches into
__host__ __device__ inline int sliced_rowsU(int mmax) { return kSlicedTile / 32 + (mmax + 31) / 32 + 1; }
// Shared-memory layout: [barrier | byte->plane map | B bit-vectors | U table].  The raw text tile is staged
// by TMA INSIDE the (not yet built) U region, so it costs no extra shared memory and three CTAs fit an SM
// for the 4-symbol DNA alphabet (76.5 KB each at m <= 224).
__host__ __device__ inline size_t sliced_smem_fixed(int nplanes, int rowsU) {
    return ((size_t)64 + 256 + (size_t)nplanes * (rowsU + 2) * 4 + 15) / 16 * 16;
}
__host__ __device__ inline size_t sliced_smem_bytes(int nplanes, int rowsU, int row_bytes = kURowBytes) {
    return sliced_smem_fixed(nplanes, rowsU) + (size_t)nplanes * rowsU * row_bytes;
}

#ifdef __CUDACC__

struct SlicedArgs {
    const uint8_t *buf;          // device text; buf[0] is global byte buf_offset
    long long buf_len, n_end;    // valid bytes; local index of the global end of text
    long long w0, w1;            // local window-start range
    const uint8_t *pat_codes;    // [npat][mcp] plane index of every pattern symbol (zero padded, mcp >= m + 2)
    const int *pat_m;            // [npat]
    const int *pat_id;           // [npat] index into counts
    const uint8_t *plane_of;     // [256] byte -> plane index or kNoPlane
    unsigned long long *counts;
    uint2 *vscratch;             // [mmax + 2][gridDim.x * kSlicedThreads] boundary deltas between column blocks
    int npat, mcp, nplanes, k;
    int nsplits;                 // pattern ranges per tile (work item = tile x pattern range)
    int rowsU;                   // rows per U plane
    int row_bytes;               // bytes per U row (kURowBytes; wider for the band kernel)
    int row_cols;                // 32-bit windows stored per U row (32; 32 + 2K for the band kernel)
    int lead;                    // text positions staged BEFORE the tile start (0; 32 for the band kernel)
    unsigned long long *work_counter;  // zeroed before the launch: dynamic (tile, range) item dispenser
    HitSink sink;                // optional match-position output
    const unsigned int *run_if;  // optional gate: the whole launch is a no-op when *run_if == 0 (filter fallback)
    uint32_t c_neg1;             // the constant 0xFFFFFFFF, opaque to ptxas (multiplier of the FMA-pipe subtractions)
};

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
constexpr int kLutNor3 = 0x01;       // ~(a | b | c)
constexpr int kLutOrAndN = 0xF4;     // a | (b & ~c)
constexpr int kLutXor3 = 0x96;       // a ^ b ^ c
constexpr int kLutAndOrN = 0xD0;     // a & (b | ~c)
constexpr int kLutNotAAndB = 0x0C;   // ~a & b
constexpr int kLutNotAAndC = 0x0A;   // ~a & c
constexpr int kLutCOrAAndNotB = 0xBA;  // c | (a & ~b)

// c - a on the FMA pipe: IMAD with the multiplier -1 held where ptxas cannot see its value (otherwise the
// subtraction becomes an IADD3 and goes back to the ALU pipe, the one the LOP3s saturate)
__device__ __forceinline__ uint32_t fma_sub(uint32_t c, uint32_t a, uint32_t neg1) {
    uint32_t d;
#ifdef APM_FMA_SUB_PLAIN
    asm("sub.u32 %0, %1, %2;" : "=r"(d) : "r"(c), "r"(a));
#else
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(neg1), "r"(c));
#endif
    return d;
}

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
//   CELL 0: five LOP3 (all on the ALU pipe).
//   CELL 3: CELL 1 with fewer register-file reads (see below); the default for m <= 224.
//   CELL 1: four LOP3 + three IMAD.  With x = ~d0 (the diagonal step D[i][j] - D[i-1][j-1], one bit) the deltas
//           obey h = x - a as integers PER BIT POSITION: h+ - h- = x - a+ + a-.  Every term is 0/1 and so is
//           h+, hence the identity also holds for the whole 32-bit words modulo 2^32 (borrows of intermediate
//           results cancel), and h+ = a- - ((a+ - x) - h-) can be evaluated by three subtractions on the
//           otherwise idle FMA pipe.  The ALU pipe, which bounds the kernel, executes 4 instead of 5
//           instructions per cell; h+ is only consumed one row later, so the longer latency is free.
template <int CELL>
__device__ __forceinline__ void sliced_cell(uint32_t q, uint32_t &ap, uint32_t &am, uint32_t &bp, uint32_t &bm,
                                            uint32_t neg1) {
    if constexpr (CELL == 0) {
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
    } else if constexpr (CELL == 1) {
        const uint32_t x = lop3<kLutNor3>(q, am, bm);
        const uint32_t vm = lop3<kLutAndOr>(bp, q, am);
        const uint32_t vp = lop3<kLutOrAndN>(bm, x, bp);
        const uint32_t hm2 = lop3<kLutAndOr>(ap, q, bm);
        const uint32_t t1 = fma_sub(ap, x, neg1);     // a+ - x
        const uint32_t t2 = fma_sub(t1, hm2, neg1);   // a+ - x - h-
        bp = fma_sub(am, t2, neg1);                   // h+ = a- - a+ + x + h-
        bm = hm2;
        ap = vp;
        am = vm;
    } else if constexpr (CELL == 3) {
        // CELL 3 ("fma3r"): CELL 1 rewritten for the REGISTER FILE, which is what bounds the 4 LOP3 + 3 IMAD cell: it delivers
        // ~2 operands per clock and SMSP (tools/ubench/issue_order.cu), CELL 1 reads 18 (ptxas' code: 18.4 per cell incl. the
        // LDS address, 9.2 clk against 8 clk of ALU pipe).  (i) v- and h- are taken from x: v = -1 iff b = +1 and the diagonal
        // step is 0 (b+ = 1 excludes b- = 1, so x = 0 <=> eq | a-), likewise h- -- two register operands instead of three.
        // (ii) x sits in slot A of the three LOP3s and of the IMAD that consume it, b+ in slot B of v- and v+, a+ in slot C of
        // h- and of that IMAD, so ptxas can flag them .reuse (operand reuse cache) and keeps the cell's instructions
        // together.  ptxas' code reads 13.5 registers per cell (profiles/r02_cell_reuse_*.jsonl: m = 64 114.2 -> 120.5 TCUPS).
        // The row recurrence a- -> a-' is two LOP3 deep here (x, then v-) instead of one; at 3 warps per scheduler the
        // latency is hidden, and unlike the straightforward x-first code (measured: spills, 91 TCUPS) this operand order
        // keeps ptxas from hoisting the whole chain ahead of the row.
        const uint32_t x = lop3<kLutNor3>(q, am, bm);
        const uint32_t vm = lop3<kLutNotAAndB>(x, bp, bp);
        const uint32_t vp = lop3<kLutCOrAAndNotB>(x, bp, bm);
        const uint32_t hm2 = lop3<kLutNotAAndC>(x, ap, ap);
        const uint32_t t1 = fma_sub(ap, x, neg1);     // a+ - x
        const uint32_t t2 = fma_sub(t1, hm2, neg1);   // a+ - x - h-
        bp = fma_sub(am, t2, neg1);                   // h+ = a- - a+ + x + h-
        bm = hm2;
        ap = vp;
        am = vm;
    } else {
        // CELL 2: four LOP3 + two IMAD, on a different encoding of the deltas:
        //   vertical   a: am = [a == -1], ap = [a != 0]        horizontal b: bp = [b == +1], bm = [b != 0]
        // v- needs only (b+, eq, a-), as before.  [v != 0] = s - u with s = [b != 0] | ~(eq | a-) (one LOP3) and
        // u = b+ & x = b+ - v- (x = ~(eq | a- | b-) is the diagonal step; bitwise exact, no borrows).  The
        // deltas obey v - h = a - b, so the parities agree: [h != 0] = [a != 0] ^ [b != 0] ^ [v != 0] (one
        // LOP3), and h+ = [h != 0] & ~[a == +1].  The row recurrence still runs through v- alone (one LOP3 deep);
        // everything behind the two subtractions is only consumed a cell / a row later.
        const uint32_t vm = lop3<kLutAndOr>(bp, q, am);
        const uint32_t s = lop3<kLutOrNor>(bm, q, am);
        const uint32_t u = fma_sub(bp, vm, neg1);
        const uint32_t vnz = fma_sub(s, u, neg1);
        const uint32_t hnz = lop3<kLutXor3>(ap, bm, vnz);
        bp = lop3<kLutAndOrN>(hnz, am, ap);
        bm = hnz;
        ap = vnz;
        am = vm;
    }
}

// second plane of a horizontal delta of +1 (the DP boundary D[0][j] - D[0][j-1], and the neutral value of
// unused columns): h- = 0 for CELL 0/1, [h != 0] = 1 for CELL 2.  The first plane (h+) is all ones, and the
// vertical boundary delta +1 is (ap, am) = (all ones, 0) in every encoding.
template <int CELL>
__device__ __forceinline__ constexpr uint32_t cell_plus_second_plane() { return CELL == 2 ? 0xFFFFFFFFu : 0u; }
// h- plane from the state planes
template <int CELL>
__device__ __forceinline__ uint32_t cell_minus_plane(uint32_t bp, uint32_t bm) { return CELL == 2 ? (bm & ~bp) : bm; }

// Row-major sweep of one column block (MC columns starting at the column the pointer `urow` is positioned
// on) over all `rows` pattern symbols.  urow = this thread's row origin inside plane 0 of the U table,
// pc = plane index of every pattern symbol in global memory (readable up to rows + 1).
// VIN / VOUT: the vertical deltas entering the block on its left edge are read from / those leaving on its
// right edge are written to vs[row * vstride] (a per-thread column of a global scratch array, L2 resident);
// without VIN the left edge is the DP boundary D[i][0] = i (+1).
// FULL (all MC columns used): branch-free and software pipelined -- the match words of the next column
// group (same row, or the first group of the next row) are requested before the current group is computed.
template <int MC, int CELL, bool FULL, bool VIN, bool VOUT>
__device__ __forceinline__ void sliced_sweep(const unsigned char *__restrict__ urow, const uint8_t *__restrict__ pc,
                                             int rows, int width, uint32_t plane_bytes, uint32_t (&hp)[MC],
                                             uint32_t (&hm)[MC], uint2 *__restrict__ vs, long long vstride,
                                             uint32_t neg1) {
    uint32_t code_next = __ldg(pc + 1);
    uint2 vin = make_uint2(0xFFFFFFFFu, 0u);
    if constexpr (VIN) vin = vs[0];
    if constexpr (FULL) {
        constexpr int G = 8;  // columns per pipeline group: 4 LDS.64 in flight
        constexpr int NG = MC / G;
        const unsigned char *e = urow + (uint32_t)__ldg(pc) * plane_bytes;
        uint2 buf[G / 2];
#pragma unroll
        for (int q = 0; q < G / 2; ++q) buf[q] = *reinterpret_cast<const uint2 *>(e + q * 8);
#pragma unroll 1
        for (int i = 0; i < rows; ++i) {
            const unsigned char *e_next = urow + code_next * plane_bytes;
            code_next = __ldg(pc + i + 2);
            uint32_t ap = vin.x, am = vin.y;  // left edge: boundary D[i][0] - D[i-1][0] = +1, or the previous block
            if constexpr (VIN) vin = vs[(long long)(i + 1) * vstride];
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
                    sliced_cell<CELL>((cc & 1) ? cur[cc >> 1].y : cur[cc >> 1].x, ap, am, hp[gi * G + cc], hm[gi * G + cc], neg1);
            }
            if constexpr (VOUT) vs[(long long)i * vstride] = make_uint2(ap, am);
            e = e_next;
        }
    } else {
        constexpr int CH = 8;  // columns per uniform early-exit check
#pragma unroll 1
        for (int i = 0; i < rows; ++i) {
            const unsigned char *e = urow + (uint32_t)__ldg(pc + i) * plane_bytes;
            uint32_t ap = vin.x, am = vin.y;
            if constexpr (VIN) vin = vs[(long long)(i + 1) * vstride];
#pragma unroll
            for (int c = 0; c < MC; c += 2) {
                if ((c % CH) == 0 && c >= width) break;  // uniform: the same for the whole CTA
                const uint2 eq = *reinterpret_cast<const uint2 *>(e + (c >> 5) * kURowBytes + (c & 31) * 4);
                sliced_cell<CELL>(eq.x, ap, am, hp[c], hm[c], neg1);
                sliced_cell<CELL>(eq.y, ap, am, hp[c + 1], hm[c + 1], neg1);
            }
            if constexpr (VOUT) vs[(long long)i * vstride] = make_uint2(ap, am);
        }
#pragma unroll
        for (int j = 0; j < MC; ++j)  // columns >= width: back to the neutral boundary value (they add a constant)
            if (j >= width) { hp[j] = 0xFFFFFFFFu; hm[j] = cell_plus_second_plane<CELL>(); }
    }
}

// Two rows per trip, skewed by SK columns: row i at column c and row i + 1 at column c - SK are independent, so a warp
// carries two recurrence chains instead of one.  The row chain (a- -> a-' of the right neighbour) then stops being the
// longest dependency path of the unrolled row, ptxas keeps the cells in program order, and the cell codes that take
// v- and h- from x (two LOP3 deep, fewer register operands) no longer blow up the live ranges.  rows must be even;
// all MC columns are used (single full block).
template <int MC, int CELL, int SK>
__device__ __forceinline__ void sliced_sweep2(const unsigned char *__restrict__ urow, const uint8_t *__restrict__ pc, int rows,
                                              uint32_t plane_bytes, uint32_t (&hp)[MC], uint32_t (&hm)[MC], uint32_t neg1) {
    constexpr int G = 4;  // columns per pipeline group and row: 2 + 2 LDS.64 in flight
    constexpr int NS = MC + SK;  // steps per row pair
    static_assert(SK % G == 0 && SK >= G, "skew must be a multiple of the group size");
    const unsigned char *eA = urow + (uint32_t)__ldg(pc) * plane_bytes;
    const unsigned char *eB = urow + (uint32_t)__ldg(pc + 1) * plane_bytes;
    uint32_t codeA = __ldg(pc + 2), codeB = __ldg(pc + 3);
    uint2 bufA[G / 2], bufB[G / 2];
#pragma unroll
    for (int q = 0; q < G / 2; ++q) bufA[q] = *reinterpret_cast<const uint2 *>(eA + q * 8);
#pragma unroll 1
    for (int i = 0; i < rows; i += 2) {
        const unsigned char *eA_next = urow + codeA * plane_bytes;
        const unsigned char *eB_next = urow + codeB * plane_bytes;
        codeA = __ldg(pc + i + 4);
        codeB = __ldg(pc + i + 5);
        uint32_t apA = 0xFFFFFFFFu, amA = 0u, apB = 0xFFFFFFFFu, amB = 0u;
#pragma unroll
        for (int s0 = 0; s0 < NS; s0 += G) {
            // current words: row A columns [s0, s0 + G), row B columns [s0 - SK, s0 - SK + G)
            uint2 curA[G / 2], curB[G / 2];
#pragma unroll
            for (int q = 0; q < G / 2; ++q) { curA[q] = bufA[q]; curB[q] = bufB[q]; }
            const int nA = s0 + G, nB = s0 + G - SK;  // first columns of the next group
#pragma unroll
            for (int q = 0; q < G / 2; ++q) {
                if (nA < MC) bufA[q] = *reinterpret_cast<const uint2 *>(eA + ((nA + 2 * q) >> 5) * kURowBytes + ((nA + 2 * q) & 31) * 4);
                else if (nA == NS) bufA[q] = *reinterpret_cast<const uint2 *>(eA_next + q * 8);
                if (nB >= 0 && nB < MC) bufB[q] = *reinterpret_cast<const uint2 *>(eB + ((nB + 2 * q) >> 5) * kURowBytes + ((nB + 2 * q) & 31) * 4);
            }
#pragma unroll
            for (int cc = 0; cc < G; ++cc) {
                const int cA = s0 + cc, cB = s0 + cc - SK;
                if (cA < MC) sliced_cell<CELL>((cc & 1) ? curA[cc >> 1].y : curA[cc >> 1].x, apA, amA, hp[cA], hm[cA], neg1);
                if (cB >= 0 && cB < MC) sliced_cell<CELL>((cc & 1) ? curB[cc >> 1].y : curB[cc >> 1].x, apB, amB, hp[cB], hm[cB], neg1);
            }
        }
        eA = eA_next;
        eB = eB_next;
    }
}

// The same pipelined sweep for a block of which only the first W columns are computed (W a compile-time multiple
// of 4, W <= MC): the ragged last (or only) column block of patterns whose length is not a multiple of MC.  The
// columns [width, W) are evaluated (they are real DP columns) and then put back to the neutral value, the columns
// [W, MC) are never touched.  The run-time-width sweep above (FULL = false) cannot be software pipelined and loses
// 9-19 % on such lengths; with W known at compile time the ragged block runs like a full one.
template <int MC, int CELL, int W, bool VIN, bool VOUT>
__device__ __forceinline__ void sliced_sweep_w(const unsigned char *__restrict__ urow, const uint8_t *__restrict__ pc,
                                               int rows, int width, uint32_t plane_bytes, uint32_t (&hp)[MC],
                                               uint32_t (&hm)[MC], uint2 *__restrict__ vs, long long vstride,
                                               uint32_t neg1) {
    static_assert(W % 4 == 0 && W >= 4 && W <= MC, "W must be a multiple of 4 in [4, MC]");
    constexpr int G = 8;                    // columns per pipeline group (the last group may hold 4)
    constexpr int NG = (W + G - 1) / G;
    constexpr int G0 = W < G ? W : G;       // size of the first group of a row
    uint32_t code_next = __ldg(pc + 1);
    uint2 vin = make_uint2(0xFFFFFFFFu, 0u);
    if constexpr (VIN) vin = vs[0];
    const unsigned char *e = urow + (uint32_t)__ldg(pc) * plane_bytes;
    uint2 buf[G / 2];
#pragma unroll
    for (int q = 0; q < G0 / 2; ++q) buf[q] = *reinterpret_cast<const uint2 *>(e + q * 8);
#pragma unroll 1
    for (int i = 0; i < rows; ++i) {
        const unsigned char *e_next = urow + code_next * plane_bytes;
        code_next = __ldg(pc + i + 2);
        uint32_t ap = vin.x, am = vin.y;
        if constexpr (VIN) vin = vs[(long long)(i + 1) * vstride];
#pragma unroll
        for (int gi = 0; gi < NG; ++gi) {
            const int gs = (W - gi * G) < G ? (W - gi * G) : G;  // compile-time after unrolling
            uint2 cur[G / 2];
#pragma unroll
            for (int q = 0; q < G / 2; ++q) cur[q] = buf[q];
            const int c1 = (gi + 1) * G;
            if (gi + 1 < NG) {
                const int ngs = (W - c1) < G ? (W - c1) : G;
#pragma unroll
                for (int q = 0; q < G / 2; ++q)
                    if (q < ngs / 2) buf[q] = *reinterpret_cast<const uint2 *>(e + ((c1 + 2 * q) >> 5) * kURowBytes + ((c1 + 2 * q) & 31) * 4);
            } else {
#pragma unroll
                for (int q = 0; q < G0 / 2; ++q) buf[q] = *reinterpret_cast<const uint2 *>(e_next + q * 8);
            }
#pragma unroll
            for (int cc = 0; cc < G; ++cc)
                if (cc < gs) sliced_cell<CELL>((cc & 1) ? cur[cc >> 1].y : cur[cc >> 1].x, ap, am, hp[gi * G + cc], hm[gi * G + cc], neg1);
        }
        if constexpr (VOUT) vs[(long long)i * vstride] = make_uint2(ap, am);
        e = e_next;
    }
#pragma unroll
    for (int j = 0; j < W; ++j)  // columns >= width: back to the neutral boundary value (they add a constant)
        if (j >= width) { hp[j] = 0xFFFFFFFFu; hm[j] = cell_plus_second_plane<CELL>(); }
}

// run-time dispatch over the instantiated widths: W = width rounded up to a multiple of STEP
template <int MC, int CELL, int STEP, bool VIN, int W = STEP>
__device__ __forceinline__ void sliced_sweep_dispatch(int wr, const unsigned char *__restrict__ urow, const uint8_t *__restrict__ pc,
                                                      int rows, int width, uint32_t plane_bytes, uint32_t (&hp)[MC],
                                                      uint32_t (&hm)[MC], uint2 *__restrict__ vs, long long vstride, uint32_t neg1) {
    if constexpr (W >= MC) {
        sliced_sweep_w<MC, CELL, MC, VIN, false>(urow, pc, rows, width, plane_bytes, hp, hm, vs, vstride, neg1);
    } else {
        if (wr <= W) sliced_sweep_w<MC, CELL, W, VIN, false>(urow, pc, rows, width, plane_bytes, hp, hm, vs, vstride, neg1);
        else sliced_sweep_dispatch<MC, CELL, STEP, VIN, W + STEP>(wr, urow, pc, rows, width, plane_bytes, hp, hm, vs, vstride, neg1);
    }
}

// tot += acc (bit-sliced): acc has NA planes, tot kSlicedTotPlanes
template <int NA>
__device__ __forceinline__ void planes_add(uint32_t (&tot)[kSlicedTotPlanes], const uint32_t (&acc)[NA]) {
    uint32_t carry = 0u;
#pragma unroll
    for (int l = 0; l < kSlicedTotPlanes; ++l) {
        if (l < NA) carry = plane_fa(tot[l], acc[l], carry);
        else {
            const uint32_t t = tot[l];
            tot[l] = t ^ carry;
            carry = t & carry;
        }
    }
}

// Stage text tile t: TMA bulk copy of the raw bytes into the (not yet built) U region, occurrence
// bit-vectors by warp ballots, then the U table.  Returns the tile geometry; ends with a __syncthreads().
__device__ __forceinline__ TileGeom sliced_stage_tile(const SlicedArgs &a, long long t, unsigned char *smem,
                                                      uint64_t *bar, uint32_t &phase) {
    const int tid = threadIdx.x;
    const int rowsU = a.rowsU;
    const int nB = rowsU + 2;  // words per occurrence bit-vector
    const int span = nB * 32;  // text positions covered by the bit-vectors of a tile
    const size_t off_U = sliced_smem_fixed(a.nplanes, rowsU);
    const uint8_t *s_map = smem + 64;
    uint32_t *s_B = reinterpret_cast<uint32_t *>(smem + 64 + 256);
    uint8_t *s_raw = smem + off_U;  // raw text tile, overwritten by the U table once the bit-vectors exist
    const uint32_t plane_bytes = (uint32_t)rowsU * a.row_bytes;

    // the staged range starts `lead` positions before the tile (positions before the buffer hold no symbol)
    TileGeom g = tile_geometry(a.buf, a.buf_len, a.w0 - a.lead, t, kSlicedTile, span - kSlicedTile);
    if (tid == 0) tile_issue(g, a.buf, s_raw, bar);  // TMA bulk copy of the raw tile
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
            if (i >= 0 && i < g.te) c = s_map[(i >= g.ta && i < g.tb) ? s_raw[i - g.a0] : a.buf[i]];
            for (int p = 0; p < a.nplanes; ++p) {
                const uint32_t bits = __ballot_sync(0xFFFFFFFFu, c == (uint32_t)p);
                if (lane == 0) s_B[p * nB + w] = bits;
            }
        }
    }
    __syncthreads();
    // ---- U table: U[p][w][s] = bits [32 w + s, +32) of B[p], s < row_cols (columns >= 32 repeat the start
    //      of the next row, so a run of row_cols - 32 + 1 consecutive offsets never has to change rows)
    for (int idx = tid; idx < a.nplanes * rowsU; idx += kSlicedThreads) {
        const int p = idx / rowsU, w = idx - p * rowsU;
        const uint32_t b0 = s_B[p * nB + w], b1 = s_B[p * nB + w + 1], b2 = s_B[p * nB + w + 2];
        uint32_t *dst = reinterpret_cast<uint32_t *>(smem + off_U + (size_t)p * plane_bytes + (size_t)w * a.row_bytes);
#pragma unroll
        for (int s = 0; s < 32; ++s) dst[s] = __funnelshift_r(b0, b1, s);
        for (int s = 32; s < a.row_cols; ++s) dst[s] = __funnelshift_r(b1, b2, s - 32);
    }
    __syncthreads();
    g.ts += a.lead;  // report the tile start proper
    return g;
}

// ------------------------------------------------------------------------------------------------
// Persistent count kernel.  Work item = (text tile, pattern range): the tile's U table is built once per
// item and every pattern of the range is swept over it; pattern symbols are streamed from global memory
// (uniform addresses, L1 broadcast).  MC = columns per register block: 32 (patterns m <= 32, one block)
// or 64 (any m <= kSlicedMaxLen, ceil(m/64) blocks chained through the global boundary scratch).
// ------------------------------------------------------------------------------------------------
// RG = 0: any pattern lengths, ragged blocks through the run-time-width sweep (the round-1 kernel, unchanged).
// RG = 1: single-block patterns (m <= MC) swept with the compile-time width m rounded up to a multiple of 4
//         (m = 50: 93.8 -> 103.3 TCUPS, m = 20: 76.6 -> 99.2; profiles/r02_len_sweep.txt).
// RG = 3: lists with m == MC, two skewed rows per trip (sliced_sweep2); used for m = 32 (110.6 -> 114.9 TCUPS with CELL 3;
//         at m = 64 the one-row sweep is faster: 120.5 vs 116.3).
// The host keeps m = 64 and all multi-block patterns on RG = 0 and routes single-block ragged lengths to RG = 1, m = 32
// to RG = 3, when the cell code is chosen automatically.  (A variant with compile-time widths for the LAST block of multi-block patterns was built
// and measured 3-10 % SLOWER than the run-time-width sweep for every m > 64 -- the extra sweep instantiations cost
// the multi-block kernel its register allocation -- and was dropped.)
template <int MC, int CELL, int RG = 0>
__global__ void __launch_bounds__(kSlicedThreads, MC == 64 ? 3 : 4) sliced_count_kernel(const SlicedArgs a) {
    static_assert(MC == 32 || MC == 64, "MC must be 32 or 64");
    if (a.run_if && *a.run_if == 0u) return;
    constexpr int LOG = MC == 32 ? 6 : 7;  // 2*MC planes are summed per block: value range [0, 2*MC]
    constexpr int NL = LOG + 1;
    extern __shared__ __align__(128) unsigned char smem[];

    const int tid = threadIdx.x;
    const int rowsU = a.rowsU;
    const size_t off_U = sliced_smem_fixed(a.nplanes, rowsU);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    uint8_t *s_map = smem + 64;
    const uint32_t plane_bytes = (uint32_t)rowsU * kURowBytes;

    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    for (int i = tid; i < 256; i += kSlicedThreads) s_map[i] = a.plane_of[i];
    __syncthreads();

    const long long nwin = a.w1 - a.w0;
    const long long ntiles = (nwin + kSlicedTile - 1) / kSlicedTile;
    const long long vstride = (long long)gridDim.x * kSlicedThreads;
    uint2 *vs = a.vscratch + ((long long)blockIdx.x * kSlicedThreads + tid);
    uint32_t phase = 0;
    const uint32_t neg1 = a.c_neg1;

    const long long nitems = ntiles * a.nsplits;
    const int per_split = (a.npat + a.nsplits - 1) / a.nsplits;
    __shared__ long long s_item;
    for (;;) {
        // dynamic scheduling: items are handed out in order by one global atomic counter, so a CTA that
        // finishes early (short patterns, ragged tiles) simply takes the next item
        if (tid == 0) s_item = (long long)atomicAdd(a.work_counter, 1ull);
        __syncthreads();
        const long long it = s_item;
        if (it >= nitems) break;
        const long long t = it / a.nsplits;
        const int split = (int)(it % a.nsplits);
        const int p_begin = split * per_split, p_end = min(a.npat, p_begin + per_split);
        const TileGeom g = sliced_stage_tile(a, t, smem, bar, phase);

        // ---- hot loop: patterns x column blocks x rows x columns, 5 LOP3 per cell
        const long long tile_end = min(g.ts + (long long)kSlicedTile, a.w1);
        const long long jbase = g.ts + 32ll * tid;
        const unsigned char *urow = smem + off_U + (size_t)tid * kURowBytes;
        for (int pi = p_begin; pi < p_end; ++pi) {
            const int m = __ldg(a.pat_m + pi);
            const long long lim = min(tile_end, a.n_end - m + 1);  // full windows only
            const long long nvalid = lim - jbase;
            const uint32_t validmask = nvalid >= 32 ? 0xFFFFFFFFu : (nvalid <= 0 ? 0u : ((1u << (int)nvalid) - 1u));
            uint32_t hits = 0;
            if (validmask != 0u) {
                const uint8_t *pc = a.pat_codes + (size_t)pi * a.mcp;
                const int nblk = (m + MC - 1) / MC;
                uint32_t tot[kSlicedTotPlanes];
#pragma unroll
                for (int l = 0; l < kSlicedTotPlanes; ++l) tot[l] = 0u;
#pragma unroll 1
                for (int b = 0; b < nblk; ++b) {
                    const int width = min(MC, m - b * MC);
                    const unsigned char *ub = urow + (size_t)b * (MC / 32) * kURowBytes;  // MC columns = MC/32 U rows
                    uint32_t hp[MC], hm[MC];
#pragma unroll
                    for (int j = 0; j < MC; ++j) { hp[j] = 0xFFFFFFFFu; hm[j] = cell_plus_second_plane<CELL>(); }  // D[0][j] - D[0][j-1] = +1
                    if constexpr (RG == 3) {
                        // two skewed rows per trip; the host routes only lists with m == MC (even) here
                        sliced_sweep2<MC, CELL, 4>(ub, pc, m, plane_bytes, hp, hm, neg1);
                    } else if constexpr (RG == 1) {
                        sliced_sweep_dispatch<MC, CELL, 4, false, (MC == 64 ? 36 : 4)>((width + 3) & ~3, ub, pc, m, width, plane_bytes, hp, hm, vs, vstride, neg1);
                    } else if (MC == 32 || nblk == 1) {
                        if (width == MC) sliced_sweep<MC, CELL, true, false, false>(ub, pc, m, width, plane_bytes, hp, hm, vs, vstride, neg1);
                        else sliced_sweep<MC, CELL, false, false, false>(ub, pc, m, width, plane_bytes, hp, hm, vs, vstride, neg1);
                    } else if (b == 0) {
                        sliced_sweep<MC, CELL, true, false, true>(ub, pc, m, width, plane_bytes, hp, hm, vs, vstride, neg1);
                    } else if (b < nblk - 1) {
                        sliced_sweep<MC, CELL, true, true, true>(ub, pc, m, width, plane_bytes, hp, hm, vs, vstride, neg1);
                    } else {
                        if (width == MC) sliced_sweep<MC, CELL, true, true, false>(ub, pc, m, width, plane_bytes, hp, hm, vs, vstride, neg1);
                        else sliced_sweep<MC, CELL, false, true, false>(ub, pc, m, width, plane_bytes, hp, hm, vs, vstride, neg1);
                    }
                    // block sum: sum_j (h+[j] + ~h-[j]) over the MC columns of the last row (unused ones add 2)
                    uint32_t acc[NL], pend[NL];
#pragma unroll
                    for (int l = 0; l < NL; ++l) { acc[l] = 0u; pend[l] = 0u; }
                    if constexpr (CELL == 2) {
#pragma unroll
                        for (int j = 0; j < MC; ++j) hm[j] = cell_minus_plane<CELL>(hp[j], hm[j]);
                    }
                    SumPlanes<MC, 0, NL>::run(hp, hm, acc, pend);
                    acc[LOG] = pend[LOG];
                    planes_add<NL>(tot, acc);
                }
                // total V = D[m][m] + 2 * (nblk * MC - m); windows with V <= T
                // (clamped: V < 2^kSlicedTotPlanes - 1 always, so a huge k simply matches everything)
                const int T = min(a.k + 2 * (nblk * MC - m), (1 << kSlicedTotPlanes) - 1);
                uint32_t lt = 0u, eqm = 0xFFFFFFFFu;
#pragma unroll
                for (int l = kSlicedTotPlanes - 1; l >= 0; --l) {
                    const uint32_t tb = ((T >> l) & 1) ? 0xFFFFFFFFu : 0u;
                    lt |= eqm & ~tot[l] & tb;
                    eqm &= ~(tot[l] ^ tb);
                }
                const uint32_t hitmask = (lt | eqm) & validmask;
                hits = __popc(hitmask);
                if (a.sink.buf && hitmask) hit_emit_mask(a.sink, __ldg(a.pat_id + pi), jbase, hitmask);
            }
            hits = __reduce_add_sync(0xFFFFFFFFu, hits);
            if ((tid & 31) == 0 && hits) atomicAdd(&a.counts[__ldg(a.pat_id + pi)], (unsigned long long)hits);
        }
        __syncthreads();  // U (and the raw staging area inside it) free for the next item
    }
}

#endif  // __CUDACC__

}  // namespace apm
