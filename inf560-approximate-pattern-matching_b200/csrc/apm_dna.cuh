// apm_dna.cuh -- exact filter mode for DNA pattern sets: 2-bit q-gram scan on a
// strided sampling grid + verification of SEED HITS by outward banded DPs.  Bit-identical counts.
//
// Pigeonhole (as apm_filter.cuh): cut a pattern of length m into k+1 pieces (offsets o_i = floor(i m / (k+1)),
// lengths l_i >= l_min).  levenshtein(P, W) <= k for a window W = T[j, j+m) forces an alignment with <= k edits in
// which some piece i is copied verbatim to T[tp, tp + l_i), tp = j + o_i + d, |d| <= k, and whose cost is
//     ed(P[0, o_i), T[j, tp))  +  ed(P[o_i + l_i, m), T[tp + l_i, j + m))  <=  k                      (*)
// Call (i, d) a VALID WITNESS of window j when the piece occurs there and (*) holds.  Every matching window has at
// least one valid witness; it is counted exactly once, by its lexicographically smallest one.
//
// Scan (dna_scan_kernel<H>): every text byte is reduced to 2 bits, code = (byte >> 1) & 3 -- injective on
// {A,C,G,T}; any other byte aliases to one of the four codes, which can only ADD false seed hits (they die in the
// byte-exact piece compare), never lose one.  A lane packs 16 bytes into one 32-bit word with 4 LOP3 + 4 IMAD +
// 3 PRMT, the warp's packed stream lives in shared memory.  Only every H-th text position is probed: each piece
// is indexed at its first H offsets (H <= l_min - q + 1), so exactly one probe of the grid falls on an indexed
// q-gram of any verbatim piece occurrence.  First level: a DIRECT bitmap over all 4^q q-grams (q <= 10: 128 KB
// of shared memory), one LDS per probe.  Second level (the nent H / 4^q of the probes that pass, compacted per
// warp so that all lanes work): CSR look-up of the q-gram's entries in L2, 2-bit compare of the whole piece
// against the packed stream, byte-exact compare, then one 64-bit SEED HIT (slot, piece, tp) is appended to the
// CTA's private segment of the candidate buffer (no global atomics).
//
// Verify, three small kernels behind the scan (stream ordered, no host round trip):
//   dna_verify_kernel       one THREAD per seed hit: a lower bound of both sides of (*) from one 32-row word of
//                           Hyyro/Myers bit-vector DP each (the 32 pattern symbols next to the seed against the
//                           32 + k text symbols next to the seed occurrence, 2-bit codes, Peq precomputed per
//                           piece).  Fixed work, no divergence; random seed hits (the vast majority) stop here.
//   dna_verify_long_kernel  one WARP per survivor: the exact last rows of both DPs (lane = 32-row word, carries
//                           across lanes by ballot carry-lookahead as in apm_tail.cuh), giving the cost (*) of all
//                           2k+1 shifts at once; every window within k is inserted into a hash SET keyed by
//                           (pattern, window) -- a window witnessed several times is one key.
//   dna_collect_kernel      counts[pattern] += 1 (and the optional position output) for every key of the set.
// Every matching window has a valid witness, every valid witness is a seed hit the scan emits, so the set holds
// exactly the matching windows.
//
// If a candidate segment, the survivor queue or the set overflows (low-complexity text) a flag turns the remaining
// kernels into no-ops and switches on the band kernel for the same patterns and windows: always exact.
#pragma once
#include "apm_common.cuh"

namespace apm {

constexpr int kDnaThreads = 1024;   // one CTA per SM: the q-gram bitmap takes up to 128 KB of its shared memory
constexpr int kDnaWarps = kDnaThreads / 32;
constexpr int kDnaQMax = 10;        // 4^10 bits = 128 KB
constexpr int kDnaQMin = 8;
constexpr int kDnaHMax = 4;         // probe stride (bytes per thread and iteration = 32 H)
constexpr int kDnaQueue = 192;      // second-level queue entries per warp (the rest is probed inline)
constexpr int kDnaBloomLog = 18;    // second shared-memory bitmap: hashed (q + H - 1)-grams of the piece starts, 32 KB
constexpr uint32_t kDnaBloomMul = 0x9E3779B1u;
constexpr int kDnaMaxK = 15;        // 2k+1 shifts <-> lanes of a warp

__host__ __device__ __forceinline__ uint32_t dna_code(uint8_t c) { return (c >> 1) & 3u; }
// bit index of a packed (q + H - 1)-gram in the second bitmap
__host__ __device__ __forceinline__ uint32_t dna_bloom_index(uint32_t gram) { return (gram * kDnaBloomMul) >> (32 - kDnaBloomLog); }
__host__ __device__ __forceinline__ int dna_piece_offset(int i, int m, int k) { return (int)((long long)i * m / (k + 1)); }
// seed hit: slot (24) | piece (5) | tp - w0 (35)
constexpr int kDnaPosBits = 35;
__host__ __device__ __forceinline__ uint64_t dna_pack(uint32_t slot, int piece, unsigned long long rel) {
    return ((uint64_t)slot << 40) | ((uint64_t)piece << kDnaPosBits) | rel;
}

// words of one warp's packed-stream slab: [look-behind word | 64 H words | look-ahead word] (+1 pad)
__host__ __device__ constexpr int dna_slab_words(int H) { return 64 * H + 3; }
__host__ __device__ inline size_t dna_smem_bytes(int table_words, int H) {
    return (size_t)table_words * 4 + ((size_t)1 << (kDnaBloomLog - 3)) + (size_t)kDnaWarps * (dna_slab_words(H) * 4 + kDnaQueue * 2) + 16;
}

#ifdef __CUDACC__

struct DnaArgs {
    const uint8_t *buf;            // device text; buf[0] is global byte buf_offset
    const uint32_t *packed;        // optional resident 2-bit copy of buf (dna_pack_kernel): word w = symbols [16 w, 16 w + 16)
    long long buf_len, n_end;      // valid bytes; local index of the global end of text
    long long w0, w1;              // local window-start range of this round
    int q, k, mmax, table_words;
    uint32_t keymask;              // 4^q - 1
    uint32_t amask;                // byte address of a bitmap word from (window >> 3): ((4^q - 1) >> 3) & ~3
    uint32_t gmask;                // 4^(q + H - 1) - 1 (all ones when q + H - 1 = 16)
    int h;                         // probe stride H
    const uint32_t *bloom;         // 2^kDnaBloomLog bits: hashed first q + H - 1 symbols of every piece
    const uint32_t *table;         // 4^q bits: q-grams of the first H offsets of every piece
    const uint32_t *first;         // [4^q + 1] CSR over the entries, by q-gram
    const uint4 *ent;              // {slot << 8 | piece << 3 | off, first 16 symbols packed, o | len << 16, m}
    const uint4 *peq32;            // [(slot * (k+1) + piece) * 2 + side] match masks by 2-bit code of the 32 pattern
                                   //   symbols before (side 0, reversed) / behind (side 1) the piece
    const int *fp_id;              // [nfp] index into counts
    const int *fp_m;               // [nfp]
    const long long *fp_off;       // [nfp] offset of the pattern in pat_bytes
    const uint8_t *pat_bytes;
    uint64_t *cand;                // [nseg][seg_cap] seed hits, one segment per scan CTA
    unsigned int seg_cap;
    unsigned int *seg_count;       // [nseg] written by the scan
    int nseg;
    unsigned int *overflow;        // zeroed before the scan; set when a segment, the survivor queue or the set is full
    uint64_t *longq;               // survivors of the lower-bound test
    unsigned long long longcap;
    unsigned long long *nlong;     // zeroed before the scan
    uint64_t *set;                 // hash set of matching (slot, window) keys, zeroed before the scan
    int set_log;                   // log2 of its slots
    unsigned long long *nset;      // zeroed before the scan
    unsigned long long *counts;
    HitSink sink;
};

// 16 text bytes -> 32 bits, symbol p of the chunk in bits [2p, 2p+2)
__device__ __forceinline__ uint32_t dna_pack16(uint4 v) {
    // x = bits 1..2 of every byte; x * (2^23 + 2^17 + 2^11 + 2^5) gathers the four codes in the top byte (the
    // partial products occupy disjoint bits, nothing carries into bits 24..31)
    const uint32_t y0 = (v.x & 0x06060606u) * 0x00820820u;
    const uint32_t y1 = (v.y & 0x06060606u) * 0x00820820u;
    const uint32_t y2 = (v.z & 0x06060606u) * 0x00820820u;
    const uint32_t y3 = (v.w & 0x06060606u) * 0x00820820u;
    return __byte_perm(__byte_perm(y0, y1, 0x0073), __byte_perm(y2, y3, 0x0073), 0x5410);
}

__device__ __noinline__ uint4 dna_load16(const DnaArgs &a, long long pos) {
    if (pos >= 0 && pos + 16 <= a.buf_len) return __ldg(reinterpret_cast<const uint4 *>(a.buf + pos));
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (pos + 16 > 0 && pos < a.buf_len) {
        for (int i = 0; i < 16; ++i) {
            const long long p = pos + i;
            if (p >= 0 && p < a.buf_len) w[i >> 2] |= (uint32_t)a.buf[p] << (8 * (i & 3));
        }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// second level: the probe at symbol `ts` of the warp's region (global text position base + ts) passed the bitmap.
// (a) shared memory only: an indexed q-gram at piece offset `off` < H means the piece starts at ts - off, and its
//     first q + H - 1 symbols are known -- their hash must be set in the second bitmap for some off (cuts the L2
//     look-ups ~4x at 4096 patterns).  (b) CSR look-up of the q-gram's entries in L2 and 2-bit compare of the
//     piece's first 16 symbols against the packed stream.  Survivors are SEED HITS; whether the text bytes are
//     really A/C/G/T (and symbols 17.. of long pieces) is checked by the byte-exact stage-2 verification.
template <int H>
__device__ __forceinline__ void dna_probe(const DnaArgs &a, const uint32_t *slab, const uint32_t *s_bloom, long long base,
                                          int ts, unsigned int *s_count) {
    bool any = false;
#pragma unroll
    for (int off = 0; off < H; ++off) {
        const int tps = ts - off + 16;  // + 16: slab index 0 is the look-behind word
        const uint32_t gram = __funnelshift_r(slab[tps >> 4], slab[(tps >> 4) + 1], 2 * (tps & 15)) & a.gmask;
        const uint32_t bi = dna_bloom_index(gram);
        any |= (s_bloom[bi >> 5] >> (bi & 31)) & 1u;
    }
    if (!any) return;
    const uint32_t key = __funnelshift_r(slab[(ts >> 4) + 1], slab[(ts >> 4) + 2], 2 * (ts & 15)) & a.keymask;
    const uint32_t lo = __ldg(a.first + key), hi = __ldg(a.first + key + 1);
    for (uint32_t e = lo; e < hi; ++e) {
        const uint4 en = __ldg(a.ent + e);
        const int off = (int)(en.x & 7u);
        const int tps = ts - off + 16;
        const uint32_t tw = __funnelshift_r(slab[tps >> 4], slab[(tps >> 4) + 1], 2 * (tps & 15));
        const int len = (int)(en.z >> 16);
        const uint32_t mask = len >= 16 ? 0xFFFFFFFFu : ((1u << (2 * len)) - 1u);
        if ((tw ^ en.y) & mask) continue;
        const long long tp = base + ts - off;
        const int o = (int)(en.z & 0xFFFFu), m = (int)en.w;
        if (tp < 0 || tp + len > min(a.n_end, a.buf_len)) continue;
        // windows this seed hit can witness: j = tp - o - d, |d| <= k, inside the round, full length, j <= tp
        const long long jlo = max(max(tp - o - a.k, a.w0), 0ll), jhi = min(min(tp - o + a.k, a.w1 - 1), min(a.n_end - m, tp));
        if (jlo > jhi) continue;
        const unsigned int pos = atomicAdd(s_count, 1u);
        if (pos < a.seg_cap) a.cand[(size_t)blockIdx.x * a.seg_cap + pos] = dna_pack(en.x >> 8, (int)((en.x >> 3) & 31u), (unsigned long long)(tp - a.w0));
        else *a.overflow = 1u;
    }
}

// PACKED: the text also exists as a resident 2-bit copy (apm_text_pack_device: searched repeatedly, packed once); the scan
// then streams a quarter of the bytes and skips the packing -- 4-byte loads of ready-made slab words.  Only the first and
// the last region of the text still pack from the raw bytes (their fringes need the bounds-checked loads).
template <int H, bool PACKED = false>
__global__ void __launch_bounds__(kDnaThreads, 1) dna_scan_kernel(const __grid_constant__ DnaArgs a) {
    extern __shared__ __align__(16) uint32_t s_mem[];
    __shared__ unsigned int s_count;
    constexpr int SLABW = dna_slab_words(H);
    constexpr int REGION = 1024 * H;  // text positions per warp and iteration
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int BLOOMW = 1 << (kDnaBloomLog - 5);
    uint32_t *s_table = s_mem;
    uint32_t *s_bloom = s_mem + a.table_words;
    uint32_t *slab = s_bloom + BLOOMW + warp * SLABW;
    uint16_t *queue = reinterpret_cast<uint16_t *>(s_bloom + BLOOMW + kDnaWarps * SLABW) + warp * kDnaQueue;
    if (threadIdx.x == 0) s_count = 0u;
    for (int i = threadIdx.x; i < a.table_words / 4; i += kDnaThreads)
        reinterpret_cast<uint4 *>(s_table)[i] = __ldg(reinterpret_cast<const uint4 *>(a.table) + i);
    for (int i = threadIdx.x; i < BLOOMW / 4; i += kDnaThreads)
        reinterpret_cast<uint4 *>(s_bloom)[i] = __ldg(reinterpret_cast<const uint4 *>(a.bloom) + i);
    __syncthreads();

    // probes: every H-th position of [t_begin, t_end); a q-gram must end inside the text
    const long long t_begin = a.w0;
    const long long t_end = min(a.w1 - 1 + a.mmax, min(a.n_end, a.buf_len)) - a.q + 1;  // exclusive
    if (t_end > t_begin) {
        const long long mis = (long long)(reinterpret_cast<uintptr_t>(a.buf + t_begin) & 15);
        const long long base0 = t_begin - mis;  // 16-byte aligned address; origin of the probe grid
        constexpr long long kTile = (long long)kDnaWarps * REGION;
        const unsigned char *tab = reinterpret_cast<const unsigned char *>(s_table);
        const uint32_t amask = a.amask;
        for (long long tile = base0 + (long long)blockIdx.x * kTile; tile < t_end; tile += (long long)gridDim.x * kTile) {
            const long long base = tile + (long long)warp * REGION;  // first text position of this warp's region
            if (base >= t_end) continue;                             // warp-uniform
            if constexpr (PACKED) {  // the region this warp reads in the NEXT iteration: REGION / 4 bytes of packed words
                const long long nb = base + (long long)gridDim.x * kTile;
                if (lane < REGION / 512 && nb + REGION <= a.buf_len)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const unsigned char *>(a.packed + (nb >> 4)) + 128 * lane));
            } else {  // one 128-byte line per lane towards L2
                const long long nb = base + (long long)gridDim.x * kTile + 128ll * lane;
                if (lane < REGION / 128 && nb + 128 <= a.buf_len) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.buf + nb));
            }
            // ---- load + pack: lane l takes the 16-byte chunks l, l + 32, ... (coalesced), one packed word each
            uint4 raw[2 * H];
            if (PACKED && base >= 16 && base + REGION + 32 <= a.buf_len) {  // warp-uniform; base is a multiple of 16
                const uint32_t *src = a.packed + (base >> 4) + lane;
                uint32_t pw[2 * H];
#pragma unroll
                for (int c = 0; c < 2 * H; ++c) pw[c] = __ldg(src + 32 * c);
#pragma unroll
                for (int c = 0; c < 2 * H; ++c) slab[1 + lane + 32 * c] = pw[c];
                if (lane < 2) slab[lane ? 1 + 64 * H : 0] = __ldg(a.packed + (base >> 4) + (lane ? 64 * H : -1));
            } else if (base >= 16 && base + REGION + 32 <= a.buf_len) {  // warp-uniform: the region and both fringes exist
                const uint4 *src = reinterpret_cast<const uint4 *>(a.buf + base) + lane;
#pragma unroll
                for (int c = 0; c < 2 * H; ++c) raw[c] = __ldg(src + 32 * c);
#pragma unroll
                for (int c = 0; c < 2 * H; ++c) slab[1 + lane + 32 * c] = dna_pack16(raw[c]);
                if (lane < 2) slab[lane ? 1 + 64 * H : 0] = dna_pack16(__ldg(src + (lane ? 64 * H - 1 : -1)));
            } else {
#pragma unroll 1
                for (int c = 0; c < 2 * H; ++c) slab[1 + lane + 32 * c] = dna_pack16(dna_load16(a, base + 16ll * (lane + 32 * c)));
                if (lane == 0) slab[0] = dna_pack16(dna_load16(a, base - 16));
                if (lane == 1) slab[1 + 64 * H] = dna_pack16(dna_load16(a, base + 16ll * 64 * H));
            }
            __syncwarp();
            // ---- first level: thread owns positions [32 H lane, +32 H) of the region, probes every H-th
            uint32_t W[2 * H + 1];
#pragma unroll
            for (int i = 0; i <= 2 * H; ++i) W[i] = slab[1 + 2 * H * lane + i];
            uint32_t maybe = 0u;
#pragma unroll
            for (int s = 0; s < 32; ++s) {
                const int bit = 2 * H * s;
                const uint32_t win = (bit & 31) ? __funnelshift_r(W[bit >> 5], W[(bit >> 5) + 1], bit & 31) : W[bit >> 5];
                const uint32_t word = *reinterpret_cast<const uint32_t *>(tab + ((win >> 3) & amask));
                maybe = __funnelshift_r(maybe, __funnelshift_r(word, 0u, win), 1);
            }
            // probes outside [t_begin, t_end) do not count (only the first and the last regions have any)
            if (base < t_begin || base + REGION > t_end) {
                const long long p0 = base + 32ll * H * lane;
                const long long lo = t_begin - p0, hi = t_end - p0;  // positions; probe s sits at H s
                if (lo > 0) {
                    const long long s0 = (lo + H - 1) / H;
                    maybe &= s0 >= 32 ? 0u : ~((1u << (int)s0) - 1u);
                }
                if (hi < 32 * H) {
                    const long long s1 = hi <= 0 ? 0 : (hi + H - 1) / H;
                    maybe &= s1 >= 32 ? 0xFFFFFFFFu : ((1u << (int)s1) - 1u);
                }
            }
            // ---- second level: compact the warp's hits into its queue (exclusive prefix sum of the hit counts),
            //      then all lanes work through the queue
            const int mine = __popc(maybe);
            int incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += y;
            }
            const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            if (total) {  // warp-uniform
                int slot = incl - mine;
                uint32_t rest = maybe;
                while (rest && slot < kDnaQueue) {
                    const int s = __ffs(rest) - 1;
                    rest &= rest - 1u;
                    queue[slot++] = (uint16_t)(32 * H * lane + H * s);
                }
                __syncwarp();
                const int n = min(total, kDnaQueue);
                for (int qi = lane; qi < n; qi += 32) dna_probe<H>(a, slab, s_bloom, base, (int)queue[qi], &s_count);
                while (rest) {  // queue full (dense bitmap): probe inline
                    const int s = __ffs(rest) - 1;
                    rest &= rest - 1u;
                    dna_probe<H>(a, slab, s_bloom, base, 32 * H * lane + H * s, &s_count);
                }
            }
            __syncwarp();  // the slab and the queue are rewritten by the next iteration
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) a.seg_count[blockIdx.x] = min(s_count, a.seg_cap);
}

// 2-bit copy of a text buffer for repeated searches: out[w] = dna_pack16 of bytes [16 w, 16 w + 16) (bytes past the end
// pack as code 0, like the scan's own bounds-checked loads).  nwords = ceil(len / 16) + 1 (one pad word).
__global__ void __launch_bounds__(256) dna_pack_kernel(const uint8_t *__restrict__ buf, long long len, uint32_t *__restrict__ out,
                                                        long long nwords) {
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (long long)gridDim.x * blockDim.x) {
        const long long pos = 16 * w;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (pos + 16 <= len) {
            v = __ldg(reinterpret_cast<const uint4 *>(buf + pos));
        } else if (pos < len) {
            uint32_t t[4] = {0u, 0u, 0u, 0u};
            for (int i = 0; i < 16 && pos + i < len; ++i) t[i >> 2] |= (uint32_t)buf[pos + i] << (8 * (i & 3));
            v = make_uint4(t[0], t[1], t[2], t[3]);
        }
        out[w] = dna_pack16(v);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Verification, stage 1: lower bound of one side of (*) from a single 32-row word.
// pq = match masks of the (up to) 32 pattern symbols next to the seed, by 2-bit code; T walks away from the seed
// with stride ts.  Returns min over the band columns c in [rows - k, rows + k] of D[rows][c], rows = min(32, part).
// Any alignment of the part within k crosses row `rows` at such a column, so this bounds the part's cost from below
// (aliasing of non-ACGT text bytes can only lower it further).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDnaInf = 1 << 20;
__device__ __forceinline__ int dna_lb32(uint4 pq, const uint8_t *__restrict__ T, int ts, int part_rows, long long avail, int k) {
    const int rows = min(32, part_rows);
    if (rows == 0) return 0;
    const int ncols = (int)min((long long)(rows + k), avail);
    if (ncols < rows - k) return kDnaInf;
    const int bo = rows - 1;
    uint32_t Pv = 0xFFFFFFFFu, Mv = 0u;
    int score = rows, best = rows <= k ? rows : kDnaInf;
    for (int c = 1; c <= ncols; ++c) {
        const uint32_t code = T[(c - 1) * ts];
        const uint32_t lo = (code & 2u) ? pq.y : pq.x, hi = (code & 2u) ? pq.w : pq.z;  // code = (byte >> 1) & 3
        const uint32_t Eq = (code & 4u) ? hi : lo;
        const uint32_t Xv = Eq | Mv;
        const uint32_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
        const uint32_t Ph = Mv | ~(Xh | Pv);
        const uint32_t Mh = Pv & Xh;
        score += (int)((Ph >> bo) & 1u) - (int)((Mh >> bo) & 1u);
        if (c >= rows - k) best = min(best, score);
        const uint32_t Phs = (Ph << 1) | 1u, Mhs = Mh << 1;
        Pv = Mhs | ~(Xv | Phs);
        Mv = Phs & Xv;
    }
    return best;
}

constexpr int kDnaVerifyThreads = 256;
// grid = nseg x parts: block (seg, part) takes every parts-th 256-entry slice of the scan CTA's segment
__global__ void __launch_bounds__(kDnaVerifyThreads) dna_verify_kernel(const __grid_constant__ DnaArgs a, int parts) {
    if (*a.overflow) return;  // the band kernel takes this round instead
    const int seg = blockIdx.x / parts, part = blockIdx.x - seg * parts;
    const unsigned int n = a.seg_count[seg];
    const uint64_t *cand = a.cand + (size_t)seg * a.seg_cap;
    const long long t_lim = min(a.n_end, a.buf_len);
    const int k = a.k;
    for (unsigned int c = part * kDnaVerifyThreads + threadIdx.x; c < n; c += parts * kDnaVerifyThreads) {
        const uint64_t e = cand[c];
        const uint32_t slot = (uint32_t)(e >> 40);
        const int piece = (int)((e >> kDnaPosBits) & 31u);
        const long long tp = a.w0 + (long long)(e & ((1ull << kDnaPosBits) - 1ull));
        const int m = __ldg(a.fp_m + slot);
        const int o = dna_piece_offset(piece, m, k), len = dna_piece_offset(piece + 1, m, k) - o;
        const uint4 *pq = a.peq32 + ((size_t)slot * (k + 1) + piece) * 2;
        const int lbl = dna_lb32(__ldg(pq), a.buf + tp - 1, -1, o, tp, k);
        if (lbl > k) continue;
        const int lbr = dna_lb32(__ldg(pq + 1), a.buf + tp + len, 1, m - o - len, t_lim - (tp + len), k);
        if (lbl + lbr > k) continue;
        const unsigned long long pos = atomicAdd(a.nlong, 1ull);
        if (pos < a.longcap) a.longq[pos] = e;
        else *a.overflow = 1u;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Verification, stage 2: exact last row of one side by a whole warp (byte-exact compares).  Lane l holds rows
// 32 l + 1 .. 32 l + 32 of the column.  Returns, in lane x <= 2k, D[rows][rows - k + x] (kDnaInf where that column
// does not exist).  All lanes call it with the same arguments.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int dna_exact_rows(const uint8_t *__restrict__ Pp, int ps, const uint8_t *__restrict__ T, int ts,
                                              int rows, long long avail, int k) {
    const int lane = threadIdx.x & 31;
    if (rows == 0) {  // D[0][c] = c
        const int c = lane - k;
        return (c >= 0 && c <= avail && lane <= 2 * k) ? c : kDnaInf;
    }
    uint32_t pA = 0u, pC = 0u, pG = 0u, pT = 0u;
    for (int b = 0; b < 32; ++b) {
        const int r = 32 * lane + b;
        if (r < rows) {
            const uint32_t ch = Pp[r * ps];
            pA |= (uint32_t)(ch == 'A') << b;
            pC |= (uint32_t)(ch == 'C') << b;
            pG |= (uint32_t)(ch == 'G') << b;
            pT |= (uint32_t)(ch == 'T') << b;
        }
    }
    const int ncols = (int)min((long long)(rows + k), avail);
    const int lo = (rows - 1) >> 5, bo = (rows - 1) & 31;
    uint32_t Pv = 0xFFFFFFFFu, Mv = 0u;
    int score = rows;
    int X = (rows <= k && lane == k - rows) ? rows : kDnaInf;  // column 0
    for (int c = 1; c <= ncols; ++c) {
        const uint32_t tc = T[(c - 1) * ts];
        const uint32_t Eq = tc == 'A' ? pA : (tc == 'C' ? pC : (tc == 'G' ? pG : (tc == 'T' ? pT : 0u)));
        const uint32_t tt = Eq & Pv;
        const uint32_t s0 = tt + Pv;
        const uint32_t G = __ballot_sync(0xFFFFFFFFu, s0 < tt);
        const uint32_t Pr = __ballot_sync(0xFFFFFFFFu, s0 == 0xFFFFFFFFu);
        const uint32_t U = G | Pr;
        const uint32_t s = s0 + ((((U + G) ^ U ^ G) >> lane) & 1u);
        const uint32_t Xv = Eq | Mv;
        const uint32_t Xh = (s ^ Pv) | Eq;
        const uint32_t Ph = Mv | ~(Xh | Pv);
        const uint32_t Mh = Pv & Xh;
        score += __shfl_sync(0xFFFFFFFFu, (int)((Ph >> bo) & 1u) - (int)((Mh >> bo) & 1u), lo);
        if (lane == c - (rows - k)) X = score;
        uint32_t tops = __shfl_up_sync(0xFFFFFFFFu, (Ph >> 31) | ((Mh >> 31) << 1), 1);
        if (lane == 0) tops = 1u;
        const uint32_t Phs = (Ph << 1) | (tops & 1u);
        const uint32_t Mhs = (Mh << 1) | (tops >> 1);
        Pv = Mhs | ~(Xv | Phs);
        Mv = Phs & Xv;
    }
    return X;
}

__device__ __forceinline__ uint64_t dna_set_key(uint32_t slot, unsigned long long rel) { return (((uint64_t)slot << kDnaPosBits) | rel) + 1ull; }

constexpr int kDnaLongThreads = 128;
__global__ void __launch_bounds__(kDnaLongThreads) dna_verify_long_kernel(const __grid_constant__ DnaArgs a) {
    if (*a.overflow) return;
    const unsigned long long n = min(*a.nlong, a.longcap);
    const int k = a.k, lane = threadIdx.x & 31;
    const long long t_lim = min(a.n_end, a.buf_len);
    const unsigned long long nwarps = (unsigned long long)gridDim.x * (kDnaLongThreads / 32);
    const uint64_t smask = (1ull << a.set_log) - 1ull;
    for (unsigned long long c = (unsigned long long)blockIdx.x * (kDnaLongThreads / 32) + (threadIdx.x >> 5); c < n; c += nwarps) {
        const uint64_t e = a.longq[c];
        const uint32_t slot = (uint32_t)(e >> 40);
        const int piece = (int)((e >> kDnaPosBits) & 31u);
        const long long tp = a.w0 + (long long)(e & ((1ull << kDnaPosBits) - 1ull));
        const int m = __ldg(a.fp_m + slot);
        const uint8_t *P = a.pat_bytes + __ldg(a.fp_off + slot);
        const int o = dna_piece_offset(piece, m, k), len = dna_piece_offset(piece + 1, m, k) - o, mr = m - o - len;
        // the scan compared 2-bit codes of the first 16 symbols: the piece must occur byte for byte
        bool same = true;
        for (int x = lane; x < len; x += 32) same = same && P[o + x] == a.buf[tp + x];
        if (!__all_sync(0xFFFFFFFFu, same)) continue;
        // lane x: L = cost of the left part for the window shifted by d = x - k, R the same for d = k - x
        const int L = dna_exact_rows(P + o - 1, -1, a.buf + tp - 1, -1, o, tp, k);
        const int R = dna_exact_rows(P + o + len, 1, a.buf + tp + len, 1, mr, t_lim - (tp + len), k);
        const int Rd = __shfl_sync(0xFFFFFFFFu, R, lane <= 2 * k ? 2 * k - lane : 0);
        if (lane > 2 * k || L + Rd > k) continue;
        const long long j = tp - o - (lane - k);
        if (j < a.w0 || j < 0 || j >= a.w1 || j + m > a.n_end) continue;
        // insert (slot, j) into the set: linear probing, bounded; a full set raises the overflow flag
        const uint64_t key = dna_set_key(slot, (unsigned long long)(j - a.w0));
        uint64_t h = (key * 0x9E3779B97F4A7C15ull) >> (64 - a.set_log);
        bool done = false;
        for (int probe = 0; probe < 256 && !done; ++probe) {
            const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long *>(a.set + h), 0ull, (unsigned long long)key);
            if (old == 0ull) {
                done = true;
                if (atomicAdd(a.nset, 1ull) > (smask >> 1)) *a.overflow = 1u;  // more than half full
            } else if (old == key) {
                done = true;
            }
            h = (h + 1) & smask;
        }
        if (!done) *a.overflow = 1u;
    }
}

__global__ void __launch_bounds__(256) dna_collect_kernel(const __grid_constant__ DnaArgs a) {
    if (*a.overflow) return;
    const unsigned long long slots = 1ull << a.set_log;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += (unsigned long long)gridDim.x * blockDim.x) {
        uint64_t key = a.set[i];
        if (!key) continue;
        key -= 1ull;
        const uint32_t slot = (uint32_t)(key >> kDnaPosBits);
        const int id = __ldg(a.fp_id + slot);
        atomicAdd(&a.counts[id], 1ull);
        if (a.sink.buf) hit_emit(a.sink, id, a.w0 + (long long)(key & ((1ull << kDnaPosBits) - 1ull)));
    }
}

#endif  // __CUDACC__

}  // namespace apm
