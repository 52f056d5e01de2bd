// apm_myers.cuh -- the hot path: bit-parallel GLOBAL edit distance (Hyyro's variant of Myers'
// algorithm) of every pattern against every full-length text window, thresholded and counted.
//
// Replaces: src/utils.c:76-99 (levenshtein) + the search loops src/sequential.c:105-144,
//           src/patterns_over_ranks.c:357-375, src/database_over_ranks.c:380-418 and their GPU clones
//           src/patterns_over_ranks.cu:19-73 / src/database_over_ranks.cu:20-134 of the reference.
// This is a from-scratch design, not a translation: the reference evaluates an O(m^2) scalar DP per
// window with a heap-allocated column; here one column of the DP is 2*NW 32-bit registers
// (vertical +1/-1 delta bit-vectors Pv/Mv; a 64-bit word is a register pair with a hardware carry
// chain) and a column step costs 10 integer instructions per 32-bit word.
//
// Mapping
//   windows  -> lanes   (thread t of a CTA owns window starts tile_start + t + s*kThreads)
//   patterns -> register blocking: a thread advances R patterns of equal length per text symbol, so
//               one text-code fetch and one address computation feed R Peq look-ups
//   text     -> CTA tile of `tile` window starts + (m_max-1)-byte halo, staged global->shared with a
//               TMA 1-D bulk copy (cp.async.bulk + mbarrier complete_tx), double buffered, then
//               re-coded in shared memory from raw bytes to the per-call compact alphabet
//   Peq      -> shared memory, layout [group][code][R][NW] words: the entry of one (group, code) is
//               contiguous, so the R*NW match words of a symbol are 16-byte vector loads (LDS.128)
//   counts   -> warp ballot + popc, one shared-memory atomic per (warp, pattern, tile), one 64-bit
//               global atomic per (CTA, pattern, chunk)
#pragma once
#include "apm_common.cuh"
#include "apm_tile.cuh"

namespace apm {

// ------------------------------------------------------------------------------------------------
// t[w] += p[w] with carry propagation across the NW 32-bit words (IADD3 / IADD3.X chain).
// ------------------------------------------------------------------------------------------------
template <int NW>
__host__ __device__ __forceinline__ void add_chain(uint32_t (&t)[NW], const uint32_t (&p)[NW]) {
#ifdef __CUDA_ARCH__
    if constexpr (NW == 1) {
        t[0] += p[0];
    } else if constexpr (NW == 2) {
        asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(t[0]), "+r"(t[1]) : "r"(p[0]), "r"(p[1]));
    } else if constexpr (NW == 3) {
        asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, %5;"
            : "+r"(t[0]), "+r"(t[1]), "+r"(t[2])
            : "r"(p[0]), "r"(p[1]), "r"(p[2]));
    } else if constexpr (NW == 4) {
        asm("add.cc.u32 %0, %0, %4;\n\taddc.cc.u32 %1, %1, %5;\n\taddc.cc.u32 %2, %2, %6;\n\taddc.u32 %3, %3, %7;"
            : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3])
            : "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]));
    } else if constexpr (NW == 5) {
        asm("add.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %6;\n\taddc.cc.u32 %2, %2, %7;\n\t"
            "addc.cc.u32 %3, %3, %8;\n\taddc.u32 %4, %4, %9;"
            : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4])
            : "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]), "r"(p[4]));
    } else if constexpr (NW == 6) {
        asm("add.cc.u32 %0, %0, %6;\n\taddc.cc.u32 %1, %1, %7;\n\taddc.cc.u32 %2, %2, %8;\n\t"
            "addc.cc.u32 %3, %3, %9;\n\taddc.cc.u32 %4, %4, %10;\n\taddc.u32 %5, %5, %11;"
            : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5])
            : "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]), "r"(p[4]), "r"(p[5]));
    } else if constexpr (NW == 7) {
        asm("add.cc.u32 %0, %0, %7;\n\taddc.cc.u32 %1, %1, %8;\n\taddc.cc.u32 %2, %2, %9;\n\t"
            "addc.cc.u32 %3, %3, %10;\n\taddc.cc.u32 %4, %4, %11;\n\taddc.cc.u32 %5, %5, %12;\n\t"
            "addc.u32 %6, %6, %13;"
            : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6])
            : "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]), "r"(p[4]), "r"(p[5]), "r"(p[6]));
    } else {
        static_assert(NW == 8, "add_chain: NW must be 1..8");
        asm("add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %9;\n\taddc.cc.u32 %2, %2, %10;\n\t"
            "addc.cc.u32 %3, %3, %11;\n\taddc.cc.u32 %4, %4, %12;\n\taddc.cc.u32 %5, %5, %13;\n\t"
            "addc.cc.u32 %6, %6, %14;\n\taddc.u32 %7, %7, %15;"
            : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7])
            : "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]), "r"(p[4]), "r"(p[5]), "r"(p[6]), "r"(p[7]));
    }
#else
    uint64_t carry = 0;  // host path: used only by the CPU unit check of this header (tests/)
    for (int w = 0; w < NW; ++w) {
        uint64_t s = (uint64_t)t[w] + (uint64_t)p[w] + carry;
        t[w] = (uint32_t)s;
        carry = s >> 32;
    }
#endif
}

// (hi << 1) | (lo >> 31): one SHF.L.W funnel shift on the device.
__host__ __device__ __forceinline__ uint32_t shl1_carry(uint32_t lo, uint32_t hi) {
#ifdef __CUDA_ARCH__
    return __funnelshift_l(lo, hi, 1);
#else
    return (hi << 1) | (lo >> 31);
#endif
}

// ------------------------------------------------------------------------------------------------
// One DP column (one text symbol) of the GLOBAL distance.  State: Pv/Mv = rows where the vertical
// delta D[r][c]-D[r-1][c] is +1 / -1.  Boundary D[0][c] = c is the constant carry-in of 1 into the
// shifted positive horizontal delta (utils.c:88 `column[0] = x`); boundary D[r][0] = r is the
// initial state Pv = ~0, Mv = 0 (utils.c:84-86).  10 integer instructions per word:
//   Xv (1) t=Eq&Pv (1) t+=Pv (1) Xh (1) Ph (1) Mh (1) Ph<<1|c (1) Mh<<1|c (1) Pv' (1) Mv' (1)
// ------------------------------------------------------------------------------------------------
template <int NW>
__host__ __device__ __forceinline__ void myers_step(uint32_t (&Pv)[NW], uint32_t (&Mv)[NW],
                                                    const uint32_t (&Eq)[NW]) {
    uint32_t t[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) t[w] = Eq[w] & Pv[w];
    add_chain<NW>(t, Pv);
    uint32_t Ph_lo = 0u, Mh_lo = 0u;  // previous (lower) word, for the cross-word funnel shift
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const uint32_t Xv = Eq[w] | Mv[w];
        const uint32_t Xh = (t[w] ^ Pv[w]) | Eq[w];
        const uint32_t Ph = Mv[w] | ~(Xh | Pv[w]);
        const uint32_t Mh = Pv[w] & Xh;
        // word 0 shifts in the +1 horizontal delta of row 0 (D[0][c] = c)
        const uint32_t Phs = (w == 0) ? ((Ph << 1) | 1u) : shl1_carry(Ph_lo, Ph);
        const uint32_t Mhs = (w == 0) ? (Mh << 1) : shl1_carry(Mh_lo, Mh);
        Ph_lo = Ph;
        Mh_lo = Mh;
        Pv[w] = Mhs | ~(Xv | Phs);
        Mv[w] = Phs & Xv;
    }
}

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// FMA-pipe variants of the column step.  On sm_100 LOP3/SHF/IADD3 issue on the ALU pipe (64 lanes/clk/SM)
// and IMAD on the FMA pipe (another 64 lanes/clk/SM).  The 7 boolean instructions per word have to be
// LOP3s, so the step is ALU-pipe bound; these variants move the three arithmetic instructions (the
// carry add and the two 1-bit shifts, including the carries that cross 32-bit words) onto the FMA pipe:
//   x << 1 | cin      ->  mad.lo (x, 2, cin)            carry out of a word = high half of mul.wide(x, 2)
//   t + Pv            ->  mad.wide(t, 1, zext(Pv))      carry out of a word = high half of the product
// The multipliers 1 and 2 come from kernel parameters so that ptxas cannot strength-reduce the IMADs
// back into ALU-pipe shifts/adds.
//   V = 1: shifts on the FMA pipe, multi-word (NW > 2) adds stay an IADD3.X chain
//   V = 2: shifts and adds on the FMA pipe for every NW
// ------------------------------------------------------------------------------------------------
struct StepConst {
    uint32_t one, two;  // 1 and 2
    uint32_t estride;   // bytes per (group, code) Peq entry
    uint64_t one64;     // 1 as a 64-bit value whose high half ptxas cannot prove to be zero
};

__device__ __forceinline__ uint32_t madlo(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint64_t madwide(uint32_t a, uint32_t b, uint64_t c) {
    uint64_t d;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mulwide(uint32_t a, uint32_t b) {
    uint64_t d;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(d) : "r"(a), "r"(b));
    return d;
}

template <int NW, int V>
__device__ __forceinline__ void myers_step_fma(uint32_t (&Pv)[NW], uint32_t (&Mv)[NW], const uint32_t (&Eq)[NW],
                                               const StepConst &K) {
    uint32_t s[NW];
    if constexpr (NW == 1) {
        s[0] = madlo(Eq[0] & Pv[0], K.one, Pv[0]);
    } else if constexpr (V == 1 && NW > 2) {
#pragma unroll
        for (int w = 0; w < NW; ++w) s[w] = Eq[w] & Pv[w];
        add_chain<NW>(s, Pv);
    } else {
        uint32_t cin = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const uint32_t t = Eq[w] & Pv[w];
            if (w == NW - 1) {
                s[w] = madlo(cin, K.one, madlo(t, K.one, Pv[w]));
            } else {
                // t + Pv (+ carry-in) as 64-bit values; every addend is the full result of a wide
                // multiply, i.e. a natural register pair, so each line is a single IMAD.WIDE
                uint64_t W = madwide(t, K.one, mulwide(Pv[w], K.one));
                if (w > 0) W = madwide(cin, K.one, W);
                s[w] = (uint32_t)W;
                cin = (uint32_t)(W >> 32);
            }
        }
    }
    uint32_t cp = 0, cm = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const uint32_t Xv = Eq[w] | Mv[w];
        const uint32_t Xh = (s[w] ^ Pv[w]) | Eq[w];
        const uint32_t Ph = Mv[w] | ~(Xh | Pv[w]);
        const uint32_t Mh = Pv[w] & Xh;
        uint32_t Phs, Mhs;
        if (w == NW - 1) {  // top word: no carry-out needed
            Phs = madlo(Ph, K.two, w == 0 ? K.one : cp);
            Mhs = w == 0 ? madlo(Mh, K.two, 0u) : madlo(Mh, K.two, cm);
        } else {
            const uint64_t Wp = w == 0 ? madwide(Ph, K.two, K.one64) : mulwide(Ph, K.two);
            const uint64_t Wm = mulwide(Mh, K.two);
            Phs = w == 0 ? (uint32_t)Wp : madlo(cp, K.one, (uint32_t)Wp);
            Mhs = w == 0 ? (uint32_t)Wm : madlo(cm, K.one, (uint32_t)Wm);
            cp = (uint32_t)(Wp >> 32);
            cm = (uint32_t)(Wm >> 32);
        }
        Pv[w] = Mhs | ~(Xv | Phs);
        Mv[w] = Phs & Xv;
    }
}
#endif  // __CUDACC__

// D[len][len] - len = sum of the vertical deltas of the last column over rows 1..len.
template <int NW>
__host__ __device__ __forceinline__ int myers_score_minus_len(const uint32_t (&Pv)[NW],
                                                              const uint32_t (&Mv)[NW], uint32_t topmask) {
    int s = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int w = 0; w < NW - 1; ++w) s += __popc(Pv[w]) - __popc(Mv[w]);
    s += __popc(Pv[NW - 1] & topmask) - __popc(Mv[NW - 1] & topmask);
#else
    for (int w = 0; w < NW; ++w) {
        uint32_t mk = (w == NW - 1) ? topmask : 0xFFFFFFFFu;
        s += __builtin_popcount(Pv[w] & mk) - __builtin_popcount(Mv[w] & mk);
    }
#endif
    return s;
}

#ifdef __CUDACC__

struct MyersArgs {
    const uint8_t *buf;          // device text; buf[0] is global byte `buf_offset` (resolved by host)
    long long buf_len;           // valid bytes in buf
    long long n_end;             // local index of the global end of the text (n_total - buf_offset)
    long long w0, w1;            // local window-start range [w0, w1)
    const uint32_t *peq;         // [ngroups][ncodes][EW]
    const int *group_m;          // [ngroups]       pattern length of the group
    const int *group_pat;        // [ngroups * R]   pattern index or -1 (padding slot)
    const uint8_t *code_of;      // [256] byte -> compact alphabet code
    unsigned long long *counts;  // [nb_patterns]
    int ngroups, ncodes, groups_per_chunk;
    int mmax;                    // largest m in this bucket
    int k;                       // approx_factor
    int tile;                    // window starts per tile (multiple of kThreads)
    uint32_t c_one, c_two;       // the constants 1 and 2, opaque to ptxas (see myers_step_fma)
    const unsigned int *run_if;  // optional gate: the whole launch is a no-op when *run_if == 0 (filter fallback)
    HitSink sink;                // optional match-position output
};

// Shared-memory layout (dynamic): see myers_smem_bytes() -- the host uses the same formula.
__host__ __device__ inline size_t myers_tile_cap(int tile, int mmax) { return tile_cap(tile, mmax - 1); }
__host__ __device__ inline size_t myers_smem_bytes(int tile, int mmax, int groups_per_chunk, int ncodes,
                                                   int R, int NW) {
    const size_t cap = myers_tile_cap(tile, mmax);
    size_t b = 64;                                             // 2 mbarriers (+pad)
    b += 256;                                                  // code_of
    b += 3 * cap;                                              // raw[2], codes
    b += (size_t)groups_per_chunk * (sizeof(int) * (1 + R) + sizeof(uint32_t) * R);  // m, pat ids, counts
    b = (b + 15) / 16 * 16;
    b += (size_t)groups_per_chunk * ncodes * entry_words(R * NW) * sizeof(uint32_t);
    return b + 16;
}

template <int N>
__device__ __forceinline__ void load_entry(const uint32_t *__restrict__ e, uint32_t *dst) {
    constexpr int EW = entry_words(N);
    uint32_t tmp[EW];
    if constexpr (EW >= 4) {
#pragma unroll
        for (int i = 0; i < EW / 4; ++i) {
            const uint4 v = reinterpret_cast<const uint4 *>(e)[i];
            tmp[4 * i + 0] = v.x; tmp[4 * i + 1] = v.y; tmp[4 * i + 2] = v.z; tmp[4 * i + 3] = v.w;
        }
    } else if constexpr (EW == 2) {
        const uint2 v = *reinterpret_cast<const uint2 *>(e);
        tmp[0] = v.x; tmp[1] = v.y;
    } else {
        tmp[0] = e[0];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) dst[i] = tmp[i];
}

template <int NW, int R, int V>
__device__ __forceinline__ void advance_columns(const uint8_t *__restrict__ tp, const unsigned char *__restrict__ sm,
                                                uint32_t pq_off, uint32_t (&Pv)[R][NW], uint32_t (&Mv)[R][NW],
                                                const StepConst &K) {
    constexpr int EW = entry_words(R * NW);
    const uint32_t c = *tp;
    // byte offset of the (group, code) entry; V > 0 keeps this address computation on the FMA pipe too
    const uint32_t off = V == 0 ? pq_off + c * (EW * 4u) : madlo(c, K.estride, pq_off);
    uint32_t Eq[R][NW];
    load_entry<R * NW>(reinterpret_cast<const uint32_t *>(sm + off), &Eq[0][0]);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if constexpr (V == 0) myers_step<NW>(Pv[r], Mv[r], Eq[r]);
        else myers_step_fma<NW, V>(Pv[r], Mv[r], Eq[r], K);
    }
}

// ------------------------------------------------------------------------------------------------
// Persistent count kernel.  grid.x strides over text tiles, grid.y over pattern chunks; the host sizes
// grid.x * grid.y to (#SMs x resident CTAs).
// ------------------------------------------------------------------------------------------------
template <int NW, int R, int V>
__global__ void __launch_bounds__(kThreads) myers_count_kernel(const MyersArgs a) {
    constexpr int EW = entry_words(R * NW);
    constexpr int U = 8;  // text symbols per unrolled inner-loop body
    if (a.run_if && *a.run_if == 0u) return;
    extern __shared__ __align__(128) unsigned char smem[];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const StepConst K = {a.c_one, a.c_two, (uint32_t)(EW * 4) * a.c_one, mulwide(a.c_one, a.c_one)};
    const size_t cap = myers_tile_cap(a.tile, a.mmax);

    // carve-up by integer offsets from the __shared__ array so every access stays an LDS/STS
    const size_t off_raw = 64 + 256;
    const size_t off_codes = off_raw + 2 * cap;
    const size_t off_gm = off_codes + cap;
    const size_t off_gpat = off_gm + sizeof(int) * a.groups_per_chunk;
    const size_t off_cnt = off_gpat + sizeof(int) * a.groups_per_chunk * R;
    const size_t off_peq = (off_cnt + sizeof(uint32_t) * a.groups_per_chunk * R + 15) & ~size_t(15);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);             // [2]
    uint8_t *s_map = smem + 64;                                       // [256]
    uint8_t *s_raw0 = smem + off_raw;                                 // [cap] x2, 16-byte aligned
    uint8_t *s_codes = smem + off_codes;                              // [cap]
    int *s_gm = reinterpret_cast<int *>(smem + off_gm);               // [gpc]
    int *s_gpat = reinterpret_cast<int *>(smem + off_gpat);           // [gpc * R]
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(smem + off_cnt);   // [gpc * R]
    uint32_t *s_peq = reinterpret_cast<uint32_t *>(smem + off_peq);   // [gpc][ncodes][EW]

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    static_assert(kThreads == 256, "s_map fill assumes 256 threads");
    s_map[tid] = a.code_of[tid];
    // codes start as 0 (a valid code): lanes past the last valid window read, but never use, stale codes
    for (size_t i = tid; i < cap / 4; i += kThreads) reinterpret_cast<uint32_t *>(s_codes)[i] = 0u;
    __syncthreads();

    const long long nwin = a.w1 - a.w0;
    const long long ntiles = (nwin + a.tile - 1) / a.tile;
    const int slots = a.tile / kThreads;
    uint32_t phase = 0;  // bit s = parity to wait for on barrier s

    // Work items = (pattern chunk, text tile), chunk-major, dealt round-robin to the persistent CTAs; the
    // text tile of the CTA's next item is prefetched by TMA while the current one is being processed.
    const int halo = a.mmax - 1;
    const int nchunks = (a.ngroups + a.groups_per_chunk - 1) / a.groups_per_chunk;
    const long long nitems = ntiles * nchunks;
    int cur_chunk = -1, gcount = 0, stage = 0;
    long long it = blockIdx.x;
    if (tid == 0 && it < nitems)
        tile_issue(tile_geometry(a.buf, a.buf_len, a.w0, it % ntiles, a.tile, halo), a.buf, s_raw0, &bars[0]);
    {
        for (; it < nitems; it += gridDim.x) {
            const long long itn = it + gridDim.x;
            if (tid == 0 && itn < nitems)  // prefetch the next item's tile
                tile_issue(tile_geometry(a.buf, a.buf_len, a.w0, itn % ntiles, a.tile, halo), a.buf,
                           s_raw0 + (stage ^ 1) * cap, &bars[stage ^ 1]);
            const int chunk = (int)(it / ntiles);
            const long long t = it % ntiles;
            if (chunk != cur_chunk) {
                // flush the CTA-local counters of the previous chunk, then stage this chunk's Peq tables,
                // lengths and pattern ids (thread i owns slot i in both loops)
                for (int i = tid; i < gcount * R; i += kThreads) {
                    const int p = s_gpat[i];
                    const uint32_t c = s_cnt[i];
                    if (p >= 0 && c) atomicAdd(&a.counts[p], (unsigned long long)c);
                }
                const int g0 = chunk * a.groups_per_chunk;
                gcount = min(a.groups_per_chunk, a.ngroups - g0);
                for (int i = tid; i < gcount * R; i += kThreads) {
                    s_gpat[i] = a.group_pat[(size_t)g0 * R + i];
                    s_cnt[i] = 0;
                }
                const uint32_t *src = a.peq + (size_t)g0 * a.ncodes * EW;
                const int nw = gcount * a.ncodes * EW;
                for (int i = tid; i < nw; i += kThreads) s_peq[i] = src[i];
                for (int i = tid; i < gcount; i += kThreads) s_gm[i] = a.group_m[g0 + i];
                cur_chunk = chunk;
            }
            const TileGeom tg = tile_geometry(a.buf, a.buf_len, a.w0, t, a.tile, halo);
            const long long ts = tg.ts, a0 = tg.a0;
            if (tg.tb > tg.ta) {
                mbar_wait(&bars[stage], (phase >> stage) & 1u);
                phase ^= (1u << stage);
            }
            // raw bytes -> compact alphabet codes (fringe bytes outside the TMA box come from global memory)
            tile_encode(tg, a.buf, a.buf_len, s_raw0 + stage * cap, s_map, s_map[0], s_codes,
                        (int)((tg.te - tg.a0 + 3) / 4 * 4), tid, kThreads);
            __syncthreads();

            // ---- the hot loop: groups (R patterns) x window slots x m columns
            const long long tile_end = min(ts + (long long)a.tile, a.w1);
            const uint8_t *tile_codes = s_codes + (ts - a0);
            for (int g = 0; g < gcount; ++g) {
                const int m = s_gm[g];
                const long long lim = min(tile_end, a.n_end - m + 1);  // full windows only
                const uint32_t pq = (uint32_t)off_peq + (uint32_t)g * (uint32_t)a.ncodes * (EW * 4u);  // byte offset
                const int topbits = m - 32 * (NW - 1);
                const uint32_t topmask = topbits >= 32 ? 0xFFFFFFFFu : ((1u << topbits) - 1u);
                const int thresh = a.k - m;
                uint32_t cnt[R];
#pragma unroll
                for (int r = 0; r < R; ++r) cnt[r] = 0;

                for (int s = 0; s < slots; ++s) {
                    const int woff = s * kThreads + tid;
                    const bool valid = ts + woff < lim;
                    if (!__any_sync(0xFFFFFFFFu, valid)) break;  // warp-uniform: later slots are past lim too
                    const uint8_t *tp = tile_codes + woff;
                    uint32_t Pv[R][NW], Mv[R][NW];
#pragma unroll
                    for (int r = 0; r < R; ++r)
#pragma unroll
                        for (int w = 0; w < NW; ++w) { Pv[r][w] = 0xFFFFFFFFu; Mv[r][w] = 0u; }
                    int x = 0;
#pragma unroll 1
                    for (; x + U <= m; x += U) {
#pragma unroll
                        for (int u = 0; u < U; ++u) advance_columns<NW, R, V>(tp + x + u, smem, pq, Pv, Mv, K);
                    }
#pragma unroll 1
                    for (; x < m; ++x) advance_columns<NW, R, V>(tp + x, smem, pq, Pv, Mv, K);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const bool hit = valid && (myers_score_minus_len<NW>(Pv[r], Mv[r], topmask) <= thresh);
                        cnt[r] += __popc(__ballot_sync(0xFFFFFFFFu, hit));
                        if (a.sink.buf && hit && s_gpat[g * R + r] >= 0) hit_emit(a.sink, s_gpat[g * R + r], ts + woff);
                    }
                }
                if (lane == 0) {
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (cnt[r]) atomicAdd(&s_cnt[g * R + r], cnt[r]);
                }
            }
            __syncthreads();  // codes + raw[stage] free for reuse
            stage ^= 1;
        }
    }
    for (int i = tid; i < gcount * R; i += kThreads) {
        const int p = s_gpat[i];
        const uint32_t c = s_cnt[i];
        if (p >= 0 && c) atomicAdd(&a.counts[p], (unsigned long long)c);
    }
}

#endif  // __CUDACC__

}  // namespace apm
