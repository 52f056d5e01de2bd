// apm_tile.cuh -- staging of a text tile (window starts + halo) from HBM into shared memory.
//
// A tile is the local byte range [ts, te) of the device text buffer.  The bulk of it is fetched with one
// TMA 1-D bulk copy (cp.async.bulk.shared.global + mbarrier complete_tx; SASS UBLKCP), which needs 16-byte
// aligned source/destination and a size that is a multiple of 16; the (at most 15 + 15) fringe bytes that
// do not fit that box are read with plain loads while the tile is re-coded from raw bytes to the per-call
// compact alphabet.  Shared buffers use a0 (16-byte aligned in GLOBAL address terms) as origin: the shared
// index of local byte i is i - a0.
#pragma once
#include "apm_common.cuh"

namespace apm {

#ifdef __CUDACC__

struct TileGeom {
    long long ts, te;  // bytes needed: [ts, te)
    long long a0;      // origin of the shared buffers (may be negative by < 16)
    long long ta, tb;  // bytes fetched by the TMA box: [ta, tb), 16-byte aligned, inside the buffer
};

// capacity (bytes) of one staging buffer for `tile` window starts and a `halo`-byte halo
__host__ __device__ inline size_t tile_cap(int tile, int halo) {
    return (size_t)((tile + halo + 15 + 15) / 16) * 16 + 16;
}

__device__ __forceinline__ TileGeom tile_geometry(const uint8_t *buf, long long buf_len, long long w0, long long t,
                                                  int tile, int halo) {
    TileGeom g;
    g.ts = w0 + t * tile;
    g.te = g.ts + tile + halo;
    if (g.te > buf_len) g.te = buf_len;
    g.a0 = g.ts - (long long)((reinterpret_cast<uintptr_t>(buf) + (uintptr_t)g.ts) & 15);
    g.ta = g.a0 < 0 ? g.a0 + ((-g.a0 + 15) / 16) * 16 : g.a0;  // first 16-byte aligned byte inside the buffer
    g.tb = g.a0 + ((g.te - g.a0 + 15) / 16) * 16;
    if (g.tb > buf_len) g.tb -= 16;
    return g;
}

// one thread: arm the barrier and start the bulk copy (no-op when the box is empty)
__device__ __forceinline__ void tile_issue(const TileGeom &g, const uint8_t *buf, uint8_t *raw, uint64_t *bar) {
    if (g.tb > g.ta) {
        const uint32_t bytes = (uint32_t)(g.tb - g.ta);
        mbar_arrive_expect_tx(bar, bytes);
        tma_bulk_g2s(raw + (g.ta - g.a0), buf + g.ta, bytes, bar);
    }
}

// all threads: raw bytes -> codes[i - a0] = map[byte] for local bytes i in [a0, a0 + span); positions that
// lie outside the buffer or beyond te get `pad`.  span must be a multiple of 4 and >= te - a0.
__device__ __forceinline__ void tile_encode(const TileGeom &g, const uint8_t *buf, long long buf_len,
                                            const uint8_t *raw, const uint8_t *map, uint8_t pad, uint8_t *codes,
                                            int span, int tid, int nthreads) {
    const int nwords = span / 4;
    for (int q = tid; q < nwords; q += nthreads) {
        const long long i0 = g.a0 + 4ll * q;
        uint32_t out;
        if (i0 >= g.ta && i0 + 4 <= g.tb && i0 + 4 <= g.te) {
            const uint32_t w = reinterpret_cast<const uint32_t *>(raw)[q];
            out = (uint32_t)map[w & 0xFF] | ((uint32_t)map[(w >> 8) & 0xFF] << 8) |
                  ((uint32_t)map[(w >> 16) & 0xFF] << 16) | ((uint32_t)map[w >> 24] << 24);
        } else {
            out = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const long long i = i0 + b;
                uint32_t code = pad;
                if (i >= 0 && i < g.te) code = map[(i >= g.ta && i < g.tb) ? raw[4 * q + b] : buf[i]];
                out |= code << (8 * b);
            }
        }
        reinterpret_cast<uint32_t *>(codes)[q] = out;
    }
}

#endif  // __CUDACC__

}  // namespace apm
