// apm_api.cu -- C-ABI launcher of libapm_b200 (see include/apm_b200.h for the contract and the
// reference interfaces each entry point replaces).  Host side only: argument checking, the per-call
// alphabet / Peq tables / pattern groups ("plan"), shard arithmetic, kernel launches.  All arithmetic
// of the hot path happens in the CUDA kernels (apm_myers.cuh, apm_dp.cuh); there is no CPU fallback.
#include "../../include/apm_b200.h"

#include <cuda_runtime.h>
#include <chrono>
#include <climits>
#include <map>
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "apm_common.cuh"
#include "apm_dp.cuh"
#include "apm_myers.cuh"
#include "apm_sliced.cuh"
#include "apm_band.cuh"
#include "apm_filter.cuh"
#include "apm_tail.cuh"
#include "apm_dna.h"
#include "apm_util_kernels.cuh"

using namespace apm;

namespace {

thread_local std::string tl_err;
std::atomic<unsigned long long> g_launches{0};

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    tl_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(APM_ECUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

enum { SHARD_AUTO = 0, SHARD_DB = 1, SHARD_PATTERNS = 2 };
enum { KERNEL_AUTO = 0, KERNEL_MYERS = 1, KERNEL_DP = 2, KERNEL_SLICED = 3 };
enum { MODE_DIRECT = 0, MODE_BAND = 1, MODE_FILTER = 2 };

struct Options {
    int gpus = 1;  // 0 = all
    int shard = SHARD_AUTO;
    int kernel = KERNEL_AUTO;
    int mode = MODE_DIRECT;
    int rblock = 0;  // 0 = auto
    int tile = 0;    // 0 = auto
    int variant = 0; // column-step code variant (see myers_step_fma)
    int cell = -1;   // DP-cell code of the sliced/band kernels: -1 auto, 0 = 5 LOP3, 1 = 4 LOP3 + 3 IMAD, 2 = 4 LOP3 + 2 IMAD,
                     // 3 = 4 LOP3 + 3 IMAD ordered for the operand reuse cache (apm_sliced.cuh)
    int filter_scan = 0;  // mode=filter: 0 auto (2-bit DNA scan when the patterns allow it), 1 hashed scan, 2 DNA scan
    int tail = 0;    // truncated tail windows: 0 = bit-parallel kernels (apm_tail.cuh), 1 = explicit DP (apm_dp.cuh)
    int reduce = 0;  // multi-GPU count reduction: 0 auto (p2p kernel, else NCCL, else host), 1 nccl, 2 host sum, 3 p2p
    long long dp_scratch_mb = 256;
    long long text_chunk_mb = 32768;  // one-shot API: a shard larger than this is streamed through two device buffers
    long long filter_cand_mb = 128;  // candidate buffer of the seed filter (mode=filter), MiB
    int ingest_threads = 0;  // reader threads per GPU of the chunked ingest (0 = auto: host threads / GPUs, 1..8)
    long long cache_mb = 4096;  // device memory kept for reuse between calls (dev_alloc / dev_free)
};
std::mutex g_opt_mu;
Options g_opt;
thread_local std::string tl_optbuf;

Options options_snapshot() {
    std::lock_guard<std::mutex> lk(g_opt_mu);
    return g_opt;
}

int device_ready(int *ndev) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(APM_ENODEVICE, "no usable CUDA device (%s); libapm_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    *ndev = n;
    return APM_OK;
}


// ---- device-memory cache -----------------------------------------------------------------------------
// cudaMalloc / cudaFree cost 0.1 - 100 ms each and cudaFree synchronises the device; the one-shot API would
// pay ~15 of them per call (measured with APM_TRACE=1: up to 126 ms in "release").  Freed blocks are kept per
// device in size classes and handed out again; everything cached can be returned to the driver with
// apm_release_cache().  Callers only free a block after the work using it has been synchronised.
struct DevPool {
    std::mutex mu;
    std::map<std::pair<int, size_t>, std::vector<void *>> free_blocks;  // (device, class bytes) -> blocks
    std::map<void *, std::pair<int, size_t>> live;                      // block -> (device, class bytes)
    size_t cached_bytes = 0;
    std::vector<uint8_t *> pinned_free;                                 // ingest staging buffers (kPinnedChunk each), not in use
};
DevPool g_pool;
constexpr size_t kPinnedChunk = (size_t)16 << 20;

// pinned staging buffers of the chunked ingest: taken from / returned to the pool (cudaHostAlloc costs milliseconds)
uint8_t *pinned_acquire() {
    {
        std::lock_guard<std::mutex> lk(g_pool.mu);
        if (!g_pool.pinned_free.empty()) {
            uint8_t *p = g_pool.pinned_free.back();
            g_pool.pinned_free.pop_back();
            return p;
        }
    }
    uint8_t *p = nullptr;
    if (cudaHostAlloc((void **)&p, kPinnedChunk, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void pinned_release(uint8_t *p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pool.mu);
    g_pool.pinned_free.push_back(p);
}

size_t pool_class(size_t bytes) {
    size_t c = 512;
    if (bytes > ((size_t)1 << 20)) return (bytes + ((size_t)2 << 20) - 1) / ((size_t)2 << 20) * ((size_t)2 << 20);
    while (c < bytes) c <<= 1;
    return c;
}

// hand every cached DEVICE block back to the driver (the pinned staging buffers stay: an ingest may be using them)
void release_device_cache() {
    std::lock_guard<std::mutex> lk(g_pool.mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto &kv : g_pool.free_blocks) {
        if (kv.second.empty()) continue;
        cudaSetDevice(kv.first.first);
        for (void *p : kv.second) cudaFree(p);
        kv.second.clear();
    }
    g_pool.free_blocks.clear();
    g_pool.cached_bytes = 0;
    cudaSetDevice(cur);
}

cudaError_t dev_alloc(void **p, size_t bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const size_t cls = pool_class(std::max<size_t>(bytes, 1));
    {
        std::lock_guard<std::mutex> lk(g_pool.mu);
        auto it = g_pool.free_blocks.find({dev, cls});
        if (it != g_pool.free_blocks.end() && !it->second.empty()) {
            *p = it->second.back();
            it->second.pop_back();
            g_pool.cached_bytes -= cls;
            g_pool.live[*p] = {dev, cls};
            return cudaSuccess;
        }
    }
    e = cudaMalloc(p, cls);
    if (e != cudaSuccess) {  // give the cache back to the driver and retry once
        cudaGetLastError();
        release_device_cache();
        e = cudaMalloc(p, cls);
    }
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lk(g_pool.mu);
        g_pool.live[*p] = {dev, cls};
    }
    return e;
}
template <typename T>
cudaError_t dev_alloc_t(T **p, size_t bytes) { return dev_alloc((void **)p, bytes); }

void dev_free(void *p) {
    if (!p) return;
    size_t limit;
    {
        std::lock_guard<std::mutex> lk(g_opt_mu);
        limit = (size_t)g_opt.cache_mb << 20;
    }
    std::lock_guard<std::mutex> lk(g_pool.mu);
    auto it = g_pool.live.find(p);
    if (it == g_pool.live.end()) {  // not ours (cannot happen) -- hand it to the driver
        cudaFree(p);
        return;
    }
    const int dev = it->second.first;
    const size_t cls = it->second.second;
    g_pool.live.erase(it);
    if (g_pool.cached_bytes + cls > limit) {
        int cur = 0;
        cudaGetDevice(&cur);
        if (cur != dev) cudaSetDevice(dev);
        cudaFree(p);
        if (cur != dev) cudaSetDevice(cur);
        return;
    }
    g_pool.free_blocks[{dev, cls}].push_back(p);
    g_pool.cached_bytes += cls;
}

// ---- dynamic shared-memory opt-in --------------------------------------------------------------------
// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the (kernel, device) pair, not to a plan: two live plans
// using the same instantiation with different sizes must not lower it under each other.  One process-wide table
// keeps the maximum ever requested per (device, kernel) and only ever raises it.
std::mutex g_smem_mu;
std::map<std::pair<int, const void *>, size_t> g_smem_set;
template <typename Fn>
cudaError_t ensure_dyn_smem(Fn fn, size_t bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(g_smem_mu);
    size_t &cur = g_smem_set[{dev, (const void *)fn}];
    if (bytes <= cur) return cudaSuccess;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
}

// ---- kernel dispatch table ------------------------------------------------------------------------
using MyersKernel = void (*)(const MyersArgs);
template <int NW, int V>
MyersKernel pick_r(int R) {
    switch (R) {
        case 1: return myers_count_kernel<NW, 1, V>;
        case 2: return myers_count_kernel<NW, 2, V>;
        default: return myers_count_kernel<NW, 4, V>;
    }
}
template <int V>
MyersKernel pick_nw(int NW, int R) {
    switch (NW) {
        case 1: return pick_r<1, V>(R);
        case 2: return pick_r<2, V>(R);
        case 3: return pick_r<3, V>(R);
        case 4: return pick_r<4, V>(R);
        case 5: return pick_r<5, V>(R);
        case 6: return pick_r<6, V>(R);
        case 7: return pick_r<7, V>(R);
        default: return pick_r<8, V>(R);
    }
}
MyersKernel pick_kernel(int NW, int R, int V) {
    switch (V) {
        case 0: return pick_nw<0>(NW, R);
        case 1: return pick_nw<1>(NW, R);
        default: return pick_nw<2>(NW, R);
    }
}

struct Bucket {
    int NW = 1, R = 1, EW = 1, ngroups = 0, mmax = 0, mmin = 0;
    std::vector<int> group_m, group_pat;
    std::vector<uint32_t> peq;
    uint32_t *d_peq = nullptr;
    int *d_group_m = nullptr, *d_group_pat = nullptr;
    MyersKernel fn = nullptr;
};

// patterns handled by the window-sliced kernel: one list for m <= 32 (MC = 32) and one for longer ones
struct SlicedList {
    int MC = 64, npat = 0, mmin = 0, mmax = 0, mcp = 0;
    int ragged = 0;  // 0: generic kernel; 1: single-block patterns with m % MC != 0 (compile-time block widths)
    std::vector<uint8_t> codes;  // [npat][mcp]
    std::vector<int> m, id;
    uint8_t *d_codes = nullptr;
    int *d_m = nullptr, *d_id = nullptr;
    uint2 *d_vscratch = nullptr;  // boundary deltas between column blocks (only when mmax > MC)
    unsigned long long *d_work = nullptr;  // item dispenser of the persistent kernel
    size_t vscratch_bytes = 0;
};


// patterns handled by the exact seed filter (apm_filter.cuh) and its device tables
struct FilterSet {
    std::vector<int> ids;  // pattern indices, slot order
    int s = 0, hb = 0, nent = 0, mmax = 0, mmin = 0;
    uint32_t bs = 0;
    uint32_t *d_bitmap = nullptr, *d_digest = nullptr, *d_ent_idx = nullptr, *d_ent_slot = nullptr;
    uint32_t *d_coarse = nullptr, *d_ent_hash = nullptr;
    uint8_t *d_ent_piece = nullptr;
    int *d_fp_id = nullptr, *d_fp_m = nullptr;
    long long *d_fp_off = nullptr;
    uint64_t *d_cand = nullptr;
    unsigned long long cap = 0;
    uint64_t *d_long = nullptr;           // long-lived candidates (finished by whole warps)
    unsigned long long long_cap = 0;
    unsigned long long *d_ctr = nullptr;  // [0] candidates, [1] (low word) overflow flag, [2] long-lived candidates
    bool use_dna = false;                 // 2-bit q-gram scan + seed-hit verification (apm_dna.cuh) instead of the hashed scan
    DnaSet dna;
};

template <typename T>
int upload(T **dptr, const std::vector<T> &h) {
    *dptr = nullptr;
    if (h.empty()) return APM_OK;
    CUDA_TRY(dev_alloc((void **)dptr, h.size() * sizeof(T)));
    CUDA_TRY(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return APM_OK;
}

}  // namespace

namespace apm {
cudaError_t pool_alloc(void **p, size_t bytes) { return dev_alloc(p, bytes); }
void pool_free(void *p) { dev_free(p); }
}  // namespace apm

struct apm_plan {
    int device = 0, P = 0, k = 0, ncodes = 0, num_sms = 0, mmax_all_patterns = 0, smem_optin = 0;
    Options opt;
    std::vector<std::string> pats;
    uint8_t code_of[256];
    int shard_rank = 0, shard_world = 1;
    std::vector<Bucket> buckets;
    std::vector<SlicedList> sliced;
    // mode=filter: patterns handled by the seed filter; fb_* = the same patterns as ordinary kernel lists, launched
    // behind the filter and gated on its overflow flag
    FilterSet filter;
    std::vector<Bucket> fb_buckets;
    std::vector<SlicedList> fb_sliced;
    int nplanes = 0;
    uint8_t *d_plane_of = nullptr;
    std::vector<int> tail_list, all_list;  // tail_list: patterns with m <= 256 first, then the longer ones
    int tail_nshort = 0;                   // patterns of tail_list served by the thread-per-window tail kernel
    int tail_width = 0, tail_mmax = 0, all_mmax = 0, tail_width_short = 0;
    uint8_t *d_code_of = nullptr, *d_pat_bytes = nullptr;
    long long *d_pat_off = nullptr;
    int *d_pat_len = nullptr, *d_tail_list = nullptr, *d_all_list = nullptr;
    unsigned long long *d_counts = nullptr;
    uint32_t *d_scratch = nullptr;
    size_t scratch_bytes = 0;
    HitSink sink;  // optional match-position output (apm_plan_set_hit_buffer); base is set per call
};

namespace {

void free_lists(std::vector<Bucket> &buckets, std::vector<SlicedList> &sliced) {
    for (auto &b : buckets) {
        dev_free(b.d_peq);
        dev_free(b.d_group_m);
        dev_free(b.d_group_pat);
    }
    buckets.clear();
    for (auto &l : sliced) {
        dev_free(l.d_codes);
        dev_free(l.d_m);
        dev_free(l.d_id);
        dev_free(l.d_vscratch);
        dev_free(l.d_work);
    }
    sliced.clear();
}

void free_work(apm_plan *pl) {
    free_lists(pl->buckets, pl->sliced);
    free_lists(pl->fb_buckets, pl->fb_sliced);
    FilterSet &f = pl->filter;
    dev_free(f.d_bitmap);
    dev_free(f.d_digest);
    dev_free(f.d_coarse);
    dev_free(f.d_ent_hash);
    dev_free(f.d_ent_idx);
    dev_free(f.d_ent_slot);
    dev_free(f.d_ent_piece);
    dev_free(f.d_fp_id);
    dev_free(f.d_fp_m);
    dev_free(f.d_fp_off);
    dev_free(f.d_cand);
    dev_free(f.d_ctr);
    dev_free(f.d_long);
    dna_free(&f.dna);
    f = FilterSet();
    dev_free(pl->d_tail_list);
    dev_free(pl->d_all_list);
    pl->d_tail_list = pl->d_all_list = nullptr;
    pl->tail_list.clear();
    pl->all_list.clear();
}

int auto_rblock(int NW) { return NW <= 2 ? 4 : (NW <= 4 ? 2 : 1); }

// Kernel lists (row-parallel buckets by word count, window-sliced lists by register-block width) for the
// bit-parallel patterns `ids`, uploaded to the device.
int build_lists(apm_plan *pl, const std::vector<int> &ids_in, std::vector<Bucket> &buckets, std::vector<SlicedList> &sliced) {
    std::vector<std::vector<int>> by_nw(kMaxWords + 1);
    // window-sliced lists: by register-block width (MC = 32 for m <= 32, else 64) and -- in direct mode with the
    // automatic cell choice -- by raggedness, so that lengths that are not a multiple of MC run on the kernels with
    // compile-time block widths (RG = 1 / 2) while multiples of MC stay on the generic kernel
    const bool split = pl->opt.mode == MODE_DIRECT && pl->opt.cell < 0;
    // {MC 32 full, MC 32 ragged, MC 64 full (m = 64), MC 64 ragged (33..63), 65..224 (three CTAs per SM), longer}
    std::vector<int> sliced_ids[6];
    for (int p : ids_in) {
        const int m = (int)pl->pats[p].size();
        const bool sliced_ok = m <= kSlicedMaxLen && pl->nplanes <= kSlicedMaxPlanes &&
                               (pl->opt.kernel == KERNEL_SLICED || pl->opt.kernel == KERNEL_AUTO);
        if (!sliced_ok) by_nw[(m + 31) / 32].push_back(p);
        else if (!split) sliced_ids[m <= 32 ? 0 : 2].push_back(p);
        else if (m <= 32) sliced_ids[m == 32 ? 0 : 1].push_back(p);
        else if (m <= 64) sliced_ids[m == 64 ? 2 : 3].push_back(p);
        else sliced_ids[m <= kSlicedThreeCtaLen ? 4 : 5].push_back(p);
    }
    for (int NW = 1; NW <= kMaxWords; ++NW) {
        auto &ids = by_nw[NW];
        if (ids.empty()) continue;
        std::stable_sort(ids.begin(), ids.end(),
                         [&](int a, int b) { return pl->pats[a].size() < pl->pats[b].size(); });
        Bucket b;
        b.NW = NW;
        b.R = pl->opt.rblock ? pl->opt.rblock : auto_rblock(NW);
        b.EW = entry_words(b.R * NW);
        b.fn = pick_kernel(NW, b.R, pl->opt.variant);
        b.mmin = (int)pl->pats[ids.front()].size();
        b.mmax = (int)pl->pats[ids.back()].size();
        size_t i = 0;
        while (i < ids.size()) {
            const int m = (int)pl->pats[ids[i]].size();
            b.group_m.push_back(m);
            for (int r = 0; r < b.R; ++r) {
                if (i < ids.size() && (int)pl->pats[ids[i]].size() == m) b.group_pat.push_back(ids[i++]);
                else b.group_pat.push_back(-1);
            }
        }
        b.ngroups = (int)b.group_m.size();
        b.peq.assign((size_t)b.ngroups * pl->ncodes * b.EW, 0u);
        for (int g = 0; g < b.ngroups; ++g)
            for (int r = 0; r < b.R; ++r) {
                const int p = b.group_pat[(size_t)g * b.R + r];
                if (p < 0) continue;
                const std::string &s = pl->pats[p];
                for (size_t x = 0; x < s.size(); ++x) {
                    const int c = pl->code_of[(uint8_t)s[x]];
                    b.peq[((size_t)g * pl->ncodes + c) * b.EW + (size_t)r * NW + (x >> 5)] |= 1u << (x & 31);
                }
            }
        int rc;
        if ((rc = upload(&b.d_peq, b.peq))) return rc;
        if ((rc = upload(&b.d_group_m, b.group_m))) return rc;
        if ((rc = upload(&b.d_group_pat, b.group_pat))) return rc;
        buckets.push_back(std::move(b));
    }
    for (int which = 0; which < 6; ++which) {
        auto &ids = sliced_ids[which];
        if (ids.empty()) continue;
        std::stable_sort(ids.begin(), ids.end(),
                         [&](int a, int b) { return pl->pats[a].size() < pl->pats[b].size(); });
        SlicedList l;
        l.MC = which <= 1 ? 32 : 64;
        l.ragged = split && (which == 1 || which == 3) ? 1 : 0;
        l.npat = (int)ids.size();
        l.mmin = (int)pl->pats[ids.front()].size();
        l.mmax = (int)pl->pats[ids.back()].size();
        l.mcp = (l.mmax + 2 + 15) / 16 * 16;  // the sweep reads two symbols ahead
        l.codes.assign((size_t)l.npat * l.mcp + 16, 0);
        for (int i = 0; i < l.npat; ++i) {
            const std::string &s = pl->pats[ids[i]];
            for (size_t x = 0; x < s.size(); ++x) l.codes[(size_t)i * l.mcp + x] = pl->code_of[(uint8_t)s[x]];
            l.m.push_back((int)s.size());
            l.id.push_back(ids[i]);
        }
        int rc2;
        if ((rc2 = upload(&l.d_codes, l.codes))) return rc2;
        if ((rc2 = upload(&l.d_m, l.m))) return rc2;
        if ((rc2 = upload(&l.d_id, l.id))) return rc2;
        sliced.push_back(std::move(l));
    }
    return APM_OK;
}

// Seed tables of the exact filter for the patterns f.ids: the 2-bit q-gram tables of apm_dna.cuh when every pattern
// is pure ACGT (and the q-gram bitmap stays sparse), the hashed tables of apm_filter.cuh otherwise.
int build_filter(apm_plan *pl) {
    FilterSet &f = pl->filter;
    const int k = pl->k;
    f.s = kFilterMaxSeed;
    f.mmax = 0;
    f.mmin = INT_MAX;
    for (int p : f.ids) {
        const int m = (int)pl->pats[p].size();
        f.s = std::min(f.s, m / (k + 1));
        f.mmax = std::max(f.mmax, m);
        f.mmin = std::min(f.mmin, m);
    }
    std::vector<int> fp_id, fp_m;
    std::vector<long long> fp_off;
    long long off = 0;
    std::vector<long long> all_off(pl->P);
    for (int p = 0; p < pl->P; ++p) {
        all_off[p] = off;
        off += (long long)pl->pats[p].size();
    }
    for (size_t slot = 0; slot < f.ids.size(); ++slot) {
        const int p = f.ids[slot];
        fp_id.push_back(p);
        fp_m.push_back((int)pl->pats[p].size());
        fp_off.push_back(all_off[p]);
    }
    int rc;
    if ((rc = upload(&f.d_fp_id, fp_id))) return rc;
    if ((rc = upload(&f.d_fp_m, fp_m))) return rc;
    if ((rc = upload(&f.d_fp_off, fp_off))) return rc;
    CUDA_TRY(dev_alloc((void **)&f.d_ctr, 3 * sizeof(unsigned long long)));
    const size_t cand_bytes = (size_t)std::max<long long>(1, pl->opt.filter_cand_mb) << 20;

    int dq = 0, dh = 0;
    f.use_dna = pl->opt.filter_scan != 1 && dna_choose(pl->pats, f.ids, k, &dq, &dh);
    if (pl->opt.filter_scan == 2 && !f.use_dna)
        return fail(APM_EINVAL, "filter_scan=dna requested but the patterns are not pure ACGT (or too many for a q-gram bitmap)");
    if (f.use_dna) {
        const int erc = dna_build(pl->pats, f.ids, k, dq, dh, pl->num_sms, cand_bytes, &f.dna);
        if (erc) return fail(APM_ECUDA, "building the q-gram tables: %s", cudaGetErrorString((cudaError_t)erc));
        return APM_OK;
    }

    f.nent = (int)f.ids.size() * (k + 1);
    int lg = 0;
    while ((1ll << lg) < f.nent) ++lg;
    f.hb = std::max(kFilterSmemLog + 1, std::min(27, lg + 12));  // <= 1/4096 of the bitmap set: the scan rarely leaves its fast path
    f.bs = 1u;
    for (int i = 0; i < f.s; ++i) f.bs *= kFilterHashB;
    std::vector<uint32_t> digest((size_t)1 << (kFilterSmemLog - 5), 0u);
    struct Ent { uint32_t idx, hash, slot; uint8_t piece; };
    std::vector<Ent> ents;
    ents.reserve((size_t)f.nent);
    for (size_t slot = 0; slot < f.ids.size(); ++slot) {
        const std::string &pat = pl->pats[f.ids[slot]];
        const int m = (int)pat.size();
        for (int i = 0; i <= k; ++i) {
            const int o = filter_piece_offset(i, m, k);
            const uint32_t hash = filter_hash((const uint8_t *)pat.data() + o, f.s);
            const uint32_t idx = filter_index(hash, f.hb);
            digest[filter_digest_word(hash)] |= filter_digest_mask(hash);
            ents.push_back({idx, hash, (uint32_t)slot, (uint8_t)i});
        }
    }
    std::stable_sort(ents.begin(), ents.end(), [](const Ent &a, const Ent &b) { return a.idx < b.idx; });
    std::vector<uint32_t> e_idx, e_slot, e_hash, coarse(((size_t)1 << 16) + 1, 0u);
    std::vector<uint8_t> e_piece;
    for (auto &e : ents) coarse[(e.idx >> (f.hb - 16)) + 1]++;
    for (size_t c = 1; c < coarse.size(); ++c) coarse[c] += coarse[c - 1];
    for (auto &e : ents) {
        e_idx.push_back(e.idx);
        e_hash.push_back(e.hash);
        e_slot.push_back(e.slot);
        e_piece.push_back(e.piece);
    }
    // the full bitmap (2^hb bits, up to 16 MiB) is zeroed and filled on the device
    const size_t bitmap_bytes = (size_t)1 << (f.hb - 3);
    CUDA_TRY(dev_alloc((void **)&f.d_bitmap, bitmap_bytes));
    CUDA_TRY(cudaMemsetAsync(f.d_bitmap, 0, bitmap_bytes, 0));
    if ((rc = upload(&f.d_digest, digest))) return rc;
    if ((rc = upload(&f.d_ent_idx, e_idx))) return rc;
    if ((rc = upload(&f.d_ent_hash, e_hash))) return rc;
    if ((rc = upload(&f.d_coarse, coarse))) return rc;
    filter_bitmap_kernel<<<(f.nent + 255) / 256, 256>>>(f.d_bitmap, f.d_ent_idx, f.nent);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(0));  // plans are used on arbitrary streams afterwards
    g_launches++;
    if ((rc = upload(&f.d_ent_slot, e_slot))) return rc;
    if ((rc = upload(&f.d_ent_piece, e_piece))) return rc;
    f.cap = (unsigned long long)(cand_bytes / sizeof(uint64_t));
    CUDA_TRY(dev_alloc((void **)&f.d_cand, f.cap * sizeof(uint64_t)));
    f.long_cap = std::max<unsigned long long>(1024, f.cap / 8);
    CUDA_TRY(dev_alloc((void **)&f.d_long, f.long_cap * sizeof(uint64_t)));
    return APM_OK;
}

// (Re)build buckets / DP lists for the active patterns (pattern shard) and upload them.
int build_work(apm_plan *pl) {
    free_work(pl);
    std::vector<int> main_ids, filter_ids;
    pl->tail_width = pl->tail_mmax = pl->all_mmax = 0;
    for (int p = 0; p < pl->P; ++p) {
        if (p % pl->shard_world != pl->shard_rank) continue;
        const int m = (int)pl->pats[p].size();
        const bool sliced_ok = m <= kSlicedMaxLen && pl->nplanes <= kSlicedMaxPlanes &&
                               (pl->opt.kernel == KERNEL_SLICED || pl->opt.kernel == KERNEL_AUTO);
        if (pl->opt.kernel == KERNEL_DP || (m > kMaxMyersLen && !sliced_ok)) {
            pl->all_list.push_back(p);
            pl->all_mmax = std::max(pl->all_mmax, m);
            continue;
        }
        // exact seed filter: worth it when every piece of the pattern still holds a selective seed
        const bool filtered = pl->opt.mode == MODE_FILTER && pl->k <= kFilterMaxK && m / (pl->k + 1) >= kFilterMinSeed &&
                              filter_ids.size() < ((size_t)1 << 24);
        (filtered ? filter_ids : main_ids).push_back(p);
        const int tw = m - 1 - pl->k;  // number of truncated tail windows (sequential.c:121,131-134)
        if (tw > 0) {
            pl->tail_list.push_back(p);
            pl->tail_width = std::max(pl->tail_width, tw);
            pl->tail_mmax = std::max(pl->tail_mmax, m);
        }
    }
    // thread-per-window tail kernel for m <= 256 (<= 8 words), warp-per-window for the longer ones
    std::stable_partition(pl->tail_list.begin(), pl->tail_list.end(),
                          [&](int p) { return (int)pl->pats[p].size() <= 32 * kTailThreadMaxWords; });
    pl->tail_nshort = 0;
    pl->tail_width_short = 0;
    for (int p : pl->tail_list) {
        const int m = (int)pl->pats[p].size();
        if (m > 32 * kTailThreadMaxWords) break;
        pl->tail_nshort++;
        pl->tail_width_short = std::max(pl->tail_width_short, m - 1 - pl->k);
    }
    int rc;
    if ((rc = build_lists(pl, main_ids, pl->buckets, pl->sliced))) return rc;
    if (!filter_ids.empty()) {
        if ((rc = build_lists(pl, filter_ids, pl->fb_buckets, pl->fb_sliced))) return rc;
        pl->filter.ids = filter_ids;
        if ((rc = build_filter(pl))) return rc;
    }
    if ((rc = upload(&pl->d_tail_list, pl->tail_list))) return rc;
    if ((rc = upload(&pl->d_all_list, pl->all_list))) return rc;
    return APM_OK;
}

int ensure_scratch(apm_plan *pl, size_t bytes) {
    if (bytes <= pl->scratch_bytes) return APM_OK;
    if (pl->d_scratch) {
        CUDA_TRY(cudaDeviceSynchronize());
        dev_free(pl->d_scratch);
        pl->d_scratch = nullptr;
        pl->scratch_bytes = 0;
    }
    CUDA_TRY(dev_alloc((void **)&pl->d_scratch, bytes));
    pl->scratch_bytes = bytes;
    return APM_OK;
}

int launch_myers(apm_plan *pl, Bucket &b, const uint8_t *d_buf, long long buf_len, long long n_end,
                 long long w0, long long w1, cudaStream_t st, const unsigned int *run_if = nullptr) {
    // windows of the shortest pattern of the bucket that are full-length
    const long long lim = std::min(w1, n_end - b.mmin + 1);
    if (lim <= w0) return APM_OK;
    const long long nwin = lim - w0;  // longer patterns have fewer full windows; the kernel clips per group

    // ---- launch geometry -----------------------------------------------------------------------
    const size_t per_group = (size_t)pl->ncodes * b.EW * 4 + (size_t)(1 + 2 * b.R) * 4;
    const size_t budget = 48 * 1024;
    int gpc_max = (int)std::max<size_t>(1, budget / per_group);
    int tile = pl->opt.tile ? pl->opt.tile : 1024;
    const int cap_guess = pl->num_sms * 4;
    while (!pl->opt.tile && tile > kThreads && (nwin + tile - 1) / tile < cap_guess) tile /= 2;
    long long ntiles = (nwin + tile - 1) / tile;

    int gpc = std::min(b.ngroups, gpc_max);
    size_t smem = myers_smem_bytes(tile, b.mmax, gpc, pl->ncodes, b.R, b.NW);
    CUDA_TRY(ensure_dyn_smem(b.fn, smem));
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, b.fn, kThreads, smem));
    if (occ < 1) return fail(APM_ECUDA, "myers kernel NW=%d R=%d does not fit an SM (smem %zu)", b.NW, b.R, smem);
    const long long capacity = (long long)pl->num_sms * occ;
    // shrink the chunk until every resident CTA gets >= ~8 (chunk, tile) items
    gpc = (int)std::max<long long>(1, std::min<long long>(gpc, ntiles * b.ngroups / (8 * capacity)));
    const long long nitems = ntiles * ((b.ngroups + gpc - 1) / gpc);
    const unsigned gx = (unsigned)std::min<long long>(nitems, capacity), gy = 1;

    MyersArgs a;
    a.buf = d_buf;
    a.buf_len = buf_len;
    a.n_end = n_end;
    a.w0 = w0;
    a.w1 = lim;
    a.peq = b.d_peq;
    a.group_m = b.d_group_m;
    a.group_pat = b.d_group_pat;
    a.code_of = pl->d_code_of;
    a.counts = pl->d_counts;
    a.ngroups = b.ngroups;
    a.ncodes = pl->ncodes;
    a.groups_per_chunk = gpc;
    a.mmax = b.mmax;
    a.k = pl->k;
    a.tile = tile;
    a.c_one = 1u;
    a.c_two = 2u;
    a.run_if = run_if;
    a.sink = pl->sink;
    b.fn<<<dim3(gx, gy), kThreads, smem, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    g_launches++;
    return APM_OK;
}

template <int MC, int CELL, int RG = 0>
int launch_sliced_mc(apm_plan *pl, SlicedList &l, SlicedArgs a, long long nwin, cudaStream_t st) {
    auto fn = sliced_count_kernel<MC, CELL, RG>;
    const long long ntiles = (nwin + kSlicedTile - 1) / kSlicedTile;
    const int rowsU = sliced_rowsU(l.mmax);
    const size_t smem = sliced_smem_bytes(pl->nplanes, rowsU);
    CUDA_TRY(ensure_dyn_smem(fn, smem));
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kSlicedThreads, smem));
    if (occ < 1) return fail(APM_ECUDA, "sliced kernel MC=%d does not fit an SM (smem %zu)", MC, smem);
    const long long capacity = (long long)pl->num_sms * occ;
    // work item = (tile, pattern range), handed out dynamically: ranges of >= 8 patterns (so the ~5 us of
    // tile staging stay ~1 % of an item), >= ~64 items per resident CTA when the job is large enough
    const long long max_splits = std::max<long long>(1, l.npat / 8);
    const long long nsplits = std::max<long long>(1, std::min<long long>(max_splits, (64 * capacity + ntiles - 1) / ntiles));
    const long long nitems = ntiles * nsplits;
    const unsigned gx = (unsigned)std::min<long long>(nitems, capacity);
    if (l.mmax > MC) {  // boundary deltas between the column blocks of long patterns
        const size_t need = (size_t)(l.mmax + 2) * gx * kSlicedThreads * sizeof(uint2);
        if (need > l.vscratch_bytes) {
            if (l.d_vscratch) {
                CUDA_TRY(cudaDeviceSynchronize());
                dev_free(l.d_vscratch);
                l.d_vscratch = nullptr;
                l.vscratch_bytes = 0;
            }
            CUDA_TRY(dev_alloc((void **)&l.d_vscratch, need));
            l.vscratch_bytes = need;
        }
    }
    if (!l.d_work) CUDA_TRY(dev_alloc((void **)&l.d_work, sizeof(unsigned long long)));
    CUDA_TRY(cudaMemsetAsync(l.d_work, 0, sizeof(unsigned long long), st));
    a.work_counter = l.d_work;
    a.vscratch = l.d_vscratch;
    a.nsplits = (int)nsplits;
    a.rowsU = rowsU;
    a.row_bytes = kURowBytes;
    a.row_cols = 32;
    fn<<<gx, kSlicedThreads, smem, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    g_launches++;
    return APM_OK;
}

// smallest instantiated band half-width >= k
using BandKernel = void (*)(const SlicedArgs);
BandKernel pick_band(int k, int cell, int *K_out) {
#define APM_BAND_CASE(KK)                                                                                  \
    if (k <= KK) {                                                                                         \
        *K_out = KK;                                                                                       \
        return cell == 2 ? band_count_kernel<KK, 2> : (cell == 1 ? band_count_kernel<KK, 1> : band_count_kernel<KK, 0>); \
    }
    APM_BAND_CASE(0) APM_BAND_CASE(1) APM_BAND_CASE(2) APM_BAND_CASE(3) APM_BAND_CASE(4) APM_BAND_CASE(5)
    APM_BAND_CASE(6) APM_BAND_CASE(8) APM_BAND_CASE(10) APM_BAND_CASE(12) APM_BAND_CASE(16)
#undef APM_BAND_CASE
    *K_out = -1;
    return nullptr;
}

// shared memory the band kernel needs for this list (its U rows are 2K windows wider than the direct kernel's)
size_t band_smem_need(const apm_plan *pl, const SlicedList &l) {
    int K = -1;
    if (!pick_band(pl->k, 0, &K)) return (size_t)-1;
    const int rowsU = sliced_rowsU(l.mmax + K) + 1;
    return sliced_smem_bytes(pl->nplanes, rowsU, 4 * ((32 + 2 * K) | 1));
}

int launch_band(apm_plan *pl, SlicedList &l, SlicedArgs a, long long nwin, cudaStream_t st) {
    int K = -1;
    BandKernel fn = pick_band(pl->k, pl->opt.cell < 0 ? 0 : pl->opt.cell, &K);  // auto: the band rows are issue bound, plain LOP3 wins
    const long long ntiles = (nwin + kSlicedTile - 1) / kSlicedTile;
    const int rowsU = sliced_rowsU(l.mmax + K) + 1;  // the band reaches K columns past the window; +1 lead row
    const int row_cols = 32 + 2 * K;
    const int row_bytes = 4 * (row_cols | 1);      // odd number of words: conflict-free LDS.32 across lanes
    const size_t smem = sliced_smem_bytes(pl->nplanes, rowsU, row_bytes);
    CUDA_TRY(ensure_dyn_smem(fn, smem));
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kSlicedThreads, smem));
    if (occ < 1) return fail(APM_ECUDA, "band kernel K=%d does not fit an SM (smem %zu)", K, smem);
    const long long capacity = (long long)pl->num_sms * occ;
    const long long max_splits = std::max<long long>(1, l.npat / 32);
    const long long nsplits = std::max<long long>(1, std::min<long long>(max_splits, (64 * capacity + ntiles - 1) / ntiles));
    const unsigned gx = (unsigned)std::min<long long>(ntiles * nsplits, capacity);
    if (!l.d_work) CUDA_TRY(dev_alloc((void **)&l.d_work, sizeof(unsigned long long)));
    CUDA_TRY(cudaMemsetAsync(l.d_work, 0, sizeof(unsigned long long), st));
    a.work_counter = l.d_work;
    a.nsplits = (int)nsplits;
    a.rowsU = rowsU;
    a.row_bytes = row_bytes;
    a.row_cols = row_cols;
    a.lead = 32;
    fn<<<gx, kSlicedThreads, smem, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    g_launches++;
    return APM_OK;
}

int launch_sliced(apm_plan *pl, SlicedList &l, const uint8_t *d_buf, long long buf_len, long long n_end, long long w0,
                  long long w1, cudaStream_t st, const unsigned int *run_if = nullptr) {
    const long long lim = std::min(w1, n_end - l.mmin + 1);
    if (lim <= w0) return APM_OK;
    SlicedArgs a;
    a.buf = d_buf;
    a.buf_len = buf_len;
    a.n_end = n_end;
    a.w0 = w0;
    a.w1 = lim;
    a.pat_codes = l.d_codes;
    a.pat_m = l.d_m;
    a.pat_id = l.d_id;
    a.plane_of = pl->d_plane_of;
    a.counts = pl->d_counts;
    a.vscratch = nullptr;
    a.npat = l.npat;
    a.mcp = l.mcp;
    a.nplanes = pl->nplanes;
    a.k = pl->k;
    a.nsplits = 1;
    a.rowsU = 0;
    a.row_bytes = kURowBytes;
    a.row_cols = 32;
    a.lead = 0;
    a.work_counter = nullptr;
    a.run_if = run_if;
    a.sink = pl->sink;
    a.c_neg1 = 0xFFFFFFFFu;
    // exact band mode: only the 2K+1 diagonals that can matter for D <= k (worth it when the band is
    // narrower than the matrix)
    // (and when its wider U table still fits the SM's shared memory: 8 symbols x long patterns x K >= 12 do not)
    if (pl->opt.mode != MODE_DIRECT && pl->k <= kBandMaxK && 2 * pl->k + 1 < l.mmin &&
        band_smem_need(pl, l) <= (size_t)pl->smem_optin)
        return launch_band(pl, l, a, lim - w0, st);
    // auto (measured on B200, profiles/r02_cell_reuse_*.jsonl): the register-file bound of the co-issued cell decides.
    // CELL 3 (4 LOP3 + 3 IMAD ordered for the operand reuse cache) wins for every list that runs at 3 CTAs per SM
    // (m <= 224 with the 4-symbol DNA alphabet): m = 64 120.5 TCUPS (CELL 1: 114.2), m = 128 116.6 (CELL 0: 109.3), ragged single blocks m = 50 113.0
    // (CELL 1: 102.7); for longer patterns (2 CTAs per SM) plain LOP3 stays ahead (m = 1000: 91.8 vs 86.4).  m = 32:
    // CELL 3 in the two-row sweep (114.9; CELL 2 one row: 110.6); ragged m < 32: CELL 2 (CELL 3 is within +-2 %).
    if (pl->opt.cell < 0) {
        if (l.MC == 32) {
            if (l.ragged == 1) return launch_sliced_mc<32, 2, 1>(pl, l, a, lim - w0, st);
            if (l.mmin == 32 && l.mmax == 32) return launch_sliced_mc<32, 3, 3>(pl, l, a, lim - w0, st);
            return launch_sliced_mc<32, 2>(pl, l, a, lim - w0, st);
        }
        if (l.ragged == 1) return launch_sliced_mc<64, 3, 1>(pl, l, a, lim - w0, st);
        // CTAs per SM of this list (registers allow three; the U table grows with the pattern length and the number of
        // symbol planes).  The longer row recurrence of CELL 3 needs the other warps of the scheduler to hide it: with
        // one CTA per SM (8 symbols) plain LOP3 is 8-15 % ahead, with two only single-block patterns keep CELL 3
        // (tools/alphabet_bench.py: 5 symbols m = 64 115.1 vs 103.5, m = 128 106.9 vs 108.0; 8 symbols 90.6 vs 97.7).
        const size_t cta_smem = sliced_smem_bytes(pl->nplanes, sliced_rowsU(l.mmax)) + 1024;
        const int ctas = (int)std::min<size_t>(3, ((size_t)pl->smem_optin + 1024) / cta_smem);
        const bool cell3 = ctas >= 3 || (ctas == 2 && l.mmax <= 64);
        return cell3 ? launch_sliced_mc<64, 3>(pl, l, a, lim - w0, st) : launch_sliced_mc<64, 0>(pl, l, a, lim - w0, st);
    }
    switch (pl->opt.cell) {
        case 3: return l.MC == 32 ? launch_sliced_mc<32, 3>(pl, l, a, lim - w0, st) : launch_sliced_mc<64, 3>(pl, l, a, lim - w0, st);
        case 2: return l.MC == 32 ? launch_sliced_mc<32, 2>(pl, l, a, lim - w0, st) : launch_sliced_mc<64, 2>(pl, l, a, lim - w0, st);
        case 1: return l.MC == 32 ? launch_sliced_mc<32, 1>(pl, l, a, lim - w0, st) : launch_sliced_mc<64, 1>(pl, l, a, lim - w0, st);
        default: return l.MC == 32 ? launch_sliced_mc<32, 0>(pl, l, a, lim - w0, st) : launch_sliced_mc<64, 0>(pl, l, a, lim - w0, st);
    }
}

// mode=filter: seed scan + verification of the filtered patterns over window starts [w0, w1) (local
// coordinates), in rounds of 2^27 windows; behind every round the same patterns' ordinary kernels are launched
// gated on the round's overflow flag (they run only when the candidate buffer was too small).  Stream ordered.
template <int S>
cudaError_t launch_filter_scan(const FilterArgs &a, unsigned blocks, cudaStream_t st, int device) {
    (void)device;
    const cudaError_t e = ensure_dyn_smem(filter_scan_kernel<S>, kFilterSmemBytes + sizeof(FilterStage) * (kFilterThreads / 32));
    if (e != cudaSuccess) return e;
    filter_scan_kernel<S><<<blocks, kFilterThreads, kFilterSmemBytes + sizeof(FilterStage) * (kFilterThreads / 32), st>>>(a);
    return cudaGetLastError();
}

int launch_filter(apm_plan *pl, const uint8_t *d_buf, long long buf_len, long long n_end, long long w0, long long w1,
                  cudaStream_t st, const uint32_t *d_packed = nullptr) {
    FilterSet &f = pl->filter;
    const long long lim = std::min(w1, n_end - f.mmin + 1);  // full windows only
    if (f.ids.empty() || lim <= w0) return APM_OK;
    if (f.use_dna) {
        // 2-bit q-gram scan + seed-hit verification (apm_dna.cuh): rounds of 2^34 window starts (35-bit positions)
        DnaRun r;
        r.buf = d_buf;
        r.packed = d_packed;
        r.buf_len = buf_len;
        r.n_end = n_end;
        r.k = pl->k;
        r.fp_id = f.d_fp_id;
        r.fp_m = f.d_fp_m;
        r.fp_off = f.d_fp_off;
        r.pat_bytes = pl->d_pat_bytes;
        r.ctr = f.d_ctr;
        r.counts = pl->d_counts;
        r.sink = pl->sink;
        const long long round = 1ll << 34;
        for (long long r0 = w0; r0 < lim; r0 += round) {
            r.w0 = r0;
            r.w1 = std::min(lim, r0 + round);
            int nl = 0;
            CUDA_TRY(dna_launch(f.dna, r, st, &nl));
            g_launches += nl;
            const unsigned int *ovf = reinterpret_cast<unsigned int *>(f.d_ctr + 1);
            int rc;
            for (auto &b : pl->fb_buckets)
                if ((rc = launch_myers(pl, b, d_buf, buf_len, n_end, r.w0, r.w1, st, ovf))) return rc;
            for (auto &l : pl->fb_sliced)
                if ((rc = launch_sliced(pl, l, d_buf, buf_len, n_end, r.w0, r.w1, st, ovf))) return rc;
        }
        return APM_OK;
    }
    FilterArgs a;
    a.buf = d_buf;
    a.buf_len = buf_len;
    a.n_end = n_end;
    a.s = f.s;
    a.k = pl->k;
    a.hb = f.hb;
    a.mmax = f.mmax;
    a.bs = f.bs;
    a.bitmap = f.d_bitmap;
    a.digest = f.d_digest;
    a.ent_idx = f.d_ent_idx;
    a.ent_hash = f.d_ent_hash;
    a.coarse = f.d_coarse;
    a.ent_slot = f.d_ent_slot;
    a.ent_piece = f.d_ent_piece;
    a.nent = f.nent;
    a.fp_id = f.d_fp_id;
    a.fp_m = f.d_fp_m;
    a.fp_off = f.d_fp_off;
    a.pat_bytes = pl->d_pat_bytes;
    a.cand = f.d_cand;
    a.cap = f.cap;
    a.ncand = f.d_ctr;
    a.overflow = reinterpret_cast<unsigned int *>(f.d_ctr + 1);
    a.nlong = f.d_ctr + 2;
    a.longq = f.d_long;
    a.longcap = f.long_cap;
    a.counts = pl->d_counts;
    a.sink = pl->sink;
    const long long round = 1ll << kFilterSlabLog;
    for (long long r0 = w0; r0 < lim; r0 += round) {
        a.w0 = r0;
        a.w1 = std::min(lim, r0 + round);
        CUDA_TRY(cudaMemsetAsync(f.d_ctr, 0, 3 * sizeof(unsigned long long), st));
        const long long positions = a.w1 - a.w0 + f.mmax;
        const long long tiles = (positions + kFilterThreads * kFilterPosPerThread - 1) / (kFilterThreads * kFilterPosPerThread);
        // persistent CTAs (3 per SM: 64 KB digest each), every CTA strides over the tiles
        const unsigned blocks = (unsigned)std::max<long long>(1, std::min<long long>(tiles, (long long)pl->num_sms * 3));
        cudaError_t le;
        switch (f.s) {
            case 8: le = launch_filter_scan<8>(a, blocks, st, pl->device); break;
            case 9: le = launch_filter_scan<9>(a, blocks, st, pl->device); break;
            case 10: le = launch_filter_scan<10>(a, blocks, st, pl->device); break;
            case 11: le = launch_filter_scan<11>(a, blocks, st, pl->device); break;
            case 12: le = launch_filter_scan<12>(a, blocks, st, pl->device); break;
            case 13: le = launch_filter_scan<13>(a, blocks, st, pl->device); break;
            case 14: le = launch_filter_scan<14>(a, blocks, st, pl->device); break;
            case 15: le = launch_filter_scan<15>(a, blocks, st, pl->device); break;
            default: le = launch_filter_scan<16>(a, blocks, st, pl->device); break;
        }
        CUDA_TRY(le);
        filter_verify_kernel<<<pl->num_sms * 16, 128, 0, st>>>(a);
        CUDA_TRY(cudaGetLastError());
        filter_verify_long_kernel<<<pl->num_sms * 8, 128, 0, st>>>(a);
        CUDA_TRY(cudaGetLastError());
        g_launches += 3;
        int rc;
        for (auto &b : pl->fb_buckets)
            if ((rc = launch_myers(pl, b, d_buf, buf_len, n_end, a.w0, a.w1, st, a.overflow))) return rc;
        for (auto &l : pl->fb_sliced)
            if ((rc = launch_sliced(pl, l, d_buf, buf_len, n_end, a.w0, a.w1, st, a.overflow))) return rc;
    }
    return APM_OK;
}

int launch_dp(apm_plan *pl, const uint8_t *d_buf, long long buf_offset, long long n_total, long long j_begin,
              long long j_end, cudaStream_t st) {
    const size_t budget = (size_t)pl->opt.dp_scratch_mb << 20;
    DpArgs a;
    a.buf = d_buf;
    a.buf_offset = buf_offset;
    a.n_total = n_total;
    a.j_begin = j_begin;
    a.j_end = j_end;
    a.pat_bytes = pl->d_pat_bytes;
    a.pat_off = pl->d_pat_off;
    a.pat_len = pl->d_pat_len;
    a.k = pl->k;
    a.counts = pl->d_counts;
    a.sink = pl->sink;
    a.sink.base = 0;  // the DP kernels work on global window starts

    // ---- truncated tail windows of the bit-parallel patterns
    const long long tail_first = std::max<long long>(0, n_total - pl->tail_mmax + 1);
    const bool have_tail = !pl->tail_list.empty() && pl->tail_width > 0 && j_end > tail_first && j_end > j_begin;
    if (have_tail && pl->opt.tail == 0) {
        // bit-parallel (apm_tail.cuh): s column steps per window of size s instead of s^2 cells; exact in every mode
        TailArgs ta;
        ta.d = a;
        ta.code_of = pl->d_code_of;
        ta.ncodes = pl->ncodes;
        if (pl->tail_nshort > 0) {
            ta.d.pat_list = pl->d_tail_list;
            ta.d.npat = pl->tail_nshort;
            ta.chunks = (pl->tail_width_short + kTailThreads - 1) / kTailThreads;
            tail_myers_kernel<<<(unsigned)(pl->tail_nshort * ta.chunks), kTailThreads, 0, st>>>(ta);
            CUDA_TRY(cudaGetLastError());
            g_launches++;
        }
        const int nlong = (int)pl->tail_list.size() - pl->tail_nshort;
        if (nlong > 0) {
            constexpr int kWarps = kTailWarpThreads / 32;
            ta.d.pat_list = pl->d_tail_list + pl->tail_nshort;
            ta.d.npat = nlong;
            ta.chunks = (pl->tail_width + kWarps - 1) / kWarps;
            const long long blocks = (long long)nlong * ta.chunks;
            if (blocks > 0x7FFFFFFFll) return fail(APM_EINVAL, "too many long patterns for one tail launch");
            tail_myers_warp_kernel<<<(unsigned)blocks, kTailWarpThreads, (size_t)pl->ncodes * 32 * sizeof(uint32_t), st>>>(ta);
            CUDA_TRY(cudaGetLastError());
            g_launches++;
        }
    } else if (have_tail) {
        const size_t per_thread = (size_t)(pl->tail_mmax + 1) * 4;
        long long max_threads = std::max<long long>(pl->tail_width, (long long)(budget / per_thread));
        int pats_per_launch = (int)std::max<long long>(1, max_threads / pl->tail_width);
        for (size_t p0 = 0; p0 < pl->tail_list.size(); p0 += pats_per_launch) {
            const int np = (int)std::min<size_t>(pats_per_launch, pl->tail_list.size() - p0);
            const long long threads = (long long)np * pl->tail_width;
            const unsigned blocks = (unsigned)((threads + 127) / 128);
            const long long stride = (long long)blocks * 128;
            const bool banded = pl->opt.mode != MODE_DIRECT && pl->k <= kBandDpMaxK;
            int rc = banded ? APM_OK : ensure_scratch(pl, (size_t)stride * per_thread);
            if (rc) return rc;
            a.pat_list = pl->d_tail_list + p0;
            a.npat = np;
            a.tail_width = pl->tail_width;
            a.scratch = pl->d_scratch;
            a.scratch_stride = stride;
            // direct mode: every cell of the (truncated) window; band / filter mode: only the band that decides
            if (banded) dp_tail_band_kernel<<<blocks, 128, 0, st>>>(a);
            else dp_tail_kernel<<<blocks, 128, 0, st>>>(a);
            CUDA_TRY(cudaGetLastError());
            g_launches++;
        }
    }
    // ---- patterns evaluated entirely by the DP kernel (too long for the bit-parallel kernel, or
    //      option kernel=dp)
    if (!pl->all_list.empty() && j_end > j_begin) {
        const size_t per_thread = (size_t)(pl->all_mmax + 1) * 4;
        const long long nwin = j_end - j_begin;
        long long max_threads = std::max<long long>(128, (long long)(budget / per_thread));
        long long nx_blocks = std::min<long long>((nwin + 127) / 128, std::max<long long>(1, max_threads / 128));
        nx_blocks = std::min<long long>(nx_blocks, (long long)pl->num_sms * 16);
        const long long nx = nx_blocks * 128;
        int pats_per_launch = (int)std::max<long long>(1, std::min<long long>(65535, max_threads / nx));
        for (size_t p0 = 0; p0 < pl->all_list.size(); p0 += pats_per_launch) {
            const int np = (int)std::min<size_t>(pats_per_launch, pl->all_list.size() - p0);
            const long long stride = nx * np;
            int rc = ensure_scratch(pl, (size_t)stride * per_thread);
            if (rc) return rc;
            a.pat_list = pl->d_all_list + p0;
            a.npat = np;
            a.tail_width = 0;
            a.scratch = pl->d_scratch;
            a.scratch_stride = stride;
            dp_all_kernel<<<dim3((unsigned)nx_blocks, (unsigned)np), 128, 0, st>>>(a);
            CUDA_TRY(cudaGetLastError());
            g_launches++;
        }
    }
    return APM_OK;
}

int check_patterns(const char *const *patterns, const int *pattern_len, int nb_patterns, int approx_factor) {
    if (approx_factor < 0) return fail(APM_EINVAL, "approx_factor must be >= 0 (got %d)", approx_factor);
    if (nb_patterns < 0) return fail(APM_EINVAL, "nb_patterns must be >= 0");
    if (nb_patterns > 0 && (!patterns || !pattern_len)) return fail(APM_EINVAL, "patterns / pattern_len is NULL");
    for (int i = 0; i < nb_patterns; ++i) {
        if (!patterns[i] || pattern_len[i] <= 0)
            return fail(APM_EINVAL, "pattern %d is empty (the reference rejects it too: sequential.c:65)", i);
    }
    return APM_OK;
}

}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

const char *apm_last_error(void) { return tl_err.c_str(); }
int apm_release_cache(void) {
    release_device_cache();
    std::lock_guard<std::mutex> lk(g_pool.mu);
    for (uint8_t *p : g_pool.pinned_free) cudaFreeHost(p);  // buffers an ingest is using are not in this list
    g_pool.pinned_free.clear();
    return APM_OK;
}

const char *apm_version(void) { return "apm_b200 0.1 (sm_100a)"; }
unsigned long long apm_launch_count(void) { return g_launches.load(); }

int apm_device_count(int *count) {
    if (!count) return fail(APM_EINVAL, "count is NULL");
    int n = 0;
    int rc = device_ready(&n);
    *count = rc ? 0 : n;
    return rc;
}

int apm_set_device(int device) {
    int n = 0;
    int rc = device_ready(&n);
    if (rc) return rc;
    if (device < 0 || device >= n) return fail(APM_EINVAL, "device %d out of range (0..%d)", device, n - 1);
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaFree(nullptr));  // create the primary context now, not inside the first timed call
    return APM_OK;
}

int apm_set_option(const char *key, const char *value) {
    if (!key || !value) return fail(APM_EINVAL, "key/value is NULL");
    std::lock_guard<std::mutex> lk(g_opt_mu);
    const std::string k = key, v = value;
    auto bad = [&]() { return fail(APM_EINVAL, "bad value '%s' for option '%s'", value, key); };
    if (k == "gpus") {
        if (v == "all") g_opt.gpus = 0;
        else {
            int n = atoi(value);
            if (n < 1 || n > 64) return bad();
            g_opt.gpus = n;
        }
    } else if (k == "shard") {
        if (v == "auto") g_opt.shard = SHARD_AUTO;
        else if (v == "db" || v == "DB_OVER_RANKS") g_opt.shard = SHARD_DB;
        else if (v == "patterns" || v == "PATTERNS_OVER_RANKS") g_opt.shard = SHARD_PATTERNS;
        else return bad();
    } else if (k == "kernel") {
        if (v == "auto") g_opt.kernel = KERNEL_AUTO;
        else if (v == "myers" || v == "rows") g_opt.kernel = KERNEL_MYERS;
        else if (v == "sliced") g_opt.kernel = KERNEL_SLICED;
        else if (v == "dp") g_opt.kernel = KERNEL_DP;
        else return bad();
    } else if (k == "mode") {
        if (v == "direct") g_opt.mode = MODE_DIRECT;
        else if (v == "band") g_opt.mode = MODE_BAND;
        else if (v == "filter") g_opt.mode = MODE_FILTER;
        else return bad();
    } else if (k == "rblock") {
        if (v == "auto") g_opt.rblock = 0;
        else if (v == "1" || v == "2" || v == "4") g_opt.rblock = atoi(value);
        else return bad();
    } else if (k == "tile") {
        if (v == "auto") g_opt.tile = 0;
        else {
            int t = atoi(value);
            if (t < kThreads || t % kThreads || t > 16384) return bad();
            g_opt.tile = t;
        }
    } else if (k == "variant") {
        if (v == "0" || v == "1" || v == "2") g_opt.variant = atoi(value);
        else return bad();
    } else if (k == "cell") {
        if (v == "auto") g_opt.cell = -1;
        else if (v == "lop3") g_opt.cell = 0;
        else if (v == "fma3") g_opt.cell = 1;
        else if (v == "fma" || v == "fma2") g_opt.cell = 2;
        else if (v == "fma3r") g_opt.cell = 3;
        else return bad();
    } else if (k == "filter_scan") {
        if (v == "auto") g_opt.filter_scan = 0;
        else if (v == "hash") g_opt.filter_scan = 1;
        else if (v == "dna") g_opt.filter_scan = 2;
        else return bad();
    } else if (k == "tail") {
        if (v == "auto" || v == "bitpar") g_opt.tail = 0;
        else if (v == "dp") g_opt.tail = 1;
        else return bad();
    } else if (k == "reduce") {
        if (v == "auto") g_opt.reduce = 0;
        else if (v == "nccl") g_opt.reduce = 1;
        else if (v == "host") g_opt.reduce = 2;
        else if (v == "p2p") g_opt.reduce = 3;
        else return bad();
    } else if (k == "text_chunk_mb") {
        long long mb = atoll(value);
        if (mb < 1 || mb > (1ll << 24)) return bad();
        g_opt.text_chunk_mb = mb;
    } else if (k == "filter_cand_mb") {
        long long mb = atoll(value);
        if (mb < 1 || mb > 16384) return bad();
        g_opt.filter_cand_mb = mb;
    } else if (k == "cache_mb") {
        long long mb = atoll(value);
        if (v.empty() || v.find_first_not_of("0123456789") != std::string::npos || mb > (1ll << 20)) return bad();
        g_opt.cache_mb = mb;
    } else if (k == "ingest_threads") {
        if (v == "auto") g_opt.ingest_threads = 0;
        else {
            int n = atoi(value);
            if (n < 1 || n > 64) return bad();
            g_opt.ingest_threads = n;
        }
    } else if (k == "dp_scratch_mb") {
        long long mb = atoll(value);
        if (mb < 1 || mb > 65536) return bad();
        g_opt.dp_scratch_mb = mb;
    } else {
        return fail(APM_EINVAL, "unknown option '%s'", key);
    }
    return APM_OK;
}

const char *apm_get_option(const char *key) {
    if (!key) return nullptr;
    const Options o = options_snapshot();
    const std::string k = key;
    if (k == "gpus") tl_optbuf = o.gpus ? std::to_string(o.gpus) : "all";
    else if (k == "shard") tl_optbuf = o.shard == SHARD_DB ? "db" : (o.shard == SHARD_PATTERNS ? "patterns" : "auto");
    else if (k == "kernel")
        tl_optbuf = o.kernel == KERNEL_DP ? "dp" : (o.kernel == KERNEL_MYERS ? "myers" : (o.kernel == KERNEL_SLICED ? "sliced" : "auto"));
    else if (k == "mode") tl_optbuf = o.mode == MODE_FILTER ? "filter" : (o.mode == MODE_BAND ? "band" : "direct");
    else if (k == "rblock") tl_optbuf = o.rblock ? std::to_string(o.rblock) : "auto";
    else if (k == "tile") tl_optbuf = o.tile ? std::to_string(o.tile) : "auto";
    else if (k == "variant") tl_optbuf = std::to_string(o.variant);
    else if (k == "cell") tl_optbuf = o.cell == 3 ? "fma3r" : (o.cell == 2 ? "fma" : (o.cell == 1 ? "fma3" : (o.cell == 0 ? "lop3" : "auto")));
    else if (k == "filter_scan") tl_optbuf = o.filter_scan == 1 ? "hash" : (o.filter_scan == 2 ? "dna" : "auto");
    else if (k == "tail") tl_optbuf = o.tail == 1 ? "dp" : "bitpar";
    else if (k == "reduce") tl_optbuf = o.reduce == 1 ? "nccl" : (o.reduce == 2 ? "host" : (o.reduce == 3 ? "p2p" : "auto"));
    else if (k == "text_chunk_mb") tl_optbuf = std::to_string(o.text_chunk_mb);
    else if (k == "filter_cand_mb") tl_optbuf = std::to_string(o.filter_cand_mb);
    else if (k == "cache_mb") tl_optbuf = std::to_string(o.cache_mb);
    else if (k == "ingest_threads") tl_optbuf = o.ingest_threads ? std::to_string(o.ingest_threads) : "auto";
    else if (k == "dp_scratch_mb") tl_optbuf = std::to_string(o.dp_scratch_mb);
    else return nullptr;
    return tl_optbuf.c_str();
}

// ---------------------------------------------------------------------------------------------------
int apm_plan_create(const char *const *patterns, const int *pattern_len, int nb_patterns, int approx_factor,
                    apm_plan **plan_out) {
    if (!plan_out) return fail(APM_EINVAL, "plan_out is NULL");
    *plan_out = nullptr;
    int rc = check_patterns(patterns, pattern_len, nb_patterns, approx_factor);
    if (rc) return rc;
    int ndev = 0;
    if ((rc = device_ready(&ndev))) return rc;

    apm_plan *pl = new (std::nothrow) apm_plan();
    if (!pl) return fail(APM_ENOMEM, "out of host memory");
    pl->opt = options_snapshot();
    pl->P = nb_patterns;
    pl->k = approx_factor;
    cudaError_t e = cudaGetDevice(&pl->device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&pl->num_sms, cudaDevAttrMultiProcessorCount, pl->device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&pl->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, pl->device);
    if (e != cudaSuccess) {
        delete pl;
        return fail(APM_ECUDA, "cudaGetDevice/attribute: %s", cudaGetErrorString(e));
    }
    // per-call compact alphabet: codes 0..A-1 for the bytes that occur in some pattern, code A =
    // "any other byte" (its Peq is all-zero); with all 256 bytes in use there is no "other".
    bool used[256] = {false};
    std::vector<uint8_t> flat;
    std::vector<long long> off(nb_patterns);
    std::vector<int> len(nb_patterns);
    for (int i = 0; i < nb_patterns; ++i) {
        pl->pats.emplace_back(patterns[i], (size_t)pattern_len[i]);
        off[i] = (long long)flat.size();
        len[i] = pattern_len[i];
        for (int x = 0; x < pattern_len[i]; ++x) {
            used[(uint8_t)patterns[i][x]] = true;
            flat.push_back((uint8_t)patterns[i][x]);
        }
        pl->mmax_all_patterns = std::max(pl->mmax_all_patterns, pattern_len[i]);
    }
    int A = 0;
    for (int b = 0; b < 256; ++b)
        if (used[b]) pl->code_of[b] = (uint8_t)A++;
    pl->ncodes = A < 256 ? A + 1 : 256;
    for (int b = 0; b < 256; ++b)
        if (!used[b]) pl->code_of[b] = (uint8_t)A;  // only reachable when A < 256

    auto cleanup_fail = [&](int code) {
        apm_plan_destroy(pl);
        return code;
    };
    std::vector<uint8_t> map(pl->code_of, pl->code_of + 256);
    if ((rc = upload(&pl->d_code_of, map))) return cleanup_fail(rc);
    pl->nplanes = A;  // one occurrence plane per pattern symbol (window-sliced kernel)
    std::vector<uint8_t> planes(256);
    for (int b = 0; b < 256; ++b) planes[b] = used[b] ? pl->code_of[b] : kNoPlane;
    if ((rc = upload(&pl->d_plane_of, planes))) return cleanup_fail(rc);
    if ((rc = upload(&pl->d_pat_bytes, flat))) return cleanup_fail(rc);
    if ((rc = upload(&pl->d_pat_off, off))) return cleanup_fail(rc);
    if ((rc = upload(&pl->d_pat_len, len))) return cleanup_fail(rc);
    e = dev_alloc((void **)&pl->d_counts, sizeof(unsigned long long) * std::max(1, nb_patterns));
    if (e == cudaSuccess) e = cudaMemset(pl->d_counts, 0, sizeof(unsigned long long) * std::max(1, nb_patterns));
    if (e != cudaSuccess) {
        fail(APM_ECUDA, "allocating counters: %s", cudaGetErrorString(e));
        return cleanup_fail(APM_ECUDA);
    }
    if ((rc = build_work(pl))) return cleanup_fail(rc);
    // the tables and the zeroed counters were written on the legacy stream; the plan is used on arbitrary
    // (non-blocking) streams afterwards
    if (cudaStreamSynchronize(0) != cudaSuccess) return cleanup_fail(fail(APM_ECUDA, "cudaStreamSynchronize failed after the plan uploads"));
    *plan_out = pl;
    return APM_OK;
}

int apm_plan_destroy(apm_plan *pl) {
    if (!pl) return APM_OK;
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(pl->device);
    cudaDeviceSynchronize();  // the blocks go back to the cache: nothing in flight may still use them
    free_work(pl);
    dev_free(pl->d_code_of);
    dev_free(pl->d_plane_of);
    dev_free(pl->d_pat_bytes);
    dev_free(pl->d_pat_off);
    dev_free(pl->d_pat_len);
    dev_free(pl->d_counts);
    dev_free(pl->d_scratch);
    cudaSetDevice(cur);
    delete pl;
    return APM_OK;
}

int apm_plan_set_pattern_shard(apm_plan *pl, int rank, int world) {
    if (!pl) return fail(APM_EINVAL, "plan is NULL");
    if (world < 1 || rank < 0 || rank >= world) return fail(APM_EINVAL, "bad pattern shard %d/%d", rank, world);
    CUDA_TRY(cudaSetDevice(pl->device));
    CUDA_TRY(cudaDeviceSynchronize());
    pl->shard_rank = rank;
    pl->shard_world = world;
    const int rc = build_work(pl);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(0));
    return APM_OK;
}

int apm_plan_set_hit_buffer(apm_plan *pl, unsigned long long *d_hits, unsigned long long capacity,
                            unsigned long long *d_n_hits) {
    if (!pl) return fail(APM_EINVAL, "plan is NULL");
    if ((d_hits == nullptr) != (d_n_hits == nullptr) || (d_hits && capacity == 0))
        return fail(APM_EINVAL, "hit buffer, its capacity and its counter go together");
    if (pl->P >= (1 << (64 - kHitPosBits))) return fail(APM_EINVAL, "too many patterns for packed hit entries");
    pl->sink.buf = d_hits;
    pl->sink.cap = d_hits ? capacity : 0;
    pl->sink.count = d_n_hits;
    return APM_OK;
}

int apm_plan_zero_counts(apm_plan *pl, void *stream) {
    if (!pl) return fail(APM_EINVAL, "plan is NULL");
    CUDA_TRY(cudaMemsetAsync(pl->d_counts, 0, sizeof(unsigned long long) * std::max(1, pl->P), (cudaStream_t)stream));
    return APM_OK;
}

int apm_plan_counts_device_ptr(apm_plan *pl, unsigned long long **d_counts) {
    if (!pl || !d_counts) return fail(APM_EINVAL, "NULL argument");
    *d_counts = pl->d_counts;
    return APM_OK;
}

int apm_plan_read_counts(apm_plan *pl, long long *n_matches, void *stream) {
    if (!pl || (!n_matches && pl->P > 0)) return fail(APM_EINVAL, "NULL argument");
    if (pl->P == 0) return APM_OK;
    CUDA_TRY(cudaMemcpyAsync(n_matches, pl->d_counts, sizeof(long long) * pl->P, cudaMemcpyDeviceToHost,
                             (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return APM_OK;
}

int apm_plan_max_pattern_len(apm_plan *pl, int *m_max) {
    if (!pl || !m_max) return fail(APM_EINVAL, "NULL argument");
    *m_max = pl->mmax_all_patterns;
    return APM_OK;
}

namespace {
int plan_count_device_impl(apm_plan *pl, const unsigned char *d_buf, const uint32_t *d_packed, unsigned long long buf_offset,
                           unsigned long long buf_len, unsigned long long n_total, unsigned long long j_begin,
                           unsigned long long j_end, void *stream) {
    if (!pl) return fail(APM_EINVAL, "plan is NULL");
    if (n_total > (1ull << 62) || buf_offset > n_total || buf_len > n_total - buf_offset)
        return fail(APM_EINVAL, "buffer [%llu, +%llu) is not inside the text of %llu bytes", buf_offset, buf_len, n_total);
    cudaStream_t st = (cudaStream_t)stream;
    const long long N = (long long)n_total;
    long long je = (long long)std::min<unsigned long long>(j_end, n_total);
    je = std::min(je, N - (long long)pl->k);  // sequential.c:121
    const long long jb = (long long)j_begin;
    if (pl->P == 0 || je <= jb) return APM_OK;
    // bytes the windows [jb, je) may touch
    const long long need_end = std::min(N, je + (long long)pl->mmax_all_patterns - 1);
    if ((long long)buf_offset > jb || (long long)(buf_offset + buf_len) < need_end)
        return fail(APM_EINVAL,
                    "device buffer [%llu, %llu) does not cover window starts [%lld, %lld) plus the %d-byte halo",
                    buf_offset, buf_offset + buf_len, jb, je, pl->mmax_all_patterns - 1);
    if (!d_buf) return fail(APM_EINVAL, "d_buf is NULL");
    int cur = -1;
    CUDA_TRY(cudaGetDevice(&cur));
    if (cur != pl->device) CUDA_TRY(cudaSetDevice(pl->device));
    pl->sink.base = (long long)buf_offset;  // the bit-parallel kernels work on buffer-local window starts
    int rc = APM_OK;
    for (auto &b : pl->buckets) {
        rc = launch_myers(pl, b, d_buf, (long long)buf_len, N - (long long)buf_offset, jb - (long long)buf_offset,
                          je - (long long)buf_offset, st);
        if (rc) break;
    }
    for (auto &l : pl->sliced) {
        if (rc) break;
        rc = launch_sliced(pl, l, d_buf, (long long)buf_len, N - (long long)buf_offset, jb - (long long)buf_offset,
                           je - (long long)buf_offset, st);
    }
    if (!rc)
        rc = launch_filter(pl, d_buf, (long long)buf_len, N - (long long)buf_offset, jb - (long long)buf_offset,
                           je - (long long)buf_offset, st, d_packed);
    if (!rc) rc = launch_dp(pl, d_buf, (long long)buf_offset, N, jb, je, st);
    if (cur != pl->device) cudaSetDevice(cur);
    return rc;
}
}  // namespace

int apm_plan_count_device(apm_plan *pl, const unsigned char *d_buf, unsigned long long buf_offset,
                          unsigned long long buf_len, unsigned long long n_total, unsigned long long j_begin,
                          unsigned long long j_end, void *stream) {
    return plan_count_device_impl(pl, d_buf, nullptr, buf_offset, buf_len, n_total, j_begin, j_end, stream);
}

// ---- resident 2-bit copy of a device text (filter mode, ACGT pattern sets): packed once, scanned many times
unsigned long long apm_text_pack_bytes(unsigned long long buf_len) { return 4ull * dna_pack_words(buf_len); }

int apm_text_pack_device(const unsigned char *d_buf, unsigned long long buf_len, void *d_packed, void *stream) {
    int ndev = 0;
    if (int rc = device_ready(&ndev)) return rc;
    if (!d_buf || !d_packed) return fail(APM_EINVAL, "d_buf / d_packed is NULL");
    if ((reinterpret_cast<uintptr_t>(d_buf) & 15) || (reinterpret_cast<uintptr_t>(d_packed) & 3))
        return fail(APM_EINVAL, "apm_text_pack_device: d_buf must be 16-byte aligned, d_packed 4-byte aligned");
    CUDA_TRY(dna_pack_text(d_buf, buf_len, reinterpret_cast<uint32_t *>(d_packed), (cudaStream_t)stream));
    g_launches++;
    return APM_OK;
}

int apm_plan_count_device_packed(apm_plan *pl, const unsigned char *d_buf, const void *d_packed,
                                 unsigned long long buf_offset, unsigned long long buf_len, unsigned long long n_total,
                                 unsigned long long j_begin, unsigned long long j_end, void *stream) {
    if (d_packed && (reinterpret_cast<uintptr_t>(d_buf) & 15))
        return fail(APM_EINVAL, "apm_plan_count_device_packed: d_buf must be the 16-byte aligned buffer the copy was packed from");
    return plan_count_device_impl(pl, d_buf, reinterpret_cast<const uint32_t *>(d_packed), buf_offset, buf_len, n_total,
                                  j_begin, j_end, stream);
}

// ---------------------------------------------------------------------------------------------------
int apm_synth_text_device(unsigned char *d_out, unsigned long long seed, unsigned long long offset,
                          unsigned long long count, void *stream) {
    int ndev = 0;
    int rc = device_ready(&ndev);
    if (rc) return rc;
    if (count == 0) return APM_OK;
    if (!d_out) return fail(APM_EINVAL, "d_out is NULL");
    const unsigned long long nvec = (count + 15) / 16;
    const unsigned blocks = (unsigned)std::min<unsigned long long>((nvec + 255) / 256, 148ull * 32);
    synth_text_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_out, seed, offset, count);
    CUDA_TRY(cudaGetLastError());
    g_launches++;
    return APM_OK;
}

int apm_int_peak(int kind, double *ops_per_sec, double *seconds) {
    int ndev = 0;
    int rc = device_ready(&ndev);
    if (rc) return rc;
    if (kind < 0 || kind > 9 || !ops_per_sec) return fail(APM_EINVAL, "bad argument");
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    uint32_t *d_out = nullptr;
    CUDA_TRY(dev_alloc((void **)&d_out, 64));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int iters = kind == 9 ? 1024 : 4096, blocks = sms * 8;
    const double ops_per_thread_iter = int_peak_ops_per_iter(kind);
    float best_ms = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {  // rep 0 = warm-up
        CUDA_TRY(cudaEventRecord(e0));
#define APM_PEAK_CASE(K) case K: int_peak_kernel<K><<<blocks, 256>>>(d_out, iters, 0x9E3779B9u + rep, 0x7F4A7C15u); break;
        switch (kind) {
            APM_PEAK_CASE(0) APM_PEAK_CASE(1) APM_PEAK_CASE(2) APM_PEAK_CASE(3) APM_PEAK_CASE(4)
            APM_PEAK_CASE(5) APM_PEAK_CASE(6) APM_PEAK_CASE(7) APM_PEAK_CASE(8)
            default: int_peak_kernel<9><<<blocks, 256>>>(d_out, iters, 0x9E3779B9u + rep, 0x7F4A7C15u); break;
        }
#undef APM_PEAK_CASE
        CUDA_TRY(cudaGetLastError());
        g_launches++;
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best_ms = std::min(best_ms, ms);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    dev_free(d_out);
    const double total_ops = ops_per_thread_iter * iters * 256.0 * blocks;
    *ops_per_sec = total_ops / (best_ms * 1e-3);
    if (seconds) *seconds = best_ms * 1e-3;
    return APM_OK;
}

// ---------------------------------------------------------------------------------------------------
// One-shot host API
// ---------------------------------------------------------------------------------------------------
namespace {

// ---- NCCL, loaded at run time (no link-time dependency): the single-process multi-GPU path sums the
//      per-GPU count vectors with one in-place ncclAllReduce per device inside a group, over NVLink.
//      This replaces the MPI_Send/MPI_Recv of per-pattern ints of the reference
//      (patterns_over_ranks.c:195,389; database_over_ranks.c:179,573).
struct Nccl {
    typedef void *comm_t;
    int (*CommInitAll)(comm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::vector<int> devs;        // devices of the cached communicators
    std::vector<comm_t> comms;
};
std::mutex g_nccl_mu;
Nccl g_nccl;
bool g_nccl_tried = false;

bool nccl_load() {
    if (g_nccl_tried) return g_nccl.ok;
    g_nccl_tried = true;
    void *h = nullptr;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
        h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return false;
    g_nccl.CommInitAll = (int (*)(Nccl::comm_t *, int, const int *))dlsym(h, "ncclCommInitAll");
    g_nccl.CommDestroy = (int (*)(Nccl::comm_t))dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, Nccl::comm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
    g_nccl.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    g_nccl.ok = g_nccl.CommInitAll && g_nccl.AllReduce && g_nccl.GroupStart && g_nccl.GroupEnd;
    return g_nccl.ok;
}

// communicators for exactly this device list (cached across calls; NCCL init costs ~100 ms)
int nccl_comms_for(const std::vector<int> &devs) {
    if (g_nccl.devs == devs && !g_nccl.comms.empty()) return APM_OK;
    if (g_nccl.CommDestroy)
        for (auto c : g_nccl.comms) g_nccl.CommDestroy(c);
    g_nccl.comms.assign(devs.size(), nullptr);
    g_nccl.devs.clear();
    const int rc = g_nccl.CommInitAll(g_nccl.comms.data(), (int)devs.size(), devs.data());
    if (rc != 0) {
        g_nccl.comms.clear();
        return fail(APM_ECUDA, "ncclCommInitAll: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
    }
    g_nccl.devs = devs;
    return APM_OK;
}

struct DevJob {
    int dev = 0;
    cudaStream_t st = nullptr, copy_st = nullptr;  // count kernels / H2D copies of the file ingest
    apm_plan *plan = nullptr;
    uint8_t *d_text = nullptr, *d_text2 = nullptr;  // second buffer: segmented shards (text_chunk_mb)
    cudaEvent_t seg_done[2] = {nullptr, nullptr};   // counting on buffer b has finished
    unsigned long long *d_hits = nullptr;  // [0] counter, [1..] packed hits (apm_find_matches)
    long long j0 = 0, j1 = 0, b0 = 0, b1 = 0;  // window-start range and byte range (global)
};

void release_jobs(std::vector<DevJob> &jobs, int restore_dev) {
    for (auto &j : jobs) {
        cudaSetDevice(j.dev);
        if (j.copy_st) cudaStreamSynchronize(j.copy_st);  // an in-flight H2D copy must not outlive its target block
        if (j.st) cudaStreamSynchronize(j.st);
        if (j.plan) apm_plan_destroy(j.plan);
        if (j.d_text) dev_free(j.d_text);
        if (j.d_text2) dev_free(j.d_text2);
        for (auto &e : j.seg_done)
            if (e) cudaEventDestroy(e);
        if (j.d_hits) dev_free(j.d_hits);
        if (j.st) cudaStreamDestroy(j.st);
        if (j.copy_st) cudaStreamDestroy(j.copy_st);
    }
    jobs.clear();
    cudaSetDevice(restore_dev);
}

// Text source: either a host buffer or a file descriptor.
struct TextSource {
    const unsigned char *host = nullptr;
    int fd = -1;
};

// Loads the global bytes [b0, b1) into d_dst and tells the caller, chunk by chunk, how far the text has arrived.
//
// Small or already page-locked host buffers: one async copy.  Everything else -- files, and pageable host buffers of
// more than a few MiB -- goes through the chunked ingest: R reader threads per GPU, each with two pinned 16 MiB
// staging buffers and its own copy stream, take the chunks round-robin (pread from the page cache / disk, or a
// memcpy out of the caller's pageable buffer, then cudaMemcpyAsync + an event per chunk).  The calling thread
// consumes the chunks IN ORDER: it makes the count stream wait for the chunk's event and calls on_chunk(bytes_end),
// which launches the counting of the windows that are complete while the readers are already several chunks ahead
// (ingest overlapped with counting; replaces read_input_file of utils.c:12-68).
// `wait_before`: optional event every copy stream waits for first (the target buffer is still being counted).
struct IngestShared {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<int> state;  // per chunk: 0 pending, 1 copy enqueued (its event is recorded), -1 failed
    std::atomic<bool> abort{false};
    std::string err;
};

int ingest_reader_count(const Options &opt, int gpus_in_flight) {
    if (opt.ingest_threads > 0) return opt.ingest_threads;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    return (int)std::max(1u, std::min(8u, hw / (unsigned)std::max(1, gpus_in_flight)));
}

bool host_pointer_is_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

extern "C++" template <typename OnChunk>
int copy_range_to_device(const TextSource &src, long long b0, long long b1, uint8_t *d_dst, cudaStream_t st,
                         cudaStream_t copy_st, cudaEvent_t wait_before, int readers, OnChunk on_chunk) {
    if (b1 <= b0) return APM_OK;
    const long long total = b1 - b0;
    if (src.host && (total <= (long long)(8 << 20) || host_pointer_is_pinned(src.host + b0))) {
        // one async copy on the copy stream; the count stream waits for it
        if (wait_before) CUDA_TRY(cudaStreamWaitEvent(copy_st, wait_before, 0));
        CUDA_TRY(cudaMemcpyAsync(d_dst, src.host + b0, (size_t)total, cudaMemcpyHostToDevice, copy_st));
        cudaEvent_t ev;
        CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        cudaError_t e = cudaEventRecord(ev, copy_st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ev, 0);
        cudaEventDestroy(ev);
        if (e != cudaSuccess) return fail(APM_ECUDA, "H2D copy: %s", cudaGetErrorString(e));
        return on_chunk(b1);
    }
    const long long chunk = (long long)kPinnedChunk;
    const long long nchunks = (total + chunk - 1) / chunk;
    const int R = (int)std::max<long long>(1, std::min<long long>(readers, nchunks));
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    std::vector<cudaEvent_t> events((size_t)nchunks, nullptr);
    for (auto &ev : events)
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
            for (auto &e2 : events)
                if (e2) cudaEventDestroy(e2);
            return fail(APM_ECUDA, "cudaEventCreate failed");
        }
    IngestShared sh;
    sh.state.assign((size_t)nchunks, 0);
    auto reader = [&](int r) {
        cudaStream_t rs = nullptr;
        uint8_t *pin[2] = {nullptr, nullptr};
        long long last[2] = {-1, -1};
        auto failed = [&](long long c, const std::string &why) {
            std::lock_guard<std::mutex> lk(sh.mu);
            if (sh.err.empty()) sh.err = why;
            sh.state[(size_t)c] = -1;
            sh.abort = true;
            sh.cv.notify_all();
        };
        bool ok = cudaSetDevice(dev) == cudaSuccess && cudaStreamCreateWithFlags(&rs, cudaStreamNonBlocking) == cudaSuccess;
        if (ok && wait_before) ok = cudaStreamWaitEvent(rs, wait_before, 0) == cudaSuccess;
        if (ok) {
            pin[0] = pinned_acquire();
            pin[1] = pinned_acquire();
            ok = pin[0] && pin[1];
        }
        int bsel = 0;
        for (long long c = r; c < nchunks; c += R, bsel ^= 1) {
            if (!ok) {
                failed(c, "ingest reader: stream / pinned buffer setup failed");
                continue;  // mark every chunk of this reader so the consumer never waits forever
            }
            if (sh.abort.load()) {
                failed(c, "");
                continue;
            }
            const long long pos = b0 + c * chunk;
            const size_t want = (size_t)std::min<long long>(chunk, b1 - pos);
            if (last[bsel] >= 0) cudaEventSynchronize(events[(size_t)last[bsel]]);  // the buffer's previous copy has left it
            if (src.host) {
                memcpy(pin[bsel], src.host + pos, want);
            } else {
                size_t got = 0;
                while (got < want) {
                    const ssize_t n = pread(src.fd, pin[bsel] + got, want - got, (off_t)(pos + (long long)got));
                    if (n <= 0) break;
                    got += (size_t)n;
                }
                if (got < want) {
                    failed(c, "short read at byte " + std::to_string(pos + (long long)got));
                    continue;
                }
            }
            cudaError_t e = cudaMemcpyAsync(d_dst + (pos - b0), pin[bsel], want, cudaMemcpyHostToDevice, rs);
            if (e == cudaSuccess) e = cudaEventRecord(events[(size_t)c], rs);
            if (e != cudaSuccess) {
                failed(c, std::string("H2D copy: ") + cudaGetErrorString(e));
                continue;
            }
            last[bsel] = c;
            std::lock_guard<std::mutex> lk(sh.mu);
            sh.state[(size_t)c] = 1;
            sh.cv.notify_all();
        }
        if (rs) {
            cudaStreamSynchronize(rs);  // the pinned buffers go back to the pool: nothing may still read them
            cudaStreamDestroy(rs);
        }
        pinned_release(pin[0]);
        pinned_release(pin[1]);
    };
    std::vector<std::thread> threads;
    for (int r = 0; r < R; ++r) threads.emplace_back(reader, r);
    int rc = APM_OK;
    for (long long c = 0; c < nchunks && !rc; ++c) {
        int stt;
        {
            std::unique_lock<std::mutex> lk(sh.mu);
            sh.cv.wait(lk, [&] { return sh.state[(size_t)c] != 0; });
            stt = sh.state[(size_t)c];
        }
        if (stt < 0) {
            std::lock_guard<std::mutex> lk(sh.mu);
            rc = fail(sh.err.rfind("short read", 0) == 0 ? APM_EIO : APM_ECUDA, "%s", sh.err.empty() ? "ingest aborted" : sh.err.c_str());
            break;
        }
        if (cudaStreamWaitEvent(st, events[(size_t)c], 0) != cudaSuccess) rc = fail(APM_ECUDA, "cudaStreamWaitEvent failed");
        if (!rc) rc = on_chunk(std::min(b1, b0 + (c + 1) * chunk));
    }
    if (rc) sh.abort = true;
    for (auto &t : threads) t.join();
    for (auto &ev : events) cudaEventDestroy(ev);
    return rc;
}

// optional match positions of the one-shot API
struct HitRequest {
    unsigned long long max_hits = 0;
    int *hit_pattern = nullptr;
    unsigned long long *hit_start = nullptr;
    unsigned long long n_hits = 0;  // out: all matching windows, may exceed max_hits
};

int count_impl(const TextSource &src, long long N, const char *const *patterns, const int *pattern_len,
               int nb_patterns, int approx_factor, long long *n_matches, HitRequest *hits = nullptr) {
    int rc = check_patterns(patterns, pattern_len, nb_patterns, approx_factor);
    if (rc) return rc;
    if (nb_patterns > 0 && !n_matches) return fail(APM_EINVAL, "n_matches is NULL");
    int ndev = 0;
    if ((rc = device_ready(&ndev))) return rc;
    for (int i = 0; i < nb_patterns; ++i) n_matches[i] = 0;
    const long long W = N - approx_factor;  // window starts (sequential.c:121)
    if (nb_patterns == 0 || W <= 0) return APM_OK;

    const Options opt = options_snapshot();
    int G = opt.gpus == 0 ? ndev : std::min(opt.gpus, ndev);
    int mmax = 0;
    for (int i = 0; i < nb_patterns; ++i) mmax = std::max(mmax, pattern_len[i]);
    int shard = opt.shard;
    if (shard == SHARD_AUTO)  // text shards when every GPU still gets a few tiles per SM (replaces main.c:88-123)
        shard = (W / G >= (long long)148 * 4 * 1024 || nb_patterns < G) ? SHARD_DB : SHARD_PATTERNS;
    if (shard == SHARD_PATTERNS) G = std::min(G, nb_patterns);
    if (shard == SHARD_DB) G = (int)std::min<long long>(G, W);

    // APM_TRACE=1: wall-clock of every phase of the one-shot call on stderr (synchronising after each phase)
    static const bool trace = getenv("APM_TRACE") && atoi(getenv("APM_TRACE")) > 0;
    auto t_last = std::chrono::steady_clock::now();
    auto mark = [&](const char *what, cudaStream_t st) {
        if (!trace) return;
        if (st) cudaStreamSynchronize(st);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[apm trace] %-22s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };

    int restore = 0;
    cudaGetDevice(&restore);
    std::vector<DevJob> jobs(G);
    auto bail = [&](int code) {
        std::string keep = tl_err;
        release_jobs(jobs, restore);
        tl_err = keep;
        return code;
    };
    const int readers = ingest_reader_count(opt, G);
    // One host thread per GPU: plan construction, text ingest and kernel launches of the GPUs proceed side by side
    // (with one thread the 8-GPU file run was one core's pread speed).  Errors are collected per job.
    std::vector<int> job_rc(G, APM_OK);
    std::vector<std::string> job_err(G);
    auto run_job = [&](int g) -> int {
        DevJob &j = jobs[g];
        int rc = APM_OK;
        j.dev = (restore + g) % ndev;  // the current device first
        if (cudaSetDevice(j.dev) != cudaSuccess) return fail(APM_ECUDA, "cudaSetDevice(%d) failed", j.dev);
        if (cudaStreamCreateWithFlags(&j.st, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&j.copy_st, cudaStreamNonBlocking) != cudaSuccess)
            return fail(APM_ECUDA, "cudaStreamCreate failed on device %d", g);
        if (g == 0) mark("stream create", nullptr);
        if ((rc = apm_plan_create(patterns, pattern_len, nb_patterns, approx_factor, &j.plan))) return rc;
        if (g == 0) mark("plan create", nullptr);
        if (hits && hits->max_hits > 0) {  // every GPU may find up to max_hits positions
            if (dev_alloc((void **)&j.d_hits, (hits->max_hits + 1) * sizeof(unsigned long long)) != cudaSuccess)
                return fail(APM_ENOMEM, "cudaMalloc of the hit buffer (%llu entries) failed", hits->max_hits);
            if (cudaMemsetAsync(j.d_hits, 0, sizeof(unsigned long long), j.st) != cudaSuccess)
                return fail(APM_ECUDA, "cudaMemsetAsync failed");
            if ((rc = apm_plan_set_hit_buffer(j.plan, j.d_hits + 1, hits->max_hits, j.d_hits))) return rc;
        }
        if (shard == SHARD_PATTERNS) {
            if (G > 1 && (rc = apm_plan_set_pattern_shard(j.plan, g, G))) return rc;
            j.j0 = 0;
            j.j1 = W;
        } else {  // database shard g owns window STARTS [j0, j1); bytes up to j1 + mmax - 1 (halo)
            j.j0 = (W * g / G) & ~15ll;
            j.j1 = g == G - 1 ? W : ((W * (g + 1) / G) & ~15ll);
        }
        j.b0 = j.j0;
        j.b1 = std::min(N, j.j1 + mmax - 1);
        // A shard of more than text_chunk_mb window starts is streamed through two device buffers, segment by
        // segment (each with its own m_max - 1 halo): the copy of segment s+1 overlaps the counting of segment s and
        // the device memory needed is bounded whatever the size of the text.
        const long long seg_w = std::max<long long>(1, opt.text_chunk_mb) << 20;
        const long long nseg = std::max<long long>(1, (j.j1 - j.j0 + seg_w - 1) / seg_w);
        const long long seg_bytes = nseg == 1 ? j.b1 - j.b0 : seg_w + mmax - 1;
        if (dev_alloc((void **)&j.d_text, (size_t)std::max<long long>(16, seg_bytes)) != cudaSuccess ||
            (nseg > 1 && dev_alloc((void **)&j.d_text2, (size_t)std::max<long long>(16, seg_bytes)) != cudaSuccess))
            return fail(APM_ENOMEM, "cudaMalloc of %lld text bytes failed on device %d", seg_bytes, g);
        if (g == 0) mark("text malloc", nullptr);
        // counting is launched for batches of complete windows while later chunks are still being read
        const long long batch = opt.mode == MODE_DIRECT ? (long long)(8 << 20) : (long long)(64 << 20);
        for (long long sgi = 0; sgi < nseg; ++sgi) {
            const long long w0 = j.j0 + sgi * seg_w, w1 = std::min(j.j1, w0 + seg_w);
            const long long sb0 = w0, sb1 = std::min(N, w1 + mmax - 1);
            const int bi = (int)(sgi & 1);
            uint8_t *d_seg = bi ? j.d_text2 : j.d_text;
            cudaEvent_t wait_before = nullptr;
            if (nseg > 1) {
                if (!j.seg_done[bi]) {
                    if (cudaEventCreateWithFlags(&j.seg_done[bi], cudaEventDisableTiming) != cudaSuccess)
                        return fail(APM_ECUDA, "cudaEventCreate failed");
                } else {
                    wait_before = j.seg_done[bi];  // the buffer is still being counted
                }
            }
            long long counted_to = w0;
            auto on_chunk = [&](long long bytes_end) -> int {
                const long long w_end = bytes_end >= sb1 ? w1 : std::min(w1, bytes_end - (mmax - 1));
                if (w_end - counted_to < (bytes_end >= sb1 ? 1 : batch)) return APM_OK;  // batch small steps
                const int r = apm_plan_count_device(j.plan, d_seg, (unsigned long long)sb0, (unsigned long long)(sb1 - sb0),
                                                    (unsigned long long)N, (unsigned long long)counted_to,
                                                    (unsigned long long)w_end, j.st);
                counted_to = w_end;
                return r;
            };
            if ((rc = copy_range_to_device(src, sb0, sb1, d_seg, j.st, j.copy_st, wait_before, readers, on_chunk))) return rc;
            if (sgi == 0 && g == 0) mark("text H2D", src.host ? j.st : nullptr);
            if (counted_to < w1 &&
                (rc = apm_plan_count_device(j.plan, d_seg, (unsigned long long)sb0, (unsigned long long)(sb1 - sb0),
                                            (unsigned long long)N, (unsigned long long)counted_to, (unsigned long long)w1, j.st)))
                return rc;
            if (nseg > 1 && cudaEventRecord(j.seg_done[bi], j.st) != cudaSuccess) return fail(APM_ECUDA, "cudaEventRecord failed");
        }
        if (g == 0) mark("count kernels", j.st);
        return APM_OK;
    };
    if (G == 1) {
        if ((rc = run_job(0))) return bail(rc);
    } else {
        std::vector<std::thread> workers;
        for (int g = 0; g < G; ++g)
            workers.emplace_back([&, g] {
                job_rc[g] = run_job(g);
                if (job_rc[g]) job_err[g] = tl_err;  // tl_err is per thread
            });
        for (auto &t : workers) t.join();
        for (int g = 0; g < G; ++g)
            if (job_rc[g]) {
                tl_err = job_err[g];
                return bail(job_rc[g]);
            }
        cudaSetDevice(restore);
    }
    // ---- combine the per-GPU count vectors
    bool reduced_on_device = false;
    if (G > 1 && (opt.reduce == 0 || opt.reduce == 3)) {
        // our own kernel over NVLink peer memory: GPU 0 pulls every other GPU's vector and adds it to its own
        bool peers = true;
        for (int g = 1; g < G && peers; ++g) {
            int can = 0;
            peers = cudaDeviceCanAccessPeer(&can, jobs[0].dev, jobs[g].dev) == cudaSuccess && can;
        }
        if (peers) {
            for (int g = 1; g < G; ++g) {
                cudaEvent_t ev = nullptr;
                cudaSetDevice(jobs[g].dev);
                cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
                if (e == cudaSuccess) e = cudaEventRecord(ev, jobs[g].st);  // behind GPU g's count kernels
                cudaSetDevice(jobs[0].dev);
                if (e == cudaSuccess) {
                    e = cudaDeviceEnablePeerAccess(jobs[g].dev, 0);
                    if (e == cudaErrorPeerAccessAlreadyEnabled) {
                        cudaGetLastError();
                        e = cudaSuccess;
                    }
                }
                if (e == cudaSuccess) e = cudaStreamWaitEvent(jobs[0].st, ev, 0);
                if (e == cudaSuccess) {
                    peer_gather_add_kernel<<<(nb_patterns + 255) / 256, 256, 0, jobs[0].st>>>(jobs[0].plan->d_counts,
                                                                                            jobs[g].plan->d_counts, nb_patterns);
                    e = cudaGetLastError();
                    g_launches++;
                }
                if (ev) cudaEventDestroy(ev);
                if (e != cudaSuccess) return bail(fail(APM_ECUDA, "peer reduction (%d <- %d): %s", jobs[0].dev, jobs[g].dev, cudaGetErrorString(e)));
            }
            reduced_on_device = true;
        } else if (opt.reduce == 3) {
            return bail(fail(APM_ECUDA, "reduce=p2p requested but the GPUs cannot access each other's memory"));
        }
    }
    if (G > 1 && !reduced_on_device && opt.reduce != 2 && opt.reduce != 3) {
        std::lock_guard<std::mutex> lk(g_nccl_mu);
        if (nccl_load()) {
            std::vector<int> devs;
            for (auto &j : jobs) devs.push_back(j.dev);
            if ((rc = nccl_comms_for(devs))) return bail(rc);
            int nrc = g_nccl.GroupStart();
            for (int g = 0; g < G && nrc == 0; ++g) {
                cudaSetDevice(jobs[g].dev);
                nrc = g_nccl.AllReduce(jobs[g].plan->d_counts, jobs[g].plan->d_counts, (size_t)nb_patterns, /*ncclUint64*/ 5,
                                       /*ncclSum*/ 0, g_nccl.comms[g], jobs[g].st);
            }
            const int erc = g_nccl.GroupEnd();
            if (nrc == 0) nrc = erc;
            if (nrc != 0)
                return bail(fail(APM_ECUDA, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "error"));
            reduced_on_device = true;
        } else if (opt.reduce == 1) {
            return bail(fail(APM_ECUDA, "reduce=nccl requested but libnccl.so.2 could not be loaded"));
        }
    }
    std::vector<long long> part(nb_patterns);
    for (int g = 0; g < G; ++g) {
        cudaSetDevice(jobs[g].dev);
        if (reduced_on_device) {  // every GPU holds the total; read it once, just drain the others
            if (g == 0) {
                if ((rc = apm_plan_read_counts(jobs[0].plan, n_matches, jobs[0].st))) return bail(rc);
            } else if (cudaStreamSynchronize(jobs[g].st) != cudaSuccess) {
                return bail(fail(APM_ECUDA, "stream synchronize failed on device %d", jobs[g].dev));
            }
            continue;
        }
        if ((rc = apm_plan_read_counts(jobs[g].plan, part.data(), jobs[g].st))) return bail(rc);
        for (int i = 0; i < nb_patterns; ++i) n_matches[i] += part[i];
    }
    if (hits && hits->max_hits > 0) {  // gather, order by (pattern, start), keep the first max_hits
        std::vector<unsigned long long> all;
        hits->n_hits = 0;
        for (int g = 0; g < G; ++g) {
            cudaSetDevice(jobs[g].dev);
            unsigned long long n = 0;
            if (cudaMemcpyAsync(&n, jobs[g].d_hits, sizeof n, cudaMemcpyDeviceToHost, jobs[g].st) != cudaSuccess ||
                cudaStreamSynchronize(jobs[g].st) != cudaSuccess)
                return bail(fail(APM_ECUDA, "reading the hit counter failed on device %d", jobs[g].dev));
            hits->n_hits += n;
            const unsigned long long take = std::min(n, hits->max_hits);
            const size_t old = all.size();
            all.resize(old + take);
            if (take && (cudaMemcpyAsync(all.data() + old, jobs[g].d_hits + 1, take * sizeof(unsigned long long),
                                         cudaMemcpyDeviceToHost, jobs[g].st) != cudaSuccess ||
                         cudaStreamSynchronize(jobs[g].st) != cudaSuccess))
                return bail(fail(APM_ECUDA, "reading the hits failed on device %d", jobs[g].dev));
        }
        std::sort(all.begin(), all.end());  // packed as pattern << 40 | start
        const size_t keep = (size_t)std::min<unsigned long long>(all.size(), hits->max_hits);
        for (size_t i = 0; i < keep; ++i) {
            hits->hit_pattern[i] = (int)(all[i] >> kHitPosBits);
            hits->hit_start[i] = all[i] & ((1ull << kHitPosBits) - 1);
        }
    }
    mark("reduce + D2H", nullptr);
    release_jobs(jobs, restore);
    mark("release", nullptr);
    return APM_OK;
}

}  // namespace

int apm_count_matches(const unsigned char *text, size_t n_bytes, const char *const *patterns,
                      const int *pattern_len, int nb_patterns, int approx_factor, long long *n_matches) {
    if (!text && n_bytes > 0) return fail(APM_EINVAL, "text is NULL");
    TextSource src;
    static const unsigned char empty = 0;
    src.host = text ? text : &empty;
    return count_impl(src, (long long)n_bytes, patterns, pattern_len, nb_patterns, approx_factor, n_matches);
}

int apm_find_matches(const unsigned char *text, size_t n_bytes, const char *const *patterns, const int *pattern_len,
                     int nb_patterns, int approx_factor, long long *n_matches, unsigned long long max_hits,
                     int *hit_pattern, unsigned long long *hit_start, unsigned long long *n_hits) {
    if (!text && n_bytes > 0) return fail(APM_EINVAL, "text is NULL");
    if (max_hits > 0 && (!hit_pattern || !hit_start)) return fail(APM_EINVAL, "hit arrays are NULL");
    if (n_bytes >= (1ull << kHitPosBits)) return fail(APM_EINVAL, "text too large for packed hit entries");
    TextSource src;
    static const unsigned char empty = 0;
    src.host = text ? text : &empty;
    HitRequest hr;
    hr.max_hits = max_hits;
    hr.hit_pattern = hit_pattern;
    hr.hit_start = hit_start;
    const int rc = count_impl(src, (long long)n_bytes, patterns, pattern_len, nb_patterns, approx_factor, n_matches, &hr);
    if (n_hits) *n_hits = hr.n_hits;
    return rc;
}

int apm_count_matches_file(const char *path, const char *const *patterns, const int *pattern_len,
                           int nb_patterns, int approx_factor, long long *n_matches,
                           unsigned long long *n_bytes_out) {
    if (!path) return fail(APM_EINVAL, "path is NULL");
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(APM_EIO, "Unable to open the text file <%s>", path);
    struct stat sb;
    if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) {
        close(fd);
        return fail(APM_EIO, "Unable to stat the text file <%s>", path);
    }
    if (n_bytes_out) *n_bytes_out = (unsigned long long)sb.st_size;
    TextSource src;
    src.fd = fd;
    const int rc = count_impl(src, (long long)sb.st_size, patterns, pattern_len, nb_patterns, approx_factor, n_matches);
    close(fd);
    return rc;
}

}  // extern "C"
