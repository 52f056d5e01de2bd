// apm_dna.cu -- host side of the DNA seed filter: table construction and kernel launches (kernels: apm_dna.cuh).
#include "apm_dna.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>

#include "apm_dna.cuh"

namespace apm {

namespace {

template <typename T>
cudaError_t upload_vec(T **dptr, const std::vector<T> &h) {
    *dptr = nullptr;
    if (h.empty()) return cudaSuccess;
    cudaError_t e = pool_alloc((void **)dptr, h.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

std::mutex g_attr_mu;
std::map<std::pair<int, const void *>, size_t> g_attr_set;
template <typename Fn>
cudaError_t ensure_smem(Fn fn, size_t bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(g_attr_mu);
    size_t &cur = g_attr_set[{dev, (const void *)fn}];
    if (bytes <= cur) return cudaSuccess;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
}

template <int H, bool PACKED>
cudaError_t launch_scan_p(const DnaArgs &a, unsigned blocks, cudaStream_t st) {
    const size_t smem = dna_smem_bytes(a.table_words, H);
    const cudaError_t e = ensure_smem(dna_scan_kernel<H, PACKED>, smem);
    if (e != cudaSuccess) return e;
    dna_scan_kernel<H, PACKED><<<blocks, kDnaThreads, smem, st>>>(a);
    return cudaGetLastError();
}
template <int H>
cudaError_t launch_scan(const DnaArgs &a, unsigned blocks, cudaStream_t st) {
    return a.packed ? launch_scan_p<H, true>(a, blocks, st) : launch_scan_p<H, false>(a, blocks, st);
}

}  // namespace

bool dna_choose(const std::vector<std::string> &pats, const std::vector<int> &ids, int k, int *q_out, int *h_out) {
    if (ids.empty() || k > kDnaMaxK) return false;
    int lmin = 1 << 30;
    for (int p : ids) {
        const std::string &s = pats[p];
        for (unsigned char c : s)
            if (c != 'A' && c != 'C' && c != 'G' && c != 'T') return false;
        lmin = std::min(lmin, (int)s.size() / (k + 1));
        if (s.size() >= 65536) return false;
    }
    if (lmin < kDnaQMin || ids.size() >= ((size_t)1 << 24)) return false;
    const int q = std::min(lmin, kDnaQMax);
    const double nent = (double)ids.size() * (k + 1), space = (double)(1ull << (2 * q));
    int h = std::min(kDnaHMax, lmin - q + 1);
    while (h > 1 && nent * h > space / 8) --h;  // keep the bitmap sparse: every set bit costs a second-level probe
    if (nent * h > space / 2) return false;     // too many patterns for a direct q-gram bitmap: hashed scan instead
    *q_out = q;
    *h_out = h;
    return true;
}

int dna_build(const std::vector<std::string> &pats, const std::vector<int> &ids, int k, int q, int h, int nseg,
              size_t cand_bytes, DnaSet *out) {
    DnaSet s;
    s.q = q;
    s.h = h;
    s.mmax = 0;
    s.mmin = 1 << 30;
    const size_t space = (size_t)1 << (2 * q);
    s.table_words = (int)(space / 32);
    struct Ent { uint32_t key; uint4 v; };
    std::vector<Ent> ents;
    ents.reserve(ids.size() * (size_t)(k + 1) * h);
    std::vector<uint32_t> table(s.table_words, 0u);
    std::vector<uint32_t> bloom((size_t)1 << (kDnaBloomLog - 5), 0u);
    const int glen = q + h - 1;  // symbols known at a probe: the piece's first q + H - 1
    std::vector<uint4> peq32(ids.size() * (size_t)(k + 1) * 2, make_uint4(0u, 0u, 0u, 0u));
    for (size_t slot = 0; slot < ids.size(); ++slot) {
        const std::string &pat = pats[ids[slot]];
        const int m = (int)pat.size();
        s.mmax = std::max(s.mmax, m);
        s.mmin = std::min(s.mmin, m);
        for (int i = 0; i <= k; ++i) {
            const int o = dna_piece_offset(i, m, k), len = dna_piece_offset(i + 1, m, k) - o;
            uint32_t pack = 0;
            for (int x = 0; x < std::min(len, 16); ++x) pack |= dna_code((uint8_t)pat[o + x]) << (2 * x);
            uint32_t gram = 0;
            for (int x = 0; x < glen; ++x) gram |= dna_code((uint8_t)pat[o + x]) << (2 * x);
            const uint32_t bi = dna_bloom_index(gram);
            bloom[bi >> 5] |= 1u << (bi & 31);
            // stage-1 masks: row r of the left side is pattern symbol o - 1 - r, of the right side o + len + r
            uint32_t lm[4] = {0u, 0u, 0u, 0u}, rm[4] = {0u, 0u, 0u, 0u};
            for (int r = 0; r < 32; ++r) {
                if (o - 1 - r >= 0) lm[dna_code((uint8_t)pat[o - 1 - r])] |= 1u << r;
                if (o + len + r < m) rm[dna_code((uint8_t)pat[o + len + r])] |= 1u << r;
            }
            peq32[(slot * (size_t)(k + 1) + i) * 2 + 0] = make_uint4(lm[0], lm[1], lm[2], lm[3]);
            peq32[(slot * (size_t)(k + 1) + i) * 2 + 1] = make_uint4(rm[0], rm[1], rm[2], rm[3]);
            for (int off = 0; off < h; ++off) {
                uint32_t key = 0;
                for (int x = 0; x < q; ++x) key |= dna_code((uint8_t)pat[o + off + x]) << (2 * x);
                table[key >> 5] |= 1u << (key & 31);
                ents.push_back({key, make_uint4(((uint32_t)slot << 8) | ((uint32_t)i << 3) | (uint32_t)off, pack,
                                                (uint32_t)o | ((uint32_t)len << 16), (uint32_t)m)});
            }
        }
    }
    std::stable_sort(ents.begin(), ents.end(), [](const Ent &a, const Ent &b) { return a.key < b.key; });
    std::vector<uint32_t> first(space + 1, 0u);
    std::vector<uint4> ev;
    ev.reserve(ents.size());
    for (auto &e : ents) {
        first[e.key + 1]++;
        ev.push_back(e.v);
    }
    for (size_t c = 1; c < first.size(); ++c) first[c] += first[c - 1];
    s.nent = (int)ents.size();
    s.nseg = nseg;
    s.nfp = (int)ids.size();
    s.seg_cap = (unsigned int)std::min<size_t>(0x7FFFFFFFu, std::max<size_t>(1024, cand_bytes / sizeof(uint64_t) / (size_t)nseg));
    s.long_cap = std::max<size_t>(4096, cand_bytes / sizeof(uint64_t) / 8);
    s.set_log_max = 12;
    while (s.set_log_max < 26 && ((size_t)16 << s.set_log_max) <= cand_bytes / 8) ++s.set_log_max;  // cand_bytes / 8 of set
    cudaError_t e = upload_vec(&s.d_table, table);
    if (e == cudaSuccess) e = upload_vec(&s.d_peq32, peq32);
    if (e == cudaSuccess) e = upload_vec(&s.d_bloom, bloom);
    if (e == cudaSuccess) e = pool_alloc((void **)&s.d_long, s.long_cap * sizeof(uint64_t));
    if (e == cudaSuccess) e = pool_alloc((void **)&s.d_set, sizeof(uint64_t) << s.set_log_max);
    if (e == cudaSuccess) e = upload_vec(&s.d_first, first);
    if (e == cudaSuccess) e = upload_vec(&s.d_ent, ev);
    if (e == cudaSuccess) e = pool_alloc((void **)&s.d_cand, (size_t)s.seg_cap * nseg * sizeof(uint64_t));
    if (e == cudaSuccess) e = pool_alloc((void **)&s.d_seg_count, (size_t)nseg * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(s.d_seg_count, 0, (size_t)nseg * sizeof(unsigned int));
    if (e != cudaSuccess) {
        dna_free(&s);
        return (int)e;
    }
    *out = s;
    return 0;
}

void dna_free(DnaSet *s) {
    pool_free(s->d_table);
    pool_free(s->d_first);
    pool_free(s->d_ent);
    pool_free(s->d_cand);
    pool_free(s->d_seg_count);
    pool_free(s->d_peq32);
    pool_free(s->d_bloom);
    pool_free(s->d_long);
    pool_free(s->d_set);
    *s = DnaSet();
}

unsigned long long dna_pack_words(unsigned long long buf_len) { return (buf_len + 15) / 16 + 1; }

cudaError_t dna_pack_text(const uint8_t *d_buf, unsigned long long buf_len, uint32_t *d_packed, cudaStream_t st) {
    const long long nwords = (long long)dna_pack_words(buf_len);
    const unsigned blocks = (unsigned)std::min<long long>((nwords + 255) / 256, 148ll * 32);
    dna_pack_kernel<<<blocks, 256, 0, st>>>(d_buf, (long long)buf_len, d_packed, nwords);
    return cudaGetLastError();
}

cudaError_t dna_launch(const DnaSet &s, const DnaRun &r, cudaStream_t st, int *launches) {
    DnaArgs a;
    a.buf = r.buf;
    a.packed = r.packed;
    a.buf_len = r.buf_len;
    a.n_end = r.n_end;
    a.w0 = r.w0;
    a.w1 = r.w1;
    a.q = s.q;
    a.k = r.k;
    a.mmax = s.mmax;
    a.table_words = s.table_words;
    a.keymask = (uint32_t)(((uint64_t)1 << (2 * s.q)) - 1);
    a.table = s.d_table;
    a.first = s.d_first;
    a.ent = s.d_ent;
    a.fp_id = r.fp_id;
    a.fp_m = r.fp_m;
    a.fp_off = r.fp_off;
    a.pat_bytes = r.pat_bytes;
    a.cand = s.d_cand;
    a.seg_cap = s.seg_cap;
    a.seg_count = s.d_seg_count;
    a.nseg = s.nseg;
    a.amask = (a.keymask >> 3) & ~3u;
    a.peq32 = s.d_peq32;
    a.h = s.h;
    a.bloom = s.d_bloom;
    a.gmask = (s.q + s.h - 1) >= 16 ? 0xFFFFFFFFu : (uint32_t)(((uint64_t)1 << (2 * (s.q + s.h - 1))) - 1);
    a.overflow = reinterpret_cast<unsigned int *>(r.ctr + 1);
    a.nset = r.ctr;
    a.nlong = r.ctr + 2;
    a.longq = s.d_long;
    a.longcap = s.long_cap;
    a.set = s.d_set;
    // the set never needs more than 2 slots per possible key of the round
    a.set_log = 12;
    const double keys = (double)(r.w1 - r.w0) * (double)s.nfp;
    while (a.set_log < s.set_log_max && (double)(1ull << a.set_log) < 2.0 * keys) ++a.set_log;
    a.counts = r.counts;
    a.sink = r.sink;
    cudaError_t e = cudaMemsetAsync(r.ctr, 0, 3 * sizeof(unsigned long long), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(s.d_set, 0, sizeof(uint64_t) << a.set_log, st);
    if (e != cudaSuccess) return e;
    switch (s.h) {
        case 1: e = launch_scan<1>(a, (unsigned)s.nseg, st); break;
        case 2: e = launch_scan<2>(a, (unsigned)s.nseg, st); break;
        case 3: e = launch_scan<3>(a, (unsigned)s.nseg, st); break;
        default: e = launch_scan<4>(a, (unsigned)s.nseg, st); break;
    }
    if (e != cudaSuccess) return e;
    const int parts = 8;
    dna_verify_kernel<<<(unsigned)(s.nseg * parts), kDnaVerifyThreads, 0, st>>>(a, parts);
    dna_verify_long_kernel<<<(unsigned)(s.nseg * 8), kDnaLongThreads, 0, st>>>(a);
    dna_collect_kernel<<<(unsigned)std::min<unsigned long long>((1ull << a.set_log) / 256, (unsigned long long)s.nseg * 8), 256, 0, st>>>(a);
    *launches += 4;
    return cudaGetLastError();
}

}  // namespace apm
