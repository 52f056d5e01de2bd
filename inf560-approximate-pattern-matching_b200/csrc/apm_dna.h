// apm_dna.h -- host interface of the DNA seed filter (apm_dna.cuh, compiled in its own translation unit apm_dna.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "apm_common.cuh"

namespace apm {

// device-memory cache of the library (apm_api.cu)
__attribute__((visibility("hidden"))) cudaError_t pool_alloc(void **p, size_t bytes);
__attribute__((visibility("hidden"))) void pool_free(void *p);

struct DnaSet {
    int q = 0, h = 0, nent = 0, mmax = 0, mmin = 0, table_words = 0, nseg = 0;
    uint32_t *d_table = nullptr, *d_first = nullptr, *d_bloom = nullptr;
    uint4 *d_ent = nullptr;
    uint4 *d_peq32 = nullptr;             // per (slot, piece): 32-row match masks left / right of the piece
    uint64_t *d_cand = nullptr;           // seed hits, one segment per scan CTA
    unsigned int seg_cap = 0;
    unsigned int *d_seg_count = nullptr;
    uint64_t *d_long = nullptr;           // survivors of the lower-bound test
    unsigned long long long_cap = 0;
    uint64_t *d_set = nullptr;            // hash set of matching (slot, window) keys
    int set_log_max = 0;
    int nfp = 0;
};

// per-launch arguments that do not live in the DnaSet
struct DnaRun {
    const uint8_t *buf;
    const uint32_t *packed = nullptr;  // optional resident 2-bit copy of buf (dna_pack_text); buf must be 16-byte aligned
    long long buf_len, n_end, w0, w1;
    int k;
    const int *fp_id, *fp_m;
    const long long *fp_off;
    const uint8_t *pat_bytes;
    unsigned long long *ctr;  // [0] keys in the set, [1] (low word) overflow flag, [2] survivors; zeroed by dna_launch
    unsigned long long *counts;
    HitSink sink;
};

// Can the 2-bit scan serve the patterns `ids` (all symbols in ACGT, the q-gram bitmap stays sparse)?  Chooses the
// q-gram length and the probe stride.
__attribute__((visibility("hidden"))) bool dna_choose(const std::vector<std::string> &pats, const std::vector<int> &ids, int k,
                                                    int *q, int *h);
// Builds and uploads bitmap / CSR / entries and allocates the candidate segments (cand_bytes in total, nseg CTAs).
// Returns 0 or a cudaError_t.
__attribute__((visibility("hidden"))) int dna_build(const std::vector<std::string> &pats, const std::vector<int> &ids, int k, int q,
                                                  int h, int nseg, size_t cand_bytes, DnaSet *out);
__attribute__((visibility("hidden"))) void dna_free(DnaSet *s);
// 2-bit copy of a device text buffer (16-byte aligned) for repeated searches: words needed, and the packing itself
__attribute__((visibility("hidden"))) unsigned long long dna_pack_words(unsigned long long buf_len);
__attribute__((visibility("hidden"))) cudaError_t dna_pack_text(const uint8_t *d_buf, unsigned long long buf_len, uint32_t *d_packed,
                                                                cudaStream_t st);
// scan + verification of one round on `st` (stream ordered).  *launches += kernels launched.
__attribute__((visibility("hidden"))) cudaError_t dna_launch(const DnaSet &s, const DnaRun &r, cudaStream_t st, int *launches);

}  // namespace apm
