/*
 * apm_b200.h -- C-ABI of the B200-native approximate-pattern-matching hot path.
 *
 * libapm_b200.so replaces, for the reference linomp/INF560-approximate-pattern-matching, the
 * evaluation of levenshtein() over every (pattern, text window) pair and the per-pattern count:
 *
 *   reference interface                                   replaced by
 *   ---------------------------------------------------   -------------------------------------------
 *   src/sequential.c:105-144 (search loop in main)        apm_count_matches / apm_count_matches_file
 *   include/utils.h:10  levenshtein(s1,s2,len,column)     (evaluated inside the CUDA kernels)
 *   include/approaches.h:4-7  patterns_over_ranks_hybrid  apm_count_matches + option shard=patterns
 *                             database_over_ranks         apm_count_matches + option shard=db
 *   src/patterns_over_ranks.c:33-36 invoke_kernel /       apm_plan_create + apm_plan_count_device
 *       write_kernel_result (async GPU launch API)        (stream-ordered; counts stay on the device)
 *   src/database_over_ranks.c:18-22 initializeGPU /       apm_plan_create + apm_plan_count_device
 *       getGPUResult                                      + apm_plan_read_counts
 *   src/utils.c:12-68  read_input_file                    apm_count_matches_file (64-bit sizes)
 *   src/main.c:18-19   getDeviceCount / setDevice         apm_device_count / option "gpus"
 *
 * Semantics (bit-exact with src/sequential.c):
 *   - text is the raw byte buffer (no FASTA parsing; '\n' is a symbol); comparison is byte equality;
 *   - for pattern i of length m and every window start j in [0, n_bytes - approx_factor):
 *       size = min(m, n_bytes - j); d = levenshtein(pattern_i[0..size), text[j..j+size));
 *       n_matches[i] += (d <= approx_factor);
 *   - approx_factor < 0, nb_patterns < 0, a NULL/empty pattern -> error (the reference is UB / exits).
 *
 * Conventions: plain C, all functions return 0 on success and a non-zero APM_E* code on failure and
 * never call exit(); apm_last_error() gives the message of the calling thread's last failure.
 * Pointers are borrowed for the duration of the call.  There is NO CPU fallback: without a CUDA
 * device every compute entry point fails with APM_ENODEVICE.
 */
#ifndef APM_B200_H
#define APM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APM_OK 0
#define APM_EINVAL 1    /* bad argument */
#define APM_ENODEVICE 2 /* no CUDA device / CUDA runtime unusable */
#define APM_ECUDA 3     /* a CUDA call failed (see apm_last_error) */
#define APM_EIO 4       /* file could not be read */
#define APM_ENOMEM 5

/* ---- one-shot host API (drop-in for the search loop; HOST pointers) --------------------------- */

/* n_matches[i] = number of windows of `text` within distance approx_factor of patterns[i].
 * patterns[i] need not be NUL-terminated; pattern_len[i] > 0 bytes are used.                      */
int apm_count_matches(const unsigned char *text, size_t n_bytes, const char *const *patterns,
                      const int *pattern_len, int nb_patterns, int approx_factor,
                      long long *n_matches);

/* Same, reading the text from `path` exactly as read_input_file() does (raw bytes), 64-bit sizes,
 * chunked through pinned staging buffers so the H2D copy overlaps the file read.                  */
int apm_count_matches_file(const char *path, const char *const *patterns, const int *pattern_len,
                           int nb_patterns, int approx_factor, long long *n_matches,
                           unsigned long long *n_bytes_out);

/* apm_count_matches that also reports WHERE the matches are (the reference only counts): the
 * first max_hits matching windows in (pattern index, window start) order go to hit_pattern[] / hit_start[];
 * *n_hits = number of matching windows found (= sum of n_matches; 0 when max_hits is 0; larger than max_hits when
 * the arrays were too small -- with several GPUs each device keeps at most max_hits, so the listed hits are then a
 * subset).
 * Truncated tail windows are reported with their start like any other window.  n_bytes < 2^40.           */
int apm_find_matches(const unsigned char *text, size_t n_bytes, const char *const *patterns,
                     const int *pattern_len, int nb_patterns, int approx_factor, long long *n_matches,
                     unsigned long long max_hits, int *hit_pattern, unsigned long long *hit_start,
                     unsigned long long *n_hits);

/* Options (process-wide; a plan and a one-shot call take a snapshot when they start):
 *   "gpus"    = "1".."8" | "all"      devices used by the one-shot API (default 1)
 *   "shard"   = "db" | "patterns" | "auto"   how work is split over several GPUs (default auto)
 *   "kernel"  = "auto" | "sliced" | "myers" | "dp"
 *                 auto (default): window-sliced bit-parallel kernel (m <= 1024, <= 8 pattern symbols),
 *                 row-parallel Hyyro/Myers kernel otherwise (m <= 256), explicit DP for the rest;
 *                 dp: explicit-DP kernel for everything (in-GPU cross-check)
 *   "tail"    = "auto" | "bitpar" | "dp"   truncated tail windows (sequential.c:131-136) of the bit-parallel patterns:
 *                 bitpar (= auto): Hyyro/Myers bit-vector kernels, one thread (m <= 256) or one warp (m <= 1024) per
 *                 window; dp: the explicit-DP kernel (its banded variant in band / filter mode)
 *   "mode"    = "direct" | "band" | "filter"
 *                 direct (default): every DP cell of every window is evaluated;
 *                 band: exact Ukkonen band |i-j| <= k (same counts, ~(2k+1)/m of the work);
 *                 filter: exact pigeonhole filter -- ONE scan of the text against the seeds of all patterns (k+1
 *                 pieces each), verification of the seed hits only; patterns whose pieces are shorter than 8
 *                 symbols (or k > 16) and rounds whose candidates overflow the buffer go through the band kernel.
 *                 Same counts; work ~ text bytes instead of text bytes x patterns.
 *   "filter_scan" = "auto" | "dna" | "hash"   scan of filter mode: dna = 2-bit q-gram scan on a strided probe grid
 *                 + seed-hit verification (pattern symbols all in ACGT, k <= 15; any text bytes); hash = rolling-hash
 *                 scan + per-shift banded verification (any alphabet); auto (default) = dna when the patterns allow
 *   "filter_cand_mb" = candidate buffers of filter mode in MiB (default 128)
 *   "ingest_threads" = "auto" | 1..64   reader threads PER GPU of the chunked ingest (files and large pageable host
 *                 buffers: pread / memcpy into pinned staging buffers, async H2D, counting overlapped); auto =
 *                 host threads / GPUs, at most 8
 *   "cell"    = "auto" | "lop3" | "fma3" | "fma" | "fma3r"   code of one DP cell in the window-sliced / band kernels:
 *                 lop3: 5 LOP3 (ALU pipe only); fma3: 4 LOP3 + 3 IMAD; fma: 4 LOP3 + 2 IMAD (FMA pipe takes the
 *                 subtractions); fma3r: 4 LOP3 + 3 IMAD with the operands ordered for the register file's reuse cache;
 *                 auto (default) picks per pattern-length class (fma3r for m <= 224).  Same results, different speed.
 *   "reduce"  = "auto" | "p2p" | "nccl" | "host"   how the per-GPU count vectors of the one-shot API are combined
 *                 when "gpus" > 1: p2p = GPU 0 pulls every other GPU's vector through NVLink peer memory and adds it to
 *                 its own (our own kernel, stream ordered, no communicator, no cross-device atomics); nccl = one in-place ncclAllReduce per device (NCCL
 *                 loaded at run time); host = host-side sum; auto (default) = p2p, else nccl, else host
 *   "text_chunk_mb" = one-shot API: a GPU's shard with more than this many Mi window starts (default 32768) is
 *                 streamed through two device buffers, segment by segment with its own halo, so the device memory
 *                 needed is bounded for any text size; the copy of the next segment overlaps the counting
 *   "cache_mb" = device memory (MiB) kept cached between calls instead of cudaFree'd (default 4096)
 *   "rblock"  = "1" | "2" | "4" | "auto"     patterns register-blocked per thread (row-parallel kernel)
 *   "tile"    = window starts per CTA tile (multiple of 256) or "auto"   (row-parallel kernel)
 *   "variant" = "0" | "1" | "2"       column-step code variant of the row-parallel kernel
 * Unknown key / value -> APM_EINVAL.                                                              */
int apm_set_option(const char *key, const char *value);
const char *apm_get_option(const char *key);

const char *apm_last_error(void);
int apm_device_count(int *count);
/* Makes `device` current for the calling thread inside the library's CUDA runtime (replaces setDevice,
 * src/cuda_utils.cu:22-35).  The one-shot API uses the current device, plus the following ones when
 * "gpus" > 1; apm_plan_create binds the plan to the current device.                                */
int apm_set_device(int device);

/* ---- planned / device-resident API (stream-ordered, no host synchronisation) ------------------- */

typedef struct apm_plan apm_plan;

/* Builds the per-call alphabet, the bit-parallel Peq tables and the pattern groups, and uploads
 * them to the CURRENT CUDA device.  The plan owns a device vector of nb_patterns 64-bit counters. */
int apm_plan_create(const char *const *patterns, const int *pattern_len, int nb_patterns,
                    int approx_factor, apm_plan **plan_out);
int apm_plan_destroy(apm_plan *plan);

/* Evaluate window starts [j_begin, j_end) (global coordinates, clamped to n_total - approx_factor)
 * of a global text of n_total bytes.  d_buf is a DEVICE pointer holding the global bytes
 * [buf_offset, buf_offset + buf_len); it must cover [j_begin, min(n_total, j_end + m_max - 1)).
 * Counts are ACCUMULATED into the plan's device counters on `stream` (a cudaStream_t passed as
 * void*; NULL = default stream).  Tail truncation happens only at n_total, never at the end of a
 * shard -- this is what makes database shards with an (m-1)-byte halo exact.                      */
int apm_plan_count_device(apm_plan *plan, const unsigned char *d_buf,
                          unsigned long long buf_offset, unsigned long long buf_len,
                          unsigned long long n_total, unsigned long long j_begin,
                          unsigned long long j_end, void *stream);

/* Resident 2-bit copy of a device text for REPEATED searches (a database that stays in HBM while query batches
 * change): apm_text_pack_device writes apm_text_pack_bytes(buf_len) bytes to d_packed -- word w holds the 2-bit codes
 * ((byte >> 1) & 3) of d_buf[16 w, 16 w + 16); d_buf must be 16-byte aligned.  apm_plan_count_device_packed is
 * apm_plan_count_device with that copy of exactly the same d_buf / buf_len at hand: in filter mode the 2-bit q-gram
 * scan (ACGT pattern sets) then streams a quarter of the bytes and skips the packing; every other kernel, and the
 * byte-exact verification, still read d_buf, so the text may contain any bytes and the counts are identical.
 * d_packed = NULL makes it the plain call.                                                          */
unsigned long long apm_text_pack_bytes(unsigned long long buf_len);
int apm_text_pack_device(const unsigned char *d_buf, unsigned long long buf_len, void *d_packed, void *stream);
int apm_plan_count_device_packed(apm_plan *plan, const unsigned char *d_buf, const void *d_packed,
                                 unsigned long long buf_offset, unsigned long long buf_len,
                                 unsigned long long n_total, unsigned long long j_begin,
                                 unsigned long long j_end, void *stream);

/* Restrict the plan to the patterns p with p % world == rank (pattern sharding, mirrors
 * src/patterns_over_ranks.c:161); counters of the other patterns stay 0.  world = 1 resets.      */
int apm_plan_set_pattern_shard(apm_plan *plan, int rank, int world);

/* Attach (or, with NULLs, detach) a DEVICE array of `capacity` packed hits (pattern index << 40 | global window
 * start) and a DEVICE counter: every later apm_plan_count_device appends the matching windows it finds
 * (unordered; the counter keeps counting past the capacity).  The caller zeroes the counter.           */
int apm_plan_set_hit_buffer(apm_plan *plan, unsigned long long *d_hits, unsigned long long capacity,
                            unsigned long long *d_n_hits);

int apm_plan_zero_counts(apm_plan *plan, void *stream);
/* Device pointer to the nb_patterns unsigned 64-bit counters (for an in-place NCCL all-reduce). */
int apm_plan_counts_device_ptr(apm_plan *plan, unsigned long long **d_counts);
/* Stream-synchronises and copies the counters to the host. */
int apm_plan_read_counts(apm_plan *plan, long long *n_matches, void *stream);
int apm_plan_max_pattern_len(apm_plan *plan, int *m_max);

/* ---- utilities used by the bench / tests ------------------------------------------------------- */

/* d_out[i] = "ACGT"[splitmix64(seed + offset + i) >> 62], i in [0, count) -- the counter-based
 * synthetic text of the BASELINE configs, generated in HBM.                                       */
int apm_synth_text_device(unsigned char *d_out, unsigned long long seed, unsigned long long offset,
                          unsigned long long count, void *stream);

/* Integer-ALU peak microbenchmark on the current device: dependency-free LOP3+IADD3 streams.
 * kind 0 = LOP3+IADD3 (the roofline denominator), 1 = LOP3 only, 2 = IADD3 only, 3 = LOP3+IMAD.
 * Returns int32 lane-operations per second and the kernel time.                                   */
int apm_int_peak(int kind, double *ops_per_sec, double *seconds);

/* Number of kernels this library has launched in this process (for bench.py's gpu_launches).    */
unsigned long long apm_launch_count(void);

/* Returns the device memory (and the pinned file-staging buffers) the library keeps cached between calls to
 * the driver.  Option "cache_mb" bounds the cache (default 4096; 0 = keep nothing).                  */
int apm_release_cache(void);

const char *apm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* APM_B200_H */
