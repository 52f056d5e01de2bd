// apm_refcompat.cu -- libapm_refcompat.so: the reference's own GPU / approach entry points, by NAME, on top of
// the C-ABI of libapm_b200.so.  Linking this library instead of the reference's patterns_over_ranks.cu,
// database_over_ranks.cu and cuda_utils.cu objects (reference Makefile:45-56) leaves main.c,
// patterns_over_ranks.c and database_over_ranks.c untouched:
//
//   symbol (reference declaration)                                         defined by the reference in
//   invoke_kernel, write_kernel_result   (patterns_over_ranks.c:33-36)     patterns_over_ranks.cu:75,115
//   initializeGPU, getGPUResult          (database_over_ranks.c:18-22)     database_over_ranks.cu:137,194
//   getDeviceCount, setDevice            (main.c:18-19)                    cuda_utils.cu:10,22
//
// and, for a driver that wants the whole approach on the GPUs (rank 0 does the job, the other ranks return),
//   patterns_over_ranks_hybrid, database_over_ranks   (include/approaches.h:4-7)
// under the names apm_patterns_over_ranks_hybrid / apm_database_over_ranks (the reference's own .c files define
// the unprefixed names, so these two cannot be exported unprefixed next to them; -DAPM_REFCOMPAT_APPROACHES
// adds the unprefixed aliases for a build that drops those .c files).
//
// Each function computes exactly what the reference function computes (same window ranges, same truncation at
// the end of the range it is given), minus its data races: the counters are exact.  No CPU fallback: without a
// device the launch functions print the library's error and leave the caller's counters untouched.
#include <cuda_runtime.h>
#include <sys/time.h>

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/apm_b200.h"

namespace {

int clamp_int(long long v) { return v > INT_MAX ? INT_MAX : (int)v; }

void report(const char *where) { fprintf(stderr, "apm_refcompat: %s: %s\n", where, apm_last_error()); }

// ---- invoke_kernel / write_kernel_result -----------------------------------------------------------
struct PorJob {
    apm_plan *plan = nullptr;
    unsigned char *d_text = nullptr;  // private copy of the text, or nullptr when the cached copy is used
    bool uses_cache = false;
    cudaStream_t st = nullptr;
    int initial = 0;
};

// The reference's worker calls invoke_kernel once per pattern with the SAME broadcast text
// (patterns_over_ranks.c:323) and its kernel wrapper re-uploads the whole buffer every time
// (patterns_over_ranks.cu:79-91): P x the PCIe traffic.  Here the device copy is kept between calls and reused
// when the next call passes the same (device, buf, n_bytes) and the buffer's fingerprint (both ends + 256 strided
// samples) is unchanged.  APM_REFCOMPAT_TEXT_CACHE=0 switches the cache off (every call uploads).
struct TextCache {
    std::mutex mu;
    int dev = -1;
    const char *buf = nullptr;
    int n_bytes = 0;
    unsigned long long fp = 0;
    unsigned char *d_text = nullptr;
    cudaEvent_t ready = nullptr;  // recorded behind the upload
    int refs = 0;                 // jobs in flight that read d_text
};
TextCache g_text;

unsigned long long text_fingerprint(const char *buf, int n) {
    unsigned long long h = 1469598103934665603ull ^ (unsigned long long)n;
    auto mix = [&](const char *p, int len) {
        for (int i = 0; i < len; ++i) h = (h ^ (unsigned char)p[i]) * 1099511628211ull;
    };
    const int edge = n < 4096 ? n : 4096;
    mix(buf, edge);
    mix(buf + n - edge, edge);
    if (n > 2 * 4096)
        for (int s = 0; s < 256; ++s) {
            const long long off = (long long)(n - 64) * s / 256;
            mix(buf + off, 64);
        }
    return h;
}

bool text_cache_enabled() {
    static const bool on = !(getenv("APM_REFCOMPAT_TEXT_CACHE") && atoi(getenv("APM_REFCOMPAT_TEXT_CACHE")) == 0);
    return on;
}

void por_release(PorJob *j) {
    if (!j) return;
    if (j->st) cudaStreamSynchronize(j->st);
    if (j->plan) apm_plan_destroy(j->plan);
    if (j->d_text) cudaFree(j->d_text);
    if (j->uses_cache) {
        std::lock_guard<std::mutex> lk(g_text.mu);
        g_text.refs--;
    }
    if (j->st) cudaStreamDestroy(j->st);
    delete j;
}

// device copy of buf[0, n_bytes) for job j, uploaded on j->st or shared with the previous calls; nullptr on failure
const unsigned char *por_text(PorJob *j, const char *buf, int n_bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (text_cache_enabled()) {
        const unsigned long long fp = text_fingerprint(buf, n_bytes);
        std::lock_guard<std::mutex> lk(g_text.mu);
        if (g_text.d_text && g_text.dev == dev && g_text.buf == buf && g_text.n_bytes == n_bytes && g_text.fp == fp) {
            if (cudaStreamWaitEvent(j->st, g_text.ready, 0) != cudaSuccess) return nullptr;
            g_text.refs++;
            j->uses_cache = true;
            return g_text.d_text;
        }
        if (g_text.refs == 0) {  // replace the cached text (nobody reads the old one any more)
            if (g_text.d_text) {
                cudaSetDevice(g_text.dev);
                cudaFree(g_text.d_text);
                cudaSetDevice(dev);
                g_text.d_text = nullptr;
            }
            if (!g_text.ready && cudaEventCreateWithFlags(&g_text.ready, cudaEventDisableTiming) != cudaSuccess) return nullptr;
            if (cudaMalloc((void **)&g_text.d_text, (size_t)n_bytes) != cudaSuccess) {
                g_text.d_text = nullptr;
                return nullptr;
            }
            if (cudaMemcpyAsync(g_text.d_text, buf, (size_t)n_bytes, cudaMemcpyHostToDevice, j->st) != cudaSuccess ||
                cudaEventRecord(g_text.ready, j->st) != cudaSuccess) {
                cudaFree(g_text.d_text);
                g_text.d_text = nullptr;
                return nullptr;
            }
            g_text.dev = dev;
            g_text.buf = buf;
            g_text.n_bytes = n_bytes;
            g_text.fp = fp;
            g_text.refs = 1;
            j->uses_cache = true;
            return g_text.d_text;
        }
    }
    // private copy: cache off, or another text is still being searched
    if (cudaMalloc((void **)&j->d_text, (size_t)n_bytes) != cudaSuccess ||
        cudaMemcpyAsync(j->d_text, buf, (size_t)n_bytes, cudaMemcpyHostToDevice, j->st) != cudaSuccess)
        return nullptr;
    return j->d_text;
}

// ---- initializeGPU / getGPUResult --------------------------------------------------------------------
struct DborState {
    std::vector<apm_plan *> plans;        // one per distinct pattern length
    std::vector<std::vector<int>> ids;    // pattern indices of every plan
    std::vector<int> result;              // numberOfMatchesInitialized, later + counts
    unsigned char *d_text = nullptr;
    cudaStream_t st = nullptr;
    bool pending = false;
};
DborState g_dbor;  // the reference keeps its result in a file-scope global too (database_over_ranks.cu:18)

void dbor_reset() {
    if (g_dbor.st) cudaStreamSynchronize(g_dbor.st);
    for (auto *p : g_dbor.plans) apm_plan_destroy(p);
    g_dbor.plans.clear();
    g_dbor.ids.clear();
    if (g_dbor.d_text) cudaFree(g_dbor.d_text);
    g_dbor.d_text = nullptr;
    if (g_dbor.st) cudaStreamDestroy(g_dbor.st);
    g_dbor.st = nullptr;
    g_dbor.pending = false;
}

// ---- whole approach on the GPUs ---------------------------------------------------------------------
int run_approach(int argc, char **argv, int rank, int world_size, const char *shard, bool por_format) {
    if (argc < 4) {  // patterns_over_ranks.c:61-68, database_over_ranks.c:46-52
        printf("Usage: %s approximation_factor dna_database pattern1 pattern2 ...\n", argv[0]);
        return 1;
    }
    if (rank != 0) return 0;  // the node's GPUs are driven by rank 0; the other ranks have nothing to do
    const int approx_factor = atoi(argv[1]);
    const char *filename = argv[2];
    const int nb_patterns = argc - 3;
    std::vector<int> len(nb_patterns);
    for (int i = 0; i < nb_patterns; i++) {
        len[i] = (int)strlen(argv[i + 3]);
        if (len[i] <= 0) {
            fprintf(stderr, "Error while parsing argument %d\n", i + 3);
            return 1;
        }
    }
    if (por_format)  // patterns_over_ranks.c:111-116 (APM_INFO)
        printf("Approximate Pattern Matching: looking for %d pattern(s) in file %s w/ distance of %d\n\n", nb_patterns,
               filename, approx_factor);
    else  // database_over_ranks.c:95-99 (sic)
        printf("Approximate Pattern Mathing: looking for %d pattern(s) in file %s w/ distance of %d\n", nb_patterns,
               filename, approx_factor);
    if (apm_set_option("shard", shard) != APM_OK || apm_set_option("gpus", "all") != APM_OK) {
        report("options");
        return 1;
    }
    std::vector<long long> n_matches(nb_patterns, 0);
    struct timeval t1, t2;
    gettimeofday(&t1, NULL);
    unsigned long long n_bytes = 0;
    const int rc = apm_count_matches_file(filename, (const char *const *)(argv + 3), len.data(), nb_patterns,
                                          approx_factor, n_matches.data(), &n_bytes);
    gettimeofday(&t2, NULL);
    if (rc != APM_OK) {
        fprintf(stderr, "%s\n", apm_last_error());
        return 1;
    }
    const double duration = (t2.tv_sec - t1.tv_sec) + ((t2.tv_usec - t1.tv_usec) / 1e6);
    const char *omp = getenv("OMP_NUM_THREADS");  // the reference calls atoi(getenv(..)) unguarded
    printf("\n(Rank %d) - TOTAL TIME using %d mpi_ranks and %d omp_thread(s) per rank: %f s\n\n", rank, world_size,
           omp ? atoi(omp) : 0, duration);  // patterns_over_ranks.c:223-226, database_over_ranks.c:200-203
    for (int i = 0; i < nb_patterns; i++) {
        if (por_format) printf("Number of matches for pattern <%.100s>: %d\n", argv[i + 3], clamp_int(n_matches[i]));
        else printf("Number of matches for pattern <%s>: %d\n", argv[i + 3], clamp_int(n_matches[i]));
    }
    return 0;
}

}  // namespace

extern "C" {

// patterns_over_ranks.cu:75-113.  Counts, for ONE pattern, the window starts [0, n_bytes - approx_factor) of
// buf[0, n_bytes) (windows truncated at n_bytes), starting from *local_matches.  Asynchronous: the returned
// pointer is an opaque job handle that must be handed to write_kernel_result (the reference returns the device
// address of its counter; callers only pass it through).
int *invoke_kernel(char *buf, int n_bytes, char *my_pattern, int pattern_length, int approx_factor, int *local_matches) {
    PorJob *j = new PorJob();
    j->initial = local_matches ? *local_matches : 0;
    const char *pats[1] = {my_pattern};
    if (apm_plan_create(pats, &pattern_length, 1, approx_factor, &j->plan) != APM_OK) {
        report("invoke_kernel");
        por_release(j);
        return NULL;
    }
    if (n_bytes > 0) {
        const unsigned char *d_text = nullptr;
        if (cudaStreamCreateWithFlags(&j->st, cudaStreamNonBlocking) != cudaSuccess || !(d_text = por_text(j, buf, n_bytes))) {
            fprintf(stderr, "apm_refcompat: invoke_kernel: %s\n", cudaGetErrorString(cudaGetLastError()));
            por_release(j);
            return NULL;
        }
        if (apm_plan_count_device(j->plan, d_text, 0, (unsigned long long)n_bytes, (unsigned long long)n_bytes, 0,
                                  (unsigned long long)n_bytes, j->st) != APM_OK) {
            report("invoke_kernel");
            por_release(j);
            return NULL;
        }
    }
    return reinterpret_cast<int *>(j);
}

// patterns_over_ranks.cu:115-134
void write_kernel_result(int *local_matches, int *d_local_matches) {
    PorJob *j = reinterpret_cast<PorJob *>(d_local_matches);
    if (!j) return;
    long long n = 0;
    if (apm_plan_read_counts(j->plan, &n, j->st) != APM_OK) report("write_kernel_result");
    else if (local_matches) *local_matches = clamp_int((long long)j->initial + n);
    por_release(j);
}

// database_over_ranks.cu:137-192 (+ the kernel :20-134).  For every pattern i < lastPatternAnalyzedByGPU:
//   end_i = indexFinishMyPieceWithoutExtra + (myRank != numberProcesses - 1 ? sizePatterns[i] - 1 : 0)
//   counts window starts [indexStartMyPiece, end_i - approx_factor) with windows truncated at end_i
// and adds them to numberOfMatchesInitialized[i]; the other patterns keep their initial value.  Returns 1.
int initializeGPU(char *buf, int n_bytes, char **pattern, int nb_patterns, int lastPatternAnalyzedByGPU, int *sizePatterns,
                  int indexFinishMyPieceWithoutExtra, int myRank, int numberProcesses, int indexStartMyPiece,
                  int approx_factor, int *numberOfMatchesInitialized) {
    dbor_reset();
    g_dbor.result.assign(numberOfMatchesInitialized, numberOfMatchesInitialized + std::max(0, nb_patterns));
    const int ngpu = std::min(lastPatternAnalyzedByGPU, nb_patterns);
    if (ngpu <= 0) return 1;
    std::map<int, std::vector<int>> by_len;
    long long max_end = 0;
    for (int i = 0; i < ngpu; i++) {
        by_len[sizePatterns[i]].push_back(i);
        long long end = indexFinishMyPieceWithoutExtra;
        if (myRank != numberProcesses - 1) end += sizePatterns[i] - 1;
        max_end = std::max(max_end, std::min<long long>(end, n_bytes));
    }
    const long long start = indexStartMyPiece;
    if (max_end <= start) return 1;
    if (cudaStreamCreateWithFlags(&g_dbor.st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc((void **)&g_dbor.d_text, (size_t)(max_end - start)) != cudaSuccess ||
        cudaMemcpyAsync(g_dbor.d_text, buf + start, (size_t)(max_end - start), cudaMemcpyHostToDevice, g_dbor.st) !=
            cudaSuccess) {
        fprintf(stderr, "apm_refcompat: initializeGPU: %s\n", cudaGetErrorString(cudaGetLastError()));
        dbor_reset();
        return 1;
    }
    for (auto &kv : by_len) {
        const int m = kv.first;
        long long end = indexFinishMyPieceWithoutExtra;
        if (myRank != numberProcesses - 1) end += m - 1;
        end = std::min<long long>(end, n_bytes);
        if (end - approx_factor <= start) continue;
        std::vector<const char *> pats;
        std::vector<int> lens;
        for (int i : kv.second) {
            pats.push_back(pattern[i]);
            lens.push_back(m);
        }
        apm_plan *pl = nullptr;
        if (apm_plan_create(pats.data(), lens.data(), (int)pats.size(), approx_factor, &pl) != APM_OK ||
            apm_plan_count_device(pl, g_dbor.d_text, (unsigned long long)start, (unsigned long long)(end - start),
                                  (unsigned long long)end, (unsigned long long)start, (unsigned long long)end,
                                  g_dbor.st) != APM_OK) {
            report("initializeGPU");
            if (pl) apm_plan_destroy(pl);
            continue;
        }
        g_dbor.plans.push_back(pl);
        g_dbor.ids.push_back(kv.second);
    }
    g_dbor.pending = true;
    return 1;
}

// database_over_ranks.cu:194-205: a malloc'd array of nb_patterns ints the caller owns.
int *getGPUResult(int nb_patterns) {
    int *out = (int *)malloc(sizeof(int) * (size_t)std::max(1, nb_patterns));
    if (!out) return NULL;
    if (g_dbor.pending) {
        for (size_t g = 0; g < g_dbor.plans.size(); g++) {
            std::vector<long long> c(g_dbor.ids[g].size());
            if (apm_plan_read_counts(g_dbor.plans[g], c.data(), g_dbor.st) != APM_OK) {
                report("getGPUResult");
                continue;
            }
            for (size_t x = 0; x < c.size(); x++) {
                int &r = g_dbor.result[(size_t)g_dbor.ids[g][x]];
                r = clamp_int((long long)r + c[x]);
            }
        }
        dbor_reset();
    }
    for (int i = 0; i < nb_patterns; i++) out[i] = i < (int)g_dbor.result.size() ? g_dbor.result[(size_t)i] : 0;
    return out;
}

// cuda_utils.cu:10-20
void getDeviceCount(int *deviceCountPtr) {
    int n = 0;
    if (apm_device_count(&n) != APM_OK) n = 0;  // the reference exits on a CUDA error; "no device" is 0 there too
    if (deviceCountPtr) *deviceCountPtr = n;
}

// cuda_utils.cu:22-35 always selects device 0; here MPI ranks of one node spread over its GPUs.
void setDevice(int rank, int deviceCount) {
    if (deviceCount == 0) {
        printf("There are no available device(s) that support CUDA\n");
        return;
    }
    const int dev = rank > 0 ? (rank - 1) % deviceCount : 0;  // rank 0 is the idle master (main.c:47-62)
    if (apm_set_device(dev) != APM_OK) report("setDevice");
    cudaSetDevice(dev);
}

// include/approaches.h:4-7
int apm_patterns_over_ranks_hybrid(int argc, char **argv, int rank, int world_size, int cuda_device_exists) {
    (void)cuda_device_exists;
    return run_approach(argc, argv, rank, world_size, "PATTERNS_OVER_RANKS", true);
}
int apm_database_over_ranks(int argc, char **argv, int myRank, int numberProcesses, int cuda_device_exists) {
    (void)cuda_device_exists;
    return run_approach(argc, argv, myRank, numberProcesses, "DB_OVER_RANKS", false);
}
#ifdef APM_REFCOMPAT_APPROACHES
int patterns_over_ranks_hybrid(int argc, char **argv, int rank, int world_size, int cuda_device_exists) {
    return apm_patterns_over_ranks_hybrid(argc, argv, rank, world_size, cuda_device_exists);
}
int database_over_ranks(int argc, char **argv, int myRank, int numberProcesses, int cuda_device_exists) {
    return apm_database_over_ranks(argc, argv, myRank, numberProcesses, cuda_device_exists);
}
#endif

}  // extern "C"
