/*
 * apm_refcompat.h -- libapm_refcompat.so: the reference's GPU entry points under their ORIGINAL names,
 * implemented on the C-ABI of libapm_b200.so (include/apm_b200.h).  Link it in place of the reference's
 * patterns_over_ranks.cu / database_over_ranks.cu / cuda_utils.cu objects; the reference's .c files that
 * declare and call these functions stay as they are.
 *
 *   declaration in the reference                    definition replaced
 *   src/patterns_over_ranks.c:33-36                 src/patterns_over_ranks.cu:75-134
 *   src/database_over_ranks.c:18-22                 src/database_over_ranks.cu:137-205
 *   src/main.c:18-19                                src/cuda_utils.cu:10-35
 *   include/approaches.h:4-7                        src/patterns_over_ranks.c:38, src/database_over_ranks.c:24
 *
 * Every function computes what the reference function computes -- same window-start range, windows truncated
 * at the end of the range the caller passes -- with exact counters (the reference kernels increment shared
 * ints without atomics, patterns_over_ranks.cu:67-69).
 */
#ifndef APM_REFCOMPAT_H
#define APM_REFCOMPAT_H
#ifdef __cplusplus
extern "C" {
#endif

/* One pattern against buf[0, n_bytes): window starts [0, n_bytes - approx_factor), size = min(pattern_length,
 * n_bytes - j).  Asynchronous; returns an opaque handle for write_kernel_result (NULL on failure).      */
int *invoke_kernel(char *buf, int n_bytes, char *my_pattern, int pattern_length, int approx_factor,
                   int *local_matches);
/* *local_matches = (value of *local_matches at invoke_kernel time) + matches; releases the handle.     */
void write_kernel_result(int *local_matches, int *d_local_matches);

/* Patterns [0, lastPatternAnalyzedByGPU) against the caller's piece of buf: window starts
 * [indexStartMyPiece, end_i - approx_factor), end_i = indexFinishMyPieceWithoutExtra (+ sizePatterns[i] - 1
 * unless myRank is the last rank), windows truncated at end_i.  Asynchronous; always returns 1.        */
int initializeGPU(char *buf, int n_bytes, char **pattern, int nb_patterns, int lastPatternAnalyzedByGPU,
                  int *sizePatterns, int indexFinishMyPieceWithoutExtra, int myRank, int numberProcesses,
                  int indexStartMyPiece, int approx_factor, int *numberOfMatchesInitialized);
/* malloc'd int[nb_patterns] owned by the caller: numberOfMatchesInitialized + the GPU's matches.       */
int *getGPUResult(int nb_patterns);

void getDeviceCount(int *deviceCountPtr);
/* rank 0 (idle master) -> device 0; worker rank r -> device (r - 1) % deviceCount.                     */
void setDevice(int rank, int deviceCount);

/* The two approaches of include/approaches.h with the whole search on the node's GPUs: rank 0 parses argv,
 * prints the reference's lines and returns 0 / 1; every other rank returns 0 immediately.  Exported with an
 * apm_ prefix because the reference's own .c files define the unprefixed names (build the library with
 * -DAPM_REFCOMPAT_APPROACHES to get the unprefixed names as well).                                     */
int apm_patterns_over_ranks_hybrid(int argc, char **argv, int rank, int world_size, int cuda_device_exists);
int apm_database_over_ranks(int argc, char **argv, int myRank, int numberProcesses, int cuda_device_exists);

#ifdef __cplusplus
}
#endif
#endif /* APM_REFCOMPAT_H */
