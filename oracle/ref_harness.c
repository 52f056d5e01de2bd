/*
 * ref_harness.c -- thin driver around the REFERENCE's own, unmodified levenshtein().
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/apm_oracle.c header).  This file contains no reference
 * code: it declares the reference's public function from include/utils.h:10 and is linked, by
 * oracle/Makefile, against /root/reference/src/utils.c compiled where it lies into
 * oracle/_ref/libapm_ref.so.  It reproduces the reference's work splits so the "host MPI+OpenMP"
 * baseline can be timed without MPI (absent from this image):
 *   mode 0  serial, the loop of src/sequential.c:105-144
 *   mode 1  positions split over OpenMP threads for one pattern at a time
 *           (src/patterns_over_ranks.c:353-357; the racy `local_matches++` at :373 replaced by a
 *           reduction)
 *   mode 2  patterns split over OpenMP threads, each thread scanning the text serially
 *           (src/database_over_ranks.c:435; correct seams because there is one piece)
 */
#include <stdlib.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* include/utils.h:10 of the reference */
int levenshtein(char *s1, char *s2, int len, int *column);

static long long scan_range(char *text, long long n_bytes, char *pattern, int m, int k,
                            long long j0, long long j1, int *column) {
    long long hits = 0;
    for (long long j = j0; j < j1; j++) {
        long long left = n_bytes - j;
        int size = left < (long long)m ? (int)left : m;
        if (levenshtein(pattern, text + j, size, column) <= k) hits++;
    }
    return hits;
}

int ref_count_matches(const unsigned char *text, long long n_bytes, const unsigned char *pat_bytes,
                      const long long *pat_off, const int *pat_len, int nb_patterns, int k,
                      int mode, int threads, long long *n_matches) {
    if (k < 0 || nb_patterns < 0) return 1;
    long long limit = n_bytes - (long long)k;
    if (limit < 0) limit = 0;
#ifndef _OPENMP
    mode = 0;
#endif
    if (threads < 1) threads = 1;
    if (mode == 0 || threads == 1) {
        for (int i = 0; i < nb_patterns; i++) {
            int m = pat_len[i];
            int *column = (int *)malloc(((size_t)m + 1) * sizeof(int));
            if (!column) return 2;
            n_matches[i] = scan_range((char *)text, n_bytes, (char *)(pat_bytes + pat_off[i]), m, k,
                                      0, limit, column);
            free(column);
        }
        return 0;
    }
#ifdef _OPENMP
    if (mode == 1) {
        for (int i = 0; i < nb_patterns; i++) {
            int m = pat_len[i];
            long long hits = 0;
#pragma omp parallel num_threads(threads) reduction(+ : hits)
            {
                int *column = (int *)malloc(((size_t)m + 1) * sizeof(int));
                int t = omp_get_thread_num(), nt = omp_get_num_threads();
                long long per = (limit + nt - 1) / nt;
                long long j0 = per * t, j1 = j0 + per;
                if (j1 > limit) j1 = limit;
                if (column && j0 < j1)
                    hits += scan_range((char *)text, n_bytes, (char *)(pat_bytes + pat_off[i]), m, k,
                                       j0, j1, column);
                free(column);
            }
            n_matches[i] = hits;
        }
        return 0;
    }
    /* mode 2 */
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int i = 0; i < nb_patterns; i++) {
        int m = pat_len[i];
        int *column = (int *)malloc(((size_t)m + 1) * sizeof(int));
        n_matches[i] = column ? scan_range((char *)text, n_bytes, (char *)(pat_bytes + pat_off[i]),
                                           m, k, 0, limit, column)
                              : -1;
        free(column);
    }
#endif
    return 0;
}

int ref_levenshtein(const unsigned char *a, const unsigned char *b, int len) {
    int *column = (int *)malloc(((size_t)(len > 0 ? len : 0) + 1) * sizeof(int));
    if (!column) return -1;
    int d = levenshtein((char *)a, (char *)b, len, column);
    free(column);
    return d;
}

int ref_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
