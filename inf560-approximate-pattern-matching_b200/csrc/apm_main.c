/*
 * apm_main.c -- the `apm` executable: same argv and stdout contract as the reference CLI
 * (src/sequential.c:35-50,79-82,151,157-160), the search itself delegated to libapm_b200.
 *
 *   apm <approx_factor> <dna_file> <pattern1> [pattern2 ...] [DB_OVER_RANKS|PATTERNS_OVER_RANKS]
 *
 * The optional trailing flag mirrors src/main.c:66-85 and selects how work is split over GPUs.
 * Extra knobs come from the environment so argv stays drop-in: APM_GPUS, APM_SHARD, APM_KERNEL, APM_MODE
 * (direct | band | filter -- all exact), APM_CELL, APM_RBLOCK, APM_TILE; APM_INFO=1 adds the
 * "(Rank 0) - TOTAL TIME ..." line of the parallel binary (patterns_over_ranks.c:223-226).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "../../include/apm_b200.h"

static int set_from_env(const char *env, const char *key) {
    const char *v = getenv(env);
    if (v && *v && apm_set_option(key, v) != APM_OK) {
        fprintf(stderr, "%s=%s: %s\n", env, v, apm_last_error());
        return 1;
    }
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 4) { /* sequential.c:35-41 */
        printf("Usage: %s approximation_factor dna_database pattern1 pattern2 ...\n", argv[0]);
        return 1;
    }
    if (set_from_env("APM_GPUS", "gpus") || set_from_env("APM_SHARD", "shard") ||
        set_from_env("APM_KERNEL", "kernel") || set_from_env("APM_RBLOCK", "rblock") ||
        set_from_env("APM_TILE", "tile") || set_from_env("APM_MODE", "mode") || set_from_env("APM_CELL", "cell"))
        return 1;

    /* main.c:66-85: an explicit approach as last argument is consumed, not searched for */
    const char *last = argv[argc - 1];
    if (argc >= 5 && (!strcmp(last, "DB_OVER_RANKS") || !strcmp(last, "PATTERNS_OVER_RANKS"))) {
        apm_set_option("shard", last);
        argc -= 1;
    }

    const int approx_factor = atoi(argv[1]); /* sequential.c:44 */
    const char *filename = argv[2];
    const int nb_patterns = argc - 3;
    int *len = (int *)malloc(sizeof(int) * nb_patterns);
    long long *n_matches = (long long *)malloc(sizeof(long long) * nb_patterns);
    if (!len || !n_matches) {
        fprintf(stderr, "Unable to allocate array of pattern of size %d\n", nb_patterns);
        return 1;
    }
    for (int i = 0; i < nb_patterns; i++) {
        len[i] = (int)strlen(argv[i + 3]);
        if (len[i] <= 0) { /* sequential.c:64-67 */
            fprintf(stderr, "Error while parsing argument %d\n", i + 3);
            return 1;
        }
    }

    printf("Approximate Pattern Mathing: looking for %d pattern(s) in file %s w/ distance of %d\n",
           nb_patterns, filename, approx_factor); /* sic -- sequential.c:79-82 */

    struct timeval t1, t2;
    gettimeofday(&t1, NULL);
    unsigned long long n_bytes = 0;
    const int rc = apm_count_matches_file(filename, (const char *const *)(argv + 3), len, nb_patterns,
                                          approx_factor, n_matches, &n_bytes);
    gettimeofday(&t2, NULL);
    if (rc != APM_OK) {
        fprintf(stderr, "%s\n", apm_last_error());
        return 1;
    }
    const double duration = (t2.tv_sec - t1.tv_sec) + ((t2.tv_usec - t1.tv_usec) / 1e6);
    printf("APM done in %lf s\n", duration);
    if (getenv("APM_INFO") && atoi(getenv("APM_INFO")) > 0) { /* patterns_over_ranks.c:223-226 */
        const char *omp = getenv("OMP_NUM_THREADS");
        printf("\n(Rank 0) - TOTAL TIME using %d mpi_ranks and %d omp_thread(s) per rank: %f s\n\n", 1, omp ? atoi(omp) : 0,
               duration);
    }
    for (int i = 0; i < nb_patterns; i++)
        printf("Number of matches for pattern <%s>: %lld\n", argv[i + 3], n_matches[i]);
    free(len);
    free(n_matches);
    return 0;
}
