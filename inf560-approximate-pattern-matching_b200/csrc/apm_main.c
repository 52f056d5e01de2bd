/*
 * apm_main.c -- the `apm` executable: same argv and stdout contract as the reference CLI
 * (src/sequential.c:35-50,79-82,151,157-160), the search itself delegated to libapm_b200.
 *
 *   apm <approx_factor> <dna_file> <pattern1> [pattern2 ...] [DB_OVER_RANKS|PATTERNS_OVER_RANKS]
 *
 * The optional trailing flag mirrors src/main.c:66-85 and selects how work is split over GPUs.
 * Extra knobs come from the environment so argv stays drop-in: APM_GPUS, APM_SHARD, APM_KERNEL, APM_MODE
 * (direct | band | filter -- all exact, same counts; the CLI defaults to filter), APM_CELL, APM_RBLOCK, APM_TILE,
 * APM_FILTER_SCAN, APM_INGEST_THREADS, APM_TAIL, APM_REDUCE (see include/apm_b200.h); APM_INFO=1 adds the
 * "(Rank 0) - TOTAL TIME ..." line of the parallel binary (patterns_over_ranks.c:223-226); APM_POSITIONS=n
 * additionally lists the first n matches as "Match of pattern <p> at byte j" (not in the reference);
 * APM_PATTERN_FILE=path appends one pattern per line of that file to the patterns of argv (argv is limited to
 * ~2 MB; with it `apm k file` without any argv pattern is accepted).
 */
#define _GNU_SOURCE
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

/* patterns of APM_PATTERN_FILE appended to argv (one per line, empty lines skipped) */
static char **with_pattern_file(int *argc, char **argv) {
    const char *path = getenv("APM_PATTERN_FILE");
    if (!path || !*path) return argv;
    FILE *fp = fopen(path, "rb");
    if (!fp) {
        fprintf(stderr, "Unable to open the pattern file <%s>\n", path);
        exit(1);
    }
    int cap = *argc + 1024, n = *argc;
    char **out = (char **)malloc(sizeof(char *) * (size_t)(cap + 2));
    memcpy(out, argv, sizeof(char *) * (size_t)n);
    char *line = NULL;
    size_t lcap = 0;
    ssize_t len;
    while ((len = getline(&line, &lcap, fp)) >= 0) {
        while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = 0;
        if (len == 0) continue;
        if (n == cap) {
            cap *= 2;
            out = (char **)realloc(out, sizeof(char *) * (size_t)(cap + 2));
        }
        out[n++] = strdup(line);
    }
    free(line);
    fclose(fp);
    out[n] = NULL;
    *argc = n;
    return out;
}

int main(int argc, char **argv) {
    if (argc >= 3) { /* an explicit approach stays the LAST argv word: take it off before appending file patterns */
        const char *lastw = argv[argc - 1];
        const int has_flag = argc >= 4 && (!strcmp(lastw, "DB_OVER_RANKS") || !strcmp(lastw, "PATTERNS_OVER_RANKS"));
        if (getenv("APM_PATTERN_FILE") && *getenv("APM_PATTERN_FILE")) {
            int n = argc - (has_flag ? 1 : 0);
            char **v = with_pattern_file(&n, argv);
            if (has_flag) { /* with_pattern_file leaves room for it */
                v[n++] = (char *)lastw;
                v[n] = NULL;
            }
            argc = n;
            argv = v;
        }
    }
    if (argc < 4) { /* sequential.c:35-41 */
        printf("Usage: %s approximation_factor dna_database pattern1 pattern2 ...\n", argv[0]);
        return 1;
    }
    /* every mode is exact; the CLI defaults to the fastest one (the library default is "direct": every DP cell) */
    apm_set_option("mode", "filter");
    if (set_from_env("APM_GPUS", "gpus") || set_from_env("APM_SHARD", "shard") ||
        set_from_env("APM_KERNEL", "kernel") || set_from_env("APM_RBLOCK", "rblock") ||
        set_from_env("APM_TILE", "tile") || set_from_env("APM_MODE", "mode") || set_from_env("APM_CELL", "cell") ||
        set_from_env("APM_FILTER_SCAN", "filter_scan") || set_from_env("APM_INGEST_THREADS", "ingest_threads") ||
        set_from_env("APM_TAIL", "tail") || set_from_env("APM_REDUCE", "reduce"))
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

    /* CUDA context creation (hundreds of ms) is start-up cost like the reference's MPI_Init / file read, which its
     * timer excludes too (sequential.c:84 vs :102): create the context before the clock starts */
    {
        int ndev = 0;
        if (apm_device_count(&ndev) == APM_OK && ndev > 0) {
            const char *g = getenv("APM_GPUS");
            int want = g && *g ? (!strcmp(g, "all") ? ndev : atoi(g)) : 1;
            if (want > ndev) want = ndev;
            for (int d = want - 1; d >= 0; --d) apm_set_device(d); /* ends on device 0 */
        }
    }
    struct timeval t1, t2;
    unsigned long long n_bytes = 0;
    const long long want_pos = getenv("APM_POSITIONS") ? atoll(getenv("APM_POSITIONS")) : 0;
    int *hit_pattern = NULL;
    unsigned long long *hit_start = NULL, n_hits = 0;
    int rc;
    if (want_pos > 0) { /* positions need the text in host memory: read it as utils.c:12-68 does, with 64-bit sizes */
        FILE *fp = fopen(filename, "rb");
        unsigned char *buf = NULL;
        if (fp && fseek(fp, 0, SEEK_END) == 0) {
            const long sz = ftell(fp);
            rewind(fp);
            buf = (unsigned char *)malloc(sz > 0 ? (size_t)sz : 1);
            if (buf && sz >= 0 && fread(buf, 1, (size_t)sz, fp) == (size_t)sz) n_bytes = (unsigned long long)sz;
            else { free(buf); buf = NULL; }
        }
        if (fp) fclose(fp);
        if (!buf) {
            fprintf(stderr, "Unable to open the text file <%s>\n", filename);
            return 1;
        }
        hit_pattern = (int *)malloc(sizeof(int) * (size_t)want_pos);
        hit_start = (unsigned long long *)malloc(sizeof(unsigned long long) * (size_t)want_pos);
        gettimeofday(&t1, NULL);
        rc = apm_find_matches(buf, (size_t)n_bytes, (const char *const *)(argv + 3), len, nb_patterns, approx_factor,
                              n_matches, (unsigned long long)want_pos, hit_pattern, hit_start, &n_hits);
        gettimeofday(&t2, NULL);
        free(buf);
    } else {
        gettimeofday(&t1, NULL);
        rc = apm_count_matches_file(filename, (const char *const *)(argv + 3), len, nb_patterns, approx_factor, n_matches,
                                    &n_bytes);
        gettimeofday(&t2, NULL);
    }
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
    for (unsigned long long h = 0; h < n_hits && h < (unsigned long long)want_pos; h++)
        printf("Match of pattern <%s> at byte %llu\n", argv[hit_pattern[h] + 3], hit_start[h]);
    free(hit_pattern);
    free(hit_start);
    free(len);
    free(n_matches);
    return 0;
}
