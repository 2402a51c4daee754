/*
 * oracle/kg_oracle_main.c -- command line of the CPU oracle, mirroring KmerGutsJava.main (KGJ:560-654).
 * TEST INFRASTRUCTURE ONLY (see kg_oracle.h).  PARITY PIN: see kg_oracle.c (the reference's own source,
 * executed through tests/java_pin/j2py.py, writes byte-identical reports; no JVM run yet).
 *
 *   kmer_guts_oracle [-a] [-d] [-m N] [-M N] [-O] [-g N] -D DataDir -q query.fasta[.gz] [-o out] [-V direct|stream]
 *
 * Divergences from KGJ.main, all deliberate: a flag error exits 2 instead of printing usage and then dying on
 * new File(null) (KGJ:616-647); -t/-l are accepted and ignored (they always throw there, KGJ:605-610); the
 * wall-clock lines (Temp. directory / Preparation / Lookup / Grouping time / Processed) are not printed.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include "kg_oracle.h"

static int exists(const char* p) {
    struct stat st;
    return stat(p, &st) == 0;
}

int main(int argc, char** argv) {
    kgo_params p = {0, 0, 5, 0, 200, 0}; /* defaults KGJ:102-107 */
    const char *dir = NULL, *query = NULL, *outp = NULL;
    int variant = KGO_LOOKUP_STREAM_JOIN;
    for (int i = 1; i < argc; i++) {
        const char* a = argv[i];
        if (a[0] != '-' || strlen(a) != 2) { fprintf(stderr, "Error: Unknown parameter: %s\n", a); return 2; }
        int need = strchr("mMgDqotlV", a[1]) != NULL;
        if (need && i + 1 >= argc) { fprintf(stderr, "Error: missing value for %s\n", a); return 2; }
        switch (a[1]) {
            case 'a': p.aa = 1; break;
            case 'd': p.debug = 1; break;
            case 'O': p.order_constraint = 1; break;
            case 'm': p.min_hits = atoi(argv[++i]); break;
            case 'M': p.min_weighted_hits = atoi(argv[++i]); break;
            case 'g': p.max_gap = atoi(argv[++i]); break;
            case 'D': dir = argv[++i]; break;
            case 'q': query = argv[++i]; break;
            case 'o': outp = argv[++i]; break;
            case 't': case 'l': ++i; break;
            case 'V': variant = strcmp(argv[++i], "direct") == 0 ? KGO_LOOKUP_DIRECT_PROBE : KGO_LOOKUP_STREAM_JOIN; break;
            default: fprintf(stderr, "Error: Unknown parameter: %s\n", a); return 2;
        }
    }
    if (!dir) { fprintf(stderr, "Error: -D parameter is required\n"); return 2; }
    if (!query) { fprintf(stderr, "Error: -q parameter is required\n"); return 2; }
    if (p.min_hits < 2) { fprintf(stderr, "Error: -m must be >= 2 (KGJ:442 indexes hits[n-2])\n"); return 2; }
    char path[4096], err[512];
    /* KGJ:749-758: the .gz variant wins when present */
    snprintf(path, sizeof path, "%s/kmer.table.mem_map.gz", dir);
    if (!exists(path)) snprintf(path, sizeof path, "%s/kmer.table.mem_map", dir);
    kgo_table* t = kgo_table_open(path, err, sizeof err);
    if (!t) { fprintf(stderr, "Error: %s\n", err); return 1; }
    snprintf(path, sizeof path, "%s/function.index.gz", dir);
    if (!exists(path)) snprintf(path, sizeof path, "%s/function.index", dir);
    kgo_functions* fn = kgo_functions_read(path, err, sizeof err);
    if (!fn) { fprintf(stderr, "Error: %s\n", err); return 1; }
    kgo_fasta* fa = kgo_fasta_read(query, err, sizeof err);
    if (!fa) { fprintf(stderr, "Error: %s\n", err); return 1; }
    kgo_result* r = kgo_run(t, &p, fa->seq, fa->off, fa->n, variant);
    FILE* out = outp ? fopen(outp, "w") : stdout;
    if (!out) { fprintf(stderr, "Error: cannot write %s\n", outp); return 1; }
    kgo_write_report(r, &p, fa, fn, t, out);
    if (outp) fclose(out);
    fprintf(stderr, "oracle: %zu sequences, %lld kmers, %zu hits, %zu calls\n", fa->n,
            (long long)kgo_result_num_kmers(r), kgo_result_num_hits(r), kgo_result_num_calls(r));
    kgo_result_free(r);
    kgo_fasta_free(fa);
    kgo_functions_free(fn);
    kgo_table_free(t);
    return 0;
}
