/*
 * oracle/kg_oracle.h -- CPU restatement of KmerGutsJava's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under kmergutsjava_b200/ may include, link
 * or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or as
 * the timed CPU baseline -- never as a product path.
 *
 * PARITY PIN: the reference (lib/src/kmergutsjava/KmerGutsJava.java, "KGJ"
 * below) ships no k-mer table, no golden output and no asserting test
 * (test/src/kmergutsjava/test/KmerGutsJavaServerTest.java:76-86), and no JVM
 * exists in this image.  What is checked against the reference ITSELF: its
 * unmodified source, transliterated statement by statement into Python by
 * tests/java_pin/j2py.py (a Java-syntax tool that knows nothing about k-mers)
 * and executed, writes reports byte-identical to this restatement's for the
 * eight configs[0] runs, malformed tables, odd FASTA text and the FSM vectors
 * (tests/test_java_transliteration.py; full-size record:
 * tests/golden/java_transliteration_pin.json).  NOT yet run on a real JVM
 * (tests/java_pin/pin_oracle.sh does that in one command).  Further pins:
 * (i) hand-traced known-answer vectors (tests/golden/), (ii) an independently
 * written pure Python restatement (oracle/kg_oracle_py.py), (iii) the equality
 * of the two lookup variants (sort-merge stream join == direct probe).
 */
#ifndef KG_ORACLE_H
#define KG_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KGO_K 8                              /* KGJ:85 */
#define KGO_CORE 1280000000LL                /* 20^7, KGJ:86 */
#define KGO_MAX_ENCODED 25600000000LL        /* 20^8, KGJ:87 */
#define KGO_MAX_HITS_PER_SEQ 40000           /* KGJ:98 */
#define KGO_OI_BUFSZ 5                       /* KGJ:99 */

/* lookup variants */
#define KGO_LOOKUP_STREAM_JOIN 0 /* faithful: comparator sort + one pass over the table stream (KGJ:944-1034) */
#define KGO_LOOKUP_DIRECT_PROBE 1 /* linear probing without wrap into the same image; proven equal in tests */

typedef struct {
    int32_t aa;                 /* -a  KGJ:102,578 */
    int32_t order_constraint;   /* -O  KGJ:103,590 */
    int32_t min_hits;           /* -m  KGJ:104,584 */
    int32_t min_weighted_hits;  /* -M  KGJ:105,587 */
    int32_t max_gap;            /* -g  KGJ:106,593 */
    int32_t debug;              /* -d  KGJ:107,581 */
} kgo_params;

/* One table hit on one (sequence, strand, frame) container: KGJ:1213-1219 plus its container. */
typedef struct {
    int32_t seq;            /* index of the FASTA record */
    int32_t strand_frame;   /* 0..5 = +0,+1,+2,-0,-1,-2 (aa mode: always 0) */
    int32_t pos;            /* from0InProt */
    int32_t oI;
    int32_t avg_off_from_end;
    int32_t fI;
    float function_wt;
} kgo_hit_rec;

/* One CALL line: KGJ:398-404. */
typedef struct {
    int32_t seq;
    int32_t strand_frame;
    int32_t start;          /* hits[0].from0InProt */
    int32_t end;            /* hits[lastHit].from0InProt + K-1 */
    int32_t count;          /* fICount */
    int32_t fI;             /* currentFI */
    float weighted;         /* weightedHits (fp32 sum in list order) */
    int32_t hits_before;    /* number of HIT lines of this container printed before this CALL (debug interleave) */
} kgo_call_rec;

/* OTU-COUNTS line of one sequence: KGJ:516-524. */
typedef struct {
    int32_t n;                      /* 0..5 */
    int32_t count[KGO_OI_BUFSZ];
    int32_t oI[KGO_OI_BUFSZ];
} kgo_otu_rec;

typedef struct kgo_table kgo_table;
typedef struct kgo_result kgo_result;

/* ---- scalar functions (one per reference function) ---- */
int kgo_to_amino_acid_off(int c);                       /* KGJ:111-175 */
int kgo_compl(int c);                                   /* KGJ:177-260 */
void kgo_rev_comp(const uint8_t* data, size_t n, uint8_t* out); /* KGJ:263-272 */
int64_t kgo_encoded_kmer(const uint8_t* codes, size_t pos);     /* KGJ:274-292 */
int kgo_dna_char(int c);                                /* KGJ:294-318 */
/* KGJ:320-343; pseq/piseq have length plen (= L/3+1 in prepareQuery) */
void kgo_translate(const uint8_t* seq, size_t L, int off, uint8_t* pseq, uint8_t* piseq, size_t plen);
char kgo_genetic_code(int idx);                         /* KGJ:88-93 */

/* ---- table (kmer.table.mem_map[.gz], KGJ:924-942, 995-999) ---- */
kgo_table* kgo_table_open(const char* path, char* err, size_t errlen);
kgo_table* kgo_table_from_memory(const void* bytes, size_t nbytes); /* copies */
void kgo_table_free(kgo_table* t);
int64_t kgo_table_num_sigs(const kgo_table* t);
int64_t kgo_table_entry_size(const kgo_table* t);
int64_t kgo_table_version(const kgo_table* t);

/* ---- the path: prepareQuery -> sort -> lookup -> gatherHits (KGJ:778-818) ---- */
kgo_result* kgo_run(const kgo_table* t, const kgo_params* p, const uint8_t* seq_bytes,
                    const uint64_t* offsets /* n+1 */, size_t n, int variant);
/* `threads` independent runs of the (single-threaded) reference algorithm on contiguous shards of the sequences. */
kgo_result* kgo_run_parallel(const kgo_table* t, const kgo_params* p, const uint8_t* seq_bytes,
                             const uint64_t* offsets, size_t n, int variant, int threads);
/* Adopt (no copy) a file image the caller keeps alive: for the 10 GB tables of the large configurations. */
kgo_table* kgo_table_borrow(const void* bytes, size_t nbytes);
void kgo_result_free(kgo_result* r);
size_t kgo_result_num_hits(const kgo_result* r);
const kgo_hit_rec* kgo_result_hits(const kgo_result* r);
size_t kgo_result_num_calls(const kgo_result* r);
const kgo_call_rec* kgo_result_calls(const kgo_result* r);
size_t kgo_result_num_otus(const kgo_result* r);           /* = n sequences */
const kgo_otu_rec* kgo_result_otus(const kgo_result* r);
int64_t kgo_result_num_kmers(const kgo_result* r);         /* valid windows enumerated (addKmers) */
int64_t kgo_result_kmers_found(const kgo_result* r);       /* "Kmers found" KGJ:1005 (stream-join only) */
int kgo_result_lookup_error(const kgo_result* r);          /* 1 = lookup aborted by EOF (KGJ:799-802) */

/* FSM alone (KGJ:457-514 + 385-455) on caller-supplied hits of ONE container; used by the KAT tests.
 * otu is in/out (shared across the frames of a contig).  Returns number of calls written (<= max_calls). */
size_t kgo_gather_hits(const kgo_params* p, kgo_hit_rec* hits, size_t nhits, kgo_otu_rec* otu,
                       kgo_call_rec* calls, size_t max_calls);

/* ---- host-side pieces of run(): function.index (KGJ:345-373), FASTA (KGJ:1132-1192), report ---- */
typedef struct {
    size_t n;
    char** id;          /* first token of the caption */
    uint8_t* seq;       /* concatenation */
    uint64_t* off;      /* n+1 */
} kgo_fasta;
kgo_fasta* kgo_fasta_read(const char* path, char* err, size_t errlen);
void kgo_fasta_free(kgo_fasta* f);

typedef struct { size_t n; char** name; } kgo_functions;
kgo_functions* kgo_functions_read(const char* path, char* err, size_t errlen);
void kgo_functions_free(kgo_functions* f);

/* Java String.format("%f"/"%1.3f") of a float argument: shortest decimal of the widened double, HALF_UP. */
void kgo_java_format_f(float v, int prec, char* out, size_t outlen);

/* Text report exactly as run() prints it to the PrintWriter (KGJ:810-818 incl. duplicate-id collapse). */
int kgo_write_report(const kgo_result* r, const kgo_params* p, const kgo_fasta* fa,
                     const kgo_functions* fn, const kgo_table* t, FILE* out);

#ifdef __cplusplus
}
#endif
#endif
