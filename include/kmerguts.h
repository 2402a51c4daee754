/*
 * include/kmerguts.h -- flat C ABI of libkmerguts_b200.so
 *
 * The B200 (sm_100a) replacement for the middle of KmerGutsJava.run()
 * (lib/src/kmergutsjava/KmerGutsJava.java, "KGJ" below): everything between the
 * FASTA callback (KGJ:778-784) and the per-sequence printing loop (KGJ:810-818),
 * i.e. prepareQuery + QueryKmerStorage + lookup + gatherHits/processSetOfHits.
 * The reference has no FFI for this path (it is one Java file); these entry
 * points are what a JNI / Panama binding inside run() would bind -- see
 * INTEGRATION.md for the Java side.
 *
 * Conventions: plain pointers and sizes only; every int-returning function gives
 * 0 on success and a negative KG_E* code on failure, with a message available
 * from kg_last_error(); the library owns every buffer it returns until the
 * matching *_free; inputs are borrowed only for the duration of the call.
 * There is NO CPU fallback: without a CUDA device kg_init fails with KG_ENODEV.
 *
 * Threading (the reference path is single-threaded, KGJ has no Thread/executor): a
 * context runs one call at a time -- it owns the streams, scratch and buffer pools --
 * so a multi-threaded host uses one context per thread (and one per GPU).  A table
 * handle is immutable after loading; kg_last_error() is thread-local.
 */
#ifndef KMERGUTS_H
#define KMERGUTS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KG_K 8                          /* KGJ:85 */
#define KG_MAX_ENCODED 25600000000LL    /* 20^8, KGJ:87 */
#define KG_OI_BUFSZ 5                   /* KGJ:99 */
#define KG_MAX_HITS_PER_SEQ 40000       /* KGJ:98 */

enum {
    KG_OK = 0,
    KG_EINVAL = -1,     /* bad argument (incl. min_hits < 2, see kg_params) */
    KG_ENODEV = -2,     /* no usable CUDA device / driver */
    KG_ECUDA = -3,      /* CUDA runtime error (message has the detail) */
    KG_EIO = -4,        /* file missing / unreadable / truncated */
    KG_EFORMAT = -5,    /* kmer.table.mem_map violates the format (entrySize != 24, ...) */
    KG_ENOMEM = -6,     /* host or device allocation failed */
    KG_ERANGE = -7      /* batch too large for one call (positions must fit 32 bits after splitting) */
};

enum { KG_MODE_DNA = 0, KG_MODE_AA = 1 }; /* "-a" flag, KGJ:102, 578, 1054 */

typedef struct kg_context kg_context; /* one GPU: device id, streams, scratch arena */
typedef struct kg_table kg_table;     /* immutable GPU-resident signature table; shareable between runs */
typedef struct kg_batch kg_batch;     /* sequences resident in HBM */
typedef struct kg_result kg_result;   /* calls / OTU counts / optional hits of one run */

/* Instance fields of KmerGutsJava that steer the path (KGJ:102-107, flags KGJ:578-595). */
typedef struct kg_params {
    int32_t min_hits;          /* -m, default 5.  Must be >= 2: KGJ:442 indexes hits[n-2] */
    int32_t min_weighted_hits; /* -M, default 0 (compared as float, KGJ:397) */
    int32_t max_gap;           /* -g, default 200 */
    int32_t order_constraint;  /* -O, default 0 */
    int32_t emit_hits;         /* 1: also return every hit (the "-d" HIT lines, KGJ:472-475) */
} kg_params;

/* One CALL line (KGJ:398-404), in the order the reference prints them: by sequence, then
 * strand/frame (+0,+1,+2,-0,-1,-2), then position. */
typedef struct kg_call {
    uint32_t seq;          /* index of the sequence in the batch */
    int32_t strand_frame;  /* 0..5 = +0,+1,+2,-0,-1,-2; always 0 in aa mode */
    int32_t start;         /* hits.get(0).from0InProt */
    int32_t end;           /* hits.get(lastHit).from0InProt + K-1 */
    int32_t count;         /* fICount */
    int32_t fI;            /* currentFI: index into function.index */
    float weighted;        /* weightedHits: fp32 sum in list order (bit-exact with the reference) */
    int32_t hits_before;   /* HIT lines of this (seq, strand_frame) printed before this CALL (for "-d" output) */
} kg_call;

/* One OTU-COUNTS line (KGJ:516-524): the top-5 buffer of KGJ:413-438 after the sequence's last frame. */
typedef struct kg_otu {
    int32_t n;                  /* entries in use, 0..5 */
    int32_t count[KG_OI_BUFSZ];
    int32_t oI[KG_OI_BUFSZ];
} kg_otu;

typedef struct kg_otu_entry { /* one used slot of the OTU buffer (OtuCount, KGJ:1221-1224) */
    int32_t count;
    int32_t oI;
} kg_otu_entry;

/* One table hit (Hit, KGJ:1213-1219, plus its container), sorted by (seq, strand_frame, pos). */
typedef struct kg_hit {
    uint32_t seq;
    int32_t strand_frame;
    int32_t pos;              /* from0InProt */
    int32_t oI;
    int32_t avg_off_from_end;
    int32_t fI;
    float function_wt;
} kg_hit;

typedef struct kg_table_info {
    int64_t num_slots;        /* "numSigs" header field (KGJ:933): slot count and hash modulus of the file */
    int64_t entry_size;       /* header field, must be 24 (KGJ:934, 992 vs 995-999) */
    int64_t version;          /* header field (KGJ:935), never checked by the reference either */
    int64_t num_signatures;   /* occupied slots that the reference's no-wrap probe chain can reach */
    int64_t num_unreachable;  /* occupied slots the reference could never return (dropped) */
    int64_t tail_run;         /* occupied slots at the very end of the file: probes into them can run off
                                 the end in the reference (EOFException, KGJ:799-802); 0 for well-formed tables */
    int64_t num_buckets;      /* 128-byte buckets (six keys + six payloads) of the GPU layout, incl. the spill tail */
    int64_t flagged_buckets;  /* buckets whose overflow flag is set (a miss there costs a second line) */
    int64_t device_bytes;     /* HBM used by the table */
} kg_table_info;

typedef struct kg_run_stats {
    uint64_t num_sequences;
    uint64_t num_positions;   /* residue positions scanned (all frames) */
    uint64_t num_kmers;       /* valid 8-mer windows = lookups (what addKmers enumerates, KGJ:912-921) */
    uint64_t num_hits;
    uint64_t num_calls;
    uint32_t num_launches;    /* kernels of this library launched for the run */
    float ms_h2d, ms_device, ms_d2h; /* CUDA-event times of the last run (0 when not applicable) */
    float ms_prepare, ms_probe, ms_group; /* inside ms_device: translate/patch, encode+probe stage, gather+FSM */
    float ms_filter, ms_refilter, ms_lines; /* inside ms_probe when the probe ran as the three-kernel cascade (else 0):
                                               encode + first prefilter, second prefilter, bucket-line probe */
    uint64_t num_survivors1, num_survivors2; /* cascade: lookups that passed the first / both prefilters (hits included) */
} kg_run_stats;

/* ---- context ---- */
int kg_init(int device, kg_context** ctx);
void kg_shutdown(kg_context* ctx);
const char* kg_last_error(void); /* thread-local, valid until the next failing call on this thread */
const char* kg_version(void);

/* ---- table: replaces readKmerTableHeader (KGJ:924-942) + the streamed table of lookup (KGJ:944-1034) ---- */
/* data_dir holds kmer.table.mem_map or kmer.table.mem_map.gz (the .gz wins when both exist, KGJ:749-753). */
int kg_table_load(kg_context* ctx, const char* data_dir, kg_table** table);
int kg_table_load_file(kg_context* ctx, const char* path, kg_table** table);
/* Same, from a file image already in memory (24-byte header + 24-byte little-endian entries). */
int kg_table_from_image(kg_context* ctx, const void* image, size_t nbytes, kg_table** table);
/* From DEVICE arrays of n distinct keys (< 20^8) and 16-byte payloads {oI, avgFromEnd, fI, wt-bits}. */
int kg_table_from_device_entries(kg_context* ctx, const uint64_t* d_keys, const void* d_payload16, size_t n,
                                 kg_table** table);
/* Cache of the built GPU layout (bucket lines + prefilter): written once, read back at file speed instead of parsing --
 * and, for .gz, inflating -- the reference-format file on every run (the reference re-streams it on every run,
 * KGJ:944-1034).  kg_table_load_cached fails with KG_EFORMAT on a file of another build / layout; fall back to kg_table_load. */
int kg_table_save(kg_context* ctx, const kg_table* table, const char* path);
int kg_table_load_cached(kg_context* ctx, const char* path, kg_table** table);
/* Same, but only if the cache was built from the kmer.table.mem_map[.gz] that data_dir holds NOW (its size and mtime are
 * recorded in the cache header); KG_EFORMAT otherwise.  Both loaders verify a checksum over the whole body. */
int kg_table_load_cached_checked(kg_context* ctx, const char* path, const char* data_dir, kg_table** table);
/* A table may serve several contexts of ITS device (one context per host thread: two threads that alternate batches
 * overlap one call's copies with the other's kernels).  Attach it to every additional context once, so that the
 * context's stream keeps the table's prefilter resident in L2 as well. */
int kg_table_attach(kg_context* ctx, const kg_table* table);
int kg_table_get_info(const kg_table* table, kg_table_info* info);
void kg_table_free(kg_table* table);

/* ---- the path ---- */
void kg_params_default(kg_params* p);

/* Host buffers in, host results out: prepareQuery + lookup + gatherHits for n sequences.
 * seq_bytes = the sequences exactly as FastaCallback.nextEntry receives them (KGJ:780), concatenated;
 * offsets[n+1] delimit them.  Includes the H2D and D2H copies (this is the end-to-end call). */
int kg_run(kg_context* ctx, const kg_table* table, int mode, const uint8_t* seq_bytes, const uint64_t* offsets,
           size_t n, const kg_params* params, kg_result** result);

/* The same call for proteins that are ALREADY residue codes, 5 bits each (the host feed, ~1 byte per lookup
 * over PCIe, is the bottleneck of the end-to-end call).  prepareQuery maps every character through toAminoAcidOff before
 * addKmers sees it (KGJ:1055-1058); kg_pack_aa does that mapping on the host and packs 8 codes into 5 bytes:
 *   code 0..19 = residue (PROT_ALPHA order, KGJ:94-96), 20 = any other character (KGJ:173), 31 = padding.
 * Sequence s occupies groups [group_offsets[s], group_offsets[s+1]) of 5 bytes, kg_pack_aa_groups(len) = ceil((len+1)/8) of
 * them: at least one padding code follows the last residue.  Results are identical to kg_run on the original characters.
 * kg_pack_aa with packed == NULL only fills group_offsets (so the caller can size `packed`: 5 * group_offsets[n] bytes). */
uint64_t kg_pack_aa_groups(uint64_t len);
int kg_pack_aa(const uint8_t* seq_bytes, const uint64_t* offsets, size_t n, uint8_t* packed, uint64_t* group_offsets, int threads);
int kg_run_packed_aa(kg_context* ctx, const kg_table* table, const uint8_t* packed, const uint64_t* group_offsets, size_t n,
                     const kg_params* params, kg_result** result);

/* The same for 6-frame input: prepareQuery / translate only ever look at dnaChar() of a nucleotide (KGJ:294-318, 320-343), so a
 * contig travels as 2 bits per nucleotide (aA 0, cC 1, gG 2, tTuU 3; four per byte, low bits first, every contig starting on
 * a byte: contig s occupies bytes [byte_offsets[s], byte_offsets[s+1]) = ceil(len / 4)) plus the sorted positions -- in the
 * space of `offsets` -- of every other character (dnaChar 4).  kg_pack_dna with packed == NULL fills byte_offsets and
 * *n_exceptions only; the second call wants *n_exceptions = the room in `exceptions`.  Results are identical to kg_run. */
int kg_pack_dna(const uint8_t* seq_bytes, const uint64_t* offsets, size_t n, uint8_t* packed, uint64_t* byte_offsets,
                uint64_t* exceptions, size_t* n_exceptions, int threads);
int kg_run_packed_dna(kg_context* ctx, const kg_table* table, const uint8_t* packed, const uint64_t* offsets,
                      const uint64_t* byte_offsets, const uint64_t* exceptions, size_t n_exceptions, size_t n,
                      const kg_params* params, kg_result** result);

/* Split form: keep the sequences resident in HBM and run the device pipeline on them (possibly many times). */
int kg_batch_upload(kg_context* ctx, int mode, const uint8_t* seq_bytes, const uint64_t* offsets, size_t n,
                    kg_batch** batch);
/* Adopt sequences that are ALREADY in device memory (d_seq_bytes is modified in place in aa mode; d_offsets is a
 * device array of n+1 uint64).  d_seq_bytes must be 16-byte aligned and followed by at least 64 readable bytes (the
 * kernels read whole 16-byte words).  The batch does not own the two buffers. */
int kg_batch_from_device(kg_context* ctx, int mode, uint8_t* d_seq_bytes, const uint64_t* d_offsets, size_t n,
                         uint64_t total_bytes, kg_batch** batch);
void kg_batch_free(kg_batch* batch);
/* Device pipeline only; results stay on the device until kg_result_fetch. */
int kg_batch_run(kg_context* ctx, const kg_table* table, kg_batch* batch, const kg_params* params,
                 kg_result** result);
/* n resident batches (the same batch may appear more than once), two in flight: the run FSM of batch i overlaps the probe of
 * batch i+1.  results[i] is what kg_batch_run(batches[i]) would have returned. */
int kg_batch_run_many(kg_context* ctx, const kg_table* table, kg_batch* const* batches, size_t n, const kg_params* params,
                      kg_result** results);
/* The same pipeline with the caller in the loop: at most two batches in flight per context; kg_batch_collect returns the
 * results in submission order.  submit(0); for i: submit(i+1); collect -> result i; use it; kg_result_free. */
int kg_batch_submit(kg_context* ctx, const kg_table* table, kg_batch* batch, const kg_params* params);
int kg_batch_collect(kg_context* ctx, kg_result** result);
int kg_result_fetch(kg_result* result); /* D2H of calls, OTU counts and (if requested) hits; idempotent */

int kg_result_stats(const kg_result* result, kg_run_stats* stats);
int kg_result_calls(kg_result* result, const kg_call** calls, size_t* n);
int kg_result_otus(kg_result* result, const kg_otu** otus, size_t* n); /* one per sequence */
/* The same counts without the unused slots: n_per_seq[s] entries belong to sequence s, in buffer order, consecutively in
 * `entries`.  This is the form kg_run / kg_run_packed_aa copy back from the GPU (most sequences use 0-2 of the 5 slots);
 * kg_result_otus expands it on first use. */
int kg_result_otus_compact(kg_result* result, const uint8_t** n_per_seq, const kg_otu_entry** entries, size_t* n_seqs, size_t* n_entries);
int kg_result_hits(kg_result* result, const kg_hit** hits, size_t* n); /* needs params.emit_hits */
void kg_result_free(kg_result* result);

#ifdef __cplusplus
}
#endif
#endif /* KMERGUTS_H */
