/*
 * include/kmerguts_host.h -- host-side mirror of the pieces of KmerGutsJava.run() that stay on the CPU
 * (lib/src/kmergutsjava/KmerGutsJava.java, "KGJ"): FASTA reading, function.index, the text report and main()'s flags.
 *
 * In a JVM deployment these stay in Java (the reference's own code keeps doing them) and only include/kmerguts.h is
 * bound.  No JVM exists in this image, so the same logic is provided here in C++ behind a C ABI: it is what the
 * kmer_guts_b200 command line and the parity tests drive.  None of it computes lookups or calls: that is the GPU's job.
 */
#ifndef KMERGUTS_HOST_H
#define KMERGUTS_HOST_H

#include "kmerguts.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kg_fasta kg_fasta;         /* parsed FASTA: ids + concatenated sequences */
typedef struct kg_functions kg_functions; /* function.index names */

/* readFasta (KGJ:1132-1192): id = first token of the caption, sequence = raw concatenation of the following lines
 * (not trimmed, KGJ:1176); ".gz" by suffix (KGJ:764).  Errors mirror the IllegalStateExceptions of KGJ:1158, 1170. */
int kg_fasta_read(const char* path, kg_fasta** out);
size_t kg_fasta_count(const kg_fasta* f);
const char* kg_fasta_id(const kg_fasta* f, size_t i);
const uint8_t* kg_fasta_bytes(const kg_fasta* f);
const uint64_t* kg_fasta_offsets(const kg_fasta* f); /* count+1 entries */
void kg_fasta_free(kg_fasta* f);

/* The same reader over a file consumed in batches of WHOLE records of about batch_bytes of text each (cut at caption
 * lines; a record longer than the batch comes alone): bounded memory for metagenome-sized inputs.  The concatenation of
 * the batches equals kg_fasta_read of the file.  kg_fasta_stream_next sets *batch = NULL at the end of the file; the
 * caller frees every batch with kg_fasta_free. */
typedef struct kg_fasta_stream kg_fasta_stream;
int kg_fasta_stream_open(const char* path, size_t batch_bytes, kg_fasta_stream** stream);
int kg_fasta_stream_next(kg_fasta_stream* stream, kg_fasta** batch);
void kg_fasta_stream_close(kg_fasta_stream* stream);

/* loadIndexedArray (KGJ:345-373): "index<TAB>name" lines, dense and in order; function.index.gz wins (KGJ:754-758). */
int kg_functions_load(const char* data_dir, kg_functions** out);
int kg_functions_read(const char* path, kg_functions** out);
size_t kg_functions_count(const kg_functions* f);
const char* kg_functions_name(const kg_functions* f, size_t i);
void kg_functions_free(kg_functions* f);

/* String.format("%f" / "%1.3f", float) as Java prints it (shortest decimal of the widened double, HALF_UP). */
int kg_format_java_f(float v, int precision, char* out, size_t outlen);

/* The per-sequence printing loop of run() (KGJ:810-818 with processAASeq 526-536, processSeq 538-558,
 * tabulateOtuDataForContig 516-524), including the collapse of repeated ids (LinkedHashMap, KGJ:772, 805-809).
 * debug != 0 adds the HIT lines (KGJ:472-475; the result must have been run with emit_hits) and the
 * "Kmer-table info" line (KGJ:951-954).  path == NULL writes to stdout. */
int kg_report_write(const char* path, int mode, int debug, const kg_fasta* fasta, const kg_functions* functions,
                    const kg_table* table, kg_result* result);

/* The same report written batch by batch (the streaming command line): kg_report_add appends the lines of one batch.
 * flags: bit 0 = debug as above, bit 1 = (extension, 6-frame runs) a "DNA-RANGE<TAB>begin<TAB>end<TAB>strand" line after
 * every CALL.  A CALL whose function index is outside function.index is KG_EFORMAT (the reference throws, KGJ:403). */
typedef struct kg_report kg_report;
int kg_report_open(const char* path, kg_report** report);
int kg_report_add(kg_report* report, int mode, int flags, const kg_fasta* fasta, const kg_functions* functions,
                  const kg_table* table, kg_result* result);
int kg_report_close(kg_report* report);

/* Extension: where a CALL of a 6-frame run lies on the contig.  The reference reports protein coordinates of the frame's
 * translation only (KGJ:398-401; the minus frames are frames of the full reverse complement, KGJ:1068-1071).  begin / end
 * are 0-based, inclusive nucleotide positions on the contig as given; strand is '+' or '-'. */
int kg_call_dna_range(const kg_call* call, uint64_t contig_len, uint64_t* begin, uint64_t* end, char* strand);

/* KmerGutsJava.main (KGJ:560-654): same flags and defaults.  Returns the process exit status. */
int kg_main(int argc, char** argv);

#ifdef __cplusplus
}
#endif
#endif
