/*
 * tools/benchlib/kmerguts_bench.h -- bench / test tooling, built as tools/benchlib/libkmerguts_bench.so.  NOT part of the
 * product library (libkmerguts_b200.so carries only the path) and not part of the drop-in surface (include/kmerguts.h):
 * the reference has no counterpart for any of this.
 *
 *   - the random-sector probe roofline microbenchmark (R_probe)
 *   - CUDA generators of the synthetic universe defined in tools/kg_synth.py (same counter-based hashing, so the
 *     bytes are identical to the numpy generator; tests/test_gpu_synth.py checks that)
 *   - a device-side writer of the REFERENCE's table format, so that the CPU oracle can be handed the very same
 *     200M-signature table the GPU probes
 */
#ifndef KMERGUTS_BENCH_H
#define KMERGUTS_BENCH_H

#include "../../include/kmerguts.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Parameters of tools/kg_synth.py::Universe.  cdf16 = 20 cumulative thresholds on a 16-bit draw, lenq = 4096 length
 * quantiles; both are computed on the host (no transcendental has to agree between CPU and GPU). */
typedef struct kg_universe {
    uint64_t n_families;
    uint64_t seed;
    uint32_t sig_keep_per_1024;
    uint32_t n_functions;
    uint32_t n_otus;
    uint32_t cdf16[20];
    uint32_t lenq[4096];
} kg_universe;

/* Signature table of the universe: every selected consensus window, first occurrence in family-major order wins,
 * truncated to max_sigs (0 = no limit).  On return *d_keys / *d_payload16 are device arrays of *n entries owned by
 * the caller (kg_device_free). */
int kg_synth_signatures(kg_context* ctx, const kg_universe* u, uint64_t max_sigs, uint64_t** d_keys, void** d_payload16,
                        uint64_t* n);
/* One shard of the same table (kmerguts_shard.h): only the signatures kg_shard_owner gives to `rank`; the union over the
 * ranks is exactly what kg_synth_signatures(max_sigs = 0) generates.  max_sigs must be 0 when nranks > 1. */
int kg_synth_signatures_sharded(kg_context* ctx, const kg_universe* u, uint64_t max_sigs, int rank, int nranks,
                                uint64_t** d_keys, void** d_payload16, uint64_t* n);

/* Query proteins first..first+n-1 of tools/kg_synth.py::Universe.proteins(seed): device byte stream + offsets. */
int kg_synth_proteins(kg_context* ctx, const kg_universe* u, uint64_t first, uint64_t n, uint64_t seed,
                      uint8_t** d_seq, uint64_t** d_off, uint64_t* total_bytes);

/* configs[2]-style genomes: n_genomes contigs of `length` bases; genes are family consensus proteins back-translated with
 * uniformly chosen synonymous codons on a random strand, separated by random spacers of uniform ACGT; ~1e-5 N. */
int kg_synth_genomes(kg_context* ctx, const kg_universe* u, uint64_t n_genomes, uint64_t length, uint64_t seed,
                     uint8_t** d_seq, uint64_t** d_off, uint64_t* total_bytes);
/* genomes first .. first+n_genomes-1 of the same job, byte for byte (a rank's share when the genomes are dealt to several GPUs) */
int kg_synth_genomes_range(kg_context* ctx, const kg_universe* u, uint64_t first, uint64_t n_genomes, uint64_t length, uint64_t seed,
                           uint8_t** d_seq, uint64_t** d_off, uint64_t* total_bytes);

/* kmer.table.mem_map image (24-byte header + num_slots 24-byte LE entries, linear probing WITHOUT wrap-around, last
 * slot empty) built on the device from (keys, payload).  num_slots = the first prime >= min_slots for which no probe
 * chain runs off the end.  *d_image is a device buffer of 24 + 24 * *num_slots bytes owned by the caller.
 * *mean_displacement (optional) = average distance of a key from its home slot key % numSigs: the extra 24-byte slots
 * the reference reads per hit, a property of the reference's own hash on skewed residue composition. */
int kg_synth_reference_image(kg_context* ctx, const uint64_t* d_keys, const void* d_payload16, uint64_t n,
                             uint64_t min_slots, uint64_t* num_slots, void** d_image, double* mean_displacement);

/* Size-independent cross-check of the probe path in protein mode: the most naive kernel possible (one thread per stream
 * position, residues read byte by byte, unfiltered full table lookup).  out3 = {valid windows, hits, order-independent
 * checksum over (stream position, payload) of all hits}.  d_seq must be the UNPATCHED protein bytes or the patched ones:
 * the kernel applies the reference's aa-mode window rule (i < len - 8) from the offsets itself. */
int kg_synth_naive_scan_aa(kg_context* ctx, const kg_table* table, const uint8_t* d_seq, const uint64_t* d_off, uint64_t n,
                           uint64_t total, uint64_t* out3);
/* The same checksum over kg_hit records (host array) of a run on those sequences. */
int kg_synth_hits_checksum(kg_context* ctx, const kg_hit* host_hits, uint64_t nhits, const uint64_t* d_off, uint64_t* out);

void kg_device_free(void* d_ptr);
int kg_device_to_host(kg_context* ctx, void* host, const void* dev, uint64_t bytes);

/* R_probe: independent uniformly random 32-byte sector loads over a buffer of `bytes` bytes (one 256-bit load each,
 * `loads_in_flight` per thread before the first use).  Returns sectors per second measured with CUDA events. */
int kg_probe_roofline(kg_context* ctx, uint64_t bytes, uint64_t n_loads, int threads_per_block, int loads_in_flight,
                      double* sectors_per_second);
/* Same access pattern over the key array of a loaded table. */
int kg_probe_roofline_table(kg_context* ctx, const kg_table* table, uint64_t n_loads, int threads_per_block,
                            int loads_in_flight, double* sectors_per_second);

#ifdef __cplusplus
}
#endif
#endif
