// kg_internal.h -- the opaque handle types behind include/kmerguts.h.
#pragma once

#include "kg_common.cuh"

struct kg_table {
    kg_context* ctx = nullptr;
    uint4* d_lines = nullptr;   // 128-byte bucket lines (kg_common.cuh)
    uint32_t num_buckets = 0;
    unsigned long long* d_filter = nullptr;
    uint32_t filter_words = 0;
    uint32_t filter2_words = 0;          // second-stage prefilter: stored right behind the first in d_filter
    uint32_t filter_halves = 0;          // d_filter = [half 0 | half 1], each filter_words long (= filter2_words): two-pass probe
    uint64_t src_size = 0;               // the reference-format file this table was parsed from (0 = not from a file):
    int64_t src_mtime_ns = 0;            //   identity recorded in the cache file (kg_table_save)
    kg_table_info info = {};
    int shard_rank = 0, shard_count = 1; // hash-sharded table: this handle holds the keys kg_owner_of() gives shard_rank
    mutable size_t l2_carve = 0;         // persisting-L2 set-aside the prefilter asked for (pin_filter)
    KgTableView view() const;
};

// Sequences resident in HBM.
//   aa mode : the residue stream IS the uploaded byte stream (the last residue of every protein is overwritten with 0,
//             which both separates proteins and reproduces the reference's dropped last window, KGJ:912/1055).
//   dna mode: the residue stream is produced by the translation kernel: six virtual proteins per contig
//             (+0,+1,+2,-0,-1,-2), each followed by at least one 0 byte.
struct kg_batch {
    kg_context* ctx = nullptr;
    int mode = KG_MODE_AA;
    uint64_t n = 0;           // sequences
    uint64_t total = 0;       // input bytes
    uint8_t* d_seq = nullptr; // input bytes (+ 64 bytes of zero padding)
    uint64_t* d_off = nullptr; // n+1
    bool owns_input = true;
    DevBuf seq_buf, off_buf;  // pooled storage behind d_seq / d_off when owns_input
    DevBuf pk_buf;            // packed input (kg_run_packed_aa: 5-bit residues; kg_run_packed_dna: 2-bit nucleotides) as uploaded, unpacked into seq_buf
    DevBuf aux_buf, exc_buf;  // kg_run_packed_dna: the slice's packed byte offsets, positions of the non-ACGT characters
    bool padded = false;      // ... in which every sequence is padded with zero bytes to a multiple of 8 positions
    // derived by prepare(): virtual sequences
    uint64_t nv = 0;          // n (aa) or 6n (dna)
    uint64_t vtotal = 0;      // residue-stream length
    DevBuf vseq;              // dna: translated stream
    DevBuf voff;              // dna: nv+1 uint64 offsets into the residue stream
    const uint8_t* stream() const { return mode == KG_MODE_AA ? d_seq : vseq.as<uint8_t>(); }
    const uint64_t* voffsets() const { return mode == KG_MODE_AA ? d_off : voff.as<uint64_t>(); }
    bool prepared = false;
    bool vtotal_known = false; // dna: vtotal already computed on the host (kg_run / kg_batch_upload)
};

struct kg_result {
    kg_context* ctx = nullptr;
    kg_params params = {};
    kg_run_stats stats = {};
    int mode = KG_MODE_AA;
    uint64_t n = 0, nv = 0;
    // device
    DevBuf d_calls;   // dense kg_call[num_calls]
    DevBuf d_otus;    // kg_otu[n]
    DevBuf d_hits;    // kg_hit[num_hits] when params.emit_hits
    // host
    HostBuf h_calls, h_otus, h_hits; // pinned, from the context's pool
    // compact OTU counts (kg_run / kg_run_packed_aa bring these home instead of 44 bytes per sequence)
    DevBuf d_otu_n, d_otu_entries;   // uint8 per sequence, kg_otu_entry per used buffer entry
    HostBuf h_otu_n, h_otu_entries;
    uint64_t num_otu_entries = 0;
    bool want_compact_otus = false;  // pipe_enqueue: also build the compact form
    bool otus_compact = false;       // the host holds the compact form (h_otus is then filled on demand)
    bool fetched = false;
};

// kg_table.cu: persisting-L2 set-aside + stream window for the first prefilter (the fused probe / k_answer want it even for
// tables whose default probe, the cascade, does not)
void kg_table_pin_filter(kg_context* ctx, const kg_table* t);
// kg_run.cu
int kg_batch_prepare(kg_batch* b, cudaStream_t st, uint32_t* launches);
