// kg_common.cuh -- shared declarations of libkmerguts_b200 (sm_100a only; no CPU fallback anywhere).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "../../include/kmerguts.h"

// ---------------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------------
void kg_set_error(const char* fmt, ...);
#define KG_FAIL(code, ...)            \
    do {                              \
        kg_set_error(__VA_ARGS__);    \
        return (code);                \
    } while (0)
#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            kg_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);   \
            return e_ == cudaErrorMemoryAllocation ? KG_ENOMEM : KG_ECUDA;                              \
        }                                                                                               \
    } while (0)
#define KG_TRY(call)              \
    do {                          \
        int r_ = (call);          \
        if (r_ != KG_OK) return r_; \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------
// GPU table layout.
//
// The reference streams 24-byte {int64 key; int32 oI; int32 avgFromEnd; int32 fI; float wt} slots of an
// open-addressing table with linear probing and no wrap-around (KGJ:944-1034, 995-999).  All a probe has to answer
// is "is k-mer v stored, and with which payload"; the layout is ours.  A key is < 20^8 < 2^35.
//
// One bucket = one 128-byte L2 line (B200 fetches the whole line from HBM on a miss, whatever the load asks for):
//
//   sector 0 (32 bytes) = 8 x uint32:  w[0..5] = low 32 bits of keys 0..5
//                                      w[6]    = bits [3i, 3i+3): bits 32..34 of key i (i = 0..5)
//                                                bit 31: overflow flag -- some key whose probe sequence passes through
//                                                        this bucket lives in a later one
//                                      w[7]    = unused (0)
//   sectors 1..3 (96 bytes)         =  six 16-byte payloads {oI, avgFromEnd, fI, float bits of wt}, one per key
//   empty slot = 35 one-bits (0x7FFFFFFFF > 20^8).
//
// A lookup reads ONE sector (the key sector) unless the flag is set; on a hit the payload is 16 more bytes of the line
// the probe has just pulled into L2, so a hit costs no second DRAM access.  slot = bucket*6 + lane.
// ---------------------------------------------------------------------------------------------------------------
constexpr int KG_BUCKET_KEYS = 6;
constexpr int KG_LINE_UINT4 = 8;              // 128-byte line = 8 x uint4: [0..1] keys, [2..7] payloads
constexpr uint32_t KG_W6_EMPTY = 0x0003FFFFu; // all six 3-bit high fields = 7, flag clear
constexpr uint32_t KG_W6_FLAG = 0x80000000u;
constexpr uint32_t KG_TAIL_BUCKETS = 4096;    // spill room past the last home bucket (no wrap-around in our layout either)

struct KgTableView {
    const uint4* lines;     // KG_LINE_UINT4 x uint4 per bucket
    uint32_t num_buckets;   // home buckets (hash range); KG_TAIL_BUCKETS more follow
    // L2-resident prefilter (kg_device.cuh): one 64-bit word per probe, two bits per key.  0 words = no filter.
    const unsigned long long* filter;
    uint32_t filter_words;
    // second prefilter of the cascade (kg_run.cu: k_refilter), an independent hash of the same keys.  It takes its turn in
    // L2 AFTER the first one: the two never have to be resident together.  0 words = no second stage.
    const unsigned long long* filter2;
    uint32_t filter2_words;
    // "halves" (kg_run.cu, k_probe_half): filter holds the keys whose hash has its top bit clear, filter2 (same size) the
    // others; the probe runs as two passes over the batch, one per half, so that each half has the L2 set-aside to itself and
    // every key gets twice the filter bits
    uint32_t halves;
};

// On B200 an L2 miss always brings in the whole 128-byte line (ncu: ~124 B of DRAM reads per random 32-byte sector
// load, whatever the load flavour or cudaLimitMaxL2FetchGranularity), so the DRAM roofline of a miss-dominant probe
// stream is (HBM bandwidth / 128 B) lookups/s.  About 85-90 % of all lookups are misses.  A blocked Bloom filter small
// enough to stay in the 126 MB L2 answers most of them without touching DRAM: random sector reads that hit L2 run at
// 2.9e11/s against 5.2e10/s from DRAM (tools/probe_bench).
// The persisting-L2 carve-out tops out at 79 MiB on B200, but a filter that fills it leaves too little ordinary L2 for
// everything else (and collides with itself): 64 MiB measured better than 76 on both the 2e8- and the 2.6e8-key tables.
constexpr uint64_t KG_FILTER_MAX_BYTES = 64ull << 20;
constexpr double KG_FILTER_BITS_PER_KEY = 3.0;

__host__ __device__ __forceinline__ uint64_t kg_mix(uint64_t k) {
    // murmur3 finaliser; the k-mer code is a base-20 number with very regular low digits
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}
__host__ __device__ __forceinline__ uint32_t kg_bucket_of_hash(uint64_t h, uint32_t num_buckets) {
    return (uint32_t)(((h >> 32) * (uint64_t)num_buckets) >> 32);
}
__host__ __device__ __forceinline__ uint32_t kg_home_bucket(uint64_t key, uint32_t num_buckets) {
    return kg_bucket_of_hash(kg_mix(key), num_buckets);
}
// Prefilter hashes.  The filter test runs once per LOOKUP (the bucket hash only once per survivor), and ncu showed the
// encode + filter stage bound by instruction issue with murmur's two 64-bit multiplies per window (~125 instructions per
// position), so the filters use ONE multiplicative hash each: m = key * odd constant (mod 2^64); the word comes from the
// top 32 bits, the two bit positions from a second 32-bit multiply of bits 11..42.  On the synthetic universes this has
// the same false-positive rate as the murmur mix (0.280 vs 0.279 at 2.68 bits/key; both stages together 0.078).
__host__ __device__ __forceinline__ uint64_t kg_fhash1(uint64_t key) { return key * 0x9E3779B97F4A7C15ull; }
__host__ __device__ __forceinline__ uint64_t kg_fhash2(uint64_t key) { return key * 0xC2B2AE3D27D4EB4Full; }
__host__ __device__ __forceinline__ uint32_t kg_filter_word(uint64_t m, uint32_t filter_words) {
    return (uint32_t)(((m >> 32) * (uint64_t)filter_words) >> 32);
}
__host__ __device__ __forceinline__ unsigned long long kg_filter_mask(uint64_t m) {
    const uint32_t b = (uint32_t)(m >> 11) * 0x85EBCA77u;
    return (1ull << (b >> 26)) | (1ull << ((b >> 20) & 63u));
}
// halves: which half a key belongs to, and its word inside that half (the bits below the top one pick the word)
__host__ __device__ __forceinline__ uint32_t kg_filter_half(uint64_t m) { return (uint32_t)(m >> 63); }
__host__ __device__ __forceinline__ uint32_t kg_filter_half_word(uint64_t m, uint32_t words) {
    return (uint32_t)((((m << 1) >> 32) * (uint64_t)words) >> 32);
}
// second-stage prefilter (the probe cascade, kg_run.cu): same construction on an independent multiplier, so that a false
// positive of the first filter is an (almost) independent draw in the second
__host__ __device__ __forceinline__ uint32_t kg_filter2_word(uint64_t m2, uint32_t filter2_words) {
    return (uint32_t)(((m2 >> 32) * (uint64_t)filter2_words) >> 32);
}
__host__ __device__ __forceinline__ unsigned long long kg_filter2_mask(uint64_t m2) {
    const uint32_t b = (uint32_t)(m2 >> 11) * 0x27D4EB2Fu;
    return (1ull << (b >> 26)) | (1ull << ((b >> 20) & 63u));
}

// hash-sharded table (configs[4]): the rank that owns a key.  A second, independent mix: the bucket and the filter
// already use both halves of kg_mix(key), and an owner derived from the same bits would leave every shard using 1/R of
// its bucket range.
constexpr int KG_MAX_RANKS = 16;
__host__ __device__ __forceinline__ uint32_t kg_owner_of(uint64_t key, uint32_t nranks) {
    const uint64_t h = kg_mix(key ^ 0x5851F42D4C957F2Dull);
    return (uint32_t)(((h >> 32) * (uint64_t)nranks) >> 32);
}

// ---------------------------------------------------------------------------------------------------------------
// device memory helpers (host side)
// ---------------------------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes); // grow-only; contents are NOT preserved across a growth
    void release();
    template <class T>
    T* as() const { return (T*)p; }
};

struct HostBuf { // pinned host memory
    void* p = nullptr;
    size_t cap = 0;
};

struct kg_context {
    int device = 0;
    cudaStream_t stream = nullptr;      // compute
    cudaStream_t copy_stream = nullptr; // H2D of the next slice in the pipelined end-to-end call (kg_run)
    cudaStream_t d2h_stream = nullptr;  // D2H of the previous slice's records
    cudaStream_t fsm_stream = nullptr;  // run FSM + call compaction of slice s while the compute stream probes slice s+1
    cudaStream_t lines_stream = nullptr; // probe cascade run in parts: bucket-line stage of part p while the filters work on part p+1
    cudaEvent_t ev[12] = {};           // 0-5: run / fetch / upload brackets, 6-9: pipeline stages
    cudaEvent_t up_ev[4] = {};         // kg_run: slice s%4 has been uploaded
    cudaEvent_t d2h_ev[3] = {};        // kg_run: records of slice s%3 have reached the host
    int sm_count = 0;
    size_t l2_bytes = 0;
    DevBuf scan_tmp, scan_tmp2;         // CUB temp storage (compute stream / fsm stream)
    // pinned staging for small device->host counters
    uint64_t* h_counters = nullptr;
    void* scratch = nullptr;            // RunScratch (kg_run.cu)
    // result buffers are recycled through these pools: cudaMalloc / cudaMallocHost cost more than a whole run
    std::vector<DevBuf> dev_pool;
    std::vector<HostBuf> host_pool;
};

// counters written by the pipeline, one block of 8 x uint64 per run
enum { KG_CTR_HITS = 0, KG_CTR_KMERS = 1, KG_CTR_CALLS = 2, KG_CTR_OVERFLOW = 3, KG_CTR_VPOS = 4, KG_CTR_CLAIM = 5 /* probe cascade: hit-chunk slots claimed = survivors of the second filter */, KG_CTR_SURV1 = 6 /* survivors of the first filter */, KG_CTR_COUNT = 8 }; // h_counters has KG_CTR_COUNT + 1 slots: the last receives the call total
