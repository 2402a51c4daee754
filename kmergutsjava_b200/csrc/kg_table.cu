// kg_table.cu -- loader for the reference's kmer.table.mem_map[.gz] and the device-side builder of the
// one-sector-per-probe bucket table (layout: kg_common.cuh).
//
// Replaces: readKmerTableHeader (KGJ:924-942), the 24-byte little-endian entry decode of lookup (KGJ:995-999,
// 1097-1130) and -- semantically -- the table side of the sort-merge join (KGJ:959-1026).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>

#include "kg_device.cuh"
#include "kg_internal.h"

// ---------------------------------------------------------------------------------------------------------------
// errors / small host helpers (shared by all translation units)
// ---------------------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void kg_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
extern "C" const char* kg_last_error(void) { return g_err; }
extern "C" const char* kg_version(void) { return "kmerguts_b200 0.1 (sm_100a)"; }

int DevBuf::ensure(size_t bytes) {
    if (bytes <= cap) return KG_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMalloc(&p, bytes);
        want = bytes;
    }
    if (e != cudaSuccess) {
        p = nullptr;
        KG_FAIL(KG_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    }
    cap = want;
    return KG_OK;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

static double filter_bits_per_key();
static int filter_stages();
static void pin_filter(kg_context* ctx, const kg_table* t, bool force);

// ---------------------------------------------------------------------------------------------------------------
// reference-format image parser (streaming, so a 10-100 GB file never has to sit in host memory twice)
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct ImageParser {
    int64_t num_slots = 0, entry_size = 0, version = 0;
    bool header_done = false;
    uint8_t carry[24];
    size_t ncarry = 0;
    int64_t slot = 0, run_start = 0;
    bool in_run = false;
    int64_t unreachable = 0, unmatchable = 0;
    int shard_rank = 0, shard_count = 1; // hash-sharded table: keep the keys this rank owns
    int64_t not_owned = 0;
    std::vector<uint64_t> keys;
    std::vector<int4> payload;
    std::string error;

    bool header(const uint8_t* h) {
        memcpy(&num_slots, h, 8);      // readLongLE x3, KGJ:933-935 (x86 is little-endian)
        memcpy(&entry_size, h + 8, 8);
        memcpy(&version, h + 16, 8);
        header_done = true;
        if (entry_size != 24) { // KGJ:992 skips by entrySize but KGJ:995-999 always reads 24 bytes
            error = "kmer table: entrySize " + std::to_string(entry_size) + " != 24 is not readable by the reference either";
            return false;
        }
        if (num_slots <= 0) {
            error = "kmer table: numSigs " + std::to_string(num_slots) + " <= 0";
            return false;
        }
        size_t guess = (size_t)std::min<int64_t>(num_slots / 2 + 16, (int64_t)1 << 33);
        keys.reserve(guess);
        payload.reserve(guess);
        return true;
    }
    inline void entry(const uint8_t* e) {
        int64_t k;
        memcpy(&k, e, 8);
        if (k > KG_MAX_ENCODED) { // empty slot, KGJ:1000: every pending probe chain ends here
            in_run = false;
        } else {
            if (!in_run) {
                in_run = true;
                run_start = slot;
            }
            if (k >= 0 && k < KG_MAX_ENCODED) {
                // The reference probes slots h, h+1, ... (h = key % numSigs) until an empty slot, with NO wrap-around
                // (KGJ:959-1026).  It can therefore return this slot iff h lies inside the occupied run that ends here.
                int64_t h = k % num_slots;
                if (h >= run_start && h <= slot) {
                    if (shard_count > 1 && (int)kg_owner_of((uint64_t)k, (uint32_t)shard_count) != shard_rank) {
                        not_owned++;
                        slot++;
                        return;
                    }
                    int4 p;
                    memcpy(&p, e + 8, 16); // otuIndex, avgFromEnd, functionIndex, functionWt bits (KGJ:996-999)
                    keys.push_back((uint64_t)k);
                    payload.push_back(p);
                } else {
                    unreachable++;
                }
            } else {
                unmatchable++; // occupies a slot (extends chains) but no valid 8-mer encodes to it
            }
        }
        slot++;
    }
    bool feed(const uint8_t* p, size_t n) {
        if (ncarry) {
            size_t need = 24 - ncarry, take = std::min(need, n);
            memcpy(carry + ncarry, p, take);
            ncarry += take;
            p += take;
            n -= take;
            if (ncarry < 24) return true;
            ncarry = 0;
            if (!header_done) {
                if (!header(carry)) return false;
            } else {
                entry(carry);
            }
        }
        if (!header_done && n >= 24) {
            if (!header(p)) return false;
            p += 24;
            n -= 24;
        }
        if (header_done) {
            size_t whole = n / 24;
            for (size_t i = 0; i < whole; i++) entry(p + 24 * i);
            p += whole * 24;
            n -= whole * 24;
        }
        if (n) {
            memcpy(carry, p, n);
            ncarry = n;
        }
        return true;
    }
    int64_t tail_run() const { return in_run ? slot - run_start : 0; }
};

// ---------------------------------------------------------------------------------------------------------------
// device builder
// ---------------------------------------------------------------------------------------------------------------
struct MaxI64 {
    __host__ __device__ __forceinline__ long long operator()(long long a, long long b) const { return a > b ? a : b; }
};

__global__ void k_make_composite(const uint64_t* __restrict__ keys, size_t n, uint32_t nb, uint64_t* __restrict__ comp,
                                 uint32_t* __restrict__ idx) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k = keys[i];
    comp[i] = ((uint64_t)kg_home_bucket(k, nb) << 35) | k;
    idx[i] = (uint32_t)i;
}

__global__ void k_flag_first(const uint64_t* __restrict__ comp, size_t n, uint8_t* __restrict__ flag) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flag[i] = (i == 0) || (comp[i] != comp[i - 1]); // equal keys are adjacent; the stable sort keeps the lowest slot first
}

__global__ void k_slot_bias(const uint64_t* __restrict__ comp, size_t n, long long* __restrict__ t) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    t[i] = (long long)(comp[i] >> 35) * KG_BUCKET_KEYS - (long long)i;
}

__global__ void k_init_buckets(uint4* __restrict__ lines, size_t nbuckets_total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; // one thread per uint4 of the table
    if (i >= nbuckets_total * KG_LINE_UINT4) return;
    const uint32_t q = (uint32_t)(i % KG_LINE_UINT4);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);                                       // payload sectors
    if (q == 0) v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu); // keys 0..3
    if (q == 1) v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, KG_W6_EMPTY, 0u);          // keys 4..5, high bits + flag, unused
    lines[i] = v;
}

// Keys sorted by home bucket take the first free slot at or after their bucket's first slot:
//   slot_r = max(slot_{r-1} + 1, 6*home_r)  <=>  slot_r = r + max_{q<=r}(6*home_q - q)     (an inclusive max-scan)
// Bucket b gets the overflow flag iff the key in the first slot of bucket b+1 has its home at or before b.
__global__ void k_scatter(const uint64_t* __restrict__ comp, const uint32_t* __restrict__ idx,
                          const long long* __restrict__ tmax, size_t n, const int4* __restrict__ payload_in,
                          uint4* __restrict__ lines, uint64_t total_slots, unsigned long long* __restrict__ err) {
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    uint64_t c = comp[r];
    uint64_t key = c & 0x7FFFFFFFFull;
    uint64_t home = c >> 35;
    uint64_t slot = (uint64_t)((long long)r + tmax[r]);
    if (slot >= total_slots) {
        atomicAdd(err, 1ull);
        return;
    }
    uint64_t b = slot / KG_BUCKET_KEYS;
    uint32_t lane = (uint32_t)(slot - b * KG_BUCKET_KEYS);
    uint32_t* words = reinterpret_cast<uint32_t*>(lines + b * KG_LINE_UINT4); // the key sector
    words[lane] = (uint32_t)key;
    uint32_t hi = (uint32_t)(key >> 32);
    atomicAnd(&words[6], ~(7u << (3 * lane)) | (hi << (3 * lane)));
    if (lane == 0 && home < b) atomicOr(reinterpret_cast<uint32_t*>(lines + (b - 1) * KG_LINE_UINT4) + 6, KG_W6_FLAG);
    reinterpret_cast<int4*>(lines + b * KG_LINE_UINT4 + 2)[lane] = payload_in[idx[r]];
}

__global__ void k_filter_build(const uint64_t* __restrict__ comp, size_t n, unsigned long long* __restrict__ filter,
                               uint32_t filter_words) {
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint64_t m = kg_fhash1(comp[r] & 0x7FFFFFFFFull);
    atomicOr(&filter[kg_filter_word(m, filter_words)], kg_filter_mask(m));
}
__global__ void k_filter2_build(const uint64_t* __restrict__ comp, size_t n, unsigned long long* __restrict__ filter2,
                                uint32_t filter2_words) {
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint64_t m = kg_fhash2(comp[r] & 0x7FFFFFFFFull);
    atomicOr(&filter2[kg_filter2_word(m, filter2_words)], kg_filter2_mask(m));
}

__global__ void k_count_flagged(const uint32_t* __restrict__ words, size_t nbuckets_total, unsigned long long* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool f = i < nbuckets_total && (words[i * (KG_LINE_UINT4 * 4) + 6] & KG_W6_FLAG);
    unsigned m = __ballot_sync(0xFFFFFFFFu, f);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}

// every input key must be found, with the payload it came with
__global__ void k_verify(KgTableView t, const uint64_t* __restrict__ keys, const int4* __restrict__ payload, size_t n,
                         const uint8_t* __restrict__ keep_flag_sorted_unused, unsigned long long* bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s = kg_lookup(t, keys[i]);
    bool ok = s != 0xFFFFFFFFu;
    if (ok) {
        int4 a = kg_load_payload(t.lines, s), b = payload[i];
        ok = a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w;
    }
    if (!ok) atomicAdd(bad, 1ull);
}

inline unsigned blocks_for(size_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

} // namespace

void kg_table_pin_filter(kg_context* ctx, const kg_table* t) { pin_filter(ctx, t, true); }

KgTableView kg_table::view() const {
    KgTableView v;
    v.lines = d_lines;
    v.num_buckets = num_buckets;
    v.filter = d_filter;
    v.filter_words = filter_words;
    v.filter2 = filter2_words ? d_filter + filter_words : nullptr;
    v.filter2_words = filter2_words;
    return v;
}

// d_keys / d_payload: device arrays in the reference's SLOT order (so that of two equal keys the earlier slot wins).
// Pin the prefilter in L2: a persisting carve-out (<= 79 MiB on B200) plus an access-policy window on the context's
// stream, so that the 128-byte lines streaming through for the probes cannot push it out.  (One table per context
// benefits; a later table takes the window over.)
static void pin_filter(kg_context* ctx, const kg_table* t, bool force) {
    if (!t->d_filter || !t->filter_words || getenv("KG_NO_L2_PERSIST")) return;
    // A table with a second prefilter is probed by the cascade (kg_run.cu): each filter has the L2 to itself during its
    // stage, and a persisting set-aside only takes L2 away from the bucket-line stage (measured: 3.5 vs 2.8 ms).  The fused
    // kernel and the hash-sharded mode's k_answer (one filter against streaming lines) keep the window.
    if (!force && t->filter2_words && t->shard_count == 1 && !getenv("KG_CASCADE_PERSIST")) return;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, ctx->device) != cudaSuccess || prop.persistingL2CacheMaxSize <= 0) return;
    const size_t fbytes = (size_t)t->filter_words * 8;
    const size_t carve = std::min<size_t>(fbytes, (size_t)prop.persistingL2CacheMaxSize);
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
    t->l2_carve = carve;
    cudaStreamAttrValue av = {};
    av.accessPolicyWindow.base_ptr = t->d_filter;
    av.accessPolicyWindow.num_bytes = std::min<size_t>(fbytes, (size_t)prop.accessPolicyMaxWindowSize);
    // a filter larger than the carve-out: that fraction of its lines persists, the rest competes normally
    av.accessPolicyWindow.hitRatio = fbytes > carve ? (float)((double)carve / (double)fbytes) : 1.0f;
    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp = fbytes > carve ? cudaAccessPropertyNormal : cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av);
    cudaGetLastError();
}

static int build_on_device(kg_context* ctx, const uint64_t* d_keys, const int4* d_payload, size_t n, double load,
                           bool verify_payload, kg_table* t) {
    if (n >= (1ull << 32)) KG_FAIL(KG_ERANGE, "table: %zu signatures do not fit the 32-bit slot index", n);
    cudaStream_t st = ctx->stream;
    uint64_t nb64 = (uint64_t)((double)n / (KG_BUCKET_KEYS * load)) + 1;
    if (nb64 < 64) nb64 = 64;
    if (nb64 + KG_TAIL_BUCKETS >= (1ull << 29)) KG_FAIL(KG_ERANGE, "table: %llu buckets exceed the 29-bit bucket index", (unsigned long long)nb64);
    const uint32_t nb = (uint32_t)nb64;
    const size_t nb_total = (size_t)nb + KG_TAIL_BUCKETS;
    const uint64_t total_slots = (uint64_t)nb_total * KG_BUCKET_KEYS;
    if (total_slots >= (1ull << 32)) KG_FAIL(KG_ERANGE, "table: %llu slots exceed the 32-bit slot index", (unsigned long long)total_slots);

    CU(cudaMalloc(&t->d_lines, nb_total * 128));
    t->num_buckets = nb;
    k_init_buckets<<<blocks_for(nb_total * KG_LINE_UINT4, 256), 256, 0, st>>>(t->d_lines, nb_total);

    unsigned long long* d_ctr = nullptr; // [0] scatter overflow, [1] flagged buckets, [2] verify failures
    CU(cudaMalloc(&d_ctr, 4 * sizeof(unsigned long long)));
    CU(cudaMemsetAsync(d_ctr, 0, 4 * sizeof(unsigned long long), st));
    size_t n_unique = n;

    if (n > 0) {
        uint64_t *comp_a = nullptr, *comp_b = nullptr;
        uint32_t *idx_a = nullptr, *idx_b = nullptr;
        long long* tb = nullptr;
        uint8_t* flag = nullptr;
        size_t* d_nsel = nullptr;
        CU(cudaMalloc(&comp_a, n * 8));
        CU(cudaMalloc(&comp_b, n * 8));
        CU(cudaMalloc(&idx_a, n * 4));
        CU(cudaMalloc(&idx_b, n * 4));
        CU(cudaMalloc(&flag, n));
        CU(cudaMalloc(&d_nsel, sizeof(size_t)));
        k_make_composite<<<blocks_for(n, 256), 256, 0, st>>>(d_keys, n, nb, comp_a, idx_a);

        int end_bit = 35;
        while (end_bit < 64 && (nb_total >> (end_bit - 35)) != 0) end_bit++;
        cub::DoubleBuffer<uint64_t> dk(comp_a, comp_b);
        cub::DoubleBuffer<uint32_t> dv(idx_a, idx_b);
        size_t tmp_bytes = 0;
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, n, 0, end_bit, st));
        KG_TRY(ctx->scan_tmp.ensure(tmp_bytes));
        CU(cub::DeviceRadixSort::SortPairs(ctx->scan_tmp.p, tmp_bytes, dk, dv, n, 0, end_bit, st));
        uint64_t* comp = dk.Current();
        uint32_t* idx = dv.Current();
        uint64_t* comp_o = dk.Alternate();
        uint32_t* idx_o = dv.Alternate();

        // drop repeated keys (a malformed file may hold one key twice; the reference returns the first on the chain)
        k_flag_first<<<blocks_for(n, 256), 256, 0, st>>>(comp, n, flag);
        size_t sel_bytes = 0;
        CU(cub::DeviceSelect::Flagged(nullptr, sel_bytes, comp, flag, comp_o, d_nsel, n, st));
        KG_TRY(ctx->scan_tmp.ensure(sel_bytes));
        CU(cub::DeviceSelect::Flagged(ctx->scan_tmp.p, sel_bytes, comp, flag, comp_o, d_nsel, n, st));
        CU(cub::DeviceSelect::Flagged(nullptr, sel_bytes, idx, flag, idx_o, d_nsel, n, st));
        KG_TRY(ctx->scan_tmp.ensure(sel_bytes));
        CU(cub::DeviceSelect::Flagged(ctx->scan_tmp.p, sel_bytes, idx, flag, idx_o, d_nsel, n, st));
        CU(cudaMemcpyAsync(&n_unique, d_nsel, sizeof(size_t), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        comp = comp_o;
        idx = idx_o;

        CU(cudaMalloc(&tb, n_unique * 8));
        k_slot_bias<<<blocks_for(n_unique, 256), 256, 0, st>>>(comp, n_unique, tb);
        size_t scan_bytes = 0;
        CU(cub::DeviceScan::InclusiveScan(nullptr, scan_bytes, tb, tb, MaxI64(), n_unique, st));
        KG_TRY(ctx->scan_tmp.ensure(scan_bytes));
        CU(cub::DeviceScan::InclusiveScan(ctx->scan_tmp.p, scan_bytes, tb, tb, MaxI64(), n_unique, st));
        k_scatter<<<blocks_for(n_unique, 256), 256, 0, st>>>(comp, idx, tb, n_unique, d_payload, t->d_lines, total_slots, d_ctr);
        k_count_flagged<<<blocks_for(nb_total, 256), 256, 0, st>>>((const uint32_t*)t->d_lines, nb_total, d_ctr + 1);
        if (filter_bits_per_key() > 0) { // L2-resident prefilter over the same keys
            uint64_t bytes = (uint64_t)((double)n_unique * filter_bits_per_key() / 8.0);
            uint64_t max_bytes = KG_FILTER_MAX_BYTES;
            if (const char* e = getenv("KG_FILTER_MAX_MB")) max_bytes = (uint64_t)atoll(e) << 20; // experiments
            if (bytes > max_bytes) bytes = max_bytes;
            if (bytes < 4096) bytes = 4096;
            t->filter_words = (uint32_t)(bytes / 8) & ~15u; // whole 128-byte lines, so that the second filter starts on one
            t->filter2_words = filter_stages() >= 2 ? t->filter_words : 0;
            const size_t fwords = (size_t)t->filter_words + t->filter2_words;
            CU(cudaMalloc(&t->d_filter, fwords * 8));
            CU(cudaMemsetAsync(t->d_filter, 0, fwords * 8, st));
            k_filter_build<<<blocks_for(n_unique, 256), 256, 0, st>>>(comp, n_unique, t->d_filter, t->filter_words);
            if (t->filter2_words)
                k_filter2_build<<<blocks_for(n_unique, 256), 256, 0, st>>>(comp, n_unique, t->d_filter + t->filter_words, t->filter2_words);
            pin_filter(ctx, t, false);
        }
        // Every stored key must be found again.  With repeated keys only the surviving copy's payload can match, so
        // the payload comparison is skipped for inputs that had duplicates.
        (void)verify_payload;
        CU(cudaStreamSynchronize(st));
        if (n_unique == n) {
            k_verify<<<blocks_for(n, 256), 256, 0, st>>>(t->view(), d_keys, d_payload, n, nullptr, d_ctr + 2);
        }
        CU(cudaStreamSynchronize(st));
        CU(cudaGetLastError());
        cudaFree(comp_a);
        cudaFree(comp_b);
        cudaFree(idx_a);
        cudaFree(idx_b);
        cudaFree(flag);
        cudaFree(d_nsel);
        cudaFree(tb);
    }
    unsigned long long h_ctr[4] = {0, 0, 0, 0};
    CU(cudaMemcpyAsync(h_ctr, d_ctr, sizeof h_ctr, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    cudaFree(d_ctr);
    if (h_ctr[0]) KG_FAIL(KG_EFORMAT, "table build: %llu keys spilled past the tail buckets (load %.2f too high?)", h_ctr[0], load);
    if (h_ctr[2]) KG_FAIL(KG_ECUDA, "table build: %llu stored keys are not found again", h_ctr[2]);
    t->info.num_signatures = (int64_t)n_unique;
    t->info.num_buckets = (int64_t)nb_total;
    t->info.flagged_buckets = (int64_t)h_ctr[1];
    t->info.device_bytes = (int64_t)(nb_total * 128 + ((size_t)t->filter_words + t->filter2_words) * 8);
    return KG_OK;
}

static double filter_bits_per_key() { // KG_FILTER_BITS=0 disables the prefilter
    const char* e = getenv("KG_FILTER_BITS");
    return e ? atof(e) : KG_FILTER_BITS_PER_KEY;
}
static int filter_stages() { // KG_FILTER_STAGES=2 also builds the second-stage prefilter; the probe then runs as the three-kernel
    const char* e = getenv("KG_FILTER_STAGES"); // cascade (kg_run.cu), which moves 40 % less DRAM traffic but measured slower than the
    return e ? atoi(e) : 1;                     // fused kernel (profiles/r02_probe_cascade.md), so one stage is the default
}
static double table_load_factor() {
    const char* e = getenv("KG_TABLE_LOAD");
    double v = e ? atof(e) : 0.0;
    return (v > 0.05 && v < 0.95) ? v : 0.50; // 3 keys per 6-key bucket on average: ~3 % of buckets overflow
}

static int table_from_parser(kg_context* ctx, ImageParser& ps, kg_table** out) {
    if (!ps.header_done) KG_FAIL(KG_EIO, "kmer table: EOF inside the 24-byte header");
    CU(cudaSetDevice(ctx->device));
    kg_table* t = new kg_table();
    t->ctx = ctx;
    t->info.num_slots = ps.num_slots;
    t->info.entry_size = ps.entry_size;
    t->info.version = ps.version;
    t->info.num_unreachable = ps.unreachable + ps.unmatchable;
    t->info.tail_run = ps.tail_run();
    t->shard_rank = ps.shard_rank;
    t->shard_count = ps.shard_count;
    size_t n = ps.keys.size();
    uint64_t* d_keys = nullptr;
    int4* d_payload = nullptr;
    int rc = KG_OK;
    do {
        if (n) {
            if (cudaMalloc(&d_keys, n * 8) != cudaSuccess || cudaMalloc(&d_payload, n * sizeof(int4)) != cudaSuccess) {
                kg_set_error("table: device allocation for %zu entries failed", n);
                rc = KG_ENOMEM;
                break;
            }
            cudaMemcpyAsync(d_keys, ps.keys.data(), n * 8, cudaMemcpyHostToDevice, ctx->stream);
            cudaMemcpyAsync(d_payload, ps.payload.data(), n * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream);
        }
        rc = build_on_device(ctx, d_keys, d_payload, n, table_load_factor(), true, t);
    } while (0);
    if (d_keys) cudaFree(d_keys);
    if (d_payload) cudaFree(d_payload);
    if (rc != KG_OK) {
        kg_table_free(t);
        return rc;
    }
    *out = t;
    return KG_OK;
}

static int check_shard(int rank, int nranks) {
    if (nranks < 1 || nranks > KG_MAX_RANKS || rank < 0 || rank >= nranks) KG_FAIL(KG_EINVAL, "shard %d of %d: need 0 <= rank < nranks <= %d", rank, nranks, KG_MAX_RANKS);
    return KG_OK;
}

extern "C" int kg_shard_owner(uint64_t key, int nranks) {
    if (nranks < 1) return -1;
    return (int)kg_owner_of(key, (uint32_t)nranks);
}

extern "C" int kg_table_from_image_sharded(kg_context* ctx, const void* image, size_t nbytes, int rank, int nranks, kg_table** table) {
    if (!ctx || !image || !table) KG_FAIL(KG_EINVAL, "kg_table_from_image: null argument");
    KG_TRY(check_shard(rank, nranks));
    ImageParser ps;
    ps.shard_rank = rank;
    ps.shard_count = nranks;
    if (!ps.feed((const uint8_t*)image, nbytes)) KG_FAIL(KG_EFORMAT, "%s", ps.error.c_str());
    return table_from_parser(ctx, ps, table);
}

extern "C" int kg_table_from_image(kg_context* ctx, const void* image, size_t nbytes, kg_table** table) {
    if (!ctx || !image || !table) KG_FAIL(KG_EINVAL, "kg_table_from_image: null argument");
    ImageParser ps;
    if (!ps.feed((const uint8_t*)image, nbytes)) KG_FAIL(KG_EFORMAT, "%s", ps.error.c_str());
    return table_from_parser(ctx, ps, table);
}

static int load_file_impl(kg_context* ctx, const char* path, int rank, int nranks, kg_table** table) {
    if (!ctx || !path || !table) KG_FAIL(KG_EINVAL, "kg_table_load_file: null argument");
    ImageParser ps;
    ps.shard_rank = rank;
    ps.shard_count = nranks;
    std::vector<uint8_t> buf(8u << 20);
    size_t len = strlen(path);
    bool gz = len > 3 && strcmp(path + len - 3, ".gz") == 0; // the reference keys on the suffix (KGJ:927)
    if (gz) {
        gzFile g = gzopen(path, "rb");
        if (!g) KG_FAIL(KG_EIO, "cannot open %s", path);
        gzbuffer(g, 1 << 20);
        for (;;) {
            int got = gzread(g, buf.data(), (unsigned)buf.size());
            if (got < 0) {
                gzclose(g);
                KG_FAIL(KG_EIO, "gzip error reading %s", path);
            }
            if (got == 0) break;
            if (!ps.feed(buf.data(), (size_t)got)) {
                gzclose(g);
                KG_FAIL(KG_EFORMAT, "%s: %s", path, ps.error.c_str());
            }
        }
        gzclose(g);
    } else {
        FILE* f = fopen(path, "rb");
        if (!f) KG_FAIL(KG_EIO, "cannot open %s", path);
        for (;;) {
            size_t got = fread(buf.data(), 1, buf.size(), f);
            if (got == 0) break;
            if (!ps.feed(buf.data(), got)) {
                fclose(f);
                KG_FAIL(KG_EFORMAT, "%s: %s", path, ps.error.c_str());
            }
        }
        fclose(f);
    }
    KG_TRY(table_from_parser(ctx, ps, table));
    struct stat sst;
    if (stat(path, &sst) == 0) {
        (*table)->src_size = (uint64_t)sst.st_size;
        (*table)->src_mtime_ns = (int64_t)sst.st_mtim.tv_sec * 1000000000ll + sst.st_mtim.tv_nsec;
    }
    return KG_OK;
}

extern "C" int kg_table_load_file(kg_context* ctx, const char* path, kg_table** table) { return load_file_impl(ctx, path, 0, 1, table); }

extern "C" int kg_table_load_sharded(kg_context* ctx, const char* data_dir, int rank, int nranks, kg_table** table) {
    if (!ctx || !data_dir || !table) KG_FAIL(KG_EINVAL, "kg_table_load: null argument");
    KG_TRY(check_shard(rank, nranks));
    std::string base = std::string(data_dir) + "/kmer.table.mem_map";
    std::string gz = base + ".gz";
    struct stat st;
    if (stat(gz.c_str(), &st) == 0) return load_file_impl(ctx, gz.c_str(), rank, nranks, table); // KGJ:750-753
    return load_file_impl(ctx, base.c_str(), rank, nranks, table);
}

extern "C" int kg_table_load(kg_context* ctx, const char* data_dir, kg_table** table) { return kg_table_load_sharded(ctx, data_dir, 0, 1, table); }

extern "C" int kg_table_from_device_entries(kg_context* ctx, const uint64_t* d_keys, const void* d_payload16, size_t n,
                                            kg_table** table) {
    if (!ctx || !table || (n && (!d_keys || !d_payload16))) KG_FAIL(KG_EINVAL, "kg_table_from_device_entries: null argument");
    CU(cudaSetDevice(ctx->device));
    kg_table* t = new kg_table();
    t->ctx = ctx;
    t->info.entry_size = 24;
    t->info.version = 1;
    int rc = build_on_device(ctx, d_keys, (const int4*)d_payload16, n, table_load_factor(), true, t);
    if (rc != KG_OK) {
        kg_table_free(t);
        return rc;
    }
    *table = t;
    return KG_OK;
}

namespace {
__global__ void k_owner_flag(const uint64_t* __restrict__ keys, size_t n, uint32_t rank, uint32_t nranks, uint8_t* __restrict__ flag) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = kg_owner_of(keys[i], nranks) == rank;
}
} // namespace

extern "C" int kg_table_from_device_entries_sharded(kg_context* ctx, const uint64_t* d_keys, const void* d_payload16, size_t n,
                                                    int rank, int nranks, kg_table** table) {
    if (!ctx || !table || (n && (!d_keys || !d_payload16))) KG_FAIL(KG_EINVAL, "kg_table_from_device_entries_sharded: null argument");
    KG_TRY(check_shard(rank, nranks));
    if (nranks == 1) return kg_table_from_device_entries(ctx, d_keys, d_payload16, n, table);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevBuf flag, keys, payload, count, tmp;
    int rc = KG_OK;
    kg_table* t = nullptr;
    do {
        // entries this rank owns, compacted in their original order (the builder wants distinct keys, not an order)
        if ((rc = flag.ensure(std::max<size_t>(n, 1))) != KG_OK || (rc = keys.ensure(std::max<size_t>(n, 1) * 8)) != KG_OK ||
            (rc = payload.ensure(std::max<size_t>(n, 1) * sizeof(int4))) != KG_OK || (rc = count.ensure(16)) != KG_OK)
            break;
        size_t owned = 0;
        if (n) {
            k_owner_flag<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_keys, n, (uint32_t)rank, (uint32_t)nranks, flag.as<uint8_t>());
            size_t tb1 = 0, tb2 = 0;
            cub::DeviceSelect::Flagged(nullptr, tb1, d_keys, flag.as<uint8_t>(), keys.as<uint64_t>(), count.as<size_t>(), n, st);
            cub::DeviceSelect::Flagged(nullptr, tb2, (const int4*)d_payload16, flag.as<uint8_t>(), payload.as<int4>(), count.as<size_t>(), n, st);
            if ((rc = tmp.ensure(std::max(tb1, tb2))) != KG_OK) break;
            cub::DeviceSelect::Flagged(tmp.p, tb1, d_keys, flag.as<uint8_t>(), keys.as<uint64_t>(), count.as<size_t>(), n, st);
            cub::DeviceSelect::Flagged(tmp.p, tb2, (const int4*)d_payload16, flag.as<uint8_t>(), payload.as<int4>(), count.as<size_t>(), n, st);
            if (cudaMemcpyAsync(&owned, count.p, sizeof(size_t), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
                kg_set_error("kg_table_from_device_entries_sharded: %s", cudaGetErrorString(cudaGetLastError()));
                rc = KG_ECUDA;
                break;
            }
        }
        flag.release();
        tmp.release();
        rc = kg_table_from_device_entries(ctx, keys.as<uint64_t>(), payload.p, owned, &t);
    } while (0);
    flag.release();
    keys.release();
    payload.release();
    count.release();
    tmp.release();
    if (rc != KG_OK) return rc;
    t->shard_rank = rank;
    t->shard_count = nranks;
    *table = t;
    return KG_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// cache file of the GPU layout: parsing a 10-100 GB reference-format file -- and inflating it when it
// is .gz -- dominates a short run; the built bucket lines + prefilter are written once and read back at file speed.
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct CacheHeader {
    char magic[8];           // "KGB200T\0"
    uint32_t format;         // bumped whenever the bucket / filter layout or the hashes change
    uint32_t bucket_keys, line_bytes, tail_buckets;
    uint32_t num_buckets, filter_words, filter2_words, pad0;
    int32_t shard_rank, shard_count;
    kg_table_info info;
    uint64_t lines_bytes, filter_bytes;
    uint64_t src_size;       // size and mtime of the reference-format file the table was parsed from (0 = unknown):
    int64_t src_mtime_ns;    //   kg_table_load_cached_checked refuses a cache made from another or an updated file
    uint64_t body_check;     // order-independent checksum over every 8-byte word of lines + filters (computed on the device)
    uint64_t check;          // kg_mix over the fields above (catches truncation / foreign files)
};
constexpr uint32_t KG_CACHE_FORMAT = 3;
uint64_t header_check(const CacheHeader& h) {
    uint64_t x = 0x4B47423230305431ull;
    const unsigned char* p = (const unsigned char*)&h;
    for (size_t i = 0; i + 8 <= offsetof(CacheHeader, check); i += 8) {
        uint64_t w;
        memcpy(&w, p + i, 8);
        x = kg_mix(x ^ w);
    }
    return x;
}
// sum over kg_mix(word ^ index): any flipped bit, swapped or missing chunk changes it; runs at HBM speed
__global__ void k_body_check(const unsigned long long* __restrict__ words, size_t n, uint64_t salt, unsigned long long* __restrict__ out) {
    unsigned long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        acc += kg_mix(words[i] ^ (salt + i));
#pragma unroll
    for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}
int body_check_of(kg_context* ctx, const kg_table* t, uint64_t* out) {
    unsigned long long* d = nullptr;
    CU(cudaMalloc(&d, 8));
    cudaStream_t st = ctx->stream;
    CU(cudaMemsetAsync(d, 0, 8, st));
    const size_t nl = ((size_t)t->num_buckets + KG_TAIL_BUCKETS) * 16, nf = (size_t)t->filter_words + t->filter2_words;
    k_body_check<<<148 * 8, 256, 0, st>>>((const unsigned long long*)t->d_lines, nl, 0x11ull << 56, d);
    if (nf) k_body_check<<<148 * 8, 256, 0, st>>>(t->d_filter, nf, 0x22ull << 56, d);
    cudaError_t e = cudaMemcpyAsync(out, d, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    CU(e);
    return KG_OK;
}
bool read_cache_header(FILE* f, CacheHeader* h) {
    struct stat st;
    bool ok = fread(h, sizeof *h, 1, f) == 1 && memcmp(h->magic, "KGB200T", 8) == 0 && h->check == header_check(*h);
    if (ok && (h->format != KG_CACHE_FORMAT || h->bucket_keys != KG_BUCKET_KEYS || h->line_bytes != 128 || h->tail_buckets != KG_TAIL_BUCKETS))
        ok = false; // written by a build with another layout: rebuild from the reference file
    if (ok) ok = h->lines_bytes == ((uint64_t)h->num_buckets + KG_TAIL_BUCKETS) * 128 &&
                 h->filter_bytes == ((uint64_t)h->filter_words + h->filter2_words) * 8 && fstat(fileno(f), &st) == 0 &&
                 (uint64_t)st.st_size == sizeof *h + h->lines_bytes + h->filter_bytes;
    return ok;
}
// the reference-format file a data directory resolves to (the .gz wins, KGJ:749-753)
bool stat_source(const char* data_dir, uint64_t* size, int64_t* mtime_ns) {
    const std::string base = std::string(data_dir) + "/kmer.table.mem_map", gz = base + ".gz";
    struct stat st;
    if (stat(gz.c_str(), &st) != 0 && stat(base.c_str(), &st) != 0) return false;
    *size = (uint64_t)st.st_size;
    *mtime_ns = (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec;
    return true;
}
} // namespace

extern "C" int kg_table_save(kg_context* ctx, const kg_table* t, const char* path) {
    if (!ctx || !t || !path) KG_FAIL(KG_EINVAL, "kg_table_save: null argument");
    CU(cudaSetDevice(ctx->device));
    CacheHeader h = {};
    memcpy(h.magic, "KGB200T", 8);
    h.format = KG_CACHE_FORMAT;
    h.bucket_keys = KG_BUCKET_KEYS;
    h.line_bytes = 128;
    h.tail_buckets = KG_TAIL_BUCKETS;
    h.num_buckets = t->num_buckets;
    h.filter_words = t->filter_words;
    h.filter2_words = t->filter2_words;
    h.src_size = t->src_size;
    h.src_mtime_ns = t->src_mtime_ns;
    KG_TRY(body_check_of(ctx, t, &h.body_check));
    h.shard_rank = t->shard_rank;
    h.shard_count = t->shard_count;
    h.info = t->info;
    h.lines_bytes = ((uint64_t)t->num_buckets + KG_TAIL_BUCKETS) * 128;
    h.filter_bytes = ((uint64_t)t->filter_words + t->filter2_words) * 8;
    h.check = header_check(h);
    std::string tmp = std::string(path) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) KG_FAIL(KG_EIO, "cannot write %s", tmp.c_str());
    const size_t CH = 64u << 20;
    void* stage = nullptr;
    if (cudaMallocHost(&stage, CH) != cudaSuccess) {
        fclose(f);
        cudaGetLastError();
        KG_FAIL(KG_ENOMEM, "kg_table_save: pinned staging buffer");
    }
    bool ok = fwrite(&h, sizeof h, 1, f) == 1;
    auto dump = [&](const void* dev, uint64_t bytes) {
        for (uint64_t o = 0; ok && o < bytes; o += CH) {
            const size_t n = (size_t)std::min<uint64_t>(CH, bytes - o);
            ok = cudaMemcpy(stage, (const char*)dev + o, n, cudaMemcpyDeviceToHost) == cudaSuccess && fwrite(stage, 1, n, f) == n;
        }
    };
    dump(t->d_lines, h.lines_bytes);
    dump(t->d_filter, h.filter_bytes);
    cudaFreeHost(stage);
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), path) != 0) {
        remove(tmp.c_str());
        cudaGetLastError();
        KG_FAIL(KG_EIO, "kg_table_save: writing %s failed", path);
    }
    return KG_OK;
}

extern "C" int kg_table_load_cached(kg_context* ctx, const char* path, kg_table** table) {
    if (!ctx || !path || !table) KG_FAIL(KG_EINVAL, "kg_table_load_cached: null argument");
    CU(cudaSetDevice(ctx->device));
    FILE* f = fopen(path, "rb");
    if (!f) KG_FAIL(KG_EIO, "cannot open %s", path);
    CacheHeader h;
    if (!read_cache_header(f, &h)) {
        fclose(f);
        KG_FAIL(KG_EFORMAT, "%s is not a table cache of this build (bad magic, layout, checksum or size)", path);
    }
    kg_table* t = new kg_table();
    t->ctx = ctx;
    t->num_buckets = h.num_buckets;
    t->filter_words = h.filter_words;
    t->filter2_words = h.filter2_words;
    t->src_size = h.src_size;
    t->src_mtime_ns = h.src_mtime_ns;
    t->shard_rank = h.shard_rank;
    t->shard_count = h.shard_count;
    t->info = h.info;
    const size_t CH = 64u << 20;
    void* stage[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int rc = KG_OK;
    do {
        if (cudaMalloc(&t->d_lines, h.lines_bytes) != cudaSuccess || (h.filter_bytes && cudaMalloc(&t->d_filter, h.filter_bytes) != cudaSuccess) ||
            cudaMallocHost(&stage[0], CH) != cudaSuccess || cudaMallocHost(&stage[1], CH) != cudaSuccess ||
            cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            kg_set_error("kg_table_load_cached: allocation of %llu bytes failed", (unsigned long long)(h.lines_bytes + h.filter_bytes));
            rc = KG_ENOMEM;
            break;
        }
        int k = 0;
        auto fill = [&](void* dev, uint64_t bytes) { // fread into one pinned buffer while the other one is in flight
            for (uint64_t o = 0; rc == KG_OK && o < bytes; o += CH, k ^= 1) {
                const size_t n = (size_t)std::min<uint64_t>(CH, bytes - o);
                cudaEventSynchronize(ev[k]);
                if (fread(stage[k], 1, n, f) != n) {
                    kg_set_error("%s: short read", path);
                    rc = KG_EIO;
                    break;
                }
                cudaMemcpyAsync((char*)dev + o, stage[k], n, cudaMemcpyHostToDevice, ctx->stream);
                cudaEventRecord(ev[k], ctx->stream);
            }
        };
        fill(t->d_lines, h.lines_bytes);
        fill(t->d_filter, h.filter_bytes);
        if (rc == KG_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            kg_set_error("kg_table_load_cached: %s", cudaGetErrorString(cudaGetLastError()));
            rc = KG_ECUDA;
        }
    } while (0);
    fclose(f);
    for (int i = 0; i < 2; i++) {
        if (stage[i]) cudaFreeHost(stage[i]);
        if (ev[i]) cudaEventDestroy(ev[i]);
    }
    if (rc == KG_OK && !getenv("KG_CACHE_NO_VERIFY")) { // bit rot / a torn write in the 8 GB body would silently change CALL lines
        uint64_t chk = 0;
        rc = body_check_of(ctx, t, &chk);
        if (rc == KG_OK && chk != h.body_check) {
            kg_set_error("%s: body checksum mismatch (file damaged); rebuild it from the reference-format table", path);
            rc = KG_EFORMAT;
        }
    }
    if (rc != KG_OK) {
        kg_table_free(t);
        return rc;
    }
    pin_filter(ctx, t, false);
    *table = t;
    return KG_OK;
}

extern "C" int kg_table_load_cached_checked(kg_context* ctx, const char* path, const char* data_dir, kg_table** table) {
    if (!ctx || !path || !data_dir || !table) KG_FAIL(KG_EINVAL, "kg_table_load_cached_checked: null argument");
    uint64_t size = 0;
    int64_t mtime = 0;
    if (!stat_source(data_dir, &size, &mtime)) KG_FAIL(KG_EIO, "no kmer.table.mem_map[.gz] in %s", data_dir);
    FILE* f = fopen(path, "rb");
    if (!f) KG_FAIL(KG_EIO, "cannot open %s", path);
    CacheHeader h;
    const bool ok = read_cache_header(f, &h);
    fclose(f);
    if (!ok) KG_FAIL(KG_EFORMAT, "%s is not a table cache of this build (bad magic, layout, checksum or size)", path);
    if (h.src_size != size || h.src_mtime_ns != mtime)
        KG_FAIL(KG_EFORMAT, "%s was built from another (or an older) kmer.table.mem_map than the one in %s", path, data_dir);
    return kg_table_load_cached(ctx, path, table);
}

extern "C" int kg_table_attach(kg_context* ctx, const kg_table* table) {
    if (!ctx || !table) KG_FAIL(KG_EINVAL, "kg_table_attach: null argument");
    if (ctx->device != table->ctx->device) KG_FAIL(KG_EINVAL, "kg_table_attach: the table lives on device %d, the context on device %d", table->ctx->device, ctx->device);
    CU(cudaSetDevice(ctx->device));
    pin_filter(ctx, table, false); // the persisting-L2 window is a per-stream attribute
    return KG_OK;
}

extern "C" int kg_table_get_info(const kg_table* table, kg_table_info* info) {
    if (!table || !info) KG_FAIL(KG_EINVAL, "kg_table_get_info: null argument");
    *info = table->info;
    return KG_OK;
}

extern "C" void kg_table_free(kg_table* t) {
    if (!t) return;
    if (t->d_lines) cudaFree(t->d_lines);
    if (t->d_filter) cudaFree(t->d_filter);
    delete t;
}
