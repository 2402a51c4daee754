// kg_table.cu -- loader for the reference's kmer.table.mem_map[.gz] and the device-side builder of the
// one-sector-per-probe bucket table (layout: kg_common.cuh).
//
// Replaces: readKmerTableHeader (KGJ:924-942), the 24-byte little-endian entry decode of lookup (KGJ:995-999,
// 1097-1130) and -- semantically -- the table side of the sort-merge join (KGJ:959-1026).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <thread>

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

namespace {

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

__global__ void k_filter_halves_build(const uint64_t* __restrict__ comp, size_t n, unsigned long long* __restrict__ filter, uint32_t words) {
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint64_t m = kg_fhash1(comp[r] & 0x7FFFFFFFFull);
    atomicOr(&filter[(size_t)kg_filter_half(m) * words + kg_filter_half_word(m, words)], kg_filter_mask(m));
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
    v.halves = filter_halves;
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
    if (!force && t->filter2_words && !t->filter_halves && t->shard_count == 1 && !getenv("KG_CASCADE_PERSIST")) return;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, ctx->device) != cudaSuccess || prop.persistingL2CacheMaxSize <= 0) return;
    const size_t fbytes = (size_t)t->filter_words * 8;
    size_t carve = std::min<size_t>(fbytes, (size_t)prop.persistingL2CacheMaxSize);
    if (const char* e = getenv("KG_L2_CARVE_MB")) // experiments: a set-aside larger than the filter (slack for uneven L2 slices)
        carve = std::min<size_t>((size_t)atoll(e) << 20, (size_t)prop.persistingL2CacheMaxSize);
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
    t->l2_carve = carve;
    if (t->filter_halves) { // the two halves take turns in the set-aside: each pass launches with its own window (kg_run.cu)
        cudaGetLastError();
        return;
    }
    cudaStreamAttrValue av = {};
    av.accessPolicyWindow.base_ptr = t->d_filter;
    av.accessPolicyWindow.num_bytes = std::min<size_t>(fbytes, (size_t)prop.accessPolicyMaxWindowSize);
    // a filter larger than the carve-out: that fraction of its lines persists, the rest competes normally
    av.accessPolicyWindow.hitRatio = fbytes > carve ? (float)((double)carve / (double)fbytes) : 1.0f;
    if (const char* e = getenv("KG_L2_HIT_RATIO")) av.accessPolicyWindow.hitRatio = (float)atof(e);
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

    struct Tmp : DevBuf { // freed on every return path (a failed build must not leak gigabytes: the caller may retry)
        ~Tmp() { release(); }
    };
    Tmp ctr_buf, comp_a_buf, comp_b_buf, idx_a_buf, idx_b_buf, flag_buf, nsel_buf, tb_buf;
    KG_TRY(ctr_buf.ensure(4 * sizeof(unsigned long long)));
    unsigned long long* d_ctr = ctr_buf.as<unsigned long long>(); // [0] scatter overflow, [1] flagged buckets, [2] verify failures
    CU(cudaMemsetAsync(d_ctr, 0, 4 * sizeof(unsigned long long), st));
    size_t n_unique = n;

    if (n > 0) {
        KG_TRY(comp_a_buf.ensure(n * 8));
        KG_TRY(comp_b_buf.ensure(n * 8));
        KG_TRY(idx_a_buf.ensure(n * 4));
        KG_TRY(idx_b_buf.ensure(n * 4));
        KG_TRY(flag_buf.ensure(n));
        KG_TRY(nsel_buf.ensure(sizeof(size_t)));
        uint64_t *comp_a = comp_a_buf.as<uint64_t>(), *comp_b = comp_b_buf.as<uint64_t>();
        uint32_t *idx_a = idx_a_buf.as<uint32_t>(), *idx_b = idx_b_buf.as<uint32_t>();
        uint8_t* flag = flag_buf.as<uint8_t>();
        size_t* d_nsel = nsel_buf.as<size_t>();
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

        KG_TRY(tb_buf.ensure(std::max<size_t>(n_unique, 1) * 8));
        long long* tb = tb_buf.as<long long>();
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
            const bool halves = getenv("KG_FILTER_HALVES") && atoi(getenv("KG_FILTER_HALVES")) > 0 && t->shard_count == 1;
            if (halves) { // two filters of up to the cap each, one per half of the keys: twice the bits per key
                bytes = std::min<uint64_t>((uint64_t)((double)n_unique * 0.5 * 2.0 * filter_bits_per_key() / 8.0), max_bytes);
                if (bytes < 4096) bytes = 4096;
            }
            t->filter_words = (uint32_t)(bytes / 8) & ~15u; // whole 128-byte lines, so that the second filter starts on one
            t->filter2_words = (halves || filter_stages() >= 2) ? t->filter_words : 0;
            t->filter_halves = halves ? 1u : 0u;
            const size_t fwords = (size_t)t->filter_words + t->filter2_words;
            CU(cudaMalloc(&t->d_filter, fwords * 8));
            CU(cudaMemsetAsync(t->d_filter, 0, fwords * 8, st));
            if (t->filter_halves)
                k_filter_halves_build<<<blocks_for(n_unique, 256), 256, 0, st>>>(comp, n_unique, t->d_filter, t->filter_words);
            else
                k_filter_build<<<blocks_for(n_unique, 256), 256, 0, st>>>(comp, n_unique, t->d_filter, t->filter_words);
            if (t->filter2_words && !t->filter_halves)
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
    }
    unsigned long long h_ctr[4] = {0, 0, 0, 0};
    CU(cudaMemcpyAsync(h_ctr, d_ctr, sizeof h_ctr, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
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

static int check_shard(int rank, int nranks) {
    if (nranks < 1 || nranks > KG_MAX_RANKS || rank < 0 || rank >= nranks) KG_FAIL(KG_EINVAL, "shard %d of %d: need 0 <= rank < nranks <= %d", rank, nranks, KG_MAX_RANKS);
    return KG_OK;
}

extern "C" int kg_shard_owner(uint64_t key, int nranks) {
    if (nranks < 1) return -1;
    return (int)kg_owner_of(key, (uint32_t)nranks);
}

// ---------------------------------------------------------------------------------------------------------------
// Reference-format loader, parsed ON THE DEVICE.  The file is a header of three little-endian int64
// (numSigs = slot count, entrySize, version; KGJ:933-935) followed by 24-byte slots {int64 key; int32 otuIndex; int32
// avgFromEnd; int32 functionIndex; float32 functionWt} (KGJ:995-999).  The host only moves bytes: reader threads pread()
// 48 MiB chunks of whole slots into a ring of pinned buffers (one inflating thread for .gz -- a gzip stream cannot be cut),
// the copy stream uploads them, and per chunk the compute stream
//   k_ld_mark     t[i] = index after slot i if the slot is empty (key > 20^8, KGJ:1000), else 0
//   max-scan      run start of every slot = index after the last empty slot before it (carried from chunk to chunk in
//                 a device-resident LdState, so no host round trip orders the chunks)
//   k_ld_flag     the reference probes h = key % numSigs, h+1, ... up to the first empty slot and never wraps
//                 (KGJ:959-1026): slot s is reachable iff h lies inside the occupied run that ends at s.  Counts the
//                 unreachable / unmatchable slots and (hash-sharded table) the keys of other ranks
//   select+gather the kept (key, payload) pairs, appended in SLOT order (of two equal keys the earlier slot wins)
// No std::vector of entries on the host (r01: 24 bytes x reachable keys of host RAM and a single-threaded push_back loop).
// ---------------------------------------------------------------------------------------------------------------
namespace {
constexpr size_t LD_SLOTS = 1u << 21;      // slots per chunk
constexpr size_t LD_CHUNK = LD_SLOTS * 24; // 48 MiB
constexpr int LD_NBUF = 4;                 // pinned host buffers
constexpr int LD_READERS = 3;

struct LdState { // device-resident parser state, carried from chunk to chunk in stream order
    long long carry;            // index after the last empty slot so far = start of the occupied run that reaches the next slot
    unsigned long long n_out;   // entries kept so far
    unsigned long long unreachable, unmatchable, not_owned;
    unsigned long long nsel;    // entries kept from the chunk in flight
    unsigned long long last_occ; // the last slot seen is occupied
    unsigned long long overflow; // the output arrays were too small (cannot happen: the host sizes them from an upper bound)
};

__global__ void k_ld_mark(const unsigned long long* __restrict__ raw, uint32_t n, long long g0, long long* __restrict__ t) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long key = (long long)raw[3 * (size_t)i];
    t[i] = key > (long long)KG_MAX_ENCODED ? g0 + (long long)i + 1 : 0;
}
__global__ void k_ld_flag(const unsigned long long* __restrict__ raw, uint32_t n, long long g0, const long long* __restrict__ rs,
                          long long num_slots, uint32_t rank, uint32_t nranks, LdState* __restrict__ st, uint8_t* __restrict__ flag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false, unreach = false, unmatch = false, foreign = false;
    if (i < n) {
        const long long key = (long long)raw[3 * (size_t)i];
        if (key <= (long long)KG_MAX_ENCODED) { // occupied
            if (key >= 0 && key < (long long)KG_MAX_ENCODED) {
                const long long run_start = max(rs[i], st->carry), h = key % num_slots, slot = g0 + (long long)i;
                if (h >= run_start && h <= slot) {
                    if (nranks > 1 && kg_owner_of((uint64_t)key, nranks) != rank) foreign = true;
                    else keep = true;
                } else {
                    unreach = true;
                }
            } else {
                unmatch = true; // occupies a slot (extends chains) but no valid 8-mer encodes to it
            }
        }
        flag[i] = keep;
    }
    const unsigned a = __ballot_sync(0xFFFFFFFFu, unreach), b = __ballot_sync(0xFFFFFFFFu, unmatch), c = __ballot_sync(0xFFFFFFFFu, foreign);
    if ((threadIdx.x & 31) == 0) {
        if (a) atomicAdd(&st->unreachable, (unsigned long long)__popc(a));
        if (b) atomicAdd(&st->unmatchable, (unsigned long long)__popc(b));
        if (c) atomicAdd(&st->not_owned, (unsigned long long)__popc(c));
    }
}
__global__ void k_ld_gather(const unsigned long long* __restrict__ raw, const uint32_t* __restrict__ sel, LdState* __restrict__ st,
                            uint64_t* __restrict__ keys, int4* __restrict__ payload, unsigned long long cap) {
    const unsigned long long nsel = st->nsel, base = st->n_out;
    if (base + nsel > cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) st->overflow = 1;
        return;
    }
    for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < nsel; j += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long* e = raw + 3 * (size_t)sel[j];
        keys[base + j] = e[0];
        const unsigned long long lo = e[1], hi = e[2]; // otuIndex | avgFromEnd << 32, functionIndex | functionWt bits << 32
        payload[base + j] = make_int4((int)(uint32_t)lo, (int)(uint32_t)(lo >> 32), (int)(uint32_t)hi, (int)(uint32_t)(hi >> 32));
    }
}
__global__ void k_ld_advance(const unsigned long long* __restrict__ raw, const long long* __restrict__ rs, uint32_t n, LdState* __restrict__ st) {
    if (!st->overflow) st->n_out += st->nsel;
    st->carry = max(st->carry, rs[n - 1]);
    st->last_occ = (long long)raw[3 * (size_t)(n - 1)] <= (long long)KG_MAX_ENCODED;
}

struct SlotLoader {
    kg_context* ctx = nullptr;
    int rank = 0, nranks = 1;
    int64_t num_slots = 0, entry_size = 0, version = 0;
    uint8_t* h_buf[LD_NBUF] = {};
    cudaEvent_t ev_h2d[LD_NBUF] = {};
    uint8_t* d_raw[2] = {};
    cudaEvent_t ev_done[2] = {};
    DevBuf t, flag, sel, keys, payload, state, tmp;
    uint64_t cap = 0, upper = 0, slots_done = 0, chunks = 0;

    ~SlotLoader() {
        for (auto& b : h_buf)
            if (b) cudaFreeHost(b);
        for (auto& e : ev_h2d)
            if (e) cudaEventDestroy(e);
        for (auto& b : d_raw)
            if (b) cudaFree(b);
        for (auto& e : ev_done)
            if (e) cudaEventDestroy(e);
        for (DevBuf* d : {&t, &flag, &sel, &keys, &payload, &state, &tmp}) d->release();
    }
    int header(const uint8_t* h) { // readLongLE x3, KGJ:933-935 (x86 is little-endian)
        memcpy(&num_slots, h, 8);
        memcpy(&entry_size, h + 8, 8);
        memcpy(&version, h + 16, 8);
        if (entry_size != 24) // KGJ:992 skips by entrySize but KGJ:995-999 always reads 24 bytes
            KG_FAIL(KG_EFORMAT, "kmer table: entrySize %lld != 24 is not readable by the reference either", (long long)entry_size);
        if (num_slots <= 0) KG_FAIL(KG_EFORMAT, "kmer table: numSigs %lld <= 0", (long long)num_slots);
        return KG_OK;
    }
    int init(kg_context* c, int rk, int nr, uint64_t expect_slots, bool exact_size) {
        ctx = c;
        rank = rk;
        nranks = nr;
        CU(cudaSetDevice(ctx->device));
        const size_t chunk = (size_t)std::min<uint64_t>(LD_CHUNK, std::max<uint64_t>(expect_slots, 1) * 24);
        for (auto& b : h_buf) CU(cudaMallocHost(&b, chunk));
        for (auto& e : ev_h2d) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& b : d_raw) CU(cudaMalloc(&b, chunk));
        for (auto& e : ev_done) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        const size_t ns = chunk / 24;
        KG_TRY(t.ensure(ns * 8));
        KG_TRY(flag.ensure(ns));
        KG_TRY(sel.ensure(ns * 4));
        KG_TRY(state.ensure(sizeof(LdState)));
        CU(cudaMemsetAsync(state.p, 0, sizeof(LdState), ctx->stream));
        // the table is an open-addressing table: normally 1/3 to 1/2 of the slots hold a key; the arrays grow if more do
        uint64_t guess = std::max<uint64_t>(expect_slots, 1) * 6 / 10 / (uint64_t)nranks + 2 * ns;
        guess = std::min<uint64_t>(guess, std::max<uint64_t>(expect_slots, 1));
        if (!exact_size) guess = std::min<uint64_t>(guess, 64ull << 20); // .gz: the header's slot count is all there is to go by
        return grow(guess);
    }
    int grow(uint64_t want) {
        if (want <= cap) return KG_OK;
        DevBuf nk, np;
        KG_TRY(nk.ensure(want * 8));
        if (int rc = np.ensure(want * sizeof(int4)); rc != KG_OK) {
            nk.release();
            return rc;
        }
        const uint64_t live = std::min<uint64_t>(cap, upper); // called with the stream drained and `upper` exact
        if (live) {
            CU(cudaMemcpyAsync(nk.p, keys.p, live * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            CU(cudaMemcpyAsync(np.p, payload.p, live * sizeof(int4), cudaMemcpyDeviceToDevice, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
        }
        keys.release();
        payload.release();
        keys = nk;
        payload = np;
        cap = want;
        return KG_OK;
    }
    // host buffer for the next chunks: buffer b is free once the upload of the chunk that used it has completed
    uint8_t* buffer(uint64_t chunk_no) {
        cudaEventSynchronize(ev_h2d[chunk_no % LD_NBUF]);
        return h_buf[chunk_no % LD_NBUF];
    }
    // nbytes of whole slots sit in buffer(chunk_no); chunks are submitted in order
    int submit(uint64_t chunk_no, size_t nbytes) {
        const uint32_t n = (uint32_t)(nbytes / 24);
        if (!n) return KG_OK;
        const int hb = (int)(chunk_no % LD_NBUF), db = (int)(chunks & 1);
        cudaStream_t cs = ctx->copy_stream, st = ctx->stream;
        // Room for everything this chunk could keep.  `upper` counts every slot submitted since the exact count was last
        // read; when it reaches the capacity the stream is drained once, the exact count read, and the arrays grow if
        // the table really is that full (one or two drains per file at the usual 1/3 - 1/2 occupancy).
        if (upper + n > cap) {
            CU(cudaStreamSynchronize(st));
            LdState now;
            CU(cudaMemcpy(&now, state.p, sizeof now, cudaMemcpyDeviceToHost));
            upper = now.n_out;
            if (upper + n > cap) KG_TRY(grow(std::max<uint64_t>(upper + n, cap + cap / 2)));
        }
        upper += n;
        CU(cudaStreamWaitEvent(cs, ev_done[db], 0)); // the kernels of chunk - 2 are done with this device buffer
        CU(cudaMemcpyAsync(d_raw[db], h_buf[hb], (size_t)n * 24, cudaMemcpyHostToDevice, cs));
        CU(cudaEventRecord(ev_h2d[hb], cs));
        CU(cudaStreamWaitEvent(st, ev_h2d[hb], 0));
        const unsigned long long* raw = (const unsigned long long*)d_raw[db];
        const long long g0 = (long long)slots_done;
        LdState* ds = state.as<LdState>();
        k_ld_mark<<<blocks_for(n, 256), 256, 0, st>>>(raw, n, g0, t.as<long long>());
        size_t b1 = 0, b2 = 0;
        CU(cub::DeviceScan::InclusiveScan(nullptr, b1, t.as<long long>(), t.as<long long>(), MaxI64(), n, st));
        thrust::counting_iterator<uint32_t> iota(0);
        CU(cub::DeviceSelect::Flagged(nullptr, b2, iota, flag.as<uint8_t>(), sel.as<uint32_t>(), &ds->nsel, n, st));
        KG_TRY(tmp.ensure(std::max(b1, b2)));
        CU(cub::DeviceScan::InclusiveScan(tmp.p, b1, t.as<long long>(), t.as<long long>(), MaxI64(), n, st));
        k_ld_flag<<<blocks_for(n, 256), 256, 0, st>>>(raw, n, g0, t.as<long long>(), (long long)num_slots, (uint32_t)rank, (uint32_t)nranks, ds,
                                                     flag.as<uint8_t>());
        CU(cub::DeviceSelect::Flagged(tmp.p, b2, iota, flag.as<uint8_t>(), sel.as<uint32_t>(), &ds->nsel, n, st));
        k_ld_gather<<<ctx->sm_count * 4, 256, 0, st>>>(raw, sel.as<uint32_t>(), ds, keys.as<uint64_t>(), payload.as<int4>(), cap);
        k_ld_advance<<<1, 1, 0, st>>>(raw, t.as<long long>(), n, ds);
        CU(cudaEventRecord(ev_done[db], st));
        slots_done += n;
        chunks++;
        CU(cudaGetLastError());
        return KG_OK;
    }
    int finish(kg_table** out) {
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaStreamSynchronize(ctx->copy_stream));
        LdState fin;
        CU(cudaMemcpy(&fin, state.p, sizeof fin, cudaMemcpyDeviceToHost));
        if (fin.overflow) KG_FAIL(KG_ECUDA, "table loader: output arrays overflowed (internal error)");
        kg_table* tb = new kg_table();
        tb->ctx = ctx;
        tb->info.num_slots = num_slots;
        tb->info.entry_size = entry_size;
        tb->info.version = version;
        tb->info.num_unreachable = (int64_t)(fin.unreachable + fin.unmatchable);
        tb->info.tail_run = fin.last_occ ? (int64_t)slots_done - (int64_t)fin.carry : 0;
        tb->shard_rank = rank;
        tb->shard_count = nranks;
        // the parser's own buffers are no longer needed while the builder runs
        for (DevBuf* d : {&t, &flag, &sel, &tmp}) d->release();
        for (auto& b : d_raw) {
            if (b) cudaFree(b);
            b = nullptr;
        }
        const int rc = build_on_device(ctx, keys.as<uint64_t>(), payload.as<int4>(), (size_t)fin.n_out, table_load_factor(), true, tb);
        if (rc != KG_OK) {
            kg_table_free(tb);
            return rc;
        }
        *out = tb;
        return KG_OK;
    }
};

// Random-access source (a plain file or an image in memory): reader threads fill the pinned ring ahead of the submitting thread.
struct RandomSource {
    int fd = -1;                  // plain file, or
    const uint8_t* image = nullptr; // image in host memory
    uint64_t body = 0;            // bytes of whole slots after the 24-byte header
    bool read(uint64_t off, uint8_t* dst, size_t n) const {
        if (image) {
            memcpy(dst, image + 24 + off, n);
            return true;
        }
        size_t got = 0;
        while (got < n) {
            const ssize_t r = pread(fd, dst + got, n - got, (off_t)(24 + off + got));
            if (r <= 0) return false;
            got += (size_t)r;
        }
        return true;
    }
};

int load_random(kg_context* ctx, const RandomSource& src, const uint8_t* header24, int rank, int nranks, kg_table** table) {
    SlotLoader ld;
    KG_TRY(ld.header(header24));
    const uint64_t nslots = src.body / 24;
    KG_TRY(ld.init(ctx, rank, nranks, nslots, true));
    const size_t chunk = (size_t)std::min<uint64_t>(LD_CHUNK, std::max<uint64_t>(nslots, 1) * 24);
    const uint64_t nchunks = (nslots * 24 + chunk - 1) / chunk;
    // ready[c % NBUF] = c + 1 once chunk c has been read; free_upto = chunks whose host buffer may be overwritten
    std::mutex m;
    std::condition_variable cv;
    uint64_t ready[LD_NBUF] = {}, free_upto = LD_NBUF;
    bool failed = false;
    auto reader = [&](int tid) {
        for (uint64_t c = (uint64_t)tid; c < nchunks; c += LD_READERS) {
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return c < free_upto || failed; });
                if (failed) return;
            }
            const uint64_t off = c * chunk;
            const size_t n = (size_t)std::min<uint64_t>(chunk, nslots * 24 - off);
            const bool ok = src.read(off, ld.h_buf[c % LD_NBUF], n);
            std::lock_guard<std::mutex> lk(m);
            if (!ok) failed = true;
            ready[c % LD_NBUF] = c + 1;
            cv.notify_all();
        }
    };
    std::vector<std::thread> th;
    const int nthreads = (int)std::min<uint64_t>(LD_READERS, nchunks);
    for (int i = 0; i < nthreads; i++) th.emplace_back(reader, i);
    int rc = KG_OK;
    for (uint64_t c = 0; c < nchunks && rc == KG_OK; c++) {
        {
            std::unique_lock<std::mutex> lk(m);
            cv.wait(lk, [&] { return ready[c % LD_NBUF] == c + 1 || failed; });
            if (failed) {
                kg_set_error("kmer table: read error");
                rc = KG_EIO;
                break;
            }
        }
        const uint64_t off = c * chunk;
        rc = ld.submit(c, (size_t)std::min<uint64_t>(chunk, nslots * 24 - off));
        if (rc != KG_OK) break;
        // the buffer of chunk c can be refilled (by chunk c + NBUF) once its upload is complete
        cudaEventSynchronize(ld.ev_h2d[c % LD_NBUF]);
        std::lock_guard<std::mutex> lk(m);
        free_upto = c + 1 + LD_NBUF;
        cv.notify_all();
    }
    {
        std::lock_guard<std::mutex> lk(m);
        if (rc != KG_OK) failed = true;
        cv.notify_all();
    }
    for (auto& t : th) t.join();
    if (rc != KG_OK) return rc;
    return ld.finish(table);
}

int load_gz(kg_context* ctx, const char* path, int rank, int nranks, kg_table** table) {
    gzFile g = gzopen(path, "rb");
    if (!g) KG_FAIL(KG_EIO, "cannot open %s", path);
    gzbuffer(g, 1 << 20);
    auto read_fully = [&](uint8_t* dst, size_t n, bool* err) -> size_t {
        size_t got = 0;
        while (got < n) {
            const int r = gzread(g, dst + got, (unsigned)std::min<size_t>(n - got, 1u << 30));
            if (r < 0) *err = true;
            if (r <= 0) break;
            got += (size_t)r;
        }
        return got;
    };
    uint8_t hdr[24];
    bool err = false;
    int rc = KG_OK;
    SlotLoader ld;
    do {
        if (read_fully(hdr, 24, &err) != 24) {
            kg_set_error(err ? "gzip error reading %s" : "kmer table: EOF inside the 24-byte header (%s)", path);
            rc = err ? KG_EIO : KG_EIO;
            break;
        }
        if ((rc = ld.header(hdr)) != KG_OK) break;
        if ((rc = ld.init(ctx, rank, nranks, (uint64_t)ld.num_slots, false)) != KG_OK) break;
        const size_t chunk = (size_t)std::min<uint64_t>(LD_CHUNK, (uint64_t)ld.num_slots * 24);
        for (uint64_t c = 0;; c++) { // the inflate of chunk c+1 overlaps the upload and the kernels of chunk c
            uint8_t* buf = ld.buffer(c);
            const size_t got = read_fully(buf, chunk, &err);
            if (err) {
                kg_set_error("gzip error reading %s", path);
                rc = KG_EIO;
                break;
            }
            if ((rc = ld.submit(c, got / 24 * 24)) != KG_OK) break;
            if (got < chunk) break;
        }
    } while (0);
    gzclose(g);
    if (rc != KG_OK) return rc;
    return ld.finish(table);
}
} // namespace

extern "C" int kg_table_from_image_sharded(kg_context* ctx, const void* image, size_t nbytes, int rank, int nranks, kg_table** table) {
    if (!ctx || !image || !table) KG_FAIL(KG_EINVAL, "kg_table_from_image: null argument");
    KG_TRY(check_shard(rank, nranks));
    if (nbytes < 24) KG_FAIL(KG_EIO, "kmer table: EOF inside the 24-byte header");
    RandomSource src;
    src.image = (const uint8_t*)image;
    src.body = (nbytes - 24) / 24 * 24;
    return load_random(ctx, src, (const uint8_t*)image, rank, nranks, table);
}

extern "C" int kg_table_from_image(kg_context* ctx, const void* image, size_t nbytes, kg_table** table) {
    return kg_table_from_image_sharded(ctx, image, nbytes, 0, 1, table);
}

static int load_file_impl(kg_context* ctx, const char* path, int rank, int nranks, kg_table** table) {
    if (!ctx || !path || !table) KG_FAIL(KG_EINVAL, "kg_table_load_file: null argument");
    const size_t len = strlen(path);
    const bool gz = len > 3 && strcmp(path + len - 3, ".gz") == 0; // the reference keys on the suffix (KGJ:927)
    int rc;
    if (gz) {
        rc = load_gz(ctx, path, rank, nranks, table);
    } else {
        const int fd = open(path, O_RDONLY);
        if (fd < 0) KG_FAIL(KG_EIO, "cannot open %s", path);
        struct stat fst;
        uint8_t hdr[24];
        if (fstat(fd, &fst) != 0 || fst.st_size < 24 || pread(fd, hdr, 24, 0) != 24) {
            close(fd);
            KG_FAIL(KG_EIO, "kmer table: EOF inside the 24-byte header (%s)", path);
        }
        posix_fadvise(fd, 0, 0, POSIX_FADV_SEQUENTIAL);
        RandomSource src;
        src.fd = fd;
        src.body = ((uint64_t)fst.st_size - 24) / 24 * 24;
        rc = load_random(ctx, src, hdr, rank, nranks, table);
        close(fd);
    }
    if (rc != KG_OK) return rc;
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
    h.pad0 = t->filter_halves;
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
    t->filter_halves = h.pad0;
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
