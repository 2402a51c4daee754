// kg_run.cu -- the device pipeline: residue stream -> encode+probe -> ordered hits -> run FSM -> CALL / OTU records.
//
// Replaces, for a whole batch of sequences at once:
//   prepareQuery  KGJ:1051-1074   (k_patch_aa / k_vlen + k_translate)
//   addKmers      KGJ:900-922     (k_probe: window enumeration, base-20 encoding)
//   sort + lookup KGJ:1076-1095, 944-1034 (k_probe: one 32-byte sector per k-mer instead of a sort-merge join)
//   gatherHits / processSetOfHits KGJ:457-514, 385-455 (k_fsm)
#include <cub/device/device_scan.cuh>

#include <algorithm>

#include "kg_device.cuh"
#include "kg_fsm.cuh"
#include "kg_internal.h"

namespace {

constexpr int PT = 16;            // positions per thread
constexpr int PROBE_BLK = 256;    // threads per block
constexpr int TILE = PT * PROBE_BLK;
constexpr int TILE_SHIFT = 12;
static_assert(TILE == (1 << TILE_SHIFT), "tile size");
constexpr int STAGE_STRIDE = PT + 1; // +1: conflict-free shared-memory stride

inline unsigned blocks_for(size_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

// toAminoAcidOff (KGJ:111-175) for 'A'..'Z'; everything else is 20
__constant__ uint8_t c_aa_code[26] = {0, 20, 1, 2, 3, 4, 5, 6, 7, 20, 8, 9, 10, 11, 20, 12, 13, 14, 15, 16, 20, 17, 18, 20, 19, 20};
// GENETIC_CODE (KGJ:88-93)
__constant__ char c_genetic_code[65] = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF";

__device__ __forceinline__ int dna_code(uint8_t c) { // dnaChar, KGJ:294-318
    switch (c) {
        case 'a': case 'A': return 0;
        case 'c': case 'C': return 1;
        case 'g': case 'G': return 2;
        case 't': case 'u': case 'T': case 'U': return 3;
        default: return 4;
    }
}

// index of the sequence that holds position g: largest s with off[s] <= g (empty sequences are skipped naturally)
__device__ __forceinline__ uint64_t seq_of(const uint64_t* __restrict__ off, uint64_t n, uint64_t g) {
    uint64_t lo = 0, hi = n; // invariant: off[lo] <= g < off[hi]
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (off[mid] <= g) lo = mid;
        else hi = mid;
    }
    return lo;
}

// ---------------------------------------------------------------------------------------------------------------
// aa mode: the protein stream is used in place.  Overwriting the last residue of every protein with 0 separates
// the proteins AND reproduces the reference's loop bound `i < pIseq.length - K` (KGJ:912 with KGJ:1055), which
// never looks up the window that starts at len-8: that window is the only one containing the last residue.
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_patch_aa(uint8_t* __restrict__ seq, const uint64_t* __restrict__ off, uint64_t n) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint64_t a = off[s], b = off[s + 1];
    if (b > a) seq[b - 1] = 0;
}

// ---------------------------------------------------------------------------------------------------------------
// dna mode: six virtual proteins per contig.  translate() (KGJ:320-343) yields floor((L-f)/3) residues for frame f;
// each virtual protein is followed by >= 1 zero byte (the reference's terminator code 21, KGJ:339-342).
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_vlen(const uint64_t* __restrict__ off, uint64_t n, uint64_t* __restrict__ vlen) {
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v > 6 * n) return;
    if (v == 6 * n) { vlen[v] = 0; return; }
    uint64_t s = v / 6, f = (v % 6) % 3;
    uint64_t L = off[s + 1] - off[s];
    uint64_t nk = L >= f + 3 ? (L - f) / 3 : 0;
    vlen[v] = (nk + 1 + 3) & ~3ull; // residues + terminator, rounded up to 4 (all padding bytes are 0 = separator)
}

__global__ void k_translate(const uint8_t* __restrict__ seq, const uint64_t* __restrict__ off, uint64_t n, uint64_t total,
                            const uint64_t* __restrict__ voff, uint8_t* __restrict__ vseq) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    uint64_t s = seq_of(off, n, g);
    uint64_t p = g - off[s], L = off[s + 1] - off[s];
    if (p + 3 > L) return;
    int c1 = dna_code(seq[g]), c2 = dna_code(seq[g + 1]), c3 = dna_code(seq[g + 2]);
    bool ok = (c1 | c2 | c3) < 4;
    // forward strand: codon at p belongs to frame p%3, residue p/3 (KGJ:323-338)
    uint8_t fw = ok ? (uint8_t)c_genetic_code[c1 * 16 + c2 * 4 + c3] : (uint8_t)'x';
    vseq[voff[6 * s + p % 3] + p / 3] = fw;
    // reverse strand: revComp (KGJ:263-272) puts compl(seq[p+2]), compl(seq[p+1]), compl(seq[p]) at q = L-3-p.
    // compl() maps ACGTU (either case) to the complementary base and never maps anything else onto ACGTU.
    uint64_t q = L - 3 - p;
    uint8_t rv = ok ? (uint8_t)c_genetic_code[(3 - c3) * 16 + (3 - c2) * 4 + (3 - c1)] : (uint8_t)'x';
    vseq[voff[6 * s + 3 + q % 3] + q / 3] = rv;
}

// ---------------------------------------------------------------------------------------------------------------
// encode + probe.  One thread owns PT consecutive stream positions; a block owns a 4096-position tile.
//   * 24 residue bytes per thread (16 + 8 halo) come in as one 16-byte and one 8-byte coalesced load
//   * residue codes via a 256-byte shared-memory LUT; 4-mer partial products give each window in one 64-bit IMAD
//   * PROBE_G independent 256-bit sector loads are in flight per thread before the first compare
//   * hits are compacted in tile order: per-thread slots staged in shared memory, block-wide exclusive scan,
//     ONE global atomic per tile to claim an output chunk; chunks are stitched into global order by k_gather.
// ---------------------------------------------------------------------------------------------------------------
constexpr int PROBE_G = 4;

__global__ __launch_bounds__(PROBE_BLK) void k_probe(const uint8_t* __restrict__ stream, uint32_t vtotal, KgTableView tab,
                                                     uint2* __restrict__ chunk, uint32_t hit_cap,
                                                     uint32_t* __restrict__ tile_base, uint32_t* __restrict__ tile_cnt,
                                                     unsigned long long* __restrict__ ctr) {
    __shared__ uint8_t lut[256];
    __shared__ uint32_t stage[PROBE_BLK * STAGE_STRIDE];
    __shared__ uint32_t warp_hits[PROBE_BLK / 32], warp_kmers[PROBE_BLK / 32];
    __shared__ uint32_t s_base;

    const int tid = threadIdx.x;
    lut[tid] = (tid >= 'A' && tid <= 'Z') ? c_aa_code[tid - 'A'] : 20;
    __syncthreads();

    const uint32_t p0 = blockIdx.x * (uint32_t)TILE + (uint32_t)tid * PT;
    uint32_t hitmask = 0, nk = 0;
    if (p0 < vtotal) {
        const uint4 a = *reinterpret_cast<const uint4*>(stream + p0);
        const uint2 b = *reinterpret_cast<const uint2*>(stream + p0 + 16);
        const uint32_t wv[6] = {a.x, a.y, a.z, a.w, b.x, b.y};
        uint32_t c[24];
        uint32_t bad = 0;
#pragma unroll
        for (int i = 0; i < 24; i++) {
            c[i] = lut[(wv[i >> 2] >> (8 * (i & 3))) & 0xFFu];
            bad |= (uint32_t)(c[i] >= 20u) << i;
        }
        const uint32_t left = vtotal - p0; // bytes of this thread's 24 that exist
        if (left < 24) bad |= ~0u << left;
        // pairs -> 4-mers -> 8-mers (first residue most significant, KGJ:274-282)
        uint32_t q[20];
#pragma unroll
        for (int i = 0; i < 20; i++) q[i] = ((c[i] * 20u + c[i + 1]) * 20u + c[i + 2]) * 20u + c[i + 3];

        uint32_t valid = 0;
#pragma unroll
        for (int i = 0; i < PT; i++) valid |= (uint32_t)(((bad >> i) & 0xFFu) == 0u) << i;
        nk = __popc(valid);

        // phase 1 -- L2-resident prefilter: one 8-byte word per window, eight loads in flight per thread.  Only the
        // windows whose two bits are both set (all true signatures + ~17 % of the rest) go on to DRAM.
        const uint64_t pol_keep = kg_policy_evict_last(), pol_stream = kg_policy_evict_first();
        uint32_t pass = valid;
        if (tab.filter_words) {
            pass = 0;
#pragma unroll
            for (int half = 0; half < PT; half += 8) {
                unsigned long long fw[8], fm[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int i = half + j;
                    const uint64_t h = kg_mix((uint64_t)q[i] * 160000ull + q[i + 4]);
                    fm[j] = kg_filter_mask(h);
                    fw[j] = 0;
                    if ((valid >> i) & 1u) fw[j] = kg_load_filter_word(tab.filter, kg_filter_word(h, tab.filter_words), pol_keep);
                }
#pragma unroll
                for (int j = 0; j < 8; j++) pass |= (uint32_t)((fw[j] & fm[j]) == fm[j]) << (half + j);
            }
            pass &= valid;
        }

        // phase 2 -- the surviving windows probe their bucket: one 256-bit sector load each, PROBE_G in flight
#pragma unroll
        for (int g = 0; g < PT; g += PROBE_G) {
            if (((pass >> g) & ((1u << PROBE_G) - 1u)) == 0u) continue;
            uint64_t key[PROBE_G];
            uint32_t bkt[PROBE_G];
            KgBucket bk[PROBE_G];
#pragma unroll
            for (int j = 0; j < PROBE_G; j++) {
                const int i = g + j;
                key[j] = (uint64_t)q[i] * 160000ull + q[i + 4];
                bkt[j] = kg_home_bucket(key[j], tab.num_buckets);
            }
#pragma unroll
            for (int j = 0; j < PROBE_G; j++)
                if ((pass >> (g + j)) & 1u) bk[j] = kg_load_bucket_hint(tab.buckets, bkt[j], pol_stream);
#pragma unroll
            for (int j = 0; j < PROBE_G; j++) {
                if (!((pass >> (g + j)) & 1u)) continue;
                uint32_t m = kg_bucket_match(bk[j], key[j]);
                uint32_t slot = 0xFFFFFFFFu;
                if (m) slot = bkt[j] * KG_BUCKET_KEYS + (__ffs(m) - 1);
                else if (bk[j].w[7] & KG_W7_FLAG) slot = kg_lookup_from(tab, key[j], bkt[j] + 1); // rare second sector
                if (slot != 0xFFFFFFFFu) {
                    hitmask |= 1u << (g + j);
                    stage[tid * STAGE_STRIDE + g + j] = slot;
                }
            }
        }
    }

    // block-wide exclusive scan of per-thread hit counts (tile order = thread order)
    const int lane = tid & 31, wid = tid >> 5;
    const uint32_t mine = __popc(hitmask);
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
    }
    const uint32_t wk = __reduce_add_sync(0xFFFFFFFFu, nk);
    if (lane == 31) warp_hits[wid] = incl;
    if (lane == 0) warp_kmers[wid] = wk;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < PROBE_BLK / 32; w++) {
        const uint32_t t = warp_hits[w];
        if (w < wid) before += t;
        total += t;
    }
    if (tid == 0) {
        uint32_t kmers = 0;
#pragma unroll
        for (int w = 0; w < PROBE_BLK / 32; w++) kmers += warp_kmers[w];
        if (kmers) atomicAdd(&ctr[KG_CTR_KMERS], (unsigned long long)kmers);
        unsigned long long base = 0;
        if (total) base = atomicAdd(&ctr[KG_CTR_HITS], (unsigned long long)total);
        uint32_t b32 = 0xFFFFFFFFu;
        if (base + total <= (unsigned long long)hit_cap) b32 = (uint32_t)base;
        else ctr[KG_CTR_OVERFLOW] = 1ull; // chunk array too small: the host re-runs with the exact size
        s_base = b32;
        tile_base[blockIdx.x] = b32;
        tile_cnt[blockIdx.x] = total;
    }
    __syncthreads();
    const uint32_t base = s_base;
    if (base != 0xFFFFFFFFu && hitmask) {
        uint32_t o = base + before + (incl - mine);
        uint32_t m = hitmask;
        while (m) {
            const int i = __ffs(m) - 1;
            m &= m - 1;
            chunk[o++] = make_uint2(p0 + (uint32_t)i, stage[tid * STAGE_STRIDE + i]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// stitch the per-tile chunks into one position-sorted hit list and attach the 16-byte payloads (one warp per tile)
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_gather(const uint2* __restrict__ chunk, const uint32_t* __restrict__ tile_base,
                         const uint32_t* __restrict__ tile_out, uint32_t ntiles, const int4* __restrict__ payload,
                         uint32_t* __restrict__ hit_pos, int4* __restrict__ hit_payload) {
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= ntiles) return;
    const uint32_t out = tile_out[t], cnt = tile_out[t + 1] - out, base = tile_base[t];
    for (uint32_t j = lane; j < cnt; j += 32) {
        const uint2 h = chunk[base + j];
        hit_pos[out + j] = h.x;
        hit_payload[out + j] = __ldg(&payload[h.y]);
    }
}

// lo[v] = index of the first hit at or after the start of virtual sequence v (v = nv gives the total)
__global__ void k_ranges(const uint64_t* __restrict__ voff, uint64_t nv, const uint32_t* __restrict__ tile_out,
                         uint32_t ntiles, const uint32_t* __restrict__ hit_pos, uint32_t* __restrict__ lo) {
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v > nv) return;
    const uint64_t x = voff[v];
    const uint64_t t = x >> TILE_SHIFT;
    if (t >= ntiles) {
        lo[v] = tile_out[ntiles];
        return;
    }
    uint32_t a = tile_out[t], b = tile_out[t + 1]; // first hit >= x lies in [a, b]
    while (a < b) {
        uint32_t mid = (a + b) >> 1;
        if ((uint64_t)hit_pos[mid] < x) a = mid + 1;
        else b = mid;
    }
    lo[v] = a;
}

// ---------------------------------------------------------------------------------------------------------------
// run FSM: one thread per sequence walks its containers in the reference's order (+0,+1,+2,-0,-1,-2) so that the
// OTU buffer sees the calls in the same order (KGJ:540-557).  Calls of container v go to the sparse slots
// [lo[v]/min_hits, lo[v+1]/min_hits): a call consumes >= min_hits counted hits and no hit is counted twice, so the
// slots cannot overflow and no second pass is needed to size them.
// ---------------------------------------------------------------------------------------------------------------
struct SparseEmit {
    KgDevCall* dst;
    __device__ __forceinline__ void operator()(int i, const KgDevCall& c) { dst[i] = c; }
};

__global__ __launch_bounds__(128) void k_fsm(const uint64_t* __restrict__ voff, uint64_t nseq, int per_seq,
                                             const uint32_t* __restrict__ lo, const uint32_t* __restrict__ hit_pos,
                                             const int4* __restrict__ hit_payload, KgFsmParams p,
                                             KgDevCall* __restrict__ sparse, uint32_t* __restrict__ call_cnt,
                                             kg_otu* __restrict__ otus) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseq) return;
    KgFsm f;
    f.begin_sequence();
    for (int k = 0; k < per_seq; k++) {
        const uint64_t v = s * per_seq + k;
        const uint32_t a = lo[v], b = lo[v + 1];
        const uint32_t base = (uint32_t)voff[v];
        f.begin_container();
        SparseEmit emit{sparse + a / (uint32_t)p.min_hits};
        for (uint32_t i = a; i < b; i++) {
            const int4 pl = hit_payload[i];
            KgHitLite h = {(int)(hit_pos[i] - base), pl.z, pl.y, pl.x, __int_as_float(pl.w)};
            f.hit(p, h, emit);
        }
        f.end_container(p, emit);
        call_cnt[v] = (uint32_t)f.ncalls;
    }
    kg_otu o;
    o.n = f.otu_c.n;
#pragma unroll
    for (int i = 0; i < KG_OI_BUFSZ; i++) {
        o.count[i] = f.otu_c.c[i];
        o.oI[i] = f.otu_c.o[i];
    }
    otus[s] = o;
}

__global__ void k_compact_calls(const KgDevCall* __restrict__ sparse, const uint32_t* __restrict__ lo,
                                const uint32_t* __restrict__ call_off, uint64_t nv, int per_seq, int min_hits,
                                kg_call* __restrict__ out) {
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const uint32_t o = call_off[v], c = call_off[v + 1] - o;
    const KgDevCall* src = sparse + lo[v] / (uint32_t)min_hits;
    for (uint32_t j = 0; j < c; j++) {
        KgDevCall d = src[j];
        kg_call r;
        r.seq = (uint32_t)(v / per_seq);
        r.strand_frame = (int32_t)(v % per_seq);
        r.start = d.start;
        r.end = d.end;
        r.count = d.count;
        r.fI = d.fI;
        r.weighted = d.weighted;
        r.hits_before = d.hits_before;
        out[o + j] = r;
    }
}

// "-d" HIT records: one thread per hit finds its container by binary search over the virtual offsets
__global__ void k_emit_hits(const uint64_t* __restrict__ voff, uint64_t nv, int per_seq, const uint32_t* __restrict__ hit_pos,
                            const int4* __restrict__ hit_payload, uint32_t nhits, kg_hit* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nhits) return;
    const uint64_t g = hit_pos[i];
    const uint64_t v = seq_of(voff, nv, g);
    const int4 pl = hit_payload[i];
    kg_hit h;
    h.seq = (uint32_t)(v / per_seq);
    h.strand_frame = (int32_t)(v % per_seq);
    h.pos = (int32_t)(g - voff[v]);
    h.oI = pl.x;
    h.avg_off_from_end = pl.y;
    h.fI = pl.z;
    h.function_wt = __int_as_float(pl.w);
    out[i] = h;
}

struct RunScratch { // grow-only device scratch kept per context (so repeated runs do not allocate)
    DevBuf tile_base, tile_cnt, tile_out, chunk, hit_pos, hit_payload, lo, sparse, call_cnt, call_off, ctr;
};
RunScratch& scratch_of(kg_context* ctx) {
    if (!ctx->scratch) ctx->scratch = new RunScratch();
    return *static_cast<RunScratch*>(ctx->scratch);
}

// ---- recycled result buffers ----
int pool_take_dev(kg_context* ctx, size_t bytes, DevBuf* out) {
    int best = -1;
    for (int i = 0; i < (int)ctx->dev_pool.size(); i++)
        if (ctx->dev_pool[i].cap >= bytes && (best < 0 || ctx->dev_pool[i].cap < ctx->dev_pool[best].cap)) best = i;
    if (best >= 0) {
        *out = ctx->dev_pool[best];
        ctx->dev_pool.erase(ctx->dev_pool.begin() + best);
        return KG_OK;
    }
    *out = DevBuf();
    return out->ensure(std::max<size_t>(bytes, 256));
}
void pool_give_dev(kg_context* ctx, DevBuf* b) {
    if (!b->p) return;
    if (ctx->dev_pool.size() >= 12) { // keep the pool bounded: drop the smallest
        size_t k = 0;
        for (size_t i = 1; i < ctx->dev_pool.size(); i++)
            if (ctx->dev_pool[i].cap < ctx->dev_pool[k].cap) k = i;
        ctx->dev_pool[k].release();
        ctx->dev_pool.erase(ctx->dev_pool.begin() + k);
    }
    ctx->dev_pool.push_back(*b);
    *b = DevBuf();
}
int pool_take_host(kg_context* ctx, size_t bytes, HostBuf* out) {
    int best = -1;
    for (int i = 0; i < (int)ctx->host_pool.size(); i++)
        if (ctx->host_pool[i].cap >= bytes && (best < 0 || ctx->host_pool[i].cap < ctx->host_pool[best].cap)) best = i;
    if (best >= 0) {
        *out = ctx->host_pool[best];
        ctx->host_pool.erase(ctx->host_pool.begin() + best);
        return KG_OK;
    }
    size_t want = std::max<size_t>(bytes + bytes / 8, 4096);
    void* p = nullptr;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        KG_FAIL(KG_ENOMEM, "cudaMallocHost(%zu bytes) failed: %s", want, cudaGetErrorString(e));
    }
    out->p = p;
    out->cap = want;
    return KG_OK;
}
void pool_give_host(kg_context* ctx, HostBuf* b) {
    if (!b->p) return;
    if (ctx->host_pool.size() >= 12) {
        size_t k = 0;
        for (size_t i = 1; i < ctx->host_pool.size(); i++)
            if (ctx->host_pool[i].cap < ctx->host_pool[k].cap) k = i;
        cudaFreeHost(ctx->host_pool[k].p);
        ctx->host_pool.erase(ctx->host_pool.begin() + k);
    }
    ctx->host_pool.push_back(*b);
    *b = HostBuf();
}

int exclusive_sum_u32(kg_context* ctx, const uint32_t* in, uint32_t* out, size_t n, cudaStream_t st) {
    size_t bytes = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, st));
    KG_TRY(ctx->scan_tmp.ensure(bytes));
    CU(cub::DeviceScan::ExclusiveSum(ctx->scan_tmp.p, bytes, in, out, n, st));
    return KG_OK;
}
int exclusive_sum_u64(kg_context* ctx, const uint64_t* in, uint64_t* out, size_t n, cudaStream_t st) {
    size_t bytes = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, st));
    KG_TRY(ctx->scan_tmp.ensure(bytes));
    CU(cub::DeviceScan::ExclusiveSum(ctx->scan_tmp.p, bytes, in, out, n, st));
    return KG_OK;
}

} // namespace

// ---------------------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------------------
extern "C" int kg_init(int device, kg_context** out) {
    if (!out) KG_FAIL(KG_EINVAL, "kg_init: null argument");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        KG_FAIL(KG_ENODEV, "no CUDA device (%s); this library has no CPU fallback", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= count) KG_FAIL(KG_EINVAL, "kg_init: device %d out of range (0..%d)", device, count - 1);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) KG_FAIL(KG_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    // Every probe wants ONE 32-byte sector of a multi-GB table.  With the default L2 fetch granularity the memory system
    // brings in the whole 128-byte line per miss (ncu: 124 B of DRAM reads per lookup); ask for sector-sized fetches.
    {
        size_t gran = 32;
        if (const char* e = getenv("KG_L2_FETCH")) gran = (size_t)atoi(e);
        if (gran == 32 || gran == 64 || gran == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        cudaGetLastError();
    }
    kg_context* ctx = new kg_context();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->l2_bytes = (size_t)prop.l2CacheSize;
    CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (auto& ev : ctx->ev) CU(cudaEventCreate(&ev));
    CU(cudaMallocHost(&ctx->h_counters, KG_CTR_COUNT * sizeof(uint64_t)));
    *out = ctx;
    return KG_OK;
}

extern "C" void kg_shutdown(kg_context* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    RunScratch& sc = scratch_of(ctx);
    for (DevBuf* b : {&sc.tile_base, &sc.tile_cnt, &sc.tile_out, &sc.chunk, &sc.hit_pos, &sc.hit_payload, &sc.lo, &sc.sparse,
                      &sc.call_cnt, &sc.call_off, &sc.ctr})
        b->release();
    delete static_cast<RunScratch*>(ctx->scratch);
    for (auto& b : ctx->dev_pool) b.release();
    for (auto& h : ctx->host_pool) cudaFreeHost(h.p);
    ctx->scan_tmp.release();
    for (auto& ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    delete ctx;
}

extern "C" void kg_params_default(kg_params* p) {
    if (!p) return;
    p->min_hits = 5;          // KGJ:104
    p->min_weighted_hits = 0; // KGJ:105
    p->max_gap = 200;         // KGJ:106
    p->order_constraint = 0;  // KGJ:103
    p->emit_hits = 0;
}

// ---------------------------------------------------------------------------------------------------------------
// batches
// ---------------------------------------------------------------------------------------------------------------
static constexpr uint64_t KG_MAX_STREAM = 0xFFFF0000ull; // residue-stream positions are 32-bit

extern "C" int kg_batch_upload(kg_context* ctx, int mode, const uint8_t* seq_bytes, const uint64_t* offsets, size_t n,
                               kg_batch** out) {
    if (!ctx || !out || !offsets || (mode != KG_MODE_AA && mode != KG_MODE_DNA)) KG_FAIL(KG_EINVAL, "kg_batch_upload: bad argument");
    if (offsets[0] != 0) KG_FAIL(KG_EINVAL, "kg_batch_upload: offsets[0] must be 0");
    for (size_t i = 0; i < n; i++)
        if (offsets[i + 1] < offsets[i]) KG_FAIL(KG_EINVAL, "kg_batch_upload: offsets must be non-decreasing (at %zu)", i);
    const uint64_t total = offsets[n];
    if (total && !seq_bytes) KG_FAIL(KG_EINVAL, "kg_batch_upload: null sequence bytes");
    if (total > KG_MAX_STREAM) KG_FAIL(KG_ERANGE, "kg_batch_upload: %llu bytes in one batch (limit %llu)", (unsigned long long)total, (unsigned long long)KG_MAX_STREAM);
    CU(cudaSetDevice(ctx->device));
    kg_batch* b = new kg_batch();
    b->ctx = ctx;
    b->mode = mode;
    b->n = n;
    b->total = total;
    cudaStream_t st = ctx->stream;
    int rc = KG_OK;
    do {
        if (pool_take_dev(ctx, total + 64, &b->seq_buf) != KG_OK || pool_take_dev(ctx, (n + 1) * 8, &b->off_buf) != KG_OK) {
            rc = KG_ENOMEM;
            break;
        }
        b->d_seq = b->seq_buf.as<uint8_t>();
        b->d_off = b->off_buf.as<uint64_t>();
        cudaMemsetAsync(b->d_seq + total, 0, 64, st);
        if (total) cudaMemcpyAsync(b->d_seq, seq_bytes, total, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(b->d_off, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            kg_set_error("kg_batch_upload: copy failed: %s", cudaGetErrorString(e));
            rc = KG_ECUDA;
        }
    } while (0);
    if (rc != KG_OK) {
        kg_batch_free(b);
        return rc;
    }
    *out = b;
    return KG_OK;
}

extern "C" int kg_batch_from_device(kg_context* ctx, int mode, uint8_t* d_seq_bytes, const uint64_t* d_offsets, size_t n,
                                    uint64_t total_bytes, kg_batch** out) {
    if (!ctx || !out || !d_offsets || (mode != KG_MODE_AA && mode != KG_MODE_DNA)) KG_FAIL(KG_EINVAL, "kg_batch_from_device: bad argument");
    if (((uintptr_t)d_seq_bytes & 15) != 0) KG_FAIL(KG_EINVAL, "kg_batch_from_device: sequence bytes must be 16-byte aligned");
    if (total_bytes > KG_MAX_STREAM) KG_FAIL(KG_ERANGE, "kg_batch_from_device: %llu bytes in one batch", (unsigned long long)total_bytes);
    kg_batch* b = new kg_batch();
    b->ctx = ctx;
    b->mode = mode;
    b->n = n;
    b->total = total_bytes;
    b->d_seq = d_seq_bytes;
    b->d_off = const_cast<uint64_t*>(d_offsets);
    b->owns_input = false;
    *out = b;
    return KG_OK;
}

extern "C" void kg_batch_free(kg_batch* b) {
    if (!b) return;
    if (b->owns_input) {
        pool_give_dev(b->ctx, &b->seq_buf);
        pool_give_dev(b->ctx, &b->off_buf);
    }
    b->vseq.release();
    b->voff.release();
    delete b;
}

// residue stream + virtual offsets.  aa: patch in place (idempotent).  dna: translate six frames.
int kg_batch_prepare(kg_batch* b, cudaStream_t st, uint32_t* launches) {
    kg_context* ctx = b->ctx;
    if (b->mode == KG_MODE_AA) {
        b->nv = b->n;
        b->vtotal = b->total;
        if (b->n) {
            k_patch_aa<<<blocks_for(b->n, 256), 256, 0, st>>>(b->d_seq, b->d_off, b->n);
            (*launches)++;
        }
        b->prepared = true;
        return KG_OK;
    }
    b->nv = 6 * b->n;
    if (!b->prepared) { // the layout depends only on the lengths: computed once per batch
        KG_TRY(b->voff.ensure((b->nv + 2) * 8 * 2));
        uint64_t* vlen = b->voff.as<uint64_t>() + (b->nv + 2);
        k_vlen<<<blocks_for(b->nv + 1, 256), 256, 0, st>>>(b->d_off, b->n, vlen);
        (*launches)++;
        KG_TRY(exclusive_sum_u64(ctx, vlen, b->voff.as<uint64_t>(), b->nv + 1, st));
        (*launches)++;
        CU(cudaMemcpyAsync(&ctx->h_counters[KG_CTR_VPOS], b->voff.as<uint64_t>() + b->nv, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        b->vtotal = ctx->h_counters[KG_CTR_VPOS];
        if (b->vtotal > KG_MAX_STREAM) KG_FAIL(KG_ERANGE, "dna batch: %llu translated residues in one batch", (unsigned long long)b->vtotal);
        KG_TRY(b->vseq.ensure(b->vtotal + 64));
        b->prepared = true;
    }
    CU(cudaMemsetAsync(b->vseq.p, 0, b->vtotal + 64, st));
    if (b->total) {
        k_translate<<<blocks_for(b->total, 256), 256, 0, st>>>(b->d_seq, b->d_off, b->n, b->total, b->voff.as<uint64_t>(),
                                                              b->vseq.as<uint8_t>());
        (*launches)++;
    }
    return KG_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// the device pipeline
// ---------------------------------------------------------------------------------------------------------------
static int run_pipeline(kg_context* ctx, const kg_table* table, kg_batch* b, const kg_params* prm, kg_result* r,
                        uint64_t hit_cap_hint) {
    cudaStream_t st = ctx->stream;
    RunScratch& sc = scratch_of(ctx);
    uint32_t launches = 0;
    cudaEventRecord(ctx->ev[6], st);
    KG_TRY(kg_batch_prepare(b, st, &launches));
    const uint64_t nv = b->nv, vtotal = b->vtotal;
    const int per_seq = b->mode == KG_MODE_AA ? 1 : 6;
    const uint32_t ntiles = (uint32_t)((vtotal + TILE - 1) >> TILE_SHIFT);
    r->n = b->n;
    r->nv = nv;
    r->mode = b->mode;
    r->params = *prm;

    KG_TRY(sc.ctr.ensure(KG_CTR_COUNT * 8));
    unsigned long long* d_ctr = sc.ctr.as<unsigned long long>();
    KG_TRY(sc.tile_base.ensure(((size_t)ntiles + 1) * 4));
    KG_TRY(sc.tile_cnt.ensure(((size_t)ntiles + 1) * 4));
    KG_TRY(sc.tile_out.ensure(((size_t)ntiles + 1) * 4));
    KG_TRY(sc.lo.ensure((nv + 1) * 4));
    KG_TRY(sc.call_cnt.ensure((nv + 1) * 4));
    KG_TRY(sc.call_off.ensure((nv + 1) * 4));

    uint64_t hit_cap = hit_cap_hint ? hit_cap_hint : std::max<uint64_t>(vtotal / 2, 1u << 16);
    if (hit_cap > vtotal) hit_cap = vtotal;
    uint64_t nhits = 0, nkmers = 0;
    for (int attempt = 0;; attempt++) {
        KG_TRY(sc.chunk.ensure(std::max<uint64_t>(hit_cap, 1) * sizeof(uint2)));
        CU(cudaMemsetAsync(d_ctr, 0, KG_CTR_COUNT * 8, st));
        CU(cudaMemsetAsync(sc.tile_cnt.p, 0, ((size_t)ntiles + 1) * 4, st));
        cudaEventRecord(ctx->ev[7], st);
        if (ntiles) {
            k_probe<<<ntiles, PROBE_BLK, 0, st>>>(b->stream(), (uint32_t)vtotal, table->view(), sc.chunk.as<uint2>(),
                                                  (uint32_t)hit_cap, sc.tile_base.as<uint32_t>(), sc.tile_cnt.as<uint32_t>(), d_ctr);
            launches++;
        }
        cudaEventRecord(ctx->ev[8], st);
        KG_TRY(exclusive_sum_u32(ctx, sc.tile_cnt.as<uint32_t>(), sc.tile_out.as<uint32_t>(), (size_t)ntiles + 1, st));
        launches++;
        CU(cudaMemcpyAsync(ctx->h_counters, d_ctr, KG_CTR_COUNT * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        nhits = ctx->h_counters[KG_CTR_HITS];
        nkmers = ctx->h_counters[KG_CTR_KMERS];
        if (!ctx->h_counters[KG_CTR_OVERFLOW]) break;
        if (attempt) KG_FAIL(KG_ECUDA, "hit buffer overflow persisted after resizing to %llu", (unsigned long long)hit_cap);
        hit_cap = nhits; // exact
    }
    if (nhits > 0xFFFFFFF0ull) KG_FAIL(KG_ERANGE, "%llu hits in one batch", (unsigned long long)nhits);

    KG_TRY(sc.hit_pos.ensure(std::max<uint64_t>(nhits, 1) * 4));
    KG_TRY(sc.hit_payload.ensure(std::max<uint64_t>(nhits, 1) * sizeof(int4)));
    if (ntiles && nhits) {
        k_gather<<<blocks_for((size_t)ntiles * 32, 256), 256, 0, st>>>(sc.chunk.as<uint2>(), sc.tile_base.as<uint32_t>(),
                                                                      sc.tile_out.as<uint32_t>(), ntiles, table->d_payload,
                                                                      sc.hit_pos.as<uint32_t>(), sc.hit_payload.as<int4>());
        launches++;
    }
    k_ranges<<<blocks_for(nv + 1, 256), 256, 0, st>>>(b->voffsets(), nv, sc.tile_out.as<uint32_t>(), ntiles,
                                                     sc.hit_pos.as<uint32_t>(), sc.lo.as<uint32_t>());
    launches++;

    const uint64_t max_calls = nhits / (uint64_t)prm->min_hits + 1;
    KG_TRY(sc.sparse.ensure(max_calls * sizeof(KgDevCall)));
    KG_TRY(pool_take_dev(ctx, std::max<uint64_t>(b->n, 1) * sizeof(kg_otu), &r->d_otus));
    KG_TRY(pool_take_dev(ctx, max_calls * sizeof(kg_call), &r->d_calls));
    KgFsmParams fp = {prm->min_hits, prm->max_gap, prm->order_constraint, (float)prm->min_weighted_hits};
    if (b->n) {
        k_fsm<<<blocks_for(b->n, 128), 128, 0, st>>>(b->voffsets(), b->n, per_seq, sc.lo.as<uint32_t>(), sc.hit_pos.as<uint32_t>(),
                                                    sc.hit_payload.as<int4>(), fp, sc.sparse.as<KgDevCall>(),
                                                    sc.call_cnt.as<uint32_t>(), r->d_otus.as<kg_otu>());
        launches++;
    }
    CU(cudaMemsetAsync(sc.call_cnt.as<uint32_t>() + nv, 0, 4, st));
    KG_TRY(exclusive_sum_u32(ctx, sc.call_cnt.as<uint32_t>(), sc.call_off.as<uint32_t>(), nv + 1, st));
    launches++;
    if (nv) {
        k_compact_calls<<<blocks_for(nv, 256), 256, 0, st>>>(sc.sparse.as<KgDevCall>(), sc.lo.as<uint32_t>(),
                                                            sc.call_off.as<uint32_t>(), nv, per_seq, prm->min_hits,
                                                            r->d_calls.as<kg_call>());
        launches++;
    }
    if (prm->emit_hits && nhits) {
        KG_TRY(pool_take_dev(ctx, nhits * sizeof(kg_hit), &r->d_hits));
        k_emit_hits<<<blocks_for(nhits, 256), 256, 0, st>>>(b->voffsets(), nv, per_seq, sc.hit_pos.as<uint32_t>(),
                                                           sc.hit_payload.as<int4>(), (uint32_t)nhits, r->d_hits.as<kg_hit>());
        launches++;
    }
    uint32_t ncalls32 = 0;
    CU(cudaMemcpyAsync(&ctx->h_counters[KG_CTR_CALLS], sc.call_off.as<uint32_t>() + nv, 4, cudaMemcpyDeviceToHost, st));
    cudaEventRecord(ctx->ev[9], st);
    CU(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&r->stats.ms_prepare, ctx->ev[6], ctx->ev[7]);
    cudaEventElapsedTime(&r->stats.ms_probe, ctx->ev[7], ctx->ev[8]);
    cudaEventElapsedTime(&r->stats.ms_group, ctx->ev[8], ctx->ev[9]);
    CU(cudaGetLastError());
    ncalls32 = *(uint32_t*)&ctx->h_counters[KG_CTR_CALLS];
    r->stats.num_sequences = b->n;
    r->stats.num_positions = vtotal;
    r->stats.num_kmers = nkmers;
    r->stats.num_hits = nhits;
    r->stats.num_calls = ncalls32;
    r->stats.num_launches = launches;
    return KG_OK;
}

static int check_params(const kg_params* p) {
    if (!p) KG_FAIL(KG_EINVAL, "null params");
    if (p->min_hits < 2) // KGJ:442 reads hits[n-2]; with min_hits <= 1 the reference dies with ArrayIndexOutOfBounds
        KG_FAIL(KG_EINVAL, "min_hits = %d: the reference requires >= 2 (KGJ:442 indexes hits[n-2])", p->min_hits);
    if (p->max_gap < 0) KG_FAIL(KG_EINVAL, "max_gap = %d must be >= 0", p->max_gap);
    return KG_OK;
}

extern "C" int kg_batch_run(kg_context* ctx, const kg_table* table, kg_batch* batch, const kg_params* params,
                            kg_result** out) {
    if (!ctx || !table || !batch || !out) KG_FAIL(KG_EINVAL, "kg_batch_run: null argument");
    KG_TRY(check_params(params));
    CU(cudaSetDevice(ctx->device));
    kg_result* r = new kg_result();
    r->ctx = ctx;
    cudaEventRecord(ctx->ev[0], ctx->stream);
    int rc = run_pipeline(ctx, table, batch, params, r, 0);
    if (rc != KG_OK) {
        kg_result_free(r);
        return rc;
    }
    cudaEventRecord(ctx->ev[1], ctx->stream);
    cudaEventSynchronize(ctx->ev[1]);
    cudaEventElapsedTime(&r->stats.ms_device, ctx->ev[0], ctx->ev[1]);
    *out = r;
    return KG_OK;
}

extern "C" int kg_result_fetch(kg_result* r) {
    if (!r) KG_FAIL(KG_EINVAL, "kg_result_fetch: null result");
    if (r->fetched) return KG_OK;
    kg_context* ctx = r->ctx;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    cudaEventRecord(ctx->ev[2], st);
    KG_TRY(pool_take_host(ctx, r->stats.num_calls * sizeof(kg_call), &r->h_calls));
    KG_TRY(pool_take_host(ctx, r->n * sizeof(kg_otu), &r->h_otus));
    if (r->stats.num_calls) CU(cudaMemcpyAsync(r->h_calls.p, r->d_calls.p, r->stats.num_calls * sizeof(kg_call), cudaMemcpyDeviceToHost, st));
    if (r->n) CU(cudaMemcpyAsync(r->h_otus.p, r->d_otus.p, r->n * sizeof(kg_otu), cudaMemcpyDeviceToHost, st));
    if (r->params.emit_hits) {
        KG_TRY(pool_take_host(ctx, r->stats.num_hits * sizeof(kg_hit), &r->h_hits));
        if (r->stats.num_hits) CU(cudaMemcpyAsync(r->h_hits.p, r->d_hits.p, r->stats.num_hits * sizeof(kg_hit), cudaMemcpyDeviceToHost, st));
    }
    cudaEventRecord(ctx->ev[3], st);
    CU(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&r->stats.ms_d2h, ctx->ev[2], ctx->ev[3]);
    r->fetched = true;
    return KG_OK;
}

extern "C" int kg_run(kg_context* ctx, const kg_table* table, int mode, const uint8_t* seq_bytes, const uint64_t* offsets,
                      size_t n, const kg_params* params, kg_result** out) {
    if (!ctx || !table || !out) KG_FAIL(KG_EINVAL, "kg_run: null argument");
    KG_TRY(check_params(params));
    kg_batch* b = nullptr;
    cudaEventRecord(ctx->ev[4], ctx->stream);
    KG_TRY(kg_batch_upload(ctx, mode, seq_bytes, offsets, n, &b));
    cudaEventRecord(ctx->ev[5], ctx->stream);
    kg_result* r = nullptr;
    int rc = kg_batch_run(ctx, table, b, params, &r);
    kg_batch_free(b);
    if (rc != KG_OK) return rc;
    cudaEventElapsedTime(&r->stats.ms_h2d, ctx->ev[4], ctx->ev[5]);
    rc = kg_result_fetch(r);
    if (rc != KG_OK) {
        kg_result_free(r);
        return rc;
    }
    *out = r;
    return KG_OK;
}

extern "C" int kg_result_stats(const kg_result* r, kg_run_stats* s) {
    if (!r || !s) KG_FAIL(KG_EINVAL, "kg_result_stats: null argument");
    *s = r->stats;
    return KG_OK;
}
extern "C" int kg_result_calls(kg_result* r, const kg_call** calls, size_t* n) {
    if (!r || !calls || !n) KG_FAIL(KG_EINVAL, "kg_result_calls: null argument");
    KG_TRY(kg_result_fetch(r));
    *calls = (const kg_call*)r->h_calls.p;
    *n = r->stats.num_calls;
    return KG_OK;
}
extern "C" int kg_result_otus(kg_result* r, const kg_otu** otus, size_t* n) {
    if (!r || !otus || !n) KG_FAIL(KG_EINVAL, "kg_result_otus: null argument");
    KG_TRY(kg_result_fetch(r));
    *otus = (const kg_otu*)r->h_otus.p;
    *n = r->n;
    return KG_OK;
}
extern "C" int kg_result_hits(kg_result* r, const kg_hit** hits, size_t* n) {
    if (!r || !hits || !n) KG_FAIL(KG_EINVAL, "kg_result_hits: null argument");
    if (!r->params.emit_hits) KG_FAIL(KG_EINVAL, "kg_result_hits: run was made without params.emit_hits");
    KG_TRY(kg_result_fetch(r));
    *hits = (const kg_hit*)r->h_hits.p;
    *n = r->stats.num_hits;
    return KG_OK;
}
extern "C" void kg_result_free(kg_result* r) {
    if (!r) return;
    // buffers go back to the context's pools (results must be freed before kg_shutdown)
    pool_give_dev(r->ctx, &r->d_calls);
    pool_give_dev(r->ctx, &r->d_otus);
    pool_give_dev(r->ctx, &r->d_hits);
    pool_give_host(r->ctx, &r->h_calls);
    pool_give_host(r->ctx, &r->h_otus);
    pool_give_host(r->ctx, &r->h_hits);
    delete r;
}
