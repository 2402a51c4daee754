// kg_shard.cuh -- the hash-sharded signature table (include/kmerguts_shard.h; BASELINE.json configs[4]).
// Included at the end of kg_run.cu: it reuses that file's window encoder, prefilter, probe loop and the whole pipeline
// downstream of the probe.
//
//   k_route              addKmers (KGJ:900-922) + "which GPU holds this key": valid windows binned by kg_owner_of()
//   keys exchange        NCCL send/recv (one process per GPU) or peer copies (all ranks in one process)
//   k_answer             lookup (KGJ:944-1034) of the received keys against this rank's shard -> replies for the hits
//   replies exchange     {index of the query in the asker's bin, payload}
//   k_mark_replies, k_word_popc + scan, k_place_replies, k_tile_meta
//                        replies -> the position-ordered hit list and per-tile view k_probe would have written; from
//                        here on the pipeline is unchanged
#include <dlfcn.h>
#include <stdlib.h>
#include <nccl.h>

#include <string>
#include <vector>

#include "../../include/kmerguts_shard.h"

namespace {

// ---- NCCL, loaded on first use: a single-GPU deployment needs no libnccl ----
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi& nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    api.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!api.h) api.h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!api.h) return api;
    bool all = true;
    auto sym = [&](const char* name) {
        void* p = dlsym(api.h, name);
        if (!p) all = false;
        return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.ok = all;
    return api;
}
#define NC(call)                                                                                                      \
    do {                                                                                                              \
        ncclResult_t e_ = (call);                                                                                     \
        if (e_ != ncclSuccess) KG_FAIL(KG_ECUDA, "%s failed: %s", #call, nccl_api().GetErrorString(e_));              \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------------
constexpr int SHARD_CNT_SLOTS = 2 * KG_MAX_RANKS; // [0, R): keys per owner; [KG_MAX_RANKS]: valid windows of the batch

// One tile of TILE positions per block, PT per thread, like k_probe.  The valid windows go to the bin of their owner:
// send_lo / send_hi[owner * cap + i] = the key (35 bits: low word + 3 high bits in a byte, 5 bytes per lookup on the
// wire), send_pos[owner * cap + i] = residue position (stays here; the reply names i).
// Slots inside a bin are claimed per tile (one global atomic per owner and tile), warp-aggregated inside the tile.  The
// tile's entries are first sorted by owner in shared memory and then copied out run by run, so that every bin receives
// one contiguous, coalesced burst per tile whatever the number of owners (scattered 4-byte stores made the kernel
// 1.9x slower with eight bins than with two).
// Where an owner's bin lives: this rank's own staging buffer (NCCL / copy transports move it afterwards) or -- the direct
// transport -- this rank's region of the OWNER's receive buffer, mapped into this process (peer memory over NVLink): the
// coalesced burst k_route writes per tile and owner then IS the exchange, and no staging copy or NCCL kernel follows.
struct RouteDst {
    uint32_t* lo[KG_MAX_RANKS];
    uint8_t* hi[KG_MAX_RANKS];
};
__global__ __launch_bounds__(PROBE_BLK) void k_route(const uint8_t* __restrict__ stream, uint32_t vtotal, uint32_t tile0, uint32_t nranks, unsigned long long cap,
                                                     RouteDst dst, uint32_t* __restrict__ send_pos,
                                                     unsigned long long* __restrict__ send_cnt) {
    __shared__ uint8_t lut[256];
    __shared__ uint32_t cnt[KG_MAX_RANKS];
    __shared__ uint32_t first[KG_MAX_RANKS + 1];      // start of every owner's run inside the staged tile
    __shared__ unsigned long long base[KG_MAX_RANKS]; // ... and inside its bin
    __shared__ uint32_t warp_kmers[PROBE_BLK / 32];
    __shared__ uint32_t st_lo[TILE], st_pos[TILE];
    __shared__ uint8_t st_hi[TILE];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 256; i += PROBE_BLK) lut[i] = (i >= 'A' && i <= 'Z') ? c_aa_code[i - 'A'] : 20;
    if (tid < KG_MAX_RANKS) cnt[tid] = 0;
    __syncthreads();
    const uint32_t p0 = (blockIdx.x + tile0) * (uint32_t)TILE + (uint32_t)tid * PT; // this launch routes tiles tile0, tile0 + 1, ...
    uint32_t q[PT + 4];
    uint32_t valid = 0;
    if (p0 < vtotal) valid = encode_windows(stream, p0, vtotal, lut, q);
    uint32_t own[PT], lr[PT];
#pragma unroll
    for (int i = 0; i < PT; i++) {
        const bool ok = (valid >> i) & 1u;
        own[i] = ok ? kg_owner_of((uint64_t)q[i] * 160000ull + q[i + 4], nranks) : nranks; // nranks = "nobody"
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, own[i]);
        const int leader = __ffs(peers) - 1;
        uint32_t b = 0;
        if (lane == leader && ok) b = atomicAdd(&cnt[own[i]], (uint32_t)__popc(peers));
        b = __shfl_sync(0xFFFFFFFFu, b, leader);
        lr[i] = b + __popc(peers & ((1u << lane) - 1u));
    }
    const uint32_t wk = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(valid));
    if (lane == 0) warp_kmers[tid >> 5] = wk;
    __syncthreads();
    if (tid < (int)nranks) base[tid] = cnt[tid] ? atomicAdd(&send_cnt[tid], (unsigned long long)cnt[tid]) : 0ull;
    if (tid == PROBE_BLK - 1) {
        uint32_t kmers = 0, acc = 0;
#pragma unroll
        for (int w = 0; w < PROBE_BLK / 32; w++) kmers += warp_kmers[w];
        if (kmers) atomicAdd(&send_cnt[KG_MAX_RANKS], (unsigned long long)kmers);
        for (uint32_t o = 0; o < nranks; o++) {
            first[o] = acc;
            acc += cnt[o];
        }
        first[nranks] = acc;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PT; i++) {
        if (!((valid >> i) & 1u)) continue;
        const uint64_t key = (uint64_t)q[i] * 160000ull + q[i + 4];
        const uint32_t at = first[own[i]] + lr[i];
        st_lo[at] = (uint32_t)key;
        st_hi[at] = (uint8_t)(key >> 32);
        st_pos[at] = p0 + (uint32_t)i;
    }
    __syncthreads();
    const uint32_t total = first[nranks];
    for (uint32_t e = tid; e < total; e += PROBE_BLK) {
        uint32_t o = 0;
        while (e >= first[o + 1]) o++;
        const unsigned long long slot = base[o] + (e - first[o]);
        if (slot < cap) { // a bin that overflows is only counted: the host repeats the pass with the exact capacity
            __stcs(dst.lo[o] + slot, st_lo[e]); // streamed: keep the L2 for the prefilter of the answer kernel
            __stcs(dst.hi[o] + slot, st_hi[e]);
            __stcs(send_pos + o * cap + slot, st_pos[e]);
        }
    }
}

// The keys received from rank s are segment s of recv_lo / recv_hi; its tiles are blocks [tile_first[s], tile_first[s + 1]).
struct AnswerPlan {
    uint32_t nseg;
    uint32_t tile_first[KG_MAX_RANKS + 1];
    unsigned long long seg_off[KG_MAX_RANKS + 1];
    // direct transport: the segments are fixed-capacity regions (seg_off[s] = s * cap) and how many keys rank s really sent
    // is only known on the device (it arrived with the keys); blocks past the end of a segment find nothing to do
    const unsigned long long* dev_cnt;
    // direct transport: block b works on segment (b + rot) % nseg, tile b / nseg.  Replies are stored straight into the
    // asker's memory, so every GPU must be writing to ALL its peers all the time: with the segments one after the other, the
    // eight GPUs of a box all answered rank 0 first, then rank 1, ... and the step waited 48 ms on one ingress link at a time
    // (r02, N = 8).  rot = this rank, so that GPU r also starts on a different peer than GPU r + 1.
    uint32_t interleave, rot;
};
// where the replies for asker s go: this rank's reply buffer (moved afterwards) or its region of the ASKER's buffer (peer memory)
struct ReplyDst {
    uint32_t* idx[KG_MAX_RANKS];
    int4* payload[KG_MAX_RANKS];
};

// k_probe with the encoder replaced by "read the key": prefilter -> survivor queue -> dense probing -> hits out.  A hit
// becomes a reply {index of the query inside its segment, payload} appended to the segment's reply region (which
// starts at the segment's own offset: there are never more replies than queries).
__global__ __launch_bounds__(PROBE_BLK, KG_PROBE_OCC) void k_answer(const uint32_t* __restrict__ recv_lo, const uint8_t* __restrict__ recv_hi, AnswerPlan plan, KgTableView tab,
                                                                    ReplyDst dst, unsigned long long* __restrict__ reply_cnt, uint32_t flags) {
    extern __shared__ int4 smem_dyn[];
    int4* pstage = smem_dyn;
    unsigned long long* queue = reinterpret_cast<unsigned long long*>(smem_dyn + TILE);
    __shared__ uint32_t hitbits[TILE / 32];
    __shared__ uint32_t warp_a[PROBE_BLK / 32], warp_b[PROBE_BLK / 32];
    __shared__ unsigned long long s_base;
    const int tid = threadIdx.x;
    uint32_t s = 0, tile;
    if (plan.interleave) {
        s = (blockIdx.x % plan.nseg + plan.rot) % plan.nseg;
        tile = blockIdx.x / plan.nseg;
    } else {
        while (s + 1 < plan.nseg && blockIdx.x >= plan.tile_first[s + 1]) s++;
        tile = blockIdx.x - plan.tile_first[s];
    }
    unsigned long long n_s = plan.seg_off[s + 1] - plan.seg_off[s];
    if (plan.dev_cnt) n_s = min(n_s, plan.dev_cnt[s]);
    if ((unsigned long long)tile * TILE >= n_s) return; // (uniform per block, before any barrier)
    if (tid < TILE / 32) hitbits[tid] = 0;
    __syncthreads();
    const uint32_t* klo = recv_lo + plan.seg_off[s];
    const uint8_t* khi = recv_hi + plan.seg_off[s];
    const unsigned long long k0 = (unsigned long long)tile * TILE + (unsigned long long)tid * PT;
    const uint64_t pol_keep = kg_policy_evict_last();
    const uint64_t pol_stream = (flags & 1u) ? kg_policy_evict_normal() : kg_policy_evict_first();

    uint64_t key[PT];
    uint32_t valid = 0;
#pragma unroll
    for (int i = 0; i < PT; i++) {
        key[i] = 0;
        if (k0 + i < n_s) {
            key[i] = (uint64_t)__ldcs(klo + k0 + i) | ((uint64_t)__ldcs(khi + k0 + i) << 32); // read once: do not displace the prefilter
            valid |= 1u << i;
        }
    }
    uint32_t pass = valid;
    if (tab.filter_words) {
        unsigned long long fw[PT];
#pragma unroll
        for (int i = 0; i < PT; i++) {
            fw[i] = 0;
            if ((valid >> i) & 1u) fw[i] = kg_load_filter_word(tab.filter, kg_filter_word(kg_fhash1(key[i]), tab.filter_words), pol_keep);
        }
        pass = 0;
#pragma unroll
        for (int i = 0; i < PT; i++) {
            const unsigned long long fm = kg_filter_mask(kg_fhash1(key[i]));
            pass |= (uint32_t)((fw[i] & fm) == fm) << i;
        }
        pass &= valid;
    }
    uint32_t nsurv;
    uint32_t qo = block_excl_scan(__popc(pass), warp_a, &nsurv);
#pragma unroll
    for (int i = 0; i < PT; i++)
        if ((pass >> i) & 1u) queue[qo++] = key[i] | ((unsigned long long)(tid * PT + i) << 35);
    __syncthreads();
    probe_queue(tab, queue, nsurv, pstage, hitbits, pol_stream);
    __syncthreads();
    const uint32_t hitmask = (hitbits[(tid * PT) >> 5] >> ((tid * PT) & 31)) & ((1u << PT) - 1u);
    uint32_t total;
    const uint32_t ho = block_excl_scan(__popc(hitmask), warp_b, &total);
    if (tid == 0) s_base = total ? atomicAdd(&reply_cnt[s], (unsigned long long)total) : 0ull;
    __syncthreads();
    // The block's replies leave as ONE contiguous run, written by consecutive threads: the destination may be the asker's
    // memory on the other side of NVLink, where a warp store that scatters 32 separate 16-byte pieces (thread t writing its
    // own hits one after the other) becomes 32 small packets -- k_answer took 16.6 instead of ~9 ms at N = 8 (r02).  The
    // hit positions are listed in block order in the (now idle) survivor queue.
    uint32_t* list = reinterpret_cast<uint32_t*>(queue);
    if (hitmask) {
        uint32_t at = ho, m = hitmask;
        while (m) {
            const int i = __ffs(m) - 1;
            m &= m - 1;
            list[at++] = (uint32_t)(tid * PT + i);
        }
    }
    __syncthreads();
    {
        uint32_t* ridx = dst.idx[s] + s_base;
        int4* rpay = dst.payload[s] + s_base;
        const unsigned long long t0 = (unsigned long long)tile * TILE;
        for (uint32_t e = tid; e < total; e += PROBE_BLK) {
            const uint32_t p = list[e];
            __stcs(ridx + e, (uint32_t)(t0 + p));
            __stcs(rpay + e, pstage[p]);
        }
    }
}

// replies of owner o: entries [o * cap, o * cap + (first[o + 1] - first[o])) of rr_idx / rr_payload.  Every owner's
// replies arrive roughly in position order and cover the whole batch, so the kernels below walk all owners' lists at
// the same pace (block b works on owner b % nseg): neighbours in the output are then written at about the same time and
// their half-sector writes merge in L2 instead of becoming read-modify-writes in DRAM.
struct ScatterPlan {
    uint32_t nseg;
    unsigned long long first[KG_MAX_RANKS + 1];
    const unsigned long long* dev_n; // direct transport: replies of owner o = dev_n[o], known on the device only
};
// Merge, step 1: one bit per residue position that has a hit (the bitmap stays in L2: one bit per position).  Every
// thread takes MERGE_U replies of one owner, a block apart, and issues their loads together: the kernels of the merge are
// chains of dependent random accesses (reply -> bin slot -> position -> bitmap word) and live on memory-level parallelism.
constexpr int MERGE_U = 4;
__global__ __launch_bounds__(256) void k_mark_replies(const uint32_t* __restrict__ rr_idx, ScatterPlan plan, unsigned long long cap,
                                                      const unsigned long long* __restrict__ send_cnt, const uint32_t* __restrict__ send_pos,
                                                      uint32_t* __restrict__ bitmap, unsigned long long* __restrict__ ctr) {
    const uint32_t o = blockIdx.x % plan.nseg; // owners interleaved block by block (see ScatterPlan)
    const unsigned long long n = plan.dev_n ? min(plan.dev_n[o], cap) : plan.first[o + 1] - plan.first[o];
    const unsigned long long j0 = (unsigned long long)(blockIdx.x / plan.nseg) * (256 * MERGE_U) + threadIdx.x;
    const unsigned long long nsent = min(send_cnt[o], cap);
    uint32_t idx[MERGE_U], pos[MERGE_U];
    bool on[MERGE_U];
#pragma unroll
    for (int u = 0; u < MERGE_U; u++) {
        on[u] = j0 + 256ull * u < n;
        idx[u] = on[u] ? __ldcs(rr_idx + o * cap + j0 + 256ull * u) : 0u;
        if (on[u] && idx[u] >= nsent) { // a reply that names no query of ours: the peer's answer is corrupt
            ctr[KG_CTR_OVERFLOW] = 2ull;
            on[u] = false;
        }
    }
#pragma unroll
    for (int u = 0; u < MERGE_U; u++) pos[u] = on[u] ? send_pos[o * cap + idx[u]] : 0u;
#pragma unroll
    for (int u = 0; u < MERGE_U; u++)
        if (on[u]) atomicOr(&bitmap[pos[u] >> 5], 1u << (pos[u] & 31));
}

// step 2: hits per 32-position word (an exclusive scan of these gives every hit its rank in position order)
__global__ void k_word_popc(const uint32_t* __restrict__ bitmap, uint32_t nwords, uint32_t* __restrict__ cnt) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w <= nwords) cnt[w] = w < nwords ? __popc(bitmap[w]) : 0u;
}

// step 3: every reply goes straight to its final place in the position-ordered hit list.  (One reply per thread: four
// per thread as in k_mark_replies measured 4.0 instead of 3.05 ms for 89 M replies.)
constexpr int PLACE_U = 1;
__global__ __launch_bounds__(256) void k_place_replies(const uint32_t* __restrict__ rr_idx, const int4* __restrict__ rr_payload, ScatterPlan plan,
                                                       unsigned long long cap, const unsigned long long* __restrict__ send_cnt,
                                                       const uint32_t* __restrict__ send_pos, const uint32_t* __restrict__ bitmap,
                                                       const uint32_t* __restrict__ word_rank, uint32_t hit_cap, uint32_t* __restrict__ chunk_pos,
                                                       int4* __restrict__ chunk_payload) {
    const uint32_t o = blockIdx.x % plan.nseg;
    const unsigned long long n = plan.dev_n ? min(plan.dev_n[o], cap) : plan.first[o + 1] - plan.first[o];
    const unsigned long long j0 = (unsigned long long)(blockIdx.x / plan.nseg) * (256 * PLACE_U) + threadIdx.x;
    const unsigned long long nsent = min(send_cnt[o], cap);
    uint32_t idx[PLACE_U], pos[PLACE_U], bits[PLACE_U], rank[PLACE_U];
    int4 pay[PLACE_U];
    bool on[PLACE_U];
#pragma unroll
    for (int u = 0; u < PLACE_U; u++) {
        on[u] = j0 + 256ull * u < n;
        idx[u] = on[u] ? rr_idx[o * cap + j0 + 256ull * u] : 0u;
        on[u] = on[u] && idx[u] < nsent;
        if (on[u]) pay[u] = rr_payload[o * cap + j0 + 256ull * u];
    }
#pragma unroll
    for (int u = 0; u < PLACE_U; u++) pos[u] = on[u] ? send_pos[o * cap + idx[u]] : 0u;
#pragma unroll
    for (int u = 0; u < PLACE_U; u++) {
        bits[u] = on[u] ? bitmap[pos[u] >> 5] : 0u;
        rank[u] = on[u] ? word_rank[pos[u] >> 5] : 0u;
    }
#pragma unroll
    for (int u = 0; u < PLACE_U; u++) {
        if (!on[u]) continue;
        const uint32_t slot = rank[u] + __popc(bits[u] & ((1u << (pos[u] & 31)) - 1u));
        if (slot < hit_cap) {
            chunk_pos[slot] = pos[u];
            chunk_payload[slot] = pay[u];
        }
    }
}

// step 4: the per-tile view of that list (what k_probe's phase D publishes), and the run's counters
__global__ void k_tile_meta(const uint32_t* __restrict__ word_rank, uint32_t ntiles, uint32_t hit_cap, uint32_t* __restrict__ tile_base,
                            uint32_t* __restrict__ tile_cnt, unsigned long long* __restrict__ ctr, unsigned long long kmers,
                            const unsigned long long* __restrict__ kmers_dev) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    constexpr uint32_t WPT = TILE / 32;
    const uint32_t b = word_rank[t * WPT], e = word_rank[(t + 1) * WPT];
    tile_base[t] = b;
    tile_cnt[t] = e - b;
    if (t == ntiles - 1) {
        ctr[KG_CTR_HITS] = e;
        ctr[KG_CTR_KMERS] = kmers_dev ? kmers_dev[0] : kmers;
        if (e > hit_cap) ctr[KG_CTR_OVERFLOW] = 1ull;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Direct transport ("the exchange hidden behind probing").  Every rank owns ONE exchange buffer, mapped into
// all its peers (CUDA IPC between processes, plain peer access inside one process):
//
//     PeerCtl                      counters and arrival flags, written by the PEERS
//     recv_lo / recv_hi  [R][cap]  region s: the keys rank s routed to this rank      <- written by s's k_route
//     rr_idx / rr_payload [R][cap] region o: the hits owner o found for this rank     <- written by o's k_answer
//
// so the coalesced burst k_route writes per tile and owner, and the run of replies a k_answer block appends, ARE the
// exchange: no staging bins, no send/recv kernels competing with the probes for SMs, no per-chunk all-gather of counters.
// The counts travel in-band: after its kernel a rank stores, into every peer's PeerCtl, how much it wrote there, fences,
// and raises that peer's flag to the step's epoch (k_publish); the consumer's stream holds a one-block kernel that waits
// for the R flags (k_wait, bounded by a timeout).  Re-use is safe without a further barrier: a rank routes step n+1 only
// after its merge of step n, i.e. after every owner's "replies of step n complete" flag, which each owner raises after it
// has finished READING that rank's keys; and an owner writes the replies of step n+1 only after the asker's keys of step
// n+1, which the asker sends after its merge of step n has read the replies of step n.
// ---------------------------------------------------------------------------------------------------------------
struct PeerCtl {
    unsigned long long key_cnt[KG_MAX_RANKS];      // [s] keys rank s sent me in the current step
    unsigned long long max_bin[KG_MAX_RANKS];      // [s] the largest bin rank s filled (for anybody): > cap means everybody repeats the step
    unsigned long long reply_cnt[KG_MAX_RANKS];    // [o] replies owner o sent me
    unsigned long long flag_keys[KG_MAX_RANKS];    // [s] epoch of the last step whose keys from s are complete
    unsigned long long flag_replies[KG_MAX_RANKS]; // [o] likewise for the replies of owner o
    unsigned long long timed_out;                  // set by this rank's own k_wait
};
struct PeerBases {
    uint8_t* base[KG_MAX_RANKS];
};
struct PeerLayout { // byte offsets inside an exchange buffer
    uint64_t cap = 0, recv_lo = 0, recv_hi = 0, rr_idx = 0, rr_payload = 0, total = 0;
    void set(uint64_t R, uint64_t c) {
        auto up = [](uint64_t x) { return (x + 4095) & ~4095ull; };
        cap = c;
        recv_lo = 4096;
        recv_hi = up(recv_lo + R * c * 4);
        rr_idx = up(recv_hi + R * c);
        rr_payload = up(rr_idx + R * c * 4);
        total = up(rr_payload + R * c * 16);
    }
};
// phase 0: keys (my_cnt = send_cnt), phase 1: replies (my_cnt = reply_cnt).  One thread per peer.
__global__ void k_publish(PeerBases pb, int me, int R, int phase, const unsigned long long* __restrict__ my_cnt, unsigned long long epoch) {
    const int p = threadIdx.x;
    if (p >= R) return;
    PeerCtl* pc = reinterpret_cast<PeerCtl*>(pb.base[p]);
    __threadfence_system(); // the data stores of the kernel before this one are ordered before what follows
    if (phase == 0) {
        unsigned long long mx = 0;
        for (int o = 0; o < R; o++) mx = max(mx, my_cnt[o]);
        pc->key_cnt[me] = my_cnt[p];
        pc->max_bin[me] = mx;
    } else {
        pc->reply_cnt[me] = my_cnt[p];
    }
    __threadfence_system();
    volatile unsigned long long* flag = phase == 0 ? &pc->flag_keys[me] : &pc->flag_replies[me];
    *flag = epoch;
}
__global__ void k_wait(PeerCtl* mine, int R, int phase, unsigned long long epoch, unsigned long long timeout_ns) {
    const int p = threadIdx.x;
    if (p >= R) return;
    volatile unsigned long long* flag = phase == 0 ? &mine->flag_keys[p] : &mine->flag_replies[p];
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*flag < epoch) {
        __nanosleep(200);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) { // a peer died or never called: report instead of hanging the GPU
            mine->timed_out = 1;
            break;
        }
    }
    __threadfence_system();
}

} // namespace

// ---------------------------------------------------------------------------------------------------------------
// communicator
// ---------------------------------------------------------------------------------------------------------------
struct KgLocalGroup {
    std::vector<kg_comm*> members;
    int alive = 0;
};

// One chunk of a step: a contiguous range of the batch's tiles with its own bins, receive / reply buffers and counters.
// A step over NCCL is cut into several chunks so that the exchange of one chunk overlaps the kernels of the others; the
// other transports use a single chunk.
constexpr int KG_MAX_CHUNKS = 4;
struct ShardChunk {
    DevBuf send_lo, send_hi, send_pos, send_cnt, recv_lo, recv_hi, reply_idx, reply_payload, reply_cnt, rr_idx, rr_payload, matrix;
    uint64_t* h = nullptr; // pinned: [0, 32) route counters, [32, 64) reply counters, [64, 64 + 32 * 16) gathered counters
    cudaEvent_t ev_route = nullptr, ev_keys = nullptr, ev_answer = nullptr, ev_replies = nullptr;
    uint32_t tile0 = 0, tile1 = 0; // tiles of the batch this chunk routes
    uint64_t cap = 0, cap_seen = 0, kmers = 0;
    uint64_t send_n[KG_MAX_RANKS] = {}, recv_n[KG_MAX_RANKS] = {}, recv_off[KG_MAX_RANKS + 1] = {};
    uint64_t reply_n[KG_MAX_RANKS] = {}, rr_n[KG_MAX_RANKS] = {};
};

struct kg_comm {
    kg_context* ctx = nullptr;
    int rank = 0, nranks = 1;
    ncclComm_t nccl = nullptr;
    KgLocalGroup* group = nullptr;
    cudaStream_t comm_stream = nullptr; // NCCL transport: the exchanges run here, next to the kernels on the compute stream
    ShardChunk ch[KG_MAX_CHUNKS];
    int nchunks = 1;                    // of the run in flight
    DevBuf bitmap, word_cnt, word_rank; // merge
    cudaEvent_t ev_begin = nullptr;
    kg_shard_stats stats = {};
    // direct transport
    DevBuf xbuf;                        // this rank's exchange buffer (PeerLayout)
    PeerLayout xl;
    PeerBases xpeers = {};              // every rank's buffer as this process sees it ([rank] = xbuf.p)
    bool xopened[KG_MAX_RANKS] = {};    // peers mapped with cudaIpcOpenMemHandle (to be closed)
    uint64_t xepoch = 0;
    bool direct_off = false;            // mapping the peers failed somewhere: this communicator uses the staged NCCL transport
    DevBuf xstage;                      // small device staging for the bootstrap all-gathers
    PeerCtl* h_ctl = nullptr;           // pinned copy of the control block, read after a step
};

namespace {

constexpr int CAP_SLOT = KG_MAX_RANKS + 1; // counter block: [0, R) keys per owner, [KG_MAX_RANKS] valid windows, [CAP_SLOT] bin capacity

int comm_alloc(kg_context* ctx, int rank, int nranks, kg_comm** out) {
    CU(cudaSetDevice(ctx->device));
    kg_comm* c = new kg_comm();
    c->ctx = ctx;
    c->rank = rank;
    c->nranks = nranks;
    bool ok = cudaEventCreate(&c->ev_begin) == cudaSuccess;
    for (auto& k : c->ch) {
        ok = ok && cudaMallocHost(&k.h, (64 + SHARD_CNT_SLOTS * KG_MAX_RANKS) * sizeof(uint64_t)) == cudaSuccess;
        for (cudaEvent_t* e : {&k.ev_route, &k.ev_keys, &k.ev_answer, &k.ev_replies}) ok = ok && cudaEventCreate(e) == cudaSuccess;
    }
    if (!ok) {
        cudaGetLastError();
        kg_comm_free(c);
        KG_FAIL(KG_ENOMEM, "kg_comm: events / pinned counters");
    }
    *out = c;
    return KG_OK;
}

// The prefilter's persisting-L2 set-aside (75 MB of the 126 MB) only pays while the answer kernels run.  Everything else
// in a step streams through L2 or keeps its own working set there (the merge: a bitmap and a rank array of the whole
// batch), so the set-aside is switched on for the answer phase only and restored when the call returns (other runs on this
// device expect it).  One rank, 311 M lookups: keys 1.13 -> 0.50 ms, merge 5.6 -> 3.4 ms, step 16.8 -> 13.8 ms.
void shard_l2_setaside(const kg_table* table, bool on) {
    if (on && !table->l2_carve) kg_table_pin_filter(table->ctx, table); // a table built for the cascade: k_answer wants the window
    if (!table->l2_carve || getenv("KG_SHARD_KEEP_CARVE")) return;
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, on ? table->l2_carve : 0);
    cudaGetLastError();
}

// ---- route: encode + bin by owner.  Enqueues only; the bin sizes are still on the device ----
int shard_route(kg_comm* c, ShardChunk& k, kg_batch* b, cudaStream_t st) {
    const uint32_t R = (uint32_t)c->nranks;
    const uint64_t vtotal = b->vtotal;
    const uint32_t ntiles = k.tile1 - k.tile0;
    KG_TRY(k.send_cnt.ensure(SHARD_CNT_SLOTS * 8));
    KG_TRY(k.send_lo.ensure(R * k.cap * 4));
    KG_TRY(k.send_hi.ensure(R * k.cap));
    KG_TRY(k.send_pos.ensure(R * k.cap * 4));
    CU(cudaMemsetAsync(k.send_cnt.p, 0, SHARD_CNT_SLOTS * 8, st));
    k.h[CAP_SLOT] = k.cap; // travels with the counters: every rank can see who overflowed
    CU(cudaMemcpyAsync(k.send_cnt.as<unsigned long long>() + CAP_SLOT, &k.h[CAP_SLOT], 8, cudaMemcpyHostToDevice, st));
    RouteDst dst = {};
    for (uint32_t o = 0; o < R; o++) {
        dst.lo[o] = k.send_lo.as<uint32_t>() + o * k.cap;
        dst.hi[o] = k.send_hi.as<uint8_t>() + o * k.cap;
    }
    if (ntiles)
        k_route<<<ntiles, PROBE_BLK, 0, st>>>(b->stream(), (uint32_t)vtotal, k.tile0, R, k.cap, dst, k.send_pos.as<uint32_t>(),
                                              k.send_cnt.as<unsigned long long>());
    cudaEventRecord(k.ev_route, st);
    return KG_OK;
}

// first guess of a chunk's bin capacity: an even split plus an eighth, or what an earlier run needed
uint64_t shard_cap_guess(const kg_comm* c, const ShardChunk& k) {
    const uint64_t R = (uint64_t)c->nranks, positions = (uint64_t)(k.tile1 - k.tile0) * TILE;
    uint64_t cap = R == 1 ? positions : std::max<uint64_t>(positions / R + positions / (8 * R) + 4096, k.cap_seen);
    return std::max<uint64_t>(std::min<uint64_t>(cap, std::max<uint64_t>(positions, 1)), 1);
}

// host-synchronous transports: bring the bin sizes back, repeat the pass with the exact capacity if a bin overflowed
int shard_route_sync(kg_comm* c, ShardChunk& k, kg_batch* b, cudaStream_t st) {
    const uint32_t R = (uint32_t)c->nranks;
    k.cap = shard_cap_guess(c, k);
    for (int attempt = 0;; attempt++) {
        KG_TRY(shard_route(c, k, b, st));
        CU(cudaMemcpyAsync(k.h, k.send_cnt.p, (KG_MAX_RANKS + 1) * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        CU(cudaGetLastError());
        uint64_t mx = 0;
        for (uint32_t r = 0; r < R; r++) mx = std::max(mx, k.h[r]);
        if (mx <= k.cap) break;
        if (attempt) KG_FAIL(KG_ECUDA, "kg_batch_run_sharded: bin overflow persisted at capacity %llu", (unsigned long long)k.cap);
        k.cap = mx; // a skewed batch (e.g. low-complexity repeats all hashing to one owner)
    }
    return KG_OK;
}

void shard_take_counts(kg_comm* c, ShardChunk& k, const uint64_t* row) {
    k.cap_seen = std::max(k.cap_seen, k.cap);
    k.kmers = row[KG_MAX_RANKS];
    for (int r = 0; r < c->nranks; r++) {
        k.send_n[r] = row[r];
        c->stats.keys_sent += k.send_n[r];
        if (r != c->rank) c->stats.keys_remote += k.send_n[r];
    }
}

// every rank's counter block -> every rank (NCCL transport), on the communication stream; ends host-synchronised
int gather_counters(kg_comm* c, ShardChunk& k, const DevBuf& mine) {
    NcclApi& nc = nccl_api();
    cudaStream_t st = c->comm_stream;
    KG_TRY(k.matrix.ensure((size_t)SHARD_CNT_SLOTS * 8 * c->nranks));
    NC(nc.AllGather(mine.p, k.matrix.p, SHARD_CNT_SLOTS, ncclUint64, c->nccl, st));
    CU(cudaMemcpyAsync(k.h + 64, k.matrix.p, (size_t)SHARD_CNT_SLOTS * 8 * c->nranks, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return KG_OK;
}

// ---- keys: learn what arrives, make room, move the bins.  `st` = the stream the exchange runs on ----
int shard_exchange_keys(kg_comm* c, ShardChunk& k, int ci, kg_batch* b, cudaStream_t st) {
    kg_context* ctx = c->ctx;
    const int R = c->nranks;
    if (c->nccl) {
        // The gathered counters carry every rank's bin capacity, so all ranks agree on whether somebody overflowed; those
        // ranks route the chunk again with the exact capacity and everybody gathers once more.
        for (int attempt = 0;; attempt++) {
            CU(cudaStreamWaitEvent(st, k.ev_route, 0));
            KG_TRY(gather_counters(c, k, k.send_cnt));
            bool any = false, mine = false;
            for (int s = 0; s < R; s++) {
                const uint64_t* row = k.h + 64 + (size_t)s * SHARD_CNT_SLOTS;
                uint64_t mx = 0;
                for (int d = 0; d < R; d++) mx = std::max(mx, row[d]);
                if (mx > row[CAP_SLOT]) {
                    any = true;
                    if (s == c->rank) {
                        mine = true;
                        k.cap = mx;
                    }
                }
            }
            if (!any) break;
            if (attempt) KG_FAIL(KG_ECUDA, "kg_batch_run_sharded: bin overflow persisted");
            if (mine) KG_TRY(shard_route(c, k, b, ctx->stream));
        }
        shard_take_counts(c, k, k.h + 64 + (size_t)c->rank * SHARD_CNT_SLOTS);
        for (int s = 0; s < R; s++) k.recv_n[s] = k.h[64 + (size_t)s * SHARD_CNT_SLOTS + c->rank];
    } else {
        shard_take_counts(c, k, k.h);
        if (c->group) {
            for (int s = 0; s < R; s++) k.recv_n[s] = c->group->members[s]->ch[ci].h[c->rank];
        } else {
            k.recv_n[0] = k.send_n[0];
        }
    }
    k.recv_off[0] = 0;
    for (int s = 0; s < R; s++) k.recv_off[s + 1] = k.recv_off[s] + k.recv_n[s];
    const uint64_t nrecv = k.recv_off[R];
    c->stats.keys_received += nrecv;
    KG_TRY(k.recv_lo.ensure(std::max<uint64_t>(nrecv, 1) * 4));
    KG_TRY(k.recv_hi.ensure(std::max<uint64_t>(nrecv, 1)));
    KG_TRY(k.reply_idx.ensure(std::max<uint64_t>(nrecv, 1) * 4));
    KG_TRY(k.reply_payload.ensure(std::max<uint64_t>(nrecv, 1) * sizeof(int4)));
    KG_TRY(k.reply_cnt.ensure(SHARD_CNT_SLOTS * 8));
    uint32_t *sl = k.send_lo.as<uint32_t>(), *rl = k.recv_lo.as<uint32_t>();
    uint8_t *sh = k.send_hi.as<uint8_t>(), *rh = k.recv_hi.as<uint8_t>();
    if (c->nccl) {
        NcclApi& nc = nccl_api();
        NC(nc.GroupStart());
        const int grc = [&]() -> int { // a failure inside the group must still close it
            for (int p = 0; p < R; p++) {
                if (p == c->rank) continue;
                if (k.send_n[p]) {
                    NC(nc.Send(sl + p * k.cap, k.send_n[p], ncclUint32, p, c->nccl, st));
                    NC(nc.Send(sh + p * k.cap, k.send_n[p], ncclUint8, p, c->nccl, st));
                }
                if (k.recv_n[p]) {
                    NC(nc.Recv(rl + k.recv_off[p], k.recv_n[p], ncclUint32, p, c->nccl, st));
                    NC(nc.Recv(rh + k.recv_off[p], k.recv_n[p], ncclUint8, p, c->nccl, st));
                }
                c->stats.bytes_sent += k.send_n[p] * 5;
            }
            return KG_OK;
        }();
        const ncclResult_t ge = nc.GroupEnd();
        if (grc != KG_OK) return grc;
        NC(ge);
        if (k.send_n[c->rank]) {
            CU(cudaMemcpyAsync(rl + k.recv_off[c->rank], sl + c->rank * k.cap, k.send_n[c->rank] * 4, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(rh + k.recv_off[c->rank], sh + c->rank * k.cap, k.send_n[c->rank], cudaMemcpyDeviceToDevice, st));
        }
    } else if (c->group) { // every member has finished its route phase (host-synchronised): pull the bins
        for (int s = 0; s < R; s++) {
            kg_comm* src = c->group->members[s];
            const ShardChunk& sk = src->ch[ci];
            if (!k.recv_n[s]) continue;
            CU(cudaMemcpyPeerAsync(rl + k.recv_off[s], ctx->device, sk.send_lo.as<uint32_t>() + c->rank * sk.cap, src->ctx->device, k.recv_n[s] * 4, st));
            CU(cudaMemcpyPeerAsync(rh + k.recv_off[s], ctx->device, sk.send_hi.as<uint8_t>() + c->rank * sk.cap, src->ctx->device, k.recv_n[s], st));
            if (s != c->rank) src->stats.bytes_sent += k.recv_n[s] * 5;
        }
    } else if (nrecv) {
        CU(cudaMemcpyAsync(rl, sl, nrecv * 4, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(rh, sh, nrecv, cudaMemcpyDeviceToDevice, st));
    }
    cudaEventRecord(k.ev_keys, st);
    return KG_OK;
}

// ---- answer: probe the received keys (compute stream).  With `sync` the reply counts come back to the host ----
int shard_answer(kg_comm* c, ShardChunk& k, const kg_table* table, bool sync) {
    kg_context* ctx = c->ctx;
    cudaStream_t st = ctx->stream;
    const int R = c->nranks;
    AnswerPlan plan = {};
    plan.nseg = (uint32_t)R;
    uint64_t tiles = 0;
    for (int s = 0; s < R; s++) {
        plan.tile_first[s] = (uint32_t)tiles;
        plan.seg_off[s] = k.recv_off[s];
        tiles += (k.recv_n[s] + TILE - 1) >> TILE_SHIFT;
    }
    plan.tile_first[R] = (uint32_t)tiles;
    plan.seg_off[R] = k.recv_off[R];
    if (tiles > 0x7FFFFFFFull) KG_FAIL(KG_ERANGE, "kg_batch_run_sharded: %llu keys received in one step", (unsigned long long)k.recv_off[R]);
    shard_l2_setaside(table, true);
    CU(cudaStreamWaitEvent(st, k.ev_keys, 0));
    CU(cudaMemsetAsync(k.reply_cnt.p, 0, SHARD_CNT_SLOTS * 8, st));
    ReplyDst rdst = {};
    for (int s = 0; s < R; s++) {
        rdst.idx[s] = k.reply_idx.as<uint32_t>() + k.recv_off[s];
        rdst.payload[s] = k.reply_payload.as<int4>() + k.recv_off[s];
    }
    if (tiles)
        k_answer<<<(unsigned)tiles, PROBE_BLK, PROBE_SMEM, st>>>(k.recv_lo.as<uint32_t>(), k.recv_hi.as<uint8_t>(), plan, table->view(), rdst,
                                                                 k.reply_cnt.as<unsigned long long>(), probe_flags());
    cudaEventRecord(k.ev_answer, st);
    if (sync) {
        CU(cudaMemcpyAsync(k.h + 32, k.reply_cnt.p, SHARD_CNT_SLOTS * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        CU(cudaGetLastError());
    }
    return KG_OK;
}

// ---- replies: back to the ranks that asked ----
int shard_exchange_replies(kg_comm* c, ShardChunk& k, int ci, cudaStream_t st) {
    kg_context* ctx = c->ctx;
    const int R = c->nranks;
    if (c->nccl) {
        CU(cudaStreamWaitEvent(st, k.ev_answer, 0));
        KG_TRY(gather_counters(c, k, k.reply_cnt));
        for (int s = 0; s < R; s++) k.reply_n[s] = k.h[64 + (size_t)c->rank * SHARD_CNT_SLOTS + s];
        for (int o = 0; o < R; o++) k.rr_n[o] = k.h[64 + (size_t)o * SHARD_CNT_SLOTS + c->rank];
    } else {
        for (int s = 0; s < R; s++) k.reply_n[s] = k.h[32 + s];
        if (c->group) {
            for (int o = 0; o < R; o++) k.rr_n[o] = c->group->members[o]->ch[ci].h[32 + c->rank];
        } else {
            k.rr_n[0] = k.reply_n[0];
        }
    }
    for (int s = 0; s < R; s++) {
        if (k.reply_n[s] > k.recv_n[s]) KG_FAIL(KG_ECUDA, "kg_batch_run_sharded: more replies than queries for rank %d", s);
        c->stats.replies_sent += k.reply_n[s];
    }
    for (int o = 0; o < R; o++) {
        if (k.rr_n[o] > k.send_n[o]) KG_FAIL(KG_ECUDA, "kg_batch_run_sharded: rank %d answers %llu of %llu queries", o, (unsigned long long)k.rr_n[o], (unsigned long long)k.send_n[o]);
        c->stats.replies_received += k.rr_n[o];
    }
    KG_TRY(k.rr_idx.ensure((size_t)R * k.cap * 4));
    KG_TRY(k.rr_payload.ensure((size_t)R * k.cap * sizeof(int4)));
    uint32_t* ri = k.rr_idx.as<uint32_t>();
    int4* rp = k.rr_payload.as<int4>();
    if (c->nccl) {
        NcclApi& nc = nccl_api();
        NC(nc.GroupStart());
        const int grc = [&]() -> int {
            for (int p = 0; p < R; p++) {
                if (p == c->rank) continue;
                if (k.reply_n[p]) {
                    NC(nc.Send(k.reply_idx.as<uint32_t>() + k.recv_off[p], k.reply_n[p], ncclUint32, p, c->nccl, st));
                    NC(nc.Send(k.reply_payload.as<int4>() + k.recv_off[p], k.reply_n[p] * sizeof(int4), ncclUint8, p, c->nccl, st));
                    c->stats.bytes_sent += k.reply_n[p] * (4 + sizeof(int4));
                }
                if (k.rr_n[p]) {
                    NC(nc.Recv(ri + p * k.cap, k.rr_n[p], ncclUint32, p, c->nccl, st));
                    NC(nc.Recv(rp + p * k.cap, k.rr_n[p] * sizeof(int4), ncclUint8, p, c->nccl, st));
                }
            }
            return KG_OK;
        }();
        const ncclResult_t ge = nc.GroupEnd();
        if (grc != KG_OK) return grc;
        NC(ge);
        const int me = c->rank;
        if (k.reply_n[me]) {
            CU(cudaMemcpyAsync(ri + me * k.cap, k.reply_idx.as<uint32_t>() + k.recv_off[me], k.reply_n[me] * 4, cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(rp + me * k.cap, k.reply_payload.as<int4>() + k.recv_off[me], k.reply_n[me] * sizeof(int4), cudaMemcpyDeviceToDevice, st));
        }
    } else if (c->group) {
        for (int o = 0; o < R; o++) {
            kg_comm* src = c->group->members[o];
            const ShardChunk& sk = src->ch[ci];
            if (!k.rr_n[o]) continue;
            CU(cudaMemcpyPeerAsync(ri + o * k.cap, ctx->device, sk.reply_idx.as<uint32_t>() + sk.recv_off[c->rank], src->ctx->device, k.rr_n[o] * 4, st));
            CU(cudaMemcpyPeerAsync(rp + o * k.cap, ctx->device, sk.reply_payload.as<int4>() + sk.recv_off[c->rank], src->ctx->device,
                                   k.rr_n[o] * sizeof(int4), st));
            if (o != c->rank) src->stats.bytes_sent += k.rr_n[o] * (4 + sizeof(int4));
        }
    } else if (k.rr_n[0]) {
        CU(cudaMemcpyAsync(ri, k.reply_idx.p, k.rr_n[0] * 4, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(rp, k.reply_payload.p, k.rr_n[0] * sizeof(int4), cudaMemcpyDeviceToDevice, st));
    }
    cudaEventRecord(k.ev_replies, st);
    return KG_OK;
}

// ---- merge: replies of all chunks -> per-tile hit chunks -> the unchanged rest of the pipeline ----
int dx_collect(kg_comm* c, ShardChunk& k);
int shard_merge(kg_comm* c, const kg_table* table, kg_batch* b, const kg_params* prm, kg_result** out, bool direct = false) {
    kg_context* ctx = c->ctx;
    const int R = c->nranks, H = c->nchunks;
    const uint64_t vtotal = b->vtotal;
    const uint32_t ntiles = (uint32_t)((vtotal + TILE - 1) >> TILE_SHIFT);
    ScatterPlan sp[KG_MAX_CHUNKS] = {};
    unsigned grid[KG_MAX_CHUNKS] = {}, grid_place[KG_MAX_CHUNKS] = {};
    uint64_t nhits = 0, kmers = 0;
    for (int h = 0; h < H; h++) {
        const ShardChunk& k = c->ch[h];
        sp[h].nseg = (uint32_t)R;
        uint64_t longest = 0;
        for (int o = 0; o < R; o++) {
            sp[h].first[o + 1] = sp[h].first[o] + k.rr_n[o];
            longest = std::max(longest, k.rr_n[o]);
        }
        if (direct) { // the reply counts are on the device (PeerCtl): walk the regions up to their capacity
            longest = k.cap;
            sp[h].dev_n = reinterpret_cast<PeerCtl*>(c->xbuf.p)->reply_cnt;
        }
        grid[h] = (unsigned)(blocks_for(longest, 256 * MERGE_U) * (uint64_t)R);
        grid_place[h] = (unsigned)(blocks_for(longest, 256 * PLACE_U) * (uint64_t)R);
        nhits += sp[h].first[R];
        kmers += k.kmers;
    }
    const uint32_t* rr_idx_of[KG_MAX_CHUNKS];
    const int4* rr_payload_of[KG_MAX_CHUNKS];
    for (int h = 0; h < H; h++) {
        rr_idx_of[h] = direct ? reinterpret_cast<const uint32_t*>(c->xbuf.as<uint8_t>() + c->xl.rr_idx) : c->ch[h].rr_idx.as<uint32_t>();
        rr_payload_of[h] = direct ? reinterpret_cast<const int4*>(c->xbuf.as<uint8_t>() + c->xl.rr_payload) : c->ch[h].rr_payload.as<int4>();
    }
    const uint32_t nwords = ntiles * (uint32_t)(TILE / 32);
    const size_t bitmap_bytes = ((size_t)nwords + 1) * 4;
    KG_TRY(c->bitmap.ensure(bitmap_bytes));
    KG_TRY(c->word_cnt.ensure(bitmap_bytes));
    KG_TRY(c->word_rank.ensure(bitmap_bytes));
    ProbeStage stage = [&](PipeSlot& sl, unsigned long long* d_ctr, uint64_t hit_cap, cudaStream_t st) -> int {
        if (!ntiles) return KG_OK;
        CU(cudaMemsetAsync(c->bitmap.p, 0, bitmap_bytes, st));
        for (int h = 0; h < H; h++) {
            const ShardChunk& k = c->ch[h];
            CU(cudaStreamWaitEvent(st, k.ev_replies, 0));
            if (grid[h])
                k_mark_replies<<<grid[h], 256, 0, st>>>(rr_idx_of[h], sp[h], k.cap, k.send_cnt.as<unsigned long long>(), k.send_pos.as<uint32_t>(),
                                                        c->bitmap.as<uint32_t>(), d_ctr);
        }
        k_word_popc<<<blocks_for((size_t)nwords + 1, 256), 256, 0, st>>>(c->bitmap.as<uint32_t>(), nwords, c->word_cnt.as<uint32_t>());
        KG_TRY(exclusive_sum_u32(ctx, c->word_cnt.as<uint32_t>(), c->word_rank.as<uint32_t>(), (size_t)nwords + 1, st));
        for (int h = 0; h < H; h++) {
            const ShardChunk& k = c->ch[h];
            if (grid_place[h])
                k_place_replies<<<grid_place[h], 256, 0, st>>>(rr_idx_of[h], rr_payload_of[h], sp[h], k.cap, k.send_cnt.as<unsigned long long>(),
                                                         k.send_pos.as<uint32_t>(), c->bitmap.as<uint32_t>(), c->word_rank.as<uint32_t>(), (uint32_t)hit_cap,
                                                         sl.chunk_pos.as<uint32_t>(), sl.chunk_payload.as<int4>());
        }
        k_tile_meta<<<blocks_for(ntiles, 256), 256, 0, st>>>(c->word_rank.as<uint32_t>(), ntiles, (uint32_t)hit_cap, sl.tile_base.as<uint32_t>(),
                                                           sl.tile_cnt.as<uint32_t>(), d_ctr, kmers,
                                                           direct ? c->ch[0].send_cnt.as<unsigned long long>() + KG_MAX_RANKS : nullptr);
        sl.launches += 3 + 2 * (uint32_t)H;
        return KG_OK;
    };
    // all answer kernels are done (their reply counts are on the host)
    shard_l2_setaside(table, false);
    kg_result* r = new kg_result();
    r->ctx = ctx;
    RunScratch& sc = scratch_of(ctx);
    int rc = pipe_enqueue(ctx, sc.slot[0], table, b, prm, r, direct ? sc.hit_cap_seen : std::max<uint64_t>(nhits, 1), 0, &stage);
    if (rc == KG_OK) rc = pipe_finish(ctx, sc.slot[0], table, b, prm, r, 0, &stage);
    if (rc == KG_OK && direct) { // the counts of the step, now that it is over
        rc = dx_collect(c, c->ch[0]);
        nhits = 0;
        for (int o = 0; o < R; o++) nhits += c->ch[0].rr_n[o];
    }
    if (rc == KG_OK && direct && c->ch[0].cap_seen > c->xl.cap) { // some bin overflowed somewhere: the caller repeats the step
        kg_result_free(r);
        *out = nullptr;
        return KG_OK;
    }
    if (rc == KG_OK && r->stats.num_hits != nhits) {
        // distinct replies always land on distinct positions; fewer bits than replies means a peer answered a query twice
        kg_set_error("kg_batch_run_sharded: %llu replies but %llu hit positions", (unsigned long long)nhits, (unsigned long long)r->stats.num_hits);
        rc = KG_ECUDA;
    }
    if (rc != KG_OK) {
        kg_result_free(r);
        return rc;
    }
    r->stats.num_launches += 2 * (uint32_t)H; // k_route, k_answer per chunk (patch / translate are counted by prepare)
    // phase END times of the last chunk since the start of the call (with several chunks the phases overlap)
    const ShardChunk& last = c->ch[H - 1];
    float t_route = 0, t_keys = 0, t_answer = 0, t_replies = 0;
    cudaEventElapsedTime(&t_route, c->ev_begin, last.ev_route);
    cudaEventElapsedTime(&t_keys, c->ev_begin, last.ev_keys);
    cudaEventElapsedTime(&t_answer, c->ev_begin, last.ev_answer);
    cudaEventElapsedTime(&t_replies, c->ev_begin, last.ev_replies);
    cudaGetLastError();
    c->stats.chunks = H;
    if (H == 1) { // one chunk: the phases run back to back, report their durations
        c->stats.ms_route = t_route;
        c->stats.ms_keys = t_keys - t_route;
        c->stats.ms_answer = t_answer - t_keys;
        c->stats.ms_replies = t_replies - t_answer;
    } else {
        c->stats.ms_route = t_route;
        c->stats.ms_keys = t_keys;
        c->stats.ms_answer = t_answer;
        c->stats.ms_replies = t_replies;
    }
    c->stats.ms_merge = r->stats.ms_device;
    *out = r;
    return KG_OK;
}

// the chunks of a run: contiguous tile ranges of the prepared batch
int shard_begin(kg_comm* c, kg_batch* b, int want_chunks) {
    kg_context* ctx = c->ctx;
    CU(cudaSetDevice(ctx->device));
    c->stats = kg_shard_stats();
    cudaEventRecord(c->ev_begin, ctx->stream);
    uint32_t launches = 0;
    KG_TRY(kg_batch_prepare(b, ctx->stream, &launches));
    const uint32_t ntiles = (uint32_t)((b->vtotal + TILE - 1) >> TILE_SHIFT);
    int H = std::max(1, std::min(want_chunks, KG_MAX_CHUNKS));
    c->nchunks = H;
    for (int h = 0; h < H; h++) {
        c->ch[h].tile0 = (uint32_t)((uint64_t)ntiles * h / H);
        c->ch[h].tile1 = (uint32_t)((uint64_t)ntiles * (h + 1) / H);
    }
    return KG_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// direct transport, host side
// ---------------------------------------------------------------------------------------------------------------
bool shard_direct_wanted() { // KG_SHARD_TRANSPORT=nccl | copy selects the staged transports (bins + send/recv or peer copies)
    const char* e = getenv("KG_SHARD_TRANSPORT");
    return !e || !strcmp(e, "direct");
}
uint64_t dx_cap_wanted(const kg_comm* c, const kg_batch* b) { // an even split plus an eighth
    const uint64_t R = (uint64_t)c->nranks, positions = ((b->vtotal + TILE - 1) >> TILE_SHIFT) * (uint64_t)TILE;
    return R == 1 ? std::max<uint64_t>(positions, 1) : positions / R + positions / (8 * R) + 4096;
}
// bootstrap all-gather of a few bytes per rank over NCCL (host-synchronous; also serves as a barrier)
int allgather_host(kg_comm* c, const void* mine, size_t bytes, void* all) {
    NcclApi& nc = nccl_api();
    KG_TRY(c->xstage.ensure(bytes * ((size_t)c->nranks + 1)));
    uint8_t* d = c->xstage.as<uint8_t>();
    cudaStream_t st = c->comm_stream;
    CU(cudaMemcpyAsync(d, mine, bytes, cudaMemcpyHostToDevice, st));
    NC(nc.AllGather(d, d + bytes, bytes, ncclUint8, c->nccl, st));
    CU(cudaMemcpyAsync(all, d + bytes, bytes * (size_t)c->nranks, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return KG_OK;
}
void dx_close(kg_comm* c) {
    for (int p = 0; p < KG_MAX_RANKS; p++) {
        if (c->xopened[p]) cudaIpcCloseMemHandle(c->xpeers.base[p]);
        c->xopened[p] = false;
        c->xpeers.base[p] = nullptr;
    }
    cudaGetLastError();
}
int dx_alloc(kg_comm* c, uint64_t cap) {
    c->xbuf.release();
    c->xl.set((uint64_t)c->nranks, cap);
    KG_TRY(c->xbuf.ensure(c->xl.total));
    CU(cudaMemset(c->xbuf.p, 0, 4096));
    if (!c->h_ctl) CU(cudaMallocHost(&c->h_ctl, sizeof(PeerCtl)));
    return KG_OK;
}
// Collective (one process per GPU): agree on the capacity, (re)allocate, map everybody's buffer.  Called before the first
// step and after a step in which some bin overflowed -- every rank sees that in its PeerCtl, so they all come here.
int dx_setup(kg_comm* c, uint64_t want_cap) {
    const int R = c->nranks;
    uint64_t caps[KG_MAX_RANKS] = {};
    KG_TRY(allgather_host(c, &want_cap, 8, caps)); // also a barrier: nobody is inside the previous step any more
    uint64_t cap = c->xl.cap;
    for (int p = 0; p < R; p++) cap = std::max(cap, caps[p]);
    if (c->xbuf.p && cap == c->xl.cap) return KG_OK;
    CU(cudaStreamSynchronize(c->ctx->stream));
    dx_close(c);
    uint64_t zero = 0, zeros[KG_MAX_RANKS];
    KG_TRY(allgather_host(c, &zero, 8, zeros)); // every rank has unmapped its peers before anybody frees
    KG_TRY(dx_alloc(c, cap));
    // Mapping a peer's buffer can fail where CUDA IPC or peer access is not available (some container set-ups, GPUs without
    // a P2P path).  No rank may then use the direct transport: the outcome is all-gathered and, if anybody failed, every rank
    // drops its buffer and the communicator falls back to the staged NCCL transport for good.
    cudaIpcMemHandle_t mine, all[KG_MAX_RANKS];
    uint64_t ok = cudaIpcGetMemHandle(&mine, c->xbuf.p) == cudaSuccess && !getenv("KG_SHARD_FAIL_IPC");
    if (!ok) memset(&mine, 0, sizeof mine);
    KG_TRY(allgather_host(c, &mine, sizeof mine, all));
    uint64_t oks[KG_MAX_RANKS] = {};
    KG_TRY(allgather_host(c, &ok, 8, oks));
    for (int p = 0; p < R; p++) ok = ok && oks[p];
    for (int p = 0; p < R && ok; p++) {
        if (p == c->rank) {
            c->xpeers.base[p] = c->xbuf.as<uint8_t>();
        } else {
            void* q = nullptr;
            if (cudaIpcOpenMemHandle(&q, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                ok = 0;
                break;
            }
            c->xpeers.base[p] = (uint8_t*)q;
            c->xopened[p] = true;
        }
    }
    cudaGetLastError();
    KG_TRY(allgather_host(c, &ok, 8, oks)); // also: everybody's control block is zeroed and mapped before anyone writes
    for (int p = 0; p < R; p++) ok = ok && oks[p];
    if (!ok) {
        dx_close(c);
        KG_TRY(allgather_host(c, &zero, 8, zeros)); // everybody has unmapped before anybody frees
        c->xbuf.release();
        c->xl = PeerLayout();
        c->direct_off = true;
    }
    return KG_OK;
}
// keys: encode, bin by owner, store every bin straight into its owner's buffer; tell the owners how much arrived
int dx_route(kg_comm* c, kg_batch* b, bool flags) {
    ShardChunk& k = c->ch[0];
    const uint32_t R = (uint32_t)c->nranks, me = (uint32_t)c->rank;
    cudaStream_t st = c->ctx->stream;
    k.cap = c->xl.cap;
    const uint32_t ntiles = k.tile1 - k.tile0;
    KG_TRY(k.send_cnt.ensure(SHARD_CNT_SLOTS * 8));
    KG_TRY(k.reply_cnt.ensure(SHARD_CNT_SLOTS * 8));
    KG_TRY(k.send_pos.ensure(R * k.cap * 4));
    CU(cudaMemsetAsync(k.send_cnt.p, 0, SHARD_CNT_SLOTS * 8, st));
    RouteDst dst = {};
    for (uint32_t o = 0; o < R; o++) {
        dst.lo[o] = reinterpret_cast<uint32_t*>(c->xpeers.base[o] + c->xl.recv_lo) + (size_t)me * k.cap;
        dst.hi[o] = c->xpeers.base[o] + c->xl.recv_hi + (size_t)me * k.cap;
    }
    if (ntiles)
        k_route<<<ntiles, PROBE_BLK, 0, st>>>(b->stream(), (uint32_t)b->vtotal, k.tile0, R, k.cap, dst, k.send_pos.as<uint32_t>(),
                                              k.send_cnt.as<unsigned long long>());
    cudaEventRecord(k.ev_route, st);
    c->xepoch++;
    k_publish<<<1, 32, 0, st>>>(c->xpeers, (int)me, (int)R, 0, k.send_cnt.as<unsigned long long>(), c->xepoch);
    if (flags) k_wait<<<1, 32, 0, st>>>(reinterpret_cast<PeerCtl*>(c->xbuf.p), (int)R, 0, c->xepoch, 20ull * 1000000000ull);
    cudaEventRecord(k.ev_keys, st);
    CU(cudaGetLastError());
    return KG_OK;
}
// answer: probe what arrived; a hit's reply goes straight into the asker's buffer
int dx_answer(kg_comm* c, const kg_table* table, bool flags) {
    ShardChunk& k = c->ch[0];
    const int R = c->nranks, me = c->rank;
    cudaStream_t st = c->ctx->stream;
    PeerCtl* ctl = reinterpret_cast<PeerCtl*>(c->xbuf.p);
    const uint64_t tps = (k.cap + TILE - 1) >> TILE_SHIFT;
    if (tps * (uint64_t)R > 0x7FFFFFFFull) KG_FAIL(KG_ERANGE, "kg_batch_run_sharded: exchange regions of %llu keys", (unsigned long long)k.cap);
    AnswerPlan plan = {};
    plan.nseg = (uint32_t)R;
    ReplyDst rdst = {};
    for (int s = 0; s <= R; s++) {
        plan.tile_first[s] = (uint32_t)(tps * (uint64_t)s);
        plan.seg_off[s] = k.cap * (uint64_t)s;
    }
    plan.dev_cnt = ctl->key_cnt;
    plan.interleave = 1;
    plan.rot = (uint32_t)me;
    for (int s = 0; s < R; s++) {
        rdst.idx[s] = reinterpret_cast<uint32_t*>(c->xpeers.base[s] + c->xl.rr_idx) + (size_t)me * k.cap;
        rdst.payload[s] = reinterpret_cast<int4*>(c->xpeers.base[s] + c->xl.rr_payload) + (size_t)me * k.cap;
    }
    shard_l2_setaside(table, true);
    CU(cudaMemsetAsync(k.reply_cnt.p, 0, SHARD_CNT_SLOTS * 8, st));
    k_answer<<<(unsigned)(tps * (uint64_t)R), PROBE_BLK, PROBE_SMEM, st>>>(reinterpret_cast<const uint32_t*>(c->xbuf.as<uint8_t>() + c->xl.recv_lo),
                                                                          c->xbuf.as<uint8_t>() + c->xl.recv_hi, plan, table->view(), rdst,
                                                                          k.reply_cnt.as<unsigned long long>(), probe_flags());
    cudaEventRecord(k.ev_answer, st);
    c->xepoch++;
    k_publish<<<1, 32, 0, st>>>(c->xpeers, me, R, 1, k.reply_cnt.as<unsigned long long>(), c->xepoch);
    if (flags) k_wait<<<1, 32, 0, st>>>(ctl, R, 1, c->xepoch, 20ull * 1000000000ull);
    cudaEventRecord(k.ev_replies, st);
    CU(cudaGetLastError());
    return KG_OK;
}
// after the step (the stream is idle): counters -> host, statistics, sanity
int dx_collect(kg_comm* c, ShardChunk& k) {
    const int R = c->nranks;
    cudaStream_t st = c->ctx->stream;
    CU(cudaMemcpyAsync(c->h_ctl, c->xbuf.p, sizeof(PeerCtl), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(k.h, k.send_cnt.p, (KG_MAX_RANKS + 1) * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(k.h + 32, k.reply_cnt.p, KG_MAX_RANKS * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (c->h_ctl->timed_out) KG_FAIL(KG_ECUDA, "kg_batch_run_sharded: a peer did not arrive within 20 s (did every rank call?)");
    uint64_t worst = 0;
    for (int s = 0; s < R; s++) worst = std::max<uint64_t>(worst, c->h_ctl->max_bin[s]);
    k.cap_seen = worst; // > capacity: somebody's bin overflowed -- every rank reads the same numbers and repeats the step
    k.cap = c->xl.cap;
    shard_take_counts(c, k, k.h);
    k.cap_seen = worst;
    for (int s = 0; s < R; s++) {
        k.recv_n[s] = std::min<uint64_t>(c->h_ctl->key_cnt[s], k.cap);
        k.reply_n[s] = k.h[32 + s];
        k.rr_n[s] = std::min<uint64_t>(c->h_ctl->reply_cnt[s], k.cap);
        c->stats.keys_received += k.recv_n[s];
        c->stats.replies_sent += k.reply_n[s];
        c->stats.replies_received += k.rr_n[s];
        if (s != c->rank) c->stats.bytes_sent += k.send_n[s] * 5 + k.reply_n[s] * (4 + sizeof(int4));
    }
    return KG_OK;
}

int shard_check(const kg_comm* c, const kg_table* t, const kg_batch* b, const kg_params* prm) {
    if (!c || !t || !b) KG_FAIL(KG_EINVAL, "kg_batch_run_sharded: null argument");
    KG_TRY(check_params(prm));
    if (t->shard_count != c->nranks || t->shard_rank != c->rank)
        KG_FAIL(KG_EINVAL, "kg_batch_run_sharded: the table is shard %d of %d, the communicator is rank %d of %d", t->shard_rank, t->shard_count,
                c->rank, c->nranks);
    if (b->ctx != c->ctx || t->ctx != c->ctx) KG_FAIL(KG_EINVAL, "kg_batch_run_sharded: table, batch and communicator must share one context");
    return KG_OK;
}

} // namespace

// ---------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------
extern "C" int kg_comm_unique_id(uint8_t id[KG_COMM_ID_BYTES]) {
    if (!id) KG_FAIL(KG_EINVAL, "kg_comm_unique_id: null argument");
    static_assert(sizeof(ncclUniqueId) == KG_COMM_ID_BYTES, "ncclUniqueId size");
    NcclApi& nc = nccl_api();
    if (!nc.ok) KG_FAIL(KG_EIO, "kg_comm_unique_id: libnccl.so.2 not loadable: %s", dlerror());
    ncclUniqueId u;
    NC(nc.GetUniqueId(&u));
    memcpy(id, &u, KG_COMM_ID_BYTES);
    return KG_OK;
}

extern "C" int kg_comm_init(kg_context* ctx, int rank, int nranks, const uint8_t id[KG_COMM_ID_BYTES], kg_comm** comm) {
    if (!ctx || !comm) KG_FAIL(KG_EINVAL, "kg_comm_init: null argument");
    if (nranks < 1 || nranks > KG_MAX_RANKS || rank < 0 || rank >= nranks) KG_FAIL(KG_EINVAL, "kg_comm_init: rank %d of %d (at most %d ranks)", rank, nranks, KG_MAX_RANKS);
    kg_comm* c = nullptr;
    KG_TRY(comm_alloc(ctx, rank, nranks, &c));
    if (nranks > 1) {
        NcclApi& nc = nccl_api();
        if (!id || !nc.ok) {
            kg_comm_free(c);
            KG_FAIL(KG_EINVAL, "kg_comm_init: %s", id ? "libnccl.so.2 not loadable" : "null id");
        }
        // The exchanges are a few large point-to-point messages per peer: more channels than NCCL's default move them faster
        // (2 GPUs, 1.25 GB each way: 8 channels 10.0 ms, default 3.6 ms, 64 channels 2.8 ms).  The caller's environment wins;
        // NCCL reads it once per process, so a host that creates its own communicators first should set it itself.
        setenv("NCCL_MIN_P2P_NCHANNELS", "64", 0);
        setenv("NCCL_MAX_P2P_NCHANNELS", "64", 0);
        ncclUniqueId u;
        memcpy(&u, id, KG_COMM_ID_BYTES);
        ncclResult_t e = nc.CommInitRank(&c->nccl, nranks, u, rank);
        if (e != ncclSuccess) {
            c->nccl = nullptr;
            kg_comm_free(c);
            KG_FAIL(KG_ECUDA, "ncclCommInitRank failed: %s", nc.GetErrorString(e));
        }
        int lo_p = 0, hi_p = 0; // the exchange kernels must not queue behind a grid of probe blocks that fills every SM
        cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
        if (cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi_p) != cudaSuccess) {
            cudaGetLastError();
            kg_comm_free(c);
            KG_FAIL(KG_ECUDA, "kg_comm_init: communication stream");
        }
    }
    *comm = c;
    return KG_OK;
}

extern "C" int kg_comm_init_local(kg_context* const* ctxs, int nranks, kg_comm** comms) {
    if (!ctxs || !comms) KG_FAIL(KG_EINVAL, "kg_comm_init_local: null argument");
    if (nranks < 1 || nranks > KG_MAX_RANKS) KG_FAIL(KG_EINVAL, "kg_comm_init_local: %d ranks (at most %d)", nranks, KG_MAX_RANKS);
    KgLocalGroup* g = new KgLocalGroup();
    for (int r = 0; r < nranks; r++) {
        kg_comm* c = nullptr;
        int rc = ctxs[r] ? comm_alloc(ctxs[r], r, nranks, &c) : KG_EINVAL;
        if (rc != KG_OK) {
            for (kg_comm* m : g->members) {
                m->group = nullptr;
                kg_comm_free(m);
            }
            delete g;
            return rc;
        }
        c->group = g;
        g->members.push_back(c);
    }
    g->alive = nranks;
    for (int r = 0; r < nranks; r++) {
        comms[r] = g->members[r];
        for (int p = 0; p < nranks; p++) { // peer copies between two devices of this process go direct when they can
            const int a = ctxs[r]->device, b = ctxs[p]->device;
            int can = 0;
            if (a != b && cudaDeviceCanAccessPeer(&can, a, b) == cudaSuccess && can) {
                cudaSetDevice(a);
                if (cudaDeviceEnablePeerAccess(b, 0) != cudaSuccess) cudaGetLastError(); // already enabled
            }
        }
    }
    return KG_OK;
}

extern "C" void kg_comm_free(kg_comm* c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    cudaDeviceSynchronize();
    if (c->nccl) nccl_api().CommDestroy(c->nccl);
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    for (auto& k : c->ch) {
        for (DevBuf* d : {&k.send_lo, &k.send_hi, &k.send_pos, &k.send_cnt, &k.recv_lo, &k.recv_hi, &k.reply_idx, &k.reply_payload, &k.reply_cnt,
                          &k.rr_idx, &k.rr_payload, &k.matrix})
            d->release();
        if (k.h) cudaFreeHost(k.h);
        for (cudaEvent_t e : {k.ev_route, k.ev_keys, k.ev_answer, k.ev_replies})
            if (e) cudaEventDestroy(e);
    }
    for (DevBuf* d : {&c->bitmap, &c->word_cnt, &c->word_rank, &c->xstage}) d->release();
    dx_close(c);
    c->xbuf.release();
    if (c->h_ctl) cudaFreeHost(c->h_ctl);
    if (c->ev_begin) cudaEventDestroy(c->ev_begin);
    if (c->group && --c->group->alive == 0) delete c->group;
    delete c;
}

extern "C" int kg_comm_last_stats(const kg_comm* c, kg_shard_stats* s) {
    if (!c || !s) KG_FAIL(KG_EINVAL, "kg_comm_last_stats: null argument");
    *s = c->stats;
    return KG_OK;
}

// Chunks of a step over NCCL (KG_SHARD_CHUNKS, default 4).  The number must not depend on the rank's own batch: every rank
// has to issue the same sequence of collectives, whatever it holds itself (an empty chunk still takes part).
static int shard_chunks_wanted(const kg_comm* c, const kg_batch*) {
    if (!c->nccl) return 1;
    int want = KG_MAX_CHUNKS;
    if (const char* e = getenv("KG_SHARD_CHUNKS")) want = atoi(e);
    return std::max(1, std::min(want, KG_MAX_CHUNKS));
}

extern "C" int kg_batch_run_sharded(kg_comm* c, const kg_table* shard, kg_batch* batch, const kg_params* params, kg_result** result) {
    if (!result) KG_FAIL(KG_EINVAL, "kg_batch_run_sharded: null argument");
    KG_TRY(shard_check(c, shard, batch, params));
    if (c->group && c->nranks > 1) KG_FAIL(KG_EINVAL, "kg_batch_run_sharded: local communicators run through kg_batch_run_sharded_local");
    const auto t0 = std::chrono::steady_clock::now();
    kg_context* ctx = c->ctx;
    shard_l2_setaside(shard, false);
    struct Restore {
        const kg_table* t;
        ~Restore() { shard_l2_setaside(t, true); }
    } restore{shard};
    if (c->nccl && shard_direct_wanted() && !c->direct_off && !c->xbuf.p) KG_TRY(dx_setup(c, dx_cap_wanted(c, batch))); // collective; may switch direct_off on
    if (c->nccl && shard_direct_wanted() && !c->direct_off) {
        // Direct transport: route -> (flags) -> answer -> (flags) -> merge, all on the compute stream, one host wait at the end.
        bool done = false;
        for (int attempt = 0;; attempt++) {
            KG_TRY(shard_begin(c, batch, 1));
            KG_TRY(dx_route(c, batch, true));
            KG_TRY(dx_answer(c, shard, true));
            // The persisting-L2 set-aside is a device-wide limit that is switched from the host, not in stream order: it is on
            // (for the prefilter) while route and answer run and must be off for the merge, so the host waits for the replies
            // here -- the one wait inside a step (the merge's own wait ends it).
            CU(cudaEventSynchronize(c->ch[0].ev_replies));
            KG_TRY(shard_merge(c, shard, batch, params, result, true));
            if (*result) {
                done = true;
                break;
            }
            if (attempt) KG_FAIL(KG_ECUDA, "kg_batch_run_sharded: bin overflow persisted at capacity %llu", (unsigned long long)c->xl.cap);
            // a skewed batch (or a larger one than the buffers were made for): all ranks have read the same max_bin values
            KG_TRY(dx_setup(c, c->ch[0].cap_seen + c->ch[0].cap_seen / 32 + 4096));
            if (c->direct_off) break; // the larger buffers could not be mapped: the staged transport takes the step
        }
        if (done) {
            c->stats.ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
            return KG_OK;
        }
    }
    KG_TRY(shard_begin(c, batch, shard_chunks_wanted(c, batch)));
    const int H = c->nchunks;
    if (!c->nccl) { // a single rank: the whole path without an interconnect
        ShardChunk& k = c->ch[0];
        KG_TRY(shard_route_sync(c, k, batch, ctx->stream));
        KG_TRY(shard_exchange_keys(c, k, 0, batch, ctx->stream));
        KG_TRY(shard_answer(c, k, shard, true));
        KG_TRY(shard_exchange_replies(c, k, 0, ctx->stream));
    } else {
        // Every rank issues the same sequence, so the collectives match up:
        //   compute stream : route(0) .. route(H-1), answer(0), answer(1), ...
        //   comm stream    : keys(0), keys(1), replies(0), keys(2), replies(1), ...
        // The host only ever waits on the comm stream (for the counters that size the next send/recv), while the kernels
        // queued on the compute stream keep the GPU busy: the exchange of a chunk overlaps the answer kernel of its neighbours.
        for (int h = 0; h < H; h++) {
            c->ch[h].cap = shard_cap_guess(c, c->ch[h]);
            KG_TRY(shard_route(c, c->ch[h], batch, ctx->stream));
        }
        for (int h = 0; h < H; h++) {
            KG_TRY(shard_exchange_keys(c, c->ch[h], h, batch, c->comm_stream));
            KG_TRY(shard_answer(c, c->ch[h], shard, false));
            if (h >= 1) KG_TRY(shard_exchange_replies(c, c->ch[h - 1], h - 1, c->comm_stream));
        }
        KG_TRY(shard_exchange_replies(c, c->ch[H - 1], H - 1, c->comm_stream));
    }
    KG_TRY(shard_merge(c, shard, batch, params, result));
    c->stats.ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return KG_OK;
}

extern "C" int kg_batch_run_sharded_local(kg_comm* const* comms, const kg_table* const* shards, kg_batch* const* batches, int nranks,
                                          const kg_params* params, kg_result** results) {
    if (!comms || !shards || !batches || !results || nranks < 1) KG_FAIL(KG_EINVAL, "kg_batch_run_sharded_local: null argument");
    for (int r = 0; r < nranks; r++) {
        KG_TRY(shard_check(comms[r], shards[r], batches[r], params));
        if (!comms[r]->group || comms[r]->nranks != nranks || comms[r]->group != comms[0]->group || comms[r]->rank != r)
            KG_FAIL(KG_EINVAL, "kg_batch_run_sharded_local: comms must be the %d members of one local group, in rank order", nranks);
        results[r] = nullptr;
    }
    const auto t0 = std::chrono::steady_clock::now();
    int rc = KG_OK;
    for (int r = 0; r < nranks; r++) {
        cudaSetDevice(comms[r]->ctx->device);
        shard_l2_setaside(shards[r], false);
    }
    struct Restore {
        kg_comm* const* comms;
        const kg_table* const* shards;
        int n;
        ~Restore() {
            for (int r = 0; r < n; r++) {
                cudaSetDevice(comms[r]->ctx->device);
                shard_l2_setaside(shards[r], true);
            }
        }
    } restore{comms, shards, nranks};
    auto each = [&](auto&& fn) {
        for (int r = 0; r < nranks && rc == KG_OK; r++) {
            cudaSetDevice(comms[r]->ctx->device);
            rc = fn(r);
        }
    };
    auto sync_all = [&]() {
        each([&](int r) {
            if (cudaStreamSynchronize(comms[r]->ctx->stream) != cudaSuccess) {
                kg_set_error("kg_batch_run_sharded_local: %s", cudaGetErrorString(cudaGetLastError()));
                return (int)KG_ECUDA;
            }
            return (int)KG_OK;
        });
    };
    if (shard_direct_wanted() && nranks > 1) {
        // The direct transport inside one process: the same kernels store into their peers' buffers (peer access between
        // devices, plain pointers on one device); the phases are separated by host synchronisation instead of arrival flags
        // (virtual ranks may share one stream, where a waiting kernel would wait for work queued behind it).
        for (int attempt = 0; rc == KG_OK; attempt++) {
            each([&](int r) { return shard_begin(comms[r], batches[r], 1); });
            uint64_t cap = 0;
            bool fresh = false;
            for (int r = 0; r < nranks; r++) {
                cap = std::max(cap, std::max(dx_cap_wanted(comms[r], batches[r]), comms[r]->xl.cap));
                if (attempt) cap = std::max<uint64_t>(cap, comms[r]->ch[0].cap_seen + comms[r]->ch[0].cap_seen / 32 + 4096);
            }
            for (int r = 0; r < nranks; r++) fresh = fresh || !comms[r]->xbuf.p || comms[r]->xl.cap != cap;
            if (fresh) {
                sync_all();
                each([&](int r) { return dx_alloc(comms[r], cap); });
                for (int r = 0; r < nranks && rc == KG_OK; r++)
                    for (int p = 0; p < nranks; p++) comms[r]->xpeers.base[p] = comms[p]->xbuf.as<uint8_t>();
            }
            each([&](int r) { return dx_route(comms[r], batches[r], false); });
            sync_all();
            each([&](int r) { return dx_answer(comms[r], shards[r], false); });
            sync_all();
            each([&](int r) { return shard_merge(comms[r], shards[r], batches[r], params, &results[r], true); });
            if (rc != KG_OK) break;
            bool again = false;
            for (int r = 0; r < nranks; r++) again = again || !results[r];
            if (!again) break;
            for (int r = 0; r < nranks; r++) {
                kg_result_free(results[r]);
                results[r] = nullptr;
            }
            if (attempt) {
                kg_set_error("kg_batch_run_sharded_local: bin overflow persisted at capacity %llu", (unsigned long long)cap);
                rc = KG_ECUDA;
            }
        }
        if (rc != KG_OK) {
            for (int r = 0; r < nranks; r++) {
                kg_result_free(results[r]);
                results[r] = nullptr;
            }
            return rc;
        }
        const float ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        for (int r = 0; r < nranks; r++) comms[r]->stats.ms_total = ms;
        return KG_OK;
    }
    // every phase ends host-synchronised on every rank before the next one reads its neighbours' buffers
    each([&](int r) { int e = shard_begin(comms[r], batches[r], 1); return e != KG_OK ? e : shard_route_sync(comms[r], comms[r]->ch[0], batches[r], comms[r]->ctx->stream); });
    each([&](int r) { return shard_exchange_keys(comms[r], comms[r]->ch[0], 0, batches[r], comms[r]->ctx->stream); });
    each([&](int r) { return shard_answer(comms[r], comms[r]->ch[0], shards[r], true); });
    each([&](int r) { return shard_exchange_replies(comms[r], comms[r]->ch[0], 0, comms[r]->ctx->stream); });
    each([&](int r) { // the reply copies read the owners' buffers: finish them before anyone merges
        if (cudaStreamSynchronize(comms[r]->ctx->stream) != cudaSuccess) {
            kg_set_error("kg_batch_run_sharded_local: %s", cudaGetErrorString(cudaGetLastError()));
            return (int)KG_ECUDA;
        }
        return (int)KG_OK;
    });
    each([&](int r) { return shard_merge(comms[r], shards[r], batches[r], params, &results[r]); });
    if (rc != KG_OK) {
        for (int r = 0; r < nranks; r++) {
            kg_result_free(results[r]);
            results[r] = nullptr;
        }
        return rc;
    }
    const float ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    for (int r = 0; r < nranks; r++) comms[r]->stats.ms_total = ms;
    return KG_OK;
}
