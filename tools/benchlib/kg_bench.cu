// kg_bench.cu -- bench/test tooling (tools/benchlib/kmerguts_bench.h), built as libkmerguts_bench.so, SEPARATE from the
// product library: probe-roofline microbenchmark, CUDA generators of the synthetic universe of tools/kg_synth.py, a
// device-side writer of the reference's table format, a naive cross-check scan.  Nothing here is on the product path;
// it links against libkmerguts_b200.so only for the context / table handles it is handed.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include <algorithm>

#include "kmerguts_bench.h"
#include "../../kmergutsjava_b200/csrc/kg_device.cuh"
#include "../../kmergutsjava_b200/csrc/kg_internal.h"

namespace {

inline unsigned blocks_for(size_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

// ---- the counter-based hashing of tools/kg_synth.py (mix64 / hash3) ----
__host__ __device__ __forceinline__ uint64_t smix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t hash3(uint64_t seed, uint64_t a, uint64_t b) {
    return smix64(smix64(seed + a) ^ (b * 0xD6E8FEB86659FD93ull));
}

struct DevUniverse {
    uint64_t n_families, seed;
    uint32_t keep, n_functions, n_otus;
    const uint32_t* lenq; // device, 4096
};
__constant__ uint32_t c_cdf16[20];

__device__ __forceinline__ uint32_t code_of_draw(uint32_t draw) { // np.searchsorted(cdf, draw, side="right")
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < 19; k++) c += (c_cdf16[k] <= draw);
    return c;
}
__device__ __forceinline__ uint32_t family_len(const DevUniverse& u, uint64_t f) {
    return u.lenq[hash3(u.seed, f, 0xFFFFFFFFull) & 4095];
}
__device__ __forceinline__ uint32_t residue_code(const DevUniverse& u, uint64_t f, uint64_t i) {
    return code_of_draw((uint32_t)(hash3(u.seed ^ 0x11, f, i) & 0xFFFF));
}
__device__ __forceinline__ uint32_t random_code(uint64_t seed, uint64_t a, uint64_t b) {
    return code_of_draw((uint32_t)(hash3(seed, a, b) & 0xFFFF));
}
__device__ __forceinline__ bool is_signature(const DevUniverse& u, uint64_t f, uint64_t i) {
    return (uint32_t)(hash3(u.seed ^ 0x22, f, i) & 1023) < u.keep;
}
__device__ __forceinline__ float weight_of_key(uint64_t key) { // 0.5 + b/256, exact in fp32
    uint32_t b = (uint32_t)((smix64(key) >> 17) & 0xFF);
    return 0.5f + (float)b * (1.0f / 256.0f);
}

__global__ void k_sig_count(DevUniverse u, uint32_t* __restrict__ cnt, uint32_t rank, uint32_t nranks) {
    uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f > u.n_families) return;
    uint32_t c = 0;
    if (f < u.n_families) {
        uint32_t L = family_len(u, f);
        if (nranks <= 1) {
            for (uint32_t i = 0; i + KG_K <= L; i++) c += is_signature(u, f, i);
        } else { // one shard of the table: only the candidates whose key this rank owns
            uint64_t enc = 0;
            for (uint32_t i = 0; i < L; i++) {
                enc = (enc % 1280000000ull) * 20ull + residue_code(u, f, i);
                if (i + 1 >= KG_K && is_signature(u, f, i + 1 - KG_K) && kg_owner_of(enc, nranks) == rank) c++;
            }
        }
    }
    cnt[f] = c;
}
__global__ void k_sig_fill(DevUniverse u, const uint64_t* __restrict__ off, uint64_t* __restrict__ keys, int4* __restrict__ payload,
                           uint32_t rank, uint32_t nranks) {
    uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= u.n_families) return;
    uint32_t L = family_len(u, f);
    uint64_t o = off[f];
    uint64_t enc = 0;
    for (uint32_t i = 0; i < L; i++) {
        enc = (enc % 1280000000ull) * 20ull + residue_code(u, f, i);
        if (i + 1 >= KG_K) {
            uint32_t w = i + 1 - KG_K; // window start
            if (is_signature(u, f, w) && (nranks <= 1 || kg_owner_of(enc, nranks) == rank)) {
                keys[o] = enc;
                payload[o] = make_int4((int)(f % u.n_otus), (int)(L - w), (int)(f % u.n_functions), __float_as_int(weight_of_key(enc)));
                o++;
            }
        }
    }
}
__global__ void k_iota(uint32_t* a, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = (uint32_t)i;
}
__global__ void k_flag_first_key(const uint64_t* __restrict__ k, size_t n, uint8_t* __restrict__ flag) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == 0) || (k[i] != k[i - 1]);
}
__global__ void k_gather_entries(const uint32_t* __restrict__ r, size_t n, const uint64_t* __restrict__ ck,
                                 const int4* __restrict__ cp, uint64_t* __restrict__ ok, int4* __restrict__ op) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ok[i] = ck[r[i]];
    op[i] = cp[r[i]];
}

// ---- proteins ----
__global__ void k_prot_len(DevUniverse u, uint64_t first, uint64_t n, uint64_t seed, uint64_t* __restrict__ len) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n) return;
    if (t == n) { len[t] = 0; return; }
    uint64_t j = first + t;
    uint64_t h = hash3(seed, j, 0xFFFFFFFFull);
    if ((h & 0xFF) < 51) len[t] = u.lenq[(h >> 8) & 4095];
    else len[t] = family_len(u, (h >> 8) % u.n_families);
}
__global__ void k_prot_fill(DevUniverse u, uint64_t first, uint64_t n, uint64_t seed, const uint64_t* __restrict__ off,
                            uint8_t* __restrict__ out) {
    const uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; // one warp per protein
    const int lane = threadIdx.x & 31;
    if (t >= n) return;
    const uint64_t j = first + t;
    const uint64_t h = hash3(seed, j, 0xFFFFFFFFull);
    const bool decoy = (h & 0xFF) < 51;
    const uint64_t f = (h >> 8) % u.n_families;
    const uint64_t o = off[t], L = off[t + 1] - o;
    const char* alpha = "ACDEFGHIKLMNPQRSTVWY";
    for (uint64_t i = lane; i < L; i += 32) {
        uint32_t code;
        if (decoy) {
            code = random_code(seed ^ 0x33, j, i);
        } else {
            bool sub = (hash3(seed ^ 0x44, j, i) & 0xFFFF) < 6554;
            code = sub ? random_code(seed ^ 0x55, j, i) : residue_code(u, f, i);
        }
        bool isx = ((hash3(seed ^ 0x66, j, i) >> 20) & 0xFFFFF) < 105;
        out[o + i] = isx ? (uint8_t)'X' : (uint8_t)alpha[code];
    }
}

// ---- genomes (configs[2]): genes = family consensus proteins back-translated with a uniformly chosen synonymous codon,
// random strand, separated by random spacers; spacer bases are uniform ACGT; ~1e-5 of all bases become N ----
__device__ __constant__ char c_gcode[65] = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF";
__global__ void k_genome_background(uint8_t* __restrict__ out, uint64_t total, uint64_t seed, uint64_t base_pos) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint64_t h = hash3(seed ^ 0x77, base_pos + i, 0);
    out[i] = ((h >> 8) % 100000ull == 0) ? (uint8_t)'N' : (uint8_t)"ACGT"[h & 3];
}
struct GenePlan { // one gene: where it starts in the concatenated genome stream, which family, which strand
    uint64_t start;
    uint32_t family;
    uint32_t minus;
};
__global__ void k_genome_genes(DevUniverse u, const GenePlan* __restrict__ plan, uint64_t ngenes, uint64_t seed,
                               uint8_t* __restrict__ out, uint64_t gene_base) {
    const uint64_t gl = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; // one warp per gene
    const int lane = threadIdx.x & 31;
    if (gl >= ngenes) return;
    const GenePlan gp = plan[gl];
    const uint64_t g = gene_base + gl; // index of the gene in the whole job: a range of genomes gets the bytes the full job gets
    const uint32_t L = family_len(u, gp.family);
    const char* alpha = "ACDEFGHIKLMNPQRSTVWY";
    for (uint32_t c = lane; c < L; c += 32) {
        const char aa = alpha[residue_code(u, gp.family, c)];
        // the k-th codon (k uniform) among the codons of this residue in GENETIC_CODE order
        uint32_t ncod = 0;
        for (int i = 0; i < 64; i++) ncod += (c_gcode[i] == aa);
        uint32_t k = (uint32_t)(hash3(seed ^ 0x88, g, c) % ncod), idx = 0;
        for (int i = 0; i < 64; i++)
            if (c_gcode[i] == aa) {
                if (k == 0) { idx = i; break; }
                k--;
            }
        const char nt[3] = {"ACGT"[idx >> 4], "ACGT"[(idx >> 2) & 3], "ACGT"[idx & 3]};
        if (!gp.minus) {
            for (int j = 0; j < 3; j++) out[gp.start + 3ull * c + j] = (uint8_t)nt[j];
        } else { // reverse complement of the whole gene
            const uint64_t end = gp.start + 3ull * L - 1;
            for (int j = 0; j < 3; j++) out[end - (3ull * c + j)] = (uint8_t)"TGCA"[(nt[j] == 'A') ? 0 : (nt[j] == 'C') ? 1 : (nt[j] == 'G') ? 2 : 3];
        }
    }
}

// ---- reference-format image: linear probing without wrap over 24-byte slots (3 x uint64 words each).
// Keys taken in home-slot order get the first free slot at or after their home:
//   slot_r = max(slot_{r-1} + 1, home_r) = r + max_{q<=r}(home_q - q)   -- the layout sequential insertion in that order
// produces.  (key % numSigs of real 8-mers is far from uniform, so a CAS-per-probe insert would crawl through runs that
// are thousands of slots long; the scan places every key in O(1).)
__global__ void k_image_init(unsigned long long* __restrict__ w, uint64_t num_slots) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= num_slots) return;
    w[3 * s] = (unsigned long long)(KG_MAX_ENCODED + 1); // whichKmer > MAX_ENCODED marks an empty slot (KGJ:1000)
    w[3 * s + 1] = 0;
    w[3 * s + 2] = 0;
}
__global__ void k_image_home(const uint64_t* __restrict__ keys, uint64_t n, uint64_t num_slots, uint32_t* __restrict__ home,
                             uint32_t* __restrict__ idx) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    home[i] = (uint32_t)(keys[i] % num_slots); // the reference's hash, KGJ:969
    idx[i] = (uint32_t)i;
}
__global__ void k_image_bias(const uint32_t* __restrict__ home, uint64_t n, long long* __restrict__ t) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) t[r] = (long long)home[r] - (long long)r;
}
struct MaxLL {
    __host__ __device__ __forceinline__ long long operator()(long long a, long long b) const { return a > b ? a : b; }
};
// ctr[0] = keys that would run off the end, ctr[1] = sum of (slot - home), ctr[2] = max of (slot - home)
__global__ void k_image_place(const uint32_t* __restrict__ home, const uint32_t* __restrict__ idx, const long long* __restrict__ tmax,
                              uint64_t n, const uint64_t* __restrict__ keys, const int4* __restrict__ payload,
                              unsigned long long* __restrict__ w, uint64_t num_slots, unsigned long long* __restrict__ ctr) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long disp = 0;
    if (r < n) {
        const uint64_t s = (uint64_t)((long long)r + tmax[r]);
        disp = s - home[r];
        if (s + 1 >= num_slots) { // the last slot stays empty so that no chain can run off the end
            atomicAdd(&ctr[0], 1ull);
        } else {
            const int4 p = payload[idx[r]];
            w[3 * s] = keys[idx[r]];
            w[3 * s + 1] = (unsigned long long)(uint32_t)p.x | ((unsigned long long)(uint32_t)p.y << 32); // otu, avgFromEnd
            w[3 * s + 2] = (unsigned long long)(uint32_t)p.z | ((unsigned long long)(uint32_t)p.w << 32); // fI, wt bits
        }
    }
    unsigned long long sum = disp, mx = disp;
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
        unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, mx, d);
        mx = o > mx ? o : mx;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&ctr[1], sum);
        atomicMax(&ctr[2], mx);
    }
}

// ---- size-independent cross-check of the probe path (tests / bench): the most naive kernel possible -- one thread per
// protein-stream position, residues re-read byte by byte, full unfiltered table lookup -- counts the valid windows, the
// hits, and a checksum over (position, payload) that the pipeline's hit records must reproduce ----
__global__ void k_naive_scan(const uint8_t* __restrict__ seq, const uint64_t* __restrict__ off, uint64_t n, uint64_t total,
                             KgTableView tab, unsigned long long* __restrict__ out) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long valid = 0, hit = 0, sum = 0;
    if (g < total) {
        // sequence of g by binary search; aa-mode rule of the reference: windows i with i < len - 8 (KGJ:912, 1055)
        uint64_t lo = 0, hi = n;
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) >> 1;
            if (off[mid] <= g) lo = mid;
            else hi = mid;
        }
        const uint64_t i = g - off[lo], len = off[lo + 1] - off[lo];
        if (len > 8 && i < len - 8) {
            uint64_t key = 0;
            bool ok = true;
            for (int k = 0; k < 8; k++) {
                const uint8_t ch = seq[g + k];
                int code = 20;
                const char* alpha = "ACDEFGHIKLMNPQRSTVWY";
                for (int a = 0; a < 20; a++)
                    if (alpha[a] == (char)ch) code = a;
                if (code >= 20) ok = false;
                key = key * 20 + (uint64_t)code;
            }
            if (ok) {
                valid = 1;
                const uint32_t slot = kg_lookup(tab, key);
                if (slot != 0xFFFFFFFFu) {
                    const int4 p = kg_load_payload(tab.lines, slot);
                    hit = 1;
                    sum = smix64(g * 0x9E3779B97F4A7C15ull ^ (uint64_t)(uint32_t)p.x ^ ((uint64_t)(uint32_t)p.z << 32)) ^
                          smix64(((uint64_t)(uint32_t)p.y << 32) | (uint32_t)p.w);
                }
            }
        }
    }
    // block reduction: sums of valid / hit, XOR... use wrapping sum for the checksum so order does not matter
    for (int d = 16; d; d >>= 1) {
        valid += __shfl_xor_sync(0xFFFFFFFFu, valid, d);
        hit += __shfl_xor_sync(0xFFFFFFFFu, hit, d);
        sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (valid) atomicAdd(&out[0], valid);
        if (hit) atomicAdd(&out[1], hit);
        if (sum) atomicAdd(&out[2], sum);
    }
}
// the same checksum over hit records (seq-relative positions are turned back into stream positions with the offsets)
__global__ void k_hits_checksum(const kg_hit* __restrict__ hits, uint64_t nhits, const uint64_t* __restrict__ off,
                                unsigned long long* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long sum = 0;
    if (i < nhits) {
        const kg_hit h = hits[i];
        const uint64_t g = off[h.seq] + (uint64_t)h.pos;
        sum = smix64(g * 0x9E3779B97F4A7C15ull ^ (uint64_t)(uint32_t)h.oI ^ ((uint64_t)(uint32_t)h.fI << 32)) ^
              smix64(((uint64_t)(uint32_t)h.avg_off_from_end << 32) | (uint32_t)__float_as_int(h.function_wt));
    }
    for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd(&out[0], sum);
}

// ---- R_probe ----
template <int U>
__global__ void k_random_sectors(const uint4* __restrict__ buf, uint64_t n_sectors, uint32_t per_thread, uint32_t* __restrict__ sink) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint32_t it = 0; it < per_thread; it += U) {
        KgBucket b[U];
#pragma unroll
        for (int k = 0; k < U; k++) {
            uint64_t h = smix64(tid * 0x100000001B3ull + it + k);
            uint32_t idx = (uint32_t)(((h >> 32) * n_sectors) >> 32);
            b[k] = kg_load_sector(buf + 2ull * idx);
        }
#pragma unroll
        for (int k = 0; k < U; k++) acc ^= b[k].w[0] ^ b[k].w[3] ^ b[k].w[7];
    }
    if (acc == 0x9E3779B9u) sink[0] = acc; // keeps the loads alive
}

int upload_universe(kg_context* ctx, const kg_universe* u, DevUniverse* du, uint32_t** d_lenq) {
    CU(cudaMalloc(d_lenq, 4096 * 4));
    CU(cudaMemcpyAsync(*d_lenq, u->lenq, 4096 * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyToSymbolAsync(c_cdf16, u->cdf16, 20 * 4, 0, cudaMemcpyHostToDevice, ctx->stream));
    du->n_families = u->n_families;
    du->seed = u->seed;
    du->keep = u->sig_keep_per_1024;
    du->n_functions = u->n_functions;
    du->n_otus = u->n_otus;
    du->lenq = *d_lenq;
    return KG_OK;
}

} // namespace

extern "C" void kg_device_free(void* p) {
    if (p) cudaFree(p);
}
extern "C" int kg_device_to_host(kg_context* ctx, void* host, const void* dev, uint64_t bytes) {
    if (!ctx || (bytes && (!host || !dev))) KG_FAIL(KG_EINVAL, "kg_device_to_host: null argument");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KG_OK;
}

extern "C" int kg_synth_signatures(kg_context* ctx, const kg_universe* u, uint64_t max_sigs, uint64_t** out_keys,
                                   void** out_payload, uint64_t* out_n) {
    return kg_synth_signatures_sharded(ctx, u, max_sigs, 0, 1, out_keys, out_payload, out_n);
}

extern "C" int kg_synth_signatures_sharded(kg_context* ctx, const kg_universe* u, uint64_t max_sigs, int rank, int nranks,
                                           uint64_t** out_keys, void** out_payload, uint64_t* out_n) {
    if (!ctx || !u || !out_keys || !out_payload || !out_n) KG_FAIL(KG_EINVAL, "kg_synth_signatures: null argument");
    if (nranks < 1 || nranks > KG_MAX_RANKS || rank < 0 || rank >= nranks) KG_FAIL(KG_EINVAL, "kg_synth_signatures_sharded: rank %d of %d", rank, nranks);
    if (nranks > 1 && max_sigs) KG_FAIL(KG_EINVAL, "kg_synth_signatures_sharded: max_sigs truncates in family order and needs the whole set; pass 0");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevUniverse du;
    uint32_t* d_lenq = nullptr;
    KG_TRY(upload_universe(ctx, u, &du, &d_lenq));
    const uint64_t F = u->n_families;
    uint32_t* cnt = nullptr;
    uint64_t* off = nullptr;
    CU(cudaMalloc(&cnt, (F + 1) * 4));
    CU(cudaMalloc(&off, (F + 1) * 8));
    k_sig_count<<<blocks_for(F + 1, 128), 128, 0, st>>>(du, cnt, (uint32_t)rank, (uint32_t)nranks);
    size_t tmp = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp, cnt, off, F + 1, st));
    KG_TRY(ctx->scan_tmp.ensure(tmp));
    CU(cub::DeviceScan::ExclusiveSum(ctx->scan_tmp.p, tmp, cnt, off, F + 1, st));
    uint64_t C = 0;
    CU(cudaMemcpyAsync(&C, off + F, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (C >= (1ull << 32)) KG_FAIL(KG_ERANGE, "kg_synth_signatures: %llu candidates", (unsigned long long)C);
    uint64_t *ck = nullptr, *ck2 = nullptr;
    int4* cp = nullptr;
    uint32_t *r1 = nullptr, *r2 = nullptr;
    uint8_t* flag = nullptr;
    size_t* d_nsel = nullptr;
    CU(cudaMalloc(&ck, std::max<uint64_t>(C, 1) * 8));
    CU(cudaMalloc(&ck2, std::max<uint64_t>(C, 1) * 8));
    CU(cudaMalloc(&cp, std::max<uint64_t>(C, 1) * sizeof(int4)));
    CU(cudaMalloc(&r1, std::max<uint64_t>(C, 1) * 4));
    CU(cudaMalloc(&r2, std::max<uint64_t>(C, 1) * 4));
    CU(cudaMalloc(&flag, std::max<uint64_t>(C, 1)));
    CU(cudaMalloc(&d_nsel, sizeof(size_t)));
    k_sig_fill<<<blocks_for(F, 128), 128, 0, st>>>(du, off, ck, cp, (uint32_t)rank, (uint32_t)nranks);
    k_iota<<<blocks_for(C, 256), 256, 0, st>>>(r1, C);
    // first occurrence (lowest candidate rank) of every key: stable sort by key, keep the head of each run
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp, ck, ck2, r1, r2, C, 0, 35, st));
    KG_TRY(ctx->scan_tmp.ensure(tmp));
    CU(cub::DeviceRadixSort::SortPairs(ctx->scan_tmp.p, tmp, ck, ck2, r1, r2, C, 0, 35, st));
    k_flag_first_key<<<blocks_for(C, 256), 256, 0, st>>>(ck2, C, flag);
    CU(cub::DeviceSelect::Flagged(nullptr, tmp, r2, flag, r1, d_nsel, C, st));
    KG_TRY(ctx->scan_tmp.ensure(tmp));
    CU(cub::DeviceSelect::Flagged(ctx->scan_tmp.p, tmp, r2, flag, r1, d_nsel, C, st));
    size_t U = 0;
    CU(cudaMemcpyAsync(&U, d_nsel, sizeof(size_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    // back to family-major order, then truncate
    CU(cub::DeviceRadixSort::SortKeys(nullptr, tmp, r1, r2, U, 0, 32, st));
    KG_TRY(ctx->scan_tmp.ensure(tmp));
    CU(cub::DeviceRadixSort::SortKeys(ctx->scan_tmp.p, tmp, r1, r2, U, 0, 32, st));
    uint64_t n = max_sigs ? std::min<uint64_t>(U, max_sigs) : U;
    uint64_t* ok = nullptr;
    int4* op = nullptr;
    CU(cudaMalloc(&ok, std::max<uint64_t>(n, 1) * 8));
    CU(cudaMalloc(&op, std::max<uint64_t>(n, 1) * sizeof(int4)));
    k_gather_entries<<<blocks_for(n, 256), 256, 0, st>>>(r2, n, ck, cp, ok, op);
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    for (void* p : {(void*)cnt, (void*)off, (void*)ck, (void*)ck2, (void*)cp, (void*)r1, (void*)r2, (void*)flag, (void*)d_nsel, (void*)d_lenq}) cudaFree(p);
    *out_keys = ok;
    *out_payload = op;
    *out_n = n;
    return KG_OK;
}

extern "C" int kg_synth_proteins(kg_context* ctx, const kg_universe* u, uint64_t first, uint64_t n, uint64_t seed,
                                 uint8_t** d_seq, uint64_t** d_off, uint64_t* total_bytes) {
    if (!ctx || !u || !d_seq || !d_off || !total_bytes) KG_FAIL(KG_EINVAL, "kg_synth_proteins: null argument");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevUniverse du;
    uint32_t* d_lenq = nullptr;
    KG_TRY(upload_universe(ctx, u, &du, &d_lenq));
    uint64_t *len = nullptr, *off = nullptr;
    CU(cudaMalloc(&len, (n + 1) * 8));
    CU(cudaMalloc(&off, (n + 1) * 8));
    k_prot_len<<<blocks_for(n + 1, 256), 256, 0, st>>>(du, first, n, seed, len);
    size_t tmp = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp, len, off, n + 1, st));
    KG_TRY(ctx->scan_tmp.ensure(tmp));
    CU(cub::DeviceScan::ExclusiveSum(ctx->scan_tmp.p, tmp, len, off, n + 1, st));
    uint64_t total = 0;
    CU(cudaMemcpyAsync(&total, off + n, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    uint8_t* seq = nullptr;
    CU(cudaMalloc(&seq, total + 64));
    CU(cudaMemsetAsync(seq + total, 0, 64, st));
    if (n) k_prot_fill<<<blocks_for(n * 32, 256), 256, 0, st>>>(du, first, n, seed, off, seq);
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    cudaFree(len);
    cudaFree(d_lenq);
    *d_seq = seq;
    *d_off = off;
    *total_bytes = total;
    return KG_OK;
}

extern "C" int kg_synth_genomes(kg_context* ctx, const kg_universe* u, uint64_t n_genomes, uint64_t length, uint64_t seed,
                                uint8_t** d_seq, uint64_t** d_off, uint64_t* total_bytes) {
    return kg_synth_genomes_range(ctx, u, 0, n_genomes, length, seed, d_seq, d_off, total_bytes);
}

extern "C" int kg_synth_genomes_range(kg_context* ctx, const kg_universe* u, uint64_t first, uint64_t n_genomes, uint64_t length, uint64_t seed,
                                      uint8_t** d_seq, uint64_t** d_off, uint64_t* total_bytes) {
    if (!ctx || !u || !d_seq || !d_off || !total_bytes || length < 64) KG_FAIL(KG_EINVAL, "kg_synth_genomes: bad argument");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevUniverse du;
    uint32_t* d_lenq = nullptr;
    KG_TRY(upload_universe(ctx, u, &du, &d_lenq));
    // gene plan on the host (a few thousand genes per genome): spacer ~ 1 + geometric(mean 120), family uniform
    std::vector<GenePlan> plan;
    std::vector<uint64_t> off(n_genomes + 1);
    uint64_t ctr = 0, gene_base = 0, ngenes_seen = 0;
    for (uint64_t gj = 0; gj < first + n_genomes; gj++) { // the plan of the genomes before `first` is walked too: the counter runs on
        const bool keep = gj >= first;
        const uint64_t gi = keep ? gj - first : 0;
        if (gj == first) gene_base = ngenes_seen;
        if (keep) off[gi] = gi * length;
        uint64_t pos = 0;
        for (;;) {
            const uint64_t h = hash3(seed ^ 0x99, gj, ctr++);
            uint64_t gap = 1;
            for (uint64_t r = h >> 20; gap < 2000 && (r % 120) != 0; r = smix64(r)) gap++;
            const uint32_t fam = (uint32_t)((h >> 1) % u->n_families);
            const uint32_t L = u->lenq[hash3(u->seed, fam, 0xFFFFFFFFull) & 4095];
            pos += gap;
            if (pos + 3ull * L > length) break;
            if (keep) plan.push_back(GenePlan{gi * length + pos, fam, (uint32_t)(h & 1)});
            ngenes_seen++;
            pos += 3ull * L;
        }
    }
    off[n_genomes] = n_genomes * length;
    const uint64_t total = n_genomes * length;
    uint8_t* seq = nullptr;
    uint64_t* doff = nullptr;
    GenePlan* dplan = nullptr;
    CU(cudaMalloc(&seq, total + 64));
    CU(cudaMalloc(&doff, (n_genomes + 1) * 8));
    CU(cudaMalloc(&dplan, std::max<size_t>(plan.size(), 1) * sizeof(GenePlan)));
    CU(cudaMemsetAsync(seq + total, 0, 64, st));
    CU(cudaMemcpyAsync(doff, off.data(), (n_genomes + 1) * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(dplan, plan.data(), plan.size() * sizeof(GenePlan), cudaMemcpyHostToDevice, st));
    k_genome_background<<<blocks_for(total, 256), 256, 0, st>>>(seq, total, seed, first * length);
    if (!plan.empty()) k_genome_genes<<<blocks_for(plan.size() * 32, 256), 256, 0, st>>>(du, dplan, plan.size(), seed, seq, gene_base);
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    cudaFree(dplan);
    cudaFree(d_lenq);
    *d_seq = seq;
    *d_off = doff;
    *total_bytes = total;
    return KG_OK;
}

extern "C" int kg_synth_naive_scan_aa(kg_context* ctx, const kg_table* table, const uint8_t* d_seq, const uint64_t* d_off, uint64_t n,
                                      uint64_t total, uint64_t* out3) {
    if (!ctx || !table || !d_off || !out3) KG_FAIL(KG_EINVAL, "kg_synth_naive_scan_aa: null argument");
    CU(cudaSetDevice(ctx->device));
    unsigned long long* d = nullptr;
    CU(cudaMalloc(&d, 24));
    CU(cudaMemsetAsync(d, 0, 24, ctx->stream));
    if (total) k_naive_scan<<<blocks_for(total, 256), 256, 0, ctx->stream>>>(d_seq, d_off, n, total, table->view(), d);
    CU(cudaMemcpyAsync(out3, d, 24, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    cudaFree(d);
    return KG_OK;
}

extern "C" int kg_synth_hits_checksum(kg_context* ctx, const kg_hit* host_hits, uint64_t nhits, const uint64_t* d_off, uint64_t* out) {
    if (!ctx || !out || (nhits && !host_hits)) KG_FAIL(KG_EINVAL, "kg_synth_hits_checksum: null argument");
    CU(cudaSetDevice(ctx->device));
    unsigned long long* d = nullptr;
    kg_hit* dh = nullptr;
    CU(cudaMalloc(&d, 8));
    CU(cudaMalloc(&dh, std::max<uint64_t>(nhits, 1) * sizeof(kg_hit)));
    CU(cudaMemsetAsync(d, 0, 8, ctx->stream));
    CU(cudaMemcpyAsync(dh, host_hits, nhits * sizeof(kg_hit), cudaMemcpyHostToDevice, ctx->stream));
    if (nhits) k_hits_checksum<<<blocks_for(nhits, 256), 256, 0, ctx->stream>>>(dh, nhits, d_off, d);
    CU(cudaMemcpyAsync(out, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    cudaFree(d);
    cudaFree(dh);
    return KG_OK;
}

static bool is_prime_u64(uint64_t n) {
    if (n < 2) return false;
    if (n % 2 == 0) return n == 2;
    for (uint64_t i = 3; i * i <= n; i += 2)
        if (n % i == 0) return false;
    return true;
}

extern "C" int kg_synth_reference_image(kg_context* ctx, const uint64_t* d_keys, const void* d_payload16, uint64_t n,
                                        uint64_t min_slots, uint64_t* num_slots_out, void** d_image,
                                        double* mean_displacement) {
    if (!ctx || !num_slots_out || !d_image || min_slots < 2) KG_FAIL(KG_EINVAL, "kg_synth_reference_image: bad argument");
    if (n >= (1ull << 32) || min_slots >= (1ull << 32) - 4096) KG_FAIL(KG_ERANGE, "kg_synth_reference_image: 32-bit slot index");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    unsigned long long* ctr = nullptr;
    uint32_t *home_a = nullptr, *home_b = nullptr, *idx_a = nullptr, *idx_b = nullptr;
    long long* tb = nullptr;
    const uint64_t n1 = std::max<uint64_t>(n, 1);
    CU(cudaMalloc(&ctr, 32));
    CU(cudaMalloc(&home_a, n1 * 4));
    CU(cudaMalloc(&home_b, n1 * 4));
    CU(cudaMalloc(&idx_a, n1 * 4));
    CU(cudaMalloc(&idx_b, n1 * 4));
    CU(cudaMalloc(&tb, n1 * 8));
    uint64_t num_slots = min_slots;
    int rc = KG_ERANGE;
    // The reference never wraps around (KGJ:959-1026): a slot count for which some probe chain would run off the end
    // is unusable.  Try successive primes until one fits.
    for (int attempt = 0; attempt < 100; attempt++) {
        while (!is_prime_u64(num_slots)) num_slots++;
        unsigned long long* w = nullptr;
        CU(cudaMalloc(&w, 24 + num_slots * 24));
        CU(cudaMemsetAsync(ctr, 0, 32, st));
        const int64_t hdr[3] = {(int64_t)num_slots, 24, 1}; // numSigs, entrySize, version (KGJ:933-935)
        CU(cudaMemcpyAsync(w, hdr, 24, cudaMemcpyHostToDevice, st));
        k_image_init<<<blocks_for(num_slots, 256), 256, 0, st>>>(w + 3, num_slots);
        if (n) {
            k_image_home<<<blocks_for(n, 256), 256, 0, st>>>(d_keys, n, num_slots, home_a, idx_a);
            cub::DoubleBuffer<uint32_t> dk(home_a, home_b), dv(idx_a, idx_b);
            size_t tmp = 0;
            CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, n, 0, 32, st)); // stable: ties keep input order
            KG_TRY(ctx->scan_tmp.ensure(tmp));
            CU(cub::DeviceRadixSort::SortPairs(ctx->scan_tmp.p, tmp, dk, dv, n, 0, 32, st));
            k_image_bias<<<blocks_for(n, 256), 256, 0, st>>>(dk.Current(), n, tb);
            CU(cub::DeviceScan::InclusiveScan(nullptr, tmp, tb, tb, MaxLL(), n, st));
            KG_TRY(ctx->scan_tmp.ensure(tmp));
            CU(cub::DeviceScan::InclusiveScan(ctx->scan_tmp.p, tmp, tb, tb, MaxLL(), n, st));
            k_image_place<<<blocks_for(n, 256), 256, 0, st>>>(dk.Current(), dv.Current(), tb, n, d_keys, (const int4*)d_payload16,
                                                             w + 3, num_slots, ctr);
        }
        unsigned long long h[4] = {0, 0, 0, 0};
        CU(cudaMemcpyAsync(h, ctr, 32, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        CU(cudaGetLastError());
        if (getenv("KG_DEBUG"))
            fprintf(stderr, "[kg] reference image: attempt %d, %llu slots, %llu keys off the end, mean displacement %.2f, max %llu\n",
                    attempt, (unsigned long long)num_slots, h[0], n ? (double)h[1] / (double)n : 0.0, h[2]);
        if (!h[0]) {
            *num_slots_out = num_slots;
            *d_image = w;
            if (mean_displacement) *mean_displacement = n ? (double)h[1] / (double)n : 0.0;
            rc = KG_OK;
            break;
        }
        cudaFree(w);
        num_slots++;
    }
    for (void* p : {(void*)ctr, (void*)home_a, (void*)home_b, (void*)idx_a, (void*)idx_b, (void*)tb}) cudaFree(p);
    if (rc != KG_OK) kg_set_error("reference image: no usable slot count found from %llu upwards", (unsigned long long)min_slots);
    return rc;
}

static int roofline_run(kg_context* ctx, const uint4* buf, uint64_t n_sectors, uint64_t n_loads, int tpb, int inflight, double* out) {
    if (tpb < 32 || tpb > 1024 || (tpb & 31)) KG_FAIL(KG_EINVAL, "threads_per_block must be a multiple of 32 in [32, 1024]");
    if (n_sectors == 0 || n_sectors >= (1ull << 32)) KG_FAIL(KG_EINVAL, "roofline buffer must hold 1..2^32-1 sectors");
    cudaStream_t st = ctx->stream;
    const uint32_t per_thread = 64;
    uint64_t threads = (n_loads + per_thread - 1) / per_thread;
    unsigned grid = blocks_for(threads, (unsigned)tpb);
    uint32_t* sink = nullptr;
    CU(cudaMalloc(&sink, 4));
    auto launch = [&]() {
        switch (inflight) {
            case 1: k_random_sectors<1><<<grid, tpb, 0, st>>>(buf, n_sectors, per_thread, sink); break;
            case 2: k_random_sectors<2><<<grid, tpb, 0, st>>>(buf, n_sectors, per_thread, sink); break;
            case 4: k_random_sectors<4><<<grid, tpb, 0, st>>>(buf, n_sectors, per_thread, sink); break;
            default: k_random_sectors<8><<<grid, tpb, 0, st>>>(buf, n_sectors, per_thread, sink); break;
        }
    };
    launch(); // warm-up
    CU(cudaEventRecord(ctx->ev[0], st));
    launch();
    CU(cudaEventRecord(ctx->ev[1], st));
    CU(cudaEventSynchronize(ctx->ev[1]));
    CU(cudaGetLastError());
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    cudaFree(sink);
    *out = (double)grid * tpb * per_thread / ((double)ms * 1e-3);
    return KG_OK;
}

extern "C" int kg_probe_roofline(kg_context* ctx, uint64_t bytes, uint64_t n_loads, int tpb, int inflight, double* out) {
    if (!ctx || !out) KG_FAIL(KG_EINVAL, "kg_probe_roofline: null argument");
    CU(cudaSetDevice(ctx->device));
    uint4* buf = nullptr;
    CU(cudaMalloc(&buf, bytes));
    CU(cudaMemsetAsync(buf, 0x5A, bytes, ctx->stream));
    int rc = roofline_run(ctx, buf, bytes / 32, n_loads, tpb, inflight, out);
    cudaFree(buf);
    return rc;
}
extern "C" int kg_probe_roofline_table(kg_context* ctx, const kg_table* table, uint64_t n_loads, int tpb, int inflight, double* out) {
    if (!ctx || !table || !out) KG_FAIL(KG_EINVAL, "kg_probe_roofline_table: null argument");
    CU(cudaSetDevice(ctx->device));
    return roofline_run(ctx, table->d_lines, ((uint64_t)table->num_buckets + KG_TAIL_BUCKETS) * 4, n_loads, tpb, inflight, out);
}
