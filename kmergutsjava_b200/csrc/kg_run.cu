// kg_run.cu -- the device pipeline: residue stream -> encode+probe -> ordered hits -> run FSM -> CALL / OTU records.
//
// Replaces, for a whole batch of sequences at once:
//   prepareQuery  KGJ:1051-1074   (k_patch_aa / k_vlen + k_translate)
//   addKmers      KGJ:900-922     (k_probe: window enumeration, base-20 encoding)
//   sort + lookup KGJ:1076-1095, 944-1034 (k_probe: one 32-byte sector per k-mer instead of a sort-merge join)
//   gatherHits / processSetOfHits KGJ:457-514, 385-455 (k_fsm)
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <string.h>

#include <algorithm>
#include <functional>
#include <thread>
#include <vector>
#include <chrono>

#include "kg_device.cuh"
#include "kg_fsm.cuh"
#include "kg_internal.h"

namespace {

constexpr int PT = 8;             // positions per thread
#ifndef KG_PROBE_BLK
#define KG_PROBE_BLK 128
#endif
constexpr int PROBE_BLK = KG_PROBE_BLK; // threads per block
constexpr int TILE = PT * PROBE_BLK;
constexpr int TILE_SHIFT = TILE == 4096 ? 12 : TILE == 2048 ? 11 : TILE == 1024 ? 10 : 9;
static_assert(TILE == (1 << TILE_SHIFT), "tile size");

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

// offsets of a slice of a larger batch -> offsets relative to the slice (scale = 8 turns group offsets of the packed
// form into residue offsets)
__global__ void k_rebase(uint64_t* __restrict__ off, uint64_t n1, uint64_t base, uint64_t scale) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n1) off[i] = (off[i] - base) * scale;
}

// ---------------------------------------------------------------------------------------------------------------
// packed protein input (include/kmerguts.h, kg_pack_aa): 8 residue codes in 5 bytes.  One thread turns one group back into
// 8 stream bytes that the encoder's LUT reads exactly like the original characters: codes 0..19 -> the letter,
// 20 (anything toAminoAcidOff maps to 20, KGJ:111-175) -> 'X', 31 (padding after the last residue) -> 0 (separator).
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_unpack_aa(const uint8_t* __restrict__ packed, uint64_t ngroups, uint2* __restrict__ stream8) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    const uint8_t* p = packed + 5 * g;
    uint64_t bits = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) bits |= (uint64_t)p[i] << (8 * i);
    uint32_t w[2] = {0u, 0u};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t c = (uint32_t)(bits >> (5 * i)) & 31u;
        const uint32_t ch = c < 20u ? (uint32_t)"ACDEFGHIKLMNPQRSTVWY"[c] : (c == 31u ? 0u : (uint32_t)'X');
        w[i >> 2] |= ch << (8 * (i & 3));
    }
    stream8[g] = make_uint2(w[0], w[1]);
}
// the padded layout of the packed form: the last residue of a protein is the last non-zero byte of its (multiple-of-8) span
// ---------------------------------------------------------------------------------------------------------------
// packed 6-frame input (kg_pack_dna / kg_run_packed_dna): 2 bits per nucleotide (dnaChar, KGJ:294-318: aA 0, cC 1, gG 2, tTuU 3),
// four per byte, every contig starting on a byte; the rare other characters (dnaChar 4) travel as a sorted list of
// positions.  Unpacked into the 1-byte-per-nucleotide stream the translation kernel reads: 'A' 'C' 'G' 'T', and 'N' at the
// listed positions -- translate() only ever looks at dnaChar() of a nucleotide, so the results cannot differ.
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_unpack_dna(const uint8_t* __restrict__ packed, uint64_t nbytes, const uint64_t* __restrict__ boff, uint64_t boff0,
                             const uint64_t* __restrict__ off, uint64_t n, uint8_t* __restrict__ seq) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nbytes) return;
    uint64_t lo = 0, hi = n; // contig c with boff[c] - boff0 <= q < boff[c + 1] - boff0 (empty contigs own no byte)
    while (hi - lo > 1) {
        const uint64_t mid = (lo + hi) >> 1;
        if (boff[mid] - boff0 <= q) lo = mid;
        else hi = mid;
    }
    const uint64_t r = (q - (boff[lo] - boff0)) * 4, len = off[lo + 1] - off[lo];
    const uint32_t b = packed[q];
    uint8_t* dst = seq + off[lo] + r; // off is already relative to the slice
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (r + k < len) dst[k] = (uint8_t)((0x54474341u >> (8 * ((b >> (2 * k)) & 3u))) & 0xFFu); // "ACGT"
}
__global__ void k_mark_dna_exceptions(const uint64_t* __restrict__ exc, uint64_t n, uint64_t base, uint8_t* __restrict__ seq) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) seq[exc[i] - base] = 'N';
}

__global__ void k_patch_aa_padded(uint8_t* __restrict__ seq, const uint64_t* __restrict__ off, uint64_t n) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint64_t a = off[s];
    uint64_t b = off[s + 1];
    for (int k = 0; k < 9 && b > a && seq[b - 1] == 0; k++) b--;
    if (b > a) seq[b - 1] = 0;
}

// ---------------------------------------------------------------------------------------------------------------
// OTU-COUNTS records for the trip home: most proteins carry 0-2 of the 5 entries, so kg_run copies one count byte per
// sequence plus the used (count, oI) pairs instead of 44 bytes per sequence.
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_otu_counts(const kg_otu* __restrict__ otus, uint64_t n, uint32_t* __restrict__ cnt, uint8_t* __restrict__ cnt8,
                             const unsigned long long* __restrict__ ctr) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n) return;
    // a pass whose hit buffers overflowed is repeated by the host: the FSM has not written its records then
    const uint32_t c = (s < n && !ctr[KG_CTR_OVERFLOW]) ? min((uint32_t)otus[s].n, (uint32_t)KG_OI_BUFSZ) : 0u;
    cnt[s] = c;
    if (s < n) cnt8[s] = (uint8_t)c;
}
__global__ void k_otu_pack(const kg_otu* __restrict__ otus, uint64_t n, const uint32_t* __restrict__ first, kg_otu_entry* __restrict__ out) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t o = first[s], c = first[s + 1] - o;
    for (uint32_t k = 0; k < c; k++) {
        kg_otu_entry e;
        e.count = otus[s].count[k];
        e.oI = otus[s].oI[k];
        out[o + k] = e;
    }
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

// One thread per 4-byte word of the translated stream (four codons of one frame in, one aligned word out: residues, then
// the terminator / padding zeros, so the stream needs no memset); a block makes TR_PASSES x 1024 consecutive residues.  Codon and
// amino-acid tables sit in shared memory.  Forward frame f: codon j starts at nucleotide f+3j (KGJ:323-338).  Reverse
// frames are frames of the reverse complement (KGJ:1068-1071): its base q is compl(seq[L-1-q]), and compl() maps ACGTU
// (either case) onto the complementary base and never maps anything else onto ACGTU, so the complemented code is simply
// 3 - code.
// Fast path (all but a few hundred blocks): the block's residues lie in ONE virtual protein, so thread 0 resolves
// (protein, contig, frame, strand) once, the 3072 nucleotides the block needs -- one contiguous span of the contig for
// either strand -- arrive as coalesced 16-byte loads and are turned into 2-bit codes in shared memory, and every thread
// builds its four codons from twelve shared-memory bytes with 32-bit arithmetic only.
constexpr int TR_BLK = 256;                    // threads = words per block
constexpr int TR_SPAN = TR_BLK * 4 * 3;        // nucleotides behind the residues of one pass
constexpr int TR_PASSES = 8;                   // passes per block: the block's set-up (a binary search, divisions) is paid once
__device__ __forceinline__ uint32_t translate_word_slow(const uint8_t* __restrict__ seq, const uint64_t* __restrict__ off,
                                                        const uint64_t* __restrict__ voff, uint64_t nv, uint64_t v, uint64_t g,
                                                        const uint8_t* s_nt, const char* s_code) {
    while (v + 1 < nv && voff[v + 1] <= g) v++;
    const uint64_t sidx = v / 6;
    const uint32_t k = (uint32_t)(v % 6), f = k % 3;
    const bool rev = k >= 3;
    const uint64_t base = off[sidx], L = off[sidx + 1] - base;
    const uint64_t nk = L >= f + 3 ? (L - f) / 3 : 0;
    const uint64_t j0 = g - voff[v];
    uint32_t word = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const uint64_t j = j0 + r;
        if (j < nk) {
            const uint64_t p = f + 3 * j; // codon start in strand coordinates
            int c1, c2, c3;
            if (!rev) {
                c1 = s_nt[seq[base + p]];
                c2 = s_nt[seq[base + p + 1]];
                c3 = s_nt[seq[base + p + 2]];
            } else {
                c1 = s_nt[seq[base + L - 1 - p]];
                c2 = s_nt[seq[base + L - 2 - p]];
                c3 = s_nt[seq[base + L - 3 - p]];
                if ((c1 | c2 | c3) < 4) {
                    c1 = 3 - c1;
                    c2 = 3 - c2;
                    c3 = 3 - c3;
                }
            }
            const uint32_t aa = ((c1 | c2 | c3) < 4) ? (uint32_t)(uint8_t)s_code[c1 * 16 + c2 * 4 + c3] : (uint32_t)'x';
            word |= aa << (8 * r);
        }
    }
    return word;
}

__global__ __launch_bounds__(TR_BLK) void k_translate(const uint8_t* __restrict__ seq, const uint64_t* __restrict__ off, uint64_t n,
                                                      const uint64_t* __restrict__ voff, uint64_t nv, uint64_t vwords,
                                                      uint32_t* __restrict__ vseq_words) {
    __shared__ uint8_t s_nt[256];   // dnaChar, KGJ:294-318
    __shared__ char s_code[64];     // GENETIC_CODE, KGJ:88-93
    __shared__ __align__(16) uint8_t s_span[TR_SPAN + 32]; // nucleotide codes of one pass (16-byte aligned chunks)
    __shared__ uint64_t s_v0, s_base, s_L, s_p0; // first virtual protein of the block; its contig; strand coordinate of the block's first codon
    __shared__ uint32_t s_fast, s_j0, s_nk, s_rev, s_f;
    const int tid = threadIdx.x;
    s_nt[tid] = (uint8_t)dna_code((uint8_t)tid);
    if (tid < 64) s_code[tid] = c_genetic_code[tid];
    const uint64_t t0 = (uint64_t)blockIdx.x * (TR_BLK * TR_PASSES);
    if (tid == 0) { // the only binary search, divisions and dependent offset loads of the block
        const uint64_t g0 = 4 * t0;
        const uint64_t v = seq_of(voff, nv, g0);
        s_v0 = v;
        const uint64_t gend = 4 * min((uint64_t)(t0 + TR_BLK * TR_PASSES), vwords); // one past the block's last residue
        uint32_t fast = v < nv && voff[v + 1] >= gend;
        if (fast) {
            const uint64_t sidx = v / 6;
            const uint32_t k = (uint32_t)(v % 6), f = k % 3;
            const uint64_t base = off[sidx], L = off[sidx + 1] - base;
            const uint64_t nk = L >= f + 3 ? (L - f) / 3 : 0;
            const uint64_t j0 = g0 - voff[v];
            s_base = base;
            s_L = L;
            s_f = f;
            s_p0 = f + 3 * j0;
            s_j0 = (uint32_t)min(j0, (uint64_t)0xFFFFFFFFu);
            s_nk = (uint32_t)min(nk, (uint64_t)0xFFFFFFFFu);
            s_rev = k >= 3;
            if (j0 > 0xFFF00000ull || nk > 0xFFF00000ull) fast = 0;
        }
        s_fast = fast;
    }
    __syncthreads();
    if (!s_fast) {
        for (int pass = 0; pass < TR_PASSES; pass++) {
            const uint64_t t = t0 + (uint64_t)pass * TR_BLK + tid;
            if (t < vwords) vseq_words[t] = translate_word_slow(seq, off, voff, nv, s_v0, 4 * t, s_nt, s_code);
        }
        return;
    }
    const uint64_t base = s_base, L = s_L;
    const uint32_t nk = s_nk;
    const bool rev = s_rev;
    const uint64_t pend = (uint64_t)s_f + 3ull * nk; // one past the frame's last whole codon, strand coordinates
    for (int pass = 0; pass < TR_PASSES; pass++) {
        // strand coordinates [p0, p1) of this pass's codons
        const uint64_t p0 = s_p0 + (uint64_t)pass * TR_SPAN, p1 = min((uint64_t)(p0 + TR_SPAN), pend);
        uint32_t first = 0, nch = 0;
        uint64_t lo_addr = 0;
        if (p1 > p0) {
            // forward: seq[base+p0 .. base+p1); reverse: seq[base+L-p1 .. base+L-p0), read downwards
            const uint64_t a = !rev ? base + p0 : base + L - p1;
            lo_addr = a & ~15ull;
            first = !rev ? (uint32_t)(a - lo_addr) : (uint32_t)(base + L - 1 - p0 - lo_addr);
            nch = (uint32_t)((a - lo_addr + (p1 - p0) + 15) / 16); // whole chunks: at most 15 bytes past the contig, inside the padding
        }
        if (pass) __syncthreads(); // the previous pass has finished reading s_span
        for (int c = tid; c < (int)nch; c += TR_BLK) {
            const uint4 w = *reinterpret_cast<const uint4*>(seq + lo_addr + 16ull * c);
            const uint32_t in[4] = {w.x, w.y, w.z, w.w};
            uint32_t out[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
                out[q] = (uint32_t)s_nt[in[q] & 0xFFu] | ((uint32_t)s_nt[(in[q] >> 8) & 0xFFu] << 8) | ((uint32_t)s_nt[(in[q] >> 16) & 0xFFu] << 16) |
                         ((uint32_t)s_nt[in[q] >> 24] << 24);
            *reinterpret_cast<uint4*>(s_span + 16 * c) = make_uint4(out[0], out[1], out[2], out[3]);
        }
        __syncthreads();
        const uint64_t t = t0 + (uint64_t)pass * TR_BLK + tid;
        if (t >= vwords) continue;
        const uint32_t jl = 4u * (uint32_t)tid; // first residue of this thread, relative to the pass
        const uint32_t j0 = s_j0 + (uint32_t)pass * (4u * TR_BLK) + jl;
        uint32_t word = 0;
        if (!rev) {
            const uint8_t* p = s_span + first + 3u * jl;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                if (j0 + r < nk) {
                    const int c1 = p[3 * r], c2 = p[3 * r + 1], c3 = p[3 * r + 2];
                    const uint32_t aa = ((c1 | c2 | c3) < 4) ? (uint32_t)(uint8_t)s_code[c1 * 16 + c2 * 4 + c3] : (uint32_t)'x';
                    word |= aa << (8 * r);
                }
            }
        } else {
            const uint8_t* p = s_span + first - 3u * jl; // codon r: p[-3r], p[-3r-1], p[-3r-2]
#pragma unroll
            for (int r = 0; r < 4; r++) {
                if (j0 + r < nk) {
                    const int c1 = *(p - 3 * r), c2 = *(p - 3 * r - 1), c3 = *(p - 3 * r - 2);
                    const uint32_t aa = ((c1 | c2 | c3) < 4) ? (uint32_t)(uint8_t)s_code[(3 - c1) * 16 + (3 - c2) * 4 + (3 - c3)] : (uint32_t)'x';
                    word |= aa << (8 * r);
                }
            }
        }
        vseq_words[t] = word;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// encode + probe.  A block (128 threads, 6 resident per SM) owns a 1024-position tile of the residue stream and works in four phases:
//   A  encode   one thread owns PT = 8 consecutive positions: 16 residue bytes (8 + 7 halo + 1) arrive as two coalesced
//               8-byte loads, codes come from a 256-byte shared-memory LUT, 4-mer partial products give
//               every 8-mer in one 64-bit IMAD (first residue most significant, KGJ:274-282)
//      filter   every valid window tests its two bits in the L2-resident Bloom prefilter: 8 independent 8-byte loads
//               in flight per thread; ~70 % of the windows end here without touching DRAM
//   B  compact  the surviving (key, position) pairs are packed into a shared-memory queue (block-wide scan)
//   C  probe    the queue is probed DENSELY: every lane busy, PROBE_U independent 256-bit sector loads in flight per
//               thread.  A hit's 16-byte payload is read at once -- it sits in the 128-byte line the probe has just
//               pulled into L2 -- and parked in shared memory at the hit's position
//   D  emit     hits leave in tile order: block-wide scan of per-thread hit counts, ONE global atomic per tile to claim
//               an output chunk, (position, payload) records written by the owning threads.
// The run FSM reads the chunks through (tile_base, tile_cnt); no global sort or re-ordering pass is needed.
// ---------------------------------------------------------------------------------------------------------------
#ifndef KG_PROBE_OCC
#define KG_PROBE_OCC 6
#endif
#ifndef KG_PROBE_U
#define KG_PROBE_U 4
#endif
constexpr int PROBE_U = KG_PROBE_U;
constexpr size_t PROBE_SMEM_PSTAGE = (size_t)TILE * sizeof(int4);                 // payload parked per position
constexpr size_t PROBE_SMEM_QUEUE = (size_t)TILE * sizeof(unsigned long long);    // survivor queue
constexpr size_t PROBE_SMEM = PROBE_SMEM_PSTAGE + PROBE_SMEM_QUEUE;               // 96 KiB: needs the opt-in limit

// exclusive scan of one value per thread across the block; returns the block total through *total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t mine, uint32_t* warp_tot, uint32_t* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    uint32_t before = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < PROBE_BLK / 32; w++) {
        const uint32_t t = warp_tot[w];
        if (w < wid) before += t;
        tot += t;
    }
    *total = tot;
    return before + incl - mine;
}

// Windows of one thread: PT consecutive positions from p0 (+ 7 bytes of halo).  q[i] = base-20 value of residues
// i..i+3, so the key of window i is q[i] * 20^4 + q[i+4] (first residue most significant, KGJ:274-292); bit i of the
// result is set when all eight residues of window i are amino-acid letters (KGJ:111-175, 283-289).
__device__ __forceinline__ uint32_t encode_windows(const uint8_t* __restrict__ stream, uint32_t p0, uint32_t vtotal,
                                                   const uint8_t* lut, uint32_t (&q)[PT + 4]) {
    constexpr int NB = PT + 8; // residue bytes a thread reads: its PT positions + 7 bytes of halo (+1 unused)
    uint32_t wv[NB / 4];
#pragma unroll
    for (int i = 0; i < NB / 8; i++) {
        const uint2 t = *reinterpret_cast<const uint2*>(stream + p0 + 8 * i);
        wv[2 * i] = t.x;
        wv[2 * i + 1] = t.y;
    }
    uint32_t c[NB];
    uint32_t bad = 0;
#pragma unroll
    for (int i = 0; i < NB; i++) {
        c[i] = lut[(wv[i >> 2] >> (8 * (i & 3))) & 0xFFu];
        bad |= (uint32_t)(c[i] >= 20u) << i;
    }
    const uint32_t left = vtotal - p0; // bytes of this thread's NB that exist
    if (left < NB) bad |= ~0u << left;
#pragma unroll
    for (int i = 0; i < PT + 4; i++) q[i] = ((c[i] * 20u + c[i + 1]) * 20u + c[i + 2]) * 20u + c[i + 3];
    uint32_t valid = 0;
#pragma unroll
    for (int i = 0; i < PT; i++) valid |= (uint32_t)(((bad >> i) & 0xFFu) == 0u) << i;
    return valid;
}

// Prefilter: which of the valid windows may be in the table (all PT filter words are in flight together).
__device__ __forceinline__ uint32_t filter_windows(const KgTableView& tab, const uint32_t (&q)[PT + 4], uint32_t valid,
                                                   uint32_t flags, uint64_t pol_keep) {
    unsigned long long fw[PT];
#pragma unroll
    for (int i = 0; i < PT; i++) {
        fw[i] = 0;
        if ((valid >> i) & 1u) {
            const uint64_t h = kg_fhash1((uint64_t)q[i] * 160000ull + q[i + 4]);
            const uint32_t w = kg_filter_word(h, tab.filter_words);
            fw[i] = (flags & 2u) ? __ldg(tab.filter + w) : kg_load_filter_word(tab.filter, w, pol_keep);
        }
    }
    uint32_t pass = 0;
#pragma unroll
    for (int i = 0; i < PT; i++) {
        const unsigned long long fm = kg_filter_mask(kg_fhash1((uint64_t)q[i] * 160000ull + q[i + 4]));
        pass |= (uint32_t)((fw[i] & fm) == fm) << i;
    }
    return pass & valid;
}

// Dense probing of a block's survivor queue (entries: key | local position << 35): every thread keeps PROBE_U bucket
// lines in flight; a hit parks its payload at the local position and sets the position's bit.
__device__ __forceinline__ void probe_queue(const KgTableView& tab, const unsigned long long* queue, uint32_t nsurv, int4* pstage,
                                            uint32_t* hitbits, uint64_t pol_stream) {
    const int tid = threadIdx.x;
    for (uint32_t k0 = 0; k0 < nsurv; k0 += PROBE_BLK * PROBE_U) {
        uint64_t key[PROBE_U];
        uint32_t lp[PROBE_U], bkt[PROBE_U], slot[PROBE_U];
        bool on[PROBE_U];
        KgBucket bk[PROBE_U];
        int4 pl[PROBE_U];
#pragma unroll
        for (int u = 0; u < PROBE_U; u++) {
            const uint32_t k = k0 + u * PROBE_BLK + tid;
            on[u] = k < nsurv;
            const unsigned long long e = on[u] ? queue[k] : 0ull;
            key[u] = e & 0x7FFFFFFFFull;
            lp[u] = (uint32_t)(e >> 35);
            bkt[u] = kg_home_bucket(key[u], tab.num_buckets);
        }
#pragma unroll
        for (int u = 0; u < PROBE_U; u++)
            if (on[u]) bk[u] = kg_load_bucket_hint(tab.lines, bkt[u], pol_stream);
#pragma unroll
        for (int u = 0; u < PROBE_U; u++) {
            slot[u] = 0xFFFFFFFFu;
            if (!on[u]) continue;
            const uint32_t m = kg_bucket_match(bk[u], key[u]);
            if (m) slot[u] = bkt[u] * KG_BUCKET_KEYS + (__ffs(m) - 1);
            else if (bk[u].w[6] & KG_W6_FLAG) slot[u] = kg_lookup_from(tab, key[u], bkt[u] + 1); // rare second line
        }
#pragma unroll
        for (int u = 0; u < PROBE_U; u++)
            if (slot[u] != 0xFFFFFFFFu) pl[u] = kg_load_payload(tab.lines, slot[u]);
#pragma unroll
        for (int u = 0; u < PROBE_U; u++)
            if (slot[u] != 0xFFFFFFFFu) {
                pstage[lp[u]] = pl[u];
                atomicOr(&hitbits[lp[u] >> 5], 1u << (lp[u] & 31));
            }
    }
}

__global__ __launch_bounds__(PROBE_BLK, KG_PROBE_OCC) void k_probe(const uint8_t* __restrict__ stream, uint32_t vtotal, KgTableView tab,
                                                        uint32_t* __restrict__ chunk_pos, int4* __restrict__ chunk_payload,
                                                        uint32_t hit_cap, uint32_t* __restrict__ tile_base,
                                                        uint32_t* __restrict__ tile_cnt, unsigned long long* __restrict__ ctr,
                                                        uint32_t flags) {
    extern __shared__ int4 smem_dyn[];
    int4* pstage = smem_dyn;                                                        // [TILE]
    unsigned long long* queue = reinterpret_cast<unsigned long long*>(smem_dyn + TILE); // [TILE]
    __shared__ uint8_t lut[256];
    __shared__ uint32_t hitbits[TILE / 32];
    __shared__ uint32_t warp_a[PROBE_BLK / 32], warp_b[PROBE_BLK / 32], warp_kmers[PROBE_BLK / 32];
    __shared__ uint32_t s_base;

    const int tid = threadIdx.x;
    for (int i = tid; i < 256; i += PROBE_BLK) lut[i] = (i >= 'A' && i <= 'Z') ? c_aa_code[i - 'A'] : 20;
    if (tid < TILE / 32) hitbits[tid] = 0;
    __syncthreads();

    const uint64_t pol_keep = kg_policy_evict_last();
    const uint64_t pol_stream = (flags & 1u) ? kg_policy_evict_normal() : kg_policy_evict_first();
    const uint32_t p0 = blockIdx.x * (uint32_t)TILE + (uint32_t)tid * PT;

    // ---- A: encode + prefilter ----
    uint32_t pass = 0, nk = 0;
    uint32_t q[PT + 4];
    if (p0 < vtotal) {
        const uint32_t valid = encode_windows(stream, p0, vtotal, lut, q);
        nk = __popc(valid); // lookups = windows the reference enumerates (KGJ:912-921)
        pass = valid;
        if (tab.filter_words) pass = filter_windows(tab, q, valid, flags, pol_keep);
    }

    // ---- B: survivors -> shared-memory queue, tile order ----
    uint32_t nsurv;
    uint32_t qo = block_excl_scan(__popc(pass), warp_a, &nsurv);
#pragma unroll
    for (int i = 0; i < PT; i++)
        if ((pass >> i) & 1u) queue[qo++] = ((uint64_t)q[i] * 160000ull + q[i + 4]) | ((unsigned long long)(tid * PT + i) << 35);
    const uint32_t wk = __reduce_add_sync(0xFFFFFFFFu, nk);
    if ((tid & 31) == 0) warp_kmers[tid >> 5] = wk;
    __syncthreads();

    // ---- C: dense probing of the queue ----
    probe_queue(tab, queue, nsurv, pstage, hitbits, pol_stream);
    __syncthreads();

    // ---- D: hits out, tile order ----
    const uint32_t hitmask = (hitbits[(tid * PT) >> 5] >> ((tid * PT) & 31)) & ((1u << PT) - 1u);
    uint32_t total;
    const uint32_t ho = block_excl_scan(__popc(hitmask), warp_b, &total);
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
        uint32_t o = base + ho;
        uint32_t m = hitmask;
        while (m) {
            const int i = __ffs(m) - 1;
            m &= m - 1;
            chunk_pos[o] = p0 + (uint32_t)i;
            chunk_payload[o] = pstage[tid * PT + i];
            o++;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// encode + probe in TWO PASSES, one per half of the key space (tables built with KG_FILTER_HALVES=1).
//
// The fused kernel's excess DRAM traffic is bucket lines fetched for false positives of the prefilter (10 of 22 GB on
// configs[1]) and prefilter words that fell out of L2 (4 GB); the false-positive rate is set by the filter bits that fit the
// L2 set-aside AT ONE TIME (2.7 bits per key: 0.28).  Here every key belongs to one of two halves (top bit of its filter
// hash) and each half has a filter of that size of its own (5.4 bits per key: ~0.10).  Pass 0 runs the whole batch against
// half 0 -- every window is encoded, the windows of the other half are dropped before they cost a load -- with filter 0 in
// the set-aside, pass 1 the same against half 1.  Unlike the cascade (profiles/r02_probe_cascade.md) no survivor queue goes
// through global memory and no window is filtered twice; the price is the second encode.  Pass 0 leaves its hits in
// per-tile chunks of its own; pass 1 loads the chunk of its tile into the shared-memory staging area next to its own
// hits (a position belongs to exactly one half) and emits the merged tile in position order, so everything downstream
// sees exactly what k_probe would have written.
// ---------------------------------------------------------------------------------------------------------------
template <int HALF>
__global__ __launch_bounds__(PROBE_BLK, KG_PROBE_OCC) void k_probe_half(const uint8_t* __restrict__ stream, uint32_t vtotal, KgTableView tab,
                                                                        uint32_t* __restrict__ chunk_pos, int4* __restrict__ chunk_payload,
                                                                        uint32_t hit_cap, uint32_t* __restrict__ tile_base,
                                                                        uint32_t* __restrict__ tile_cnt, unsigned long long* __restrict__ ctr,
                                                                        const uint32_t* __restrict__ prev_pos, const int4* __restrict__ prev_payload,
                                                                        const uint32_t* __restrict__ prev_base, const uint32_t* __restrict__ prev_cnt,
                                                                        uint32_t flags) {
    extern __shared__ int4 smem_dyn[];
    int4* pstage = smem_dyn;
    unsigned long long* queue = reinterpret_cast<unsigned long long*>(smem_dyn + TILE);
    __shared__ uint8_t lut[256];
    __shared__ uint32_t hitbits[TILE / 32];
    __shared__ uint32_t warp_a[PROBE_BLK / 32], warp_b[PROBE_BLK / 32], warp_kmers[PROBE_BLK / 32];
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x;
    for (int i = tid; i < 256; i += PROBE_BLK) lut[i] = (i >= 'A' && i <= 'Z') ? c_aa_code[i - 'A'] : 20;
    if (tid < TILE / 32) hitbits[tid] = 0;
    __syncthreads();
    const uint64_t pol_keep = kg_policy_evict_last();
    const uint64_t pol_stream = (flags & 1u) ? kg_policy_evict_normal() : kg_policy_evict_first();
    const uint32_t p0 = blockIdx.x * (uint32_t)TILE + (uint32_t)tid * PT;
    const unsigned long long* filt = HALF ? tab.filter2 : tab.filter;

    // pass 1: the hits pass 0 found in this tile (issued first: their latency hides behind the encode)
    uint32_t pv_cnt = 0, pv_base = 0, pv_pos = 0;
    int4 pv_pl = make_int4(0, 0, 0, 0);
    if (HALF) {
        pv_cnt = prev_cnt[blockIdx.x];
        pv_base = prev_base[blockIdx.x];
        if ((uint32_t)tid < pv_cnt && pv_base != 0xFFFFFFFFu) {
            pv_pos = __ldcs(prev_pos + pv_base + tid);
            pv_pl = __ldcs(prev_payload + pv_base + tid);
        }
    }

    // ---- A: encode + this half's prefilter ----
    uint32_t pass = 0, nk = 0;
    uint32_t q[PT + 4];
    if (p0 < vtotal) {
        const uint32_t valid = encode_windows(stream, p0, vtotal, lut, q);
        uint32_t mine = 0;
        unsigned long long fw[PT];
#pragma unroll
        for (int i = 0; i < PT; i++) {
            fw[i] = 0;
            if ((valid >> i) & 1u) {
                const uint64_t h = kg_fhash1((uint64_t)q[i] * 160000ull + q[i + 4]);
                if (kg_filter_half(h) == (uint32_t)HALF) {
                    mine |= 1u << i;
                    fw[i] = kg_load_filter_word(filt, kg_filter_half_word(h, tab.filter_words), pol_keep);
                }
            }
        }
        nk = __popc(mine); // every valid window is counted in exactly one of the two passes
#pragma unroll
        for (int i = 0; i < PT; i++) {
            const unsigned long long fm = kg_filter_mask(kg_fhash1((uint64_t)q[i] * 160000ull + q[i + 4]));
            pass |= (uint32_t)((fw[i] & fm) == fm) << i;
        }
        pass &= mine;
    }
    // ---- B ----
    uint32_t nsurv;
    uint32_t qo = block_excl_scan(__popc(pass), warp_a, &nsurv);
#pragma unroll
    for (int i = 0; i < PT; i++)
        if ((pass >> i) & 1u) queue[qo++] = ((uint64_t)q[i] * 160000ull + q[i + 4]) | ((unsigned long long)(tid * PT + i) << 35);
    const uint32_t wk = __reduce_add_sync(0xFFFFFFFFu, nk);
    if ((tid & 31) == 0) warp_kmers[tid >> 5] = wk;
    __syncthreads();
    // ---- C ----
    probe_queue(tab, queue, nsurv, pstage, hitbits, pol_stream);
    if (HALF && pv_base != 0xFFFFFFFFu) { // merge pass 0's hits of this tile (disjoint positions)
        const uint32_t t0 = blockIdx.x * (uint32_t)TILE;
        if ((uint32_t)tid < pv_cnt) {
            const uint32_t lp = pv_pos - t0;
            pstage[lp] = pv_pl;
            atomicOr(&hitbits[lp >> 5], 1u << (lp & 31));
        }
        for (uint32_t e = tid + PROBE_BLK; e < pv_cnt; e += PROBE_BLK) { // tiles with more than 128 hits in pass 0
            const uint32_t lp = __ldcs(prev_pos + pv_base + e) - t0;
            pstage[lp] = __ldcs(prev_payload + pv_base + e);
            atomicOr(&hitbits[lp >> 5], 1u << (lp & 31));
        }
    }
    __syncthreads();
    // ---- D ----
    const uint32_t hitmask = (hitbits[(tid * PT) >> 5] >> ((tid * PT) & 31)) & ((1u << PT) - 1u);
    uint32_t total;
    const uint32_t ho = block_excl_scan(__popc(hitmask), warp_b, &total);
    if (tid == 0) {
        uint32_t kmers = 0;
#pragma unroll
        for (int w = 0; w < PROBE_BLK / 32; w++) kmers += warp_kmers[w];
        if (kmers) atomicAdd(&ctr[KG_CTR_KMERS], (unsigned long long)kmers);
        unsigned long long base = 0;
        if (total) base = atomicAdd(&ctr[HALF ? KG_CTR_HITS : KG_CTR_CLAIM], (unsigned long long)total); // pass 0 claims from its own counter
        uint32_t b32 = 0xFFFFFFFFu;
        if (base + total <= (unsigned long long)hit_cap) b32 = (uint32_t)base;
        else ctr[KG_CTR_OVERFLOW] = 1ull;
        s_base = b32;
        tile_base[blockIdx.x] = b32;
        tile_cnt[blockIdx.x] = total;
    }
    __syncthreads();
    const uint32_t base = s_base;
    if (base != 0xFFFFFFFFu && hitmask) {
        uint32_t o = base + ho;
        uint32_t m = hitmask;
        while (m) {
            const int i = __ffs(m) - 1;
            m &= m - 1;
            chunk_pos[o] = p0 + (uint32_t)i;
            chunk_payload[o] = pstage[tid * PT + i];
            o++;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// encode + probe as a CASCADE of three kernels (the default when the table carries a second prefilter).
//
// Why: the fused kernel above keeps one 64 MiB prefilter in L2 WHILE 15 GB of bucket lines stream through the same L2;
// ncu (profiles/r01) shows 10 % of its filter loads missing L2 (4 GB of DRAM reads), 30 % false positives (9.5 GB of
// bucket lines fetched for nothing) and more resident blocks evicting the filter.  L2 capacity bounds the bits per key
// that can be resident AT ONE TIME -- so the stages take turns:
//   k_filter    A  encode every window + first prefilter (L2-resident, nothing else competes)  -> survivors (key, local
//                  position) of the tile, tile order, into the tile's own 8 KB slice of a global queue
//   k_refilter  B  survivors re-tested against a SECOND, independently hashed prefilter that now has the L2 to itself;
//                  compacted in place, order kept.  Two 2.7-bit filters in sequence: false positives 0.30 -> ~0.09
//   k_probe2    C  what is left (hits + ~9 % of the misses) fetches its 128-byte bucket line: pure DRAM streaming of
//                  random lines, no filter to protect, so occupancy is free; hits leave in tile order as before.
// The queue costs 8 B per survivor written and read (3 GB per 3.1e8 lookups), the lines saved are ~7 GB and the filter
// misses ~4 GB.  Every stage keeps tile order, so the run FSM downstream is unchanged.
// ---------------------------------------------------------------------------------------------------------------
constexpr int CAS_BLK = 256;                 // threads of the two queue kernels (eight warps = eight tiles)

template <int BLK>
__device__ __forceinline__ uint32_t block_excl_scan_t(uint32_t mine, uint32_t* warp_tot, uint32_t* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    uint32_t before = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < BLK / 32; w++) {
        const uint32_t t = warp_tot[w];
        if (w < wid) before += t;
        tot += t;
    }
    *total = tot;
    return before + incl - mine;
}

__device__ __forceinline__ void kg_store_stream_2xu64(unsigned long long* p, unsigned long long a, unsigned long long b, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v2.u64 [%0], {%1, %2}, %3;" ::"l"(p), "l"(a), "l"(b), "l"(policy) : "memory");
}
__device__ __forceinline__ void kg_store_stream_u64(unsigned long long* p, unsigned long long a, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(a), "l"(policy) : "memory");
}
__device__ __forceinline__ unsigned long long kg_load_stream_u64(const unsigned long long* p, uint64_t policy) {
    unsigned long long v;
    asm volatile("ld.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(policy));
    return v;
}

#ifndef KG_FILTER_OCC
#define KG_FILTER_OCC 8
#endif
__global__ __launch_bounds__(PROBE_BLK, KG_FILTER_OCC) void k_filter(const uint8_t* __restrict__ stream, uint32_t vtotal, uint32_t tile0, KgTableView tab,
                                                          unsigned long long* __restrict__ queue, uint32_t* __restrict__ tile_qcnt,
                                                          unsigned long long* __restrict__ ctr, uint32_t flags) {
    const uint32_t tile = tile0 + blockIdx.x;
    __shared__ __align__(16) unsigned long long sq[TILE];
    __shared__ uint8_t lut[256];
    __shared__ uint32_t warp_a[PROBE_BLK / 32], warp_kmers[PROBE_BLK / 32];
    const int tid = threadIdx.x;
    for (int i = tid; i < 256; i += PROBE_BLK) lut[i] = (i >= 'A' && i <= 'Z') ? c_aa_code[i - 'A'] : 20;
    __syncthreads();
    const uint64_t pol_keep = kg_policy_evict_last();
    const uint64_t pol_stream = kg_policy_evict_first();
    const uint32_t p0 = tile * (uint32_t)TILE + (uint32_t)tid * PT;
    uint32_t pass = 0, nk = 0;
    uint32_t q[PT + 4];
    if (p0 < vtotal) {
        const uint32_t valid = encode_windows(stream, p0, vtotal, lut, q);
        nk = __popc(valid);
        pass = valid;
        if (tab.filter_words) pass = filter_windows(tab, q, valid, flags, pol_keep);
    }
    uint32_t nsurv;
    uint32_t qo = block_excl_scan(__popc(pass), warp_a, &nsurv);
#pragma unroll
    for (int i = 0; i < PT; i++)
        if ((pass >> i) & 1u) sq[qo++] = ((uint64_t)q[i] * 160000ull + q[i + 4]) | ((unsigned long long)(tid * PT + i) << 35);
    const uint32_t wk = __reduce_add_sync(0xFFFFFFFFu, nk);
    if ((tid & 31) == 0) warp_kmers[tid >> 5] = wk;
    __syncthreads();
    if (tid == 0) {
        uint32_t kmers = 0;
#pragma unroll
        for (int w = 0; w < PROBE_BLK / 32; w++) kmers += warp_kmers[w];
        if (kmers) atomicAdd(&ctr[KG_CTR_KMERS], (unsigned long long)kmers);
        if (nsurv) atomicAdd(&ctr[KG_CTR_SURV1], (unsigned long long)nsurv);
        tile_qcnt[tile] = nsurv;
    }
    // the tile's survivors leave as whole 16-byte stores, streaming through L2 (they must not displace the filter)
    unsigned long long* dst = queue + (size_t)tile * TILE;
    for (uint32_t i = 2u * tid; i < nsurv; i += 2u * PROBE_BLK) {
        if (i + 1 < nsurv) kg_store_stream_2xu64(dst + i, sq[i], sq[i + 1], pol_stream);
        else kg_store_stream_u64(dst + i, sq[i], pol_stream);
    }
}

// Stages B and C give every tile to ONE WARP (eight tiles per block): a tile holds only a few hundred survivors, and with a
// block per tile the chain count -> entries -> filter word / bucket line -> scan -> store is five dependent round trips
// for one load per thread (measured: 0.96 and 3.2 ms).  A lane owns CAS_U consecutive entries of a 32 x CAS_U round, so
// CAS_U independent loads per lane are in flight, and order is kept by a shuffle scan of the lanes' counts: no barrier,
// no shared memory.
constexpr int CAS_U = 8;
constexpr int CAS_ROUND = 32 * CAS_U;
__device__ __forceinline__ void kg_load_entries8(const unsigned long long* p, unsigned long long (&e)[CAS_U], uint64_t policy) {
    // 64 consecutive bytes (the slice is 8 KB aligned and a lane's offset a multiple of 64): two 256-bit loads
    asm volatile("ld.global.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(e[0]), "=l"(e[1]), "=l"(e[2]), "=l"(e[3]) : "l"(p), "l"(policy));
    asm volatile("ld.global.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(e[4]), "=l"(e[5]), "=l"(e[6]), "=l"(e[7]) : "l"(p + 4), "l"(policy));
}
__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t mine, uint32_t* total) {
    const int lane = threadIdx.x & 31;
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
    }
    *total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    return incl - mine;
}

__global__ __launch_bounds__(CAS_BLK) void k_refilter(unsigned long long* __restrict__ queue, uint32_t* __restrict__ tile_qcnt, uint32_t tile0,
                                                      uint32_t tile_end, KgTableView tab) {
    const uint32_t tile = tile0 + blockIdx.x * (CAS_BLK / 32) + (threadIdx.x >> 5);
    if (tile >= tile_end) return;
    const uint32_t cnt = tile_qcnt[tile];
    if (cnt == 0) return;
    const int lane = threadIdx.x & 31;
    unsigned long long* base = queue + (size_t)tile * TILE;
    const uint64_t pol_keep = kg_policy_evict_last();
    const uint64_t pol_stream = kg_policy_evict_first();
    uint32_t out = 0;
    for (uint32_t r0 = 0; r0 < cnt; r0 += CAS_ROUND) {
        const uint32_t i0 = r0 + (uint32_t)lane * CAS_U;
        unsigned long long e[CAS_U], fw[CAS_U];
        kg_load_entries8(base + i0, e, pol_stream); // entries past cnt are stale but inside the tile's slice
#pragma unroll
        for (int u = 0; u < CAS_U; u++) {
            const uint64_t m = kg_fhash2(e[u] & 0x7FFFFFFFFull);
            fw[u] = i0 + u < cnt ? kg_load_filter_word(tab.filter2, kg_filter2_word(m, tab.filter2_words), pol_keep) : 0ull;
        }
        uint32_t pass = 0;
#pragma unroll
        for (int u = 0; u < CAS_U; u++) {
            const unsigned long long fm = kg_filter2_mask(kg_fhash2(e[u] & 0x7FFFFFFFFull));
            pass |= (uint32_t)(i0 + u < cnt && (fw[u] & fm) == fm) << u;
        }
        uint32_t total;
        // The scan needs every lane's loaded entries, so all reads of this round are complete before the first store; the
        // stores land at or below the positions already read (compaction in place, order kept).
        uint32_t o = out + warp_excl_scan(__popc(pass), &total);
#pragma unroll
        for (int u = 0; u < CAS_U; u++)
            if ((pass >> u) & 1u) kg_store_stream_u64(base + o++, e[u], pol_stream);
        out += total;
    }
    if (lane == 0) tile_qcnt[tile] = out;
}

// Stage C.  Warp-cooperative probing: FOUR LANES fetch one 128-byte bucket line, one 32-byte sector each, in the same load
// instruction -- lane 0 of the group gets the key sector, lanes 1-3 the payload sectors.  A hit's payload is then already
// in a register of a sibling lane (four shuffles), not a second, dependent access to a line that L2 may have dropped in the
// meantime (ncu, first version with a separate payload load: 14.3 GB of DRAM reads for 68 M survivors -- the 47 M payload
// loads missed L2 and fetched their lines AGAIN; hit rate 9.7 %).  PR_U lines are in flight per group, 8 x PR_U per warp.
// The tile's hit chunk is claimed UP FRONT for all its survivors (one atomic whose round trip overlaps the line fetches;
// hits <= survivors), so hits can be written round by round in order (a ballot over the group leaders); tile_cnt gets the
// real number, the gap behind it stays unused (downstream addresses chunks through tile_base / tile_cnt only).
constexpr int PR_PASS = 8; // entries per lane and pass: 256 entries of the tile are fetched at once, then probed in rounds
__device__ __forceinline__ KgBucket kg_load_line_sector(const uint4* lines, uint32_t b, uint32_t sector, uint64_t policy) {
    KgBucket r;
    const uint4* p = lines + (size_t)KG_LINE_UINT4 * b + 2u * sector;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p), "l"(policy));
    return r;
}
// PR_U = bucket lines in flight per four-lane group (8 x PR_U per warp); a round covers 8 x PR_U consecutive entries
template <int PR_U>
__global__ __launch_bounds__(CAS_BLK) void k_probe2(const unsigned long long* __restrict__ queue, const uint32_t* __restrict__ tile_qcnt,
                                                    uint32_t tile0, uint32_t tile_end, KgTableView tab, uint32_t* __restrict__ chunk_pos,
                                                    int4* __restrict__ chunk_payload, uint32_t hit_cap, uint32_t* __restrict__ tile_base,
                                                    uint32_t* __restrict__ tile_cnt, unsigned long long* __restrict__ ctr) {
    constexpr int LPR = 32 / (8 * PR_U) ? 32 / (8 * PR_U) : 1; // (unused when a round needs more than one entry register)
    constexpr int EPR = (8 * PR_U + 31) / 32;                   // entry registers a round consumes (1 for U <= 4, 2 for U = 8)
    static_assert(PR_U == 2 || PR_U == 4 || PR_U == 8, "lines in flight per group");
    (void)LPR;
    const uint32_t tile = tile0 + blockIdx.x * (CAS_BLK / 32) + (threadIdx.x >> 5);
    if (tile >= tile_end) return;
    const uint32_t cnt = tile_qcnt[tile];
    const int lane = threadIdx.x & 31;
    if (cnt == 0) {
        if (lane == 0) {
            tile_base[tile] = 0;
            tile_cnt[tile] = 0;
        }
        return;
    }
    unsigned long long claim = 0;
    if (lane == 0) claim = atomicAdd(&ctr[KG_CTR_CLAIM], (unsigned long long)cnt);
    const unsigned long long* base = queue + (size_t)tile * TILE;
    const uint64_t pol_stream = kg_policy_evict_first();
    const int grp = lane >> 2, sub = lane & 3, lead = lane & ~3;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t out = 0, cb = 0;
    bool fits = true, have_claim = false;
    for (uint32_t p0 = 0; p0 < cnt; p0 += 32 * PR_PASS) {
        // the pass's entries, one coalesced 256-byte load per register: entry p0 + 32 j + lane sits in e[j] of `lane`
        unsigned long long e[PR_PASS];
#pragma unroll
        for (int j = 0; j < PR_PASS; j++) {
            e[j] = 0x7FFFFFFFFull; // the empty-slot pattern: matches nothing
            if (p0 + 32 * j + lane < cnt) e[j] = kg_load_stream_u64(base + p0 + 32 * j + lane, pol_stream);
        }
        // rounds of 8 x PR_U entries, in order; the entry registers shift down as they are consumed
        for (uint32_t r0 = p0; r0 < cnt && r0 < p0 + 32 * PR_PASS; r0 += 8 * PR_U) {
            const uint32_t off = (r0 - p0) & 31u; // first lane of this round inside e[0] (always 0 unless PR_U < 4)
            uint32_t mybkt[EPR];
#pragma unroll
            for (int q = 0; q < EPR; q++) mybkt[q] = kg_home_bucket(e[q] & 0x7FFFFFFFFull, tab.num_buckets);
            KgBucket bk[PR_U];
            bool valid[PR_U];
#pragma unroll
            for (int u = 0; u < PR_U; u++) {
                const int idx = u * 8 + grp;               // entry r0 + idx of the round
                const int src = (int)((off + idx) & 31u);
                const uint32_t b = __shfl_sync(0xFFFFFFFFu, mybkt[idx >> 5], src);
                valid[u] = r0 + idx < cnt;
                if (valid[u]) bk[u] = kg_load_line_sector(tab.lines, b, (uint32_t)sub, pol_stream);
            }
#pragma unroll
            for (int u = 0; u < PR_U; u++) {
                const int idx = u * 8 + grp;
                const int src = (int)((off + idx) & 31u);
                const unsigned long long eu = __shfl_sync(0xFFFFFFFFu, e[idx >> 5], src);
                const uint64_t key = eu & 0x7FFFFFFFFull;
                int s = -1;
                bool slow = false;
                if (sub == 0 && valid[u]) {
                    const uint32_t m = kg_bucket_match(bk[u], key);
                    if (m) s = __ffs(m) - 1;
                    else slow = (bk[u].w[6] & KG_W6_FLAG) != 0; // rare: the key may live in a later bucket
                }
                s = __shfl_sync(0xFFFFFFFFu, s, lead);
                // payload s sits in sector 1 + s/2 (lane lead + 1 + s/2), words 4*(s&1) .. +3
                const bool hi = (s & 1) != 0;
                const uint32_t v0 = hi ? bk[u].w[4] : bk[u].w[0], v1 = hi ? bk[u].w[5] : bk[u].w[1];
                const uint32_t v2 = hi ? bk[u].w[6] : bk[u].w[2], v3 = hi ? bk[u].w[7] : bk[u].w[3];
                const int srcl = lead + 1 + (max(s, 0) >> 1);
                int4 pl;
                pl.x = (int)__shfl_sync(0xFFFFFFFFu, v0, srcl);
                pl.y = (int)__shfl_sync(0xFFFFFFFFu, v1, srcl);
                pl.z = (int)__shfl_sync(0xFFFFFFFFu, v2, srcl);
                pl.w = (int)__shfl_sync(0xFFFFFFFFu, v3, srcl);
                bool hit = sub == 0 && s >= 0;
                if (slow) {
                    const uint32_t slot = kg_lookup_from(tab, key, kg_home_bucket(key, tab.num_buckets) + 1);
                    if (slot != 0xFFFFFFFFu) {
                        pl = kg_load_payload(tab.lines, slot);
                        hit = true;
                    }
                }
                const uint32_t ball = __ballot_sync(0xFFFFFFFFu, hit);
                if (!have_claim) { // the claim is needed only now: its latency ran under the line fetches
                    claim = __shfl_sync(0xFFFFFFFFu, claim, 0);
                    fits = claim + cnt <= (unsigned long long)hit_cap;
                    cb = (uint32_t)claim;
                    have_claim = true;
                }
                if (hit && fits) {
                    const uint32_t o = cb + out + __popc(ball & lt_mask);
                    chunk_pos[o] = tile * (uint32_t)TILE + (uint32_t)(eu >> 35);
                    chunk_payload[o] = pl;
                }
                out += __popc(ball);
            }
            // consume: after a round has used up the 32 entries of e[0] (and e[1] for PR_U = 8), shift the registers down
            if (((r0 - p0 + 8 * PR_U) & 31u) == 0 || PR_U == 8) {
#pragma unroll
                for (int j = 0; j + EPR < PR_PASS; j++) e[j] = e[j + EPR];
            }
        }
    }
    if (lane == 0) {
        if (!fits) ctr[KG_CTR_OVERFLOW] = 1ull; // chunk array too small: the host re-runs with the exact size
        tile_base[tile] = fits ? cb : 0xFFFFFFFFu;
        tile_cnt[tile] = out;
        if (out) atomicAdd(&ctr[KG_CTR_HITS], (unsigned long long)out);
    }
}

// Variant of stage C without the four-lane cooperation: a lane owns PO_U consecutive entries of a 32 x PO_U round, loads
// their key sectors (PO_U lines in flight per LANE, 32 x PO_U per warp) and reads a hit's payload with a second load from
// the same line.  `pol` = L2 eviction policy of the key-sector load: with evict_first the line is gone again before the
// payload load arrives (ncu: 80 % of the payload loads fetched their line a second time).
template <int PO_U>
__global__ __launch_bounds__(CAS_BLK) void k_probe2_own(const unsigned long long* __restrict__ queue, const uint32_t* __restrict__ tile_qcnt,
                                                        uint32_t tile0, uint32_t tile_end, KgTableView tab, uint32_t* __restrict__ chunk_pos,
                                                        int4* __restrict__ chunk_payload, uint32_t hit_cap, uint32_t* __restrict__ tile_base,
                                                        uint32_t* __restrict__ tile_cnt, unsigned long long* __restrict__ ctr, uint32_t polsel) {
    static_assert(PO_U == 4, "entries per lane");
    const uint32_t tile = tile0 + blockIdx.x * (CAS_BLK / 32) + (threadIdx.x >> 5);
    if (tile >= tile_end) return;
    const uint32_t cnt = tile_qcnt[tile];
    const int lane = threadIdx.x & 31;
    if (cnt == 0) {
        if (lane == 0) {
            tile_base[tile] = 0;
            tile_cnt[tile] = 0;
        }
        return;
    }
    unsigned long long claim = 0;
    if (lane == 0) claim = atomicAdd(&ctr[KG_CTR_CLAIM], (unsigned long long)cnt);
    const unsigned long long* base = queue + (size_t)tile * TILE;
    const uint64_t pol_stream = kg_policy_evict_first();
    const uint64_t pol_line = (polsel & 3u) == 1 ? kg_policy_evict_normal() : (polsel & 3u) == 2 ? kg_policy_evict_last() : kg_policy_evict_first();
    const bool whole = (polsel & 4u) != 0;
    uint32_t out = 0, cb = 0;
    bool fits = true;
    for (uint32_t r0 = 0; r0 < cnt; r0 += 32 * PO_U) {
        const uint32_t i0 = r0 + (uint32_t)lane * PO_U;
        unsigned long long e[PO_U];
        asm volatile("ld.global.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(e[0]), "=l"(e[1]), "=l"(e[2]), "=l"(e[3]) : "l"(base + i0), "l"(pol_stream));
        uint32_t bkt[PO_U], slot[PO_U];
        KgBucket bk[PO_U];
#pragma unroll
        for (int u = 0; u < PO_U; u++) {
            bkt[u] = kg_home_bucket(e[u] & 0x7FFFFFFFFull, tab.num_buckets);
            if (i0 + u < cnt) bk[u] = whole ? kg_load_bucket_line(tab.lines, bkt[u], pol_line) : kg_load_bucket_hint(tab.lines, bkt[u], pol_line);
        }
        uint32_t hm = 0;
#pragma unroll
        for (int u = 0; u < PO_U; u++) {
            slot[u] = 0xFFFFFFFFu;
            if (i0 + u >= cnt) continue;
            const uint64_t key = e[u] & 0x7FFFFFFFFull;
            const uint32_t m = kg_bucket_match(bk[u], key);
            if (m) slot[u] = bkt[u] * KG_BUCKET_KEYS + (__ffs(m) - 1);
            else if (bk[u].w[6] & KG_W6_FLAG) slot[u] = kg_lookup_from(tab, key, bkt[u] + 1); // rare second line
            hm |= (uint32_t)(slot[u] != 0xFFFFFFFFu) << u;
        }
        int4 pl[PO_U];
#pragma unroll
        for (int u = 0; u < PO_U; u++)
            if ((hm >> u) & 1u) pl[u] = kg_load_payload(tab.lines, slot[u]);
        if (r0 == 0) {
            claim = __shfl_sync(0xFFFFFFFFu, claim, 0);
            fits = claim + cnt <= (unsigned long long)hit_cap;
            cb = (uint32_t)claim;
        }
        uint32_t total;
        uint32_t o = cb + out + warp_excl_scan(__popc(hm), &total);
        if (fits) {
#pragma unroll
            for (int u = 0; u < PO_U; u++)
                if ((hm >> u) & 1u) {
                    chunk_pos[o] = tile * (uint32_t)TILE + (uint32_t)(e[u] >> 35);
                    chunk_payload[o] = pl[u];
                    o++;
                }
        }
        out += total;
    }
    if (lane == 0) {
        if (!fits) ctr[KG_CTR_OVERFLOW] = 1ull;
        tile_base[tile] = fits ? cb : 0xFFFFFFFFu;
        tile_cnt[tile] = out;
        if (out) atomicAdd(&ctr[KG_CTR_HITS], (unsigned long long)out);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// segment path and "-d": stitch the per-tile chunks into one position-sorted hit list.  A block takes GATHER_TILES
// consecutive tiles: their (start, chunk base) pairs go to shared memory in one coalesced load, then the block walks the
// hits of all its tiles as ONE flat range -- every lane busy, the writes coalesced -- and finds a hit's tile by a
// binary search in shared memory.  (One warp per tile was a chain of dependent loads for ~65 hits: 6 % issue-active.)
// ---------------------------------------------------------------------------------------------------------------
constexpr int GATHER_TILES = 64;
__global__ __launch_bounds__(256) void k_gather(const uint32_t* __restrict__ chunk_pos, const int4* __restrict__ chunk_payload,
                                                const uint32_t* __restrict__ tile_base, const uint32_t* __restrict__ tile_out, uint32_t ntiles,
                                                uint32_t* __restrict__ hit_pos, int4* __restrict__ hit_payload,
                                                const unsigned long long* __restrict__ ctr) {
    __shared__ uint32_t s_out[GATHER_TILES + 1], s_base[GATHER_TILES];
    const uint32_t t0 = blockIdx.x * GATHER_TILES;
    if (t0 >= ntiles || ctr[KG_CTR_OVERFLOW]) return;
    const uint32_t nt = min((uint32_t)GATHER_TILES, ntiles - t0);
    if (threadIdx.x <= nt) s_out[threadIdx.x] = tile_out[t0 + threadIdx.x];
    if (threadIdx.x < nt) s_base[threadIdx.x] = tile_base[t0 + threadIdx.x];
    __syncthreads();
    const uint32_t o0 = s_out[0], total = s_out[nt] - o0;
    for (uint32_t j = threadIdx.x; j < total; j += blockDim.x) {
        const uint32_t g = o0 + j; // rank of the hit in position order
        uint32_t lo = 0, hi = nt;  // the last tile that starts at or before g (empty tiles share their successor's start)
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_out[mid] <= g) lo = mid;
            else hi = mid;
        }
        const uint32_t src = s_base[lo] + (g - s_out[lo]);
        hit_pos[g] = chunk_pos[src];
        hit_payload[g] = chunk_payload[src];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// run FSM: one thread per sequence walks its containers in the reference's order (+0,+1,+2,-0,-1,-2) so that the
// OTU buffer sees the calls in the same order (KGJ:540-557).  The hits of a container are read straight from the
// per-tile chunks k_probe wrote: tiles in order, each chunk already sorted by position.
// lo[v] = rank of the container's first hit in the global position order (tile_out = exclusive scan of tile_cnt).
// Calls of container v go to the sparse slots [lo[v]/min_hits, lo[v+1]/min_hits): a call consumes >= min_hits counted
// hits and no hit is counted twice, so the slots cannot overflow and no second pass is needed to size them.
// ---------------------------------------------------------------------------------------------------------------
struct SparseEmit {
    KgDevCall* dst;
    __device__ __forceinline__ void operator()(int i, const KgDevCall& c) { dst[i] = c; }
};

// lo[v] = rank (in the global position order) of the first hit at or after the start of container v, found by a binary
// search inside the one tile chunk that holds that position; lo[nv] = number of hits.
__global__ void k_lo_tiles(const uint64_t* __restrict__ voff, uint64_t nv, const uint32_t* __restrict__ tile_base,
                           const uint32_t* __restrict__ tile_out, uint32_t ntiles, const uint32_t* __restrict__ chunk_pos,
                           uint32_t* __restrict__ lo, const unsigned long long* __restrict__ ctr) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v > nv) return;
    if (ctr[KG_CTR_OVERFLOW]) {
        lo[v] = 0;
        return;
    }
    const uint32_t x0 = (uint32_t)voff[v];
    const uint32_t t = x0 >> TILE_SHIFT;
    uint32_t rank = tile_out[ntiles];
    if (v < nv && t < ntiles) {
        const uint32_t o = tile_out[t], cnt = tile_out[t + 1] - o, base = tile_base[t];
        uint32_t a = 0, b = cnt;
        while (a < b) {
            const uint32_t mid = (a + b) >> 1;
            if (chunk_pos[base + mid] < x0) a = mid + 1;
            else b = mid;
        }
        rank = o + a;
    }
    lo[v] = rank;
}
// Sequences are handed to the FSM threads grouped by hit count (64 classes, largest first), so that the 32 lanes of a
// warp walk about the same number of hits (ungrouped, the average lane was active in 18 % of the issued instructions:
// ncu, r01).  A counting sort by class: histogram, then scatter with one cursor per class; the order inside a class is
// arbitrary, which is fine -- every sequence writes to places fixed by its own ranks, not by the thread that runs it.
constexpr int FSM_CLASSES = 64;
constexpr int FSM_LONG_CLASSES = 4;  // classes 0..3 = 256 hits or more
__device__ __forceinline__ uint32_t fsm_class(uint32_t hits) { // 0 = most hits
    const uint32_t c = hits < 32 ? hits : 32 + min(31u, (hits - 32) >> 3); // 0..31 exact, then steps of 8 up to 280+
    return (FSM_CLASSES - 1) - c;
}
__global__ void k_fsm_hist(const uint32_t* __restrict__ lo, uint64_t nseq, int per_seq, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[FSM_CLASSES];
    if (threadIdx.x < FSM_CLASSES) sh[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < nseq) atomicAdd(&sh[fsm_class(lo[(s + 1) * per_seq] - lo[s * per_seq])], 1u);
    __syncthreads();
    if (threadIdx.x < FSM_CLASSES && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}
__global__ void k_fsm_scatter(const uint32_t* __restrict__ lo, uint64_t nseq, int per_seq, const uint32_t* __restrict__ hist,
                              uint32_t* __restrict__ cursor, uint32_t* __restrict__ perm) {
    __shared__ uint32_t start[FSM_CLASSES], sh_cnt[FSM_CLASSES], sh_base[FSM_CLASSES];
    __shared__ uint32_t s_nlong;
    if (threadIdx.x < FSM_CLASSES) sh_cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        uint32_t acc = 0, nl = 0;
        for (int c = 0; c < FSM_CLASSES; c++) {
            start[c] = acc;
            acc += hist[c];
            if (c < FSM_LONG_CLASSES) nl = acc;
        }
        s_nlong = min(nl, (uint32_t)(nseq / 32)); // long sequences that get a warp's lane 0 to themselves
    }
    __syncthreads();
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t cls = 0, mine = 0;
    if (s < nseq) {
        cls = fsm_class(lo[(s + 1) * per_seq] - lo[s * per_seq]);
        mine = atomicAdd(&sh_cnt[cls], 1u); // rank inside this block's share of the class
    }
    __syncthreads();
    if (threadIdx.x < FSM_CLASSES && sh_cnt[threadIdx.x]) sh_base[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], sh_cnt[threadIdx.x]);
    __syncthreads();
    if (s < nseq) {
        // Ordinal j in class order (most hits first).  A warp full of long sequences would execute the union of 32
        // divergent paths on every hit and stretch the kernel's critical path; so the nl longest sequences are dealt one
        // per warp (lane 0) and the others fill lanes 1..31 and then the remaining warps, still grouped by class.
        const uint32_t j = start[cls] + sh_base[cls] + mine, nl = s_nlong;
        uint32_t slot;
        if (j < nl) {
            slot = 32 * j;
        } else {
            const uint32_t q = j - nl;
            slot = q < 31 * nl ? (q / 31) * 32 + 1 + q % 31 : 32 * nl + (q - 31 * nl);
        }
        perm[slot] = (uint32_t)s;
    }
}

// The hits of a sequence are read through a shared-memory window that the warp refills for all its 32 sequences together:
// FW hits per lane, fetched with coalesced loads (eight lanes read the next eight positions / payloads of one sequence: one
// sector / one line) instead of 32 lanes each pulling on a line of their own -- with ~10^5 such streams open at once the
// lines were evicted before their other hits were used and the kernel waited on memory at 41 % occupancy (ncu, r01).
constexpr int FW = 8, FW_STRIDE = FW + 1, FSM_BLK = 128;
__global__ __launch_bounds__(FSM_BLK) void k_fsm(const uint64_t* __restrict__ voff, uint64_t nseq, int per_seq,
                                                 const uint32_t* __restrict__ tile_base, const uint32_t* __restrict__ tile_out,
                                                 uint32_t ntiles, const uint32_t* __restrict__ chunk_pos,
                                                 const int4* __restrict__ chunk_payload, KgFsmParams p,
                                                 KgDevCall* __restrict__ sparse, const uint32_t* __restrict__ lo,
                                                 const uint32_t* __restrict__ perm, uint32_t* __restrict__ call_cnt,
                                                 kg_otu* __restrict__ otus, const unsigned long long* __restrict__ ctr) {
    __shared__ uint32_t s_pos[FSM_BLK / 32][32 * FW_STRIDE];
    __shared__ int4 s_pl[FSM_BLK / 32][32 * FW_STRIDE];
    const uint64_t tix = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool have = tix < nseq;
    const uint64_t s = have ? perm[tix] : 0;
    if (ctr[KG_CTR_OVERFLOW]) { // some tile could not claim its chunk: the host repeats the pass with larger buffers
        if (have)
            for (int k = 0; k < per_seq; k++) call_cnt[s * per_seq + k] = 0;
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane >> 3, k8 = lane & 7; // refill: lane = (which of four sequences, which of its next eight hits)
    uint32_t* wpos = s_pos[warp];
    int4* wpl = s_pl[warp];
    KgFsm f;
    f.begin_sequence();
    for (int k = 0; k < per_seq; k++) {
        const uint64_t v = s * per_seq + k;
        uint32_t x0 = 0, x1 = 0, t = 0, rank = 0, e = 0, cnt = 0, base = 0;
        bool done = true;
        if (have) {
            x0 = (uint32_t)voff[v];
            x1 = (uint32_t)voff[v + 1];
            t = x0 >> TILE_SHIFT;
            rank = lo[v]; // k_lo_tiles
            if (t < ntiles && x1 > x0) {
                const uint32_t o = tile_out[t];
                cnt = tile_out[t + 1] - o;
                base = tile_base[t];
                e = rank - o;
                done = false;
            }
        }
        f.begin_container();
        SparseEmit emit{sparse + rank / (uint32_t)p.min_hits};
        while (__any_sync(0xFFFFFFFFu, !done)) {
            const uint32_t rem = done ? 0u : cnt - e;
#pragma unroll
            for (int q = 0; q < 32 / 4; q++) {
                const int src = 4 * q + sub;
                const uint32_t from = __shfl_sync(0xFFFFFFFFu, base + e, src), rs = __shfl_sync(0xFFFFFFFFu, rem, src);
                if ((uint32_t)k8 < rs) {
                    wpos[src * FW_STRIDE + k8] = chunk_pos[from + k8];
                    wpl[src * FW_STRIDE + k8] = chunk_payload[from + k8];
                }
            }
            __syncwarp();
            const uint32_t wn = min(rem, (uint32_t)FW);
#pragma unroll 1
            for (int j = 0; j < FW; j++) {
                if (!done && (uint32_t)j < wn) {
                    const uint32_t g = wpos[lane * FW_STRIDE + j];
                    if (g >= x1) {
                        done = true; // the chunk goes on with the next sequence's hits
                    } else {
                        const int4 pl = wpl[lane * FW_STRIDE + j];
                        KgHitLite h = {(int)(g - x0), pl.z, pl.y, pl.x, __int_as_float(pl.w)};
                        f.hit(p, h, emit);
                    }
                }
            }
            __syncwarp();
            if (!done) {
                e += wn;
                if (e >= cnt) { // next tile of this sequence, if any
                    t++;
                    if (t >= ntiles || ((uint64_t)t << TILE_SHIFT) >= x1) {
                        done = true;
                    } else {
                        const uint32_t o = tile_out[t];
                        cnt = tile_out[t + 1] - o;
                        base = tile_base[t];
                        e = 0;
                    }
                }
            }
        }
        if (have) {
            f.end_container(p, emit);
            call_cnt[v] = (uint32_t)f.ncalls;
        }
    }
    if (have) {
        kg_otu o;
        o.n = f.otu_c.n;
#pragma unroll
        for (int i = 0; i < KG_OI_BUFSZ; i++) {
            o.count[i] = f.otu_c.c[i];
            o.oI[i] = f.otu_c.o[i];
        }
        otus[s] = o;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// segment path (long contigs; kg_fsm.cuh explains why a container can be cut at gaps > max_gap)
// ---------------------------------------------------------------------------------------------------------------
// All kernels of this path size their work from counters that live on the device (hits, segments): their grids are
// fixed (tiles of the batch, or a few blocks per SM walking a grid-stride loop), never "one thread per slot of the hit
// buffer" -- the buffer is sized for the worst case and walking it cost more than the work itself (r02 launch list:
// 1.2 ms of 3.6 ms in kernels whose threads found nothing to do).
//
// k_tile_bnd: which tiles contain the start of a container.  Two hits less than max_gap apart whose tiles hold no
// container start belong to the same container; only the others need the binary search over the offsets.
__global__ void k_tile_bnd(const uint64_t* __restrict__ voff, uint64_t nv, uint32_t ntiles, uint8_t* __restrict__ tile_bnd) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const uint64_t t = voff[v] >> TILE_SHIFT;
    if (t < ntiles) tile_bnd[t] = 1;
}

// k_gather_seg: the gather of k_gather (GATHER_TILES tiles per block, walked as one flat range) fused with the segment
// cut: hit g starts a segment when it is the first hit of its container or lies more than max_gap behind hit g-1
// (KGJ:477-478; kg_fsm.cuh explains why the FSM state is empty there).  The block's segment starts are compacted, in
// order, into seg_tmp[o0 ...] (o0 = rank of the block's first hit: a block never has more starts than hits) and counted
// in blk_cnt; k_seg_offsets / k_seg_place turn that into the dense seg_begin list.
constexpr int GS_BLK = 256;
__global__ __launch_bounds__(GS_BLK) void k_gather_seg(const uint32_t* __restrict__ chunk_pos, const int4* __restrict__ chunk_payload,
                                                       const uint32_t* __restrict__ tile_base, const uint32_t* __restrict__ tile_out, uint32_t ntiles,
                                                       const uint64_t* __restrict__ voff, uint64_t nv, const uint8_t* __restrict__ tile_bnd, int max_gap,
                                                       uint32_t* __restrict__ hit_pos, int4* __restrict__ hit_payload,
                                                       uint32_t* __restrict__ seg_tmp, uint32_t* __restrict__ blk_cnt,
                                                       const unsigned long long* __restrict__ ctr) {
    __shared__ uint32_t s_out[GATHER_TILES + 1], s_base[GATHER_TILES];
    __shared__ uint32_t s_pos[GS_BLK], s_wcnt[GS_BLK / 32], s_prev, s_run;
    const uint32_t t0 = blockIdx.x * GATHER_TILES;
    if (ctr[KG_CTR_OVERFLOW]) { // the host repeats the pass with larger buffers
        if (threadIdx.x == 0) blk_cnt[blockIdx.x] = 0;
        return;
    }
    const uint32_t nt = min((uint32_t)GATHER_TILES, ntiles - t0);
    if (threadIdx.x <= nt) s_out[threadIdx.x] = tile_out[t0 + threadIdx.x];
    if (threadIdx.x < nt) s_base[threadIdx.x] = tile_base[t0 + threadIdx.x];
    __syncthreads();
    const uint32_t o0 = s_out[0], total = s_out[nt] - o0;
    if (threadIdx.x == 0) {
        s_run = 0;
        uint32_t prev = 0xFFFFFFFFu; // position of hit o0 - 1: the last hit of the last non-empty tile before this block
        if (o0 > 0) {
            uint32_t lo = 0, hi = t0; // largest tp < t0 with tile_out[tp] < o0  (tile_out[0] = 0 < o0)
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (tile_out[mid] < o0) lo = mid;
                else hi = mid;
            }
            prev = chunk_pos[tile_base[lo] + (o0 - 1 - tile_out[lo])];
        }
        s_prev = prev;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t j0 = 0; j0 < total; j0 += GS_BLK) {
        const uint32_t j = j0 + threadIdx.x;
        const bool valid = j < total;
        const uint32_t g = o0 + j; // rank of the hit in position order
        uint32_t pos = 0;
        if (valid) {
            uint32_t lo = 0, hi = nt; // the last tile that starts at or before g (empty tiles share their successor's start)
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (s_out[mid] <= g) lo = mid;
                else hi = mid;
            }
            const uint32_t src = s_base[lo] + (g - s_out[lo]);
            pos = chunk_pos[src];
            const int4 pl = chunk_payload[src];
            hit_pos[g] = pos;
            hit_payload[g] = pl;
        }
        s_pos[threadIdx.x] = pos;
        __syncthreads();
        bool start = false;
        if (valid) {
            const uint32_t prev = threadIdx.x ? s_pos[threadIdx.x - 1] : s_prev;
            if (prev == 0xFFFFFFFFu || pos - prev > (uint32_t)max_gap) {
                start = true;
            } else {
                const uint32_t tp = prev >> TILE_SHIFT, tc = pos >> TILE_SHIFT;
                if (tc - tp > 1 || tile_bnd[tp] || tile_bnd[tc]) start = (uint64_t)prev < voff[seq_of(voff, nv, pos)];
            }
        }
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, start);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        uint32_t before = s_run, all = 0;
#pragma unroll
        for (int w = 0; w < GS_BLK / 32; w++) {
            before += w < warp ? s_wcnt[w] : 0u;
            all += s_wcnt[w];
        }
        if (start) seg_tmp[o0 + before + __popc(bal & ((1u << lane) - 1u))] = g;
        __syncthreads();
        if (threadIdx.x == GS_BLK - 1) s_prev = pos; // only read again when the next pass exists, i.e. this one was full
        if (threadIdx.x == 0) s_run += all;
    }
    __syncthreads();
    if (threadIdx.x == 0) blk_cnt[blockIdx.x] = s_run;
}
// one block: exclusive scan of the per-block segment counts; nseg and the closing seg_begin entry
__global__ __launch_bounds__(1024) void k_seg_offsets(const uint32_t* __restrict__ blk_cnt, uint32_t nblk, uint32_t* __restrict__ blk_off,
                                                      const uint32_t* __restrict__ tile_out, uint32_t ntiles, uint32_t* __restrict__ seg_begin,
                                                      uint32_t* __restrict__ nseg_out, const unsigned long long* __restrict__ ctr) {
    __shared__ uint32_t s_w[32], s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t i0 = 0; i0 < nblk; i0 += 1024) {
        const uint32_t i = i0 + threadIdx.x;
        const uint32_t v = i < nblk ? blk_cnt[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) s_w[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_w[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, w, d);
                if (lane >= d) w += y;
            }
            s_w[lane] = w; // inclusive over warps
        }
        __syncthreads();
        const uint32_t carry = s_carry, wbase = warp ? s_w[warp - 1] : 0u;
        if (i < nblk) blk_off[i] = carry + wbase + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + wbase + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const uint32_t ns = s_carry;
        nseg_out[0] = ns;
        seg_begin[ns] = ctr[KG_CTR_OVERFLOW] ? 0u : tile_out[ntiles];
    }
}
__global__ __launch_bounds__(GS_BLK) void k_seg_place(const uint32_t* __restrict__ seg_tmp, const uint32_t* __restrict__ blk_cnt,
                                                      const uint32_t* __restrict__ blk_off, const uint32_t* __restrict__ tile_out,
                                                      uint32_t* __restrict__ seg_begin) {
    const uint32_t cnt = blk_cnt[blockIdx.x], off = blk_off[blockIdx.x];
    if (!cnt) return;
    const uint32_t o0 = tile_out[blockIdx.x * GATHER_TILES];
    for (uint32_t k = threadIdx.x; k < cnt; k += GS_BLK) seg_begin[off + k] = seg_tmp[o0 + k];
}
// lo[v] = rank of the first hit at or after the start of container v (v = nv gives the number of hits)
__global__ void k_lo(const uint64_t* __restrict__ voff, uint64_t nv, const uint32_t* __restrict__ hit_pos,
                     const uint32_t* __restrict__ tile_out, uint32_t ntiles, uint32_t* __restrict__ lo,
                     const unsigned long long* __restrict__ ctr) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v > nv) return;
    const uint32_t nhits = ctr[KG_CTR_OVERFLOW] ? 0u : tile_out[ntiles];
    const uint64_t x = voff[v];
    uint32_t a = 0, b = nhits;
    while (a < b) {
        const uint32_t mid = (a + b) >> 1;
        if ((uint64_t)hit_pos[mid] < x) a = mid + 1;
        else b = mid;
    }
    lo[v] = a;
}
// Segments are handed to the FSM threads grouped by length (a gene of 300 hits next to a stray segment of one hit left the
// average lane idle 82 % of the time: ncu, r02).  A counting sort over a count that lives on the device (grid-stride
// loops); the classes are exact up to 16 hits and then four per octave, so the lanes of a warp differ by < 25 %; the
// longest segments come first.  (Unlike the per-sequence path no segment gets a warp to itself: 3 % of the segments of a
// genome have 256 hits or more, and a warp per long segment was most of the kernel's instruction count.)
constexpr int SEG_CLASSES = 96;
__device__ __forceinline__ uint32_t seg_class(uint32_t hits) { // 0 = most hits
    uint32_t c = hits;
    if (hits >= 16) {
        const int e = 31 - __clz(hits);                       // 4 ..
        c = 16 + 4 * (e - 4) + ((hits >> (e - 2)) & 3u);      // 16 .. 16 + 4*27 + 3
    }
    return (SEG_CLASSES - 1) - min(c, (uint32_t)SEG_CLASSES - 1);
}
__global__ __launch_bounds__(256) void k_seg_hist(const uint32_t* __restrict__ seg_begin, const uint32_t* __restrict__ nseg,
                                                  uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[SEG_CLASSES];
    if (threadIdx.x < SEG_CLASSES) sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t n = nseg[0];
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
        atomicAdd(&sh[seg_class(seg_begin[j + 1] - seg_begin[j])], 1u);
    __syncthreads();
    if (threadIdx.x < SEG_CLASSES && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}
__global__ __launch_bounds__(256) void k_seg_scatter(const uint32_t* __restrict__ seg_begin, const uint32_t* __restrict__ nseg,
                                                     const uint32_t* __restrict__ hist, uint32_t* __restrict__ cursor, uint32_t* __restrict__ perm) {
    __shared__ uint32_t start[SEG_CLASSES], sh_cnt[SEG_CLASSES], sh_base[SEG_CLASSES];
    const uint32_t n = nseg[0];
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int c = 0; c < SEG_CLASSES; c++) {
            start[c] = acc;
            acc += hist[c];
        }
    }
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t j0 = blockIdx.x * blockDim.x; j0 < n; j0 += stride) {
        if (threadIdx.x < SEG_CLASSES) sh_cnt[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t s = j0 + threadIdx.x;
        uint32_t cls = 0, mine = 0;
        if (s < n) {
            cls = seg_class(seg_begin[s + 1] - seg_begin[s]);
            mine = atomicAdd(&sh_cnt[cls], 1u);
        }
        __syncthreads();
        if (threadIdx.x < SEG_CLASSES && sh_cnt[threadIdx.x]) sh_base[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], sh_cnt[threadIdx.x]);
        __syncthreads();
        if (s < n) perm[start[cls] + sh_base[cls] + mine] = s;
        __syncthreads();
    }
}
// One thread per segment, 32 segments of about the same length per warp -- but a thread that walks its own hit list
// straight from global memory touches a different 128-byte line than its 31 neighbours on every step, and with ~10^5 such
// streams open at once the lines are evicted from L1 and L2 before their other seven hits are used (ncu r02: better
// balanced warps made the kernel SLOWER, 1.1 -> 1.6 ms).  So the warp refills a small shared-memory window for all its
// segments together: FS_CHUNK hits per lane, fetched with coalesced loads (eight lanes read the eight positions / payloads
// of one segment: one sector / one line), then every lane steps its FSM through its own window.  Each line of the hit
// arrays is read once.
constexpr int FS_CHUNK = 8, FS_BLK = 128;
constexpr int FS_PSTRIDE = FS_CHUNK + 1; // padded strides: conflict-free shared-memory reads by lane
__global__ __launch_bounds__(FS_BLK) void k_fsm_seg(const uint64_t* __restrict__ voff, uint64_t nv, const uint32_t* __restrict__ hit_pos,
                                                    const int4* __restrict__ hit_payload, const uint32_t* __restrict__ seg_begin,
                                                    const uint32_t* __restrict__ nseg, const uint32_t* __restrict__ perm,
                                                    const uint32_t* __restrict__ lo, KgFsmParams p, KgDevCall* __restrict__ sparse,
                                                    uint2* __restrict__ seg_cnt, uint32_t* __restrict__ seg_v, int2* __restrict__ run) {
    __shared__ uint32_t s_pos[FS_BLK / 32][32 * FS_PSTRIDE];
    __shared__ int4 s_pl[FS_BLK / 32][32 * FS_PSTRIDE];
    const uint32_t ns = nseg[0];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane >> 3, k8 = lane & 7; // refill: lane = (which of four segments, which of its next eight hits)
    uint32_t* wpos = s_pos[warp];
    int4* wpl = s_pl[warp];
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t t0 = blockIdx.x * blockDim.x + warp * 32; t0 < ns; t0 += stride) { // t0 is warp-uniform
        const uint32_t tix = t0 + (uint32_t)lane;
        const bool have = tix < ns;
        uint32_t j = 0, a = 0, b = 0, v = 0, base = 0;
        KgFsmSeg f;
        f.begin(0);
        if (have) {
            j = perm[tix];
            a = seg_begin[j];
            b = seg_begin[j + 1];
            v = (uint32_t)seq_of(voff, nv, hit_pos[a]);
            base = (uint32_t)voff[v];
            f.begin((int)(a - lo[v])); // HIT lines of this container printed before the segment
        }
        SparseEmit emit{sparse + a / (uint32_t)p.min_hits};
        KgSegRuns runs{run + a, 0u};
        // software pipeline: the loads of window w+1 are in flight (in registers) while the lanes step through window w
        uint32_t i = a;
        uint32_t rp[32 / 4];
        int4 rl[32 / 4];
        auto fetch = [&](uint32_t from) {
#pragma unroll
            for (int t = 0; t < 32 / 4; t++) {
                const int src = 4 * t + sub;
                const uint32_t is = __shfl_sync(0xFFFFFFFFu, from, src), bs = __shfl_sync(0xFFFFFFFFu, b, src);
                const uint32_t idx = min(is + (uint32_t)k8, bs ? bs - 1 : 0u); // clamped: a valid hit, never used past the end
                rp[t] = hit_pos[idx];
                rl[t] = hit_payload[idx];
            }
        };
        fetch(i);
        while (__any_sync(0xFFFFFFFFu, i < b)) {
#pragma unroll
            for (int t = 0; t < 32 / 4; t++) {
                wpos[(4 * t + sub) * FS_PSTRIDE + k8] = rp[t];
                wpl[(4 * t + sub) * FS_PSTRIDE + k8] = rl[t];
            }
            __syncwarp();
            const uint32_t inext = min(b, i + FS_CHUNK);
            if (__any_sync(0xFFFFFFFFu, inext < b)) fetch(inext);
#pragma unroll 1
            for (int k = 0; k < FS_CHUNK; k++) {
                if (i + k < b) {
                    const uint32_t pos = wpos[lane * FS_PSTRIDE + k];
                    const int4 pl = wpl[lane * FS_PSTRIDE + k];
                    KgHitLite h = {(int)(pos - base), pl.z, pl.y, pl.x, __int_as_float(pl.w)};
                    f.hit(p, h, emit, runs);
                }
            }
            i = inext;
            __syncwarp();
        }
        if (have) {
            // In the reference the run that ends at a gap is processed when the NEXT hit of the container arrives, after that
            // hit's HIT line (KGJ:472-480); only the container's last run is processed after the loop (KGJ:511-513).
            if (j + 1 < ns && (uint64_t)hit_pos[b] < voff[v + 1]) f.consumed++;
            f.end(p, emit, runs);
            seg_cnt[j] = make_uint2((uint32_t)f.ncalls, runs.n);
            seg_v[j] = v;
        }
    }
}
// Exclusive scan of the (calls, OTU runs) pairs over a count that lives on the device: three small launches with a fixed grid
// (per-block sums over contiguous chunks, one block scans the sums, every block scans its chunk).
constexpr int SCAN_BLOCKS = 256, SCAN_BLK = 256;
__device__ __forceinline__ uint2 add2(uint2 a, uint2 b) { return make_uint2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ uint2 shfl_up2(uint2 v, int d) {
    return make_uint2(__shfl_up_sync(0xFFFFFFFFu, v.x, d), __shfl_up_sync(0xFFFFFFFFu, v.y, d));
}
// inclusive scan across the block; returns the block total in `total`
__device__ __forceinline__ uint2 block_scan2(uint2 v, uint2* s_w /* SCAN_BLK / 32 */, uint2& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint2 x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint2 y = shfl_up2(x, d);
        if (lane >= d) x = add2(x, y);
    }
    __syncthreads(); // s_w may still be read from the previous call
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    uint2 wbase = make_uint2(0, 0), all = make_uint2(0, 0);
#pragma unroll
    for (int w = 0; w < SCAN_BLK / 32; w++) {
        if (w < warp) wbase = add2(wbase, s_w[w]);
        all = add2(all, s_w[w]);
    }
    total = all;
    return add2(x, wbase);
}
__device__ __forceinline__ void scan_chunk(uint32_t n, uint32_t& c0, uint32_t& c1) {
    const uint32_t per = ((n + SCAN_BLOCKS - 1) / SCAN_BLOCKS + SCAN_BLK - 1) / SCAN_BLK * SCAN_BLK;
    c0 = min(n, blockIdx.x * per);
    c1 = min(n, c0 + per);
}
__global__ __launch_bounds__(SCAN_BLK) void k_scan2_sums(const uint2* __restrict__ in, const uint32_t* __restrict__ nptr, uint2* __restrict__ part) {
    __shared__ uint2 s_w[SCAN_BLK / 32];
    uint32_t c0, c1;
    scan_chunk(nptr[0], c0, c1);
    uint2 acc = make_uint2(0, 0);
    for (uint32_t i = c0 + threadIdx.x; i < c1; i += SCAN_BLK) acc = add2(acc, in[i]);
    uint2 total;
    block_scan2(acc, s_w, total);
    if (threadIdx.x == 0) part[blockIdx.x] = total;
}
__global__ __launch_bounds__(SCAN_BLK) void k_scan2_top(uint2* __restrict__ part, const uint32_t* __restrict__ nptr, uint2* __restrict__ out,
                                                        uint32_t* __restrict__ totals) {
    static_assert(SCAN_BLOCKS == SCAN_BLK, "one pass");
    __shared__ uint2 s_w[SCAN_BLK / 32];
    const uint2 v = part[threadIdx.x];
    uint2 total;
    const uint2 inc = block_scan2(v, s_w, total);
    part[threadIdx.x] = make_uint2(inc.x - v.x, inc.y - v.y);
    if (threadIdx.x == 0) {
        out[nptr[0]] = total;
        totals[0] = total.x;
        totals[1] = total.y;
    }
}
__global__ __launch_bounds__(SCAN_BLK) void k_scan2_apply(const uint2* __restrict__ in, const uint32_t* __restrict__ nptr, const uint2* __restrict__ part,
                                                          uint2* __restrict__ out) {
    __shared__ uint2 s_w[SCAN_BLK / 32];
    uint32_t c0, c1;
    scan_chunk(nptr[0], c0, c1);
    uint2 run = part[blockIdx.x];
    for (uint32_t i0 = c0; i0 < c1; i0 += SCAN_BLK) {
        const uint32_t i = i0 + threadIdx.x;
        const uint2 v = i < c1 ? in[i] : make_uint2(0, 0);
        uint2 total;
        const uint2 inc = block_scan2(v, s_w, total);
        if (i < c1) out[i] = make_uint2(run.x + inc.x - v.x, run.y + inc.y - v.y);
        run = add2(run, total);
    }
}
// calls (tagged with sequence and frame) and OTU runs, packed in segment order = print order (+0,+1,+2,-0,-1,-2 and
// ascending positions, KGJ:540-557)
__global__ __launch_bounds__(256) void k_compact_seg(const KgDevCall* __restrict__ sparse, const uint32_t* __restrict__ seg_begin,
                                                     const uint32_t* __restrict__ nseg, const uint32_t* __restrict__ seg_v,
                                                     const uint2* __restrict__ seg_off, const int2* __restrict__ run, int per_seq,
                                                     int min_hits, uint32_t seq_base, kg_call* __restrict__ out, int2* __restrict__ dense) {
    const uint32_t ns = nseg[0];
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < ns; j += gridDim.x * blockDim.x) {
        const uint2 o = seg_off[j], e = seg_off[j + 1];
        if (o.x == e.x) continue; // no call, hence no runs either
        const uint32_t a = seg_begin[j], v = seg_v[j];
        const KgDevCall* src = sparse + a / (uint32_t)min_hits;
        for (uint32_t k = 0; k < e.x - o.x; k++) {
            const KgDevCall d = src[k];
            kg_call r;
            r.seq = seq_base + v / (uint32_t)per_seq;
            r.strand_frame = (int32_t)(v % (uint32_t)per_seq);
            r.start = d.start;
            r.end = d.end;
            r.count = d.count;
            r.fI = d.fI;
            r.weighted = d.weighted;
            r.hits_before = d.hits_before;
            out[o.x + k] = r;
        }
        for (uint32_t k = 0; k < e.y - o.y; k++) dense[o.y + k] = run[a + k];
    }
}
// OTU-COUNTS (KGJ:413-438, 516-524): replay, in order, the OTU index of every hit a CALL counted.  Only the fold of the
// five-entry buffer is inherently sequential per sequence.  k_fsm_seg leaves, per segment, the run-length encoded OTU
// indices of its calls; k_compact_seg packs them in print order; here one warp folds one sequence: 32 runs arrive per
// coalesced load (the next 32 are already on their way) and wait in shared memory, every lane applies them in order (all
// lanes hold the same buffer).  The update has no branch (kg_otu_update_n) and the shared-memory reads do not depend on
// the buffer, so what is left is the dependent chain of the updates themselves (~5 k per 5 Mbp contig).
__global__ __launch_bounds__(32) void k_otu_fold(const uint32_t* __restrict__ lo, uint64_t nseq, int per_seq, const uint32_t* __restrict__ seg_begin,
                                                 const uint32_t* __restrict__ nseg, const uint2* __restrict__ seg_off,
                                                 const int2* __restrict__ dense, kg_otu* __restrict__ otus,
                                                 const unsigned long long* __restrict__ ctr) {
    __shared__ int2 s_run[2][32];
    const uint64_t s = blockIdx.x;
    const int lane = threadIdx.x;
    if (s >= nseq || ctr[KG_CTR_OVERFLOW]) return;
    const uint32_t ns = nseg[0];
    // the first hit of a container always starts a segment: the sequence's runs are those of the segments from the one that
    // begins at lo[first container] up to the one that begins at lo[first container of the next sequence]
    uint32_t rr[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
        const uint32_t a = lo[(s + e) * per_seq];
        uint32_t x = 0, y = ns; // first segment with seg_begin >= a (seg_begin[ns] = number of hits)
        while (x < y) {
            const uint32_t mid = (x + y) >> 1;
            if (seg_begin[mid] < a) x = mid + 1;
            else y = mid;
        }
        rr[e] = seg_off[x].y;
    }
    const uint32_t r0 = rr[0], r1 = rr[1];
    KgOtuBuf u;
    kg_otu_clear(u);
    int2 nxt = r0 + lane < r1 ? dense[r0 + lane] : make_int2(0, 0);
    int buf = 0;
    for (uint32_t r = r0; r < r1; r += 32, buf ^= 1) {
        s_run[buf][lane] = nxt;
        __syncwarp();
        const uint32_t idx = r + 32 + (uint32_t)lane;
        nxt = idx < r1 ? dense[idx] : make_int2(0, 0);
        const int cnt = (int)min(32u, r1 - r);
        if (cnt == 32) {
#pragma unroll 8
            for (int k = 0; k < 32; k++) {
                const int2 q = s_run[buf][k];
                kg_otu_update_n(u, q.x, q.y);
            }
        } else {
            for (int k = 0; k < cnt; k++) {
                const int2 q = s_run[buf][k];
                kg_otu_update_n(u, q.x, q.y);
            }
        }
    }
    if (lane == 0) {
        kg_otu o;
        o.n = u.n;
#pragma unroll
        for (int k = 0; k < KG_OI_BUFSZ; k++) {
            o.count[k] = u.c[k];
            o.oI[k] = u.o[k];
        }
        otus[s] = o;
    }
}

__global__ void k_compact_calls(const KgDevCall* __restrict__ sparse, const uint32_t* __restrict__ lo,
                                const uint32_t* __restrict__ call_off, uint64_t nv, int per_seq, int min_hits,
                                uint32_t seq_base, kg_call* __restrict__ out) {
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const uint32_t o = call_off[v], c = call_off[v + 1] - o;
    const KgDevCall* src = sparse + lo[v] / (uint32_t)min_hits;
    for (uint32_t j = 0; j < c; j++) {
        KgDevCall d = src[j];
        kg_call r;
        r.seq = seq_base + (uint32_t)(v / per_seq);
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
                            const int4* __restrict__ hit_payload, uint32_t nhits, uint32_t seq_base, kg_hit* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nhits) return;
    const uint64_t g = hit_pos[i];
    const uint64_t v = seq_of(voff, nv, g);
    const int4 pl = hit_payload[i];
    kg_hit h;
    h.seq = seq_base + (uint32_t)(v / per_seq);
    h.strand_frame = (int32_t)(v % per_seq);
    h.pos = (int32_t)(g - voff[v]);
    h.oI = pl.x;
    h.avg_off_from_end = pl.y;
    h.fI = pl.z;
    h.function_wt = __int_as_float(pl.w);
    out[i] = h;
}

constexpr int KG_MAX_PARTS = 16;
struct PipeSlot { // everything ONE in-flight pipeline run needs; two of them let slice s+1 queue up behind slice s
    DevBuf tile_base, tile_cnt, tile_out, chunk_pos, chunk_payload, lo, sparse, call_cnt, call_off, ctr;
    DevBuf otu_cnt, otu_first;                                   // kg_run: OTU entries per sequence and their exclusive scan
    DevBuf queue, tile_qcnt;                                     // probe cascade: survivors (8 KB slice per tile) and their counts
    DevBuf half_pos, half_payload, half_base, half_cnt;          // two-pass probe: hits of pass 0 (per-tile chunks)
    bool halves = false;
    DevBuf hit_pos, hit_payload;                                 // position-ordered hits (segment path, "-d")
    DevBuf fk_a, fi_a;                                           // per-sequence path: class histogram / cursors, sequence permutation
    DevBuf seg_tmp, seg_v, seg_perm, seg_begin, seg_cnt, seg_off, nseg, tile_bnd, blk_cnt, blk_off, scan_part; // segment path
    DevBuf o_run, o_dense;                                       // segment path: OTU runs (oI, length) per segment (sparse) and packed in print order
    bool seg = false;                                            // which FSM path the enqueued run uses
    uint64_t* h_ctr = nullptr; // pinned, KG_CTR_COUNT + 1 (the last slot receives the call total)
    cudaEvent_t ev[6] = {};    // begin, probe begin, probe end, end, cascade: first filter done, second filter done
    cudaEvent_t part_ev[KG_MAX_PARTS + 1] = {}; // cascade in parts: filters of part p done; [KG_MAX_PARTS]: last line stage done
    bool cascade = false;      // the enqueued run used the three-kernel probe
    uint32_t cascade_parts = 1;
    uint64_t hit_cap = 0;
    uint32_t launches = 0;
};
struct RunScratch { // grow-only device scratch kept per context (so repeated runs do not allocate)
    PipeSlot slot[2];
    uint64_t hit_cap_seen = 0;   // hits of the largest run so far (+ slack): sizes the next run's buffers
    uint64_t calls_seen = 0;     // calls of the largest kg_run so far: sizes the pinned result buffer
    uint64_t otu_entries_seen = 0; // likewise, (count, oI) pairs
    struct InFlight {              // kg_batch_submit / kg_batch_collect
        kg_result* r = nullptr;
        kg_batch* b = nullptr;
        const kg_table* t = nullptr;
        kg_params p = {};
    } inflight[2];
    int inflight_head = 0, inflight_n = 0;
};
RunScratch& scratch_of(kg_context* ctx) {
    if (!ctx->scratch) {
        RunScratch* sc = new RunScratch();
        for (auto& sl : sc->slot) {
            cudaMallocHost(&sl.h_ctr, (KG_CTR_COUNT + 2) * sizeof(uint64_t)); // + call total, + OTU entry total
            for (auto& e : sl.ev) cudaEventCreate(&e);
            for (auto& e : sl.part_ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        }
        ctx->scratch = sc;
    }
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
    // size classes (next power of two, refined to eighths) so that slices of slightly different sizes reuse each other's
    // buffers: a cudaMalloc / cudaFree inside a run synchronises the device
    size_t cls = 256;
    while (cls < bytes) cls <<= 1;
    if (cls >= 2048) {
        const size_t eighth = cls >> 4; // refine within [cls/2, cls] in steps of cls/16
        cls = (bytes + eighth - 1) / eighth * eighth;
    }
    *out = DevBuf();
    return out->ensure(cls);
}
void pool_give_dev(kg_context* ctx, DevBuf* b) {
    if (!b->p) return;
    if (ctx->dev_pool.size() >= 64) { // keep the pool bounded: drop the smallest
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
    if (ctx->host_pool.size() >= 32) {
        size_t k = 0;
        for (size_t i = 1; i < ctx->host_pool.size(); i++)
            if (ctx->host_pool[i].cap < ctx->host_pool[k].cap) k = i;
        cudaFreeHost(ctx->host_pool[k].p);
        ctx->host_pool.erase(ctx->host_pool.begin() + k);
    }
    ctx->host_pool.push_back(*b);
    *b = HostBuf();
}

// CUB temporary storage: one buffer per stream, because work on the two streams overlaps
DevBuf& scan_tmp_of(kg_context* ctx, cudaStream_t st) { return st == ctx->fsm_stream ? ctx->scan_tmp2 : ctx->scan_tmp; }
int exclusive_sum_u32(kg_context* ctx, const uint32_t* in, uint32_t* out, size_t n, cudaStream_t st) {
    size_t bytes = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, st));
    KG_TRY(scan_tmp_of(ctx, st).ensure(bytes));
    CU(cub::DeviceScan::ExclusiveSum(scan_tmp_of(ctx, st).p, bytes, in, out, n, st));
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
    CU(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    {   // the FSM's critical path is one long sequence on one thread: let its blocks jump the queue of probe blocks
        int lo_p = 0, hi_p = 0;
        cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
        CU(cudaStreamCreateWithPriority(&ctx->fsm_stream, cudaStreamNonBlocking, hi_p));
        CU(cudaStreamCreateWithPriority(&ctx->lines_stream, cudaStreamNonBlocking, hi_p)); // cascade in parts: the bucket-line stage
    }
    for (auto& ev : ctx->ev) CU(cudaEventCreate(&ev));
    for (auto& ev : ctx->d2h_ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : ctx->up_ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CU(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PROBE_SMEM));
    CU(cudaFuncSetAttribute(k_probe_half<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PROBE_SMEM));
    CU(cudaFuncSetAttribute(k_probe_half<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PROBE_SMEM));
    CU(cudaMallocHost(&ctx->h_counters, (KG_CTR_COUNT + 1) * sizeof(uint64_t)));
    memset(ctx->h_counters, 0, (KG_CTR_COUNT + 1) * sizeof(uint64_t));
    *out = ctx;
    return KG_OK;
}

extern "C" void kg_shutdown(kg_context* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    RunScratch& sc = scratch_of(ctx);
    for (auto& sl : sc.slot) {
        for (DevBuf* b : {&sl.tile_base, &sl.tile_cnt, &sl.tile_out, &sl.chunk_pos, &sl.chunk_payload, &sl.lo, &sl.sparse, &sl.call_cnt,
                          &sl.call_off, &sl.ctr, &sl.otu_cnt, &sl.otu_first, &sl.queue, &sl.tile_qcnt, &sl.half_pos, &sl.half_payload, &sl.half_base, &sl.half_cnt, &sl.hit_pos, &sl.hit_payload, &sl.seg_tmp, &sl.seg_v, &sl.seg_perm, &sl.seg_begin,
                          &sl.seg_cnt, &sl.seg_off, &sl.tile_bnd, &sl.blk_cnt, &sl.blk_off, &sl.scan_part,
                          &sl.nseg, &sl.fk_a, &sl.fi_a, &sl.o_run, &sl.o_dense})
            b->release();
        if (sl.h_ctr) cudaFreeHost(sl.h_ctr);
        for (auto& e : sl.ev)
            if (e) cudaEventDestroy(e);
        for (auto& e : sl.part_ev)
            if (e) cudaEventDestroy(e);
    }
    delete static_cast<RunScratch*>(ctx->scratch);
    for (auto& b : ctx->dev_pool) b.release();
    for (auto& h : ctx->host_pool) cudaFreeHost(h.p);
    ctx->scan_tmp.release();
    ctx->scan_tmp2.release();
    for (auto& ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->d2h_ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->up_ev)
        if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    if (ctx->fsm_stream) cudaStreamDestroy(ctx->fsm_stream);
    if (ctx->lines_stream) cudaStreamDestroy(ctx->lines_stream);
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
    pool_give_dev(b->ctx, &b->vseq);
    pool_give_dev(b->ctx, &b->voff);
    pool_give_dev(b->ctx, &b->pk_buf);
    pool_give_dev(b->ctx, &b->aux_buf);
    pool_give_dev(b->ctx, &b->exc_buf);
    delete b;
}

// residue stream + virtual offsets.  aa: patch in place (idempotent).  dna: translate six frames.
int kg_batch_prepare(kg_batch* b, cudaStream_t st, uint32_t* launches) {
    kg_context* ctx = b->ctx;
    if (b->mode == KG_MODE_AA) {
        b->nv = b->n;
        b->vtotal = b->total;
        if (b->n) {
            // k_patch_aa is idempotent (a repeated pass after a hit-buffer overflow patches again, harmlessly); the padded
            // variant finds the last residue by its value, so it must run exactly once per batch
            if (!b->padded) k_patch_aa<<<blocks_for(b->n, 256), 256, 0, st>>>(b->d_seq, b->d_off, b->n);
            else if (!b->prepared) k_patch_aa_padded<<<blocks_for(b->n, 256), 256, 0, st>>>(b->d_seq, b->d_off, b->n);
            (*launches)++;
        }
        b->prepared = true;
        return KG_OK;
    }
    b->nv = 6 * b->n;
    auto need = [&](DevBuf* buf, size_t bytes) -> int { // pooled: slices of a kg_run come and go
        if (buf->cap >= bytes) return KG_OK;
        pool_give_dev(ctx, buf);
        return pool_take_dev(ctx, bytes, buf);
    };
    if (!b->prepared) { // the layout depends only on the lengths: computed once per batch
        KG_TRY(need(&b->voff, (b->nv + 2) * 8 * 2));
        uint64_t* vlen = b->voff.as<uint64_t>() + (b->nv + 2);
        k_vlen<<<blocks_for(b->nv + 1, 256), 256, 0, st>>>(b->d_off, b->n, vlen);
        (*launches)++;
        KG_TRY(exclusive_sum_u64(ctx, vlen, b->voff.as<uint64_t>(), b->nv + 1, st));
        (*launches)++;
        if (!b->vtotal_known) { // sequences adopted from device memory: the total has to come back
            CU(cudaMemcpyAsync(&ctx->h_counters[KG_CTR_VPOS], b->voff.as<uint64_t>() + b->nv, 8, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            b->vtotal = ctx->h_counters[KG_CTR_VPOS];
            b->vtotal_known = true;
        }
        if (b->vtotal > KG_MAX_STREAM) KG_FAIL(KG_ERANGE, "dna batch: %llu translated residues in one batch", (unsigned long long)b->vtotal);
        KG_TRY(need(&b->vseq, b->vtotal + 64));
        b->prepared = true;
    }
    CU(cudaMemsetAsync(b->vseq.as<uint8_t>() + b->vtotal, 0, 64, st)); // tail padding; every stream word is written below
    if (b->vtotal) {
        k_translate<<<blocks_for(b->vtotal / 4, TR_BLK * TR_PASSES), TR_BLK, 0, st>>>(b->d_seq, b->d_off, b->n, b->voff.as<uint64_t>(), b->nv,
                                                                   b->vtotal / 4, b->vseq.as<uint32_t>());
        (*launches)++;
    }
    return KG_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// the device pipeline
// ---------------------------------------------------------------------------------------------------------------
// Which run FSM: one thread per sequence (many short sequences: proteins) or one thread per gap-delimited segment plus an
// OTU replay (long contigs: 6-frame mode).  KG_FSM=seq|seg overrides (tests drive both paths through the same cases).
static bool use_segment_path(int mode) {
    if (const char* e = getenv("KG_FSM")) return strcmp(e, "seg") == 0;
    return mode == KG_MODE_DNA;
}

static uint32_t probe_flags() { // experiment switches: bit 0 = no evict_first on bucket lines, bit 1 = unhinted filter loads
    const char* e = getenv("KG_PROBE_FLAGS");
    return e ? (uint32_t)atoi(e) : 0u;
}

// Which probe: the three-kernel cascade (tables with a second prefilter) or the fused kernel.  KG_PROBE=fused|cascade.
static bool use_cascade(const kg_table* table) {
    if (!table->filter_words || !table->filter2_words || table->filter_halves) return false;
    if (const char* e = getenv("KG_PROBE")) return strcmp(e, "fused") != 0;
    return true;
}
static int probe2_u() { // experiment switch: bucket lines in flight per four-lane group in the cascade's last stage
    static const int u = [] { const char* e = getenv("KG_PROBE2_U"); return e ? atoi(e) : 4; }();
    return u;
}
// In how many parts the cascade runs (KG_CASCADE_PARTS; 1 = the three stages one after the other over the whole batch).
static uint32_t cascade_parts(uint32_t ntiles) {
    uint32_t p = 1;
    if (const char* e = getenv("KG_CASCADE_PARTS")) p = (uint32_t)atoi(e);
    p = std::max(1u, std::min<uint32_t>(p, KG_MAX_PARTS));
    while (p > 1 && ntiles / p < 4096) p--; // parts of less than ~4 M positions only add launches
    return p;
}
// A kernel launch with its own persisting-L2 access-policy window (bytes = 0: no window): the cascade's filters take
// turns in the set-aside part of L2, which a stream-wide window cannot express.
template <class... KArgs, class... Args>
static cudaError_t launch_windowed(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, const void* win_base,
                                   size_t win_bytes, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeAccessPolicyWindow;
    at[0].val.accessPolicyWindow.base_ptr = const_cast<void*>(win_base);
    at[0].val.accessPolicyWindow.num_bytes = win_bytes;
    at[0].val.accessPolicyWindow.hitRatio = 1.0f;
    at[0].val.accessPolicyWindow.hitProp = win_bytes ? cudaAccessPropertyPersisting : cudaAccessPropertyNormal;
    at[0].val.accessPolicyWindow.missProp = win_bytes ? cudaAccessPropertyStreaming : cudaAccessPropertyNormal;
    cfg.attrs = at;
    cfg.numAttrs = getenv("KG_NO_L2_PERSIST") ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Enqueue one whole pass (prepare -> probe -> scan -> FSM -> scan -> compact) on the compute stream; no host
// synchronisation.  The hit buffers are sized from a guess (half the positions, or what an earlier run needed); if a
// tile cannot claim its chunk the kernels downstream skip their work and pipe_finish repeats the pass with the exact size.
// A custom probe stage (the hash-sharded table, kg_shard.cuh) stands in for k_probe: it must fill chunk_pos /
// chunk_payload / tile_base / tile_cnt and the HITS and KMERS counters on the given stream; the batch is then already prepared.
using ProbeStage = std::function<int(PipeSlot& sl, unsigned long long* d_ctr, uint64_t hit_cap, cudaStream_t st)>;

static int pipe_enqueue(kg_context* ctx, PipeSlot& sl, const kg_table* table, kg_batch* b, const kg_params* prm, kg_result* r,
                        uint64_t hit_cap_hint, uint32_t seq_base, const ProbeStage* custom = nullptr) {
    cudaStream_t st = ctx->stream;
    sl.launches = 0;
    sl.halves = false;
    cudaEventRecord(sl.ev[0], st);
    if (!custom) KG_TRY(kg_batch_prepare(b, st, &sl.launches));
    const uint64_t nv = b->nv, vtotal = b->vtotal;
    const int per_seq = b->mode == KG_MODE_AA ? 1 : 6;
    const uint32_t ntiles = (uint32_t)((vtotal + TILE - 1) >> TILE_SHIFT);
    r->n = b->n;
    r->nv = nv;
    r->mode = b->mode;
    r->params = *prm;

    KG_TRY(sl.ctr.ensure(KG_CTR_COUNT * 8));
    unsigned long long* d_ctr = sl.ctr.as<unsigned long long>();
    KG_TRY(sl.tile_base.ensure(((size_t)ntiles + 1) * 4));
    KG_TRY(sl.tile_cnt.ensure(((size_t)ntiles + 1) * 4));
    KG_TRY(sl.tile_out.ensure(((size_t)ntiles + 1) * 4));
    KG_TRY(sl.lo.ensure((nv + 2) * 4));
    KG_TRY(sl.call_cnt.ensure((nv + 1) * 4));
    KG_TRY(sl.call_off.ensure((nv + 1) * 4));
    uint64_t hit_cap = hit_cap_hint ? hit_cap_hint : std::max<uint64_t>(vtotal / 2, 1u << 16);
    if (hit_cap > vtotal) hit_cap = std::max<uint64_t>(vtotal, 1);
    sl.hit_cap = hit_cap;
    const uint64_t max_calls = hit_cap / (uint64_t)prm->min_hits + 1;
    KG_TRY(sl.chunk_pos.ensure(hit_cap * 4));
    KG_TRY(sl.chunk_payload.ensure(hit_cap * sizeof(int4)));
    KG_TRY(sl.sparse.ensure(max_calls * sizeof(KgDevCall)));
    if (r->d_otus.cap < std::max<uint64_t>(b->n, 1) * sizeof(kg_otu)) {
        pool_give_dev(ctx, &r->d_otus);
        KG_TRY(pool_take_dev(ctx, std::max<uint64_t>(b->n, 1) * sizeof(kg_otu), &r->d_otus));
    }
    if (r->d_calls.cap < max_calls * sizeof(kg_call)) {
        pool_give_dev(ctx, &r->d_calls);
        KG_TRY(pool_take_dev(ctx, max_calls * sizeof(kg_call), &r->d_calls));
    }
    const KgFsmParams fp = {prm->min_hits, prm->max_gap, prm->order_constraint, (float)prm->min_weighted_hits};
    CU(cudaMemsetAsync(d_ctr, 0, KG_CTR_COUNT * 8, st));
    CU(cudaMemsetAsync(sl.tile_cnt.p, 0, ((size_t)ntiles + 1) * 4, st));
    cudaEventRecord(sl.ev[1], st);
    if (custom) {
        KG_TRY((*custom)(sl, d_ctr, hit_cap, st));
    } else if (ntiles && use_cascade(table)) {
        sl.cascade = true;
        KG_TRY(sl.queue.ensure((size_t)ntiles * TILE * 8));
        KG_TRY(sl.tile_qcnt.ensure((size_t)ntiles * 4));
        const KgTableView tv = table->view();
        const size_t carve = table->l2_carve; // 0 unless KG_CASCADE_PERSIST: the filters have L2 to themselves in their stages
        // The stage that streams bucket lines (DRAM-bound) can overlap the filter stages of the NEXT part of the batch
        // (bound by L1 wavefronts and L2 hits): parts > 1 puts it on a second, high-priority stream.
        uint32_t parts = cascade_parts(ntiles);
        sl.cascade_parts = parts;
        const uint32_t per = (ntiles + parts - 1) / parts;
        cudaStream_t ls = parts > 1 ? ctx->lines_stream : st;
        for (uint32_t p = 0; p < parts; p++) {
            const uint32_t t0 = p * per, t1 = std::min(ntiles, t0 + per);
            if (t1 <= t0) break;
            const unsigned wgrid = blocks_for(t1 - t0, CAS_BLK / 32);
            CU(launch_windowed(k_filter, dim3(t1 - t0), dim3(PROBE_BLK), 0, st, tv.filter, std::min<size_t>((size_t)tv.filter_words * 8, carve),
                               b->stream(), (uint32_t)vtotal, t0, tv, sl.queue.as<unsigned long long>(), sl.tile_qcnt.as<uint32_t>(), d_ctr, probe_flags()));
            if (parts == 1) cudaEventRecord(sl.ev[4], st);
            CU(launch_windowed(k_refilter, dim3(wgrid), dim3(CAS_BLK), 0, st, tv.filter2, std::min<size_t>((size_t)tv.filter2_words * 8, carve),
                               sl.queue.as<unsigned long long>(), sl.tile_qcnt.as<uint32_t>(), t0, t1, tv));
            if (parts == 1) cudaEventRecord(sl.ev[5], st);
            if (parts > 1) {
                cudaEventRecord(sl.part_ev[p], st);
                cudaStreamWaitEvent(ls, sl.part_ev[p], 0);
            }
            if (const char* kind = getenv("KG_PROBE2_KIND"); !kind || strcmp(kind, "coop")) { // default: the faster of the two variants
                const char* pe = getenv("KG_PROBE2_POLICY");
                CU(launch_windowed(k_probe2_own<4>, dim3(wgrid), dim3(CAS_BLK), 0, ls, nullptr, 0, (const unsigned long long*)sl.queue.as<unsigned long long>(),
                                   (const uint32_t*)sl.tile_qcnt.as<uint32_t>(), t0, t1, tv, sl.chunk_pos.as<uint32_t>(), sl.chunk_payload.as<int4>(),
                                   (uint32_t)hit_cap, sl.tile_base.as<uint32_t>(), sl.tile_cnt.as<uint32_t>(), d_ctr, (uint32_t)(pe ? atoi(pe) : 1)));
                sl.launches += 3;
                continue;
            }
            auto kp2 = probe2_u() == 2 ? k_probe2<2> : k_probe2<4>;
            CU(launch_windowed(kp2, dim3(wgrid), dim3(CAS_BLK), 0, ls, nullptr, 0, (const unsigned long long*)sl.queue.as<unsigned long long>(),
                               (const uint32_t*)sl.tile_qcnt.as<uint32_t>(), t0, t1, tv, sl.chunk_pos.as<uint32_t>(), sl.chunk_payload.as<int4>(),
                               (uint32_t)hit_cap, sl.tile_base.as<uint32_t>(), sl.tile_cnt.as<uint32_t>(), d_ctr));
            sl.launches += 3;
        }
        if (parts > 1) {
            cudaEventRecord(sl.part_ev[KG_MAX_PARTS], ls);
            cudaStreamWaitEvent(st, sl.part_ev[KG_MAX_PARTS], 0);
        }
    } else if (ntiles && table->filter_halves) { // two passes, one per half of the key space (k_probe_half)
        sl.cascade = false;
        const KgTableView tv = table->view();
        const size_t wbytes = (size_t)tv.filter_words * 8;
        KG_TRY(sl.half_pos.ensure(hit_cap * 4));
        KG_TRY(sl.half_payload.ensure(hit_cap * sizeof(int4)));
        KG_TRY(sl.half_base.ensure(((size_t)ntiles + 1) * 4));
        KG_TRY(sl.half_cnt.ensure(((size_t)ntiles + 1) * 4));
        CU(launch_windowed(k_probe_half<0>, dim3(ntiles), dim3(PROBE_BLK), PROBE_SMEM, st, tv.filter, wbytes, b->stream(), (uint32_t)vtotal, tv,
                           sl.half_pos.as<uint32_t>(), sl.half_payload.as<int4>(), (uint32_t)hit_cap, sl.half_base.as<uint32_t>(),
                           sl.half_cnt.as<uint32_t>(), d_ctr, (const uint32_t*)nullptr, (const int4*)nullptr, (const uint32_t*)nullptr,
                           (const uint32_t*)nullptr, probe_flags()));
        cudaEventRecord(sl.ev[4], st);
        CU(launch_windowed(k_probe_half<1>, dim3(ntiles), dim3(PROBE_BLK), PROBE_SMEM, st, tv.filter2, wbytes, b->stream(), (uint32_t)vtotal, tv,
                           sl.chunk_pos.as<uint32_t>(), sl.chunk_payload.as<int4>(), (uint32_t)hit_cap, sl.tile_base.as<uint32_t>(),
                           sl.tile_cnt.as<uint32_t>(), d_ctr, (const uint32_t*)sl.half_pos.as<uint32_t>(), (const int4*)sl.half_payload.as<int4>(),
                           (const uint32_t*)sl.half_base.as<uint32_t>(), (const uint32_t*)sl.half_cnt.as<uint32_t>(), probe_flags()));
        sl.halves = true;
        sl.launches += 2;
    } else if (ntiles) {
        sl.cascade = false;
        k_probe<<<ntiles, PROBE_BLK, PROBE_SMEM, st>>>(b->stream(), (uint32_t)vtotal, table->view(), sl.chunk_pos.as<uint32_t>(),
                                                      sl.chunk_payload.as<int4>(), (uint32_t)hit_cap, sl.tile_base.as<uint32_t>(),
                                                      sl.tile_cnt.as<uint32_t>(), d_ctr, probe_flags());
        sl.launches++;
    }
    cudaEventRecord(sl.ev[2], st);
    // Everything downstream of the probe runs on a second stream: in kg_run the run FSM of slice s (whose tail is a few
    // long sequences on a mostly idle GPU) then overlaps the probe of slice s+1, which is queued on the compute stream.
    st = ctx->fsm_stream;
    cudaStreamWaitEvent(st, sl.ev[2], 0);
    KG_TRY(exclusive_sum_u32(ctx, sl.tile_cnt.as<uint32_t>(), sl.tile_out.as<uint32_t>(), (size_t)ntiles + 1, st));
    sl.launches++;
    CU(cudaMemcpyAsync(sl.h_ctr, d_ctr, KG_CTR_COUNT * 8, cudaMemcpyDeviceToHost, st));
    sl.seg = use_segment_path(b->mode);
    if (!sl.seg) { // many short sequences: one thread per sequence straight off the per-tile chunks
        k_lo_tiles<<<blocks_for(nv + 1, 256), 256, 0, st>>>(b->voffsets(), nv, sl.tile_base.as<uint32_t>(), sl.tile_out.as<uint32_t>(), ntiles,
                                                           sl.chunk_pos.as<uint32_t>(), sl.lo.as<uint32_t>(), d_ctr);
        sl.launches++;
        if (b->n) {
            KG_TRY(sl.fi_a.ensure(b->n * 4));
            KG_TRY(sl.fk_a.ensure(2 * FSM_CLASSES * 4));
            uint32_t* hist = sl.fk_a.as<uint32_t>();
            CU(cudaMemsetAsync(hist, 0, 2 * FSM_CLASSES * 4, st));
            k_fsm_hist<<<blocks_for(b->n, 256), 256, 0, st>>>(sl.lo.as<uint32_t>(), b->n, per_seq, hist);
            k_fsm_scatter<<<blocks_for(b->n, 256), 256, 0, st>>>(sl.lo.as<uint32_t>(), b->n, per_seq, hist, hist + FSM_CLASSES, sl.fi_a.as<uint32_t>());
            k_fsm<<<blocks_for(b->n, FSM_BLK), FSM_BLK, 0, st>>>(b->voffsets(), b->n, per_seq, sl.tile_base.as<uint32_t>(),
                                                        sl.tile_out.as<uint32_t>(), ntiles, sl.chunk_pos.as<uint32_t>(),
                                                        sl.chunk_payload.as<int4>(), fp, sl.sparse.as<KgDevCall>(), sl.lo.as<uint32_t>(),
                                                        sl.fi_a.as<uint32_t>(), sl.call_cnt.as<uint32_t>(), r->d_otus.as<kg_otu>(), d_ctr);
            sl.launches += 3;
        }
        CU(cudaMemsetAsync(sl.call_cnt.as<uint32_t>() + nv, 0, 4, st));
        KG_TRY(exclusive_sum_u32(ctx, sl.call_cnt.as<uint32_t>(), sl.call_off.as<uint32_t>(), nv + 1, st));
        sl.launches++;
        if (nv) {
            k_compact_calls<<<blocks_for(nv, 256), 256, 0, st>>>(sl.sparse.as<KgDevCall>(), sl.lo.as<uint32_t>(), sl.call_off.as<uint32_t>(),
                                                                nv, per_seq, prm->min_hits, seq_base, r->d_calls.as<kg_call>());
            sl.launches++;
        }
        CU(cudaMemcpyAsync(&sl.h_ctr[KG_CTR_COUNT], sl.call_off.as<uint32_t>() + nv, 4, cudaMemcpyDeviceToHost, st));
    } else { // long contigs: position-ordered hits -> segments at gaps > max_gap -> one thread per segment -> OTU replay
        const uint32_t cap = (uint32_t)hit_cap;
        const uint32_t nblk = blocks_for(ntiles, GATHER_TILES);
        KG_TRY(sl.hit_pos.ensure((size_t)cap * 4));
        KG_TRY(sl.hit_payload.ensure((size_t)cap * sizeof(int4)));
        KG_TRY(sl.seg_tmp.ensure((size_t)cap * 4));
        KG_TRY(sl.seg_v.ensure((size_t)cap * 4));
        KG_TRY(sl.seg_perm.ensure((size_t)cap * 4));
        KG_TRY(sl.seg_begin.ensure(((size_t)cap + 1) * 4));
        KG_TRY(sl.seg_cnt.ensure(((size_t)cap + 1) * 8));
        KG_TRY(sl.seg_off.ensure(((size_t)cap + 1) * 8));
        KG_TRY(sl.nseg.ensure(16));
        KG_TRY(sl.tile_bnd.ensure((size_t)ntiles + 1));
        KG_TRY(sl.blk_cnt.ensure(((size_t)nblk + 1) * 4));
        KG_TRY(sl.blk_off.ensure(((size_t)nblk + 1) * 4));
        KG_TRY(sl.scan_part.ensure(SCAN_BLOCKS * 8));
        KG_TRY(sl.fk_a.ensure(2 * SEG_CLASSES * 4));
        KG_TRY(sl.o_run.ensure(((size_t)cap + 1) * 8));
        KG_TRY(sl.o_dense.ensure(((size_t)cap + 1) * 8));
        uint32_t* nseg = sl.nseg.as<uint32_t>(); // [0] segments, [2] calls, [3] OTU runs
        uint32_t* hist = sl.fk_a.as<uint32_t>();
        const unsigned wide = (unsigned)ctx->sm_count * 8; // grid of the kernels that walk a device-side count
        CU(cudaMemsetAsync(sl.tile_bnd.p, 0, (size_t)ntiles + 1, st));
        CU(cudaMemsetAsync(hist, 0, 2 * SEG_CLASSES * 4, st));
        CU(cudaMemsetAsync(nseg, 0, 16, st));
        if (nv) k_tile_bnd<<<blocks_for(nv, 256), 256, 0, st>>>(b->voffsets(), nv, ntiles, sl.tile_bnd.as<uint8_t>());
        if (ntiles) {
            k_gather_seg<<<nblk, GS_BLK, 0, st>>>(sl.chunk_pos.as<uint32_t>(), sl.chunk_payload.as<int4>(), sl.tile_base.as<uint32_t>(),
                                                  sl.tile_out.as<uint32_t>(), ntiles, b->voffsets(), nv, sl.tile_bnd.as<uint8_t>(), prm->max_gap,
                                                  sl.hit_pos.as<uint32_t>(), sl.hit_payload.as<int4>(), sl.seg_tmp.as<uint32_t>(),
                                                  sl.blk_cnt.as<uint32_t>(), d_ctr);
            k_seg_offsets<<<1, 1024, 0, st>>>(sl.blk_cnt.as<uint32_t>(), nblk, sl.blk_off.as<uint32_t>(), sl.tile_out.as<uint32_t>(), ntiles,
                                              sl.seg_begin.as<uint32_t>(), nseg, d_ctr);
            k_seg_place<<<nblk, GS_BLK, 0, st>>>(sl.seg_tmp.as<uint32_t>(), sl.blk_cnt.as<uint32_t>(), sl.blk_off.as<uint32_t>(),
                                                 sl.tile_out.as<uint32_t>(), sl.seg_begin.as<uint32_t>());
        }
        k_lo<<<blocks_for(nv + 1, 256), 256, 0, st>>>(b->voffsets(), nv, sl.hit_pos.as<uint32_t>(), sl.tile_out.as<uint32_t>(), ntiles,
                                                     sl.lo.as<uint32_t>(), d_ctr);
        k_seg_hist<<<wide, 256, 0, st>>>(sl.seg_begin.as<uint32_t>(), nseg, hist);
        k_seg_scatter<<<wide, 256, 0, st>>>(sl.seg_begin.as<uint32_t>(), nseg, hist, hist + SEG_CLASSES, sl.seg_perm.as<uint32_t>());
        k_fsm_seg<<<wide * 2, FS_BLK, 0, st>>>(b->voffsets(), nv, sl.hit_pos.as<uint32_t>(), sl.hit_payload.as<int4>(), sl.seg_begin.as<uint32_t>(),
                                            nseg, sl.seg_perm.as<uint32_t>(), sl.lo.as<uint32_t>(), fp, sl.sparse.as<KgDevCall>(),
                                            sl.seg_cnt.as<uint2>(), sl.seg_v.as<uint32_t>(), sl.o_run.as<int2>());
        k_scan2_sums<<<SCAN_BLOCKS, SCAN_BLK, 0, st>>>(sl.seg_cnt.as<uint2>(), nseg, sl.scan_part.as<uint2>());
        k_scan2_top<<<1, SCAN_BLK, 0, st>>>(sl.scan_part.as<uint2>(), nseg, sl.seg_off.as<uint2>(), nseg + 2);
        k_scan2_apply<<<SCAN_BLOCKS, SCAN_BLK, 0, st>>>(sl.seg_cnt.as<uint2>(), nseg, sl.scan_part.as<uint2>(), sl.seg_off.as<uint2>());
        k_compact_seg<<<wide, 256, 0, st>>>(sl.sparse.as<KgDevCall>(), sl.seg_begin.as<uint32_t>(), nseg, sl.seg_v.as<uint32_t>(),
                                            sl.seg_off.as<uint2>(), sl.o_run.as<int2>(), per_seq, prm->min_hits, seq_base,
                                            r->d_calls.as<kg_call>(), sl.o_dense.as<int2>());
        // OTU-COUNTS: the runs k_fsm_seg listed per segment, packed in print order, folded one sequence per thread
        if (b->n)
            k_otu_fold<<<(unsigned)b->n, 32, 0, st>>>(sl.lo.as<uint32_t>(), b->n, per_seq, sl.seg_begin.as<uint32_t>(), nseg,
                                                      sl.seg_off.as<uint2>(), sl.o_dense.as<int2>(), r->d_otus.as<kg_otu>(), d_ctr);
        sl.launches += 13;
        CU(cudaMemcpyAsync(&sl.h_ctr[KG_CTR_COUNT], nseg + 2, 4, cudaMemcpyDeviceToHost, st));
    }
    sl.h_ctr[KG_CTR_COUNT + 1] = 0;
    if (r->want_compact_otus && b->n) { // kg_run: the OTU counts go home as a count byte per sequence + the used pairs
        const uint64_t ns = b->n;
        KG_TRY(sl.otu_cnt.ensure((ns + 1) * 4));
        KG_TRY(sl.otu_first.ensure((ns + 1) * 4));
        auto need = [&](DevBuf* buf, size_t bytes) -> int {
            if (buf->cap >= bytes) return KG_OK;
            pool_give_dev(ctx, buf);
            return pool_take_dev(ctx, bytes, buf);
        };
        KG_TRY(need(&r->d_otu_n, ns));
        KG_TRY(need(&r->d_otu_entries, std::max<uint64_t>(std::min<uint64_t>(ns * KG_OI_BUFSZ, hit_cap), 1024) * sizeof(kg_otu_entry)));
        k_otu_counts<<<blocks_for(ns + 1, 256), 256, 0, st>>>(r->d_otus.as<kg_otu>(), ns, sl.otu_cnt.as<uint32_t>(), r->d_otu_n.as<uint8_t>(), d_ctr);
        KG_TRY(exclusive_sum_u32(ctx, sl.otu_cnt.as<uint32_t>(), sl.otu_first.as<uint32_t>(), ns + 1, st));
        k_otu_pack<<<blocks_for(ns, 256), 256, 0, st>>>(r->d_otus.as<kg_otu>(), ns, sl.otu_first.as<uint32_t>(), r->d_otu_entries.as<kg_otu_entry>());
        CU(cudaMemcpyAsync(&sl.h_ctr[KG_CTR_COUNT + 1], sl.otu_first.as<uint32_t>() + ns, 4, cudaMemcpyDeviceToHost, st));
        sl.launches += 3;
    }
    cudaEventRecord(sl.ev[3], st);
    return KG_OK;
}

static uint32_t ntiles_of(const kg_batch* b) { return (uint32_t)((b->vtotal + TILE - 1) >> TILE_SHIFT); }

// Wait for the pass, repeat it once if the hit buffers were too small, fill in the statistics, and (for "-d") build the
// position-ordered HIT records.
static int pipe_finish(kg_context* ctx, PipeSlot& sl, const kg_table* table, kg_batch* b, const kg_params* prm, kg_result* r,
                       uint32_t seq_base, const ProbeStage* custom = nullptr) {
    cudaStream_t st = ctx->stream;
    RunScratch& sc = scratch_of(ctx);
    uint32_t launches = 0;
    for (int attempt = 0;; attempt++) {
        CU(cudaEventSynchronize(sl.ev[3]));
        CU(cudaGetLastError());
        launches += sl.launches;
        if (!sl.h_ctr[KG_CTR_OVERFLOW]) break;
        if (attempt) KG_FAIL(KG_ECUDA, "hit buffer overflow persisted after resizing to %llu", (unsigned long long)sl.hit_cap);
        KG_TRY(pipe_enqueue(ctx, sl, table, b, prm, r, std::max(sl.h_ctr[KG_CTR_HITS], sl.h_ctr[KG_CTR_CLAIM]), seq_base, custom)); // exact size
    }
    const uint64_t nhits = sl.h_ctr[KG_CTR_HITS], nkmers = sl.h_ctr[KG_CTR_KMERS];
    const uint64_t nslots = std::max<uint64_t>(nhits, sl.h_ctr[KG_CTR_CLAIM]); // the cascade claims a slot per second-filter survivor
    sc.hit_cap_seen = std::max<uint64_t>(sc.hit_cap_seen, nslots + nslots / 16 + 1024);
    if (nhits > 0xFFFFFFF0ull) KG_FAIL(KG_ERANGE, "%llu hits in one batch", (unsigned long long)nhits);
    cudaEventElapsedTime(&r->stats.ms_prepare, sl.ev[0], sl.ev[1]);
    cudaEventElapsedTime(&r->stats.ms_probe, sl.ev[1], sl.ev[2]);
    cudaEventElapsedTime(&r->stats.ms_group, sl.ev[2], sl.ev[3]);
    r->stats.ms_filter = r->stats.ms_refilter = r->stats.ms_lines = 0.f;
    if (sl.halves && !custom && ntiles_of(b)) { // two-pass probe: the passes, in the cascade's fields
        cudaEventElapsedTime(&r->stats.ms_filter, sl.ev[1], sl.ev[4]);
        cudaEventElapsedTime(&r->stats.ms_lines, sl.ev[4], sl.ev[2]);
    }
    if (sl.cascade && sl.cascade_parts == 1 && !custom && ntiles_of(b)) {
        cudaEventElapsedTime(&r->stats.ms_filter, sl.ev[1], sl.ev[4]);
        cudaEventElapsedTime(&r->stats.ms_refilter, sl.ev[4], sl.ev[5]);
        cudaEventElapsedTime(&r->stats.ms_lines, sl.ev[5], sl.ev[2]);
    }
    cudaEventElapsedTime(&r->stats.ms_device, sl.ev[0], sl.ev[3]);
    const uint64_t nv = b->nv;
    const int per_seq = b->mode == KG_MODE_AA ? 1 : 6;
    const uint32_t ntiles = (uint32_t)((b->vtotal + TILE - 1) >> TILE_SHIFT);
    if (prm->emit_hits && nhits) { // "-d": position-ordered HIT records
        if (r->d_hits.cap < nhits * sizeof(kg_hit)) {
            pool_give_dev(ctx, &r->d_hits);
            KG_TRY(pool_take_dev(ctx, nhits * sizeof(kg_hit), &r->d_hits));
        }
        if (!sl.seg) { // the segment path has the position-ordered arrays already
            KG_TRY(sl.hit_pos.ensure(nhits * 4));
            KG_TRY(sl.hit_payload.ensure(nhits * sizeof(int4)));
            k_gather<<<blocks_for(ntiles, GATHER_TILES), 256, 0, st>>>(sl.chunk_pos.as<uint32_t>(), sl.chunk_payload.as<int4>(),
                                                                          sl.tile_base.as<uint32_t>(), sl.tile_out.as<uint32_t>(), ntiles,
                                                                          sl.hit_pos.as<uint32_t>(), sl.hit_payload.as<int4>(), sl.ctr.as<unsigned long long>());
        }
        k_emit_hits<<<blocks_for(nhits, 256), 256, 0, st>>>(b->voffsets(), nv, per_seq, sl.hit_pos.as<uint32_t>(),
                                                           sl.hit_payload.as<int4>(), (uint32_t)nhits, seq_base, r->d_hits.as<kg_hit>());
        launches += 2;
        CU(cudaStreamSynchronize(st));
        CU(cudaGetLastError());
    }
    if (sl.seg && getenv("KG_DEBUG_SEG")) { // developer aid: segment length distribution of this run
        uint32_t ns = 0, tot[2] = {0, 0};
        cudaMemcpy(&ns, sl.nseg.p, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(tot, sl.nseg.as<uint32_t>() + 2, 8, cudaMemcpyDeviceToHost);
        std::vector<uint32_t> sb((size_t)ns + 1);
        cudaMemcpy(sb.data(), sl.seg_begin.p, ((size_t)ns + 1) * 4, cudaMemcpyDeviceToHost);
        std::vector<uint32_t> len(ns);
        for (uint32_t i = 0; i < ns; i++) len[i] = sb[i + 1] - sb[i];
        std::sort(len.begin(), len.end());
        if (ns)
            fprintf(stderr, "[kg seg] %u segments, %u calls, %u OTU runs; hits/segment median %u p99 %u p99.9 %u max %u\n", ns, tot[0], tot[1],
                    len[ns / 2], len[(size_t)ns * 99 / 100], len[(size_t)ns * 999 / 1000], len[ns - 1]);
    }
    r->stats.num_sequences = b->n;
    r->stats.num_positions = b->vtotal;
    r->stats.num_kmers = nkmers;
    r->stats.num_hits = nhits;
    r->stats.num_calls = *(uint32_t*)&sl.h_ctr[KG_CTR_COUNT];
    r->stats.num_launches = launches;
    r->stats.num_survivors1 = sl.h_ctr[KG_CTR_SURV1];
    r->stats.num_survivors2 = sl.h_ctr[KG_CTR_CLAIM];
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
    if (table->shard_count > 1) KG_FAIL(KG_EINVAL, "kg_batch_run: the table is shard %d of %d; use kg_batch_run_sharded", table->shard_rank, table->shard_count);
    KG_TRY(check_params(params));
    CU(cudaSetDevice(ctx->device));
    kg_result* r = new kg_result();
    r->ctx = ctx;
    RunScratch& sc = scratch_of(ctx);
    int rc = pipe_enqueue(ctx, sc.slot[0], table, batch, params, r, sc.hit_cap_seen, 0);
    if (rc == KG_OK) rc = pipe_finish(ctx, sc.slot[0], table, batch, params, r, 0);
    if (rc != KG_OK) {
        kg_result_free(r);
        return rc;
    }
    *out = r;
    return KG_OK;
}

// Submit / collect: the same two-in-flight pipeline with the caller in the loop -- it consumes (and frees) result i while
// batch i+1 runs, so the result buffers cycle through the context's pool instead of piling up.
extern "C" int kg_batch_submit(kg_context* ctx, const kg_table* table, kg_batch* batch, const kg_params* params) {
    if (!ctx || !table || !batch) KG_FAIL(KG_EINVAL, "kg_batch_submit: null argument");
    if (table->shard_count > 1) KG_FAIL(KG_EINVAL, "kg_batch_submit: the table is shard %d of %d; use kg_batch_run_sharded", table->shard_rank, table->shard_count);
    KG_TRY(check_params(params));
    CU(cudaSetDevice(ctx->device));
    RunScratch& sc = scratch_of(ctx);
    if (sc.inflight_n >= 2) KG_FAIL(KG_EINVAL, "kg_batch_submit: two batches are in flight already; collect one first");
    const int slot = (sc.inflight_head + sc.inflight_n) & 1;
    RunScratch::InFlight& f = sc.inflight[slot];
    f.r = new kg_result();
    f.r->ctx = ctx;
    f.b = batch;
    f.t = table;
    f.p = *params;
    const int rc = pipe_enqueue(ctx, sc.slot[slot], table, batch, params, f.r, sc.hit_cap_seen, 0);
    if (rc != KG_OK) {
        cudaDeviceSynchronize();
        kg_result_free(f.r);
        f.r = nullptr;
        return rc;
    }
    sc.inflight_n++;
    return KG_OK;
}
extern "C" int kg_batch_collect(kg_context* ctx, kg_result** out) {
    if (!ctx || !out) KG_FAIL(KG_EINVAL, "kg_batch_collect: null argument");
    *out = nullptr;
    CU(cudaSetDevice(ctx->device));
    RunScratch& sc = scratch_of(ctx);
    if (!sc.inflight_n) KG_FAIL(KG_EINVAL, "kg_batch_collect: nothing in flight");
    const int slot = sc.inflight_head & 1;
    RunScratch::InFlight& f = sc.inflight[slot];
    const int rc = pipe_finish(ctx, sc.slot[slot], f.t, f.b, &f.p, f.r, 0);
    sc.inflight_head ^= 1;
    sc.inflight_n--;
    if (rc != KG_OK) {
        cudaDeviceSynchronize();
        kg_result_free(f.r);
        f.r = nullptr;
        return rc;
    }
    *out = f.r;
    f.r = nullptr;
    return KG_OK;
}

// Several resident batches, two in flight: the pipeline of batch i+1 is enqueued (second pipeline slot) before the host waits
// for batch i, so the run FSM and the call compaction of batch i -- on their own stream, with a tail of a few long sequences
// on a mostly idle GPU -- overlap the probe of batch i+1.  Results are those of n separate kg_batch_run calls.
extern "C" int kg_batch_run_many(kg_context* ctx, const kg_table* table, kg_batch* const* batches, size_t n, const kg_params* params,
                                 kg_result** results) {
    if (!ctx || !table || (n && (!batches || !results))) KG_FAIL(KG_EINVAL, "kg_batch_run_many: null argument");
    if (table->shard_count > 1) KG_FAIL(KG_EINVAL, "kg_batch_run_many: the table is shard %d of %d; use kg_batch_run_sharded", table->shard_rank, table->shard_count);
    KG_TRY(check_params(params));
    CU(cudaSetDevice(ctx->device));
    for (size_t i = 0; i < n; i++) {
        if (!batches[i]) KG_FAIL(KG_EINVAL, "kg_batch_run_many: batch %zu is null", i);
        results[i] = nullptr;
    }
    RunScratch& sc = scratch_of(ctx);
    int rc = KG_OK;
    auto enqueue = [&](size_t i) -> int {
        results[i] = new kg_result();
        results[i]->ctx = ctx;
        return pipe_enqueue(ctx, sc.slot[i & 1], table, batches[i], params, results[i], sc.hit_cap_seen, 0);
    };
    if (n) rc = enqueue(0);
    for (size_t i = 0; i < n && rc == KG_OK; i++) {
        if (i + 1 < n) rc = enqueue(i + 1);
        if (rc == KG_OK) rc = pipe_finish(ctx, sc.slot[i & 1], table, batches[i], params, results[i], 0);
    }
    if (rc != KG_OK) {
        cudaDeviceSynchronize(); // nothing may still be writing into a result that is about to be freed
        for (size_t i = 0; i < n; i++) {
            kg_result_free(results[i]);
            results[i] = nullptr;
        }
        return rc;
    }
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

// Host buffers in, host results out.  The batch is cut into slices of a few tens of MB at sequence boundaries (sequences
// are independent, KGJ:528/540); slice i+1 is copied to the device (copy stream) while slice i runs (compute stream) and
// the records of slice i-1 travel back (third stream), so the end-to-end time approaches max(H2D, device) instead of
// their sum.  Pinned caller buffers make the copies truly asynchronous; pageable ones still work.
// `packed`: seq_bytes / offsets are the 5-bit form of kg_pack_aa (offsets in groups of 8 residues = 5 bytes); the slice is
// unpacked on the device into the same residue stream (every sequence padded to a multiple of 8 positions).
struct DnaPacked { // kg_run_packed_dna: seq_bytes is the 2-bit form
    const uint64_t* byte_offsets; // n + 1
    const uint64_t* exceptions;   // sorted nucleotide positions (in the caller's offset space) whose dnaChar is 4
    size_t n_exceptions;
};
static int run_host_impl(kg_context* ctx, const kg_table* table, int mode, bool packed, const uint8_t* seq_bytes, const uint64_t* offsets,
                         size_t n, const kg_params* params, kg_result** out, const DnaPacked* dp = nullptr) {
    if (!ctx || !table || !out || !offsets || (mode != KG_MODE_AA && mode != KG_MODE_DNA)) KG_FAIL(KG_EINVAL, "kg_run: bad argument");
    if (packed && mode != KG_MODE_AA) KG_FAIL(KG_EINVAL, "kg_run_packed_aa: protein mode only");
    const uint64_t unit_pos = packed ? 8 : 1; // stream positions per offset unit (a group of the packed form = 5 bytes = 8 positions)
    if (table->shard_count > 1) KG_FAIL(KG_EINVAL, "kg_run: the table is shard %d of %d; use kg_batch_run_sharded", table->shard_rank, table->shard_count);
    KG_TRY(check_params(params));
    if (offsets[0] != 0) KG_FAIL(KG_EINVAL, "kg_run: offsets[0] must be 0");
    // offsets are checked slice by slice (validate, below) while the GPU already works on the slices before
    if (offsets[n] && !seq_bytes) KG_FAIL(KG_EINVAL, "kg_run: null sequence bytes");
    if (n >= 0xFFFFFFF0ull) KG_FAIL(KG_ERANGE, "kg_run: %zu sequences in one call", n);
    CU(cudaSetDevice(ctx->device));
    RunScratch& sc = scratch_of(ctx);

    // slice plan
    // about six slices per call, 24..96 MB of residues each (a slice costs ~0.4 ms of host-side API calls, so many
    // small slices would make the host the bottleneck; few large ones expose the first upload)
    uint64_t target = std::min<uint64_t>(std::max<uint64_t>(offsets[n] * unit_pos / 6, 24ull << 20), 96ull << 20); // in stream positions
    uint64_t ramp = 8;
    if (mode == KG_MODE_DNA) {
        // Long contigs: every slice pays ~1.5 ms of latency-bound tails (the longest segment in k_fsm_seg, the OTU fold
        // of the largest contig), whatever its size, so few big slices win (configs[2]: 14 slices 32 ms, 3 slices 12 ms).
        target = std::min<uint64_t>(std::max<uint64_t>(offsets[n] * 2 / 5, 32ull << 20), 128ull << 20);
        ramp = 2;
    }
    if (const char* e = getenv("KG_SLICE_MB")) target = (uint64_t)atoll(e) << 20;
    if (target < 65536) target = 65536;
    uint64_t hard = mode == KG_MODE_AA ? KG_MAX_STREAM : KG_MAX_STREAM / 2 - 64 * (uint64_t)n; // dna: 2 residues per nucleotide
    target /= unit_pos; // from here on in offset units (groups when packed)
    hard /= unit_pos;
    // the first slices are small (target/ramp, doubling): nothing can overlap the very first upload, so keep it short
    std::vector<size_t> cut{0};
    if (const char* e = getenv("KG_SLICE_RAMP")) ramp = std::max(1, atoi(e));
    uint64_t step = std::max<uint64_t>(target / ramp, std::min<uint64_t>(target, 2ull << 20));
    for (size_t i = 0; i < n;) {
        // last j with offsets[j] - offsets[i] <= step (binary search: a linear walk over a million offsets costs ~1 ms)
        size_t j = (size_t)(std::upper_bound(offsets + i, offsets + n + 1, offsets[i] + step) - offsets) - 1;
        if (j <= i) j = i + 1;
        if (offsets[n] - offsets[j] < step / 2 && offsets[n] - offsets[i] <= hard) j = n; // no tiny last slice: its FSM tail would be exposed
        step = std::min<uint64_t>(step * 2, target);
        if (offsets[j] < offsets[i]) KG_FAIL(KG_EINVAL, "kg_run: offsets must be non-decreasing (between %zu and %zu)", i, j);
        if (offsets[j] - offsets[i] > hard) KG_FAIL(KG_ERANGE, "kg_run: sequence %zu alone exceeds the per-call limit", i);
        cut.push_back(j);
        i = j;
    }
    if (n == 0) cut.push_back(0);
    const size_t nslices = cut.size() - 1;

    kg_result* R = new kg_result();
    R->ctx = ctx;
    R->n = n;
    R->mode = mode;
    R->params = *params;
    R->nv = (uint64_t)n * (mode == KG_MODE_AA ? 1 : 6);
    // device records of finished slices stay alive until their D2H copies are done: a ring of KEEP slices, each fenced
    // by an event on the D2H stream, so that the buffers return to the pool (no cudaMalloc in steady state)
    constexpr int KEEP = 3;
    DevBuf keep[KEEP][5];
    bool keep_used[KEEP] = {false, false, false};
    // uploads run UP-1 slices ahead of the compute stream: the copy engine never waits for a slice to finish
    constexpr size_t UP = 4;
    kg_batch* slot[UP] = {nullptr, nullptr, nullptr, nullptr};
    int rc = KG_OK;
    uint64_t ncalls = 0, nhits = 0;
    auto fail = [&](int code) {
        cudaDeviceSynchronize();
        for (auto& k3 : keep)
            for (auto& d : k3) pool_give_dev(ctx, &d);
        for (auto*& b : slot)
            if (b) { kg_batch_free(b); b = nullptr; }
        kg_result_free(R);
        return code;
    };
    auto upload = [&](size_t s) -> int { // slice s -> slot[s % UP], asynchronously on the copy stream
        const size_t a = cut[s], b = cut[s + 1], cnt = b - a;
        for (size_t i = a; i < b; i++)
            if (offsets[i + 1] < offsets[i]) KG_FAIL(KG_EINVAL, "kg_run: offsets must be non-decreasing (at %zu)", i);
        const uint64_t units = offsets[b] - offsets[a];
        if (units > hard) KG_FAIL(KG_ERANGE, "kg_run: sequence %zu alone exceeds the per-call limit", a);
        const uint64_t bytes = units * unit_pos; // residue-stream bytes of the slice
        kg_batch* bt = new kg_batch();
        bt->ctx = ctx;
        bt->mode = mode;
        bt->n = cnt;
        bt->total = bytes;
        slot[s % UP] = bt;
        KG_TRY(pool_take_dev(ctx, bytes + 64, &bt->seq_buf));
        KG_TRY(pool_take_dev(ctx, (cnt + 1) * 8, &bt->off_buf));
        bt->d_seq = bt->seq_buf.as<uint8_t>();
        bt->d_off = bt->off_buf.as<uint64_t>();
        if (mode == KG_MODE_DNA) { // residue-stream length of the six translations, as k_vlen lays them out
            uint64_t vt = 0;
            for (size_t i = a; i < b; i++) {
                const uint64_t L = offsets[i + 1] - offsets[i];
                for (uint64_t f = 0; f < 3; f++) vt += 2 * (((L >= f + 3 ? (L - f) / 3 : 0) + 1 + 3) & ~3ull);
            }
            bt->vtotal = vt;
            bt->vtotal_known = true;
        }
        cudaStream_t cs = ctx->copy_stream;
        CU(cudaMemsetAsync(bt->d_seq + bytes, 0, 64, cs));
        if (packed) {
            bt->padded = true;
            KG_TRY(pool_take_dev(ctx, units * 5 + 64, &bt->pk_buf));
            if (units) {
                CU(cudaMemcpyAsync(bt->pk_buf.p, seq_bytes + 5 * offsets[a], units * 5, cudaMemcpyHostToDevice, cs));
                k_unpack_aa<<<blocks_for(units, 256), 256, 0, cs>>>(bt->pk_buf.as<uint8_t>(), units, reinterpret_cast<uint2*>(bt->d_seq));
            }
        } else if (bytes && !dp) {
            CU(cudaMemcpyAsync(bt->d_seq, seq_bytes + offsets[a], bytes, cudaMemcpyHostToDevice, cs));
        }
        CU(cudaMemcpyAsync(bt->d_off, offsets + a, (cnt + 1) * 8, cudaMemcpyHostToDevice, cs));
        if (offsets[a] || packed) k_rebase<<<blocks_for(cnt + 1, 256), 256, 0, cs>>>(bt->d_off, cnt + 1, offsets[a], packed ? 8 : 1);
        if (dp && bytes) { // 2-bit nucleotides + the positions of the other characters: a quarter of the bytes cross the host link
            const uint64_t pb0 = dp->byte_offsets[a], pbytes = dp->byte_offsets[b] - pb0;
            if (dp->byte_offsets[b] < pb0) KG_FAIL(KG_EINVAL, "kg_run_packed_dna: byte offsets must be non-decreasing");
            KG_TRY(pool_take_dev(ctx, pbytes + 64, &bt->pk_buf));
            KG_TRY(pool_take_dev(ctx, (cnt + 1) * 8, &bt->aux_buf));
            CU(cudaMemcpyAsync(bt->pk_buf.p, seq_bytes + pb0, pbytes, cudaMemcpyHostToDevice, cs));
            CU(cudaMemcpyAsync(bt->aux_buf.p, dp->byte_offsets + a, (cnt + 1) * 8, cudaMemcpyHostToDevice, cs));
            if (pbytes)
                k_unpack_dna<<<blocks_for(pbytes, 256), 256, 0, cs>>>(bt->pk_buf.as<uint8_t>(), pbytes, bt->aux_buf.as<uint64_t>(), pb0, bt->d_off, cnt,
                                                                      bt->d_seq);
            const uint64_t* e0 = std::lower_bound(dp->exceptions, dp->exceptions + dp->n_exceptions, offsets[a]);
            const uint64_t* e1 = std::lower_bound(e0, dp->exceptions + dp->n_exceptions, offsets[b]);
            if (e1 > e0) {
                const uint64_t ne = (uint64_t)(e1 - e0);
                KG_TRY(pool_take_dev(ctx, ne * 8, &bt->exc_buf));
                CU(cudaMemcpyAsync(bt->exc_buf.p, e0, ne * 8, cudaMemcpyHostToDevice, cs));
                k_mark_dna_exceptions<<<blocks_for(ne, 256), 256, 0, cs>>>(bt->exc_buf.as<uint64_t>(), ne, offsets[a], bt->d_seq);
            }
        }
        CU(cudaEventRecord(ctx->up_ev[s % UP], cs));
        return KG_OK;
    };

    // OTU counts travel home compact: a count byte per sequence + the used (count, oI) pairs (kg_result_otus expands on demand)
    if ((rc = pool_take_host(ctx, std::max<size_t>(n, 1), &R->h_otu_n)) != KG_OK) return fail(rc);
    if ((rc = pool_take_host(ctx, std::max<uint64_t>(sc.otu_entries_seen + sc.otu_entries_seen / 4, 1u << 16) * sizeof(kg_otu_entry), &R->h_otu_entries)) != KG_OK) return fail(rc);
    uint64_t nentries = 0;
    if ((rc = pool_take_host(ctx, std::max<uint64_t>(sc.calls_seen + sc.calls_seen / 4, 1u << 16) * sizeof(kg_call), &R->h_calls)) != KG_OK) return fail(rc);
    if (params->emit_hits && (rc = pool_take_host(ctx, std::max<uint64_t>(sc.hit_cap_seen, 1u << 16) * sizeof(kg_hit), &R->h_hits)) != KG_OK) return fail(rc);
    auto grow = [&](HostBuf* hb, uint64_t need_bytes, uint64_t used_bytes) -> int { // rare: first run, or a batch unlike the last
        if (need_bytes <= hb->cap) return KG_OK;
        CU(cudaStreamSynchronize(ctx->d2h_stream));
        HostBuf nb;
        KG_TRY(pool_take_host(ctx, need_bytes + need_bytes / 2, &nb));
        if (used_bytes) memcpy(nb.p, hb->p, used_bytes);
        pool_give_host(ctx, hb);
        *hb = nb;
        return KG_OK;
    };

    const bool dbg = getenv("KG_DEBUG") != nullptr;
    auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now_ms();
    kg_result part[2];
    part[0].ctx = part[1].ctx = ctx;
    auto drop_parts = [&]() {
        for (auto& pt : part) {
            pool_give_dev(ctx, &pt.d_calls);
            pool_give_dev(ctx, &pt.d_otus);
            pool_give_dev(ctx, &pt.d_hits);
            pool_give_dev(ctx, &pt.d_otu_n);
            pool_give_dev(ctx, &pt.d_otu_entries);
        }
    };
    auto enqueue = [&](size_t s) -> int { // slice s runs as soon as its bytes have landed; nothing here waits for the GPU
        cudaStreamWaitEvent(ctx->stream, ctx->up_ev[s % UP], 0);
        part[s & 1].want_compact_otus = true;
        return pipe_enqueue(ctx, sc.slot[s & 1], table, slot[s % UP], params, &part[s & 1], sc.hit_cap_seen, (uint32_t)cut[s]);
    };
    for (size_t s = 0; s < std::min(nslices, UP - 1); s++)
        if ((rc = upload(s)) != KG_OK) return fail(rc);
    if (nslices && (rc = enqueue(0)) != KG_OK) {
        drop_parts();
        return fail(rc);
    }
    for (size_t s = 0; s < nslices; s++) {
        const double t0 = now_ms();
        if (keep_used[s % KEEP]) { // the ring slot this slice will take: its copies are long done
            cudaEventSynchronize(ctx->d2h_ev[s % KEEP]);
            for (auto& d : keep[s % KEEP]) pool_give_dev(ctx, &d);
            keep_used[s % KEEP] = false;
        }
        // queue slice s+1 behind slice s BEFORE waiting for slice s: the GPU never idles between slices
        if ((s + UP - 1 < nslices && (rc = upload(s + UP - 1)) != KG_OK) || (s + 1 < nslices && (rc = enqueue(s + 1)) != KG_OK)) {
            drop_parts();
            return fail(rc);
        }
        kg_batch* bt = slot[s % UP];
        kg_result& pr = part[s & 1];
        rc = pipe_finish(ctx, sc.slot[s & 1], table, bt, params, &pr, (uint32_t)cut[s]);
        if (rc != KG_OK) {
            drop_parts();
            return fail(rc);
        }
        const float ms = pr.stats.ms_device;
        // records of this slice go home on the third stream while the next slice computes
        const kg_run_stats ps = pr.stats;
        cudaStream_t ds = ctx->d2h_stream;
        if ((rc = grow(&R->h_calls, (ncalls + ps.num_calls) * sizeof(kg_call), ncalls * sizeof(kg_call))) != KG_OK) { drop_parts(); return fail(rc); }
        if (ps.num_calls) cudaMemcpyAsync((kg_call*)R->h_calls.p + ncalls, pr.d_calls.p, ps.num_calls * sizeof(kg_call), cudaMemcpyDeviceToHost, ds);
        const uint64_t slice_entries = sc.slot[s & 1].h_ctr[KG_CTR_COUNT + 1];
        if ((rc = grow(&R->h_otu_entries, (nentries + slice_entries) * sizeof(kg_otu_entry), nentries * sizeof(kg_otu_entry))) != KG_OK) { drop_parts(); return fail(rc); }
        if (bt->n) cudaMemcpyAsync((uint8_t*)R->h_otu_n.p + cut[s], pr.d_otu_n.p, bt->n, cudaMemcpyDeviceToHost, ds);
        if (slice_entries) cudaMemcpyAsync((kg_otu_entry*)R->h_otu_entries.p + nentries, pr.d_otu_entries.p, slice_entries * sizeof(kg_otu_entry), cudaMemcpyDeviceToHost, ds);
        nentries += slice_entries;
        if (params->emit_hits) {
            if ((rc = grow(&R->h_hits, (nhits + ps.num_hits) * sizeof(kg_hit), nhits * sizeof(kg_hit))) != KG_OK) { drop_parts(); return fail(rc); }
            if (ps.num_hits) cudaMemcpyAsync((kg_hit*)R->h_hits.p + nhits, pr.d_hits.p, ps.num_hits * sizeof(kg_hit), cudaMemcpyDeviceToHost, ds);
        }
        {
            const int k = (int)(s % KEEP);
            keep[k][0] = pr.d_calls;
            keep[k][1] = pr.d_otus;
            keep[k][2] = pr.d_hits;
            keep[k][3] = pr.d_otu_n;
            keep[k][4] = pr.d_otu_entries;
            keep_used[k] = true;
            cudaEventRecord(ctx->d2h_ev[k], ds);
            pr.d_calls = pr.d_otus = pr.d_hits = pr.d_otu_n = pr.d_otu_entries = DevBuf();
        }
        ncalls += ps.num_calls;
        nhits += ps.num_hits;
        R->stats.num_positions += ps.num_positions;
        R->stats.num_kmers += ps.num_kmers;
        R->stats.num_launches += ps.num_launches;
        R->stats.num_survivors1 += ps.num_survivors1;
        R->stats.num_survivors2 += ps.num_survivors2;
        R->stats.ms_device += ms;
        R->stats.ms_prepare += ps.ms_prepare;
        R->stats.ms_probe += ps.ms_probe;
        R->stats.ms_group += ps.ms_group;
        R->stats.ms_filter += ps.ms_filter;
        R->stats.ms_refilter += ps.ms_refilter;
        R->stats.ms_lines += ps.ms_lines;
        kg_batch_free(bt);
        slot[s % UP] = nullptr;
        if (dbg) fprintf(stderr, "[kg] slice %zu: %zu seqs, host %.3f ms (device %.3f = prepare %.3f + probe %.3f + group %.3f ms), since start %.3f ms\n", s, (size_t)(cut[s + 1] - cut[s]), now_ms() - t0, ms, ps.ms_prepare, ps.ms_probe, ps.ms_group, now_ms() - t_begin);
    }
    CU(cudaStreamSynchronize(ctx->d2h_stream));
    if (dbg) fprintf(stderr, "[kg] kg_run total %.3f ms\n", now_ms() - t_begin);
    for (auto& k3 : keep)
        for (auto& d : k3) pool_give_dev(ctx, &d);
    R->stats.num_sequences = n;
    R->stats.num_calls = ncalls;
    R->stats.num_hits = nhits;
    sc.calls_seen = std::max<uint64_t>(sc.calls_seen, ncalls);
    sc.otu_entries_seen = std::max<uint64_t>(sc.otu_entries_seen, nentries);
    R->num_otu_entries = nentries;
    R->otus_compact = true;
    R->fetched = true;
    *out = R;
    return KG_OK;
}

extern "C" int kg_run(kg_context* ctx, const kg_table* table, int mode, const uint8_t* seq_bytes, const uint64_t* offsets,
                      size_t n, const kg_params* params, kg_result** out) {
    return run_host_impl(ctx, table, mode, false, seq_bytes, offsets, n, params, out);
}

extern "C" int kg_run_packed_aa(kg_context* ctx, const kg_table* table, const uint8_t* packed, const uint64_t* group_offsets, size_t n,
                                const kg_params* params, kg_result** out) {
    return run_host_impl(ctx, table, KG_MODE_AA, true, packed, group_offsets, n, params, out);
}

// toAminoAcidOff (KGJ:111-175) on the host -- what prepareQuery does before addKmers (KGJ:1055-1058) -- then 8 codes in 5 bytes
extern "C" uint64_t kg_pack_aa_groups(uint64_t len) { return (len + 1 + 7) / 8; }

extern "C" int kg_pack_aa(const uint8_t* seq_bytes, const uint64_t* offsets, size_t n, uint8_t* packed, uint64_t* group_offsets, int threads) {
    if (!offsets || !group_offsets || (offsets[n] && !seq_bytes)) KG_FAIL(KG_EINVAL, "kg_pack_aa: null argument");
    uint8_t lut[256];
    for (int i = 0; i < 256; i++) lut[i] = 20;
    for (int i = 0; i < 20; i++) lut[(unsigned char)"ACDEFGHIKLMNPQRSTVWY"[i]] = (uint8_t)i;
    uint64_t g = 0;
    for (size_t s = 0; s < n; s++) { // the layout first (cheap), so that the packing itself can run on any number of threads
        if (offsets[s + 1] < offsets[s]) KG_FAIL(KG_EINVAL, "kg_pack_aa: offsets must be non-decreasing (at %zu)", s);
        group_offsets[s] = g;
        g += kg_pack_aa_groups(offsets[s + 1] - offsets[s]);
    }
    group_offsets[n] = g;
    if (!packed) return KG_OK; // layout only: the caller sizes its buffer with group_offsets[n] * 5
    auto work = [&](size_t s0, size_t s1) {
        for (size_t s = s0; s < s1; s++) {
            const uint8_t* src = seq_bytes + offsets[s];
            const uint64_t len = offsets[s + 1] - offsets[s];
            uint8_t* dst = packed + 5 * group_offsets[s];
            const uint64_t ng = group_offsets[s + 1] - group_offsets[s];
            for (uint64_t q = 0; q < ng; q++) {
                uint64_t bits = 0;
                for (int i = 0; i < 8; i++) {
                    const uint64_t r = 8 * q + i;
                    bits |= (uint64_t)(r < len ? lut[src[r]] : 31) << (5 * i);
                }
                for (int i = 0; i < 5; i++) dst[5 * q + i] = (uint8_t)(bits >> (8 * i));
            }
        }
    };
    const int T = std::max(1, std::min<int>(threads, 64));
    if (T == 1 || n < 1024) {
        work(0, n);
        return KG_OK;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) { // equal numbers of groups per thread
        const uint64_t g0 = g * t / T, g1 = g * (t + 1) / T;
        const size_t s0 = (size_t)(std::lower_bound(group_offsets, group_offsets + n, g0) - group_offsets);
        const size_t s1 = t + 1 == T ? n : (size_t)(std::lower_bound(group_offsets, group_offsets + n, g1) - group_offsets);
        th.emplace_back(work, s0, s1);
    }
    for (auto& x : th) x.join();
    return KG_OK;
}

extern "C" int kg_pack_dna(const uint8_t* seq_bytes, const uint64_t* offsets, size_t n, uint8_t* packed, uint64_t* byte_offsets,
                           uint64_t* exceptions, size_t* n_exceptions, int threads) {
    if (!offsets || !byte_offsets || !n_exceptions || (offsets[n] && !seq_bytes)) KG_FAIL(KG_EINVAL, "kg_pack_dna: null argument");
    uint8_t lut[256]; // dnaChar, KGJ:294-318
    for (int i = 0; i < 256; i++) lut[i] = 4;
    lut['a'] = lut['A'] = 0;
    lut['c'] = lut['C'] = 1;
    lut['g'] = lut['G'] = 2;
    lut['t'] = lut['T'] = lut['u'] = lut['U'] = 3;
    uint64_t pb = 0;
    for (size_t s = 0; s < n; s++) {
        if (offsets[s + 1] < offsets[s]) KG_FAIL(KG_EINVAL, "kg_pack_dna: offsets must be non-decreasing (at %zu)", s);
        byte_offsets[s] = pb;
        pb += (offsets[s + 1] - offsets[s] + 3) / 4;
    }
    byte_offsets[n] = pb;
    // threads take equal ranges of nucleotides; a range may start inside a contig, but always on one of its packed bytes
    const int T = std::max(1, std::min<int>(threads, 64));
    const uint64_t total = offsets[n];
    std::vector<std::vector<uint64_t>> exc((size_t)T);
    std::vector<size_t> s_first((size_t)T + 1, n);
    std::vector<uint64_t> p_first((size_t)T + 1, total);
    for (int t = 0; t < T; t++) { // first nucleotide of thread t: rounded down to a packed-byte boundary of its contig
        uint64_t x = total * (uint64_t)t / (uint64_t)T;
        size_t sidx = (size_t)(std::upper_bound(offsets, offsets + n + 1, x) - offsets);
        sidx = sidx ? sidx - 1 : 0;
        if (sidx >= n) {
            s_first[t] = n;
            p_first[t] = total;
            continue;
        }
        x = offsets[sidx] + (x - offsets[sidx]) / 4 * 4;
        s_first[t] = sidx;
        p_first[t] = x;
    }
    p_first[0] = 0;
    s_first[0] = 0;
    auto work = [&](int t) {
        size_t sidx = s_first[t];
        uint64_t x = p_first[t];
        const uint64_t end = p_first[t + 1];
        while (x < end && sidx < n) {
            const uint64_t a = offsets[sidx], b = offsets[sidx + 1];
            if (x >= b) {
                sidx++;
                continue;
            }
            const uint64_t stop = std::min(b, end);
            for (uint64_t g0 = x; g0 < stop; g0 += 4) { // one packed byte (g0 - a is a multiple of 4)
                uint32_t v = 0;
                const uint64_t g1 = std::min(g0 + 4, b);
                for (uint64_t g = g0; g < g1; g++) {
                    const uint8_t c = lut[seq_bytes[g]];
                    if (c == 4) exc[t].push_back(g);
                    else v |= (uint32_t)c << (2 * (g - g0));
                }
                if (packed) packed[byte_offsets[sidx] + (g0 - a) / 4] = (uint8_t)v;
            }
            x = stop;
        }
    };
    if (T == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    size_t ne = 0;
    for (auto& v : exc) ne += v.size();
    if (packed) {
        if (ne > *n_exceptions) KG_FAIL(KG_EINVAL, "kg_pack_dna: room for %zu exceptions, %zu found (call with packed = NULL first)", *n_exceptions, ne);
        if (ne && !exceptions) KG_FAIL(KG_EINVAL, "kg_pack_dna: null exceptions");
        size_t k = 0;
        for (auto& v : exc)
            for (uint64_t g : v) exceptions[k++] = g; // ascending: the threads' ranges ascend
    }
    *n_exceptions = ne;
    return KG_OK;
}

extern "C" int kg_run_packed_dna(kg_context* ctx, const kg_table* table, const uint8_t* packed, const uint64_t* offsets,
                                 const uint64_t* byte_offsets, const uint64_t* exceptions, size_t n_exceptions, size_t n,
                                 const kg_params* params, kg_result** out) {
    if (!byte_offsets || (n_exceptions && !exceptions)) KG_FAIL(KG_EINVAL, "kg_run_packed_dna: null argument");
    DnaPacked dp = {byte_offsets, exceptions, n_exceptions};
    return run_host_impl(ctx, table, KG_MODE_DNA, false, packed, offsets, n, params, out, &dp);
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
    if (r->otus_compact && !r->h_otus.p) { // kg_run brought the counts home compact: expand once, on demand
        KG_TRY(pool_take_host(r->ctx, std::max<uint64_t>(r->n, 1) * sizeof(kg_otu), &r->h_otus));
        kg_otu* o = (kg_otu*)r->h_otus.p;
        const uint8_t* cn = (const uint8_t*)r->h_otu_n.p;
        const kg_otu_entry* en = (const kg_otu_entry*)r->h_otu_entries.p;
        memset(o, 0, r->n * sizeof(kg_otu));
        uint64_t e = 0;
        for (uint64_t s = 0; s < r->n; s++) {
            o[s].n = cn[s];
            for (int k = 0; k < cn[s]; k++, e++) {
                o[s].count[k] = en[e].count;
                o[s].oI[k] = en[e].oI;
            }
        }
    }
    *otus = (const kg_otu*)r->h_otus.p;
    *n = r->n;
    return KG_OK;
}
extern "C" int kg_result_otus_compact(kg_result* r, const uint8_t** n_per_seq, const kg_otu_entry** entries, size_t* n_seqs, size_t* n_entries) {
    if (!r || !n_per_seq || !entries || !n_seqs || !n_entries) KG_FAIL(KG_EINVAL, "kg_result_otus_compact: null argument");
    KG_TRY(kg_result_fetch(r));
    if (!r->otus_compact) { // a kg_batch_run result holds full records: compact them once
        const kg_otu* o = (const kg_otu*)r->h_otus.p;
        uint64_t tot = 0;
        for (uint64_t s = 0; s < r->n; s++) tot += (uint64_t)o[s].n;
        KG_TRY(pool_take_host(r->ctx, std::max<uint64_t>(r->n, 1), &r->h_otu_n));
        KG_TRY(pool_take_host(r->ctx, std::max<uint64_t>(tot, 1) * sizeof(kg_otu_entry), &r->h_otu_entries));
        uint8_t* cn = (uint8_t*)r->h_otu_n.p;
        kg_otu_entry* en = (kg_otu_entry*)r->h_otu_entries.p;
        uint64_t e = 0;
        for (uint64_t s = 0; s < r->n; s++) {
            cn[s] = (uint8_t)o[s].n;
            for (int k = 0; k < o[s].n; k++, e++) {
                en[e].count = o[s].count[k];
                en[e].oI = o[s].oI[k];
            }
        }
        r->num_otu_entries = tot;
        r->otus_compact = true;
    }
    *n_per_seq = (const uint8_t*)r->h_otu_n.p;
    *entries = (const kg_otu_entry*)r->h_otu_entries.p;
    *n_seqs = r->n;
    *n_entries = r->num_otu_entries;
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
    pool_give_dev(r->ctx, &r->d_otu_n);
    pool_give_dev(r->ctx, &r->d_otu_entries);
    pool_give_host(r->ctx, &r->h_calls);
    pool_give_host(r->ctx, &r->h_otus);
    pool_give_host(r->ctx, &r->h_hits);
    pool_give_host(r->ctx, &r->h_otu_n);
    pool_give_host(r->ctx, &r->h_otu_entries);
    delete r;
}

#include "kg_shard.cuh" // hash-sharded table (same translation unit: it reuses the probe helpers and the pipeline)
